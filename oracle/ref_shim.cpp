// TEST INFRASTRUCTURE ONLY -- never part of the product path.
//
// extern "C" shim over the UNMODIFIED reference library (FEniCS/basix C++ core)
// so that tests/ and bench.py's cpu_baseline / --impl reference legs can call
// the reference's own implementation of the hot path through ctypes.
//
// Every entry point forwards to the public reference API:
//   basix::create_element<double>        cpp/basix/finite-element.h:1623-1628
//   FiniteElement::tabulate              cpp/basix/finite-element.h:482-483
//   polyset::tabulate                    cpp/basix/polyset.h:181-183
//   FiniteElement::push_forward          cpp/basix/finite-element.h:577-579
//   FiniteElement::pull_back             cpp/basix/finite-element.h:592-594
//   FiniteElement::T_apply & 7 siblings  cpp/basix/finite-element.h:1067-1159
//   FiniteElement::permute/permute_inv   cpp/basix/finite-element.h:800-843
// Nothing here re-implements reference arithmetic; the batched/threaded loops
// only call the reference once per cell/chunk (the reference itself is serial).
#include <basix/cell.h>
#include <basix/finite-element.h>
#include <basix/interpolation.h>
#include <basix/polyset.h>

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <span>
#include <string>
#include <thread>
#include <vector>

using namespace basix;
using FE = FiniteElement<double>;
template <typename T, std::size_t d>
using mds = md::mdspan<T, md::dextents<std::size_t, d>>;

namespace
{
thread_local std::string g_err;

template <typename F>
int guarded(F&& f)
{
  try
  {
    f();
    return 0;
  }
  catch (const std::exception& e)
  {
    g_err = e.what();
    return 1;
  }
}

template <typename F>
void parallel_for(std::size_t n, int nthreads, F&& f)
{
  if (nthreads <= 1 or n < 2)
  {
    f(std::size_t(0), n);
    return;
  }
  std::vector<std::thread> pool;
  const std::size_t chunk = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; ++t)
  {
    const std::size_t b = std::min(n, t * chunk), e = std::min(n, b + chunk);
    if (b < e)
      pool.emplace_back([=, &f]() { f(b, e); });
  }
  for (auto& th : pool)
    th.join();
}
} // namespace

extern "C"
{
const char* bxref_last_error() { return g_err.c_str(); }

void* bxref_create_element(int family, int cell, int degree, int lvariant,
                           int dvariant, int discontinuous,
                           const int* dof_ordering, int n_dof_ordering)
{
  FE* e = nullptr;
  guarded(
      [&]()
      {
        std::vector<int> ord(dof_ordering, dof_ordering + n_dof_ordering);
        e = new FE(basix::create_element<double>(
            static_cast<element::family>(family),
            static_cast<cell::type>(cell), degree,
            static_cast<element::lagrange_variant>(lvariant),
            static_cast<element::dpc_variant>(dvariant), discontinuous != 0,
            ord));
      });
  return e;
}

void bxref_destroy_element(void* h) { delete static_cast<FE*>(h); }

// info[0..11]: cell_type, polyset_type, tdim, ndofs, value_size, map_type,
// embedded_superdegree, perms flag, identity flag, coeffs cols, degree,
// n_dof_ordering
void bxref_element_info(void* h, int* info)
{
  const FE& e = *static_cast<FE*>(h);
  std::size_t vs = 1;
  for (auto s : e.value_shape())
    vs *= s;
  info[0] = static_cast<int>(e.cell_type());
  info[1] = static_cast<int>(e.polyset_type());
  info[2] = cell::topological_dimension(e.cell_type());
  info[3] = e.dim();
  info[4] = static_cast<int>(vs);
  info[5] = static_cast<int>(e.map_type());
  info[6] = e.embedded_superdegree();
  info[7] = e.dof_transformations_are_permutations();
  info[8] = e.dof_transformations_are_identity();
  info[9] = static_cast<int>(e.coefficient_matrix().second[1]);
  info[10] = e.degree();
  info[11] = static_cast<int>(e.dof_ordering().size());
}

int bxref_value_shape(void* h, int* shape)
{
  const FE& e = *static_cast<FE*>(h);
  int i = 0;
  for (auto s : e.value_shape())
    shape[i++] = static_cast<int>(s);
  return i;
}

void bxref_coeffs(void* h, double* out)
{
  const auto& c = static_cast<FE*>(h)->coefficient_matrix();
  std::copy(c.first.begin(), c.first.end(), out);
}

void bxref_dof_ordering(void* h, int* out)
{
  const auto& o = static_cast<FE*>(h)->dof_ordering();
  std::copy(o.begin(), o.end(), out);
}

// Number of entities of dimension d
int bxref_num_entities(void* h, int d)
{
  const auto& ed = static_cast<FE*>(h)->entity_dofs();
  return d < static_cast<int>(ed.size()) ? static_cast<int>(ed[d].size()) : 0;
}

// Dofs of entity (d, i); returns count, writes into out if non-null
int bxref_entity_dofs(void* h, int d, int i, int* out)
{
  const auto& v = static_cast<FE*>(h)->entity_dofs()[d][i];
  if (out)
    std::copy(v.begin(), v.end(), out);
  return static_cast<int>(v.size());
}

int bxref_sub_entity_type(int cell, int d, int i)
{
  return static_cast<int>(
      cell::sub_entity_type(static_cast<cell::type>(cell), d, i));
}

// Entity transformation matrices for sub-entity cell type `ct`:
// shape (ntrans, n, n).  Returns 0 and shape={0,0,0} if absent.
int bxref_entity_transformations(void* h, int ct, int* shape, double* out)
{
  const auto& m = static_cast<FE*>(h)->entity_transformations();
  auto it = m.find(static_cast<cell::type>(ct));
  if (it == m.end())
  {
    shape[0] = shape[1] = shape[2] = 0;
    return 0;
  }
  for (int i = 0; i < 3; ++i)
    shape[i] = static_cast<int>(it->second.second[i]);
  if (out)
    std::copy(it->second.first.begin(), it->second.first.end(), out);
  return 1;
}

// FiniteElement::base_transformations() (finite-element.cpp:1902-1969): shape (nt, ndofs, ndofs)
void bxref_base_transformations(void* h, int* shape, double* out)
{
  auto [data, shp] = static_cast<FE*>(h)->base_transformations();
  for (int i = 0; i < 3; ++i)
    shape[i] = static_cast<int>(shp[i]);
  if (out)
    std::copy(data.begin(), data.end(), out);
}

// FiniteElement::entity_closure_dofs()[d][i] (finite-element.h:686-690); returns the count
int bxref_entity_closure_dofs(void* h, int d, int i, int* out)
{
  const auto& v = static_cast<FE*>(h)->entity_closure_dofs()[d][i];
  if (out)
    std::copy(v.begin(), v.end(), out);
  return static_cast<int>(v.size());
}

// FiniteElement::points() (finite-element.h:1167-1170): shape (npts, tdim)
void bxref_points(void* h, int* shape, double* out)
{
  const auto& p = static_cast<FE*>(h)->points();
  shape[0] = static_cast<int>(p.second[0]);
  shape[1] = static_cast<int>(p.second[1]);
  if (out)
    std::copy(p.first.begin(), p.first.end(), out);
}

// FiniteElement::interpolation_matrix() (finite-element.h:1222-1226): shape (ndofs, value_size * npts)
void bxref_interpolation_matrix(void* h, int* shape, double* out)
{
  const auto& m = static_cast<FE*>(h)->interpolation_matrix();
  shape[0] = static_cast<int>(m.second[0]);
  shape[1] = static_cast<int>(m.second[1]);
  if (out)
    std::copy(m.first.begin(), m.first.end(), out);
}

int bxref_polyset_dim(int cell, int ptype, int d)
{
  return polyset::dim(static_cast<cell::type>(cell),
                      static_cast<polyset::type>(ptype), d);
}

int bxref_polyset_nderivs(int cell, int n)
{
  return polyset::nderivs(static_cast<cell::type>(cell), n);
}

// P[nderivs][psize][npts]
int bxref_polyset_tabulate(int cell, int ptype, int d, int n, const double* x,
                           std::size_t npts, int tdim, double* P)
{
  return guarded(
      [&]()
      {
        const auto ct = static_cast<cell::type>(cell);
        const auto pt = static_cast<polyset::type>(ptype);
        mds<double, 3> Pm(P, polyset::nderivs(ct, n), polyset::dim(ct, pt, d),
                          npts);
        mds<const double, 2> xm(x, npts, tdim);
        polyset::tabulate(Pm, ct, pt, d, n, xm);
      });
}

// basis[nderivs][npts][ndofs][vs]; points processed in chunks of `chunk`
// (the reference allocates 2-3x the output in temporaries per call) and the
// chunks are spread over `nthreads` std::threads (reference itself is serial).
int bxref_tabulate(void* h, int nd, const double* x, std::size_t npts,
                   int xdim, double* basis, std::size_t chunk, int nthreads)
{
  const FE& e = *static_cast<FE*>(h);
  return guarded(
      [&]()
      {
        const auto full = e.tabulate_shape(nd, npts);
        if (chunk == 0 or (chunk >= npts and nthreads <= 1))
        {
          e.tabulate(nd, std::span<const double>(x, npts * xdim),
                     {npts, static_cast<std::size_t>(xdim)},
                     std::span<double>(basis,
                                       full[0] * full[1] * full[2] * full[3]));
          return;
        }
        const std::size_t nchunks = (npts + chunk - 1) / chunk;
        const std::size_t row = full[2] * full[3];
        std::string err;
        parallel_for(
            nchunks, nthreads,
            [&](std::size_t c0, std::size_t c1)
            {
              std::vector<double> tmp;
              for (std::size_t c = c0; c < c1; ++c)
              {
                const std::size_t p0 = c * chunk;
                const std::size_t np = std::min(chunk, npts - p0);
                tmp.resize(full[0] * np * row);
                try
                {
                  e.tabulate(nd,
                             std::span<const double>(x + p0 * xdim, np * xdim),
                             {np, static_cast<std::size_t>(xdim)},
                             std::span<double>(tmp));
                }
                catch (const std::exception& ex)
                {
                  err = ex.what();
                  return;
                }
                for (std::size_t d = 0; d < full[0]; ++d)
                  std::memcpy(basis + (d * npts + p0) * row,
                              tmp.data() + d * np * row,
                              np * row * sizeof(double));
              }
            });
        if (!err.empty())
          throw std::runtime_error(err);
      });
}

// out[nc][rows][out_vs] ; returns out_vs through *out_vs.  If out == nullptr
// only the value size is computed (by a one-cell dry run).
int bxref_map(void* h, int pull, const double* U, const double* J,
              const double* detJ, const double* K, std::size_t nc,
              std::size_t rows, std::size_t in_vs, std::size_t gdim,
              std::size_t tdim, double* out, int* out_vs, int nthreads)
{
  const FE& e = *static_cast<FE*>(h);
  return guarded(
      [&]()
      {
        std::string err;
        int ovs = 0;
        parallel_for(
            nc, nthreads,
            [&](std::size_t c0, std::size_t c1)
            {
              try
              {
                const std::size_t n = c1 - c0;
                mds<const double, 3> Um(U + c0 * rows * in_vs, n, rows, in_vs);
                mds<const double, 3> Jm(J + c0 * gdim * tdim, n, gdim, tdim);
                mds<const double, 3> Km(K + c0 * gdim * tdim, n, tdim, gdim);
                std::span<const double> dj(detJ + c0, n);
                auto [data, shape] = pull ? e.pull_back(Um, Jm, dj, Km)
                                          : e.push_forward(Um, Jm, dj, Km);
                if (c0 == 0)
                  ovs = static_cast<int>(shape[2]);
                if (out)
                  std::copy(data.begin(), data.end(),
                            out + c0 * rows * shape[2]);
              }
              catch (const std::exception& ex)
              {
                err = ex.what();
              }
            });
        if (!err.empty())
          throw std::runtime_error(err);
        *out_vs = ovs;
      });
}

// op: 0 T_apply, 1 Tt_apply, 2 Tt_inv_apply, 3 Tinv_apply,
//     4 Tt_apply_right, 5 Tinv_apply_right, 6 T_apply_right,
//     7 Tt_inv_apply_right.   data[nc][cell_size] transformed in place, one
// reference call per cell.
int bxref_T_apply_f64(void* h, int op, double* data, std::size_t nc,
                      std::size_t cell_size, int n, const std::uint32_t* ci,
                      int nthreads)
{
  const FE& e = *static_cast<FE*>(h);
  return guarded(
      [&]()
      {
        parallel_for(nc, nthreads,
                     [&](std::size_t c0, std::size_t c1)
                     {
                       for (std::size_t c = c0; c < c1; ++c)
                       {
                         std::span<double> u(data + c * cell_size, cell_size);
                         switch (op)
                         {
                         case 0: e.T_apply(u, n, ci[c]); break;
                         case 1: e.Tt_apply(u, n, ci[c]); break;
                         case 2: e.Tt_inv_apply(u, n, ci[c]); break;
                         case 3: e.Tinv_apply(u, n, ci[c]); break;
                         case 4: e.Tt_apply_right(u, n, ci[c]); break;
                         case 5: e.Tinv_apply_right(u, n, ci[c]); break;
                         case 6: e.T_apply_right(u, n, ci[c]); break;
                         case 7: e.Tt_inv_apply_right(u, n, ci[c]); break;
                         default: throw std::runtime_error("bad op");
                         }
                       }
                     });
      });
}

int bxref_T_apply_f32(void* h, int op, float* data, std::size_t nc,
                      std::size_t cell_size, int n, const std::uint32_t* ci)
{
  const FE& e = *static_cast<FE*>(h);
  return guarded(
      [&]()
      {
        for (std::size_t c = 0; c < nc; ++c)
        {
          std::span<float> u(data + c * cell_size, cell_size);
          switch (op)
          {
          case 0: e.T_apply(u, n, ci[c]); break;
          case 1: e.Tt_apply(u, n, ci[c]); break;
          case 2: e.Tt_inv_apply(u, n, ci[c]); break;
          case 3: e.Tinv_apply(u, n, ci[c]); break;
          case 4: e.Tt_apply_right(u, n, ci[c]); break;
          case 5: e.Tinv_apply_right(u, n, ci[c]); break;
          case 6: e.T_apply_right(u, n, ci[c]); break;
          case 7: e.Tt_inv_apply_right(u, n, ci[c]); break;
          default: throw std::runtime_error("bad op");
          }
        }
      });
}

int bxref_T_apply_i32(void* h, int op, std::int32_t* data, std::size_t nc,
                      std::size_t cell_size, int n, const std::uint32_t* ci)
{
  const FE& e = *static_cast<FE*>(h);
  return guarded(
      [&]()
      {
        for (std::size_t c = 0; c < nc; ++c)
        {
          std::span<std::int32_t> u(data + c * cell_size, cell_size);
          switch (op)
          {
          case 0: e.T_apply(u, n, ci[c]); break;
          case 1: e.Tt_apply(u, n, ci[c]); break;
          case 2: e.Tt_inv_apply(u, n, ci[c]); break;
          case 3: e.Tinv_apply(u, n, ci[c]); break;
          default: throw std::runtime_error("bad op");
          }
        }
      });
}

// d[nc][ndofs] int32, permuted in place, one reference call per cell
int bxref_permute_i32(void* h, int inverse, std::int32_t* d, std::size_t nc,
                      std::size_t ndofs, const std::uint32_t* ci, int nthreads)
{
  const FE& e = *static_cast<FE*>(h);
  return guarded(
      [&]()
      {
        std::string err;
        parallel_for(nc, nthreads,
                     [&](std::size_t c0, std::size_t c1)
                     {
                       try
                       {
                         for (std::size_t c = c0; c < c1; ++c)
                         {
                           std::span<std::int32_t> s(d + c * ndofs, ndofs);
                           if (inverse)
                             e.permute_inv(s, ci[c]);
                           else
                             e.permute(s, ci[c]);
                         }
                       }
                       catch (const std::exception& ex)
                       {
                         err = ex.what();
                       }
                     });
        if (!err.empty())
          throw std::runtime_error(err);
      });
}
// d[nc][m] int32 (dofs on the closure of one sub-entity), permuted in place, one reference call per cell.
// entity_index < 0: info[c] is the entity's own bits (finite-element.h:942-985 / 997-1035);
// entity_index >= 0: info[c] is the cell's cell_info (finite-element.h:860-930).
int bxref_permute_subentity_closure_i32(void* h, int inverse, std::int32_t* d, std::size_t nc, std::size_t m,
                                        const std::uint32_t* info, int entity_type, int entity_index)
{
  const FE& e = *static_cast<FE*>(h);
  return guarded(
      [&]()
      {
        const auto et = static_cast<basix::cell::type>(entity_type);
        for (std::size_t c = 0; c < nc; ++c)
        {
          std::span<std::int32_t> s(d + c * m, m);
          if (entity_index < 0)
          {
            if (inverse)
              e.permute_subentity_closure_inv(s, info[c], et);
            else
              e.permute_subentity_closure(s, info[c], et);
          }
          else
          {
            if (inverse)
              e.permute_subentity_closure_inv(s, info[c], et, entity_index);
            else
              e.permute_subentity_closure(s, info[c], et, entity_index);
          }
        }
      });
}
// cell::entity_jacobians (cell.cpp:662-689): shape (n_entities, tdim, e_dim)
int bxref_entity_jacobians(int cell, int e_dim, int* shape, double* out)
{
  return guarded(
      [&]()
      {
        auto [data, shp] = cell::entity_jacobians<double>(static_cast<cell::type>(cell), e_dim);
        for (int i = 0; i < 3; ++i)
          shape[i] = static_cast<int>(shp[i]);
        if (out)
          std::copy(data.begin(), data.end(), out);
      });
}

// cell::sub_entity_geometry (cell.cpp): vertex coordinates of sub-entity (dim, index), shape (nvertices, tdim)
int bxref_sub_entity_geometry(int cell, int dim, int index, int* shape, double* out)
{
  return guarded(
      [&]()
      {
        auto [data, shp] = cell::sub_entity_geometry<double>(static_cast<cell::type>(cell), dim, index);
        shape[0] = static_cast<int>(shp[0]);
        shape[1] = static_cast<int>(shp[1]);
        if (out)
          std::copy(data.begin(), data.end(), out);
      });
}

// basix::compute_interpolation_operator (interpolation.cpp:16-116): shape (rows, cols)
int bxref_compute_interpolation_operator(void* from, void* to, int* shape, double* out)
{
  return guarded(
      [&]()
      {
        auto [data, shp] = basix::compute_interpolation_operator<double>(*static_cast<FE*>(from), *static_cast<FE*>(to));
        shape[0] = static_cast<int>(shp[0]);
        shape[1] = static_cast<int>(shp[1]);
        if (out)
          std::copy(data.begin(), data.end(), out);
      });
}

// ---- the reference's float instantiation: FiniteElement<float> (finite-element.cpp:2051), polyset::tabulate<float>
// (polyset.cpp:2995-2997).  Same forwarding as above, single-threaded (parity checks only). ----
using FE32 = FiniteElement<float>;

void* bxref32_create_element(int family, int cell, int degree, int lvariant, int dvariant, int discontinuous,
                             const int* dof_ordering, int n_dof_ordering)
{
  FE32* e = nullptr;
  guarded(
      [&]()
      {
        std::vector<int> ord(dof_ordering, dof_ordering + n_dof_ordering);
        e = new FE32(basix::create_element<float>(static_cast<element::family>(family), static_cast<cell::type>(cell),
                                                  degree, static_cast<element::lagrange_variant>(lvariant),
                                                  static_cast<element::dpc_variant>(dvariant), discontinuous != 0, ord));
      });
  return e;
}

void bxref32_destroy_element(void* h) { delete static_cast<FE32*>(h); }

int bxref32_polyset_tabulate(int cell, int ptype, int d, int n, const float* x, std::size_t npts, int tdim, float* P)
{
  return guarded(
      [&]()
      {
        const auto ct = static_cast<cell::type>(cell);
        const auto pt = static_cast<polyset::type>(ptype);
        mds<float, 3> Pm(P, polyset::nderivs(ct, n), polyset::dim(ct, pt, d), npts);
        mds<const float, 2> xm(x, npts, tdim);
        polyset::tabulate(Pm, ct, pt, d, n, xm);
      });
}

int bxref32_tabulate(void* h, int nd, const float* x, std::size_t npts, int xdim, float* basis)
{
  const FE32& e = *static_cast<FE32*>(h);
  return guarded(
      [&]()
      {
        const auto full = e.tabulate_shape(nd, npts);
        e.tabulate(nd, std::span<const float>(x, npts * xdim), {npts, static_cast<std::size_t>(xdim)},
                   std::span<float>(basis, full[0] * full[1] * full[2] * full[3]));
      });
}

int bxref32_map(void* h, int pull, const float* U, const float* J, const float* detJ, const float* K, std::size_t nc,
                std::size_t rows, std::size_t in_vs, std::size_t gdim, std::size_t tdim, float* out, int* out_vs)
{
  const FE32& e = *static_cast<FE32*>(h);
  return guarded(
      [&]()
      {
        mds<const float, 3> Um(U, nc, rows, in_vs);
        mds<const float, 3> Jm(J, nc, gdim, tdim);
        mds<const float, 3> Km(K, nc, tdim, gdim);
        std::span<const float> dj(detJ, nc);
        auto [data, shape] = pull ? e.pull_back(Um, Jm, dj, Km) : e.push_forward(Um, Jm, dj, Km);
        *out_vs = static_cast<int>(shape[2]);
        if (out)
          std::copy(data.begin(), data.end(), out);
      });
}

int bxref32_T_apply(void* h, int op, float* data, std::size_t nc, std::size_t cell_size, int n, const std::uint32_t* ci)
{
  const FE32& e = *static_cast<FE32*>(h);
  return guarded(
      [&]()
      {
        for (std::size_t c = 0; c < nc; ++c)
        {
          std::span<float> u(data + c * cell_size, cell_size);
          switch (op)
          {
          case 0: e.T_apply(u, n, ci[c]); break;
          case 1: e.Tt_apply(u, n, ci[c]); break;
          case 2: e.Tt_inv_apply(u, n, ci[c]); break;
          case 3: e.Tinv_apply(u, n, ci[c]); break;
          case 4: e.Tt_apply_right(u, n, ci[c]); break;
          case 5: e.Tinv_apply_right(u, n, ci[c]); break;
          case 6: e.T_apply_right(u, n, ci[c]); break;
          case 7: e.Tt_inv_apply_right(u, n, ci[c]); break;
          default: throw std::runtime_error("bad op");
          }
        }
      });
}
} // extern "C"

// Host-side construction of bx_element (tables + device mirror).
//
// Reference behaviour restated here (set-up only, runs once per element on the host):
//   permutation / identity detection        cpp/basix/finite-element.cpp:1260-1295
//   _eperm / _eperm_inv from the matrices   cpp/basix/finite-element.cpp:1302-1347
//   _etrans, _etransT, _etrans_inv, _etrans_invT (M^-1 = M^2 / M^3 for rotations)
//                                           cpp/basix/finite-element.cpp:1349-1445
//   order of application per cell_info      cpp/basix/finite-element.h:1703-1837
#include "element.cuh"
#include "jets.cuh"

#include <algorithm>
#include <cstring>
#include <cmath>
#include <limits>
#include <memory>
#include <numeric>

namespace bx
{
namespace
{
// ---- tiny dense helpers (n <= a few dozen) -------------------------------------------------------
using Mat = std::vector<double>; // row-major n x n

Mat mat_identity(int n)
{
  Mat I(size_t(n) * n, 0.0);
  for (int i = 0; i < n; ++i)
    I[size_t(i) * n + i] = 1.0;
  return I;
}
Mat mat_mul(const Mat& A, const Mat& B, int n)
{
  Mat C(size_t(n) * n, 0.0);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < n; ++k)
    {
      const double a = A[size_t(i) * n + k];
      if (a == 0.0)
        continue;
      for (int j = 0; j < n; ++j)
        C[size_t(i) * n + j] += a * B[size_t(k) * n + j];
    }
  return C;
}
Mat mat_transpose(const Mat& A, int n)
{
  Mat T(size_t(n) * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j)
      T[size_t(j) * n + i] = A[size_t(i) * n + j];
  return T;
}

template <typename T>
T* upload(const std::vector<T>& v)
{
  if (v.empty())
    return nullptr;
  T* d = nullptr;
  BX_CUDA(cudaMalloc(&d, v.size() * sizeof(T)));
  BX_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

int num_states(int mask) { return mask + 1; }
} // namespace

namespace
{
void factor_tensor_product(bx_element& e)
{
  const int tdim = e.tdim, n1 = e.superdegree + 1, N = e.N;
  constexpr int kMaxFun = 16;
  if (n1 > 9)
    return;
  auto& tp = e.tp;
  std::vector<std::vector<double>> funs[3];
  tp.ijk.assign(N, 0);
  tp.scale.assign(N, 0.0);
  double cmax = 0;
  for (double v : e.Cp)
    cmax = std::max(cmax, std::abs(v));
  const double tol = 1e-13 * cmax;
  const int stride[3] = {tdim == 2 ? n1 : n1 * n1, tdim == 2 ? 1 : n1, 1};
  for (int n = 0; n < N; ++n)
  {
    const double* T = e.Cp.data() + size_t(n) * e.Kpad;
    int kmax = 0;
    for (int k = 1; k < e.K; ++k)
      if (std::abs(T[k]) > std::abs(T[kmax]))
        kmax = k;
    const double T0 = T[kmax];
    if (std::abs(T0) <= tol)
      continue; // zero row: scale 0
    int p0[3] = {0, 0, 0};
    {
      int rem = kmax;
      for (int a = 0; a < tdim; ++a)
      {
        p0[a] = rem / stride[a];
        rem -= p0[a] * stride[a];
      }
    }
    // fibres through the pivot, normalised by the pivot
    std::vector<double> f[3];
    for (int a = 0; a < tdim; ++a)
    {
      f[a].resize(n1);
      for (int q = 0; q < n1; ++q)
      {
        int k = 0;
        for (int b = 0; b < tdim; ++b)
          k += (b == a ? q : p0[b]) * stride[b];
        f[a][q] = T[k] / T0;
      }
    }
    // rank-1 check
    for (int k = 0; k < e.K; ++k)
    {
      int rem = k;
      double r = T0;
      for (int a = 0; a < tdim; ++a)
      {
        const int q = rem / stride[a];
        rem -= q * stride[a];
        r *= f[a][q];
      }
      if (std::abs(r - T[k]) > tol)
        return; // not a tensor product
    }
    int id[3] = {0, 0, 0};
    for (int a = 0; a < tdim; ++a)
    {
      int found = -1;
      for (size_t j = 0; j < funs[a].size() and found < 0; ++j)
      {
        double dmax = 0;
        for (int q = 0; q < n1; ++q)
          dmax = std::max(dmax, std::abs(funs[a][j][q] - f[a][q]));
        if (dmax <= 1e-10)
          found = int(j);
      }
      if (found < 0)
      {
        if (int(funs[a].size()) >= kMaxFun)
          return;
        funs[a].push_back(f[a]);
        found = int(funs[a].size()) - 1;
      }
      id[a] = found;
    }
    tp.ijk[n] = id[0] | (id[1] << 8) | (id[2] << 16);
    tp.scale[n] = T0;
  }
  tp.fstride = 1;
  for (int a = 0; a < tdim; ++a)
  {
    tp.nfun[a] = int(funs[a].size());
    tp.fstride = std::max(tp.fstride, tp.nfun[a]);
  }
  tp.A.assign(size_t(tdim) * tp.fstride * n1, 0.0);
  for (int a = 0; a < tdim; ++a)
    for (int j = 0; j < tp.nfun[a]; ++j)
      for (int q = 0; q < n1; ++q)
        tp.A[(size_t(a) * tp.fstride + j) * n1 + q] = funs[a][j][q] * std::sqrt(double(2 * q + 1));
  tp.valid = true;
  // lexicographic (i, j, k) -> column table for the staged kernel, when the columns are a full tensor grid
  {
    const int nf = tp.nfun[0];
    bool grid = nf >= 2;
    for (int a = 1; a < tdim; ++a)
      grid = grid and tp.nfun[a] == nf;
    int total = 1;
    for (int a = 0; a < tdim; ++a)
      total *= nf;
    grid = grid and total == N;
    if (grid)
    {
      std::vector<double2> order(size_t(N), make_double2(0.0, -1.0)); // .y: -1 = unset, then the column's raw bits
      for (int n = 0; n < N and grid; ++n)
      {
        const int i = tp.ijk[n] & 255, j = (tp.ijk[n] >> 8) & 255, k = (tp.ijk[n] >> 16) & 255;
        const int idx = tdim == 2 ? i * nf + j : (i * nf + j) * nf + k;
        if (idx >= N or order[idx].y >= 0.0)
          grid = false;
        else
          order[idx] = make_double2(tp.scale[n], double(n));
      }
      for (double2& o : order) // the kernel reads the column as an integer without a conversion instruction
      {
        const long long col = (long long)(o.y);
        std::memcpy(&o.y, &col, sizeof(col));
      }
      if (grid)
      {
        tp.nf = nf;
        tp.order = order;
      }
    }
  }
  if (e.on_device)
  {
    tp.d_A = upload(tp.A);
    tp.d_ijk = upload(tp.ijk);
    tp.d_scale = upload(tp.scale);
    if (!tp.order.empty())
      tp.d_order = upload(tp.order);
  }
}
} // namespace

// seq -> (basis slot, norm) in the order jets.cuh emits basis functions
void jets::production_order(int tdim, int n, int* slot, double* norm)
{
  int s = 0;
  if (tdim == 1)
  {
    for (int p = 0; p <= n; ++p, ++s)
    {
      slot[s] = p;
      norm[s] = std::sqrt(double(2 * p + 1));
    }
  }
  else if (tdim == 2)
  {
    for (int p = 0; p <= n; ++p)
      for (int q = 0; q <= n - p; ++q, ++s)
      {
        slot[s] = idx2(q, p);
        norm[s] = std::sqrt((p + 0.5) * (p + q + 1)) * 2;
      }
  }
  else
  {
    for (int p = 0; p <= n; ++p)
      for (int q = 0; q <= n - p; ++q)
        for (int r = 0; r <= n - p - q; ++r, ++s)
        {
          slot[s] = idx3(r, q, p);
          norm[s] = std::sqrt(2 * (p + 0.5) * (p + q + 1.0) * (p + q + r + 1.5)) * 2;
        }
  }
}

// ---- sub-entity closure permutations ------------------------------------------------------------------
using StepMaps = std::array<std::array<std::vector<int>, 2>, 8>; // [sub-entity type][i]: new[r] = old[map[r]]
void build_closure_tables(bx_element& e, const StepMaps& fwd, const StepMaps& inv)
{
  // finite-element.cpp:1449-1705 builds, for edge 0 and for one face of each face type, the permutation of the dofs
  // on the entity's CLOSURE (vertex blocks, edge blocks, interior) under a rotation / reflection of the entity,
  // and finite-element.h:942-1035 applies them per entity_info: forward = rotate x r then reflect, inverse =
  // reflect then inverse-rotate x r.  Here the three generators are built the same way as index maps and the net
  // gather of every state is composed once.
  if (e.perms)
  {
    struct Conn
    {
      int cell, sub;
      std::vector<int> v, ed, f;
    };
    // cell.cpp:143-292 sub_entity_connectivity of edge 0 / the face used by the reference (face_n)
    static const Conn conns[] = {
        {BX_CELL_TRIANGLE, BX_CELL_INTERVAL, {1, 2}, {0}, {}},
        {BX_CELL_QUADRILATERAL, BX_CELL_INTERVAL, {0, 1}, {0}, {}},
        {BX_CELL_TETRAHEDRON, BX_CELL_INTERVAL, {2, 3}, {0}, {}},
        {BX_CELL_HEXAHEDRON, BX_CELL_INTERVAL, {0, 1}, {0}, {}},
        {BX_CELL_PRISM, BX_CELL_INTERVAL, {0, 1}, {0}, {}},
        {BX_CELL_PYRAMID, BX_CELL_INTERVAL, {0, 1}, {0}, {}},
        {BX_CELL_TETRAHEDRON, BX_CELL_TRIANGLE, {1, 2, 3}, {0, 1, 2}, {0}},
        {BX_CELL_PRISM, BX_CELL_TRIANGLE, {0, 1, 2}, {0, 1, 3}, {0}},
        {BX_CELL_PYRAMID, BX_CELL_TRIANGLE, {0, 1, 4}, {0, 2, 4}, {1}},
        {BX_CELL_HEXAHEDRON, BX_CELL_QUADRILATERAL, {0, 1, 2, 3}, {0, 1, 3, 5}, {0}},
        {BX_CELL_PRISM, BX_CELL_QUADRILATERAL, {0, 1, 3, 4}, {0, 2, 4, 6}, {1}},
        {BX_CELL_PYRAMID, BX_CELL_QUADRILATERAL, {0, 1, 2, 3}, {0, 1, 3, 5}, {0}},
    };
    using Idx = std::vector<int>;
    // data[offset + i] <- data[offset + map[i]] (the effect of precompute::apply_permutation with the prepared map)
    auto apply_at = [](const Idx& map, Idx& data, int offset)
    {
      Idx tmp(map.size());
      for (size_t i = 0; i < map.size(); ++i)
        tmp[i] = data[offset + map[i]];
      std::copy(tmp.begin(), tmp.end(), data.begin() + offset);
    };
    auto apply_all = [&](const Idx& map, Idx& data) { apply_at(map, data, 0); };
    for (const Conn& cn : conns)
    {
      if (cn.cell != e.cell_type)
        continue;
      std::vector<std::vector<Idx>> dofs(3);
      int n = 0;
      const std::vector<int>* ents[3] = {&cn.v, &cn.ed, &cn.f};
      for (int dim = 0; dim < 3; ++dim)
        for (int ent : *ents[dim])
        {
          Idx blk(e.edofs[dim][ent].size());
          for (int& b : blk)
            b = n++;
          dofs[dim].push_back(blk);
        }
      auto cat = [](std::initializer_list<const Idx*> blocks)
      {
        Idx out;
        for (const Idx* b : blocks)
          out.insert(out.end(), b->begin(), b->end());
        return out;
      };
      const auto &V = dofs[0], &E = dofs[1], &F = dofs[2];
      const bool move = !e.identity;
      const Idx& t1 = fwd[BX_CELL_INTERVAL][0];
      Idx rot, rot_inv, ref;
      if (cn.sub == BX_CELL_INTERVAL)
      {
        ref = cat({&V[1], &V[0], &E[0]});
        if (move and !E[0].empty())
          apply_at(t1, ref, E[0][0]);
      }
      else if (cn.sub == BX_CELL_TRIANGLE)
      {
        rot = cat({&V[1], &V[2], &V[0], &E[1], &E[2], &E[0], &F[0]});
        rot_inv = cat({&V[2], &V[0], &V[1], &E[2], &E[0], &E[1], &F[0]});
        ref = cat({&V[0], &V[2], &V[1], &E[0], &E[2], &E[1], &F[0]});
        if (move and !E[0].empty())
        {
          apply_at(t1, rot, E[0][0]);
          apply_at(t1, rot, E[1][0]);
          apply_at(t1, ref, E[0][0]);
          apply_at(t1, rot_inv, E[1][0]);
          apply_at(t1, rot_inv, E[2][0]);
        }
        if (move and !F[0].empty())
        {
          apply_at(fwd[BX_CELL_TRIANGLE][0], rot_inv, F[0][0]);
          apply_at(fwd[BX_CELL_TRIANGLE][1], ref, F[0][0]);
          apply_at(inv[BX_CELL_TRIANGLE][0], rot, F[0][0]);
        }
      }
      else
      {
        rot = cat({&V[1], &V[3], &V[0], &V[2], &E[2], &E[0], &E[3], &E[1], &F[0]});
        rot_inv = cat({&V[2], &V[0], &V[3], &V[1], &E[1], &E[3], &E[0], &E[2], &F[0]});
        ref = cat({&V[0], &V[2], &V[1], &V[3], &E[1], &E[0], &E[3], &E[2], &F[0]});
        if (move and !E[0].empty())
        {
          apply_at(t1, rot, E[1][0]);
          apply_at(t1, rot, E[2][0]);
          apply_at(t1, rot_inv, E[0][0]);
          apply_at(t1, rot_inv, E[3][0]);
        }
        if (move and !F[0].empty())
        {
          apply_at(fwd[BX_CELL_QUADRILATERAL][0], rot_inv, F[0][0]);
          apply_at(fwd[BX_CELL_QUADRILATERAL][1], ref, F[0][0]);
          apply_at(inv[BX_CELL_QUADRILATERAL][0], rot, F[0][0]);
        }
      }
      bx_element::Closure& cl = e.closure[cn.sub];
      cl.size = n;
      cl.states = cn.sub == BX_CELL_INTERVAL ? 2 : 8;
      cl.src.assign(size_t(2) * cl.states * n, 0);
      for (int dir = 0; dir < 2; ++dir)
        for (int st = 0; st < cl.states; ++st)
        {
          Idx net(n);
          for (int i = 0; i < n; ++i)
            net[i] = i;
          if (cn.sub == BX_CELL_INTERVAL)
          {
            if (st & 1)
              apply_all(ref, net);
          }
          else if (dir == 0)
          {
            for (int r = 0; r < ((st >> 1) & 3); ++r)
              apply_all(rot, net);
            if (st & 1)
              apply_all(ref, net);
          }
          else
          {
            if (st & 1)
              apply_all(ref, net);
            for (int r = 0; r < ((st >> 1) & 3); ++r)
              apply_all(rot_inv, net);
          }
          for (int i = 0; i < n; ++i)
            cl.src[(size_t(dir) * cl.states + st) * n + i] = int16_t(net[i]);
        }
    }
  }
  if (e.on_device)
    for (auto& cl : e.closure)
      if (cl.size > 0)
        cl.d_src = upload(cl.src);
}

int tabulate_path(const bx_element& e, int nd);
namespace
{
// The tabulate operands of derivative orders 0..2 (derivative-operator kernel: composed per order) are built and
// uploaded at creation, so device-pointer calls only launch kernels (bx_element_prepare does the same for higher orders).
void prepare_tabulate(const bx_element& e)
{
  if (!e.on_device)
    return;
  for (int nd = 0; nd <= 2; ++nd)
    (void)tabulate_path(e, nd);
}
} // namespace

bx_element* element_from_desc(const bx_element_desc& d)
{
  if (d.cell_type < BX_CELL_POINT or d.cell_type > BX_CELL_PYRAMID)
    fail(BX_ERR_INVALID_ARG, "create_element: unknown cell type");
  if (d.ndofs <= 0 or d.value_size <= 0 or !d.coeffs)
    fail(BX_ERR_INVALID_ARG, "create_element: empty coefficient matrix");
  auto e = std::make_unique<bx_element>();
  e->cell_type = d.cell_type;
  e->polyset_type = d.polyset_type;
  e->degree = d.degree;
  e->superdegree = d.embedded_superdegree;
  e->map_type = d.map_type;
  e->ndofs = d.ndofs;
  e->vs = d.value_size;
  e->tdim = cell_tdim(d.cell_type);
  e->psize = polyset_dim(d.cell_type, d.polyset_type, d.embedded_superdegree);
  {
    // device is optional at construction: host tables are always built
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) == cudaSuccess and ndev > 0)
    {
      BX_CUDA(cudaGetDevice(&e->device));
      e->on_device = true;
    }
    else
      cudaGetLastError();
  }

  const int ndofs = e->ndofs, vs = e->vs, psize = e->psize;
  e->coeffs.assign(d.coeffs, d.coeffs + size_t(ndofs) * vs * psize);
  if (d.n_dof_ordering != 0)
  {
    if (d.n_dof_ordering != ndofs or !d.dof_ordering)
      fail(BX_ERR_INVALID_ARG, "Incorrect number of dofs in ordering.");
    e->dof_ordering.assign(d.dof_ordering, d.dof_ordering + ndofs);
    std::vector<int> check(ndofs, 0);
    for (int q : e->dof_ordering)
    {
      if (q < 0 or q >= ndofs)
        fail(BX_ERR_INVALID_ARG, "Out of range: dof_ordering.");
      check[q] += 1;
    }
    for (int q : check)
      if (q != 1)
        fail(BX_ERR_INVALID_ARG, "Dof ordering not a permutation.");
  }

  // ---- contraction operand: rows = output column (dof_ordering[dof]*vs + j), zero padded to 8 x 4 ---
  e->N = ndofs * vs;
  e->K = psize;
  e->Npad = int(ceil_div(e->N, 8) * 8);
  e->Kpad = int(ceil_div(e->K, 4) * 4);
  e->Cp.assign(size_t(e->Npad) * e->Kpad, 0.0);
  for (int dof = 0; dof < ndofs; ++dof)
  {
    const int tgt = e->dof_ordering.empty() ? dof : e->dof_ordering[dof];
    for (int j = 0; j < vs; ++j)
      for (int k = 0; k < psize; ++k)
        e->Cp[size_t(tgt * vs + j) * e->Kpad + k] = e->coeffs[size_t(dof) * vs * psize + size_t(j) * psize + k];
  }
  if (e->on_device)
    e->d_Cp = upload(e->Cp);
  e->dop_cache = dop_cache_create();

  // ---- fused simplex kernel operand (tabulate_fused.cuh): Cs[Npad8][KS], column s = production-order
  //      position of the basis function, scaled by its norm; only when all output columns fit ----
  if (e->polyset_type == BX_POLYSET_STANDARD
      and (e->cell_type == BX_CELL_INTERVAL or e->cell_type == BX_CELL_TRIANGLE or e->cell_type == BX_CELL_TETRAHEDRON))
  {
    const int K = e->K;
    const int npad = int(ceil_div(K, 8) * 8);
    const int kpad = e->Kpad;
    const int ks = kpad + ((4 - kpad % 16) + 16) % 16;
    std::vector<int> slot(K);
    std::vector<double> norm(K);
    jets::production_order(e->tdim, e->superdegree, slot.data(), norm.data());
    if (K <= 10 and size_t(e->N) * K <= 4096)
    {
      // register-path operand (tabulate_small.cu): any number of output columns
      e->Cq.assign(size_t(e->N) * K, 0.0);
      for (int n = 0; n < e->N; ++n)
        for (int s = 0; s < K; ++s)
          e->Cq[size_t(n) * K + s] = e->Cp[size_t(n) * e->Kpad + slot[s]] * norm[s];
      if (e->on_device)
        e->d_Cq = upload(e->Cq);
    }
    if (e->N <= npad)
    {
      e->Cs.assign(size_t(npad) * ks, 0.0);
      for (int n = 0; n < e->N; ++n)
        for (int s = 0; s < K; ++s)
          e->Cs[size_t(n) * ks + s] = e->Cp[size_t(n) * e->Kpad + slot[s]] * norm[s];
      if (e->on_device)
        e->d_Cs = upload(e->Cs);
    }
  }

  // ---- tensor-product factorisation of the rows of Cp (quad / hex): Cp[n][(px,py,pz)] =
  //      s_n a_i[px] b_j[py] c_k[pz].  Detected numerically; any row that is not rank-1 disables the path.
  if (e->polyset_type == BX_POLYSET_STANDARD and (e->cell_type == BX_CELL_QUADRILATERAL or e->cell_type == BX_CELL_HEXAHEDRON))
    factor_tensor_product(*e);

  // ---- entity dofs -----------------------------------------------------------------------------------
  if (!d.entity_dof_offsets or !d.entity_dofs)
    fail(BX_ERR_INVALID_ARG, "create_element: entity dofs missing");
  e->edofs.resize(e->tdim + 1);
  {
    int ent = 0;
    for (int dim = 0; dim <= e->tdim; ++dim)
    {
      const int ne = cell_num_entities(d.cell_type, dim);
      e->edofs[dim].resize(ne);
      for (int i = 0; i < ne; ++i, ++ent)
      {
        const int b = d.entity_dof_offsets[ent], f = d.entity_dof_offsets[ent + 1];
        if (b < 0 or f < b)
          fail(BX_ERR_INVALID_ARG, "create_element: bad entity dof offsets");
        for (int k = b; k < f; ++k)
        {
          if (d.entity_dofs[k] < 0 or d.entity_dofs[k] >= ndofs)
            fail(BX_ERR_INVALID_ARG, "create_element: entity dof out of range");
          e->edofs[dim][i].push_back(d.entity_dofs[k]);
        }
      }
    }
  }

  // ---- interpolation data (optional) ----------------------------------------------------------------
  if (d.interpolation_matrix and d.points and d.npts > 0)
  {
    e->interp_npts = d.npts;
    e->points.assign(d.points, d.points + size_t(d.npts) * e->tdim);
    const int K = e->vs * d.npts;
    e->interp_M.assign(d.interpolation_matrix, d.interpolation_matrix + size_t(e->ndofs) * K);
    // device operands, one per value layout: B[n][k'] with k' the position of (component j, point p) in the layout
    const int kp = int(ceil_div(K, 48) * 48); // k-steps of 4, unrolled by 2, 4 or 6 in interpolate.cu: zero columns
    e->interp_kstride = kp % 8 == 4 ? kp : kp + 4;
    e->interp_npad = int(ceil_div(e->ndofs, 8) * 8);
    if (e->on_device)
      for (int layout = 0; layout < 2; ++layout)
      {
        std::vector<double> B(size_t(e->interp_npad) * e->interp_kstride, 0.0);
        for (int n = 0; n < e->ndofs; ++n)
          for (int j = 0; j < e->vs; ++j)
            for (int p = 0; p < d.npts; ++p)
              B[size_t(n) * e->interp_kstride + (layout == 0 ? j * d.npts + p : p * e->vs + j)]
                  = e->interp_M[size_t(n) * K + j * d.npts + p];
        e->d_interp[layout] = upload(B);
      }
    // blocked component-major operand of the fused pull_back + interpolation (interpolate.cu): Piola elements only
    if (e->on_device and e->vs == e->tdim and (e->map_type == BX_MAP_COVARIANT_PIOLA or e->map_type == BX_MAP_CONTRAVARIANT_PIOLA))
    {
      e->interp_pb = int(ceil_div(d.npts, 8) * 8);
      const int kb = e->vs * e->interp_pb;
      e->interp_kstride_blocked = kb % 8 == 4 ? kb : kb + 4;
      std::vector<double> B(size_t(e->interp_npad) * e->interp_kstride_blocked, 0.0);
      for (int n = 0; n < e->ndofs; ++n)
        for (int j = 0; j < e->vs; ++j)
          for (int p = 0; p < d.npts; ++p)
            B[size_t(n) * e->interp_kstride_blocked + j * e->interp_pb + p] = e->interp_M[size_t(n) * K + j * d.npts + p];
      e->d_interp_blocked = upload(B);
    }
  }

  // ---- entity transformations ------------------------------------------------------------------------
  auto take = [&](int cls, const double* p, int n, int ntrans)
  {
    e->etrans_dim[cls] = p ? n : -1;
    if (p and n > 0)
      e->etrans[cls].assign(p, p + size_t(ntrans) * n * n);
  };
  for (auto& v : e->etrans_dim)
    v = -1;
  take(BX_CELL_INTERVAL, d.etrans_interval, d.etrans_interval_dim, 1);
  take(BX_CELL_TRIANGLE, d.etrans_triangle, d.etrans_triangle_dim, 2);
  take(BX_CELL_QUADRILATERAL, d.etrans_quadrilateral, d.etrans_quadrilateral_dim, 2);

  // Permutation / identity detection, same test as finite-element.cpp:1260-1295
  e->perms = true;
  e->identity = true;
  {
    const double eps = 10.0 * d.degree * d.degree * std::numeric_limits<double>::epsilon();
    for (int cls : {BX_CELL_INTERVAL, BX_CELL_TRIANGLE, BX_CELL_QUADRILATERAL})
    {
      const int n = e->etrans_dim[cls];
      if (n <= 0)
        continue;
      const int ntrans = cls == BX_CELL_INTERVAL ? 1 : 2;
      for (int i = 0; e->perms and i < ntrans; ++i)
      {
        const double* M = e->etrans[cls].data() + size_t(i) * n * n;
        for (int row = 0; row < n; ++row)
        {
          double rmin = 0, rmax = 0, rtot = 0;
          for (int k = 0; k < n; ++k)
          {
            const double r = M[size_t(row) * n + k];
            rmin = std::min(r, rmin);
            rmax = std::max(r, rmax);
            rtot += r;
          }
          if ((n != 1 and std::abs(rmin) > eps) or std::abs(rmax - 1.0) > eps or std::abs(rtot - 1.0) > eps)
          {
            e->perms = false;
            e->identity = false;
            break;
          }
          if (std::abs(M[size_t(row) * n + row] - 1) > eps)
            e->identity = false;
        }
      }
      if (!e->perms)
        break;
    }
  }

  e->dof_slot.assign(ndofs, -1);
  e->dof_loc.assign(ndofs, 0);
  e->lo_dof = e->hi_dof = 0;
  if (e->identity or e->tdim < 2)
  {
    if (e->tdim >= 2)
      build_closure_tables(*e, {}, {}); // vertex and edge blocks still move with the entity
    prepare_tabulate(*e);
    return e.release();
  }

  // ---- slots in the reference's processing order ------------------------------------------------------
  const int nfaces = e->tdim == 3 ? int(e->edofs[2].size()) : 0;
  const int face_start = e->tdim == 3 ? 3 * nfaces : 0;
  int dofstart = 0;
  for (auto& v : e->edofs[0])
    dofstart += int(v.size());
  auto add_slot = [&](int shift, int mask, int cls, const std::vector<int>& dofs)
  {
    bx_element::Slot s;
    s.shift = shift;
    s.mask = mask;
    s.cls = cls;
    s.first = dofstart;
    s.dofs = dofs;
    s.dim = int(dofs.size());
    dofstart += s.dim;
    if (s.dim == 0)
      return;
    if (e->etrans_dim[cls] != s.dim)
      fail(BX_ERR_INVALID_ARG, "create_element: entity transformation size does not match entity dofs");
    if (!e->perms)
    {
      // transform_data walks contiguous blocks (running dofstart), finite-element.h:1772-1835
      s.dofs.resize(s.dim);
      std::iota(s.dofs.begin(), s.dofs.end(), s.first);
      if (s.first + s.dim > ndofs)
        fail(BX_ERR_INVALID_ARG, "create_element: entity dof blocks exceed the element dimension");
    }
    e->slots.push_back(std::move(s));
  };
  for (size_t ed = 0; ed < e->edofs[1].size(); ++ed)
    add_slot(face_start + int(ed), 1, BX_CELL_INTERVAL, e->edofs[1][ed]);
  if (e->tdim == 3)
    for (int f = 0; f < nfaces; ++f)
      add_slot(3 * f, 7, cell_sub_entity_type(d.cell_type, 2, f), e->edofs[2][f]);

  int lo = ndofs, hi = 0;
  for (size_t s = 0; s < e->slots.size(); ++s)
    for (int l = 0; l < e->slots[s].dim; ++l)
    {
      const int dof = e->slots[s].dofs[l];
      e->dof_slot[dof] = int16_t(s);
      e->dof_loc[dof] = int16_t(l);
      lo = std::min(lo, dof);
      hi = std::max(hi, dof + 1);
    }
  e->lo_dof = lo < hi ? lo : 0;
  e->hi_dof = lo < hi ? hi : 0;

  // ---- per-class step operators -----------------------------------------------------------------------
  // index maps (permutation elements): new[r] = old[map[r]]
  StepMaps fwd, inv;
  // matrices (general elements): new = M old ; [cls][i] for the four stored flavours
  std::array<std::array<Mat, 2>, 8> M_, MT_, Minv_, MinvT_;
  for (int cls : {BX_CELL_INTERVAL, BX_CELL_TRIANGLE, BX_CELL_QUADRILATERAL})
  {
    const int n = e->etrans_dim[cls];
    if (n <= 0)
      continue;
    const int ntrans = cls == BX_CELL_INTERVAL ? 1 : 2;
    for (int i = 0; i < ntrans; ++i)
    {
      const double* T = e->etrans[cls].data() + size_t(i) * n * n;
      if (e->perms)
      {
        std::vector<int> perm(n, 0), iperm(n, 0);
        for (int row = 0; row < n; ++row)
          for (int col = 0; col < n; ++col)
            if (T[size_t(row) * n + col] > 0.5)
            {
              perm[row] = col;
              iperm[col] = row;
              break;
            }
        fwd[cls][i] = perm;
        inv[cls][i] = iperm;
      }
      else
      {
        Mat M(T, T + size_t(n) * n);
        Mat Minv = M;
        if (cls == BX_CELL_QUADRILATERAL and i == 0)
          Minv = mat_mul(mat_mul(M, M, n), M, n);
        else if (cls == BX_CELL_TRIANGLE and i == 0)
          Minv = mat_mul(M, M, n);
        M_[cls][i] = M;
        MT_[cls][i] = mat_transpose(M, n);
        Minv_[cls][i] = Minv;
        MinvT_[cls][i] = mat_transpose(Minv, n);
      }
    }
  }

  build_closure_tables(*e, fwd, inv);

  // ---- compose net operator per (op, slot/class, state) --------------------------------------------
  // op kinds: 0 T, 1 Tt, 2 Tt_inv, 3 Tinv.  `post` (reflect after rotate) for 1 and 3.
  std::vector<int4> slot_info(e->slots.size());
  std::vector<int2> slot_mat(e->slots.size());
  int gsize = 0;
  for (size_t s = 0; s < e->slots.size(); ++s)
  {
    slot_info[s] = make_int4(e->slots[s].shift, e->slots[s].mask, e->slots[s].dim, gsize);
    gsize += num_states(e->slots[s].mask) * e->slots[s].dim;
  }
  std::array<int, 8> cls_base{};
  int msize = 0;
  if (!e->perms)
    for (int cls : {BX_CELL_INTERVAL, BX_CELL_TRIANGLE, BX_CELL_QUADRILATERAL})
    {
      const int n = e->etrans_dim[cls];
      cls_base[cls] = msize;
      if (n > 0)
        msize += (cls == BX_CELL_INTERVAL ? 2 : 8) * mat_state_stride(n);
    }
  for (size_t s = 0; s < e->slots.size(); ++s)
    slot_mat[s] = make_int2(e->slots[s].first, cls_base[e->slots[s].cls]);

  std::vector<int16_t> gather(e->perms ? size_t(4) * gsize : 0);
  std::vector<double> mats(e->perms ? 0 : size_t(4) * msize);
  for (int op = 0; op < 4; ++op)
  {
    const bool post = (op == 1 or op == 3);
    if (e->perms)
    {
      const auto& steps = (op == 0 or op == 2) ? fwd : inv; // _eperm for T/Tt_inv, _eperm_inv for Tt/Tinv
      for (size_t s = 0; s < e->slots.size(); ++s)
      {
        const auto& sl = e->slots[s];
        const int n = sl.dim;
        for (int state = 0; state < num_states(sl.mask); ++state)
        {
          std::vector<int> src(n);
          std::iota(src.begin(), src.end(), 0);
          auto apply = [&](const std::vector<int>& m)
          {
            std::vector<int> t(n);
            for (int r = 0; r < n; ++r)
              t[r] = src[m[r]];
            src.swap(t);
          };
          if (sl.mask == 1)
          {
            if (state & 1)
              apply(steps[sl.cls][0]);
          }
          else
          {
            const int refl = state & 1, rots = (state >> 1) & 3;
            if (!post and refl)
              apply(steps[sl.cls][1]);
            for (int r = 0; r < rots; ++r)
              apply(steps[sl.cls][0]);
            if (post and refl)
              apply(steps[sl.cls][1]);
          }
          for (int r = 0; r < n; ++r)
            gather[size_t(op) * gsize + slot_info[s].w + state * n + r] = int16_t(sl.dofs[src[r]]);
        }
      }
    }
    else
    {
      const auto& steps = op == 0 ? M_ : op == 1 ? MT_ : op == 2 ? MinvT_ : Minv_;
      for (int cls : {BX_CELL_INTERVAL, BX_CELL_TRIANGLE, BX_CELL_QUADRILATERAL})
      {
        const int n = e->etrans_dim[cls];
        if (n <= 0)
          continue;
        const int nst = cls == BX_CELL_INTERVAL ? 2 : 8;
        for (int state = 0; state < nst; ++state)
        {
          Mat A = mat_identity(n);
          if (cls == BX_CELL_INTERVAL)
          {
            if (state & 1)
              A = mat_mul(steps[cls][0], A, n);
          }
          else
          {
            const int refl = state & 1, rots = (state >> 1) & 3;
            if (!post and refl)
              A = mat_mul(steps[cls][1], A, n);
            for (int r = 0; r < rots; ++r)
              A = mat_mul(steps[cls][0], A, n);
            if (post and refl)
              A = mat_mul(steps[cls][1], A, n);
          }
          double* dst = mats.data() + size_t(op) * msize + cls_base[cls] + size_t(state) * mat_state_stride(n);
          for (int row = 0; row < n; ++row)
            std::copy(A.begin() + size_t(row) * n, A.begin() + size_t(row + 1) * n, dst + size_t(row) * mat_row_stride(n));
        }
      }
    }
  }
  // ---- column band of every operator row (matrix elements) -----------------------------------------------
  // Legendre-variant moment dofs transform block-diagonally (one block per polynomial degree), so most of a
  // dense dim x dim operator is zero; the kernel only walks [first, last) of each row.  The band is the union
  // over the four operator kinds and all states, so one table serves every T-flavour.
  std::vector<int2> band;
  std::vector<int> slot_band(e->slots.size(), 0);
  if (!e->perms)
  {
    std::array<int, 8> cls_band{};
    for (int cls : {BX_CELL_INTERVAL, BX_CELL_TRIANGLE, BX_CELL_QUADRILATERAL})
    {
      const int n = e->etrans_dim[cls];
      if (n <= 0)
        continue;
      cls_band[cls] = int(band.size());
      const int nst = cls == BX_CELL_INTERVAL ? 2 : 8;
      double amax = 0.0;
      for (int op = 0; op < 4; ++op)
        for (int state = 0; state < nst; ++state)
          for (int k = 0; k < n * mat_row_stride(n); ++k)
            amax = std::max(amax, std::abs(mats[size_t(op) * msize + cls_base[cls] + size_t(state) * mat_state_stride(n) + k]));
      for (int row = 0; row < n; ++row)
      {
        int lo = n, hi = 0;
        for (int op = 0; op < 4; ++op)
          for (int state = 0; state < nst; ++state)
          {
            const double* r = mats.data() + size_t(op) * msize + cls_base[cls] + size_t(state) * mat_state_stride(n)
                              + size_t(row) * mat_row_stride(n);
            for (int j = 0; j < n; ++j)
              if (std::abs(r[j]) > 1e-14 * amax)
              {
                lo = std::min(lo, j);
                hi = std::max(hi, j + 1);
              }
          }
        band.push_back(lo < hi ? make_int2(lo, hi) : make_int2(0, 0));
      }
    }
    for (size_t s = 0; s < e->slots.size(); ++s)
      slot_band[s] = cls_band[e->slots[s].cls];
  }
  e->h_band = band;
  e->h_slot_band = slot_band;
  e->band_size = int(band.size());
  e->gather_size = gsize;
  e->mats_size = msize;
  e->h_slot_info = slot_info;
  e->h_slot_mat = slot_mat;
  e->h_gather = gather;
  e->h_mats = mats;
  if (e->on_device)
  {
    e->d_dof_slot = upload(e->dof_slot);
    e->d_dof_loc = upload(e->dof_loc);
    e->d_slot_info = upload(slot_info);
    e->d_slot_mat = upload(slot_mat);
    e->d_gather = upload(gather);
    e->d_mats = upload(mats);
    e->d_band = upload(band);
    if (e->perms and gsize < 65536)
    {
      // packed per-dof descriptor of the permutation kernel (dof_transform_tma.cuh); mask 0 = never moves
      std::vector<uint32_t> dinfo(ndofs, 0u);
      bool ok = true;
      for (size_t s = 0; s < e->slots.size(); ++s)
      {
        const auto& sl = e->slots[s];
        ok = ok and sl.dim < 256 and sl.shift < 32;
        for (int l = 0; l < sl.dim; ++l)
          dinfo[sl.dofs[l]] = uint32_t(sl.shift) | uint32_t(sl.mask) << 5 | uint32_t(sl.dim) << 8
                              | uint32_t(slot_info[s].w + l) << 16;
      }
      if (ok)
        e->d_dinfo = upload(dinfo);
    }
    e->d_slot_band = upload(slot_band);
  }
  prepare_tabulate(*e);
  return e.release();
}

void require_device(const bx_element& e)
{
  if (!e.on_device)
    fail(BX_ERR_CUDA, "basix_b200: no usable CUDA device (the element holds host tables only; there is no CPU fallback)");
  // the device mirror lives on the GPU that was current at creation: a kernel launched on another device would
  // dereference pointers of the wrong address space (sticky illegal-address fault), so refuse here instead
  int dev = -1;
  BX_CUDA(cudaGetDevice(&dev));
  if (dev != e.device)
    fail(BX_ERR_INVALID_ARG, "basix_b200: the element was created on CUDA device " + std::to_string(e.device)
                                 + " but device " + std::to_string(dev)
                                 + " is current (create one element per GPU, or call bx_set_device first)");
}

void element_net_operator(const bx_element& e, int kind, uint32_t cell_info, int32_t* src, double* A)
{
  if (kind < 0 or kind > 3)
    fail(BX_ERR_INVALID_ARG, "net_operator: kind must be 0..3");
  const int n = e.ndofs;
  if (src)
  {
    if (!e.perms)
      fail(BX_ERR_RUNTIME, "The DOF transformations for this element are not permutations");
    for (int i = 0; i < n; ++i)
      src[i] = i;
  }
  if (A)
  {
    std::fill(A, A + size_t(n) * n, 0.0);
    for (int i = 0; i < n; ++i)
      A[size_t(i) * n + i] = 1.0;
  }
  if (e.identity or e.tdim < 2)
    return;
  for (size_t s = 0; s < e.slots.size(); ++s)
  {
    const auto& sl = e.slots[s];
    const int state = int((cell_info >> sl.shift) & uint32_t(sl.mask));
    if (state == 0)
      continue;
    for (int l = 0; l < sl.dim; ++l)
    {
      const int dof = sl.dofs[l];
      if (e.perms)
      {
        const int sd = e.h_gather[size_t(kind) * e.gather_size + e.h_slot_info[s].w + state * sl.dim + l];
        if (src)
          src[dof] = sd;
        if (A)
        {
          A[size_t(dof) * n + dof] = 0.0;
          A[size_t(dof) * n + sd] = 1.0;
        }
      }
      else if (A)
      {
        const double* row
            = e.h_mats.data() + size_t(kind) * e.mats_size + e.h_slot_mat[s].y + size_t(state) * mat_state_stride(sl.dim)
              + size_t(l) * mat_row_stride(sl.dim);
        for (int i = 0; i < sl.dim; ++i)
          A[size_t(dof) * n + sl.first + i] = row[i];
      }
    }
  }
}
} // namespace bx

bx_element::~bx_element()
{
  // best effort; errors at teardown are ignored
  cudaFree(d_Cp);
  cudaFree(d_Cs);
  cudaFree(d_Cq);
  bx::dop_cache_destroy(dop_cache);
  cudaFree(tp.d_A);
  cudaFree(tp.d_ijk);
  cudaFree(tp.d_scale);
  cudaFree(tp.d_order);
  cudaFree(d_dof_slot);
  cudaFree(d_dof_loc);
  cudaFree(d_slot_info);
  cudaFree(d_gather);
  cudaFree(d_slot_mat);
  cudaFree(d_mats);
  cudaFree(d_band);
  cudaFree(d_dinfo);
  cudaFree(d_slot_band);
  for (auto& cl : closure)
    cudaFree(cl.d_src);
  cudaFree(d_interp[0]);
  cudaFree(d_interp[1]);
  cudaFree(d_interp_blocked);
}

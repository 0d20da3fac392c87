// basix_b200.hpp -- C++ host mirror of the hot-path subset of basix::FiniteElement<F> over the C ABI
// (include/basix_b200.h).  Header-only, C++17; link with basix_b200/libbasix_b200.so.
//
// Same method names, argument meaning and error behaviour as the reference class
// (cpp/basix/finite-element.h): failures surface as std::runtime_error carrying the library's message,
// exactly like the reference's `throw std::runtime_error` (finite-element.cpp:1838-1843,
// finite-element.h:802-806, 653-654).  Shapes are passed explicitly instead of std::mdspan so the header has no
// dependency on the vendored mdspan; every array is row-major and contiguous as in the reference.
//
//   reference                                                     here
//   FiniteElement::tabulate_shape        finite-element.h:364-376      tabulate_shape(nd, npts)
//   FiniteElement::tabulate              finite-element.h:399-483      tabulate(nd, x, npts[, out])
//   FiniteElement::push_forward          finite-element.h:577-579      push_forward(U, J, detJ, K, nc, rows, ...)
//   FiniteElement::pull_back             finite-element.h:592-594      pull_back(...)
//   FiniteElement::T_apply (+7 siblings) finite-element.h:1067-1159    T_apply(u, n, cell_info) ...
//   FiniteElement::permute{,_inv}        finite-element.h:800-843      permute(d, cell_info) ...
//   polyset::tabulate                    polyset.h:181-183             polyset::tabulate(...)
// plus the batched forms the reference cannot express (one cell_info per cell): *_batch.
//
// Memory kinds: every call takes `mem` (BX_MEM_HOST: host pointers, the library stages through the GPU and
// returns when the result is in host memory -- the drop-in behaviour; BX_MEM_DEVICE: device pointers, work is
// enqueued on `stream` and the call returns without synchronising).  There is no CPU fallback.
#pragma once

#include "basix_b200.h"

#include <array>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace basix_b200
{
inline void check(int status)
{
  if (status != BX_OK)
    throw std::runtime_error(bx_last_error());
}

namespace polyset
{
inline int dim(int cell_type, int polyset_type, int degree) { return bx_polyset_dim(cell_type, polyset_type, degree); }
inline int nderivs(int cell_type, int nderiv) { return bx_polyset_nderivs(cell_type, nderiv); }
/// P[nderivs][psize][npts] (polyset.cpp:2914-2992).  x[npts][tdim].
inline std::vector<double> tabulate(int cell_type, int polyset_type, int degree, int nderiv, const double* x,
                                    std::size_t npts)
{
  std::vector<double> P(std::size_t(nderivs(cell_type, nderiv)) * dim(cell_type, polyset_type, degree) * npts);
  check(bx_polyset_tabulate_f64(cell_type, polyset_type, degree, nderiv, x, npts, P.data(), BX_MEM_HOST, nullptr));
  return P;
}
} // namespace polyset

/// Arrays that define an element for the hot path: what basix::FiniteElement exposes through
/// coefficient_matrix(), entity_dofs(), entity_transformations(), dof_ordering(), value_shape(), map_type() ...
struct ElementData
{
  int cell_type = 0, polyset_type = 0, degree = 0, embedded_superdegree = 0, map_type = 0;
  std::vector<std::size_t> value_shape;
  std::vector<double> coefficient_matrix;                  // [ndofs][value_size * psize]
  std::vector<std::vector<std::vector<int>>> entity_dofs;  // [dim][entity][k]
  std::vector<int> dof_ordering;                           // empty or [ndofs]
  // entity_transformations(): [ntrans][n][n] per sub-entity type; n = 0 when the type is present without dofs,
  // n = -1 when the cell has no such sub-entity
  std::vector<double> etrans_interval, etrans_triangle, etrans_quadrilateral;
  int etrans_interval_dim = -1, etrans_triangle_dim = -1, etrans_quadrilateral_dim = -1;
};

class FiniteElement
{
public:
  explicit FiniteElement(const ElementData& d)
  {
    std::size_t vs = 1;
    for (std::size_t s : d.value_shape)
      vs *= s;
    const int psize = polyset::dim(d.cell_type, d.polyset_type, d.embedded_superdegree);
    if (psize <= 0 or d.coefficient_matrix.size() % (vs * std::size_t(psize)) != 0)
      throw std::runtime_error("coefficient_matrix has the wrong number of columns for this polyset");
    std::vector<std::int32_t> offs{0}, flat;
    for (const auto& row : d.entity_dofs)
      for (const auto& e : row)
      {
        flat.insert(flat.end(), e.begin(), e.end());
        offs.push_back(std::int32_t(flat.size()));
      }
    if (flat.empty())
      flat.push_back(0);
    static const double none = 0.0;
    bx_element_desc desc{};
    desc.cell_type = d.cell_type;
    desc.polyset_type = d.polyset_type;
    desc.degree = d.degree;
    desc.embedded_superdegree = d.embedded_superdegree;
    desc.map_type = d.map_type;
    desc.ndofs = int(d.coefficient_matrix.size() / (vs * std::size_t(psize)));
    desc.value_size = int(vs);
    desc.coeffs = d.coefficient_matrix.data();
    desc.dof_ordering = d.dof_ordering.empty() ? nullptr : d.dof_ordering.data();
    desc.n_dof_ordering = int(d.dof_ordering.size());
    desc.entity_dof_offsets = offs.data();
    desc.entity_dofs = flat.data();
    auto set = [](const std::vector<double>& m, int n, const double*& p, int& dim)
    {
      dim = n < 0 ? 0 : n;
      p = n < 0 ? nullptr : (m.empty() ? &none : m.data());
    };
    set(d.etrans_interval, d.etrans_interval_dim, desc.etrans_interval, desc.etrans_interval_dim);
    set(d.etrans_triangle, d.etrans_triangle_dim, desc.etrans_triangle, desc.etrans_triangle_dim);
    set(d.etrans_quadrilateral, d.etrans_quadrilateral_dim, desc.etrans_quadrilateral, desc.etrans_quadrilateral_dim);
    check(bx_element_create_from_arrays(&desc, &_h));
    read_info();
    _value_shape = d.value_shape;
  }
  /// basix::create_element (finite-element.h:1623-1628): built natively for the P family (equispaced /
  /// gll_warped; interval, triangle, tetrahedron, quadrilateral, hexahedron) and for RT / N1E on triangle and
  /// tetrahedron; throws otherwise (see bx_create_element in basix_b200.h).
  static FiniteElement create(int family, int cell_type, int degree, int lagrange_variant = 0, int dpc_variant = 0,
                              bool discontinuous = false, const std::vector<int>& dof_ordering = {})
  {
    bx_element* h = nullptr;
    check(bx_create_element(family, cell_type, degree, lagrange_variant, dpc_variant, discontinuous ? 1 : 0,
                            dof_ordering.empty() ? nullptr : dof_ordering.data(), int(dof_ordering.size()), &h));
    FiniteElement e;
    e._h = h;
    e.read_info();
    e._value_shape = e._vs == 1 ? std::vector<std::size_t>{} : std::vector<std::size_t>{std::size_t(e._vs)};
    return e;
  }
  /// coefficient_matrix() (finite-element.h:1342-1346): [ndofs][value_size * psize]
  std::vector<double> coefficient_matrix() const
  {
    std::int32_t info[10];
    check(bx_element_info(_h, info));
    std::vector<double> c(std::size_t(info[3]) * info[4] * info[9]);
    check(bx_element_coefficient_matrix(_h, c.data()));
    return c;
  }
  ~FiniteElement() { bx_element_destroy(_h); }
  FiniteElement(const FiniteElement&) = delete;
  FiniteElement& operator=(const FiniteElement&) = delete;
  FiniteElement(FiniteElement&& o) noexcept { *this = std::move(o); }
  FiniteElement& operator=(FiniteElement&& o) noexcept
  {
    std::swap(_h, o._h);
    _tdim = o._tdim, _ndofs = o._ndofs, _vs = o._vs, _map = o._map, _perms = o._perms, _identity = o._identity;
    _value_shape = o._value_shape;
    return *this;
  }

  // ---- accessors (finite-element.h:487-561) ---------------------------------------------------------------
  int dim() const { return _ndofs; }
  int value_size() const { return _vs; }
  const std::vector<std::size_t>& value_shape() const { return _value_shape; }
  int map_type() const { return _map; }
  bool dof_transformations_are_permutations() const { return _perms; }
  bool dof_transformations_are_identity() const { return _identity; }
  const bx_element* handle() const { return _h; }

  /// {nderivs, npts, ndofs, value_size}  (finite-element.h:364-376)
  std::array<std::size_t, 4> tabulate_shape(int nd, std::size_t npts) const
  {
    std::array<std::size_t, 4> s;
    check(bx_element_tabulate_shape(_h, nd, npts, s.data()));
    return s;
  }

  /// basis[nderivs][npts][ndofs][value_size]; x[npts][tdim_x]  (finite-element.h:454-455)
  void tabulate(int nd, const double* x, std::size_t npts, int tdim_x, double* basis, int mem = BX_MEM_HOST,
                bx_stream_t stream = nullptr) const
  {
    check(bx_tabulate_f64(_h, nd, x, npts, tdim_x, basis, mem, stream));
  }
  /// Allocating overload (finite-element.h:399-401), host memory.
  std::pair<std::vector<double>, std::array<std::size_t, 4>> tabulate(int nd, const double* x, std::size_t npts,
                                                                      int tdim_x) const
  {
    auto shape = tabulate_shape(nd, npts);
    std::vector<double> out(shape[0] * shape[1] * shape[2] * shape[3]);
    tabulate(nd, x, npts, tdim_x, out.data());
    return {std::move(out), shape};
  }

  /// u[nc][rows][physical value size] = map(U[nc][rows][in_vs]; J[nc][gdim][tdim], detJ[nc], K[nc][tdim][gdim])
  /// (finite-element.h:577-579).  Returns the physical value size.
  int push_forward(const double* U, const double* J, const double* detJ, const double* K, std::size_t nc,
                   std::size_t rows, int in_vs, int gdim, int tdim, double* u, int mem = BX_MEM_HOST,
                   bx_stream_t stream = nullptr) const
  {
    check(bx_push_forward_f64(_h, U, J, detJ, K, nc, rows, in_vs, gdim, tdim, u, mem, stream));
    return bx_map_value_size(_map, 0, in_vs, gdim, tdim);
  }
  /// push_forward of ONE slab U[rows][in_vs] of reference values onto every cell (same result as push_forward on the
  /// array replicated nc times, which is what the reference API requires).
  int push_forward_broadcast(const double* U, const double* J, const double* detJ, const double* K, std::size_t nc,
                             std::size_t rows, int in_vs, int gdim, int tdim, double* u, int mem = BX_MEM_HOST,
                             bx_stream_t stream = nullptr) const
  {
    check(bx_push_forward_broadcast_f64(_h, U, J, detJ, K, nc, rows, in_vs, gdim, tdim, u, mem, stream));
    return bx_map_value_size(_map, 0, in_vs, gdim, tdim);
  }
  /// (finite-element.h:592-594)
  int pull_back(const double* u, const double* J, const double* detJ, const double* K, std::size_t nc,
                std::size_t rows, int in_vs, int gdim, int tdim, double* U, int mem = BX_MEM_HOST,
                bx_stream_t stream = nullptr) const
  {
    check(bx_pull_back_f64(_h, u, J, detJ, K, nc, rows, in_vs, gdim, tdim, U, mem, stream));
    return bx_map_value_size(_map, 1, in_vs, gdim, tdim);
  }

  // ---- DOF transformations: single cell, reference signatures (finite-element.h:1067-1159) -----------------
#define BASIX_B200_T_OP(NAME, OP)                                                                          \
  template <typename T>                                                                                    \
  void NAME(T* u, std::size_t size, int n, std::uint32_t cell_info) const                                  \
  {                                                                                                        \
    transform(OP, u, 1, size, n, &cell_info, BX_MEM_HOST, nullptr);                                        \
  }
  BASIX_B200_T_OP(T_apply, BX_T_APPLY)
  BASIX_B200_T_OP(Tt_apply, BX_TT_APPLY)
  BASIX_B200_T_OP(Tt_inv_apply, BX_TT_INV_APPLY)
  BASIX_B200_T_OP(Tinv_apply, BX_TINV_APPLY)
  BASIX_B200_T_OP(Tt_apply_right, BX_TT_APPLY_RIGHT)
  BASIX_B200_T_OP(Tinv_apply_right, BX_TINV_APPLY_RIGHT)
  BASIX_B200_T_OP(T_apply_right, BX_T_APPLY_RIGHT)
  BASIX_B200_T_OP(Tt_inv_apply_right, BX_TT_INV_APPLY_RIGHT)
#undef BASIX_B200_T_OP

  /// Batched: data[nc][cell_size] in place, cell c with cell_info[c]; op is a bx_top.
  template <typename T>
  void T_apply_batch(int op, T* data, std::size_t nc, std::size_t cell_size, int n, const std::uint32_t* cell_info,
                     int mem = BX_MEM_HOST, bx_stream_t stream = nullptr) const
  {
    transform(op, data, nc, cell_size, n, cell_info, mem, stream);
  }

  /// (finite-element.h:800-812, 831-843)
  void permute(std::int32_t* d, std::uint32_t cell_info) const
  {
    check(bx_permute_i32(_h, 0, d, 1, &cell_info, BX_MEM_HOST, nullptr));
  }
  void permute_inv(std::int32_t* d, std::uint32_t cell_info) const
  {
    check(bx_permute_i32(_h, 1, d, 1, &cell_info, BX_MEM_HOST, nullptr));
  }
  void permute_batch(std::int32_t* d, std::size_t nc, const std::uint32_t* cell_info, bool inverse = false,
                     int mem = BX_MEM_HOST, bx_stream_t stream = nullptr) const
  {
    check(bx_permute_i32(_h, inverse ? 1 : 0, d, nc, cell_info, mem, stream));
  }

  /// (finite-element.h:860-985 / 997-1035): `d` = the dofs on the closure of one sub-entity.
  /// entity_info overloads: the entity's own bits; cell_info overloads: the cell's bits + the entity's index.
  void permute_subentity_closure(std::int32_t* d, std::size_t m, std::uint32_t entity_info, int entity_type) const
  {
    check(bx_permute_subentity_closure_i32(_h, 0, d, 1, m, &entity_info, entity_type, -1, BX_MEM_HOST, nullptr));
  }
  void permute_subentity_closure_inv(std::int32_t* d, std::size_t m, std::uint32_t entity_info, int entity_type) const
  {
    check(bx_permute_subentity_closure_i32(_h, 1, d, 1, m, &entity_info, entity_type, -1, BX_MEM_HOST, nullptr));
  }
  void permute_subentity_closure(std::int32_t* d, std::size_t m, std::uint32_t cell_info, int entity_type,
                                 int entity_index) const
  {
    check(bx_permute_subentity_closure_i32(_h, 0, d, 1, m, &cell_info, entity_type, entity_index, BX_MEM_HOST, nullptr));
  }
  void permute_subentity_closure_inv(std::int32_t* d, std::size_t m, std::uint32_t cell_info, int entity_type,
                                     int entity_index) const
  {
    check(bx_permute_subentity_closure_i32(_h, 1, d, 1, m, &cell_info, entity_type, entity_index, BX_MEM_HOST, nullptr));
  }
  void permute_subentity_closure_batch(std::int32_t* d, std::size_t nc, std::size_t m, const std::uint32_t* info,
                                       int entity_type, int entity_index = -1, bool inverse = false,
                                       int mem = BX_MEM_HOST, bx_stream_t stream = nullptr) const
  {
    check(bx_permute_subentity_closure_i32(_h, inverse ? 1 : 0, d, nc, m, info, entity_type, entity_index, mem, stream));
  }
  int subentity_closure_size(int entity_type) const { return bx_element_subentity_closure_size(_h, entity_type); }

  /// points() / interpolation_matrix() (finite-element.h:1167-1170, 1222-1226) as {data, shape}
  std::pair<std::vector<double>, std::array<std::size_t, 2>> points() const
  {
    std::int32_t s[4];
    check(bx_element_interpolation_shape(_h, s));
    std::vector<double> x(std::size_t(s[0]) * s[1]);
    if (!x.empty())
      check(bx_element_points(_h, x.data()));
    return {std::move(x), {std::size_t(s[0]), std::size_t(s[1])}};
  }
  std::pair<std::vector<double>, std::array<std::size_t, 2>> interpolation_matrix() const
  {
    std::int32_t s[4];
    check(bx_element_interpolation_shape(_h, s));
    std::vector<double> m(std::size_t(s[0] ? s[2] : 0) * s[3]);
    if (!m.empty())
      check(bx_element_interpolation_matrix(_h, m.data()));
    return {std::move(m), {std::size_t(s[0] ? s[2] : 0), std::size_t(s[3])}};
  }
  /// coefficients[c] = interpolation_matrix() * values[c] over nc cells (values[nc][value_size * npts])
  void interpolate_batch(const double* values, std::size_t nc, double* coefficients,
                         int layout = BX_INTERP_COMPONENT_MAJOR, int mem = BX_MEM_HOST,
                         bx_stream_t stream = nullptr) const
  {
    check(bx_interpolate_f64(_h, layout, values, nc, coefficients, mem, stream));
  }

  /// coefficients[c] = interpolation_matrix() * pull_back(u[c], J[c], detJ[c], K[c]) in one pass (u[nc][npts][gdim])
  void interpolate_pulled_batch(const double* u, const double* J, const double* detJ, const double* K, std::size_t nc,
                                int gdim, double* coefficients, int mem = BX_MEM_HOST, bx_stream_t stream = nullptr) const
  {
    check(bx_interpolate_pulled_f64(_h, u, J, detJ, K, nc, gdim, coefficients, mem, stream));
  }
  /// out[c][p][dof][:] = push_forward(T_op(cell_info[c]) tabulate(0, X)[0][p], geometry of cell c): tabulate once, T_apply
  /// and push_forward fused.  Geometry: coords[nc][tdim + 1][gdim] of affine simplex cells (J, detJ, K computed on chip),
  /// or J / detJ / K per cell (coords == nullptr), or -- per_point -- per (cell, point) for non-affine cells.
  void physical_basis(const double* X, std::size_t npts, const double* coords, const double* J, const double* detJ,
                      const double* K, const std::uint32_t* cell_info, std::size_t nc, int gdim, double* out,
                      int op = BX_T_APPLY, bool per_point = false, int mem = BX_MEM_HOST, bx_stream_t stream = nullptr) const
  {
    if (per_point)
      check(bx_physical_basis_pointwise_f64(_h, op, X, npts, _tdim, J, detJ, K, cell_info, nc, gdim, out, mem, stream));
    else
      check(bx_physical_basis_f64(_h, op, X, npts, _tdim, coords, J, detJ, K, cell_info, nc, gdim, out, mem, stream));
  }
  /// J, detJ, K of nc cells at npts points from node coordinates and the coordinate element's derivatives
  /// (dphi == nullptr: affine simplex cells, npts = 1)
  static void geometry(const double* coords, const double* dphi, std::size_t nc, int nv, int gdim, int tdim, int npts,
                       double* J, double* detJ, double* K, int mem = BX_MEM_HOST, bx_stream_t stream = nullptr)
  {
    check(bx_geometry_f64(coords, dphi, nc, nv, gdim, tdim, npts, J, detJ, K, mem, stream));
  }

private:
  FiniteElement() = default;
  void read_info()
  {
    std::int32_t info[10];
    check(bx_element_info(_h, info));
    _tdim = info[2];
    _ndofs = info[3];
    _vs = info[4];
    _map = info[5];
    _perms = info[7] != 0;
    _identity = info[8] != 0;
  }
  void transform(int op, double* u, std::size_t nc, std::size_t cs, int n, const std::uint32_t* ci, int mem,
                 bx_stream_t s) const
  {
    check(bx_T_apply_f64(_h, op, u, nc, cs, n, ci, mem, s));
  }
  void transform(int op, float* u, std::size_t nc, std::size_t cs, int n, const std::uint32_t* ci, int mem,
                 bx_stream_t s) const
  {
    check(bx_T_apply_f32(_h, op, u, nc, cs, n, ci, mem, s));
  }
  void transform(int op, std::int32_t* u, std::size_t nc, std::size_t cs, int n, const std::uint32_t* ci, int mem,
                 bx_stream_t s) const
  {
    check(bx_T_apply_i32(_h, op, u, nc, cs, n, ci, mem, s));
  }

  bx_element* _h = nullptr;
  int _tdim = 0, _ndofs = 0, _vs = 1, _map = 0;
  bool _perms = false, _identity = true;
  std::vector<std::size_t> _value_shape;
};
} // namespace basix_b200

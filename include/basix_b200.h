/* basix_b200.h -- C ABI of the B200-native (sm_100a) basix hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types.
 * Each entry point cites the reference interface (FEniCS/basix, file:line under
 * /root/reference) it replaces.  All functions return BX_OK (0) or an error
 * code; the message for the calling thread is available via bx_last_error().
 * There is NO CPU fallback anywhere below this boundary: if no CUDA device is
 * usable every compute entry point fails with BX_ERR_CUDA.
 *
 * Pointer kinds: every compute call takes `mem`:
 *   BX_MEM_DEVICE  pointers are device pointers valid on the current device;
 *                  work is enqueued on `stream` and the call returns without
 *                  synchronising (stream-ordered, re-entrant).
 *   BX_MEM_HOST    pointers are host pointers (pageable or pinned); the library
 *                  stages host<->device in chunks through pinned buffers on its
 *                  own streams and returns when the result is in host memory.
 *
 * Integer enums keep the reference's values:
 *   cell type     basix::cell::type            cpp/basix/cell.h:20-30
 *   polyset type  basix::polyset::type         cpp/basix/polyset.h:136-140
 *   map type      basix::maps::type            cpp/basix/maps.h:39-47
 *   family        basix::element::family       cpp/basix/element-families.h:44-61
 *   lagrange/dpc variant                       cpp/basix/element-families.h:12-41
 */
#ifndef BASIX_B200_H
#define BASIX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C"
{
#endif

  typedef struct bx_element bx_element;
  typedef void* bx_stream_t; /* a cudaStream_t; NULL = legacy default stream */

  enum bx_status
  {
    BX_OK = 0,
    BX_ERR_INVALID_ARG = 1,
    BX_ERR_CUDA = 2,
    BX_ERR_UNSUPPORTED = 3,
    BX_ERR_RUNTIME = 4
  };

  enum bx_mem
  {
    BX_MEM_HOST = 0,
    BX_MEM_DEVICE = 1
  };

  enum bx_cell
  {
    BX_CELL_POINT = 0,
    BX_CELL_INTERVAL = 1,
    BX_CELL_TRIANGLE = 2,
    BX_CELL_TETRAHEDRON = 3,
    BX_CELL_QUADRILATERAL = 4,
    BX_CELL_HEXAHEDRON = 5,
    BX_CELL_PRISM = 6,
    BX_CELL_PYRAMID = 7
  };

  enum bx_polyset
  {
    BX_POLYSET_STANDARD = 0,
    BX_POLYSET_MACROEDGE = 1
  };

  enum bx_map
  {
    BX_MAP_IDENTITY = 0,
    BX_MAP_L2_PIOLA = 1,
    BX_MAP_COVARIANT_PIOLA = 2,
    BX_MAP_CONTRAVARIANT_PIOLA = 3,
    BX_MAP_DOUBLE_COVARIANT_PIOLA = 4,
    BX_MAP_DOUBLE_CONTRAVARIANT_PIOLA = 5
  };

  /* The eight DOF-transformation operators, finite-element.h:1839-1998. */
  enum bx_top
  {
    BX_T_APPLY = 0,           /* T_apply            finite-element.h:1839-1854 */
    BX_TT_APPLY = 1,          /* Tt_apply           finite-element.h:1856-1870 */
    BX_TT_INV_APPLY = 2,      /* Tt_inv_apply       finite-element.h:1872-1886 */
    BX_TINV_APPLY = 3,        /* Tinv_apply         finite-element.h:1888-1902 */
    BX_TT_APPLY_RIGHT = 4,    /* Tt_apply_right     finite-element.h:1904-1926 */
    BX_TINV_APPLY_RIGHT = 5,  /* Tinv_apply_right   finite-element.h:1928-1950 */
    BX_T_APPLY_RIGHT = 6,     /* T_apply_right      finite-element.h:1952-1974 */
    BX_TT_INV_APPLY_RIGHT = 7 /* Tt_inv_apply_right finite-element.h:1976-1998 */
  };

  /* ---- library / device ---------------------------------------------------------------- */

  /* Message of the last failing call made by the calling thread ("" if none). */
  const char* bx_last_error(void);
  /* Library version string. */
  const char* bx_version(void);
  /* Number of visible CUDA devices (0 if the driver/device is missing). */
  int bx_device_count(void);
  /* cudaSetDevice for the calling thread. */
  int bx_set_device(int device);
  /* Block until `stream` has drained (cudaStreamSynchronize). */
  int bx_stream_synchronize(bx_stream_t stream);
  /* Number of kernels this library has launched since load (all threads);
     bench.py reports it as gpu_launches. */
  uint64_t bx_kernel_launch_count(void);

  /* ---- polynomial sets -------------------------------------------------------------------
   * Replaces polyset::dim (polyset.cpp:3003-3050), polyset::nderivs (polyset.cpp:3052-3075)
   * and polyset::tabulate (dispatch polyset.cpp:2914-2978; python binding
   * tabulate_polynomial_set_float64, basix_wrappers.h:467-475).
   * x[npts][tdim] row-major; P[nderivs][psize][npts] (point index fastest).
   * Standard polysets of every cell type (point, interval, triangle, tetrahedron, quadrilateral, hexahedron, prism,
   * pyramid); the macroedge polysets return BX_ERR_UNSUPPORTED. */
  int bx_polyset_dim(int cell_type, int polyset_type, int degree);
  int bx_polyset_nderivs(int cell_type, int nderiv);
  int bx_polyset_tabulate_f64(int cell_type, int polyset_type, int degree, int nderiv, const double* x,
                              size_t npts, double* P, int mem, bx_stream_t stream);
  int bx_polyset_tabulate_f32(int cell_type, int polyset_type, int degree, int nderiv, const float* x,
                              size_t npts, float* P, int mem, bx_stream_t stream);

  /* ---- element object ---------------------------------------------------------------------
   * Mirrors the immutable data members of basix::FiniteElement<F> that the hot path reads
   * (finite-element.h:1429-1554): _coeffs, _edofs, _entity_transformations (from which
   * _eperm/_etrans* are derived as in finite-element.cpp:1260-1447), _dof_ordering,
   * _value_shape, _map_type, cell/polyset type, _embedded_superdegree. */
  typedef struct bx_element_desc
  {
    int cell_type;            /* bx_cell */
    int polyset_type;         /* bx_polyset */
    int degree;               /* FiniteElement::degree() (tolerance in perm detection) */
    int embedded_superdegree; /* polyset degree, FiniteElement::embedded_superdegree() */
    int map_type;             /* bx_map */
    int ndofs;                /* FiniteElement::dim() */
    int value_size;           /* prod(value_shape) */
    const double* coeffs;     /* coefficient_matrix(): [ndofs][value_size * psize] row-major */
    const int32_t* dof_ordering; /* dof_ordering() or NULL */
    int n_dof_ordering;          /* 0 or ndofs */
    /* entity_dofs(): entities enumerated dimension-major (all vertices, all edges, ...);
       dofs of entity k are entity_dofs[entity_dof_offsets[k] .. entity_dof_offsets[k+1]) */
    const int32_t* entity_dof_offsets; /* length (total number of sub-entities)+1 */
    const int32_t* entity_dofs;
    /* entity_transformations() (finite-element.h:772-776) per sub-entity cell type:
       interval [1][n][n]; triangle / quadrilateral [2][n][n] = {rotation, reflection}.
       NULL / 0 when the cell has no such sub-entity. */
    const double* etrans_interval;
    int etrans_interval_dim;
    const double* etrans_triangle;
    int etrans_triangle_dim;
    const double* etrans_quadrilateral;
    int etrans_quadrilateral_dim;
    /* Optional (NULL / 0 when the element is only tabulated / mapped / transformed): points()
       (finite-element.h:1167-1170), [npts][tdim], and interpolation_matrix() (finite-element.h:1222-1226),
       [ndofs][value_size * npts] with column j * npts + p = value component j at point p. */
    const double* points;
    int npts;
    const double* interpolation_matrix;
  } bx_element_desc;

  /* Build an element (host tables + device mirror on the current device) from arrays. */
  int bx_element_create_from_arrays(const bx_element_desc* desc, bx_element** out);
  void bx_element_destroy(bx_element* e);

  /* Native basix::create_element (finite-element.cpp create_element<T>, python basix.create_element,
     finite_element.py:595-634): same argument order and enum values.  Built natively (host set-up code, as in
     the reference; only the resulting tables go to the device):
       family P   (e-lagrange.cpp:457-666)       variants equispaced / gll_warped (unset allowed for degree < 3) on
                                                  interval, triangle, tetrahedron, quadrilateral, hexahedron;
                                                  continuous or discontinuous; optional dof_ordering
       family RT  (e-raviart-thomas.cpp:20-173)   triangle, tetrahedron
       family N1E (e-nedelec.cpp:24-380)          triangle, tetrahedron
     RT / N1E take the variant of their moment spaces: legendre, equispaced or gll_warped (unset resolves per
     moment space as the reference does).  Anything else returns BX_ERR_UNSUPPORTED: describe such elements
     with bx_element_create_from_arrays. */
  int bx_create_element(int family, int cell_type, int degree, int lagrange_variant, int dpc_variant,
                        int discontinuous, const int32_t* dof_ordering, int n_dof_ordering, bx_element** out);

  /* Defining arrays of an element (reference accessors finite-element.h:665-668, 772-776, 1342-1346, 1397):
     coefficient_matrix() [ndofs][value_size * psize]; entity_dofs()[dim][index] (returns the count, fills
     out when non-NULL); entity_transformations() of one sub-entity cell type (shape {ntrans, n, n}; returns
     1 when the cell has that sub-entity type, 0 otherwise; fills out when non-NULL); dof_ordering()
     (returns the length, 0 when none). */
  int bx_element_coefficient_matrix(const bx_element* e, double* out);
  int bx_element_entity_dofs(const bx_element* e, int dim, int index, int32_t* out);
  int bx_element_entity_transformations(const bx_element* e, int entity_cell_type, int32_t shape[3], double* out);
  int bx_element_dof_ordering(const bx_element* e, int32_t* out);

  /* Accessors (reference: finite-element.h:487-561). info[0..9] = cell_type, polyset_type,
     tdim, ndofs, value_size, map_type, embedded_superdegree, dof_transformations_are_permutations,
     dof_transformations_are_identity, psize. */
  int bx_element_info(const bx_element* e, int32_t info[10]);
  /* FiniteElement::tabulate_shape, finite-element.h:364-376: {nderivs, npts, ndofs, value_size}. */
  int bx_element_tabulate_shape(const bx_element* e, int nd, size_t npts, size_t shape[4]);
  /* Host-side inspection of the composed DOF-transformation tables (no device needed): the net
     operator that op (a bx_top; the *_right variants share the operator of their left twin) applies
     for one cell_info.  src[ndofs] (permutation elements only): out[i] = in[src[i]];
     A[ndofs][ndofs]: out = A in.  Either may be NULL.  Generalises
     FiniteElement::base_transformations (finite-element.cpp:1902-1969) to arbitrary cell_info. */
  int bx_element_net_operator(const bx_element* e, int op, uint32_t cell_info, int32_t* src, double* A);
  /* Build and upload, now, every operand tabulate() needs for derivative orders 0..nd (the derivative-operator
     kernel composes its operators per order on first use, with a synchronous allocation and copy).  Elements are
     prepared for nd <= 2 when they are created; call this for higher orders before capturing bx_tabulate_* into a
     CUDA graph or timing a first call.  After it, BX_MEM_DEVICE calls only launch kernels and stream-ordered
     allocations (capturable). */
  int bx_element_prepare(const bx_element* e, int nd);
  /* Which device path tabulate() will take for derivative order nd:
     0 generic (polyset kernel + DMMA contraction), 1 fused simplex kernel,
     2 tensor-product kernel, 3 register path (low-degree simplex elements),
     4 derivative-operator kernel (simplex elements: all slabs from the values of the orthonormal set). */
  int bx_element_tabulate_path(const bx_element* e, int nd);
  /* Testing / profiling hook: bit i of `mask` allows path i above (the generic path 0 is always available);
     bit 5 allows the staged variant of the tensor-product kernel (tabulate_tp_staged.cu, same path number 2).
     Returns the previous mask; the default allows everything. */
  unsigned bx_tabulate_path_mask(unsigned mask);

  /* ---- FiniteElement::tabulate ------------------------------------------------------------
   * Replaces FiniteElement<F>::tabulate (finite-element.cpp:1834-1888; python
   * FiniteElement.tabulate, finite_element.py:63-95, binding basix_wrappers.h:117-124).
   * x[npts][tdim]; basis[nderivs][npts][ndofs][value_size].  `tdim_x` is x.shape[1] and is
   * validated against the element ("Point dim (..) does not match element dim (..)."). */
  int bx_tabulate_f64(const bx_element* e, int nd, const double* x, size_t npts, int tdim_x, double* basis,
                      int mem, bx_stream_t stream);
  /* float32 points and table (the reference's FiniteElement<float>, finite-element.cpp:2051).  The element data
     stays float64 and the arithmetic is float64; loads widen and stores narrow inside the kernels. */
  int bx_tabulate_f32(const bx_element* e, int nd, const float* x, size_t npts, int tdim_x, float* basis,
                      int mem, bx_stream_t stream);

  /* ---- push_forward / pull_back -----------------------------------------------------------
   * Replaces FiniteElement<F>::push_forward (finite-element.cpp:1971-2005), pull_back
   * (finite-element.cpp:2007-2042) and the kernels of maps.h:56-192.
   * U[nc][rows][in_vs], J[nc][gdim][tdim], detJ[nc], K[nc][tdim][gdim] -> out[nc][rows][out_vs].
   * bx_map_value_size gives out_vs (compute_value_size, finite-element.cpp:42-59, for push;
   * the reference value size for pull). */
  int bx_map_value_size(int map_type, int pull, int in_vs, int gdim, int tdim);
  int bx_map_f64(int map_type, int pull, const double* U, const double* J, const double* detJ,
                 const double* K, size_t nc, size_t rows, int in_vs, int gdim, int tdim, double* out,
                 int mem, bx_stream_t stream);
  int bx_push_forward_f64(const bx_element* e, const double* U, const double* J, const double* detJ,
                          const double* K, size_t nc, size_t rows, int in_vs, int gdim, int tdim,
                          double* out, int mem, bx_stream_t stream);
  /* push_forward of ONE slab of reference values U[rows][in_vs] onto every cell (the tabulated reference basis is
     the same on all cells: the reference API makes the caller replicate it nc times, finite-element.cpp:1971-2005).
     Same result as bx_push_forward_* on the replicated array; only J / detJ / K and the result cross HBM. */
  int bx_push_forward_broadcast_f64(const bx_element* e, const double* U, const double* J, const double* detJ,
                                    const double* K, size_t nc, size_t rows, int in_vs, int gdim, int tdim, double* out,
                                    int mem, bx_stream_t stream);
  int bx_push_forward_broadcast_f32(const bx_element* e, const float* U, const float* J, const float* detJ, const float* K,
                                    size_t nc, size_t rows, int in_vs, int gdim, int tdim, float* out, int mem,
                                    bx_stream_t stream);
  int bx_pull_back_f64(const bx_element* e, const double* u, const double* J, const double* detJ,
                       const double* K, size_t nc, size_t rows, int in_vs, int gdim, int tdim, double* out,
                       int mem, bx_stream_t stream);
  /* float32 instantiations (float arithmetic in the reference's operation order, as FiniteElement<float>). */
  int bx_map_f32(int map_type, int pull, const float* U, const float* J, const float* detJ, const float* K,
                 size_t nc, size_t rows, int in_vs, int gdim, int tdim, float* out, int mem, bx_stream_t stream);
  int bx_push_forward_f32(const bx_element* e, const float* U, const float* J, const float* detJ, const float* K,
                          size_t nc, size_t rows, int in_vs, int gdim, int tdim, float* out, int mem,
                          bx_stream_t stream);
  int bx_pull_back_f32(const bx_element* e, const float* u, const float* J, const float* detJ, const float* K,
                       size_t nc, size_t rows, int in_vs, int gdim, int tdim, float* out, int mem,
                       bx_stream_t stream);

  /* ---- cell geometry and the fused physical basis -----------------------------------------
   * bx_geometry: J, detJ and K of nc cells at npts points each, from the cells' node coordinates and the derivatives of
   * the coordinate element -- what every caller computes before push_forward / pull_back (the reference's own analogue
   * is the affine case, cell::entity_jacobians, cell.cpp:662-689):
   *     J[c][p][g][t] = sum_v coords[c][v][g] * dphi[t][p][v]
   *     detJ = det J (gdim == tdim) or sqrt(det(J^T J));  K = J^-1 (adjugate / det) or (J^T J)^-1 J^T.
   * coords[nc][nv][gdim]; dphi[tdim][npts][nv] = slabs 1..tdim of tabulate(1, X) of the scalar coordinate element, or
   * NULL for affine simplex cells (nv = tdim + 1, npts = 1: J(:, t) = x[1 + t] - x[0], bit-identical to
   * cell::entity_jacobians).  Outputs J[nc][npts][gdim][tdim], detJ[nc][npts], K[nc][npts][tdim][gdim]; any may be NULL. */
  int bx_geometry_f64(const double* coords, const double* dphi, size_t nc, int nv, int gdim, int tdim, int npts, double* J,
                      double* detJ, double* K, int mem, bx_stream_t stream);
  int bx_geometry_f32(const float* coords, const float* dphi, size_t nc, int nv, int gdim, int tdim, int npts, float* J,
                      float* detJ, float* K, int mem, bx_stream_t stream);
  /* Fused tabulate-once -> T_apply -> push_forward: the three reference calls a caller makes to get the basis functions
   * of every physical cell (FiniteElement::tabulate finite-element.cpp:1834-1888 once; per cell T_apply
   * finite-element.h:1839-1854 on the reference values with block size = value size, then push_forward
   * finite-element.cpp:1971-2005), in one pass that writes only the result:
   *     out[c][p][dof][:] = push_forward( T_op(cell_info[c]) U[p] , J_c, detJ_c, K_c )
   * U[npts][ndofs][value_size] = tabulate(0, X)[0] (bx_push_forward_fused) or the points X[npts][tdim] themselves
   * (bx_physical_basis: tabulated on the device first).  Geometry: coords[nc][tdim + 1][gdim] (affine triangle /
   * tetrahedron cells; J, detJ, K are computed on chip as bx_geometry computes them and never touch memory) or, with
   * coords == NULL, the per-cell arrays J[nc][gdim][tdim], detJ[nc], K[nc][tdim][gdim] (only the ones the map reads
   * are dereferenced).  cell_info[nc] may be NULL (no transformation).  op: BX_T_APPLY, BX_TT_APPLY, BX_TT_INV_APPLY or
   * BX_TINV_APPLY.  Identity, covariant and contravariant Piola maps; elements of up to 256 dofs.
   * out[nc][npts][ndofs][physical value size].  Equal, bit for bit, to the library's own three-call sequence. */
  int bx_push_forward_fused_f64(const bx_element* e, int op, const double* U, size_t npts, const double* coords,
                                const double* J, const double* detJ, const double* K, const uint32_t* cell_info, size_t nc,
                                int gdim, double* out, int mem, bx_stream_t stream);
  int bx_push_forward_fused_f32(const bx_element* e, int op, const float* U, size_t npts, const float* coords,
                                const float* J, const float* detJ, const float* K, const uint32_t* cell_info, size_t nc,
                                int gdim, float* out, int mem, bx_stream_t stream);
  int bx_physical_basis_f64(const bx_element* e, int op, const double* X, size_t npts, int tdim_x, const double* coords,
                            const double* J, const double* detJ, const double* K, const uint32_t* cell_info, size_t nc,
                            int gdim, double* out, int mem, bx_stream_t stream);

  /* ---- DOF transformations ----------------------------------------------------------------
   * Replaces T_apply and its 7 siblings (finite-element.h:1839-1998), i.e. permute_data
   * (finite-element.h:1703-1760) / transform_data (finite-element.h:1762-1837) with
   * precompute::apply_* (precompute.h:136-374), batched over cells: data[nc][cell_size] is
   * transformed in place, cell c using cell_info[c]; cell_size = ndofs * n.  The reference
   * API is the nc == 1 case.  op is a bx_top. */
  int bx_T_apply_f64(const bx_element* e, int op, double* data, size_t nc, size_t cell_size, int n,
                     const uint32_t* cell_info, int mem, bx_stream_t stream);
  int bx_T_apply_f32(const bx_element* e, int op, float* data, size_t nc, size_t cell_size, int n,
                     const uint32_t* cell_info, int mem, bx_stream_t stream);
  int bx_T_apply_i32(const bx_element* e, int op, int32_t* data, size_t nc, size_t cell_size, int n,
                     const uint32_t* cell_info, int mem, bx_stream_t stream);
  /* Replaces FiniteElement::permute / permute_inv (finite-element.h:800-843): d[nc][ndofs]. */
  int bx_permute_i32(const bx_element* e, int inverse, int32_t* d, size_t nc, const uint32_t* cell_info,
                     int mem, bx_stream_t stream);

  /* ---- permute_subentity_closure ------------------------------------------------------------
   * Replaces FiniteElement::permute_subentity_closure / _inv (finite-element.h:860-1035; tables built at
   * finite-element.cpp:1449-1705; python/basix/finite_element.py:322-374), batched over cells:
   * d[nc][m] holds, per cell, the dofs on the closure of ONE sub-entity of type `entity_type` (a bx_cell:
   * interval, triangle or quadrilateral; m = bx_element_subentity_closure_size); row c is permuted in place.
   * entity_index < 0: info[c] is the entity's own bits (bit 0 reflected, bits 1-2 rotations) -- the reference's
   * (d, entity_info, entity_type) overload; entity_index >= 0: info[c] is the cell's cell_info and the bits of
   * entity `entity_index` are extracted as the reference does -- the (d, cell_info, entity_type, entity_index)
   * overload.  Permutation elements only (same error text as the reference otherwise).  The reference API is
   * the nc == 1 case. */
  int bx_permute_subentity_closure_i32(const bx_element* e, int inverse, int32_t* d, size_t nc, size_t m,
                                       const uint32_t* info, int entity_type, int entity_index, int mem,
                                       bx_stream_t stream);
  /* Number of dofs on the closure of a sub-entity of type `entity_type` (0: none on this cell; -1: bad argument). */
  int bx_element_subentity_closure_size(const bx_element* e, int entity_type);
  /* Host table behind the call above: src[m] with out[k] = in[src[k]] for one entity_info (testing / inspection). */
  int bx_element_subentity_closure_map(const bx_element* e, int entity_type, int inverse, uint32_t entity_info,
                                       int32_t* src);


  /* ---- interpolation ------------------------------------------------------------------------
   * points() / interpolation_matrix() of the reference (finite-element.h:1167-1170, 1222-1226) and the batched
   * application of the matrix -- the step that follows pull_back when a function is interpolated into the
   * element: the reference leaves `coefficients = interpolation_matrix * values` to the caller, one cell at a
   * time (finite-element.h:1183-1216).  values[nc][value_size * npts] -> coefficients[nc][ndofs].
   * layout BX_INTERP_COMPONENT_MAJOR: values[c][j * npts + p], the reference's column order;
   * layout BX_INTERP_POINT_MAJOR:     values[c][p * value_size + j], the order pull_back writes
   *                                   (rows = points), so its output feeds this call without a transpose. */
  enum
  {
    BX_INTERP_COMPONENT_MAJOR = 0,
    BX_INTERP_POINT_MAJOR = 1
  };
  /* shape[0] = npts (0 when the element was built without interpolation data), shape[1] = tdim,
     shape[2] = ndofs, shape[3] = value_size * npts */
  int bx_element_interpolation_shape(const bx_element* e, int32_t* shape);
  int bx_element_points(const bx_element* e, double* out);               /* [npts][tdim] */
  int bx_element_interpolation_matrix(const bx_element* e, double* out); /* [ndofs][value_size * npts] */
  int bx_interpolate_f64(const bx_element* e, int layout, const double* values, size_t nc, double* coefficients,
                         int mem, bx_stream_t stream);
  /* pull_back fused into the interpolation: coefficients[c] = interpolation_matrix * pull_back(u[c], J_c, detJ_c, K_c) --
     FiniteElement::pull_back (finite-element.cpp:2007-2042) followed by the product with interpolation_matrix()
     (finite-element.h:1183-1226) -- in one pass that reads the physical values once and never writes the pulled-back
     ones.  u[nc][npts][gdim]: values of the function at the mapped interpolation points, the layout pull_back takes.
     Covariant (needs J) and contravariant (needs K and detJ) Piola elements with gdim == tdim == value size; other
     elements: BX_ERR_UNSUPPORTED (call bx_pull_back and bx_interpolate).  Equal to those two calls to rounding (fused
     multiply-adds in the map, another summation order in the contraction: <= 1e-13 relative). */
  int bx_interpolate_pulled_f64(const bx_element* e, const double* u, const double* J, const double* detJ, const double* K,
                                size_t nc, int gdim, double* coefficients, int mem, bx_stream_t stream);
  /* float32 arrays (loads widen, stores narrow; matrix and arithmetic float64) */
  int bx_interpolate_pulled_f32(const bx_element* e, const float* u, const float* J, const float* detJ, const float* K, size_t nc,
                                int gdim, float* coefficients, int mem, bx_stream_t stream);
  /* float32 values / coefficients (the reference's FiniteElement<float>); the matrix and the arithmetic are float64 */
  int bx_interpolate_f32(const bx_element* e, int layout, const float* values, size_t nc, float* coefficients,
                         int mem, bx_stream_t stream);

  /* The same fused pass for NON-AFFINE cells (hexahedra, curved simplices): the Jacobian varies from point to point, so the
     geometry is J[nc][npts][gdim][tdim], detJ[nc][npts], K[nc][npts][tdim][gdim] -- exactly what bx_geometry writes when it
     is given the derivatives of the coordinate element --
         out[c][p][dof][:] = push_forward( T_op(cell_info[c]) U[p] , J[c][p], detJ[c][p], K[c][p] ).
     Covariant and contravariant Piola maps (the identity map has no geometry: use bx_push_forward_fused); the state tables
     of all points must fit in shared memory (npts * ndofs below about 900 rows for 3-component float64 values).  Equal,
     bit for bit, to T_apply on the replicated reference values followed by push_forward with nc * npts Jacobians. */
  int bx_physical_basis_pointwise_f64(const bx_element* e, int op, const double* X, size_t npts, int tdim_x, const double* J,
                                      const double* detJ, const double* K, const uint32_t* cell_info, size_t nc, int gdim,
                                      double* out, int mem, bx_stream_t stream);
  int bx_push_forward_fused_pointwise_f64(const bx_element* e, int op, const double* U, size_t npts, const double* J,
                                          const double* detJ, const double* K, const uint32_t* cell_info, size_t nc, int gdim,
                                          double* out, int mem, bx_stream_t stream);
  int bx_push_forward_fused_pointwise_f32(const bx_element* e, int op, const float* U, size_t npts, const float* J,
                                          const float* detJ, const float* K, const uint32_t* cell_info, size_t nc, int gdim,
                                          float* out, int mem, bx_stream_t stream);
  /* Replaces basix::compute_interpolation_operator (interpolation.cpp:16-116; python
     basix.compute_interpolation_operator): the matrix taking the coefficients of a function in `from` to those of its
     interpolant in `to` -- `from` tabulated at to's points(), to's interpolation_matrix() applied -- composed on the
     device from bx_tabulate and bx_interpolate.  shape: {dim_to, dim_from} (equal value sizes), {dim_to * vs_from,
     dim_from} (scalar target) or {dim_to, dim_from * vs_to} (scalar source); same errors as the reference otherwise.
     out[shape[0]][shape[1]] row-major; out == NULL returns the shape only. */
  int bx_compute_interpolation_operator_f64(const bx_element* from, const bx_element* to, int32_t shape[2], double* out,
                                            int mem, bx_stream_t stream);

  /* Replaces lattice::create (lattice.cpp:636-700; python basix.create_lattice) for the lattices the native factory
     builds its elements from: lattice_type 0 equispaced, 1 gll (simplices: warp method, simplex_method = 1).  Host
     set-up code.  out[npts][tdim] (NULL: only *npts is returned). */
  int bx_create_lattice(int cell_type, int n, int lattice_type, int exterior, int simplex_method, double* out, size_t* npts);

  /* quadrature::make_quadrature with the Gauss-Jacobi scheme (quadrature.cpp:215-372; python basix.make_quadrature(...,
     rule=QuadratureType.gauss_jacobi)): a rule exact for polynomials of degree `degree` on any cell.  Host set-up code
     (the native element factory integrates its moments with it).  points[npts][tdim], weights[npts]; NULL outputs
     return *npts only. */
  int bx_make_quadrature(int cell_type, int degree, double* points, double* weights, size_t* npts);
  /* quadrature::gauss_jacobi_rule (quadrature.cpp:4955-4965; python basix.quadrature.gauss_jacobi_rule): the npoints-point
     Gauss-Jacobi rule on [0, 1] for the weight (1 - x)^alpha, alpha > -1 real.  points[npoints], weights[npoints].  Host
     set-up code. */
  int bx_gauss_jacobi_rule(double alpha, int npoints, double* points, double* weights);

#ifdef __cplusplus
}
#endif
#endif /* BASIX_B200_H */

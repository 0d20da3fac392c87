// Host-side element object + its device mirror.
//
// Mirrors the immutable members of the reference's basix::FiniteElement<F> that the hot path
// reads (cpp/basix/finite-element.h:1429-1554).  The DOF-transformation data is NOT kept in the
// reference's "prepared swap sequence / in-place LU" form (precompute.h); instead, at creation
// time, the host composes -- for each transformable sub-entity and each value of that entity's
// cell_info bit-field -- the net operator the reference's sequence of swaps / matrix sweeps
// produces, so a device thread can produce one output entry with one table lookup (gather for
// permutation elements, one short dot product for matrix elements) instead of replaying a
// data-dependent sequence of in-place updates.
#pragma once

#include "common.cuh"

#include <array>
#include <vector>

namespace bx
{
// Per-element cache of the derivative-operator tabulate operands, one per derivative order (tabulate_dop.cu).
struct bx_element_dop;
bx_element_dop* dop_cache_create();
void dop_cache_destroy(bx_element_dop* c);
} // namespace bx

struct bx_element
{
  // ---- scalar facts -------------------------------------------------------------------------
  int cell_type = 0, polyset_type = 0, degree = 0, superdegree = 0, map_type = 0;
  int ndofs = 0, vs = 1, tdim = 0, psize = 0;
  bool perms = false, identity = true; // dof_transformations_are_{permutations,identity}
  int device = 0;

  // ---- host copies ----------------------------------------------------------------------------
  std::vector<double> coeffs;                           // [ndofs][vs*psize]
  std::vector<int> dof_ordering;                        // empty or [ndofs]
  std::vector<std::vector<std::vector<int>>> edofs;     // [dim][entity][k]
  std::array<std::vector<double>, 8> etrans;            // per sub-entity cell type: [ntrans][n][n]
  std::array<int, 8> etrans_dim{};                      // n per sub-entity cell type

  // ---- tabulate: contraction operand ----------------------------------------------------------
  // Cp[Npad][Kpad]: row (dof_ordering[dof]*vs + j) = coeffs[dof][j*psize .. (j+1)*psize), zero padded
  int N = 0, K = 0, Npad = 0, Kpad = 0;
  std::vector<double> Cp;
  std::vector<double> Cs; // fused-kernel operand (host copy), empty when not applicable
  double* d_Cp = nullptr;
  double* d_Cs = nullptr; // fused-kernel operand (norm-scaled, production order, padded) or null
  std::vector<double> Cq; // register-path operand [N][K] (norm-scaled, production order), low-degree simplices
  double* d_Cq = nullptr;
  bx::bx_element_dop* dop_cache = nullptr; // derivative-operator path operands, built lazily per nd

  // ---- tabulate: tensor-product factorisation (quad / hex), see tabulate_tp.cu -------------------------
  struct TensorProduct
  {
    bool valid = false;
    int nfun[3] = {0, 0, 0};     // distinct 1-D functions per axis
    int fstride = 0;             // max(nfun)
    std::vector<double> A;       // [tdim][fstride][n1] 1-D coefficients (w.r.t. un-normalised Legendre)
    std::vector<int32_t> ijk;    // [N] packed i | j << 8 | k << 16
    std::vector<double> scale;   // [N]
    double* d_A = nullptr;
    int32_t* d_ijk = nullptr;
    double* d_scale = nullptr;
    // staged kernel (tabulate_tp_staged.cu): when every axis has nf functions and the columns are exactly the
    // nf^tdim index triples, entry (i * nf + j) * nf + k = {scale, column} of that triple; else empty
    int nf = 0;
    std::vector<double2> order; // .x = scale, .y = raw int64 bits of the column
    double2* d_order = nullptr;
  } tp;

  // ---- DOF transformations ----------------------------------------------------------------------
  // "slots" = transformable sub-entities in the reference's processing order (edges, then faces)
  struct Slot
  {
    int shift;      // first cell_info bit of this entity
    int mask;       // 1 (edge: reversed?) or 7 (face: reflect | rotations<<1)
    int dim;        // number of dofs on the entity
    int cls;        // sub-entity cell type (interval / triangle / quadrilateral)
    int first;      // matrix path: first dof of the contiguous block (running dofstart)
    std::vector<int> dofs; // permutation path: _edofs[dim][entity]
  };
  std::vector<Slot> slots;
  std::vector<int16_t> dof_slot, dof_loc; // per dof: owning slot (-1: never moves) and local index
  int lo_dof = 0, hi_dof = 0;             // [lo, hi): smallest dof range containing all movable dofs

  // Host tables (all four operator kinds packed), uploaded verbatim to the device.
  std::vector<int4> h_slot_info;  // [nslots] {shift, mask, dim, gather base}
  std::vector<int2> h_slot_mat;   // [nslots] {first dof, matrix base of the slot's class}
  std::vector<int16_t> h_gather;  // [4][gather_size]
  std::vector<double> h_mats;     // [4][mats_size]
  std::vector<int2> h_band;       // per (class, row): [first, last) column of the non-zero band, union over all
                                  // operator kinds and states (entries below 1e-14 * max|M| count as zero)
  std::vector<int> h_slot_band;   // [nslots] index of the slot's class in h_band
  bool on_device = false;         // false when the element was built without a usable CUDA device

  // Device tables (all ops packed).  See dof_transform.cu for the layout.
  int16_t* d_dof_slot = nullptr;  // [ndofs]
  int16_t* d_dof_loc = nullptr;   // [ndofs]
  int4* d_slot_info = nullptr;    // [nslots] {shift, mask, dim, table base}
  int16_t* d_gather = nullptr;    // [4 ops][gather_size]: source dof for (slot, state, loc)
  int gather_size = 0;
  int2* d_slot_mat = nullptr;     // [nslots] {first dof, matrix base of its class}
  double* d_mats = nullptr;       // [4 ops][mats_size]: dense [state][dim][dim] per class
  int mats_size = 0;
  int2* d_band = nullptr;         // [band_size]
  int* d_slot_band = nullptr;     // [nslots]
  int band_size = 0;
  uint32_t* d_dinfo = nullptr;    // [ndofs] shift | mask << 5 | dim << 8 | (gather base + loc) << 16, or null

  // ---- sub-entity closure permutations (permutation elements; finite-element.cpp:1449-1705) -----------
  struct Closure
  {
    int size = 0;             // dofs on the closure of a sub-entity of this type; 0: the cell has none
    int states = 0;           // 2 (interval: reversed?) or 8 (faces: reflect | rotations << 1)
    std::vector<int16_t> src; // [2 directions][states][size] net gather: out[k] = in[src[k]]
    int16_t* d_src = nullptr;
  };
  std::array<Closure, 8> closure; // by sub-entity cell type (interval, triangle, quadrilateral)

  // ---- interpolation (optional): points(), interpolation_matrix() and the device operands of interpolate.cu ----
  int interp_npts = 0;
  std::vector<double> points;   // [npts][tdim]
  std::vector<double> interp_M; // [ndofs][vs * npts], reference column order (component major)
  int interp_kstride = 0, interp_npad = 0; // operand rows: ndofs padded to 8; row stride == 4 (mod 8) doubles
  double* d_interp[2] = {nullptr, nullptr}; // [layout][npad][kstride]: column order of each value layout
  // fused pull_back + interpolation (Piola elements): columns k' = i * interp_pb + p, interp_pb = npts padded to 8
  int interp_pb = 0, interp_kstride_blocked = 0;
  double* d_interp_blocked = nullptr; // [npad][interp_kstride_blocked]

  ~bx_element();
  bx_element() = default;
  bx_element(const bx_element&) = delete;
  bx_element& operator=(const bx_element&) = delete;
};

namespace bx
{
// Layout of the dense net operators of one entity class: [state][row][mat_row_stride(dim)], states
// mat_state_stride(dim) doubles apart.  Rows are padded to an even length (zero fill) so a thread can fetch
// two entries with one 16-byte shared-memory load; the state stride is rounded up to == 2 (mod 16) doubles,
// so lanes whose cells are in different states hit different banks when they read the same entry, while
// lanes in the same state broadcast (dof_transform.cu).
__host__ __device__ constexpr int mat_row_stride(int dim) { return dim + (dim & 1); }
__host__ __device__ constexpr int mat_state_stride(int dim)
{
  const int n = dim * mat_row_stride(dim);
  return n + ((2 - n % 16) + 16) % 16;
}
// Build host tables and upload the device mirror (current device).  Without a usable device the
// host tables are still built (so they can be inspected) and every compute call fails.
bx_element* element_from_desc(const bx_element_desc& d);
// Throws BX_ERR_CUDA unless the device mirror exists.
void require_device(const bx_element& e);
// Net operator of T-kind `kind` (0 T, 1 Tt, 2 Tt_inv, 3 Tinv) for one cell_info, from the host tables:
// permutation elements fill src[ndofs] (out[i] = in[src[i]]); all elements can fill the dense
// A[ndofs][ndofs] (out = A in).  Either pointer may be null.
void element_net_operator(const bx_element& e, int kind, uint32_t cell_info, int32_t* src, double* A);
} // namespace bx

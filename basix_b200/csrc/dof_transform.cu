// Per-cell DOF transformations driven by cell_info bits, batched over cells.
//
// Replaces T_apply and its seven siblings (reference cpp/basix/finite-element.h:1839-1998), i.e.
// permute_data (finite-element.h:1703-1760) / transform_data (finite-element.h:1762-1837) and the
// precompute::apply_* kernels (precompute.h:136-374), plus permute / permute_inv
// (finite-element.h:800-843).
//
// B200 design (HBM-bound, integer-indexed): a CTA stages a tile of whole cells into shared memory
// with fully coalesced loads, then every thread produces ONE output entry from the staged tile:
//   * permutation elements: out[dof] = tile[src] with src from a gather table indexed by the
//     owning sub-entity's cell_info bit-field (bit-exact for any data type);
//   * matrix elements: out[dof] = sum_i A[state][loc][i] * tile[first + i] with the dense net
//     operator of that bit-field composed on the host.
// Each global byte is read once and written at most once; entries that cannot change are not
// written back.
#include "element.cuh"

#include <algorithm>
#include <type_traits>

namespace bx
{
namespace
{
constexpr int kThreads = 256;

struct FastDiv
{
  unsigned long long m;
  unsigned d;
  __host__ explicit FastDiv(unsigned d_ = 1) : m((1ull << 32) / d_ + 1), d(d_) {}
  // exact for q < 2^16 and d < 2^16
  __device__ __forceinline__ unsigned div(unsigned q) const { return unsigned((q * m) >> 32); }
};

struct TransformParams
{
  size_t nc;
  int cell_size, n, ndofs, cells_per_tile;
  FastDiv by_cell, by_n, by_ndofs;
  const uint32_t* cell_info;
  const int16_t* dof_slot;
  const int16_t* dof_loc;
  const int4* slot_info;
  const int16_t* gather; // op offset applied
  const int2* slot_mat;
  const double* mats; // op offset applied
};

template <typename T, bool RIGHT, bool MATRIX>
__global__ void __launch_bounds__(kThreads) dof_transform_kernel(T* __restrict__ data, TransformParams p)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* tile = reinterpret_cast<T*>(smem_raw);
  const int cs = p.cell_size;
  const size_t ntiles = ceil_div(p.nc, size_t(p.cells_per_tile));
  for (size_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x)
  {
    const size_t c0 = tl * p.cells_per_tile;
    const unsigned ncell = unsigned(min(size_t(p.cells_per_tile), p.nc - c0));
    const unsigned total = ncell * cs;
    T* g = data + c0 * cs;
    // ---- stage the tile (coalesced; 16-byte vectors when the tile base is aligned) ----
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0)
    {
      constexpr unsigned V = 16 / sizeof(T);
      const unsigned nv = total / V;
      const uint4* g4 = reinterpret_cast<const uint4*>(g);
      uint4* t4 = reinterpret_cast<uint4*>(tile);
      for (unsigned q = threadIdx.x; q < nv; q += kThreads)
        t4[q] = g4[q];
      for (unsigned q = nv * V + threadIdx.x; q < total; q += kThreads)
        tile[q] = g[q];
    }
    else
    {
      for (unsigned q = threadIdx.x; q < total; q += kThreads)
        tile[q] = g[q];
    }
    __syncthreads();
    // ---- one output entry per thread-iteration ----
    for (unsigned q = threadIdx.x; q < total; q += kThreads)
    {
      const unsigned c = p.by_cell.div(q);
      const unsigned r = q - c * cs;
      unsigned dof, b;
      if constexpr (RIGHT)
      {
        b = p.by_ndofs.div(r);
        dof = r - b * p.ndofs;
      }
      else
      {
        dof = p.by_n.div(r);
        b = r - dof * p.n;
      }
      const int slot = __ldg(p.dof_slot + dof);
      if (slot < 0)
        continue;
      const int4 si = __ldg(p.slot_info + slot);
      const unsigned state = (__ldg(p.cell_info + c0 + c) >> si.x) & unsigned(si.y);
      if (state == 0)
        continue;
      const int loc = __ldg(p.dof_loc + dof);
      const T* cell = tile + c * cs;
      if constexpr (!MATRIX)
      {
        const unsigned sd = unsigned(__ldg(p.gather + si.w + state * si.z + loc));
        const unsigned src = RIGHT ? b * p.ndofs + sd : sd * p.n + b;
        g[q] = cell[src];
      }
      else
      {
        const int2 sm = __ldg(p.slot_mat + slot);
        const int dim = si.z;
        const double* row = p.mats + sm.y + (size_t(state) * dim + loc) * dim;
        double acc = 0.0;
        if constexpr (RIGHT)
        {
          const T* v = cell + b * p.ndofs + sm.x;
          for (int i = 0; i < dim; ++i)
            acc += __ldg(row + i) * static_cast<double>(v[i]);
        }
        else
        {
          const T* v = cell + size_t(sm.x) * p.n + b;
          for (int i = 0; i < dim; ++i)
            acc += __ldg(row + i) * static_cast<double>(v[size_t(i) * p.n]);
        }
        g[q] = static_cast<T>(acc);
      }
    }
    __syncthreads();
  }
}

template <typename T, bool RIGHT, bool MATRIX>
void launch(T* data, const TransformParams& p, size_t smem, cudaStream_t s)
{
  auto kern = dof_transform_kernel<T, RIGHT, MATRIX>;
  if (smem > 48 * 1024)
    BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const size_t ntiles = ceil_div(p.nc, size_t(p.cells_per_tile));
  const size_t cap = size_t(sm_count()) * 8;
  const int grid = int(ntiles < cap ? ntiles : cap);
  kern<<<grid, kThreads, smem, s>>>(data, p);
  BX_LAUNCHED();
}
} // namespace

// data[nc][cell_size] on the device, transformed in place.  op is a bx_top (0..7).
template <typename T>
void dof_transform_device(const bx_element& e, int op, T* data, size_t nc, size_t cell_size, int n,
                          const uint32_t* cell_info, cudaStream_t s)
{
  if (op < 0 or op > 7)
    fail(BX_ERR_INVALID_ARG, "T_apply: unknown operator");
  if (n <= 0)
    fail(BX_ERR_INVALID_ARG, "T_apply: block size must be positive");
  if (cell_size != size_t(e.ndofs) * n)
    fail(BX_ERR_INVALID_ARG, "T_apply: data size per cell (" + std::to_string(cell_size)
                                 + ") is not ndofs * block_size (" + std::to_string(size_t(e.ndofs) * n) + ")");
  if (e.identity or e.tdim < 2 or e.slots.empty() or nc == 0)
    return;
  require_device(e);
  if (std::is_integral<T>::value and !e.perms)
    fail(BX_ERR_UNSUPPORTED, "T_apply: integer data needs an element whose DOF transformations are permutations");
  // reference operator -> (composed operator kind, data layout); see element.cu
  static const int kind_of[8] = {0, 1, 2, 3, 0, 2, 1, 3};
  const bool right = op >= 4;
  const int kind = kind_of[op];
  if (cell_size >= 65536)
    fail(BX_ERR_UNSUPPORTED, "T_apply: more than 65535 entries per cell");

  TransformParams p;
  p.nc = nc;
  p.cell_size = int(cell_size);
  p.n = n;
  p.ndofs = e.ndofs;
  const size_t bytes_per_cell = cell_size * sizeof(T);
  size_t cpt = std::max<size_t>(1, (32 * 1024) / bytes_per_cell);
  cpt = std::min<size_t>(cpt, 65535 / cell_size);
  cpt = std::min<size_t>(cpt, nc);
  p.cells_per_tile = int(cpt);
  const size_t smem = cpt * bytes_per_cell;
  if (smem > 200 * 1024)
    fail(BX_ERR_UNSUPPORTED, "T_apply: a single cell does not fit in shared memory");
  p.by_cell = FastDiv(unsigned(cell_size));
  p.by_n = FastDiv(unsigned(n));
  p.by_ndofs = FastDiv(unsigned(e.ndofs));
  p.cell_info = cell_info;
  p.dof_slot = e.d_dof_slot;
  p.dof_loc = e.d_dof_loc;
  p.slot_info = e.d_slot_info;
  p.gather = e.d_gather ? e.d_gather + size_t(kind) * e.gather_size : nullptr;
  p.slot_mat = e.d_slot_mat;
  p.mats = e.d_mats ? e.d_mats + size_t(kind) * e.mats_size : nullptr;

  if (e.perms)
  {
    if (right)
      launch<T, true, false>(data, p, smem, s);
    else
      launch<T, false, false>(data, p, smem, s);
  }
  else
  {
    if constexpr (std::is_integral<T>::value)
      fail(BX_ERR_UNSUPPORTED, "T_apply: integer data on a non-permutation element");
    else
    {
      if (right)
        launch<T, true, true>(data, p, smem, s);
      else
        launch<T, false, true>(data, p, smem, s);
    }
  }
}

template void dof_transform_device<double>(const bx_element&, int, double*, size_t, size_t, int, const uint32_t*,
                                           cudaStream_t);
template void dof_transform_device<float>(const bx_element&, int, float*, size_t, size_t, int, const uint32_t*,
                                          cudaStream_t);
template void dof_transform_device<int32_t>(const bx_element&, int, int32_t*, size_t, size_t, int,
                                            const uint32_t*, cudaStream_t);
} // namespace bx

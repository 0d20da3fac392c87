// Batched push_forward / pull_back (Piola maps) over many cells.
//
// Replaces FiniteElement::push_forward / pull_back (reference cpp/basix/finite-element.cpp:1971-2042)
// and the per-cell kernels of cpp/basix/maps.h:56-192.  pull_back is the same kernel with
// (K, 1/detJ, J) in place of (J, detJ, K), exactly as finite-element.cpp:2038 does.
//
// Bandwidth-bound: every thread owns one (cell, row) pair, reads the row's reference values and the
// cell's small matrix, writes the mapped row.  The accumulation order of maps.h is kept and the file
// is compiled with -fmad=false, so results are bit-identical to a reference built without FMA
// contraction.
#include "common.cuh"

#include <algorithm>

namespace bx
{
namespace
{
constexpr int kThreads = 256;
constexpr int kBcastBatch = 4; // cells per warp iteration of map_bcast_kernel

// One row through the map (the arithmetic of maps.h in its order; `det` is detJ, or 1 / detJ for pull_back, and is
// only read by the contravariant maps).
template <typename T, int MAP, int RA, int CA, int IN_VS, int OUT_VS>
__device__ __forceinline__ void map_row(const T (&a)[RA][CA], const T (&u)[IN_VS], T det, T (&r)[OUT_VS])
{
  if constexpr (MAP == BX_MAP_COVARIANT_PIOLA)
  {
#pragma unroll
    for (int i = 0; i < CA; ++i)
    {
      T acc = 0;
#pragma unroll
      for (int k = 0; k < RA; ++k)
        acc += a[k][i] * u[k];
      r[i] = acc;
    }
  }
  else if constexpr (MAP == BX_MAP_CONTRAVARIANT_PIOLA)
  {
#pragma unroll
    for (int i = 0; i < RA; ++i)
    {
      T acc = 0;
#pragma unroll
      for (int k = 0; k < CA; ++k)
        acc += a[i][k] * u[k];
      r[i] = acc / det;
    }
  }
  else if constexpr (MAP == BX_MAP_DOUBLE_COVARIANT_PIOLA)
  {
#pragma unroll
    for (int i = 0; i < CA; ++i)
#pragma unroll
      for (int j = 0; j < CA; ++j)
      {
        T acc = 0;
#pragma unroll
        for (int k = 0; k < RA; ++k)
#pragma unroll
          for (int l = 0; l < RA; ++l)
            acc += a[k][i] * u[k * RA + l] * a[l][j];
        r[i * CA + j] = acc;
      }
  }
  else
  {
    const T det2 = det * det;
#pragma unroll
    for (int i = 0; i < RA; ++i)
#pragma unroll
      for (int j = 0; j < RA; ++j)
      {
        T acc = 0;
#pragma unroll
        for (int k = 0; k < CA; ++k)
#pragma unroll
          for (int l = 0; l < CA; ++l)
            acc += a[i][k] * u[k * CA + l] * a[j][l];
        r[i * RA + j] = acc / det2;
      }
  }
}

// A is the matrix the reference kernel actually reads (RA x CA, row-major):
//   covariant          A = "K" argument:  out[i] = sum_k A[k][i] in[k]                 maps.h:73-94
//   contravariant      A = "J" argument:  out[i] = (sum_k A[i][k] in[k]) / det         maps.h:100-124
//   double covariant   A = "K" argument:  out = A^T in A                                maps.h:128-160
//   double contravar.  A = "J" argument:  out = A in A^T / det^2                        maps.h:163-192
// STAGE (used when every cell has several rows): the RA x CA matrices of the cells a CTA tile of kThreads rows
// touches are first copied to shared memory with one coalesced pass, so each thread issues global loads for its
// own row only -- with 45 rows per cell the per-thread matrix loads otherwise outnumber the row loads 3 : 1 and
// the load/store unit, not HBM, sets the pace.
template <typename T, int MAP, int RA, int CA, bool STAGE>
__global__ void __launch_bounds__(kThreads)
    map_kernel(const T* __restrict__ in, const T* __restrict__ A, const T* __restrict__ detJ,
               int invert_det, size_t nc, size_t rows, int bcast, T* __restrict__ out)
{
  constexpr int IN_VS = MAP == BX_MAP_COVARIANT_PIOLA            ? RA
                        : MAP == BX_MAP_CONTRAVARIANT_PIOLA      ? CA
                        : MAP == BX_MAP_DOUBLE_COVARIANT_PIOLA   ? RA * RA
                                                                 : CA * CA;
  constexpr int OUT_VS = MAP == BX_MAP_COVARIANT_PIOLA            ? CA
                         : MAP == BX_MAP_CONTRAVARIANT_PIOLA      ? RA
                         : MAP == BX_MAP_DOUBLE_COVARIANT_PIOLA   ? CA * CA
                                                                  : RA * RA;
  const size_t total = nc * rows;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  extern __shared__ __align__(16) unsigned char map_smem[];
  T* sA = reinterpret_cast<T*>(map_smem); // STAGE: [cells of this tile][RA * CA]
  // every thread of a CTA runs the same number of iterations (tiles), so the barriers below are uniform
  for (size_t t0 = size_t(blockIdx.x) * blockDim.x; t0 < total; t0 += stride)
  {
    const size_t t = t0 + threadIdx.x;
    size_t c_first = 0;
    if constexpr (STAGE)
    {
      c_first = t0 / rows;
      const size_t t_last = min(t0 + blockDim.x, total) - 1;
      const size_t nA = (t_last / rows - c_first + 1) * (RA * CA);
      __syncthreads(); // the previous tile's matrices are no longer read
      for (size_t i = threadIdx.x; i < nA; i += blockDim.x)
        sA[i] = A[c_first * (RA * CA) + i];
      __syncthreads();
    }
    if (t >= total)
      continue;
    const size_t c = t / rows;
    T a[RA][CA];
#pragma unroll
    for (int i = 0; i < RA; ++i)
#pragma unroll
      for (int j = 0; j < CA; ++j)
        a[i][j] = STAGE ? sA[(c - c_first) * (RA * CA) + i * CA + j] : A[c * (RA * CA) + i * CA + j];
    // bcast: one slab of reference values in[rows][IN_VS] serves every cell (the tabulated reference basis is the
    // same on all cells); it stays in L1 / L2 and only the cell matrices and the result cross HBM
    const size_t tin = bcast ? t - c * rows : t;
    T u[IN_VS];
#pragma unroll
    for (int i = 0; i < IN_VS; ++i)
      u[i] = in[tin * IN_VS + i];
    T r[OUT_VS];
    T det = T(1);
    if constexpr (MAP == BX_MAP_CONTRAVARIANT_PIOLA or MAP == BX_MAP_DOUBLE_CONTRAVARIANT_PIOLA)
      det = invert_det ? T(1.0) / detJ[c] : detJ[c];
    map_row<T, MAP, RA, CA, IN_VS, OUT_VS>(a, u, det, r);
#pragma unroll
    for (int i = 0; i < OUT_VS; ++i)
      out[t * OUT_VS + i] = r[i];
  }
}

// Several rows per cell, per-cell input rows (push_forward / pull_back of a [cell][row][value] array): the thread-per-row
// form without block barriers.  A WARP owns 32 consecutive rows of the flat (cell, row) space per iteration; the
// matrices of the cells those rows belong to (two at most when rows >= 32) are copied into a warp-private shared slot
// by the warp itself (one coalesced load), so the tile needs only __syncwarp and the warps of a CTA run independently
// (the block-staged form above spends a third of its stall samples at its two barriers per tile).  The position of a
// tile in (cell, row) terms advances incrementally (no 64-bit division in the loop).
// kMaxCells: cells a 32-row tile can touch; rows >= 2 -> at most 17.
constexpr int kWarpTileCells = 17;
template <typename T, int MAP, int RA, int CA>
__global__ void __launch_bounds__(kThreads)
    map_warp_kernel(const T* __restrict__ in, const T* __restrict__ A, const T* __restrict__ detJ, int invert_det, size_t nc,
                    uint32_t rows, T* __restrict__ out)
{
  constexpr int IN_VS = MAP == BX_MAP_COVARIANT_PIOLA            ? RA
                        : MAP == BX_MAP_CONTRAVARIANT_PIOLA      ? CA
                        : MAP == BX_MAP_DOUBLE_COVARIANT_PIOLA   ? RA * RA
                                                                 : CA * CA;
  constexpr int OUT_VS = MAP == BX_MAP_COVARIANT_PIOLA            ? CA
                         : MAP == BX_MAP_CONTRAVARIANT_PIOLA      ? RA
                         : MAP == BX_MAP_DOUBLE_COVARIANT_PIOLA   ? CA * CA
                                                                  : RA * RA;
  constexpr int MA = RA * CA;
  constexpr bool kDet = MAP == BX_MAP_CONTRAVARIANT_PIOLA or MAP == BX_MAP_DOUBLE_CONTRAVARIANT_PIOLA;
  __shared__ T sA_all[kThreads / 32][kWarpTileCells * (MA + 1)];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* sA = sA_all[warp];
  const size_t total = nc * size_t(rows);
  const size_t nwarps = size_t(gridDim.x) * (kThreads / 32);
  size_t t0 = (size_t(blockIdx.x) * (kThreads / 32) + warp) * 32;
  if (t0 >= total)
    return;
  // (cell, row) of t0 and of the tile stride, advanced incrementally
  size_t c0 = t0 / rows;
  uint32_t r0 = uint32_t(t0 - c0 * rows);
  const size_t stride = nwarps * 32;
  const size_t dc = stride / rows;
  const uint32_t dr = uint32_t(stride - dc * rows);
  const bool wide = rows >= 32; // a tile touches at most two cells
  for (; t0 < total; t0 += stride)
  {
    const uint32_t nrow = uint32_t(min(size_t(32), total - t0));
    const uint32_t ncell = (r0 + nrow - 1) / rows + 1;
    // the warp's matrices (and determinants): one coalesced pass
    for (uint32_t i = lane; i < ncell * MA; i += 32)
      sA[i] = A[c0 * MA + i];
    if constexpr (kDet)
    {
      if (lane < int(ncell))
        sA[kWarpTileCells * MA + lane] = detJ[c0 + lane];
    }
    const uint32_t rl = r0 + min(uint32_t(lane), nrow - 1); // lanes past the last row repeat it and do not store
    const uint32_t off = wide ? (rl >= rows ? 1u : 0u) : rl / rows;
    const size_t t = t0 + min(uint32_t(lane), nrow - 1);
    T u[IN_VS];
#pragma unroll
    for (int i = 0; i < IN_VS; ++i)
      u[i] = in[t * IN_VS + i];
    __syncwarp();
    T a[RA][CA];
#pragma unroll
    for (int i = 0; i < RA; ++i)
#pragma unroll
      for (int j = 0; j < CA; ++j)
        a[i][j] = sA[off * MA + i * CA + j];
    T det = T(1);
    if constexpr (kDet)
    {
      det = sA[kWarpTileCells * MA + off];
      if (invert_det)
        det = T(1.0) / det;
    }
    T r[OUT_VS];
    map_row<T, MAP, RA, CA, IN_VS, OUT_VS>(a, u, det, r);
    if (uint32_t(lane) < nrow)
    {
#pragma unroll
      for (int i = 0; i < OUT_VS; ++i)
        out[t * OUT_VS + i] = r[i];
    }
    __syncwarp(); // the slot is overwritten by the next tile
    c0 += dc;
    r0 += dr;
    if (r0 >= rows)
    {
      r0 -= rows;
      ++c0;
    }
  }
}

// Broadcast push (one slab of reference values in[rows][IN_VS] for every cell), covariant / contravariant Piola:
// a WARP owns a cell.  Lanes own ROWS for the arithmetic (the row's reference values stay in registers for the whole
// kernel, the cell matrix is read with warp-uniform loads, every component index is compile-time) and write the
// mapped rows into a warp-private shared-memory slice; the slice -- the cell's rows * OUT_VS contiguous output
// entries -- then leaves with lane-consecutive stores, 256 contiguous bytes per instruction.  The thread-per-row
// kernel above spends 63 LSU wavefronts per 45-row cell (8-byte accesses at a 24-byte stride, matrix reads from
// shared memory) and is bound by the load/store pipe.  Same accumulation order as maps.h, bit-identical results.
// Measured (45 rows, 8 M cells): 2.98e9 cells/s with the thread-per-row kernel, 3.70e9 with one cell per warp
// iteration, 4.45e9 (0.78 of the HBM roofline) with batches of four cells (batches of eight: the same).
template <typename T, int MAP, int RA, int CA, int PASSES>
__global__ void __launch_bounds__(kThreads)
    map_bcast_kernel(const T* __restrict__ in, const T* __restrict__ A, const T* __restrict__ detJ, size_t nc, int rows,
                     T* __restrict__ out)
{
  constexpr int IN_VS = MAP == BX_MAP_COVARIANT_PIOLA ? RA : CA;
  constexpr int OUT_VS = MAP == BX_MAP_COVARIANT_PIOLA ? CA : RA;
  extern __shared__ __align__(16) unsigned char bcast_smem[];
  const int lane = threadIdx.x & 31;
  const int E = rows * OUT_VS;
  constexpr int kExtra = kBcastBatch * (RA * CA + 1); // staged matrices + determinants of a batch
  T* slice = reinterpret_cast<T*>(bcast_smem) + (threadIdx.x >> 5) * (E + kExtra);
  T u[PASSES][IN_VS];
#pragma unroll
  for (int p = 0; p < PASSES; ++p)
  {
    const int r = min(lane + 32 * p, rows - 1);
#pragma unroll
    for (int k = 0; k < IN_VS; ++k)
      u[p][k] = in[r * IN_VS + k];
  }
  // A warp takes kBatch consecutive cells per iteration: their matrices (and determinants) are one contiguous run,
  // fetched with one coalesced load per 32 entries into the warp's shared slice, so the global-memory latency is
  // paid once per batch instead of once per cell (long-scoreboard stalls dominated the one-cell-per-iteration form).
  constexpr int kBatch = kBcastBatch;
  constexpr int MA = RA * CA;
  T* sm = slice + E; // [kBatch][MA] matrices, then [kBatch] determinants
  const size_t warps = size_t(gridDim.x) * (kThreads / 32);
  for (size_t c0 = (size_t(blockIdx.x) * (kThreads / 32) + (threadIdx.x >> 5)) * kBatch; c0 < nc; c0 += warps * kBatch)
  {
    const int nb = int(min(size_t(kBatch), nc - c0));
    for (int q = lane; q < nb * MA; q += 32)
      sm[q] = A[c0 * MA + q];
    if constexpr (MAP == BX_MAP_CONTRAVARIANT_PIOLA)
      if (lane < nb)
        sm[kBatch * MA + lane] = detJ[c0 + lane];
    __syncwarp();
    for (int bi = 0; bi < nb; ++bi)
    {
      const size_t c = c0 + bi;
      T a[RA][CA];
#pragma unroll
      for (int i = 0; i < RA; ++i)
#pragma unroll
        for (int j = 0; j < CA; ++j)
          a[i][j] = sm[bi * MA + i * CA + j];
      [[maybe_unused]] T det = T(1);
      if constexpr (MAP == BX_MAP_CONTRAVARIANT_PIOLA)
        det = sm[kBatch * MA + bi];
#pragma unroll
      for (int p = 0; p < PASSES; ++p)
      {
        const int r = lane + 32 * p;
        if (r < rows)
        {
#pragma unroll
          for (int i = 0; i < OUT_VS; ++i)
          {
            T acc = 0;
            if constexpr (MAP == BX_MAP_COVARIANT_PIOLA)
            {
              // out[i] = sum_k A[k][i] in[k]   (maps.h:73-94)
#pragma unroll
              for (int k = 0; k < RA; ++k)
                acc += a[k][i] * u[p][k];
            }
            else
            {
              // out[i] = (sum_k A[i][k] in[k]) / det   (maps.h:100-124)
#pragma unroll
              for (int k = 0; k < CA; ++k)
                acc += a[i][k] * u[p][k];
              acc = acc / det;
            }
            slice[r * OUT_VS + i] = acc;
          }
        }
      }
      __syncwarp();
      T* o = out + c * size_t(E);
      for (int e = lane; e < E; e += 32)
        o[e] = slice[e];
      __syncwarp(); // the slice is overwritten by the next cell
    }
  }
}

// identity map: u(i, j) = U(i, j)  (finite-element.h:632-640)
template <typename T>
__global__ void __launch_bounds__(kThreads) copy_kernel(const T* __restrict__ in, size_t n, size_t period, T* __restrict__ out)
{
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += stride)
    out[t] = in[period ? t % period : t];
}

// Persistent grid-stride launch: exactly the CTAs that are resident at once (no ragged last wave).
template <typename K>
int grid_for(K kern, size_t n, size_t smem)
{
  int per_sm = 1;
  BX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  const size_t want = ceil_div(n, size_t(kThreads));
  const size_t cap = size_t(sm_count()) * std::max(per_sm, 1);
  return int(want < cap ? (want == 0 ? 1 : want) : cap);
}

template <typename T, int MAP, int RA, int CA>
void launch_one(const T* in, const T* A, const T* detJ, int inv, size_t nc, size_t rows, int bcast, T* out, cudaStream_t s)
{
  if constexpr (MAP == BX_MAP_COVARIANT_PIOLA or MAP == BX_MAP_CONTRAVARIANT_PIOLA)
  {
    constexpr int OUT_VS = MAP == BX_MAP_COVARIANT_PIOLA ? CA : RA;
    const size_t E = rows * OUT_VS;
    if (bcast and !inv and rows <= 256)
    {
      const int slots = int((rows + 31) / 32);
      const size_t smem = size_t(kThreads / 32) * (E + kBcastBatch * (RA * CA + 1)) * sizeof(T);
      auto go = [&](auto kern)
      {
        // 243..256 rows of three doubles need more than the 48 KB a kernel gets without opting in
        if (smem > 48 * 1024)
          BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        int per_sm = 1;
        BX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
        const size_t want = ceil_div(nc, size_t(kThreads / 32) * kBcastBatch);
        const size_t cap = size_t(sm_count()) * std::max(per_sm, 1);
        kern<<<unsigned(std::max<size_t>(1, std::min(want, cap))), kThreads, smem, s>>>(in, A, detJ, nc, int(rows), out);
      };
      switch (slots)
      {
      case 1: go(map_bcast_kernel<T, MAP, RA, CA, 1>); break;
      case 2: go(map_bcast_kernel<T, MAP, RA, CA, 2>); break;
      case 3: go(map_bcast_kernel<T, MAP, RA, CA, 3>); break;
      case 4: go(map_bcast_kernel<T, MAP, RA, CA, 4>); break;
      case 5: go(map_bcast_kernel<T, MAP, RA, CA, 5>); break;
      case 6: go(map_bcast_kernel<T, MAP, RA, CA, 6>); break;
      case 7: go(map_bcast_kernel<T, MAP, RA, CA, 7>); break;
      default: go(map_bcast_kernel<T, MAP, RA, CA, 8>); break;
      }
      return;
    }
  }
  static const int warp_form = []
  {
    const char* v = std::getenv("BX_MAP_WARP"); // development knob: 0 selects the block-staged kernel
    return v ? std::atoi(v) : 1;
  }();
  if (rows >= 2 and !bcast and warp_form and rows < (size_t(1) << 31))
  {
    auto kern = map_warp_kernel<T, MAP, RA, CA>;
    int per_sm = 1;
    BX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, 0));
    const size_t want = ceil_div(nc * rows, size_t(kThreads));
    const size_t cap = size_t(sm_count()) * std::max(per_sm, 1);
    kern<<<unsigned(std::max<size_t>(1, std::min(want, cap))), kThreads, 0, s>>>(in, A, detJ, inv, nc, uint32_t(rows), out);
  }
  else if (rows >= 2)
  {
    auto kern = map_kernel<T, MAP, RA, CA, true>;
    // a tile of kThreads rows touches at most kThreads / rows + 2 cells
    const size_t smem = (kThreads / rows + 2) * (RA * CA) * sizeof(T);
    kern<<<grid_for(kern, nc * rows, smem), kThreads, smem, s>>>(in, A, detJ, inv, nc, rows, bcast, out);
  }
  else
  {
    auto kern = map_kernel<T, MAP, RA, CA, false>;
    kern<<<grid_for(kern, nc * rows, 0), kThreads, 0, s>>>(in, A, detJ, inv, nc, rows, bcast, out);
  }
}

template <typename T, int MAP, int RA>
void launch_ca(int ca, const T* in, const T* A, const T* detJ, int inv, size_t nc, size_t rows, int bcast, T* out,
               cudaStream_t s)
{
  switch (ca)
  {
  case 1: launch_one<T, MAP, RA, 1>(in, A, detJ, inv, nc, rows, bcast, out, s); break;
  case 2: launch_one<T, MAP, RA, 2>(in, A, detJ, inv, nc, rows, bcast, out, s); break;
  case 3: launch_one<T, MAP, RA, 3>(in, A, detJ, inv, nc, rows, bcast, out, s); break;
  default: fail(BX_ERR_INVALID_ARG, "map: Jacobian dimensions must be between 1 and 3");
  }
  BX_LAUNCHED();
}

template <typename T, int MAP>
void launch_ra(int ra, int ca, const T* in, const T* A, const T* detJ, int inv, size_t nc, size_t rows, int bcast,
               T* out, cudaStream_t s)
{
  switch (ra)
  {
  case 1: launch_ca<T, MAP, 1>(ca, in, A, detJ, inv, nc, rows, bcast, out, s); break;
  case 2: launch_ca<T, MAP, 2>(ca, in, A, detJ, inv, nc, rows, bcast, out, s); break;
  case 3: launch_ca<T, MAP, 3>(ca, in, A, detJ, inv, nc, rows, bcast, out, s); break;
  default: fail(BX_ERR_INVALID_ARG, "map: Jacobian dimensions must be between 1 and 3");
  }
}
} // namespace

// Value size of the mapped data: push -> compute_value_size (finite-element.cpp:42-59) with
// dim = gdim; pull -> the reference value size (tdim-based).
int map_value_size(int map_type, int pull, int in_vs, int gdim, int tdim)
{
  const int dim = pull ? tdim : gdim;
  switch (map_type)
  {
  case BX_MAP_IDENTITY: return pull ? in_vs : 1;
  case BX_MAP_COVARIANT_PIOLA:
  case BX_MAP_CONTRAVARIANT_PIOLA: return dim;
  case BX_MAP_DOUBLE_COVARIANT_PIOLA:
  case BX_MAP_DOUBLE_CONTRAVARIANT_PIOLA: return dim * dim;
  case BX_MAP_L2_PIOLA: fail(BX_ERR_UNSUPPORTED, "Map not implemented");
  default: fail(BX_ERR_INVALID_ARG, "Mapping not yet implemented");
  }
}

template <typename T>
bool push_forward_broadcast_flat(int map_type, const T* U, const T* J, const T* detJ, const T* K, size_t nc, size_t rows,
                                 int gdim, int tdim, T* out, cudaStream_t s); // geometry.cu

// Device pointers.  U[nc][rows][in_vs] (bcast: U[rows][in_vs], shared by all cells), J[nc][gdim][tdim], detJ[nc],
// K[nc][tdim][gdim].
template <typename T>
void map_device(int map_type, int pull, const T* U, const T* J, const T* detJ, const T* K, size_t nc, size_t rows,
                int in_vs, int gdim, int tdim, T* out, cudaStream_t s, int bcast)
{
  const int out_vs = map_value_size(map_type, pull, in_vs, gdim, tdim);
  if (nc == 0 or rows == 0)
    return;
  // The reference kernel's "J" and "K" arguments: push (J, K); pull (K, J).
  const T* Jarg = pull ? K : J;
  const T* Karg = pull ? J : K;
  const int Jr = pull ? tdim : gdim, Jc = pull ? gdim : tdim; // shape of the "J" argument
  const int Kr = Jc, Kc = Jr;                                  // shape of the "K" argument
  const int dim_in = pull ? gdim : tdim;
  auto need = [&](int want)
  {
    if (in_vs != want)
      fail(BX_ERR_INVALID_ARG, "map: value size of the input (" + std::to_string(in_vs)
                                   + ") does not match the map (" + std::to_string(want) + ")");
  };
  // push_forward_broadcast with a Piola map: the flat (cell, row) walk of the fused physical-basis kernel
  if (bcast and !pull and in_vs == tdim and (map_type == BX_MAP_COVARIANT_PIOLA ? K != nullptr : (J != nullptr and detJ != nullptr))
      and push_forward_broadcast_flat<T>(map_type, U, J, detJ, K, nc, rows, gdim, tdim, out, s))
    return;
  switch (map_type)
  {
  case BX_MAP_IDENTITY:
    if (in_vs != out_vs)
      fail(BX_ERR_INVALID_ARG, "map: identity map needs matching value sizes");
    copy_kernel<T><<<grid_for(copy_kernel<T>, nc * rows * in_vs, 0), kThreads, 0, s>>>(U, nc * rows * in_vs,
                                                                                      bcast ? rows * in_vs : 0, out);
    BX_LAUNCHED();
    return;
  case BX_MAP_COVARIANT_PIOLA:
    need(dim_in);
    launch_ra<T, BX_MAP_COVARIANT_PIOLA>(Kr, Kc, U, Karg, detJ, pull, nc, rows, bcast, out, s);
    return;
  case BX_MAP_CONTRAVARIANT_PIOLA:
    need(dim_in);
    launch_ra<T, BX_MAP_CONTRAVARIANT_PIOLA>(Jr, Jc, U, Jarg, detJ, pull, nc, rows, bcast, out, s);
    return;
  case BX_MAP_DOUBLE_COVARIANT_PIOLA:
    need(dim_in * dim_in);
    launch_ra<T, BX_MAP_DOUBLE_COVARIANT_PIOLA>(Kr, Kc, U, Karg, detJ, pull, nc, rows, bcast, out, s);
    return;
  case BX_MAP_DOUBLE_CONTRAVARIANT_PIOLA:
    need(dim_in * dim_in);
    launch_ra<T, BX_MAP_DOUBLE_CONTRAVARIANT_PIOLA>(Jr, Jc, U, Jarg, detJ, pull, nc, rows, bcast, out, s);
    return;
  default: fail(BX_ERR_UNSUPPORTED, "Map not implemented");
  }
}

template void map_device<double>(int, int, const double*, const double*, const double*, const double*, size_t, size_t, int,
                                 int, int, double*, cudaStream_t, int);
template void map_device<float>(int, int, const float*, const float*, const float*, const float*, size_t, size_t, int, int,
                                int, float*, cudaStream_t, int);
} // namespace bx

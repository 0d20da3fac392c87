// Register-path tabulate for low-degree simplex elements (P1 / P2 and friends).
//
// Same contract as the other paths of FiniteElement::tabulate (reference cpp/basix/finite-element.cpp:1834-1888):
//   basis[d][pt][n] = sum_k C[n][k] * P[d][k][pt],   n = dof_ordering[dof] * vs + j.
// For these elements the whole polynomial-set jet of a point (K basis functions x NDER derivatives, at most 40
// doubles) fits in one thread's registers and the contraction is a few dozen FMAs, far below what a tensor-core
// tile needs to pay off: the kernel is pure HBM streaming (16-24 B in, NDER*N*8 B out per point).  One thread
// owns one point: jets by the register recurrences of jets.cuh (polyset.cpp:66-120, 1601-1743, 1745-2020),
// contraction against the norm-scaled coefficient operand broadcast from shared memory, results staged through
// a warp-private shared tile that holds, per derivative slab, exactly one contiguous run of the [d][pt][n] table and
// leaves by TMA bulk copies.
#include "element.cuh"
#include "jets.cuh"

#include <algorithm>
#include <cstdlib>

namespace bx
{
namespace small
{
constexpr int kThreads = 128; // points per CTA tile

template <int K, int NDER>
struct RegSink
{
  double (&J)[K][NDER];
  __device__ __forceinline__ void operator()(int s, const double (&jet)[NDER]) const
  {
#pragma unroll
    for (int d = 0; d < NDER; ++d)
      J[s][d] = jet[d];
  }
};

// T = type of the points and of the table (double, or float for the float32 API: loads widen, stores narrow,
// the arithmetic is double either way).
//
// Warps are independent (no block barrier after the operand is in shared memory): a warp owns R * 32 consecutive
// points -- lane l the points l, l + 32, ... -- and a private shared tile [NDER][R * 32][N] typed like the table, i.e.
// for every derivative slab exactly the bytes of one contiguous run of the [d][pt][n] table.  All point loads of the
// warp are issued before the first recurrence; each finished slab leaves as ONE cp.async.bulk shared -> global copy
// issued by its own lane, and the warp retires as soon as the copy engine has read the tile.  The grid is one warp
// per tile (not persistent): with 1 M points the kernel lives for ~20 us and a persistent grid loses a quarter of
// that to the imbalance of 2.2 tiles per resident warp, which the block scheduler does not have.
template <int TDIM, int DEG, int ND, int R, typename T>
__global__ void __launch_bounds__(kThreads)
    tabulate_small_kernel(const T* __restrict__ x, size_t npts, const double* __restrict__ Cq, T* __restrict__ basis, int N,
                          int bulk)
{
  constexpr int NDER = jets::nder(TDIM, ND);
  constexpr int K = jets::psize(TDIM, DEG);
  constexpr int PW = 32 * R; // points per warp
  extern __shared__ __align__(128) unsigned char smem_small[];
  double* s_C = reinterpret_cast<double*>(smem_small); // [N][K]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* tile = reinterpret_cast<T*>(smem_small + ((size_t(N) * K * 8 + 127) & ~size_t(127))) + size_t(warp) * NDER * PW * N;
  for (int i = threadIdx.x; i < N * K; i += kThreads)
    s_C[i] = Cq[i];
  __syncthreads();

  const size_t pt0 = (size_t(blockIdx.x) * (kThreads / 32) + warp) * PW;
  if (pt0 >= npts)
    return;
  const unsigned np = unsigned(min(size_t(PW), npts - pt0));
  double xr[R][TDIM];
#pragma unroll
  for (int r = 0; r < R; ++r)
  {
    const size_t pt = min(pt0 + r * 32 + lane, npts - 1); // ragged tail: recompute the last point, never stored
#pragma unroll
    for (int a = 0; a < TDIM; ++a)
      xr[r][a] = double(x[TDIM * pt + a]);
  }
#pragma unroll
  for (int r = 0; r < R; ++r)
  {
    double J[K][NDER];
    RegSink<K, NDER> sink{J};
    if constexpr (TDIM == 1)
      jets::interval<DEG, ND>(xr[r][0], sink);
    else if constexpr (TDIM == 2)
      jets::triangle<DEG, ND>(xr[r][0], xr[r][1], sink);
    else
      jets::tetrahedron<DEG, ND>(xr[r][0], xr[r][1], xr[r][2], sink);
    T* mine = tile + (r * 32 + lane) * N;
    for (int n = 0; n < N; ++n)
    {
      double acc[NDER];
#pragma unroll
      for (int d = 0; d < NDER; ++d)
        acc[d] = 0.0;
#pragma unroll
      for (int s = 0; s < K; ++s)
      {
        const double c = s_C[n * K + s];
#pragma unroll
        for (int d = 0; d < NDER; ++d)
          acc[d] = fma(c, J[s][d], acc[d]);
      }
#pragma unroll
      for (int d = 0; d < NDER; ++d)
        mine[d * (PW * N) + n] = static_cast<T>(acc[d]);
    }
  }
  // each derivative slab of the warp's tile is one contiguous run of np * N entries in global memory
  const unsigned run = np * unsigned(N);
  if (bulk and np == PW)
  {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane < NDER)
    {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                       basis + (size_t(lane) * npts + pt0) * N),
                   "r"(uint32_t(__cvta_generic_to_shared(tile + lane * (PW * N)))), "r"(run * uint32_t(sizeof(T)))
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    return;
  }
  __syncwarp();
#pragma unroll 1
  for (int d = 0; d < NDER; ++d)
  {
    T* g = basis + (size_t(d) * npts + pt0) * N;
    const T* t = tile + d * (PW * N);
    for (unsigned i = lane; i < run; i += 32)
      g[i] = t[i];
  }
}

template <int TDIM, int DEG, int ND, int R, typename T>
void launch_r(const bx_element& e, const T* x, size_t npts, T* basis, size_t smem, int bulk, cudaStream_t s)
{
  auto kern = tabulate_small_kernel<TDIM, DEG, ND, R, T>;
  if (smem > 48 * 1024)
    BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const size_t nblocks = ceil_div(npts, size_t(kThreads) * R);
  kern<<<unsigned(nblocks), kThreads, smem, s>>>(x, npts, e.d_Cq, basis, e.N, bulk);
  BX_LAUNCHED();
}

// shared memory of one CTA with R * 32 points per warp
template <typename T>
size_t smem_bytes(int N, int K, int NDER, int R)
{
  return ((size_t(N) * K * 8 + 127) & ~size_t(127)) + size_t(kThreads / 32) * NDER * 32 * R * N * sizeof(T);
}

template <int TDIM, int DEG, int ND, typename T>
bool launch(const bx_element& e, const T* x, size_t npts, T* basis, cudaStream_t s, bool dry_run)
{
  constexpr int NDER = jets::nder(TDIM, ND);
  constexpr int K = jets::psize(TDIM, DEG);
  static_assert(K * NDER <= 40, "jet does not fit in registers");
  const int N = e.N;
  if (e.Cq.empty() or smem_bytes<T>(N, K, NDER, 1) > 96 * 1024)
    return false;
  if (dry_run)
    return true;
  if (!e.d_Cq)
    return false;
  if (ceil_div(npts, size_t(kThreads)) > 0x7fffffffu)
    fail(BX_ERR_INVALID_ARG, "tabulate (register path): too many points for one launch");
  // points per warp: the largest of 128 / 64 / 32 whose tiles leave room for several CTAs per SM
  static const int forced = []
  {
    const char* v = std::getenv("BX_SMALL_R");
    return v ? std::atoi(v) : 0;
  }();
  int R = 1;
  for (int r : {4, 2})
    if (smem_bytes<T>(N, K, NDER, r) <= 36 * 1024)
    {
      R = r;
      break;
    }
  if (forced == 1 or forced == 2 or forced == 4)
    if (smem_bytes<T>(N, K, NDER, forced) <= 96 * 1024)
      R = forced;
  // TMA bulk copy-out needs every slab run of the table to start on a 16-byte boundary (the runs are 32 * R * N
  // entries long, a multiple of 16 bytes): base aligned and slab pitch npts * N a multiple of 16 bytes
  const size_t per16 = 16 / sizeof(T);
  const int bulk = (reinterpret_cast<uintptr_t>(basis) % 16 == 0 and (npts * size_t(N)) % per16 == 0
                    and (size_t(32) * N) % per16 == 0)
                       ? 1
                       : 0;
  const size_t smem = smem_bytes<T>(N, K, NDER, R);
  if (R == 4)
    launch_r<TDIM, DEG, ND, 4, T>(e, x, npts, basis, smem, bulk, s);
  else if (R == 2)
    launch_r<TDIM, DEG, ND, 2, T>(e, x, npts, basis, smem, bulk, s);
  else
    launch_r<TDIM, DEG, ND, 1, T>(e, x, npts, basis, smem, bulk, s);
  return true;
}

#define BX_SMALL(TD, DEG, ND)                                                                            \
  if (e.tdim == TD and e.superdegree == DEG and nd == ND)                                                   \
    return launch<TD, DEG, ND, T>(e, x, npts, basis, s, dry_run);

// Menu: every (cell, degree, nd) whose jet is at most 40 doubles.  Returns false when off the menu.
template <typename T>
bool launch_any(const bx_element& e, int nd, const T* x, size_t npts, T* basis, cudaStream_t s, bool dry_run)
{
  if (e.polyset_type != BX_POLYSET_STANDARD or e.N < 2)
    return false;
  if (e.cell_type != BX_CELL_INTERVAL and e.cell_type != BX_CELL_TRIANGLE and e.cell_type != BX_CELL_TETRAHEDRON)
    return false;
  BX_SMALL(1, 1, 0) BX_SMALL(1, 1, 1) BX_SMALL(1, 1, 2) BX_SMALL(1, 1, 3)
  BX_SMALL(1, 2, 0) BX_SMALL(1, 2, 1) BX_SMALL(1, 2, 2) BX_SMALL(1, 2, 3)
  BX_SMALL(2, 1, 0) BX_SMALL(2, 1, 1) BX_SMALL(2, 1, 2) BX_SMALL(2, 1, 3)
  BX_SMALL(2, 2, 0) BX_SMALL(2, 2, 1) BX_SMALL(2, 2, 2)
  BX_SMALL(3, 1, 0) BX_SMALL(3, 1, 1) BX_SMALL(3, 1, 2)
  BX_SMALL(3, 2, 0) BX_SMALL(3, 2, 1)
  return false;
}
template bool launch_any<double>(const bx_element&, int, const double*, size_t, double*, cudaStream_t, bool);
template bool launch_any<float>(const bx_element&, int, const float*, size_t, float*, cudaStream_t, bool);
} // namespace small
} // namespace bx

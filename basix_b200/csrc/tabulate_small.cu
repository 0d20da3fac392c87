// Register-path tabulate for low-degree simplex elements (P1 / P2 and friends).
//
// Same contract as the other paths of FiniteElement::tabulate (reference cpp/basix/finite-element.cpp:1834-1888):
//   basis[d][pt][n] = sum_k C[n][k] * P[d][k][pt],   n = dof_ordering[dof] * vs + j.
// For these elements the whole polynomial-set jet of a point (K basis functions x NDER derivatives, at most 40
// doubles) fits in one thread's registers and the contraction is a few dozen FMAs, far below what a tensor-core
// tile needs to pay off: the kernel is pure HBM streaming (16-24 B in, NDER*N*8 B out per point).  One thread
// owns one point: jets by the register recurrences of jets.cuh (polyset.cpp:66-120, 1601-1743, 1745-2020),
// contraction against the norm-scaled coefficient operand broadcast from shared memory, results staged through
// an odd-pitch shared tile so that every global store instruction writes one contiguous, fully coalesced run
// of the [d][pt][n] table.
#include "element.cuh"
#include "jets.cuh"

#include <algorithm>

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
template <int TDIM, int DEG, int ND, typename T>
__global__ void __launch_bounds__(kThreads)
    tabulate_small_kernel(const T* __restrict__ x, size_t npts, const double* __restrict__ Cq,
                          T* __restrict__ basis, int N, int NP, unsigned inv_n, int bulk)
{
  constexpr int NDER = jets::nder(TDIM, ND);
  constexpr int K = jets::psize(TDIM, DEG);
  extern __shared__ __align__(16) double smem_small[];
  double* s_C = smem_small;                     // [N][K]
  double* tile = smem_small + ((N * K + 1) & ~1); // [NDER][kThreads][NP]
  for (int i = threadIdx.x; i < N * K; i += kThreads)
    s_C[i] = Cq[i];
  __syncthreads();

  const size_t ntiles = ceil_div(npts, size_t(kThreads));
  for (size_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x)
  {
    const size_t pt0 = tl * kThreads;
    const unsigned np = unsigned(min(size_t(kThreads), npts - pt0));
    const size_t pt = min(pt0 + threadIdx.x, npts - 1); // ragged tail: recompute the last point, never stored
    double J[K][NDER];
    RegSink<K, NDER> sink{J};
    if constexpr (TDIM == 1)
      jets::interval<DEG, ND>(double(x[pt]), sink);
    else if constexpr (TDIM == 2)
      jets::triangle<DEG, ND>(double(x[2 * pt]), double(x[2 * pt + 1]), sink);
    else
      jets::tetrahedron<DEG, ND>(double(x[3 * pt]), double(x[3 * pt + 1]), double(x[3 * pt + 2]), sink);

    double* mine = tile + threadIdx.x * NP;
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
        mine[d * (kThreads * NP) + n] = acc[d];
    }
    // each derivative slab of the tile is one contiguous run of np * N doubles in global memory
    const unsigned run = np * unsigned(N);
    if constexpr (sizeof(T) == 8)
    {
      if (bulk and (run & 1) == 0)
      {
        // dense tile rows (odd N), 16-byte aligned slabs: one TMA bulk copy shared -> global per slab; the CTA only
        // waits until the engine has read the tile, the HBM write completes in the background
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncthreads();
        if (threadIdx.x == 0)
        {
#pragma unroll
          for (int d = 0; d < NDER; ++d)
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(
                             basis + (size_t(d) * npts + pt0) * N),
                         "r"(uint32_t(__cvta_generic_to_shared(tile + d * (kThreads * NP)))), "r"(run * 8u)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
        continue;
      }
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < NDER; ++d)
    {
      T* g = basis + (size_t(d) * npts + pt0) * N;
      const double* t = tile + d * (kThreads * NP);
      for (unsigned i = threadIdx.x; i < run; i += kThreads)
      {
        const unsigned p = __umulhi(i, inv_n), n = i - p * unsigned(N); // p = i / N (exact: i < 2^16)
        g[i] = static_cast<T>(t[p * NP + n]);
      }
    }
    __syncthreads();
  }
}

template <int TDIM, int DEG, int ND, typename T>
bool launch(const bx_element& e, const T* x, size_t npts, T* basis, cudaStream_t s, bool dry_run)
{
  constexpr int NDER = jets::nder(TDIM, ND);
  constexpr int K = jets::psize(TDIM, DEG);
  static_assert(K * NDER <= 40, "jet does not fit in registers");
  const int N = e.N, NP = N | 1;
  const size_t smem = (size_t((N * K + 1) & ~1) + size_t(NDER) * kThreads * NP) * sizeof(double);
  if (e.Cq.empty() or smem > 96 * 1024 or size_t(kThreads) * N >= 65536)
    return false;
  if (dry_run)
    return true;
  if (!e.d_Cq)
    return false;
  auto kern = tabulate_small_kernel<TDIM, DEG, ND, T>;
  if (smem > 48 * 1024)
    BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 0;
  BX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  const size_t ntiles = ceil_div(npts, size_t(kThreads));
  const size_t cap = size_t(sm_count()) * std::max(per_sm, 1);
  const unsigned inv_n = N > 1 ? unsigned((1ull << 32) / unsigned(N) + 1) : 0u;
  if (N == 1)
    fail(BX_ERR_RUNTIME, "tabulate (register path): single-column element"); // never on the menu (see below)
  // TMA bulk copy-out: dense tile rows and every slab region 16-byte aligned
  const int bulk = (NP == N and reinterpret_cast<uintptr_t>(basis) % 16 == 0 and (npts * size_t(N)) % 2 == 0
                    and (size_t(NDER) * kThreads * NP) % 2 == 0 and ((N * K + 1) & ~1) % 2 == 0)
                       ? 1
                       : 0;
  kern<<<unsigned(std::min(ntiles, cap)), kThreads, smem, s>>>(x, npts, e.d_Cq, basis, N, NP, inv_n, bulk);
  BX_LAUNCHED();
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

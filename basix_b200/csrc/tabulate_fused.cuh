// Fused FiniteElement::tabulate for simplex elements: polyset recurrences + coefficient contraction +
// transposed store in ONE persistent, warp-specialised kernel.  The table P[d][k][pt] of the reference
// (cpp/basix/finite-element.cpp:1845-1851) never exists in global memory.
//
// CTA = 12 warps (3 per SM sub-partition), one CTA per SM, 32-point tiles, grid-stride over tiles:
//   producer warps 8..11        lane = point.  Each walks the jet recurrences (jets.cuh) in registers for
//   (one per sub-partition)     ITS share of the index tree (first-level sub-trees split over the four
//                               producers, so the FP64 pipe of every sub-partition carries the same load)
//                               and streams every finished basis function (all derivatives, 32 points)
//                               into a shared-memory ring of K-chunks.
//   consumer warps 0..7         FP64 tensor cores (mma.sync m8n8k4 = SASS DMMA).  Warp w owns the 8-point
//                               row tile (w & 3) and half of the derivative slabs (w >> 2); its accumulators
//                               for ALL output columns of those slabs stay in registers for the K sweep:
//                                   D[pt][dof] += P[d][k][pt] * Cs[dof][k]        (A = ring, B = Cs)
//                               A-fragments come from the ring, B-fragments from the (norm-scaled,
//                               production-ordered, zero-padded) coefficient matrix resident in shared
//                               memory.  After the last chunk the fragment is stored straight to
//                               basis[d][pt][dof] -- the accumulator layout already is (point, dof), so the
//                               reference's transpose pass (finite-element.cpp:1873-1886) vanishes.
//   ring                        full/empty mbarriers per stage; producers run up to a whole tile ahead.
//
// Shared-memory layouts are padded so every fragment load is bank-conflict free:
//   ring chunk  [d][kk (8)][36]   point stride 1, basis stride 36 (== 4 mod 16)
//   Cs          [Npad][KS]        KS == 4 mod 16
#pragma once

#include "element.cuh"
#include "jets.cuh"

#include <algorithm>

namespace bx
{
namespace fused
{
constexpr int kTilePts = 32;
constexpr int kConsumers = 8;
constexpr int kProducers = 4;
constexpr int kThreads = 32 * (kConsumers + kProducers);
constexpr int kChunkK = 8;    // basis functions per ring chunk (two k-steps of 4)
constexpr int kPtStride = 36; // doubles between consecutive basis functions inside a chunk

__host__ __device__ constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }
// row stride of Cs: >= Kpad and == 4 (mod 16)
__host__ __device__ constexpr int cs_stride(int kpad) { return kpad + ((4 - kpad % 16) + 16) % 16; }

template <int TDIM, int DEG, int ND>
struct Cfg
{
  static constexpr int NDER = jets::nder(TDIM, ND);
  static constexpr int K = jets::psize(TDIM, DEG);
  static constexpr int KPAD = round_up(K, 4);
  static constexpr int KSTEPS = KPAD / 4;
  static constexpr int NCHUNK = (KPAD + kChunkK - 1) / kChunkK;
  static constexpr int NT = (K + 7) / 8; // 8-column output tiles (N <= K for scalar elements)
  static constexpr int NPAD = NT * 8;
  static constexpr int KS = cs_stride(KPAD);
  static constexpr int CHUNK_DOUBLES = NDER * kChunkK * kPtStride;
  static constexpr int CS_DOUBLES = NPAD * KS;
  // consumer halves split the derivative slabs when there are at least two, else the column tiles
  static constexpr bool SPLIT_D = NDER >= 2;
  static constexpr int D_LO(int h) { return SPLIT_D ? (h == 0 ? 0 : (NDER + 1) / 2) : 0; }
  static constexpr int D_HI(int h) { return SPLIT_D ? (h == 0 ? (NDER + 1) / 2 : NDER) : NDER; }
  static constexpr int T_LO(int h) { return SPLIT_D ? 0 : (h == 0 ? 0 : (NT + 1) / 2); }
  static constexpr int T_HI(int h) { return SPLIT_D ? NT : (h == 0 ? (NT + 1) / 2 : NT); }
  static constexpr int stages(int smem_budget)
  {
    int s = (smem_budget - CS_DOUBLES * 8 - 256) / (CHUNK_DOUBLES * 8);
    s = s > 2 * NCHUNK ? 2 * NCHUNK : s;
    return s > 16 ? 16 : s; // the barrier block holds 2 x 16 mbarriers
  }
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a)
{
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity)
{
  uint32_t ok;
  do
  {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(a), "r"(parity)
                 : "memory");
  } while (!ok);
}
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// Producer-side sink: writes finished jets into the ring and walks the chunk barriers in order.  Every
// producer visits every chunk of a tile (wait empty -> write its entries, possibly none -> arrive full), so
// the full barriers always expect kProducers * 32 arrivals.
template <typename C>
struct RingSink
{
  double* ring;   // stage 0 of the ring
  uint32_t full0; // shared address of full[0]; empty[s] follows the S full barriers
  int nstages;
  int lane;
  uint32_t stage, phase; // ring slot of chunk `cur` (or of the next chunk to open when cur < 0)
  int cur;               // chunk of the current tile that is open, -1 before the first

  __device__ __forceinline__ void advance()
  {
    if (++stage == uint32_t(nstages))
    {
      stage = 0;
      phase ^= 1;
    }
  }
  __device__ __forceinline__ void open_next()
  {
    if (cur >= 0)
    {
      mbar_arrive(full0 + 8 * stage);
      advance();
    }
    ++cur;
    mbar_wait(full0 + 8 * (nstages + stage), phase ^ 1);
  }
  __device__ __forceinline__ void operator()(int seq, const double (&J)[C::NDER])
  {
    const int chunk = seq / kChunkK;
    while (cur < chunk)
      open_next();
    double* dst = ring + size_t(stage) * C::CHUNK_DOUBLES + (seq % kChunkK) * kPtStride + lane;
#pragma unroll
    for (int d = 0; d < C::NDER; ++d)
      dst[d * (kChunkK * kPtStride)] = J[d];
  }
  __device__ __forceinline__ void finish_tile()
  {
    while (cur < C::NCHUNK - 1)
      open_next();
    mbar_arrive(full0 + 8 * stage);
    advance();
    cur = -1;
  }
};

template <int TDIM, int DEG, int ND, int ROLE>
__device__ __forceinline__ void producer_loop(const double* __restrict__ x, size_t npts, size_t ntiles, double* ring,
                                              uint32_t full0, int nstages, int lane)
{
  using C = Cfg<TDIM, DEG, ND>;
  using Own = jets::OwnSplit<TDIM, DEG, kProducers, ROLE>;
  RingSink<C> sink{ring, full0, nstages, lane, 0u, 0u, -1};
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
  {
    size_t pt = tile * kTilePts + lane;
    pt = pt < npts ? pt : npts - 1; // ragged tail: recompute the last point, stores are predicated
    if constexpr (TDIM == 1)
      jets::interval<DEG, ND, Own>(x[pt], sink);
    else if constexpr (TDIM == 2)
      jets::triangle<DEG, ND, Own>(x[2 * pt], x[2 * pt + 1], sink);
    else
      jets::tetrahedron<DEG, ND, Own>(x[3 * pt], x[3 * pt + 1], x[3 * pt + 2], sink);
    // rows K..KPAD-1 of the last k-step: B is zero there, A must be finite
    if constexpr (C::KPAD > C::K and ROLE == 0)
    {
      double Z[C::NDER];
#pragma unroll
      for (int d = 0; d < C::NDER; ++d)
        Z[d] = 0.0;
#pragma unroll
      for (int s = C::K; s < C::KPAD; ++s)
        sink(s, Z);
    }
    sink.finish_tile();
  }
}

template <int TDIM, int DEG, int ND, int H>
__device__ __forceinline__ void consumer_loop(const double* ring, const double* Cs, uint32_t full0, int nstages,
                                              int mt, int lane, double* __restrict__ basis, size_t npts, int N,
                                              size_t ntiles)
{
  using C = Cfg<TDIM, DEG, ND>;
  constexpr int DLO = C::D_LO(H), DHI = C::D_HI(H), TLO = C::T_LO(H), THI = C::T_HI(H);
  constexpr int DN = DHI - DLO, TN = THI - TLO;
  const int g = lane >> 2, t = lane & 3;
  uint32_t stage = 0, phase = 0;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
  {
    double acc[DN > 0 ? DN : 1][TN > 0 ? TN : 1][2];
#pragma unroll
    for (int d = 0; d < DN; ++d)
#pragma unroll
      for (int j = 0; j < TN; ++j)
        acc[d][j][0] = acc[d][j][1] = 0.0;
#pragma unroll 1
    for (int c = 0; c < C::NCHUNK; ++c)
    {
      mbar_wait(full0 + 8 * stage, phase);
      const double* chunk = ring + size_t(stage) * C::CHUNK_DOUBLES + mt * 8 + g;
#pragma unroll
      for (int ks = 0; ks < kChunkK / 4; ++ks)
      {
        const int kstep = c * (kChunkK / 4) + ks;
        if (kstep < C::KSTEPS)
        {
          double b[TN > 0 ? TN : 1];
#pragma unroll
          for (int j = 0; j < TN; ++j)
            b[j] = Cs[((TLO + j) * 8 + g) * C::KS + kstep * 4 + t];
#pragma unroll
          for (int d = 0; d < DN; ++d)
          {
            const double a = chunk[((DLO + d) * kChunkK + ks * 4 + t) * kPtStride];
#pragma unroll
            for (int j = 0; j < TN; ++j)
              dmma(acc[d][j][0], acc[d][j][1], a, b[j]);
          }
        }
      }
      __syncwarp();
      if (lane == 0)
        mbar_arrive(full0 + 8 * (nstages + stage));
      if (++stage == uint32_t(nstages))
      {
        stage = 0;
        phase ^= 1;
      }
    }
    // ---- epilogue: accumulator fragment -> basis[d][pt][n] ----
    const size_t pt = tile * kTilePts + mt * 8 + g;
    if (pt < npts)
    {
#pragma unroll
      for (int d = 0; d < DN; ++d)
      {
        double* row = basis + (size_t(DLO + d) * npts + pt) * N;
#pragma unroll
        for (int j = 0; j < TN; ++j)
        {
          const int n = (TLO + j) * 8 + 2 * t;
          if ((N & 1) == 0)
          {
            if (n < N)
              __stcs(reinterpret_cast<double2*>(row + n), make_double2(acc[d][j][0], acc[d][j][1]));
          }
          else
          {
            if (n < N)
              __stcs(row + n, acc[d][j][0]);
            if (n + 1 < N)
              __stcs(row + n + 1, acc[d][j][1]);
          }
        }
      }
    }
  }
}

// x[npts][TDIM]; Cs_g: [NPAD][KS] prepared on the host; basis[NDER][npts][N]
template <int TDIM, int DEG, int ND>
__global__ void __launch_bounds__(kThreads, 1)
    tabulate_fused_kernel(const double* __restrict__ x, size_t npts, const double* __restrict__ Cs_g,
                          double* __restrict__ basis, int N, int nstages)
{
  using C = Cfg<TDIM, DEG, ND>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [ barriers: full[S], empty[S] (256 B) ][ Cs ][ ring: S chunks ]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  double* Cs = reinterpret_cast<double*>(smem_raw + 256);
  double* ring = Cs + C::CS_DOUBLES;
  const uint32_t full0 = smem_u32(bars);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < C::CS_DOUBLES; i += kThreads)
    Cs[i] = Cs_g[i];
  if (threadIdx.x == 0)
  {
    for (int s = 0; s < nstages; ++s)
    {
      mbar_init(full0 + 8 * s, 32 * kProducers);        // every producer lane arrives
      mbar_init(full0 + 8 * (nstages + s), kConsumers); // one elected lane per consumer warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const size_t ntiles = ceil_div(npts, size_t(kTilePts));
  if (warp >= kConsumers)
  {
    switch (warp - kConsumers)
    {
    case 0: producer_loop<TDIM, DEG, ND, 0>(x, npts, ntiles, ring, full0, nstages, lane); break;
    case 1: producer_loop<TDIM, DEG, ND, 1>(x, npts, ntiles, ring, full0, nstages, lane); break;
    case 2: producer_loop<TDIM, DEG, ND, 2>(x, npts, ntiles, ring, full0, nstages, lane); break;
    default: producer_loop<TDIM, DEG, ND, 3>(x, npts, ntiles, ring, full0, nstages, lane); break;
    }
  }
  else
  {
    const int mt = warp & 3;
    if ((warp >> 2) == 0)
      consumer_loop<TDIM, DEG, ND, 0>(ring, Cs, full0, nstages, mt, lane, basis, npts, N, ntiles);
    else
      consumer_loop<TDIM, DEG, ND, 1>(ring, Cs, full0, nstages, mt, lane, basis, npts, N, ntiles);
  }
}

// Host: launch one instantiation.
template <int TDIM, int DEG, int ND>
void launch(const bx_element& e, const double* x, size_t npts, double* basis, cudaStream_t s)
{
  using C = Cfg<TDIM, DEG, ND>;
  constexpr int budget = 200 * 1024;
  constexpr int S = C::stages(budget);
  static_assert(S >= 2, "ring does not fit in shared memory");
  constexpr size_t smem = 256 + size_t(C::CS_DOUBLES) * 8 + size_t(S) * C::CHUNK_DOUBLES * 8;
  auto kern = tabulate_fused_kernel<TDIM, DEG, ND>;
  static thread_local int configured_dev = -1;
  int dev = 0;
  BX_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev)
  {
    BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    configured_dev = dev;
  }
  const size_t ntiles = ceil_div(npts, size_t(kTilePts));
  const int grid = int(std::min<size_t>(ntiles, size_t(sm_count())));
  kern<<<grid, kThreads, smem, s>>>(x, npts, e.d_Cs, basis, e.N, S);
  BX_LAUNCHED();
}
} // namespace fused
} // namespace bx

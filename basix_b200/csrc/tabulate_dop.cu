// Derivative-operator tabulate for simplex elements: every derivative slab from the VALUES of the orthonormal set.
//
// Contract (reference cpp/basix/finite-element.cpp:1834-1888):
//   basis[d][pt][n] = sum_k C[n][k] * d^alpha P_k(x_pt),   alpha = derivative multi-index of slab d.
// The reference tabulates all derivatives of all K basis functions (polyset.cpp:1601-2020) and contracts each of
// the nder slabs with the full K-column coefficient matrix.  On a simplex the orthonormal set is hierarchical by
// total degree (indexing.h:13-34 orders by p+q(+r)), and a derivative of order m of a polynomial of degree n lies
// in P_{n-m}; so with D_a[k][j] = int d_a P_k P_j (first-derivative operators in the orthonormal basis, exact
// Gauss-Jacobi quadrature on the host)
//   d^alpha P_k = sum_{j < dim P_{n-m}} (prod_a D_a^{alpha_a})[k][j] P_j
//   basis[d][pt][n] = sum_{j < dim P_{n-m}} B_d[n][j] P_j(x_pt),     B_d = C * prod_a D_a^{alpha_a}   (host, once).
// Consequences on the device:
//   * only VALUES of the orthonormal set are evaluated (no derivative jets: 10x less recurrence work for nd = 2),
//     once per point, into shared memory;
//   * slab d contracts over dim P_{n-m} columns instead of K: for P5 on the tetrahedron with nd = 2 the k-extent
//     summed over the 10 slabs is 56 + 3*35 + 6*20 = 281 instead of 560 -- half the FP64 tensor-core work;
//   * all B_d stay resident in shared memory for the whole kernel; vector-valued elements (N > K) fit too.
// Kernel: persistent, one CTA per SM, twelve INDEPENDENT warps (three per SM sub-partition; eight when the operands
// leave less room).  Each warp walks its own 16-point chunks: lanes 0..15 evaluate the recurrences for one point each
// (compile-time (p, q, r): literal coefficients, immediate store offsets) into the warp's private, XOR-swizzled
// shared-memory slab P[k][16]; then the warp contracts its two m8 row tiles with every slab operand on the FP64 tensor
// cores (each B-fragment load feeds two DMMAs) and stores each m8n8 accumulator fragment straight as
// basis[d][pt][n].  The contraction is ONE flat loop over the k-steps of all slabs, unrolled by two over two fragment
// register sets: the fragments of step s+1 are loaded while the DMMAs of step s issue (the compiler spreads the loads
// between the DMMAs, where they cost nothing), so a slab boundary costs only its stores, not a pipeline drain (a
// separate loop per slab reached 65 % of the DMMA peak for one warp per sub-partition and 80 % for two,
// tools/dmma_loop.cu).  The slab operands are stored fragment-major in the order the loop consumes them (one pointer,
// immediate offsets, no per-step address arithmetic) and enter shared memory as TMA bulk copies on one mbarrier.  No
// block barrier and no mbarrier after set-up:
// DFMA and DMMA share one execution pipe on B200 (tools/mixed_fp64.cu: 37 TFLOP/s alone or mixed), so dedicated
// producer warps starve behind the DMMA stream and stall every consumer waiting on their tile (measured: 2 ms of
// 12.5 ms per 10 M points); with independent warps a warp that is producing or storing simply leaves the pipe to
// the other warps of its sub-partition.  Measured on P5 / nd = 2, 10 M points: 9.9-10.0 ms (0.855 of the DMMA peak on
// the flops this algorithm executes), against 20.9 ms for the dense fused kernel; with the value production removed
// 9.42, with the stores removed 9.50, with both removed 9.15 (the loop itself: 0.94 of the executed-flop floor).
// Parity: 2.8e-14 * max|ref| per slab against the reference on P5 / nd = 2 (bar: 1e-12).
#include "element.cuh"
#include "hostmath.cuh"
#include "jets.cuh"

#include <algorithm>
#include <type_traits>
#include <cstdlib>
#include <map>
#include <mutex>

namespace bx
{
namespace dop
{
constexpr int kChunkPts = 16; // points per warp chunk: two m8 row tiles
constexpr int kMaxNT = 7;     // 8-column output tiles per column group

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// One k-step of the flat contraction loop = one packed table entry: rows 4 ks .. 4 ks + 3 of the point slab against
// the operand of one derivative slab.  The operands are stored FRAGMENT-MAJOR, in the order the loop consumes them:
//   Bf[column group][step][j < NT][lane] = B_slab[n0 + 8 j + (lane >> 2)][4 ks + (lane & 3)]
// so every B-fragment load of a step is one conflict-free LDS at an immediate offset from ONE pointer that advances
// by NT * 32 doubles per step: no per-step address arithmetic, no padding of the operand rows.
using Step = int;
constexpr int kStepKsMask = 0xff;   // bits 0..7: ks (slab row 4 ks of the point slab = offset 64 ks doubles)
constexpr int kStepLastShift = 8;   // bits 8..: 0, or 1 + derivative slab index when this is the last k-step of it

// T = type of the points and of the table (double, or float for the float32 API: loads widen, stores narrow, the
// operands and the arithmetic are double either way)
template <typename T>
struct Params
{
  const T* x;
  size_t npts;
  const double* B;   // all slab operands, fragment-major (+ one step of zero padding: prefetch target)
  const Step* steps; // nsteps + 2 entries (the last two repeat the last step: prefetch targets)
  T* basis;
  size_t slab_stride; // npts * N: elements between derivative slabs of the table
  int N, K, kpad;     // output columns, basis functions, rows of a point slab (multiple of 4)
  int b_doubles, nsteps, steps_pad; // steps_pad: table entries rounded up to a multiple of 4 (16-byte granules)
};

// Point slab: element (k, c), c = point 0..15, lives at 16 k + (c ^ 4 (k & 3)); both the producer's row writes
// (16 lanes, fixed k) and the consumers' A-fragment reads (k = 4 ks + t, c = g or g + 8) are conflict free.
struct SlabSink
{
  double* base[4]; // slab + (c ^ 4 m), m = k & 3: row k is one store at a compile-time offset
  template <int SLOT>
  __device__ __forceinline__ void put(double v) const
  {
    base[SLOT & 3][16 * SLOT] = v;
  }
};

// ---- values of the orthonormal set, un-normalised (norms are folded into the operands) --------------------
// The same three-term recurrences as jets.cuh with ND = 0 (reference polyset.cpp:66-120, 1601-1743, 1745-2020),
// written with compile-time (p, q, r) so every recurrence coefficient is a literal and every value goes to its
// natural basis slot (indexing.h:13-34) at an immediate shared-memory offset.
template <int I, int E, typename F>
__device__ __forceinline__ void static_for(F&& f)
{
  if constexpr (I < E)
  {
    f(std::integral_constant<int, I>{});
    static_for<I + 1, E>(f);
  }
}

template <int N>
__device__ __forceinline__ void values_interval(double x0, const SlabSink& out)
{
  const double X = x0 * 2.0 - 1.0;
  double J1 = 1.0, J2 = 0.0;
  out.put<0>(J1);
  static_for<1, N + 1>(
      [&](auto P)
      {
        constexpr int p = decltype(P)::value;
        constexpr double a = 1.0 - 1.0 / double(p);
        double v = (X * (a + 1.0)) * J1;
        if constexpr (p > 1)
          v -= a * J2;
        J2 = J1;
        J1 = v;
        out.put<p>(J1);
      });
}

template <int N>
__device__ __forceinline__ void values_triangle(double x0, double x1, const SlabSink& out)
{
  const double X = x0 * 2.0 - 1.0, Y = x1 * 2.0 - 1.0;
  const double lin = X + 0.5 * Y + 0.5;
  const double f3 = 1.0 - x1;
  const double f3sq = f3 * f3;
  double Jp = 1.0, Jpm = 0.0;
  static_for<0, N + 1>(
      [&](auto P)
      {
        constexpr int p = decltype(P)::value;
        if constexpr (p > 0)
        {
          constexpr double a = double(2 * p - 1) / double(p);
          double v = (lin * a) * Jp;
          if constexpr (p > 1)
            v -= (f3sq * (a - 1.0)) * Jpm;
          Jpm = Jp;
          Jp = v;
        }
        out.put<idx2(0, p)>(Jp);
        double Jq = Jp, Jqm = 0.0;
        static_for<1, N - p + 1>(
            [&](auto Q)
            {
              constexpr int q = decltype(Q)::value;
              double v;
              if constexpr (q == 1)
                v = (Y * (1.5 + p) + (0.5 + p)) * Jq;
              else
              {
                constexpr double a1 = jets::jrc_a(2 * p + 1, q - 1), a2 = jets::jrc_b(2 * p + 1, q - 1),
                                 a3 = jets::jrc_c(2 * p + 1, q - 1);
                v = (Y * a1 + a2) * Jq - a3 * Jqm;
              }
              Jqm = Jq;
              Jq = v;
              out.put<idx2(q, p)>(Jq);
            });
      });
}

template <int N>
__device__ __forceinline__ void values_tetrahedron(double x0, double x1, double x2, const SlabSink& out)
{
  const double X = x0 * 2.0 - 1.0, Y = x1 * 2.0 - 1.0, Z = x2 * 2.0 - 1.0;
  const double lin = X + 0.5 * (Y + Z) + 1.0;
  const double f2 = x1 + x2 - 1.0;
  const double f2sq = f2 * f2;
  const double f4 = 1.0 - x2;
  const double f4sq = f4 * f4;
  const double f3 = Y + x2;
  double Jp = 1.0, Jpm = 0.0;
  static_for<0, N + 1>(
      [&](auto P)
      {
        constexpr int p = decltype(P)::value;
        if constexpr (p > 0)
        {
          constexpr double a = double(2 * p - 1) / double(p);
          double v = (lin * a) * Jp;
          if constexpr (p > 1)
            v -= (f2sq * (a - 1.0)) * Jpm;
          Jpm = Jp;
          Jp = v;
        }
        double Jq = Jp, Jqm = 0.0;
        static_for<0, N - p + 1>(
            [&](auto Q)
            {
              constexpr int q = decltype(Q)::value;
              if constexpr (q == 1)
              {
                const double v = ((1.0 + Y) * p + (2.0 + Y * 3.0 + Z) * 0.5) * Jq;
                Jqm = Jq;
                Jq = v;
              }
              else if constexpr (q > 1)
              {
                constexpr double aq = jets::jrc_a(2 * p + 1, q - 1), bq = jets::jrc_b(2 * p + 1, q - 1),
                                 cq = jets::jrc_c(2 * p + 1, q - 1);
                const double v = (f3 * aq + f4 * bq) * Jq - (f4sq * cq) * Jqm;
                Jqm = Jq;
                Jq = v;
              }
              out.put<idx3(0, q, p)>(Jq);
              double Jr = Jq, Jrm = 0.0;
              static_for<1, N - p - q + 1>(
                  [&](auto R)
                  {
                    constexpr int r = decltype(R)::value;
                    double v;
                    if constexpr (r == 1)
                      v = ((1.0 + p + q) + Z * (2.0 + p + q)) * Jr;
                    else
                    {
                      constexpr double ar = jets::jrc_a(2 * p + 2 * q + 2, r - 1), br = jets::jrc_b(2 * p + 2 * q + 2, r - 1),
                                       cr = jets::jrc_c(2 * p + 2 * q + 2, r - 1);
                      v = (Z * ar + br) * Jr - cr * Jrm;
                    }
                    Jrm = Jr;
                    Jr = v;
                    out.put<idx3(r, q, p)>(Jr);
                  });
            });
      });
}

// Everything run_group needs to place an accumulator fragment in the table.
template <typename T>
struct Out
{
  T* row0;            // LEAN: basis + pt * N + n0 + 2 t (row g of the first row tile, this lane's column pair)
  T* basis;           // general form
  size_t slab_stride; // npts * N
  size_t npts, pt;
  int N, n0, t;
};

// One k-step on this warp's two row tiles: 2 NT DMMAs, then -- on the last step of a slab -- the stores.
template <int NT, bool LEAN, typename T>
__device__ __forceinline__ void step_compute(int st, double (&acc)[2][NT][2], double a0, double a1,
                                             const double (&b)[NT], const Out<T>& o)
{
#pragma unroll
  for (int j = 0; j < NT; ++j)
  {
    dmma(acc[0][j][0], acc[0][j][1], a0, b[j]);
    dmma(acc[1][j][0], acc[1][j][1], a1, b[j]);
  }
  const int last = st >> kStepLastShift;
  if (last)
  {
    // accumulator fragment: row g = point, columns 2t, 2t+1 of each 8-column tile
    const size_t slab = size_t(last - 1);
    if constexpr (LEAN)
    {
      // whole chunk inside the table, N even: 16-byte stores at compile-time offsets from one row pointer;
      // only the last column tile of the operand can run past N
      T* row = o.row0 + slab * o.slab_stride;
      using T2 = typename std::conditional<sizeof(T) == 8, double2, float2>::type;
      auto pair = [](double x, double y)
      {
        T2 v;
        v.x = static_cast<T>(x);
        v.y = static_cast<T>(y);
        return v;
      };
      const bool last_ok = o.n0 + (NT - 1) * 8 + 2 * o.t < o.N;
#pragma unroll
      for (int m = 0; m < 2; ++m)
      {
#pragma unroll
        for (int j = 0; j < NT - 1; ++j)
          __stcs(reinterpret_cast<T2*>(row + j * 8), pair(acc[m][j][0], acc[m][j][1]));
        if (last_ok)
          __stcs(reinterpret_cast<T2*>(row + (NT - 1) * 8), pair(acc[m][NT - 1][0], acc[m][NT - 1][1]));
        row += size_t(8) * o.N;
      }
    }
    else
    {
#pragma unroll
      for (int m = 0; m < 2; ++m)
      {
        const size_t q = o.pt + 8 * m;
        if (q < o.npts)
        {
          T* row = o.basis + slab * o.slab_stride + q * o.N;
#pragma unroll
          for (int j = 0; j < NT; ++j)
          {
            const int n = o.n0 + j * 8 + 2 * o.t;
            if (n < o.N)
              __stcs(row + n, static_cast<T>(acc[m][j][0]));
            if (n + 1 < o.N)
              __stcs(row + n + 1, static_cast<T>(acc[m][j][1]));
          }
        }
      }
    }
    // the next step opens the next slab (a branch-free DMMA block lets the compiler spread the fragment loads of the
    // next step between the DMMAs; a zero-C first step in its own branch measured the same)
#pragma unroll
    for (int m = 0; m < 2; ++m)
#pragma unroll
      for (int j = 0; j < NT; ++j)
        acc[m][j][0] = acc[m][j][1] = 0.0;
  }
}

// All slabs for one column group of this warp's 16 points.  Two fragment register sets used alternately (the loop is
// unrolled by two, so nothing is moved between them): the fragments of step s + 1 are in flight during the DMMAs of
// step s, across slab boundaries too.
template <int NT, bool LEAN, typename T>
__device__ __forceinline__ void run_group(const Step* __restrict__ steps, int nsteps, const double* __restrict__ A0,
                                          const double* __restrict__ A1, const double* __restrict__ bp,
                                          const Out<T>& o)
{
  double acc[2][NT][2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < NT; ++j)
      acc[m][j][0] = acc[m][j][1] = 0.0;
  double a[2][2], b[2][NT];
  auto fetch = [&](int set, int st, const double* p)
  {
    const int aoff = (st & kStepKsMask) * 64;
    a[set][0] = A0[aoff];
    a[set][1] = A1[aoff];
#pragma unroll
    for (int j = 0; j < NT; ++j)
      b[set][j] = p[j * 32];
  };
  // table entries are read two steps ahead of their use, fragments one step ahead
  int st0 = steps[0], st1 = steps[1];
  fetch(0, st0, bp);
  int s = 0;
#pragma unroll 1
  for (; s + 1 < nsteps; s += 2)
  {
    const int st2 = steps[s + 2];
    fetch(1, st1, bp + NT * 32);
    step_compute<NT, LEAN, T>(st0, acc, a[0][0], a[0][1], b[0], o);
    bp += 2 * NT * 32;
    const int st3 = steps[s + 3];
    fetch(0, st2, bp);
    step_compute<NT, LEAN, T>(st1, acc, a[1][0], a[1][1], b[1], o);
    st0 = st2;
    st1 = st3;
  }
  if (s < nsteps)
    step_compute<NT, LEAN, T>(st0, acc, a[0][0], a[0][1], b[0], o);
}

template <bool LEAN, typename T>
__device__ __forceinline__ void run_groups(const Step* steps, int nsteps, const double* A0, const double* A1,
                                           const double* Bl, int g, int t, T* basis, size_t npts, size_t slab_stride,
                                           int N, size_t pt)
{
  Out<T> o;
  o.basis = basis;
  o.slab_stride = slab_stride;
  o.npts = npts;
  o.pt = pt;
  o.N = N;
  o.t = t;
  for (int n0 = 0; n0 < N; n0 += 8 * kMaxNT)
  {
    const int nt = min(kMaxNT, (N - n0 + 7) >> 3);
    o.n0 = n0;
    o.row0 = basis + pt * N + n0 + 2 * t;
    switch (nt)
    {
    case 1: run_group<1, LEAN, T>(steps, nsteps, A0, A1, Bl, o); break;
    case 2: run_group<2, LEAN, T>(steps, nsteps, A0, A1, Bl, o); break;
    case 3: run_group<3, LEAN, T>(steps, nsteps, A0, A1, Bl, o); break;
    case 4: run_group<4, LEAN, T>(steps, nsteps, A0, A1, Bl, o); break;
    case 5: run_group<5, LEAN, T>(steps, nsteps, A0, A1, Bl, o); break;
    case 6: run_group<6, LEAN, T>(steps, nsteps, A0, A1, Bl, o); break;
    default: run_group<7, LEAN, T>(steps, nsteps, A0, A1, Bl, o); break;
    }
    Bl += size_t(nsteps) * nt * 32; // fragment block of the next column group
  }
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }

template <int TDIM, int DEG, int WARPS, typename T>
__global__ void __launch_bounds__(32 * WARPS, 1) tabulate_dop_kernel(const Params<T> p)
{
  constexpr int kThreads = 32 * WARPS;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // [ mbarrier ][ steps ][ B ][ slab of warp 0 ] ... [ slab of warp WARPS-1 ]
  unsigned long long* bar = reinterpret_cast<unsigned long long*>(smem_raw);
  Step* steps = reinterpret_cast<Step*>(smem_raw + 16);
  double* Bs = reinterpret_cast<double*>(smem_raw + 16 + size_t(p.steps_pad) * sizeof(Step));
  const int slab_doubles = p.kpad * 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  double* slab = Bs + p.b_doubles + warp * slab_doubles;

  // the operands (one contiguous block, up to ~190 KB) and the step table enter shared memory as TMA bulk copies
  // tracked by one mbarrier; meanwhile all threads clear the point slabs
  if (threadIdx.x == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    const uint32_t b_bytes = uint32_t(p.b_doubles) * 8u, s_bytes = uint32_t(p.steps_pad) * uint32_t(sizeof(Step));
    asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(smem_addr(bar)),
                 "r"(b_bytes + s_bytes)
                 : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_addr(steps)),
                 "l"(p.steps), "r"(s_bytes), "r"(smem_addr(bar))
                 : "memory");
    constexpr uint32_t kPiece = 32768;
    for (uint32_t off = 0; off < b_bytes; off += kPiece)
    {
      const uint32_t n = min(kPiece, b_bytes - off);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                       smem_addr(Bs) + off),
                   "l"(reinterpret_cast<const char*>(p.B) + off), "r"(n), "r"(smem_addr(bar))
                   : "memory");
    }
  }
  for (int i = threadIdx.x; i < WARPS * slab_doubles; i += kThreads)
    Bs[p.b_doubles + i] = 0.0; // rows K .. kpad-1 are read by the last k-step of the value slab and must stay finite
  __syncthreads(); // barrier initialised and visible, slabs cleared
  {
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n .reg .pred q;\n mbarrier.try_wait.parity.shared::cta.b64 q, [%1], %2;\n selp.u32 %0, 1, 0, q;\n}"
                   : "=r"(done)
                   : "r"(smem_addr(bar)), "r"(0)
                   : "memory");
  }

  // chunk c of this warp: points [16 c, 16 c + 16); consecutive warps / CTAs take consecutive chunks
  const size_t nchunks = ceil_div(p.npts, size_t(kChunkPts));
  const size_t stride = size_t(gridDim.x) * WARPS;
  const int g = lane >> 2, t = lane & 3;
  const bool n_even = (p.N & 1) == 0;
  const double* A0 = slab + t * 16 + (g ^ (t << 2));
  const double* A1 = slab + t * 16 + ((g + 8) ^ (t << 2));
  const double* Bl = Bs + lane;
  const int c = lane & 15;
  const SlabSink sink{{slab + c, slab + (c ^ 4), slab + (c ^ 8), slab + (c ^ 12)}};

  auto load_x = [&](size_t chunk, double (&xr)[TDIM])
  {
    size_t pt = chunk * kChunkPts + (lane & 15);
    pt = pt < p.npts ? pt : p.npts - 1; // ragged tail: recompute the last point, stores are predicated
#pragma unroll
    for (int a = 0; a < TDIM; ++a)
      xr[a] = double(__ldg(p.x + TDIM * pt + a));
  };

  size_t chunk = size_t(blockIdx.x) * WARPS + warp;
  double xr[TDIM];
  if (chunk < nchunks)
    load_x(chunk, xr);
  for (; chunk < nchunks; chunk += stride)
  {
    // ---- values of the orthonormal set: lane = point (lanes 16..31 idle) ----
    if (lane < kChunkPts)
    {
      if constexpr (TDIM == 1)
        values_interval<DEG>(xr[0], sink);
      else if constexpr (TDIM == 2)
        values_triangle<DEG>(xr[0], xr[1], sink);
      else
        values_tetrahedron<DEG>(xr[0], xr[1], xr[2], sink);
    }
    __syncwarp();
    if (chunk + stride < nchunks)
      load_x(chunk + stride, xr); // in flight during the contraction

    // ---- FP64 tensor cores: every slab on this warp's two row tiles ----
    const size_t pt = chunk * kChunkPts + g;
    const bool lean = n_even and (chunk + 1) * kChunkPts <= p.npts;
    if (lean)
      run_groups<true, T>(steps, p.nsteps, A0, A1, Bl, g, t, p.basis, p.npts, p.slab_stride, p.N, pt);
    else
      run_groups<false, T>(steps, p.nsteps, A0, A1, Bl, g, t, p.basis, p.npts, p.slab_stride, p.N, pt);
    __syncwarp(); // the slab is overwritten by the next chunk
  }
}

// ---- host: operand per (element, nd), built once and cached on the element --------------------------------
struct Operand
{
  double* d_B = nullptr;
  Step* d_steps = nullptr;
  int b_doubles = 0, nsteps = 0, steps_pad = 0, kpad = 0, warps = 0;
  bool usable = false;
  size_t smem = 0;
};
} // namespace dop

struct bx_element_dop
{
  std::mutex mutex;
  std::map<int, dop::Operand> by_nd;
  ~bx_element_dop()
  {
    for (auto& kv : by_nd)
    {
      cudaFree(kv.second.d_B);
      cudaFree(kv.second.d_steps);
    }
  }
};
bx_element_dop* dop_cache_create() { return new bx_element_dop; }
void dop_cache_destroy(bx_element_dop* c) { delete c; }

namespace dop
{
namespace
{
using Vec = std::vector<double>;

int simplex_dim(int tdim, int n) { return n < 0 ? 0 : jets::psize(tdim, n); }

template <typename T>
T* upload(const std::vector<T>& v)
{
  T* d = nullptr;
  BX_CUDA(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty())
    BX_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

Operand build_operand(const bx_element& e, int nd, bool upload_to_device)
{
  Operand op;
  const int tdim = e.tdim, n = e.superdegree, K = e.K, N = e.N;
  const int cell = e.cell_type;
  // first-derivative operators in the orthonormal basis: D_a[k][j] = int d_a P_k P_j, j < dim P_{n-1}
  std::vector<Vec> D(tdim, Vec(size_t(K) * K, 0.0));
  if (nd > 0 and n > 0)
  {
    Vec qx, qw;
    host::quadrature(cell, 2 * n, qx, qw);
    const size_t nq = qw.size();
    const Vec P = host::polyset_tabulate(cell, n, 1, qx, nq); // [1 + tdim][K][nq]
    const int Km1 = simplex_dim(tdim, n - 1);
    for (int a = 0; a < tdim; ++a)
    {
      // slot of the first derivative along axis a: idx(1,0[,0]), idx(0,1[,0]), idx(0,0,1)
      const int slot = tdim == 1 ? 1 : tdim == 2 ? (a == 0 ? idx2(1, 0) : idx2(0, 1))
                                                 : (a == 0 ? idx3(1, 0, 0) : a == 1 ? idx3(0, 1, 0) : idx3(0, 0, 1));
      const double* dP = P.data() + size_t(slot) * K * nq;
      for (int k = 0; k < K; ++k)
        for (int j = 0; j < Km1; ++j)
        {
          double s = 0.0;
          for (size_t q = 0; q < nq; ++q)
            s += qw[q] * dP[size_t(k) * nq + q] * P[size_t(j) * nq + q];
          D[a][size_t(k) * K + j] = s;
        }
    }
  }
  // natural-index norms and the production-order row map of the jets
  std::vector<int> slot(K);
  Vec norm_prod(K), norm(K);
  jets::production_order(tdim, n, slot.data(), norm_prod.data());
  for (int s = 0; s < K; ++s)
    norm[slot[s]] = norm_prod[s];

  const int nder = polyset_nderivs(cell, nd);
  const int npad = int(ceil_div(N, 8) * 8);
  op.kpad = int(ceil_div(K, 4) * 4);
  // dense, zero-padded operand of every slab in step order: Bd[npad][4 ksteps]
  struct SlabOp
  {
    Vec Bd;
    int ksteps, d;
  };
  std::vector<SlabOp> slabs;
  std::vector<Step> steps;
  auto matmul = [K](const Vec& A, const Vec& Bm)
  {
    Vec Cm(size_t(K) * K, 0.0);
    for (int i = 0; i < K; ++i)
      for (int k = 0; k < K; ++k)
      {
        const double a = A[size_t(i) * K + k];
        if (a == 0.0)
          continue;
        for (int j = 0; j < K; ++j)
          Cm[size_t(i) * K + j] += a * Bm[size_t(k) * K + j];
      }
    return Cm;
  };
  for (int kx = 0; kx <= nd; ++kx)
    for (int ky = 0; ky <= (tdim >= 2 ? nd - kx : 0); ++ky)
      for (int kz = 0; kz <= (tdim >= 3 ? nd - kx - ky : 0); ++kz)
      {
        const int m = kx + ky + kz;
        const int d = tdim == 1 ? kx : tdim == 2 ? idx2(kx, ky) : idx3(kx, ky, kz);
        if (d >= nder)
          continue;
        Vec M(size_t(K) * K, 0.0);
        for (int i = 0; i < K; ++i)
          M[size_t(i) * K + i] = 1.0;
        const int cnt[3] = {kx, ky, kz};
        for (int a = 0; a < tdim; ++a)
          for (int r = 0; r < cnt[a]; ++r)
            M = matmul(M, D[a]);
        const int Kd = simplex_dim(tdim, n - m);
        SlabOp so;
        so.ksteps = std::max(1, int(ceil_div(Kd, 4))); // derivative order above the degree: one zero step
        so.d = d;
        const int kp = so.ksteps * 4;
        so.Bd.assign(size_t(npad) * kp, 0.0);
        for (int row = 0; row < N; ++row)
          for (int j = 0; j < Kd; ++j)
          {
            double sum = 0.0;
            for (int k = 0; k < K; ++k)
              sum += e.Cp[size_t(row) * e.Kpad + k] * M[size_t(k) * K + j];
            so.Bd[size_t(row) * kp + j] = sum * norm[j];
          }
        for (int ks = 0; ks < so.ksteps; ++ks)
          steps.push_back(ks | (ks + 1 == so.ksteps ? (1 + d) << kStepLastShift : 0));
        slabs.push_back(std::move(so));
      }
  if (nder > 255 or op.kpad / 4 > 256)
    return op; // outside the packed step encoding
  op.nsteps = int(steps.size());
  steps.push_back(steps.back()); // prefetch targets of the last steps
  steps.push_back(steps.back());
  op.steps_pad = int(ceil_div(steps.size(), 4) * 4);
  steps.resize(op.steps_pad, steps.back());
  // fragment-major: Bf[column group][step][j < NT][lane] = B_slab[n0 + 8 j + (lane >> 2)][4 ks + (lane & 3)]
  Vec B;
  B.reserve(size_t(op.nsteps + 1) * (npad / 8) * 32);
  for (int n0 = 0; n0 < N; n0 += 8 * kMaxNT)
  {
    const int nt = std::min(kMaxNT, (N - n0 + 7) / 8);
    for (const SlabOp& so : slabs)
      for (int ks = 0; ks < so.ksteps; ++ks)
        for (int j = 0; j < nt; ++j)
          for (int lane = 0; lane < 32; ++lane)
            B.push_back(so.Bd[size_t(n0 + 8 * j + (lane >> 2)) * (4 * so.ksteps) + 4 * ks + (lane & 3)]);
  }
  B.resize(B.size() + size_t(2) * kMaxNT * 32, 0.0); // the loop prefetches up to two steps past the last group
  op.b_doubles = int(B.size());
  // three warps per SM sub-partition when their slabs fit beside the operands, else two
  static const int warps_env = std::getenv("BX_DOP_WARPS") ? std::atoi(std::getenv("BX_DOP_WARPS")) : 0;
  for (int warps : {12, 8})
  {
    if (warps_env and warps != warps_env)
      continue;
    const size_t smem = 16 + sizeof(Step) * steps.size() + size_t(op.b_doubles) * 8
                        + size_t(warps) * op.kpad * 16 * 8;
    if (smem <= 227 * 1024)
    {
      op.smem = smem;
      op.warps = warps;
      break;
    }
  }
  if (op.warps == 0)
    return op;
  if (upload_to_device)
  {
    op.d_B = upload(B);
    op.d_steps = upload(steps);
  }
  op.usable = true;
  return op;
}

template <int TDIM, int DEG, int WARPS, typename T>
void launch_warps(const Operand& op, const bx_element& e, const T* x, size_t npts, T* basis, cudaStream_t s)
{
  auto kern = tabulate_dop_kernel<TDIM, DEG, WARPS, T>;
  BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(op.smem)));
  Params<T> p;
  p.x = x;
  p.npts = npts;
  p.B = op.d_B;
  p.steps = op.d_steps;
  p.basis = basis;
  p.N = e.N;
  p.K = e.K;
  p.kpad = op.kpad;
  p.b_doubles = op.b_doubles;
  p.nsteps = op.nsteps;
  p.steps_pad = op.steps_pad;
  p.slab_stride = npts * size_t(e.N);
  const size_t nblocks = ceil_div(npts, size_t(kChunkPts) * WARPS);
  kern<<<unsigned(std::min<size_t>(nblocks, size_t(sm_count()))), 32 * WARPS, op.smem, s>>>(p);
  BX_LAUNCHED();
}

template <int TDIM, int DEG, typename T>
void launch(const Operand& op, const bx_element& e, const T* x, size_t npts, T* basis, cudaStream_t s)
{
  if (op.warps == 12)
    launch_warps<TDIM, DEG, 12, T>(op, e, x, npts, basis, s);
  else
    launch_warps<TDIM, DEG, 8, T>(op, e, x, npts, basis, s);
}

#define BX_DOP(T, DEG)                                                                                     \
  if (e.tdim == T and e.superdegree == DEG)                                                                \
  {                                                                                                        \
    if (!dry_run)                                                                                          \
      launch<T, DEG, U>(*op, e, x, npts, basis, s);                                                           \
    return true;                                                                                           \
  }
} // namespace

// Returns false when the element / derivative order is outside this path (caller falls through).
template <typename U>
bool launch_any(const bx_element& e, int nd, const U* x, size_t npts, U* basis, cudaStream_t s, bool dry_run)
{
  if (e.polyset_type != BX_POLYSET_STANDARD or nd < 0 or nd > 4 or !e.dop_cache)
    return false;
  if (e.cell_type != BX_CELL_INTERVAL and e.cell_type != BX_CELL_TRIANGLE and e.cell_type != BX_CELL_TETRAHEDRON)
    return false;
  const int max_deg = e.tdim == 1 ? 10 : e.tdim == 2 ? 8 : 6;
  if (e.superdegree < 1 or e.superdegree > max_deg)
    return false;
  const Operand* op = nullptr;
  {
    std::lock_guard<std::mutex> lock(e.dop_cache->mutex);
    auto it = e.dop_cache->by_nd.find(nd);
    if (it == e.dop_cache->by_nd.end())
      it = e.dop_cache->by_nd.emplace(nd, build_operand(e, nd, e.on_device)).first;
    op = &it->second;
  }
  if (!op->usable or (!dry_run and !op->d_B))
    return false;
#ifdef BX_DOP_DEV // development builds: only the BASELINE element (the full menu takes minutes to compile)
  BX_DOP(3, 5)
#else
  BX_DOP(1, 1) BX_DOP(1, 2) BX_DOP(1, 3) BX_DOP(1, 4) BX_DOP(1, 5) BX_DOP(1, 6) BX_DOP(1, 7) BX_DOP(1, 8)
  BX_DOP(1, 9) BX_DOP(1, 10)
  BX_DOP(2, 1) BX_DOP(2, 2) BX_DOP(2, 3) BX_DOP(2, 4) BX_DOP(2, 5) BX_DOP(2, 6) BX_DOP(2, 7) BX_DOP(2, 8)
  BX_DOP(3, 1) BX_DOP(3, 2) BX_DOP(3, 3) BX_DOP(3, 4) BX_DOP(3, 5) BX_DOP(3, 6)
#endif
  return false;
}
template bool launch_any<double>(const bx_element&, int, const double*, size_t, double*, cudaStream_t, bool);
template bool launch_any<float>(const bx_element&, int, const float*, size_t, float*, cudaStream_t, bool);
} // namespace dop
} // namespace bx

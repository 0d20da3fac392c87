// Batched application of the interpolation matrix: coefficients[c][n] = sum_k M[n][k] values[c][k].
//
// The reference exposes interpolation_matrix() (cpp/basix/finite-element.h:1222-1226) and leaves the product
// "coefficients = i_m * values" to the caller, one cell at a time (finite-element.h:1183-1216).  Over many cells
// it is a tall-skinny FP64 GEMM -- [nc x K] by [K x ndofs], K = value_size * npts -- and the step that follows
// pull_back when a function is interpolated into the element.  For N1curl degree 3 on the tetrahedron (K = 141,
// ndofs = 45) that is 12.7 kflop and 1488 B per cell: 8.5 flop/B against a machine balance of 5.7 (37 TFLOP/s
// DMMA, 6.5 TB/s HBM), so the kernel is bound by the FP64 tensor pipe with HBM close behind.
//
// Kernel (the contraction half of tabulate_dop.cu, with the A operand streamed from HBM instead of produced):
// persistent, one CTA per SM, sixteen independent warps (four per SM sub-partition).  The operand M (rows padded to 8, row stride == 4 mod 8
// doubles, column order matching the value layout) stays in shared memory.  A warp owns 16 cells = two m8 row
// tiles; lane (g, t) loads its A-fragment entries values[c0 + g (+8)][4 ks + t] straight from global memory --
// the four t-lanes of a row read 32 consecutive bytes -- four k-steps ahead in registers, B fragments one step ahead
// from shared memory, and every B fragment feeds two DMMAs.  The k loop is unrolled with compile-time register
// roles: a rolled loop rotates its prefetch registers with two dozen moves per step and is issue-bound (measured
// 14.5 ms per 20 M cells against 8.3 ms unrolled).  Accumulator fragments are stored straight as
// coefficients[c][n].  Measured on N1curl degree 3, 20 M cells: 8.3 ms = 90 % of the DMMA peak on the executed
// (padded 48 x 144) flops.
#include "element.cuh"

#include <algorithm>

namespace bx
{
namespace
{
constexpr int kWarps = 16;
constexpr int kThreads = 32 * kWarps;
constexpr int kCells = 16; // cells per warp chunk
constexpr int kMaxNT = 7;  // 8-column tiles per column group

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// T = type of the values and coefficients (double, or float for the float32 API: loads widen, stores narrow, the
// operand and the arithmetic are double either way)
template <typename T>
struct Params
{
  const T* values;      // [nc][K]
  const double* B;      // [npad][kstride]
  T* out;               // [nc][N]
  size_t nc;
  int K, N, npad, kstride, ksteps;
  int n_base, out_stride; // this launch computes columns [n_base, n_base + N) of out[nc][out_stride]
};

// Column group [n0, n0 + 8 NT) for the 16 cells of one chunk.
// kDepth = k-steps of A fragments in flight per lane; kStreaming = evict-first loads of the values
template <int NT, int kDepth, bool kStreaming, typename T>
__device__ __forceinline__ void run_group(const Params<T>& p, const double* __restrict__ Bs, size_t c0, int n0, int g, int t)
{
  double acc[2][NT][2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < NT; ++j)
      acc[m][j][0] = acc[m][j][1] = 0.0;
  // rows of this lane (clamped: a row past the end is computed and never stored)
  const size_t r0 = min(c0 + g, p.nc - 1), r1 = min(c0 + g + 8, p.nc - 1);
  const T* __restrict__ v0 = p.values + r0 * p.K + t;
  const T* __restrict__ v1 = p.values + r1 * p.K + t;
  const double* __restrict__ bp = Bs + (n0 + g) * p.kstride + t;
  const int ts = 8 * p.kstride;
  const int kt = p.K - t; // entries 4 ks + t < K  <=>  4 ks < kt
  auto lda = [&](const T* v, int ks) -> double
  { return 4 * ks < kt ? double(kStreaming ? __ldcs(v + 4 * ks) : __ldg(v + 4 * ks)) : 0.0; };

  // The k loop is unrolled by kDepth with compile-time register roles, so nothing is copied between steps (a
  // rolled loop spends two dozen register moves per step rotating its prefetch queue and is issue-bound, not
  // DMMA-bound): A fragments live in a queue kDepth k-steps deep (global latency is several steps long), B
  // fragments ping-pong one step ahead.  ksteps is a multiple of kDepth (zero columns pad the operand).
  double q0[kDepth], q1[kDepth], b[2][NT];
#pragma unroll
  for (int d = 0; d < kDepth; ++d)
  {
    q0[d] = lda(v0, d);
    q1[d] = lda(v1, d);
  }
#pragma unroll
  for (int j = 0; j < NT; ++j)
    b[0][j] = bp[j * ts];
#pragma unroll 1
  for (int ks = 0; ks < p.ksteps; ks += kDepth)
  {
#pragma unroll
    for (int u = 0; u < kDepth; ++u)
    {
      const int kn = min(ks + u + 1, p.ksteps - 1);
#pragma unroll
      for (int j = 0; j < NT; ++j)
        b[(u + 1) & 1][j] = bp[j * ts + 4 * kn]; // shared: one step ahead
      const double a0 = q0[u], a1 = q1[u];
      q0[u] = lda(v0, ks + u + kDepth); // global: kDepth steps ahead
      q1[u] = lda(v1, ks + u + kDepth);
#pragma unroll
      for (int j = 0; j < NT; ++j)
      {
        dmma(acc[0][j][0], acc[0][j][1], a0, b[u & 1][j]);
        dmma(acc[1][j][0], acc[1][j][1], a1, b[u & 1][j]);
      }
    }
  }
  // accumulator fragment: row g = cell, columns 2t, 2t+1 of each 8-column tile
  const bool even = ((p.out_stride | p.n_base) & 1) == 0;
#pragma unroll
  for (int m = 0; m < 2; ++m)
  {
    const size_t c = c0 + g + 8 * m;
    if (c < p.nc)
    {
      T* row = p.out + c * size_t(p.out_stride) + p.n_base;
#pragma unroll
      for (int j = 0; j < NT; ++j)
      {
        const int n = n0 + j * 8 + 2 * t;
        if constexpr (sizeof(T) == 8)
        {
          if (even)
          {
            if (n < p.N)
              __stcs(reinterpret_cast<double2*>(row + n), make_double2(acc[m][j][0], acc[m][j][1]));
            continue;
          }
        }
        if (n < p.N)
          __stcs(row + n, static_cast<T>(acc[m][j][0]));
        if (n + 1 < p.N)
          __stcs(row + n + 1, static_cast<T>(acc[m][j][1]));
      }
    }
  }
}

template <int kDepth, bool kStreaming, typename T>
__global__ void __launch_bounds__(kThreads, 1) interpolate_kernel(const Params<T> p)
{
  extern __shared__ __align__(16) double interp_smem[];
  double* Bs = interp_smem;
  for (int i = threadIdx.x; i < p.npad * p.kstride; i += kThreads)
    Bs[i] = p.B[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const size_t nchunks = ceil_div(p.nc, size_t(kCells));
  const size_t stride = size_t(gridDim.x) * kWarps;
  for (size_t chunk = size_t(blockIdx.x) * kWarps + warp; chunk < nchunks; chunk += stride)
  {
    const size_t c0 = chunk * kCells;
    for (int n0 = 0; n0 < p.N; n0 += 8 * kMaxNT)
    {
      const int nt = min(kMaxNT, (p.N - n0 + 7) >> 3);
      switch (nt)
      {
      case 1: run_group<1, kDepth, kStreaming, T>(p, Bs, c0, n0, g, t); break;
      case 2: run_group<2, kDepth, kStreaming, T>(p, Bs, c0, n0, g, t); break;
      case 3: run_group<3, kDepth, kStreaming, T>(p, Bs, c0, n0, g, t); break;
      case 4: run_group<4, kDepth, kStreaming, T>(p, Bs, c0, n0, g, t); break;
      case 5: run_group<5, kDepth, kStreaming, T>(p, Bs, c0, n0, g, t); break;
      case 6: run_group<6, kDepth, kStreaming, T>(p, Bs, c0, n0, g, t); break;
      default: run_group<7, kDepth, kStreaming, T>(p, Bs, c0, n0, g, t); break;
      }
    }
  }
}
} // namespace

// values[nc][vs * npts] (layout 0: component major, 1: point major) -> coefficients[nc][ndofs]
template <typename T>
void interpolate_device(const bx_element& e, int layout, const T* values, size_t nc, T* coefficients, cudaStream_t s)
{
  if (e.interp_npts == 0)
    fail(BX_ERR_RUNTIME, "interpolate: the element was built without points() / interpolation_matrix()");
  if (layout != 0 and layout != 1)
    fail(BX_ERR_INVALID_ARG, "interpolate: layout must be BX_INTERP_COMPONENT_MAJOR or BX_INTERP_POINT_MAJOR");
  if (nc == 0)
    return;
  require_device(e);
  Params<T> p;
  p.values = values;
  p.out = coefficients;
  p.nc = nc;
  p.K = e.vs * e.interp_npts;
  p.kstride = e.interp_kstride;
  p.out_stride = e.ndofs;
  p.ksteps = int(ceil_div(size_t(p.K), size_t(4))); // rounded up to the unroll factor below
  // A fragments four k-steps ahead, default caching (measured best of depth 2 / 4 / 6 and streaming / default loads)
  constexpr int kDepth = 4;
  p.ksteps = int(ceil_div(size_t(p.ksteps), size_t(kDepth)) * kDepth);
  if (4 * p.ksteps > p.kstride)
    fail(BX_ERR_RUNTIME, "interpolate: operand padding too small for the unroll factor");
  // The operand [npad][kstride] stays resident in shared memory.  A matrix that does not fit (a degree-5 hexahedron
  // element interpolated into itself is 216 x 216) is applied in blocks of rows: each block of the operand is a
  // contiguous sub-array and fills its own columns of the result.
  const size_t row_bytes = size_t(p.kstride) * sizeof(double);
  const int rows_fit = int(std::min<size_t>(size_t(e.interp_npad), (size_t(200) * 1024 / row_bytes) / 8 * 8));
  if (rows_fit < 8)
    fail(BX_ERR_UNSUPPORTED, "interpolate: a block of eight rows of the interpolation matrix does not fit in shared memory");
  const size_t nblocks = ceil_div(nc, size_t(kCells) * kWarps);
  const unsigned grid = unsigned(std::min<size_t>(nblocks, size_t(sm_count())));
  auto kern = interpolate_kernel<kDepth, false, T>;
  for (int n0 = 0; n0 < e.ndofs; n0 += rows_fit)
  {
    p.n_base = n0;
    p.N = std::min(rows_fit, e.ndofs - n0);
    p.npad = std::min(rows_fit, e.interp_npad - n0);
    p.B = e.d_interp[layout] + size_t(n0) * p.kstride;
    const size_t smem = size_t(p.npad) * row_bytes;
    BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kern<<<grid, kThreads, smem, s>>>(p);
    BX_LAUNCHED();
  }
}
template void interpolate_device<double>(const bx_element&, int, const double*, size_t, double*, cudaStream_t);
template void interpolate_device<float>(const bx_element&, int, const float*, size_t, float*, cudaStream_t);
// ---- pull_back fused into the interpolation -------------------------------------------------------------------------
// coefficients[c] = interpolation_matrix * pull_back(u[c], J_c, detJ_c, K_c): the two steps a caller takes to interpolate
// a physical function into a Piola-mapped element (FiniteElement::pull_back, finite-element.cpp:2007-2042, then the
// product with interpolation_matrix(), finite-element.h:1183-1226), without the pulled-back values ever touching memory:
// u[nc][npts][D] (physical values at the mapped interpolation points, the layout push_forward / pull_back use) is read
// once, 1560 B per N1curl3 cell move instead of 3816.
//
// The k axis of the contraction is taken COMPONENT-MAJOR in blocks (k' = i * PB + p, PB = npts padded to 8): during the
// PB / 4 k-steps of block i every lane needs the same column i of its cell's matrix, so the pulled-back A fragment
//     covariant       U[p][i] =  sum_k J[k][i] u[p][k]                      (maps.h:73-94 with (K, 1 / detJ, J))
//     contravariant   U[p][i] = (sum_k K[i][k] u[p][k]) / (1 / detJ)        (maps.h:100-124 with (K, 1 / detJ, J))
// costs three loads per POINT step (shared by the D blocks) and D * D multiply-adds with the cell's matrix in registers --
// no selects, no shared-memory staging.  (Folding the map into a point-major k axis needs a different matrix column at
// every k and is issue-bound.)  The loop walks the point steps; each feeds D k-steps (one per block) of 2 NT DMMAs.
// The map is evaluated with fused multiply-adds (and the division by 1 / detJ as a multiplication by detJ) and the
// contraction sums in another k order than bx_interpolate, so the result equals bx_pull_back + bx_interpolate to
// rounding (observed <= 1e-15 relative, tested at 1e-13), not bit for bit.  Twelve warps (the cell matrices and the
// three-wide value queue need 160 registers).
namespace
{
constexpr int kPullWarps = 12;
constexpr int kPullThreads = 32 * kPullWarps;

template <typename T>
struct PullParams
{
  const T* u;      // [nc][npts][D]
  const T* A;      // [nc][D][D]: J (covariant) or K (contravariant)
  const T* detJ;   // [nc] (contravariant)
  const double* B; // [npad][kstride], columns k' = i * PB + p
  T* out;          // [nc][out_stride]
  size_t nc;
  int npts, PB, N, npad, kstride, n_base, out_stride;
};

template <int NT, int MAP, int D, typename T>
__device__ __forceinline__ void run_group_pulled(const PullParams<T>& p, const double* __restrict__ Bs, size_t c0, int n0, int g,
                                                 int t)
{
  double acc[2][NT][2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < NT; ++j)
      acc[m][j][0] = acc[m][j][1] = 0.0;
  const size_t r[2] = {min(c0 + g, p.nc - 1), min(c0 + g + 8, p.nc - 1)}; // a row past the end is computed, never stored
  double Am[2][D * D];
  [[maybe_unused]] double det[2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
  {
#pragma unroll
    for (int k = 0; k < D * D; ++k)
      Am[m][k] = double(__ldg(p.A + r[m] * (D * D) + k));
    if constexpr (MAP == BX_MAP_CONTRAVARIANT_PIOLA)
    {
      det[m] = double(__ldg(p.detJ + r[m])); // pull_back passes 1 / detJ as the kernel's detJ (finite-element.cpp:2038)
    }
  }
  const T* __restrict__ up[2] = {p.u + r[0] * size_t(p.npts) * D, p.u + r[1] * size_t(p.npts) * D};
  const int last_pt = p.npts - 1;
  auto ldu = [&](int m, int ks, double (&dst)[D])
  {
    const int pt = min(4 * ks + t, last_pt); // padded points: any finite value, their operand rows are zero
#pragma unroll
    for (int k = 0; k < D; ++k)
      dst[k] = double(__ldg(up[m] + pt * D + k));
  };
  const double* __restrict__ bp = Bs + (n0 + g) * p.kstride + t;
  const int ts = 8 * p.kstride;
  const int nks = p.PB >> 2; // point steps (even)
  double uq[2][2][D]; // [slot][row tile][component]: values two point steps ahead
  double b[2][NT];
#pragma unroll
  for (int m = 0; m < 2; ++m)
  {
    ldu(m, 0, uq[0][m]);
    ldu(m, 1, uq[1][m]);
  }
#pragma unroll
  for (int j = 0; j < NT; ++j)
    b[0][j] = bp[j * ts];
#pragma unroll 1
  for (int ks = 0; ks < nks; ks += 2)
  {
#pragma unroll
    for (int s = 0; s < 2; ++s)
    {
      // pulled-back A fragments of this point step, all D blocks
      double a[2][D];
#pragma unroll
      for (int m = 0; m < 2; ++m)
      {
#pragma unroll
        for (int i = 0; i < D; ++i)
        {
          double v = (MAP == BX_MAP_COVARIANT_PIOLA ? Am[m][i] : Am[m][i * D]) * uq[s][m][0];
#pragma unroll
          for (int k = 1; k < D; ++k)
            v = fma(MAP == BX_MAP_COVARIANT_PIOLA ? Am[m][k * D + i] : Am[m][i * D + k], uq[s][m][k], v);
          if constexpr (MAP == BX_MAP_CONTRAVARIANT_PIOLA)
            v *= det[m]; // the reference divides by 1 / detJ
          a[m][i] = v;
        }
        ldu(m, ks + s + 2, uq[s][m]); // two point steps ahead (clamped inside ldu)
      }
#pragma unroll
      for (int i = 0; i < D; ++i)
      {
        const int cur = (s * D + i) & 1;
        // next k-step: the next block of this point step, or block 0 of the next point step
        const int nb = i + 1 < D ? i + 1 : 0;
        const int nk = i + 1 < D ? ks + s : min(ks + s + 1, nks - 1);
#pragma unroll
        for (int j = 0; j < NT; ++j)
          b[cur ^ 1][j] = bp[j * ts + 4 * (nb * nks + nk)];
#pragma unroll
        for (int j = 0; j < NT; ++j)
        {
          dmma(acc[0][j][0], acc[0][j][1], a[0][i], b[cur][j]);
          dmma(acc[1][j][0], acc[1][j][1], a[1][i], b[cur][j]);
        }
      }
    }
  }
  const bool even = ((p.out_stride | p.n_base) & 1) == 0;
#pragma unroll
  for (int m = 0; m < 2; ++m)
  {
    const size_t c = c0 + g + 8 * m;
    if (c < p.nc)
    {
      T* row = p.out + c * size_t(p.out_stride) + p.n_base;
#pragma unroll
      for (int j = 0; j < NT; ++j)
      {
        const int n = n0 + j * 8 + 2 * t;
        if constexpr (sizeof(T) == 8)
        {
          if (even)
          {
            if (n < p.N)
              __stcs(reinterpret_cast<double2*>(row + n), make_double2(acc[m][j][0], acc[m][j][1]));
            continue;
          }
        }
        if (n < p.N)
          __stcs(row + n, static_cast<T>(acc[m][j][0]));
        if (n + 1 < p.N)
          __stcs(row + n + 1, static_cast<T>(acc[m][j][1]));
      }
    }
  }
}

template <int MAP, int D, typename T>
__global__ void __launch_bounds__(kPullThreads, 1) interpolate_pulled_kernel(const PullParams<T> p)
{
  extern __shared__ __align__(16) double interp_smem[];
  double* Bs = interp_smem;
  for (int i = threadIdx.x; i < p.npad * p.kstride; i += kPullThreads)
    Bs[i] = p.B[i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const size_t nchunks = ceil_div(p.nc, size_t(kCells));
  const size_t stride = size_t(gridDim.x) * kPullWarps;
  for (size_t chunk = size_t(blockIdx.x) * kPullWarps + warp; chunk < nchunks; chunk += stride)
  {
    const size_t c0 = chunk * kCells;
    for (int n0 = 0; n0 < p.N; n0 += 8 * kMaxNT)
    {
      const int nt = min(kMaxNT, (p.N - n0 + 7) >> 3);
      switch (nt)
      {
      case 1: run_group_pulled<1, MAP, D, T>(p, Bs, c0, n0, g, t); break;
      case 2: run_group_pulled<2, MAP, D, T>(p, Bs, c0, n0, g, t); break;
      case 3: run_group_pulled<3, MAP, D, T>(p, Bs, c0, n0, g, t); break;
      case 4: run_group_pulled<4, MAP, D, T>(p, Bs, c0, n0, g, t); break;
      case 5: run_group_pulled<5, MAP, D, T>(p, Bs, c0, n0, g, t); break;
      case 6: run_group_pulled<6, MAP, D, T>(p, Bs, c0, n0, g, t); break;
      default: run_group_pulled<7, MAP, D, T>(p, Bs, c0, n0, g, t); break;
      }
    }
  }
}

template <int MAP, int D, typename T>
void launch_pulled(const bx_element& e, PullParams<T> p, cudaStream_t s)
{
  const size_t row_bytes = size_t(p.kstride) * sizeof(double);
  const int rows_fit = int(std::min<size_t>(size_t(e.interp_npad), (size_t(200) * 1024 / row_bytes) / 8 * 8));
  if (rows_fit < 8)
    fail(BX_ERR_UNSUPPORTED, "interpolate: a block of eight rows of the interpolation matrix does not fit in shared memory");
  const size_t nblocks = ceil_div(p.nc, size_t(kCells) * kPullWarps);
  const unsigned grid = unsigned(std::min<size_t>(nblocks, size_t(sm_count())));
  auto kern = interpolate_pulled_kernel<MAP, D, T>;
  for (int n0 = 0; n0 < e.ndofs; n0 += rows_fit)
  {
    p.n_base = n0;
    p.N = std::min(rows_fit, e.ndofs - n0);
    p.npad = std::min(rows_fit, e.interp_npad - n0);
    p.B = e.d_interp_blocked + size_t(n0) * p.kstride;
    const size_t smem = size_t(p.npad) * row_bytes;
    BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kern<<<grid, kPullThreads, smem, s>>>(p);
    BX_LAUNCHED();
  }
}
} // namespace

// u[nc][npts][gdim] physical values at the interpolation points; J[nc][gdim][tdim], detJ[nc], K[nc][tdim][gdim] per cell
// -> coefficients[nc][ndofs] = interpolation_matrix * pull_back(u).  Covariant and contravariant Piola maps on square
// Jacobians (gdim == tdim == value size); other elements: bx_pull_back + bx_interpolate.
template <typename T>
void interpolate_pulled_device(const bx_element& e, const T* u, const T* J, const T* detJ, const T* K, size_t nc, int gdim,
                               T* coefficients, cudaStream_t s)
{
  if (e.interp_npts == 0)
    fail(BX_ERR_RUNTIME, "interpolate: the element was built without points() / interpolation_matrix()");
  const int map = e.map_type, D = e.tdim;
  if ((map != BX_MAP_COVARIANT_PIOLA and map != BX_MAP_CONTRAVARIANT_PIOLA) or gdim != D or e.vs != D or (D != 2 and D != 3)
      or !e.d_interp_blocked)
    fail(BX_ERR_UNSUPPORTED, "interpolate_pulled: covariant / contravariant Piola elements on square Jacobians (gdim == tdim "
                             "== value size) are fused; use pull_back and interpolate separately otherwise");
  if (map == BX_MAP_COVARIANT_PIOLA ? !J : (!K or !detJ))
    fail(BX_ERR_INVALID_ARG, "interpolate_pulled: the covariant pull_back needs J, the contravariant one K and detJ");
  if (nc == 0)
    return;
  if (!u or !coefficients)
    fail(BX_ERR_INVALID_ARG, "interpolate_pulled: null pointer");
  require_device(e);
  PullParams<T> p;
  p.u = u;
  p.A = map == BX_MAP_COVARIANT_PIOLA ? J : K;
  p.detJ = detJ;
  p.out = coefficients;
  p.nc = nc;
  p.npts = e.interp_npts;
  p.PB = e.interp_pb;
  p.kstride = e.interp_kstride_blocked;
  p.out_stride = e.ndofs;
  if (map == BX_MAP_COVARIANT_PIOLA)
  {
    if (D == 3)
      launch_pulled<BX_MAP_COVARIANT_PIOLA, 3, T>(e, p, s);
    else
      launch_pulled<BX_MAP_COVARIANT_PIOLA, 2, T>(e, p, s);
  }
  else
  {
    if (D == 3)
      launch_pulled<BX_MAP_CONTRAVARIANT_PIOLA, 3, T>(e, p, s);
    else
      launch_pulled<BX_MAP_CONTRAVARIANT_PIOLA, 2, T>(e, p, s);
  }
}
template void interpolate_pulled_device<double>(const bx_element&, const double*, const double*, const double*, const double*,
                                                size_t, int, double*, cudaStream_t);
template void interpolate_pulled_device<float>(const bx_element&, const float*, const float*, const float*, const float*, size_t,
                                               int, float*, cudaStream_t);

// ---- compute_interpolation_operator (reference cpp/basix/interpolation.cpp:16-116) --------------------------------
// The matrix that maps the coefficients of a function in `from` to the coefficients of its interpolant in `to`:
// tabulate `from` at the interpolation points of `to`, apply `to`'s interpolation matrix.  Composition of two device
// paths of this library: FiniteElement::tabulate (any kernel of tabulate.cu) and the batched interpolation above, with
// the columns of the operator playing the role of cells; two small kernels reorder the table into `to`'s value layout
// and the coefficients into the reference's output layout.
namespace
{
// values[(j, i)][c]: one "cell" per column of the operator, in the component-major layout of `to`
//   equal value sizes   (cells = dim_from, i = 0):          values[j][k * npts + l] = tab[l][j][k]
//   vs_to == 1          (cells = dim_from * vs_from):       values[(j, i)][l]       = tab[l][j][i]
//   vs_from == 1        (cells = dim_from * vs_to):         values[(j, i)][c]       = c in block i ? tab[l][j][0] : 0
__global__ void interp_op_gather_kernel(const double* __restrict__ tab, int npts, int dim_from, int vs_from, int vs_to,
                                        double* __restrict__ values)
{
  const int K = vs_to * npts;
  const int per = vs_from == vs_to ? 1 : (vs_to == 1 ? vs_from : vs_to);
  const size_t total = size_t(dim_from) * per * K;
  for (size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += size_t(gridDim.x) * blockDim.x)
  {
    const int c = int(t % K);
    const int cell = int(t / K), j = cell / per, i = cell % per;
    const int k = c / npts, l = c % npts;
    double v;
    if (vs_from == vs_to)
      v = tab[(size_t(l) * dim_from + j) * vs_from + k];
    else if (vs_to == 1)
      v = tab[(size_t(l) * dim_from + j) * vs_from + i];
    else
      v = k == i ? tab[size_t(l) * dim_from + j] : 0.0;
    values[t] = v;
  }
}

// coefficients[(j, i)][d] -> out: equal value sizes out[d][j]; vs_to == 1 out[i + d * vs_from][j]; vs_from == 1
// out[d][i + j * vs_to]
__global__ void interp_op_scatter_kernel(const double* __restrict__ coef, int dim_to, int dim_from, int vs_from, int vs_to,
                                         double* __restrict__ out)
{
  const int per = vs_from == vs_to ? 1 : (vs_to == 1 ? vs_from : vs_to);
  const size_t total = size_t(dim_from) * per * dim_to;
  for (size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += size_t(gridDim.x) * blockDim.x)
  {
    const int d = int(t % dim_to);
    const int cell = int(t / dim_to), j = cell / per, i = cell % per;
    size_t o;
    if (vs_from == vs_to)
      o = size_t(d) * dim_from + j;
    else if (vs_to == 1)
      o = (size_t(i) + size_t(d) * vs_from) * dim_from + j;
    else
      o = size_t(d) * (size_t(dim_from) * vs_to) + i + size_t(j) * vs_to;
    out[o] = coef[t];
  }
}
} // namespace

void interpolation_operator_shape(const bx_element& from, const bx_element& to, int32_t shape[2])
{
  if (from.cell_type != to.cell_type)
    fail(BX_ERR_RUNTIME, "Cannot interpolate between elements defined on different cell types.");
  if (from.vs == to.vs)
  {
    shape[0] = to.ndofs;
    shape[1] = from.ndofs;
  }
  else if (to.vs == 1)
  {
    shape[0] = to.ndofs * from.vs;
    shape[1] = from.ndofs;
  }
  else if (from.vs == 1)
  {
    shape[0] = to.ndofs;
    shape[1] = from.ndofs * to.vs;
  }
  else
    fail(BX_ERR_RUNTIME, "Cannot interpolate between elements with this combination of value sizes.");
}

template <typename T>
void tabulate_device(const bx_element& e, int nd, const T* x, size_t npts, T* basis, size_t out_pts, cudaStream_t s);

// out: device pointer, shape as interpolation_operator_shape, row-major
void interpolation_operator_device(const bx_element& from, const bx_element& to, double* out, cudaStream_t s)
{
  int32_t shape[2];
  interpolation_operator_shape(from, to, shape);
  if (to.interp_npts == 0)
    fail(BX_ERR_RUNTIME, "compute_interpolation_operator: the target element was built without points() / "
                         "interpolation_matrix()");
  require_device(from);
  require_device(to);
  const int npts = to.interp_npts, K = to.vs * npts;
  const int per = from.vs == to.vs ? 1 : (to.vs == 1 ? from.vs : to.vs);
  const size_t ncells = size_t(from.ndofs) * per;
  double *x = nullptr, *tab = nullptr, *values = nullptr, *coef = nullptr;
  BX_CUDA(cudaMallocAsync(&x, size_t(npts) * to.tdim * sizeof(double), s));
  BX_CUDA(cudaMallocAsync(&tab, size_t(npts) * from.ndofs * from.vs * sizeof(double), s));
  BX_CUDA(cudaMallocAsync(&values, (ncells * K + 64) * sizeof(double), s)); // slack: the kernel prefetches A fragments
  BX_CUDA(cudaMallocAsync(&coef, ncells * to.ndofs * sizeof(double), s));
  auto release = [&]()
  {
    cudaFreeAsync(x, s);
    cudaFreeAsync(tab, s);
    cudaFreeAsync(values, s);
    cudaFreeAsync(coef, s);
  };
  try
  {
    BX_CUDA(cudaMemcpyAsync(x, to.points.data(), size_t(npts) * to.tdim * sizeof(double), cudaMemcpyHostToDevice, s));
    tabulate_device<double>(from, 0, x, size_t(npts), tab, size_t(npts), s);
    BX_CUDA(cudaMemsetAsync(values + ncells * K, 0, 64 * sizeof(double), s));
    const int grid = int(std::min<size_t>(ceil_div(ncells * K, size_t(256)), size_t(sm_count()) * 8));
    interp_op_gather_kernel<<<std::max(grid, 1), 256, 0, s>>>(tab, npts, from.ndofs, from.vs, to.vs, values);
    BX_LAUNCHED();
    interpolate_device<double>(to, BX_INTERP_COMPONENT_MAJOR, values, ncells, coef, s);
    const int grid2 = int(std::min<size_t>(ceil_div(ncells * to.ndofs, size_t(256)), size_t(sm_count()) * 8));
    interp_op_scatter_kernel<<<std::max(grid2, 1), 256, 0, s>>>(coef, to.ndofs, from.ndofs, from.vs, to.vs, out);
    BX_LAUNCHED();
  }
  catch (...)
  {
    release();
    throw;
  }
  release();
}
} // namespace bx

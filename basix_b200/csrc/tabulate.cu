// FiniteElement::tabulate on the device.
//
// Replaces FiniteElement<F>::tabulate (reference cpp/basix/finite-element.cpp:1834-1888):
//   basis[d][pt][dof][j] = sum_k C[dof*vs + j][k] * P[d][k][pt]      (+ optional dof_ordering)
// i.e. polyset::tabulate (polyset.cpp) -> math::dot/dgemm (math.h:323-366) -> transposing scatter
// (finite-element.cpp:1873-1886).
//
// Generic path (any element the polyset kernels support): polyset kernel writes P[d][k][pt] for a
// chunk of points into a stream-ordered scratch buffer, then an FP64 tensor-core (DMMA m8n8k4)
// kernel contracts it with the padded coefficient matrix and stores straight into the
// [d][pt][dof*vs] layout -- the reference's separate transpose pass does not exist here because the
// MMA accumulator fragment already has (point, dof) orientation.
#include "element.cuh"
#include "fused_menu.cuh"

#include <algorithm>
#include <atomic>

namespace bx
{
template <typename T>
void polyset_tabulate_device(int cell, int ptype, int degree, int nd, const T* x, size_t npts, T* P,
                             cudaStream_t stream);

template <typename T>
void tabulate_device(const bx_element& e, int nd, const T* x, size_t npts, T* basis, size_t out_pts, cudaStream_t s);

namespace
{
constexpr int kWarps = 4;
constexpr int kNG = 4; // n-tiles (8 columns each) accumulated per pass

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b)
{
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
}

// One warp: 8 points (m) x all N columns of one derivative slab.
//   A (8 x 4, row)  = P[d][k0 + t][pt0 + g]          t = lane & 3, g = lane >> 2
//   B (4 x 8, col)  = Cp[n0 + g][k0 + t]
//   D (8 x 8)       : lane holds (row g, cols 2t, 2t+1)  ->  out[d][pt0 + g][n0 + 2t ..]
__global__ void __launch_bounds__(kWarps * 32)
    contract_dmma_kernel(const double* __restrict__ P, size_t ldp, const double* __restrict__ Cp, double* out,
                         size_t out_pts, size_t npts, int nder, int K, int Kpad, int N, int Npad)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, t = lane & 3;
  const size_t mtiles = ceil_div(npts, size_t(8 * kWarps)); // CTA tiles of 32 points
  const size_t total = mtiles * nder;
  for (size_t w = blockIdx.x; w < total; w += gridDim.x)
  {
    const int d = int(w / mtiles);
    const size_t pt0 = (w % mtiles) * (8 * kWarps) + warp * 8;
    if (pt0 >= npts)
      continue;
    const size_t pt = pt0 + g;
    const bool pt_ok = pt < npts;
    const double* Pd = P + size_t(d) * K * ldp + (pt_ok ? pt : npts - 1);
    double* od = out + (size_t(d) * out_pts + pt) * N;
    for (int n0 = 0; n0 < Npad; n0 += 8 * kNG)
    {
      double acc[kNG][2];
#pragma unroll
      for (int j = 0; j < kNG; ++j)
        acc[j][0] = acc[j][1] = 0.0;
      for (int k0 = 0; k0 < Kpad; k0 += 4)
      {
        const int k = k0 + t;
        const double a = (k < K) ? Pd[size_t(k) * ldp] : 0.0;
#pragma unroll
        for (int j = 0; j < kNG; ++j)
        {
          const int nrow = n0 + 8 * j + g;
          const double b = (nrow < Npad) ? __ldg(Cp + size_t(nrow) * Kpad + k) : 0.0;
          dmma884(acc[j][0], acc[j][1], a, b);
        }
      }
      if (pt_ok)
      {
#pragma unroll
        for (int j = 0; j < kNG; ++j)
        {
          const int n = n0 + 8 * j + 2 * t;
          if ((N & 1) == 0)
          {
            if (n < N)
              *reinterpret_cast<double2*>(od + n) = make_double2(acc[j][0], acc[j][1]);
          }
          else
          {
            if (n < N)
              od[n] = acc[j][0];
            if (n + 1 < N)
              od[n + 1] = acc[j][1];
          }
        }
      }
    }
  }
}
} // namespace

// x[npts][tdim] -> basis[nder][out_pts][N] rows [0, npts) of each slab (out_pts >= npts lets a caller
// tabulate a chunk into a larger table).  Generic two-kernel path.
void tabulate_generic_device(const bx_element& e, int nd, const double* x, size_t npts, double* basis,
                             size_t out_pts, cudaStream_t s)
{
  const int nder = polyset_nderivs(e.cell_type, nd);
  // scratch for P is bounded to ~1 GiB: chunk the points
  const size_t per_pt = size_t(nder) * e.K * sizeof(double);
  size_t chunk = (size_t(1) << 30) / per_pt;
  chunk = std::max<size_t>(chunk / 32 * 32, 32);
  chunk = std::min(chunk, size_t(ceil_div(npts, 32) * 32));
  double* P = nullptr;
  BX_CUDA(cudaMallocAsync(&P, chunk * per_pt, s));
  try
  {
    for (size_t p0 = 0; p0 < npts; p0 += chunk)
    {
      const size_t np = std::min(chunk, npts - p0);
      polyset_tabulate_device<double>(e.cell_type, e.polyset_type, e.superdegree, nd, x + p0 * e.tdim, np, P, s);
      const size_t work = ceil_div(np, size_t(8 * kWarps)) * nder;
      const size_t cap = size_t(sm_count()) * 16;
      const int grid = int(work < cap ? work : cap);
      contract_dmma_kernel<<<grid, kWarps * 32, 0, s>>>(P, np, e.d_Cp, basis + p0 * e.N, out_pts, np, nder, e.K,
                                                        e.Kpad, e.N, e.Npad);
      BX_LAUNCHED();
    }
  }
  catch (...)
  {
    cudaFreeAsync(P, s);
    throw;
  }
  BX_CUDA(cudaFreeAsync(P, s));
}
} // namespace bx

namespace bx
{
template <typename T>
bool tabulate_tp_device(const bx_element& e, int nd, const T* x, size_t npts, T* basis, cudaStream_t s, bool dry_run);
namespace small
{
template <typename T>
bool launch_any(const bx_element& e, int nd, const T* x, size_t npts, T* basis, cudaStream_t s, bool dry_run);
}
namespace dop
{
template <typename T>
bool launch_any(const bx_element& e, int nd, const T* x, size_t npts, T* basis, cudaStream_t s, bool dry_run);
}
namespace
{
// Fused simplex kernel (tabulate_fused.cuh) when the element has the operand and (degree, nd) is on the menu.
bool try_fused(const bx_element& e, int nd, const double* x, size_t npts, double* basis, cudaStream_t s, bool dry_run)
{
  if (e.Cs.empty() or (!dry_run and !e.d_Cs))
    return false;
  switch (e.cell_type)
  {
  case BX_CELL_INTERVAL: return fused::launch_interval(e, nd, x, npts, basis, s, dry_run);
  case BX_CELL_TRIANGLE: return fused::launch_triangle(e, nd, x, npts, basis, s, dry_run);
  case BX_CELL_TETRAHEDRON: return fused::launch_tetrahedron(e, nd, x, npts, basis, s, dry_run);
  default: return false;
  }
}

// float32 API on the paths that have no float instantiation (fused simplex, generic): widen the points, run the
// float64 path on a bounded chunk, narrow the table.  All on the device, stream ordered.
__global__ void widen_kernel(const float* __restrict__ in, size_t n, double* __restrict__ out)
{
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = double(in[i]);
}
// in[nder][np * N] (dense chunk) -> out[(d * out_pts + p0) * N + i]
__global__ void narrow_slabs_kernel(const double* __restrict__ in, size_t run, int nder, float* __restrict__ out,
                                    size_t slab)
{
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < run * nder; i += stride)
  {
    const size_t d = i / run, k = i - d * run;
    out[d * slab + k] = float(in[i]);
  }
}
} // namespace

// Which single-kernel paths tabulate() may take (bit i = path i below; the generic path is always available).
// All enabled by default; tests and profiling runs narrow the mask to reach a specific kernel.
std::atomic<unsigned> g_path_mask{~0u};
unsigned tabulate_path_mask(unsigned mask) { return g_path_mask.exchange(mask); }
bool tabulate_path_allowed(int path) { return (g_path_mask.load(std::memory_order_relaxed) >> path) & 1u; }
namespace
{
bool allowed(int path) { return tabulate_path_allowed(path); }
} // namespace

// Which device path tabulate() takes: 0 generic, 1 fused simplex, 2 tensor product, 3 register path,
// 4 derivative-operator path.
int tabulate_path(const bx_element& e, int nd)
{
  if (allowed(3) and small::launch_any<double>(e, nd, nullptr, 0, nullptr, nullptr, true))
    return 3;
  if (allowed(4) and dop::launch_any<double>(e, nd, nullptr, 0, nullptr, nullptr, true))
    return 4;
  if (allowed(1) and try_fused(e, nd, nullptr, 0, nullptr, nullptr, true))
    return 1;
  if (allowed(2) and tabulate_tp_device<double>(e, nd, nullptr, 0, nullptr, nullptr, true))
    return 2;
  return 0;
}

template <>
void tabulate_device<double>(const bx_element& e, int nd, const double* x, size_t npts, double* basis, size_t out_pts,
                             cudaStream_t s)
{
  require_device(e);
  // the single-kernel paths write a dense [nder][npts][N] table; a chunk of a larger table takes the generic path
  if (out_pts == npts and allowed(3) and small::launch_any<double>(e, nd, x, npts, basis, s, false))
    return;
  if (out_pts == npts and allowed(4) and dop::launch_any<double>(e, nd, x, npts, basis, s, false))
    return;
  if (out_pts == npts and allowed(1) and try_fused(e, nd, x, npts, basis, s, false))
    return;
  if (out_pts == npts and allowed(2) and tabulate_tp_device<double>(e, nd, x, npts, basis, s, false))
    return;
  tabulate_generic_device(e, nd, x, npts, basis, out_pts, s);
}

template <>
void tabulate_device<float>(const bx_element& e, int nd, const float* x, size_t npts, float* basis, size_t out_pts,
                            cudaStream_t s)
{
  require_device(e);
  if (out_pts == npts and allowed(3) and small::launch_any<float>(e, nd, x, npts, basis, s, false))
    return;
  if (out_pts == npts and allowed(4) and dop::launch_any<float>(e, nd, x, npts, basis, s, false))
    return;
  if (out_pts == npts and allowed(2) and tabulate_tp_device<float>(e, nd, x, npts, basis, s, false))
    return;
  const int nder = polyset_nderivs(e.cell_type, nd);
  const size_t per_pt = (size_t(e.tdim) + size_t(nder) * e.N) * sizeof(double);
  size_t chunk = std::max<size_t>(32, ((size_t(512) << 20) / per_pt) / 32 * 32);
  chunk = std::min(chunk, npts);
  double *x64 = nullptr, *b64 = nullptr;
  BX_CUDA(cudaMallocAsync(&x64, chunk * e.tdim * sizeof(double), s));
  BX_CUDA(cudaMallocAsync(&b64, chunk * size_t(nder) * e.N * sizeof(double), s));
  try
  {
    const int grid = sm_count() * 8;
    for (size_t p0 = 0; p0 < npts; p0 += chunk)
    {
      const size_t np = std::min(chunk, npts - p0);
      widen_kernel<<<grid, 256, 0, s>>>(x + p0 * e.tdim, np * e.tdim, x64);
      BX_LAUNCHED();
      tabulate_device<double>(e, nd, x64, np, b64, np, s);
      narrow_slabs_kernel<<<grid, 256, 0, s>>>(b64, np * e.N, nder, basis + p0 * e.N, out_pts * e.N);
      BX_LAUNCHED();
    }
  }
  catch (...)
  {
    cudaFreeAsync(x64, s);
    cudaFreeAsync(b64, s);
    throw;
  }
  BX_CUDA(cudaFreeAsync(x64, s));
  BX_CUDA(cudaFreeAsync(b64, s));
}
} // namespace bx

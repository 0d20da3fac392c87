// Tensor-product fast path of FiniteElement::tabulate for quadrilateral / hexahedron elements.
//
// On these cells the orthonormal set is a product of 1-D Legendre polynomials (reference
// cpp/basix/polyset.cpp:2241-2702).  When, in addition, every row of the coefficient matrix is a
// rank-1 tensor (true for Lagrange-type Q elements; detected numerically on the host, element.cu:
// factor_tensor_product) the dense contraction of finite-element.cpp:1867-1872,
//     basis[d][pt][n] = sum_{px,py,pz} C[n][px,py,pz] L^(kx)_px(x) L^(ky)_py(y) L^(kz)_pz(z),
// collapses to  s_n * phi_x[kx][i_n](x) * phi_y[ky][j_n](y) * phi_z[kz][k_n](z)  with a handful of 1-D
// functions per axis.  That turns a compute-bound 125 kflop/point GEMM (Q4 hex) into ~1.5 kflop/point
// and leaves the kernel bound by the HBM write of the table.
//
// Kernel: a warp owns 32 consecutive points.
//   phase 1  lane = point: 1-D Legendre recurrence per axis -> 1-D basis values/derivatives -> shared memory
//   phase 2  lane = output column: for each of the 32 points the warp writes the N columns of every
//            derivative slab with fully coalesced stores (consecutive lanes = consecutive columns).
#include "element.cuh"

#include <algorithm>
#include <type_traits>
#include <cstdlib>
#include <string>

namespace bx
{
namespace
{
constexpr int kWarps = 8;
constexpr int kMaxN1 = 9;   // degree + 1

// T = type of the points and of the table (double, or float for the float32 API; arithmetic is double)
template <typename T>
struct TPParams
{
  const T* x;
  size_t npts;
  T* basis;
  const double* A;       // [TDIM][fstride][n1] 1-D expansion coefficients, sqrt(2q+1) norms folded in
  const int32_t* ijk;    // [N] packed i | j << 8 | k << 16
  const double* scale;   // [N]
  int n1, N, fstride;    // fstride = max number of 1-D functions over the axes
  int nfun[3];
  int plain_stores; // tuning aid (BX_TP_TUNE=plain): default write-back stores instead of streaming stores
};

// Un-normalised shifted Legendre values and derivatives L[k][p] = d^k/dx^k P_p(2x - 1), p < n1
// (recurrence of polyset.cpp:66-110; the sqrt(2p+1) norms are folded into A on the host).
template <int NK>
__device__ __forceinline__ void legendre_raw(double (&L)[NK][kMaxN1], int n1, double x)
{
  const double X = x * 2.0 - 1.0;
#pragma unroll
  for (int k = 0; k < NK; ++k)
    L[k][0] = k == 0 ? 1.0 : 0.0;
#pragma unroll
  for (int q = 1; q < kMaxN1; ++q)
  {
    if (q < n1)
    {
      const double a = 1.0 - 1.0 / double(q);
#pragma unroll
      for (int k = 0; k < NK; ++k)
      {
        double v = (X * (a + 1.0)) * L[k][q - 1];
        if (k > 0)
          v += (2.0 * k * (a + 1.0)) * L[k - 1][q - 1];
        if (q > 1)
          v -= a * L[k][q - 2];
        L[k][q] = v;
      }
    }
  }
}

// Shared-memory row of one point: [axis][function][NKP] with NKP = derivative orders rounded up to even, so a
// lane fetches the value and first derivative of a 1-D factor with ONE 16-byte load.  Row stride: even
// (16-byte aligned rows) with an odd number of 16-byte units (8 lanes on 8 points hit 8 different bank groups).
__host__ __device__ constexpr int tp_nkp(int nk) { return nk == 1 ? 1 : (nk + 1) / 2 * 2; }
__host__ __device__ constexpr int tp_pstride(int tdim, int nk, int F)
{
  const int base = tdim * F * tp_nkp(nk);
  const int even = base + (base & 1);
  return (even / 2) % 2 == 1 ? even : even + 2;
}

template <int TDIM, int ND, int MD, typename T>
__global__ void __launch_bounds__(kWarps * 32) tabulate_tp_kernel(TPParams<T> p)
{
  constexpr int NK = ND + 1;
  constexpr int NDER = TDIM == 2 ? (ND + 1) * (ND + 2) / 2 : (ND + 1) * (ND + 2) * (ND + 3) / 6;
  extern __shared__ __align__(16) double smem[];
  const int F = p.fstride;
  constexpr int NKP = tp_nkp(NK);
  const int pstride = tp_pstride(TDIM, NK, F);
  double* As = smem;                              // [TDIM][F][n1]
  double* Phi = As + ((TDIM * F * p.n1 + 1) & ~1); // [kWarps*32][pstride], 16-byte aligned
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < TDIM * F * p.n1; i += blockDim.x)
    As[i] = p.A[i];
  __syncthreads();

  // per-lane output columns (fixed for the whole kernel)
  int ci[MD], cj[MD], ck[MD];
  double cs[MD];
#pragma unroll
  for (int m = 0; m < MD; ++m)
  {
    const int n = lane + 32 * m;
    const int32_t v = n < p.N ? p.ijk[n] : 0;
    ci[m] = v & 255;
    cj[m] = (v >> 8) & 255;
    ck[m] = (v >> 16) & 255;
    cs[m] = n < p.N ? p.scale[n] : 0.0;
  }

  double* myPhi = Phi + size_t(warp * 32) * pstride;
  const size_t ntiles = ceil_div(p.npts, size_t(32));
  for (size_t tile = size_t(blockIdx.x) * kWarps + warp; tile < ntiles; tile += size_t(gridDim.x) * kWarps)
  {
    // ---- phase 1: lane = point ----
    {
      size_t pt = tile * 32 + lane;
      pt = pt < p.npts ? pt : p.npts - 1;
      double* out = myPhi + lane * pstride;
#pragma unroll
      for (int a = 0; a < TDIM; ++a)
      {
        double L[NK][kMaxN1];
        legendre_raw<NK>(L, p.n1, double(p.x[pt * TDIM + a]));
        for (int f = 0; f < p.nfun[a]; ++f)
        {
          const double* Af = As + (a * F + f) * p.n1;
#pragma unroll
          for (int k = 0; k < NK; ++k)
          {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < kMaxN1; ++q)
              if (q < p.n1)
                acc += Af[q] * L[k][q];
            out[(a * F + f) * NKP + k] = acc;
          }
        }
      }
    }
    __syncwarp();
    // ---- phase 2: lane = output column ----
    // Point q outermost, column group m innermost, so the N columns of one table row are written back to back
    // (rows complete in L2 before they are evicted; the other loop order halves the throughput).  Everything
    // that depends only on the column group -- the shared-memory offsets of its three 1-D factors, its scale --
    // is hoisted out of the point loop, and one running pointer per derivative slab replaces the 64-bit
    // address arithmetic.
    const int npt = int(min(size_t(32), p.npts - tile * 32));
    const bool plain = p.plain_stores != 0;
    auto store = [plain](T* a, double v)
    {
      if (plain)
        *a = static_cast<T>(v);
      else
        __stcs(a, static_cast<T>(v));
    };
    int ox[MD], oy[MD], oz[MD];
#pragma unroll
    for (int m = 0; m < MD; ++m)
    {
      ox[m] = ci[m] * NKP;
      oy[m] = (F + cj[m]) * NKP;
      oz[m] = (2 * F + ck[m]) * NKP;
    }
    T* os[NDER];
#pragma unroll
    for (int d = 0; d < NDER; ++d)
      os[d] = p.basis + size_t(d) * p.npts * p.N + tile * 32 * p.N + lane;
    const double* ph = myPhi;
    for (int q = 0; q < npt; ++q, ph += pstride)
    {
#pragma unroll
      for (int m = 0; m < MD; ++m)
      {
        if (lane + 32 * m < p.N)
        {
          double vx[NKP], vy[NKP], vz[NKP];
          if constexpr (NKP == 1)
          {
            vx[0] = ph[ox[m]];
            vy[0] = ph[oy[m]];
            if constexpr (TDIM == 3)
              vz[0] = ph[oz[m]];
          }
          else
          {
#pragma unroll
            for (int k = 0; k < NKP; k += 2)
            {
              const double2 tx = *reinterpret_cast<const double2*>(ph + ox[m] + k);
              const double2 ty = *reinterpret_cast<const double2*>(ph + oy[m] + k);
              vx[k] = tx.x, vx[k + 1] = tx.y;
              vy[k] = ty.x, vy[k + 1] = ty.y;
              if constexpr (TDIM == 3)
              {
                const double2 tz = *reinterpret_cast<const double2*>(ph + oz[m] + k);
                vz[k] = tz.x, vz[k + 1] = tz.y;
              }
            }
          }
#pragma unroll
          for (int k = 0; k < NK; ++k)
            vx[k] *= cs[m];
          if constexpr (TDIM == 2)
          {
#pragma unroll
            for (int kx = 0; kx <= ND; ++kx)
#pragma unroll
              for (int ky = 0; ky <= ND - kx; ++ky)
                store(os[idx2(kx, ky)] + 32 * m, vx[kx] * vy[ky]);
          }
          else
          {
#pragma unroll
            for (int kx = 0; kx <= ND; ++kx)
#pragma unroll
              for (int ky = 0; ky <= ND - kx; ++ky)
              {
                const double vxy = vx[kx] * vy[ky];
#pragma unroll
                for (int kz = 0; kz <= ND - kx - ky; ++kz)
                  store(os[idx3(kx, ky, kz)] + 32 * m, vxy * vz[kz]);
              }
          }
        }
      }
#pragma unroll
      for (int d = 0; d < NDER; ++d)
        os[d] += p.N;
    }
    __syncwarp();
  }
  (void)NDER;
}

template <int TDIM, int ND, typename T>
void launch_md(int md, const TPParams<T>& p, size_t smem, int grid, cudaStream_t s)
{
#define BX_TP_CASE(MD)                                                                                     \
  case MD:                                                                                                 \
  {                                                                                                        \
    auto kern = tabulate_tp_kernel<TDIM, ND, MD, T>;                                                          \
    if (smem > 48 * 1024)                                                                                  \
      BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));         \
    /* persistent grid = exactly the resident CTAs, so the grid-stride loop has no ragged last wave */      \
    int per_sm = 1;                                                                                        \
    BX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kWarps * 32, smem));              \
    grid = std::min(grid, sm_count() * std::max(per_sm, 1));                                               \
    kern<<<grid, kWarps * 32, smem, s>>>(p);                                                               \
    break;                                                                                                 \
  }
  switch (md)
  {
    BX_TP_CASE(1)
    BX_TP_CASE(2)
    BX_TP_CASE(4)
    BX_TP_CASE(8)
    BX_TP_CASE(16)
    BX_TP_CASE(24)
  default: fail(BX_ERR_UNSUPPORTED, "tensor-product tabulate: too many output columns");
  }
#undef BX_TP_CASE
  BX_LAUNCHED();
}
} // namespace

bool tabulate_path_allowed(int path);
namespace staged
{
template <typename T>
bool launch_any(const bx_element& e, int nd, const T* x, size_t npts, T* basis, cudaStream_t s, bool dry_run);
}

// Returns false when the element has no tensor-product factorisation or the shape is outside the menu.
template <typename T>
bool tabulate_tp_device(const bx_element& e, int nd, const T* x, size_t npts, T* basis, cudaStream_t s, bool dry_run)
{
  if (!e.tp.valid or nd < 0 or nd > 2)
    return false;
  // full tensor grids (Lagrange-type Q elements): one slab of a point tile at a time, staged through shared
  // memory and written as contiguous runs by the TMA engine (tabulate_tp_staged.cu); bit 5 of the path mask
  // (bx_tabulate_path_mask) disables it for tests / profiling of the kernel below
  if (tabulate_path_allowed(5) and staged::launch_any<T>(e, nd, x, npts, basis, s, dry_run))
    return true;
  const int tdim = e.tdim;
  const int n1 = e.superdegree + 1;
  if (n1 > kMaxN1 or e.N > 24 * 32)
    return false;
  if (dry_run)
    return true;
  if (!e.tp.d_A)
    return false;
  TPParams<T> p;
  p.x = x;
  p.npts = npts;
  p.basis = basis;
  p.A = e.tp.d_A;
  p.ijk = e.tp.d_ijk;
  p.scale = e.tp.d_scale;
  p.n1 = n1;
  p.N = e.N;
  p.fstride = e.tp.fstride;
  {
    static const bool plain = []() { const char* v = std::getenv("BX_TP_TUNE"); return v and std::string(v) == "plain"; }();
    p.plain_stores = plain ? 1 : 0;
  }
  for (int a = 0; a < 3; ++a)
    p.nfun[a] = a < tdim ? e.tp.nfun[a] : 0;
  const int per_lane = int(ceil_div(e.N, 32));
  const int md = per_lane <= 1 ? 1 : per_lane <= 2 ? 2 : per_lane <= 4 ? 4 : per_lane <= 8 ? 8 : per_lane <= 16 ? 16 : 24;
  const int pstride = tp_pstride(tdim, nd + 1, p.fstride);
  const size_t smem = (((size_t(tdim) * p.fstride * n1 + 1) & ~size_t(1)) + size_t(kWarps) * 32 * pstride) * sizeof(double);
  if (smem > 200 * 1024)
    return false;
  const size_t ntiles = ceil_div(npts, size_t(32));
  const size_t want = ceil_div(ntiles, size_t(kWarps));
  const size_t cap = size_t(sm_count()) * std::max<size_t>(1, std::min<size_t>(8, (200 * 1024) / std::max<size_t>(smem, 1)));
  const int grid = int(std::max<size_t>(1, std::min(want, cap)));
  if (tdim == 2)
  {
    if (nd == 0)
      launch_md<2, 0, T>(md, p, smem, grid, s);
    else if (nd == 1)
      launch_md<2, 1, T>(md, p, smem, grid, s);
    else
      launch_md<2, 2, T>(md, p, smem, grid, s);
  }
  else
  {
    if (nd == 0)
      launch_md<3, 0, T>(md, p, smem, grid, s);
    else if (nd == 1)
      launch_md<3, 1, T>(md, p, smem, grid, s);
    else
      launch_md<3, 2, T>(md, p, smem, grid, s);
  }
  return true;
}
template bool tabulate_tp_device<double>(const bx_element&, int, const double*, size_t, double*, cudaStream_t, bool);
template bool tabulate_tp_device<float>(const bx_element&, int, const float*, size_t, float*, cudaStream_t, bool);
} // namespace bx

// Staged tensor-product tabulate for Lagrange-type Q elements on quadrilateral / hexahedron (float64 and float32
// tables; the arithmetic is float64 either way).
//
// Same factorisation as tabulate_tp.cu (reference cpp/basix/polyset.cpp:2241-2702 + finite-element.cpp:1867-1872
// collapse to  basis[d][pt][n] = s_n phi_x[kx][i_n](x) phi_y[ky][j_n](y) phi_z[kz][k_n](z)), different data flow.
// The table is pure HBM write traffic (4024 B per point for Q4 on the hexahedron, nd = 1), and what the write
// bandwidth depends on is the store pattern (tools/write_pattern.cu, 10 M points of Q4 / nd = 1):
//     every slab of a point written together, 8-byte stores (tabulate_tp.cu)      5100 GB/s
//     one slab of a 32-point tile at a time, 8-byte stores                          5982 GB/s
//     one slab of a 32-point tile as ONE contiguous run of 16-byte stores           6213 GB/s  (95 % of the copy peak)
// So here a warp owns 32 points and finishes one derivative slab at a time: lane = point keeps the 1-D function
// values and derivatives of its point in REGISTERS (no shared-memory round trip for the factors), walks the
// (i, j, k) index grid with compile-time indices, and writes each product into a warp-private shared-memory tile
// [point][column] (odd pitch: conflict free); the finished 32 x N tile is exactly the contiguous run
// basis[d][32 tile .. 32 tile + 32)[0 .. N) and leaves as ONE TMA bulk copy (cp.async.bulk shared -> global): the
// warp waits only until the engine has read the tile, the HBM write completes in the background.  Measured on Q4
// hexahedron, nd = 1, 50 M points: 32.0 ms = 96 % of the HBM copy peak (tabulate_tp.cu: 39.2 ms, 78 %).
#include "element.cuh"

#include <algorithm>

namespace bx
{
namespace staged
{
template <typename T>
struct Params
{
  const T* x;
  size_t npts;
  T* basis;
  const double* A;      // [TDIM][NF][NF] 1-D expansion coefficients w.r.t. un-normalised Legendre (norms folded in)
  const double2* order; // [NF^TDIM] {scale, column as raw int64 bits} of index triple (i, j, k), lexicographic
  int N, pitch, warps;
  int aligned16;        // every slab region starts on a 16-byte boundary
};

// Un-normalised shifted Legendre values and derivatives L[k][q] = d^k/dx^k P_q(2x - 1), q < NF
// (recurrence of polyset.cpp:66-110; the sqrt(2q+1) norms are folded into A on the host).
template <int NK, int NF>
__device__ __forceinline__ void legendre(double (&L)[NK][NF], double x)
{
  const double X = x * 2.0 - 1.0;
#pragma unroll
  for (int k = 0; k < NK; ++k)
    L[k][0] = k == 0 ? 1.0 : 0.0;
#pragma unroll
  for (int q = 1; q < NF; ++q)
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

// warps per CTA: as many 32 x pitch tiles as fit in 200 KB beside the tables, at most 16
template <int TDIM, int NF, typename T>
constexpr int warps_for()
{
  constexpr int N = TDIM == 2 ? NF * NF : NF * NF * NF;
  constexpr int pitch = (N & 1) ? N : N + 1;
  constexpr long fixed = (2L * N + ((TDIM * NF * NF + 1) & ~1)) * 8;
  constexpr long w = (200L * 1024 - fixed) / (32L * pitch * long(sizeof(T)));
  return int(w > 16 ? 16 : w);
}

template <int TDIM, int ND, int NF, typename T>
__global__ void __launch_bounds__(32 * warps_for<TDIM, NF, T>()) tabulate_tp_staged_kernel(const Params<T> p)
{
  constexpr int NK = ND + 1;
  constexpr int NGRID = TDIM == 2 ? NF * NF : NF * NF * NF;
  extern __shared__ __align__(16) double smem_staged[];
  double2* s_order = reinterpret_cast<double2*>(smem_staged);  // [NGRID]
  double* s_A = smem_staged + 2 * NGRID;                        // [TDIM][NF][NF]
  T* tiles = reinterpret_cast<T*>(s_A + ((TDIM * NF * NF + 1) & ~1)); // [warps][32 * pitch]
  for (int i = threadIdx.x; i < NGRID; i += blockDim.x)
    s_order[i] = p.order[i];
  for (int i = threadIdx.x; i < TDIM * NF * NF; i += blockDim.x)
    s_A[i] = p.A[i];
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.N, pitch = p.pitch;
  T* tile = tiles + size_t(warp) * 32 * pitch;
  T* mine = tile + lane * pitch;
  const size_t ntiles = ceil_div(p.npts, size_t(32));
  for (size_t tl = size_t(blockIdx.x) * p.warps + warp; tl < ntiles; tl += size_t(gridDim.x) * p.warps)
  {
    // ---- 1-D functions of this lane's point, in registers: f[axis][derivative order][function] ----
    size_t pt = tl * 32 + lane;
    pt = pt < p.npts ? pt : p.npts - 1; // ragged tail: recompute the last point, its rows are never copied out
    double f[TDIM][NK][NF];
#pragma unroll
    for (int a = 0; a < TDIM; ++a)
    {
      double L[NK][NF];
      legendre<NK, NF>(L, double(p.x[pt * TDIM + a]));
#pragma unroll
      for (int fn = 0; fn < NF; ++fn)
#pragma unroll
        for (int k = 0; k < NK; ++k)
        {
          double acc = 0.0;
#pragma unroll
          for (int q = 0; q < NF; ++q)
            acc += s_A[(a * NF + fn) * NF + q] * L[k][q];
          f[a][k][fn] = acc;
        }
    }
    const unsigned nrows = unsigned(min(size_t(32), p.npts - tl * 32));
    const unsigned run = nrows * unsigned(N); // doubles of one slab region of this tile
    // ---- one derivative slab at a time (slab index = idx2 / idx3 of the derivative multi-index) ----
#pragma unroll
    for (int kx = 0; kx <= ND; ++kx)
#pragma unroll
      for (int ky = 0; ky <= ND - kx; ++ky)
#pragma unroll
        for (int kz = 0; kz <= (TDIM == 3 ? ND - kx - ky : 0); ++kz)
        {
          const int d = TDIM == 2 ? idx2(kx, ky) : idx3(kx, ky, kz);
#pragma unroll
          for (int i = 0; i < NF; ++i)
#pragma unroll
            for (int j = 0; j < NF; ++j)
            {
              const double pxy = f[0][kx][i] * f[1][ky][j];
              if constexpr (TDIM == 2)
              {
                const double2 o = s_order[i * NF + j];
                mine[int(__double_as_longlong(o.y))] = static_cast<T>(pxy * o.x);
              }
              else
              {
#pragma unroll
                for (int k = 0; k < NF; ++k)
                {
                  const double2 o = s_order[(i * NF + j) * NF + k];
                  mine[int(__double_as_longlong(o.y))] = static_cast<T>((pxy * f[2][kz][k]) * o.x);
                }
              }
            }
          T* g = p.basis + (size_t(d) * p.npts + tl * 32) * N;
          if (pitch == N and p.aligned16 and (run * sizeof(T)) % 16 == 0)
          {
            // the tile IS the contiguous run g[0 .. run): one bulk copy shared -> global by the TMA engine.  The
            // warp only waits until the engine has READ the tile (then the next slab may overwrite it); the
            // write to HBM completes in the background.
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0)
            {
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g),
                           "r"(uint32_t(__cvta_generic_to_shared(tile))), "r"(run * unsigned(sizeof(T)))
                           : "memory");
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            __syncwarp();
            continue;
          }
          __syncwarp();
          if (pitch == N)
          {
#pragma unroll 4
            for (unsigned q = lane; q < run; q += 32)
              __stcs(g + q, tile[q]);
          }
          else
          {
            // padded rows (even N): row by row
            for (unsigned r = 0; r < nrows; ++r)
              for (int n = lane; n < N; n += 32)
                __stcs(g + size_t(r) * N + n, tile[r * pitch + n]);
          }
          __syncwarp();
        }
  }
}

template <int TDIM, int ND, int NF, typename T>
void launch(const bx_element& e, const T* x, size_t npts, T* basis, cudaStream_t s)
{
  Params<T> p;
  p.x = x;
  p.npts = npts;
  p.basis = basis;
  p.A = e.tp.d_A;
  p.order = e.tp.d_order;
  p.N = e.N;
  p.pitch = (e.N & 1) ? e.N : e.N + 1;
  p.aligned16 = (reinterpret_cast<uintptr_t>(basis) % 16 == 0 and (npts * size_t(e.N) * sizeof(T)) % 16 == 0) ? 1 : 0;
  const size_t fixed = (size_t(2) * e.N + ((TDIM * NF * NF + 1) & ~1)) * sizeof(double);
  const size_t per_warp = size_t(32) * p.pitch * sizeof(T);
  p.warps = warps_for<TDIM, NF, T>();
  const size_t smem = fixed + size_t(p.warps) * per_warp;
  auto kern = tabulate_tp_staged_kernel<TDIM, ND, NF, T>;
  BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const size_t ntiles = ceil_div(npts, size_t(32));
  const size_t want = ceil_div(ntiles, size_t(p.warps));
  kern<<<unsigned(std::min<size_t>(want, size_t(sm_count()))), 32 * p.warps, smem, s>>>(p);
  BX_LAUNCHED();
}

#define BX_STAGED(TD, ND_, NF_)                                                                          \
  if (e.tdim == TD and nd == ND_ and e.tp.nf == NF_)                                                      \
  {                                                                                                      \
    if (!dry_run)                                                                                        \
      launch<TD, ND_, NF_, T>(e, x, npts, basis, s);                                                        \
    return true;                                                                                         \
  }

// Returns false when the element / derivative order is outside this kernel's menu (caller uses tabulate_tp.cu).
template <typename T>
bool launch_any(const bx_element& e, int nd, const T* x, size_t npts, T* basis, cudaStream_t s, bool dry_run)
{
  if (!e.tp.valid or e.tp.order.empty() or e.tp.nf != e.superdegree + 1 or e.tp.fstride != e.tp.nf)
    return false;
  if (!dry_run and !e.tp.d_order)
    return false;
  // at least four warps must fit beside each other (tile = 32 x N doubles per warp)
  if (size_t(32) * (e.N | 1) * sizeof(T) * 4 > size_t(196) * 1024)
    return false;
  BX_STAGED(2, 0, 2) BX_STAGED(2, 0, 3) BX_STAGED(2, 0, 4) BX_STAGED(2, 0, 5) BX_STAGED(2, 0, 6) BX_STAGED(2, 0, 7)
  BX_STAGED(2, 1, 2) BX_STAGED(2, 1, 3) BX_STAGED(2, 1, 4) BX_STAGED(2, 1, 5) BX_STAGED(2, 1, 6) BX_STAGED(2, 1, 7)
  BX_STAGED(3, 0, 2) BX_STAGED(3, 0, 3) BX_STAGED(3, 0, 4) BX_STAGED(3, 0, 5)
  BX_STAGED(3, 1, 2) BX_STAGED(3, 1, 3) BX_STAGED(3, 1, 4) BX_STAGED(3, 1, 5)
  BX_STAGED(2, 2, 2) BX_STAGED(2, 2, 3) BX_STAGED(2, 2, 4) BX_STAGED(2, 2, 5) BX_STAGED(2, 2, 6) BX_STAGED(2, 2, 7)
  BX_STAGED(3, 2, 2) BX_STAGED(3, 2, 3) BX_STAGED(3, 2, 4) BX_STAGED(3, 2, 5)
  return false;
}
template bool launch_any<double>(const bx_element&, int, const double*, size_t, double*, cudaStream_t, bool);
template bool launch_any<float>(const bx_element&, int, const float*, size_t, float*, cudaStream_t, bool);
} // namespace staged
} // namespace bx

// Generic polynomial-set tabulation kernels: one point per thread, table written straight to
// global memory as P[deriv][basis][point] (point index fastest => every store and every
// re-load of a lower-order entry is coalesced across the warp).  Any degree / derivative order.
//
// Replaces polyset::tabulate, reference cpp/basix/polyset.cpp:2914-2978.
// Compiled with -fmad=false (see polyset.cuh).
#include "polyset.cuh"

namespace bx
{
namespace
{
constexpr int kThreads = 128;
constexpr int kMaxTPDegree = 20; // quad/hex generic path: per-thread 1-D tables
constexpr int kMaxTPDeriv = 5;

template <typename T>
__global__ void __launch_bounds__(kThreads)
    polyset_simplex_kernel(int cell, int n, int nd, int psize, const T* __restrict__ x, size_t npts, T* P)
{
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < npts; i += stride)
  {
    PointView<T> v{P + i, npts, psize};
    if (cell == BX_CELL_INTERVAL)
      polyset_interval<T>(v, n, nd, x[i]);
    else if (cell == BX_CELL_TRIANGLE)
      polyset_triangle<T>(v, n, nd, x[2 * i], x[2 * i + 1]);
    else if (cell == BX_CELL_TETRAHEDRON)
      polyset_tetrahedron<T>(v, n, nd, x[3 * i], x[3 * i + 1], x[3 * i + 2]);
    else if (cell == BX_CELL_PRISM)
      polyset_prism<T>(v, n, nd, x[3 * i], x[3 * i + 1], x[3 * i + 2]);
    else
      polyset_pyramid<T>(v, n, nd, x[3 * i], x[3 * i + 1], x[3 * i + 2]);
  }
}

// point cell: value 1, derivatives 0 (reference polyset.cpp:45-58)
template <typename T>
__global__ void polyset_point_kernel(int nder, size_t npts, T* P)
{
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < npts * nder; i += stride)
    P[i] = i < npts ? T(1.0) : T(0.0);
}

// quad / hex: products of per-axis normalised Legendre tables.
// basis slot: quad (n+1) px + py; hex (n+1)^2 px + (n+1) py + pz  (polyset.cpp:2252, 2420-2422)
template <typename T, int TDIM>
__global__ void __launch_bounds__(kThreads)
    polyset_tp_kernel(int n, int nd, const T* __restrict__ x, size_t npts, T* P)
{
  T L[TDIM][(kMaxTPDeriv + 1) * (kMaxTPDegree + 1)];
  const int n1 = n + 1;
  const int psize = TDIM == 2 ? n1 * n1 : n1 * n1 * n1;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < npts; i += stride)
  {
    for (int a = 0; a < TDIM; ++a)
      legendre_table<T>(L[a], n, nd, x[TDIM * i + a]);
    T* out = P + i;
    if constexpr (TDIM == 2)
    {
      for (int kx = 0; kx <= nd; ++kx)
        for (int ky = 0; ky <= nd - kx; ++ky)
        {
          T* o = out + size_t(idx2(kx, ky)) * psize * npts;
          for (int px = 0; px < n1; ++px)
          {
            const T lx = L[0][kx * n1 + px];
            for (int py = 0; py < n1; ++py)
              o[size_t(px * n1 + py) * npts] = lx * L[1][ky * n1 + py];
          }
        }
    }
    else
    {
      for (int kx = 0; kx <= nd; ++kx)
        for (int ky = 0; ky <= nd - kx; ++ky)
          for (int kz = 0; kz <= nd - kx - ky; ++kz)
          {
            T* o = out + size_t(idx3(kx, ky, kz)) * psize * npts;
            for (int px = 0; px < n1; ++px)
              for (int py = 0; py < n1; ++py)
              {
                const T lxy = L[0][kx * n1 + px] * L[1][ky * n1 + py];
                for (int pz = 0; pz < n1; ++pz)
                  o[size_t((px * n1 + py) * n1 + pz) * npts] = lxy * L[2][kz * n1 + pz];
              }
          }
    }
  }
}

int grid_for(size_t n, int threads, int per_sm)
{
  const size_t want = ceil_div(n, size_t(threads));
  const size_t cap = size_t(sm_count()) * per_sm;
  return static_cast<int>(want < cap ? (want == 0 ? 1 : want) : cap);
}
} // namespace

template <typename T>
void polyset_macroedge_device(int cell, int degree, int nd, const T* x, size_t npts, T* P, cudaStream_t s);

// Device-pointer entry: x[npts][tdim] and P[nderivs][psize][npts] on the device.
template <typename T>
void polyset_tabulate_device(int cell, int ptype, int degree, int nd, const T* x, size_t npts, T* P,
                             cudaStream_t stream)
{
  if (degree < 0 or nd < 0)
    fail(BX_ERR_INVALID_ARG, "Polynomial set: negative degree or derivative order");
  if (ptype == BX_POLYSET_MACROEDGE and cell != BX_CELL_POINT)
  {
    polyset_macroedge_device<T>(cell, degree, nd, x, npts, P, stream);
    return;
  }
  if (ptype != BX_POLYSET_STANDARD and ptype != BX_POLYSET_MACROEDGE)
    fail(BX_ERR_INVALID_ARG, "Polynomial set: unknown polyset type");
  if (npts == 0)
    return;
  const int psize = polyset_dim(cell, ptype, degree);
  switch (cell)
  {
  case BX_CELL_POINT:
  {
    const int nder = polyset_nderivs(cell, nd);
    polyset_point_kernel<T><<<grid_for(npts * nder, 256, 8), 256, 0, stream>>>(nder, npts, P);
    BX_LAUNCHED();
    return;
  }
  case BX_CELL_INTERVAL:
  case BX_CELL_TRIANGLE:
  case BX_CELL_TETRAHEDRON:
  case BX_CELL_PRISM:
  case BX_CELL_PYRAMID:
    polyset_simplex_kernel<T>
        <<<grid_for(npts, kThreads, 16), kThreads, 0, stream>>>(cell, degree, nd, psize, x, npts, P);
    BX_LAUNCHED();
    return;
  case BX_CELL_QUADRILATERAL:
  case BX_CELL_HEXAHEDRON:
    if (degree > kMaxTPDegree or nd > kMaxTPDeriv)
      fail(BX_ERR_UNSUPPORTED, "Polynomial set: quad/hex device path supports degree <= 20 and nderiv <= 5");
    if (cell == BX_CELL_QUADRILATERAL)
      polyset_tp_kernel<T, 2><<<grid_for(npts, kThreads, 8), kThreads, 0, stream>>>(degree, nd, x, npts, P);
    else
      polyset_tp_kernel<T, 3><<<grid_for(npts, kThreads, 8), kThreads, 0, stream>>>(degree, nd, x, npts, P);
    BX_LAUNCHED();
    return;
  default:
    fail(BX_ERR_INVALID_ARG, "Polynomial set: unsupported cell type");
  }
}

template void polyset_tabulate_device<double>(int, int, int, int, const double*, size_t, double*,
                                              cudaStream_t);
template void polyset_tabulate_device<float>(int, int, int, int, const float*, size_t, float*, cudaStream_t);
} // namespace bx

// Macroedge polynomial sets (polyset::type::macroedge): piecewise polynomials on the cell split by halving every edge.
//
// Replaces the macroedge arms of polyset::tabulate (reference cpp/basix/polyset.cpp:126-1591, dispatch 2914-2978):
//   interval / quadrilateral / hexahedron   any degree: products of the 1-D piecewise functions (polyset.cpp:126-438,
//                                           1263-1591), restated here operation for operation
//   triangle (degree <= 2), tetrahedron (degree <= 1)   polynomials per sub-cell (polyset.cpp:440-1261), evaluated from
//                                           the coefficient tables of macroedge_tables.inc (see tools/macroedge_tables.py)
// "Identical to the reference" includes the reference's derivative slabs as they are: the 1-D code differentiates the
// two factors of a product separately and multiplies by `-i` for an unsigned i (= 2^64 - i), so slabs of order >= 1 of
// the interval / quadrilateral / hexahedron sets hold values of magnitude 1e19 where the reference holds them.
// One point per thread; compiled with -fmad=false.
#include "common.cuh"

#include <cmath>
#include <vector>

namespace bx
{
namespace
{
constexpr int kThreads = 128;
constexpr int kMaxN = 10;  // degree (2 n + 1 <= 21 functions per axis)
constexpr int kMaxND = 3;  // derivative order

// the reference's `factorials` vectors (polyset.cpp:142-146, 178-184), computed on the host in its expression order
struct MacroFactors
{
  double f0[kMaxN + 1];
  double fj[kMaxN][kMaxN + 1];
};

int single_choose(int n, int k) // polyset.cpp:20-28 (int arithmetic there; degrees <= kMaxN stay far from overflow)
{
  long long out = 1;
  for (int i = n + 1 - k; i <= n; ++i)
    out *= i;
  for (int i = 1; i <= k; ++i)
    out /= i;
  return int(out);
}

template <typename T>
MacroFactors make_factors(int n)
{
  MacroFactors m{};
  for (int k = 0; k <= n; ++k)
    m.f0[k] = double(T((k % 2 == 0 ? 1 : -1) * single_choose(2 * n + 1 - k, n - k) * single_choose(n, k) * std::pow(2, n - k)));
  for (int j = 0; j < n; ++j)
    for (int k = 0; k <= j; ++k)
      m.fj[j][k] = double(T((k % 2 == 0 ? 1 : -1) * single_choose(2 * n + 1 - k, j - k) * single_choose(j, k) * std::pow(2, j - k)
                            * std::pow(2, n - j) * std::sqrt(4 * (n - j) + 2)));
  return m;
}

// T(-i) for an unsigned 64-bit i, as `x_term *= -i` evaluates it
template <typename T>
__device__ __forceinline__ T neg_unsigned(int i)
{
  return T(0ull - (unsigned long long)(i));
}

// F[d * (2 n + 1) + j], d <= nd: the 1-D functions and their "derivatives" at t (polyset.cpp:147-224)
template <typename T>
__device__ void macro_1d(T t, int n, int nd, const MacroFactors& mf, T* F)
{
  const int m = 2 * n + 1;
  for (int i = 0; i < (nd + 1) * m; ++i)
    F[i] = T(0);
  const bool left = t <= T(0.5);
  const double base = left ? double(t) : 1.0 - double(t);
  for (int d = 0; d <= nd; ++d)
  {
    T value = T(0);
    for (int k = 0; k + d <= n; ++k)
    {
      T x_term = T(pow(base, double(n - k - d)));
      for (int i = n - k; i > n - k - d; --i)
        x_term *= left ? T(i) : neg_unsigned<T>(i);
      value += T(mf.f0[k]) * x_term;
    }
    F[d * m] = value;
  }
  for (int j = 0; j < n; ++j)
    for (int d = 0; d <= nd; ++d)
    {
      T value = T(0);
      for (int k = 0; k + d <= j; ++k)
      {
        T x_term = T(pow(base, double(j - k - d)));
        for (int i = j - k; i > j - k - d; --i)
          x_term *= left ? T(i) : neg_unsigned<T>(i);
        value += T(mf.fj[j][k]) * x_term;
      }
      // exponent n - j - d in unsigned arithmetic: wraps to 2^64 - (d - (n - j)) when d > n - j, and then the loop
      // over i below does not run
      const bool wrapped = d > n - j;
      const double e = wrapped ? double(0ull - (unsigned long long)(d - (n - j))) : double(n - j - d);
      if (left)
      {
        value *= T(pow(0.5 - double(t), e));
        if (!wrapped)
          for (int i = n - j; i > n - j - d; --i)
            value *= neg_unsigned<T>(i);
        F[d * m + j + 1] = value;
      }
      else
      {
        value *= T(pow(double(t) - 0.5, e));
        if (!wrapped)
          for (int i = n - j; i > n - j - d; --i)
            value *= T(i);
        F[d * m + j + n + 1] = value;
      }
    }
}

template <typename T, int TDIM>
__global__ void __launch_bounds__(kThreads)
    macroedge_tp_kernel(int n, int nd, const MacroFactors mf, const T* __restrict__ x, size_t npts, T* __restrict__ P)
{
  T F[TDIM][(kMaxND + 1) * (2 * kMaxN + 1)];
  const int m = 2 * n + 1;
  const size_t psize = TDIM == 1 ? m : TDIM == 2 ? m * m : m * m * m;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t p = size_t(blockIdx.x) * blockDim.x + threadIdx.x; p < npts; p += stride)
  {
    for (int a = 0; a < TDIM; ++a)
      macro_1d<T>(x[TDIM * p + a], n, nd, mf, F[a]);
    T* out = P + p;
    if constexpr (TDIM == 1)
    {
      for (int d = 0; d <= nd; ++d)
        for (int j = 0; j < m; ++j)
          out[(size_t(d) * psize + j) * npts] = F[0][d * m + j];
    }
    else if constexpr (TDIM == 2)
    {
      for (int dx = 0; dx <= nd; ++dx)
        for (int dy = 0; dy <= nd - dx; ++dy)
        {
          T* o = out + size_t(idx2(dx, dy)) * psize * npts;
          for (int px = 0; px < m; ++px)
            for (int py = 0; py < m; ++py)
              o[size_t(px * m + py) * npts] = F[0][dx * m + px] * F[1][dy * m + py];
        }
    }
    else
    {
      for (int dx = 0; dx <= nd; ++dx)
        for (int dy = 0; dy <= nd - dx; ++dy)
          for (int dz = 0; dz <= nd - dx - dy; ++dz)
          {
            T* o = out + size_t(idx3(dx, dy, dz)) * psize * npts;
            for (int px = 0; px < m; ++px)
              for (int py = 0; py < m; ++py)
              {
                const T xy = F[0][dx * m + px] * F[1][dy * m + py];
                for (int pz = 0; pz < m; ++pz)
                  o[size_t((px * m + py) * m + pz) * npts] = xy * F[2][dz * m + pz];
              }
          }
    }
  }
}

// ---- triangle / tetrahedron: polynomials per sub-cell from coefficient tables ------------------------------------
#include "macroedge_tables.inc"

// Sub-cell of a point, tested in the reference's order (a point on an inner edge belongs to the first match).
__device__ __forceinline__ int triangle_region(double x0, double x1)
{
  if (x0 + x1 < 0.5)
    return 0;
  if (x0 > 0.5)
    return 1;
  if (x1 > 0.5)
    return 2;
  return 3;
}
__device__ __forceinline__ int tetrahedron_region(double x0, double x1, double x2)
{
  if (x0 + x1 + x2 < 0.5)
    return 0;
  if (x0 > 0.5)
    return 1;
  if (x1 > 0.5)
    return 2;
  if (x2 > 0.5)
    return 3;
  if (x1 + x2 < 0.5 and x0 + x1 < 0.5)
    return 4;
  if (x1 + x2 < 0.5)
    return 5;
  if (x0 + x1 > 0.5)
    return 6;
  return 7;
}

// The reference evaluates these sets in double whatever the table type (its literals are doubles) and narrows on
// assignment; sums run left to right over the monomials 1, y, x, y y, x y, x x, each product as (c * a) * b.
template <typename T>
__global__ void __launch_bounds__(kThreads)
    macroedge_triangle_kernel(int n, int nd, const T* __restrict__ xs, size_t npts, T* __restrict__ P)
{
  const int psize = (n + 1) * (2 * n + 1);
  const int nder = (nd + 1) * (nd + 2) / 2;
  const int nmono = n == 1 ? 3 : 6;
  const double* tab = n == 1 ? kMacroTri1 : kMacroTri2;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t p = size_t(blockIdx.x) * blockDim.x + threadIdx.x; p < npts; p += stride)
  {
    const double x = double(xs[2 * p]), y = double(xs[2 * p + 1]);
    for (int d = 0; d < nder; ++d)
      for (int f = 0; f < psize; ++f)
        P[(size_t(d) * psize + f) * npts + p] = T(0);
    if (n == 0)
    {
      P[p] = T(sqrt(2.0));
      continue;
    }
    const int r = triangle_region(x, y);
    for (int f = 0; f < psize; ++f)
    {
      const double* c = tab + (size_t(r) * psize + f) * nmono;
      T* o = P + size_t(f) * npts + p;
      const size_t slab = size_t(psize) * npts;
      if (n == 1)
      {
        o[0] = T(c[0] + c[1] * y + c[2] * x);
        if (nd >= 1)
        {
          o[idx2(0, 1) * slab] = T(c[1]);
          o[idx2(1, 0) * slab] = T(c[2]);
        }
        continue;
      }
      o[0] = T(c[0] + c[1] * y + c[2] * x + c[3] * y * y + c[4] * x * y + c[5] * x * x);
      if (nd >= 1)
      {
        o[idx2(0, 1) * slab] = T(c[1] + 2.0 * c[3] * y + c[4] * x);
        o[idx2(1, 0) * slab] = T(c[2] + c[4] * y + 2.0 * c[5] * x);
      }
      if (nd >= 2)
      {
        o[idx2(0, 2) * slab] = T(2.0 * c[3]);
        o[idx2(1, 1) * slab] = T(c[4]);
        o[idx2(2, 0) * slab] = T(2.0 * c[5]);
      }
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    macroedge_tetrahedron_kernel(int n, int nd, const T* __restrict__ xs, size_t npts, T* __restrict__ P)
{
  const int psize = n == 0 ? 1 : 10;
  const int nder = (nd + 1) * (nd + 2) * (nd + 3) / 6;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t p = size_t(blockIdx.x) * blockDim.x + threadIdx.x; p < npts; p += stride)
  {
    const double x[3] = {double(xs[3 * p]), double(xs[3 * p + 1]), double(xs[3 * p + 2])};
    for (int d = 0; d < nder; ++d)
      for (int f = 0; f < psize; ++f)
        P[(size_t(d) * psize + f) * npts + p] = T(0);
    if (n == 0)
    {
      P[p] = T(sqrt(6.0));
      continue;
    }
    const int r = tetrahedron_region(x[0], x[1], x[2]);
    const size_t slab = size_t(psize) * npts;
    for (int f = 0; f < psize; ++f)
    {
      const double* c = kMacroTet1 + (size_t(r) * psize + f) * 4; // 1, x, y, z
      T* o = P + size_t(f) * npts + p;
      double acc = c[0];
#pragma unroll
      for (int k = 0; k < 3; ++k)
      {
        const int a = kMacroTetOrder[k];
        acc += c[1 + a] * x[a];
      }
      o[0] = T(acc);
      if (nd >= 1)
      {
        o[idx3(1, 0, 0) * slab] = T(c[1]);
        o[idx3(0, 1, 0) * slab] = T(c[2]);
        o[idx3(0, 0, 1) * slab] = T(c[3]);
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
void polyset_macroedge_device(int cell, int degree, int nd, const T* x, size_t npts, T* P, cudaStream_t s)
{
  if (npts == 0)
    return;
  switch (cell)
  {
  case BX_CELL_INTERVAL:
  case BX_CELL_QUADRILATERAL:
  case BX_CELL_HEXAHEDRON:
  {
    if (degree > kMaxN or nd > kMaxND)
      fail(BX_ERR_UNSUPPORTED, "Polynomial set: the macroedge device path supports degree <= 10 and nderiv <= 3");
    const MacroFactors mf = make_factors<T>(degree);
    const int grid = grid_for(npts, kThreads, 8);
    if (cell == BX_CELL_INTERVAL)
      macroedge_tp_kernel<T, 1><<<grid, kThreads, 0, s>>>(degree, nd, mf, x, npts, P);
    else if (cell == BX_CELL_QUADRILATERAL)
      macroedge_tp_kernel<T, 2><<<grid, kThreads, 0, s>>>(degree, nd, mf, x, npts, P);
    else
      macroedge_tp_kernel<T, 3><<<grid, kThreads, 0, s>>>(degree, nd, mf, x, npts, P);
    BX_LAUNCHED();
    return;
  }
  case BX_CELL_TRIANGLE:
    // the reference's own limits and messages (polyset.cpp:921-925, 1254-1258)
    if (degree > 2)
      fail(BX_ERR_RUNTIME, "Only degree 0 to 2 macro polysets are currently implemented on a triangle.");
    macroedge_triangle_kernel<T><<<grid_for(npts, kThreads, 8), kThreads, 0, s>>>(degree, nd, x, npts, P);
    BX_LAUNCHED();
    return;
  case BX_CELL_TETRAHEDRON:
    if (degree > 1)
      fail(BX_ERR_RUNTIME, "Only degree 0 and 1 macro polysets are currently implemented on a tetrahedron.");
    macroedge_tetrahedron_kernel<T><<<grid_for(npts, kThreads, 8), kThreads, 0, s>>>(degree, nd, x, npts, P);
    BX_LAUNCHED();
    return;
  case BX_CELL_POINT:
  {
    // a point has no edges: the standard set (value 1)
    fail(BX_ERR_UNSUPPORTED, "Polynomial set: macroedge polyset of a point");
  }
  default:
    fail(BX_ERR_RUNTIME, "Polynomial set: unsupported polyset type on this cell type");
  }
}
template void polyset_macroedge_device<double>(int, int, int, const double*, size_t, double*, cudaStream_t);
template void polyset_macroedge_device<float>(int, int, int, const float*, size_t, float*, cudaStream_t);
} // namespace bx

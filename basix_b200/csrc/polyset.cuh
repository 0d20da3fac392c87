// Orthonormal polynomial-set recurrences, one point per thread (device side).
//
// What is computed follows the reference's polyset::tabulate (cpp/basix/polyset.cpp):
//   interval     tabulate_polyset_line_derivs         polyset.cpp:66-120
//   triangle     tabulate_polyset_triangle_derivs     polyset.cpp:1601-1743
//   tetrahedron  tabulate_polyset_tetrahedron_derivs  polyset.cpp:1745-2020
//   quad / hex   tensor products of the interval set  polyset.cpp:2241-2702
//   prism        tabulate_polyset_prism_derivs        polyset.cpp:2704-2911
//   pyramid      tabulate_polyset_pyramid_derivs      polyset.cpp:2022-2239
// How it is computed is different: each thread owns one point and walks the three-term
// recurrences through a strided view `P(d, k)` whose backing store is either global memory
// (generic kernel: element stride = npts, so a warp's accesses are coalesced) or a shared
// memory tile (fused kernels).  The arithmetic of every entry is written in the same
// operation order as the reference and this file is compiled with -fmad=false, so the
// simplex/interval tables agree with the reference (built without FMA contraction) to the
// last bit; quad/hex use the product form and agree to rounding.
#pragma once

#include "common.cuh"

namespace bx
{
// View of one point's column of the table P[deriv][basis][point].
template <typename T>
struct PointView
{
  T* p;       // address of P[0][0][this point]
  size_t ld;  // distance between consecutive basis functions (= #points in the table)
  int psize;  // basis functions per derivative
  __host__ __device__ __forceinline__ T& operator()(int d, int k) const { return p[(size_t(d) * psize + k) * ld]; }
};

// Jacobi recurrence coefficients, reference polyset.cpp:31-41 (jrc).
template <typename T>
__host__ __device__ __forceinline__ void jacobi_abc(int a, int n, T& an, T& bn, T& cn)
{
  an = (a + 2 * n + 1) * (a + 2 * n + 2) / static_cast<T>(2 * (n + 1) * (a + n + 1));
  bn = a * a * (a + 2 * n + 1) / static_cast<T>(2 * (n + 1) * (a + n + 1) * (a + 2 * n));
  cn = n * (a + n) * (a + 2 * n + 2) / static_cast<T>((n + 1) * (a + n + 1) * (a + 2 * n));
}

// ---- interval: shifted Legendre on [0,1] --------------------------------------------------------
template <typename T, typename View>
__host__ __device__ void polyset_interval(View P, int n, int nd, T x0)
{
  const T X = x0 * T(2.0) - T(1.0);
  P(0, 0) = T(1.0);
  for (int k = 1; k <= nd; ++k)
    P(k, 0) = T(0.0);
  if (n > 0)
  {
    P(0, 1) = X * P(0, 0);
    for (int p = 2; p <= n; ++p)
    {
      const T a = T(1.0) - T(1.0) / static_cast<T>(p);
      P(0, p) = X * P(0, p - 1) * (a + T(1.0)) - P(0, p - 2) * a;
    }
    for (int k = 1; k <= nd; ++k)
    {
      // p = 1: the P_{p-2} term has coefficient a = 0
      P(k, 1) = X * P(k, 0) * T(1.0) + static_cast<T>(2 * k) * P(k - 1, 0) * T(1.0);
      for (int p = 2; p <= n; ++p)
      {
        const T a = T(1.0) - T(1.0) / static_cast<T>(p);
        P(k, p) = X * P(k, p - 1) * (a + T(1.0)) + static_cast<T>(2 * k) * P(k - 1, p - 1) * (a + T(1.0))
                  - P(k, p - 2) * a;
      }
    }
  }
  for (int p = 0; p <= n; ++p)
  {
    const T norm = static_cast<T>(sqrt(static_cast<double>(2 * p + 1)));
    for (int k = 0; k <= nd; ++k)
      P(k, p) *= norm;
  }
}

// ---- triangle: Dubiner / Sherwin-Karniadakis --------------------------------------------------
// basis slot idx2(q, p); derivative slot idx2(kx, ky)
template <typename T, typename View>
__host__ __device__ void polyset_triangle(View P, int n, int nd, T x0, T x1)
{
  const int nder = (nd + 1) * (nd + 2) / 2;
  if (n == 0)
  {
    P(0, 0) = static_cast<T>(sqrt(2.0));
    for (int d = 1; d < nder; ++d)
      P(d, 0) = T(0.0);
    return;
  }
  const T X = x0 * T(2.0) - T(1.0);
  const T Y = x1 * T(2.0) - T(1.0);
  const T lin = X + T(0.5) * Y + T(0.5);
  const T f3 = T(1.0) - x1;
  for (int kx = 0; kx <= nd; ++kx)
  {
    for (int ky = 0; ky <= nd - kx; ++ky)
    {
      const int d = idx2(kx, ky);
      P(d, 0) = (d == 0) ? T(1.0) : T(0.0);
      // q = 0 row: recurrence in p
      for (int p = 1; p <= n; ++p)
      {
        const T a = static_cast<T>(2 * p - 1) / static_cast<T>(p);
        T v = lin * P(d, idx2(0, p - 1)) * a;
        if (kx > 0)
          v += static_cast<T>(2 * kx) * a * P(idx2(kx - 1, ky), idx2(0, p - 1));
        if (ky > 0)
          v += static_cast<T>(ky) * a * P(idx2(kx, ky - 1), idx2(0, p - 1));
        if (p > 1)
        {
          v -= f3 * f3 * P(d, idx2(0, p - 2)) * (a - T(1.0));
          if (ky > 0)
            v -= static_cast<T>(ky) * (Y - T(1.0)) * P(idx2(kx, ky - 1), idx2(0, p - 2)) * (a - T(1.0));
          if (ky > 1)
            v -= static_cast<T>(ky * (ky - 1)) * P(idx2(kx, ky - 2), idx2(0, p - 2)) * (a - T(1.0));
        }
        P(d, idx2(0, p)) = v;
      }
      // extend in q
      for (int p = 0; p < n; ++p)
      {
        T v = P(d, idx2(0, p)) * (Y * (T(1.5) + p) + T(0.5) + p);
        if (ky > 0)
          v += static_cast<T>(2 * ky) * (T(1.5) + p) * P(idx2(kx, ky - 1), idx2(0, p));
        P(d, idx2(1, p)) = v;
        for (int q = 1; q < n - p; ++q)
        {
          T a1, a2, a3;
          jacobi_abc<T>(2 * p + 1, q, a1, a2, a3);
          T w = P(d, idx2(q, p)) * (Y * a1 + a2) - P(d, idx2(q - 1, p)) * a3;
          if (ky > 0)
            w += static_cast<T>(2 * ky) * a1 * P(idx2(kx, ky - 1), idx2(q, p));
          P(d, idx2(q + 1, p)) = w;
        }
      }
    }
  }
  for (int p = 0; p <= n; ++p)
    for (int q = 0; q <= n - p; ++q)
    {
      const T norm = static_cast<T>(sqrt((p + 0.5) * (p + q + 1)) * 2);
      const int j = idx2(q, p);
      for (int d = 0; d < nder; ++d)
        P(d, j) *= norm;
    }
}

// ---- tetrahedron ------------------------------------------------------------------------------
// basis slot idx3(r, q, p); derivative slot idx3(kx, ky, kz)
template <typename T, typename View>
__host__ __device__ void polyset_tetrahedron(View P, int n, int nd, T x0, T x1, T x2)
{
  const int nder = (nd + 1) * (nd + 2) * (nd + 3) / 6;
  if (n == 0)
  {
    P(0, 0) = static_cast<T>(sqrt(6.0));
    for (int d = 1; d < nder; ++d)
      P(d, 0) = T(0.0);
    return;
  }
  const T X = x0 * T(2.0) - T(1.0);
  const T Y = x1 * T(2.0) - T(1.0);
  const T Z = x2 * T(2.0) - T(1.0);
  const T lin = X + T(0.5) * (Y + Z) + T(1.0);
  const T f2 = x1 + x2 - T(1.0);
  const T YZ = Y + Z;
  const T f4 = T(1.0) - x2;
  const T f3 = x1 * T(2.0) - T(1.0) + x2;
  for (int kx = 0; kx <= nd; ++kx)
  {
    for (int ky = 0; ky <= nd - kx; ++ky)
    {
      for (int kz = 0; kz <= nd - kx - ky; ++kz)
      {
        const int d = idx3(kx, ky, kz);
        P(d, 0) = (d == 0) ? T(1.0) : T(0.0);
        // (q, r) = (0, 0): recurrence in p
        for (int p = 1; p <= n; ++p)
        {
          const T a = static_cast<T>(2 * p - 1) / static_cast<T>(p);
          const int k1 = idx3(0, 0, p - 1);
          T v = lin * a * P(d, k1);
          if (kx > 0)
            v += static_cast<T>(2 * kx) * a * P(idx3(kx - 1, ky, kz), k1);
          if (ky > 0)
            v += static_cast<T>(ky) * a * P(idx3(kx, ky - 1, kz), k1);
          if (kz > 0)
            v += static_cast<T>(kz) * a * P(idx3(kx, ky, kz - 1), k1);
          if (p > 1)
          {
            const int k2 = idx3(0, 0, p - 2);
            v -= f2 * f2 * P(d, k2) * (a - T(1.0));
            if (ky > 0)
              v -= static_cast<T>(ky) * YZ * P(idx3(kx, ky - 1, kz), k2) * (a - T(1.0));
            if (ky > 1)
              v -= static_cast<T>(ky * (ky - 1)) * P(idx3(kx, ky - 2, kz), k2) * (a - T(1.0));
            if (kz > 0)
              v -= static_cast<T>(kz) * YZ * P(idx3(kx, ky, kz - 1), k2) * (a - T(1.0));
            if (kz > 1)
              v -= static_cast<T>(kz * (kz - 1)) * P(idx3(kx, ky, kz - 2), k2) * (a - T(1.0));
            if (ky > 0 and kz > 0)
              v -= T(2.0) * static_cast<T>(ky) * static_cast<T>(kz) * P(idx3(kx, ky - 1, kz - 1), k2)
                   * (a - T(1.0));
          }
          P(d, idx3(0, 0, p)) = v;
        }
        // r = 0: recurrence in q
        for (int p = 0; p < n; ++p)
        {
          const int k0 = idx3(0, 0, p);
          T v = P(d, k0) * ((T(1.0) + Y) * static_cast<T>(p) + (T(2.0) + Y * T(3.0) + Z) * T(0.5));
          if (ky > 0)
            v += static_cast<T>(2 * ky) * P(idx3(kx, ky - 1, kz), k0) * (T(1.5) + p);
          if (kz > 0)
            v += static_cast<T>(kz) * P(idx3(kx, ky, kz - 1), k0);
          P(d, idx3(0, 1, p)) = v;
          for (int q = 1; q < n - p; ++q)
          {
            T aq, bq, cq;
            jacobi_abc<T>(2 * p + 1, q, aq, bq, cq);
            const int kq = idx3(0, q, p), kqm = idx3(0, q - 1, p);
            T w = P(d, kq) * (f3 * aq + f4 * bq) - P(d, kqm) * f4 * f4 * cq;
            if (ky > 0)
              w += static_cast<T>(2 * ky) * P(idx3(kx, ky - 1, kz), kq) * aq;
            if (kz > 0)
              w += static_cast<T>(kz) * P(idx3(kx, ky, kz - 1), kq) * (aq - bq)
                   + static_cast<T>(kz) * (T(1.0) - Z) * P(idx3(kx, ky, kz - 1), kqm) * cq;
            if (kz > 1)
              w -= static_cast<T>(kz * (kz - 1)) * P(idx3(kx, ky, kz - 2), kqm) * cq;
            P(d, idx3(0, q + 1, p)) = w;
          }
        }
        // r = 1
        for (int p = 0; p < n; ++p)
          for (int q = 0; q < n - p; ++q)
          {
            const int k0 = idx3(0, q, p);
            T v = P(d, k0) * ((T(1.0) + p + q) + Z * (T(2.0) + p + q));
            if (kz > 0)
              v += static_cast<T>(2 * kz) * (T(2.0) + p + q) * P(idx3(kx, ky, kz - 1), k0);
            P(d, idx3(1, q, p)) = v;
          }
        // recurrence in r
        for (int p = 0; p + 1 < n; ++p)
          for (int q = 0; q + 1 < n - p; ++q)
            for (int r = 1; r < n - p - q; ++r)
            {
              T ar, br, cr;
              jacobi_abc<T>(2 * p + 2 * q + 2, r, ar, br, cr);
              const int kr = idx3(r, q, p);
              T v = P(d, kr) * (Z * ar + br) - P(d, idx3(r - 1, q, p)) * cr;
              if (kz > 0)
                v += static_cast<T>(2 * kz) * ar * P(idx3(kx, ky, kz - 1), kr);
              P(d, idx3(r + 1, q, p)) = v;
            }
      }
    }
  }
  for (int p = 0; p <= n; ++p)
    for (int q = 0; q <= n - p; ++q)
      for (int r = 0; r <= n - p - q; ++r)
      {
        const T norm = static_cast<T>(sqrt(2 * (p + 0.5) * (p + q + 1.0) * (p + q + r + 1.5)) * 2);
        const int j = idx3(r, q, p);
        for (int d = 0; d < nder; ++d)
          P(d, j) *= norm;
      }
}

// ---- prism: triangle (x, y) set extended by the interval set in z ---------------------------------
// reference tabulate_polyset_prism_derivs, polyset.cpp:2704-2911
// basis slot (n+1) idx2(q, p) + r; derivative slot idx3(kx, ky, kz)
template <typename T, typename View>
__host__ __device__ void polyset_prism(View P, int n, int nd, T x0, T x1, T x2)
{
  const int nder = (nd + 1) * (nd + 2) * (nd + 3) / 6;
  const int n1 = n + 1;
  const int psize = n1 * n1 * (n + 2) / 2;
  for (int d = 0; d < nder; ++d)
    for (int k = 0; k < psize; ++k)
      P(d, k) = T(0.0);
  if (n == 0)
  {
    P(0, 0) = static_cast<T>(sqrt(2.0));
    return;
  }
  P(0, 0) = T(1.0);
  const T X = x0 * T(2.0) - T(1.0), Y = x1 * T(2.0) - T(1.0), Z = x2 * T(2.0) - T(1.0);
  const T lin = X + T(0.5) * Y + T(0.5);
  const T f2 = T(1.0) - x1;
  auto slot = [n1](int p, int q, int r) { return n1 * idx2(q, p) + r; };
  // triangle part (r = 0, kz = 0)
  for (int kx = 0; kx <= nd; ++kx)
    for (int ky = 0; ky <= nd - kx; ++ky)
    {
      const int d = idx3(kx, ky, 0);
      for (int p = 1; p <= n; ++p)
      {
        const T a = static_cast<T>(2 * p - 1) / static_cast<T>(p);
        T v = lin * P(d, slot(p - 1, 0, 0)) * a;
        if (kx > 0)
          v += static_cast<T>(2 * kx) * a * P(idx3(kx - 1, ky, 0), slot(p - 1, 0, 0));
        if (ky > 0)
          v += static_cast<T>(ky) * a * P(idx3(kx, ky - 1, 0), slot(p - 1, 0, 0));
        if (p > 1)
        {
          v -= f2 * f2 * P(d, slot(p - 2, 0, 0)) * (a - T(1.0));
          if (ky > 0)
            v -= static_cast<T>(ky) * (Y - T(1.0)) * P(idx3(kx, ky - 1, 0), slot(p - 2, 0, 0)) * (a - T(1.0));
          if (ky > 1)
            v -= static_cast<T>(ky * (ky - 1)) * P(idx3(kx, ky - 2, 0), slot(p - 2, 0, 0)) * (a - T(1.0));
        }
        P(d, slot(p, 0, 0)) = v;
      }
      for (int p = 0; p < n; ++p)
      {
        T v = P(d, slot(p, 0, 0)) * (Y * (T(1.5) + p) + T(0.5) + p);
        if (ky > 0)
          v += static_cast<T>(2 * ky) * (T(1.5) + p) * P(idx3(kx, ky - 1, 0), slot(p, 0, 0));
        P(d, slot(p, 1, 0)) = v;
        for (int q = 1; q < n - p; ++q)
        {
          T a1, a2, a3;
          jacobi_abc<T>(2 * p + 1, q, a1, a2, a3);
          T w = P(d, slot(p, q, 0)) * (Y * a1 + a2) - P(d, slot(p, q - 1, 0)) * a3;
          if (ky > 0)
            w += static_cast<T>(2 * ky) * a1 * P(idx3(kx, ky - 1, 0), slot(p, q, 0));
          P(d, slot(p, q + 1, 0)) = w;
        }
      }
    }
  // interval recurrence in z on top of every triangle function
  for (int kz = 0; kz <= nd; ++kz)
    for (int r = 1; r <= n; ++r)
    {
      const T a = T(1.0) - T(1.0) / static_cast<T>(r);
      for (int kx = 0; kx <= nd - kz; ++kx)
        for (int ky = 0; ky <= nd - kx - kz; ++ky)
        {
          const int d = idx3(kx, ky, kz);
          for (int p = 0; p <= n; ++p)
            for (int q = 0; q <= n - p; ++q)
            {
              T v = Z * P(d, slot(p, q, r - 1)) * (a + T(1.0));
              if (kz > 0)
                v += static_cast<T>(2 * kz) * P(idx3(kx, ky, kz - 1), slot(p, q, r - 1)) * (a + T(1.0));
              if (r > 1)
                v -= P(d, slot(p, q, r - 2)) * a;
              P(d, slot(p, q, r)) = v;
            }
        }
    }
  for (int p = 0; p <= n; ++p)
    for (int q = 0; q <= n - p; ++q)
      for (int r = 0; r <= n; ++r)
      {
        const T norm = static_cast<T>(sqrt((p + 0.5) * (p + q + 1) * (2 * r + 1)) * 2);
        for (int d = 0; d < nder; ++d)
          P(d, slot(p, q, r)) *= norm;
      }
}

// ---- pyramid: rational recurrences in x / (1 - z), y / (1 - z), Jacobi (2, 0) in z ---------------------------
// reference tabulate_polyset_pyramid_derivs, polyset.cpp:2022-2239 (the apex z = 1 takes the same special
// values as there); basis slot (n+1)^2 p + (n+1) q + r; derivative slot idx3(kx, ky, kz)
template <typename T, typename View>
__host__ __device__ void polyset_pyramid(View P, int n, int nd, T x0, T x1, T x2)
{
  const int nder = (nd + 1) * (nd + 2) * (nd + 3) / 6;
  const int n1 = n + 1;
  const int psize = n1 * n1 * n1;
  for (int d = 0; d < nder; ++d)
    for (int k = 0; k < psize; ++k)
      P(d, k) = T(0.0);
  if (n == 0)
  {
    P(0, 0) = static_cast<T>(sqrt(3.0));
    return;
  }
  P(0, 0) = T(1.0);
  auto slot = [n1](int p, int q, int r) { return n1 * n1 * p + n1 * q + r; };
  const bool apex = x2 == T(1.0);
  const T x0over = apex ? T(0.0) : x0 / (T(1.0) - x2);
  const T x1over = apex ? T(0.0) : x1 / (T(1.0) - x2);
  const T over_z = apex ? T(1.0) : T(1.0) / (T(1.0) - x2);
  for (int kx = 0; kx <= nd; ++kx)
    for (int ky = 0; ky <= nd - kx; ++ky)
      for (int kz = 0; kz <= nd - kx - ky; ++kz)
      {
        const int d = idx3(kx, ky, kz);
        for (int p = 0; p <= n; ++p)
        {
          if (p > 0)
          {
            const T a = static_cast<T>(p - 1) / static_cast<T>(p);
            T v = (a + T(1.0)) * (T(2.0) * x0over - T(1.0)) * P(d, slot(p - 1, 0, 0));
            if (kx > 0)
              v += (a + T(1.0)) * T(2.0) * static_cast<T>(kx) * P(idx3(kx - 1, ky, kz), slot(p - 1, 0, 0)) * over_z;
            if (kz > 0)
              v += static_cast<T>(kz)
                   * ((a + T(1.0)) * P(idx3(kx, ky, kz - 1), slot(p - 1, 0, 0)) + P(idx3(kx, ky, kz - 1), slot(p, 0, 0)))
                   * over_z;
            if (p > 1)
            {
              v -= a * P(d, slot(p - 2, 0, 0));
              if (kz > 0)
                v += a * static_cast<T>(kz) * P(idx3(kx, ky, kz - 1), slot(p - 2, 0, 0)) * over_z;
            }
            P(d, slot(p, 0, 0)) = v;
          }
          for (int q = 1; q <= n; ++q)
          {
            const T a = static_cast<T>(q - 1) / static_cast<T>(q);
            T v = (a + T(1.0)) * (T(2.0) * x1over - T(1.0)) * P(d, slot(p, q - 1, 0));
            if (ky > 0)
              v += T(2.0) * (a + T(1.0)) * static_cast<T>(ky) * P(idx3(kx, ky - 1, kz), slot(p, q - 1, 0)) * over_z;
            if (kz > 0)
              v += static_cast<T>(kz)
                   * ((a + T(1.0)) * P(idx3(kx, ky, kz - 1), slot(p, q - 1, 0)) + P(idx3(kx, ky, kz - 1), slot(p, q, 0)))
                   * over_z;
            if (q > 1)
            {
              v -= a * P(d, slot(p, q - 2, 0));
              if (kz > 0)
                v += a * static_cast<T>(kz) * P(idx3(kx, ky, kz - 1), slot(p, q - 2, 0)) * over_z;
            }
            P(d, slot(p, q, 0)) = v;
          }
        }
        for (int p = 0; p <= n; ++p)
          for (int q = 0; q <= n; ++q)
            for (int r = 1; r <= n; ++r)
            {
              T ar, br, cr;
              jacobi_abc<T>(2, r - 1, ar, br, cr);
              T v = P(d, slot(p, q, r - 1)) * (T(2.0) * ar * x2 + br - ar);
              if (r > 1)
                v -= P(d, slot(p, q, r - 2)) * cr;
              if (kz > 0)
                v += static_cast<T>(2 * kz) * ar * P(idx3(kx, ky, kz - 1), slot(p, q, r - 1));
              P(d, slot(p, q, r)) = v;
            }
      }
  for (int p = 0; p <= n; ++p)
    for (int q = 0; q <= n; ++q)
      for (int r = 0; r <= n; ++r)
      {
        const T norm = static_cast<T>(sqrt(2 * (q + 0.5) * (p + 0.5) * (r + 1.5)) * 2);
        for (int d = 0; d < nder; ++d)
          P(d, slot(p, q, r)) *= norm;
      }
}

// ---- 1-D normalised Legendre values and derivatives into a small local table -----------------
// L[k * (n + 1) + p] = sqrt(2p+1) d^k/dx^k P_p(2x-1), k <= nd
template <typename T>
__host__ __device__ void legendre_table(T* L, int n, int nd, T x0)
{
  struct LocalView
  {
    T* p;
    int n1;
    __host__ __device__ __forceinline__ T& operator()(int k, int q) const { return p[k * n1 + q]; }
  };
  polyset_interval<T>(LocalView{L, n + 1}, n, nd, x0);
}
} // namespace bx

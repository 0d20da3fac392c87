// Register-resident "jet" recurrences for the orthonormal simplex polynomial sets.
//
// A jet is the vector of all partial derivatives (up to order ND) of one basis function at one
// point.  The reference (cpp/basix/polyset.cpp:66-120, 1601-1743, 1745-2020) sweeps the whole
// table derivative-by-derivative, re-reading lower derivatives from memory.  Here one thread owns
// one point and walks the (p, q, r) index tree depth first, carrying only two jets per tree level
// in registers; every finished jet is handed to a sink exactly once (`emit(seq, jet)`), in the
// fixed "production order"
//     interval     seq = p
//     triangle     for p: for q <= n-p                  -> basis slot idx2(q, p)
//     tetrahedron  for p: for q <= n-p: for r <= n-p-q  -> basis slot idx3(r, q, p)
// Jets are emitted UN-normalised: the norm of each basis function is folded into the coefficient
// matrix on the host (element.cu: fused_operand), which removes one multiply per table entry.
// Everything is meant to be fully unrolled (N, ND compile-time), so all recurrence coefficients
// (a_p, Jacobi A/B/C) are folded at compile time and all jet indices are register names.
#pragma once

#include "common.cuh"

namespace bx
{
namespace jets
{
__host__ __device__ constexpr double jrc_a(int a, int n)
{
  return double((a + 2 * n + 1) * (a + 2 * n + 2)) / double(2 * (n + 1) * (a + n + 1));
}
__host__ __device__ constexpr double jrc_b(int a, int n)
{
  return double(a * a * (a + 2 * n + 1)) / double(2 * (n + 1) * (a + n + 1) * (a + 2 * n));
}
__host__ __device__ constexpr double jrc_c(int a, int n)
{
  return double(n * (a + n) * (a + 2 * n + 2)) / double((n + 1) * (a + n + 1) * (a + 2 * n));
}

__host__ __device__ constexpr int nder(int tdim, int nd)
{
  return tdim == 1 ? nd + 1 : tdim == 2 ? (nd + 1) * (nd + 2) / 2 : (nd + 1) * (nd + 2) * (nd + 3) / 6;
}
__host__ __device__ constexpr int psize(int tdim, int n)
{
  return tdim == 1 ? n + 1 : tdim == 2 ? (n + 1) * (n + 2) / 2 : (n + 1) * (n + 2) * (n + 3) / 6;
}

// Ownership policy: lets several producer warps share one point tile by splitting the index tree at
// its first level.  `Own::owns(p)` says whether this instantiation emits the sub-tree rooted at (p,0,0);
// the (cheap) p chain itself is advanced by every instantiation up to its last owned p.
struct OwnAll
{
  static constexpr bool owns(int) { return true; }
};
// number of basis functions in the sub-tree rooted at first index p
__host__ __device__ constexpr int subtree_size(int tdim, int n, int p)
{
  return tdim == 1 ? 1 : tdim == 2 ? (n - p + 1) : (n - p + 1) * (n - p + 2) / 2;
}
// Greedy (largest first) split of the sub-trees p = 0..n over `nroles` producers.
__host__ __device__ constexpr int subtree_owner(int tdim, int n, int nroles, int p)
{
  int load[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int owner = 0;
  for (int q = 0; q <= n; ++q) // sub-tree sizes decrease with p, so p order is largest-first
  {
    int best = 0;
    for (int r = 1; r < nroles; ++r)
      if (load[r] < load[best])
        best = r;
    load[best] += subtree_size(tdim, n, q);
    if (q == p)
      owner = best;
  }
  return owner;
}
template <int TDIM, int N, int NROLES, int ROLE>
struct OwnSplit
{
  static constexpr bool owns(int p) { return TDIM == 1 ? ROLE == 0 : subtree_owner(TDIM, N, NROLES, p) == ROLE; }
};
template <typename Own>
__host__ __device__ constexpr int last_owned(int n)
{
  int last = -1;
  for (int p = 0; p <= n; ++p)
    if (Own::owns(p))
      last = p;
  return last;
}

// ---- interval -----------------------------------------------------------------------------------
template <int N, int ND, typename Own = OwnAll, typename Emit>
__device__ __forceinline__ void interval(double x0, Emit&& emit)
{
  if (!Own::owns(0))
    return;
  constexpr int NDER = ND + 1;
  const double X = x0 * 2.0 - 1.0;
  double J1[NDER], J2[NDER];
#pragma unroll
  for (int k = 0; k < NDER; ++k)
  {
    J1[k] = k == 0 ? 1.0 : 0.0;
    J2[k] = 0.0;
  }
  emit(0, J1);
#pragma unroll
  for (int p = 1; p <= N; ++p)
  {
    const double a = 1.0 - 1.0 / double(p);
    double Jn[NDER];
#pragma unroll
    for (int k = 0; k < NDER; ++k)
    {
      double v = (X * (a + 1.0)) * J1[k];
      if (k > 0)
        v += (2.0 * k * (a + 1.0)) * J1[k - 1];
      if (p > 1)
        v -= a * J2[k];
      Jn[k] = v;
    }
#pragma unroll
    for (int k = 0; k < NDER; ++k)
    {
      J2[k] = J1[k];
      J1[k] = Jn[k];
    }
    emit(p, J1);
  }
}

// ---- triangle -------------------------------------------------------------------------------------
template <int N, int ND, typename Own = OwnAll, typename Emit>
__device__ __forceinline__ void triangle(double x0, double x1, Emit&& emit)
{
  constexpr int LAST = last_owned<Own>(N);
  constexpr int NDER = (ND + 1) * (ND + 2) / 2;
  const double X = x0 * 2.0 - 1.0, Y = x1 * 2.0 - 1.0;
  const double lin = X + 0.5 * Y + 0.5;
  const double f3 = 1.0 - x1;
  const double f3sq = f3 * f3;
  double Jp[NDER], Jpm[NDER];
#pragma unroll
  for (int d = 0; d < NDER; ++d)
  {
    Jp[d] = d == 0 ? 1.0 : 0.0;
    Jpm[d] = 0.0;
  }
  int seq = 0;
#pragma unroll
  for (int p = 0; p <= LAST; ++p)
  {
    if (p > 0)
    {
      const double a = double(2 * p - 1) / double(p);
      double Jn[NDER];
#pragma unroll
      for (int kx = 0; kx <= ND; ++kx)
#pragma unroll
        for (int ky = 0; ky <= ND - kx; ++ky)
        {
          const int d = idx2(kx, ky);
          double v = (lin * a) * Jp[d];
          if (kx > 0)
            v += (2.0 * kx * a) * Jp[idx2(kx - 1, ky)];
          if (ky > 0)
            v += (ky * a) * Jp[idx2(kx, ky - 1)];
          if (p > 1)
          {
            v -= (f3sq * (a - 1.0)) * Jpm[d];
            if (ky > 0)
              v -= (ky * (a - 1.0) * (Y - 1.0)) * Jpm[idx2(kx, ky - 1)];
            if (ky > 1)
              v -= (ky * (ky - 1) * (a - 1.0)) * Jpm[idx2(kx, ky - 2)];
          }
          Jn[d] = v;
        }
#pragma unroll
      for (int d = 0; d < NDER; ++d)
      {
        Jpm[d] = Jp[d];
        Jp[d] = Jn[d];
      }
    }
    if (!Own::owns(p))
    {
      seq += subtree_size(2, N, p);
      continue;
    }
    emit(seq++, Jp);
    // q chain
    double Jq[NDER], Jqm[NDER];
#pragma unroll
    for (int d = 0; d < NDER; ++d)
    {
      Jq[d] = Jp[d];
      Jqm[d] = 0.0;
    }
#pragma unroll
    for (int q = 1; q <= N - p; ++q)
    {
      double Jn[NDER];
      if (q == 1)
      {
        const double c0 = Y * (1.5 + p) + 0.5 + p;
#pragma unroll
        for (int kx = 0; kx <= ND; ++kx)
#pragma unroll
          for (int ky = 0; ky <= ND - kx; ++ky)
          {
            const int d = idx2(kx, ky);
            double v = c0 * Jq[d];
            if (ky > 0)
              v += (2.0 * ky * (1.5 + p)) * Jq[idx2(kx, ky - 1)];
            Jn[d] = v;
          }
      }
      else
      {
        const double a1 = jrc_a(2 * p + 1, q - 1), a2 = jrc_b(2 * p + 1, q - 1), a3 = jrc_c(2 * p + 1, q - 1);
        const double c0 = Y * a1 + a2;
#pragma unroll
        for (int kx = 0; kx <= ND; ++kx)
#pragma unroll
          for (int ky = 0; ky <= ND - kx; ++ky)
          {
            const int d = idx2(kx, ky);
            double v = c0 * Jq[d] - a3 * Jqm[d];
            if (ky > 0)
              v += (2.0 * ky * a1) * Jq[idx2(kx, ky - 1)];
            Jn[d] = v;
          }
      }
#pragma unroll
      for (int d = 0; d < NDER; ++d)
      {
        Jqm[d] = Jq[d];
        Jq[d] = Jn[d];
      }
      emit(seq++, Jq);
    }
  }
}

// ---- tetrahedron -----------------------------------------------------------------------------------
template <int N, int ND, typename Own = OwnAll, typename Emit>
__device__ __forceinline__ void tetrahedron(double x0, double x1, double x2, Emit&& emit)
{
  constexpr int LAST = last_owned<Own>(N);
  constexpr int NDER = (ND + 1) * (ND + 2) * (ND + 3) / 6;
  const double X = x0 * 2.0 - 1.0, Y = x1 * 2.0 - 1.0, Z = x2 * 2.0 - 1.0;
  const double lin = X + 0.5 * (Y + Z) + 1.0;
  const double f2 = x1 + x2 - 1.0;
  const double f2sq = f2 * f2;
  const double YZ = Y + Z;
  const double f4 = 1.0 - x2;
  const double f4sq = f4 * f4;
  const double f3 = Y + x2;
  double Jp[NDER], Jpm[NDER];
#pragma unroll
  for (int d = 0; d < NDER; ++d)
  {
    Jp[d] = d == 0 ? 1.0 : 0.0;
    Jpm[d] = 0.0;
  }
  int seq = 0;
#pragma unroll
  for (int p = 0; p <= LAST; ++p)
  {
    if (p > 0)
    {
      const double a = double(2 * p - 1) / double(p);
      double Jn[NDER];
#pragma unroll
      for (int kx = 0; kx <= ND; ++kx)
#pragma unroll
        for (int ky = 0; ky <= ND - kx; ++ky)
#pragma unroll
          for (int kz = 0; kz <= ND - kx - ky; ++kz)
          {
            const int d = idx3(kx, ky, kz);
            double v = (lin * a) * Jp[d];
            if (kx > 0)
              v += (2.0 * kx * a) * Jp[idx3(kx - 1, ky, kz)];
            if (ky > 0)
              v += (ky * a) * Jp[idx3(kx, ky - 1, kz)];
            if (kz > 0)
              v += (kz * a) * Jp[idx3(kx, ky, kz - 1)];
            if (p > 1)
            {
              v -= (f2sq * (a - 1.0)) * Jpm[d];
              if (ky > 0)
                v -= (ky * (a - 1.0) * YZ) * Jpm[idx3(kx, ky - 1, kz)];
              if (ky > 1)
                v -= (ky * (ky - 1) * (a - 1.0)) * Jpm[idx3(kx, ky - 2, kz)];
              if (kz > 0)
                v -= (kz * (a - 1.0) * YZ) * Jpm[idx3(kx, ky, kz - 1)];
              if (kz > 1)
                v -= (kz * (kz - 1) * (a - 1.0)) * Jpm[idx3(kx, ky, kz - 2)];
              if (ky > 0 and kz > 0)
                v -= (2.0 * ky * kz * (a - 1.0)) * Jpm[idx3(kx, ky - 1, kz - 1)];
            }
            Jn[d] = v;
          }
#pragma unroll
      for (int d = 0; d < NDER; ++d)
      {
        Jpm[d] = Jp[d];
        Jp[d] = Jn[d];
      }
    }
    if (!Own::owns(p))
    {
      seq += subtree_size(3, N, p);
      continue;
    }
    // q chain (r = 0), each entry followed by its r chain
    double Jq[NDER], Jqm[NDER];
#pragma unroll
    for (int d = 0; d < NDER; ++d)
    {
      Jq[d] = Jp[d];
      Jqm[d] = 0.0;
    }
#pragma unroll
    for (int q = 0; q <= N - p; ++q)
    {
      if (q > 0)
      {
        double Jn[NDER];
        if (q == 1)
        {
          const double c0 = (1.0 + Y) * p + (2.0 + Y * 3.0 + Z) * 0.5;
#pragma unroll
          for (int kx = 0; kx <= ND; ++kx)
#pragma unroll
            for (int ky = 0; ky <= ND - kx; ++ky)
#pragma unroll
              for (int kz = 0; kz <= ND - kx - ky; ++kz)
              {
                const int d = idx3(kx, ky, kz);
                double v = c0 * Jq[d];
                if (ky > 0)
                  v += (2.0 * ky * (1.5 + p)) * Jq[idx3(kx, ky - 1, kz)];
                if (kz > 0)
                  v += double(kz) * Jq[idx3(kx, ky, kz - 1)];
                Jn[d] = v;
              }
        }
        else
        {
          const double aq = jrc_a(2 * p + 1, q - 1), bq = jrc_b(2 * p + 1, q - 1), cq = jrc_c(2 * p + 1, q - 1);
          const double c0 = f3 * aq + f4 * bq;
          const double c1 = f4sq * cq;
          const double c2 = (1.0 - Z) * cq;
#pragma unroll
          for (int kx = 0; kx <= ND; ++kx)
#pragma unroll
            for (int ky = 0; ky <= ND - kx; ++ky)
#pragma unroll
              for (int kz = 0; kz <= ND - kx - ky; ++kz)
              {
                const int d = idx3(kx, ky, kz);
                double v = c0 * Jq[d] - c1 * Jqm[d];
                if (ky > 0)
                  v += (2.0 * ky * aq) * Jq[idx3(kx, ky - 1, kz)];
                if (kz > 0)
                {
                  v += (kz * (aq - bq)) * Jq[idx3(kx, ky, kz - 1)];
                  v += (kz * c2) * Jqm[idx3(kx, ky, kz - 1)];
                }
                if (kz > 1)
                  v -= (kz * (kz - 1) * cq) * Jqm[idx3(kx, ky, kz - 2)];
                Jn[d] = v;
              }
        }
#pragma unroll
        for (int d = 0; d < NDER; ++d)
        {
          Jqm[d] = Jq[d];
          Jq[d] = Jn[d];
        }
      }
      emit(seq++, Jq);
      // r chain
      double Jr[NDER], Jrm[NDER];
#pragma unroll
      for (int d = 0; d < NDER; ++d)
      {
        Jr[d] = Jq[d];
        Jrm[d] = 0.0;
      }
#pragma unroll
      for (int r = 1; r <= N - p - q; ++r)
      {
        double Jn[NDER];
        if (r == 1)
        {
          const double c0 = (1.0 + p + q) + Z * (2.0 + p + q);
#pragma unroll
          for (int kx = 0; kx <= ND; ++kx)
#pragma unroll
            for (int ky = 0; ky <= ND - kx; ++ky)
#pragma unroll
              for (int kz = 0; kz <= ND - kx - ky; ++kz)
              {
                const int d = idx3(kx, ky, kz);
                double v = c0 * Jr[d];
                if (kz > 0)
                  v += (2.0 * kz * (2.0 + p + q)) * Jr[idx3(kx, ky, kz - 1)];
                Jn[d] = v;
              }
        }
        else
        {
          const double ar = jrc_a(2 * p + 2 * q + 2, r - 1), br = jrc_b(2 * p + 2 * q + 2, r - 1),
                       cr = jrc_c(2 * p + 2 * q + 2, r - 1);
          const double c0 = Z * ar + br;
#pragma unroll
          for (int kx = 0; kx <= ND; ++kx)
#pragma unroll
            for (int ky = 0; ky <= ND - kx; ++ky)
#pragma unroll
              for (int kz = 0; kz <= ND - kx - ky; ++kz)
              {
                const int d = idx3(kx, ky, kz);
                double v = c0 * Jr[d] - cr * Jrm[d];
                if (kz > 0)
                  v += (2.0 * kz * ar) * Jr[idx3(kx, ky, kz - 1)];
                Jn[d] = v;
              }
        }
#pragma unroll
        for (int d = 0; d < NDER; ++d)
        {
          Jrm[d] = Jr[d];
          Jr[d] = Jn[d];
        }
        emit(seq++, Jr);
      }
    }
  }
}

// ---- host mirror of the production order: seq -> (basis slot, norm) ---------------------------------------
// norms: interval sqrt(2p+1) (polyset.cpp:113-119); triangle 2 sqrt((p+1/2)(p+q+1)) (polyset.cpp:1728-1742,
// degree 0: sqrt 2); tetrahedron 2 sqrt(2 (p+1/2)(p+q+1)(p+q+r+3/2)) (polyset.cpp:2004-2019, degree 0: sqrt 6)
void production_order(int tdim, int n, int* slot, double* norm);
} // namespace jets
} // namespace bx

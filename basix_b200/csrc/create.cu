// Native element construction (host side, runs once per element): basix.create_element for the Lagrange (P)
// family on interval / triangle / tetrahedron / quadrilateral / hexahedron, and for Raviart-Thomas and Nedelec
// (first kind) on triangle / tetrahedron.
//
// This is set-up code, not the accelerated path: it produces the immutable arrays the hot path consumes
// (coefficient matrix, entity dofs, entity transformations) and hands them to element_from_desc().  What is
// computed follows the reference; how it is organised is this library's own:
//   lattice points            cpp/basix/lattice.cpp:42-71 (interval, equispaced / GLL), 201-216 (warp function),
//                             218-280 (quad / hex), 282-350 (triangle), 503-535 + 583-633 (tetrahedron)
//   GLL nodes                 cpp/basix/quadrature.cpp:470-510 (there: Golub-Welsch with fixed end nodes; here:
//                             Newton on (1 - x^2) P'_n, the same nodes to rounding)
//   Gauss-Jacobi quadrature   cpp/basix/quadrature.cpp:215-372 (collapsed rules on simplices)
//   dof points per entity     cpp/basix/e-lagrange.cpp:457-666 (vertices, then edges, faces, interior; the
//                             sub-entity lattice mapped by v0 + sum_k (v_{k+1} - v0) t_k)
//   RT / N1curl spaces        cpp/basix/e-raviart-thomas.cpp:20-173, e-nedelec.cpp:24-380; math.h:386-444
//   integral moment dofs      cpp/basix/moments.cpp:100-190 (vector pad-out), 262-330 (tangent), 336-440 (normal)
//   coefficient matrix        cpp/basix/finite-element.cpp:84-175, 1105-1119: C = (B D)^-1 B
//   entity transformations    cpp/basix/dof-transformations.cpp:71-330 (the cell-level maps that reverse /
//                             rotate / reflect the first sub-entity of each type), 340-470 (tabulate the
//                             element basis at the mapped dof points of that entity, push it forward with the
//                             element's Piola map, interpolate with the entity's dofs)
//   discontinuous variant     cpp/basix/finite-element.cpp make_discontinuous: every dof moves to the interior
// The orthonormal polynomial sets are evaluated with the SAME recurrence code the device kernels use
// (polyset.cuh, compiled for the host here).
#include "element.cuh"
#include "polyset.cuh"

#include <algorithm>
#include <cmath>
#include <functional>
#include <limits>
#include <memory>
#include <numeric>

namespace bx
{
namespace
{
using Vec = std::vector<double>;

// ---- host polyset tabulate: P[nder][psize][npts] (the device recurrences of polyset.cuh, run on the host) ----
Vec host_polyset_nd(int cell, int degree, int nd, const Vec& x, size_t npts)
{
  const int tdim = cell_tdim(cell);
  const int psize = polyset_dim(cell, BX_POLYSET_STANDARD, degree);
  const int nder = polyset_nderivs(cell, nd);
  Vec P(size_t(nder) * psize * npts, 0.0);
  const int n1 = degree + 1;
  Vec L(size_t(3) * (nd + 1) * n1);
  for (size_t i = 0; i < npts; ++i)
  {
    PointView<double> v{P.data() + i, npts, psize};
    switch (cell)
    {
    case BX_CELL_POINT: P[i] = 1.0; break;
    case BX_CELL_INTERVAL: polyset_interval<double>(v, degree, nd, x[i]); break;
    case BX_CELL_TRIANGLE: polyset_triangle<double>(v, degree, nd, x[2 * i], x[2 * i + 1]); break;
    case BX_CELL_TETRAHEDRON: polyset_tetrahedron<double>(v, degree, nd, x[3 * i], x[3 * i + 1], x[3 * i + 2]); break;
    case BX_CELL_PRISM: polyset_prism<double>(v, degree, nd, x[3 * i], x[3 * i + 1], x[3 * i + 2]); break;
    case BX_CELL_PYRAMID: polyset_pyramid<double>(v, degree, nd, x[3 * i], x[3 * i + 1], x[3 * i + 2]); break;
    case BX_CELL_QUADRILATERAL:
    case BX_CELL_HEXAHEDRON:
    {
      const int tsz = (nd + 1) * n1; // per-axis table L[k][p]
      for (int a = 0; a < tdim; ++a)
        legendre_table<double>(L.data() + size_t(a) * tsz, degree, nd, x[tdim * i + a]);
      if (tdim == 2)
      {
        for (int kx = 0; kx <= nd; ++kx)
          for (int ky = 0; ky <= nd - kx; ++ky)
            for (int px = 0; px < n1; ++px)
              for (int py = 0; py < n1; ++py)
                v(idx2(kx, ky), px * n1 + py) = L[kx * n1 + px] * L[tsz + ky * n1 + py];
      }
      else
      {
        for (int kx = 0; kx <= nd; ++kx)
          for (int ky = 0; ky <= nd - kx; ++ky)
            for (int kz = 0; kz <= nd - kx - ky; ++kz)
              for (int px = 0; px < n1; ++px)
                for (int py = 0; py < n1; ++py)
                  for (int pz = 0; pz < n1; ++pz)
                    v(idx3(kx, ky, kz), (px * n1 + py) * n1 + pz)
                        = L[kx * n1 + px] * L[tsz + ky * n1 + py] * L[2 * tsz + kz * n1 + pz];
      }
      break;
    }
    default: fail(BX_ERR_UNSUPPORTED, "create_element: cell type not supported by the native factory");
    }
  }
  return P;
}

Vec host_polyset(int cell, int degree, const Vec& x, size_t npts) { return host_polyset_nd(cell, degree, 0, x, npts); }

// ---- dense LU solve with partial pivoting: returns X with A X = B (A n x n, B n x m, row-major) -------------
Vec lu_solve(Vec A, Vec B, int n, int m)
{
  for (int k = 0; k < n; ++k)
  {
    int piv = k;
    for (int i = k + 1; i < n; ++i)
      if (std::abs(A[size_t(i) * n + k]) > std::abs(A[size_t(piv) * n + k]))
        piv = i;
    if (A[size_t(piv) * n + k] == 0.0)
      fail(BX_ERR_RUNTIME, "Dual matrix is singular, there is an error in your inputs");
    if (piv != k)
    {
      for (int j = 0; j < n; ++j)
        std::swap(A[size_t(k) * n + j], A[size_t(piv) * n + j]);
      for (int j = 0; j < m; ++j)
        std::swap(B[size_t(k) * m + j], B[size_t(piv) * m + j]);
    }
    const double inv = 1.0 / A[size_t(k) * n + k];
    for (int i = k + 1; i < n; ++i)
    {
      const double f = A[size_t(i) * n + k] * inv;
      if (f == 0.0)
        continue;
      for (int j = k + 1; j < n; ++j)
        A[size_t(i) * n + j] -= f * A[size_t(k) * n + j];
      for (int j = 0; j < m; ++j)
        B[size_t(i) * m + j] -= f * B[size_t(k) * m + j];
    }
  }
  for (int k = n - 1; k >= 0; --k)
  {
    const double inv = 1.0 / A[size_t(k) * n + k];
    for (int j = 0; j < m; ++j)
    {
      double s = B[size_t(k) * m + j];
      for (int i = k + 1; i < n; ++i)
        s -= A[size_t(k) * n + i] * B[size_t(i) * m + j];
      B[size_t(k) * m + j] = s * inv;
    }
  }
  return B;
}

// ---- Gauss-Lobatto-Legendre nodes on [0, 1], ordered as the reference does: both ends first, then the
//      interior nodes ascending (quadrature.cpp:494-497) -------------------------------------------------------
Vec gll_points(int m)
{
  if (m < 2)
    fail(BX_ERR_INVALID_ARG, "Gauss-Lobatto-Legendre quadrature invalid for fewer than 2 points");
  const int n = m - 1; // interior nodes are the roots of P'_n
  Vec x(m);
  x[0] = 0.0;
  x[1] = 1.0;
  for (int k = 1; k < n; ++k)
  {
    // Chebyshev-Gauss-Lobatto start, Newton on q(t) = P'_n(t), q' from the Legendre ODE
    double t = -std::cos(M_PI * k / n);
    for (int it = 0; it < 100; ++it)
    {
      double p0 = 1.0, p1 = t;
      for (int j = 2; j <= n; ++j)
      {
        const double p2 = ((2 * j - 1) * t * p1 - (j - 1) * p0) / j;
        p0 = p1;
        p1 = p2;
      }
      const double dp = n * (t * p1 - p0) / (t * t - 1.0);              // P'_n
      const double ddp = (2.0 * t * dp - n * (n + 1.0) * p1) / (1.0 - t * t); // P''_n
      const double dt = dp / ddp;
      t -= dt;
      if (std::abs(dt) < 1e-16)
        break;
    }
    x[k + 1] = 0.5 * (t + 1.0);
  }
  // symmetrise (the exact nodes are symmetric about 1/2)
  for (int k = 1; k < n; ++k)
  {
    const int mirror = n - k;
    if (k < mirror)
    {
      const double a = 0.5 * (x[k + 1] + (1.0 - x[mirror + 1]));
      x[k + 1] = a;
      x[mirror + 1] = 1.0 - a;
    }
    else if (k == mirror)
      x[k + 1] = 0.5;
  }
  return x;
}

Vec linspace(double x0, double x1, int n)
{
  if (n <= 0)
    return {};
  if (n == 1)
    return {x0};
  Vec p(n, x0);
  p.back() = x1;
  const double delta = (x1 - x0) / (n - 1);
  for (int i = 1; i < n - 1; ++i)
    p[i] += i * delta;
  return p;
}

// lattice.cpp:42-71, 140-166
Vec interval_lattice(int n, bool gll, bool exterior)
{
  if (n == 0)
    return {0.5};
  if (!gll)
  {
    const double h = exterior ? 0.0 : 1.0 / n;
    return linspace(h, 1.0 - h, exterior ? n + 1 : n - 1);
  }
  const Vec pts = gll_points(n + 1);
  const int b = exterior ? 0 : 1;
  Vec x(n + 1 - 2 * b);
  if (exterior)
  {
    x[0] = pts[0];
    x[n] = pts[1];
  }
  for (int j = 2; j < n + 1; ++j)
    x[j - 1 - b] = pts[j];
  return x;
}

// lattice.cpp:168-216: displacement of the GLL nodes from the equispaced nodes, interpolated (degree n
// Lagrange interpolant through the equispaced nodes) at the points r
Vec warp_function(int n, const Vec& r)
{
  Vec d = interval_lattice(n, true, true);
  for (int i = 0; i <= n; ++i)
    d[i] -= double(i) / n;
  // interpolation matrix through the orthonormal Legendre set, as the reference does: solve P(equi) c = P(r)
  Vec equi(n + 1);
  for (int i = 0; i <= n; ++i)
    equi[i] = double(i) / n;
  const Vec V = host_polyset(BX_CELL_INTERVAL, n, equi, n + 1);   // [n+1][n+1]
  const Vec R = host_polyset(BX_CELL_INTERVAL, n, r, r.size());   // [n+1][nr]
  const Vec c = lu_solve(V, R, n + 1, int(r.size()));             // c[i][j] = l_i(r_j)
  Vec w(r.size(), 0.0);
  for (int i = 0; i <= n; ++i)
    for (size_t j = 0; j < r.size(); ++j)
      w[j] += c[size_t(i) * r.size() + j] * d[i];
  return w;
}

// Interior (exterior = false) or full lattice of a cell, row-major [npts][tdim]; lattice.cpp:218-350, 503-633.
// Simplices with GLL spacing use the warp method (the gll_warped variant).
Vec lattice_create(int cell, int n, bool gll, bool exterior)
{
  const int b = exterior ? 0 : 1;
  switch (cell)
  {
  case BX_CELL_POINT: return {};
  case BX_CELL_INTERVAL: return interval_lattice(n, gll, exterior);
  case BX_CELL_QUADRILATERAL:
  {
    if (n == 0)
      return {0.5, 0.5};
    const Vec r = interval_lattice(n, gll, exterior);
    Vec x;
    for (size_t j = 0; j < r.size(); ++j)
      for (size_t i = 0; i < r.size(); ++i)
      {
        x.push_back(r[i]);
        x.push_back(r[j]);
      }
    return x;
  }
  case BX_CELL_HEXAHEDRON:
  {
    if (n == 0)
      return {0.5, 0.5, 0.5};
    const Vec r = interval_lattice(n, gll, exterior);
    Vec x;
    for (size_t k = 0; k < r.size(); ++k)
      for (size_t j = 0; j < r.size(); ++j)
        for (size_t i = 0; i < r.size(); ++i)
        {
          x.push_back(r[i]);
          x.push_back(r[j]);
          x.push_back(r[k]);
        }
    return x;
  }
  case BX_CELL_TRIANGLE:
  {
    if (n == 0)
      return {1.0 / 3.0, 1.0 / 3.0};
    const Vec r = linspace(0.0, 1.0, 2 * n + 1);
    Vec wbar;
    if (gll)
    {
      wbar = warp_function(n, r);
      for (int i = 1; i < 2 * n - 1; ++i)
        wbar[i] /= r[i] * (1.0 - r[i]);
    }
    Vec p;
    for (int j = b; j < n - b + 1; ++j)
      for (int i = b; i < n - b + 1 - j; ++i)
      {
        const double x = r[2 * i], y = r[2 * j];
        double px = x, py = y;
        if (gll)
        {
          const int l = n - j - i;
          const double a = r[2 * l];
          px += x * (a * wbar[n + i - l] + y * wbar[n + i - j]);
          py += y * (a * wbar[n + j - l] + x * wbar[n + j - i]);
        }
        p.push_back(px);
        p.push_back(py);
      }
    return p;
  }
  case BX_CELL_TETRAHEDRON:
  {
    if (n == 0)
      return {0.25, 0.25, 0.25};
    const Vec r = linspace(0.0, 1.0, 2 * n + 1);
    Vec wbar;
    if (gll)
    {
      wbar = warp_function(n, r);
      for (int i = 1; i < 2 * n; ++i)
        wbar[i] /= r[i] * (1.0 - r[i]);
    }
    Vec p;
    for (int k = b; k < n - b + 1; ++k)
      for (int j = b; j < n - b + 1 - k; ++j)
        for (int i = b; i < n - b + 1 - j - k; ++i)
        {
          const int l = n - k - j - i;
          const double x = r[2 * i], y = r[2 * j], z = r[2 * k], a = r[2 * l];
          double px = x, py = y, pz = z;
          if (gll)
          {
            px += x * (a * wbar[n + i - l] + y * wbar[n + i - j] + z * wbar[n + i - k]);
            py += y * (a * wbar[n + j - l] + z * wbar[n + j - k] + x * wbar[n + j - i]);
            pz += z * (a * wbar[n + k - l] + x * wbar[n + k - i] + y * wbar[n + k - j]);
          }
          p.push_back(px);
          p.push_back(py);
          p.push_back(pz);
        }
    return p;
  }
  default: fail(BX_ERR_UNSUPPORTED, "create_element: cell type not supported by the native factory");
  }
}

// ---- reference cells: vertices and vertex lists of the sub-entities (cell.cpp:18-130) ----------------------
Vec cell_geometry(int cell)
{
  switch (cell)
  {
  case BX_CELL_INTERVAL: return {0.0, 1.0};
  case BX_CELL_TRIANGLE: return {0, 0, 1, 0, 0, 1};
  case BX_CELL_QUADRILATERAL: return {0, 0, 1, 0, 0, 1, 1, 1};
  case BX_CELL_TETRAHEDRON: return {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1};
  case BX_CELL_HEXAHEDRON: return {0, 0, 0, 1, 0, 0, 0, 1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 1, 0, 1, 1, 1, 1, 1};
  default: fail(BX_ERR_UNSUPPORTED, "create_element: cell type not supported by the native factory");
  }
}

std::vector<std::vector<std::vector<int>>> cell_topology(int cell)
{
  switch (cell)
  {
  case BX_CELL_INTERVAL: return {{{0}, {1}}, {{0, 1}}};
  case BX_CELL_TRIANGLE: return {{{0}, {1}, {2}}, {{1, 2}, {0, 2}, {0, 1}}, {{0, 1, 2}}};
  case BX_CELL_QUADRILATERAL: return {{{0}, {1}, {2}, {3}}, {{0, 1}, {0, 2}, {1, 3}, {2, 3}}, {{0, 1, 2, 3}}};
  case BX_CELL_TETRAHEDRON:
    return {{{0}, {1}, {2}, {3}},
            {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}},
            {{1, 2, 3}, {0, 2, 3}, {0, 1, 3}, {0, 1, 2}},
            {{0, 1, 2, 3}}};
  case BX_CELL_HEXAHEDRON:
    return {{{0}, {1}, {2}, {3}, {4}, {5}, {6}, {7}},
            {{0, 1}, {0, 2}, {0, 4}, {1, 3}, {1, 5}, {2, 3}, {2, 6}, {3, 7}, {4, 5}, {4, 6}, {5, 7}, {6, 7}},
            {{0, 1, 2, 3}, {0, 1, 4, 5}, {0, 2, 4, 6}, {1, 3, 5, 7}, {2, 3, 6, 7}, {4, 5, 6, 7}},
            {{0, 1, 2, 3, 4, 5, 6, 7}}};
  default: fail(BX_ERR_UNSUPPORTED, "create_element: cell type not supported by the native factory");
  }
}

// cell-level maps that act on the first sub-entity of each type (dof-transformations.cpp:71-330)
struct EntityMap
{
  int entity_type;
  std::function<std::array<double, 3>(const double*)> map;
};
std::vector<EntityMap> entity_maps(int cell)
{
  using A = std::array<double, 3>;
  switch (cell)
  {
  case BX_CELL_TRIANGLE: return {{BX_CELL_INTERVAL, [](const double* p) { return A{p[1], p[0], 0}; }}};
  case BX_CELL_QUADRILATERAL: return {{BX_CELL_INTERVAL, [](const double* p) { return A{1 - p[0], p[1], 0}; }}};
  case BX_CELL_TETRAHEDRON:
    return {{BX_CELL_INTERVAL, [](const double* p) { return A{p[0], p[2], p[1]}; }},
            {BX_CELL_TRIANGLE, [](const double* p) { return A{p[2], p[0], p[1]}; }},   // rotation
            {BX_CELL_TRIANGLE, [](const double* p) { return A{p[0], p[2], p[1]}; }}};  // reflection
  case BX_CELL_HEXAHEDRON:
    return {{BX_CELL_INTERVAL, [](const double* p) { return A{1 - p[0], p[1], p[2]}; }},
            {BX_CELL_QUADRILATERAL, [](const double* p) { return A{1 - p[1], p[0], p[2]}; }}, // rotation
            {BX_CELL_QUADRILATERAL, [](const double* p) { return A{p[1], p[0], p[2]}; }}};    // reflection
  default: return {};
  }
}

// ---- Gauss-Jacobi quadrature (quadrature.cpp:215-372; Karniadakis & Sherwin collapsed rules) ----------------
// P_m^{(a,0)} and its derivative at x by the three-term recurrence
void jacobi_a0(double a, int m, double x, double& p, double& dp)
{
  double p0 = 1.0, d0 = 0.0;
  if (m == 0)
  {
    p = 1.0, dp = 0.0;
    return;
  }
  double p1 = (x * (a + 2.0) + a) * 0.5, d1 = a * 0.5 + 1.0;
  for (int j = 2; j <= m; ++j)
  {
    const double a1 = 2.0 * j * (j + a) * (2.0 * j + a - 2.0);
    const double a2 = (2.0 * j + a - 1.0) * (a * a) / a1;
    const double a3 = (2.0 * j + a - 1.0) * (2.0 * j + a) / (2.0 * j * (j + a));
    const double a4 = 2.0 * (j + a - 1.0) * (j - 1.0) * (2.0 * j + a) / a1;
    const double p2 = p1 * (x * a3 + a2) - p0 * a4;
    const double d2 = d1 * (x * a3 + a2) - d0 * a4 + a3 * p1;
    p0 = p1, p1 = p2, d0 = d1, d1 = d2;
  }
  p = p1, dp = d1;
}

// m-point Gauss-Jacobi rule for the weight (1 - x)^a on [-1, 1]: Newton with deflation from Chebyshev guesses
void gauss_jacobi(double a, int m, Vec& pts, Vec& wts)
{
  pts.assign(m, 0.0);
  wts.assign(m, 0.0);
  for (int k = 0; k < m; ++k)
  {
    double x = -std::cos((2.0 * k + 1.0) * M_PI / (2.0 * m));
    if (k > 0)
      x = 0.5 * (x + pts[k - 1]);
    for (int it = 0; it < 100; ++it)
    {
      double s = 0.0;
      for (int i = 0; i < k; ++i)
        s += 1.0 / (x - pts[i]);
      double f, df;
      jacobi_a0(a, m, x, f, df);
      const double delta = f / (df - f * s);
      x -= delta;
      if (std::abs(delta) < 1e-15)
        break;
    }
    pts[k] = x;
  }
  for (int k = 0; k < m; ++k)
  {
    double f, df;
    jacobi_a0(a, m, pts[k], f, df);
    wts[k] = std::pow(2.0, a + 1.0) / (1.0 - pts[k] * pts[k]) / (df * df);
  }
}

// Quadrature exact for polynomials of degree q on interval / triangle / tetrahedron: x[n][tdim], w[n].  The
// reference's default for these cells is a tabulated Xiao-Gimbutas rule below degree 30 (quadrature.cpp:4960-4990);
// every integrand formed here is a polynomial within the exactness degree, so any exact rule yields the same
// moments and the same element to rounding.
void make_quadrature(int cell, int q, Vec& x, Vec& w)
{
  const int m = (q + 2) / 2;
  Vec px, wx, py, wy, pz, wz;
  gauss_jacobi(0.0, m, px, wx);
  x.clear();
  w.clear();
  if (cell == BX_CELL_INTERVAL)
  {
    for (int i = 0; i < m; ++i)
    {
      x.push_back(0.5 * (px[i] + 1.0));
      w.push_back(0.5 * wx[i]);
    }
    return;
  }
  gauss_jacobi(1.0, m, py, wy);
  if (cell == BX_CELL_TRIANGLE)
  {
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j)
      {
        x.push_back(0.25 * (1.0 + px[i]) * (1.0 - py[j]));
        x.push_back(0.5 * (1.0 + py[j]));
        w.push_back(wx[i] * wy[j] * 0.125);
      }
    return;
  }
  if (cell == BX_CELL_QUADRILATERAL or cell == BX_CELL_HEXAHEDRON)
  {
    // tensor Gauss-Legendre, first coordinate slowest
    const int tdim = cell == BX_CELL_QUADRILATERAL ? 2 : 3;
    const int n = tdim == 2 ? m * m : m * m * m;
    for (int q3 = 0; q3 < n; ++q3)
    {
      const int i = tdim == 2 ? q3 / m : q3 / (m * m), j = tdim == 2 ? q3 % m : (q3 / m) % m, k = q3 % m;
      x.push_back(0.5 * (px[i] + 1.0));
      x.push_back(0.5 * (px[j] + 1.0));
      double wt = 0.25 * wx[i] * wx[j];
      if (tdim == 3)
      {
        x.push_back(0.5 * (px[k] + 1.0));
        wt *= 0.5 * wx[k];
      }
      w.push_back(wt);
    }
    return;
  }
  if (cell == BX_CELL_PRISM)
  {
    // collapsed triangle rule in (x, y), Gauss-Legendre in z
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j)
        for (int k = 0; k < m; ++k)
        {
          x.push_back(0.25 * (1.0 + px[i]) * (1.0 - py[j]));
          x.push_back(0.5 * (1.0 + py[j]));
          x.push_back(0.5 * (px[k] + 1.0));
          w.push_back(wx[i] * wy[j] * 0.125 * 0.5 * wx[k]);
        }
    return;
  }
  if (cell == BX_CELL_PYRAMID)
  {
    // x = (1 + xi)(1 - z) / 2, y = (1 + eta)(1 - z) / 2, z = (1 + zeta) / 2: weight (1 - zeta)^2, two more points in z
    // (the pyramid's polynomial set is rational in z)
    gauss_jacobi(2.0, m + 2, pz, wz);
    for (int i = 0; i < m; ++i)
      for (int j = 0; j < m; ++j)
        for (int k = 0; k < m + 2; ++k)
        {
          x.push_back(0.25 * (1.0 + px[i]) * (1.0 - pz[k]));
          x.push_back(0.25 * (1.0 + px[j]) * (1.0 - pz[k]));
          x.push_back(0.5 * (1.0 + pz[k]));
          w.push_back(wx[i] * wx[j] * wz[k] * 0.125 * 0.25);
        }
    return;
  }
  if (cell != BX_CELL_TETRAHEDRON)
    fail(BX_ERR_UNSUPPORTED, "create_element: quadrature on this cell is not available in the native factory");
  gauss_jacobi(2.0, m, pz, wz);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j)
      for (int k = 0; k < m; ++k)
      {
        x.push_back(0.125 * (1.0 + px[i]) * (1.0 - py[j]) * (1.0 - pz[k]));
        x.push_back(0.25 * (1.0 + py[j]) * (1.0 - pz[k]));
        x.push_back(0.5 * (1.0 + pz[k]));
        w.push_back(wx[i] * wy[j] * wz[k] * 0.125 * 0.125);
      }
}

// math::orthogonalise (math.h:386-444): Gram-Schmidt on the rows [start, n) of B[n][cols]
void orthogonalise(Vec& B, int n, int cols, int start)
{
  for (int i = start; i < n; ++i)
  {
    double* ri = B.data() + size_t(i) * cols;
    double norm = 0.0;
    for (int k = 0; k < cols; ++k)
      norm += ri[k] * ri[k];
    norm = std::sqrt(norm);
    if (norm < 2 * std::numeric_limits<double>::epsilon())
      fail(BX_ERR_RUNTIME, "Cannot orthogonalise the rows of a matrix with incomplete row rank");
    for (int k = 0; k < cols; ++k)
      ri[k] /= norm;
    for (int j = i + 1; j < n; ++j)
    {
      double* rj = B.data() + size_t(j) * cols;
      double a = 0.0;
      for (int k = 0; k < cols; ++k)
        a += rj[k] * ri[k];
      for (int k = 0; k < cols; ++k)
        rj[k] -= a * ri[k];
    }
  }
}

// Dofs of one element: x[d][e] points [np][tdim]; M[d][e] interpolation weights [ndofs_e][vs][np]
struct DofSet
{
  std::vector<std::vector<Vec>> x, M;
  std::vector<std::vector<int>> ndofs;
};

// Values of the moment space (discontinuous Lagrange of `degree` on `cell`, variant `lvariant`) at points:
// phi[i][q].  legendre variant: the orthonormal set itself (e-lagrange.cpp:211-259); equispaced / gll_warped:
// the nodal basis built by create_lagrange.
bx_element* create_lagrange_host(int cell, int degree, int variant, bool discontinuous);
Vec moment_space(int cell, int degree, int lvariant, const Vec& pts, size_t npts)
{
  const Vec P = host_polyset(cell, degree, pts, npts);
  if (lvariant == 11) // legendre
    return P;
  std::unique_ptr<bx_element> e(create_lagrange_host(cell, degree, lvariant, true));
  const int n = e->ndofs, psize = e->psize;
  Vec phi(size_t(n) * npts, 0.0);
  for (int i = 0; i < n; ++i)
    for (int k = 0; k < psize; ++k)
    {
      const double c = e->coeffs[size_t(i) * psize + k];
      for (size_t q = 0; q < npts; ++q)
        phi[size_t(i) * npts + q] += c * P[size_t(k) * npts + q];
    }
  return phi;
}

struct Entity
{
  Vec v0;                // first vertex
  std::vector<Vec> axes; // v_{k+1} - v0 (simplex sub-entities)
};
Entity entity_frame(int cell, int dim, int index)
{
  const int tdim = cell_tdim(cell);
  const Vec geom = cell_geometry(cell);
  const auto topo = cell_topology(cell);
  const auto& verts = topo[dim][index];
  Entity f;
  f.v0.assign(geom.begin() + verts[0] * tdim, geom.begin() + (verts[0] + 1) * tdim);
  for (int k = 0; k < dim; ++k)
  {
    Vec a(tdim);
    for (int q = 0; q < tdim; ++q)
      a[q] = geom[verts[k + 1] * tdim + q] - f.v0[q];
    f.axes.push_back(a);
  }
  return f;
}

// The three moment kinds of moments.cpp on the entities of dimension `edim` of a simplex cell:
//   kind 0  make_integral_moments against a vector-valued pad-out of a scalar space (moments.cpp:100-190):
//           dof i*edim + d integrates  phi_i  v . axis_d
//   kind 1  make_tangent_integral_moments (moments.cpp:262-330), edges
//   kind 2  make_normal_integral_moments (moments.cpp:336-440), facets
void add_moments(DofSet& dofs, int cell, int edim, int kind, int mdegree, int lvariant, int qdeg)
{
  const int tdim = cell_tdim(cell);
  const int ecell = edim == 1 ? BX_CELL_INTERVAL : edim == 2 ? BX_CELL_TRIANGLE : BX_CELL_TETRAHEDRON;
  Vec qx, qw;
  make_quadrature(ecell, qdeg, qx, qw);
  const size_t nq = qw.size();
  const Vec phi = moment_space(ecell, mdegree, lvariant, qx, nq);
  const int nphi = int(phi.size() / nq);
  const int nent = cell_num_entities(cell, edim);
  for (int e = 0; e < nent; ++e)
  {
    const Entity f = entity_frame(cell, edim, e);
    Vec pts(nq * tdim);
    for (size_t p = 0; p < nq; ++p)
      for (int q = 0; q < tdim; ++q)
      {
        double v = f.v0[q];
        for (int k = 0; k < edim; ++k)
          v += qx[p * edim + k] * f.axes[k][q];
        pts[p * tdim + q] = v;
      }
    const int nd = kind == 0 ? nphi * edim : nphi;
    Vec M(size_t(nd) * tdim * nq, 0.0);
    Vec dir(tdim, 0.0);
    if (kind == 1)
      dir = f.axes[0];
    else if (kind == 2)
    {
      if (tdim == 2)
        dir = {-f.axes[0][1], f.axes[0][0]};
      else
      {
        const Vec &a = f.axes[0], &b = f.axes[1];
        dir = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
      }
    }
    for (int i = 0; i < nphi; ++i)
      for (int d = 0; d < (kind == 0 ? edim : 1); ++d)
      {
        const int dof = kind == 0 ? i * edim + d : i;
        const Vec& v = kind == 0 ? f.axes[d] : dir;
        for (int j = 0; j < tdim; ++j)
          for (size_t k = 0; k < nq; ++k)
            M[(size_t(dof) * tdim + j) * nq + k] = phi[size_t(i) * nq + k] * qw[k] * v[j];
      }
    dofs.x[edim].push_back(std::move(pts));
    dofs.M[edim].push_back(std::move(M));
    dofs.ndofs[edim].push_back(nd);
  }
}

void add_empty(DofSet& dofs, int cell, int dim)
{
  for (int e = 0; e < cell_num_entities(cell, dim); ++e)
  {
    dofs.x[dim].push_back({});
    dofs.M[dim].push_back({});
    dofs.ndofs[dim].push_back(0);
  }
}

// cell-level Jacobians of the entity maps (dof-transformations.cpp:71-200): {J, detJ, K} per map, in the order
// of entity_maps()
struct MapJac
{
  Vec J;
  double detJ;
  Vec K;
};
std::vector<MapJac> entity_map_jacobians(int cell)
{
  switch (cell)
  {
  case BX_CELL_TRIANGLE: return {{{0, 1, 1, 0}, -1.0, {0, 1, 1, 0}}};
  case BX_CELL_TETRAHEDRON:
    return {{{1, 0, 0, 0, 0, 1, 0, 1, 0}, -1.0, {1, 0, 0, 0, 0, 1, 0, 1, 0}},
            {{0, 1, 0, 0, 0, 1, 1, 0, 0}, 1.0, {0, 0, 1, 1, 0, 0, 0, 1, 0}},
            {{1, 0, 0, 0, 0, 1, 0, 1, 0}, -1.0, {1, 0, 0, 0, 0, 1, 0, 1, 0}}};
  default: fail(BX_ERR_UNSUPPORTED, "create_element: vector-valued elements are built natively on simplices only");
  }
}

// The general construction of finite-element.cpp:84-175, 1105-1119 and dof-transformations.cpp:340-470 for a
// vector-valued element on a simplex: B = wcoeffs [n][tdim*psize], dofs = integral moments, Piola map `map_type`.
bx_element* build_vector_element(int cell, int degree, int map_type, Vec B, DofSet dofs, bool discontinuous)
{
  const int tdim = cell_tdim(cell), vs = tdim;
  const int psize = polyset_dim(cell, BX_POLYSET_STANDARD, degree);
  const auto topo = cell_topology(cell);
  int ndofs = 0;
  for (int d = 0; d <= tdim; ++d)
    ndofs += std::accumulate(dofs.ndofs[d].begin(), dofs.ndofs[d].end(), 0);
  if (size_t(ndofs) * vs * psize != B.size())
    fail(BX_ERR_RUNTIME, "create_element: internal error, space and dof counts differ");

  // D[(j, m)][dof] = sum_k M(dof, j, k) P_m(x_k)
  Vec D(size_t(vs) * psize * ndofs, 0.0);
  {
    int dof0 = 0;
    for (int d = 0; d <= tdim; ++d)
      for (size_t e = 0; e < dofs.x[d].size(); ++e)
      {
        const int nd = dofs.ndofs[d][e];
        if (nd == 0)
          continue;
        const size_t np = dofs.x[d][e].size() / tdim;
        const Vec P = host_polyset(cell, degree, dofs.x[d][e], np);
        const Vec& M = dofs.M[d][e];
        for (int i = 0; i < nd; ++i)
          for (int j = 0; j < vs; ++j)
            for (int m = 0; m < psize; ++m)
            {
              double s = 0.0;
              for (size_t k = 0; k < np; ++k)
                s += M[(size_t(i) * vs + j) * np + k] * P[size_t(m) * np + k];
              D[(size_t(j) * psize + m) * ndofs + dof0 + i] = s;
            }
        dof0 += nd;
      }
  }
  // dual = B D, C = dual^-1 B
  const int cols = vs * psize;
  Vec dual(size_t(ndofs) * ndofs, 0.0);
  for (int i = 0; i < ndofs; ++i)
    for (int k = 0; k < cols; ++k)
    {
      const double b = B[size_t(i) * cols + k];
      if (b == 0.0)
        continue;
      for (int j = 0; j < ndofs; ++j)
        dual[size_t(i) * ndofs + j] += b * D[size_t(k) * ndofs + j];
    }
  const Vec C = lu_solve(dual, B, ndofs, cols); // [ndofs][vs*psize]

  // entity transformations
  std::array<Vec, 8> etrans;
  std::array<int, 8> edim;
  edim.fill(-1);
  const auto maps = entity_maps(cell);
  const auto jacs = entity_map_jacobians(cell);
  for (size_t mi = 0; mi < maps.size(); ++mi)
  {
    const EntityMap& em = maps[mi];
    const MapJac& mj = jacs[mi];
    const int ed = cell_tdim(em.entity_type);
    const int entity = 0; // every sub-entity of a simplex has the same type
    const int n = discontinuous ? 0 : dofs.ndofs[ed][entity];
    edim[em.entity_type] = n;
    if (n == 0)
      continue;
    int dofstart = 0;
    for (int d = 0; d < ed; ++d)
      dofstart += std::accumulate(dofs.ndofs[d].begin(), dofs.ndofs[d].end(), 0);
    const Vec& pts = dofs.x[ed][entity];
    const size_t np = pts.size() / tdim;
    Vec mapped(np * tdim);
    for (size_t p = 0; p < np; ++p)
    {
      double pt[3] = {0, 0, 0};
      for (int q = 0; q < tdim; ++q)
        pt[q] = pts[p * tdim + q];
      const auto mp = em.map(pt);
      for (int q = 0; q < tdim; ++q)
        mapped[p * tdim + q] = mp[q];
    }
    const Vec Pm = host_polyset(cell, degree, mapped, np);
    // pushed(pt, dof k1, :) = Piola map of the reference values of basis function dofstart + k1 at the mapped point
    const Vec& M = dofs.M[ed][entity];
    Vec T(size_t(n) * n, 0.0);
    for (int k1 = 0; k1 < n; ++k1)
      for (size_t p = 0; p < np; ++p)
      {
        double U[3] = {0, 0, 0}, u[3] = {0, 0, 0};
        for (int j = 0; j < vs; ++j)
          for (int m = 0; m < psize; ++m)
            U[j] += C[size_t(dofstart + k1) * cols + j * psize + m] * Pm[size_t(m) * np + p];
        if (map_type == BX_MAP_COVARIANT_PIOLA) // u = K^T U  (maps.h:73-94)
        {
          for (int i = 0; i < tdim; ++i)
            for (int k = 0; k < tdim; ++k)
              u[i] += mj.K[k * tdim + i] * U[k];
        }
        else // contravariant: u = J U / detJ  (maps.h:100-124)
        {
          for (int i = 0; i < tdim; ++i)
          {
            for (int k = 0; k < tdim; ++k)
              u[i] += mj.J[i * tdim + k] * U[k];
            u[i] /= mj.detJ;
          }
        }
        // T(k1, k0) += sum_i imat(k0, i, p) pushed(p, k1, i)
        for (int k0 = 0; k0 < n; ++k0)
        {
          double s = 0.0;
          for (int i = 0; i < vs; ++i)
            s += M[(size_t(k0) * vs + i) * np + p] * u[i];
          T[size_t(k1) * n + k0] += s;
        }
      }
    etrans[em.entity_type].insert(etrans[em.entity_type].end(), T.begin(), T.end());
  }

  std::vector<int32_t> offs{0}, flat;
  {
    int dof = 0;
    for (int d = 0; d <= tdim; ++d)
      for (size_t ent = 0; ent < topo[d].size(); ++ent)
      {
        int cnt = dofs.ndofs[d][ent];
        if (discontinuous)
          cnt = (d == tdim) ? ndofs : 0;
        for (int i = 0; i < cnt; ++i)
          flat.push_back(dof++);
        offs.push_back(int32_t(flat.size()));
      }
  }
  if (flat.empty())
    flat.push_back(0);
  // points() and interpolation_matrix() (finite-element.cpp:1121-1188): the dof points of all entities in entity
  // order, and M[dof][j * npts + p] assembled from the per-entity moment matrices
  size_t npts_all = 0;
  for (int d = 0; d <= tdim; ++d)
    for (const Vec& xe : dofs.x[d])
      npts_all += xe.size() / tdim;
  Vec allx, Mall(size_t(ndofs) * vs * npts_all, 0.0);
  {
    size_t pt0 = 0;
    int dof0 = 0;
    for (int d = 0; d <= tdim; ++d)
      for (size_t e = 0; e < dofs.x[d].size(); ++e)
      {
        const Vec& xe = dofs.x[d][e];
        const size_t np = xe.size() / tdim;
        allx.insert(allx.end(), xe.begin(), xe.end());
        const int nd = dofs.ndofs[d][e];
        for (int i = 0; i < nd; ++i)
          for (int j = 0; j < vs; ++j)
            for (size_t k = 0; k < np; ++k)
              Mall[(size_t(dof0 + i) * vs + j) * npts_all + pt0 + k] = dofs.M[d][e][(size_t(i) * vs + j) * np + k];
        dof0 += nd;
        pt0 += np;
      }
  }

  static const double none = 0.0;
  bx_element_desc desc{};
  desc.points = allx.data();
  desc.npts = int(npts_all);
  desc.interpolation_matrix = Mall.data();
  desc.cell_type = cell;
  desc.polyset_type = BX_POLYSET_STANDARD;
  desc.degree = degree;
  desc.embedded_superdegree = degree;
  desc.map_type = map_type;
  desc.ndofs = ndofs;
  desc.value_size = vs;
  desc.coeffs = C.data();
  desc.entity_dof_offsets = offs.data();
  desc.entity_dofs = flat.data();
  auto set = [&](int cls, const double*& p, int& dim)
  {
    dim = edim[cls] < 0 ? 0 : edim[cls];
    p = edim[cls] < 0 ? nullptr : (etrans[cls].empty() ? &none : etrans[cls].data());
  };
  set(BX_CELL_INTERVAL, desc.etrans_interval, desc.etrans_interval_dim);
  set(BX_CELL_TRIANGLE, desc.etrans_triangle, desc.etrans_triangle_dim);
  set(BX_CELL_QUADRILATERAL, desc.etrans_quadrilateral, desc.etrans_quadrilateral_dim);
  return element_from_desc(desc);
}

void check_moment_variant(int lvariant)
{
  if (lvariant != 11 and lvariant != 1 and lvariant != 2)
    fail(BX_ERR_UNSUPPORTED, "create_element: the native factory builds RT / N1curl with the legendre, equispaced "
                             "or gll_warped variant of the moment spaces");
}
} // namespace

// basix::element::create_lagrange for the point-evaluation variants (equispaced, gll_warped).
bx_element* create_lagrange(int cell, int degree, int variant, bool discontinuous, const int32_t* dof_ordering,
                            int n_dof_ordering)
{
  constexpr int kUnset = 0, kEquispaced = 1, kGllWarped = 2;
  if (cell != BX_CELL_INTERVAL and cell != BX_CELL_TRIANGLE and cell != BX_CELL_TETRAHEDRON
      and cell != BX_CELL_QUADRILATERAL and cell != BX_CELL_HEXAHEDRON)
    fail(BX_ERR_UNSUPPORTED, "create_element: the native factory supports interval, triangle, tetrahedron, "
                             "quadrilateral and hexahedron cells");
  if (degree < 0)
    fail(BX_ERR_INVALID_ARG, "create_element: negative degree");
  if (variant == kUnset)
  {
    if (degree < 3)
      variant = kGllWarped;
    else
      fail(BX_ERR_RUNTIME, "Lagrange elements of degree > 2 need to be given a variant.");
  }
  if (variant != kEquispaced and variant != kGllWarped)
    fail(BX_ERR_UNSUPPORTED, "create_element: the native factory builds the equispaced and gll_warped Lagrange "
                             "variants; describe other elements with bx_element_create_from_arrays");
  if (degree == 0 and !discontinuous)
    fail(BX_ERR_RUNTIME, "Cannot create a continuous order 0 Lagrange basis function");
  const bool gll = variant == kGllWarped;
  if (n_dof_ordering != 0 and (!dof_ordering or n_dof_ordering != polyset_dim(cell, BX_POLYSET_STANDARD, degree)))
    fail(BX_ERR_INVALID_ARG, "Incorrect number of dofs in ordering.");
  const int tdim = cell_tdim(cell);
  const Vec geom = cell_geometry(cell);
  const auto topo = cell_topology(cell);

  // ---- dof points per entity (e-lagrange.cpp:560-640) ----
  std::vector<std::vector<Vec>> x(tdim + 1); // x[dim][entity] = [npts][tdim]
  if (degree == 0)
  {
    for (int d = 0; d < tdim; ++d)
      x[d].assign(topo[d].size(), Vec{});
    x[tdim].push_back(lattice_create(cell, 0, gll, true));
  }
  else
  {
    for (int d = 0; d <= tdim; ++d)
      for (size_t ent = 0; ent < topo[d].size(); ++ent)
      {
        const auto& verts = topo[d][ent];
        if (d == 0)
          x[d].push_back(Vec(geom.begin() + verts[0] * tdim, geom.begin() + (verts[0] + 1) * tdim));
        else if (d == tdim)
          x[d].push_back(lattice_create(cell, degree, gll, false));
        else
        {
          const int et = cell_sub_entity_type(cell, d, int(ent));
          const Vec lat = lattice_create(et, degree, gll, false);
          const size_t np = lat.size() / d;
          Vec pts(np * tdim);
          const double* v0 = geom.data() + verts[0] * tdim;
          for (size_t j = 0; j < np; ++j)
            for (int q = 0; q < tdim; ++q)
            {
              double v = v0[q];
              for (int k = 0; k < d; ++k)
                v += (geom[verts[k + 1] * tdim + q] - v0[q]) * lat[j * d + k];
              pts[j * tdim + q] = v;
            }
          x[d].push_back(std::move(pts));
        }
      }
  }

  // ---- all dof points in dof order, entity dof counts ----
  Vec allx;
  std::vector<std::vector<int>> ecount(tdim + 1);
  for (int d = 0; d <= tdim; ++d)
    for (const Vec& xe : x[d])
    {
      allx.insert(allx.end(), xe.begin(), xe.end());
      ecount[d].push_back(int(xe.size() / tdim));
    }
  const int ndofs = int(allx.size() / tdim);
  const int psize = polyset_dim(cell, BX_POLYSET_STANDARD, degree);
  if (ndofs != psize)
    fail(BX_ERR_RUNTIME, "create_element: internal error, dof count does not match the polynomial space");

  // ---- coefficient matrix: C P(x) = I  <=>  P(x)^T C^T = I ----
  const Vec P = host_polyset(cell, degree, allx, ndofs); // [psize][ndofs]
  Vec Pt(size_t(ndofs) * psize);
  for (int k = 0; k < psize; ++k)
    for (int j = 0; j < ndofs; ++j)
      Pt[size_t(j) * psize + k] = P[size_t(k) * ndofs + j];
  Vec I(size_t(ndofs) * ndofs, 0.0);
  for (int i = 0; i < ndofs; ++i)
    I[size_t(i) * ndofs + i] = 1.0;
  const Vec Ct = lu_solve(Pt, I, ndofs, ndofs); // [psize][ndofs]
  Vec C(size_t(ndofs) * psize);
  for (int k = 0; k < psize; ++k)
    for (int j = 0; j < ndofs; ++j)
      C[size_t(j) * psize + k] = Ct[size_t(k) * ndofs + j];

  // ---- entity transformations (dof-transformations.cpp:340-470 with M = point evaluations, identity map):
  //      T(k1, k0) = phi_{dofstart + k1}(map(x_{k0})) for the dof points of the first sub-entity of each type ----
  std::array<Vec, 8> etrans;
  std::array<int, 8> edim;
  edim.fill(-1);
  for (const EntityMap& em : entity_maps(cell))
  {
    const int ed = cell_tdim(em.entity_type);
    int entity = -1;
    for (size_t i = 0; i < topo[ed].size() and entity < 0; ++i)
      if (cell_sub_entity_type(cell, ed, int(i)) == em.entity_type)
        entity = int(i);
    const bool moved = discontinuous; // discontinuous: no dofs are left on sub-entities
    const Vec& pts = x[ed][entity];
    const int n = moved ? 0 : int(pts.size() / tdim);
    edim[em.entity_type] = n;
    if (n == 0)
      continue;
    int dofstart = 0;
    for (int d = 0; d < ed; ++d)
      dofstart += std::accumulate(ecount[d].begin(), ecount[d].end(), 0);
    for (int i = 0; i < entity; ++i)
      dofstart += ecount[ed][i];
    Vec mapped(size_t(n) * tdim);
    for (int p = 0; p < n; ++p)
    {
      double pt[3] = {0, 0, 0};
      for (int q = 0; q < tdim; ++q)
        pt[q] = pts[size_t(p) * tdim + q];
      const auto mp = em.map(pt);
      for (int q = 0; q < tdim; ++q)
        mapped[size_t(p) * tdim + q] = mp[q];
    }
    const Vec Pm = host_polyset(cell, degree, mapped, n); // [psize][n]
    Vec T(size_t(n) * n, 0.0);
    for (int k1 = 0; k1 < n; ++k1)
      for (int k0 = 0; k0 < n; ++k0)
      {
        double s = 0.0;
        for (int k = 0; k < psize; ++k)
          s += C[size_t(dofstart + k1) * psize + k] * Pm[size_t(k) * n + k0];
        T[size_t(k1) * n + k0] = s;
      }
    etrans[em.entity_type].insert(etrans[em.entity_type].end(), T.begin(), T.end());
  }

  // ---- entity dofs ----
  std::vector<int32_t> offs{0}, flat;
  {
    int dof = 0;
    for (int d = 0; d <= tdim; ++d)
      for (size_t ent = 0; ent < topo[d].size(); ++ent)
      {
        int cnt = ecount[d][ent];
        if (discontinuous)
          cnt = (d == tdim) ? ndofs : 0; // make_discontinuous: everything belongs to the interior
        for (int i = 0; i < cnt; ++i)
        {
          const int q = dof++;
          flat.push_back(n_dof_ordering ? dof_ordering[q] : q);
        }
        offs.push_back(int32_t(flat.size()));
      }
  }
  if (flat.empty())
    flat.push_back(0);

  // points() follow the dof order (finite-element.cpp:1228-1236); interpolation_matrix() is the identity
  Vec xdof(allx.size());
  for (int q = 0; q < ndofs; ++q)
    for (int i = 0; i < tdim; ++i)
      xdof[size_t(n_dof_ordering ? dof_ordering[q] : q) * tdim + i] = allx[size_t(q) * tdim + i];
  Vec Mid(size_t(ndofs) * ndofs, 0.0);
  for (int q = 0; q < ndofs; ++q)
    Mid[size_t(q) * ndofs + q] = 1.0;

  static const double none = 0.0;
  bx_element_desc desc{};
  desc.points = xdof.data();
  desc.npts = ndofs;
  desc.interpolation_matrix = Mid.data();
  desc.cell_type = cell;
  desc.polyset_type = BX_POLYSET_STANDARD;
  desc.degree = degree;
  desc.embedded_superdegree = degree;
  desc.map_type = BX_MAP_IDENTITY;
  desc.ndofs = ndofs;
  desc.value_size = 1;
  desc.coeffs = C.data();
  desc.dof_ordering = n_dof_ordering ? dof_ordering : nullptr;
  desc.n_dof_ordering = n_dof_ordering;
  desc.entity_dof_offsets = offs.data();
  desc.entity_dofs = flat.data();
  auto set = [&](int cls, const double*& p, int& dim)
  {
    dim = edim[cls] < 0 ? 0 : edim[cls];
    p = edim[cls] < 0 ? nullptr : (etrans[cls].empty() ? &none : etrans[cls].data());
  };
  set(BX_CELL_INTERVAL, desc.etrans_interval, desc.etrans_interval_dim);
  set(BX_CELL_TRIANGLE, desc.etrans_triangle, desc.etrans_triangle_dim);
  set(BX_CELL_QUADRILATERAL, desc.etrans_quadrilateral, desc.etrans_quadrilateral_dim);
  return element_from_desc(desc);
}

namespace
{
bx_element* create_lagrange_host(int cell, int degree, int variant, bool discontinuous)
{
  return create_lagrange(cell, degree, variant, discontinuous, nullptr, 0);
}

// "unset" moment-space variant: resolved per space as create_lagrange does (e-lagrange.cpp:491-500)
int resolve_variant(int lvariant, int space_degree)
{
  if (lvariant != 0)
    return lvariant;
  if (space_degree < 3)
    return 2; // gll_warped
  fail(BX_ERR_RUNTIME, "Lagrange elements of degree > 2 need to be given a variant.");
}

// phi[k][q] of the orthonormal set of `degree` on `cell` at a quadrature exact to degree 2*degree, with weights
void space_quadrature(int cell, int degree, Vec& pts, Vec& wts, Vec& phi)
{
  make_quadrature(cell, 2 * degree, pts, wts);
  phi = host_polyset(cell, degree, pts, wts.size());
}
} // namespace

// basix::element::create_rt (e-raviart-thomas.cpp:20-173), triangle / tetrahedron
bx_element* create_rt(int cell, int degree, int lvariant, bool discontinuous)
{
  if (cell == BX_CELL_QUADRILATERAL or cell == BX_CELL_HEXAHEDRON)
    fail(BX_ERR_UNSUPPORTED, "create_element: the native factory builds RT on triangles and tetrahedra (the reference's "
                             "quadrilateral / hexahedron variants: describe them with bx_element_create_from_arrays)");
  if (cell != BX_CELL_TRIANGLE and cell != BX_CELL_TETRAHEDRON)
    fail(BX_ERR_RUNTIME, "Unsupported cell type");
  if (degree < 1)
    fail(BX_ERR_RUNTIME, "Degree must be at least 1");
  const int tdim = cell_tdim(cell);
  const int facet = tdim == 2 ? BX_CELL_INTERVAL : BX_CELL_TRIANGLE;
  const int nv = polyset_dim(cell, BX_POLYSET_STANDARD, degree - 1);
  const int ns0 = degree >= 2 ? polyset_dim(cell, BX_POLYSET_STANDARD, degree - 2) : 0;
  const int ns = polyset_dim(facet, BX_POLYSET_STANDARD, degree - 1);
  const int psize = polyset_dim(cell, BX_POLYSET_STANDARD, degree);
  Vec pts, wts, phi;
  space_quadrature(cell, degree, pts, wts, phi);
  const size_t npts = wts.size();
  const int rows = nv * tdim + ns, cols = psize * tdim;
  Vec B(size_t(rows) * cols, 0.0);
  for (int i = 0; i < tdim; ++i)
    for (int j = 0; j < nv; ++j)
      B[size_t(nv * i + j) * cols + psize * i + j] = 1.0;
  // x_j * (homogeneous degree-1 part), expanded in the orthonormal set by quadrature
  for (int j = 0; j < tdim; ++j)
    for (int i = 0; i < ns; ++i)
      for (int kc = 0; kc < psize - nv; ++kc)
      {
        double s = 0.0;
        for (size_t k = 0; k < npts; ++k)
          s += wts[k] * pts[k * tdim + j] * phi[size_t(ns0 + i) * npts + k] * phi[size_t(nv + kc) * npts + k];
        B[size_t(nv * tdim + i) * cols + nv + kc + psize * j] = s;
      }
  orthogonalise(B, rows, cols, nv * tdim);

  DofSet dofs;
  dofs.x.resize(tdim + 1), dofs.M.resize(tdim + 1), dofs.ndofs.resize(tdim + 1);
  for (int d = 0; d < tdim - 1; ++d)
    add_empty(dofs, cell, d);
  const int fv = resolve_variant(lvariant, degree - 1);
  check_moment_variant(fv);
  add_moments(dofs, cell, tdim - 1, 2, degree - 1, fv, 2 * degree - 1);
  if (degree > 1)
  {
    const int iv = resolve_variant(lvariant, degree - 2);
    check_moment_variant(iv);
    add_moments(dofs, cell, tdim, 0, degree - 2, iv, 2 * degree - 2);
  }
  else
    add_empty(dofs, cell, tdim);
  return build_vector_element(cell, degree, BX_MAP_CONTRAVARIANT_PIOLA, std::move(B), std::move(dofs), discontinuous);
}

// basix::element::create_nedelec (e-nedelec.cpp:24-380), first kind, triangle / tetrahedron
bx_element* create_nedelec(int cell, int degree, int lvariant, bool discontinuous)
{
  if (cell == BX_CELL_QUADRILATERAL or cell == BX_CELL_HEXAHEDRON)
    fail(BX_ERR_UNSUPPORTED, "create_element: the native factory builds N1E on triangles and tetrahedra (the reference's "
                             "quadrilateral / hexahedron variants: describe them with bx_element_create_from_arrays)");
  if (cell != BX_CELL_TRIANGLE and cell != BX_CELL_TETRAHEDRON)
    fail(BX_ERR_RUNTIME, "Invalid celltype in Nedelec");
  if (degree < 1)
    fail(BX_ERR_RUNTIME, "Degree must be at least 1");
  const int tdim = cell_tdim(cell);
  const int psize = polyset_dim(cell, BX_POLYSET_STANDARD, degree);
  const int nv = polyset_dim(cell, BX_POLYSET_STANDARD, degree - 1);
  const int ns0 = degree >= 2 ? polyset_dim(cell, BX_POLYSET_STANDARD, degree - 2) : 0;
  Vec pts, wts, phi;
  space_quadrature(cell, degree, pts, wts, phi);
  const size_t npts = wts.size();
  const int ncols = psize - nv;
  const int cols = psize * tdim;
  // W[dim](i, jc) = int x_dim phi_{ns0+i} phi_{nv+jc}
  auto moment = [&](int dim, int i, int jc)
  {
    double s = 0.0;
    for (size_t k = 0; k < npts; ++k)
      s += wts[k] * pts[k * tdim + dim] * phi[size_t(ns0 + i) * npts + k] * phi[size_t(nv + jc) * npts + k];
    return s;
  };
  Vec B;
  int rows = 0;
  if (tdim == 2)
  {
    const int ns = degree;
    rows = 2 * nv + ns;
    B.assign(size_t(rows) * cols, 0.0);
    for (int i = 0; i < nv; ++i)
    {
      B[size_t(i) * cols + i] = 1.0;
      B[size_t(nv + i) * cols + psize + i] = 1.0;
    }
    for (int i = 0; i < ns; ++i)
      for (int jc = 0; jc < ncols; ++jc)
      {
        B[size_t(2 * nv + i) * cols + nv + jc] = moment(1, i, jc);
        B[size_t(2 * nv + i) * cols + nv + jc + psize] = -moment(0, i, jc);
      }
    orthogonalise(B, rows, cols, 2 * nv);
  }
  else
  {
    const int ns = degree * (degree + 1) / 2;
    const int ns_remove = degree * (degree - 1) / 2;
    rows = 6 * degree + 4 * degree * (degree - 1) + (degree - 2) * (degree - 1) * degree / 2;
    B.assign(size_t(rows) * cols, 0.0);
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < nv; ++j)
        B[size_t(i * nv + j) * cols + i * psize + j] = 1.0;
    for (int i = 0; i < ns; ++i)
      for (int jc = 0; jc < ncols; ++jc)
      {
        const int j = nv + jc;
        const double w0 = moment(0, i, jc), w1 = moment(1, i, jc), w2 = moment(2, i, jc);
        if (i >= ns_remove)
          B[size_t(3 * nv + i - ns_remove) * cols + psize + j] = -w2;
        B[size_t(3 * nv + i + ns - ns_remove) * cols + j] = w2;
        B[size_t(3 * nv + i + 2 * ns - ns_remove) * cols + j] = -w1;
        if (i >= ns_remove)
          B[size_t(3 * nv + i - ns_remove) * cols + 2 * psize + j] = w1;
        B[size_t(3 * nv + i + ns - ns_remove) * cols + 2 * psize + j] = -w0;
        B[size_t(3 * nv + i + 2 * ns - ns_remove) * cols + psize + j] = w0;
      }
    orthogonalise(B, rows, cols, 3 * nv);
  }

  DofSet dofs;
  dofs.x.resize(tdim + 1), dofs.M.resize(tdim + 1), dofs.ndofs.resize(tdim + 1);
  add_empty(dofs, cell, 0);
  const int ev = resolve_variant(lvariant, degree - 1);
  check_moment_variant(ev);
  add_moments(dofs, cell, 1, 1, degree - 1, ev, 2 * degree - 1);
  if (degree > 1)
  {
    const int fv = resolve_variant(lvariant, degree - 2);
    check_moment_variant(fv);
    add_moments(dofs, cell, 2, 0, degree - 2, fv, 2 * degree - 2);
  }
  else
    add_empty(dofs, cell, 2);
  if (tdim == 3)
  {
    if (degree > 2)
    {
      const int iv = resolve_variant(lvariant, degree - 3);
      check_moment_variant(iv);
      add_moments(dofs, cell, 3, 0, degree - 3, iv, 2 * degree - 3);
    }
    else
      add_empty(dofs, cell, 3);
  }
  return build_vector_element(cell, degree, BX_MAP_COVARIANT_PIOLA, std::move(B), std::move(dofs), discontinuous);
}

// ---- host helpers shared with other set-up code (hostmath.cuh) ------------------------------------------------
namespace host
{
std::vector<double> polyset_tabulate(int cell, int degree, int nd, const std::vector<double>& x, size_t npts)
{
  return host_polyset_nd(cell, degree, nd, x, npts);
}
void quadrature(int cell, int q, std::vector<double>& x, std::vector<double>& w) { make_quadrature(cell, q, x, w); }
} // namespace host
// lattice::create (cpp/basix/lattice.cpp:636-700) for the lattice types this factory builds its elements from:
// equispaced (0) and gll (1; simplices by the warp method).  Returns the points row-major [npts][tdim].
std::vector<double> create_lattice_points(int cell, int n, int lattice_type, int exterior, int simplex_method)
{
  if (cell < BX_CELL_POINT or cell > BX_CELL_PYRAMID)
    fail(BX_ERR_INVALID_ARG, "Unsupported cell type");
  if (n < 0)
    fail(BX_ERR_INVALID_ARG, "create_lattice: negative size");
  if (lattice_type != 0 and lattice_type != 1)
    fail(BX_ERR_UNSUPPORTED, "create_lattice: equispaced and gll lattices are built natively");
  const bool simplex = cell == BX_CELL_TRIANGLE or cell == BX_CELL_TETRAHEDRON;
  if (lattice_type == 1 and simplex and simplex_method != 1)
    fail(BX_ERR_UNSUPPORTED, "create_lattice: gll lattices on simplices are built with the warp method only");
  if (cell == BX_CELL_PRISM or cell == BX_CELL_PYRAMID)
    fail(BX_ERR_UNSUPPORTED, "create_lattice: prism and pyramid lattices are not built natively");
  return lattice_create(cell, n, lattice_type == 1, exterior != 0);
}
// quadrature::make_quadrature (quadrature.cpp:4960-5020) with the Gauss-Jacobi scheme (quadrature.cpp:215-372): points
// [npts][tdim] and weights of a rule exact for polynomials of degree q.
// quadrature::gauss_jacobi_rule (quadrature.cpp:4955-4965): m points on [0, 1] for the weight (1 - x)^a
void gauss_jacobi_rule_unit(double a, int m, std::vector<double>& x, std::vector<double>& w)
{
  if (m < 1)
    fail(BX_ERR_INVALID_ARG, "gauss_jacobi_rule: at least one point is needed");
  gauss_jacobi(a, m, x, w);
  const double scale = std::pow(2.0, a + 1.0);
  for (int i = 0; i < m; ++i)
  {
    x[i] = (x[i] + 1.0) / 2.0;
    w[i] /= scale;
  }
}

void make_quadrature_gauss_jacobi(int cell, int q, std::vector<double>& x, std::vector<double>& w)
{
  if (cell < BX_CELL_INTERVAL or cell > BX_CELL_PYRAMID)
    fail(BX_ERR_INVALID_ARG, "make_quadrature: unsupported cell type");
  if (q < 0)
    fail(BX_ERR_INVALID_ARG, "make_quadrature: negative degree");
  make_quadrature(cell, q, x, w);
}
} // namespace bx

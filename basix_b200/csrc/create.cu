// Native element construction (host side, runs once per element): basix.create_element for the Lagrange (P)
// family on interval / triangle / tetrahedron / quadrilateral / hexahedron.
//
// This is set-up code, not the accelerated path: it produces the immutable arrays the hot path consumes
// (coefficient matrix, entity dofs, entity transformations) and hands them to element_from_desc().  What is
// computed follows the reference; how it is organised is this library's own:
//   lattice points            cpp/basix/lattice.cpp:42-71 (interval, equispaced / GLL), 201-216 (warp function),
//                             218-280 (quad / hex), 282-350 (triangle), 503-535 + 583-633 (tetrahedron)
//   GLL nodes                 cpp/basix/quadrature.cpp:470-510 (there: Golub-Welsch with fixed end nodes; here:
//                             Newton on (1 - x^2) P'_n, the same nodes to rounding)
//   dof points per entity     cpp/basix/e-lagrange.cpp:457-666 (vertices, then edges, faces, interior; the
//                             sub-entity lattice mapped by v0 + sum_k (v_{k+1} - v0) t_k)
//   coefficient matrix        cpp/basix/finite-element.cpp:84-175, 1105-1119: C = (B D^T)^-1 B with B = I, i.e.
//                             C P(x_dofs) = I
//   entity transformations    cpp/basix/dof-transformations.cpp:71-330 (the cell-level maps that reverse /
//                             rotate / reflect the first sub-entity of each type), 340-470 (tabulate the
//                             element basis at the mapped dof points of that entity)
//   discontinuous variant     cpp/basix/finite-element.cpp make_discontinuous: every dof moves to the interior
// The orthonormal polynomial sets are evaluated with the SAME recurrence code the device kernels use
// (polyset.cuh, compiled for the host here).
#include "element.cuh"
#include "polyset.cuh"

#include <algorithm>
#include <cmath>
#include <functional>
#include <numeric>

namespace bx
{
namespace
{
using Vec = std::vector<double>;

// ---- host polyset tabulate, nd = 0: P[psize][npts] ---------------------------------------------------------
Vec host_polyset(int cell, int degree, const Vec& x, size_t npts)
{
  const int tdim = cell_tdim(cell);
  const int psize = polyset_dim(cell, BX_POLYSET_STANDARD, degree);
  Vec P(size_t(psize) * npts, 0.0);
  const int n1 = degree + 1;
  Vec L(size_t(3) * n1);
  for (size_t i = 0; i < npts; ++i)
  {
    PointView<double> v{P.data() + i, npts, psize};
    switch (cell)
    {
    case BX_CELL_POINT: P[i] = 1.0; break;
    case BX_CELL_INTERVAL: polyset_interval<double>(v, degree, 0, x[i]); break;
    case BX_CELL_TRIANGLE: polyset_triangle<double>(v, degree, 0, x[2 * i], x[2 * i + 1]); break;
    case BX_CELL_TETRAHEDRON: polyset_tetrahedron<double>(v, degree, 0, x[3 * i], x[3 * i + 1], x[3 * i + 2]); break;
    case BX_CELL_QUADRILATERAL:
    case BX_CELL_HEXAHEDRON:
    {
      for (int a = 0; a < tdim; ++a)
        legendre_table<double>(L.data() + size_t(a) * n1, degree, 0, x[tdim * i + a]);
      if (tdim == 2)
      {
        for (int px = 0; px < n1; ++px)
          for (int py = 0; py < n1; ++py)
            P[size_t(px * n1 + py) * npts + i] = L[px] * L[n1 + py];
      }
      else
      {
        for (int px = 0; px < n1; ++px)
          for (int py = 0; py < n1; ++py)
            for (int pz = 0; pz < n1; ++pz)
              P[size_t((px * n1 + py) * n1 + pz) * npts + i] = L[px] * L[n1 + py] * L[2 * n1 + pz];
      }
      break;
    }
    default: fail(BX_ERR_UNSUPPORTED, "create_element: cell type not supported by the native factory");
    }
  }
  return P;
}

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

  static const double none = 0.0;
  bx_element_desc desc{};
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
} // namespace bx

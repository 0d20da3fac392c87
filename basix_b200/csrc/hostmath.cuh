// Host-side numerical helpers used by element set-up code (defined in create.cu).
#pragma once

#include <cstddef>
#include <vector>

namespace bx
{
namespace host
{
// Orthonormal polynomial set with derivatives on the host: P[nderivs][psize][npts], the same recurrence code as
// the device kernels (polyset.cuh).  x[npts][tdim].
std::vector<double> polyset_tabulate(int cell, int degree, int nd, const std::vector<double>& x, size_t npts);
// Gauss-Jacobi quadrature exact to polynomial degree q on interval / triangle / tetrahedron: x[n][tdim], w[n].
void quadrature(int cell, int q, std::vector<double>& x, std::vector<double>& w);
} // namespace host
} // namespace bx

// Which (cell, degree, nderiv) combinations have a fused-kernel instantiation, and their launchers.
// One translation unit per cell type (fused_interval.cu, fused_triangle.cu, fused_tetrahedron.cu) keeps the
// heavily unrolled instantiations compiling in parallel.
#pragma once

#include "element.cuh"

namespace bx
{
namespace fused
{
// Each returns false when (degree, nd) is not on its menu; otherwise launches and returns true.
bool launch_interval(const bx_element& e, int nd, const double* x, size_t npts, double* basis, cudaStream_t s,
                     bool dry_run);
bool launch_triangle(const bx_element& e, int nd, const double* x, size_t npts, double* basis, cudaStream_t s,
                     bool dry_run);
bool launch_tetrahedron(const bx_element& e, int nd, const double* x, size_t npts, double* basis, cudaStream_t s,
                        bool dry_run);
} // namespace fused
} // namespace bx

// Fused tabulate instantiations for the triangle (see tabulate_fused.cuh).
#include "fused_menu.cuh"
#include "tabulate_fused.cuh"

namespace bx
{
namespace fused
{
#define BX_CASE(DEG, ND)                                                                                   \
  if (e.superdegree == DEG and nd == ND)                                                                   \
  {                                                                                                        \
    if (!dry_run)                                                                                          \
      launch<2, DEG, ND>(e, x, npts, basis, s);                                                            \
    return true;                                                                                           \
  }

bool launch_triangle(const bx_element& e, int nd, const double* x, size_t npts, double* basis, cudaStream_t s,
                     bool dry_run)
{
  BX_CASE(3, 0) BX_CASE(3, 1) BX_CASE(3, 2) BX_CASE(3, 3)
  BX_CASE(4, 0) BX_CASE(4, 1) BX_CASE(4, 2) BX_CASE(4, 3)
  BX_CASE(5, 0) BX_CASE(5, 1) BX_CASE(5, 2)
  BX_CASE(6, 0) BX_CASE(6, 1) BX_CASE(6, 2)
  BX_CASE(7, 0) BX_CASE(7, 1) BX_CASE(7, 2)
  BX_CASE(8, 0) BX_CASE(8, 1) BX_CASE(8, 2)
  return false;
}
} // namespace fused
} // namespace bx

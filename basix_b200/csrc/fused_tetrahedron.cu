// Fused tabulate instantiations for the tetrahedron (see tabulate_fused.cuh).
// Menu limit: the consumer accumulators ceil(NDER/2) x ceil(K/8) x 2 doubles must fit beside the
// fragments in 168 registers, and Cs + a >= 3-stage ring in 200 KB of shared memory.
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
      launch<3, DEG, ND>(e, x, npts, basis, s);                                                            \
    return true;                                                                                           \
  }

bool launch_tetrahedron(const bx_element& e, int nd, const double* x, size_t npts, double* basis, cudaStream_t s,
                        bool dry_run)
{
  BX_CASE(2, 0) BX_CASE(2, 1) BX_CASE(2, 2) BX_CASE(2, 3)
  BX_CASE(3, 0) BX_CASE(3, 1) BX_CASE(3, 2) BX_CASE(3, 3)
  BX_CASE(4, 0) BX_CASE(4, 1) BX_CASE(4, 2)
  BX_CASE(5, 0) BX_CASE(5, 1) BX_CASE(5, 2)
  BX_CASE(6, 0) BX_CASE(6, 1)
  return false;
}
} // namespace fused
} // namespace bx

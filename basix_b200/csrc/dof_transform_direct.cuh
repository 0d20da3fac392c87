// Direct (register) form of the matrix DOF transformation for the layout the BASELINE configuration uses: float64,
// left operators, block size 1, every transformed entity starting on an even entry of 16-byte aligned cells
// (transform_data, finite-element.h:1762-1837; apply_matrix, precompute.h:251-321).
//
// One thread per (cell, sub-entity) item, sub-entity fastest: the thread fetches ITS entity's values straight from
// global memory with 16-byte loads (consecutive lanes = consecutive entities of consecutive cells, so a warp reads a
// few hundred contiguous bytes per cell), multiplies by the net operator of the entity's cell_info state (shared
// memory, inside the non-zero column band of each row) and stores the entity back with 16-byte stores.  Against the
// tensor-map tile kernel (dof_transform_tma.cuh) the request granularity is the 32-byte sector instead of the
// 128-byte line the strided TMA box fetches, and entities whose state is the identity are neither read nor written.
// Measured (RT4 tetrahedron, 4 M cells, ncu): DRAM traffic 403 + 278 = 681 B/cell against 436 + 325 = 761 B/cell for
// the tile kernel, 582 us against 597 us; 50 M cells: 7.33 ms either way.  Two mechanisms with different traffic and
// the same time: the strided in-place pattern (320 of every 560 bytes, read and written) is paced by the DRAM pages
// it opens, not by the bytes it uses.  The tile kernel stays the default; this one is selected with
// BX_DOF_TUNE="direct=1" (development knob, covered by tests/test_gpu_parity.py::test_T_apply_direct_kernel).
#pragma once

#include "dof_transform_tma.cuh"

namespace bx
{
namespace direct
{
constexpr int kThreads = 256;

struct Params
{
  size_t nc;
  size_t cell_size; // entries per cell
  int nslots, table_size, band_size;
  const uint32_t* cell_info;
  const int4* slot_info; // {shift, mask, dim, gather base}
  const int2* slot_mat;  // {first dof, matrix base of the slot's class}
  const int* slot_band;  // band table base of the slot's class
  const int2* band;      // per (class, row): [first, last) non-zero column
  const double* mats;    // operator kind offset applied
};

template <int REG>
__global__ void __launch_bounds__(kThreads) dof_matrix_direct_kernel(double* __restrict__ data, const Params p)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_mats = reinterpret_cast<double*>(smem_raw);
  int2* s_band = reinterpret_cast<int2*>(s_mats + p.table_size);
  int4* s_info = reinterpret_cast<int4*>(s_band + p.band_size + (p.band_size & 1));
  int4* s_slot = s_info + p.nslots; // {first, matrix base, band base, 0}
  for (int i = threadIdx.x; i < p.table_size; i += kThreads)
    s_mats[i] = p.mats[i];
  for (int i = threadIdx.x; i < p.band_size; i += kThreads)
    s_band[i] = p.band[i];
  for (int i = threadIdx.x; i < p.nslots; i += kThreads)
  {
    s_info[i] = p.slot_info[i];
    const int2 sm = p.slot_mat[i];
    s_slot[i] = make_int4(sm.x, sm.y, p.slot_band[i], 0);
  }
  __syncthreads();

  const size_t nitems = p.nc * size_t(p.nslots);
  const size_t stride = size_t(gridDim.x) * kThreads;
  for (size_t it = size_t(blockIdx.x) * kThreads + threadIdx.x; it < nitems; it += stride)
  {
    // items of one cell are consecutive: it = cell * nslots + s (the per-CTA offset keeps the 32-bit division exact)
    const size_t cell = p.nslots == 4 ? it >> 2 : it / size_t(p.nslots);
    const int s = int(it - cell * size_t(p.nslots));
    const int4 si = s_info[s];
    const unsigned state = (__ldg(p.cell_info + cell) >> si.x) & unsigned(si.y);
    if (state == 0)
      continue;
    const int dim = si.z;
    const int4 sl = s_slot[s];
    const int rs = mat_row_stride(dim);
    const double* M = s_mats + sl.y + state * mat_state_stride(dim);
    const int2* band = s_band + sl.z;
    double* ent = data + cell * p.cell_size + sl.x;
    double2* x = reinterpret_cast<double2*>(ent);
    double v[REG];
#pragma unroll
    for (int jj = 0; jj < REG / 2; ++jj)
    {
      if (2 * jj + 1 < dim)
      {
        const double2 t = x[jj];
        v[2 * jj] = t.x;
        v[2 * jj + 1] = t.y;
      }
      else
      {
        v[2 * jj] = 2 * jj < dim ? ent[2 * jj] : 0.0;
        v[2 * jj + 1] = 0.0;
      }
    }
    for (int i0 = 0; i0 < dim; i0 += 2)
    {
      const bool two = i0 + 1 < dim;
      const int2 b0 = band[i0], b1 = two ? band[i0 + 1] : make_int2(0, 0);
      const int jlo = two ? min(b0.x, b1.x) : b0.x, jhi = two ? max(b0.y, b1.y) : b0.y;
      const double2* r0p = reinterpret_cast<const double2*>(M + i0 * rs);
      const double2* r1p = reinterpret_cast<const double2*>(M + (two ? i0 + 1 : i0) * rs);
      double a0 = 0.0, a1 = 0.0, c0a = 0.0, c1a = 0.0;
#pragma unroll
      for (int jj = 0; jj < REG / 2; ++jj)
        if (2 * jj + 1 >= jlo and 2 * jj < jhi)
        {
          const double2 m0 = r0p[jj], m1 = r1p[jj];
          a0 = fma(m0.x, v[2 * jj], a0);
          a1 = fma(m0.y, v[2 * jj + 1], a1);
          c0a = fma(m1.x, v[2 * jj], c0a);
          c1a = fma(m1.y, v[2 * jj + 1], c1a);
        }
      if (two)
        x[i0 >> 1] = make_double2(a0 + a1, c0a + c1a);
      else
        ent[i0] = a0 + a1;
    }
  }
}
} // namespace direct
} // namespace bx

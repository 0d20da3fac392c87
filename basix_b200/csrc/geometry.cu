// Cell geometry on the device and the fused "physical basis" kernel.
//
//   geometry_kernel     J, detJ, K of every (cell, point) from the cell's node coordinates and the derivatives of the
//                       coordinate element -- what every caller of push_forward / pull_back computes first.  The
//                       reference's own analogue is the affine case, cell::entity_jacobians (cpp/basix/cell.cpp:662-689:
//                       J(:, j) = x[1 + j] - x[0]), which this kernel reproduces bit for bit with the derivatives of the
//                       P1 coordinate element.
//   physical_basis_kernel
//                       tabulate-once -> T_apply -> push_forward in one pass over the cells (the three calls a caller
//                       makes per cell, finite-element.cpp:1971-2005 and finite-element.h:1839-1854): a warp owns a
//                       cell, computes J / detJ / K of the affine cell from its vertex coordinates ON CHIP, picks for
//                       every dof row the reference values already transformed for the state of the row's sub-entity
//                       in this cell's cell_info (the 8 possible states are transformed once per call by the library's
//                       own T_apply kernel), maps the row (maps.h:73-124) and writes only the final
//                       [cell][point][dof][value] rows: 100 B in and rows * 24 B out per cell instead of the ~3.3 KB
//                       the three separate calls move.
//
// Compiled with -fmad=false: the arithmetic below (sums in index order, one rounding per operation, IEEE division) is
// restated operation for operation by oracle/restate.py (geometry_*), and the map arithmetic is that of maps.cu, so
// the fused result equals push_forward(T_apply(U), geometry(coords)) bit for bit.
#include "element.cuh"

#include <algorithm>
#include <vector>

namespace bx
{
template <typename T>
void dof_transform_device(const bx_element& e, int op, T* data, size_t nc, size_t cell_size, int n,
                          const uint32_t* cell_info, cudaStream_t s);
template <typename T>
void tabulate_device(const bx_element& e, int nd, const T* x, size_t npts, T* basis, size_t out_pts, cudaStream_t s);

namespace geo
{
constexpr int kThreads = 256;
constexpr int kBatch = 4; // cells per warp iteration of the fused kernel

// Unscaled inverse and its divisor: K = adj / div, detJ = div (square) or sqrt(div) (non-square, div = det(J^T J)).
//   1 x 1, 2 x 2, 3 x 3: adjugate and determinant (first-row cofactor expansion)
//   G > TD:              adj = adj(J^T J) J^T, div = det(J^T J)   (Moore-Penrose left inverse)
template <typename T, int G, int TD>
__device__ __forceinline__ void adjugate(const T (&J)[G][TD], T (&adj)[TD][G], T& div)
{
  if constexpr (G == TD and TD == 1)
  {
    adj[0][0] = T(1);
    div = J[0][0];
  }
  else if constexpr (G == TD and TD == 2)
  {
    adj[0][0] = J[1][1];
    adj[0][1] = -J[0][1];
    adj[1][0] = -J[1][0];
    adj[1][1] = J[0][0];
    div = J[0][0] * J[1][1] - J[0][1] * J[1][0];
  }
  else if constexpr (G == TD and TD == 3)
  {
    adj[0][0] = J[1][1] * J[2][2] - J[1][2] * J[2][1];
    adj[0][1] = J[0][2] * J[2][1] - J[0][1] * J[2][2];
    adj[0][2] = J[0][1] * J[1][2] - J[0][2] * J[1][1];
    adj[1][0] = J[1][2] * J[2][0] - J[1][0] * J[2][2];
    adj[1][1] = J[0][0] * J[2][2] - J[0][2] * J[2][0];
    adj[1][2] = J[0][2] * J[1][0] - J[0][0] * J[1][2];
    adj[2][0] = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    adj[2][1] = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    adj[2][2] = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    div = J[0][0] * adj[0][0] + J[0][1] * adj[1][0] + J[0][2] * adj[2][0];
  }
  else
  {
    static_assert(G > TD and TD <= 2, "Jacobian shape");
    T M[TD][TD]; // J^T J
#pragma unroll
    for (int a = 0; a < TD; ++a)
#pragma unroll
      for (int b = 0; b < TD; ++b)
      {
        T acc = 0;
#pragma unroll
        for (int g = 0; g < G; ++g)
          acc += J[g][a] * J[g][b];
        M[a][b] = acc;
      }
    if constexpr (TD == 1)
    {
      div = M[0][0];
#pragma unroll
      for (int g = 0; g < G; ++g)
        adj[0][g] = J[g][0];
    }
    else
    {
      div = M[0][0] * M[1][1] - M[0][1] * M[1][0];
#pragma unroll
      for (int g = 0; g < G; ++g)
      {
        adj[0][g] = M[1][1] * J[g][0] - M[0][1] * J[g][1];
        adj[1][g] = M[0][0] * J[g][1] - M[1][0] * J[g][0];
      }
    }
  }
}

template <typename T, int G, int TD>
__device__ __forceinline__ T det_from_div(T div)
{
  if constexpr (G == TD)
    return div;
  else
    return sqrt(div);
}

// coords[nc][nv][G], dphi[TD][npts][nv] (= tabulate(1, X)[1 : 1 + TD] of the scalar coordinate element)
//   -> J[nc][npts][G][TD], detJ[nc][npts], K[nc][npts][TD][G]; null outputs are skipped.
template <typename T, int G, int TD>
__global__ void __launch_bounds__(kThreads)
    geometry_kernel(const T* __restrict__ coords, const T* __restrict__ dphi, size_t nc, int nv, int npts, int dphi_in_smem,
                    T* __restrict__ Jout, T* __restrict__ detout, T* __restrict__ Kout)
{
  extern __shared__ __align__(16) unsigned char geo_smem[];
  const T* d = dphi; // null: affine simplex cells (nv = TD + 1, one point)
  if (dphi_in_smem and dphi)
  {
    T* s = reinterpret_cast<T*>(geo_smem);
    for (int i = threadIdx.x; i < TD * npts * nv; i += kThreads)
      s[i] = dphi[i];
    __syncthreads();
    d = s;
  }
  const size_t total = nc * size_t(npts), stride = size_t(gridDim.x) * kThreads;
  for (size_t t = size_t(blockIdx.x) * kThreads + threadIdx.x; t < total; t += stride)
  {
    const size_t c = t / size_t(npts);
    const int p = int(t - c * size_t(npts));
    const T* xc = coords + c * size_t(nv) * G;
    T J[G][TD];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int a = 0; a < TD; ++a)
        J[g][a] = d ? T(0) : xc[(1 + a) * G + g] - xc[g]; // no derivatives given: affine simplex, cell.cpp:680-686
    for (int v = 0; d and v < nv; ++v)
    {
      T dv[TD];
#pragma unroll
      for (int a = 0; a < TD; ++a)
        dv[a] = d[(size_t(a) * npts + p) * nv + v];
#pragma unroll
      for (int g = 0; g < G; ++g)
      {
        const T xg = xc[v * G + g];
#pragma unroll
        for (int a = 0; a < TD; ++a)
          J[g][a] += xg * dv[a];
      }
    }
    T adj[TD][G], div;
    adjugate<T, G, TD>(J, adj, div);
    if (Jout)
    {
#pragma unroll
      for (int g = 0; g < G; ++g)
#pragma unroll
        for (int a = 0; a < TD; ++a)
          Jout[t * (G * TD) + g * TD + a] = J[g][a];
    }
    if (detout)
      detout[t] = det_from_div<T, G, TD>(div);
    if (Kout)
    {
#pragma unroll
      for (int a = 0; a < TD; ++a)
#pragma unroll
        for (int g = 0; g < G; ++g)
          Kout[t * (G * TD) + a * G + g] = adj[a][g] / div;
    }
  }
}

// ---- fused physical basis ------------------------------------------------------------------------------------
struct FusedParams
{
  size_t nc;
  int rows;         // npts * ndofs rows per cell
  int ndofs;
  int tab_stride;   // entries between the tables of consecutive states
  int nstates;      // 8, or 1 when nothing is transformed (no cell_info, identity transformations)
  size_t out_pitch; // entries between consecutive cells of `out` (>= rows * OUT_VS: a chunk of a larger point set)
  const int16_t* dof_slot;  // [ndofs] owning slot or -1
  const int4* slot_info;    // [nslots] {shift, mask, dim, _}
  const uint32_t* cell_info; // [nc] or null
};

// MAP: BX_MAP_IDENTITY (value size VS, no matrix), BX_MAP_COVARIANT_PIOLA, BX_MAP_CONTRAVARIANT_PIOLA.
// GEOM: J / detJ / K of the affine simplex cell from coords[nc][TD + 1][G]; otherwise A / detJ are read per cell:
// A = K[nc][TD][G] (covariant) or J[nc][G][TD] (contravariant).
// tab[8][rows][IN_VS]: the reference values of every row, transformed for state 0..7 of the row's sub-entity.
template <typename T, int MAP, int G, int TD, bool GEOM, int PASSES, int VS>
__global__ void __launch_bounds__(kThreads)
    physical_basis_kernel(const T* __restrict__ tab, const T* __restrict__ coords, const T* __restrict__ A,
                          const T* __restrict__ detJ, const FusedParams p, T* __restrict__ out)
{
  constexpr int IN_VS = MAP == BX_MAP_IDENTITY ? VS : TD;
  constexpr int OUT_VS = MAP == BX_MAP_IDENTITY ? VS : G;
  constexpr int MA = G * TD;
  constexpr int NX = (TD + 1) * G; // coordinates of an affine simplex cell
  extern __shared__ __align__(16) unsigned char fused_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows = p.rows, E = rows * OUT_VS;
  T* s_tab = reinterpret_cast<T*>(fused_smem); // [nstates][tab_stride]
  constexpr int kExtra = kBatch * (MA + 1) + (GEOM ? kBatch * NX : 0);
  T* slice = s_tab + p.nstates * p.tab_stride + warp * (E + kExtra);
  T* sm = slice + E;                 // [kBatch][MA] matrices, then [kBatch] determinants
  T* sx = sm + kBatch * (MA + 1);    // GEOM: [kBatch][NX] vertex coordinates
  for (int i = threadIdx.x; i < p.nstates * p.tab_stride; i += kThreads)
    s_tab[i] = tab[i];
  // per-lane descriptors of the rows this lane owns: first cell_info bit and mask of the row's sub-entity
  int shift[PASSES], mask[PASSES];
#pragma unroll
  for (int q = 0; q < PASSES; ++q)
  {
    const int r = lane + 32 * q;
    shift[q] = mask[q] = 0;
    if (r < rows and p.cell_info)
    {
      const int slot = p.dof_slot[r % p.ndofs];
      if (slot >= 0)
      {
        const int4 si = p.slot_info[slot];
        shift[q] = si.x;
        mask[q] = si.y;
      }
    }
  }
  __syncthreads();

  const size_t warps = size_t(gridDim.x) * (kThreads / 32);
  for (size_t c0 = (size_t(blockIdx.x) * (kThreads / 32) + warp) * kBatch; c0 < p.nc; c0 += warps * kBatch)
  {
    const int nb = int(min(size_t(kBatch), p.nc - c0));
    uint32_t my_ci = 0;
    if (p.cell_info and lane < nb)
      my_ci = p.cell_info[c0 + lane];
    if constexpr (GEOM)
    {
      for (int q = lane; q < nb * NX; q += 32)
        sx[q] = coords[c0 * NX + q];
      __syncwarp();
      if (lane < nb)
      {
        // J(:, a) = x[1 + a] - x[0]: the P1 coordinate element's derivatives are -1 at vertex 0 and +1 at vertex
        // 1 + a, so the general sum  0 + x0 * (-1) + ... + x[1 + a] * 1  reduces to this difference exactly
        const T* x = sx + lane * NX;
        T J[G][TD];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int a = 0; a < TD; ++a)
            J[g][a] = x[(1 + a) * G + g] - x[g];
        T adj[TD][G], div;
        adjugate<T, G, TD>(J, adj, div);
        T* m = sm + lane * MA;
        if constexpr (MAP == BX_MAP_COVARIANT_PIOLA)
        {
#pragma unroll
          for (int a = 0; a < TD; ++a)
#pragma unroll
            for (int g = 0; g < G; ++g)
              m[a * G + g] = adj[a][g] / div;
        }
        else if constexpr (MAP == BX_MAP_CONTRAVARIANT_PIOLA)
        {
#pragma unroll
          for (int g = 0; g < G; ++g)
#pragma unroll
            for (int a = 0; a < TD; ++a)
              m[g * TD + a] = J[g][a];
          sm[kBatch * MA + lane] = det_from_div<T, G, TD>(div);
        }
      }
    }
    else if constexpr (MAP != BX_MAP_IDENTITY)
    {
      for (int q = lane; q < nb * MA; q += 32)
        sm[q] = A[c0 * MA + q];
      if constexpr (MAP == BX_MAP_CONTRAVARIANT_PIOLA)
        if (lane < nb)
          sm[kBatch * MA + lane] = detJ[c0 + lane];
    }
    __syncwarp();
    for (int bi = 0; bi < nb; ++bi)
    {
      const uint32_t ci = __shfl_sync(0xffffffffu, my_ci, bi);
      [[maybe_unused]] T a[MAP == BX_MAP_COVARIANT_PIOLA ? TD : G][MAP == BX_MAP_COVARIANT_PIOLA ? G : TD];
      [[maybe_unused]] T det = T(1);
      if constexpr (MAP == BX_MAP_COVARIANT_PIOLA)
      {
#pragma unroll
        for (int i = 0; i < TD; ++i)
#pragma unroll
          for (int j = 0; j < G; ++j)
            a[i][j] = sm[bi * MA + i * G + j];
      }
      else if constexpr (MAP == BX_MAP_CONTRAVARIANT_PIOLA)
      {
#pragma unroll
        for (int i = 0; i < G; ++i)
#pragma unroll
          for (int j = 0; j < TD; ++j)
            a[i][j] = sm[bi * MA + i * TD + j];
        det = sm[kBatch * MA + bi];
      }
#pragma unroll
      for (int q = 0; q < PASSES; ++q)
      {
        const int r = lane + 32 * q;
        if (r < rows)
        {
          const int st = int(ci >> shift[q]) & mask[q];
          const T* u = s_tab + st * p.tab_stride + r * IN_VS;
          if constexpr (MAP == BX_MAP_IDENTITY)
          {
#pragma unroll
            for (int i = 0; i < VS; ++i)
              slice[r * VS + i] = u[i];
          }
          else
          {
            T uu[IN_VS];
#pragma unroll
            for (int k = 0; k < IN_VS; ++k)
              uu[k] = u[k];
#pragma unroll
            for (int i = 0; i < OUT_VS; ++i)
            {
              T acc = 0;
              if constexpr (MAP == BX_MAP_COVARIANT_PIOLA)
              {
                // out[i] = sum_k K[k][i] in[k]   (maps.h:73-94)
#pragma unroll
                for (int k = 0; k < TD; ++k)
                  acc += a[k][i] * uu[k];
              }
              else
              {
                // out[i] = (sum_k J[i][k] in[k]) / det   (maps.h:100-124)
#pragma unroll
                for (int k = 0; k < TD; ++k)
                  acc += a[i][k] * uu[k];
                acc = acc / det;
              }
              slice[r * OUT_VS + i] = acc;
            }
          }
        }
      }
      __syncwarp();
      T* o = out + (c0 + bi) * p.out_pitch;
      for (int e = lane; e < E; e += 32)
        o[e] = slice[e];
      __syncwarp(); // the slice is overwritten by the next cell
    }
  }
}

// ---- launchers -------------------------------------------------------------------------------------------------
template <typename K>
int persistent_grid(K kern, size_t want, size_t smem)
{
  if (smem > 48 * 1024)
    BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 1;
  BX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  if (per_sm < 1)
    fail(BX_ERR_UNSUPPORTED, "geometry: the kernel does not fit in shared memory");
  const size_t cap = size_t(sm_count()) * size_t(per_sm);
  return int(std::max<size_t>(1, std::min(want, cap)));
}

template <typename T, int G, int TD>
void launch_geometry(const T* coords, const T* dphi, size_t nc, int nv, int npts, T* J, T* detJ, T* K, cudaStream_t s)
{
  auto kern = geometry_kernel<T, G, TD>;
  const size_t tbl = size_t(TD) * npts * nv * sizeof(T);
  const int in_smem = tbl <= 96 * 1024;
  const size_t smem = in_smem ? tbl : 0;
  const int grid = persistent_grid(kern, ceil_div(nc * size_t(npts), size_t(kThreads)), smem);
  kern<<<grid, kThreads, smem, s>>>(coords, dphi, nc, nv, npts, in_smem, J, detJ, K);
  BX_LAUNCHED();
}

template <typename T, int MAP, int G, int TD, bool GEOM, int VS>
void launch_fused(const T* tab, const T* coords, const T* A, const T* detJ, const FusedParams& p, T* out, cudaStream_t s)
{
  constexpr int OUT_VS = MAP == BX_MAP_IDENTITY ? VS : G;
  constexpr int kExtra = kBatch * (G * TD + 1) + (GEOM ? kBatch * (TD + 1) * G : 0);
  const size_t smem = (size_t(p.nstates) * p.tab_stride + size_t(kThreads / 32) * (size_t(p.rows) * OUT_VS + kExtra)) * sizeof(T);
  if (smem > 200 * 1024)
    fail(BX_ERR_UNSUPPORTED, "physical_basis: the point chunk does not fit in shared memory");
  const size_t want = ceil_div(p.nc, size_t(kThreads / 32) * kBatch);
  auto go = [&](auto kern)
  {
    const int grid = persistent_grid(kern, want, smem);
    kern<<<grid, kThreads, smem, s>>>(tab, coords, A, detJ, p, out);
    BX_LAUNCHED();
  };
  const int passes = (p.rows + 31) / 32;
  if (passes <= 1)
    go(physical_basis_kernel<T, MAP, G, TD, GEOM, 1, VS>);
  else if (passes <= 2)
    go(physical_basis_kernel<T, MAP, G, TD, GEOM, 2, VS>);
  else if (passes <= 4)
    go(physical_basis_kernel<T, MAP, G, TD, GEOM, 4, VS>);
  else
    go(physical_basis_kernel<T, MAP, G, TD, GEOM, 8, VS>);
}

template <typename T, int MAP, bool GEOM>
void launch_fused_shape(int gdim, int tdim, int vs, const T* tab, const T* coords, const T* A, const T* detJ,
                        const FusedParams& p, T* out, cudaStream_t s)
{
  if constexpr (MAP == BX_MAP_IDENTITY)
  {
    // no matrix: the shape of J is irrelevant, the value size is the template parameter
    switch (vs)
    {
    case 1: launch_fused<T, MAP, 1, 1, false, 1>(tab, coords, A, detJ, p, out, s); return;
    case 2: launch_fused<T, MAP, 1, 1, false, 2>(tab, coords, A, detJ, p, out, s); return;
    case 3: launch_fused<T, MAP, 1, 1, false, 3>(tab, coords, A, detJ, p, out, s); return;
    default: fail(BX_ERR_UNSUPPORTED, "physical_basis: identity map with value size above 3");
    }
  }
  else
  {
    if (gdim == 2 and tdim == 2)
      launch_fused<T, MAP, 2, 2, GEOM, 1>(tab, coords, A, detJ, p, out, s);
    else if (gdim == 3 and tdim == 3)
      launch_fused<T, MAP, 3, 3, GEOM, 1>(tab, coords, A, detJ, p, out, s);
    else if (gdim == 3 and tdim == 2)
      launch_fused<T, MAP, 3, 2, GEOM, 1>(tab, coords, A, detJ, p, out, s);
    else
      fail(BX_ERR_UNSUPPORTED, "physical_basis: supported cell embeddings are 2D in 2D, 3D in 3D and 2D in 3D");
  }
}

// states 0..7 of every sub-entity at once: "cell" st of the T_apply batch has every face in state st and every edge
// in state st & 1
__global__ void state_info_kernel(uint32_t* ci, int npts, const uint32_t w0, const uint32_t w1, const uint32_t w2,
                                  const uint32_t w3, const uint32_t w4, const uint32_t w5, const uint32_t w6, const uint32_t w7)
{
  const uint32_t w[8] = {w0, w1, w2, w3, w4, w5, w6, w7};
  for (int i = threadIdx.x; i < 8 * npts; i += blockDim.x)
    ci[i] = w[i / npts];
}

template <typename T>
__global__ void replicate_kernel(const T* __restrict__ in, size_t period, int copies, T* __restrict__ out)
{
  const size_t n = period * size_t(copies), stride = size_t(gridDim.x) * blockDim.x;
  for (size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x; t < n; t += stride)
    out[t] = in[t % period];
}
} // namespace geo

// ---- device-pointer entry points -------------------------------------------------------------------------------
template <typename T>
void geometry_device(const T* coords, const T* dphi, size_t nc, int nv, int gdim, int tdim, int npts, T* J, T* detJ, T* K,
                     cudaStream_t s)
{
  if (nv < 1 or npts < 1)
    fail(BX_ERR_INVALID_ARG, "geometry: need at least one node and one point");
  if (tdim < 1 or tdim > 3 or gdim < tdim or gdim > 3)
    fail(BX_ERR_INVALID_ARG, "geometry: need 1 <= tdim <= gdim <= 3");
  if (nc == 0)
    return;
  if (!coords)
    fail(BX_ERR_INVALID_ARG, "geometry: null pointer");
  if (!dphi and (nv != tdim + 1 or npts != 1))
    fail(BX_ERR_INVALID_ARG, "geometry: without coordinate-element derivatives the cells are affine simplices (tdim + 1 "
                             "vertices, one point)");
  if (nc * size_t(npts) / size_t(npts) != nc)
    fail(BX_ERR_INVALID_ARG, "geometry: too many (cell, point) pairs");
  switch (gdim * 4 + tdim)
  {
  case 1 * 4 + 1: geo::launch_geometry<T, 1, 1>(coords, dphi, nc, nv, npts, J, detJ, K, s); break;
  case 2 * 4 + 1: geo::launch_geometry<T, 2, 1>(coords, dphi, nc, nv, npts, J, detJ, K, s); break;
  case 3 * 4 + 1: geo::launch_geometry<T, 3, 1>(coords, dphi, nc, nv, npts, J, detJ, K, s); break;
  case 2 * 4 + 2: geo::launch_geometry<T, 2, 2>(coords, dphi, nc, nv, npts, J, detJ, K, s); break;
  case 3 * 4 + 2: geo::launch_geometry<T, 3, 2>(coords, dphi, nc, nv, npts, J, detJ, K, s); break;
  default: geo::launch_geometry<T, 3, 3>(coords, dphi, nc, nv, npts, J, detJ, K, s); break;
  }
}

// Rows of one kernel pass: at most 256 (eight per lane).  Larger point sets are cut into chunks of whole points.
int physical_basis_chunk(const bx_element& e, size_t npts)
{
  if (e.ndofs > 256)
    fail(BX_ERR_UNSUPPORTED, "physical_basis: elements with more than 256 dofs are not supported (use tabulate, T_apply and "
                             "push_forward separately)");
  return int(std::min<size_t>(npts, size_t(256 / e.ndofs)));
}

// U[npts][ndofs][vs] on the device (one slab of reference values for all cells); geometry either as coords[nc][tdim + 1]
// [gdim] (affine simplex cells) or as J / detJ / K per cell; cell_info[nc] or null; out[nc][npts][ndofs][out_vs].
template <typename T>
void physical_basis_device(const bx_element& e, int op, const T* U, size_t npts, const T* coords, const T* J,
                           const T* detJ, const T* K, const uint32_t* cell_info, size_t nc, int gdim, T* out,
                           cudaStream_t s)
{
  const int tdim = e.tdim, vs = e.vs, map = e.map_type;
  if (op < 0 or op > 3)
    fail(BX_ERR_INVALID_ARG, "physical_basis: the operator must be one of T_apply, Tt_apply, Tt_inv_apply, Tinv_apply");
  if (map != BX_MAP_IDENTITY and map != BX_MAP_COVARIANT_PIOLA and map != BX_MAP_CONTRAVARIANT_PIOLA)
    fail(BX_ERR_UNSUPPORTED, "physical_basis: identity, covariant Piola and contravariant Piola maps are fused; use "
                             "tabulate, T_apply and push_forward separately for the double Piola maps");
  if (gdim < tdim or gdim > 3 or tdim < 2)
    fail(BX_ERR_INVALID_ARG, "physical_basis: need 2 <= tdim <= gdim <= 3");
  if (map != BX_MAP_IDENTITY and vs != tdim)
    fail(BX_ERR_INVALID_ARG, "physical_basis: the reference value size of a Piola-mapped element must equal tdim");
  const bool geom = coords != nullptr;
  if (geom and e.cell_type != BX_CELL_TRIANGLE and e.cell_type != BX_CELL_TETRAHEDRON)
    fail(BX_ERR_UNSUPPORTED, "physical_basis: vertex coordinates describe affine simplex cells only; pass J, detJ and K "
                             "for other cells");
  if (!geom and map == BX_MAP_COVARIANT_PIOLA and !K)
    fail(BX_ERR_INVALID_ARG, "physical_basis: the covariant Piola map needs K (or vertex coordinates)");
  if (!geom and map == BX_MAP_CONTRAVARIANT_PIOLA and (!J or !detJ))
    fail(BX_ERR_INVALID_ARG, "physical_basis: the contravariant Piola map needs J and detJ (or vertex coordinates)");
  if (nc == 0 or npts == 0)
    return;
  if (!U or !out)
    fail(BX_ERR_INVALID_ARG, "physical_basis: null pointer");
  const bool transform = cell_info != nullptr and !e.identity and tdim >= 2 and !e.slots.empty();
  if (transform)
    require_device(e);
  const int out_vs = map == BX_MAP_IDENTITY ? vs : gdim;
  const int chunk = physical_basis_chunk(e, npts);
  const size_t slab = size_t(chunk) * e.ndofs * vs; // entries of one state's table
  // scratch: the 8 state tables and their cell_info words (stream-ordered allocation: capturable, no host sync)
  T* tab = nullptr;
  uint32_t* st_ci = nullptr;
  BX_CUDA(cudaMallocAsync(&tab, 8 * slab * sizeof(T), s));
  BX_CUDA(cudaMallocAsync(&st_ci, size_t(8) * chunk * sizeof(uint32_t), s));
  try
  {
    uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (transform)
      for (int st = 0; st < 8; ++st)
        for (const auto& sl : e.slots)
          w[st] |= uint32_t(st & sl.mask) << sl.shift;
    geo::FusedParams p;
    p.nc = nc;
    p.ndofs = e.ndofs;
    p.out_pitch = npts * size_t(e.ndofs) * out_vs;
    p.dof_slot = e.d_dof_slot;
    p.slot_info = e.d_slot_info;
    p.cell_info = transform ? cell_info : nullptr;
    p.nstates = transform ? 8 : 1;
    for (size_t p0 = 0; p0 < npts; p0 += size_t(chunk))
    {
      const int np = int(std::min<size_t>(size_t(chunk), npts - p0));
      const size_t period = size_t(np) * e.ndofs * vs;
      geo::replicate_kernel<T><<<std::max(1, int(std::min<size_t>(64, ceil_div(p.nstates * period, 256)))), 256, 0, s>>>(
          U + p0 * e.ndofs * vs, period, p.nstates, tab);
      BX_LAUNCHED();
      if (transform)
      {
        geo::state_info_kernel<<<1, 256, 0, s>>>(st_ci, np, w[0], w[1], w[2], w[3], w[4], w[5], w[6], w[7]);
        BX_LAUNCHED();
        // the library's own T_apply on the 8 * np (state, point) "cells", block size = value size: identical values to
        // what bx_T_apply gives a caller who transforms the replicated reference values cell by cell
        dof_transform_device<T>(e, op, tab, size_t(8) * np, size_t(e.ndofs) * vs, vs, st_ci, s);
      }
      p.rows = np * e.ndofs;
      p.tab_stride = int(period);
      T* o = out + p0 * size_t(e.ndofs) * out_vs;
      const T* A = map == BX_MAP_COVARIANT_PIOLA ? K : J;
      if (map == BX_MAP_IDENTITY)
        geo::launch_fused_shape<T, BX_MAP_IDENTITY, false>(gdim, tdim, vs, tab, nullptr, nullptr, nullptr, p, o, s);
      else if (map == BX_MAP_COVARIANT_PIOLA)
      {
        if (geom)
          geo::launch_fused_shape<T, BX_MAP_COVARIANT_PIOLA, true>(gdim, tdim, vs, tab, coords, nullptr, nullptr, p, o, s);
        else
          geo::launch_fused_shape<T, BX_MAP_COVARIANT_PIOLA, false>(gdim, tdim, vs, tab, nullptr, A, detJ, p, o, s);
      }
      else
      {
        if (geom)
          geo::launch_fused_shape<T, BX_MAP_CONTRAVARIANT_PIOLA, true>(gdim, tdim, vs, tab, coords, nullptr, nullptr, p, o, s);
        else
          geo::launch_fused_shape<T, BX_MAP_CONTRAVARIANT_PIOLA, false>(gdim, tdim, vs, tab, nullptr, A, detJ, p, o, s);
      }
    }
  }
  catch (...)
  {
    cudaFreeAsync(tab, s);
    cudaFreeAsync(st_ci, s);
    throw;
  }
  BX_CUDA(cudaFreeAsync(tab, s));
  BX_CUDA(cudaFreeAsync(st_ci, s));
}

template void geometry_device<double>(const double*, const double*, size_t, int, int, int, int, double*, double*, double*,
                                      cudaStream_t);
template void geometry_device<float>(const float*, const float*, size_t, int, int, int, int, float*, float*, float*,
                                     cudaStream_t);
template void physical_basis_device<double>(const bx_element&, int, const double*, size_t, const double*, const double*,
                                            const double*, const double*, const uint32_t*, size_t, int, double*,
                                            cudaStream_t);
template void physical_basis_device<float>(const bx_element&, int, const float*, size_t, const float*, const float*,
                                           const float*, const float*, const uint32_t*, size_t, int, float*, cudaStream_t);
} // namespace bx

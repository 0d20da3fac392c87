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
#include "fpdiv.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
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
#ifndef BX_FUSED_PAIR
#define BX_FUSED_PAIR 1
#endif
constexpr int kFusedMaxThreads = 512; // the fused kernel runs 4, 8 or 16 warps per CTA (launch_fused)
constexpr int kPair = BX_FUSED_PAIR; // row passes per trip of the fused kernel's row loop
#ifndef BX_FLAT_PAIR
#define BX_FLAT_PAIR 2
#endif
constexpr int kFlatPair = BX_FLAT_PAIR; // 32-row passes per trip of the flat row walk

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

// ---- affine simplex cells: J(:, a) = x[1 + a] - x[0] -------------------------------------------------------------
// The general kernel above is a thread per (cell, point) with strided 8-byte accesses (12 loads and 19 stores per
// thread, each touching 32 different sectors per warp instruction): 2.4 ms per 20 M tetrahedra, a third of the HBM
// roofline.  Here a warp owns 32 consecutive cells: their coordinates arrive with coalesced loads into an odd-pitch
// shared tile, lane = cell computes from registers, and J, K and detJ of the 32 cells -- one contiguous run of each
// array -- leave as one TMA bulk copy each.
template <typename T, int G, int TD>
__global__ void __launch_bounds__(kThreads)
    geometry_affine_kernel(const T* __restrict__ coords, size_t nc, T* __restrict__ Jout, T* __restrict__ detout,
                           T* __restrict__ Kout, int bulk)
{
  constexpr int NX = (TD + 1) * G, MA = G * TD;
  constexpr int XP = NX | 1;                               // odd pitch: lane-per-cell reads are conflict free
  constexpr int MP = (MA & 1) ? MA : MA + 1;               // same for the staged results ...
  constexpr int kWarpT = ((32 * (XP > MA ? XP : MA) + 3) & ~3) + ((32 * MA + 3) & ~3) + 32; // x / J (dense), K, det
  extern __shared__ __align__(128) unsigned char aff_smem[];
  (void)MP;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  T* sx = reinterpret_cast<T*>(aff_smem) + size_t(warp) * kWarpT; // coordinates [32][XP], then J [32][MA] dense
  T* sK = sx + ((32 * (XP > MA ? XP : MA) + 3) & ~3);              // K [32][MA] dense
  T* sd = sK + ((32 * MA + 3) & ~3);                               // detJ [32]
  const size_t warps = size_t(gridDim.x) * (kThreads / 32);
  bool pending = false;
  for (size_t c0 = (size_t(blockIdx.x) * (kThreads / 32) + warp) * 32; c0 < nc; c0 += warps * 32)
  {
    const int nb = int(min(size_t(32), nc - c0));
    if (pending)
    {
      if (lane < 3)
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      pending = false;
    }
    for (int q = lane; q < nb * NX; q += 32)
      sx[(q / NX) * XP + q % NX] = coords[c0 * NX + q];
    __syncwarp();
    T x[NX];
#pragma unroll
    for (int k = 0; k < NX; ++k)
      x[k] = sx[min(lane, nb - 1) * XP + k];
    __syncwarp(); // the coordinate tile is overwritten by J below
    T J[G][TD], adj[TD][G], div;
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int a = 0; a < TD; ++a)
        J[g][a] = x[(1 + a) * G + g] - x[g];
    adjugate<T, G, TD>(J, adj, div);
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int a = 0; a < TD; ++a)
        sx[lane * MA + g * TD + a] = J[g][a];
#pragma unroll
    for (int a = 0; a < TD; ++a)
#pragma unroll
      for (int g = 0; g < G; ++g)
        sK[lane * MA + a * G + g] = adj[a][g] / div;
    sd[lane] = det_from_div<T, G, TD>(div);
    if (bulk and nb == 32)
    {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      T* dst = lane == 0 ? Jout : lane == 1 ? Kout : detout;
      const T* src = lane == 0 ? sx : lane == 1 ? sK : sd;
      const uint32_t per = lane < 2 ? MA : 1;
      if (lane < 3 and dst)
      {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + c0 * per),
                     "r"(uint32_t(__cvta_generic_to_shared(src))), "r"(uint32_t(32 * per * sizeof(T)))
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      pending = true;
    }
    else
    {
      __syncwarp();
      for (int q = lane; q < nb * MA; q += 32)
      {
        if (Jout)
          Jout[c0 * MA + q] = sx[q];
        if (Kout)
          Kout[c0 * MA + q] = sK[q];
      }
      if (detout and lane < nb)
        detout[c0 + lane] = sd[lane];
      __syncwarp();
    }
  }
  if (pending and lane < 3)
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- fused physical basis ------------------------------------------------------------------------------------
struct FusedParams
{
  size_t nc;
  int rows;         // npts * ndofs rows per cell
  int ndofs;
  int tab_stride;   // entries between the tables of consecutive states
  int nstates;      // 8, or 1 when nothing is transformed (no cell_info, identity transformations)
  int batch;        // cells per slice (one TMA bulk copy-out each)
  int nbuf;         // slices per warp (1 or 2)
  int bulk;         // the rows of `batch` consecutive cells are one 16-byte aligned run of `out`: TMA bulk copy-out
  int vec_coords;   // coords is 16-byte aligned: 16-byte loads
  int ppc;          // points per cell when the kernel's "cells" are (cell, point) items with their own matrix (else 1)
  uint32_t rows_magic; // flat row walk: floor(2^32 / rows) + 1 (0 when rows == 1): cell = umulhi(f, rows_magic) for f < 2^13
  size_t out_pitch; // entries between consecutive cells of `out` (>= rows * OUT_VS: a chunk of a larger point set)
  const int16_t* dof_slot;  // [ndofs] owning slot or -1
  const int4* slot_info;    // [nslots] {shift, mask, dim, _}
  const uint32_t* cell_info; // [nc] or null
};

using fpdiv::div_by;
using fpdiv::safe_reciprocal;

// ---- any coordinate element: the warp form of geometry_kernel --------------------------------------------------
// A warp owns 32 consecutive (cell, point) items of the flat space (their J, K and detJ are one contiguous run of each
// output array).  The node coordinates of the cells those items belong to -- one contiguous run of `coords` -- arrive
// with one coalesced pass into a warp-private odd-pitch tile; every lane forms its J from that tile and the derivative
// table in shared memory (same summation order as geometry_kernel: nodes outermost), inverts it, and the three result
// tiles leave as one TMA bulk copy each (or lane-consecutive stores when the run is ragged or unaligned).  K = adj / div
// uses one reciprocal and FMA-corrected quotients (div_by below: bit-equal to IEEE division).  No block
// barrier after the table is in shared memory; the tile position advances incrementally.
// xcap: entries of one coordinate buffer (>= cells per tile * odd pitch); a warp has two, filled by cp.async one tile ahead.
template <typename T, int G, int TD>
__global__ void __launch_bounds__(kThreads)
    geometry_warp_kernel(const T* __restrict__ coords, const T* __restrict__ dphi, size_t nc, int nv, uint32_t npts, int xcap,
                         T* __restrict__ Jout, T* __restrict__ detout, T* __restrict__ Kout, int bulk)
{
  constexpr int MA = G * TD;
  extern __shared__ __align__(128) unsigned char gw_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int ntab = TD * int(npts) * nv;
  T* s_d = reinterpret_cast<T*>(gw_smem); // [TD][nv][npts]
  constexpr int TP = (32 * MA + 3) & ~3;
  const int per_warp = 2 * xcap + 2 * TP + 32;
  T* cb = s_d + ((ntab + 3) & ~3) + size_t(warp) * per_warp; // coordinates [2][cells][XP]: this tile's and the next one's
  T* sx = cb + 2 * xcap;                                      // J [32][MA] dense
  T* sK = sx + TP;                                            // K [32][MA] dense
  T* sd = sK + TP;                                            // detJ [32]
  // the table is kept node-major, [TD][nv][npts]: lanes = consecutive points read consecutive entries (no bank conflicts)
  for (int i = threadIdx.x; i < ntab; i += kThreads)
  {
    const int v = i % nv, ap = i / nv, pp = ap % int(npts), a = ap / int(npts);
    s_d[(a * nv + v) * int(npts) + pp] = dphi[i];
  }
  __syncthreads();
  const int NX = nv * G, XP = NX | 1;
  const size_t total = nc * size_t(npts);
  const size_t stride = size_t(gridDim.x) * (kThreads / 32) * 32;
  size_t t0 = (size_t(blockIdx.x) * (kThreads / 32) + warp) * 32;
  if (t0 >= total)
    return;
  size_t c0 = t0 / npts;
  uint32_t p0 = uint32_t(t0 - c0 * npts);
  const size_t dc = stride / npts;
  const uint32_t dp = uint32_t(stride - dc * npts);
  // node coordinates of a tile's cells: global -> shared by cp.async, one tile ahead of the arithmetic
  auto fetch = [&](size_t ft0, size_t fc0, uint32_t fp0, T* buf)
  {
    const uint32_t fr = uint32_t(min(size_t(32), total - ft0));
    const uint32_t fcells = (fp0 + fr - 1) / npts + 1;
    for (uint32_t q = lane; q < fcells * uint32_t(NX); q += 32)
      asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(uint32_t(__cvta_generic_to_shared(buf + (q / NX) * XP + q % NX))),
                   "l"(coords + fc0 * NX + q), "n"(sizeof(T))
                   : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  fetch(t0, c0, p0, cb);
  bool pending = false;
  int cur = 0;
  for (; t0 < total; t0 += stride)
  {
    const uint32_t nrow = uint32_t(min(size_t(32), total - t0));
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncwarp();
    // position of the next tile; its coordinates travel while this tile is computed
    size_t cn = c0 + dc;
    uint32_t pn = p0 + dp;
    if (pn >= npts)
    {
      pn -= npts;
      ++cn;
    }
    if (t0 + stride < total)
      fetch(t0 + stride, cn, pn, cb + (cur ^ 1) * xcap);
    const T* sxc = cb + cur * xcap;
    const uint32_t il = p0 + min(uint32_t(lane), nrow - 1); // lanes past the last item repeat it and do not store
    const uint32_t off = il / npts, pt = il - off * npts;
    const T* xc = sxc + off * XP;
    T J[G][TD];
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int a = 0; a < TD; ++a)
        J[g][a] = T(0);
    for (int v = 0; v < nv; ++v)
    {
      T dv[TD];
#pragma unroll
      for (int a = 0; a < TD; ++a)
        dv[a] = s_d[(a * nv + v) * int(npts) + int(pt)];
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
    if (pending)
    {
      // the result tiles of the previous item tile have been read by the copy engine
      if (lane < 3)
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
      pending = false;
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
      for (int a = 0; a < TD; ++a)
        sx[lane * MA + g * TD + a] = J[g][a];
    {
      // K = adj / div, correctly rounded: one reciprocal and FMA-corrected quotients (div_by: equal to the IEEE
      // division bit for bit, tools/div_check.cu), instead of G * TD divisions
      T q[MA];
#pragma unroll
      for (int a = 0; a < TD; ++a)
#pragma unroll
        for (int g = 0; g < G; ++g)
          q[a * G + g] = adj[a][g];
      div_by<T, MA>(q, div, safe_reciprocal(div));
#pragma unroll
      for (int k = 0; k < MA; ++k)
        sK[lane * MA + k] = q[k];
    }
    sd[lane] = det_from_div<T, G, TD>(div);
    if (bulk and nrow == 32)
    {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      T* dst = lane == 0 ? Jout : lane == 1 ? Kout : detout;
      const T* src = lane == 0 ? sx : lane == 1 ? sK : sd;
      const uint32_t per = lane < 2 ? MA : 1;
      if (lane < 3 and dst)
      {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + t0 * per),
                     "r"(uint32_t(__cvta_generic_to_shared(src))), "r"(uint32_t(32 * per * sizeof(T)))
                     : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      pending = true;
    }
    else
    {
      __syncwarp();
      for (uint32_t q = lane; q < nrow * MA; q += 32)
      {
        if (Jout)
          Jout[t0 * MA + q] = sx[q];
        if (Kout)
          Kout[t0 * MA + q] = sK[q];
      }
      if (detout and uint32_t(lane) < nrow)
        detout[t0 + lane] = sd[lane];
      __syncwarp();
    }
    c0 = cn;
    p0 = pn;
    cur ^= 1;
  }
  if (pending and lane < 3)
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

__host__ __device__ constexpr int pad4(int n) { return (n + 3) & ~3; }
// per-cell record of the fused kernel in shared memory: the matrix each row is multiplied by, detJ and 1 / detJ;
// an odd number of entries, so lane = cell writes are free of bank conflicts
template <int G, int TD>
__host__ __device__ constexpr int cell_record() { return (G * TD + 2) | 1; }
// entries a lane prefetches for its cell: vertex coordinates, or the given matrix (and detJ)
template <int MAP, int G, int TD, bool GEOM>
__host__ __device__ constexpr int prefetch_entries()
{
  return GEOM ? (TD + 1) * G : (MAP == BX_MAP_IDENTITY ? 0 : G * TD + (MAP == BX_MAP_CONTRAVARIANT_PIOLA ? 1 : 0));
}

// MAP: BX_MAP_IDENTITY (value size VS, no matrix), BX_MAP_COVARIANT_PIOLA, BX_MAP_CONTRAVARIANT_PIOLA.
// GEOM: J / detJ / K of the affine simplex cell from coords[nc][TD + 1][G]; otherwise A / detJ are read per cell:
// A = K[nc][TD][G] (covariant) or J[nc][G][TD] (contravariant).
// tab[nstates][rows][IN_VS]: the reference values of every row, transformed for state 0..7 of the row's sub-entity.
//
// A warp takes 32 consecutive cells per iteration.
//   phase A, lane = cell: the cell's vertex coordinates (loaded into registers one iteration ahead), J, its adjugate and
//     determinant, K = adj / det or 1 / detJ -- all 32 lanes busy, the IEEE divisions of 32 cells for the price of
//     one -- written as the cell's record to shared memory; cell_info stays in the lane's register.
//   phase B, `batch` cells at a time: per cell the record is read back with warp-uniform loads into registers, the
//     lanes walk the cell's rows (descriptor of a row = first cell_info bit and mask of its sub-entity, one word of a
//     shared table), pick the reference values of the row's state, map them and write the slice; the rows of the batch
//     are one contiguous run of `out` and leave as ONE TMA bulk copy, the warp going on with the next batch in the
//     other slice while the copy engine reads.
template <typename T, int MAP, int G, int TD, bool GEOM, int VS, bool FLAT>
__global__ void __launch_bounds__(kFusedMaxThreads, 1)
    physical_basis_kernel(const T* __restrict__ tab, const T* __restrict__ coords, const T* __restrict__ A,
                          const T* __restrict__ detJ, const FusedParams p, T* __restrict__ out)
{
  constexpr int IN_VS = MAP == BX_MAP_IDENTITY ? VS : TD;
  constexpr int OUT_VS = MAP == BX_MAP_IDENTITY ? VS : G;
  constexpr int MA = G * TD;
  constexpr int NX = (TD + 1) * G; // coordinates of an affine simplex cell
  constexpr int REC = cell_record<G, TD>();
  constexpr int NPRE_ALL = prefetch_entries<MAP, G, TD, GEOM>();
  extern __shared__ __align__(128) unsigned char fused_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int rows = p.rows, E = rows * OUT_VS, B = p.batch;
  const int slice_stride = pad4(B * E);
  T* s_tab = reinterpret_cast<T*>(fused_smem);                                            // [nstates][tab_stride]
  uint32_t* s_desc = reinterpret_cast<uint32_t*>(s_tab + pad4(p.nstates * p.tab_stride)); // [rows]
  T* wbase = reinterpret_cast<T*>(s_desc + pad4(rows)) + size_t(warp) * (p.nbuf * slice_stride + pad4(32 * REC) + pad4(32 * NPRE_ALL));
  T* slices = wbase;                       // [nbuf][batch][E]
  T* rec = wbase + p.nbuf * slice_stride;  // [32][REC]
  for (int i = threadIdx.x; i < p.nstates * p.tab_stride; i += blockDim.x)
    s_tab[i] = tab[i];
  for (int r = threadIdx.x; r < rows; r += blockDim.x)
  {
    uint32_t shift = 0, mask = 0;
    if (p.cell_info)
    {
      const int slot = p.dof_slot[r % p.ndofs];
      if (slot >= 0)
      {
        const int4 si = p.slot_info[slot];
        shift = uint32_t(si.x);
        mask = uint32_t(si.y);
      }
    }
    s_desc[r] = shift | mask << 5;
  }
  __syncthreads();

  const size_t warps = size_t(gridDim.x) * (blockDim.x / 32);
  const size_t first = (size_t(blockIdx.x) * (blockDim.x / 32) + warp) * 32;
  // ---- prefetch: the cell of this lane for the NEXT iteration travels global -> shared by cp.async (no registers
  // held across phase B); every lane reads back only what it copied itself ----
  constexpr int NPRE = NPRE_ALL;
  [[maybe_unused]] T* pf = rec + pad4(32 * REC) + lane * NPRE; // [32][NPRE]
  auto fetch = [&](size_t c0)
  {
    if constexpr (NPRE > 0)
    {
      const size_t c = min(c0 + lane, p.nc - 1);
      const uint32_t dst = uint32_t(__cvta_generic_to_shared(pf));
      if constexpr (GEOM)
      {
        const T* src = coords + c * NX;
        if ((NX * sizeof(T)) % 16 == 0 and p.vec_coords)
        {
#pragma unroll
          for (int k = 0; k < int(NX * sizeof(T) / 16); ++k)
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * k), "l"(reinterpret_cast<const char*>(src) + 16 * k) : "memory");
        }
        else
        {
#pragma unroll
          for (int k = 0; k < NX; ++k)
            asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst + uint32_t(sizeof(T)) * k), "l"(src + k), "n"(sizeof(T)) : "memory");
        }
      }
      else
      {
        const T* src = A + c * MA;
#pragma unroll
        for (int k = 0; k < MA; ++k)
          asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst + uint32_t(sizeof(T)) * k), "l"(src + k), "n"(sizeof(T)) : "memory");
        if constexpr (MAP == BX_MAP_CONTRAVARIANT_PIOLA)
          asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(dst + uint32_t(sizeof(T)) * MA), "l"(detJ + c), "n"(sizeof(T)) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    }
  };
  if (first < p.nc)
    fetch(first);
  int in_flight = 0, cur = 0; // committed bulk groups not yet waited for; slice in use
  for (size_t c0 = first; c0 < p.nc; c0 += warps * 32)
  {
    const int nb = int(min(size_t(32), p.nc - c0));
    // items of a per-point geometry share the cell_info of their cell and take the table rows of their point
    const uint32_t my_ci = (p.cell_info and lane < nb) ? p.cell_info[p.ppc > 1 ? (c0 + lane) / size_t(p.ppc) : c0 + lane] : 0u;
    [[maybe_unused]] const uint32_t pt0 = p.ppc > 1 ? uint32_t(c0 % size_t(p.ppc)) : 0u;
    // ---- phase A: lane = cell ----
    if constexpr (MAP != BX_MAP_IDENTITY)
    {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
      T pre[NPRE];
#pragma unroll
      for (int k = 0; k < NPRE; ++k)
        pre[k] = pf[k];
      if (c0 + warps * 32 < p.nc)
        fetch(c0 + warps * 32); // the next iteration's cell: in flight during the rest of this one
      T* mine = rec + lane * REC;
      if constexpr (GEOM)
      {
        T J[G][TD];
#pragma unroll
        for (int g = 0; g < G; ++g)
#pragma unroll
          for (int a = 0; a < TD; ++a)
            J[g][a] = pre[(1 + a) * G + g] - pre[g];
        T adj[TD][G], div;
        adjugate<T, G, TD>(J, adj, div);
        if constexpr (MAP == BX_MAP_COVARIANT_PIOLA)
        {
#pragma unroll
          for (int a = 0; a < TD; ++a)
#pragma unroll
            for (int g = 0; g < G; ++g)
              mine[a * G + g] = adj[a][g] / div;
        }
        else
        {
#pragma unroll
          for (int g = 0; g < G; ++g)
#pragma unroll
            for (int a = 0; a < TD; ++a)
              mine[g * TD + a] = J[g][a];
          const T det = det_from_div<T, G, TD>(div);
          mine[MA] = det;
          mine[MA + 1] = safe_reciprocal(det);
        }
      }
      else
      {
#pragma unroll
        for (int k = 0; k < MA; ++k)
          mine[k] = pre[k];
        if constexpr (MAP == BX_MAP_CONTRAVARIANT_PIOLA)
        {
          mine[MA] = pre[MA];
          mine[MA + 1] = safe_reciprocal(pre[MA]);
        }
      }
    }
    __syncwarp();
    if constexpr (FLAT)
    {
      // ---- phase B, flat: the (cell, row) pairs of the warp's cells are ONE index space walked 32 (or 64) at a time, so
      // every lane has a row whatever rows % 32 is (70 rows: 9 passes per 4 cells instead of 12); the cell's record is
      // read per lane (at most two different cells per pass when rows >= 32: broadcast loads), the mapped row goes
      // straight to global memory: a pass writes one contiguous run of 32 * OUT_VS entries of `out` ----
      static_assert(MAP != BX_MAP_IDENTITY, "the flat row walk is for the Piola maps");
      const uint32_t total = uint32_t(nb) * uint32_t(rows);
      T* const obase = out + c0 * p.out_pitch;
      auto fpass = [&](auto nh, uint32_t f0)
      {
        constexpr int NH = decltype(nh)::value;
        bool ok[NH];
        const T* u[NH];
        const T* rc[NH];
        T* g[NH];
#pragma unroll
        for (int h = 0; h < NH; ++h)
        {
          const uint32_t f = f0 + 32u * h + lane;
          ok[h] = f < total;
          const uint32_t fc = min(f, total - 1u); // lanes past the last row repeat it and do not store
          const uint32_t cell = p.rows_magic ? __umulhi(fc, p.rows_magic) : fc;
          const uint32_t row = fc - cell * uint32_t(rows);
          const uint32_t ci = __shfl_sync(0xffffffffu, my_ci, int(cell));
          const uint32_t d = s_desc[row];
          const int st = int(ci >> (d & 31u)) & int(d >> 5);
          const uint32_t trow = p.ppc > 1 ? ((pt0 + cell) % uint32_t(p.ppc)) * uint32_t(rows) + row : row;
          u[h] = s_tab + st * p.tab_stride + trow * IN_VS;
          rc[h] = rec + cell * REC;
          g[h] = obase + cell * p.out_pitch + row * OUT_VS;
        }
        T uu[NH][IN_VS], m[NH][MA], acc[NH][OUT_VS];
#pragma unroll
        for (int h = 0; h < NH; ++h)
        {
#pragma unroll
          for (int k = 0; k < IN_VS; ++k)
            uu[h][k] = u[h][k];
#pragma unroll
          for (int k = 0; k < MA; ++k)
            m[h][k] = rc[h][k];
        }
#pragma unroll
        for (int i = 0; i < OUT_VS; ++i)
#pragma unroll
          for (int h = 0; h < NH; ++h)
          {
            T a = 0;
            if constexpr (MAP == BX_MAP_COVARIANT_PIOLA)
            {
#pragma unroll
              for (int k = 0; k < TD; ++k)
                a += m[h][k * G + i] * uu[h][k]; // out[i] = sum_k K[k][i] in[k]   (maps.h:73-94)
            }
            else
            {
#pragma unroll
              for (int k = 0; k < TD; ++k)
                a += m[h][i * TD + k] * uu[h][k]; // out[i] = (sum_k J[i][k] in[k]) / det   (maps.h:100-124)
            }
            acc[h][i] = a;
          }
        if constexpr (MAP == BX_MAP_CONTRAVARIANT_PIOLA)
        {
#pragma unroll
          for (int h = 0; h < NH; ++h)
            div_by<T, OUT_VS>(acc[h], rc[h][MA], rc[h][MA + 1]);
        }
#pragma unroll
        for (int h = 0; h < NH; ++h)
          if (ok[h])
          {
#pragma unroll
            for (int i = 0; i < OUT_VS; ++i)
              g[h][i] = acc[h][i];
          }
      };
#pragma unroll 1
      for (uint32_t f0 = 0; f0 < total; f0 += 32 * kFlatPair)
      {
        if (kFlatPair == 2 and f0 + 32 < total)
          fpass(std::integral_constant<int, kFlatPair>{}, f0);
        else
          fpass(std::integral_constant<int, 1>{}, f0);
      }
    }
    // ---- phase B ----
    if constexpr (!FLAT)
    for (int b0 = 0; b0 < nb; b0 += B)
    {
      const int cells = min(B, nb - b0);
      T* slice = slices + cur * slice_stride;
      if (in_flight >= p.nbuf)
      {
        // the copy that last used this slice has been read out of shared memory
        if (lane == 0)
        {
          if (p.nbuf == 2)
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          else
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        in_flight = p.nbuf - 1;
        __syncwarp();
      }
      for (int bi = 0; bi < cells; ++bi)
      {
        const int cell = b0 + bi;
        const uint32_t ci = __shfl_sync(0xffffffffu, my_ci, cell);
        [[maybe_unused]] T m[MA];
        [[maybe_unused]] T det = T(1), inv = T(1);
        if constexpr (MAP != BX_MAP_IDENTITY)
        {
          const T* rc = rec + cell * REC;
#pragma unroll
          for (int k = 0; k < MA; ++k)
            m[k] = rc[k];
          if constexpr (MAP == BX_MAP_CONTRAVARIANT_PIOLA)
          {
            det = rc[MA];
            inv = rc[MA + 1];
          }
        }
        T* o = slice + bi * E;
        // two row passes per trip when both exist (independent chains for the FP64 pipe), else one
        auto pass = [&](auto nh, int r0)
        {
          constexpr int NH = decltype(nh)::value;
          int r[NH];
          const T* u[NH];
#pragma unroll
          for (int h = 0; h < NH; ++h)
          {
            r[h] = r0 + 32 * h + lane;
            const int rc = min(r[h], rows - 1); // lanes past the last row repeat it and do not store
            const uint32_t d = s_desc[rc];
            const int st = int(ci >> (d & 31u)) & int(d >> 5);
            u[h] = s_tab + st * p.tab_stride + rc * IN_VS;
          }
          if constexpr (MAP == BX_MAP_IDENTITY)
          {
#pragma unroll
            for (int h = 0; h < NH; ++h)
              if (r[h] < rows)
              {
#pragma unroll
                for (int i = 0; i < VS; ++i)
                  o[r[h] * VS + i] = u[h][i];
              }
          }
          else
          {
            T uu[NH][IN_VS], acc[NH][OUT_VS];
#pragma unroll
            for (int h = 0; h < NH; ++h)
#pragma unroll
              for (int k = 0; k < IN_VS; ++k)
                uu[h][k] = u[h][k];
#pragma unroll
            for (int i = 0; i < OUT_VS; ++i)
#pragma unroll
              for (int h = 0; h < NH; ++h)
              {
                T a = 0;
                if constexpr (MAP == BX_MAP_COVARIANT_PIOLA)
                {
                  // out[i] = sum_k K[k][i] in[k]   (maps.h:73-94)
#pragma unroll
                  for (int k = 0; k < TD; ++k)
                    a += m[k * G + i] * uu[h][k];
                }
                else
                {
                  // out[i] = (sum_k J[i][k] in[k]) / det   (maps.h:100-124)
#pragma unroll
                  for (int k = 0; k < TD; ++k)
                    a += m[i * TD + k] * uu[h][k];
                }
                acc[h][i] = a;
              }
            if constexpr (MAP == BX_MAP_CONTRAVARIANT_PIOLA)
            {
#pragma unroll
              for (int h = 0; h < NH; ++h)
                div_by<T, OUT_VS>(acc[h], det, inv);
            }
#pragma unroll
            for (int h = 0; h < NH; ++h)
              if (r[h] < rows)
              {
#pragma unroll
                for (int i = 0; i < OUT_VS; ++i)
                  o[r[h] * OUT_VS + i] = acc[h][i];
              }
          }
        };
#pragma unroll 1
        for (int r0 = 0; r0 < rows; r0 += 32 * kPair)
        {
          if (kPair == 2 and r0 + 32 < rows)
            pass(std::integral_constant<int, kPair>{}, r0);
          else
            pass(std::integral_constant<int, 1>{}, r0);
        }
      }
      if (p.bulk and cells == B)
      {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0)
        {
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(out + (c0 + b0) * size_t(E)),
                       "r"(uint32_t(__cvta_generic_to_shared(slice))), "r"(uint32_t(B * E * sizeof(T)))
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++in_flight;
        cur = cur + 1 == p.nbuf ? 0 : cur + 1;
      }
      else
      {
        __syncwarp();
        for (int bi = 0; bi < cells; ++bi)
        {
          T* g = out + (c0 + b0 + bi) * p.out_pitch;
          for (int e = lane; e < E; e += 32)
            g[e] = slice[bi * E + e];
        }
        __syncwarp(); // the slice is overwritten by the next batch
      }
    }
    __syncwarp(); // the records are overwritten by the next iteration's phase A
  }
  if (in_flight > 0 and lane == 0)
    asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
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
  if (!dphi)
  {
    // affine simplex cells: warp-per-32-cells kernel with coalesced loads and bulk copy-out
    constexpr int NX = (TD + 1) * G, MA = G * TD, XP = NX | 1;
    constexpr int kWarpT = ((32 * (XP > MA ? XP : MA) + 3) & ~3) + ((32 * MA + 3) & ~3) + 32;
    auto kern = geometry_affine_kernel<T, G, TD>;
    const size_t smem = size_t(kThreads / 32) * kWarpT * sizeof(T);
    auto aligned = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
    const int bulk = aligned(J) and aligned(K) and aligned(detJ) and (32 * MA * sizeof(T)) % 16 == 0;
    const int grid = persistent_grid(kern, ceil_div(nc, size_t(kThreads)), smem);
    kern<<<grid, kThreads, smem, s>>>(coords, nc, J, detJ, K, bulk);
    BX_LAUNCHED();
    return;
  }
  {
    // warp form: derivative table and the warps' tiles in shared memory (falls through when they do not fit)
    constexpr int MA = G * TD;
    const int cells_per_tile = std::min(32, 31 / npts + 2);
    const int xcap = (cells_per_tile * ((nv * G) | 1) + 3) & ~3; // one coordinate buffer (two per warp)
    const size_t ntab = (size_t(TD) * npts * nv + 3) & ~size_t(3);
    const size_t smem = (ntab + size_t(kThreads / 32) * (2 * xcap + 2 * ((32 * MA + 3) & ~3) + 32)) * sizeof(T);
    static const bool warp_form = []
    {
      const char* v = std::getenv("BX_GEOM_WARP"); // development knob: 0 selects the thread-per-item kernel
      return v ? std::atoi(v) != 0 : true;
    }();
    if (warp_form and smem <= 100 * 1024 and npts < (1 << 20))
    {
      auto kern = geometry_warp_kernel<T, G, TD>;
      auto aligned = [](const void* q) { return reinterpret_cast<uintptr_t>(q) % 16 == 0; };
      const int bulk = aligned(J) and aligned(K) and aligned(detJ) and (32 * MA * sizeof(T)) % 16 == 0;
      const int grid = persistent_grid(kern, ceil_div(nc * size_t(npts), size_t(kThreads)), smem);
      kern<<<grid, kThreads, smem, s>>>(coords, dphi, nc, nv, uint32_t(npts), xcap, J, detJ, K, bulk);
      BX_LAUNCHED();
      return;
    }
  }
  auto kern = geometry_kernel<T, G, TD>;
  const size_t tbl = size_t(TD) * npts * nv * sizeof(T);
  const int in_smem = tbl <= 96 * 1024;
  const size_t smem = in_smem ? tbl : 0;
  const int grid = persistent_grid(kern, ceil_div(nc * size_t(npts), size_t(kThreads)), smem);
  kern<<<grid, kThreads, smem, s>>>(coords, dphi, nc, nv, npts, in_smem, J, detJ, K);
  BX_LAUNCHED();
}

// Tuning aid (development only): BX_FUSED_TUNE="b=2,nbuf=2,nw=8" overrides the cells per slice, the slices per warp and the warps per CTA.
struct FusedTune
{
  int b = 0, nbuf = 0, nw = 0, flat = -1;
};
inline const FusedTune& fused_tune()
{
  static FusedTune t = []()
  {
    FusedTune v;
    if (const char* e = std::getenv("BX_FUSED_TUNE"))
    {
      std::string str(e);
      auto get = [&](const char* key, int& out)
      {
        const size_t k = str.find(key);
        if (k != std::string::npos)
          out = std::atoi(str.c_str() + k + std::strlen(key));
      };
      get("b=", v.b);
      get("nbuf=", v.nbuf);
      get("nw=", v.nw);
      get("flat=", v.flat);
    }
    return v;
  }();
  return t;
}

template <typename T, int MAP, int G, int TD, bool GEOM, int VS>
void launch_fused(const T* tab, const T* coords, const T* A, const T* detJ, FusedParams p, T* out, cudaStream_t s)
{
  constexpr int OUT_VS = MAP == BX_MAP_IDENTITY ? VS : G;
  constexpr int REC = cell_record<G, TD>();
  const int E = p.rows * OUT_VS;
  auto smem_of = [&](int batch, int nbuf, int nw)
  {
    return size_t(pad4(p.nstates * p.tab_stride)) * sizeof(T) + size_t(pad4(p.rows)) * 4
           + size_t(nw) * (size_t(nbuf) * pad4(batch * E) + pad4(32 * REC) + pad4(32 * prefetch_entries<MAP, G, TD, GEOM>())) * sizeof(T);
  };
  const bool contiguous = p.out_pitch == size_t(E) and reinterpret_cast<uintptr_t>(out) % 16 == 0;
  static const size_t smem_max = [&]
  {
    int dev = 0, v = 0;
    BX_CUDA(cudaGetDevice(&dev));
    BX_CUDA(cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    return size_t(v);
  }();
  p.rows_magic = p.rows > 1 ? uint32_t((uint64_t(1) << 32) / uint64_t(p.rows)) + 1u : 0u;
  if constexpr (MAP != BX_MAP_IDENTITY)
  {
    // Piola maps: flat walk over the (cell, row) pairs, rows written straight to global memory (no slices)
    if (fused_tune().flat != 0 or p.ppc > 1)
    {
      auto flat = physical_basis_kernel<T, MAP, G, TD, GEOM, VS, true>;
      const int nw = (fused_tune().nw == 4 or fused_tune().nw == 8 or fused_tune().nw == 16) ? fused_tune().nw : 8;
      p.batch = 0;
      p.nbuf = 0;
      p.bulk = 0;
      p.vec_coords = reinterpret_cast<uintptr_t>(coords) % 16 == 0;
      const size_t smem = smem_of(0, 0, nw);
      if (smem <= smem_max)
      {
        if (smem > 48 * 1024)
          BX_CUDA(cudaFuncSetAttribute(flat, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        int per_sm = 1;
        BX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, flat, nw * 32, smem));
        if (per_sm >= 1)
        {
          const size_t want = ceil_div(p.nc, size_t(nw) * 32);
          const int grid = int(std::max<size_t>(1, std::min(want, size_t(sm_count()) * per_sm)));
          flat<<<grid, nw * 32, smem, s>>>(tab, coords, A, detJ, p, out);
          BX_LAUNCHED();
          return;
        }
      }
      if (p.ppc > 1)
        fail(BX_ERR_UNSUPPORTED, "physical_basis: the state tables of this many points do not fit in shared memory (split "
                                 "the point set)");
    }
  }
  auto kern = physical_basis_kernel<T, MAP, G, TD, GEOM, VS, false>;
  cudaFuncAttributes fa;
  BX_CUDA(cudaFuncGetAttributes(&fa, kern));
  // Warps per CTA (4 / 8 / 16), cells per slice and slices per warp: the configuration with the most resident warps per
  // SM (the state tables are per CTA: 34 KB for 180 rows), then the largest batch, then double buffering.  A batch
  // whose rows are a whole number of 16-byte units leaves by TMA bulk copy.
  int best_nw = 0, best_warps = -1, best_rank = -1;
  p.batch = 1;
  p.nbuf = 1;
  static const int menu[][2] = {{1, 1}, {1, 2}, {2, 1}, {4, 1}, {2, 2}, {4, 2}};
  for (int nw : {4, 8, 16})
    for (int mi = 0; mi < 6; ++mi)
    {
      const int b = menu[mi][0], nbuf = menu[mi][1];
      if (mi > 0 and (size_t(b) * E * sizeof(T)) % 16 != 0)
        continue;
      const size_t sm = smem_of(b, nbuf, nw);
      if (sm > smem_max)
        continue;
      const int by_smem = int((size_t(228) * 1024 - 1024) / (sm + 1024));
      const int by_regs = int(65536 / (std::max(fa.numRegs, 1) * nw * 32));
      const int resident = nw * std::min({by_smem, by_regs, 64 / nw});
      // prefer more resident warps up to 24 per SM (beyond that the kernel does not gain), then the richer slicing
      const int capped = std::min(resident, 24);
      if (capped > best_warps or (capped == best_warps and mi > best_rank))
      {
        best_warps = capped;
        best_rank = mi;
        best_nw = nw;
        p.batch = b;
        p.nbuf = nbuf;
      }
    }
  if (best_nw == 0)
    fail(BX_ERR_UNSUPPORTED, "physical_basis: the point chunk does not fit in shared memory");
  if (fused_tune().b > 0 and fused_tune().b <= 32)
    p.batch = fused_tune().b;
  if (fused_tune().nbuf == 1 or fused_tune().nbuf == 2)
    p.nbuf = fused_tune().nbuf;
  if (fused_tune().nw == 4 or fused_tune().nw == 8 or fused_tune().nw == 16)
    best_nw = fused_tune().nw;
  p.bulk = contiguous and (size_t(p.batch) * E * sizeof(T)) % 16 == 0;
  p.vec_coords = reinterpret_cast<uintptr_t>(coords) % 16 == 0;
  const size_t smem = smem_of(p.batch, p.nbuf, best_nw);
  if (smem > smem_max)
    fail(BX_ERR_UNSUPPORTED, "physical_basis: the point chunk does not fit in shared memory");
  if (smem > 48 * 1024)
    BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 1;
  BX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, best_nw * 32, smem));
  if (per_sm < 1)
    fail(BX_ERR_UNSUPPORTED, "physical_basis: the kernel does not fit on an SM");
  const size_t want = ceil_div(p.nc, size_t(best_nw) * 32);
  const int grid = int(std::max<size_t>(1, std::min(want, size_t(sm_count()) * per_sm)));
  kern<<<grid, best_nw * 32, smem, s>>>(tab, coords, A, detJ, p, out);
  BX_LAUNCHED();
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

// push_forward_broadcast through the flat walk of the fused kernel (one slab of reference values U[rows][tdim] for all
// cells, A = K (covariant) or J (contravariant) per cell, no transformation): the same arithmetic as maps.cu, every lane
// busy whatever rows % 32 is.  Returns false when the shape is outside the fused kernel's menu (the caller falls back to
// map_bcast_kernel).
template <typename T>
bool push_forward_broadcast_flat(int map_type, const T* U, const T* J, const T* detJ, const T* K, size_t nc, size_t rows,
                                 int gdim, int tdim, T* out, cudaStream_t s)
{
  if (map_type != BX_MAP_COVARIANT_PIOLA and map_type != BX_MAP_CONTRAVARIANT_PIOLA)
    return false;
  if (rows > 256 or geo::fused_tune().flat == 0)
    return false;
  if (!((gdim == 2 and tdim == 2) or (gdim == 3 and tdim == 3) or (gdim == 3 and tdim == 2)))
    return false;
  geo::FusedParams p;
  p.nc = nc;
  p.rows = int(rows);
  p.ndofs = int(rows);
  p.tab_stride = int(rows) * tdim;
  p.nstates = 1;
  p.ppc = 1;
  p.out_pitch = rows * size_t(gdim);
  p.dof_slot = nullptr;
  p.slot_info = nullptr;
  p.cell_info = nullptr;
  if (map_type == BX_MAP_COVARIANT_PIOLA)
    geo::launch_fused_shape<T, BX_MAP_COVARIANT_PIOLA, false>(gdim, tdim, tdim, U, nullptr, K, detJ, p, out, s);
  else
    geo::launch_fused_shape<T, BX_MAP_CONTRAVARIANT_PIOLA, false>(gdim, tdim, tdim, U, nullptr, J, detJ, p, out, s);
  return true;
}
template bool push_forward_broadcast_flat<double>(int, const double*, const double*, const double*, const double*, size_t,
                                                  size_t, int, int, double*, cudaStream_t);
template bool push_forward_broadcast_flat<float>(int, const float*, const float*, const float*, const float*, size_t, size_t,
                                                 int, int, float*, cudaStream_t);

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
                           cudaStream_t s, int per_point)
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
  // per-point geometry (non-affine cells): J[nc][npts][gdim][tdim], detJ[nc][npts], K[nc][npts][tdim][gdim].  The kernel's
  // "cells" are the (cell, point) items: rows = ndofs, one state table for all points, every item with its own matrix.
  const bool items = per_point != 0 and map != BX_MAP_IDENTITY and npts > 1;
  if (per_point and geom)
    fail(BX_ERR_INVALID_ARG, "physical_basis: per-point geometry is given as J / detJ / K, not as vertex coordinates");
  if (items and npts * size_t(e.ndofs) >= 8192)
    fail(BX_ERR_UNSUPPORTED, "physical_basis: too many points for one pass (split the point set)");
  const int chunk = items ? int(npts) : physical_basis_chunk(e, npts);
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
      p.rows = items ? e.ndofs : np * e.ndofs;
      p.tab_stride = int(period);
      p.ppc = items ? np : 1;
      if (items)
      {
        p.nc = nc * npts;
        p.out_pitch = size_t(e.ndofs) * out_vs;
      }
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
                                            cudaStream_t, int);
template void physical_basis_device<float>(const bx_element&, int, const float*, size_t, const float*, const float*,
                                           const float*, const float*, const uint32_t*, size_t, int, float*, cudaStream_t,
                                           int);
} // namespace bx

// Pipelined DOF-transformation kernels.  data[nc][cell_size] is described to the TMA unit as a 2-D tensor
// (cuTensorMapEncodeTiled, built per call in dof_transform.cu); ONE cp.async.bulk.tensor instruction moves the
// movable column range of a whole tile of cells between HBM and shared memory, issued by a dedicated DMA warp
// and tracked with mbarriers.  The seven compute warps never execute a global load, entries that can never move
// are never read, and the HBM stream of tiles i+1 / i-1 overlaps the arithmetic of tile i (three buffers).
// Used whenever the cell pitch is a multiple of 16 bytes and the range fits a TMA box (dof_transform.cu
// decides); the register-staged kernel in dof_transform.cu is the general fallback.
//
//   dof_matrix_tma_kernel   matrix elements (transform_data, finite-element.h:1762-1837): one thread per
//                           (cell, sub-entity, block) item, lanes <-> consecutive cells; the entity's values sit
//                           in registers, the dense net operator of the item's cell_info state is read from
//                           shared memory, two entries per 16-byte load, and only inside the non-zero column
//                           band of each row; results go back in place and leave as one tensor store per tile.
//   dof_perm_tma_kernel     permutation elements (permute_data, finite-element.h:1703-1760): a warp walks the
//                           cells of the tile, lanes <-> entries of the movable range, each lane keeping the
//                           packed descriptor of its entries in registers; the gather table gives the source
//                           entry, the value is read from the staged tile and stored straight to global memory
//                           (coalesced), bit for bit.  (General block sizes / right layout: one thread per
//                           entry with the descriptor looked up in shared memory.)
#pragma once

#include "element.cuh"

#include <cuda.h>

namespace bx
{
namespace tma
{
constexpr int kComputeThreads = 256;            // warps 0..7 compute
constexpr int kThreads = kComputeThreads + 32;  // warp 8 is the DMA warp
constexpr int kMaxBuffers = 4;                  // tile buffers per CTA (Params::nbuf <= kMaxBuffers)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a)
{
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(a) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t a, uint32_t bytes)
{
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n}" ::"r"(a), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t a, uint32_t parity)
{
  uint32_t ok;
  do
  {
    asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                 : "=r"(ok)
                 : "r"(a), "r"(parity)
                 : "memory");
  } while (!ok);
}
// 2-D tensor tile, global -> shared; completion counted in bytes on an mbarrier.  (x, y) = (entry, cell)
__device__ __forceinline__ void tensor_load(uint32_t dst, const CUtensorMap* map, int x, int y, uint32_t bar)
{
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
               "l"(map), "r"(x), "r"(y), "r"(bar)
               : "memory");
}
// 2-D tensor tile, shared -> global (rows past the end of the tensor are clipped); tracked in bulk groups
__device__ __forceinline__ void tensor_store(const CUtensorMap* map, int x, int y, uint32_t src)
{
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(x), "r"(y),
               "r"(src)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read()
{
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

struct FastDiv
{
  unsigned m = 0, d = 1;
  FastDiv() = default;
  // multiply-high form, exact whenever q * d < 2^32 (`qmax` = largest dividend); otherwise a real division
  __host__ FastDiv(unsigned d_, unsigned long long qmax)
      : m(d_ > 1 and qmax * d_ < (1ull << 32) ? unsigned((1ull << 32) / d_ + 1) : 0), d(d_)
  {
  }
  __device__ __forceinline__ unsigned div(unsigned q) const { return d > 1 ? (m ? __umulhi(q, m) : q / d) : q; }
};

struct Params
{
  size_t nc;
  int cell_size, n, ndofs;
  int cpt;              // cells per tile
  int r0, len;          // movable range [r0, r0 + len) of each cell, in entries; r0 is 16-byte aligned
  int pitch;            // TMA box width = shared-memory distance between cells, in entries (>= len)
  int nbuf;             // tile buffers in flight per CTA (2..kMaxBuffers)
  int skip_apply;       // tuning aid: move the data only (BX_DOF_TUNE)
  int nslots, table_size, band_size;
  FastDiv by_cpt, by_len, by_n, by_ndofs;
  const uint32_t* cell_info;
  const int4* slot_info; // {shift, mask, dim, gather base}
  const int2* slot_mat;  // {first dof, matrix base of the slot's class}
  const int* slot_band;  // band table base of the slot's class
  const int2* band;      // per (class, row): [first, last) non-zero column
  const double* mats;    // operator kind offset applied
  const int16_t* gather; // operator kind offset applied
  const uint32_t* dinfo; // per dof: shift | mask << 5 | dim << 8 | (gather base + loc) << 16
};

template <bool RIGHT>
__device__ __forceinline__ int entry(int dof, int b, int n, int ndofs)
{
  return RIGHT ? b * ndofs + dof : dof * n + b;
}

// ---- the DMA warp ---------------------------------------------------------------------------------------
// Loads tile `tl` into a buffer: cell_info by plain loads (all lanes), the data by one tensor copy (lane 0).
// The `full` barrier (count 32) completes when every lane has arrived and all bytes of the box have landed.
__device__ __forceinline__ void dma_load(const Params& p, const CUtensorMap* map, size_t tl, int lane, uint32_t* s_ci,
                                         uint32_t tile_addr, uint32_t box_bytes, uint32_t full_bar)
{
  const size_t c0 = tl * p.cpt;
  const unsigned ncell = unsigned(min(size_t(p.cpt), p.nc - c0));
  for (unsigned c = lane; c < ncell; c += 32)
    s_ci[c] = p.cell_info[c0 + c];
  if (lane == 0)
  {
    mbar_arrive_expect_tx(full_bar, box_bytes);
    tensor_load(tile_addr, map, p.r0, int(c0), full_bar);
  }
  else
    mbar_arrive(full_bar);
}

// =====================================================================================================
// Shared-memory carve-up helper: tiles start on a 128-byte boundary (TMA), one tile every `tile_stride` bytes.
__host__ __device__ inline uint32_t tile_stride_bytes(int cpt, int pitch, int elem)
{
  return (uint32_t(cpt) * uint32_t(pitch) * uint32_t(elem) + 127u) & ~127u;
}

template <typename T, bool RIGHT, int REG, bool PAIRS>
__global__ void __launch_bounds__(kThreads)
    dof_matrix_tma_kernel(const __grid_constant__ CUtensorMap tmap, const Params p)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // ---- carve: mats | band | slot tables | barriers | cell_info[kBuffers] | (128 B aligned) tiles[kBuffers] ----
  double* s_mats = reinterpret_cast<double*>(smem_raw);
  int2* s_band = reinterpret_cast<int2*>(s_mats + p.table_size);
  int4* s_info = reinterpret_cast<int4*>(s_band + p.band_size + (p.band_size & 1));
  int4* s_slot = s_info + p.nslots; // {first, matrix base, band base, 0}
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_slot + p.nslots);
  uint32_t* s_ci = reinterpret_cast<uint32_t*>(s_bar + 2 * kMaxBuffers);
  const int ci_stride = (p.cpt + 3) & ~3;
  const uint32_t head = (uint32_t(reinterpret_cast<unsigned char*>(s_ci + kMaxBuffers * ci_stride) - smem_raw) + 127u) & ~127u;
  unsigned char* tiles = smem_raw + head;
  const uint32_t tile_bytes = tile_stride_bytes(p.cpt, p.pitch, sizeof(T));
  const uint32_t box_bytes = uint32_t(p.cpt) * uint32_t(p.pitch) * sizeof(T);
  const uint32_t bar0 = smem_u32(s_bar);

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
  if (threadIdx.x == 0)
  {
    for (int b = 0; b < kMaxBuffers; ++b)
    {
      mbar_init(bar0 + 8 * b, 32);                             // full: the DMA lanes
      mbar_init(bar0 + 8 * (kMaxBuffers + b), kComputeThreads); // done: every compute thread
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const size_t ntiles = (p.nc + p.cpt - 1) / p.cpt;
  const size_t mine = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kThreads / 32 - 1)
  {
    // ---- DMA warp: loads run nbuf-1 tiles ahead; stores follow the compute warps --------------------------
    const int nbuf = p.nbuf;
    for (int i = 0; size_t(i) < mine and i < nbuf - 1; ++i)
      dma_load(p, &tmap, blockIdx.x + size_t(i) * gridDim.x, lane, s_ci + i * ci_stride,
               smem_u32(tiles) + uint32_t(i) * tile_bytes, box_bytes, bar0 + 8 * uint32_t(i));
    int b = 0, bj = nbuf - 1; // buffer of tile i, buffer of tile j = i + nbuf - 1
    uint32_t phase = 0;
    for (size_t i = 0; i < mine; ++i)
    {
      // tile j reuses the buffer of tile i - 1, whose store was committed one iteration ago: once it has read
      // shared memory, refill the buffer -- before waiting for tile i, so that the loads stay nbuf - 1 tiles
      // ahead of the compute warps
      const size_t j = i + nbuf - 1;
      if (j < mine)
      {
        if (lane == 0)
          bulk_wait_read<0>();
        __syncwarp();
        dma_load(p, &tmap, blockIdx.x + j * gridDim.x, lane, s_ci + bj * ci_stride,
                 smem_u32(tiles) + uint32_t(bj) * tile_bytes, box_bytes, bar0 + 8 * uint32_t(bj));
      }
      const size_t c0 = (blockIdx.x + i * gridDim.x) * size_t(p.cpt);
      mbar_wait(bar0 + 8 * (kMaxBuffers + b), phase); // compute warps are done with tile i
      if (lane == 0)
      {
        tensor_store(&tmap, p.r0, int(c0), smem_u32(tiles) + uint32_t(b) * tile_bytes);
        bulk_commit();
      }
      bj = bj + 1 == nbuf ? 0 : bj + 1;
      if (++b == nbuf)
      {
        b = 0;
        phase ^= 1;
      }
    }
    if (lane == 0)
      bulk_wait_all();
    return;
  }

  // ---- compute warps ---------------------------------------------------------------------------------------
  const int n = p.n, ndofs = p.ndofs, pitch = p.pitch, r0 = p.r0;
  const unsigned nitems = p.skip_apply ? 0u : unsigned(p.nslots * n) * unsigned(p.cpt);
  int b = 0;
  uint32_t phase = 0;
  for (size_t i = 0; i < mine; ++i)
  {
    const size_t c0 = (blockIdx.x + i * gridDim.x) * size_t(p.cpt);
    const unsigned ncell = unsigned(min(size_t(p.cpt), p.nc - c0));
    T* tile = reinterpret_cast<T*>(tiles + size_t(b) * tile_bytes);
    const uint32_t* ci = s_ci + b * ci_stride;
    mbar_wait(bar0 + 8 * b, phase);
    for (unsigned it = threadIdx.x; it < nitems; it += kComputeThreads)
    {
      const unsigned sb = p.by_cpt.div(it), c = it - sb * p.cpt;
      if (c >= ncell)
        continue;
      const unsigned s = p.by_n.div(sb), bl = sb - s * n;
      const int4 si = s_info[s];
      const unsigned state = (ci[c] >> si.x) & unsigned(si.y);
      if (state == 0)
        continue;
      const int dim = si.z;
      const int4 sl = s_slot[s];
      const int rs = mat_row_stride(dim);
      const double* M = s_mats + sl.y + state * mat_state_stride(dim);
      const int2* band = s_band + sl.z;
      T* cell = tile + c * pitch - r0;
      double v[REG];
      if constexpr (PAIRS)
      {
        // n == 1, left layout, every entity starts on an even entry: 16-byte shared accesses throughout
        double2* x = reinterpret_cast<double2*>(cell + sl.x);
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
            v[2 * jj] = 2 * jj < dim ? cell[sl.x + 2 * jj] : 0.0;
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
            cell[sl.x + i0] = a0 + a1;
        }
      }
      else
      {
#pragma unroll
        for (int j = 0; j < REG; ++j)
          v[j] = j < dim ? static_cast<double>(cell[entry<RIGHT>(sl.x + j, bl, n, ndofs)]) : 0.0;
        for (int i0 = 0; i0 < dim; ++i0)
        {
          const int2 bd = band[i0];
          const double2* row = reinterpret_cast<const double2*>(M + i0 * rs);
          double a0 = 0.0, a1 = 0.0;
#pragma unroll
          for (int jj = 0; jj < REG / 2; ++jj)
            if (2 * jj + 1 >= bd.x and 2 * jj < bd.y)
            {
              const double2 m = row[jj];
              a0 = fma(m.x, v[2 * jj], a0);
              a1 = fma(m.y, v[2 * jj + 1], a1);
            }
          cell[entry<RIGHT>(sl.x + i0, bl, n, ndofs)] = static_cast<T>(a0 + a1);
        }
      }
    }
    fence_async_smem(); // generic-proxy writes of this thread -> visible to the tensor store
    mbar_arrive(bar0 + 8 * (kMaxBuffers + b));
    if (++b == p.nbuf)
    {
      b = 0;
      phase ^= 1;
    }
  }
}

// =====================================================================================================
// U is the unsigned integer type of the entry size (bit-exact moves for any 4- or 8-byte data type).
// EPL > 0: warp-per-cell form for block size 1 / left layout, EPL entries per lane (len <= 32 * EPL);
// EPL == 0: one thread per entry, any block size and layout.
template <typename U, bool RIGHT, int EPL>
__global__ void __launch_bounds__(kThreads)
    dof_perm_tma_kernel(const __grid_constant__ CUtensorMap tmap, U* __restrict__ data, const Params p)
{
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // ---- carve: per-dof descriptors | gather table | barriers | cell_info[kBuffers] | tiles[kBuffers] --------
  uint32_t* s_dinfo = reinterpret_cast<uint32_t*>(smem_raw);
  int16_t* s_gather = reinterpret_cast<int16_t*>(s_dinfo + p.ndofs);
  const size_t tables = (size_t(p.ndofs) * 4 + size_t(p.table_size) * 2 + 15) & ~size_t(15);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + tables);
  uint32_t* s_ci = reinterpret_cast<uint32_t*>(s_bar + 2 * kMaxBuffers);
  const int ci_stride = (p.cpt + 3) & ~3;
  const uint32_t head = (uint32_t(reinterpret_cast<unsigned char*>(s_ci + kMaxBuffers * ci_stride) - smem_raw) + 127u) & ~127u;
  unsigned char* tiles = smem_raw + head;
  const uint32_t tile_bytes = tile_stride_bytes(p.cpt, p.pitch, sizeof(U));
  const uint32_t box_bytes = uint32_t(p.cpt) * uint32_t(p.pitch) * sizeof(U);
  const uint32_t bar0 = smem_u32(s_bar);

  for (int i = threadIdx.x; i < p.ndofs; i += kThreads)
    s_dinfo[i] = p.dinfo[i];
  for (int i = threadIdx.x; i < p.table_size; i += kThreads)
    s_gather[i] = p.gather[i];
  if (threadIdx.x == 0)
  {
    for (int b = 0; b < kMaxBuffers; ++b)
    {
      mbar_init(bar0 + 8 * b, 32);                             // full
      mbar_init(bar0 + 8 * (kMaxBuffers + b), kComputeThreads); // empty
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const size_t ntiles = (p.nc + p.cpt - 1) / p.cpt;
  const size_t mine = blockIdx.x < ntiles ? (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kThreads / 32 - 1)
  {
    int b = 0;
    uint32_t phase = 1; // parity of the previous use of the buffer
    for (size_t i = 0; i < mine; ++i)
    {
      if (i >= size_t(p.nbuf))
        mbar_wait(bar0 + 8 * (kMaxBuffers + b), phase); // tile i - nbuf has been read
      dma_load(p, &tmap, blockIdx.x + i * gridDim.x, lane, s_ci + b * ci_stride,
               smem_u32(tiles) + uint32_t(b) * tile_bytes, box_bytes, bar0 + 8 * uint32_t(b));
      if (++b == p.nbuf)
      {
        b = 0;
        phase ^= 1;
      }
    }
    return;
  }

  const int n = p.n, ndofs = p.ndofs, pitch = p.pitch, r0 = p.r0, len = p.len, cs = p.cell_size;
  // warp-per-cell form: the descriptors of this lane's entries never change -> registers
  [[maybe_unused]] uint32_t shift[EPL > 0 ? EPL : 1], mask[EPL > 0 ? EPL : 1], gbase[EPL > 0 ? EPL : 1],
      gdim[EPL > 0 ? EPL : 1];
  if constexpr (EPL > 0)
  {
#pragma unroll
    for (int k = 0; k < EPL; ++k)
    {
      const int e = lane + 32 * k;
      const uint32_t info = e < len ? s_dinfo[r0 + e] : 0u;
      shift[k] = info & 31u;
      mask[k] = (info >> 5) & 7u;
      gdim[k] = (info >> 8) & 0xFFu;
      gbase[k] = info >> 16;
    }
  }
  int b = 0;
  uint32_t phase = 0;
  for (size_t i = 0; i < mine; ++i)
  {
    const size_t c0 = (blockIdx.x + i * gridDim.x) * size_t(p.cpt);
    const unsigned ncell = p.skip_apply ? 0u : unsigned(min(size_t(p.cpt), p.nc - c0));
    const U* tile = reinterpret_cast<const U*>(tiles + size_t(b) * tile_bytes);
    const uint32_t* ci = s_ci + b * ci_stride;
    mbar_wait(bar0 + 8 * b, phase);
    if constexpr (EPL > 0)
    {
      U* g = data + c0 * cs + r0 + lane;
      for (unsigned c = warp; c < ncell; c += kComputeThreads / 32)
      {
        const uint32_t info = ci[c];
        const U* row = tile + c * pitch - r0;
#pragma unroll
        for (int k = 0; k < EPL; ++k)
        {
          const unsigned state = (info >> shift[k]) & mask[k];
          if (state != 0)
            g[size_t(c) * cs + 32 * k] = row[s_gather[gbase[k] + state * gdim[k]]];
        }
      }
    }
    else
    {
      U* g = data + c0 * cs;
      for (unsigned q = threadIdx.x; q < ncell * len; q += kComputeThreads)
      {
        const unsigned c = p.by_len.div(q), e = r0 + (q - c * len);
        unsigned dof, bl;
        if constexpr (RIGHT)
        {
          bl = p.by_ndofs.div(e);
          dof = e - bl * ndofs;
        }
        else
        {
          dof = p.by_n.div(e);
          bl = e - dof * n;
        }
        const uint32_t info = s_dinfo[dof];
        const unsigned state = (ci[c] >> (info & 31u)) & ((info >> 5) & 7u);
        if (state == 0) // also every dof that never moves (mask 0)
          continue;
        const unsigned src = unsigned(s_gather[(info >> 16) + state * ((info >> 8) & 0xFFu)]);
        g[size_t(c) * cs + e] = tile[c * pitch + entry<RIGHT>(src, bl, n, ndofs) - r0];
      }
    }
    mbar_arrive(bar0 + 8 * (kMaxBuffers + b));
    if (++b == p.nbuf)
    {
      b = 0;
      phase ^= 1;
    }
  }
}
} // namespace tma
} // namespace bx

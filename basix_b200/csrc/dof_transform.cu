// Per-cell DOF transformations driven by cell_info bits, batched over cells.
//
// Replaces T_apply and its seven siblings (reference cpp/basix/finite-element.h:1839-1998), i.e.
// permute_data (finite-element.h:1703-1760) / transform_data (finite-element.h:1762-1837) and the
// precompute::apply_* kernels (precompute.h:136-374), plus permute / permute_inv
// (finite-element.h:800-843).
//
// Two implementations (HBM-bound, integer-indexed):
//   * dof_transform_tma.cuh   pipelined kernels fed by bulk-async (TMA) copies; taken whenever the per-cell
//                             ranges are 16-byte aligned -- the case for every BASELINE.json configuration;
//   * dof_transform_kernel    (this file) the general fallback for any alignment, block size and entity
//                             dimension.  A persistent CTA walks tiles of whole cells:
//   1. stage   the smallest contiguous range of each cell that holds every movable dof is copied to shared
//              memory with coalesced loads (entries that can never move are not read at all); cells are laid
//              out with an ODD pitch, so that 32 lanes working on 32 consecutive cells never share a bank;
//   2. apply   one thread owns one (cell, sub-entity, block) item, lanes <-> consecutive cells.  It reads the
//              entity's `dim` values into registers, then
//                * permutation elements: writes them back through the gather table of the entity's cell_info
//                  bit-field (bit-exact for any data type);
//                * matrix elements: multiplies them by the dense net operator of that bit-field, composed on the
//                  host (element.cu).  All states of all entity classes sit in shared memory, `mat_state_stride`
//                  apart, so lanes in the same state broadcast and lanes in different states hit different
//                  banks: one 16-byte shared load feeds two FMAs;
//   3. store   the range goes back with coalesced stores.
// Each global byte of the movable range is read once and written once; nothing else is touched.
#include "dof_transform_tma.cuh"
#include "dof_transform_direct.cuh"
#include "element.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>

namespace bx
{
namespace
{
constexpr int kThreads = 256;
constexpr int kMaxRegDim = 32; // largest entity dimension held in registers

struct FastDiv
{
  unsigned m = 0, d = 1;
  FastDiv() = default;
  // multiply-high form: exact whenever q * d < 2^32; `qmax` is the largest dividend the caller will pass.
  // Outside that range (huge block sizes) m stays 0 and div() falls back to a hardware division.
  __host__ FastDiv(unsigned d_, unsigned long long qmax)
      : m(d_ > 1 and qmax * d_ < (1ull << 32) ? unsigned((1ull << 32) / d_ + 1) : 0), d(d_)
  {
  }
  __device__ __forceinline__ unsigned div(unsigned q) const { return d > 1 ? (m ? __umulhi(q, m) : q / d) : q; }
};

struct TransformParams
{
  size_t nc;
  int cell_size, n, ndofs;
  int cpt_log2;       // cells per tile = 1 << cpt_log2
  int r0, len, pitch; // staged range [r0, r0 + len) of each cell (in entries), shared-memory pitch (odd)
  int nslots, table_size;
  FastDiv by_len, by_n;
  const uint32_t* cell_info;
  const int4* slot_info;  // {shift, mask, dim, gather base}
  const int2* slot_mat;   // {first dof, matrix base of the slot's class}
  const int16_t* gather;  // op offset applied (permutation elements)
  const double* mats;     // op offset applied (matrix elements)
};

// Entry index (within a cell) of (dof, block b): the left operators see u[ndofs][n], the right ones
// n contiguous blocks of ndofs entries (finite-element.h:1904-1998).
template <bool RIGHT>
__device__ __forceinline__ int entry(int dof, int b, int n, int ndofs)
{
  return RIGHT ? b * ndofs + dof : dof * n + b;
}

// REG = number of entity values held in registers (>= the largest entity dimension), 0 = any dimension:
// results then go straight to global memory and the staged tile stays read-only.
template <typename T, bool RIGHT, bool MATRIX, int REG>
__global__ void __launch_bounds__(kThreads) dof_transform_kernel(T* __restrict__ data, const TransformParams p)
{
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // ---- carve shared memory: operator tables, slot descriptors, cell_info, tile --------------------------
  const int table_bytes = (MATRIX ? p.table_size * 8 : p.table_size * 2 + 15) & ~15;
  double* s_mats = reinterpret_cast<double*>(smem_raw);
  int16_t* s_gather = reinterpret_cast<int16_t*>(smem_raw);
  int4* s_info = reinterpret_cast<int4*>(smem_raw + table_bytes);
  int2* s_mat = reinterpret_cast<int2*>(s_info + p.nslots);
  const int cpt = 1 << p.cpt_log2;
  uint32_t* s_ci = reinterpret_cast<uint32_t*>(s_mat + p.nslots + (p.nslots & 1));
  T* tile = reinterpret_cast<T*>(s_ci + ((cpt + 3) & ~3));

  if constexpr (MATRIX)
    for (int i = threadIdx.x; i < p.table_size; i += kThreads)
      s_mats[i] = p.mats[i];
  else
    for (int i = threadIdx.x; i < p.table_size; i += kThreads)
      s_gather[i] = p.gather[i];
  for (int i = threadIdx.x; i < p.nslots; i += kThreads)
  {
    s_info[i] = p.slot_info[i];
    s_mat[i] = p.slot_mat[i];
  }

  const int cs = p.cell_size, n = p.n, ndofs = p.ndofs, pitch = p.pitch, r0 = p.r0, len = p.len;
  const size_t ntiles = (p.nc + cpt - 1) >> p.cpt_log2;
  const unsigned nitems = unsigned(p.nslots * n) << p.cpt_log2;
  for (size_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x)
  {
    const size_t c0 = tl << p.cpt_log2;
    const unsigned ncell = unsigned(min(size_t(cpt), p.nc - c0));
    T* g = data + c0 * cs + r0;
    __syncthreads(); // previous tile fully stored (and, first time, tables loaded)
    // ---- 1. stage -------------------------------------------------------------------------------------
    for (unsigned q = threadIdx.x; q < ncell * len; q += kThreads)
    {
      const unsigned c = p.by_len.div(q), k = q - c * len;
      tile[c * pitch + k] = g[size_t(c) * cs + k];
    }
    for (unsigned c = threadIdx.x; c < ncell; c += kThreads)
      s_ci[c] = p.cell_info[c0 + c];
    __syncthreads();
    // ---- 2. apply: item = (slot, block, cell), cell fastest -----------------------------------------------
    for (unsigned it = threadIdx.x; it < nitems; it += kThreads)
    {
      const unsigned c = it & (cpt - 1), sb = it >> p.cpt_log2;
      if (c >= ncell)
        continue;
      const unsigned s = p.by_n.div(sb), b = sb - s * n;
      const int4 si = s_info[s];
      const unsigned state = (s_ci[c] >> si.x) & unsigned(si.y);
      if (state == 0)
        continue;
      const int dim = si.z;
      T* cell = tile + c * pitch - r0;
      if constexpr (!MATRIX)
      {
        const int16_t* src = s_gather + si.w + state * dim; // source dofs of this state
        const int16_t* dst = s_gather + si.w;               // state 0 is the identity: the entity's own dofs
        if constexpr (REG > 0)
        {
          T v[REG];
#pragma unroll
          for (int l = 0; l < REG; ++l)
            if (l < dim)
              v[l] = cell[entry<RIGHT>(src[l], b, n, ndofs)];
#pragma unroll
          for (int l = 0; l < REG; ++l)
            if (l < dim)
              cell[entry<RIGHT>(dst[l], b, n, ndofs)] = v[l];
        }
        else
        {
          T* gc = data + (c0 + c) * cs;
          for (int l = 0; l < dim; ++l)
            gc[entry<RIGHT>(dst[l], b, n, ndofs)] = cell[entry<RIGHT>(src[l], b, n, ndofs)];
        }
      }
      else
      {
        const int2 sm = s_mat[s];
        const int rs = mat_row_stride(dim);
        const double* M = s_mats + sm.y + state * mat_state_stride(dim);
        if constexpr (REG > 0)
        {
          double v[REG];
#pragma unroll
          for (int j = 0; j < REG; ++j)
            v[j] = j < dim ? static_cast<double>(cell[entry<RIGHT>(sm.x + j, b, n, ndofs)]) : 0.0;
          for (int i = 0; i < dim; ++i)
          {
            const double2* row = reinterpret_cast<const double2*>(M + i * rs);
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int jj = 0; jj < REG / 2; ++jj)
              if (2 * jj < dim)
              {
                const double2 m = row[jj];
                a0 = fma(m.x, v[2 * jj], a0);
                a1 = fma(m.y, v[2 * jj + 1], a1);
              }
            cell[entry<RIGHT>(sm.x + i, b, n, ndofs)] = static_cast<T>(a0 + a1);
          }
        }
        else
        {
          T* gc = data + (c0 + c) * cs;
          for (int i = 0; i < dim; ++i)
          {
            double a = 0.0;
            for (int j = 0; j < dim; ++j)
              a = fma(M[i * rs + j], static_cast<double>(cell[entry<RIGHT>(sm.x + j, b, n, ndofs)]), a);
            gc[entry<RIGHT>(sm.x + i, b, n, ndofs)] = static_cast<T>(a);
          }
        }
      }
    }
    // ---- 3. store ---------------------------------------------------------------------------------------
    if constexpr (REG > 0)
    {
      __syncthreads();
      for (unsigned q = threadIdx.x; q < ncell * len; q += kThreads)
      {
        const unsigned c = p.by_len.div(q), k = q - c * len;
        g[size_t(c) * cs + k] = tile[c * pitch + k];
      }
    }
  }
}

size_t smem_bytes(const TransformParams& p, bool matrix, size_t elem)
{
  const size_t table = (matrix ? size_t(p.table_size) * 8 : size_t(p.table_size) * 2 + 15) & ~size_t(15);
  const size_t cpt = size_t(1) << p.cpt_log2;
  return table + size_t(p.nslots) * 16 + size_t(p.nslots + (p.nslots & 1)) * 8 + ((cpt + 3) & ~size_t(3)) * 4
         + cpt * p.pitch * elem;
}

template <typename T, bool RIGHT, bool MATRIX, int REG>
void launch(T* data, const TransformParams& p, cudaStream_t s)
{
  auto kern = dof_transform_kernel<T, RIGHT, MATRIX, REG>;
  const size_t smem = smem_bytes(p, MATRIX, sizeof(T));
  if (smem > 48 * 1024)
    BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 0;
  BX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, kThreads, smem));
  const size_t ntiles = (p.nc + (size_t(1) << p.cpt_log2) - 1) >> p.cpt_log2;
  const size_t cap = size_t(sm_count()) * std::max(per_sm, 1);
  kern<<<unsigned(std::min(ntiles, cap)), kThreads, smem, s>>>(data, p);
  BX_LAUNCHED();
}

template <typename T, bool RIGHT, bool MATRIX>
void launch_by_dim(T* data, const TransformParams& p, int maxdim, cudaStream_t s)
{
  if (maxdim <= 8)
    launch<T, RIGHT, MATRIX, 8>(data, p, s);
  else if (maxdim <= 16)
    launch<T, RIGHT, MATRIX, 16>(data, p, s);
  else if (maxdim <= kMaxRegDim)
    launch<T, RIGHT, MATRIX, kMaxRegDim>(data, p, s);
  else
    launch<T, RIGHT, MATRIX, 0>(data, p, s);
}

// ---- pipelined (TMA) path --------------------------------------------------------------------------------
using EncodeTiled = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                 const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
EncodeTiled encode_tiled()
{
  static EncodeTiled fn = []() -> EncodeTiled
  {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess
        or q != cudaDriverEntryPointSuccess)
    {
      cudaGetLastError();
      return nullptr;
    }
    return reinterpret_cast<EncodeTiled>(f);
  }();
  return fn;
}

// data[nc][cell_size] as a 2-D tensor, box = [cpt cells][box_w entries]
// Tuning aid (development only): BX_DOF_TUNE="cpt=64,nbuf=3,promo=0,skip=0,pad=1" overrides the tile shape, the
// number of tile buffers, the L2 promotion of the tensor map, can skip the arithmetic (data movement only) and sets the
// number of 16-byte units the box is padded by (pad=0: never).
struct Tune
{
  int cpt = 0, nbuf = 0, promo = -1, skip = 0, pad = -1, direct = -1;
};
const Tune& tune()
{
  static Tune t = []()
  {
    Tune v;
    if (const char* e = std::getenv("BX_DOF_TUNE"))
    {
      std::string str(e);
      auto get = [&](const char* key, int& out)
      {
        const size_t k = str.find(key);
        if (k != std::string::npos)
          out = std::atoi(str.c_str() + k + std::strlen(key));
      };
      get("cpt=", v.cpt);
      get("nbuf=", v.nbuf);
      get("promo=", v.promo);
      get("skip=", v.skip);
      get("pad=", v.pad);
      get("direct=", v.direct);
    }
    return v;
  }();
  return t;
}

bool make_tensor_map(CUtensorMap& map, void* data, size_t nc, size_t cell_size, size_t elem, int box_w, int cpt)
{
  EncodeTiled enc = encode_tiled();
  if (!enc)
    return false;
  const cuuint64_t gdim[2] = {cuuint64_t(cell_size), cuuint64_t(nc)};
  const cuuint64_t gstride[1] = {cuuint64_t(cell_size * elem)};
  const cuuint32_t box[2] = {cuuint32_t(box_w), cuuint32_t(cpt)};
  const cuuint32_t estride[2] = {1, 1};
  const CUtensorMapDataType dt = elem == 8 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32;
  return enc(&map, dt, 2, data, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             tune().promo >= 0 ? CUtensorMapL2promotion(tune().promo) : CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE)
         == CUDA_SUCCESS;
}

template <typename K>
unsigned tma_grid(K kern, const tma::Params& p, size_t smem)
{
  if (smem > 48 * 1024)
    BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 0;
  BX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, tma::kThreads, smem));
  const size_t ntiles = (p.nc + p.cpt - 1) / p.cpt;
  return unsigned(std::min(ntiles, size_t(sm_count()) * std::max(per_sm, 1)));
}

template <typename T, bool RIGHT, bool PAIRS, int REG>
void launch_matrix_tma_reg(const CUtensorMap& map, const tma::Params& p, size_t smem, cudaStream_t s)
{
  auto kern = tma::dof_matrix_tma_kernel<T, RIGHT, REG, PAIRS>;
  kern<<<tma_grid(kern, p, smem), tma::kThreads, smem, s>>>(map, p);
  BX_LAUNCHED();
}

template <typename T, bool RIGHT, bool PAIRS>
void launch_matrix_tma(const CUtensorMap& map, const tma::Params& p, int maxdim, size_t smem, cudaStream_t s)
{
  if (maxdim <= 4)
    launch_matrix_tma_reg<T, RIGHT, PAIRS, 4>(map, p, smem, s);
  else if (maxdim <= 8)
    launch_matrix_tma_reg<T, RIGHT, PAIRS, 8>(map, p, smem, s);
  else if (maxdim <= 12)
    launch_matrix_tma_reg<T, RIGHT, PAIRS, 12>(map, p, smem, s);
  else if (maxdim <= 16)
    launch_matrix_tma_reg<T, RIGHT, PAIRS, 16>(map, p, smem, s);
  else if (maxdim <= 24)
    launch_matrix_tma_reg<T, RIGHT, PAIRS, 24>(map, p, smem, s);
  else
    launch_matrix_tma_reg<T, RIGHT, PAIRS, 32>(map, p, smem, s);
}

template <typename U, bool RIGHT, int EPL>
void launch_perm_tma(const CUtensorMap& map, void* data, const tma::Params& p, size_t smem, cudaStream_t s)
{
  auto kern = tma::dof_perm_tma_kernel<U, RIGHT, EPL>;
  kern<<<tma_grid(kern, p, smem), tma::kThreads, smem, s>>>(map, static_cast<U*>(data), p);
  BX_LAUNCHED();
}

template <typename U>
void launch_perm_tma_any(const CUtensorMap& map, void* data, const tma::Params& p, bool right, size_t smem,
                         cudaStream_t s)
{
  if (right)
    launch_perm_tma<U, true, 0>(map, data, p, smem, s);
  else if (p.n == 1 and p.len <= 32)
    launch_perm_tma<U, false, 1>(map, data, p, smem, s);
  else if (p.n == 1 and p.len <= 64)
    launch_perm_tma<U, false, 2>(map, data, p, smem, s);
  else if (p.n == 1 and p.len <= 128)
    launch_perm_tma<U, false, 4>(map, data, p, smem, s);
  else
    launch_perm_tma<U, false, 0>(map, data, p, smem, s);
}

// Returns false when the layout does not meet the TMA rules (the caller then takes the register-staged kernel).
template <int REG>
void launch_matrix_direct_reg(double* data, const direct::Params& p, size_t smem, cudaStream_t s)
{
  auto kern = direct::dof_matrix_direct_kernel<REG>;
  if (smem > 48 * 1024)
    BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  int per_sm = 0;
  BX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, direct::kThreads, smem));
  const size_t nitems = p.nc * size_t(p.nslots);
  const size_t nblocks = ceil_div(nitems, size_t(direct::kThreads));
  // a few waves of resident CTAs: each CTA pays the table set-up once
  const size_t cap = size_t(sm_count()) * size_t(std::max(per_sm, 1)) * 8;
  kern<<<unsigned(std::min(nblocks, cap)), direct::kThreads, smem, s>>>(data, p);
  BX_LAUNCHED();
}

// float64, left operators, block size 1, even entity starts, 16-byte aligned cells: the register kernel
bool try_direct(const bx_element& e, int kind, double* data, size_t nc, size_t cell_size, int maxdim,
                const uint32_t* cell_info, cudaStream_t s)
{
  if (reinterpret_cast<uintptr_t>(data) % 16 != 0 or (cell_size * 8) % 16 != 0 or maxdim > kMaxRegDim or !e.d_mats)
    return false;
  for (const auto& sl : e.slots)
    if (sl.first % 2 != 0)
      return false;
  direct::Params p;
  p.nc = nc;
  p.cell_size = cell_size;
  p.nslots = int(e.slots.size());
  p.table_size = e.mats_size;
  p.band_size = e.band_size;
  p.cell_info = cell_info;
  p.slot_info = e.d_slot_info;
  p.slot_mat = e.d_slot_mat;
  p.slot_band = e.d_slot_band;
  p.band = e.d_band;
  p.mats = e.d_mats + size_t(kind) * e.mats_size;
  const size_t smem = size_t(e.mats_size) * 8 + size_t(e.band_size + (e.band_size & 1)) * 8 + size_t(p.nslots) * 32;
  if (smem > 200 * 1024)
    return false;
  if (maxdim <= 4)
    launch_matrix_direct_reg<4>(data, p, smem, s);
  else if (maxdim <= 8)
    launch_matrix_direct_reg<8>(data, p, smem, s);
  else if (maxdim <= 12)
    launch_matrix_direct_reg<12>(data, p, smem, s);
  else if (maxdim <= 16)
    launch_matrix_direct_reg<16>(data, p, smem, s);
  else if (maxdim <= 24)
    launch_matrix_direct_reg<24>(data, p, smem, s);
  else
    launch_matrix_direct_reg<32>(data, p, smem, s);
  return true;
}

template <typename T>
bool try_tma(const bx_element& e, int kind, bool right, T* data, size_t nc, size_t cell_size, int n, int r0, int len,
             int maxdim, const uint32_t* cell_info, cudaStream_t s)
{
  constexpr size_t sz = sizeof(T);
  constexpr int per16 = int(16 / sz);
  if constexpr (std::is_same<T, double>::value)
    if (!e.perms and !right and n == 1 and tune().direct > 0 and !tune().skip
        and try_direct(e, kind, data, nc, cell_size, maxdim, cell_info, s))
      return true;
  if (reinterpret_cast<uintptr_t>(data) % 16 != 0 or (cell_size * sz) % 16 != 0 or nc >= (size_t(1) << 31))
    return false;
  if (maxdim > kMaxRegDim or e.gather_size >= 65536 or cell_size >= (size_t(1) << 20))
    return false;
  if (e.perms and !e.d_dinfo)
    return false;
  tma::Params p;
  p.nc = nc;
  p.cell_size = int(cell_size);
  p.n = n;
  p.ndofs = e.ndofs;
  p.r0 = r0 / per16 * per16;
  p.len = r0 + len - p.r0;
  // box width = shared-memory distance between consecutive cells.  Lanes are consecutive cells, so the number of
  // distinct bank offsets they hit is 128 / gcd(pitch bytes, 128).  The matrix kernel with 8-byte accesses (block size
  // > 1, right layout, float data) runs 16 lanes per wavefront: at an 8-way conflict the shared-memory pipe, not HBM,
  // sets the pace (RT4 tetrahedron, block size 3, 960-byte rows: 29.4 -> 20.2 ms per 50 M cells with one more 16-byte
  // unit).  With 16-byte accesses (block size 1, left layout, double) the conflict is at most 4-way and the kernel
  // is bound by the HBM stream: padding only adds bytes (RT4, 320-byte rows: 7.3 -> 7.6 ms), so it is not padded.
  size_t units = (size_t(p.len) * sz + 15) / 16;
  {
    auto gcd = [](size_t a, size_t b)
    {
      while (b)
      {
        const size_t t = a % b;
        a = b;
        b = t;
      }
      return a;
    };
    const bool wide = std::is_same<T, double>::value and !right and n == 1; // 16-byte accesses (PAIRS form)
    const size_t distinct = 128 / gcd(units * 16, 128), lanes = wide ? 8 : 128 / sz > 32 ? 32 : 128 / sz;
    const size_t ways = lanes / std::min(lanes, distinct);
    if (!e.perms and tune().pad < 0 and ways >= 8)
      units += 1;
  }
  if (tune().pad > 0)
    units += size_t(tune().pad);
  const size_t pitch_bytes = units * 16;
  p.pitch = int(pitch_bytes / sz);
  if (p.pitch > 256)
    return false;
  p.nslots = int(e.slots.size());
  p.table_size = e.perms ? e.gather_size : e.mats_size;
  p.band_size = e.band_size;
  const size_t tables = e.perms ? ((size_t(e.ndofs) * 4 + size_t(e.gather_size) * 2 + 15) & ~size_t(15))
                                : size_t(e.mats_size) * 8 + size_t(e.band_size + (e.band_size & 1)) * 8
                                      + size_t(p.nslots) * 32;
  // cells per tile: ~24 KB per buffer (three buffers per CTA, several CTAs per SM), whole warps of cells
  // buffers: measured on B200 (tools/sweep_dof.sh): 2 is best for the matrix kernel (4 CTAs per SM), 3 for the
  // permutation kernel
  p.nbuf = tune().nbuf >= 2 and tune().nbuf <= tma::kMaxBuffers ? tune().nbuf : (e.perms ? 3 : 2);
  p.skip_apply = tune().skip;
  size_t cpt = std::min<size_t>(256, (24 * 1024) / pitch_bytes);
  if (cpt >= 32)
    cpt = cpt / 32 * 32;
  auto total = [&](size_t c)
  {
    return tables + 2 * tma::kMaxBuffers * 8 + tma::kMaxBuffers * ((c + 3) & ~size_t(3)) * 4 + 128
           + p.nbuf * size_t(tma::tile_stride_bytes(int(c), p.pitch, int(sz)));
  };
  while (cpt > 1 and total(cpt) > 100 * 1024)
    cpt /= 2;
  if (tune().cpt > 0 and tune().cpt <= 256)
    cpt = size_t(tune().cpt);
  if (cpt < 8 or total(cpt) > 200 * 1024)
    return false;
  p.cpt = int(cpt);
  const unsigned long long nitems = (unsigned long long)(p.nslots) * n * cpt;
  if (nitems >= (1ull << 31) or (unsigned long long)(p.len) * cpt >= (1ull << 31))
    return false;
  p.by_cpt = tma::FastDiv(unsigned(cpt), nitems);
  p.by_len = tma::FastDiv(unsigned(p.len), (unsigned long long)(p.len) * cpt);
  p.by_n = tma::FastDiv(unsigned(n), std::max<unsigned long long>((unsigned long long)(p.nslots) * n, cell_size));
  p.by_ndofs = tma::FastDiv(unsigned(e.ndofs), cell_size);
  p.cell_info = cell_info;
  p.slot_info = e.d_slot_info;
  p.slot_mat = e.d_slot_mat;
  p.slot_band = e.d_slot_band;
  p.band = e.d_band;
  p.mats = e.d_mats ? e.d_mats + size_t(kind) * e.mats_size : nullptr;
  p.gather = e.d_gather ? e.d_gather + size_t(kind) * e.gather_size : nullptr;
  p.dinfo = e.d_dinfo;
  const size_t smem = total(cpt);
  CUtensorMap map;
  if (!make_tensor_map(map, data, nc, cell_size, sz, p.pitch, p.cpt))
    return false;

  if (e.perms)
  {
    if constexpr (sz == 4)
      launch_perm_tma_any<uint32_t>(map, data, p, right, smem, s);
    else
      launch_perm_tma_any<unsigned long long>(map, data, p, right, smem, s);
    return true;
  }
  if constexpr (std::is_integral<T>::value)
    return false;
  else
  {
    bool pairs = std::is_same<T, double>::value and !right and n == 1;
    for (const auto& sl : e.slots)
      pairs = pairs and sl.first % 2 == 0;
    if (right)
      launch_matrix_tma<T, true, false>(map, p, maxdim, smem, s);
    else if constexpr (std::is_same<T, double>::value)
    {
      if (pairs)
        launch_matrix_tma<T, false, true>(map, p, maxdim, smem, s);
      else
        launch_matrix_tma<T, false, false>(map, p, maxdim, smem, s);
    }
    else
      launch_matrix_tma<T, false, false>(map, p, maxdim, smem, s);
    return true;
  }
}
} // namespace

// data[nc][cell_size] on the device, transformed in place.  op is a bx_top (0..7).
template <typename T>
void dof_transform_device(const bx_element& e, int op, T* data, size_t nc, size_t cell_size, int n,
                          const uint32_t* cell_info, cudaStream_t s)
{
  if (op < 0 or op > 7)
    fail(BX_ERR_INVALID_ARG, "T_apply: unknown operator");
  if (n <= 0)
    fail(BX_ERR_INVALID_ARG, "T_apply: block size must be positive");
  if (cell_size != size_t(e.ndofs) * n)
    fail(BX_ERR_INVALID_ARG, "T_apply: data size per cell (" + std::to_string(cell_size)
                                 + ") is not ndofs * block_size (" + std::to_string(size_t(e.ndofs) * n) + ")");
  if (e.identity or e.tdim < 2 or e.slots.empty() or nc == 0)
    return;
  require_device(e);
  if (std::is_integral<T>::value and !e.perms)
    fail(BX_ERR_UNSUPPORTED, "T_apply: integer data needs an element whose DOF transformations are permutations");
  // reference operator -> (composed operator kind, data layout); see element.cu
  static const int kind_of[8] = {0, 1, 2, 3, 0, 2, 1, 3};
  const bool right = op >= 4;
  const int kind = kind_of[op];

  TransformParams p;
  p.nc = nc;
  p.cell_size = int(cell_size);
  p.n = n;
  p.ndofs = e.ndofs;
  // staged range: left layout u[dof][b] -> [lo*n, hi*n); right layout u[b][dof] -> [lo, (n-1)*ndofs + hi)
  p.r0 = right ? e.lo_dof : e.lo_dof * n;
  p.len = right ? (n - 1) * e.ndofs + e.hi_dof - e.lo_dof : (e.hi_dof - e.lo_dof) * n;
  p.pitch = p.len | 1;
  if (p.len >= 65536 or n >= 65536)
    fail(BX_ERR_UNSUPPORTED, "T_apply: more than 65535 entries per cell");
  int maxdim = 0;
  for (const auto& sl : e.slots)
    maxdim = std::max(maxdim, sl.dim);
  if (try_tma<T>(e, kind, right, data, nc, cell_size, n, p.r0, p.len, maxdim, cell_info, s))
    return;

  p.nslots = int(e.slots.size());
  p.table_size = e.perms ? e.gather_size : e.mats_size;
  p.cell_info = cell_info;
  p.slot_info = e.d_slot_info;
  p.slot_mat = e.d_slot_mat;
  p.gather = e.d_gather ? e.d_gather + size_t(kind) * e.gather_size : nullptr;
  p.mats = e.d_mats ? e.d_mats + size_t(kind) * e.mats_size : nullptr;

  // cells per tile: a power of two, tile <= ~40 KB (several CTAs per SM hide the stage/apply/store phases
  // of one another), at least one cell, fast-division range respected, no larger than the batch
  const size_t cell_bytes = size_t(p.pitch) * sizeof(T);
  int lg = 8;
  while (lg > 0 and ((size_t(1) << lg) * cell_bytes > 40 * 1024 or (size_t(p.len) << lg) >= 65536
                     or (size_t(p.nslots) * n << lg) >= (size_t(1) << 31)))
    --lg;
  while (lg > 0 and (size_t(1) << (lg - 1)) >= nc)
    --lg;
  p.cpt_log2 = lg;
  p.by_len = FastDiv(unsigned(p.len), (unsigned long long)(p.len) << lg);
  p.by_n = FastDiv(unsigned(n), (unsigned long long)(p.nslots) * n);
  if (smem_bytes(p, !e.perms, sizeof(T)) > 200 * 1024)
  {
    // One cell is larger than shared memory: a caller transforming whole element tensors passes block sizes as large as
    // the element (T_apply(data, n = ndofs, ...): 216 x 216 doubles for a degree-5 hexahedron).  The operators act on
    // the dof axis and leave the block index alone, so the blocks are processed a slab at a time through a contiguous
    // copy: left layout u[dof][b] -> columns [b0, b0 + w) of every row; right layout u[b][dof] -> w consecutive blocks.
    const size_t row = size_t(e.ndofs) * sizeof(T);
    int w = int(std::min<size_t>(size_t(n), std::max<size_t>(1, (32 * 1024) / row)));
    if (n == 1 or row > 150 * 1024)
      fail(BX_ERR_UNSUPPORTED, "T_apply: a single block of the cell does not fit in shared memory");
    T* scratch = nullptr;
    BX_CUDA(cudaMallocAsync(&scratch, nc * size_t(e.ndofs) * w * sizeof(T), s));
    try
    {
      for (int b0 = 0; b0 < n; b0 += w)
      {
        const int wb = std::min(w, n - b0);
        if (right)
        {
          BX_CUDA(cudaMemcpy2DAsync(scratch, size_t(wb) * row, data + size_t(b0) * e.ndofs, cell_size * sizeof(T), size_t(wb) * row,
                                    nc, cudaMemcpyDeviceToDevice, s));
          dof_transform_device<T>(e, op, scratch, nc, size_t(e.ndofs) * wb, wb, cell_info, s);
          BX_CUDA(cudaMemcpy2DAsync(data + size_t(b0) * e.ndofs, cell_size * sizeof(T), scratch, size_t(wb) * row, size_t(wb) * row,
                                    nc, cudaMemcpyDeviceToDevice, s));
        }
        else
        {
          BX_CUDA(cudaMemcpy2DAsync(scratch, size_t(wb) * sizeof(T), data + b0, size_t(n) * sizeof(T), size_t(wb) * sizeof(T),
                                    nc * size_t(e.ndofs), cudaMemcpyDeviceToDevice, s));
          dof_transform_device<T>(e, op, scratch, nc, size_t(e.ndofs) * wb, wb, cell_info, s);
          BX_CUDA(cudaMemcpy2DAsync(data + b0, size_t(n) * sizeof(T), scratch, size_t(wb) * sizeof(T), size_t(wb) * sizeof(T),
                                    nc * size_t(e.ndofs), cudaMemcpyDeviceToDevice, s));
        }
      }
    }
    catch (...)
    {
      cudaFreeAsync(scratch, s);
      throw;
    }
    BX_CUDA(cudaFreeAsync(scratch, s));
    return;
  }

  if (e.perms)
  {
    if (right)
      launch_by_dim<T, true, false>(data, p, maxdim, s);
    else
      launch_by_dim<T, false, false>(data, p, maxdim, s);
  }
  else
  {
    if constexpr (std::is_integral<T>::value)
      fail(BX_ERR_UNSUPPORTED, "T_apply: integer data on a non-permutation element");
    else
    {
      if (right)
        launch_by_dim<T, true, true>(data, p, maxdim, s);
      else
        launch_by_dim<T, false, true>(data, p, maxdim, s);
    }
  }
}

// ---- permute_subentity_closure{,_inv} (finite-element.h:860-1035), batched over cells ----------------------
// d[nc][m]: the dofs on the closure of one sub-entity per cell; info[c] is the entity's own bits (shift < 0) or
// the cell's cell_info, from which the entity's bits are (info >> shift) & mask.  The net gather of every state
// was composed on the host (element.cu); a CTA stages `cpb` cells through shared memory so the in-place gather
// reads every entry before any is overwritten, global accesses are coalesced runs of cpb * m entries.
namespace
{
constexpr int kClosureThreads = 256;
__global__ void __launch_bounds__(kClosureThreads)
    closure_permute_kernel(int32_t* __restrict__ d, size_t nc, int m, int cpb, const uint32_t* __restrict__ info,
                           int shift, uint32_t mask, const int16_t* __restrict__ src, int states, unsigned inv_m)
{
  extern __shared__ int32_t closure_smem[];
  int16_t* s_src = reinterpret_cast<int16_t*>(closure_smem);                     // [states][m]
  int32_t* tile = closure_smem + ((states * m + 1) / 2 + 3) / 4 * 4;              // [cpb][m]
  uint32_t* s_state = reinterpret_cast<uint32_t*>(tile + size_t(cpb) * m);       // [cpb]
  for (int i = threadIdx.x; i < states * m; i += kClosureThreads)
    s_src[i] = src[i];
  const size_t ntiles = (nc + cpb - 1) / cpb;
  for (size_t tl = blockIdx.x; tl < ntiles; tl += gridDim.x)
  {
    const size_t c0 = tl * cpb;
    const unsigned ncell = unsigned(min(size_t(cpb), nc - c0));
    __syncthreads(); // previous tile fully written back (and s_src ready on the first pass)
    for (unsigned c = threadIdx.x; c < ncell; c += kClosureThreads)
    {
      const uint32_t v = info[c0 + c];
      s_state[c] = (shift < 0 ? v : (v >> shift)) & mask;
    }
    int32_t* g = d + c0 * m;
    const unsigned run = ncell * unsigned(m);
    for (unsigned q = threadIdx.x; q < run; q += kClosureThreads)
      tile[q] = g[q];
    __syncthreads();
    for (unsigned q = threadIdx.x; q < run; q += kClosureThreads)
    {
      const unsigned c = inv_m ? __umulhi(q, inv_m) : q, k = q - c * unsigned(m); // c = q / m (exact: q < 2^16)
      const unsigned st = s_state[c];
      if (st != 0)
        g[q] = tile[c * m + s_src[st * m + k]];
    }
  }
}
} // namespace

void closure_permute_device(const bx_element& e, int inverse, int32_t* d, size_t nc, size_t m, const uint32_t* info,
                            int entity_type, int entity_index, cudaStream_t s)
{
  // entity_dim == 0: nothing moves; the cell_info overload returns before looking at the element
  // (finite-element.h:871-875), the entity_info overload after the permutation check (950-960)
  if (entity_index >= 0 and entity_type == BX_CELL_POINT)
    return;
  if (!e.perms)
    fail(BX_ERR_RUNTIME, "The DOF transformations for this element are not permutations");
  if (entity_type == BX_CELL_POINT)
    return;
  if (entity_type != BX_CELL_INTERVAL and entity_type != BX_CELL_TRIANGLE and entity_type != BX_CELL_QUADRILATERAL)
    fail(BX_ERR_RUNTIME, entity_index < 0 ? "Invalid dimension for permute_subentity_closure" : "Unsupported cell dimension");
  const bx_element::Closure& cl = e.closure[entity_type];
  if (cl.size == 0)
    fail(BX_ERR_RUNTIME, "permute_subentity_closure: the cell has no sub-entity of this type");
  if (m != size_t(cl.size))
    fail(BX_ERR_INVALID_ARG, "permute_subentity_closure: row length does not match the closure of the sub-entity");
  int shift = -1;
  if (entity_index >= 0)
  {
    const int edim = entity_type == BX_CELL_INTERVAL ? 1 : 2;
    if (edim >= int(e.edofs.size()) or entity_index >= int(e.edofs[edim].size()))
      fail(BX_ERR_INVALID_ARG, "permute_subentity_closure: entity index out of range");
    const int face_start = e.tdim == 3 ? 3 * int(e.edofs[2].size()) : 0;
    shift = edim == 1 ? face_start + entity_index : 3 * entity_index;
  }
  if (nc == 0)
    return;
  require_device(e);
  const int cpb = std::max(1, std::min(256, 16384 / cl.size)); // cpb * m < 2^16 entries per tile
  const size_t smem = (size_t((cl.states * cl.size + 1) / 2 + 3) / 4 * 4 + size_t(cpb) * cl.size + cpb) * 4;
  if (smem > 96 * 1024)
    fail(BX_ERR_UNSUPPORTED, "permute_subentity_closure: closure too large");
  if (smem > 48 * 1024)
    BX_CUDA(cudaFuncSetAttribute(closure_permute_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const size_t ntiles = (nc + cpb - 1) / cpb;
  const unsigned grid = unsigned(std::min(ntiles, size_t(sm_count()) * 4));
  const unsigned inv_m = cl.size > 1 ? unsigned((1ull << 32) / unsigned(cl.size) + 1) : 0u;
  closure_permute_kernel<<<grid, kClosureThreads, smem, s>>>(d, nc, cl.size, cpb, info, shift,
                                                             entity_type == BX_CELL_INTERVAL ? 1u : 7u,
                                                             cl.d_src + size_t(inverse ? 1 : 0) * cl.states * cl.size,
                                                             cl.states, inv_m);
  BX_LAUNCHED();
}

template void dof_transform_device<double>(const bx_element&, int, double*, size_t, size_t, int, const uint32_t*,
                                           cudaStream_t);
template void dof_transform_device<float>(const bx_element&, int, float*, size_t, size_t, int, const uint32_t*,
                                          cudaStream_t);
template void dof_transform_device<int32_t>(const bx_element&, int, int32_t*, size_t, size_t, int,
                                            const uint32_t*, cudaStream_t);
} // namespace bx

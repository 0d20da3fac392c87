// Derivative-operator tabulate for simplex elements: every derivative slab from the VALUES of the orthonormal set.
//
// Contract (reference cpp/basix/finite-element.cpp:1834-1888):
//   basis[d][pt][n] = sum_k C[n][k] * d^alpha P_k(x_pt),   alpha = derivative multi-index of slab d.
// The reference tabulates all derivatives of all K basis functions (polyset.cpp:1601-2020) and contracts each of
// the nder slabs with the full K-column coefficient matrix.  On a simplex the orthonormal set is hierarchical by
// total degree (indexing.h:13-34 orders by p+q(+r)), and a derivative of order m of a polynomial of degree n lies
// in P_{n-m}; so with D_a[k][j] = int d_a P_k P_j (first-derivative operators in the orthonormal basis, exact
// Gauss-Jacobi quadrature on the host)
//   d^alpha P_k = sum_{j < dim P_{n-m}} (prod_a D_a^{alpha_a})[k][j] P_j
//   basis[d][pt][n] = sum_{j < dim P_{n-m}} B_d[n][j] P_j(x_pt),     B_d = C * prod_a D_a^{alpha_a}   (host, once).
// Consequences on the device:
//   * the producers evaluate only VALUES of the orthonormal set (no derivative jets: 10x less recurrence work
//     for nd = 2), once per point tile, into shared memory;
//   * slab d contracts over dim P_{n-m} columns instead of K: for P5 on the tetrahedron with nd = 2 the k-extent
//     summed over the 10 slabs is 56 + 3*35 + 6*20 = 281 instead of 560 -- half the FP64 tensor-core work;
//   * all B_d stay resident in shared memory for the whole kernel; vector-valued elements (N > K) fit too.
// Kernel: persistent, warp specialised.  2 producer warps (lane = point, register recurrences of jets.cuh) fill a
// double-buffered 64-point tile P[k][pt]; 8 consumer warps each own 16 points (two m8 row tiles, so every
// B-fragment load feeds two DMMAs) and half of the (slab, column-group) work items; accumulators live in
// registers for one item at a time and the m8n8 fragment is stored straight as basis[d][pt][n].
// Parity: 2.8e-14 * max|ref| per slab against the reference on P5 / nd = 2 (bar: 1e-12).
#include "element.cuh"
#include "hostmath.cuh"
#include "jets.cuh"

#include <algorithm>
#include <map>
#include <mutex>

namespace bx
{
namespace dop
{
constexpr int kTilePts = 64;
constexpr int kConsumers = 8;
constexpr int kProducers = 2;
constexpr int kThreads = 32 * (kConsumers + kProducers);
constexpr int kPStride = 68; // doubles between consecutive basis rows of a point tile (== 4 mod 16: conflict free)
constexpr int kMaxNT = 8;    // 8-column output tiles per work item
constexpr int kMaxItems = 96;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t a, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(a), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t a)
{
  asm volatile("{\n .reg .b64 st;\n mbarrier.arrive.shared::cta.b64 st, [%0];\n}" ::"r"(a) : "memory");
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
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

// One work item: NT column tiles of slab `slab`, contraction over 4 * ksteps rows of the point tile.
struct Item
{
  int boff;    // offset (doubles) of B_slab row 0 inside the operand block
  int ksteps;  // ceil(dim P_{n-m} / 4)
  int kstride; // row stride of B_slab (>= 4 * ksteps, == 4 mod 16)
  int n0;      // first output column
  int nt;      // column tiles (<= kMaxNT)
  int slab;    // derivative slab index
  int pad0, pad1;
};

struct Params
{
  const double* x;
  size_t npts;
  const double* B;       // all slabs
  const Item* items;     // half 0 first, then half 1
  const int* rowmap;     // production order -> natural basis index
  double* basis;
  int N, K, kpad;        // output columns, basis functions, rows of the point tile (multiple of 4)
  int b_doubles, nitems, split; // items [0, split) belong to half 0, [split, nitems) to half 1
};

template <int K>
struct TileSink
{
  double* col;       // &P[0][this point]
  const int* rowmap; // shared memory
  __device__ __forceinline__ void operator()(int s, const double (&J)[1]) const { col[rowmap[s] * kPStride] = J[0]; }
};

template <int NT>
__device__ __forceinline__ void run_item(const Item& it, const double* __restrict__ Pa, const double* __restrict__ Bs,
                                         double* __restrict__ basis, size_t npts, int N, size_t pt, int g, int t)
{
  double acc[2][NT][2];
#pragma unroll
  for (int m = 0; m < 2; ++m)
#pragma unroll
    for (int j = 0; j < NT; ++j)
      acc[m][j][0] = acc[m][j][1] = 0.0;
  const double* Bp = Bs + it.boff + (it.n0 + g) * it.kstride + t;
  const int tile_stride = 8 * it.kstride;
#pragma unroll 2
  for (int ks = 0; ks < it.ksteps; ++ks)
  {
    const double a0 = Pa[(ks * 4 + t) * kPStride];
    const double a1 = Pa[(ks * 4 + t) * kPStride + 8];
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
      const double b = Bp[j * tile_stride + ks * 4];
      dmma(acc[0][j][0], acc[0][j][1], a0, b);
      dmma(acc[1][j][0], acc[1][j][1], a1, b);
    }
  }
  // accumulator fragment: row g = point, columns 2t, 2t+1 of each 8-column tile
#pragma unroll
  for (int m = 0; m < 2; ++m)
  {
    const size_t p = pt + 8 * m;
    if (p < npts)
    {
      double* row = basis + (size_t(it.slab) * npts + p) * N;
#pragma unroll
      for (int j = 0; j < NT; ++j)
      {
        const int n = it.n0 + j * 8 + 2 * t;
        if ((N & 1) == 0)
        {
          if (n < N)
            __stcs(reinterpret_cast<double2*>(row + n), make_double2(acc[m][j][0], acc[m][j][1]));
        }
        else
        {
          if (n < N)
            __stcs(row + n, acc[m][j][0]);
          if (n + 1 < N)
            __stcs(row + n + 1, acc[m][j][1]);
        }
      }
    }
  }
}

template <int TDIM, int DEG>
__global__ void __launch_bounds__(kThreads, 1) tabulate_dop_kernel(const Params p)
{
  constexpr int K = jets::psize(TDIM, DEG);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // [ barriers 64 B ][ items ][ rowmap ][ B ][ P tile 0 ][ P tile 1 ]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw);
  Item* items = reinterpret_cast<Item*>(smem_raw + 64);
  int* rowmap = reinterpret_cast<int*>(items + kMaxItems);
  double* Bs = reinterpret_cast<double*>(rowmap + ((K + 3) & ~3));
  double* Pt = Bs + p.b_doubles;
  const int tile_doubles = p.kpad * kPStride;
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 16;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  for (int i = threadIdx.x; i < p.b_doubles; i += kThreads)
    Bs[i] = p.B[i];
  for (int i = threadIdx.x; i < p.nitems; i += kThreads)
    items[i] = p.items[i];
  for (int i = threadIdx.x; i < K; i += kThreads)
    rowmap[i] = p.rowmap[i];
  for (int i = threadIdx.x; i < 2 * tile_doubles; i += kThreads)
    Pt[i] = 0.0; // rows K .. kpad-1 are read by the last k-step of the value slab and must stay finite
  if (threadIdx.x == 0)
  {
    for (int b = 0; b < 2; ++b)
    {
      mbar_init(full0 + 8 * b, kProducers);
      mbar_init(empty0 + 8 * b, kConsumers);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  const size_t ntiles = ceil_div(p.npts, size_t(kTilePts));
  if (warp >= kConsumers)
  {
    // ---- producers: values of the orthonormal set, lane = point ----
    const int pw = warp - kConsumers;
    uint32_t phase = 1; // parity of the previous use of the buffer
    int b = 0;
    size_t i = 0;
    for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++i)
    {
      if (i >= 2)
        mbar_wait(empty0 + 8 * b, phase);
      size_t pt = tile * kTilePts + pw * 32 + lane;
      pt = pt < p.npts ? pt : p.npts - 1; // ragged tail: recompute the last point, stores are predicated
      TileSink<K> sink{Pt + b * tile_doubles + pw * 32 + lane, rowmap};
      if constexpr (TDIM == 1)
        jets::interval<DEG, 0>(p.x[pt], sink);
      else if constexpr (TDIM == 2)
        jets::triangle<DEG, 0>(p.x[2 * pt], p.x[2 * pt + 1], sink);
      else
        jets::tetrahedron<DEG, 0>(p.x[3 * pt], p.x[3 * pt + 1], p.x[3 * pt + 2], sink);
      __syncwarp();
      if (lane == 0)
        mbar_arrive(full0 + 8 * b);
      if (++b == 2)
      {
        b = 0;
        phase ^= 1;
      }
    }
    return;
  }

  // ---- consumers: FP64 tensor cores ----
  const int g = lane >> 2, t = lane & 3;
  const int mp = warp & 3, half = warp >> 2;
  const int first = half == 0 ? 0 : p.split, last = half == 0 ? p.split : p.nitems;
  uint32_t phase = 0;
  int b = 0;
  for (size_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x)
  {
    mbar_wait(full0 + 8 * b, phase);
    const double* Pa = Pt + b * tile_doubles + mp * 16 + g;
    const size_t pt = tile * kTilePts + mp * 16 + g;
    for (int i = first; i < last; ++i)
    {
      const Item it = items[i];
      switch (it.nt)
      {
      case 1: run_item<1>(it, Pa, Bs, p.basis, p.npts, p.N, pt, g, t); break;
      case 2: run_item<2>(it, Pa, Bs, p.basis, p.npts, p.N, pt, g, t); break;
      case 3: run_item<3>(it, Pa, Bs, p.basis, p.npts, p.N, pt, g, t); break;
      case 4: run_item<4>(it, Pa, Bs, p.basis, p.npts, p.N, pt, g, t); break;
      case 5: run_item<5>(it, Pa, Bs, p.basis, p.npts, p.N, pt, g, t); break;
      case 6: run_item<6>(it, Pa, Bs, p.basis, p.npts, p.N, pt, g, t); break;
      case 7: run_item<7>(it, Pa, Bs, p.basis, p.npts, p.N, pt, g, t); break;
      default: run_item<8>(it, Pa, Bs, p.basis, p.npts, p.N, pt, g, t); break;
      }
    }
    __syncwarp();
    if (lane == 0)
      mbar_arrive(empty0 + 8 * b);
    if (++b == 2)
    {
      b = 0;
      phase ^= 1;
    }
  }
}

// ---- host: operand per (element, nd), built once and cached on the element --------------------------------
struct Operand
{
  double* d_B = nullptr;
  Item* d_items = nullptr;
  int* d_rowmap = nullptr;
  int b_doubles = 0, nitems = 0, split = 0, kpad = 0;
  bool usable = false;
  size_t smem = 0;
};
} // namespace dop

struct bx_element_dop
{
  std::mutex mutex;
  std::map<int, dop::Operand> by_nd;
  ~bx_element_dop()
  {
    for (auto& kv : by_nd)
    {
      cudaFree(kv.second.d_B);
      cudaFree(kv.second.d_items);
      cudaFree(kv.second.d_rowmap);
    }
  }
};
bx_element_dop* dop_cache_create() { return new bx_element_dop; }
void dop_cache_destroy(bx_element_dop* c) { delete c; }

namespace dop
{
namespace
{
using Vec = std::vector<double>;

int simplex_dim(int tdim, int n) { return n < 0 ? 0 : jets::psize(tdim, n); }

template <typename T>
T* upload(const std::vector<T>& v)
{
  T* d = nullptr;
  BX_CUDA(cudaMalloc(&d, std::max<size_t>(v.size(), 1) * sizeof(T)));
  if (!v.empty())
    BX_CUDA(cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
  return d;
}

Operand build_operand(const bx_element& e, int nd, bool upload_to_device)
{
  Operand op;
  const int tdim = e.tdim, n = e.superdegree, K = e.K, N = e.N;
  const int cell = e.cell_type;
  // first-derivative operators in the orthonormal basis: D_a[k][j] = int d_a P_k P_j, j < dim P_{n-1}
  std::vector<Vec> D(tdim, Vec(size_t(K) * K, 0.0));
  if (nd > 0 and n > 0)
  {
    Vec qx, qw;
    host::quadrature(cell, 2 * n, qx, qw);
    const size_t nq = qw.size();
    const Vec P = host::polyset_tabulate(cell, n, 1, qx, nq); // [1 + tdim][K][nq]
    const int Km1 = simplex_dim(tdim, n - 1);
    for (int a = 0; a < tdim; ++a)
    {
      // slot of the first derivative along axis a: idx(1,0[,0]), idx(0,1[,0]), idx(0,0,1)
      const int slot = tdim == 1 ? 1 : tdim == 2 ? (a == 0 ? idx2(1, 0) : idx2(0, 1))
                                                 : (a == 0 ? idx3(1, 0, 0) : a == 1 ? idx3(0, 1, 0) : idx3(0, 0, 1));
      const double* dP = P.data() + size_t(slot) * K * nq;
      for (int k = 0; k < K; ++k)
        for (int j = 0; j < Km1; ++j)
        {
          double s = 0.0;
          for (size_t q = 0; q < nq; ++q)
            s += qw[q] * dP[size_t(k) * nq + q] * P[size_t(j) * nq + q];
          D[a][size_t(k) * K + j] = s;
        }
    }
  }
  // natural-index norms and the production-order row map of the jets
  std::vector<int> slot(K);
  Vec norm_prod(K), norm(K);
  jets::production_order(tdim, n, slot.data(), norm_prod.data());
  for (int s = 0; s < K; ++s)
    norm[slot[s]] = norm_prod[s];

  const int nder = polyset_nderivs(cell, nd);
  const int npad = int(ceil_div(N, 8) * 8);
  op.kpad = int(ceil_div(K, 4) * 4);
  Vec B;
  std::vector<Item> items;
  auto matmul = [K](const Vec& A, const Vec& Bm)
  {
    Vec Cm(size_t(K) * K, 0.0);
    for (int i = 0; i < K; ++i)
      for (int k = 0; k < K; ++k)
      {
        const double a = A[size_t(i) * K + k];
        if (a == 0.0)
          continue;
        for (int j = 0; j < K; ++j)
          Cm[size_t(i) * K + j] += a * Bm[size_t(k) * K + j];
      }
    return Cm;
  };
  for (int kx = 0; kx <= nd; ++kx)
    for (int ky = 0; ky <= (tdim >= 2 ? nd - kx : 0); ++ky)
      for (int kz = 0; kz <= (tdim >= 3 ? nd - kx - ky : 0); ++kz)
      {
        const int m = kx + ky + kz;
        const int d = tdim == 1 ? kx : tdim == 2 ? idx2(kx, ky) : idx3(kx, ky, kz);
        if (d >= nder)
          continue;
        Vec M(size_t(K) * K, 0.0);
        for (int i = 0; i < K; ++i)
          M[size_t(i) * K + i] = 1.0;
        const int cnt[3] = {kx, ky, kz};
        for (int a = 0; a < tdim; ++a)
          for (int r = 0; r < cnt[a]; ++r)
            M = matmul(M, D[a]);
        const int Kd = simplex_dim(tdim, n - m);
        const int ksteps = int(ceil_div(Kd, 4));
        const int kp = ksteps * 4;
        const int kstride = kp + ((4 - kp % 16) + 16) % 16;
        const int boff = int(B.size());
        B.resize(B.size() + size_t(npad) * kstride, 0.0);
        for (int row = 0; row < N; ++row)
          for (int j = 0; j < Kd; ++j)
          {
            double s = 0.0;
            for (int k = 0; k < K; ++k)
              s += e.Cp[size_t(row) * e.Kpad + k] * M[size_t(k) * K + j];
            B[size_t(boff) + size_t(row) * kstride + j] = s * norm[j];
          }
        for (int n0 = 0; n0 < npad; n0 += 8 * kMaxNT)
        {
          Item it{};
          it.boff = boff;
          it.ksteps = ksteps;
          it.kstride = kstride;
          it.n0 = n0;
          it.nt = std::min(kMaxNT, (npad - n0) / 8);
          it.slab = d;
          items.push_back(it);
        }
      }
  if (int(items.size()) > kMaxItems)
    return op;
  // balance the items over the two consumer halves (largest first)
  std::sort(items.begin(), items.end(),
            [](const Item& a, const Item& b) { return a.ksteps * a.nt > b.ksteps * b.nt; });
  std::vector<Item> h0, h1;
  long c0 = 0, c1 = 0;
  for (const Item& it : items)
  {
    const long cost = long(it.ksteps) * it.nt + 1;
    if (c0 <= c1)
    {
      h0.push_back(it);
      c0 += cost;
    }
    else
    {
      h1.push_back(it);
      c1 += cost;
    }
  }
  op.split = int(h0.size());
  h0.insert(h0.end(), h1.begin(), h1.end());
  op.nitems = int(h0.size());
  op.b_doubles = int((B.size() + 1) & ~size_t(1));
  B.resize(op.b_doubles, 0.0);
  op.smem = 64 + sizeof(Item) * kMaxItems + size_t((K + 3) & ~3) * sizeof(int) + size_t(op.b_doubles) * 8
            + 2 * size_t(op.kpad) * kPStride * 8;
  if (op.smem > 225 * 1024)
    return op;
  if (upload_to_device)
  {
    op.d_B = upload(B);
    op.d_items = upload(h0);
    op.d_rowmap = upload(slot);
  }
  op.usable = true;
  return op;
}

template <int TDIM, int DEG>
void launch(const Operand& op, const bx_element& e, const double* x, size_t npts, double* basis, cudaStream_t s)
{
  auto kern = tabulate_dop_kernel<TDIM, DEG>;
  BX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(op.smem)));
  Params p;
  p.x = x;
  p.npts = npts;
  p.B = op.d_B;
  p.items = op.d_items;
  p.rowmap = op.d_rowmap;
  p.basis = basis;
  p.N = e.N;
  p.K = e.K;
  p.kpad = op.kpad;
  p.b_doubles = op.b_doubles;
  p.nitems = op.nitems;
  p.split = op.split;
  const size_t ntiles = ceil_div(npts, size_t(kTilePts));
  kern<<<unsigned(std::min<size_t>(ntiles, size_t(sm_count()))), kThreads, op.smem, s>>>(p);
  BX_LAUNCHED();
}

#define BX_DOP(T, DEG)                                                                                     \
  if (e.tdim == T and e.superdegree == DEG)                                                                \
  {                                                                                                        \
    if (!dry_run)                                                                                          \
      launch<T, DEG>(*op, e, x, npts, basis, s);                                                           \
    return true;                                                                                           \
  }
} // namespace

// Returns false when the element / derivative order is outside this path (caller falls through).
bool launch_any(const bx_element& e, int nd, const double* x, size_t npts, double* basis, cudaStream_t s, bool dry_run)
{
  if (e.polyset_type != BX_POLYSET_STANDARD or nd < 0 or nd > 4 or !e.dop_cache)
    return false;
  if (e.cell_type != BX_CELL_INTERVAL and e.cell_type != BX_CELL_TRIANGLE and e.cell_type != BX_CELL_TETRAHEDRON)
    return false;
  const int max_deg = e.tdim == 1 ? 10 : e.tdim == 2 ? 8 : 6;
  if (e.superdegree < 1 or e.superdegree > max_deg)
    return false;
  const Operand* op = nullptr;
  {
    std::lock_guard<std::mutex> lock(e.dop_cache->mutex);
    auto it = e.dop_cache->by_nd.find(nd);
    if (it == e.dop_cache->by_nd.end())
      it = e.dop_cache->by_nd.emplace(nd, build_operand(e, nd, e.on_device)).first;
    op = &it->second;
  }
  if (!op->usable or (!dry_run and !op->d_B))
    return false;
  BX_DOP(1, 1) BX_DOP(1, 2) BX_DOP(1, 3) BX_DOP(1, 4) BX_DOP(1, 5) BX_DOP(1, 6) BX_DOP(1, 7) BX_DOP(1, 8)
  BX_DOP(1, 9) BX_DOP(1, 10)
  BX_DOP(2, 1) BX_DOP(2, 2) BX_DOP(2, 3) BX_DOP(2, 4) BX_DOP(2, 5) BX_DOP(2, 6) BX_DOP(2, 7) BX_DOP(2, 8)
  BX_DOP(3, 1) BX_DOP(3, 2) BX_DOP(3, 3) BX_DOP(3, 4) BX_DOP(3, 5) BX_DOP(3, 6)
  return false;
}
} // namespace dop
} // namespace bx

// extern "C" boundary (include/basix_b200.h): argument validation, error plumbing and the
// host-pointer staging pipelines.  No arithmetic lives here; every compute call ends in a CUDA
// kernel of this library (there is no CPU fallback).
#include "element.cuh"

#include <cstdlib>
#include <algorithm>
#include <memory>
#include <mutex>
#include <vector>

namespace bx
{
// ---- globals -------------------------------------------------------------------------------------
std::atomic<uint64_t> g_launch_count{0};
namespace
{
thread_local std::string t_error;
}
void set_last_error(const std::string& msg) { t_error = msg; }

int sm_count()
{
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  BX_CUDA(cudaGetDevice(&dev));
  if (dev != cached_dev)
  {
    BX_CUDA(cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev));
    cached_dev = dev;
  }
  return cached;
}

// ---- cell topology: counts per dimension and face types (reference cell.cpp topology) ---------------
int cell_tdim(int c)
{
  switch (c)
  {
  case BX_CELL_POINT: return 0;
  case BX_CELL_INTERVAL: return 1;
  case BX_CELL_TRIANGLE:
  case BX_CELL_QUADRILATERAL: return 2;
  case BX_CELL_TETRAHEDRON:
  case BX_CELL_HEXAHEDRON:
  case BX_CELL_PRISM:
  case BX_CELL_PYRAMID: return 3;
  default: fail(BX_ERR_INVALID_ARG, "Unsupported cell type");
  }
}

int cell_num_entities(int c, int dim)
{
  static const int counts[8][4] = {
      {1, 0, 0, 0},  // point
      {2, 1, 0, 0},  // interval
      {3, 3, 1, 0},  // triangle
      {4, 6, 4, 1},  // tetrahedron
      {4, 4, 1, 0},  // quadrilateral
      {8, 12, 6, 1}, // hexahedron
      {6, 9, 5, 1},  // prism
      {5, 8, 5, 1},  // pyramid
  };
  if (c < 0 or c > 7 or dim < 0 or dim > 3)
    fail(BX_ERR_INVALID_ARG, "Unsupported cell type");
  return counts[c][dim];
}

int cell_sub_entity_type(int c, int dim, int index)
{
  const int tdim = cell_tdim(c);
  if (dim == 0)
    return BX_CELL_POINT;
  if (dim == 1)
    return BX_CELL_INTERVAL;
  if (dim == tdim)
    return c;
  // dim == 2 faces of a 3-D cell
  switch (c)
  {
  case BX_CELL_TETRAHEDRON: return BX_CELL_TRIANGLE;
  case BX_CELL_HEXAHEDRON: return BX_CELL_QUADRILATERAL;
  case BX_CELL_PRISM: return (index == 0 or index == 4) ? BX_CELL_TRIANGLE : BX_CELL_QUADRILATERAL;
  case BX_CELL_PYRAMID: return index == 0 ? BX_CELL_QUADRILATERAL : BX_CELL_TRIANGLE;
  default: fail(BX_ERR_INVALID_ARG, "Unsupported cell type");
  }
}

// reference polyset.cpp:3003-3050
int polyset_dim(int c, int ptype, int d)
{
  if (ptype == BX_POLYSET_STANDARD)
  {
    switch (c)
    {
    case BX_CELL_POINT: return 1;
    case BX_CELL_INTERVAL: return d + 1;
    case BX_CELL_TRIANGLE: return (d + 1) * (d + 2) / 2;
    case BX_CELL_TETRAHEDRON: return (d + 1) * (d + 2) * (d + 3) / 6;
    case BX_CELL_QUADRILATERAL: return (d + 1) * (d + 1);
    case BX_CELL_HEXAHEDRON: return (d + 1) * (d + 1) * (d + 1);
    case BX_CELL_PRISM: return (d + 1) * (d + 1) * (d + 2) / 2;
    case BX_CELL_PYRAMID: return (d + 1) * (d + 1) * (d + 1);
    default: return 1;
    }
  }
  if (ptype == BX_POLYSET_MACROEDGE)
  {
    switch (c)
    {
    case BX_CELL_POINT: return 1;
    case BX_CELL_INTERVAL: return 2 * d + 1;
    case BX_CELL_TRIANGLE: return (d + 1) * (2 * d + 1);
    case BX_CELL_TETRAHEDRON: return (d + 1) * (2 * d + 1) * (2 * d + 3) / 3;
    case BX_CELL_QUADRILATERAL: return (2 * d + 1) * (2 * d + 1);
    case BX_CELL_HEXAHEDRON: return (2 * d + 1) * (2 * d + 1) * (2 * d + 1);
    default: return 1;
    }
  }
  return 1;
}

// reference polyset.cpp:3052-3075
int polyset_nderivs(int c, int n)
{
  switch (c)
  {
  case BX_CELL_POINT: return 1;
  case BX_CELL_INTERVAL: return n + 1;
  case BX_CELL_TRIANGLE:
  case BX_CELL_QUADRILATERAL: return (n + 1) * (n + 2) / 2;
  case BX_CELL_TETRAHEDRON:
  case BX_CELL_HEXAHEDRON:
  case BX_CELL_PRISM:
  case BX_CELL_PYRAMID: return (n + 1) * (n + 2) * (n + 3) / 6;
  default: return 1;
  }
}

// ---- forward declarations of the device-pointer implementations ---------------------------------------
template <typename T>
void polyset_tabulate_device(int cell, int ptype, int degree, int nd, const T* x, size_t npts, T* P,
                             cudaStream_t stream);
template <typename T>
void tabulate_device(const bx_element& e, int nd, const T* x, size_t npts, T* basis, size_t out_pts, cudaStream_t s);
int tabulate_path(const bx_element& e, int nd);
template <typename T>
void interpolate_device(const bx_element& e, int layout, const T* values, size_t nc, T* coefficients, cudaStream_t s);
template <typename T>
void interpolate_pulled_device(const bx_element& e, const T* u, const T* J, const T* detJ, const T* K, size_t nc, int gdim,
                               T* coefficients, cudaStream_t s);
void closure_permute_device(const bx_element& e, int inverse, int32_t* d, size_t nc, size_t m, const uint32_t* info,
                            int entity_type, int entity_index, cudaStream_t s);
unsigned tabulate_path_mask(unsigned mask);
int map_value_size(int map_type, int pull, int in_vs, int gdim, int tdim);
template <typename T>
void map_device(int map_type, int pull, const T* U, const T* J, const T* detJ, const T* K, size_t nc, size_t rows,
                int in_vs, int gdim, int tdim, T* out, cudaStream_t s, int bcast = 0);
template <typename T>
void dof_transform_device(const bx_element& e, int op, T* data, size_t nc, size_t cell_size, int n,
                          const uint32_t* cell_info, cudaStream_t s);
template <typename T>
void geometry_device(const T* coords, const T* dphi, size_t nc, int nv, int gdim, int tdim, int npts, T* J, T* detJ, T* K,
                     cudaStream_t s);
template <typename T>
void physical_basis_device(const bx_element& e, int op, const T* U, size_t npts, const T* coords, const T* J,
                           const T* detJ, const T* K, const uint32_t* cell_info, size_t nc, int gdim, T* out,
                           cudaStream_t s, int per_point = 0);
void interpolation_operator_shape(const bx_element& from, const bx_element& to, int32_t shape[2]);
void interpolation_operator_device(const bx_element& from, const bx_element& to, double* out, cudaStream_t s);
void gauss_jacobi_rule_unit(double a, int m, std::vector<double>& x, std::vector<double>& w);
void make_quadrature_gauss_jacobi(int cell, int q, std::vector<double>& x, std::vector<double>& w);
std::vector<double> create_lattice_points(int cell, int n, int lattice_type, int exterior, int simplex_method);
bx_element* create_lagrange(int cell, int degree, int variant, bool discontinuous, const int32_t* dof_ordering,
                            int n_dof_ordering);
bx_element* create_rt(int cell, int degree, int lvariant, bool discontinuous);
bx_element* create_nedelec(int cell, int degree, int lvariant, bool discontinuous);

namespace
{
// ---- host-pointer staging -----------------------------------------------------------------------------
// Two streams, each with its own device buffers; chunk i runs H2D -> kernels -> D2H in order on
// stream i % 2, so the copies of one chunk overlap the kernels of the other.
// Per-thread staging resources, created on first use and kept: two streams and a set of grow-only device
// buffers.  (Creating streams and cudaMalloc-ing per call costs more than a small call itself; the reference API
// is one cell / a handful of points per call.)  The pool follows the calling thread's current device.
struct StagePool
{
  static constexpr int kSlots = 5;
  int device = -1;
  cudaStream_t s[2] = {nullptr, nullptr};
  void* buf[kSlots][2] = {};
  size_t cap[kSlots][2] = {};

  void release()
  {
    for (auto& st : s)
      if (st)
      {
        cudaStreamSynchronize(st);
        cudaStreamDestroy(st);
        st = nullptr;
      }
    for (int i = 0; i < kSlots; ++i)
      for (int b = 0; b < 2; ++b)
      {
        if (buf[i][b])
          cudaFree(buf[i][b]);
        buf[i][b] = nullptr;
        cap[i][b] = 0;
      }
    device = -1;
  }
  void attach()
  {
    int dev = 0;
    BX_CUDA(cudaGetDevice(&dev));
    if (dev == device)
      return;
    release();
    BX_CUDA(cudaStreamCreateWithFlags(&s[0], cudaStreamNonBlocking));
    BX_CUDA(cudaStreamCreateWithFlags(&s[1], cudaStreamNonBlocking));
    device = dev;
  }
  // buffer `slot` of pipeline half b, at least `bytes` large; buffers above 1 GiB are not kept between calls
  template <typename T>
  T* get(int slot, int b, size_t bytes)
  {
    if (cap[slot][b] < bytes)
    {
      if (buf[slot][b])
      {
        BX_CUDA(cudaStreamSynchronize(s[b]));
        BX_CUDA(cudaFree(buf[slot][b]));
        buf[slot][b] = nullptr;
        cap[slot][b] = 0;
      }
      if (bytes)
      {
        BX_CUDA(cudaMalloc(&buf[slot][b], bytes));
        cap[slot][b] = bytes;
      }
    }
    return static_cast<T*>(buf[slot][b]);
  }
  void sync()
  {
    BX_CUDA(cudaStreamSynchronize(s[0]));
    BX_CUDA(cudaStreamSynchronize(s[1]));
  }
  // give large buffers back so an occasional huge call does not pin device memory for the thread's lifetime
  void trim()
  {
    for (int i = 0; i < kSlots; ++i)
      for (int b = 0; b < 2; ++b)
        if (cap[i][b] > (size_t(1) << 30))
        {
          cudaFree(buf[i][b]);
          buf[i][b] = nullptr;
          cap[i][b] = 0;
        }
  }
  ~StagePool() { release(); }
};

// Drains both staging streams when a host-pointer call leaves its scope, also on an exception: the caller may
// free its host buffers as soon as the call returns.
struct DrainOnExit
{
  StagePool& p;
  ~DrainOnExit()
  {
    for (auto st : p.s)
      if (st)
        cudaStreamSynchronize(st);
  }
};

StagePool& stage_pool()
{
  static thread_local StagePool pool;
  pool.attach();
  return pool;
}

// Target size of one staged chunk (all arrays of the chunk that cross the bus).  BX_STAGE_MB overrides (tuning).
size_t stage_bytes()
{
  static const size_t v = []
  {
    const char* e = std::getenv("BX_STAGE_MB");
    const long mb = e ? std::atol(e) : 0;
    return size_t(mb > 0 ? mb : 256) << 20;
  }();
  return v;
}

size_t chunk_units(size_t nunits, size_t bytes_per_unit, size_t align)
{
  size_t c = std::max<size_t>(1, stage_bytes() / std::max<size_t>(1, bytes_per_unit));
  c = std::max(align, c / align * align);
  return std::min(c, nunits);
}

void require(bool ok, const char* msg)
{
  if (!ok)
    fail(BX_ERR_INVALID_ARG, msg);
}

void check_mem(int mem) { require(mem == BX_MEM_HOST or mem == BX_MEM_DEVICE, "mem must be BX_MEM_HOST or BX_MEM_DEVICE"); }

template <typename T>
void polyset_any(int cell, int ptype, int degree, int nd, const T* x, size_t npts, T* P, int mem, bx_stream_t stream)
{
  check_mem(mem);
  if (npts == 0)
    return;
  require(x and P, "polyset tabulate: null pointer");
  const int tdim = cell_tdim(cell);
  if (mem == BX_MEM_DEVICE)
  {
    polyset_tabulate_device<T>(cell, ptype, degree, nd, x, npts, P, static_cast<cudaStream_t>(stream));
    return;
  }
  const size_t rows = size_t(polyset_nderivs(cell, nd)) * polyset_dim(cell, ptype, degree);
  const size_t chunk = chunk_units(npts, rows * sizeof(T), 32);
  StagePool& st = stage_pool();
  DrainOnExit drain{st};
  T *dx[2], *dP[2];
  for (int b = 0; b < 2; ++b)
  {
    dx[b] = st.get<T>(0, b, chunk * std::max(tdim, 1) * sizeof(T));
    dP[b] = st.get<T>(1, b, chunk * rows * sizeof(T));
  }
  size_t i = 0;
  for (size_t p0 = 0; p0 < npts; p0 += chunk, ++i)
  {
    const int b = int(i & 1);
    const size_t np = std::min(chunk, npts - p0);
    if (tdim > 0)
      BX_CUDA(cudaMemcpyAsync(dx[b], x + p0 * tdim, np * tdim * sizeof(T), cudaMemcpyHostToDevice, st.s[b]));
    polyset_tabulate_device<T>(cell, ptype, degree, nd, dx[b], np, dP[b], st.s[b]);
    // table rows are npts apart on the host, np apart on the device.  A 2-D copy is limited to pitches below
    // cudaDeviceProp::memPitch (2 GiB), which huge host tables exceed: those go row by row (the rows are few).
    if (npts * sizeof(T) < (size_t(1) << 31))
      BX_CUDA(cudaMemcpy2DAsync(P + p0, npts * sizeof(T), dP[b], np * sizeof(T), np * sizeof(T), rows,
                                cudaMemcpyDeviceToHost, st.s[b]));
    else
      for (size_t r = 0; r < rows; ++r)
        BX_CUDA(cudaMemcpyAsync(P + r * npts + p0, dP[b] + r * np, np * sizeof(T), cudaMemcpyDeviceToHost, st.s[b]));
  }
  st.sync();
  st.trim();
}

template <typename T>
void tabulate_any(const bx_element& e, int nd, const T* x, size_t npts, int tdim_x, T* basis, int mem, bx_stream_t stream)
{
  check_mem(mem);
  if (tdim_x != e.tdim)
    fail(BX_ERR_INVALID_ARG, "Point dim (" + std::to_string(tdim_x) + ") does not match element dim ("
                                 + std::to_string(e.tdim) + ").");
  require(nd >= 0, "tabulate: negative derivative order");
  if (npts == 0)
    return;
  require(x and basis, "tabulate: null pointer");
  if (mem == BX_MEM_DEVICE)
  {
    tabulate_device<T>(e, nd, x, npts, basis, npts, static_cast<cudaStream_t>(stream));
    return;
  }
  const int nder = polyset_nderivs(e.cell_type, nd);
  const size_t per_pt = size_t(nder) * e.N * sizeof(T);
  const size_t chunk = chunk_units(npts, per_pt, 32);
  StagePool& st = stage_pool();
  DrainOnExit drain{st};
  T *dx[2], *dout[2];
  for (int b = 0; b < 2; ++b)
  {
    dx[b] = st.get<T>(0, b, chunk * e.tdim * sizeof(T));
    dout[b] = st.get<T>(1, b, chunk * per_pt);
  }
  size_t i = 0;
  for (size_t p0 = 0; p0 < npts; p0 += chunk, ++i)
  {
    const int b = int(i & 1);
    const size_t np = std::min(chunk, npts - p0);
    BX_CUDA(cudaMemcpyAsync(dx[b], x + p0 * e.tdim, np * e.tdim * sizeof(T), cudaMemcpyHostToDevice, st.s[b]));
    tabulate_device<T>(e, nd, dx[b], np, dout[b], np, st.s[b]);
    // each derivative slab of the chunk is one contiguous run of np * N values on both sides
    for (int d = 0; d < nder; ++d)
      BX_CUDA(cudaMemcpyAsync(basis + (size_t(d) * npts + p0) * e.N, dout[b] + size_t(d) * np * e.N,
                              np * e.N * sizeof(T), cudaMemcpyDeviceToHost, st.s[b]));
  }
  st.sync();
  st.trim();
}

template <typename T>
void map_any(int map_type, int pull, const T* U, const T* J, const T* detJ, const T* K, size_t nc, size_t rows, int in_vs,
             int gdim, int tdim, T* out, int mem, bx_stream_t stream, int bcast = 0)
{
  check_mem(mem);
  require(in_vs > 0 and gdim >= 1 and gdim <= 3 and tdim >= 1 and tdim <= 3, "map: bad value size or Jacobian shape");
  const int out_vs = map_value_size(map_type, pull, in_vs, gdim, tdim);
  if (nc == 0 or rows == 0)
    return;
  require(U and J and detJ and K and out, "map: null pointer");
  if (mem == BX_MEM_DEVICE)
  {
    map_device<T>(map_type, pull, U, J, detJ, K, nc, rows, in_vs, gdim, tdim, out, static_cast<cudaStream_t>(stream),
                  bcast);
    return;
  }
  const size_t jsz = size_t(gdim) * tdim;
  // Only the arrays the map reads cross the PCIe bus: the covariant maps use the reference kernel's "K" argument
  // alone, the contravariant maps its "J" argument and detJ (maps.h:73-192); pull_back swaps J and K
  // (finite-element.cpp:2038).  For the covariant push of BASELINE config 4 that is 96 instead of 176 bytes per cell.
  const bool uses_Karg = map_type == BX_MAP_COVARIANT_PIOLA or map_type == BX_MAP_DOUBLE_COVARIANT_PIOLA;
  const bool uses_Jarg = map_type == BX_MAP_CONTRAVARIANT_PIOLA or map_type == BX_MAP_DOUBLE_CONTRAVARIANT_PIOLA;
  const bool need_det = uses_Jarg;
  const bool need_J = pull ? uses_Karg : uses_Jarg;
  const bool need_K = pull ? uses_Jarg : uses_Karg;
  const size_t per_cell = (rows * ((bcast ? 0 : in_vs) + out_vs) + (need_J ? jsz : 0) + (need_K ? jsz : 0) + (need_det ? 1 : 0))
                          * sizeof(T);
  const size_t chunk = chunk_units(nc, per_cell, 1);
  StagePool& st = stage_pool();
  DrainOnExit drain{st};
  T *dU[2], *dJ[2], *dd[2], *dK[2], *dout[2];
  for (int b = 0; b < 2; ++b)
  {
    dU[b] = st.get<T>(0, b, (bcast ? 1 : chunk) * rows * in_vs * sizeof(T));
    dJ[b] = st.get<T>(1, b, chunk * jsz * sizeof(T));
    dd[b] = st.get<T>(2, b, chunk * sizeof(T));
    dK[b] = st.get<T>(3, b, chunk * jsz * sizeof(T));
    dout[b] = st.get<T>(4, b, chunk * rows * out_vs * sizeof(T));
  }
  size_t i = 0;
  for (size_t c0 = 0; c0 < nc; c0 += chunk, ++i)
  {
    const int b = int(i & 1);
    const size_t n = std::min(chunk, nc - c0);
    cudaStream_t s = st.s[b];
    if (!bcast)
      BX_CUDA(cudaMemcpyAsync(dU[b], U + c0 * rows * in_vs, n * rows * in_vs * sizeof(T), cudaMemcpyHostToDevice, s));
    else if (i < 2) // the shared slab of reference values goes up once per pipeline half
      BX_CUDA(cudaMemcpyAsync(dU[b], U, rows * in_vs * sizeof(T), cudaMemcpyHostToDevice, s));
    if (need_J)
      BX_CUDA(cudaMemcpyAsync(dJ[b], J + c0 * jsz, n * jsz * sizeof(T), cudaMemcpyHostToDevice, s));
    if (need_det)
      BX_CUDA(cudaMemcpyAsync(dd[b], detJ + c0, n * sizeof(T), cudaMemcpyHostToDevice, s));
    if (need_K)
      BX_CUDA(cudaMemcpyAsync(dK[b], K + c0 * jsz, n * jsz * sizeof(T), cudaMemcpyHostToDevice, s));
    map_device<T>(map_type, pull, dU[b], dJ[b], dd[b], dK[b], n, rows, in_vs, gdim, tdim, dout[b], s, bcast);
    BX_CUDA(cudaMemcpyAsync(out + c0 * rows * out_vs, dout[b], n * rows * out_vs * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
  st.sync();
  st.trim();
}

template <typename T>
void transform_any(const bx_element* e, int op, T* data, size_t nc, size_t cell_size, int n, const uint32_t* cell_info,
                   int mem, bx_stream_t stream)
{
  check_mem(mem);
  require(e != nullptr, "T_apply: null element");
  if (nc == 0)
    return;
  require(data and cell_info, "T_apply: null pointer");
  if (mem == BX_MEM_DEVICE)
  {
    dof_transform_device<T>(*e, op, data, nc, cell_size, n, cell_info, static_cast<cudaStream_t>(stream));
    return;
  }
  // validate once (throws on bad arguments) and return early for identity elements, before staging
  dof_transform_device<T>(*e, op, static_cast<T*>(nullptr), 0, cell_size, n, nullptr, nullptr);
  if (e->identity or e->tdim < 2 or e->slots.empty())
    return;
  const size_t chunk = chunk_units(nc, cell_size * sizeof(T), 1);
  // movable range of a cell (the kernels never touch anything else): left layout u[dof][b], right layout u[b][dof]
  const bool right = op >= 4;
  size_t r0 = right ? size_t(e->lo_dof) : size_t(e->lo_dof) * n;
  size_t len = right ? size_t(n - 1) * e->ndofs + e->hi_dof - e->lo_dof : size_t(e->hi_dof - e->lo_dof) * n;
  // Strided (2-D) PCIe copies pay off only for wide rows that skip a good part of the cell: measured on B200,
  // 320-byte rows at a 560-byte pitch (RT4) move 1.5x more cells per second than whole cells, 192-byte rows at a
  // 224-byte pitch (P5 permute) 15x fewer.  Otherwise ship whole cells as one contiguous copy.
  if (len * sizeof(T) < 256 or 3 * len > 2 * cell_size)
  {
    r0 = 0;
    len = cell_size;
  }
  StagePool& st = stage_pool();
  DrainOnExit drain{st};
  T* dd[2];
  uint32_t* dc[2];
  for (int b = 0; b < 2; ++b)
  {
    dd[b] = st.get<T>(0, b, chunk * cell_size * sizeof(T));
    dc[b] = st.get<uint32_t>(1, b, chunk * sizeof(uint32_t));
  }
  size_t i = 0;
  for (size_t c0 = 0; c0 < nc; c0 += chunk, ++i)
  {
    const int b = int(i & 1);
    const size_t m = std::min(chunk, nc - c0);
    cudaStream_t s = st.s[b];
    // only the entries that can move travel: columns [r0, r0 + len) of every cell, same pitch on both sides
    if (len == cell_size)
      BX_CUDA(cudaMemcpyAsync(dd[b], data + c0 * cell_size, m * cell_size * sizeof(T), cudaMemcpyHostToDevice, s));
    else
      BX_CUDA(cudaMemcpy2DAsync(dd[b] + r0, cell_size * sizeof(T), data + c0 * cell_size + r0,
                                cell_size * sizeof(T), len * sizeof(T), m, cudaMemcpyHostToDevice, s));
    BX_CUDA(cudaMemcpyAsync(dc[b], cell_info + c0, m * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    dof_transform_device<T>(*e, op, dd[b], m, cell_size, n, dc[b], s);
    if (len == cell_size)
      BX_CUDA(cudaMemcpyAsync(data + c0 * cell_size, dd[b], m * cell_size * sizeof(T), cudaMemcpyDeviceToHost, s));
    else
      BX_CUDA(cudaMemcpy2DAsync(data + c0 * cell_size + r0, cell_size * sizeof(T), dd[b] + r0,
                                cell_size * sizeof(T), len * sizeof(T), m, cudaMemcpyDeviceToHost, s));
  }
  st.sync();
  st.trim();
}
template <typename T>
void interpolate_any(const bx_element* e, int layout, const T* values, size_t nc, T* out, int mem, bx_stream_t stream)
{
  check_mem(mem);
  require(e != nullptr, "interpolate: null element");
  require(nc == 0 or (values and out), "interpolate: null pointer");
  if (mem == BX_MEM_DEVICE)
  {
    interpolate_device<T>(*e, layout, values, nc, out, static_cast<cudaStream_t>(stream));
    return;
  }
  interpolate_device<T>(*e, layout, nullptr, 0, nullptr, nullptr); // validate before staging
  if (nc == 0)
    return;
  const size_t K = size_t(e->vs) * e->interp_npts, N = size_t(e->ndofs);
  const size_t chunk = chunk_units(nc, (K + N) * sizeof(T), 1);
  StagePool& st = stage_pool();
  DrainOnExit drain{st};
  T *dv[2], *dout[2];
  for (int b = 0; b < 2; ++b)
  {
    dv[b] = st.get<T>(0, b, chunk * K * sizeof(T));
    dout[b] = st.get<T>(1, b, chunk * N * sizeof(T));
  }
  size_t i = 0;
  for (size_t c0 = 0; c0 < nc; c0 += chunk, ++i)
  {
    const int b = int(i & 1);
    const size_t k = std::min(chunk, nc - c0);
    cudaStream_t s = st.s[b];
    BX_CUDA(cudaMemcpyAsync(dv[b], values + c0 * K, k * K * sizeof(T), cudaMemcpyHostToDevice, s));
    interpolate_device<T>(*e, layout, dv[b], k, dout[b], s);
    BX_CUDA(cudaMemcpyAsync(out + c0 * N, dout[b], k * N * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
  st.sync();
  st.trim();
}

// pull_back fused into the interpolation; host pointers are staged in chunks of cells like every other call
template <typename T>
void interpolate_pulled_any(const bx_element* e, const T* u, const T* J, const T* detJ, const T* K, size_t nc, int gdim, T* out,
                            int mem, bx_stream_t stream)
{
  check_mem(mem);
  require(e != nullptr, "interpolate_pulled: null element");
  if (mem == BX_MEM_DEVICE)
  {
    interpolate_pulled_device<T>(*e, u, J, detJ, K, nc, gdim, out, static_cast<cudaStream_t>(stream));
    return;
  }
  {
    // validate the element / shapes before staging (no cells: nothing is dereferenced)
    const T one = T(1);
    interpolate_pulled_device<T>(*e, &one, J ? &one : nullptr, detJ ? &one : nullptr, K ? &one : nullptr, 0, gdim, nullptr, nullptr);
  }
  if (nc == 0)
    return;
  require(u and out, "interpolate_pulled: null pointer");
  const bool cov = e->map_type == BX_MAP_COVARIANT_PIOLA;
  const size_t D = size_t(e->tdim), U = size_t(e->interp_npts) * D, N = size_t(e->ndofs), M = D * D;
  const size_t chunk = chunk_units(nc, (U + N + M + 1) * sizeof(T), 1);
  StagePool& st = stage_pool();
  DrainOnExit drain{st};
  T *du[2], *dA[2], *dd[2], *dout[2];
  for (int b = 0; b < 2; ++b)
  {
    du[b] = st.get<T>(0, b, chunk * U * sizeof(T));
    dA[b] = st.get<T>(1, b, chunk * M * sizeof(T));
    dd[b] = st.get<T>(2, b, chunk * sizeof(T));
    dout[b] = st.get<T>(3, b, chunk * N * sizeof(T));
  }
  size_t i = 0;
  for (size_t c0 = 0; c0 < nc; c0 += chunk, ++i)
  {
    const int b = int(i & 1);
    const size_t k = std::min(chunk, nc - c0);
    cudaStream_t s = st.s[b];
    BX_CUDA(cudaMemcpyAsync(du[b], u + c0 * U, k * U * sizeof(T), cudaMemcpyHostToDevice, s));
    BX_CUDA(cudaMemcpyAsync(dA[b], (cov ? J : K) + c0 * M, k * M * sizeof(T), cudaMemcpyHostToDevice, s));
    if (!cov)
      BX_CUDA(cudaMemcpyAsync(dd[b], detJ + c0, k * sizeof(T), cudaMemcpyHostToDevice, s));
    interpolate_pulled_device<T>(*e, du[b], cov ? dA[b] : nullptr, cov ? nullptr : dd[b], cov ? nullptr : dA[b], k, gdim, dout[b], s);
    BX_CUDA(cudaMemcpyAsync(out + c0 * N, dout[b], k * N * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
  st.sync();
  st.trim();
}

// J / detJ / K of nc cells at npts points each; any output may be null.
template <typename T>
void geometry_any(const T* coords, const T* dphi, size_t nc, int nv, int gdim, int tdim, int npts, T* J, T* detJ, T* K,
                  int mem, bx_stream_t stream)
{
  check_mem(mem);
  if (mem == BX_MEM_DEVICE)
  {
    geometry_device<T>(coords, dphi, nc, nv, gdim, tdim, npts, J, detJ, K, static_cast<cudaStream_t>(stream));
    return;
  }
  geometry_device<T>(coords, dphi, 0, nv, gdim, tdim, npts, J, detJ, K, nullptr); // validate the shapes
  if (nc == 0)
    return;
  require(coords != nullptr, "geometry: null pointer");
  const size_t xsz = size_t(nv) * gdim, jsz = size_t(gdim) * tdim * npts;
  const size_t per_cell = (xsz + (J ? jsz : 0) + (K ? jsz : 0) + (detJ ? size_t(npts) : 0)) * sizeof(T);
  const size_t chunk = chunk_units(nc, per_cell, 1);
  StagePool& st = stage_pool();
  DrainOnExit drain{st};
  T *dx[2], *dJ[2], *dd[2], *dK[2], *dphi_d = nullptr;
  for (int b = 0; b < 2; ++b)
  {
    dx[b] = st.get<T>(0, b, chunk * xsz * sizeof(T));
    dJ[b] = J ? st.get<T>(1, b, chunk * jsz * sizeof(T)) : nullptr;
    dd[b] = detJ ? st.get<T>(2, b, chunk * npts * sizeof(T)) : nullptr;
    dK[b] = K ? st.get<T>(3, b, chunk * jsz * sizeof(T)) : nullptr;
  }
  if (dphi)
  {
    // the derivatives of the coordinate element: one small table for all chunks, uploaded on stream 0 and read by
    // both streams (stream 1 waits for the upload through an event)
    dphi_d = st.get<T>(4, 0, size_t(tdim) * npts * nv * sizeof(T));
    BX_CUDA(cudaMemcpyAsync(dphi_d, dphi, size_t(tdim) * npts * nv * sizeof(T), cudaMemcpyHostToDevice, st.s[0]));
    cudaEvent_t ev;
    BX_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    BX_CUDA(cudaEventRecord(ev, st.s[0]));
    BX_CUDA(cudaStreamWaitEvent(st.s[1], ev, 0));
    BX_CUDA(cudaEventDestroy(ev));
  }
  size_t i = 0;
  for (size_t c0 = 0; c0 < nc; c0 += chunk, ++i)
  {
    const int b = int(i & 1);
    const size_t n = std::min(chunk, nc - c0);
    cudaStream_t s = st.s[b];
    BX_CUDA(cudaMemcpyAsync(dx[b], coords + c0 * xsz, n * xsz * sizeof(T), cudaMemcpyHostToDevice, s));
    geometry_device<T>(dx[b], dphi_d, n, nv, gdim, tdim, npts, dJ[b], dd[b], dK[b], s);
    if (J)
      BX_CUDA(cudaMemcpyAsync(J + c0 * jsz, dJ[b], n * jsz * sizeof(T), cudaMemcpyDeviceToHost, s));
    if (detJ)
      BX_CUDA(cudaMemcpyAsync(detJ + c0 * npts, dd[b], n * npts * sizeof(T), cudaMemcpyDeviceToHost, s));
    if (K)
      BX_CUDA(cudaMemcpyAsync(K + c0 * jsz, dK[b], n * jsz * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
  st.sync();
  st.trim();
}

// Fused tabulate-once -> T_apply -> push_forward.  X == nullptr: U[npts][ndofs][vs] holds the reference values;
// otherwise they are tabulated here from the points X[npts][tdim] (on the device; U is ignored).
template <typename T>
void physical_basis_any(const bx_element* e, int op, const T* X, const T* U, size_t npts, const T* coords, const T* J,
                        const T* detJ, const T* K, const uint32_t* cell_info, size_t nc, int gdim, T* out, int mem,
                        bx_stream_t stream, int per_point = 0)
{
  check_mem(mem);
  require(e != nullptr, "physical_basis: null element");
  require(X or U or npts == 0, "physical_basis: neither points nor reference values given");
  const size_t urow = size_t(e->ndofs) * e->vs;
  if (mem == BX_MEM_DEVICE)
  {
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!X or npts == 0 or nc == 0)
    {
      physical_basis_device<T>(*e, op, U, npts, coords, J, detJ, K, cell_info, nc, gdim, out, s, per_point);
      return;
    }
    T* dU = nullptr;
    BX_CUDA(cudaMallocAsync(&dU, npts * urow * sizeof(T), s));
    try
    {
      tabulate_device<T>(*e, 0, X, npts, dU, npts, s);
      physical_basis_device<T>(*e, op, dU, npts, coords, J, detJ, K, cell_info, nc, gdim, out, s, per_point);
    }
    catch (...)
    {
      cudaFreeAsync(dU, s);
      throw;
    }
    BX_CUDA(cudaFreeAsync(dU, s));
    return;
  }
  physical_basis_device<T>(*e, op, U, 0, coords, J, detJ, K, cell_info, 0, gdim, out, nullptr, per_point); // validate
  if (nc == 0 or npts == 0)
    return;
  require(out != nullptr, "physical_basis: null pointer");
  const int tdim = e->tdim, map = e->map_type;
  const int out_vs = map == BX_MAP_IDENTITY ? e->vs : gdim;
  // per-point geometry: every cell brings npts matrices (and determinants)
  const size_t pp = per_point ? npts : 1;
  const size_t xsz = coords ? size_t(tdim + 1) * gdim : 0, jsz = size_t(gdim) * tdim * pp;
  const bool need_K = !coords and map == BX_MAP_COVARIANT_PIOLA, need_J = !coords and map == BX_MAP_CONTRAVARIANT_PIOLA;
  const size_t orow = npts * size_t(e->ndofs) * out_vs;
  const size_t per_cell = (xsz + ((need_K or need_J) ? jsz : 0) + (need_J ? pp : 0) + orow) * sizeof(T) + sizeof(uint32_t);
  const size_t chunk = chunk_units(nc, per_cell, 1);
  StagePool& st = stage_pool();
  DrainOnExit drain{st};
  T *dg[2], *dd[2], *dout[2];
  uint32_t* dci[2];
  for (int b = 0; b < 2; ++b)
  {
    dg[b] = st.get<T>(0, b, chunk * std::max(xsz, jsz) * sizeof(T));
    dd[b] = need_J ? st.get<T>(1, b, chunk * pp * sizeof(T)) : nullptr;
    dci[b] = cell_info ? st.get<uint32_t>(2, b, chunk * sizeof(uint32_t)) : nullptr;
    dout[b] = st.get<T>(3, b, chunk * orow * sizeof(T));
  }
  // reference values: one slab for all chunks, produced on stream 0 (uploaded or tabulated there)
  T* dU = st.get<T>(4, 0, npts * urow * sizeof(T));
  if (X)
  {
    T* dX = st.get<T>(4, 1, npts * tdim * sizeof(T));
    BX_CUDA(cudaMemcpyAsync(dX, X, npts * tdim * sizeof(T), cudaMemcpyHostToDevice, st.s[0]));
    tabulate_device<T>(*e, 0, dX, npts, dU, npts, st.s[0]);
  }
  else
    BX_CUDA(cudaMemcpyAsync(dU, U, npts * urow * sizeof(T), cudaMemcpyHostToDevice, st.s[0]));
  {
    cudaEvent_t ev;
    BX_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    BX_CUDA(cudaEventRecord(ev, st.s[0]));
    BX_CUDA(cudaStreamWaitEvent(st.s[1], ev, 0));
    BX_CUDA(cudaEventDestroy(ev));
  }
  size_t i = 0;
  for (size_t c0 = 0; c0 < nc; c0 += chunk, ++i)
  {
    const int b = int(i & 1);
    const size_t n = std::min(chunk, nc - c0);
    cudaStream_t s = st.s[b];
    if (coords)
      BX_CUDA(cudaMemcpyAsync(dg[b], coords + c0 * xsz, n * xsz * sizeof(T), cudaMemcpyHostToDevice, s));
    else if (need_K)
      BX_CUDA(cudaMemcpyAsync(dg[b], K + c0 * jsz, n * jsz * sizeof(T), cudaMemcpyHostToDevice, s));
    else if (need_J)
    {
      BX_CUDA(cudaMemcpyAsync(dg[b], J + c0 * jsz, n * jsz * sizeof(T), cudaMemcpyHostToDevice, s));
      BX_CUDA(cudaMemcpyAsync(dd[b], detJ + c0 * pp, n * pp * sizeof(T), cudaMemcpyHostToDevice, s));
    }
    if (cell_info)
      BX_CUDA(cudaMemcpyAsync(dci[b], cell_info + c0, n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    physical_basis_device<T>(*e, op, dU, npts, coords ? dg[b] : nullptr, need_J ? dg[b] : nullptr, dd[b],
                             need_K ? dg[b] : nullptr, dci[b], n, gdim, dout[b], s, per_point);
    BX_CUDA(cudaMemcpyAsync(out + c0 * orow, dout[b], n * orow * sizeof(T), cudaMemcpyDeviceToHost, s));
  }
  st.sync();
  st.trim();
}

void closure_permute_any(const bx_element* e, int inverse, int32_t* d, size_t nc, size_t m, const uint32_t* info,
                         int entity_type, int entity_index, int mem, bx_stream_t stream)
{
  check_mem(mem);
  require(e != nullptr, "permute_subentity_closure: null element");
  require(nc == 0 or (d and info), "permute_subentity_closure: null pointer");
  if (mem == BX_MEM_DEVICE)
  {
    closure_permute_device(*e, inverse, d, nc, m, info, entity_type, entity_index, static_cast<cudaStream_t>(stream));
    return;
  }
  // validate once (throws on bad arguments / returns for vertices) before staging
  closure_permute_device(*e, inverse, nullptr, 0, m, nullptr, entity_type, entity_index, nullptr);
  if (nc == 0 or entity_type == BX_CELL_POINT)
    return;
  const size_t chunk = chunk_units(nc, 2 * m * sizeof(int32_t) + sizeof(uint32_t), 1);
  StagePool& st = stage_pool();
  DrainOnExit drain{st};
  int32_t* dd[2];
  uint32_t* dc[2];
  for (int b = 0; b < 2; ++b)
  {
    dd[b] = st.get<int32_t>(0, b, chunk * m * sizeof(int32_t));
    dc[b] = st.get<uint32_t>(1, b, chunk * sizeof(uint32_t));
  }
  size_t i = 0;
  for (size_t c0 = 0; c0 < nc; c0 += chunk, ++i)
  {
    const int b = int(i & 1);
    const size_t k = std::min(chunk, nc - c0);
    cudaStream_t s = st.s[b];
    BX_CUDA(cudaMemcpyAsync(dd[b], d + c0 * m, k * m * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    BX_CUDA(cudaMemcpyAsync(dc[b], info + c0, k * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    closure_permute_device(*e, inverse, dd[b], k, m, dc[b], entity_type, entity_index, s);
    BX_CUDA(cudaMemcpyAsync(d + c0 * m, dd[b], k * m * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  }
  st.sync();
  st.trim();
}
} // namespace
} // namespace bx

// =====================================================================================================
extern "C"
{
  const char* bx_last_error(void) { return bx::t_error.c_str(); }
  const char* bx_version(void) { return "basix_b200 0.1.0 (sm_100a)"; }

  int bx_device_count(void)
  {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess)
    {
      cudaGetLastError();
      return 0;
    }
    return n;
  }

  int bx_set_device(int device)
  {
    return bx::guarded([&]() { BX_CUDA(cudaSetDevice(device)); });
  }

  int bx_stream_synchronize(bx_stream_t stream)
  {
    return bx::guarded([&]() { BX_CUDA(cudaStreamSynchronize(static_cast<cudaStream_t>(stream))); });
  }

  uint64_t bx_kernel_launch_count(void) { return bx::g_launch_count.load(); }

  int bx_polyset_dim(int cell_type, int polyset_type, int degree)
  {
    return bx::polyset_dim(cell_type, polyset_type, degree);
  }
  int bx_polyset_nderivs(int cell_type, int nderiv) { return bx::polyset_nderivs(cell_type, nderiv); }

  int bx_polyset_tabulate_f64(int cell_type, int polyset_type, int degree, int nderiv, const double* x, size_t npts,
                              double* P, int mem, bx_stream_t stream)
  {
    return bx::guarded([&]() { bx::polyset_any<double>(cell_type, polyset_type, degree, nderiv, x, npts, P, mem, stream); });
  }
  int bx_polyset_tabulate_f32(int cell_type, int polyset_type, int degree, int nderiv, const float* x, size_t npts,
                              float* P, int mem, bx_stream_t stream)
  {
    return bx::guarded([&]() { bx::polyset_any<float>(cell_type, polyset_type, degree, nderiv, x, npts, P, mem, stream); });
  }

  int bx_element_create_from_arrays(const bx_element_desc* desc, bx_element** out)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(desc and out, "create_element: null argument");
          *out = nullptr;
          *out = bx::element_from_desc(*desc);
        });
  }

  void bx_element_destroy(bx_element* e) { delete e; }

  int bx_create_element(int family, int cell_type, int degree, int lagrange_variant, int dpc_variant,
                        int discontinuous, const int32_t* dof_ordering, int n_dof_ordering, bx_element** out)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(out != nullptr, "create_element: null argument");
          *out = nullptr;
          (void)dpc_variant;
          if (family == 1) // element::family::P
            *out = bx::create_lagrange(cell_type, degree, lagrange_variant, discontinuous != 0, dof_ordering,
                                       n_dof_ordering);
          else if (family == 2 or family == 3) // RT, N1E
          {
            if (n_dof_ordering != 0)
              bx::fail(BX_ERR_UNSUPPORTED, "create_element: dof_ordering is supported for the P family only");
            *out = family == 2 ? bx::create_rt(cell_type, degree, lagrange_variant, discontinuous != 0)
                               : bx::create_nedelec(cell_type, degree, lagrange_variant, discontinuous != 0);
          }
          else
            bx::fail(BX_ERR_UNSUPPORTED,
                     "create_element: the P, RT and N1E families are built natively; describe other families "
                     "with bx_element_create_from_arrays");
        });
  }

  int bx_element_coefficient_matrix(const bx_element* e, double* out)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e and out, "coefficient_matrix: null argument");
          std::copy(e->coeffs.begin(), e->coeffs.end(), out);
        });
  }

  int bx_element_entity_dofs(const bx_element* e, int dim, int index, int32_t* out)
  {
    if (!e or dim < 0 or dim >= int(e->edofs.size()) or index < 0 or index >= int(e->edofs[dim].size()))
      return -1;
    const auto& d = e->edofs[dim][index];
    if (out)
      std::copy(d.begin(), d.end(), out);
    return int(d.size());
  }

  int bx_element_entity_transformations(const bx_element* e, int entity_cell_type, int32_t shape[3], double* out)
  {
    if (!e or !shape or entity_cell_type < 0 or entity_cell_type > 7 or e->etrans_dim[entity_cell_type] < 0)
      return 0;
    const int n = e->etrans_dim[entity_cell_type];
    shape[0] = entity_cell_type == BX_CELL_INTERVAL ? 1 : 2;
    shape[1] = shape[2] = n;
    if (out)
      std::copy(e->etrans[entity_cell_type].begin(), e->etrans[entity_cell_type].end(), out);
    return 1;
  }

  int bx_element_dof_ordering(const bx_element* e, int32_t* out)
  {
    if (!e)
      return -1;
    if (out)
      std::copy(e->dof_ordering.begin(), e->dof_ordering.end(), out);
    return int(e->dof_ordering.size());
  }

  int bx_element_info(const bx_element* e, int32_t info[10])
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e and info, "element_info: null argument");
          info[0] = e->cell_type;
          info[1] = e->polyset_type;
          info[2] = e->tdim;
          info[3] = e->ndofs;
          info[4] = e->vs;
          info[5] = e->map_type;
          info[6] = e->superdegree;
          info[7] = e->perms;
          info[8] = e->identity;
          info[9] = e->psize;
        });
  }

  int bx_element_tabulate_shape(const bx_element* e, int nd, size_t npts, size_t shape[4])
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e and shape, "tabulate_shape: null argument");
          shape[0] = size_t(bx::polyset_nderivs(e->cell_type, nd));
          shape[1] = npts;
          shape[2] = size_t(e->ndofs);
          shape[3] = size_t(e->vs);
        });
  }

  int bx_element_net_operator(const bx_element* e, int op, uint32_t cell_info, int32_t* src, double* A)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr, "net_operator: null element");
          bx::require(op >= 0 and op <= 7, "net_operator: unknown operator");
          static const int kind_of[8] = {0, 1, 2, 3, 0, 2, 1, 3};
          bx::element_net_operator(*e, kind_of[op], cell_info, src, A);
        });
  }

  int bx_element_tabulate_path(const bx_element* e, int nd) { return e ? bx::tabulate_path(*e, nd) : -1; }
  unsigned bx_tabulate_path_mask(unsigned mask) { return bx::tabulate_path_mask(mask); }

  int bx_tabulate_f64(const bx_element* e, int nd, const double* x, size_t npts, int tdim_x, double* basis, int mem,
                      bx_stream_t stream)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr, "tabulate: null element");
          bx::tabulate_any<double>(*e, nd, x, npts, tdim_x, basis, mem, stream);
        });
  }

  int bx_tabulate_f32(const bx_element* e, int nd, const float* x, size_t npts, int tdim_x, float* basis, int mem,
                      bx_stream_t stream)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr, "tabulate: null element");
          bx::tabulate_any<float>(*e, nd, x, npts, tdim_x, basis, mem, stream);
        });
  }

  int bx_map_value_size(int map_type, int pull, int in_vs, int gdim, int tdim)
  {
    int v = -1;
    bx::guarded([&]() { v = bx::map_value_size(map_type, pull, in_vs, gdim, tdim); });
    return v;
  }

  int bx_map_f64(int map_type, int pull, const double* U, const double* J, const double* detJ, const double* K,
                 size_t nc, size_t rows, int in_vs, int gdim, int tdim, double* out, int mem, bx_stream_t stream)
  {
    return bx::guarded([&]() { bx::map_any<double>(map_type, pull, U, J, detJ, K, nc, rows, in_vs, gdim, tdim, out, mem, stream); });
  }

  int bx_push_forward_f64(const bx_element* e, const double* U, const double* J, const double* detJ, const double* K,
                          size_t nc, size_t rows, int in_vs, int gdim, int tdim, double* out, int mem,
                          bx_stream_t stream)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr, "push_forward: null element");
          bx::map_any<double>(e->map_type, 0, U, J, detJ, K, nc, rows, in_vs, gdim, tdim, out, mem, stream);
        });
  }

  int bx_push_forward_broadcast_f64(const bx_element* e, const double* U, const double* J, const double* detJ,
                                    const double* K, size_t nc, size_t rows, int in_vs, int gdim, int tdim, double* out,
                                    int mem, bx_stream_t stream)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr, "push_forward: null element");
          bx::map_any<double>(e->map_type, 0, U, J, detJ, K, nc, rows, in_vs, gdim, tdim, out, mem, stream, 1);
        });
  }

  int bx_pull_back_f64(const bx_element* e, const double* u, const double* J, const double* detJ, const double* K,
                       size_t nc, size_t rows, int in_vs, int gdim, int tdim, double* out, int mem, bx_stream_t stream)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr, "pull_back: null element");
          bx::map_any<double>(e->map_type, 1, u, J, detJ, K, nc, rows, in_vs, gdim, tdim, out, mem, stream);
        });
  }

  int bx_map_f32(int map_type, int pull, const float* U, const float* J, const float* detJ, const float* K, size_t nc,
                 size_t rows, int in_vs, int gdim, int tdim, float* out, int mem, bx_stream_t stream)
  {
    return bx::guarded([&]() { bx::map_any<float>(map_type, pull, U, J, detJ, K, nc, rows, in_vs, gdim, tdim, out, mem, stream); });
  }

  int bx_push_forward_f32(const bx_element* e, const float* U, const float* J, const float* detJ, const float* K,
                          size_t nc, size_t rows, int in_vs, int gdim, int tdim, float* out, int mem,
                          bx_stream_t stream)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr, "push_forward: null element");
          bx::map_any<float>(e->map_type, 0, U, J, detJ, K, nc, rows, in_vs, gdim, tdim, out, mem, stream);
        });
  }

  int bx_push_forward_broadcast_f32(const bx_element* e, const float* U, const float* J, const float* detJ, const float* K,
                                    size_t nc, size_t rows, int in_vs, int gdim, int tdim, float* out, int mem,
                                    bx_stream_t stream)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr, "push_forward: null element");
          bx::map_any<float>(e->map_type, 0, U, J, detJ, K, nc, rows, in_vs, gdim, tdim, out, mem, stream, 1);
        });
  }

  int bx_pull_back_f32(const bx_element* e, const float* u, const float* J, const float* detJ, const float* K,
                       size_t nc, size_t rows, int in_vs, int gdim, int tdim, float* out, int mem, bx_stream_t stream)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr, "pull_back: null element");
          bx::map_any<float>(e->map_type, 1, u, J, detJ, K, nc, rows, in_vs, gdim, tdim, out, mem, stream);
        });
  }

  int bx_T_apply_f64(const bx_element* e, int op, double* data, size_t nc, size_t cell_size, int n,
                     const uint32_t* cell_info, int mem, bx_stream_t stream)
  {
    return bx::guarded([&]() { bx::transform_any<double>(e, op, data, nc, cell_size, n, cell_info, mem, stream); });
  }
  int bx_T_apply_f32(const bx_element* e, int op, float* data, size_t nc, size_t cell_size, int n,
                     const uint32_t* cell_info, int mem, bx_stream_t stream)
  {
    return bx::guarded([&]() { bx::transform_any<float>(e, op, data, nc, cell_size, n, cell_info, mem, stream); });
  }
  int bx_T_apply_i32(const bx_element* e, int op, int32_t* data, size_t nc, size_t cell_size, int n,
                     const uint32_t* cell_info, int mem, bx_stream_t stream)
  {
    return bx::guarded([&]() { bx::transform_any<int32_t>(e, op, data, nc, cell_size, n, cell_info, mem, stream); });
  }

  int bx_element_interpolation_shape(const bx_element* e, int32_t* shape)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr and shape != nullptr, "interpolation_shape: null pointer");
          shape[0] = e->interp_npts;
          shape[1] = e->tdim;
          shape[2] = e->ndofs;
          shape[3] = e->vs * e->interp_npts;
        });
  }
  int bx_element_points(const bx_element* e, double* out)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr and out != nullptr, "points: null pointer");
          std::copy(e->points.begin(), e->points.end(), out);
        });
  }
  int bx_element_interpolation_matrix(const bx_element* e, double* out)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr and out != nullptr, "interpolation_matrix: null pointer");
          std::copy(e->interp_M.begin(), e->interp_M.end(), out);
        });
  }
  int bx_interpolate_f64(const bx_element* e, int layout, const double* values, size_t nc, double* coefficients, int mem,
                         bx_stream_t stream)
  {
    return bx::guarded([&]() { bx::interpolate_any<double>(e, layout, values, nc, coefficients, mem, stream); });
  }
  int bx_interpolate_pulled_f64(const bx_element* e, const double* u, const double* J, const double* detJ, const double* K,
                                size_t nc, int gdim, double* coefficients, int mem, bx_stream_t stream)
  {
    return bx::guarded([&]() { bx::interpolate_pulled_any<double>(e, u, J, detJ, K, nc, gdim, coefficients, mem, stream); });
  }
  int bx_interpolate_pulled_f32(const bx_element* e, const float* u, const float* J, const float* detJ, const float* K, size_t nc,
                                int gdim, float* coefficients, int mem, bx_stream_t stream)
  {
    return bx::guarded([&]() { bx::interpolate_pulled_any<float>(e, u, J, detJ, K, nc, gdim, coefficients, mem, stream); });
  }
  int bx_interpolate_f32(const bx_element* e, int layout, const float* values, size_t nc, float* coefficients, int mem,
                         bx_stream_t stream)
  {
    return bx::guarded([&]() { bx::interpolate_any<float>(e, layout, values, nc, coefficients, mem, stream); });
  }

  int bx_element_subentity_closure_size(const bx_element* e, int entity_type)
  {
    if (!e or entity_type < 0 or entity_type >= 8)
      return -1;
    return e->closure[entity_type].size;
  }

  int bx_element_subentity_closure_map(const bx_element* e, int entity_type, int inverse, uint32_t entity_info, int32_t* src)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr and src != nullptr, "subentity_closure_map: null pointer");
          bx::require(entity_type >= 0 and entity_type < 8 and e->closure[entity_type].size > 0,
                      "subentity_closure_map: the element has no closure permutation for this sub-entity type");
          const auto& cl = e->closure[entity_type];
          const uint32_t st = entity_info & uint32_t(cl.states - 1);
          for (int k = 0; k < cl.size; ++k)
            src[k] = cl.src[(size_t(inverse ? 1 : 0) * cl.states + st) * cl.size + k];
        });
  }

  int bx_permute_subentity_closure_i32(const bx_element* e, int inverse, int32_t* d, size_t nc, size_t m,
                                       const uint32_t* info, int entity_type, int entity_index, int mem,
                                       bx_stream_t stream)
  {
    return bx::guarded(
        [&]() { bx::closure_permute_any(e, inverse, d, nc, m, info, entity_type, entity_index, mem, stream); });
  }

  int bx_geometry_f64(const double* coords, const double* dphi, size_t nc, int nv, int gdim, int tdim, int npts, double* J,
                      double* detJ, double* K, int mem, bx_stream_t stream)
  {
    return bx::guarded([&]() { bx::geometry_any<double>(coords, dphi, nc, nv, gdim, tdim, npts, J, detJ, K, mem, stream); });
  }
  int bx_geometry_f32(const float* coords, const float* dphi, size_t nc, int nv, int gdim, int tdim, int npts, float* J,
                      float* detJ, float* K, int mem, bx_stream_t stream)
  {
    return bx::guarded([&]() { bx::geometry_any<float>(coords, dphi, nc, nv, gdim, tdim, npts, J, detJ, K, mem, stream); });
  }
  int bx_push_forward_fused_f64(const bx_element* e, int op, const double* U, size_t npts, const double* coords,
                                const double* J, const double* detJ, const double* K, const uint32_t* cell_info, size_t nc,
                                int gdim, double* out, int mem, bx_stream_t stream)
  {
    return bx::guarded([&]()
                       { bx::physical_basis_any<double>(e, op, nullptr, U, npts, coords, J, detJ, K, cell_info, nc, gdim, out, mem, stream); });
  }
  int bx_push_forward_fused_f32(const bx_element* e, int op, const float* U, size_t npts, const float* coords,
                                const float* J, const float* detJ, const float* K, const uint32_t* cell_info, size_t nc,
                                int gdim, float* out, int mem, bx_stream_t stream)
  {
    return bx::guarded([&]()
                       { bx::physical_basis_any<float>(e, op, nullptr, U, npts, coords, J, detJ, K, cell_info, nc, gdim, out, mem, stream); });
  }
  int bx_physical_basis_f64(const bx_element* e, int op, const double* X, size_t npts, int tdim_x, const double* coords,
                            const double* J, const double* detJ, const double* K, const uint32_t* cell_info, size_t nc,
                            int gdim, double* out, int mem, bx_stream_t stream)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr, "physical_basis: null element");
          if (tdim_x != e->tdim)
            bx::fail(BX_ERR_INVALID_ARG, "Point dim (" + std::to_string(tdim_x) + ") does not match element dim ("
                                             + std::to_string(e->tdim) + ").");
          bx::require(X != nullptr or npts == 0, "physical_basis: null points");
          bx::physical_basis_any<double>(e, op, X, nullptr, npts, coords, J, detJ, K, cell_info, nc, gdim, out, mem, stream);
        });
  }

  int bx_physical_basis_pointwise_f64(const bx_element* e, int op, const double* X, size_t npts, int tdim_x, const double* J,
                                      const double* detJ, const double* K, const uint32_t* cell_info, size_t nc, int gdim,
                                      double* out, int mem, bx_stream_t stream)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr, "physical_basis: null element");
          if (tdim_x != e->tdim)
            bx::fail(BX_ERR_INVALID_ARG, "Point dim (" + std::to_string(tdim_x) + ") does not match element dim ("
                                             + std::to_string(e->tdim) + ").");
          bx::require(X != nullptr or npts == 0, "physical_basis: null points");
          bx::physical_basis_any<double>(e, op, X, nullptr, npts, nullptr, J, detJ, K, cell_info, nc, gdim, out, mem, stream, 1);
        });
  }
  int bx_push_forward_fused_pointwise_f32(const bx_element* e, int op, const float* U, size_t npts, const float* J,
                                          const float* detJ, const float* K, const uint32_t* cell_info, size_t nc, int gdim,
                                          float* out, int mem, bx_stream_t stream)
  {
    return bx::guarded(
        [&]()
        { bx::physical_basis_any<float>(e, op, nullptr, U, npts, nullptr, J, detJ, K, cell_info, nc, gdim, out, mem, stream, 1); });
  }
  int bx_push_forward_fused_pointwise_f64(const bx_element* e, int op, const double* U, size_t npts, const double* J,
                                          const double* detJ, const double* K, const uint32_t* cell_info, size_t nc, int gdim,
                                          double* out, int mem, bx_stream_t stream)
  {
    return bx::guarded(
        [&]()
        { bx::physical_basis_any<double>(e, op, nullptr, U, npts, nullptr, J, detJ, K, cell_info, nc, gdim, out, mem, stream, 1); });
  }

  int bx_compute_interpolation_operator_f64(const bx_element* from, const bx_element* to, int32_t shape[2], double* out,
                                            int mem, bx_stream_t stream)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(from and to and shape, "compute_interpolation_operator: null argument");
          bx::check_mem(mem);
          bx::interpolation_operator_shape(*from, *to, shape);
          if (!out)
            return;
          if (mem == BX_MEM_DEVICE)
          {
            bx::interpolation_operator_device(*from, *to, out, static_cast<cudaStream_t>(stream));
            return;
          }
          bx::StagePool& st = bx::stage_pool();
          bx::DrainOnExit drain{st};
          const size_t bytes = size_t(shape[0]) * shape[1] * sizeof(double);
          double* d = st.get<double>(0, 0, bytes);
          bx::interpolation_operator_device(*from, *to, d, st.s[0]);
          BX_CUDA(cudaMemcpyAsync(out, d, bytes, cudaMemcpyDeviceToHost, st.s[0]));
          st.sync();
        });
  }

  int bx_create_lattice(int cell_type, int n, int lattice_type, int exterior, int simplex_method, double* out, size_t* npts)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(npts != nullptr, "create_lattice: null argument");
          const std::vector<double> x = bx::create_lattice_points(cell_type, n, lattice_type, exterior, simplex_method);
          const int tdim = bx::cell_tdim(cell_type);
          *npts = tdim > 0 ? x.size() / size_t(tdim) : 1;
          if (out)
            std::copy(x.begin(), x.end(), out);
        });
  }

  int bx_make_quadrature(int cell_type, int degree, double* points, double* weights, size_t* npts)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(npts != nullptr, "make_quadrature: null argument");
          std::vector<double> x, w;
          bx::make_quadrature_gauss_jacobi(cell_type, degree, x, w);
          *npts = w.size();
          if (points)
            std::copy(x.begin(), x.end(), points);
          if (weights)
            std::copy(w.begin(), w.end(), weights);
        });
  }

  int bx_gauss_jacobi_rule(double alpha, int npoints, double* points, double* weights)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(points != nullptr and weights != nullptr, "gauss_jacobi_rule: null argument");
          std::vector<double> x, w;
          bx::gauss_jacobi_rule_unit(alpha, npoints, x, w);
          std::copy(x.begin(), x.end(), points);
          std::copy(w.begin(), w.end(), weights);
        });
  }

  int bx_element_prepare(const bx_element* e, int nd)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr and nd >= 0, "element_prepare: bad argument");
          for (int k = 0; k <= nd; ++k)
            (void)bx::tabulate_path(*e, k); // builds and uploads the operands of order k on first use
        });
  }

  int bx_permute_i32(const bx_element* e, int inverse, int32_t* d, size_t nc, const uint32_t* cell_info, int mem,
                     bx_stream_t stream)
  {
    return bx::guarded(
        [&]()
        {
          bx::require(e != nullptr, "permute: null element");
          if (!e->perms)
            bx::fail(BX_ERR_RUNTIME, "The DOF transformations for this element are not permutations");
          // permute -> permute_data<false>(_eperm) = T_apply's operator; permute_inv ->
          // permute_data<true>(_eperm_inv) = Tt_apply's operator (finite-element.h:800-843)
          bx::transform_any<int32_t>(e, inverse ? BX_TT_APPLY : BX_T_APPLY, d, nc, size_t(e->ndofs), 1, cell_info, mem,
                                     stream);
        });
  }
}

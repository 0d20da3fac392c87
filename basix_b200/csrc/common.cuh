// Shared host/device helpers for the basix_b200 CUDA library (sm_100a only).
#pragma once

#include <nvtx3/nvToolsExt.h>
#include <cuda_runtime.h>

#include <atomic>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/basix_b200.h"

namespace bx
{
// ---- error plumbing: int status + thread-local message (SURVEY 8b "errors") -----------------
struct Error : std::runtime_error
{
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string& msg);
extern std::atomic<uint64_t> g_launch_count;

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line)
{
  if (e != cudaSuccess)
  {
    throw Error(BX_ERR_CUDA, std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what + " ("
                                 + file + ":" + std::to_string(line) + ")");
  }
}
#define BX_CUDA(expr) ::bx::cuda_check((expr), #expr, __FILE__, __LINE__)
// After a kernel launch: count it and surface launch-configuration errors.
#define BX_LAUNCHED()                                                                                    \
  do                                                                                                     \
  {                                                                                                      \
    ::bx::g_launch_count.fetch_add(1, std::memory_order_relaxed);                                        \
    ::bx::cuda_check(cudaGetLastError(), "kernel launch", __FILE__, __LINE__);                           \
  } while (0)

// NVTX range named after the C-ABI entry point (header-only NVTX 3: a no-op unless a tool is attached), so that
// ncu --nvtx / Nsight timelines show which library call a kernel belongs to.
struct NvtxRange
{
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};

template <typename F>
int guarded(F&& f, const char* entry = __builtin_FUNCTION())
{
  NvtxRange range(entry);
  try
  {
    f();
    return BX_OK;
  }
  catch (const Error& e)
  {
    set_last_error(e.what());
    return e.code;
  }
  catch (const std::exception& e)
  {
    set_last_error(e.what());
    return BX_ERR_RUNTIME;
  }
}

[[noreturn]] inline void fail(int code, const std::string& msg) { throw Error(code, msg); }

// ---- index helpers: reference indexing.h:13-34 ------------------------------------------------
__host__ __device__ constexpr int idx2(int p, int q) { return (p + q + 1) * (p + q) / 2 + q; }
__host__ __device__ constexpr int idx3(int p, int q, int r)
{
  return (p + q + r) * (p + q + r + 1) * (p + q + r + 2) / 6 + (q + r) * (q + r + 1) / 2 + r;
}

__host__ __device__ constexpr size_t ceil_div(size_t a, size_t b) { return (a + b - 1) / b; }

// ---- cell topology (reference cell.cpp topology tables; only counts and face types needed) ---
int cell_tdim(int cell_type);
int cell_num_entities(int cell_type, int dim);
// sub-entity cell type of entity (dim, index)
int cell_sub_entity_type(int cell_type, int dim, int index);

int polyset_dim(int cell_type, int polyset_type, int degree);
int polyset_nderivs(int cell_type, int nderiv);

int sm_count();
} // namespace bx

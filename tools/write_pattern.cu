// HBM write bandwidth of the tabulate output pattern, without any arithmetic: does the store pattern of the
// tensor-product kernel (each warp writes 32 rows of N doubles into each of NDER slabs, 8 bytes per lane, rows
// not sector aligned) cap the bandwidth, and what does a staged 16-byte / row-aligned variant reach?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/write_pattern tools/write_pattern.cu
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#define CK(x)                                                                                     \
  do                                                                                              \
  {                                                                                               \
    cudaError_t e_ = (x);                                                                         \
    if (e_ != cudaSuccess)                                                                        \
    {                                                                                             \
      fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);    \
      exit(1);                                                                                    \
    }                                                                                             \
  } while (0)

// mode 0: tp pattern (point outer, column group inner, all slabs per point), 8-byte stores
// mode 1: slab outer (one slab region of 32 rows at a time), 8-byte stores
// mode 2: slab outer, the 32 x N region written as one contiguous run with 16-byte stores
__global__ void __launch_bounds__(256) pattern_kernel(double* out, size_t npts, int N, int nder, int mode)
{
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const size_t ntiles = npts / 32;
  const double v = 1.0 + lane;
  for (size_t tile = size_t(blockIdx.x) * 8 + warp; tile < ntiles; tile += size_t(gridDim.x) * 8)
  {
    if (mode == 0)
    {
      for (int q = 0; q < 32; ++q)
        for (int m = 0; m * 32 < N; ++m)
          if (lane + 32 * m < N)
            for (int d = 0; d < nder; ++d)
              __stcs(out + (size_t(d) * npts + tile * 32 + q) * N + lane + 32 * m, v);
    }
    else if (mode == 1)
    {
      for (int d = 0; d < nder; ++d)
        for (int q = 0; q < 32; ++q)
          for (int m = 0; m * 32 < N; ++m)
            if (lane + 32 * m < N)
              __stcs(out + (size_t(d) * npts + tile * 32 + q) * N + lane + 32 * m, v);
    }
    else
    {
      for (int d = 0; d < nder; ++d)
      {
        double2* base = reinterpret_cast<double2*>(out + (size_t(d) * npts + tile * 32) * N);
        for (int i = lane; i < 16 * N; i += 32)
          __stcs(base + i, make_double2(v, v));
      }
    }
  }
}

int main()
{
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int N = 125, nder = 4;
  const size_t npts = 10'000'000;
  double* out;
  CK(cudaMalloc(&out, npts * N * nder * sizeof(double)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int per_sm : {2, 4, 8})
    for (int mode = 0; mode < 3; ++mode)
    {
      float best = 1e30f;
      for (int r = 0; r < 4; ++r)
      {
        CK(cudaEventRecord(e0));
        pattern_kernel<<<prop.multiProcessorCount * per_sm, 256>>>(out, npts, N, nder, mode);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 and ms < best)
          best = ms;
      }
      printf("ctas/sm %d mode %d: %.3f ms, %.0f GB/s\n", per_sm, mode, best, double(npts) * N * nder * 8 / (best * 1e-3) / 1e9);
    }
  return 0;
}

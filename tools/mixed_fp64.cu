// Do DMMA (FP64 tensor, mma.sync m8n8k4) and DFMA share one execution pipe on B200?
// Half of the warps of each CTA issue DMMA chains, the other half DFMA chains; the combined rate is compared
// with each kind running alone with the same number of warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/mixed_fp64 tools/mixed_fp64.cu
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

// mode: 0 = every warp DMMA, 1 = every warp DFMA, 2 = warps alternate (even DMMA, odd DFMA),
//       3 = warps 0..n/2-1 DMMA, the rest DFMA (so each SM sub-partition sees both kinds)
__global__ void mixed_kernel(double* out, int iters_mma, int iters_fma, int mode, double a, double b)
{
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  bool mma = mode == 0 ? true : mode == 1 ? false : mode == 2 ? (warp & 1) == 0 : warp < nw / 2;
  double s = 0;
  if (mma)
  {
    double d0[8], d1[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      d0[c] = d1[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters_mma; ++i)
    {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                     : "+d"(d0[c]), "+d"(d1[c])
                     : "d"(a), "d"(b));
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
      s += d0[c] + d1[c];
  }
  else
  {
    double acc[8];
#pragma unroll
    for (int c = 0; c < 8; ++c)
      acc[c] = threadIdx.x * 1e-3 + c;
    for (int i = 0; i < iters_fma; ++i)
    {
#pragma unroll
      for (int c = 0; c < 8; ++c)
        acc[c] = fma(acc[c], a, b);
    }
#pragma unroll
    for (int c = 0; c < 8; ++c)
      s += acc[c];
  }
  if (s == 123.456)
    out[0] = s;
}

int main()
{
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double* d;
  CK(cudaMalloc(&d, 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("{\"gpu\": \"%s\"", prop.name);
  for (int warps : {8, 16, 32})
  {
    for (int mode = 0; mode < 4; ++mode)
    {
      // one DMMA = 512 flop per warp; one DFMA = 64 flop per warp: 8 DFMA iterations per DMMA iteration keep
      // the two kinds' run times comparable when the pipes are separate and of similar peak
      const int im = 4000, ifm = 32000;
      float best = 1e30f;
      for (int r = 0; r < 4; ++r)
      {
        CK(cudaEventRecord(e0));
        mixed_kernel<<<sms, warps * 32>>>(d, im, ifm, mode, 1.0000001, 1e-9);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r > 0 and ms < best)
          best = ms;
      }
      const double wm = mode == 0 ? warps : mode == 1 ? 0 : warps / 2.0;
      const double wf = warps - wm;
      const double fl_m = wm * sms * double(im) * 8 * 512, fl_f = wf * sms * double(ifm) * 8 * 64;
      printf(", \"w%d_mode%d\": {\"ms\": %.3f, \"dmma_tf\": %.2f, \"dfma_tf\": %.2f, \"total_tf\": %.2f}", warps, mode, best,
             fl_m / (best * 1e-3) / 1e12, fl_f / (best * 1e-3) / 1e12, (fl_m + fl_f) / (best * 1e-3) / 1e12);
    }
  }
  printf("}\n");
  return 0;
}

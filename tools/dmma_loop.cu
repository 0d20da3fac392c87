// Loop-shape study for the derivative-operator tabulate kernel: how close to the FP64 tensor peak does one warp's
// contraction loop (A and B fragments from shared memory, 2 x NT accumulator tiles) get, as a function of the
// number of warps per SM sub-partition and of the software pipelining of the fragment loads?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_loop tools/dmma_loop.cu
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

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b)
{
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(d0), "+d"(d1)
      : "d"(a), "d"(b));
}

constexpr int NT = 7;
__host__ __device__ constexpr int pstride(int mt) { return mt == 2 ? 20 : 36; }

// VARIANT 0: loads at the top of each k-step (unroll 2).  1: register double buffering (loads of step ks+1 issued
// before the DMMAs of step ks).  2: as 0 with unroll 1.  3: as 1 but B fragment reused by 4 row tiles (32 points).
template <int VARIANT, int MT>
__device__ __forceinline__ void item(const double* __restrict__ Pa, const double* __restrict__ Bp, int ksteps, int kstride,
                                     double (&sink)[2])
{
  double acc[MT][NT][2];
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int j = 0; j < NT; ++j)
      acc[m][j][0] = acc[m][j][1] = 0.0;
  const int ts = 8 * kstride;
  if constexpr (VARIANT == 0 or VARIANT == 2)
  {
#pragma unroll(VARIANT == 0 ? 2 : 1)
    for (int ks = 0; ks < ksteps; ++ks)
    {
      double a[MT];
#pragma unroll
      for (int m = 0; m < MT; ++m)
        a[m] = Pa[ks * 4 * pstride(MT) + 8 * m];
#pragma unroll
      for (int j = 0; j < NT; ++j)
      {
        const double b = Bp[j * ts + ks * 4];
#pragma unroll
        for (int m = 0; m < MT; ++m)
          dmma(acc[m][j][0], acc[m][j][1], a[m], b);
      }
    }
  }
  else if constexpr (VARIANT == 3 or VARIANT == 4)
  {
    // 3: fragments loaded once per item, none inside the loop.  4: only the A fragments are reloaded per k-step.
    double a[MT], b[NT];
#pragma unroll
    for (int m = 0; m < MT; ++m)
      a[m] = Pa[8 * m];
#pragma unroll
    for (int j = 0; j < NT; ++j)
      b[j] = Bp[j * ts];
#pragma unroll 1
    for (int ks = 0; ks < ksteps; ++ks)
    {
      if constexpr (VARIANT == 4)
      {
#pragma unroll
        for (int m = 0; m < MT; ++m)
          a[m] = Pa[ks * 4 * pstride(MT) + 8 * m];
      }
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int m = 0; m < MT; ++m)
          dmma(acc[m][j][0], acc[m][j][1], a[m], b[j]);
    }
  }
  else
  {
    double a[MT], b[NT];
#pragma unroll
    for (int m = 0; m < MT; ++m)
      a[m] = Pa[8 * m];
#pragma unroll
    for (int j = 0; j < NT; ++j)
      b[j] = Bp[j * ts];
#pragma unroll 1
    for (int ks = 0; ks < ksteps; ++ks)
    {
      double an[MT], bn[NT];
      const int kn = ks + 1 < ksteps ? ks + 1 : ks;
#pragma unroll
      for (int m = 0; m < MT; ++m)
        an[m] = Pa[kn * 4 * pstride(MT) + 8 * m];
#pragma unroll
      for (int j = 0; j < NT; ++j)
        bn[j] = Bp[j * ts + kn * 4];
#pragma unroll
      for (int j = 0; j < NT; ++j)
#pragma unroll
        for (int m = 0; m < MT; ++m)
          dmma(acc[m][j][0], acc[m][j][1], a[m], b[j]);
#pragma unroll
      for (int m = 0; m < MT; ++m)
        a[m] = an[m];
#pragma unroll
      for (int j = 0; j < NT; ++j)
        b[j] = bn[j];
    }
  }
#pragma unroll
  for (int m = 0; m < MT; ++m)
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
      sink[0] += acc[m][j][0];
      sink[1] += acc[m][j][1];
    }
}

template <int VARIANT, int MT>
__global__ void __launch_bounds__(MT == 4 ? 256 : 512, 1) loop_kernel(double* out, int iters, int long_k)
{
  extern __shared__ __align__(16) double smem[];
  // B: 10 slabs as in P5 / nd = 2: strides 60, 36 x 3, 20 x 6, 56 rows each; then one P slab per warp
  const int kst[10] = {14, 9, 9, 9, 5, 5, 5, 5, 5, 5};
  const int strd[10] = {60, 36, 36, 36, 20, 20, 20, 20, 20, 20};
  int boff[10], total = 0;
  for (int i = 0; i < 10; ++i)
  {
    boff[i] = total;
    total += 56 * strd[i];
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  for (int i = threadIdx.x; i < total + nw * 56 * pstride(MT); i += blockDim.x)
    smem[i] = 1.0 / (1 + i % 97);
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  const double* Pa = smem + total + warp * 56 * pstride(MT) + t * pstride(MT) + g;
  double sink[2] = {0, 0};
  if (long_k)
    item<VARIANT, MT>(Pa, smem + boff[9] + g * strd[9] + t, 71 * iters, 0, sink); // one long chain, fragments wrap in place
  else
  for (int it = 0; it < iters; ++it)
  {
#pragma unroll 1
    for (int i = 0; i < 10; ++i)
      item<VARIANT, MT>(Pa, smem + boff[i] + g * strd[i] + t, kst[i], strd[i], sink);
  }
  if (sink[0] + sink[1] == 123.456)
    out[0] = sink[0];
}

template <int VARIANT, int MT>
void run(const char* name, int sms, int warps, double* d, int long_k = 0)
{
  const int iters = 200;
  const size_t smem = (56 * (60 + 3 * 36 + 6 * 20) + warps * 56 * pstride(MT)) * sizeof(double);
  if (smem > 227 * 1024 or (MT == 4 and warps > 8))
    return;
  auto k = loop_kernel<VARIANT, MT>;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for (int r = 0; r < 4; ++r)
  {
    CK(cudaEventRecord(e0));
    k<<<sms, warps * 32, smem>>>(d, iters, long_k);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (r > 0 and ms < best)
      best = ms;
  }
  CK(cudaGetLastError());
  const double dmmas = double(iters) * (14 + 27 + 30) * NT * MT;
  printf("%s%s MT=%d warps=%d: %.3f ms, %.2f TFLOP/s\n", name, long_k ? " [one long item]" : "", MT, warps, best,
         dmmas * 512 * warps * sms / (best * 1e-3) / 1e12);
}

int main()
{
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double* d;
  CK(cudaMalloc(&d, 1024));
  for (int warps : {4, 8})
  {
    run<3, 2>("no loads in loop", sms, warps, d, 1);
    run<3, 2>("no loads in loop", sms, warps, d);
    run<4, 2>("A loads only", sms, warps, d);
    run<0, 2>("loads-at-top unroll2", sms, warps, d);
    run<2, 2>("loads-at-top unroll1", sms, warps, d);
    run<1, 2>("register double buffer", sms, warps, d);
    run<0, 4>("loads-at-top unroll2", sms, warps, d);
    run<1, 4>("register double buffer", sms, warps, d);
  }
  return 0;
}

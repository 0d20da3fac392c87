// Microbenchmarks that fix the roofline denominators MEASURED_PEAKS.json does not hold:
//   * FP64 FMA (DFMA) peak, * FP64 tensor (DMMA m8n8k4) peak with 1..8 independent chains,
//   * HBM copy / write-only / read-only bandwidth with this library's access widths.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/peaks tools/peaks.cu
// Output: one JSON object on stdout.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

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

template <int CHAINS>
__global__ void dfma_kernel(double* out, int iters, double a, double b)
{
  double acc[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c)
    acc[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < iters; ++i)
  {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
      acc[c] = fma(acc[c], a, b);
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c)
    s += acc[c];
  if (s == 123.456)
    out[0] = s;
}

template <int CHAINS>
__global__ void dmma_kernel(double* out, int iters, double a, double b)
{
  double d0[CHAINS], d1[CHAINS];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c)
    d0[c] = d1[c] = threadIdx.x * 1e-3 + c;
  for (int i = 0; i < iters; ++i)
  {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                   : "+d"(d0[c]), "+d"(d1[c])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c)
    s += d0[c] + d1[c];
  if (s == 123.456)
    out[0] = s;
}

__global__ void copy_kernel(const double2* __restrict__ in, double2* __restrict__ out, size_t n)
{
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = in[i];
}
__global__ void write_kernel(double2* __restrict__ out, size_t n, double v)
{
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = make_double2(v, v);
}
__global__ void read_kernel(const double2* __restrict__ in, size_t n, double* sink)
{
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  double s = 0;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
  {
    double2 v = in[i];
    s += v.x + v.y;
  }
  if (s == 123.456)
    sink[0] = s;
}

template <typename F>
float time_ms(F&& f, int reps)
{
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; ++r)
  {
    CK(cudaEventRecord(a));
    f();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    best = ms < best ? ms : best;
  }
  return best;
}

template <int CH>
double dfma_tflops(double* d, int sms, int warps)
{
  const int iters = 20000;
  float ms = time_ms([&]() { dfma_kernel<CH><<<sms * 2, warps * 32 / 2>>>(d, iters, 1.0000001, 1e-9); }, 3);
  return 2.0 * CH * double(iters) * sms * warps * 32 / (ms * 1e-3) / 1e12;
}
template <int CH>
double dmma_tflops(double* d, int sms, int warps)
{
  const int iters = 10000;
  float ms = time_ms([&]() { dmma_kernel<CH><<<sms * 2, warps * 32 / 2>>>(d, iters, 1.0000001, 1e-9); }, 3);
  return 2.0 * 256 * CH * double(iters) * sms * warps / (ms * 1e-3) / 1e12;
}

int main()
{
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  double* d;
  CK(cudaMalloc(&d, 1024));
  printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
  printf(", \"dfma_tflops\": {\"w8_c4\": %.2f, \"w16_c4\": %.2f, \"w32_c4\": %.2f, \"w16_c8\": %.2f}",
         dfma_tflops<4>(d, sms, 8), dfma_tflops<4>(d, sms, 16), dfma_tflops<4>(d, sms, 32), dfma_tflops<8>(d, sms, 16));
  printf(", \"dmma_tflops\": {\"w4_c1\": %.2f, \"w4_c4\": %.2f, \"w4_c8\": %.2f, \"w8_c4\": %.2f, \"w8_c8\": %.2f, "
         "\"w16_c4\": %.2f, \"w16_c8\": %.2f, \"w32_c4\": %.2f}",
         dmma_tflops<1>(d, sms, 4), dmma_tflops<4>(d, sms, 4), dmma_tflops<8>(d, sms, 4), dmma_tflops<4>(d, sms, 8),
         dmma_tflops<8>(d, sms, 8), dmma_tflops<4>(d, sms, 16), dmma_tflops<8>(d, sms, 16), dmma_tflops<4>(d, sms, 32));
  const size_t bytes = size_t(4) << 30;
  double2 *a, *b;
  CK(cudaMalloc(&a, bytes));
  CK(cudaMalloc(&b, bytes));
  CK(cudaMemset(a, 0, bytes));
  CK(cudaMemset(b, 0, bytes));
  const size_t n = bytes / sizeof(double2);
  const int grid = sms * 16;
  float ms = time_ms([&]() { copy_kernel<<<grid, 256>>>(a, b, n); }, 5);
  printf(", \"copy_gbs\": %.1f", 2.0 * bytes / (ms * 1e-3) / 1e9);
  ms = time_ms([&]() { write_kernel<<<grid, 256>>>(b, n, 1.0); }, 5);
  printf(", \"write_gbs\": %.1f", double(bytes) / (ms * 1e-3) / 1e9);
  ms = time_ms([&]() { read_kernel<<<grid, 256>>>(a, n, d); }, 5);
  printf(", \"read_gbs\": %.1f", double(bytes) / (ms * 1e-3) / 1e9);
  ms = time_ms([&]() { CK(cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice)); }, 5);
  printf(", \"memcpy_d2d_gbs\": %.1f", 2.0 * bytes / (ms * 1e-3) / 1e9);
  // host <-> device over PCIe with pinned memory
  void* h;
  const size_t hb = size_t(1) << 30;
  CK(cudaMallocHost(&h, hb));
  ms = time_ms([&]() { CK(cudaMemcpyAsync(a, h, hb, cudaMemcpyHostToDevice)); }, 3);
  printf(", \"h2d_pinned_gbs\": %.1f", double(hb) / (ms * 1e-3) / 1e9);
  ms = time_ms([&]() { CK(cudaMemcpyAsync(h, a, hb, cudaMemcpyDeviceToHost)); }, 3);
  printf(", \"d2h_pinned_gbs\": %.1f", double(hb) / (ms * 1e-3) / 1e9);
  printf("}\n");
  return 0;
}

// What does an in-place update of a strided part of every cell cost on B200?  (T_apply / permute)
// Cells of CS bytes, the movable part is bytes [OFF, OFF + LEN) of each; one thread per 16-byte chunk of the movable part.
//   read       sum the movable part (no stores)
//   write      store the movable part (no loads)
//   inplace    load, modify, store to the same place                       (what T_apply / permute do)
//   outplace   load from one array, store to the same offsets of another   (a copy with the same strides)
//   whole      load and store WHOLE cells in place (contiguous stream)
//   nvcc -arch=sm_100a -O3 -o build/bin/inplace_pattern tools/inplace_pattern.cu && ./build/bin/inplace_pattern
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int ST = 0>
__global__ void k(const int4* __restrict__ src, int4* __restrict__ dst, size_t nc, int cs16, int off16, int len16, int* sink)
{
  const size_t total = nc * size_t(len16), stride = size_t(gridDim.x) * blockDim.x;
  int acc = 0;
  for (size_t q = size_t(blockIdx.x) * blockDim.x + threadIdx.x; q < total; q += stride)
  {
    const size_t c = q / len16;
    const size_t i = c * cs16 + off16 + (q - c * len16);
    int4 v = make_int4(int(q), 1, 2, 3);
    if (MODE != 1)
      v = src[i];
    if (MODE == 0)
      acc += v.x + v.y + v.z + v.w;
    else
    {
      v.x += 1;
      if (ST == 1)
        __stcs(dst + i, v); // evict-first
      else if (ST == 2)
        __stwt(dst + i, v); // write-through
      else if (ST == 3)
        __stcg(dst + i, v); // cache at L2 only
      else
        dst[i] = v;
    }
  }
  if (MODE == 0 and acc == 0x7fffffff)
    *sink = acc;
}

template <int MODE, int ST = 0>
float run(const int4* src, int4* dst, size_t nc, int cs16, int off16, int len16, int* sink)
{
  cudaEvent_t a, b;
  cudaEventCreate(&a);
  cudaEventCreate(&b);
  for (int i = 0; i < 3; ++i)
    k<MODE, ST><<<148 * 8, 256>>>(src, dst, nc, cs16, off16, len16, sink);
  cudaEventRecord(a);
  for (int i = 0; i < 10; ++i)
    k<MODE, ST><<<148 * 8, 256>>>(src, dst, nc, cs16, off16, len16, sink);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms / 10;
}

int main()
{
  struct Case
  {
    const char* name;
    size_t nc;
    int cs, off, len;
  } cases[] = {{"RT4 f64 block 1 (560 B cells, 320 B movable)", 50000000, 560, 0, 320},
               {"P5 i32 (224 B cells, 192 B movable)", 50000000, 224, 16, 192},
               {"RT4 f64 block 3 (1680 B cells, 960 B movable)", 20000000, 1680, 0, 960}};
  printf("[\n");
  for (int ci = 0; ci < 3; ++ci)
  {
    const Case& c = cases[ci];
    int4 *a, *b;
    int* sink;
    cudaMalloc(&a, c.nc * c.cs);
    cudaMalloc(&b, c.nc * c.cs);
    cudaMalloc(&sink, 4);
    cudaMemset(a, 1, c.nc * c.cs);
    cudaMemset(b, 1, c.nc * c.cs);
    const int cs16 = c.cs / 16, off16 = c.off / 16, len16 = c.len / 16;
    const float rd = run<0>(a, a, c.nc, cs16, off16, len16, sink), wr = run<1>(a, a, c.nc, cs16, off16, len16, sink),
                in = run<2>(a, a, c.nc, cs16, off16, len16, sink), out = run<2>(a, b, c.nc, cs16, off16, len16, sink),
                whole = run<2>(a, a, c.nc, cs16, 0, cs16, sink), in_cs = run<2, 1>(a, a, c.nc, cs16, off16, len16, sink),
                in_wt = run<2, 2>(a, a, c.nc, cs16, off16, len16, sink), in_cg = run<2, 3>(a, a, c.nc, cs16, off16, len16, sink),
                wr_cs = run<1, 1>(a, a, c.nc, cs16, off16, len16, sink), wr_wt = run<1, 2>(a, a, c.nc, cs16, off16, len16, sink);
    const double gb = double(c.nc) * c.len / 1e9;
    printf(" {\"case\": \"%s\", \"cells\": %zu, \"read_ms\": %.3f, \"write_ms\": %.3f, \"inplace_ms\": %.3f, \"outplace_ms\": %.3f, "
           "\"whole_cells_inplace_ms\": %.3f, \"inplace_stcs_ms\": %.3f, \"inplace_stwt_ms\": %.3f, \"inplace_stcg_ms\": %.3f, "
           "\"write_stcs_ms\": %.3f, \"write_stwt_ms\": %.3f, \"inplace_GBps_on_movable_bytes\": %.0f, \"cuda\": \"%s\"}%s\n",
           c.name, c.nc, rd, wr, in, out, whole, in_cs, in_wt, in_cg, wr_cs, wr_wt, 2 * gb / (in * 1e-3),
           cudaGetErrorString(cudaGetLastError()), ci < 2 ? "," : "");
    cudaFree(a);
    cudaFree(b);
    cudaFree(sink);
  }
  printf("]\n");
  return 0;
}

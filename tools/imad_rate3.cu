// Row-scanning digit product in isolation: ND x ND digit products into sliding
// 64-bit lanes (exactly the access pattern of mpfw::mul29_rows), operands in
// registers, no normalisation.  Tells whether the lane pattern itself sustains
// the IMAD.WIDE rate.  Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/imad_rate3 tools/imad_rate3.cu
#include <cstdint>
#include <cstdio>

template <int ND, int C0, int MINB> __global__ void __launch_bounds__(256, MINB) rows(uint32_t *out, int iters, uint32_t seed)
{
  uint32_t b[ND];
#pragma unroll
  for(int j = 0; j < ND; ++j)
    b[j] = (threadIdx.x * 2654435761u + seed * (j + 3)) & 0x1FFFFFFFu;
  constexpr int NLN = 2 * ND - 1 - C0;
  uint64_t lane[NLN];
#pragma unroll
  for(int c = 0; c < NLN; ++c)
    lane[c] = 0;
  uint32_t x = blockIdx.x * 40503u + seed;
  for(int it = 0; it < iters; ++it)
    {
#pragma unroll
      for(int i = 0; i < ND; ++i)
        {
          x = x * 1664525u + 1013904223u;
          const uint32_t ai = x & 0x1FFFFFFFu;
#pragma unroll
          for(int j = 0; j < ND; ++j)
            if(i + j >= C0)
              lane[i + j - C0] += (uint64_t)ai * b[j];
        }
#pragma unroll
      for(int c = 0; c < NLN; ++c)
        lane[c] &= 0x00FFFFFFFFFFFFFFull; // keep lanes from overflowing (1 LOP3 per lane)
    }
  uint32_t s = 0;
#pragma unroll
  for(int c = 0; c < NLN; ++c)
    s += (uint32_t)lane[c] + (uint32_t)(lane[c] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ND, int C0, int MINB> void run(const char *name, int grid)
{
  uint32_t *d;
  const int iters = 400;
  cudaMalloc(&d, grid * 256 * 4);
  rows<ND, C0, MINB><<<grid, 256>>>(d, 4, 3);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rows<ND, C0, MINB><<<grid, 256>>>(d, iters, 3);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  int products = 0;
  for(int i = 0; i < ND; ++i)
    for(int j = 0; j < ND; ++j)
      products += (i + j >= C0);
  const double ops = (double)grid * 256 * iters * products;
  printf("{\"bench\": \"%s\", \"ND\": %d, \"C0\": %d, \"grid\": %d, \"products\": %d, \"ms\": %.3f, "
         "\"products_per_clk_per_sm_at_1.965GHz\": %.2f, \"err\": \"%s\"}\n",
         name, ND, C0, grid, products, ms, ops / (ms * 1e-3) / (148 * 1.965e9), cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main()
{
  run<29, 22, 2>("rows29 short 2cta", 296);
  run<29, 22, 1>("rows29 short 1cta", 148);
  run<8, 0, 2>("rows8 full 2cta", 296);
  run<16, 0, 2>("rows16 full 2cta", 296);
  run<16, 10, 2>("rows16 short 2cta", 296);
  return 0;
}

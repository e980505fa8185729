// Definitive IMAD pipe rates on sm_100a with operands that change every row, so
// that ptxas cannot hoist the products out of the loop (the first-round rate
// kernels used loop-invariant operands and measured the 64-bit adds instead).
//   wide  : lane[k] (64-bit) += a * b[k]      IMAD.WIDE.U32 Rd, Ra, Rb, Rc
//   lo    : r[k]    (32-bit) += a * b[k]      IMAD
//   hi    : r[k]    (32-bit) += hi(a * b[k])  IMAD.HI.U32
//   widez : p = a * b[k] (64-bit, no accumulate), r[k] ^= lo ^ hi
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/imad_rate4 tools/imad_rate4.cu
#include <cstdint>
#include <cstdio>

template <int MODE, int NLN> __global__ void __launch_bounds__(256, 2) rate(uint32_t *out, int iters, uint32_t seed)
{
  uint32_t b[NLN];
  uint64_t lane[NLN];
  uint32_t r[NLN];
#pragma unroll
  for(int k = 0; k < NLN; ++k)
    {
      b[k] = threadIdx.x * 2654435761u + seed * (k + 3);
      lane[k] = k;
      r[k] = k;
    }
  uint32_t x = blockIdx.x * 40503u + seed;
  for(int it = 0; it < iters; ++it)
    {
#pragma unroll
      for(int i = 0; i < 8; ++i)
        {
          x = x * 1664525u + 1013904223u;
#pragma unroll
          for(int k = 0; k < NLN; ++k)
            {
              if(MODE == 0)
                lane[k] += (uint64_t)x * b[k];
              else if(MODE == 1)
                r[k] += x * b[k];
              else if(MODE == 2)
                r[k] += __umulhi(x, b[k]);
              else
                {
                  const uint64_t p = (uint64_t)x * b[k];
                  r[k] ^= (uint32_t)p ^ (uint32_t)(p >> 32);
                }
            }
        }
    }
  uint32_t s = 0;
#pragma unroll
  for(int k = 0; k < NLN; ++k)
    s += (uint32_t)lane[k] + (uint32_t)(lane[k] >> 32) + r[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NLN> void run(const char *name)
{
  uint32_t *d;
  const int grid = 296, iters = 2000;
  cudaMalloc(&d, grid * 256 * 4);
  rate<MODE, NLN><<<grid, 256>>>(d, 4, 3);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate<MODE, NLN><<<grid, 256>>>(d, iters, 3);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)grid * 256 * iters * 8 * NLN;
  printf("{\"bench\": \"%s\", \"lanes\": %d, \"ms\": %.3f, \"ops_per_clk_per_sm_at_1.965GHz\": %.2f, \"err\": \"%s\"}\n", name,
         NLN, ms, ops / (ms * 1e-3) / (148 * 1.965e9), cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main()
{
  run<0, 16>("IMAD.WIDE.U32 64-bit accumulate, a varies per row");
  run<0, 8>("IMAD.WIDE.U32 64-bit accumulate, a varies per row");
  run<1, 16>("IMAD (lo) 32-bit accumulate");
  run<2, 16>("IMAD.HI.U32 32-bit accumulate");
  run<3, 16>("IMAD.WIDE.U32 no accumulate + 2 LOP3");
  return 0;
}

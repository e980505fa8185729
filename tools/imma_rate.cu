// Issue rate of the warp-level integer MMA (mma.sync.m16n8k32 u8 x u8 -> s32, SASS IMMA) on sm_100a:
// the candidate for the exact residue syrk (Q' mod p as byte-slice products).  Prints one JSON line.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/imma_rate tools/imma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void mma_u8(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int ACC> __global__ void __launch_bounds__(256) rate_kernel(int iters, int *out)
{
  int c[ACC][4];
  uint32_t a[4], b[2];
  for(int i = 0; i < 4; ++i)
    a[i] = threadIdx.x * 0x01010101u + i;
  b[0] = threadIdx.x + 7;
  b[1] = threadIdx.x * 3;
  for(int q = 0; q < ACC; ++q)
    for(int i = 0; i < 4; ++i)
      c[q][i] = 0;
  for(int it = 0; it < iters; ++it)
    {
#pragma unroll
      for(int q = 0; q < ACC; ++q)
        mma_u8(c[q], a, b);
      a[0] += 1; // operands change between iterations
    }
  int s = 0;
  for(int q = 0; q < ACC; ++q)
    for(int i = 0; i < 4; ++i)
      s += c[q][i];
  if(s == 123456789)
    out[0] = s;
}
template <int ACC> double run(int ctas_per_sm, int iters, int *d)
{
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  rate_kernel<ACC><<<sms * ctas_per_sm, 256>>>(iters / 8, d);
  cudaEventRecord(e0);
  rate_kernel<ACC><<<sms * ctas_per_sm, 256>>>(iters, d);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  const double macs = (double)sms * ctas_per_sm * 8 /*warps*/ * (double)iters * ACC * 16 * 8 * 32;
  return 2 * macs / (ms * 1e-3) / 1e12; // TOPS
}
int main()
{
  int *d;
  cudaMalloc(&d, 64);
  const double t1 = run<4>(1, 1 << 16, d), t2 = run<8>(2, 1 << 16, d), t3 = run<8>(4, 1 << 15, d), t4 = run<16>(2, 1 << 15, d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("{\"what\": \"mma.sync.m16n8k32 u8.u8.s32 issue rate, TOPS (2 ops per MAC)\", \"acc4_1cta\": %.1f, \"acc8_2cta\": %.1f, "
         "\"acc8_4cta\": %.1f, \"acc16_2cta\": %.1f, \"cuda\": \"%s\"}\n",
         t1, t2, t3, t4, cudaGetErrorString(e));
  return 0;
}

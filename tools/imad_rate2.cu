// Which IMAD.WIDE forms sustain the measured 63.6 lanes/clk/SM?  Independent
// streams with varying operand patterns, with and without the carry predicate.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o build/imad_rate2 tools/imad_rate2.cu
#include <cstdint>
#include <cstdio>
#include <vector>

template <int MODE> __global__ void __launch_bounds__(256, 2) rate(uint32_t *out, int iters, uint32_t seed)
{
  uint32_t a[8], b[8], c[16];
#pragma unroll
  for(int k = 0; k < 8; ++k)
    {
      a[k] = threadIdx.x * 2654435761u + seed * (k + 1);
      b[k] = blockIdx.x * 40503u + seed * (k + 3) + 7u;
    }
#pragma unroll
  for(int k = 0; k < 16; ++k)
    c[k] = seed + k;
  for(int it = 0; it < iters; ++it)
    {
      if(MODE == 0) // 8 independent 64-bit lanes, same a, distinct b: one row without carries
        {
#pragma unroll
          for(int u = 0; u < 4; ++u)
#pragma unroll
            for(int k = 0; k < 8; ++k)
              asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;"
                           : "+r"(c[2 * k]), "+r"(c[2 * k + 1])
                           : "r"(a[u]), "r"(b[k]));
        }
      else if(MODE == 1) // one carry chain of 8 lanes per row (IMAD.WIDE.U32.X), 4 rows
        {
#pragma unroll
          for(int u = 0; u < 4; ++u)
            {
              asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                           : "+r"(c[0]), "+r"(c[1])
                           : "r"(a[u]), "r"(b[0]));
#pragma unroll
              for(int k = 1; k < 8; ++k)
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                             : "+r"(c[2 * k]), "+r"(c[2 * k + 1])
                             : "r"(a[u]), "r"(b[k]));
            }
        }
      else if(MODE == 2) // mad.wide.u32 64-bit accumulate, same a, distinct b
        {
          uint64_t *w = reinterpret_cast<uint64_t *>(c);
#pragma unroll
          for(int u = 0; u < 4; ++u)
#pragma unroll
            for(int k = 0; k < 8; ++k)
              asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(a[u]), "r"(b[k]));
        }
      else if(MODE == 3) // mad.wide.u32, a and b both vary per instruction
        {
          uint64_t *w = reinterpret_cast<uint64_t *>(c);
#pragma unroll
          for(int u = 0; u < 4; ++u)
#pragma unroll
            for(int k = 0; k < 8; ++k)
              asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(a[(u + k) & 7]), "r"(b[k]));
        }
      else if(MODE == 4) // mul.wide (no accumulate) + separate 64-bit add
        {
          uint64_t *w = reinterpret_cast<uint64_t *>(c);
#pragma unroll
          for(int u = 0; u < 4; ++u)
#pragma unroll
            for(int k = 0; k < 8; ++k)
              {
                uint64_t p;
                asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(a[u]), "r"(b[k]));
                asm volatile("add.u64 %0, %0, %1;" : "+l"(w[k]) : "l"(p));
              }
        }
      else if(MODE == 5) // mac3: IMAD.WIDE with carry-out + IADD3.X collector, one column
        {
#pragma unroll
          for(int u = 0; u < 4; ++u)
#pragma unroll
            for(int k = 0; k < 8; ++k)
              asm volatile("mad.lo.cc.u32 %0, %3, %4, %0;\n\tmadc.hi.cc.u32 %1, %3, %4, %1;\n\taddc.u32 %2, %2, 0;"
                           : "+r"(c[0]), "+r"(c[1]), "+r"(c[2])
                           : "r"(a[(u + k) & 7]), "r"(b[k]));
        }
      else if(MODE == 6) // 32-bit IMAD lo only, distinct b
        {
#pragma unroll
          for(int u = 0; u < 4; ++u)
#pragma unroll
            for(int k = 0; k < 8; ++k)
              asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(c[k]) : "r"(a[u]), "r"(b[k]));
        }
      else if(MODE == 7) // mad.hi.u32
        {
#pragma unroll
          for(int u = 0; u < 4; ++u)
#pragma unroll
            for(int k = 0; k < 8; ++k)
              asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(c[k]) : "r"(a[u]), "r"(b[k]));
        }
      else if(MODE == 8) // two interleaved carry chains of 8 (even/odd lanes of one row)
        {
#pragma unroll
          for(int u = 0; u < 4; u += 2)
            {
              uint32_t p0, p1;
              asm volatile("{\n\t.reg .u32 x;\n\t"
                           "mad.lo.cc.u32 %0, %16, %18, %0;\n\tmadc.hi.cc.u32 %1, %16, %18, %1;\n\t"
                           "madc.lo.cc.u32 %2, %16, %19, %2;\n\tmadc.hi.cc.u32 %3, %16, %19, %3;\n\t"
                           "madc.lo.cc.u32 %4, %16, %20, %4;\n\tmadc.hi.cc.u32 %5, %16, %20, %5;\n\t"
                           "madc.lo.cc.u32 %6, %16, %21, %6;\n\tmadc.hi.cc.u32 %7, %16, %21, %7;\n\t"
                           "madc.lo.cc.u32 %8, %17, %18, %8;\n\tmadc.hi.cc.u32 %9, %17, %18, %9;\n\t"
                           "madc.lo.cc.u32 %10, %17, %19, %10;\n\tmadc.hi.cc.u32 %11, %17, %19, %11;\n\t"
                           "madc.lo.cc.u32 %12, %17, %20, %12;\n\tmadc.hi.cc.u32 %13, %17, %20, %13;\n\t"
                           "madc.lo.cc.u32 %14, %17, %21, %14;\n\tmadc.hi.u32 %15, %17, %21, %15;\n\t}"
                           : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]), "+r"(c[4]), "+r"(c[5]), "+r"(c[6]),
                             "+r"(c[7]), "+r"(c[8]), "+r"(c[9]), "+r"(c[10]), "+r"(c[11]), "+r"(c[12]),
                             "+r"(c[13]), "+r"(c[14]), "+r"(c[15])
                           : "r"(a[u]), "r"(a[u + 1]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
              (void)p0;
              (void)p1;
            }
#pragma unroll
          for(int u = 0; u < 2; ++u)
            {
            }
        }
    }
  uint32_t s = 0;
#pragma unroll
  for(int k = 0; k < 16; ++k)
    s += c[k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, int per_iter)
{
  uint32_t *d;
  const int grid = 296, iters = 4000;
  cudaMalloc(&d, grid * 256 * 4);
  rate<MODE><<<grid, 256>>>(d, 10, 3);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  rate<MODE><<<grid, 256>>>(d, iters, 3);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double ops = (double)grid * 256 * iters * per_iter;
  printf("{\"bench\": \"%s\", \"ms\": %.3f, \"products_per_clk_per_sm_at_1.965GHz\": %.2f, \"err\": \"%s\"}\n", name, ms,
         ops / (ms * 1e-3) / (148 * 1.965e9), cudaGetErrorString(cudaGetLastError()));
  cudaFree(d);
}
int main()
{
  run<0>("wide_pairs_carry_within_pair (lo.cc + madc.hi), same a distinct b", 32);
  run<1>("row carry chain of 8 lanes (IMAD.WIDE.U32.X)", 32);
  run<2>("mad.wide.u32 64-bit accumulate, same a distinct b", 32);
  run<3>("mad.wide.u32, a and b vary", 32);
  run<4>("mul.wide.u32 + add.u64", 32);
  run<5>("mac3 column (IMAD.WIDE P-out + IADD3.X)", 32);
  run<6>("mad.lo.u32 distinct b", 32);
  run<7>("mad.hi.u32 distinct b", 32);
  run<8>("chain of 8 lanes as one asm block (2 rows)", 16);
  return 0;
}

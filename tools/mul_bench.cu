// Microbenchmark of the short-product inner loop of mpfw::mac (768-bit class):
//   V1  product scanning by columns, (t2:t1:t0) += a_i b_j   (mpfw::mul_columns)
//   V2  operand scanning by rows, even/odd 64-bit lanes, IMAD.WIDE.U32.X carry chains
// Reports products (32x32->64) per clock per SM against the 63.6 measured IMAD.WIDE peak.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o build/mul_bench tools/mul_bench.cu
#include "../sdpb_b200/csrc/mpfw.h"
#include <cstdio>
#include <vector>

template <int W, int C0>
__device__ __forceinline__ void mul_rows(uint32_t (&p)[2 * W - C0], const uint32_t *a, const uint32_t (&b)[W])
{
  mpfw::mul_rows_short<W, C0>(p, a, b);
}

template <int W, int C0, int V, int MINB>
__global__ void __launch_bounds__(256, MINB) k(const uint32_t *A, uint32_t *O, int K)
{
  extern __shared__ uint32_t sm[];
  for(int i = threadIdx.x; i < 64 * 36; i += 256)
    sm[i] = A[i];
  __syncthreads();
  constexpr int NO = 2 * W - C0;
  uint32_t acc[NO];
#pragma unroll
  for(int q = 0; q < NO; ++q)
    acc[q] = 0;
  const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
  for(int kk = 0; kk < K; ++kk)
    {
      const uint32_t *pa = sm + ((kk & 1) * 32 + ti) * 36, *pb = sm + ((kk & 1) * 32 + 16 + tj) * 36;
      uint32_t p[NO];
      if(V == 1)
        {
          uint32_t a[W], b[W];
#pragma unroll
          for(int q = 0; q < W; ++q)
            {
              a[q] = pa[q + 2];
              b[q] = pb[q + 2];
            }
          mpfw::mul_columns<W, C0, C0>(p, a, b);
        }
      else if(V == 2)
        {
          uint32_t b[W];
#pragma unroll
          for(int q = 0; q < W; ++q)
            b[q] = pb[q + 2];
          mul_rows<W, C0>(p, pa + 2, b);
        }
      else
        {
          // radix-2^29 carry-free lanes (mpfw::mul29_words); words GW.. = C0+1.. of the same product
          constexpr int NL = W / 2 + 1;
          typedef mpfw::Mul29Geom<NL> G;
          uint32_t b[W];
#pragma unroll
          for(int q = 0; q < W; ++q)
            b[q] = pb[q + 2];
          uint32_t bd[G::ND];
          mpfw::digits29<G::W, G::ND>(bd, b);
          uint32_t out[G::NOUT];
          mpfw::mul29_words<NL>(out, pa + 2, bd);
          p[0] = 0;
#pragma unroll
          for(int q = 0; q < G::NOUT; ++q)
            p[q + 1] = out[q];
        }
#pragma unroll
      for(int q = 0; q < NO; ++q)
        acc[q] ^= p[q];
    }
#pragma unroll
  for(int q = 0; q < NO; ++q)
    O[(size_t)(blockIdx.x * 256 + threadIdx.x) * NO + q] = acc[q];
}

template <int W, int C0, int V, int MINB> void run(const char *name, int grid, int K, std::vector<uint32_t> *ref)
{
  constexpr int NO = 2 * W - C0;
  std::vector<uint32_t> h(64 * 36);
  uint64_t s = 88172645463325252ull;
  for(auto &x : h)
    {
      s ^= s << 13;
      s ^= s >> 7;
      s ^= s << 17;
      x = (uint32_t)s;
    }
  uint32_t *dA, *dO;
  cudaMalloc(&dA, h.size() * 4);
  cudaMalloc(&dO, (size_t)grid * 256 * NO * 4);
  cudaMemcpy(dA, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  k<W, C0, V, MINB><<<grid, 256, 64 * 36 * 4>>>(dA, dO, 3);
  std::vector<uint32_t> out(256 * NO);
  cudaMemcpy(out.data(), dO, out.size() * 4, cudaMemcpyDeviceToHost);
  bool same = true;
  if(ref->empty())
    *ref = out;
  else
    same = (*ref == out);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<W, C0, V, MINB><<<grid, 256, 64 * 36 * 4>>>(dA, dO, K);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  int products = 0;
  for(int i = 0; i < W; ++i)
    for(int j = 0; j < W; ++j)
      products += (i + j >= C0);
  const double ops = (double)grid * 256 * K;
  printf("{\"bench\": \"%s\", \"W\": %d, \"C0\": %d, \"grid\": %d, \"minb\": %d, \"K\": %d, \"ms\": %.3f, \"muls_per_s\": %.4e, "
         "\"products_per_clk_per_sm_at_1.965GHz\": %.2f, \"matches_V1\": %s, \"err\": \"%s\"}\n",
         name, W, C0, grid, MINB, K, ms, ops / (ms * 1e-3), ops * products / (ms * 1e-3) / (148 * 1.965e9),
         same ? "true" : "false", cudaGetErrorString(cudaGetLastError()));
  cudaFree(dA);
  cudaFree(dO);
}

int main()
{
  std::vector<uint32_t> ref;
  run<26, 20, 1, 2>("mul_V1_columns", 296, 400, &ref);
  run<26, 20, 2, 2>("mul_V2_rows_evenodd", 296, 400, &ref);
  run<26, 20, 2, 1>("mul_V2_rows_evenodd_1cta", 148, 400, &ref);
  std::vector<uint32_t> ref29;
  run<26, 20, 3, 2>("mul_V3_radix29_lanes", 296, 400, &ref29);
  run<26, 20, 3, 1>("mul_V3_radix29_lanes_1cta", 148, 400, &ref29);
  std::vector<uint32_t> ref2;
  run<10, 4, 1, 2>("mul_V1_columns", 296, 2000, &ref2);
  run<10, 4, 2, 2>("mul_V2_rows_evenodd", 296, 2000, &ref2);
  std::vector<uint32_t> ref3;

  return 0;
}

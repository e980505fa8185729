// Microbenchmark of the mpf multiply-accumulate / division / sqrt / reciprocal
// device routines (mpfw.h): latency of one warp alone and throughput at full
// occupancy, in SM clocks per operation per thread.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o build/mac_bench tools/mac_bench.cu
#include "../sdpb_b200/csrc/tile.cuh"

#include <cstdio>
#include <vector>
using namespace sdpb_b200;

// the single-thread Newton routines of mpfw.h (the library itself now takes pivots with the
// warp-cooperative coop::pivot; these stay here as the baseline they are measured against)
template <int NL> __device__ __noinline__ Reg<NL> sqrt_nl(Reg<NL> a)
{
  Reg<NL> r;
  mpfw::sqrt_fast<NL>(r, a);
  return r;
}
template <int NL> struct RecipWords
{
  uint32_t w[2 * NL + 4];
};
template <int NL> __device__ __noinline__ RecipWords<NL> recip_nl(Reg<NL> a)
{
  RecipWords<NL> r;
  mpfw::reciprocal_fast<NL>(r.w, a);
  return r;
}

template <int NL, int OP>
__global__ void __launch_bounds__(256, 2) bench(const uint32_t *A, uint32_t *O, int K, long long *clk)
{
  typedef TileGeom<NL> G;
  extern __shared__ uint32_t sm[];
  for(int i = threadIdx.x; i < 64 * G::SW; i += blockDim.x)
    sm[i] = A[i];
  __syncthreads();
  const int ti = threadIdx.x & 15, tj = (threadIdx.x >> 4) & 15;
  Reg<NL> acc;
  mpfw::load(acc, A + (threadIdx.x & 63) * G::SW);
  const long long t0 = clock64();
  for(int kk = 0; kk < K; ++kk)
    {
      const uint32_t *a = sm + ((kk & 1) * 32 + ti) * G::SW, *b = sm + ((kk & 1) * 32 + 16 + tj) * G::SW;
      if(OP == 0)
        acc = mac_nl<NL>(acc, a, b, (kk & 3) == 0);
      else if(OP == 1)
        {
          mpfw::mac<NL>(acc, a, b, (kk & 3) == 0);
        }
      else if(OP >= 10) // the tile kernels' k-loop: a CTA-wide barrier every OP-10 operations
        {
          acc = mac_nl<NL>(acc, a, b, (kk & 3) == 0);
          if((kk + 1) % (OP - 10) == 0)
            __syncthreads();
        }
      else if(OP == 2)
        acc = div_nl<NL>(acc, a, O + 8 * (kk & 1)); // reciprocal words: arbitrary (timing only)
      else if(OP == 3)
        {
          acc.sign = 1;
          acc = sqrt_nl<NL>(acc);
          acc.w[3] ^= a[5];
        }
      else if(OP == 4)
        {
          const RecipWords<NL> r = recip_nl<NL>(acc);
          acc.w[3] ^= r.w[7] ^ a[5];
          acc.w[2 * NL - 1] |= 1;
        }
    }
  const long long t1 = clock64();
  if(threadIdx.x == 0)
    clk[blockIdx.x] = t1 - t0;
  mpfw::store(O + 64 + (size_t)(blockIdx.x * blockDim.x + threadIdx.x) * G::EW, acc);
}

template <int NL, int OP> void run(const char *name, int grid, int block, int K)
{
  typedef TileGeom<NL> G;
  std::vector<uint32_t> h(64 * G::SW);
  uint64_t s = 88172645463325252ull;
  for(auto &x : h)
    {
      s ^= s << 13;
      s ^= s >> 7;
      s ^= s << 17;
      x = (uint32_t)s;
    }
  for(int e = 0; e < 64; ++e)
    {
      h[e * G::SW] = (uint32_t)(e % 3) - 1; // exp
      h[e * G::SW + 1] = (e & 1) ? 1u : 0xFFFFFFFFu; // sign
    }
  uint32_t *dA, *dO;
  long long *dclk;
  cudaMalloc(&dA, h.size() * 4);
  cudaMalloc(&dO, (size_t)(64 + (size_t)grid * block * G::EW) * 4);
  cudaMalloc(&dclk, grid * 8);
  cudaMemcpy(dA, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  cudaMemset(dO, 0x5a, 64 * 4);
  const size_t smem = 64 * G::SW * 4;
  bench<NL, OP><<<grid, block, smem>>>(dA, dO, 4, dclk);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaEventRecord(e0);
  bench<NL, OP><<<grid, block, smem>>>(dA, dO, K, dclk);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  std::vector<long long> c(grid);
  cudaMemcpy(c.data(), dclk, grid * 8, cudaMemcpyDeviceToHost);
  double avg = 0;
  for(auto x : c)
    avg += x;
  avg /= grid;
  const double ops = (double)grid * block * K;
  std::vector<uint32_t> ho((size_t)block * G::EW);
  cudaMemcpy(ho.data(), dO + 64, ho.size() * 4, cudaMemcpyDeviceToHost);
  uint64_t sum = 1469598103934665603ull;
  for(auto x : ho)
    sum = (sum ^ x) * 1099511628211ull;
  printf("{\"bench\": \"%s\", \"NL\": %d, \"grid\": %d, \"block\": %d, \"K\": %d, \"ms\": %.3f, "
         "\"clk_per_op_per_warp\": %.0f, \"ns_per_op_chip\": %.4f, \"checksum\": \"%016llx\", \"err\": \"%s\"}\n",
         name, NL, grid, block, K, ms, avg / K, ms * 1e6 / ops, (unsigned long long)sum,
         cudaGetErrorString(cudaGetLastError()));
  cudaFree(dA);
  cudaFree(dO);
  cudaFree(dclk);
}

int main(int argc, char **argv)
{
  const bool quick = argc > 1; // any argument: only the 768-bit multiply-accumulate lines
  run<14, 0>("mac_noinline latency (1 warp)", 1, 32, 200);
  run<14, 0>("mac_noinline 1 CTA x 256", 1, 256, 200);
  run<14, 0>("mac_noinline full chip 148x256", 148, 256, 200);
  run<14, 0>("mac_noinline full chip 296x256", 296, 256, 200);
  run<14, 1>("mac_inline full chip 296x256", 296, 256, 200);
  run<14, 14>("mac + __syncthreads every 4, 296x256", 296, 256, 200);
  run<14, 18>("mac + __syncthreads every 8, 296x256", 296, 256, 200);
  run<14, 26>("mac + __syncthreads every 16, 296x256", 296, 256, 208);
  run<14, 11>("mac + __syncthreads every 1, 296x256", 296, 256, 200);
  if(quick)
    return 0;
#ifndef MAC_BENCH_QUICK
  run<14, 2>("div_recip latency (1 warp)", 1, 32, 100);
  run<14, 2>("div_recip full chip", 296, 256, 100);
  run<14, 3>("sqrt_fast latency (1 warp)", 1, 32, 50);
  run<14, 4>("reciprocal_fast latency (1 warp)", 1, 32, 50);
  run<6, 0>("mac_noinline full chip 296x256", 296, 256, 400);
  run<26, 0>("mac_noinline full chip 296x256", 296, 256, 100);
#endif
  return 0;
}

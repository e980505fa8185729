// Step-0 microbenchmark: what does the sm_100a integer
// multiply pipe deliver, and which exact multi-word multiply-accumulate inner
// loop should the big-integer syrk use?
//
//   rate_*   : independent IMAD.LO / IMAD.HI / IMAD.WIDE streams, ops/clk/SM
//   mac_A<W> : W x W 32-bit words, mad.lo.cc / madc.hi.cc carry chains into a
//              (2W+2)-word accumulator (full ripple)             [2 W^2 IMAD]
//   mac_B<W> : W x W radix-2^28 limbs, product scanning with mad.wide.u32 into
//              one 64-bit column accumulator, 32-bit lazy result words
//                                                                 [W^2 IMAD.WIDE]
// Each mac variant is verified against a host big-integer computation.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo
//        -o build/imad_bench tools/imad_bench.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                 \
  do                                                                          \
    {                                                                         \
      cudaError_t e = (x);                                                    \
      if(e != cudaSuccess)                                                    \
        {                                                                     \
          printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, \
                 __LINE__);                                                   \
          exit(1);                                                            \
        }                                                                     \
    }                                                                         \
  while(0)

// ------------------------------------------------------------- raw pipe rate
template <int MODE> __global__ void rate_kernel(uint32_t *out, int iters)
{
  uint32_t a = threadIdx.x * 2654435761u + 12345u, b = blockIdx.x + 7u;
  uint32_t r[8];
  uint64_t w[8];
#pragma unroll
  for(int k = 0; k < 8; ++k)
    {
      r[k] = a + k;
      w[k] = a * 3ull + k;
    }
  for(int it = 0; it < iters; ++it)
    {
#pragma unroll
      for(int u = 0; u < 8; ++u)
        {
#pragma unroll
          for(int k = 0; k < 8; ++k)
            {
              if(MODE == 0)
                asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(r[k]) : "r"(a), "r"(b));
              else if(MODE == 1)
                asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(r[k]) : "r"(a), "r"(b));
              else if(MODE == 2)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[k]) : "r"(a), "r"(b));
              else
                asm volatile("add.u32 %0, %0, %1;" : "+r"(r[k]) : "r"(b));
            }
        }
    }
  uint32_t s = 0;
#pragma unroll
  for(int k = 0; k < 8; ++k)
    s += r[k] + (uint32_t)w[k] + (uint32_t)(w[k] >> 32);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// -------------------------------------------------------- variant A: chains
// acc[0..2W+1] += a[0..W-1] * b[0..W-1]
template <int W>
__device__ __forceinline__ void mac_chain(uint32_t (&acc)[2 * W + 2],
                                          const uint32_t (&a)[W],
                                          const uint32_t (&b)[W])
{
#pragma unroll
  for(int i = 0; i < W; ++i)
    {
      // low halves: acc[i+j] += lo(a_i b_j)
      asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(a[i]), "r"(b[0]));
#pragma unroll
      for(int j = 1; j < W; ++j)
        asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(acc[i + j]) : "r"(a[i]), "r"(b[j]));
#pragma unroll
      for(int k = i + W; k < 2 * W + 1; ++k)
        asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(acc[k]));
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[2 * W + 1]));
      // high halves: acc[i+j+1] += hi(a_i b_j)
      asm volatile("mad.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[i + 1]) : "r"(a[i]), "r"(b[0]));
#pragma unroll
      for(int j = 1; j < W; ++j)
        asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(acc[i + j + 1]) : "r"(a[i]), "r"(b[j]));
#pragma unroll
      for(int k = i + W + 1; k < 2 * W + 1; ++k)
        asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(acc[k]));
      asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[2 * W + 1]));
    }
}

// operands live in shared memory as [row][col][W]; every thread owns one
// output (ti, tj) of a 16x16 tile and loops over `rows` rows.
template <int W>
__global__ void __launch_bounds__(256)
mac_A_kernel(const uint32_t *__restrict__ A, uint32_t *__restrict__ out,
             int rows, int reps)
{
  extern __shared__ uint32_t sm[];
  const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
  for(int k = threadIdx.x; k < rows * 32 * W; k += 256)
    sm[k] = A[k];
  __syncthreads();
  uint32_t acc[2 * W + 2];
#pragma unroll
  for(int k = 0; k < 2 * W + 2; ++k)
    acc[k] = 0;
  for(int rep = 0; rep < reps; ++rep)
    for(int r = 0; r < rows; ++r)
      {
        uint32_t a[W], b[W];
        const uint32_t *pa = sm + (r * 32 + ti) * W;
        const uint32_t *pb = sm + (r * 32 + 16 + tj) * W;
#pragma unroll
        for(int k = 0; k < W; ++k)
          {
            a[k] = pa[k];
            b[k] = pb[k];
          }
        mac_chain<W>(acc, a, b);
      }
  uint32_t *o = out + (size_t)(blockIdx.x * 256 + threadIdx.x) * (2 * W + 2);
#pragma unroll
  for(int k = 0; k < 2 * W + 2; ++k)
    o[k] = acc[k];
}

// ----------------------------------------------- variant B: radix 2^28 wide
// limbs < 2^28.  res[c] (32-bit, lazily normalised) += column sums.
template <int W>
__device__ __forceinline__ void mac_wide28(uint32_t (&res)[2 * W],
                                           const uint32_t (&a)[W],
                                           const uint32_t (&b)[W])
{
  uint64_t carry = 0;
#pragma unroll
  for(int c = 0; c < 2 * W - 1; ++c)
    {
      uint64_t t = carry;
#pragma unroll
      for(int i = 0; i < W; ++i)
        {
          const int j = c - i;
          if(j >= 0 && j < W)
            asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(t) : "r"(a[i]), "r"(b[j]));
        }
      res[c] += (uint32_t)t & 0x0FFFFFFFu;
      carry = t >> 28;
    }
  res[2 * W - 1] += (uint32_t)carry;
}
template <int W>
__device__ __forceinline__ void normalize28(uint32_t (&res)[2 * W])
{
  uint32_t c = 0;
#pragma unroll
  for(int k = 0; k < 2 * W - 1; ++k)
    {
      const uint32_t v = res[k] + c;
      res[k] = v & 0x0FFFFFFFu;
      c = v >> 28;
    }
  res[2 * W - 1] += c;
}
template <int W>
__global__ void __launch_bounds__(256)
mac_B_kernel(const uint32_t *__restrict__ A, uint32_t *__restrict__ out,
             int rows, int reps)
{
  extern __shared__ uint32_t sm[];
  const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
  for(int k = threadIdx.x; k < rows * 32 * W; k += 256)
    sm[k] = A[k];
  __syncthreads();
  uint32_t res[2 * W];
#pragma unroll
  for(int k = 0; k < 2 * W; ++k)
    res[k] = 0;
  int since = 0;
  for(int rep = 0; rep < reps; ++rep)
    for(int r = 0; r < rows; ++r)
      {
        uint32_t a[W], b[W];
        const uint32_t *pa = sm + (r * 32 + ti) * W;
        const uint32_t *pb = sm + (r * 32 + 16 + tj) * W;
#pragma unroll
        for(int k = 0; k < W; ++k)
          {
            a[k] = pa[k];
            b[k] = pb[k];
          }
        mac_wide28<W>(res, a, b);
        if(++since == 8) // 28-bit digits + 8 additions < 2^32
          {
            normalize28<W>(res);
            since = 0;
          }
      }
  // the top word keeps everything above 28*(2W-1) bits; 32 bits suffice for
  // the benchmark's row count
  normalize28<W>(res);
  uint32_t *o = out + (size_t)(blockIdx.x * 256 + threadIdx.x) * (2 * W);
#pragma unroll
  for(int k = 0; k < 2 * W; ++k)
    o[k] = res[k];
}

// ------------------------------- variant C: radix 2^RB, 64-bit column lanes
// operand scanning: col[i+j] += a_i * b_j with mad.wide.u32; every lane is an
// independent 64-bit accumulator, no carries inside a row pair; lanes are
// normalised every NORM row pairs (RB=28: product < 2^56, W products/lane/row).
template <int W, int RB, int NORM>
__global__ void __launch_bounds__(256)
mac_C_kernel(const uint32_t *__restrict__ A, uint32_t *__restrict__ out,
             int rows, int reps)
{
  extern __shared__ uint32_t sm[];
  const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
  for(int k = threadIdx.x; k < rows * 32 * W; k += 256)
    sm[k] = A[k];
  __syncthreads();
  uint64_t col[2 * W];
#pragma unroll
  for(int k = 0; k < 2 * W; ++k)
    col[k] = 0;
  int since = 0;
  for(int rep = 0; rep < reps; ++rep)
    for(int r = 0; r < rows; ++r)
      {
        uint32_t b[W];
        const uint32_t *pa = sm + (r * 32 + ti) * W;
        const uint32_t *pb = sm + (r * 32 + 16 + tj) * W;
#pragma unroll
        for(int k = 0; k < W; ++k)
          b[k] = pb[k];
#pragma unroll
        for(int i = 0; i < W; ++i)
          {
            const uint32_t ai = pa[i];
#pragma unroll
            for(int j = 0; j < W; ++j)
              asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(col[i + j]) : "r"(ai), "r"(b[j]));
          }
        if(++since == NORM)
          {
            since = 0;
#pragma unroll
            for(int k = 0; k < 2 * W - 1; ++k)
              {
                col[k + 1] += col[k] >> RB;
                col[k] &= ((1ull << RB) - 1);
              }
          }
      }
#pragma unroll
  for(int k = 0; k < 2 * W - 1; ++k)
    {
      col[k + 1] += col[k] >> RB;
      col[k] &= ((1ull << RB) - 1);
    }
  uint32_t *o = out + (size_t)(blockIdx.x * 256 + threadIdx.x) * (2 * W + 1);
#pragma unroll
  for(int k = 0; k < 2 * W - 1; ++k)
    o[k] = (uint32_t)col[k];
  o[2 * W - 1] = (uint32_t)col[2 * W - 1];
  o[2 * W] = (uint32_t)(col[2 * W - 1] >> 32);
}

// ------------------------------------------------------------ host checking
typedef unsigned __int128 u128;
static void host_mac_words(std::vector<uint64_t> &acc, const uint32_t *a,
                           const uint32_t *b, int W, int radix_bits)
{
  // acc holds base-2^radix digits in 64-bit slots (lazy), normalised by caller
  for(int i = 0; i < W; ++i)
    for(int j = 0; j < W; ++j)
      {
        u128 p = (u128)a[i] * b[j];
        int pos = i + j;
        while(p)
          {
            acc[pos] += (uint64_t)(p & (((u128)1 << radix_bits) - 1));
            p >>= radix_bits;
            pos++;
          }
      }
}
static void host_norm(std::vector<uint64_t> &acc, int radix_bits)
{
  uint64_t c = 0;
  for(size_t k = 0; k < acc.size(); ++k)
    {
      uint64_t v = acc[k] + c;
      acc[k] = v & (((uint64_t)1 << radix_bits) - 1);
      c = v >> radix_bits;
    }
}

template <int W, int RB, int NORM> static void run_mac_C(int sms, float clock_ghz)
{
  const int rows = 32, reps = 40;
  std::vector<uint32_t> h((size_t)rows * 32 * W);
  uint64_t s = 88172645463325252ull;
  for(auto &x : h)
    {
      s ^= s << 13;
      s ^= s >> 7;
      s ^= s << 17;
      x = (uint32_t)(s >> 11) & ((1u << RB) - 1);
    }
  uint32_t *dA, *dOut;
  const int ctas = sms * 2;
  const int accw = 2 * W + 1;
  CK(cudaMalloc(&dA, h.size() * 4));
  CK(cudaMalloc(&dOut, (size_t)ctas * 256 * accw * 4));
  CK(cudaMemcpy(dA, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)rows * 32 * W * 4;
  CK(cudaFuncSetAttribute(mac_C_kernel<W, RB, NORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for(int t = 0; t < 5; ++t)
    {
      CK(cudaEventRecord(e0));
      mac_C_kernel<W, RB, NORM><<<ctas, 256, smem>>>(dA, dOut, rows, reps);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if(t > 0 && ms < best)
        best = ms;
    }
  CK(cudaGetLastError());
  std::vector<uint32_t> got(accw);
  const int tid = 5 * 16 + 3;
  CK(cudaMemcpy(got.data(), dOut + (size_t)tid * accw, accw * 4, cudaMemcpyDeviceToHost));
  std::vector<uint64_t> acc(2 * W + 6, 0);
  for(int rep = 0; rep < reps; ++rep)
    for(int r = 0; r < rows; ++r)
      {
        host_mac_words(acc, &h[(size_t)(r * 32 + 3) * W], &h[(size_t)(r * 32 + 16 + 5) * W], W, RB);
        host_norm(acc, RB);
      }
  bool ok = true;
  for(int k = 0; k < 2 * W - 1; ++k)
    ok &= (acc[k] == got[k]);
  uint64_t top = 0;
  for(int k = (int)acc.size() - 1; k >= 2 * W - 1; --k)
    top = (top << RB) | acc[k];
  ok &= (top == ((uint64_t)got[2 * W - 1] | ((uint64_t)got[2 * W] << 32)));
  const double macs = (double)ctas * 256 * rows * reps;
  const double imads = macs * (double)W * W;
  printf("{\"bench\": \"mac_C_lanes%d\", \"W\": %d, \"bits\": %d, \"norm_every\": %d, \"ms\": %.4f, \"elem_macs_per_s\": %.4e, "
         "\"imad_per_clk_per_sm_at_%.3fGHz\": %.2f, \"verified\": %s}\n",
         RB, W, W * RB, NORM, best, macs / (best * 1e-3), clock_ghz,
         imads / (best * 1e-3) / (clock_ghz * 1e9) / sms, ok ? "true" : "false");
  CK(cudaFree(dA));
  CK(cudaFree(dOut));
}

template <int W, bool WIDE> static void run_mac(int sms, float clock_ghz)
{
  const int rows = 32, reps = WIDE ? 40 : 40;
  const int radix = WIDE ? 28 : 32;
  std::vector<uint32_t> h((size_t)rows * 32 * W);
  uint64_t s = 88172645463325252ull;
  for(auto &x : h)
    {
      s ^= s << 13;
      s ^= s >> 7;
      s ^= s << 17;
      x = (uint32_t)(s >> 11) & (WIDE ? 0x0FFFFFFFu : 0xFFFFFFFFu);
    }
  uint32_t *dA, *dOut;
  const int ctas = sms * 2;
  const int accw = WIDE ? 2 * W : 2 * W + 2;
  CK(cudaMalloc(&dA, h.size() * 4));
  CK(cudaMalloc(&dOut, (size_t)ctas * 256 * accw * 4));
  CK(cudaMemcpy(dA, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)rows * 32 * W * 4;
  if(WIDE)
    CK(cudaFuncSetAttribute(mac_B_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  else
    CK(cudaFuncSetAttribute(mac_A_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for(int t = 0; t < 5; ++t)
    {
      CK(cudaEventRecord(e0));
      if(WIDE)
        mac_B_kernel<W><<<ctas, 256, smem>>>(dA, dOut, rows, reps);
      else
        mac_A_kernel<W><<<ctas, 256, smem>>>(dA, dOut, rows, reps);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if(t > 0 && ms < best)
        best = ms;
    }
  CK(cudaGetLastError());
  // verify thread (ti=3, tj=5) of CTA 0
  std::vector<uint32_t> got(accw);
  const int tid = 5 * 16 + 3;
  CK(cudaMemcpy(got.data(), dOut + (size_t)tid * accw, accw * 4, cudaMemcpyDeviceToHost));
  std::vector<uint64_t> acc(2 * W + 4, 0);
  for(int rep = 0; rep < reps; ++rep)
    for(int r = 0; r < rows; ++r)
      {
        host_mac_words(acc, &h[(size_t)(r * 32 + 3) * W], &h[(size_t)(r * 32 + 16 + 5) * W], W, radix);
        host_norm(acc, radix);
      }
  bool ok = true;
  if(WIDE)
    {
      // device top word holds digits 2W-1 and above
      for(int k = 0; k < 2 * W - 1; ++k)
        ok &= (acc[k] == got[k]);
      uint64_t top = 0;
      for(int k = (int)acc.size() - 1; k >= 2 * W - 1; --k)
        top = (top << 28) | acc[k];
      ok &= (top == got[2 * W - 1]);
    }
  else
    for(int k = 0; k < 2 * W + 2; ++k)
      ok &= (acc[k] == got[k]);
  const double macs = (double)ctas * 256 * rows * reps;
  const double imads = macs * (WIDE ? (double)W * W : 2.0 * W * W);
  printf("{\"bench\": \"mac_%s\", \"W\": %d, \"bits\": %d, \"ms\": %.4f, \"elem_macs_per_s\": %.4e, "
         "\"imad_per_clk_per_sm_at_%.3fGHz\": %.2f, \"verified\": %s}\n",
         WIDE ? "B_wide28" : "A_chain32", W, W * radix, best, macs / (best * 1e-3),
         clock_ghz, imads / (best * 1e-3) / (clock_ghz * 1e9) / sms, ok ? "true" : "false");
  CK(cudaFree(dA));
  CK(cudaFree(dOut));
}

template <int MODE> static void run_rate(const char *name, int sms, float clock_ghz)
{
  uint32_t *d;
  const int ctas = sms * 4, thr = 512, iters = 2000;
  CK(cudaMalloc(&d, (size_t)ctas * thr * 4));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  float best = 1e30f;
  for(int t = 0; t < 4; ++t)
    {
      CK(cudaEventRecord(e0));
      rate_kernel<MODE><<<ctas, thr>>>(d, iters);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if(t > 0 && ms < best)
        best = ms;
    }
  const double ops = (double)ctas * thr * iters * 64.0;
  printf("{\"bench\": \"rate_%s\", \"ms\": %.4f, \"ops_per_s\": %.4e, \"ops_per_clk_per_sm_at_%.3fGHz\": %.2f}\n",
         name, best, ops / (best * 1e-3), clock_ghz, ops / (best * 1e-3) / (clock_ghz * 1e9) / sms);
  CK(cudaFree(d));
}

int main()
{
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0;
  CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  const float ghz = clk_khz * 1e-6f;
  printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz_nominal\": %.3f}\n", p.name, p.multiProcessorCount, ghz);
  const int sms = p.multiProcessorCount;
  run_rate<0>("imad_lo", sms, ghz);
  run_rate<1>("imad_hi", sms, ghz);
  run_rate<2>("imad_wide", sms, ghz);
  run_rate<3>("iadd", sms, ghz);
  run_mac<10, false>(sms, ghz);
  run_mac<25, false>(sms, ghz);
  run_mac<26, false>(sms, ghz);
  run_mac<12, true>(sms, ghz);
  run_mac<28, true>(sms, ghz);
  run_mac_C<12, 28, 4>(sms, ghz);
  run_mac_C<28, 28, 4>(sms, ghz);
  run_mac_C<27, 29, 2>(sms, ghz);
  run_mac_C<18, 28, 4>(sms, ghz);
  return 0;
}

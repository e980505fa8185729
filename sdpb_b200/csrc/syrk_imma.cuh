// The exact integer syrk on the integer tensor path (row a9 of SURVEY §8):
//     Qres[p][i][j] = sum_rows R[p][row][i] R[p][row][j] mod p,   i <= j,
// the reference's own formulation -- residues modulo word-sized primes, one dense contraction per
// prime, CRT (bigint_syrk/Readme.md:27-55; there: fp64 dsyrk) -- with the contraction done by
// warp-level integer MMAs (mma.sync.m16n8k32 u8 x u8 -> s32, SASS IMMA.16832.U8.U8; measured
// 1140 TOPS on this B200, profiles/imma_rate_r02.json) instead of IMAD.WIDE (syrk_mod_kernel).
// This is an integer contraction modulo p, not the multi-limb floating-point kernels the north
// star keeps off the tensor cores; the result is exact, hence identical bit for bit.
//
// A residue r < p < 2^28 is four byte slices r = b0 + 2^8 b1 + 2^16 b2 + 2^24 b3 (b3 < 16), so
//     sum_k r_ki r_kj = sum_{sa, sb} 2^(8 (sa + sb)) sum_k b_sa(k,i) b_sb(k,j):
// 16 byte products per pair, grouped by their 7 weights d = sa + sb into 7 s32 accumulators.  A weight
// collects at most 3 full byte products per row (d = 2: 3 * 255^2 = 195 075), so 8192 rows stay
// below 2^31; every 8192 rows the accumulators are folded into a running residue
// (sum_d acc_d (2^(8d) mod p), below 2^62, one Barrett reduction).
//
// Layout.  An MMA fragment register holds the same byte slice of FOUR consecutive rows k of one
// column, so the residue planes are first re-packed in place (syrk_pack_kernel): the four words
// R[4g .. 4g+3][c] become the four slice words P[4g + s][c] = {b_s(4g,c), b_s(4g+1,c), b_s(4g+2,c),
// b_s(4g+3,c)} -- a 4x4 byte transpose that needs no second buffer.  The syrk then streams 64-row x
// 64-column tiles of P with cp.async into a 4-stage shared-memory ring (rows re-ordered slice-major,
// stride 72 words: conflict-free fragment loads) and every fragment is one LDS.32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdpb_b200
{
constexpr int SI_TILE = 64;     // output tile (columns of R on either side)
constexpr int SI_KS = 64;       // rows of R per pipeline stage
constexpr int SI_STAGES = 4;
constexpr int SI_CS = 72;       // shared row stride, words (== 8 mod 32)
constexpr int SI_FOLD = 128;    // stages between folds: 8192 rows
constexpr size_t SI_SMEM = (size_t)SI_STAGES * 2 * SI_KS * SI_CS * 4;

// in-place 4x4 byte transpose of every group of four rows of every plane (KR rows each, a multiple
// of 4).  Rows from K on are pad: they read as zero whatever they hold -- the slice words of a
// group that straddles K land in them, and the next step's residues only overwrite the rows below K.
template <int ROWS /* = 4 */>
__global__ void __launch_bounds__(256) syrk_pack_kernel(uint32_t *R, long groups, long K, long KR, int NS)
{
  // one thread per (row group, column); consecutive threads on consecutive columns
  const long total = groups * NS, per_plane = KR / 4;
  for(long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
    {
      const long g = e / NS;
      const int c = (int)(e % NS);
      uint32_t *p = R + (g * 4) * NS + c;
      const long row = (g % per_plane) * 4;
      if(row >= K)
        continue; // the group is all pad: zeros, left alone
      const uint32_t w0 = p[0], w1 = row + 1 < K ? p[NS] : 0u, w2 = row + 2 < K ? p[2 * (long)NS] : 0u,
                     w3 = row + 3 < K ? p[3 * (long)NS] : 0u;
      // slice s = byte s of w0..w3
      const uint32_t lo01 = __byte_perm(w0, w1, 0x5140); // b0(w0) b0(w1) b1(w0) b1(w1)
      const uint32_t hi01 = __byte_perm(w0, w1, 0x7362); // b2(w0) b2(w1) b3(w0) b3(w1)
      const uint32_t lo23 = __byte_perm(w2, w3, 0x5140);
      const uint32_t hi23 = __byte_perm(w2, w3, 0x7362);
      p[0] = __byte_perm(lo01, lo23, 0x5410);
      p[NS] = __byte_perm(lo01, lo23, 0x7632);
      p[2 * (long)NS] = __byte_perm(hi01, hi23, 0x5410);
      p[3 * (long)NS] = __byte_perm(hi01, hi23, 0x7632);
    }
}

__device__ __forceinline__ void si_mma(int (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
  asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void si_cp_async16(uint32_t smem_addr, const void *gptr, int src_bytes)
{
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_addr), "l"(gptr), "r"(src_bytes));
}
__device__ __forceinline__ void si_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void si_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// P: the packed residue planes, KR rows (a multiple of SI_KS; rows beyond the data are zero) of NS
// words per prime.  grid: (tile pairs ti <= tj of 64 columns, primes); 256 threads = 8 warps as
// 4 (rows) x 2 (columns), a warp owns 16 x 32 outputs.
template <int STAGES /* = SI_STAGES */>
__global__ void __launch_bounds__(256, 1)
syrk_imma_kernel(const uint32_t *__restrict__ P, long KR, int N, int NS, const uint32_t *__restrict__ primes,
                 const uint64_t *__restrict__ inv64, uint32_t *Qres)
{
  extern __shared__ __align__(16) uint32_t si_smem[];
  int tpair = blockIdx.x, tj = 0;
  while(tpair > tj)
    {
      tpair -= tj + 1;
      ++tj;
    }
  const int ti = tpair; // ti <= tj
  const int pi = blockIdx.y;
  const uint32_t p = primes[pi];
  const uint64_t inv = inv64[pi];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int wi = warp & 3, wj = warp >> 2;
  const uint32_t *plane = P + (size_t)pi * KR * NS;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(si_smem);

  // 2^(8 d) mod p, d = 0 .. 6
  uint32_t wgt[7];
  wgt[0] = 1;
#pragma unroll
  for(int d = 1; d < 7; ++d)
    wgt[d] = (uint32_t)(((uint64_t)wgt[d - 1] << 8) % p);

  // a stage = rows [64 st, 64 st + 64) of the column blocks ti and tj: 2 x 64 rows x 16 chunks of 16 B
  auto issue = [&](long st, int buf) {
#pragma unroll
    for(int q = 0; q < 8; ++q)
      {
        const int e = tid + 256 * q;        // 0 .. 2047
        const int side = e >> 10, r = (e >> 4) & 63, ch = e & 15;
        const int col = (side ? tj : ti) * SI_TILE + 4 * ch;
        const uint32_t *src = plane + (size_t)(st * SI_KS + r) * NS + col;
        // packed row r = 4 kg + s goes to shared row s * 16 + kg
        const int srow = (r & 3) * 16 + (r >> 2);
        const uint32_t dst = smem_base + (uint32_t)(((buf * 2 + side) * SI_KS + srow) * SI_CS + 4 * ch) * 4u;
        const bool ok = col < NS;
        si_cp_async16(dst, ok ? (const void *)src : (const void *)plane, ok ? 16 : 0);
      }
  };

  int acc[7][4][4];
#pragma unroll
  for(int d = 0; d < 7; ++d)
#pragma unroll
    for(int n = 0; n < 4; ++n)
#pragma unroll
      for(int q = 0; q < 4; ++q)
        acc[d][n][q] = 0;
  uint32_t run[4][4];
#pragma unroll
  for(int n = 0; n < 4; ++n)
#pragma unroll
    for(int q = 0; q < 4; ++q)
      run[n][q] = 0;
  auto fold = [&]() {
#pragma unroll
    for(int n = 0; n < 4; ++n)
#pragma unroll
      for(int q = 0; q < 4; ++q)
        {
          uint64_t v = run[n][q];
#pragma unroll
          for(int d = 0; d < 7; ++d)
            {
              v += (uint64_t)(uint32_t)acc[d][n][q] * wgt[d];
              acc[d][n][q] = 0;
            }
          // v < 2^28 + 7 * 2^31 * 2^28 < 2^62
          const uint64_t qq = __umul64hi(v, inv);
          uint32_t r = (uint32_t)v - (uint32_t)qq * p;
          r = r >= p ? r - p : r;
          r = r >= p ? r - p : r;
          run[n][q] = r;
        }
  };

  const long nstages = KR / SI_KS;
#pragma unroll
  for(int s = 0; s < STAGES - 1; ++s)
    {
      if(s < nstages)
        issue(s, s);
      si_commit();
    }
  int since = 0;
  for(long st = 0; st < nstages; ++st)
    {
      si_wait<STAGES - 2>();
      __syncthreads();
      {
        const long nx = st + STAGES - 1;
        if(nx < nstages)
          issue(nx, (int)(nx % STAGES));
        si_commit();
      }
      const int buf = (int)(st % STAGES);
      const uint32_t *Ai = si_smem + (size_t)(buf * 2 + 0) * SI_KS * SI_CS + wi * 16 + g;
      const uint32_t *Bj = si_smem + (size_t)(buf * 2 + 1) * SI_KS * SI_CS + wj * 32 + g;
#pragma unroll
      for(int kk = 0; kk < 2; ++kk)
        {
          uint32_t a[4][4];
#pragma unroll
          for(int s = 0; s < 4; ++s)
            {
              const uint32_t *r0 = Ai + (s * 16 + kk * 8 + t) * SI_CS, *r1 = r0 + 4 * SI_CS;
              a[s][0] = r0[0];
              a[s][1] = r0[8];
              a[s][2] = r1[0];
              a[s][3] = r1[8];
            }
          // two column tiles at a time; the 16 slice pairs in an order that keeps the updates of
          // one accumulator (same weight sa + sb, same tile) at least six instructions apart -- a
          // warp issues in order, and two warps per scheduler do not hide an IMMA's latency
#pragma unroll
          for(int n2 = 0; n2 < 4; n2 += 2)
            {
              uint32_t b[2][4][2];
#pragma unroll
              for(int q = 0; q < 2; ++q)
#pragma unroll
                for(int s = 0; s < 4; ++s)
                  {
                    const uint32_t *r0 = Bj + (s * 16 + kk * 8 + t) * SI_CS + (n2 + q) * 8;
                    b[q][s][0] = r0[0];
                    b[q][s][1] = r0[4 * SI_CS];
                  }
              constexpr int ORDER[16][2] = {{0, 0}, {0, 1}, {0, 2}, {0, 3}, {1, 3}, {2, 3}, {3, 3}, {1, 0},
                                            {1, 1}, {1, 2}, {2, 2}, {3, 2}, {2, 0}, {2, 1}, {3, 1}, {3, 0}};
#pragma unroll
              for(int o = 0; o < 16; ++o)
#pragma unroll
                for(int q = 0; q < 2; ++q)
                  si_mma(acc[ORDER[o][0] + ORDER[o][1]][n2 + q], a[ORDER[o][0]], b[q][ORDER[o][1]][0],
                         b[q][ORDER[o][1]][1]);
            }
        }
      if(++since == SI_FOLD)
        {
          fold();
          since = 0;
        }
    }
  si_wait<0>();
  fold();
  // c0, c1: row g, columns 2t, 2t+1; c2, c3: row g + 8
#pragma unroll
  for(int n = 0; n < 4; ++n)
#pragma unroll
    for(int q = 0; q < 4; ++q)
      {
        const int i = ti * SI_TILE + wi * 16 + g + (q >= 2 ? 8 : 0);
        const int j = tj * SI_TILE + wj * 32 + n * 8 + 2 * t + (q & 1);
        if(i < N && j < N && i <= j)
          Qres[((size_t)pi * N + i) * N + j] = run[n][q];
      }
}
} // namespace sdpb_b200

// sm_100a kernels of the Schur-complement step.  All arithmetic is mpfx
// (GMP-mpf-exact fixed-limb floating point, mpfx.h) except the exact integer
// syrk, which works on residues modulo 28-bit primes on the INT32 multiply
// pipe (IMAD.WIDE) and reconstructs the integer by Garner's CRT.
//
// Canonical operation order: every matrix element receives its
// updates in ascending k, one mpf operation per reference operator, exactly as
// oracle/hotpath_core.hpp spells out.  That makes right-looking, left-looking
// and blocked schedules bit-identical, which is what lets these kernels pick
// the schedule with the most parallelism.
#pragma once
#include "mpfx.h"
#include "tile.cuh"
#include "syrk_imma.cuh"

#include <climits>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sdpb_b200
{
using mpfx::limb_t;
using mpfx::Num;

template <int NL> struct Fmt
{
  static constexpr int ES = (NL + 2) & ~1; // 64-bit words per element
};

struct SchurDesc
{
  const limb_t *AX[2];
  const limb_t *AY[2];
  limb_t *S;
  int m, n;
};
struct BandDesc // P band of one SDP block inside the stacked K x N matrix
{
  limb_t *P; // P_j x N, column-major, ld = rows
  int rows;
  int row0; // first stacked row
  int gidx; // global block index (row of the norm partials; == local index on one GPU)
};

// ------------------------------------------------------------- small helpers
template <int NL>
__device__ __forceinline__ void ld(Num<NL> &x, const limb_t *base, size_t idx)
{
  mpfx::load(x, base + idx * Fmt<NL>::ES);
}
template <int NL>
__device__ __forceinline__ void st(limb_t *base, size_t idx, const Num<NL> &x)
{
  mpfx::store(base + idx * Fmt<NL>::ES, x);
}

// ------------------------------------------------------------ Schur assembly
// compute_schur_complement.cxx:31-124, lower triangle + mirror; one thread per
// element, the eight products in the order the reference spells out.
template <int NL>
__global__ void __launch_bounds__(128, 4) schur_kernel(const SchurDesc *descs)
{
  const SchurDesc d = descs[blockIdx.x];
  const int n = d.n, m = d.m, mn = m * n;
  const int P = n * m * (m + 1) / 2;
  const long total = (long)P * P;
  constexpr int EW = 2 * Fmt<NL>::ES;
  for(long e = (long)blockIdx.y * blockDim.x + threadIdx.x; e < total;
      e += (long)gridDim.y * blockDim.x)
    {
      const int I = (int)(e % P), J = (int)(e / P);
      if(I < J)
        continue;
      // I = (c0(c0+1)/2 + r0) n + row ; J = (c1(c1+1)/2 + r1) n + col
      const int row = I % n, col = J % n;
      int t0 = I / n, t1 = J / n;
      int c0 = 0, c1 = 0;
      while((c0 + 1) * (c0 + 2) / 2 <= t0)
        ++c0;
      while((c1 + 1) * (c1 + 2) / 2 <= t1)
        ++c1;
      const int r0 = t0 - c0 * (c0 + 1) / 2, r1 = t1 - c1 * (c1 + 1) / 2;
      Reg<NL> element;
      mpfw::set_zero(element);
      // ax(p,cb,rb) = AX[p](cb n + row, rb n + col); ay = AY[p](cb n + col, rb n + row)
#define SDPB_AX(cb, rb) ((size_t)((rb) * n + col) * mn + ((cb) * n + row))
#define SDPB_AY(cb, rb) ((size_t)((rb) * n + row) * mn + ((cb) * n + col))
#define SDPB_TERM(p, xa, xb, ya, yb)                                          \
  element = mac_nl<NL>(                                                       \
    element, reinterpret_cast<const uint32_t *>(d.AX[p]) + SDPB_AX(xa, xb) * EW, \
    reinterpret_cast<const uint32_t *>(d.AY[p]) + SDPB_AY(ya, yb) * EW, false);
      for(int p = 0; p < 2; ++p)
        {
          SDPB_TERM(p, c0, r1, c1, r0)
          SDPB_TERM(p, r0, r1, c1, c0)
          SDPB_TERM(p, c0, c1, r1, r0)
          SDPB_TERM(p, r0, c1, r1, c0)
        }
#undef SDPB_TERM
#undef SDPB_AX
#undef SDPB_AY
      mpfw::div4<NL>(element);
      stg_reg<NL>(d.S + ((size_t)J * P + I) * Fmt<NL>::ES, element);
      if(I != J)
        stg_reg<NL>(d.S + ((size_t)I * P + J) * Fmt<NL>::ES, element);
    }
}

// --------------------------------------------------------------- normaliser
// Matrix_Normalizer.cxx:75-139.  part[j*N + c] = sum over the rows of band j
// of v^2 (rows ascending); norms[c] = sqrt(sum_j part) in block order.
template <int NL>
__global__ void __launch_bounds__(64, 8) norm_partial_kernel(const BandDesc *bands, int N, limb_t *part)
{
  const BandDesc b = bands[blockIdx.x];
  for(int c = blockIdx.y * blockDim.x + threadIdx.x; c < N;
      c += gridDim.y * blockDim.x)
    {
      Reg<NL> acc;
      mpfw::set_zero(acc);
      const uint32_t *col = reinterpret_cast<const uint32_t *>(b.P + (size_t)c * b.rows * Fmt<NL>::ES);
      for(int r = 0; r < b.rows; ++r) // acc += v * v: one mpf_mul, one mpf_add (Matrix_Normalizer.cxx:82-88)
        acc = mac_nl<NL>(acc, col + (size_t)r * 2 * Fmt<NL>::ES, col + (size_t)r * 2 * Fmt<NL>::ES, false);
      stg_reg<NL>(part + ((size_t)b.gidx * N + c) * Fmt<NL>::ES, acc);
    }
}
template <int NL> __device__ __noinline__ Reg<NL> add_nl(Reg<NL> acc, Reg<NL> v)
{
  mpfw::add_signed<NL>(acc, v, v.sign);
  return acc;
}
// The canonical sum of per-block rows over the GLOBAL blocks (oracle: ordered_block_sum): groups
// of BLOCK_SUM_GROUP consecutive blocks are summed from an exact zero in ascending order -- one
// group per lane -- and lane 0 then adds the group sums in order.  The dependent chain is
// 64 + J/64 mpf_adds instead of J (J grows with the number of GPUs).  Called by a whole warp;
// the total is returned in lane 0.  scratch: 32 * TileGeom::SW words of shared memory.
constexpr int BLOCK_SUM_GROUP = 64;
template <int NL>
__device__ __forceinline__ Reg<NL> ordered_block_sum(const limb_t *part, int J, int N, int c, uint32_t *scratch)
{
  typedef TileGeom<NL> G;
  const int lane = threadIdx.x & 31;
  const int ngroups = (J + BLOCK_SUM_GROUP - 1) / BLOCK_SUM_GROUP;
  Reg<NL> total;
  mpfw::set_zero(total);
  for(int base = 0; base < ngroups; base += 32)
    {
      const int g = base + lane;
      Reg<NL> acc, v;
      mpfw::set_zero(acc);
      if(g < ngroups)
        {
          const int j1 = min(J, (g + 1) * BLOCK_SUM_GROUP);
          for(int j = g * BLOCK_SUM_GROUP; j < j1; ++j)
            {
              if(j + 4 < j1) // a sequential sum, but the loads need not wait for it
                asm volatile("prefetch.global.L1 [%0];" ::"l"(part + ((size_t)(j + 4) * N + c) * Fmt<NL>::ES));
              ldg_reg<NL>(v, part + ((size_t)j * N + c) * Fmt<NL>::ES);
              acc = add_nl<NL>(acc, v);
            }
        }
      mpfw::store<NL>(scratch + lane * G::SW, acc);
      __syncwarp();
      if(lane == 0)
        for(int q = 0; q < 32 && base + q < ngroups; ++q)
          {
            mpfw::load<NL>(v, scratch + q * G::SW);
            total = add_nl<NL>(total, v);
          }
      __syncwarp();
    }
  return total;
}
// one warp per column: the canonical sum of the per-block partials, then the warp takes sqrt and
// its reciprocal together (coop.cuh)
template <int NL>
__global__ void __launch_bounds__(32) norm_final_kernel(const limb_t *part, int J, int N,
                                                       limb_t *norms, uint32_t *recip)
{
  typedef TileGeom<NL> G;
  extern __shared__ __align__(16) unsigned char coop_raw[];
  coop::Work<NL> &ws = *reinterpret_cast<coop::Work<NL> *>(coop_raw);
  uint32_t *slot = reinterpret_cast<uint32_t *>(coop_raw + ((sizeof(coop::Work<NL>) + 15) & ~(size_t)15));
  uint32_t *scratch = slot + G::SW;
  const int c = blockIdx.x;
  if(c >= N)
    return;
  const Reg<NL> acc = ordered_block_sum<NL>(part, J, N, c, scratch);
  if(threadIdx.x == 0)
    {
      mpfw::store<NL>(slot, acc);
      ws.flag = 0;
    }
  __syncwarp();
  if((int32_t)slot[1] > 0)
    coop::pivot<NL>(ws, slot, nullptr, recip + (size_t)c * G::RS);
  uint32_t *dst = reinterpret_cast<uint32_t *>(norms + (size_t)c * Fmt<NL>::ES);
  for(int i = threadIdx.x; i < G::EW; i += 32)
    dst[i] = slot[i];
}

// residue tables for the CRT syrk
struct CrtTables
{
  const uint32_t *primes; // [np]
  const uint32_t *pow28;  // [np][nd]  2^(28k) mod p
  const uint32_t *ginv;   // [np][np]  ginv[i*np+j] = p_j^{-1} mod p_i (j < i)
  const uint32_t *M;      // [mw] product of all primes, 32-bit words
  const uint32_t *Mhalf;  // [mw] (M+1)/2
  int np, nd, mw;
  const uint32_t *pow28p; // [np][ndp] the same powers, rows zero-padded to ndp (a multiple of 4)
  const uint64_t *inv64;  // [np] floor((2^64 - 1) / p): Barrett constant
  int ndp;
};

// x mod p for x < 2^63, p < 2^28, inv = floor((2^64 - 1) / p)
__device__ __forceinline__ uint32_t barrett_mod(uint64_t x, uint32_t p, uint64_t inv)
{
  const uint64_t q = __umul64hi(x, inv); // floor(x / p) - 2 <= q <= floor(x / p)
  uint32_t r = (uint32_t)x - (uint32_t)q * p;
  if(r >= p)
    r -= p;
  if(r >= p)
    r -= p;
  return r;
}

// P' = (P / norm) << prec (Matrix_Normalizer.cxx:174-190) held in registers, the
// residues of trunc(P') (fmpz_set_mpf truncates toward zero,
// fmpz_BigFloat_convert.hxx:13) modulo every prime:
// R[(p*K + row)*NS + col], NS = row stride (N rounded up to 16; the pad
// columns stay zero).  flags[0] is raised if |trunc(P')| does not fit.
//
// One thread per element, consecutive threads on consecutive COLUMNS of one
// row, so the residue stores coalesce (an element is a whole 128-byte line
// either way).  The integer part is cut into 28-bit digits held in registers;
// a residue is sum_k digit_k * (2^(28k) mod p) -- IMAD.WIDE against a table in
// shared memory -- followed by one Barrett reduction.
template <int NL>
__device__ __noinline__ Reg<NL> mul_nl(Reg<NL> a, const uint32_t *b)
{
  Reg<NL> r;
  if(a.sign == 0 || (int32_t)b[1] == 0)
    {
      mpfw::set_zero(r);
      return r;
    }
  int32_t bexp, bsign;
  uint32_t bw[2 * NL];
  mpfw::load_packed<NL>(bexp, bsign, bw, b);
  const uint32_t *aw = a.w;
  mpfw::mul<NL>(r, a.sign, a.exp, aw, bsign, bexp, bw);
  return r;
}
template <int NL> struct NormGeom
{
  static constexpr int ND = (64 * (NL - 2) + 2 + 27) / 28; // digits of an integer below 2^(prec+2), prec <= 64 (NL-2)
  static constexpr int NDP = (ND + 3) & ~3;
};
template <int NL>
__global__ void __launch_bounds__(128, 4)
normalize_kernel(const BandDesc *bands, int N, int NS, long K, const limb_t *norms,
                 const uint32_t *recip, int prec, CrtTables T, uint32_t *R, int *flags)
{
  typedef NormGeom<NL> G;
  extern __shared__ __align__(16) uint32_t nsm[]; // [np][NDP] powers, then [np] primes, then [np] inv64
  uint32_t *s_pow = nsm;
  uint32_t *s_p = nsm + (size_t)T.np * G::NDP;
  uint64_t *s_inv = reinterpret_cast<uint64_t *>(s_p + ((T.np + 1) & ~1));
  for(int q = threadIdx.x; q < T.np * G::NDP; q += blockDim.x)
    {
      const int pi = q / G::NDP, k = q % G::NDP;
      s_pow[q] = k < T.ndp ? T.pow28p[(size_t)pi * T.ndp + k] : 0u;
    }
  for(int q = threadIdx.x; q < T.np; q += blockDim.x)
    {
      s_p[q] = T.primes[q];
      s_inv[q] = T.inv64[q];
    }
  __syncthreads();
  const BandDesc b = bands[blockIdx.x];
  const long total = (long)b.rows * N;
  const int nd = T.nd;
  for(long e = (long)blockIdx.y * blockDim.x + threadIdx.x; e < total;
      e += (long)gridDim.y * blockDim.x)
    {
      const int c = (int)(e % N), r = (int)(e / N);
      limb_t *elem = b.P + ((size_t)c * b.rows + r) * Fmt<NL>::ES;
      const uint32_t *nrm = reinterpret_cast<const uint32_t *>(norms + (size_t)c * Fmt<NL>::ES);
      Reg<NL> v;
      ldg_reg<NL>(v, elem);
      if((int32_t)nrm[1] != 0)
        {
          v = div_nl<NL>(v, nrm, recip + (size_t)c * TileGeom<NL>::RS);
          if((prec & 63) == 0)
            {
              if(v.sign != 0)
                v.exp += prec >> 6; // mpf_mul_2exp by whole limbs: exponent only
            }
          else
            {
              Num<NL> t;
              mpfw::to_num(t, v);
              mpfx::mul_2exp(t, t, (uint32_t)prec);
              mpfw::from_num(v, t);
            }
          // restore_P (Matrix_Normalizer.cxx:210-226) fused: the stored band goes straight to
          // P = (P' >> prec) * norm; P' itself is only needed for the residues below
          Reg<NL> back = v;
          if((prec & 63) == 0)
            {
              if(back.sign != 0)
                back.exp -= prec >> 6;
            }
          else
            {
              Num<NL> t;
              mpfw::to_num(t, back);
              mpfx::div_2exp(t, t, (uint32_t)prec);
              mpfw::from_num(back, t);
            }
          back = mul_nl<NL>(back, nrm);
          stg_reg<NL>(elem, back);
        }
      // integer part: the mantissa shifted right by NL - exp limbs
      uint32_t iw[2 * NL];
#pragma unroll
      for(int i = 0; i < 2 * NL; ++i)
        iw[i] = v.w[i];
      bool fits = true;
      if(v.sign == 0 || v.exp <= 0)
        {
#pragma unroll
          for(int i = 0; i < 2 * NL; ++i)
            iw[i] = 0;
        }
      else if(v.exp > NL)
        fits = false;
      else
        mpfw::shr_limbs<NL>(iw, NL - v.exp);
      // 28-bit digits
      uint32_t dg[G::NDP];
#pragma unroll
      for(int k = 0; k < G::NDP; ++k)
        {
          const int bit = 28 * k, wi = bit >> 5, sh = bit & 31;
          uint32_t x = 0;
          if(k < G::ND && wi < 2 * NL)
            {
              x = iw[wi] >> sh;
              if(sh > 4 && wi + 1 < 2 * NL)
                x |= iw[wi + 1] << (32 - sh);
              x &= 0x0FFFFFFFu;
            }
          dg[k] = x;
          if(k >= nd && x)
            fits = false;
        }
#pragma unroll
      for(int i = 0; i < 2 * NL; ++i) // bits beyond the last digit
        {
          const int lo = 32 * i, cut = 28 * G::ND;
          if(lo >= cut)
            fits = fits && iw[i] == 0;
          else if(lo + 32 > cut)
            fits = fits && (iw[i] >> (cut - lo)) == 0;
        }
      if(!fits)
        atomicExch(&flags[0], 1);
      uint32_t *out = R + ((size_t)b.row0 + r) * NS + c;
      const bool neg = v.sign < 0;
      for(int pi = 0; pi < T.np; ++pi)
        {
          const uint4 *pw = reinterpret_cast<const uint4 *>(s_pow + (size_t)pi * G::NDP);
          uint64_t acc = 0; // < ND * 2^56
#pragma unroll
          for(int k4 = 0; k4 < G::NDP / 4; ++k4)
            {
              const uint4 t = pw[k4];
              acc += (uint64_t)dg[4 * k4] * t.x;
              acc += (uint64_t)dg[4 * k4 + 1] * t.y;
              acc += (uint64_t)dg[4 * k4 + 2] * t.z;
              acc += (uint64_t)dg[4 * k4 + 3] * t.w;
            }
          const uint32_t p = s_p[pi];
          uint32_t res = barrett_mod(acc, p, s_inv[pi]);
          if(neg && res)
            res = p - res;
          out[(size_t)pi * K * NS] = res;
        }
    }
}

// Qres[p][i][j] = sum_rows R[p][row][i] R[p][row][j] mod p, i <= j.
// A CTA owns one 16x16 output tile of one prime.  Its 256 threads form 16
// groups of 16; group g takes the rows congruent to g modulo 16 and each of
// its threads a 4x4 patch of the tile in registers, fed by two 128-bit loads
// per row (16 IMAD.WIDE per 2 loads).  Products are below 2^56; every 128 rows
// the 64-bit sums are folded with 2^32 mod p, which keeps them below 2^61.
// The 16 partial tiles meet in shared memory at the end.
template <int UNROLL>
__global__ void __launch_bounds__(256, 2)
syrk_mod_kernel(const uint32_t *__restrict__ R, long K, int N, int NS,
                const uint32_t *__restrict__ primes, uint32_t *Qres)
{
  int t = blockIdx.x, tj = 0;
  while(t > tj)
    {
      t -= tj + 1;
      ++tj;
    }
  const int ti = t; // ti <= tj
  const int pi = blockIdx.y;
  const uint32_t p = primes[pi];
  const uint32_t c32 = (uint32_t)((1ull << 32) % p);
  const int g = threadIdx.x >> 4, l = threadIdx.x & 15;
  const int tx = l & 3, ty = l >> 2;
  const uint32_t *Ra = R + (size_t)pi * K * NS + ti * 16 + 4 * tx;
  const uint32_t *Rb = R + (size_t)pi * K * NS + tj * 16 + 4 * ty;
  uint64_t acc[4][4];
#pragma unroll
  for(int a = 0; a < 4; ++a)
#pragma unroll
    for(int b = 0; b < 4; ++b)
      acc[a][b] = 0;
  int since = 0;
  long row = g;
  auto step = [&](const uint4 &va, const uint4 &vb) {
    const uint32_t xa[4] = {va.x, va.y, va.z, va.w}, xb[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
    for(int a = 0; a < 4; ++a)
#pragma unroll
      for(int b = 0; b < 4; ++b)
        acc[a][b] += (uint64_t)xa[a] * xb[b];
  };
  auto fold = [&]() {
#pragma unroll
    for(int a = 0; a < 4; ++a)
#pragma unroll
      for(int b = 0; b < 4; ++b)
        acc[a][b] = (acc[a][b] & 0xFFFFFFFFull) + (acc[a][b] >> 32) * c32;
  };
  for(; row + 16 * (UNROLL - 1) < K; row += 16 * UNROLL)
    {
      uint4 va[UNROLL], vb[UNROLL];
#pragma unroll
      for(int u = 0; u < UNROLL; ++u)
        {
          va[u] = __ldg(reinterpret_cast<const uint4 *>(Ra + (size_t)(row + 16 * u) * NS));
          vb[u] = __ldg(reinterpret_cast<const uint4 *>(Rb + (size_t)(row + 16 * u) * NS));
        }
#pragma unroll
      for(int u = 0; u < UNROLL; ++u)
        step(va[u], vb[u]);
      since += UNROLL;
      if(since >= 128)
        {
          fold();
          since = 0;
        }
    }
  for(; row < K; row += 16)
    {
      const uint4 va = __ldg(reinterpret_cast<const uint4 *>(Ra + (size_t)row * NS));
      const uint4 vb = __ldg(reinterpret_cast<const uint4 *>(Rb + (size_t)row * NS));
      step(va, vb);
    }
  // <= 2^61 + 3 * 2^56 each: reduce, then add the 16 groups' partial tiles
  __shared__ uint32_t part[16][256];
#pragma unroll
  for(int a = 0; a < 4; ++a)
#pragma unroll
    for(int b = 0; b < 4; ++b)
      part[g][(4 * ty + b) * 16 + 4 * tx + a] = (uint32_t)(acc[a][b] % p);
  __syncthreads();
  uint32_t sum = 0; // 16 * (p - 1) < 2^32
#pragma unroll
  for(int q = 0; q < 16; ++q)
    sum += part[q][threadIdx.x];
  const int i = ti * 16 + (threadIdx.x & 15), j = tj * 16 + (threadIdx.x >> 4);
  if(i < N && j < N && i <= j)
    Qres[((size_t)pi * N + i) * N + j] = sum % p;
}

// Garner reconstruction of the signed integer Q'_ij from its residues,
// conversion with fmpz_get_mpf semantics, check_normalized_Q_diagonal
// (compute_Q.cxx:65-91) and restore_Q (Matrix_Normalizer.cxx:245-265):
// Q_ij = ((Q'_ij >> 2 prec) * n_i) * n_j for i <= j; lower triangle zeroed.
template <int NL>
__global__ void __launch_bounds__(64)
crt_restore_kernel(const uint32_t *__restrict__ Qres, int N, int prec,
                   CrtTables T, const limb_t *norms, limb_t *Q, int *flags)
{
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if(e >= (long)N * N)
    return;
  const int i = (int)(e % N), j = (int)(e / N);
  Num<NL> q;
  if(i > j)
    {
      mpfx::set_zero(q);
      st(Q, (size_t)j * N + i, q);
      return;
    }
  constexpr int MAXP = 8 * NL + 8; // primes: ~ (2*64*(NL-2)+44)/28
  constexpr int MAXMW = 8 * NL + 8;
  const int np = T.np, mw = T.mw;
  uint32_t v[MAXP];
  for(int a = 0; a < np; ++a)
    {
      const uint32_t pa = T.primes[a];
      const uint64_t inva = T.inv64[a]; // Barrett: the 64-bit `%` by a run-time prime is a ~70-instruction
                                        // routine, and this recurrence is 2 np^2 / 2 of them in a row
      // a sum of per-rank residues when the blocks are sharded over GPUs
      uint32_t t = barrett_mod(Qres[((size_t)a * N + i) * N + j], pa, inva);
      for(int b = 0; b < a; ++b)
        {
          // t = (t - v_b) * p_b^{-1} mod p_a
          const uint32_t vb = barrett_mod(v[b], pa, inva);
          const uint32_t diff = t >= vb ? t - vb : t + pa - vb;
          t = barrett_mod((uint64_t)diff * T.ginv[(size_t)a * np + b], pa, inva);
        }
      v[a] = t;
    }
  // x = v0 + p0 (v1 + p1 (v2 + ...)) in 32-bit words
  uint32_t x[MAXMW];
  for(int k = 0; k < mw; ++k)
    x[k] = 0;
  for(int a = np - 1; a >= 0; --a)
    {
      const uint32_t pa = T.primes[a];
      uint64_t carry = v[a];
      for(int k = 0; k < mw; ++k)
        {
          const uint64_t z = (uint64_t)x[k] * pa + carry;
          x[k] = (uint32_t)z;
          carry = z >> 32;
        }
    }
  // centre: x >= (M+1)/2  ->  x - M (negative)
  int sign = 1;
  {
    int c = 0;
    for(int k = mw - 1; k >= 0 && c == 0; --k)
      if(x[k] != T.Mhalf[k])
        c = x[k] > T.Mhalf[k] ? 1 : -1;
    if(c >= 0)
      {
        // x = M - x
        uint64_t borrow = 0;
        for(int k = 0; k < mw; ++k)
          {
            const uint64_t z = (uint64_t)T.M[k] - x[k] - borrow;
            x[k] = (uint32_t)z;
            borrow = (z >> 32) & 1;
          }
        sign = -1;
      }
  }
  limb_t zl[MAXMW / 2 + 1];
  const int zn = (mw + 1) / 2;
  for(int k = 0; k < zn; ++k)
    {
      const uint64_t lo = x[2 * k];
      const uint64_t hi = (2 * k + 1 < mw) ? x[2 * k + 1] : 0;
      zl[k] = lo | (hi << 32);
    }
  mpfx::from_limbs(q, zl, zn, sign);
  mpfx::div_2exp(q, q, 2u * (uint32_t)prec);
  if(i == j)
    {
      Num<NL> one, diff, eps;
      mpfx::set_zero(one);
      one.sign = 1;
      one.exp = 1;
      one.d[NL - 1] = 1;
      mpfx::sub(diff, q, one);
      if(diff.sign < 0)
        diff.sign = 1;
      mpfx::div_2exp(eps, one, (uint32_t)(prec / 2));
      if(!(mpfx::cmp(diff, eps) < 0))
        atomicMin(&flags[1], i);
    }
  Num<NL> ni, nj;
  ld(ni, norms, i);
  ld(nj, norms, j);
  mpfx::mul(q, q, ni);
  mpfx::mul(q, q, nj);
  st(Q, (size_t)j * N + i, q);
}

// fail[0] = 1 if any status word of this rank's step reports a failed pivot, or the residue
// overflow / Q-diagonal flags are raised (summed over the ranks afterwards)
static __global__ void fail_flag_kernel(const int *status, int n, const int *flags, int *fail)
{
  __shared__ int bad;
  if(threadIdx.x == 0)
    bad = (flags[0] != 0 || flags[1] != INT_MAX) ? 1 : 0;
  __syncthreads();
  for(int i = threadIdx.x; i < n; i += blockDim.x)
    if(status[i] >= 0)
      bad = 1;
  __syncthreads();
  if(threadIdx.x == 0)
    fail[0] = bad;
}
// out[e] = element (i, i) of matrix m for the stacked diagonal index e; per matrix: base pointer,
// size, and the element offset of its first diagonal entry in `out`
struct DiagDesc
{
  const limb_t *A;
  int s;
  long ld;   // elements between consecutive columns
  long out0; // first output element
};
static __global__ void diag_gather_kernel(const DiagDesc *descs, int count, int es, limb_t *out)
{
  for(int m = blockIdx.x; m < count; m += gridDim.x)
    {
      const DiagDesc d = descs[m];
      for(int e = threadIdx.x; e < d.s * es; e += blockDim.x)
        {
          const int i = e / es, w = e % es;
          out[(d.out0 + i) * es + w] = d.A[((long)i * d.ld + i) * es + w];
        }
    }
}

// scalar-op kernel for device-vs-libgmp parity tests (same op codes as
// oracle_scalar_op): 0 mul 1 add 2 sub 3 div 4 sqrt 5 <<k 6 >>k 7 /4
template <int NL>
__global__ void scalar_op_kernel(int op, int k, long count, const limb_t *a,
                                 const limb_t *b, limb_t *r)
{
  const long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if(e >= count)
    return;
  Num<NL> x, y, z;
  ld(x, a, e);
  ld(y, b, e);
  mpfx::set_zero(z);
  switch(op)
    {
    case 0: mpfx::mul(z, x, y); break;
    case 1: mpfx::add(z, x, y); break;
    case 2: mpfx::sub(z, x, y); break;
    case 3:
      if(y.sign != 0)
        mpfx::div(z, x, y);
      break;
    case 4:
      if(x.sign >= 0)
        mpfx::sqrt(z, x);
      break;
    case 5: mpfx::mul_2exp(z, x, (uint32_t)k); break;
    case 6: mpfx::div_2exp(z, x, (uint32_t)k); break;
    case 7: mpfx::div4(z, x); break;
    }
  st(r, e, z);
}
// test hook for the warp-cooperative pivot (coop.cuh): one warp per element.
// op 8: r = mpf_sqrt(a);  op 9: r = [sign 1, exp 1, w0 = number of words in which the
// cooperative reciprocal of sqrt(a) differs from mpfw::reciprocal (Knuth), other words 0]
template <int NL>
__global__ void __launch_bounds__(32) coop_test_kernel(int op, long count, const limb_t *a, limb_t *r)
{
  typedef TileGeom<NL> G;
  extern __shared__ __align__(16) unsigned char coop_raw[];
  coop::Work<NL> &ws = *reinterpret_cast<coop::Work<NL> *>(coop_raw);
  uint32_t *slot = reinterpret_cast<uint32_t *>(coop_raw + ((sizeof(coop::Work<NL>) + 15) & ~(size_t)15));
  uint32_t *R = slot + G::SW;
  const long e = blockIdx.x;
  if(e >= count)
    return;
  const uint32_t *src = reinterpret_cast<const uint32_t *>(a + e * Fmt<NL>::ES);
  for(int i = threadIdx.x; i < G::EW; i += 32)
    slot[i] = src[i];
  if(threadIdx.x == 0)
    ws.flag = 0;
  __syncwarp();
  Reg<NL> out;
  mpfw::set_zero(out);
  if((int32_t)slot[1] > 0)
    {
      coop::pivot<NL>(ws, slot, R, nullptr);
      if(threadIdx.x == 0)
        {
          mpfw::load<NL>(out, slot);
          if(op == 9)
            {
              Num<NL> x;
              mpfw::to_num(x, out);
              uint32_t rw[2 * NL + 4];
              mpfw::reciprocal<NL>(rw, x); // Knuth long division: independent of the Newton code
              uint32_t bad = ws.flag;      // a fallback counts as a mismatch: the fast path must close
              for(int w = 0; w < G::RW; ++w)
                bad += rw[w] != R[w];
              mpfw::set_zero(out);
              out.sign = 1;
              out.exp = 1;
              out.w[0] = bad;
            }
        }
    }
  if(threadIdx.x == 0)
    stg_reg<NL>(r + e * Fmt<NL>::ES, out);
}
} // namespace sdpb_b200

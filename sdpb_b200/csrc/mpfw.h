// mpfw — the hot-loop form of mpfx: the same GMP-mpf-exact arithmetic, but on
// 32-bit words held entirely in registers, built around the operation every
// GEMM-class kernel of the Schur step repeats ~10^9 times per Newton
// iteration:    acc <- mpf_add/sub(acc, mpf_mul(a, b)).
//
// Multiply (GMP mpf/mul.c): the top P limbs of both operands, exact product,
// top P+1 limbs kept.  Here: product scanning over 32-bit words with
// IMAD.WIDE.U32 + carry (one instruction per partial product on sm_100a), and
// only the columns that can reach the kept window are formed (a "short
// product", ~31 % fewer partial products at 768 bits).  The dropped columns
// can change the kept words only through a carry that has to cross a whole
// guard word; when that guard word is within 2^-24 of overflowing the exact
// full product is formed instead, so the result is always the truncated exact
// product.
//
// Add/sub (GMP mpf/add.c, mpf/sub.c) on full-size operands reduce to: align the
// operand with the smaller exponent by a whole-limb right shift (a barrel
// shifter on registers, no dynamic indexing), add or subtract exactly over the
// window GMP uses (P limbs for add, P+1 for sub), then strip leading zero
// limbs.  mpf_sub's "operands extremely close" paths coincide with the exact
// subtraction except when the exponents differ by one limb; that rare case
// (and nothing else) is handed to the verified generic mpfx::sub.
//
// Every routine compiles for the host as well (portable carry primitives), and
// tests/cpp/mpfw_fuzz.cpp checks them against mpfx (itself fuzzed against
// libgmp) on random and adversarial operands.
#pragma once
#include "mpfx.h"

namespace mpfw
{
#if defined(__CUDA_ARCH__)
#define MPFW_D __device__ __forceinline__
#define MPFW_NOINLINE __device__ __noinline__
#elif defined(__CUDACC__)
#define MPFW_D __host__ __device__ inline
#define MPFW_NOINLINE __host__ __device__
#else
#define MPFW_D inline
#define MPFW_NOINLINE
#endif

#ifdef MPFW_COUNT_RARE
static long rare_mul_count = 0, rare_sub_count = 0, recip_fallbacks = 0, sqrt_fallbacks = 0, div_exact_count = 0, div_fast_count = 0; // host fuzz: rare paths exercised?
#endif
// ------------------------------------------------------------------ carries
// (t2:t1:t0) += a * b
MPFW_D void mac3(uint32_t &t0, uint32_t &t1, uint32_t &t2, uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
  asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
      "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
      "addc.u32 %2, %2, 0;"
      : "+r"(t0), "+r"(t1), "+r"(t2)
      : "r"(a), "r"(b));
#else
  const uint64_t p = (uint64_t)a * b;
  const uint64_t s0 = (uint64_t)t0 + (uint32_t)p;
  const uint64_t s1 = (uint64_t)t1 + (uint32_t)(p >> 32) + (s0 >> 32);
  t0 = (uint32_t)s0;
  t1 = (uint32_t)s1;
  t2 += (uint32_t)(s1 >> 32);
#endif
}

// r = u + v (N words), returns the carry out
template <int N>
MPFW_D uint32_t add_n(uint32_t (&r)[N], const uint32_t (&u)[N], const uint32_t (&v)[N])
{
#if defined(__CUDA_ARCH__)
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r[0]) : "r"(u[0]), "r"(v[0]));
#pragma unroll
  for(int i = 1; i < N; ++i)
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r[i]) : "r"(u[i]), "r"(v[i]));
  uint32_t c;
  asm volatile("addc.u32 %0, 0, 0;" : "=r"(c));
  return c;
#else
  uint64_t c = 0;
  for(int i = 0; i < N; ++i)
    {
      const uint64_t s = (uint64_t)u[i] + v[i] + c;
      r[i] = (uint32_t)s;
      c = s >> 32;
    }
  return (uint32_t)c;
#endif
}
// r = u - v (N words), returns the borrow out (0 or 1)
template <int N>
MPFW_D uint32_t sub_n(uint32_t (&r)[N], const uint32_t (&u)[N], const uint32_t (&v)[N])
{
#if defined(__CUDA_ARCH__)
  asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r[0]) : "r"(u[0]), "r"(v[0]));
#pragma unroll
  for(int i = 1; i < N; ++i)
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r[i]) : "r"(u[i]), "r"(v[i]));
  uint32_t b;
  asm volatile("subc.u32 %0, 0, 0;" : "=r"(b));
  return b & 1u;
#else
  uint64_t b = 0;
  for(int i = 0; i < N; ++i)
    {
      const uint64_t s = (uint64_t)u[i] - v[i] - b;
      r[i] = (uint32_t)s;
      b = (s >> 32) & 1;
    }
  return (uint32_t)b;
#endif
}
// r = -r (two's complement over N words)
template <int N> MPFW_D void neg_n(uint32_t (&r)[N])
{
#if defined(__CUDA_ARCH__)
  asm volatile("sub.cc.u32 %0, 0, %0;" : "+r"(r[0]));
#pragma unroll
  for(int i = 1; i < N; ++i)
    asm volatile("subc.cc.u32 %0, 0, %0;" : "+r"(r[i]));
#else
  uint64_t b = 0;
  for(int i = 0; i < N; ++i)
    {
      const uint64_t s = (uint64_t)0 - r[i] - b;
      r[i] = (uint32_t)s;
      b = (s >> 32) & 1;
    }
#endif
}

// v >>= k limbs (k in [0, 2^ceil(log2 NL))), zeros shifted in; register barrel shifter
template <int NL> MPFW_D void shr_limbs(uint32_t (&v)[2 * NL], int k)
{
#pragma unroll
  for(int s = 1; s < NL; s <<= 1)
    {
      const bool on = (k & s) != 0;
#pragma unroll
      for(int i = 0; i < 2 * NL; ++i)
        {
          const uint32_t hi = (i + 2 * s < 2 * NL) ? v[i + 2 * s] : 0u;
          v[i] = on ? hi : v[i];
        }
    }
}
// v <<= k limbs, zeros shifted in at the bottom
template <int NL> MPFW_D void shl_limbs(uint32_t (&v)[2 * NL], int k)
{
#pragma unroll
  for(int s = 1; s < NL; s <<= 1)
    {
      const bool on = (k & s) != 0;
#pragma unroll
      for(int i = 2 * NL - 1; i >= 0; --i)
        {
          const uint32_t lo = (i - 2 * s >= 0) ? v[i - 2 * s] : 0u;
          v[i] = on ? lo : v[i];
        }
    }
}

// ------------------------------------------------------------------- number
template <int NL> struct Reg
{
  int32_t sign; // -1, 0, +1
  int32_t exp;  // in 64-bit limbs
  uint32_t w[2 * NL]; // the NL limbs as 32-bit words, little-endian, top-aligned
};

template <int NL> MPFW_D void set_zero(Reg<NL> &r)
{
  r.sign = 0;
  r.exp = 0;
#pragma unroll
  for(int i = 0; i < 2 * NL; ++i)
    r.w[i] = 0;
}
// packed element (mpfx::load layout) viewed as 32-bit words:
// [exp, sign, w0 .. w(2NL-1), pad]
template <int NL>
MPFW_D void load_packed(int32_t &exp, int32_t &sign, uint32_t (&w)[2 * NL], const uint32_t *p);
template <int NL> MPFW_D void load(Reg<NL> &r, const uint32_t *p)
{
  load_packed<NL>(r.exp, r.sign, r.w, p);
}
template <int NL> MPFW_D void store(uint32_t *p, const Reg<NL> &r)
{
  p[0] = (uint32_t)r.exp;
  p[1] = (uint32_t)r.sign;
#pragma unroll
  for(int i = 0; i < 2 * NL; ++i)
    p[2 + i] = r.w[i];
  if((NL + 1) & 1)
    {
      p[2 + 2 * NL] = 0;
      p[3 + 2 * NL] = 0;
    }
}
template <int NL> MPFW_D void to_num(mpfx::Num<NL> &n, const Reg<NL> &r)
{
  n.sign = r.sign;
  n.exp = r.exp;
#pragma unroll
  for(int i = 0; i < NL; ++i)
    n.d[i] = (uint64_t)r.w[2 * i] | ((uint64_t)r.w[2 * i + 1] << 32);
}
template <int NL> MPFW_D void from_num(Reg<NL> &r, const mpfx::Num<NL> &n)
{
  r.sign = n.sign;
  r.exp = n.exp;
#pragma unroll
  for(int i = 0; i < NL; ++i)
    {
      r.w[2 * i] = (uint32_t)n.d[i];
      r.w[2 * i + 1] = (uint32_t)(n.d[i] >> 32);
    }
}

// ----------------------------------------------------------------- multiply
template <int NL> struct MulGeom
{
  static constexpr int W = 2 * (NL - 1);             // operand words used
  static constexpr int C0 = (W - 6 > 0) ? W - 6 : 0; // first column formed
  static constexpr int NO = 2 * W - C0;              // product words kept
};

// out[c - C0] = word c of the exact product of a[0..W) and b[0..W) for
// c >= FROM, assuming nothing below column FROM carries in.
template <int W, int FROM, int C0>
MPFW_D void mul_columns(uint32_t (&out)[2 * W - C0], const uint32_t *a, const uint32_t *b)
{
  uint32_t t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
  for(int c = FROM; c < 2 * W - 1; ++c)
    {
#pragma unroll
      for(int i = 0; i < W; ++i)
        {
          const int j = c - i;
          if(j >= 0 && j < W)
            mac3(t0, t1, t2, a[i], b[j]);
        }
      if(c >= C0)
        out[c - C0] = t0;
      t0 = t1;
      t1 = t2;
      t2 = 0;
    }
  out[2 * W - 1 - C0] = t0;
}
// Same result as mul_columns<W, C0, C0>, formed by rows (operand scanning):
// for each word a_i the products a_i b_j go into two sets of 64-bit lanes, the
// ones with i+j even into `ev`, the ones with i+j odd into `od`, so that every
// product is ONE IMAD.WIDE.U32.X whose carry feeds the next lane of the same
// row (ptxas fuses each mad.lo.cc / madc.hi.cc pair).  No per-product carry
// bookkeeping, and `a` is read one word at a time (it may live in shared
// memory), which keeps the live registers at b + two lane sets.
// Rows ascend, so the word a chain's final carry lands in has only ever
// received carries: it cannot overflow.
template <int W, int C0, class APtr>
MPFW_D void mul_rows_short(uint32_t (&out)[2 * W - C0], APtr a, const uint32_t (&b)[W])
{
  static_assert((C0 & 1) == 0 && (W & 1) == 0, "lane pairing needs even W and C0");
#if defined(__CUDA_ARCH__)
  constexpr int NO = 2 * W - C0;
  uint32_t ev[NO], od[NO]; // od[k] = word k+1
#pragma unroll
  for(int k = 0; k < NO; ++k)
    ev[k] = od[k] = 0;
#if !defined(MPFW_MUL_MANYCHAINS)
  // ONE carry chain through the whole product.  A row's chain ends in a word that has only
  // ever received carries (or past the top word), so its carry-out is always zero; handing
  // that zero to the next chain as its carry-in changes nothing arithmetically, but it makes
  // every chain depend on the previous one.  ptxas then keeps a single carry predicate live
  // instead of interleaving ~20 independent chains and spilling their predicates into a
  // bit-mask register (3 LOP3 per product and a serial dependency through that register):
  // mac_nl inside trsm_gemm_level 1604 -> 1501 instructions, LOP3 215 -> 74; step 118.8 ->
  // 117.2 ms at c3.  Four warps per scheduler hide the chain's latency: the microbenchmark
  // (tools/mac_bench.cu) measures the same rate for both forms.  MPFW_MUL_MANYCHAINS keeps
  // the independent-chain form for A/B measurements.
  bool started = false;
#pragma unroll
  for(int i = 0; i < W; ++i)
    {
      const uint32_t ai = a[i];
      const int jmin = (C0 - i > 0) ? C0 - i : 0;
#pragma unroll
      for(int par = 0; par < 2; ++par)
        {
          int ctop = -1;
#pragma unroll
          for(int j = 0; j < W; ++j)
            {
              const int c = i + j - C0;
              if(j < jmin || (c & 1) != par)
                continue;
              uint32_t &lo = par ? od[c - 1] : ev[c];
              uint32_t &hi = par ? od[c] : ev[c + 1];
              if(!started)
                asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                             : "+r"(lo), "+r"(hi)
                             : "r"(ai), "r"(b[j]));
              else
                asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                             : "+r"(lo), "+r"(hi)
                             : "r"(ai), "r"(b[j]));
              started = true;
              ctop = c;
            }
          if(par == 0 && ctop >= 0 && ctop + 2 < NO)
            asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(ev[ctop + 2]));
          if(par == 1 && ctop >= 0 && ctop + 1 < NO - 1)
            asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(od[ctop + 1]));
        }
    }
#else
#pragma unroll
  for(int i = 0; i < W; ++i)
    {
      const uint32_t ai = a[i];
      const int jmin = (C0 - i > 0) ? C0 - i : 0;
      bool first = true;
      int ctop = -1;
#pragma unroll
      for(int j = 0; j < W; ++j)
        {
          const int c = i + j - C0;
          if(j < jmin || (c & 1))
            continue;
          if(first)
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                         : "+r"(ev[c]), "+r"(ev[c + 1])
                         : "r"(ai), "r"(b[j]));
          else
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                         : "+r"(ev[c]), "+r"(ev[c + 1])
                         : "r"(ai), "r"(b[j]));
          first = false;
          ctop = c;
        }
      if(ctop >= 0 && ctop + 2 < NO)
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(ev[ctop + 2]));
      first = true;
      ctop = -1;
#pragma unroll
      for(int j = 0; j < W; ++j)
        {
          const int c = i + j - C0;
          if(j < jmin || !(c & 1))
            continue;
          if(first)
            asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                         : "+r"(od[c - 1]), "+r"(od[c])
                         : "r"(ai), "r"(b[j]));
          else
            asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;"
                         : "+r"(od[c - 1]), "+r"(od[c])
                         : "r"(ai), "r"(b[j]));
          first = false;
          ctop = c;
        }
      if(ctop >= 0 && ctop + 1 < NO - 1)
        asm volatile("addc.u32 %0, %0, 0;" : "+r"(od[ctop + 1]));
    }
#endif
  out[0] = ev[0];
  asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(out[1]) : "r"(ev[1]), "r"(od[0]));
#pragma unroll
  for(int q = 2; q < NO - 1; ++q)
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(out[q]) : "r"(ev[q]), "r"(od[q - 1]));
  asm volatile("addc.u32 %0, %1, %2;" : "=r"(out[NO - 1]) : "r"(ev[NO - 1]), "r"(od[NO - 2]));
#else
  uint32_t aw[W];
  for(int i = 0; i < W; ++i)
    aw[i] = a[i];
  mul_columns<W, C0, C0>(out, aw, b);
#endif
}

// the exact product, all columns (rare path; kept out of line, operands and
// result by value so that the caller's arrays never have their address taken
// and stay in registers; a pointer-based interface through a local buffer was
// measured slower, profiles/mac_bench_r01_v2.jsonl)
template <int NL> struct MulWords
{
  uint32_t w[2 * NL];
};
template <int NL> struct MulProd
{
  uint32_t p[MulGeom<NL>::NO];
};
template <int NL>
MPFW_NOINLINE MulProd<NL> mul_full(MulWords<NL> a, MulWords<NL> b)
{
  MulProd<NL> r;
  mul_columns<MulGeom<NL>::W, 0, MulGeom<NL>::C0>(r.p, a.w + 2, b.w + 2);
  return r;
}

// ---- radix-2^29 short product (the hot multiply) ---------------------------
// On sm_100a IMAD.WIDE.U32 with a plain 64-bit accumulate issues at the full
// 64 lanes/clk/SM, but every carry-using form (carry-out predicate, .X, or
// IMAD.HI) runs at half that (profiles/imad_rate2_r01.jsonl).  So the product
// is formed without carries: both operands are re-cut into 29-bit digits,
// digit products (< 2^58) are summed in independent 64-bit column lanes (at
// most 2^5 products per lane), and only the lanes are normalised afterwards.
// Lanes below C0 are not formed; what they would add is below one unit of the
// guard word GW, so the kept words are exact unless the guard word is all
// ones (probability 2^-32), in which case the exact full product is used.
template <int NL> struct Mul29Geom
{
  static constexpr int W = 2 * (NL - 1);       // operand words used
  static constexpr int ND = (32 * W + 28) / 29; // 29-bit digits per operand
  static constexpr int NLANE = 2 * ND - 1;
  static constexpr bool SHORT = W > 6;
  static constexpr int GW = SHORT ? W - 5 : 0; // guard word (product word index)
  static constexpr int NOUT = 2 * W - GW;      // product words GW .. 2W-1
  static constexpr int bits_of(int c) { return c == 0 ? 0 : 1 + bits_of(c >> 1); }
  // largest c0 with c0 * 2^(29 c0 + 29) <= 2^(32 GW): omitted lanes < one unit of word GW
  static constexpr int find_c0(int c)
  {
    return (29 * (c + 1) + 29 + bits_of(c + 1) <= 32 * GW) ? find_c0(c + 1) : c;
  }
  static constexpr int C0 = SHORT ? find_c0(0) : 0;
  static constexpr int T = 32 * GW - 29 * C0; // bits of the lane stream below word GW
  static constexpr int NL29 = NLANE - C0;     // lanes formed
  static_assert(!SHORT || (29 * C0 + 29 + bits_of(C0) <= 32 * GW && T >= 0), "guard bound");
};

// digit k (29 bits) of the W-word integer w[0..W)
template <int W, int K, class WPtr> MPFW_D uint32_t digit29(WPtr w)
{
  constexpr int bit = 29 * K, q = bit >> 5, r = bit & 31;
  const uint32_t lo = w[q];
  uint32_t v = lo >> r;
  if constexpr(r > 3 && q + 1 < W)
    v |= w[q + 1] << (32 - r);
  return v & 0x1FFFFFFFu;
}
template <int W, int ND, int K = 0, class WPtr> MPFW_D void digits29(uint32_t (&d)[ND], WPtr w)
{
  if constexpr(K < ND)
    {
      d[K] = digit29<W, K>(w);
      digits29<W, ND, K + 1>(d, w);
    }
}

// one row of the digit product: lanes[i + j - C0] += a_i * b_j for i + j >= C0
template <int NL, int I, class APtr>
MPFW_D void mul29_rows(uint64_t (&lane)[Mul29Geom<NL>::NL29], APtr aw,
                       const uint32_t (&bd)[Mul29Geom<NL>::ND])
{
  typedef Mul29Geom<NL> G;
  if constexpr(I < G::ND)
    {
      const uint32_t ai = digit29<G::W, I>(aw);
#pragma unroll
      for(int j = 0; j < G::ND; ++j)
        if(I + j >= G::C0)
          lane[I + j - G::C0] += (uint64_t)ai * bd[j];
      mul29_rows<NL, I + 1>(lane, aw, bd);
    }
}

// out[k] = word GW + k of  sum_{i+j >= C0} a_i b_j 2^(29 (i+j)),  k in [0, NOUT)
template <int NL, class APtr>
MPFW_D void mul29_words(uint32_t (&out)[Mul29Geom<NL>::NOUT], APtr aw, const uint32_t (&bd)[Mul29Geom<NL>::ND])
{
  typedef Mul29Geom<NL> G;
  uint64_t lane[G::NL29];
#pragma unroll
  for(int c = 0; c < G::NL29; ++c)
    lane[c] = 0;
  mul29_rows<NL, 0>(lane, aw, bd);
  // normalise: 29-bit digits, the top one keeps everything that is left
  uint32_t dg[G::NL29];
  uint64_t carry = 0;
#pragma unroll
  for(int c = 0; c < G::NL29 - 1; ++c)
    {
      const uint64_t x = lane[c] + carry;
      dg[c] = (uint32_t)x & 0x1FFFFFFFu;
      carry = x >> 29;
    }
  const uint64_t top = lane[G::NL29 - 1] + carry;
  // re-cut into 32-bit words starting T bits into the stream
#pragma unroll
  for(int k = 0; k < G::NOUT; ++k)
    {
      const int bit0 = G::T + 32 * k;
      uint32_t v = 0;
#pragma unroll
      for(int q = 0; q < G::NL29; ++q)
        {
          const int s = 29 * q - bit0; // digit q sits s bits above the word's bit 0
          if(q < G::NL29 - 1)
            {
              if(s >= 0 && s < 32)
                v |= dg[q] << s;
              else if(s < 0 && -s < 29)
                v |= dg[q] >> (-s);
            }
          else
            {
              if(s >= 0 && s < 32)
                v |= (uint32_t)(top << s);
              else if(s < 0 && -s < 64)
                v |= (uint32_t)(top >> (-s));
            }
        }
      out[k] = v;
    }
}

// assemble mpf_mul's result from product words GW.. (out) : top limb zero -> one limb lower
template <int NL>
MPFW_D void mul_finish(Reg<NL> &r, const uint32_t (&out)[Mul29Geom<NL>::NOUT], int32_t asign, int32_t aexp,
                       int32_t bsign, int32_t bexp)
{
  typedef Mul29Geom<NL> G;
  constexpr int W = G::W;
  const bool adj = (out[G::NOUT - 1] | out[G::NOUT - 2]) == 0;
#pragma unroll
  for(int i = 0; i < 2 * NL; ++i)
    {
      const uint32_t hi = out[W - 2 + i - G::GW];
      const uint32_t lo = out[W - 4 + i - G::GW];
      r.w[i] = adj ? lo : hi;
    }
  r.exp = aexp + bexp - (adj ? 1 : 0);
  r.sign = asign * bsign;
}

// r = mpf_mul(a, b); aw: the mantissa words w[0..2NL) of a (registers or
// memory), bw likewise in registers; both non-zero (their lowest limb is
// ignored, as GMP does).
// Default form: operand scanning by rows with IMAD.WIDE.U32.X carry chains
// (mul_rows_short), a's words read as the rows need them.  IMAD.WIDE issues at
// 32 lanes/clk/SM on sm_100a whether or not it carries (profiles/
// imad_rate4_r01.jsonl), so the carry-free radix-2^29 form (MPFW_MUL_RADIX29,
// 26 % more partial products) loses; it and the column form (MPFW_MUL_COLUMNS)
// are kept for A/B measurements only.
template <int NL, class APtr>
MPFW_D void mul(Reg<NL> &r, int32_t asign, int32_t aexp, APtr aw, int32_t bsign, int32_t bexp,
                const uint32_t (&bw)[2 * NL])
{
  typedef Mul29Geom<NL> G;
  uint32_t out[G::NOUT];
  bool rare;
#if defined(MPFW_MUL_RADIX29)
  {
    uint32_t bd[G::ND];
    digits29<G::W, G::ND>(bd, bw + 2);
    mul29_words<NL>(out, aw + 2, bd);
    rare = G::SHORT && out[0] == 0xFFFFFFFFu;
  }
#else
  {
    uint32_t p[MulGeom<NL>::NO], br[G::W];
#pragma unroll
    for(int i = 0; i < G::W; ++i)
      br[i] = bw[2 + i];
#if defined(MPFW_MUL_COLUMNS)
    uint32_t ar[G::W];
#pragma unroll
    for(int i = 0; i < G::W; ++i)
      ar[i] = aw[2 + i];
    mul_columns<G::W, MulGeom<NL>::C0, MulGeom<NL>::C0>(p, ar, br);
#else
    mul_rows_short<G::W, MulGeom<NL>::C0>(p, aw + 2, br);
#endif
#pragma unroll
    for(int i = 0; i < G::NOUT; ++i)
      out[i] = p[i + G::GW - MulGeom<NL>::C0];
    // columns below C0 add less than 2^8 units of word GW - 1 ... conservatively: guard word near overflow
    rare = G::SHORT && out[0] >= 0xFFFFFF00u;
  }
#endif
  if(rare)
    {
#ifdef MPFW_COUNT_RARE
      ++rare_mul_count;
#endif
      MulWords<NL> ca, cb;
#pragma unroll
      for(int i = 0; i < 2 * NL; ++i)
        {
          ca.w[i] = aw[i];
          cb.w[i] = bw[i];
        }
      const MulProd<NL> full = mul_full<NL>(ca, cb); // words from column MulGeom::C0 = GW - 1
#pragma unroll
      for(int i = 0; i < G::NOUT; ++i)
        out[i] = full.p[i + G::GW - MulGeom<NL>::C0];
    }
  mul_finish<NL>(r, out, asign, aexp, bsign, bexp);
}

// ------------------------------------------------------------------ add/sub
template <int NL>
MPFW_NOINLINE Reg<NL> addsub_generic(Reg<NL> acc, Reg<NL> v, int vsign)
{
  mpfx::Num<NL> a, b, r;
  to_num(a, acc);
  to_num(b, v);
  b.sign = vsign;
  mpfx::add(r, a, b);
  Reg<NL> out;
  from_num(out, r);
  return out;
}

// r = u + (s ^ mask) + cin over N words (mask = 0 / ~0 and cin = 0 / 1 give
// u + s and u - s); returns the carry out
template <int N>
MPFW_D uint32_t addx_n(uint32_t (&r)[N], const uint32_t (&u)[N], const uint32_t (&s)[N],
                       uint32_t mask)
{
#if defined(__CUDA_ARCH__)
  uint32_t dummy;
  asm volatile("add.cc.u32 %0, %1, 1;" : "=r"(dummy) : "r"(mask)); // CC = (mask == ~0)
#pragma unroll
  for(int i = 0; i < N; ++i)
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r[i]) : "r"(u[i]), "r"(s[i] ^ mask));
  uint32_t c;
  asm volatile("addc.u32 %0, 0, 0;" : "=r"(c));
  return c;
#else
  uint64_t c = mask & 1u;
  for(int i = 0; i < N; ++i)
    {
      const uint64_t t = (uint64_t)u[i] + (s[i] ^ mask) + c;
      r[i] = (uint32_t)t;
      c = t >> 32;
    }
  return (uint32_t)c;
#endif
}

// acc <- mpf_add(acc, v) with v's sign replaced by vsign (so the same routine
// serves mpf_sub).  One code path for both magnitude-add and magnitude-subtract
// (the lanes of a warp disagree on the sign all the time); only the rare
// post-processing steps branch.
template <int NL> MPFW_D void add_signed(Reg<NL> &acc, const Reg<NL> &v, int vsign)
{
  if(vsign == 0)
    return; // x + 0 = x exactly
  if(acc.sign == 0)
    {
      acc = v;
      acc.sign = vsign;
      return;
    }
  // u = operand with the larger exponent, s = the other one
  const bool swap = acc.exp < v.exp;
  const bool sub = acc.sign != vsign;
  const int32_t uexp = swap ? v.exp : acc.exp;
  const int32_t usign = swap ? vsign : acc.sign;
  const int64_t ediff64 = swap ? (int64_t)v.exp - acc.exp : (int64_t)acc.exp - v.exp;
  // add: s is ignored from ediff = NL-1 on; sub: from ediff = NL on (r = u):
  // s is cleared explicitly for ediff >= NL, smaller shifts go through the shifter
  const int ediff = ediff64 > NL ? NL : (int)ediff64;
  uint32_t u[2 * NL], s[2 * NL];
#pragma unroll
  for(int i = 0; i < 2 * NL; ++i)
    {
      u[i] = swap ? v.w[i] : acc.w[i];
      s[i] = swap ? acc.w[i] : v.w[i];
    }
  if(sub && ediff == 1)
    {
      // "extremely close" path of mpf_sub keeps v's lowest limb; rare
      const bool close = u[2 * NL - 1] == 0 && u[2 * NL - 2] == 1
                         && (NL < 2 || (u[2 * NL - 3] | u[2 * NL - 4]) == 0)
                         && s[2 * NL - 1] == 0xFFFFFFFFu && s[2 * NL - 2] == 0xFFFFFFFFu;
      if(close)
        {
#ifdef MPFW_COUNT_RARE
          ++rare_sub_count;
#endif
          acc = addsub_generic<NL>(acc, v, vsign);
          return;
        }
    }
  // align: s >>= ediff limbs (the two low stages inline, the high ones only when needed)
#pragma unroll
  for(int st = 1; st < NL && st <= 2; st <<= 1)
    {
      const bool on = (ediff & st) != 0;
#pragma unroll
      for(int i = 0; i < 2 * NL; ++i)
        {
          const uint32_t hi = (i + 2 * st < 2 * NL) ? s[i + 2 * st] : 0u;
          s[i] = on ? hi : s[i];
        }
    }
  if(ediff >= 4)
    {
      if(ediff64 >= NL)
        {
#pragma unroll
          for(int i = 0; i < 2 * NL; ++i)
            s[i] = 0;
        }
#pragma unroll
      for(int st = 4; st < NL; st <<= 1)
        {
          const bool on = (ediff & st) != 0;
#pragma unroll
          for(int i = 0; i < 2 * NL; ++i)
            {
              const uint32_t hi = (i + 2 * st < 2 * NL) ? s[i + 2 * st] : 0u;
              s[i] = on ? hi : s[i];
            }
        }
    }
  // mpf_add works on the top P limbs only: drop the lowest limb of both
  if(!sub)
    u[0] = u[1] = s[0] = s[1] = 0;
  uint32_t r[2 * NL];
  const uint32_t cy = addx_n<2 * NL>(r, u, s, sub ? 0xFFFFFFFFu : 0u);
  int32_t rexp = uexp, rsign = usign;
  if(!sub)
    {
      if(cy)
        {
          // a new top limb "1": everything moves down one limb
#pragma unroll
          for(int i = 0; i < 2 * NL - 2; ++i)
            r[i] = r[i + 2];
          r[2 * NL - 2] = 1u;
          r[2 * NL - 1] = 0u;
          rexp += 1;
        }
    }
  else
    {
      if(!cy)
        {
          // borrow: only possible for ediff == 0 with |s| > |u|
          neg_n<2 * NL>(r);
          rsign = -usign;
        }
      if((r[2 * NL - 1] | r[2 * NL - 2]) == 0)
        {
          // strip leading zero limbs
          int z = 0;
          bool run = true;
#pragma unroll
          for(int l = NL - 1; l >= 0; --l)
            {
              run = run && (r[2 * l] | r[2 * l + 1]) == 0;
              z += run ? 1 : 0;
            }
          if(z == NL)
            {
              set_zero(acc);
              return;
            }
          shl_limbs<NL>(r, z);
          rexp -= z;
        }
    }
#pragma unroll
  for(int i = 0; i < 2 * NL; ++i)
    acc.w[i] = r[i];
  acc.exp = rexp;
  acc.sign = rsign;
}

// acc <- acc + a*b (negate == false) or acc - a*b (negate == true), one
// mpf_mul followed by one mpf_add / mpf_sub.  a, b: packed elements as 32-bit
// words (shared or global memory).
// header and mantissa words of a packed element (16-byte aligned): 128-bit loads
template <int NL>
MPFW_D void load_packed(int32_t &exp, int32_t &sign, uint32_t (&w)[2 * NL], const uint32_t *p)
{
#if defined(__CUDA_ARCH__)
  constexpr int EW = 2 * ((NL + 2) & ~1);
  uint32_t t[EW];
  const uint4 *p4 = reinterpret_cast<const uint4 *>(p);
#pragma unroll
  for(int q = 0; q < EW / 4; ++q)
    {
      const uint4 x = p4[q];
      t[4 * q] = x.x;
      t[4 * q + 1] = x.y;
      t[4 * q + 2] = x.z;
      t[4 * q + 3] = x.w;
    }
  exp = (int32_t)t[0];
  sign = (int32_t)t[1];
#pragma unroll
  for(int i = 0; i < 2 * NL; ++i)
    w[i] = t[2 + i];
#else
  exp = (int32_t)p[0];
  sign = (int32_t)p[1];
  for(int i = 0; i < 2 * NL; ++i)
    w[i] = p[2 + i];
#endif
}
template <int NL>
MPFW_D void mac(Reg<NL> &acc, const uint32_t *a, const uint32_t *b, bool negate)
{
  int32_t bexp, bsign;
  uint32_t bw[2 * NL];
  load_packed<NL>(bexp, bsign, bw, b);
  const int32_t aexp = (int32_t)a[0], asign = (int32_t)a[1];
  if(asign == 0 || bsign == 0)
    return; // 0 * x = 0 and c + 0 = c exactly in mpf
  Reg<NL> p;
  mul<NL>(p, asign, aexp, a + 2, bsign, bexp, bw); // a's words are read from memory as the rows need them
  add_signed<NL>(acc, p, negate ? -p.sign : p.sign);
}
template <int NL>
MPFW_D void mac(Reg<NL> &acc, const Reg<NL> &a, const Reg<NL> &b, bool negate)
{
  if(a.sign == 0 || b.sign == 0)
    return;
  Reg<NL> p;
  const uint32_t *aw = a.w;
  mul<NL>(p, a.sign, a.exp, aw, b.sign, b.exp, b.w);
  add_signed<NL>(acc, p, negate ? -p.sign : p.sign);
}

// ----------------------------------------------------------------- division
// mpf_div(u, v) = floor(U * B^P / D) with U, D the NL-limb mantissas (GMP
// mpf/div.c on full-size operands, restated in mpfx::div).  When many numbers
// are divided by the same D (a Cholesky pivot, a column norm) the quotient is
// formed from a precomputed reciprocal
//     R = floor(beta^(4n+1) / D),   beta = 2^32, n = NL,   2n+4 words,
// as  q~ = floor(U R / beta^(2n+3))  (a short product of the high columns),
// which is q, q-1 or q-2; the exact remainder  U B^P - q~ D  (its low 2n+1
// words, a short product of the low columns) then fixes q~ up.  The result is
// the exact truncated quotient, i.e. bit-identical to mpf_div.
template <int NL> struct DivGeom
{
  static constexpr int RW = 2 * NL + 4; // words of R
};

// R = floor(beta^(4n+1) / D) by Knuth long division (slow, once per divisor)
template <int NL>
MPFW_D void reciprocal(uint32_t (&R)[2 * NL + 4], const mpfx::Num<NL> &d)
{
  constexpr int NN = 2 * NL + 2; // limbs of the numerator beta^(4n+1) = 2^32 * B^(2n)
  mpfx::limb_t np[NN], q[NN - NL + 1];
  for(int i = 0; i < NN; ++i)
    np[i] = 0;
  np[2 * NL] = (mpfx::limb_t)1 << 32;
  mpfx::n_div_q<2 * NN + 2>(q, np, 2 * NL + 1, d.d, NL);
  // quotient has (2NL+1) - NL + 1 = NL + 2 limbs = 2NL + 4 words
#pragma unroll
  for(int i = 0; i < NL + 2; ++i)
    {
      R[2 * i] = (uint32_t)q[i];
      R[2 * i + 1] = (uint32_t)(q[i] >> 32);
    }
}

// `x /= 4` of compute_schur_complement.cxx:102 (mpfx::div4): mantissa shifted
// right by two bits, a vanished top limb stripped
template <int NL> MPFW_D void div4(Reg<NL> &x)
{
  if(x.sign == 0)
    return;
  uint32_t q[2 * NL];
#pragma unroll
  for(int i = 0; i < 2 * NL; ++i)
    q[i] = (x.w[i] >> 2) | (i + 1 < 2 * NL ? (x.w[i + 1] << 30) : 0u);
  const bool adj = (q[2 * NL - 1] | q[2 * NL - 2]) == 0;
#pragma unroll
  for(int i = 2 * NL - 1; i >= 2; --i)
    x.w[i] = adj ? q[i - 2] : q[i];
  x.w[1] = adj ? 0u : q[1];
  x.w[0] = adj ? 0u : q[0];
  x.exp -= adj ? 1 : 0;
}

// ------------------------------------------------- fast reciprocal and sqrt
// Newton iterations on fixed-point fractions with statically sized word
// arrays, precision roughly doubling per level (w -> 2w-1 words, the lost word
// absorbs the accumulated error), finished by an EXACT correction against the
// remainder, so the results equal the Knuth / integer-Newton reference
// routines of mpfx bit for bit.  They replace those on the critical path of
// the Cholesky factorisations (one sqrt and one reciprocal per pivot).

// out[0..NOUT) = low NOUT words of a*b (exact)
template <int KA, int KB, int NOUT>
MPFW_D void mul_low(uint32_t (&out)[NOUT], const uint32_t (&a)[KA], const uint32_t (&b)[KB])
{
  uint32_t t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
  for(int c = 0; c < NOUT; ++c)
    {
#pragma unroll
      for(int i = 0; i < KA; ++i)
        {
          const int j = c - i;
          if(j >= 0 && j < KB)
            mac3(t0, t1, t2, a[i], b[j]);
        }
      out[c] = t0;
      t0 = t1;
      t1 = t2;
      t2 = 0;
    }
}
// out[c - FROM] = word c of a*b for c in [FROM, KA+KB), columns below FROM-2
// are not formed (the result may be low by a few units of word FROM-1)
template <int KA, int KB, int FROM>
MPFW_D void mul_high(uint32_t (&out)[KA + KB - FROM], const uint32_t (&a)[KA],
                     const uint32_t (&b)[KB])
{
  constexpr int C0 = FROM - 2 > 0 ? FROM - 2 : 0;
  uint32_t t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
  for(int c = C0; c < KA + KB - 1; ++c)
    {
#pragma unroll
      for(int i = 0; i < KA; ++i)
        {
          const int j = c - i;
          if(j >= 0 && j < KB)
            mac3(t0, t1, t2, a[i], b[j]);
        }
      if(c >= FROM)
        out[c - FROM] = t0;
      t0 = t1;
      t1 = t2;
      t2 = 0;
    }
  out[KA + KB - 1 - FROM] = t0;
}
// out[c - FROM] = word c of a*b for c in [FROM, TO); columns below FROM-2 are
// not formed, the carry out of column TO-1 is dropped
template <int KA, int KB, int FROM, int TO>
MPFW_D void mul_mid(uint32_t (&out)[TO - FROM], const uint32_t (&a)[KA], const uint32_t (&b)[KB])
{
  constexpr int C0 = FROM - 2 > 0 ? FROM - 2 : 0;
  uint32_t t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
  for(int c = C0; c < TO; ++c)
    {
#pragma unroll
      for(int i = 0; i < KA; ++i)
        {
          const int j = c - i;
          if(j >= 0 && j < KB)
            mac3(t0, t1, t2, a[i], b[j]);
        }
      if(c >= FROM)
        out[c - FROM] = t0;
      t0 = t1;
      t1 = t2;
      t2 = 0;
    }
}
// r -= small constant (borrow chain)
template <int N> MPFW_D void sub_small(uint32_t (&r)[N], uint32_t k)
{
  uint32_t bw = k;
#pragma unroll
  for(int i = 0; i < N; ++i)
    {
      const uint32_t a = r[i];
      r[i] = a - bw;
      bw = a < bw ? 1u : 0u;
    }
}
// v <<= sh bits, sh in [0, 32*N), zeros shifted in (static word stages + one bit stage)
template <int N> MPFW_D void shl_bits(uint32_t (&v)[N], int sh)
{
  const int ws = sh >> 5, bs = sh & 31;
#pragma unroll
  for(int s = 1; s < N; s <<= 1)
    {
      const bool on = (ws & s) != 0;
#pragma unroll
      for(int i = N - 1; i >= 0; --i)
        {
          const uint32_t lo = (i - s >= 0) ? v[i - s] : 0u;
          v[i] = on ? lo : v[i];
        }
    }
  if(bs)
    {
#pragma unroll
      for(int i = N - 1; i >= 0; --i)
        {
          const uint32_t lo = i > 0 ? v[i - 1] : 0u;
          v[i] = (v[i] << bs) | (lo >> (32 - bs));
        }
    }
}
template <int N> MPFW_D void shr_bits(uint32_t (&v)[N], int sh)
{
  const int ws = sh >> 5, bs = sh & 31;
#pragma unroll
  for(int s = 1; s < N; s <<= 1)
    {
      const bool on = (ws & s) != 0;
#pragma unroll
      for(int i = 0; i < N; ++i)
        {
          const uint32_t hi = (i + s < N) ? v[i + s] : 0u;
          v[i] = on ? hi : v[i];
        }
    }
  if(bs)
    {
#pragma unroll
      for(int i = 0; i < N; ++i)
        {
          const uint32_t hi = i + 1 < N ? v[i + 1] : 0u;
          v[i] = (v[i] >> bs) | (hi << (32 - bs));
        }
    }
}
template <int N> MPFW_D int clz_words(const uint32_t (&v)[N])
{
  int z = 0;
  bool run = true;
#pragma unroll
  for(int i = N - 1; i >= 0; --i)
    {
      const int c = mpfx::clz32(v[i]); // 32 for a zero word
      z += run ? c : 0;
      run = run && v[i] == 0;
    }
  return z;
}

// Y (W words) ~ beta^W / (2 d), d = Dn / beta^ND in [1/2, 1); 0 < Y/beta^W < 1/(2d),
// short by at most a few dozen units.
template <int ND, int W> struct RecipLevel
{
  static constexpr int WP = (W <= 2) ? 1 : (W + 2) / 2;
  MPFW_D static void run(uint32_t (&Y)[W], const uint32_t (&Dn)[ND])
  {
    if constexpr(W == 1)
      {
        // floor(2^63 / (top word + 1)): below 2^31/d_top by < 2 units of 2^-32
        const uint64_t dt = (uint64_t)Dn[ND - 1] + 1;
        uint64_t q = ((uint64_t)1 << 63) / dt;
        if(q > 0xFFFFFFFFull)
          q = 0xFFFFFFFFull;
        Y[0] = (uint32_t)q - 1u;
      }
    else
      {
        uint32_t Yp[WP];
        RecipLevel<ND, WP>::run(Yp, Dn);
        uint32_t Dt[W];
#pragma unroll
        for(int i = 0; i < W; ++i)
          Dt[i] = (ND - W + i >= 0) ? Dn[ND - W + i] : 0u;
        // G = beta^(W+WP) - 2 Dt Yp, low W+1 words (the rest is zero)
        uint32_t G[W + 1];
        mul_low<W, WP, W + 1>(G, Dt, Yp);
#pragma unroll
        for(int i = W; i >= 0; --i)
          G[i] = (G[i] << 1) | (i > 0 ? (G[i - 1] >> 31) : 0u);
        neg_n<W + 1>(G);
        // Y = Yp beta^(W-WP) + floor(Yp G / beta^(2 WP))
        uint32_t Q[WP + W + 1 - 2 * WP];
        mul_high<WP, W + 1, 2 * WP>(Q, Yp, G);
        uint32_t Ye[W + 1], Qe[W + 1];
#pragma unroll
        for(int i = 0; i <= W; ++i)
          {
            Ye[i] = (i >= W - WP && i < W) ? Yp[i - (W - WP)] : 0u;
            Qe[i] = (i < W - WP + 1) ? Q[i] : 0u;
          }
        uint32_t S[W + 1];
        add_n<W + 1>(S, Ye, Qe);
#pragma unroll
        for(int i = 0; i < W; ++i)
          Y[i] = S[i];
        if(S[W]) // cannot happen for a convergent iterate; saturate
          {
#pragma unroll
            for(int i = 0; i < W; ++i)
              Y[i] = 0xFFFFFFFFu;
          }
        sub_small<W>(Y, 4u);
      }
  }
};

// R = floor(beta^(4n+1) / D), same value as reciprocal() above
template <int NL>
MPFW_D void reciprocal_fast(uint32_t (&R)[2 * NL + 4], const Reg<NL> &d)
{
  constexpr int n2 = 2 * NL, WF = 2 * NL + 5;
  uint32_t Dn[n2];
#pragma unroll
  for(int i = 0; i < n2; ++i)
    Dn[i] = d.w[i];
  const int s = clz_words<n2>(Dn); // < 64: the top limb is non-zero
  shl_bits<n2>(Dn, s);
  uint32_t Y[WF];
  RecipLevel<n2, WF>::run(Y, Dn);
  // R~ = floor(Y 2^(s+1) / beta^4) = Y >> (127 - s) bits
  shr_bits<WF>(Y, 127 - s);
  uint32_t Rt[n2 + 4];
#pragma unroll
  for(int i = 0; i < n2 + 4; ++i)
    Rt[i] = Y[i];
  // exact remainder beta^(4n+1) - R~ D, low 2n+2 words
  uint32_t rem[n2 + 2];
  {
    uint32_t dw[n2];
#pragma unroll
    for(int i = 0; i < n2; ++i)
      dw[i] = d.w[i];
    mul_low<n2 + 4, n2, n2 + 2>(rem, Rt, dw);
    neg_n<n2 + 2>(rem); // beta^(4n+1) = 0 mod beta^(2n+2)
  }
  uint32_t dext[n2 + 2];
#pragma unroll
  for(int i = 0; i < n2 + 2; ++i)
    dext[i] = i < n2 ? d.w[i] : 0u;
  bool ok = false;
#pragma unroll
  for(int round = 0; round < 4; ++round)
    {
      uint32_t tmp[n2 + 2];
      const uint32_t bw = sub_n<n2 + 2>(tmp, rem, dext);
      const bool ge = bw == 0 && !ok;
      ok = ok || bw != 0;
#pragma unroll
      for(int c = 0; c < n2 + 2; ++c)
        rem[c] = ge ? tmp[c] : rem[c];
      uint32_t carry = ge ? 1u : 0u;
#pragma unroll
      for(int c = 0; c < n2 + 4; ++c)
        {
          const uint32_t t = Rt[c] + carry;
          carry = (t < carry) ? 1u : 0u;
          Rt[c] = t;
        }
    }
  if(!ok)
    {
      // the iterate was further off than the analysis allows: exact slow path
#ifdef MPFW_COUNT_RARE
      ++recip_fallbacks;
#endif
      mpfx::Num<NL> dn;
      to_num(dn, d);
      reciprocal<NL>(R, dn);
      return;
    }
#pragma unroll
  for(int i = 0; i < n2 + 4; ++i)
    R[i] = Rt[i];
}

// V (W words) ~ beta^W / (2 sqrt(t)), t = Tn / beta^NT in [1/4, 1), from below
template <int NT, int W> struct RsqrtLevel
{
  static constexpr int WP = (W <= 2) ? 1 : (W + 2) / 2;
  MPFW_D static void run(uint32_t (&V)[W], const uint32_t (&Tn)[NT])
  {
    if constexpr(W == 1)
      {
        // 32-bit estimate from the top 64 bits in double precision, then lowered
        const double t = ((double)Tn[NT - 1] * 4294967296.0 + (double)Tn[NT - 2])
                         / 18446744073709551616.0;
        double v = 0.5 / ::sqrt(t);
        v = v * 4294967296.0 - 8.0;
        if(v > 4294967295.0)
          v = 4294967295.0;
        V[0] = (uint32_t)v;
      }
    else
      {
        uint32_t Vp[WP];
        RsqrtLevel<NT, WP>::run(Vp, Tn);
        uint32_t Tt[W];
#pragma unroll
        for(int i = 0; i < W; ++i)
          Tt[i] = (NT - W + i >= 0) ? Tn[NT - W + i] : 0u;
        // F = Tt Vp^2 ; G = beta^(2WP+W) - 4F < beta^(WP+W+1).  Only G's words
        // from GL up matter for the quotient below, so only that slice of F is formed.
        uint32_t V2[2 * WP];
        mul_low<WP, WP, 2 * WP>(V2, Vp, Vp);
        constexpr int GL = (2 * WP - 3 > 0) ? 2 * WP - 3 : 0;
        constexpr int NG = WP + W + 1 - GL;
        uint32_t G[NG];
        mul_mid<2 * WP, W, GL, WP + W + 1>(G, V2, Tt);
#pragma unroll
        for(int i = NG - 1; i >= 0; --i)
          G[i] = (G[i] << 2) | (i > 0 ? (G[i - 1] >> 30) : 0u);
        neg_n<NG>(G);
        // V = Vp beta^(W-WP) + floor(Vp G / (2 beta^(3 WP)))
        constexpr int QF = 3 * WP - GL - 1; // keep one extra low word for the halving
        constexpr int NQ = WP + NG - QF;
        uint32_t Q[NQ];
        mul_high<WP, NG, QF>(Q, Vp, G);
        shr_bits<NQ>(Q, 1);
        uint32_t Ve[W + 1], Qe[W + 1];
#pragma unroll
        for(int i = 0; i <= W; ++i)
          {
            Ve[i] = (i >= W - WP && i < W) ? Vp[i - (W - WP)] : 0u;
            Qe[i] = (i + 1 < NQ) ? Q[i + 1] : 0u;
          }
        uint32_t S[W + 1];
        add_n<W + 1>(S, Ve, Qe);
#pragma unroll
        for(int i = 0; i < W; ++i)
          V[i] = S[i];
        if(S[W])
          {
#pragma unroll
            for(int i = 0; i < W; ++i)
              V[i] = 0xFFFFFFFFu;
          }
        sub_small<W>(V, 4u);
      }
  }
};

// r = mpf_sqrt(u), u > 0: same value as mpfx::sqrt
template <int NL> MPFW_D void sqrt_fast(Reg<NL> &r, const Reg<NL> &u)
{
  constexpr int P = NL - 1, NT = 4 * P, NR = 2 * P, WF = 2 * P + 3;
  const int expodd = u.exp & 1;
  // T: u's NL limbs top-aligned in 2P - expodd limbs, viewed in a 2P-limb frame
  uint32_t T[NT];
#pragma unroll
  for(int i = 0; i < NT; ++i)
    {
      // expodd == 0: T word i = u.w[i - (NT - 2NL)]; expodd == 1: two words lower
      const int a = i - (NT - 2 * NL), b = i - (NT - 2 * NL) + 2;
      const uint32_t w0 = (a >= 0 && a < 2 * NL) ? u.w[a] : 0u;
      const uint32_t w1 = (b >= 0 && b < 2 * NL) ? u.w[b] : 0u;
      T[i] = expodd ? w1 : w0;
    }
  uint32_t Tn[NT];
#pragma unroll
  for(int i = 0; i < NT; ++i)
    Tn[i] = T[i];
  const int s = clz_words<NT>(Tn) & ~1; // even, < 128
  shl_bits<NT>(Tn, s);
  uint32_t V[WF];
  RsqrtLevel<NT, WF>::run(V, Tn);
  // sqrt(Tn) = 2 Tn V / (beta^(NT/2) beta^WF): top NR words of the product, then >> s/2
  // use the top NR+3 words of Tn
  uint32_t Tt[NR + 3];
#pragma unroll
  for(int i = 0; i < NR + 3; ++i)
    Tt[i] = Tn[NT - (NR + 3) + i];
  // Tn V / beta^(NT/2 + WF) = Tt V / beta^(NR + 3 - NT/2 ... ) : NT/2 = NR, so
  // Tn ~ Tt beta^(NT-NR-3) and the quotient is Tt V / beta^(WF + 3 - (NT - 2 NR)) = Tt V / beta^(WF+3)
  uint32_t Sx[NR + 3 + WF - (WF + 2)];
  mul_high<NR + 3, WF, WF + 2>(Sx, Tt, V);
  // Sx = floor(Tt V / beta^(WF+2)) = 2*beta * sqrt(Tn)/2 ... : sqrt(Tn) = 2 Tt V / beta^(WF+3)
  // so sqrt(Tn) = Sx * 2 / beta; and sqrt(T) = sqrt(Tn) >> (s/2)
  // S~ = (Sx >> (31 + s/2)) : Sx has NR+1 words
  shr_bits<NR + 1>(Sx, 31 + (s >> 1));
  uint32_t S[NR];
#pragma unroll
  for(int i = 0; i < NR; ++i)
    S[i] = Sx[i];
  // keep S~ <= floor(sqrt(T)): lower it by a few units, the correction walks up
  sub_small<NR>(S, 2u);
  // rem = T - S^2, low NR+2 words
  uint32_t rem[NR + 2];
  {
    uint32_t sq[NR + 2];
    mul_low<NR, NR, NR + 2>(sq, S, S);
    uint32_t tl[NR + 2];
#pragma unroll
    for(int i = 0; i < NR + 2; ++i)
      tl[i] = T[i];
    sub_n<NR + 2>(rem, tl, sq);
  }
  bool ok = false;
#pragma unroll
  for(int round = 0; round < 8; ++round)
    {
      // (S+1)^2 <= T  <=>  rem >= 2S + 1
      uint32_t step[NR + 2];
#pragma unroll
      for(int i = 0; i < NR + 2; ++i)
        {
          const uint32_t lo = (i < NR) ? S[i] : 0u;
          const uint32_t below = (i > 0 && i - 1 < NR) ? S[i - 1] : 0u;
          step[i] = (lo << 1) | (below >> 31);
        }
      step[0] |= 1u;
      uint32_t tmp[NR + 2];
      const uint32_t bw = sub_n<NR + 2>(tmp, rem, step);
      const bool ge = bw == 0 && !ok;
      ok = ok || bw != 0;
#pragma unroll
      for(int i = 0; i < NR + 2; ++i)
        rem[i] = ge ? tmp[i] : rem[i];
      uint32_t carry = ge ? 1u : 0u;
#pragma unroll
      for(int i = 0; i < NR; ++i)
        {
          const uint32_t t = S[i] + carry;
          carry = (t < carry) ? 1u : 0u;
          S[i] = t;
        }
    }
  // rem must now be in [0, 2S]; a negative start (S~ too large) shows as a huge rem
  if(!ok || rem[NR + 1] != 0)
    {
#ifdef MPFW_COUNT_RARE
      ++sqrt_fallbacks;
#endif
      mpfx::Num<NL> a, b;
      to_num(a, u);
      mpfx::sqrt(b, a);
      from_num(r, b);
      return;
    }
  r.w[0] = r.w[1] = 0;
#pragma unroll
  for(int i = 0; i < NR; ++i)
    r.w[i + 2] = S[i];
  r.sign = 1;
  r.exp = (u.exp + expodd) / 2;
}

// x <- mpf_div(x, d);  dw = mantissa words of d (2n), R = reciprocal(d).
template <int NL>
MPFW_D void div_recip(Reg<NL> &x, int32_t dsign, int32_t dexp, const uint32_t *dw,
                      const uint32_t *R)
{
  if(x.sign == 0)
    return;
  constexpr int n2 = 2 * NL, RW = 2 * NL + 4;
  // ---- q~ = words [2n+3, 4n+3) of U*R, columns >= 2n+1 only ----
  uint32_t q[n2];
  uint32_t guard = 0; // word n2+2 of the column sum: the first fraction word below q~
  {
    uint32_t t0 = 0, t1 = 0, t2 = 0;
    constexpr int C0 = n2 + 1;
#pragma unroll
    for(int c = C0; c < n2 + RW - 1; ++c)
      {
#pragma unroll
        for(int i = 0; i < n2; ++i)
          {
            const int j = c - i;
            if(j >= 0 && j < RW)
              mac3(t0, t1, t2, x.w[i], R[j]);
          }
        if(c == n2 + 2)
          guard = t0;
        if(c >= n2 + 3 && c - (n2 + 3) < n2)
          q[c - (n2 + 3)] = t0;
        t0 = t1;
        t1 = t2;
        t2 = 0;
      }
    // column n2+RW-1 = 4n+3 is word 2n of q~: always zero (q < beta^(2n))
  }
  // With W = the columns >= 2n+1 that were formed, Q = U beta^(2n-2) / D (real) and R = floor(
  // beta^(4n+1) / D):  W / beta^(2n+3) <= Q < W / beta^(2n+3) + 2n / beta + beta^-3  (the dropped
  // columns hold fewer than 2n beta^(2n+2); R's own floor costs less than beta^-3), and
  // W / beta^(2n+3) < q~ + (guard + 1) / beta.  So guard + 2n + 2 <= beta already proves
  // floor(Q) = q~: the exact remainder (a second half product) is formed only on the 2^-26 of
  // the inputs whose fraction sits within 64 units of a whole number (exact quotients among them).
  static_assert(2 * NL + 2 <= 64, "the guard-word test of div_recip assumes 2n + 2 <= 64");
#ifdef MPFW_COUNT_RARE
  ++(guard > 0xFFFFFFFFu - 64u ? div_exact_count : div_fast_count);
#endif
  if(guard > 0xFFFFFFFFu - 64u)
  {
  // ---- rem = (U beta^(2n-2) - q~ D) mod beta^(2n+1) ----
  uint32_t rem[n2 + 1];
  {
    uint32_t t0 = 0, t1 = 0, t2 = 0;
#pragma unroll
    for(int c = 0; c <= n2; ++c)
      {
#pragma unroll
        for(int i = 0; i < n2; ++i)
          {
            const int j = c - i;
            if(j >= 0 && j < n2)
              mac3(t0, t1, t2, q[i], dw[j]);
          }
        rem[c] = t0;
        t0 = t1;
        t1 = t2;
        t2 = 0;
      }
    uint32_t ulow[n2 + 1];
#pragma unroll
    for(int c = 0; c <= n2; ++c)
      ulow[c] = (c >= n2 - 2) ? x.w[c - (n2 - 2)] : 0u;
    uint32_t tmp[n2 + 1];
    sub_n<n2 + 1>(tmp, ulow, rem);
#pragma unroll
    for(int c = 0; c <= n2; ++c)
      rem[c] = tmp[c];
  }
  // ---- at most two corrections: while rem >= D: rem -= D, q~ += 1 ----
  uint32_t dext[n2 + 1];
#pragma unroll
  for(int c = 0; c < n2; ++c)
    dext[c] = dw[c];
  dext[n2] = 0;
#pragma unroll
  for(int round = 0; round < 2; ++round)
    {
      uint32_t tmp[n2 + 1];
      const uint32_t bw = sub_n<n2 + 1>(tmp, rem, dext);
      const bool ge = bw == 0;
#pragma unroll
      for(int c = 0; c <= n2; ++c)
        rem[c] = ge ? tmp[c] : rem[c];
      // q += ge
      uint32_t carry = ge ? 1u : 0u;
#pragma unroll
      for(int c = 0; c < n2; ++c)
        {
          const uint32_t s = q[c] + carry;
          carry = (s < carry) ? 1u : 0u;
          q[c] = s;
        }
    }
  }
  // ---- assemble (mpfx::div): strip one leading zero limb ----
  const bool adj = (q[n2 - 1] | q[n2 - 2]) == 0;
#pragma unroll
  for(int i = n2 - 1; i >= 2; --i)
    x.w[i] = adj ? q[i - 2] : q[i];
  x.w[1] = adj ? 0u : q[1];
  x.w[0] = adj ? 0u : q[0];
  x.exp = x.exp - dexp + (adj ? 0 : 1);
  x.sign = x.sign * dsign;
}
} // namespace mpfw

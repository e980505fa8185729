// mpfx — fixed-limb binary floating point that reproduces GMP `mpf_*` results
// bit for bit, usable from host and from sm_100a device code.
//
// Why: the reference's scalar type El::BigFloat wraps GMP's mpf_t (reference:
// src/sdp_solve/SDP_Solver/run/bigint_syrk/fmpz/fmpz_BigFloat_convert.hxx:9,13;
// src/sdpb_util/Boost_Float.cxx:22-29).  mpf arithmetic is limb-granular and
// truncating.  For a working precision of `prec_bits` GMP keeps
// P = (prec_bits+63)/64 + 1 "precision limbs" and stores up to NL = P+1 limbs.
//
// Representation used here (and in HBM): every number owns exactly NL 64-bit
// limbs, little-endian, TOP-ALIGNED (d[NL-1] != 0 unless the number is zero)
// and zero padded at the bottom, plus a sign in {-1,0,+1} and an exponent
// counted in limbs:  value = sign * sum_i d[i] * B^(i - NL + exp),  B = 2^64.
// Every mpf operation is a function of the operand VALUES only (trailing zero
// limbs never change a result), so an mpf_t with _mp_size < NL maps to this
// format by padding, and results compare equal limb for limb after the same
// padding.  tests/test_mpfx_vs_gmp.py fuzzes every routine below against the
// real libgmp on variable-size operands, including the cancellation paths.
//
// The routines restate the published GMP algorithms (mpf/mul.c, add.c, sub.c,
// div.c, sqrt.c, mul_2exp.c, div_2exp.c of GMP 6.x) on this fixed format.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MPFX_HD __host__ __device__ inline
#else
#define MPFX_HD inline
#endif

namespace mpfx
{
typedef uint64_t limb_t;

// number of limbs stored for a given --precision (reference formula restated in
// test/src/integration_tests/util/Float.cxx:36-40)
MPFX_HD int prec_limbs(int prec_bits) { return (prec_bits + 63) / 64 + 1; }
MPFX_HD int stored_limbs(int prec_bits) { return prec_limbs(prec_bits) + 1; }
// element stride in 64-bit words inside HBM / the C-ABI: 1 header word + NL
// limbs, rounded up to an even count so every element is 16-byte aligned.
MPFX_HD int elem_words(int nl) { return (nl + 1 + 1) & ~1; }

template <int NL> struct Num
{
  int32_t sign; // -1, 0, +1
  int32_t exp;  // in limbs
  limb_t d[NL];
};

// ---------------------------------------------------------------- primitives
MPFX_HD void mul64(uint64_t a, uint64_t b, uint64_t &hi, uint64_t &lo)
{
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umul64hi(a, b);
#else
  unsigned __int128 p = (unsigned __int128)a * b;
  lo = (uint64_t)p;
  hi = (uint64_t)(p >> 64);
#endif
}
MPFX_HD int clz64(uint64_t x)
{
#if defined(__CUDA_ARCH__)
  return __clzll((long long)x);
#else
  return x ? __builtin_clzll(x) : 64;
#endif
}
MPFX_HD int clz32(uint32_t x)
{
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}

// r = u + v over n limbs, returns carry
MPFX_HD limb_t n_add_n(limb_t *rp, const limb_t *up, const limb_t *vp, int n)
{
  limb_t cy = 0;
  for(int i = 0; i < n; ++i)
    {
      limb_t a = up[i], s = a + vp[i];
      limb_t c1 = s < a;
      limb_t t = s + cy;
      limb_t c2 = t < s;
      rp[i] = t;
      cy = c1 | c2;
    }
  return cy;
}
// r = u - v - bin over n limbs, returns borrow
MPFX_HD limb_t n_sub_nc(limb_t *rp, const limb_t *up, const limb_t *vp, int n,
                        limb_t bin)
{
  limb_t bw = bin;
  for(int i = 0; i < n; ++i)
    {
      limb_t a = up[i], b = vp[i];
      limb_t s = a - b;
      limb_t b1 = a < b;
      limb_t t = s - bw;
      limb_t b2 = s < bw;
      rp[i] = t;
      bw = b1 | b2;
    }
  return bw;
}
// r = u - b over n limbs, returns borrow
MPFX_HD limb_t n_sub_1(limb_t *rp, const limb_t *up, int n, limb_t b)
{
  limb_t bw = b;
  for(int i = 0; i < n; ++i)
    {
      limb_t a = up[i];
      rp[i] = a - bw;
      bw = a < bw;
    }
  return bw;
}
// r = u + b over n limbs, returns carry
MPFX_HD limb_t n_add_1(limb_t *rp, const limb_t *up, int n, limb_t b)
{
  limb_t cy = b;
  for(int i = 0; i < n; ++i)
    {
      limb_t s = up[i] + cy;
      cy = s < cy;
      rp[i] = s;
    }
  return cy;
}
// r = -u over n limbs (two's complement), returns 1 iff u != 0
MPFX_HD limb_t n_neg(limb_t *rp, const limb_t *up, int n)
{
  limb_t bw = 0;
  for(int i = 0; i < n; ++i)
    {
      limb_t a = up[i];
      limb_t t = (limb_t)0 - a - bw;
      bw = (a != 0) | bw;
      rp[i] = t;
    }
  return bw;
}
// r[0..un) = u[0..un) - v[0..vn), un >= vn, returns borrow
MPFX_HD limb_t n_sub(limb_t *rp, const limb_t *up, int un, const limb_t *vp,
                     int vn)
{
  limb_t bw = n_sub_nc(rp, up, vp, vn, 0);
  return n_sub_1(rp + vn, up + vn, un - vn, bw);
}
MPFX_HD int n_cmp(const limb_t *up, const limb_t *vp, int n)
{
  for(int i = n - 1; i >= 0; --i)
    {
      if(up[i] != vp[i])
        return up[i] > vp[i] ? 1 : -1;
    }
  return 0;
}

template <int NL> MPFX_HD void set_zero(Num<NL> &r)
{
  r.sign = 0;
  r.exp = 0;
  for(int i = 0; i < NL; ++i)
    r.d[i] = 0;
}

// place `rsize` limbs tp[0..rsize) (tp[rsize-1] != 0) top-aligned into r
template <int NL>
MPFX_HD void place_top(Num<NL> &r, const limb_t *tp, int rsize)
{
  // rsize <= NL always
  limb_t out[NL];
  const int shift = NL - rsize;
  for(int i = 0; i < NL; ++i)
    out[i] = (i >= shift) ? tp[i - shift] : 0;
  for(int i = 0; i < NL; ++i)
    r.d[i] = out[i];
}

// ----------------------------------------------------------------------- mul
// GMP mpf_mul: both operands cut to their top P limbs, exact product, one
// possible leading zero limb stripped, top P+1 limbs kept.
template <int NL>
MPFX_HD void mul(Num<NL> &r, const Num<NL> &u, const Num<NL> &v)
{
  constexpr int P = NL - 1;
  if(u.sign == 0 || v.sign == 0)
    {
      set_zero(r);
      return;
    }
  limb_t t[2 * P];
#pragma unroll
  for(int i = 0; i < 2 * P; ++i)
    t[i] = 0;
#pragma unroll
  for(int i = 0; i < P; ++i)
    {
      const limb_t a = u.d[i + 1];
      limb_t carry = 0;
#pragma unroll
      for(int j = 0; j < P; ++j)
        {
          limb_t hi, lo;
          mul64(a, v.d[j + 1], hi, lo);
          lo += carry;
          hi += (lo < carry);
          limb_t s = t[i + j] + lo;
          hi += (s < lo);
          t[i + j] = s;
          carry = hi;
        }
      t[i + P] = carry;
    }
  const int adj = (t[2 * P - 1] == 0);
  const int32_t e = u.exp + v.exp - adj;
  const int32_t s = u.sign * v.sign;
  // top NL limbs of the (2P - adj)-limb product; 2P-1 >= NL needs P >= 2
  if(adj)
    {
#pragma unroll
      for(int i = 0; i < NL; ++i)
        r.d[i] = t[2 * P - 1 - NL + i];
    }
  else
    {
#pragma unroll
      for(int i = 0; i < NL; ++i)
        r.d[i] = t[2 * P - NL + i];
    }
  r.exp = e;
  r.sign = s;
}

// ------------------------------------------------------------------- add/sub
// |u| + |v| with the sign of u; GMP mpf_add same-sign path: window of P limbs
// below the top of the operand with the larger exponent.
template <int NL>
MPFX_HD void add_mag(Num<NL> &r, const Num<NL> &u_in, const Num<NL> &v_in)
{
  constexpr int P = NL - 1;
  const int32_t sgn = u_in.sign;
  const Num<NL> *u = &u_in, *v = &v_in;
  if(u->exp < v->exp)
    {
      const Num<NL> *t = u;
      u = v;
      v = t;
    }
  const int64_t ediff = (int64_t)u->exp - (int64_t)v->exp;
  limb_t s[NL]; // s[1..NL-1] = window sum
  limb_t cy = 0;
  for(int i = 1; i < NL; ++i)
    {
      const int64_t k = i + ediff;
      const limb_t b = (k < NL) ? v->d[k] : 0;
      const limb_t a = u->d[i];
      limb_t x = a + b;
      limb_t c1 = x < a;
      limb_t y = x + cy;
      limb_t c2 = y < x;
      s[i] = y;
      cy = c1 | c2;
    }
  const int32_t e = u->exp;
  if(cy)
    {
      for(int i = 0; i < P; ++i)
        r.d[i] = s[i + 1];
      r.d[P] = 1;
      r.exp = e + 1;
    }
  else
    {
      r.d[0] = 0;
      for(int i = 1; i < NL; ++i)
        r.d[i] = s[i];
      r.exp = e;
    }
  r.sign = sgn;
}

// |u| - |v| times the sign of u; GMP mpf_sub same-sign path restated on
// operands of full size NL (prec+1 limbs), including the close-operand paths
// that keep extra low limbs under cancellation.
template <int NL>
MPFX_HD void sub_mag(Num<NL> &r, const Num<NL> &u_in, const Num<NL> &v_in)
{
  constexpr int prec = NL; // mpf_sub works with PREC(r)+1
  int negate = u_in.sign < 0;
  const Num<NL> *u = &u_in, *v = &v_in;
  if(u->exp < v->exp)
    {
      const Num<NL> *t = u;
      u = v;
      v = t;
      negate ^= 1;
    }
  const limb_t *up = u->d, *vp = v->d;
  int usize = NL, vsize = NL;
  int64_t exp = u->exp;
  const int64_t ediff = exp - (int64_t)v->exp;
  limb_t tp[NL + 1];
  int rsize = 0;
  bool have_result = false; // result already in tp[0..rsize), normalised
  bool need_normalize = false;

  if(ediff <= 1)
    {
      bool general = false;
      if(ediff == 0)
        {
          // skip leading limbs that are equal
          bool cancelled = false;
          while(up[usize - 1] == vp[vsize - 1])
            {
              usize--;
              vsize--;
              exp--;
              if(usize == 0)
                {
                  // sizes are equal here, so v is exhausted as well: u == v
                  cancelled = true;
                  break;
                }
            }
          if(cancelled)
            {
              set_zero(r);
              return;
            }
          if(up[usize - 1] < vp[vsize - 1])
            {
              const limb_t *t = up;
              up = vp;
              vp = t;
              negate ^= 1;
            }
          if(up[usize - 1] != vp[vsize - 1] + 1)
            general = true;
          else
            {
              usize--;
              vsize--;
              exp--;
            }
        }
      else // ediff == 1
        {
          if(up[usize - 1] != 1 || vp[vsize - 1] != ~(limb_t)0
             || (usize >= 2 && up[usize - 2] != 0))
            general = true;
          else
            {
              usize--;
              exp--;
            }
        }
      if(!general)
        {
          // skip sequences of 00000000/ffffffff
          while(vsize != 0 && usize != 0 && up[usize - 1] == 0
                && vp[vsize - 1] == ~(limb_t)0)
            {
              usize--;
              vsize--;
              exp--;
            }
          if(usize == 0)
            {
              while(vsize != 0 && vp[vsize - 1] == ~(limb_t)0)
                {
                  vsize--;
                  exp--;
                }
            }
          else if(usize > prec - 1)
            {
              up += usize - (prec - 1);
              usize = prec - 1;
            }
          if(vsize > prec - 1)
            {
              vp += vsize - (prec - 1);
              vsize = prec - 1;
            }
          limb_t cy_limb;
          if(vsize == 0)
            {
              for(int i = 0; i < usize; ++i)
                tp[i] = up[i];
              tp[usize] = 1;
              rsize = usize + 1;
              exp++;
              have_result = true;
            }
          else
            {
              if(usize == 0)
                {
                  cy_limb = n_neg(tp, vp, vsize);
                  rsize = vsize;
                }
              else if(usize >= vsize)
                {
                  const int size = usize - vsize;
                  for(int i = 0; i < size; ++i)
                    tp[i] = up[i];
                  cy_limb = n_sub_nc(tp + size, up + size, vp, vsize, 0);
                  rsize = usize;
                }
              else
                {
                  const int size = vsize - usize;
                  cy_limb = n_neg(tp, vp, size);
                  cy_limb = n_sub_nc(tp + size, up, vp + size, usize, cy_limb);
                  rsize = vsize;
                }
              if(cy_limb == 0)
                {
                  tp[rsize] = 1;
                  rsize++;
                  exp++;
                  have_result = true;
                }
              else
                need_normalize = true;
            }
        }
    }

  if(!have_result && !need_normalize)
    {
      // general case
      if(usize > prec)
        {
          up += usize - prec;
          usize = prec;
        }
      if(vsize + ediff > prec)
        {
          vp += vsize + ediff - prec;
          vsize = (int)(prec - ediff); // may be <= 0
        }
      if(ediff >= prec)
        {
          // v completely below the window: r = u (cut to prec limbs)
          for(int i = 0; i < usize; ++i)
            tp[i] = up[i];
          rsize = usize;
          have_result = true;
        }
      else
        {
          // exact difference of the aligned operands; u > v is guaranteed.
          // u occupies [0,usize) of a frame whose top is usize; v's top sits
          // ediff limbs lower.  Let the frame bottom be the lower of the two.
          const int vtop = usize - (int)ediff; // index one past v's top limb
          const int vbot = vtop - vsize;       // may be negative
          const int low = vbot < 0 ? -vbot : 0; // limbs of v below u's bottom
          rsize = usize + low;
          limb_t bw = 0;
          for(int i = 0; i < rsize; ++i)
            {
              const int ui = i - low; // index into up
              const int vi = i - low - vbot; // index into vp
              const limb_t a = (ui >= 0 && ui < usize) ? up[ui] : 0;
              const limb_t b = (vi >= 0 && vi < vsize) ? vp[vi] : 0;
              limb_t s = a - b;
              limb_t b1 = a < b;
              limb_t t = s - bw;
              limb_t b2 = s < bw;
              tp[i] = t;
              bw = b1 | b2;
            }
          need_normalize = true;
        }
    }
  if(need_normalize)
    {
      while(rsize != 0 && tp[rsize - 1] == 0)
        {
          rsize--;
          exp--;
        }
    }
  if(rsize == 0)
    {
      set_zero(r);
      return;
    }
  place_top(r, tp, rsize);
  r.sign = negate ? -1 : 1;
  r.exp = (int32_t)exp;
}

template <int NL>
MPFX_HD void add(Num<NL> &r, const Num<NL> &u, const Num<NL> &v)
{
  if(u.sign == 0)
    {
      r = v;
      return;
    }
  if(v.sign == 0)
    {
      r = u;
      return;
    }
  if(u.sign == v.sign)
    add_mag(r, u, v);
  else
    {
      // u + v = u - (-v): same-sign subtraction of magnitudes, sign of u
      Num<NL> nv = v;
      nv.sign = -nv.sign;
      sub_mag(r, u, nv);
    }
}

template <int NL>
MPFX_HD void sub(Num<NL> &r, const Num<NL> &u, const Num<NL> &v)
{
  if(u.sign == 0)
    {
      r = v;
      r.sign = -r.sign;
      return;
    }
  if(v.sign == 0)
    {
      r = u;
      return;
    }
  if(u.sign == v.sign)
    sub_mag(r, u, v);
  else
    {
      Num<NL> nv = v;
      nv.sign = -nv.sign;
      add_mag(r, u, nv);
    }
}

// ------------------------------------------------------------------ division
// Quotient of multi-limb integers, floor(N / D).  np has nn limbs, dp has dn
// limbs with dp[dn-1] != 0, nn >= dn.  Writes nn-dn+1 limbs to qp.  Knuth
// algorithm D on 32-bit digits (a 64/32 hardware-friendly estimate step).
// MAXW = capacity in 32-bit words of the scratch arrays.
template <int MAXW>
MPFX_HD void n_div_q(limb_t *qp, const limb_t *np, int nn, const limb_t *dp,
                     int dn)
{
  uint32_t un[MAXW + 2];
  uint32_t vn[MAXW];
  const int qn = nn - dn + 1;
  for(int i = 0; i < qn; ++i)
    qp[i] = 0;
  const int s = clz64(dp[dn - 1]);
  // normalised divisor, 2*dn digits, top bit set
  for(int i = dn - 1; i >= 0; --i)
    {
      limb_t w = dp[i] << s;
      if(s && i > 0)
        w |= dp[i - 1] >> (64 - s);
      vn[2 * i] = (uint32_t)w;
      vn[2 * i + 1] = (uint32_t)(w >> 32);
    }
  // shifted numerator, nn+1 limbs = 2*nn+2 digits
  {
    const limb_t top = s ? (np[nn - 1] >> (64 - s)) : 0;
    un[2 * nn] = (uint32_t)top;
    un[2 * nn + 1] = (uint32_t)(top >> 32);
    for(int i = nn - 1; i >= 0; --i)
      {
        limb_t w = np[i] << s;
        if(s && i > 0)
          w |= np[i - 1] >> (64 - s);
        un[2 * i] = (uint32_t)w;
        un[2 * i + 1] = (uint32_t)(w >> 32);
      }
  }
  const int n = 2 * dn; // divisor digits
  const int m = 2 * qn; // quotient digits j = m-1 .. 0
  const uint64_t vtop = vn[n - 1];
  const uint64_t vsec = vn[n - 2];
  for(int j = m - 1; j >= 0; --j)
    {
      const uint64_t num = ((uint64_t)un[j + n] << 32) | un[j + n - 1];
      uint64_t qhat, rhat;
      if(un[j + n] >= vtop)
        qhat = 0xFFFFFFFFull;
      else
        qhat = num / vtop;
      rhat = num - qhat * vtop;
      while(rhat <= 0xFFFFFFFFull
            && qhat * vsec > ((rhat << 32) | un[j + n - 2]))
        {
          qhat--;
          rhat += vtop;
        }
      // multiply and subtract
      uint64_t borrow = 0, carry = 0;
      for(int i = 0; i < n; ++i)
        {
          const uint64_t p = qhat * vn[i] + carry;
          carry = p >> 32;
          const uint64_t d = (uint64_t)un[i + j] - (uint32_t)p - borrow;
          un[i + j] = (uint32_t)d;
          borrow = (d >> 32) & 1;
        }
      const uint64_t d = (uint64_t)un[j + n] - carry - borrow;
      un[j + n] = (uint32_t)d;
      if((d >> 32) & 1)
        {
          // estimate was one too large: add the divisor back
          qhat--;
          uint64_t c = 0;
          for(int i = 0; i < n; ++i)
            {
              const uint64_t t = (uint64_t)un[i + j] + vn[i] + c;
              un[i + j] = (uint32_t)t;
              c = t >> 32;
            }
          un[j + n] += (uint32_t)c;
        }
      qp[j >> 1] |= qhat << (32 * (j & 1));
    }
}

// GMP mpf_div: quotient of P+1 limbs, floor(U * B^zeros / V), one possible
// leading zero limb stripped.
template <int NL>
MPFX_HD void div(Num<NL> &r, const Num<NL> &u, const Num<NL> &v)
{
  constexpr int P = NL - 1;
  if(u.sign == 0)
    {
      set_zero(r);
      return;
    }
  // v.sign == 0 is a caller error (GMP raises SIGFPE); callers check first.
  limb_t np[NL + P];
  for(int i = 0; i < P; ++i)
    np[i] = 0;
  for(int i = 0; i < NL; ++i)
    np[P + i] = u.d[i];
  limb_t q[NL];
  n_div_q<2 * (NL + P) + 2>(q, np, NL + P, v.d, NL);
  const int32_t sgn = u.sign * v.sign;
  int32_t e = u.exp - v.exp + 1;
  if(q[NL - 1] == 0)
    {
      e -= 1;
      r.d[0] = 0;
      for(int i = 1; i < NL; ++i)
        r.d[i] = q[i - 1];
    }
  else
    {
      for(int i = 0; i < NL; ++i)
        r.d[i] = q[i];
    }
  r.sign = sgn;
  r.exp = e;
}

// `x /= 4` of compute_schur_complement.cxx:102 — an mpf_div by the one-limb
// value 4: the NL-limb mantissa is shifted right by two bits (low bits are
// dropped), and a vanished top limb is stripped.
template <int NL> MPFX_HD void div4(Num<NL> &r, const Num<NL> &u)
{
  if(u.sign == 0)
    {
      set_zero(r);
      return;
    }
  limb_t q[NL];
  for(int i = 0; i < NL; ++i)
    {
      limb_t w = u.d[i] >> 2;
      if(i + 1 < NL)
        w |= u.d[i + 1] << 62;
      q[i] = w;
    }
  int32_t e = u.exp; // exp(u) - exp(4) + 1 = exp(u)
  if(q[NL - 1] == 0)
    {
      e -= 1;
      r.d[0] = 0;
      for(int i = 1; i < NL; ++i)
        r.d[i] = q[i - 1];
    }
  else
    {
      for(int i = 0; i < NL; ++i)
        r.d[i] = q[i];
    }
  r.sign = u.sign;
  r.exp = e;
}

// ---------------------------------------------------------------------- sqrt
// floor(sqrt(T)) of a tn-limb integer (tp[tn-1] != 0) -> (tn+1)/2 limbs.
// Integer Newton iteration from an over-estimate; the result is independent of
// the starting point.
template <int MAXL>
MPFX_HD void n_sqrt(limb_t *rp, const limb_t *tp, int tn)
{
  const int rn = (tn + 1) / 2;
  // bit length of T
  const int topbits = 64 - clz64(tp[tn - 1]);
  const int bits = 64 * (tn - 1) + topbits;
  // leading 62..64 bits with an even shift
  int sh = bits - 64;
  if(sh < 0)
    sh = 0;
  if(sh & 1)
    sh += 1;
  uint64_t lead;
  {
    const int ls = sh / 64, bs = sh % 64;
    lead = tp[ls] >> bs;
    if(bs && ls + 1 < tn)
      lead |= tp[ls + 1] << (64 - bs);
  }
  // 32-bit root of lead, rounded up
  uint64_t r0 = (uint64_t)::sqrt((double)lead);
  while(r0 * r0 > lead || r0 > 0xFFFFFFFFull)
    r0--;
  while(r0 < 0xFFFFFFFFull && (r0 + 1) * (r0 + 1) <= lead)
    r0++;
  r0 += 1; // now r0 > sqrt(lead + 1 - eps)  =>  (r0 << sh/2) > sqrt(T)
  limb_t x[MAXL + 1], y[MAXL + 2], q[MAXL + 2];
  int xn;
  {
    // x = r0 << (sh/2)
    const int hs = sh / 2;
    const int ls = hs / 64, bs = hs % 64;
    xn = ls + 2;
    for(int i = 0; i < xn; ++i)
      x[i] = 0;
    x[ls] = r0 << bs;
    if(bs)
      x[ls + 1] = r0 >> (64 - bs);
    while(xn > 0 && x[xn - 1] == 0)
      xn--;
  }
  for(int it = 0; it < 200; ++it)
    {
      // q = T / x
      int qn;
      if(xn > tn)
        {
          qn = 1;
          q[0] = 0;
        }
      else
        {
          qn = tn - xn + 1;
          n_div_q<2 * MAXL + 2>(q, tp, tn, x, xn);
        }
      // y = (x + q) >> 1
      int yn = xn > qn ? xn : qn;
      limb_t cy = 0;
      for(int i = 0; i < yn; ++i)
        {
          const limb_t a = i < xn ? x[i] : 0, b = i < qn ? q[i] : 0;
          limb_t s = a + b;
          limb_t c1 = s < a;
          limb_t t = s + cy;
          limb_t c2 = t < s;
          y[i] = t;
          cy = c1 | c2;
        }
      y[yn] = cy;
      yn++;
      for(int i = 0; i < yn; ++i)
        {
          limb_t w = y[i] >> 1;
          if(i + 1 < yn)
            w |= y[i + 1] << 63;
          y[i] = w;
        }
      while(yn > 0 && y[yn - 1] == 0)
        yn--;
      // stop when y >= x
      bool ge;
      if(yn != xn)
        ge = yn > xn;
      else
        ge = n_cmp(y, x, xn) >= 0;
      if(ge)
        break;
      xn = yn;
      for(int i = 0; i < yn; ++i)
        x[i] = y[i];
    }
  for(int i = 0; i < rn; ++i)
    rp[i] = i < xn ? x[i] : 0;
}

// GMP mpf_sqrt: mantissa top-aligned in 2P - (exp odd) limbs, integer square
// root of P limbs.  Caller guarantees u > 0.
template <int NL> MPFX_HD void sqrt(Num<NL> &r, const Num<NL> &u)
{
  constexpr int P = NL - 1;
  if(u.sign == 0)
    {
      set_zero(r);
      return;
    }
  const int expodd = u.exp & 1;
  const int tsize = 2 * P - expodd;
  const int32_t rexp = (u.exp + expodd) / 2;
  limb_t t[2 * P];
  // NL <= tsize needs P >= 2
  for(int i = 0; i < tsize - NL; ++i)
    t[i] = 0;
  for(int i = 0; i < NL; ++i)
    t[tsize - NL + i] = u.d[i];
  limb_t root[P + 1];
  n_sqrt<2 * P>(root, t, tsize);
  r.d[0] = 0;
  for(int i = 0; i < P; ++i)
    r.d[i + 1] = root[i];
  r.sign = 1;
  r.exp = rexp;
}

// ------------------------------------------------------------- 2^k scalings
// shared tail of mpf_mul_2exp / mpf_div_2exp for a bit count that is not a
// multiple of 64: top P limbs shifted left by `lsh` bits into P+1 limbs.
template <int NL>
MPFX_HD void shift_tail(Num<NL> &r, const Num<NL> &u, int lsh, int32_t ebase)
{
  constexpr int P = NL - 1;
  limb_t w[NL];
  // w = (u.d[1..P]) << lsh, P+1 limbs
  w[P] = u.d[P] >> (64 - lsh);
  for(int i = P - 1; i >= 0; --i)
    {
      limb_t x = u.d[i + 1] << lsh;
      if(i > 0)
        x |= u.d[i] >> (64 - lsh);
      w[i] = x;
    }
  if(w[P] != 0)
    {
      for(int i = 0; i < NL; ++i)
        r.d[i] = w[i];
      r.exp = ebase + 1;
    }
  else
    {
      limb_t o[NL];
      o[0] = 0;
      for(int i = 0; i < P; ++i)
        o[i + 1] = w[i];
      for(int i = 0; i < NL; ++i)
        r.d[i] = o[i];
      r.exp = ebase;
    }
  r.sign = u.sign;
}
template <int NL>
MPFX_HD void mul_2exp(Num<NL> &r, const Num<NL> &u, uint32_t k)
{
  if(u.sign == 0)
    {
      set_zero(r);
      return;
    }
  if(k % 64 == 0)
    {
      r = u;
      r.exp = u.exp + (int32_t)(k / 64);
      return;
    }
  shift_tail(r, u, (int)(k % 64), u.exp + (int32_t)(k / 64));
}
template <int NL>
MPFX_HD void div_2exp(Num<NL> &r, const Num<NL> &u, uint32_t k)
{
  if(u.sign == 0)
    {
      set_zero(r);
      return;
    }
  if(k % 64 == 0)
    {
      r = u;
      r.exp = u.exp - (int32_t)(k / 64);
      return;
    }
  shift_tail(r, u, 64 - (int)(k % 64), u.exp - (int32_t)(k / 64) - 1);
}

// ------------------------------------------------------------------ compare
template <int NL> MPFX_HD int cmp(const Num<NL> &u, const Num<NL> &v)
{
  if(u.sign != v.sign)
    return u.sign > v.sign ? 1 : -1;
  if(u.sign == 0)
    return 0;
  int c;
  if(u.exp != v.exp)
    c = u.exp > v.exp ? 1 : -1;
  else
    c = n_cmp(u.d, v.d, NL);
  return u.sign > 0 ? c : -c;
}

// ---------------------------------------------------- integer conversions
// mpz_set_f (as used by fmpz_set_mpf, reference fmpz_BigFloat_convert.hxx:13):
// integer part, truncated toward zero.  Writes nw 32-bit words (magnitude);
// returns false if the integer does not fit.
template <int NL>
MPFX_HD bool trunc_to_words(uint32_t *out, int nw, const Num<NL> &u)
{
  for(int i = 0; i < nw; ++i)
    out[i] = 0;
  if(u.sign == 0 || u.exp <= 0)
    return true;
  // integer limb k (k = 0 .. exp-1) is mantissa limb k + NL - exp
  const int nlimbs = u.exp;
  bool ok = true;
  for(int k = 0; k < nlimbs; ++k)
    {
      const int src = k + NL - u.exp;
      const limb_t w = (src >= 0 && src < NL) ? u.d[src] : 0;
      if(2 * k < nw)
        out[2 * k] = (uint32_t)w;
      else if((uint32_t)w)
        ok = false;
      if(2 * k + 1 < nw)
        out[2 * k + 1] = (uint32_t)(w >> 32);
      else if((uint32_t)(w >> 32))
        ok = false;
    }
  return ok;
}
// mpf_set_z (as used by fmpz_get_mpf, fmpz_BigFloat_convert.hxx:9): top NL
// limbs of an n-limb integer, exponent n.
template <int NL>
MPFX_HD void from_limbs(Num<NL> &r, const limb_t *zp, int zn, int sign)
{
  while(zn > 0 && zp[zn - 1] == 0)
    zn--;
  if(zn == 0 || sign == 0)
    {
      set_zero(r);
      return;
    }
  for(int i = 0; i < NL; ++i)
    {
      const int src = zn - NL + i;
      r.d[i] = src >= 0 ? zp[src] : 0;
    }
  r.sign = sign;
  r.exp = zn;
}

// ------------------------------------------------------ packed memory form
// word 0: low 32 bits = exponent (int32), high 32 bits = sign (int32);
// words 1..NL: limbs.  Stride = elem_words(NL).
template <int NL> MPFX_HD void load(Num<NL> &r, const limb_t *p)
{
  const limb_t h = p[0];
  r.exp = (int32_t)(uint32_t)h;
  r.sign = (int32_t)(uint32_t)(h >> 32);
#pragma unroll
  for(int i = 0; i < NL; ++i)
    r.d[i] = p[1 + i];
}
template <int NL> MPFX_HD void store(limb_t *p, const Num<NL> &r)
{
  p[0] = (limb_t)(uint32_t)r.exp | ((limb_t)(uint32_t)r.sign << 32);
#pragma unroll
  for(int i = 0; i < NL; ++i)
    p[1 + i] = r.d[i];
  if(((NL + 1) & 1))
    p[NL + 1] = 0;
}
} // namespace mpfx

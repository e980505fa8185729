// Host fuzz of mpfw (register-form mpf arithmetic, sdpb_b200/csrc/mpfw.h)
// against mpfx (sdpb_b200/csrc/mpfx.h, itself fuzzed against libgmp by
// mpfx_fuzz.cpp): acc +- a*b on random and adversarial operands must agree in
// every bit.  Build: g++ -O2 -std=c++17 -o build/mpfw_fuzz tests/cpp/mpfw_fuzz.cpp
#define MPFW_COUNT_RARE 1
#include "../../sdpb_b200/csrc/mpfw.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint64_t rnd()
{
  uint64_t z = (rng_state += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}

template <int NL> static void random_num(mpfx::Num<NL> &x, int mode)
{
  for(int i = 0; i < NL; ++i)
    x.d[i] = rnd();
  switch(mode % 7)
    {
    case 1: // short (trailing zero limbs)
      for(int i = 0; i < NL - 1 - (int)(rnd() % (NL - 1)); ++i)
        x.d[i] = 0;
      break;
    case 2: // small top limb
      x.d[NL - 1] = 1 + rnd() % 3;
      break;
    case 3: // runs of ones
      for(int i = 0; i < NL; ++i)
        x.d[i] = (rnd() & 1) ? ~0ull : 0ull;
      break;
    case 4: // all ones
      for(int i = 0; i < NL; ++i)
        x.d[i] = ~0ull;
      if(rnd() & 1)
        x.d[0] = rnd();
      break;
    case 5: // 1 000...
      for(int i = 0; i < NL; ++i)
        x.d[i] = 0;
      x.d[NL - 1] = 1;
      if(rnd() & 1)
        x.d[0] = rnd();
      break;
    default: break;
    }
  if(x.d[NL - 1] == 0)
    x.d[NL - 1] = 1;
  x.sign = (rnd() & 1) ? 1 : -1;
  x.exp = (rnd() & 3) ? (int32_t)(rnd() % 7) - 3 : (int32_t)(rnd() % (4 * NL)) - 2 * NL;
  if(rnd() % 37 == 0)
    mpfx::set_zero(x);
}

template <int NL> static bool same(const mpfx::Num<NL> &a, const mpfx::Num<NL> &b)
{
  if(a.sign != b.sign)
    return false;
  if(a.sign == 0)
    return true;
  return a.exp == b.exp && memcmp(a.d, b.d, sizeof(a.d)) == 0;
}
template <int NL> static void dump(const char *n, const mpfx::Num<NL> &a)
{
  printf("  %s: sign %d exp %d ", n, a.sign, a.exp);
  for(int i = NL - 1; i >= 0; --i)
    printf("%016llx ", (unsigned long long)a.d[i]);
  printf("\n");
}

template <int NL> static long run(long iters)
{
  long bad = 0;
  for(long it = 0; it < iters; ++it)
    {
      mpfx::Num<NL> a, b, c, want, prod, got;
      random_num(a, (int)rnd());
      random_num(b, (int)rnd());
      random_num(c, (int)rnd());
      const bool negate = rnd() & 1;
      const int kind = (int)(rnd() % 8);
      if(kind == 0 && a.sign && b.sign)
        {
          // force the accumulator close to the product (cancellation paths)
          mpfx::mul(c, a, b);
          c.d[rnd() % NL] ^= 1ull << (rnd() % 64);
          if(c.d[NL - 1] == 0)
            c.d[NL - 1] = 1;
          c.sign = negate ? c.sign : -c.sign;
          if(kind == 0 && (rnd() & 1))
            c.exp += (rnd() & 1) ? 1 : -1;
        }
      else if(kind == 1 && a.sign && b.sign)
        {
          // exponent gap of one limb with 1 000 / fff patterns
          mpfx::mul(prod, a, b);
          for(int i = 0; i < NL; ++i)
            c.d[i] = 0;
          c.d[NL - 1] = 1;
          c.d[0] = rnd() & 3;
          c.exp = prod.exp + 1;
          c.sign = negate ? prod.sign : -prod.sign;
        }
      // reference: mpf_mul then mpf_add/sub via mpfx
      if(a.sign == 0 || b.sign == 0)
        want = c;
      else
        {
          mpfx::mul(prod, a, b);
          if(negate)
            mpfx::sub(want, c, prod);
          else
            mpfx::add(want, c, prod);
        }
      mpfw::Reg<NL> ra, rb, rc;
      mpfw::from_num(ra, a);
      mpfw::from_num(rb, b);
      mpfw::from_num(rc, c);
      mpfw::mac(rc, ra, rb, negate);
      mpfw::to_num(got, rc);
      // memory-form entry point as well
      uint32_t ma[2 * NL + 4], mb[2 * NL + 4];
      mpfw::store(ma, ra);
      mpfw::store(mb, rb);
      mpfw::Reg<NL> rc2;
      mpfw::from_num(rc2, c);
      mpfw::mac(rc2, ma, mb, negate);
      mpfx::Num<NL> got2;
      mpfw::to_num(got2, rc2);
      if(!same(got, want) || !same(got2, want))
        {
          if(bad < 5)
            {
              printf("MISMATCH NL=%d negate=%d\n", NL, (int)negate);
              dump("a", a);
              dump("b", b);
              dump("c", c);
              dump("want", want);
              dump("got", got);
            }
          ++bad;
        }
    }
  // division by reciprocal against mpfx::div
  long dbad = 0;
  for(long it = 0; it < iters / 4; ++it)
    {
      mpfx::Num<NL> u, d, want, got;
      random_num(u, (int)rnd());
      do
        random_num(d, (int)rnd());
      while(d.sign == 0);
      if(it % 11 == 0)
        {
          // exact multiples and near misses
          mpfx::Num<NL> k;
          random_num(k, 2);
          if(k.sign)
            {
              mpfx::mul(u, d, k);
              if(it % 22 == 0)
                u.d[0] ^= 1;
            }
        }
      {
        mpfx::Num<NL> w4, g4;
        mpfx::div4(w4, u);
        mpfw::Reg<NL> r4;
        mpfw::from_num(r4, u);
        mpfw::div4<NL>(r4);
        mpfw::to_num(g4, r4);
        if(!same(g4, w4))
          {
            if(dbad < 5)
              printf("DIV4 MISMATCH NL=%d\n", NL);
            ++dbad;
          }
      }
      mpfx::div(want, u, d);
      uint32_t R[2 * NL + 4];
      mpfw::reciprocal<NL>(R, d);
      mpfw::Reg<NL> ru, rd;
      mpfw::from_num(ru, u);
      mpfw::from_num(rd, d);
      mpfw::div_recip<NL>(ru, rd.sign, rd.exp, rd.w, R);
      mpfw::to_num(got, ru);
      if(!same(got, want))
        {
          if(dbad < 5)
            {
              printf("DIV MISMATCH NL=%d\n", NL);
              dump("u", u);
              dump("d", d);
              dump("want", want);
              dump("got", got);
            }
          ++dbad;
        }
    }
  // quotients whose fraction sits m / 2^32 below a whole number, m = 0 .. 160: the band in which
  // div_recip's guard word switches between "decided" and "form the exact remainder"
  for(int m = 0; m <= 160; ++m)
    for(int rep = 0; rep < 8; ++rep)
      {
        constexpr int n2 = 2 * NL;
        uint32_t K[n2], D[n2], prod[2 * n2];
        for(int i = 0; i < n2; ++i)
          {
            K[i] = (uint32_t)rnd();
            D[i] = (uint32_t)rnd();
          }
        K[n2 - 1] = 0; // K = q beta + (beta - m), q of n2 - 2 words
        if(K[n2 - 2] == 0)
          K[n2 - 2] = 1;
        K[0] = (uint32_t)(0u - (uint32_t)m);
        if(D[n2 - 1] == 0)
          D[n2 - 1] = 1 + (uint32_t)(rnd() % 5);
        if(rep & 1)
          D[n2 - 1] = 1 + (uint32_t)(rnd() % 3); // small top word: coarse quotient grid
        for(int i = 0; i < 2 * n2; ++i)
          prod[i] = 0;
        for(int i = 0; i < n2; ++i)
          {
            uint64_t carry = 0;
            for(int j = 0; j < n2; ++j)
              {
                const uint64_t t = (uint64_t)K[i] * D[j] + prod[i + j] + carry;
                prod[i + j] = (uint32_t)t;
                carry = t >> 32;
              }
            prod[i + n2] = (uint32_t)carry;
          }
        // U = floor(K D / beta^(n2-1)), needs to fit n2 words with a non-zero top limb
        if(prod[2 * n2 - 1] != 0)
          continue;
        mpfx::Num<NL> u, d, want, got;
        for(int i = 0; i < NL; ++i)
          {
            u.d[i] = (uint64_t)prod[n2 - 1 + 2 * i] | ((uint64_t)prod[n2 + 2 * i] << 32);
            d.d[i] = (uint64_t)D[2 * i] | ((uint64_t)D[2 * i + 1] << 32);
          }
        if(u.d[NL - 1] == 0 || d.d[NL - 1] == 0)
          continue;
        u.sign = (rnd() & 1) ? 1 : -1;
        d.sign = (rnd() & 1) ? 1 : -1;
        u.exp = (int32_t)(rnd() % 7) - 3;
        d.exp = (int32_t)(rnd() % 7) - 3;
        mpfx::div(want, u, d);
        uint32_t R[2 * NL + 4];
        mpfw::reciprocal<NL>(R, d);
        mpfw::Reg<NL> ru, rd;
        mpfw::from_num(ru, u);
        mpfw::from_num(rd, d);
        mpfw::div_recip<NL>(ru, rd.sign, rd.exp, rd.w, R);
        mpfw::to_num(got, ru);
        if(!same(got, want))
          {
            if(dbad < 5)
              {
                printf("DIV MISMATCH (guard band, m = %d) NL=%d\n", m, NL);
                dump("u", u);
                dump("d", d);
              }
            ++dbad;
          }
      }
  // fast reciprocal / sqrt against the reference routines
  long rbad = 0, sbad = 0;
  for(long it = 0; it < iters / 4; ++it)
    {
      mpfx::Num<NL> d;
      do
        random_num(d, (int)rnd());
      while(d.sign == 0);
      d.sign = 1;
      uint32_t R0[2 * NL + 4], R1[2 * NL + 4];
      mpfw::reciprocal<NL>(R0, d);
      mpfw::Reg<NL> rd;
      mpfw::from_num(rd, d);
      mpfw::reciprocal_fast<NL>(R1, rd);
      if(memcmp(R0, R1, sizeof(R0)) != 0)
        {
          if(rbad < 3)
            {
              printf("RECIP MISMATCH NL=%d\n", NL);
              dump("d", d);
            }
          ++rbad;
        }
      mpfx::Num<NL> want, got;
      mpfx::sqrt(want, d);
      mpfw::Reg<NL> rr;
      mpfw::sqrt_fast<NL>(rr, rd);
      mpfw::to_num(got, rr);
      if(!same(got, want))
        {
          if(sbad < 3)
            {
              printf("SQRT MISMATCH NL=%d\n", NL);
              dump("u", d);
              dump("want", want);
              dump("got", got);
            }
          ++sbad;
        }
    }
  printf("NL=%d: reciprocal_fast %ld mismatches, sqrt_fast %ld mismatches\n", NL, rbad, sbad);
  bad += rbad + sbad;
  printf("NL=%d: %ld mac cases, %ld mismatches; %ld div cases, %ld mismatches\n", NL, iters, bad,
         iters / 4, dbad);
  return bad + dbad;
}

int main(int argc, char **argv)
{
  const long iters = argc > 1 ? atol(argv[1]) : 200000;
  long bad = 0;
  bad += run<4>(iters);
  bad += run<6>(iters);
  bad += run<9>(iters);
  bad += run<13>(iters);
  bad += run<14>(iters);
  bad += run<17>(iters);
  bad += run<26>(iters / 4);
  printf("rare paths exercised: mul guard fallback %ld, sub close-operand fallback %ld\n",
         mpfw::rare_mul_count, mpfw::rare_sub_count);
  printf("slow-path fallbacks (must be 0: the Newton paths are exact by themselves): reciprocal %ld, sqrt %ld\n",
         mpfw::recip_fallbacks, mpfw::sqrt_fallbacks);
  if(mpfw::recip_fallbacks || mpfw::sqrt_fallbacks)
    return 3;
  printf("division: guard word decided %ld, exact remainder formed %ld\n", mpfw::div_fast_count, mpfw::div_exact_count);
  if(mpfw::rare_mul_count == 0 || mpfw::rare_sub_count == 0 || mpfw::div_exact_count == 0 || mpfw::div_fast_count == 0)
    {
      printf("rare paths not covered\n");
      return 2;
    }
  return bad ? 1 : 0;
}

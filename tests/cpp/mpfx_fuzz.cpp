// Fuzz mpfx (the fixed-limb host/device arithmetic) against the real libgmp
// mpf_* on variable-size operands.  Build: see tests/test_mpfx_vs_gmp.py.
// Usage: mpfx_fuzz <iterations> <seed>
#include "../../sdpb_b200/csrc/host/gmp_abi.h"
#include "../../sdpb_b200/csrc/mpfx.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

using mpfx::limb_t;
using mpfx::Num;

static std::mt19937_64 rng;

template <int NL> static void from_mpf(Num<NL> &r, mpf_srcptr x)
{
  int size = x->_mp_size;
  int asz = size < 0 ? -size : size;
  if(asz == 0)
    {
      mpfx::set_zero(r);
      return;
    }
  if(asz > NL)
    {
      fprintf(stderr, "operand larger than NL\n");
      exit(2);
    }
  for(int i = 0; i < NL; ++i)
    r.d[i] = 0;
  for(int i = 0; i < asz; ++i)
    r.d[NL - asz + i] = x->_mp_d[i];
  r.sign = size < 0 ? -1 : 1;
  r.exp = (int32_t)x->_mp_exp;
}

template <int NL> static bool same(const Num<NL> &a, const Num<NL> &b)
{
  if(a.sign != b.sign)
    return false;
  if(a.sign == 0)
    return true;
  if(a.exp != b.exp)
    return false;
  return memcmp(a.d, b.d, sizeof(a.d)) == 0;
}

template <int NL> static void dump(const char *name, const Num<NL> &a)
{
  fprintf(stderr, "%s: sign %d exp %d limbs(hi..lo):", name, a.sign, a.exp);
  for(int i = NL - 1; i >= 0; --i)
    fprintf(stderr, " %016lx", (unsigned long)a.d[i]);
  fprintf(stderr, "\n");
}

static limb_t special_limb()
{
  switch(rng() % 8)
    {
    case 0: return 0;
    case 1: return ~(limb_t)0;
    case 2: return 1;
    case 3: return rng() & 0xFF;
    case 4: return ~(limb_t)0 - (rng() & 0xFF);
    case 5: return (limb_t)1 << 63;
    default: return rng();
    }
}

// fill an mpf with a random value of random size; structured patterns make the
// cancellation paths of mpf_sub likely
static void random_mpf(mpf_ptr x, int NL, int mode)
{
  int size = 1 + (int)(rng() % NL);
  if(rng() % 3 == 0)
    size = NL;
  if(rng() % 5 == 0)
    size = NL - 1;
  if(size < 1)
    size = 1;
  for(int i = 0; i < size; ++i)
    x->_mp_d[i] = (mode == 0) ? rng() : special_limb();
  if(x->_mp_d[size - 1] == 0)
    x->_mp_d[size - 1] = 1 + (rng() & 3);
  x->_mp_size = (rng() & 1) ? size : -size;
  x->_mp_exp = (long)(rng() % 7) - 3;
  if(rng() % 6 == 0)
    x->_mp_exp += (long)(rng() % (2 * NL)) - NL;
  if(rng() % 40 == 0)
    {
      x->_mp_size = 0;
      x->_mp_exp = 0;
    }
}

// v := something close to u so that u - v cancels heavily
static void near_copy(mpf_ptr v, mpf_srcptr u, int NL)
{
  int asz = u->_mp_size < 0 ? -u->_mp_size : u->_mp_size;
  if(asz == 0)
    return;
  memcpy(v->_mp_d, u->_mp_d, asz * sizeof(limb_t));
  v->_mp_size = u->_mp_size;
  v->_mp_exp = u->_mp_exp;
  int k = (int)(rng() % asz);
  switch(rng() % 4)
    {
    case 0: v->_mp_d[k] += 1; break;
    case 1: v->_mp_d[k] -= 1; break;
    case 2: v->_mp_d[k] ^= (limb_t)1 << (rng() % 64); break;
    default: break;
    }
  if(v->_mp_d[asz - 1] == 0)
    v->_mp_d[asz - 1] = 1;
  if(rng() % 4 == 0 && asz > 1)
    {
      // x+1 0000.. vs x ffff.. pattern
      for(int i = 0; i < asz - 1; ++i)
        {
          v->_mp_d[i] = (rng() % 3) ? ~(limb_t)0 : rng();
          const_cast<mp_limb_t *>(u->_mp_d)[i] = (rng() % 3) ? 0 : rng();
        }
      v->_mp_d[asz - 1] = u->_mp_d[asz - 1] - 1;
      if(v->_mp_d[asz - 1] == 0)
        {
          const_cast<mp_limb_t *>(u->_mp_d)[asz - 1] = 2;
          v->_mp_d[asz - 1] = 1;
        }
    }
  if(rng() % 6 == 0 && asz > 1)
    {
      // 1 0000.. (exp e+1)  vs  ffff.. (exp e)
      mpf_ptr uu = const_cast<mpf_ptr>(u);
      for(int i = 0; i < asz; ++i)
        {
          uu->_mp_d[i] = (rng() % 4) ? 0 : rng();
          v->_mp_d[i] = (rng() % 4) ? ~(limb_t)0 : rng();
        }
      uu->_mp_d[asz - 1] = 1;
      v->_mp_d[asz - 1] = ~(limb_t)0;
      uu->_mp_exp = v->_mp_exp + 1;
    }
}

template <int NL> static long run(long iters)
{
  constexpr int P = NL - 1;
  const int bits = 64 * (P - 1);
  mpf_t a, b, c;
  mpf_init2(a, bits);
  mpf_init2(b, bits);
  mpf_init2(c, bits);
  if(a->_mp_prec != P)
    {
      fprintf(stderr, "unexpected _mp_prec %d for NL %d\n", a->_mp_prec, NL);
      exit(2);
    }
  long fails = 0;
  for(long it = 0; it < iters; ++it)
    {
      const int mode = (int)(rng() % 2);
      random_mpf(a, NL, mode);
      random_mpf(b, NL, mode);
      if(rng() % 3 == 0)
        near_copy(b, a, NL);
      Num<NL> A, B, R, G;
      from_mpf(A, a);
      from_mpf(B, b);
      const int nops = 9;
      for(int op = 0; op < nops; ++op)
        {
          const char *name = "";
          uint32_t k = (uint32_t)(rng() % 2000);
          if(rng() % 4 == 0)
            k = 64 * (uint32_t)(rng() % 30);
          switch(op)
            {
            case 0:
              name = "mul";
              mpf_mul(c, a, b);
              mpfx::mul(R, A, B);
              break;
            case 1:
              name = "add";
              mpf_add(c, a, b);
              mpfx::add(R, A, B);
              break;
            case 2:
              name = "sub";
              mpf_sub(c, a, b);
              mpfx::sub(R, A, B);
              break;
            case 3:
              name = "div";
              if(b->_mp_size == 0)
                continue;
              mpf_div(c, a, b);
              mpfx::div(R, A, B);
              break;
            case 4:
              name = "sqrt";
              if(a->_mp_size <= 0)
                continue;
              mpf_sqrt(c, a);
              mpfx::sqrt(R, A);
              break;
            case 5:
              name = "mul_2exp";
              mpf_mul_2exp(c, a, k);
              mpfx::mul_2exp(R, A, k);
              break;
            case 6:
              name = "div_2exp";
              mpf_div_2exp(c, a, k);
              mpfx::div_2exp(R, A, k);
              break;
            case 7:
              {
                name = "div4";
                mpf_t four;
                mpf_init2(four, bits);
                mpf_set_ui(four, 4);
                mpf_div(c, a, four);
                mpfx::div4(R, A);
                // mpf_div_ui must agree as well
                mpf_t c2;
                mpf_init2(c2, bits);
                mpf_div_ui(c2, a, 4);
                Num<NL> G2;
                from_mpf(G2, c2);
                from_mpf(G, c);
                if(!same(G, G2))
                  {
                    fprintf(stderr, "note: mpf_div(4) != mpf_div_ui(4)\n");
                    fails++;
                  }
                mpf_clear(four);
                mpf_clear(c2);
                break;
              }
            case 8:
              {
                name = "cmp";
                int g = mpf_cmp(a, b);
                int m = mpfx::cmp(A, B);
                g = g > 0 ? 1 : (g < 0 ? -1 : 0);
                if(g != m)
                  {
                    fprintf(stderr, "MISMATCH cmp NL=%d: gmp %d mpfx %d\n", NL,
                            g, m);
                    dump("a", A);
                    dump("b", B);
                    fails++;
                  }
                continue;
              }
            }
          from_mpf(G, c);
          if(!same(G, R))
            {
              fails++;
              if(fails < 6)
                {
                  fprintf(stderr, "MISMATCH %s NL=%d (k=%u)\n", name, NL, k);
                  dump("a   ", A);
                  dump("b   ", B);
                  dump("gmp ", G);
                  dump("mpfx", R);
                }
            }
        }
      // mul with aliased operands (u == v pointer): GMP takes its squaring
      // branch; values must not differ from the general one
      {
        mpf_mul(c, a, a);
        mpfx::mul(R, A, A);
        from_mpf(G, c);
        if(!same(G, R))
          {
            fails++;
            if(fails < 6)
              {
                fprintf(stderr, "MISMATCH sqr NL=%d\n", NL);
                dump("a   ", A);
                dump("gmp ", G);
                dump("mpfx", R);
              }
          }
      }
      // integer conversions: mpz_set_f / mpf_set_z
      {
        mpz_t z;
        mpz_init(z);
        mpz_set_f(z, a);
        uint32_t w[4 * NL + 8];
        const int nw = 4 * NL + 8;
        bool ok = mpfx::trunc_to_words(w, nw, A);
        int zs = z->_mp_size < 0 ? -z->_mp_size : z->_mp_size;
        bool good = true;
        if(ok)
          {
            for(int i = 0; i < nw / 2; ++i)
              {
                limb_t want = i < zs ? z->_mp_d[i] : 0;
                limb_t got = (limb_t)w[2 * i] | ((limb_t)w[2 * i + 1] << 32);
                if(want != got)
                  good = false;
              }
          }
        else if(zs <= nw / 2)
          good = false;
        if(!good)
          {
            fails++;
            if(fails < 6)
              {
                fprintf(stderr, "MISMATCH trunc NL=%d\n", NL);
                dump("a", A);
              }
          }
        if(ok && zs > 0)
          {
            mpf_set_z(c, z);
            from_mpf(G, c);
            limb_t zl[2 * NL + 4];
            for(int i = 0; i < nw / 2; ++i)
              zl[i] = (limb_t)w[2 * i] | ((limb_t)w[2 * i + 1] << 32);
            mpfx::from_limbs(R, zl, nw / 2, A.sign);
            if(!same(G, R))
              {
                fails++;
                if(fails < 6)
                  {
                    fprintf(stderr, "MISMATCH set_z NL=%d\n", NL);
                    dump("gmp ", G);
                    dump("mpfx", R);
                  }
              }
          }
        mpz_clear(z);
      }
    }
  mpf_clear(a);
  mpf_clear(b);
  mpf_clear(c);
  return fails;
}

int main(int argc, char **argv)
{
  long iters = argc > 1 ? atol(argv[1]) : 20000;
  unsigned long seed = argc > 2 ? strtoul(argv[2], 0, 10) : 1;
  rng.seed(seed);
  long fails = 0;
  fails += run<3>(iters);
  fails += run<4>(iters);
  fails += run<6>(iters);
  fails += run<9>(iters);
  fails += run<12>(iters);
  fails += run<14>(iters);
  fails += run<17>(iters);
  fails += run<26>(iters / 4 + 1);
  printf("mpfx_fuzz: %ld mismatches\n", fails);
  return fails ? 1 : 0;
}

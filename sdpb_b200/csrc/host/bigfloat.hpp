// Host-side BigFloat: a thin RAII wrapper over GMP's mpf_t, i.e. the same
// scalar the reference uses (El::BigFloat wraps mpf_class; reference:
// src/sdp_solve/SDP_Solver/run/bigint_syrk/fmpz/fmpz_BigFloat_convert.hxx:9,13).
// Every operator is exactly one mpf_* call, so host arithmetic truncates the
// same way the reference's does.  Also: conversion to/from the packed
// fixed-limb element format that lives in HBM and crosses the C-ABI
// (include/sdpb_b200.h, sdpb_b200/csrc/mpfx.h).
#pragma once
#include "gmp_abi.h"

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace sdpb_host
{
inline int &working_precision_bits()
{
  static int bits = 768;
  return bits;
}
// mirrors Environment::set_precision (reference src/sdpb_util/Environment.cxx:29-36)
inline void set_precision(int bits)
{
  working_precision_bits() = bits;
  mpf_set_default_prec((mp_bitcnt_t)bits);
}
inline int prec_limbs() { return (working_precision_bits() + 63) / 64 + 1; }
inline int stored_limbs() { return prec_limbs() + 1; }
inline int elem_words() { return (stored_limbs() + 2) & ~1; }

class BigFloat
{
public:
  mpf_t v;
  BigFloat() { mpf_init2(v, (mp_bitcnt_t)working_precision_bits()); }
  BigFloat(const BigFloat &o)
  {
    mpf_init2(v, (mp_bitcnt_t)working_precision_bits());
    mpf_set(v, o.v);
  }
  BigFloat(BigFloat &&o) noexcept
  {
    v[0] = o.v[0];
    o.v[0]._mp_d = nullptr;
  }
  BigFloat(double d)
  {
    mpf_init2(v, (mp_bitcnt_t)working_precision_bits());
    mpf_set_d(v, d);
  }
  BigFloat(int i)
  {
    mpf_init2(v, (mp_bitcnt_t)working_precision_bits());
    mpf_set_si(v, i);
  }
  BigFloat(long i)
  {
    mpf_init2(v, (mp_bitcnt_t)working_precision_bits());
    mpf_set_si(v, i);
  }
  explicit BigFloat(const std::string &dec)
  {
    mpf_init2(v, (mp_bitcnt_t)working_precision_bits());
    if(mpf_set_str(v, dec.c_str(), 10) != 0)
      throw std::runtime_error("BigFloat: cannot parse '" + dec + "'");
  }
  ~BigFloat()
  {
    if(v[0]._mp_d)
      mpf_clear(v);
  }
  BigFloat &operator=(const BigFloat &o)
  {
    if(this != &o)
      mpf_set(v, o.v);
    return *this;
  }
  BigFloat &operator=(BigFloat &&o) noexcept
  {
    if(this != &o)
      {
        __mpf_struct t = v[0];
        v[0] = o.v[0];
        o.v[0] = t;
      }
    return *this;
  }
  BigFloat &operator+=(const BigFloat &o)
  {
    mpf_add(v, v, o.v);
    return *this;
  }
  BigFloat &operator-=(const BigFloat &o)
  {
    mpf_sub(v, v, o.v);
    return *this;
  }
  BigFloat &operator*=(const BigFloat &o)
  {
    mpf_mul(v, v, o.v);
    return *this;
  }
  BigFloat &operator/=(const BigFloat &o)
  {
    mpf_div(v, v, o.v);
    return *this;
  }
  BigFloat operator-() const
  {
    BigFloat r;
    mpf_neg(r.v, v);
    return r;
  }
  void zero() { mpf_set_ui(v, 0); }
  int sgn() const { return v[0]._mp_size > 0 ? 1 : (v[0]._mp_size < 0 ? -1 : 0); }
  double to_double() const { return mpf_get_d(v); }
  // decimal string with `digits` significant digits (0 = exact)
  std::string str(int digits = 0) const
  {
    mp_exp_t e;
    char *s = mpf_get_str(nullptr, &e, 10, (size_t)digits, v);
    std::string m(s);
    free(s);
    if(m.empty())
      return "0";
    std::string out;
    size_t pos = 0;
    if(m[0] == '-')
      {
        out = "-";
        pos = 1;
      }
    out += "0." + m.substr(pos) + "e" + std::to_string((long)e);
    return out;
  }
};
inline BigFloat operator+(const BigFloat &a, const BigFloat &b)
{
  BigFloat r;
  mpf_add(r.v, a.v, b.v);
  return r;
}
inline BigFloat operator-(const BigFloat &a, const BigFloat &b)
{
  BigFloat r;
  mpf_sub(r.v, a.v, b.v);
  return r;
}
inline BigFloat operator*(const BigFloat &a, const BigFloat &b)
{
  BigFloat r;
  mpf_mul(r.v, a.v, b.v);
  return r;
}
inline BigFloat operator/(const BigFloat &a, const BigFloat &b)
{
  BigFloat r;
  mpf_div(r.v, a.v, b.v);
  return r;
}
inline BigFloat operator<<(const BigFloat &a, unsigned k)
{
  BigFloat r;
  mpf_mul_2exp(r.v, a.v, k);
  return r;
}
inline BigFloat operator>>(const BigFloat &a, unsigned k)
{
  BigFloat r;
  mpf_div_2exp(r.v, a.v, k);
  return r;
}
inline bool operator<(const BigFloat &a, const BigFloat &b) { return mpf_cmp(a.v, b.v) < 0; }
inline bool operator>(const BigFloat &a, const BigFloat &b) { return mpf_cmp(a.v, b.v) > 0; }
inline bool operator<=(const BigFloat &a, const BigFloat &b) { return mpf_cmp(a.v, b.v) <= 0; }
inline bool operator>=(const BigFloat &a, const BigFloat &b) { return mpf_cmp(a.v, b.v) >= 0; }
inline bool operator==(const BigFloat &a, const BigFloat &b) { return mpf_cmp(a.v, b.v) == 0; }
inline bool operator!=(const BigFloat &a, const BigFloat &b) { return mpf_cmp(a.v, b.v) != 0; }
inline BigFloat Sqrt(const BigFloat &a)
{
  BigFloat r;
  mpf_sqrt(r.v, a.v);
  return r;
}
inline BigFloat Abs(const BigFloat &a)
{
  BigFloat r;
  mpf_abs(r.v, a.v);
  return r;
}
inline BigFloat Min(const BigFloat &a, const BigFloat &b) { return a < b ? a : b; }
inline BigFloat Max(const BigFloat &a, const BigFloat &b) { return a > b ? a : b; }

// ---- packed element format (see mpfx.h): header word + NL limbs, top aligned
inline void pack(const BigFloat &x, uint64_t *out)
{
  const int nl = stored_limbs(), ew = elem_words();
  const int size = x.v[0]._mp_size;
  const int asz = size < 0 ? -size : size;
  for(int i = 0; i < ew; ++i)
    out[i] = 0;
  if(asz == 0)
    return;
  if(asz > nl)
    throw std::runtime_error("pack: mpf has more limbs than the format");
  const int32_t sign = size < 0 ? -1 : 1;
  const int32_t e = (int32_t)x.v[0]._mp_exp;
  out[0] = (uint64_t)(uint32_t)e | ((uint64_t)(uint32_t)sign << 32);
  for(int i = 0; i < asz; ++i)
    out[1 + nl - asz + i] = x.v[0]._mp_d[i];
}
inline void unpack(BigFloat &x, const uint64_t *in)
{
  const int nl = stored_limbs();
  const int32_t e = (int32_t)(uint32_t)in[0];
  const int32_t sign = (int32_t)(uint32_t)(in[0] >> 32);
  if(sign == 0)
    {
      mpf_set_ui(x.v, 0);
      return;
    }
  // strip trailing zero limbs so the mpf looks the way GMP would have left it
  int lo = 0;
  while(lo < nl - 1 && in[1 + lo] == 0)
    lo++;
  const int asz = nl - lo;
  if(x.v[0]._mp_prec + 1 < asz)
    throw std::runtime_error("unpack: target precision too small");
  for(int i = 0; i < asz; ++i)
    x.v[0]._mp_d[i] = in[1 + lo + i];
  x.v[0]._mp_size = sign < 0 ? -asz : asz;
  x.v[0]._mp_exp = e;
}

// column-major dense matrix of BigFloat
struct Matrix
{
  int h = 0, w = 0;
  std::vector<BigFloat> a;
  Matrix() {}
  Matrix(int h_, int w_) : h(h_), w(w_), a((size_t)h_ * w_) {}
  void resize(int h_, int w_)
  {
    h = h_;
    w = w_;
    a.assign((size_t)h_ * w_, BigFloat());
  }
  BigFloat &operator()(int i, int j) { return a[(size_t)j * h + i]; }
  const BigFloat &operator()(int i, int j) const { return a[(size_t)j * h + i]; }
  void zero()
  {
    for(auto &x : a)
      x.zero();
  }
};
inline void pack_matrix(const Matrix &m, uint64_t *out)
{
  const int ew = elem_words();
  for(size_t k = 0; k < m.a.size(); ++k)
    pack(m.a[k], out + k * ew);
}
inline void unpack_matrix(Matrix &m, int h, int w, const uint64_t *in)
{
  m.resize(h, w);
  const int ew = elem_words();
  for(size_t k = 0; k < m.a.size(); ++k)
    unpack(m.a[k], in + k * ew);
}
} // namespace sdpb_host

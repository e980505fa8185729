// Hand-declared ABI of the GMP 6.x runtime (libgmp.so.10).
//
// This image ships the libgmp runtime but not <gmp.h>.  The reference's
// El::BigFloat is a thin wrapper over GMP's mpf_t (reference:
// src/sdp_solve/SDP_Solver/run/bigint_syrk/fmpz/fmpz_BigFloat_convert.hxx:9,13),
// so the host side of this repo talks to the same library through the
// prototypes below.  Link with  -l:libgmp.so.10 .
//
// Only the entry points that the host solver and the oracle use are declared.
#pragma once
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned long mp_limb_t; // 64-bit limbs on x86-64
typedef long mp_size_t;
typedef long mp_exp_t;
typedef unsigned long mp_bitcnt_t;

typedef struct
{
  int _mp_prec;    // precision in limbs; allocation is _mp_prec + 1 limbs
  int _mp_size;    // |size| = limbs in use, sign = sign of the number
  mp_exp_t _mp_exp; // exponent in limbs
  mp_limb_t *_mp_d;
} __mpf_struct;
typedef __mpf_struct mpf_t[1];
typedef __mpf_struct *mpf_ptr;
typedef const __mpf_struct *mpf_srcptr;

typedef struct
{
  int _mp_alloc;
  int _mp_size;
  mp_limb_t *_mp_d;
} __mpz_struct;
typedef __mpz_struct mpz_t[1];
typedef __mpz_struct *mpz_ptr;
typedef const __mpz_struct *mpz_srcptr;

#define mpf_init2 __gmpf_init2
#define mpf_clear __gmpf_clear
#define mpf_set __gmpf_set
#define mpf_set_ui __gmpf_set_ui
#define mpf_set_si __gmpf_set_si
#define mpf_set_d __gmpf_set_d
#define mpf_set_str __gmpf_set_str
#define mpf_set_z __gmpf_set_z
#define mpf_get_str __gmpf_get_str
#define mpf_get_d __gmpf_get_d
#define mpf_add __gmpf_add
#define mpf_sub __gmpf_sub
#define mpf_mul __gmpf_mul
#define mpf_div __gmpf_div
#define mpf_div_ui __gmpf_div_ui
#define mpf_mul_ui __gmpf_mul_ui
#define mpf_sqrt __gmpf_sqrt
#define mpf_neg __gmpf_neg
#define mpf_abs __gmpf_abs
#define mpf_cmp __gmpf_cmp
#define mpf_cmp_ui __gmpf_cmp_ui
#define mpf_cmp_si __gmpf_cmp_si
#define mpf_cmp_d __gmpf_cmp_d
#define mpf_mul_2exp __gmpf_mul_2exp
#define mpf_div_2exp __gmpf_div_2exp
#define mpf_set_default_prec __gmpf_set_default_prec
#define mpf_get_default_prec __gmpf_get_default_prec
#define mpf_get_prec __gmpf_get_prec
#define mpz_init __gmpz_init
#define mpz_clear __gmpz_clear
#define mpz_set_f __gmpz_set_f
#define mpz_mul __gmpz_mul
#define mpz_add __gmpz_add
#define mpz_addmul __gmpz_addmul
#define mpz_set_ui __gmpz_set_ui
#define mpz_import __gmpz_import
#define mpz_export __gmpz_export
#define mpz_sizeinbase __gmpz_sizeinbase

void __gmpf_init2(mpf_ptr, mp_bitcnt_t);
void __gmpf_clear(mpf_ptr);
void __gmpf_set(mpf_ptr, mpf_srcptr);
void __gmpf_set_ui(mpf_ptr, unsigned long);
void __gmpf_set_si(mpf_ptr, long);
void __gmpf_set_d(mpf_ptr, double);
int __gmpf_set_str(mpf_ptr, const char *, int);
void __gmpf_set_z(mpf_ptr, mpz_srcptr);
char *__gmpf_get_str(char *, mp_exp_t *, int, size_t, mpf_srcptr);
double __gmpf_get_d(mpf_srcptr);
void __gmpf_add(mpf_ptr, mpf_srcptr, mpf_srcptr);
void __gmpf_sub(mpf_ptr, mpf_srcptr, mpf_srcptr);
void __gmpf_mul(mpf_ptr, mpf_srcptr, mpf_srcptr);
void __gmpf_div(mpf_ptr, mpf_srcptr, mpf_srcptr);
void __gmpf_div_ui(mpf_ptr, mpf_srcptr, unsigned long);
void __gmpf_mul_ui(mpf_ptr, mpf_srcptr, unsigned long);
void __gmpf_sqrt(mpf_ptr, mpf_srcptr);
void __gmpf_neg(mpf_ptr, mpf_srcptr);
void __gmpf_abs(mpf_ptr, mpf_srcptr);
int __gmpf_cmp(mpf_srcptr, mpf_srcptr);
int __gmpf_cmp_ui(mpf_srcptr, unsigned long);
int __gmpf_cmp_si(mpf_srcptr, long);
int __gmpf_cmp_d(mpf_srcptr, double);
void __gmpf_mul_2exp(mpf_ptr, mpf_srcptr, mp_bitcnt_t);
void __gmpf_div_2exp(mpf_ptr, mpf_srcptr, mp_bitcnt_t);
void __gmpf_set_default_prec(mp_bitcnt_t);
mp_bitcnt_t __gmpf_get_default_prec(void);
mp_bitcnt_t __gmpf_get_prec(mpf_srcptr);

void __gmpz_init(mpz_ptr);
void __gmpz_clear(mpz_ptr);
void __gmpz_set_f(mpz_ptr, mpf_srcptr);
void __gmpz_mul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_add(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_addmul(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_set_ui(mpz_ptr, unsigned long);
void __gmpz_import(mpz_ptr, size_t, int, size_t, int, size_t, const void *);
void *__gmpz_export(void *, size_t *, int, size_t, int, size_t, mpz_srcptr);
size_t __gmpz_sizeinbase(mpz_srcptr, int);

#ifdef __cplusplus
}
#endif

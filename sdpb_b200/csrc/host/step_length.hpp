// step_length: the canonical host restatement of
//   step_length                          run/step/step_length/step_length.cxx:27-46
//   lower_triangular_inverse_congruence  run/step/step_length/lower_triangular_inverse_congruence.cxx:5-18
//   min_eigenvalue                       run/step/step_length/min_eigenvalue.cxx:8-33
//
// min_eigenvalue calls El::HermitianEig of the un-vendored Elemental fork (tridiagonalisation +
// divide and conquer); neither its operation order nor its stopping rule is visible from the
// reference tree, and the reference's goldens pin the step lengths at 2^-99 only
// (test/src/integration_tests/cases/end-to-end.test.cxx:24-27).  This file fixes an order that
// a GPU can follow operation by operation, so that csrc/eig.cuh reproduces it bit for bit:
//
//   * congruence: A <- A L^-T by rows, then A <- L^-1 A by columns, both forward substitutions with
//     k ascending and one division by the pivot (the rule of direction.hpp);
//   * Householder tridiagonalisation, one reflector per column k, written so that every quantity is
//     either a per-row chain (row i of the trailing matrix is owned by one thread) or one
//     sequential sum from an exact zero:
//         x = A(k+1.., k);  tail2 = sum_{i>k+1} x_i^2;  norm2 = tail2 + x_{k+1}^2;  sigma = sqrt(norm2)
//         alpha = -sign(x_{k+1}) sigma;  v = x - alpha e_1;  H = norm2 - alpha x_{k+1}  (= v.v / 2)
//         p_i = (sum_j A(i,j) v_j) / H;  K = ((sum_i p_i v_i) * 0.5) / H;  w = p - K v
//         A(i,j) -= v_hi w_lo;  A(i,j) -= w_hi v_lo,  hi = max(i,j), lo = min(i,j)  (symmetric by construction)
//   * the smallest eigenvalue of the tridiagonal matrix by Laguerre's iteration on its
//     characteristic polynomial, started from the Gershgorin lower bound: for a polynomial with
//     real roots the iterates increase monotonically to the smallest root, the three-term
//     recurrences for p, p', p'' need no division (mpf exponents cannot overflow here), and an
//     iteration costs one division and one square root on top of 6 n multiplications -- where
//     the QL iteration pays a square root and two divisions per rotation.
// Only the smallest eigenvalue is needed (step_length.cxx:38-45).
#pragma once
#include "bigfloat.hpp"

#include <vector>

namespace sdpb_host
{
// B <- B L^{-T}: row r of B solves x L^T = b, i.e. a forward substitution along the row
inline void trsm_lower_transpose_right(const Matrix &L, Matrix &B)
{
  BigFloat t;
  for(int r = 0; r < B.h; ++r)
    for(int j = 0; j < B.w; ++j)
      {
        for(int k = 0; k < j; ++k)
          {
            t = L(j, k);
            t *= B(r, k);
            B(r, j) -= t;
          }
        B(r, j) /= L(j, j);
      }
}
// B <- L^{-1} B, forward substitution down every column (same loop as direction.hpp::trsm_lower_left)
inline void trsm_lower_left_columns(const Matrix &L, Matrix &B)
{
  BigFloat t;
  for(int c = 0; c < B.w; ++c)
    for(int i = 0; i < B.h; ++i)
      {
        for(int k = 0; k < i; ++k)
          {
            t = L(i, k);
            t *= B(k, c);
            B(i, c) -= t;
          }
        B(i, c) /= L(i, i);
      }
}
// lower_triangular_inverse_congruence.cxx:5-18: A := L^{-1} A L^{-T}
inline void lower_triangular_inverse_congruence(const Matrix &L, Matrix &A)
{
  trsm_lower_transpose_right(L, A);
  trsm_lower_left_columns(L, A);
}

// Householder reduction of the symmetric matrix given by the LOWER triangle of A (as
// El::HermitianEig(LOWER, ...) of min_eigenvalue.cxx:28 reads it; the upper triangle is overwritten
// by the mirror image first, A is destroyed) to tridiagonal form: diagonal d[0..n), sub-diagonal
// e[0..n-1).
inline void tridiagonalize(Matrix &A, std::vector<BigFloat> &d, std::vector<BigFloat> &e)
{
  const int n = A.h;
  d.assign(n, BigFloat());
  e.assign(n > 0 ? n - 1 : 0, BigFloat());
  for(int j = 0; j < n; ++j)
    for(int i = 0; i < j; ++i)
      A(i, j) = A(j, i);
  std::vector<BigFloat> v(n), p(n), w(n);
  BigFloat t, tail2, norm2, sigma, alpha, H, K;
  const BigFloat half(0.5);
  for(int k = 0; k + 2 < n; ++k)
    {
      tail2.zero();
      for(int i = k + 2; i < n; ++i)
        {
          t = A(i, k);
          t *= A(i, k);
          tail2 += t;
        }
      d[k] = A(k, k);
      if(tail2.sgn() == 0)
        {
          e[k] = A(k + 1, k); // the column is tridiagonal already
          continue;
        }
      const BigFloat &x1 = A(k + 1, k);
      t = x1;
      t *= x1;
      norm2 = tail2 + t;
      sigma = Sqrt(norm2);
      alpha = x1.sgn() > 0 ? -sigma : sigma;
      for(int i = k + 2; i < n; ++i)
        v[i] = A(i, k);
      v[k + 1] = x1 - alpha;
      t = alpha;
      t *= x1;
      H = norm2 - t;
      for(int i = k + 1; i < n; ++i)
        {
          BigFloat &acc = p[i];
          acc.zero();
          for(int j = k + 1; j < n; ++j)
            {
              t = A(i, j);
              t *= v[j];
              acc += t;
            }
          acc /= H;
        }
      K.zero();
      for(int i = k + 1; i < n; ++i)
        {
          t = p[i];
          t *= v[i];
          K += t;
        }
      K *= half;
      K /= H;
      for(int i = k + 1; i < n; ++i)
        {
          t = K;
          t *= v[i];
          w[i] = p[i] - t;
        }
      for(int lo = k + 1; lo < n; ++lo)
        for(int hi = lo; hi < n; ++hi)
          {
            t = v[hi];
            t *= w[lo];
            A(hi, lo) -= t;
            t = w[hi];
            t *= v[lo];
            A(hi, lo) -= t;
            if(hi != lo)
              A(lo, hi) = A(hi, lo);
          }
      e[k] = alpha;
    }
  if(n >= 2)
    {
      d[n - 2] = A(n - 2, n - 2);
      e[n - 2] = A(n - 1, n - 2);
    }
  if(n >= 1)
    d[n - 1] = A(n - 1, n - 1);
}

// 2^-(prec - 16): the relative size of a Laguerre step at which the iteration stops
inline BigFloat laguerre_epsilon() { return BigFloat(1) >> (unsigned)(working_precision_bits() - 16); }
constexpr int LAGUERRE_MAX_ITERATIONS = 4096;

// Smallest eigenvalue of the symmetric tridiagonal matrix (d, e), n = d.size() >= 1.
// `iterations` (optional) receives the number of Laguerre steps taken.
inline BigFloat tridiagonal_min_eigenvalue(const std::vector<BigFloat> &d, const std::vector<BigFloat> &e,
                                           int *iterations = nullptr)
{
  const int n = (int)d.size();
  if(iterations)
    *iterations = 0;
  if(n == 1)
    return d[0];
  // Gershgorin: lo <= lambda_min, and scale >= |lambda| for every eigenvalue
  BigFloat lo, scale, r, g, t;
  for(int k = 0; k < n; ++k)
    {
      r.zero();
      if(k > 0)
        r += Abs(e[k - 1]);
      if(k + 1 < n)
        r += Abs(e[k]);
      g = d[k] - r;
      if(k == 0 || g < lo)
        lo = g;
      t = Abs(d[k]) + r;
      if(t > scale)
        scale = t;
    }
  if(scale.sgn() == 0)
    return BigFloat();
  std::vector<BigFloat> e2(n - 1);
  for(int k = 0; k + 1 < n; ++k)
    {
      e2[k] = e[k];
      e2[k] *= e[k];
    }
  const BigFloat tol = scale * laguerre_epsilon();
  const BigFloat big_n((long)n), big_n1((long)(n - 1));
  BigFloat x = lo - tol;
  BigFloat p, q, s, pp, qp, sp, pn, qn, sn, dk, qq, ps, disc, den, a;
  for(int it = 0; it < LAGUERRE_MAX_ITERATIONS; ++it)
    {
      // p = det(T_k - x), q = p', s = p'' by the three-term recurrence
      pp = BigFloat(1);
      qp.zero();
      sp.zero();
      p = d[0] - x;
      q = BigFloat(-1);
      s.zero();
      for(int k = 1; k < n; ++k)
        {
          dk = d[k] - x;
          pn = dk * p;
          t = e2[k - 1];
          t *= pp;
          pn -= t;
          qn = dk * q;
          qn -= p;
          t = e2[k - 1];
          t *= qp;
          qn -= t;
          sn = dk * s;
          sn -= q;
          sn -= q;
          t = e2[k - 1];
          t *= sp;
          sn -= t;
          pp = p;
          qp = q;
          sp = s;
          p = pn;
          q = qn;
          s = sn;
        }
      // left of the spectrum p > 0 > q; anything else means x has reached the smallest root to
      // within the rounding of the recurrence
      if(p.sgn() <= 0 || q.sgn() >= 0)
        break;
      if(iterations)
        ++*iterations;
      // Laguerre: x += n p / (sqrt((n-1) (n (q^2 - p s) - q^2)) - q)
      qq = q * q;
      ps = p * s;
      disc = qq - ps;
      disc *= big_n;
      disc -= qq;
      disc *= big_n1;
      if(disc.sgn() < 0)
        disc.zero();
      den = Sqrt(disc);
      den -= q;
      a = big_n * p;
      a /= den;
      x += a;
      if(a <= tol)
        break;
    }
  return x;
}

// min eigenvalue of L^{-1} dM L^{-T} for one block (dM is copied)
inline BigFloat block_min_eigenvalue(const Matrix &L, const Matrix &dM, int *iterations = nullptr)
{
  Matrix A(dM);
  lower_triangular_inverse_congruence(L, A);
  std::vector<BigFloat> d, e;
  tridiagonalize(A, d, e);
  return tridiagonal_min_eigenvalue(d, e, iterations);
}
} // namespace sdpb_host

// Smallest eigenvalue of a symmetric BigFloat matrix, for step_length
// (reference src/sdp_solve/SDP_Solver/run/step/step_length/min_eigenvalue.cxx:8-33,
// which calls El::HermitianEig from the un-vendored Elemental fork).  Elemental
// reduces to tridiagonal form and runs a divide-and-conquer / QR eigen-solver;
// neither its operation order nor its stopping rule is visible from the
// reference tree, and the reference's own goldens pin the result to 2^-99 only
// (test/src/integration_tests/cases/end-to-end.test.cxx:24-27).  This file does
// the textbook equivalent at working precision: Householder tridiagonalisation
// followed by the implicit-shift QL iteration on the tridiagonal matrix, run to
// the working precision.  It is host code off the hot path.
#pragma once
#include "bigfloat.hpp"

#include <vector>

namespace sdpb_host
{
// Householder reduction of the symmetric matrix A (only its lower triangle is
// read; A is destroyed) to tridiagonal form: diagonal d[0..n), sub-diagonal
// e[0..n-1).
inline void tridiagonalize(Matrix &A, std::vector<BigFloat> &d, std::vector<BigFloat> &e)
{
  const int n = A.h;
  d.assign(n, BigFloat());
  e.assign(n > 0 ? n - 1 : 0, BigFloat());
  // work on a full symmetric copy
  for(int j = 0; j < n; ++j)
    for(int i = 0; i < j; ++i)
      A(i, j) = A(j, i);
  std::vector<BigFloat> v(n), p(n), w(n);
  BigFloat t, two(2);
  for(int k = 0; k + 2 < n; ++k)
    {
      // x = A[k+1.., k]
      BigFloat norm2;
      for(int i = k + 1; i < n; ++i)
        {
          t = A(i, k);
          t *= A(i, k);
          norm2 += t;
        }
      BigFloat tail2 = norm2;
      t = A(k + 1, k);
      t *= A(k + 1, k);
      tail2 -= t;
      if(norm2.sgn() == 0 || tail2.sgn() <= 0)
        {
          // already tridiagonal in this column
          d[k] = A(k, k);
          e[k] = A(k + 1, k);
          continue;
        }
      BigFloat alpha = Sqrt(norm2);
      if(A(k + 1, k).sgn() > 0)
        alpha = -alpha;
      // v = x - alpha e1, normalised
      for(int i = k + 1; i < n; ++i)
        v[i] = A(i, k);
      v[k + 1] -= alpha;
      BigFloat vn2;
      for(int i = k + 1; i < n; ++i)
        {
          t = v[i];
          t *= v[i];
          vn2 += t;
        }
      const BigFloat vn = Sqrt(vn2);
      for(int i = k + 1; i < n; ++i)
        v[i] /= vn;
      // p = A_sub v ; K = v.p ; w = p - K v ; A_sub -= 2 v w^T + 2 w v^T
      BigFloat K;
      for(int i = k + 1; i < n; ++i)
        {
          BigFloat acc;
          for(int j = k + 1; j < n; ++j)
            {
              t = A(i, j);
              t *= v[j];
              acc += t;
            }
          p[i] = acc;
          t = acc;
          t *= v[i];
          K += t;
        }
      for(int i = k + 1; i < n; ++i)
        {
          t = K;
          t *= v[i];
          w[i] = p[i] - t;
        }
      for(int j = k + 1; j < n; ++j)
        for(int i = k + 1; i < n; ++i)
          {
            t = v[i];
            t *= w[j];
            BigFloat u = w[i];
            u *= v[j];
            t += u;
            t *= two;
            A(i, j) -= t;
          }
      d[k] = A(k, k);
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

// All eigenvalues of the symmetric tridiagonal matrix (d, e) by the QL
// iteration with implicit (Wilkinson) shifts; d is overwritten by the
// eigenvalues (unordered).  Returns false if an eigenvalue failed to converge.
inline bool tridiagonal_eigenvalues(std::vector<BigFloat> &d, std::vector<BigFloat> e_in)
{
  const int n = (int)d.size();
  if(n <= 1)
    return true;
  std::vector<BigFloat> e(n);
  for(int i = 0; i + 1 < n; ++i)
    e[i] = e_in[i];
  const BigFloat eps = BigFloat(1) >> (unsigned)(working_precision_bits() - 2);
  const BigFloat one(1), two(2);
  const int max_iter = 60 + working_precision_bits() / 8;
  for(int l = 0; l < n; ++l)
    {
      int iter = 0;
      for(;;)
        {
          int m = l;
          for(; m + 1 < n; ++m)
            {
              const BigFloat dd = Abs(d[m]) + Abs(d[m + 1]);
              if(Abs(e[m]) <= eps * dd)
                break;
            }
          if(m == l)
            break;
          if(++iter > max_iter)
            return false;
          // shift
          BigFloat g = (d[l + 1] - d[l]) / (two * e[l]);
          BigFloat r = Sqrt(g * g + one);
          g = d[m] - d[l] + e[l] / (g + (g.sgn() >= 0 ? r : -r));
          BigFloat s(1), c(1), p(0);
          int i = m - 1;
          bool underflow = false;
          for(; i >= l; --i)
            {
              BigFloat f = s * e[i];
              const BigFloat b = c * e[i];
              r = Sqrt(f * f + g * g);
              e[i + 1] = r;
              if(r.sgn() == 0)
                {
                  d[i + 1] -= p;
                  e[m].zero();
                  underflow = true;
                  break;
                }
              s = f / r;
              c = g / r;
              g = d[i + 1] - p;
              r = (d[i] - g) * s + two * c * b;
              p = s * r;
              d[i + 1] = g + p;
              g = c * r - b;
            }
          if(underflow)
            continue;
          d[l] -= p;
          e[l] = g;
          e[m].zero();
        }
    }
  return true;
}

// min eigenvalue of the symmetric matrix A (destroyed)
inline BigFloat min_eigenvalue_symmetric(Matrix &A)
{
  std::vector<BigFloat> d, e;
  tridiagonalize(A, d, e);
  if(!tridiagonal_eigenvalues(d, e))
    throw std::runtime_error("min_eigenvalue: the QL iteration did not converge");
  BigFloat m = d[0];
  for(size_t i = 1; i < d.size(); ++i)
    if(d[i] < m)
      m = d[i];
  return m;
}
} // namespace sdpb_host

// ORACLE — TEST INFRASTRUCTURE ONLY.  Nothing in the product path may include,
// link or call this file; only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs do.
//
// CPU restatement of the reference's Schur-complement hot path on GMP mpf
// scalars (the real libgmp, i.e. the arithmetic the reference itself uses).
// Each function cites the reference source it follows.  Where the reference
// delegates to the (un-vendored, un-pinned) Elemental fork — Cholesky, Trsm,
// Syrk, Gemm on BigFloat — the operation ORDER is not observable from
// /root/reference; this file fixes the canonical order documented in
// DESIGN.md §3 ("every element is updated in ascending k, one mpf op per
// reference operator").  Last-limb parity with a real Elemental build is
// therefore unpinned; parity with the reference's own 2^-99 goldens is pinned
// by tests/test_golden_trajectory.py through the host solver.
#pragma once
#include "../sdpb_b200/csrc/host/bigfloat.hpp"

#include <array>
#include <chrono>
#include <string>
#include <vector>

namespace oracle
{
using sdpb_host::BigFloat;
using sdpb_host::Matrix;

struct BlockShape
{
  int m; // dimensions[j]            (reference Block_Info.hxx:23)
  int n; // num_points[j] = degree+1 (reference Block_Info.hxx:24)
  int schur_size() const { return n * m * (m + 1) / 2; } // Block_Info.hxx:54-58
  int psd_size(int parity) const                          // Block_Info.hxx:83-95
  {
    const int even = m * ((n + 1) / 2);
    return parity == 0 ? even : m * n - even;
  }
  int pairing_size() const { return m * n; } // Block_Info.hxx:69-74
  int basis_height(int parity) const         // Block_Info.hxx:110-115
  {
    const int degree = n - 1;
    return (degree + parity) / 2 + 1 - parity;
  }
};

// A = L L^T, lower, in place.  Canonical order: right-looking; column j is
// finished by  l_jj = sqrt(a_jj), l_ij = a_ij / l_jj, then
// a_ik -= l_ij * l_kj for j < k <= i.  Every element therefore receives its
// updates in ascending j.  Returns -1 on success, else the index of the first
// non-positive pivot (El::Cholesky throws there; cholesky_decomposition.cxx:14-26).
inline int cholesky_lower(Matrix &A)
{
  const int s = A.h;
  for(int j = 0; j < s; ++j)
    {
      if(A(j, j).sgn() <= 0)
        return j;
      A(j, j) = Sqrt(A(j, j));
      for(int i = j + 1; i < s; ++i)
        A(i, j) /= A(j, j);
#pragma omp parallel if(s - j > 48)
      {
        BigFloat tprod;
#pragma omp for schedule(dynamic, 4)
        for(int k = j + 1; k < s; ++k)
          for(int i = k; i < s; ++i)
            {
              tprod = A(i, j);
              tprod *= A(k, j);
              A(i, k) -= tprod;
            }
      }
    }
  for(int j = 0; j < s; ++j)
    for(int i = 0; i < j; ++i)
      A(i, j).zero();
  return -1;
}

// A = U^T U, upper, in place (initialize_schur_complement_solver.cxx:98).
// Same recurrence on the transposed storage.
inline int cholesky_upper(Matrix &A)
{
  const int s = A.h;
  for(int j = 0; j < s; ++j)
    {
      if(A(j, j).sgn() <= 0)
        return j;
      A(j, j) = Sqrt(A(j, j));
      for(int i = j + 1; i < s; ++i)
        A(j, i) /= A(j, j);
#pragma omp parallel if(s - j > 48)
      {
        BigFloat tprod;
#pragma omp for schedule(dynamic, 4)
        for(int k = j + 1; k < s; ++k)
          for(int i = k; i < s; ++i)
            {
              tprod = A(j, i);
              tprod *= A(j, k);
              A(k, i) -= tprod;
            }
      }
    }
  for(int j = 0; j < s; ++j)
    for(int i = j + 1; i < s; ++i)
      A(i, j).zero();
  return -1;
}

// B <- L^{-1} B, L lower non-unit (El::Trsm LEFT LOWER NORMAL NON_UNIT,
// compute_Q.cxx:48-51, compute_A_X_inv.cxx:21).  Forward substitution,
// x_i = (b_i - sum_{k<i} l_ik x_k) / l_ii with k ascending.
inline void trsm_lower(const Matrix &L, Matrix &B)
{
#pragma omp parallel
  {
    BigFloat prod;
#pragma omp for schedule(static)
    for(int c = 0; c < B.w; ++c)
      for(int i = 0; i < B.h; ++i)
        {
          for(int k = 0; k < i; ++k)
            {
              prod = L(i, k);
              prod *= B(k, c);
              B(i, c) -= prod;
            }
          B(i, c) /= L(i, i);
        }
  }
}

// C = A^T A, lower triangle by dot products (l ascending, starting from an
// exact zero), then mirrored (El::Syrk LOWER TRANSPOSE + MakeSymmetric,
// compute_A_X_inv.cxx:28-30).
inline void syrk_lower_transpose(const Matrix &A, Matrix &C)
{
  C.resize(A.w, A.w);
  BigFloat prod;
  for(int j = 0; j < A.w; ++j)
    for(int i = j; i < A.w; ++i)
      {
        BigFloat &c = C(i, j);
        for(int l = 0; l < A.h; ++l)
          {
            prod = A(l, i);
            prod *= A(l, j);
            c += prod;
          }
      }
  for(int j = 0; j < A.w; ++j)
    for(int i = 0; i < j; ++i)
      C(i, j) = C(j, i);
}

// C = op(A) B by dot products, l ascending (El::Gemm, compute_A_Y.cxx:32,35)
inline void gemm(bool transposeA, const Matrix &A, const Matrix &B, Matrix &C)
{
  const int M = transposeA ? A.w : A.h, K = transposeA ? A.h : A.w;
  C.resize(M, B.w);
  BigFloat prod;
  for(int j = 0; j < B.w; ++j)
    for(int i = 0; i < M; ++i)
      {
        BigFloat &c = C(i, j);
        for(int l = 0; l < K; ++l)
          {
            prod = transposeA ? A(l, i) : A(i, l);
            prod *= B(l, j);
            c += prod;
          }
      }
}

// bases_blocks[p] = I_m (x) bilinear_bases[p]  (SDP/set_bases_blocks.cxx:24-47)
inline void make_bases_block(const BlockShape &sh, int parity,
                             const Matrix &basis, Matrix &V)
{
  V.resize(sh.psd_size(parity), sh.pairing_size());
  for(int row = 0; row < V.h; ++row)
    for(int col = 0; col < V.w; ++col)
      if(row / basis.h == col / basis.w)
        V(row, col) = basis(row % basis.h, col % basis.w);
}

// A_X_inv matrix = V^T X^{-1} V for one block/parity
// (compute_A_X_inv.cxx:17-30): T = L_X^{-1} V, A = T^T T.
inline void compute_A_X_inv(const Matrix &X_cholesky, const Matrix &V,
                            Matrix &out)
{
  Matrix T = V;
  trsm_lower(X_cholesky, T);
  syrk_lower_transpose(T, out);
}
// A_Y matrix = V^T Y V (compute_A_Y.cxx:30-45): YV, V^T(YV), lower mirrored.
inline void compute_A_Y(const Matrix &Y, const Matrix &V, Matrix &out)
{
  Matrix YV;
  gemm(false, Y, V, YV);
  gemm(true, V, YV, out);
  for(int j = 0; j < out.w; ++j)
    for(int i = 0; i < j; ++i)
      out(i, j) = out(j, i);
}

// Schur complement block (compute_schur_complement.cxx:31-124).  AX[p], AY[p]
// are the mn x mn matrices above; the reference's tile accessors are
//   A_X_inv[p][j][cb][rb](row,col) = AX[p](cb*n+row, rb*n+col)  (compute_A_X_inv.cxx:45-55)
//   A_Y   [p][j][cb][rb](row,col) = AY[p](cb*n+col, rb*n+row)  (compute_A_Y.cxx:51-63, transposed)
// Only the lower triangle survives MakeSymmetric(LOWER) (:121).
inline void compute_schur_block(const BlockShape &sh,
                                const std::array<Matrix, 2> &AX,
                                const std::array<Matrix, 2> &AY, Matrix &S)
{
  const int n = sh.n, m = sh.m;
  S.resize(sh.schur_size(), sh.schur_size());
  BigFloat element, product, four(4);
  auto ax = [&](int p, int cb, int rb, int row, int col) -> const BigFloat & {
    return AX[p](cb * n + row, rb * n + col);
  };
  auto ay = [&](int p, int cb, int rb, int row, int col) -> const BigFloat & {
    return AY[p](cb * n + col, rb * n + row);
  };
  for(int c0 = 0; c0 < m; ++c0)
    for(int r0 = 0; r0 <= c0; ++r0)
      {
        const int roff = (c0 * (c0 + 1) / 2 + r0) * n;
        for(int c1 = 0; c1 < m; ++c1)
          for(int r1 = 0; r1 <= c1; ++r1)
            {
              const int coff = (c1 * (c1 + 1) / 2 + r1) * n;
              for(int row = 0; row < n; ++row)
                for(int col = 0; col < n; ++col)
                  {
                    if(roff + row < coff + col)
                      continue; // upper triangle is overwritten by the mirror
                    element.zero();
                    for(int p = 0; p < 2; ++p)
                      {
                        product = ax(p, c0, r1, row, col);
                        product *= ay(p, c1, r0, row, col);
                        element += product;
                        product = ax(p, r0, r1, row, col);
                        product *= ay(p, c1, c0, row, col);
                        element += product;
                        product = ax(p, c0, c1, row, col);
                        product *= ay(p, r1, r0, row, col);
                        element += product;
                        product = ax(p, r0, c1, row, col);
                        product *= ay(p, r1, c0, row, col);
                        element += product;
                      }
                    element /= four;
                    S(roff + row, coff + col) = element;
                  }
            }
      }
  for(int j = 0; j < S.w; ++j)
    for(int i = 0; i < j; ++i)
      S(i, j) = S(j, i);
}

// Matrix_Normalizer (bigint_syrk/Matrix_Normalizer.cxx:75-139): per-block
// partial sums of squares (rows ascending) ...
inline void column_norm_partials(const std::vector<Matrix> &P_blocks, int N,
                                 std::vector<std::vector<BigFloat>> &part)
{
  const int J = (int)P_blocks.size();
  part.assign(J, std::vector<BigFloat>());
#pragma omp parallel
  {
    BigFloat prod;
#pragma omp for schedule(dynamic)
    for(int j = 0; j < J; ++j)
      {
        const Matrix &blk = P_blocks[j];
        part[j].assign(N, BigFloat());
        for(int c = 0; c < N; ++c)
          for(int r = 0; r < blk.h; ++r)
            {
              prod = blk(r, c);
              prod *= blk(r, c);
              part[j][c] += prod;
            }
      }
  }
}
// ... partials added in (global) block order (the reference's AllReduce leaves
// the cross-rank order open, Matrix_Normalizer.cxx:131), then sqrt (:136).
// The canonical order of a sum over the (GLOBAL) blocks of per-block rows: the blocks are
// taken in groups of BLOCK_SUM_GROUP consecutive global indices; every group is summed from an
// exact zero in ascending block order, then the group sums are added, again from zero and
// ascending.  (The reference's single-rank loop runs one accumulator through all blocks,
// Matrix_Normalizer.cxx:82-88,126-129, and leaves the cross-rank order to MPI_Allreduce, :131;
// a single chain is the one thing a GPU cannot shorten, and it grows with the number of GPUs --
// 4800 dependent mpf_adds per column at 8 x 600 blocks.  Two levels keep the order fixed and
// independent of the sharding, and cut the chain to 64 + J/64.)
constexpr int BLOCK_SUM_GROUP = 64;
inline void ordered_block_sum(const std::vector<std::vector<BigFloat>> &part, int N, std::vector<BigFloat> &total)
{
  total.assign(N, BigFloat());
  std::vector<BigFloat> group(N);
  for(size_t j0 = 0; j0 < part.size(); j0 += BLOCK_SUM_GROUP)
    {
      for(int c = 0; c < N; ++c)
        group[c].zero();
      for(size_t j = j0; j < part.size() && j < j0 + BLOCK_SUM_GROUP; ++j)
        for(int c = 0; c < N; ++c)
          group[c] += part[j][c];
      for(int c = 0; c < N; ++c)
        total[c] += group[c];
    }
}
inline void norms_from_partials(const std::vector<std::vector<BigFloat>> &part, int N,
                                std::vector<BigFloat> &norms)
{
  std::vector<BigFloat> total;
  ordered_block_sum(part, N, total);
  norms.assign(N, BigFloat());
  for(int c = 0; c < N; ++c)
    if(total[c].sgn() > 0)
      norms[c] = Sqrt(total[c]);
}
inline void column_norms(const std::vector<Matrix> &P_blocks, int N,
                         std::vector<BigFloat> &norms)
{
  std::vector<std::vector<BigFloat>> part;
  column_norm_partials(P_blocks, N, part);
  norms_from_partials(part, N, norms);
}

// Exact Q' = P'^T P' on truncated integers (bigint_syrk_blas; the CRT/BLAS
// machinery computes exactly this integer, Readme.md:27-55), upper triangle
// Qz[j*N + i], i <= j (caller owns the mpz's: N*N initialised entries).
inline void exact_syrk_upper_integer(const std::vector<Matrix> &Pn_blocks, int N,
                                     std::vector<__mpz_struct> &Qz)
{
  size_t rows = 0;
  for(const auto &b : Pn_blocks)
    rows += b.h;
  std::vector<__mpz_struct> z(rows * (size_t)N);
  size_t r0 = 0;
  for(const auto &b : Pn_blocks)
    {
#pragma omp parallel for schedule(static)
      for(int r = 0; r < b.h; ++r)
        for(int c = 0; c < N; ++c)
          {
            __mpz_struct *p = &z[(r0 + r) * N + c];
            mpz_init(p);
            mpz_set_f(p, b(r, c).v); // truncates toward zero (:13)
          }
      r0 += b.h;
    }
  for(int j = 0; j < N; ++j)
    for(int i = 0; i <= j; ++i)
      mpz_set_ui(&Qz[(size_t)j * N + i], 0);
  // exact integer sums are order-free: rows are taken in chunks that stay in the caches while
  // all N(N+1)/2 pairs of columns visit them (one pass over all rows per pair streams the whole
  // matrix from DRAM N^2/2 times: 27 s instead of 13 s at K = 36 000, N = 300 on 16 cores)
  const size_t chunk = 96;
  for(size_t rb = 0; rb < rows; rb += chunk)
    {
      const size_t re = std::min(rows, rb + chunk);
#pragma omp parallel for schedule(dynamic, 1)
      for(int j = 0; j < N; ++j)
        for(int i = 0; i <= j; ++i)
          {
            __mpz_struct *acc = &Qz[(size_t)j * N + i];
            for(size_t r = rb; r < re; ++r)
              mpz_addmul(acc, &z[r * N + i], &z[r * N + j]);
          }
    }
  for(auto &p : z)
    mpz_clear(&p);
}
// ... converted back with fmpz_get_mpf semantics (fmpz_BigFloat_convert.hxx:9).
inline void exact_syrk_upper(const std::vector<Matrix> &Pn_blocks, int N,
                             Matrix &Q)
{
  Q.resize(N, N);
  std::vector<__mpz_struct> Qz((size_t)N * N);
  for(auto &q : Qz)
    mpz_init(&q);
  exact_syrk_upper_integer(Pn_blocks, N, Qz);
  for(int j = 0; j < N; ++j)
    for(int i = 0; i <= j; ++i)
      mpf_set_z(Q(i, j).v, &Qz[(size_t)j * N + i]);
  for(auto &q : Qz)
    mpz_clear(&q);
}

struct SchurOutputs
{
  std::vector<Matrix> schur_complement_cholesky; // L_j
  std::vector<Matrix> schur_off_diagonal;        // P_j = L_j^{-1} B_j (after the normalise/restore round trip)
  Matrix Q;                                      // upper Cholesky factor of Q
  std::vector<BigFloat> norms;
  std::string error;                             // empty on success
  double cholesky_Q_ms = 0, syrk_ms = 0, block_ms = 0; // wall clock of the three big parts
};

// ---- compute_Q + Cholesky(Q) in stages (compute_Q.cxx:134-151,
// initialize_schur_complement_solver.cxx:89-103).  The single-process driver
// compute_Q_and_factor below runs them back to back; the sharded model used by
// the world_size-2 tests runs stage 1 and 2 per rank and exchanges in between.

// stage 1 (compute_Q.cxx:20-54): per local block Cholesky(S_j), P_j = L_j^{-1} B_j
inline void factor_and_solve_blocks(const std::vector<Matrix> &S, const std::vector<Matrix> &B,
                                    SchurOutputs &out)
{
  const size_t J = S.size();
  const auto t_begin = std::chrono::steady_clock::now();
  out.schur_complement_cholesky.resize(J);
  out.schur_off_diagonal.resize(J);
  std::vector<int> failed(J, 0);
#pragma omp parallel for schedule(dynamic)
  for(size_t j = 0; j < J; ++j)
    {
      out.schur_complement_cholesky[j] = S[j];
      const int bad = cholesky_lower(out.schur_complement_cholesky[j]);
      if(bad >= 0)
        {
          failed[j] = 1;
          continue;
        }
      out.schur_off_diagonal[j] = B[j];
      trsm_lower(out.schur_complement_cholesky[j], out.schur_off_diagonal[j]);
    }
  for(size_t j = 0; j < J; ++j)
    if(failed[j])
      {
        out.error = "Error when computing Cholesky decomposition of block_"
                    + std::to_string(j);
        return;
      }
  out.block_ms = std::chrono::duration<double, std::milli>(
                   std::chrono::steady_clock::now() - t_begin)
                   .count();
}
// stage 2 (Matrix_Normalizer.cxx:174-190): P' = (P / norm) << prec in place
inline void normalize_and_shift(std::vector<Matrix> &P, const std::vector<BigFloat> &norms, int N)
{
  const int prec = sdpb_host::working_precision_bits();
#pragma omp parallel for schedule(dynamic)
  for(size_t jb = 0; jb < P.size(); ++jb)
    for(int c = 0; c < N; ++c)
      {
        Matrix &blk = P[jb];
        if(norms[c].sgn() == 0)
          continue;
        for(int r = 0; r < blk.h; ++r)
          blk(r, c) = (blk(r, c) / norms[c]) << (unsigned)prec;
      }
}
// stage 3: check_normalized_Q_diagonal (compute_Q.cxx:65-91), restore_P
// (Matrix_Normalizer.cxx:210-226), restore_Q upper (:245-265), Cholesky(UPPER, Q);
// out.Q holds Q' (as BigFloat) on entry.
inline void restore_and_factor_Q(SchurOutputs &out, int N)
{
  const int prec = sdpb_host::working_precision_bits();
  {
    const BigFloat one(1), eps = BigFloat(1) >> (unsigned)(prec / 2);
    for(int i = 0; i < N; ++i)
      {
        const BigFloat should_be_one = out.Q(i, i) >> (unsigned)(2 * prec);
        const BigFloat diff = Abs(should_be_one - one);
        if(!(diff < eps))
          {
            out.error = "Normalized Q should have ones on diagonal. For i = "
                        + std::to_string(i);
            return;
          }
      }
  }
#pragma omp parallel for schedule(dynamic)
  for(size_t jb = 0; jb < out.schur_off_diagonal.size(); ++jb)
    for(int c = 0; c < N; ++c)
      {
        Matrix &blk = out.schur_off_diagonal[jb];
        if(out.norms[c].sgn() == 0)
          continue;
        for(int r = 0; r < blk.h; ++r)
          blk(r, c) = (blk(r, c) >> (unsigned)prec) * out.norms[c];
      }
  for(int j = 0; j < N; ++j)
    for(int i = 0; i <= j; ++i)
      out.Q(i, j) = (out.Q(i, j) >> (unsigned)(2 * prec)) * out.norms[i]
                    * out.norms[j];
  const auto tq = std::chrono::steady_clock::now();
  const int bad = cholesky_upper(out.Q);
  out.cholesky_Q_ms = std::chrono::duration<double, std::milli>(
                        std::chrono::steady_clock::now() - tq)
                        .count();
  if(bad >= 0)
    out.error = "Error when computing Cholesky(Q)";
}

inline void compute_Q_and_factor(const std::vector<Matrix> &S,
                                 const std::vector<Matrix> &B, int N,
                                 SchurOutputs &out)
{
  factor_and_solve_blocks(S, B, out);
  if(!out.error.empty())
    return;
  // syrk_Q (compute_Q.cxx:94-132)
  const auto ts = std::chrono::steady_clock::now();
  column_norms(out.schur_off_diagonal, N, out.norms);
  normalize_and_shift(out.schur_off_diagonal, out.norms, N);
  exact_syrk_upper(out.schur_off_diagonal, N, out.Q);
  restore_and_factor_Q(out, N);
  out.syrk_ms = std::chrono::duration<double, std::milli>(
                  std::chrono::steady_clock::now() - ts)
                  .count()
                - out.cholesky_Q_ms;
}

// ---- solve_schur_complement_equation -------------------
// Reference: compute_search_direction/solve_schur_complement_equation.cxx:16-79
// (lower_triangular_solve :23, Gemv TRANSPOSE -1 per block and the dy sum
// :31-60, cholesky::SolveAfter(UPPER, Q) :64-65, Gemv NORMAL +1 :69-75,
// lower_triangular_transpose_solve :78).  Trsv/Gemv/SolveAfter live in the
// un-vendored Elemental fork, so the order inside them is this repo's canonical
// one: dot products start from an exact zero and ascend; a
// substitution applies the solved unknowns in the order they become available
// (forward: k ascending, backward: k descending), then divides by the pivot.
//
// stage A, per local block: dx_j <- L_j^{-1} dx_j ; part_j[c] = -(sum_r P_j(r,c) dx_j(r))
inline void schur_solve_forward(const SchurOutputs &f, std::vector<Matrix> &dx, int N,
                                std::vector<std::vector<BigFloat>> &part)
{
  const size_t J = dx.size();
  part.assign(J, std::vector<BigFloat>(N));
#pragma omp parallel for schedule(dynamic)
  for(size_t j = 0; j < J; ++j)
    {
      const Matrix &L = f.schur_complement_cholesky[j], &P = f.schur_off_diagonal[j];
      Matrix &x = dx[j];
      BigFloat prod, acc;
      for(int i = 0; i < x.h; ++i)
        {
          for(int k = 0; k < i; ++k)
            {
              prod = L(i, k);
              prod *= x(k, 0);
              x(i, 0) -= prod;
            }
          x(i, 0) /= L(i, i);
        }
      for(int c = 0; c < N; ++c)
        {
          acc.zero();
          for(int r = 0; r < P.h; ++r)
            {
              prod = P(r, c);
              prod *= x(r, 0);
              acc += prod;
            }
          part[j][c] = -acc;
        }
    }
}
// stage B, replicated: dy += sum_j part_j (ordered_block_sum over the GLOBAL blocks);
// dy <- U^{-1} U^{-T} dy, Q = U^T U
inline void schur_solve_Q(const Matrix &U, const std::vector<std::vector<BigFloat>> &part_global,
                          Matrix &dy)
{
  const int n = U.h;
  {
    std::vector<BigFloat> total;
    ordered_block_sum(part_global, n, total); // two-level, in global block order (see there)
    for(int c = 0; c < n; ++c)
      dy(c, 0) += total[c];
  }
  BigFloat prod;
  for(int i = 0; i < n; ++i)
    {
      for(int k = 0; k < i; ++k)
        {
          prod = U(k, i);
          prod *= dy(k, 0);
          dy(i, 0) -= prod;
        }
      dy(i, 0) /= U(i, i);
    }
  for(int i = n - 1; i >= 0; --i)
    {
      for(int k = n - 1; k > i; --k)
        {
          prod = U(i, k);
          prod *= dy(k, 0);
          dy(i, 0) -= prod;
        }
      dy(i, 0) /= U(i, i);
    }
}
// stage C, per local block: dx_j += P_j dy ; dx_j <- L_j^{-T} dx_j
inline void schur_solve_backward(const SchurOutputs &f, const Matrix &dy, std::vector<Matrix> &dx)
{
  const size_t J = dx.size();
  const int N = dy.h;
#pragma omp parallel for schedule(dynamic)
  for(size_t j = 0; j < J; ++j)
    {
      const Matrix &L = f.schur_complement_cholesky[j], &P = f.schur_off_diagonal[j];
      Matrix &x = dx[j];
      BigFloat prod, acc;
      for(int r = 0; r < P.h; ++r)
        {
          acc.zero();
          for(int c = 0; c < N; ++c)
            {
              prod = P(r, c);
              prod *= dy(c, 0);
              acc += prod;
            }
          x(r, 0) += acc;
        }
      for(int i = x.h - 1; i >= 0; --i)
        {
          for(int k = x.h - 1; k > i; --k)
            {
              prod = L(k, i);
              prod *= x(k, 0);
              x(i, 0) -= prod;
            }
          x(i, 0) /= L(i, i);
        }
    }
}
inline void solve_schur_complement_equation(const SchurOutputs &f, std::vector<Matrix> &dx, Matrix &dy)
{
  std::vector<std::vector<BigFloat>> part;
  schur_solve_forward(f, dx, dy.h, part);
  schur_solve_Q(f.Q, part, dy);
  schur_solve_backward(f, dy, dx);
}

// ---- scale_multiply_add (the block GEMMs of step() / compute_search_direction)
// Reference: compute_search_direction/scale_multiply_add.cxx:4-16 -- per block
// El::Gemm(NORMAL, NORMAL, alpha, A_b, B_b, beta, C_b); call sites step.cxx:137 (-X Y),
// compute_search_direction.cxx:28 (1, 0) and :60 (-1, 1).  Canonical order (Gemm is
// Elemental's): the dot product from an exact zero with l ascending, then `acc *= alpha`,
// then C = acc (beta == 0) or `C *= beta; C += acc`.
inline void scale_multiply_add_block(const BigFloat &alpha, const Matrix &A, const Matrix &B,
                                     const BigFloat &beta, Matrix &C)
{
  BigFloat acc, prod;
  const bool beta_zero = beta.sgn() == 0;
  for(int j = 0; j < B.w; ++j)
    for(int i = 0; i < A.h; ++i)
      {
        acc.zero();
        for(int l = 0; l < A.w; ++l)
          {
            prod = A(i, l);
            prod *= B(l, j);
            acc += prod;
          }
        acc *= alpha;
        if(beta_zero)
          C(i, j) = acc;
        else
          {
            C(i, j) *= beta;
            C(i, j) += acc;
          }
      }
}
} // namespace oracle

// The search direction of one Newton iteration: the canonical host
// restatement of
//   compute_search_direction      run/step/compute_search_direction.cxx:44-90
//   cholesky_solve                run/step/compute_search_direction/cholesky_solve.cxx:4-13
//   compute_schur_RHS             run/step/compute_search_direction/compute_schur_RHS.cxx:21-86
//   constraint_matrix_weighted_sum run/constraint_matrix_weighted_sum.cxx:14-66
//   Block_Diagonal_Matrix::symmetrize  Block_Diagonal_Matrix.hxx:95-109
// and of the scalar reductions of step() around it (mu, R error, corrector centering:
// step.cxx:137-160, compute_R_error.hxx, frobenius_product_of_sums.cxx).
//
// The operation order inside Elemental's Trsm / Gemm / Dotu is not visible from the reference
// (un-vendored fork), and its results are pinned at 2^-99 only; the order written out here is this
// repo's canonical one, chosen so that no dependent chain is longer than a matrix dimension:
//   * a triangular solve applies the solved unknowns in the order they become available (forward:
//     k ascending, transposed: k DESCENDING), then divides by the pivot -- same rule as the Schur
//     solves (oracle/hotpath_core.hpp);
//   * a sum over all blocks (trace, Frobenius product) is formed per block from an exact zero and
//     the per-block values are added in the two-level order of ordered_sum() -- the rule the
//     column norms use;
//   * inside a block the Frobenius product is summed by columns (rows ascending), then the column
//     sums ascending.
// The CUDA kernels (csrc/direction.cuh) follow the same order; tests compare byte for byte.
#pragma once
#include "sdp.hpp"

#include <functional>

namespace sdpb_host
{
constexpr int SUM_GROUP = 64; // == BLOCK_SUM_GROUP of the kernels and of oracle/hotpath_core.hpp

// sum of `parts` in the canonical two-level order: groups of SUM_GROUP consecutive entries, each
// from an exact zero and ascending, then the group sums, again from zero and ascending
inline BigFloat ordered_sum(const std::vector<BigFloat> &parts)
{
  BigFloat total, group;
  for(size_t i0 = 0; i0 < parts.size(); i0 += SUM_GROUP)
    {
      group.zero();
      for(size_t i = i0; i < parts.size() && i < i0 + SUM_GROUP; ++i)
        group += parts[i];
      total += group;
    }
  return total;
}

// B <- L^{-1} B (forward substitution, k ascending)
inline void trsm_lower_left(const Matrix &L, Matrix &B)
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
// B <- L^{-T} B (back substitution; unknown i receives x_k for k = h-1 down to i+1)
inline void trsm_lower_transpose_left(const Matrix &L, Matrix &B)
{
  BigFloat t;
  for(int c = 0; c < B.w; ++c)
    for(int i = B.h - 1; i >= 0; --i)
      {
        for(int k = B.h - 1; k > i; --k)
          {
            t = L(k, i);
            t *= B(k, c);
            B(i, c) -= t;
          }
        B(i, c) /= L(i, i);
      }
}
// cholesky_solve.cxx: Z <- L^{-T} L^{-1} Z per block
inline void cholesky_solve(const std::vector<Matrix> &L, std::vector<Matrix> &Z)
{
#pragma omp parallel for schedule(dynamic)
  for(size_t b = 0; b < Z.size(); ++b)
    {
      trsm_lower_left(L[b], Z[b]);
      trsm_lower_transpose_left(L[b], Z[b]);
    }
}
// Block_Diagonal_Matrix::symmetrize (Block_Diagonal_Matrix.hxx:95-109): A *= 0.5; A += A^T
inline void symmetrize(Matrix &A)
{
  const BigFloat half(0.5);
  for(auto &x : A.a)
    x *= half;
  for(int j = 0; j < A.w; ++j)
    for(int i = 0; i < j; ++i)
      {
        const BigFloat s = A(i, j) + A(j, i);
        A(i, j) = s;
        A(j, i) = s;
      }
  for(int i = 0; i < A.h; ++i)
    A(i, i) += A(i, i);
}

// constraint_matrix_weighted_sum.cxx:14-66: result[2j+parity] = sum_p a_p A_p on block j
inline void constraint_matrix_weighted_sum(const Block_Info &block_info, const std::vector<Matrix> &bilinear_bases,
                                           const std::vector<Matrix> &a, std::vector<Matrix> &result)
{
  const int J = block_info.num_blocks();
#pragma omp parallel for schedule(dynamic)
  for(int j = 0; j < J; ++j)
    {
      const int n = block_info.num_points[j], m = block_info.dimensions[j];
      BigFloat acc, t;
      const BigFloat half(0.5);
      for(int parity = 0; parity < 2; ++parity)
        {
          Matrix &R = result[2 * j + parity];
          const Matrix &bases = bilinear_bases[2 * j + parity];
          const int h = bases.h;
          R.zero();
          for(int cb = 0; cb < m; ++cb)
            for(int rb = 0; rb <= cb; ++rb)
              {
                const int voff = (cb * (cb + 1) / 2 + rb) * n;
                for(int c = 0; c < h; ++c)
                  for(int r = 0; r < h; ++r)
                    {
                      acc.zero();
                      for(int k = 0; k < n; ++k)
                        {
                          t = bases(c, k);
                          t *= a[j](voff + k, 0);
                          t *= bases(r, k);
                          acc += t;
                        }
                      if(cb != rb)
                        acc *= half;
                      R(rb * h + r, cb * h + c) = acc;
                    }
              }
          if(m > 1)
            for(int c = 0; c < R.w; ++c)
              for(int r = c + 1; r < R.h; ++r)
                R(r, c) = R(c, r); // MakeSymmetric(UPPER)
        }
    }
}

// compute_schur_RHS.cxx:21-86: dx = -dual_residues - Tr(A_p Z)
inline void compute_schur_RHS(const Block_Info &block_info, const std::vector<Matrix> &bilinear_bases,
                              const std::vector<Matrix> &dual_residues, const std::vector<Matrix> &Z,
                              std::vector<Matrix> &dx)
{
  const int J = block_info.num_blocks();
#pragma omp parallel for schedule(dynamic)
  for(int j = 0; j < J; ++j)
    {
      const int n = block_info.num_points[j], m = block_info.dimensions[j];
      dx[j] = dual_residues[j];
      for(auto &e : dx[j].a)
        e = -e;
      BigFloat acc, t, zq;
      for(int parity = 0; parity < 2; ++parity)
        {
          const Matrix &bases = bilinear_bases[2 * j + parity];
          const Matrix &Zb = Z[2 * j + parity];
          const int h = bases.h;
          for(int cb = 0; cb < m; ++cb)
            for(int rb = 0; rb <= cb; ++rb)
              {
                const int off = (cb * (cb + 1) / 2 + rb) * n;
                for(int k = 0; k < n; ++k)
                  {
                    // sum_a bases(a,k) * (Z_sub bases)(a,k), Z_sub = Z[rb h .., cb h ..]
                    acc.zero();
                    for(int a = 0; a < h; ++a)
                      {
                        zq.zero();
                        for(int b = 0; b < h; ++b)
                          {
                            t = Zb(rb * h + a, cb * h + b);
                            t *= bases(b, k);
                            zq += t;
                          }
                        zq *= bases(a, k);
                        acc += zq;
                      }
                    dx[j](off + k, 0) -= acc;
                  }
              }
        }
    }
}

// ---- per-block pieces of the scalar reductions of step() -------------------
// trace of every block (from an exact zero, i ascending); mu = -ordered_sum(traces) / rows
inline void block_traces(const std::vector<Matrix> &M, std::vector<BigFloat> &traces)
{
  traces.assign(M.size(), BigFloat());
  for(size_t b = 0; b < M.size(); ++b)
    for(int i = 0; i < M[b].h; ++i)
      traces[b] += M[b](i, i);
}
// compute_R_error.hxx: max |minus_XY + mu I| per block (maxima are exact: order-free)
inline void block_R_errors(const std::vector<Matrix> &minus_XY, const BigFloat &mu, std::vector<BigFloat> &maxima)
{
  maxima.assign(minus_XY.size(), BigFloat());
#pragma omp parallel for schedule(dynamic)
  for(size_t b = 0; b < minus_XY.size(); ++b)
    {
      const Matrix &blk = minus_XY[b];
      BigFloat v;
      for(int j = 0; j < blk.w; ++j)
        for(int i = 0; i < blk.h; ++i)
          {
            v = blk(i, j);
            if(i == j)
              v += mu;
            v = Abs(v);
            if(v > maxima[b])
              maxima[b] = v;
          }
    }
}
// frobenius_product_of_sums.cxx: sum (X + dX)_ij (Y + dY)_ij per block -- by columns (rows
// ascending from an exact zero), then the column sums ascending
inline void block_frobenius_products(const std::vector<Matrix> &X, const std::vector<Matrix> &dX,
                                     const std::vector<Matrix> &Y, const std::vector<Matrix> &dY,
                                     std::vector<BigFloat> &products)
{
  products.assign(X.size(), BigFloat());
#pragma omp parallel for schedule(dynamic)
  for(size_t b = 0; b < X.size(); ++b)
    {
      BigFloat t, u, col;
      for(int j = 0; j < X[b].w; ++j)
        {
          col.zero();
          for(int i = 0; i < X[b].h; ++i)
            {
              t = X[b](i, j) + dX[b](i, j);
              u = Y[b](i, j) + dY[b](i, j);
              t *= u;
              col += t;
            }
          products[b] += col;
        }
    }
}

// C_b = alpha A_b B_b + beta C_b per block (scale_multiply_add.cxx:4-16), alpha in {1,-1},
// beta in {0,1}: dot products from an exact zero, l ascending
typedef std::function<void(int, const std::vector<Matrix> &, const std::vector<Matrix> &, int, std::vector<Matrix> &)>
  Scale_Multiply_Add;
// solve_schur_complement_equation.cxx:16-79 on the factors of the current step
typedef std::function<void(std::vector<Matrix> &, Matrix &)> Schur_Solve;

// compute_search_direction.cxx:44-90.  dX, dY are inputs too in the corrector phase.
inline void compute_search_direction(const Block_Info &block_info, const std::vector<Matrix> &bilinear_bases,
                                     const std::vector<Matrix> &X, const std::vector<Matrix> &Y,
                                     const std::vector<Matrix> &X_cholesky, const std::vector<Matrix> &minus_XY,
                                     const std::vector<Matrix> &primal_residues,
                                     const std::vector<Matrix> &dual_residues, const Matrix &primal_residue_p,
                                     const BigFloat &beta_mu, bool is_corrector_phase,
                                     const Scale_Multiply_Add &scale_multiply_add, const Schur_Solve &schur_solve,
                                     std::vector<Matrix> &dx, std::vector<Matrix> &dX, Matrix &dy,
                                     std::vector<Matrix> &dY)
{
  // R = beta mu I - X Y (predictor) or beta mu I - X Y - dX dY (corrector)
  std::vector<Matrix> R(minus_XY);
  if(is_corrector_phase)
    scale_multiply_add(-1, dX, dY, 1, R);
  for(auto &blk : R)
    for(int i = 0; i < blk.h; ++i)
      blk(i, i) += beta_mu;
  // Z = Symmetrize(X^{-1} (PrimalResidues Y - R))
  std::vector<Matrix> Z(X);
  scale_multiply_add(1, primal_residues, Y, 0, Z);
  for(size_t b = 0; b < Z.size(); ++b)
    for(size_t i = 0; i < Z[b].a.size(); ++i)
      Z[b].a[i] -= R[b].a[i];
  cholesky_solve(X_cholesky, Z);
  for(auto &blk : Z)
    symmetrize(blk);
  compute_schur_RHS(block_info, bilinear_bases, dual_residues, Z, dx);
  dy = primal_residue_p;
  schur_solve(dx, dy);
  // dX = PrimalResidues + sum_p A_p dx[p]
  constraint_matrix_weighted_sum(block_info, bilinear_bases, dx, dX);
  for(size_t b = 0; b < dX.size(); ++b)
    for(size_t i = 0; i < dX[b].a.size(); ++i)
      dX[b].a[i] += primal_residues[b].a[i];
  // dY = Symmetrize(X^{-1} (R - dX Y))
  scale_multiply_add(1, dX, Y, 0, dY);
  for(size_t b = 0; b < dY.size(); ++b)
    for(size_t i = 0; i < dY[b].a.size(); ++i)
      dY[b].a[i] -= R[b].a[i];
  cholesky_solve(X_cholesky, dY);
  for(auto &blk : dY)
    {
      symmetrize(blk);
      for(auto &e : blk.a)
        e = -e;
    }
}
} // namespace sdpb_host

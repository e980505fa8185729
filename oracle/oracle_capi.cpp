// ORACLE — TEST INFRASTRUCTURE ONLY (see hotpath_core.hpp).
// The same call surface as include/sdpb_b200.h, implemented on the CPU with
// libgmp, so tests drive both through identical inputs and compare bytes.
// Build: make -C oracle   (-> oracle/liboracle.so)
#include "hotpath_core.hpp"
#include "../sdpb_b200/csrc/host/direction.hpp"
#include "../sdpb_b200/csrc/host/step_length.hpp"

#include <chrono>
#include <cmath>
#include <cstdio>
#include <memory>

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace oracle;
using sdpb_host::elem_words;
using sdpb_host::pack_matrix;
using sdpb_host::unpack_matrix;

struct oracle_ctx
{
  int prec, N, J;
  std::vector<BlockShape> shapes;
  std::vector<Matrix> B;               // J
  std::vector<Matrix> V;               // 2J bases_blocks
  std::vector<Matrix> X_cholesky;      // 2J
  std::vector<Matrix> Y_cholesky;      // 2J (step_length reads it, row N3)
  std::vector<Matrix> AX, AY;          // 2J
  std::string error;
  double stage_ms[9];
  SchurOutputs shard; // state between the stages of the sharded model; after a step: L_j, P_j, chol(Q)
  // search direction (row N2): what compute_search_direction.cxx:44-90 reads and writes
  sdpb_host::Block_Info block_info;
  std::vector<Matrix> bases;                  // 2J bilinear bases (h x n)
  std::vector<Matrix> X, Y;                   // 2J, pristine inputs of the last step
  std::vector<Matrix> minus_XY, primal_residues, dX, dY; // 2J
  std::vector<Matrix> dual_residues, dx;      // J
  Matrix primal_residue_p, dy;                // N x 1
  bool have_minus_XY = false, have_residues = false, have_direction = false;
};

static void pack_out(const Matrix &m, uint64_t *out)
{
  if(out && m.a.size())
    pack_matrix(m, out);
}

extern "C" {

int oracle_elem_words(int prec_bits)
{
  return (((prec_bits + 63) / 64 + 2) + 2) & ~1;
}

int oracle_create(oracle_ctx **out, int prec_bits, int num_blocks,
                  const int *dims, const int *num_points, int N)
{
  sdpb_host::set_precision(prec_bits);
  auto *c = new oracle_ctx();
  c->prec = prec_bits;
  c->N = N;
  c->J = num_blocks;
  for(int j = 0; j < num_blocks; ++j)
    {
      c->shapes.push_back(BlockShape{dims[j], num_points[j]});
      c->block_info.dimensions.push_back(dims[j]);
      c->block_info.num_points.push_back(num_points[j]);
    }
  c->bases.resize(2 * num_blocks);
  c->X.resize(2 * num_blocks);
  c->Y.resize(2 * num_blocks);
  c->B.resize(num_blocks);
  c->V.resize(2 * num_blocks);
  c->X_cholesky.resize(2 * num_blocks);
  c->Y_cholesky.resize(2 * num_blocks);
  c->AX.resize(2 * num_blocks);
  c->AY.resize(2 * num_blocks);
  for(double &x : c->stage_ms)
    x = 0;
  *out = c;
  return 0;
}
void oracle_destroy(oracle_ctx *c) { delete c; }
const char *oracle_last_error(const oracle_ctx *c) { return c->error.c_str(); }

int oracle_set_block(oracle_ctx *c, int j, const uint64_t *B,
                     const uint64_t *bases_even, const uint64_t *bases_odd)
{
  sdpb_host::set_precision(c->prec);
  const BlockShape &sh = c->shapes[j];
  unpack_matrix(c->B[j], sh.schur_size(), c->N, B);
  for(int p = 0; p < 2; ++p)
    {
      Matrix basis;
      const int h = sh.basis_height(p);
      unpack_matrix(basis, h, sh.n, p == 0 ? bases_even : bases_odd);
      c->bases[2 * j + p] = basis;
      if(h > 0)
        make_bases_block(sh, p, basis, c->V[2 * j + p]);
      else
        c->V[2 * j + p].resize(0, sh.pairing_size());
    }
  return 0;
}

int oracle_cholesky_decomposition(oracle_ctx *c, int which,
                                  const uint64_t *const *A, uint64_t *const *L)
{
  sdpb_host::set_precision(c->prec);
  auto t0 = std::chrono::steady_clock::now();
  int rc = 0;
  std::vector<Matrix> out(2 * c->J);
#pragma omp parallel for schedule(dynamic)
  for(int b = 0; b < 2 * c->J; ++b)
    {
      const int s = c->shapes[b / 2].psd_size(b % 2);
      if(s == 0)
        continue;
      unpack_matrix(out[b], s, s, A[b]);
      (which == 0 ? c->X : c->Y)[b] = out[b]; // the pristine matrix: -XY and the Frobenius product read it
      const int bad = cholesky_lower(out[b]);
      if(bad >= 0)
        {
#pragma omp critical
          if(rc == 0)
            {
              rc = 3;
              c->error
                = std::string("Error when computing Cholesky decomposition of "
                              "Block_Diagonal_Matrix ")
                  + (which == 0 ? "X" : "Y")
                  + ", block index = " + std::to_string(b / 2)
                  + ", parity = " + std::to_string(b % 2);
            }
        }
    }
  if(rc)
    return rc;
  for(int b = 0; b < 2 * c->J; ++b)
    {
      if(L && L[b])
        pack_out(out[b], L[b]);
      (which == 0 ? c->X_cholesky : c->Y_cholesky)[b] = std::move(out[b]);
    }
  c->stage_ms[0] += std::chrono::duration<double, std::milli>(
                      std::chrono::steady_clock::now() - t0)
                      .count();
  return 0;
}

int oracle_compute_bilinear_pairings(oracle_ctx *c, const uint64_t *const *Y,
                                     uint64_t *const *A_X_inv,
                                     uint64_t *const *A_Y)
{
  sdpb_host::set_precision(c->prec);
  auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic)
  for(int b = 0; b < 2 * c->J; ++b)
    {
      const BlockShape &sh = c->shapes[b / 2];
      const int s = sh.psd_size(b % 2), mn = sh.pairing_size();
      if(s == 0)
        {
          c->AX[b].resize(mn, mn);
          c->AY[b].resize(mn, mn);
        }
      else
        {
          Matrix Yb;
          unpack_matrix(Yb, s, s, Y[b]);
          compute_A_X_inv(c->X_cholesky[b], c->V[b], c->AX[b]);
          compute_A_Y(Yb, c->V[b], c->AY[b]);
        }
      if(A_X_inv && A_X_inv[b])
        pack_out(c->AX[b], A_X_inv[b]);
      if(A_Y && A_Y[b])
        pack_out(c->AY[b], A_Y[b]);
    }
  c->stage_ms[1] = std::chrono::duration<double, std::milli>(
                     std::chrono::steady_clock::now() - t0)
                     .count();
  return 0;
}

int oracle_initialize_schur_complement_solver(
  oracle_ctx *c, uint64_t *const *schur_complement_cholesky,
  uint64_t *const *schur_off_diagonal, uint64_t *Q, int32_t *block_timings_ms)
{
  (void)block_timings_ms;
  sdpb_host::set_precision(c->prec);
  auto t0 = std::chrono::steady_clock::now();
  std::vector<Matrix> S(c->J);
#pragma omp parallel for schedule(dynamic)
  for(int j = 0; j < c->J; ++j)
    {
      std::array<Matrix, 2> AX{c->AX[2 * j], c->AX[2 * j + 1]};
      std::array<Matrix, 2> AY{c->AY[2 * j], c->AY[2 * j + 1]};
      compute_schur_block(c->shapes[j], AX, AY, S[j]);
    }
  c->stage_ms[2] = std::chrono::duration<double, std::milli>(
                     std::chrono::steady_clock::now() - t0)
                     .count();
  c->shard = SchurOutputs();
  SchurOutputs &out = c->shard; // kept: oracle_solve_schur_complement_equation reads L_j, P_j, chol(Q)
  compute_Q_and_factor(S, c->B, c->N, out);
  c->stage_ms[3] = out.block_ms;
  c->stage_ms[5] = out.syrk_ms;
  c->stage_ms[7] = out.cholesky_Q_ms;
  if(!out.error.empty())
    {
      c->error = out.error;
      return out.error.find("Normalized Q") != std::string::npos ? 4 : 3;
    }
  for(int j = 0; j < c->J; ++j)
    {
      if(schur_complement_cholesky && schur_complement_cholesky[j])
        pack_out(out.schur_complement_cholesky[j],
                 schur_complement_cholesky[j]);
      if(schur_off_diagonal && schur_off_diagonal[j])
        pack_out(out.schur_off_diagonal[j], schur_off_diagonal[j]);
    }
  if(Q)
    pack_out(out.Q, Q);
  c->stage_ms[8] = std::chrono::duration<double, std::milli>(
                     std::chrono::steady_clock::now() - t0)
                     .count();
  return 0;
}

int oracle_schur_step(oracle_ctx *c, const uint64_t *const *X,
                      const uint64_t *const *Y, uint64_t *const *X_cholesky,
                      uint64_t *const *Y_cholesky, uint64_t *const *A_X_inv,
                      uint64_t *const *A_Y,
                      uint64_t *const *schur_complement_cholesky,
                      uint64_t *const *schur_off_diagonal, uint64_t *Q,
                      int32_t *block_timings_ms)
{
  for(double &x : c->stage_ms)
    x = 0;
  int rc = oracle_cholesky_decomposition(c, 0, X, X_cholesky);
  if(rc)
    return rc;
  rc = oracle_cholesky_decomposition(c, 1, Y, Y_cholesky);
  if(rc)
    return rc;
  rc = oracle_compute_bilinear_pairings(c, Y, A_X_inv, A_Y);
  if(rc)
    return rc;
  return oracle_initialize_schur_complement_solver(
    c, schur_complement_cholesky, schur_off_diagonal, Q, block_timings_ms);
}

// ---- sharded model (multi-GPU semantics on the CPU) ---------------------
// The blocks of one SDP are split over several oracle contexts ("ranks"); the
// two exchanges of DESIGN.md §7 happen in the caller (tests: torch.distributed
// gloo, world_size 2) between these stages.  Q' partials travel as exact
// integers: per entry  [sign (0, 1, ~0), magnitude limbs...]  oracle_qprime_words(prec) words.
int oracle_qprime_words(int prec_bits) { return 2 * ((prec_bits + 63) / 64 + 2) + 4; }

// stage 1: Cholesky X, Y, pairings, S, Cholesky(S_j), L_j^{-1} B_j for the local
// blocks; part[j*N + c] = sum over the rows of local block j of P(r,c)^2
int oracle_shard_stage1(oracle_ctx *c, const uint64_t *const *X, const uint64_t *const *Y,
                        uint64_t *part)
{
  int rc = oracle_cholesky_decomposition(c, 0, X, nullptr);
  if(rc)
    return rc;
  rc = oracle_cholesky_decomposition(c, 1, Y, nullptr);
  if(rc)
    return rc;
  rc = oracle_compute_bilinear_pairings(c, Y, nullptr, nullptr);
  if(rc)
    return rc;
  std::vector<Matrix> S(c->J);
#pragma omp parallel for schedule(dynamic)
  for(int j = 0; j < c->J; ++j)
    {
      std::array<Matrix, 2> AX{c->AX[2 * j], c->AX[2 * j + 1]};
      std::array<Matrix, 2> AY{c->AY[2 * j], c->AY[2 * j + 1]};
      compute_schur_block(c->shapes[j], AX, AY, S[j]);
    }
  c->shard = SchurOutputs();
  factor_and_solve_blocks(S, c->B, c->shard);
  if(!c->shard.error.empty())
    {
      c->error = c->shard.error;
      return 3;
    }
  std::vector<std::vector<BigFloat>> partials;
  column_norm_partials(c->shard.schur_off_diagonal, c->N, partials);
  const int ew = elem_words();
  for(int j = 0; j < c->J; ++j)
    for(int col = 0; col < c->N; ++col)
      sdpb_host::pack(partials[j][col], part + ((size_t)j * c->N + col) * ew);
  return 0;
}
// stage 2: norms from ALL blocks' partials in global block order; normalise the
// local P; exact integer Q' of the local rows
int oracle_shard_stage2(oracle_ctx *c, const uint64_t *part_global, int J_global, uint64_t *qprime)
{
  sdpb_host::set_precision(c->prec);
  const int ew = elem_words(), N = c->N;
  std::vector<std::vector<BigFloat>> partials(J_global, std::vector<BigFloat>(N));
  for(int j = 0; j < J_global; ++j)
    for(int col = 0; col < N; ++col)
      sdpb_host::unpack(partials[j][col], part_global + ((size_t)j * N + col) * ew);
  norms_from_partials(partials, N, c->shard.norms);
  normalize_and_shift(c->shard.schur_off_diagonal, c->shard.norms, N);
  std::vector<__mpz_struct> Qz((size_t)N * N);
  for(auto &q : Qz)
    mpz_init(&q);
  exact_syrk_upper_integer(c->shard.schur_off_diagonal, N, Qz);
  const int W = oracle_qprime_words(c->prec);
  for(size_t e = 0; e < (size_t)N * N; ++e)
    {
      uint64_t *o = qprime + e * W;
      for(int k = 0; k < W; ++k)
        o[k] = 0;
      const int sz = Qz[e]._mp_size, asz = sz < 0 ? -sz : sz;
      if(asz > W - 1)
        {
          c->error = "Q' entry too large for the exchange record";
          return 1;
        }
      o[0] = sz == 0 ? 0 : (sz > 0 ? 1 : ~(uint64_t)0);
      for(int k = 0; k < asz; ++k)
        o[1 + k] = Qz[e]._mp_d[k];
    }
  for(auto &q : Qz)
    mpz_clear(&q);
  return 0;
}
// stage 3: Q' = sum of every rank's partial (exact, order-free), back to mpf,
// diagonal check, restore the local P and Q, Cholesky(UPPER, Q)
int oracle_shard_stage3(oracle_ctx *c, int nparts, const uint64_t *const *qprimes,
                        uint64_t *const *schur_complement_cholesky, uint64_t *const *schur_off_diagonal,
                        uint64_t *Q)
{
  sdpb_host::set_precision(c->prec);
  const int N = c->N, W = oracle_qprime_words(c->prec);
  c->shard.Q.resize(N, N);
  mpz_t acc, term;
  mpz_init(acc);
  mpz_init(term);
  for(int j = 0; j < N; ++j)
    for(int i = 0; i <= j; ++i)
      {
        const size_t e = (size_t)j * N + i;
        mpz_set_ui(acc, 0);
        for(int p = 0; p < nparts; ++p)
          {
            const uint64_t *o = qprimes[p] + e * W;
            if(o[0] == 0)
              continue;
            mpz_import(term, W - 1, -1, 8, 0, 0, o + 1);
            if(o[0] != 1)
              term[0]._mp_size = -term[0]._mp_size;
            mpz_add(acc, acc, term);
          }
        mpf_set_z(c->shard.Q(i, j).v, acc);
      }
  mpz_clear(acc);
  mpz_clear(term);
  restore_and_factor_Q(c->shard, N);
  if(!c->shard.error.empty())
    {
      c->error = c->shard.error;
      return c->error.find("Normalized Q") != std::string::npos ? 4 : 3;
    }
  for(int j = 0; j < c->J; ++j)
    {
      if(schur_complement_cholesky && schur_complement_cholesky[j])
        pack_out(c->shard.schur_complement_cholesky[j], schur_complement_cholesky[j]);
      if(schur_off_diagonal && schur_off_diagonal[j])
        pack_out(c->shard.schur_off_diagonal[j], schur_off_diagonal[j]);
    }
  if(Q)
    pack_out(c->shard.Q, Q);
  return 0;
}

// ---- solve_schur_complement_equation (row N1) -----------------------------
// dx[j]: P_j elements (in: r_x, out: dx), dy: N elements (in: r_y, out: dy);
// uses L_j, L_j^-1 B_j and chol(Q) of the last initialize_schur_complement_solver
// (or oracle_shard_stage3) on this context.
static int unpack_dx(oracle_ctx *c, uint64_t *const *dx, std::vector<Matrix> &x)
{
  if((int)c->shard.schur_complement_cholesky.size() != c->J || c->shard.Q.h != c->N)
    {
      c->error = "solve_schur_complement_equation before initialize_schur_complement_solver";
      return 5;
    }
  x.resize(c->J);
  for(int j = 0; j < c->J; ++j)
    unpack_matrix(x[j], c->shapes[j].schur_size(), 1, dx[j]);
  return 0;
}
int oracle_solve_schur_complement_equation(oracle_ctx *c, uint64_t *const *dx, uint64_t *dy)
{
  sdpb_host::set_precision(c->prec);
  std::vector<Matrix> x;
  if(int rc = unpack_dx(c, dx, x))
    return rc;
  Matrix y;
  unpack_matrix(y, c->N, 1, dy);
  solve_schur_complement_equation(c->shard, x, y);
  for(int j = 0; j < c->J; ++j)
    pack_out(x[j], dx[j]);
  pack_out(y, dy);
  return 0;
}
// sharded model: stage A on the local blocks (dx updated in place, one row of N
// partials per local block), the caller gathers the rows in global block order,
// then stage B+C with the gathered rows.
int oracle_shard_solve_stage1(oracle_ctx *c, uint64_t *const *dx, uint64_t *part)
{
  sdpb_host::set_precision(c->prec);
  std::vector<Matrix> x;
  if(int rc = unpack_dx(c, dx, x))
    return rc;
  std::vector<std::vector<BigFloat>> p;
  schur_solve_forward(c->shard, x, c->N, p);
  const int ew = elem_words();
  for(int j = 0; j < c->J; ++j)
    {
      pack_out(x[j], dx[j]);
      for(int col = 0; col < c->N; ++col)
        sdpb_host::pack(p[j][col], part + ((size_t)j * c->N + col) * ew);
    }
  return 0;
}
int oracle_shard_solve_stage2(oracle_ctx *c, const uint64_t *part_global, int J_global,
                              uint64_t *const *dx, uint64_t *dy)
{
  sdpb_host::set_precision(c->prec);
  std::vector<Matrix> x;
  if(int rc = unpack_dx(c, dx, x))
    return rc;
  const int ew = elem_words(), N = c->N;
  std::vector<std::vector<BigFloat>> p(J_global, std::vector<BigFloat>(N));
  for(int j = 0; j < J_global; ++j)
    for(int col = 0; col < N; ++col)
      sdpb_host::unpack(p[j][col], part_global + ((size_t)j * N + col) * ew);
  Matrix y;
  unpack_matrix(y, N, 1, dy);
  schur_solve_Q(c->shard.Q, p, y);
  schur_solve_backward(c->shard, y, x);
  for(int j = 0; j < c->J; ++j)
    pack_out(x[j], dx[j]);
  pack_out(y, dy);
  return 0;
}

// ---- scale_multiply_add (row N2) ------------------------------------------
// C_b = alpha A_b B_b + beta C_b for the 2J PSD-shaped blocks (s x s each); alpha in {1,-1},
// beta in {0,1} (the reference's call sites)
int oracle_scale_multiply_add(oracle_ctx *c, int alpha, const uint64_t *const *A, const uint64_t *const *B,
                              int beta, uint64_t *const *C)
{
  sdpb_host::set_precision(c->prec);
  if((alpha != 1 && alpha != -1) || (beta != 0 && beta != 1))
    {
      c->error = "scale_multiply_add: alpha must be 1 or -1, beta 0 or 1";
      return 1;
    }
  const BigFloat al(alpha), be(beta);
#pragma omp parallel for schedule(dynamic)
  for(int b = 0; b < 2 * c->J; ++b)
    {
      const int s = c->shapes[b / 2].psd_size(b % 2);
      if(s == 0)
        continue;
      Matrix Am, Bm, Cm;
      unpack_matrix(Am, s, s, A[b]);
      unpack_matrix(Bm, s, s, B[b]);
      if(beta)
        unpack_matrix(Cm, s, s, C[b]);
      else
        Cm.resize(s, s);
      scale_multiply_add_block(al, Am, Bm, be, Cm);
      pack_out(Cm, C[b]);
    }
  return 0;
}

// ---- the search direction (row N2): same call surface as sdpb_b200_direction_* ----
// compute_search_direction.cxx:44-90 and the per-block reductions of step.cxx:137-160, by the
// canonical host restatement of csrc/host/direction.hpp on this context's X, Y and factors.
static void oracle_sma(oracle_ctx *c, int alpha, const std::vector<Matrix> &A, const std::vector<Matrix> &B, int beta,
                       std::vector<Matrix> &C)
{
  const BigFloat al(alpha), be(beta);
  C.resize(A.size());
#pragma omp parallel for schedule(dynamic)
  for(size_t b = 0; b < A.size(); ++b)
    {
      if(A[b].h == 0)
        continue;
      if(!beta)
        C[b].resize(A[b].h, A[b].w);
      scale_multiply_add_block(al, A[b], B[b], be, C[b]);
    }
}
static void pack_scalars(const std::vector<BigFloat> &v, uint64_t *out)
{
  const size_t ew = (size_t)elem_words();
  for(size_t i = 0; i < v.size(); ++i)
    sdpb_host::pack(v[i], out + i * ew);
}
int oracle_direction_begin(oracle_ctx *c, uint64_t *block_traces)
{
  sdpb_host::set_precision(c->prec);
  if((int)c->shard.schur_complement_cholesky.size() != c->J || c->shard.Q.h != c->N)
    {
      c->error = "direction_begin called out of order (needs a successful Schur-complement step)";
      return 5;
    }
  oracle_sma(c, -1, c->X, c->Y, 0, c->minus_XY);
  std::vector<BigFloat> traces;
  sdpb_host::block_traces(c->minus_XY, traces);
  if(block_traces)
    pack_scalars(traces, block_traces);
  c->have_minus_XY = true;
  c->have_direction = false;
  return 0;
}
int oracle_direction_R_errors(oracle_ctx *c, const uint64_t *mu, uint64_t *block_maxima)
{
  sdpb_host::set_precision(c->prec);
  if(!c->have_minus_XY)
    {
      c->error = "direction_R_errors called out of order";
      return 5;
    }
  BigFloat m;
  sdpb_host::unpack(m, mu);
  std::vector<BigFloat> maxima;
  sdpb_host::block_R_errors(c->minus_XY, m, maxima);
  pack_scalars(maxima, block_maxima);
  return 0;
}
int oracle_direction_set_residues(oracle_ctx *c, const uint64_t *const *primal_residues,
                                  const uint64_t *const *dual_residues, const uint64_t *primal_residue_p)
{
  sdpb_host::set_precision(c->prec);
  c->primal_residues.resize(2 * c->J);
  c->dual_residues.resize(c->J);
  for(int b = 0; b < 2 * c->J; ++b)
    {
      const int s = c->shapes[b / 2].psd_size(b % 2);
      if(s)
        unpack_matrix(c->primal_residues[b], s, s, primal_residues[b]);
      else
        c->primal_residues[b].resize(0, 0);
    }
  for(int j = 0; j < c->J; ++j)
    unpack_matrix(c->dual_residues[j], c->shapes[j].schur_size(), 1, dual_residues[j]);
  unpack_matrix(c->primal_residue_p, c->N, 1, primal_residue_p);
  c->have_residues = true;
  return 0;
}
int oracle_compute_search_direction(oracle_ctx *c, const uint64_t *beta_mu, int is_corrector)
{
  sdpb_host::set_precision(c->prec);
  if(!c->have_minus_XY || !c->have_residues || (is_corrector && !c->have_direction))
    {
      c->error = "compute_search_direction called out of order";
      return 5;
    }
  BigFloat bm;
  sdpb_host::unpack(bm, beta_mu);
  if(!is_corrector)
    {
      c->dX = c->X;
      c->dY = c->Y;
      c->dx.resize(c->J);
    }
  sdpb_host::compute_search_direction(
    c->block_info, c->bases, c->X, c->Y, c->X_cholesky, c->minus_XY, c->primal_residues, c->dual_residues,
    c->primal_residue_p, bm, is_corrector != 0,
    [&](int alpha, const std::vector<Matrix> &A, const std::vector<Matrix> &B, int beta, std::vector<Matrix> &C) {
      oracle_sma(c, alpha, A, B, beta, C);
    },
    [&](std::vector<Matrix> &x, Matrix &y) { solve_schur_complement_equation(c->shard, x, y); }, c->dx, c->dX,
    c->dy, c->dY);
  c->have_direction = true;
  return 0;
}
int oracle_direction_frobenius(oracle_ctx *c, uint64_t *block_products)
{
  sdpb_host::set_precision(c->prec);
  if(!c->have_direction)
    {
      c->error = "direction_frobenius called out of order";
      return 5;
    }
  std::vector<BigFloat> products;
  sdpb_host::block_frobenius_products(c->X, c->dX, c->Y, c->dY, products);
  pack_scalars(products, block_products);
  return 0;
}
int oracle_direction_get(oracle_ctx *c, uint64_t *const *dx, uint64_t *const *dX, uint64_t *dy, uint64_t *const *dY)
{
  sdpb_host::set_precision(c->prec);
  if(!c->have_direction)
    {
      c->error = "direction_get called out of order";
      return 5;
    }
  for(int j = 0; j < c->J; ++j)
    if(dx && dx[j])
      pack_out(c->dx[j], dx[j]);
  for(int b = 0; b < 2 * c->J; ++b)
    {
      if(dX && dX[b])
        pack_out(c->dX[b], dX[b]);
      if(dY && dY[b])
        pack_out(c->dY[b], dY[b]);
    }
  if(dy)
    pack_out(c->dy, dy);
  return 0;
}

int oracle_direction_put(oracle_ctx *c, const uint64_t *const *dX, const uint64_t *const *dY)
{
  sdpb_host::set_precision(c->prec);
  if((int)c->shard.schur_complement_cholesky.size() != c->J || c->shard.Q.h != c->N)
    {
      c->error = "direction_put called out of order (needs a successful Schur-complement step)";
      return 5;
    }
  c->dX.resize(2 * c->J);
  c->dY.resize(2 * c->J);
  for(int b = 0; b < 2 * c->J; ++b)
    {
      const int s = c->shapes[b / 2].psd_size(b % 2);
      if(dX)
        unpack_matrix(c->dX[b], s, s, dX[b]);
      if(dY)
        unpack_matrix(c->dY[b], s, s, dY[b]);
    }
  if(dX && dY)
    c->have_direction = true;
  return 0;
}
// ---- step_length (row N3): same call surface as sdpb_b200_step_length ----
// step_length.cxx:27-46 up to the reduction over the blocks: block_min_eigenvalues[b] = smallest
// eigenvalue of L_b^-1 dM_b L_b^-T (csrc/host/step_length.hpp), M = X (which 0) or Y (which 1) of
// the last step and dM = dX / dY of the last compute_search_direction; empty blocks give 0.
int oracle_step_length(oracle_ctx *c, int which, uint64_t *block_min_eigenvalues)
{
  sdpb_host::set_precision(c->prec);
  if(!c->have_direction)
    {
      c->error = "step_length called out of order (needs compute_search_direction)";
      return 5;
    }
  const std::vector<Matrix> &L = which == 0 ? c->X_cholesky : c->Y_cholesky;
  const std::vector<Matrix> &dM = which == 0 ? c->dX : c->dY;
  std::vector<BigFloat> mins(2 * (size_t)c->J);
#pragma omp parallel for schedule(dynamic)
  for(int b = 0; b < 2 * c->J; ++b)
    if(dM[b].h)
      mins[b] = sdpb_host::block_min_eigenvalue(L[b], dM[b]);
  pack_scalars(mins, block_min_eigenvalues);
  return 0;
}
// unit test hooks: the smallest eigenvalue of L^-1 A L^-T for one s x s pair (L == NULL: of A
// itself), and the number of Laguerre steps it took
int oracle_min_eigenvalue(int prec, int s, const uint64_t *L, const uint64_t *A, uint64_t *out, int *iterations)
{
  sdpb_host::set_precision(prec);
  Matrix Am, Lm;
  unpack_matrix(Am, s, s, A);
  if(L)
    {
      unpack_matrix(Lm, s, s, L);
      sdpb_host::lower_triangular_inverse_congruence(Lm, Am);
    }
  std::vector<BigFloat> d, e;
  sdpb_host::tridiagonalize(Am, d, e);
  const BigFloat m = sdpb_host::tridiagonal_min_eigenvalue(d, e, iterations);
  sdpb_host::pack(m, out);
  return 0;
}

// ---- single-kernel entry points for unit parity tests -----------------
int oracle_potrf(int prec, int s, int upper, const uint64_t *A, uint64_t *L)
{
  sdpb_host::set_precision(prec);
  Matrix M;
  unpack_matrix(M, s, s, A);
  const int bad = upper ? cholesky_upper(M) : cholesky_lower(M);
  pack_out(M, L);
  return bad;
}
int oracle_trsm(int prec, int p, int ncols, const uint64_t *L,
                const uint64_t *B, uint64_t *X)
{
  sdpb_host::set_precision(prec);
  Matrix Lm, Bm;
  unpack_matrix(Lm, p, p, L);
  unpack_matrix(Bm, p, ncols, B);
  trsm_lower(Lm, Bm);
  pack_out(Bm, X);
  return 0;
}
// scalar ops on packed elements, for device-vs-libgmp scalar parity
// op: 0 mul 1 add 2 sub 3 div 4 sqrt(a) 5 a<<k 6 a>>k 7 a/4
int oracle_scalar_op(int prec, int op, int k, long count, const uint64_t *a,
                     const uint64_t *b, uint64_t *r)
{
  sdpb_host::set_precision(prec);
  const int ew = elem_words();
  BigFloat x, y, z, four(4);
  for(long i = 0; i < count; ++i)
    {
      sdpb_host::unpack(x, a + i * ew);
      sdpb_host::unpack(y, b + i * ew);
      switch(op)
        {
        case 0: mpf_mul(z.v, x.v, y.v); break;
        case 1: mpf_add(z.v, x.v, y.v); break;
        case 2: mpf_sub(z.v, x.v, y.v); break;
        case 3:
          if(y.sgn() == 0)
            z.zero();
          else
            mpf_div(z.v, x.v, y.v);
          break;
        case 4:
          if(x.sgn() < 0)
            z.zero();
          else
            mpf_sqrt(z.v, x.v);
          break;
        case 5: mpf_mul_2exp(z.v, x.v, (unsigned)k); break;
        case 6: mpf_div_2exp(z.v, x.v, (unsigned)k); break;
        case 7: mpf_div(z.v, x.v, four.v); break;
        default: return 1;
        }
      sdpb_host::pack(z, r + i * ew);
    }
  return 0;
}
// decimal string -> packed element (mpf_set_str at `prec`), for fixtures
int oracle_from_decimal(int prec, const char *dec, uint64_t *out)
{
  sdpb_host::set_precision(prec);
  try
    {
      BigFloat x{std::string(dec)};
      sdpb_host::pack(x, out);
    }
  catch(...)
    {
      return 1;
    }
  return 0;
}
double oracle_to_double(int prec, const uint64_t *in)
{
  sdpb_host::set_precision(prec);
  BigFloat x;
  sdpb_host::unpack(x, in);
  return x.to_double();
}
int oracle_stage_ms(const oracle_ctx *c, double *ms, int n)
{
  for(int i = 0; i < n && i < 9; ++i)
    ms[i] = c->stage_ms[i];
  return 0;
}
int oracle_num_threads()
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
// torchrun exports OMP_NUM_THREADS=1 to its workers: the CPU arm of bench.py sets the thread
// count explicitly so that the baseline always uses every host core it is allowed to run on
void oracle_set_num_threads(int n)
{
#ifdef _OPENMP
  if(n > 0)
    omp_set_num_threads(n);
#else
  (void)n;
#endif
}
} // extern "C"


// ---- the reference's own formulation of the exact syrk (row a9), on the CPU ----------
// bigint_syrk/Readme.md:27-55: residues of trunc(P') modulo primes chosen as in
// fmpz/Fmpz_Comb.cxx:22-68 (adapted there from FLINT's fmpz_mat/mul_blas.c), stored as symmetric
// doubles (fmpz_mul_blas_util.hxx:16-25), one fp64 dsyrk per prime (the caller supplies the BLAS:
// tests/oracle_lib.py uses scipy's OpenBLAS), CRT back to the integer.  Used (a) to check that the
// direct mpz sum above and the CRT route give the same Q' bit for bit, as the reference's own
// calculate_matrix_square.test.cxx:44-86 does against fmpz_mat_mul_blas, and (b) as the second CPU
// syrk variant SURVEY §8d asks bench.py to time.
extern "C" {
unsigned long __gmpz_fdiv_ui(mpz_srcptr, unsigned long);
void __gmpz_mul_ui(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_addmul_ui(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_divexact_ui(mpz_ptr, mpz_srcptr, unsigned long);
void __gmpz_mod(mpz_ptr, mpz_srcptr, mpz_srcptr);
void __gmpz_sub(mpz_ptr, mpz_srcptr, mpz_srcptr);
int __gmpz_cmp(mpz_srcptr, mpz_srcptr);
void __gmpz_mul_2exp(mpz_ptr, mpz_srcptr, unsigned long);
}
struct syrk_crt_state
{
  int prec, N;
  long K;
  std::vector<__mpz_struct> z; // trunc(P'), column-major K x N
};
static bool small_prime(uint64_t n)
{
  if(n < 2 || n % 2 == 0)
    return n == 2;
  for(uint64_t d = 3; d * d <= n; d += 2)
    if(n % d == 0)
      return false;
  return true;
}
static uint64_t inv_mod(uint64_t a, uint64_t p) // p prime
{
  int64_t t = 0, nt = 1, r = (int64_t)p, nr = (int64_t)(a % p);
  while(nr)
    {
      const int64_t q = r / nr;
      int64_t x = t - q * nt;
      t = nt;
      nt = x;
      x = r - q * nr;
      r = nr;
      nr = x;
    }
  return (uint64_t)(t < 0 ? t + (int64_t)p : t);
}
extern "C" {
// Fmpz_Comb.cxx:22-68 with bits = 2 prec + bits(k) + 1 (calculate_output_bits :14-18)
int oracle_syrk_crt_primes(int prec_bits, long k, uint64_t *primes, int max)
{
  int kbits = 0;
  for(long t = k; t; t >>= 1)
    ++kbits;
  const long bits = 2L * prec_bits + kbits + 1;
  uint64_t root = (uint64_t)std::sqrt((double)(((uint64_t)1 << 53) - 1) / (double)k);
  while(root * root > (((uint64_t)1 << 53) - 1) / (uint64_t)k)
    --root;
  uint64_t p = 2 + 2 * root;
  if(bits > 200 && p > 1664544)
    p = 1664544;
  mpz_t prod;
  mpz_init(prod);
  mpz_set_ui(prod, 1);
  int n = 0;
  do
    {
      do
        {
          if(p < 1000)
            {
              mpz_clear(prod);
              return 0;
            }
          --p;
        }
      while(!small_prime(p));
      if(n >= max)
        {
          mpz_clear(prod);
          return -1;
        }
      primes[n++] = p;
      __gmpz_mul_ui(prod, prod, p);
    }
  while((long)mpz_sizeinbase(prod, 2) <= bits);
  mpz_clear(prod);
  return n;
}
// BigFloat -> integer once (fmpz_BigFloat_convert.hxx:13: truncation toward zero)
syrk_crt_state *oracle_syrk_crt_begin(int prec, long K, int N, const uint64_t *Pn)
{
  sdpb_host::set_precision(prec);
  auto *st = new syrk_crt_state{prec, N, K, {}};
  st->z.resize((size_t)K * N);
  const int ew = elem_words();
#pragma omp parallel
  {
    BigFloat x;
#pragma omp for schedule(static)
    for(long e = 0; e < K * (long)N; ++e)
      {
        sdpb_host::unpack(x, Pn + (size_t)e * ew);
        mpz_init(&st->z[e]);
        mpz_set_f(&st->z[e], x.v);
      }
  }
  return st;
}
// residues modulo p as symmetric doubles in (-p/2, p/2], column-major K x N
int oracle_syrk_crt_residues(const syrk_crt_state *st, uint64_t p, double *out)
{
#pragma omp parallel for schedule(static)
  for(long e = 0; e < st->K * (long)st->N; ++e)
    {
      const uint64_t r = __gmpz_fdiv_ui(&st->z[e], p);
      out[e] = r > p / 2 ? (double)((int64_t)r - (int64_t)p) : (double)r;
    }
  return 0;
}
void oracle_syrk_crt_end(syrk_crt_state *st)
{
  for(auto &q : st->z)
    mpz_clear(&q);
  delete st;
}
// res[k][i + j*N] = (dsyrk result for prime k)(i,j) mod p_k in [0, p_k), upper triangle used;
// Q'(i,j) = symmetric lift of the CRT value, converted like fmpz_get_mpf (:9)
int oracle_syrk_crt_reconstruct(int prec, int N, int np, const uint64_t *primes, const int64_t *res,
                                uint64_t *Qout)
{
  sdpb_host::set_precision(prec);
  mpz_t M, half;
  mpz_init(M);
  mpz_init(half);
  mpz_set_ui(M, 1);
  for(int k = 0; k < np; ++k)
    __gmpz_mul_ui(M, M, primes[k]);
  std::vector<__mpz_struct> coef(np);
  for(int k = 0; k < np; ++k)
    {
      mpz_init(&coef[k]);
      __gmpz_divexact_ui(&coef[k], M, primes[k]);
      const uint64_t inv = inv_mod(__gmpz_fdiv_ui(&coef[k], primes[k]), primes[k]);
      __gmpz_mul_ui(&coef[k], &coef[k], inv);
    }
  Matrix Q(N, N);
#pragma omp parallel
  {
    mpz_t x, twice;
    mpz_init(x);
    mpz_init(twice);
#pragma omp for schedule(dynamic, 1)
    for(int j = 0; j < N; ++j)
      for(int i = 0; i <= j; ++i)
        {
          mpz_set_ui(x, 0);
          for(int k = 0; k < np; ++k)
            __gmpz_addmul_ui(x, &coef[k], (unsigned long)res[(size_t)k * N * N + (size_t)j * N + i]);
          __gmpz_mod(x, x, M);
          __gmpz_mul_2exp(twice, x, 1);
          if(__gmpz_cmp(twice, M) > 0)
            __gmpz_sub(x, x, M);
          mpf_set_z(Q(i, j).v, x);
        }
    mpz_clear(x);
    mpz_clear(twice);
  }
  pack_out(Q, Qout);
  for(auto &q : coef)
    mpz_clear(&q);
  mpz_clear(M);
  mpz_clear(half);
  return 0;
}
// the direct mpz sum on the same input, for comparison and timing
int oracle_syrk_direct(int prec, long K, int N, const uint64_t *Pn, uint64_t *Qout)
{
  sdpb_host::set_precision(prec);
  std::vector<Matrix> blocks(1);
  unpack_matrix(blocks[0], (int)K, N, Pn);
  Matrix Q;
  exact_syrk_upper(blocks, N, Q);
  pack_out(Q, Qout);
  return 0;
}
}

// ---- deterministic synthetic inputs (seeded) ---------------
static inline uint64_t splitmix64(uint64_t &s)
{
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
extern "C" {
// h x w matrix, entries uniform in (-1,1) with full-length mantissas
int oracle_random_matrix(int prec, int h, int w, uint64_t seed, uint64_t *out)
{
  const int nl = (prec + 63) / 64 + 2, ew = (nl + 2) & ~1;
  uint64_t s = seed * 0x2545F4914F6CDD1Dull + 0x5D9B0000ull;
  for(long e = 0; e < (long)h * w; ++e)
    {
      uint64_t *p = out + e * ew;
      for(int i = 0; i < ew; ++i)
        p[i] = 0;
      for(int i = 0; i < nl; ++i)
        p[1 + i] = splitmix64(s);
      if(p[nl] == 0)
        p[nl] = 1;
      const int32_t sign = (splitmix64(s) & 1) ? 1 : -1;
      p[0] = (uint64_t)(uint32_t)0 | ((uint64_t)(uint32_t)sign << 32);
    }
  return 0;
}
// symmetric positive definite s x s: G G^T / s + I, G random as above
int oracle_random_spd(int prec, int s, uint64_t seed, uint64_t *out)
{
  sdpb_host::set_precision(prec);
  std::vector<uint64_t> buf((size_t)s * s * elem_words());
  oracle_random_matrix(prec, s, s, seed, buf.data());
  Matrix G, A(s, s);
  unpack_matrix(G, s, s, buf.data());
  BigFloat prod, inv_s = BigFloat(1) / BigFloat(s > 0 ? s : 1);
  for(int j = 0; j < s; ++j)
    for(int i = j; i < s; ++i)
      {
        BigFloat acc;
        for(int l = 0; l < s; ++l)
          {
            prod = G(i, l);
            prod *= G(j, l);
            acc += prod;
          }
        acc *= inv_s;
        if(i == j)
          acc += BigFloat(1);
        A(i, j) = acc;
        A(j, i) = acc;
      }
  pack_out(A, out);
  return 0;
}
// out = scale * in  (scale as a double), for building ill-scaled test inputs
int oracle_scale_matrix(int prec, long count, double scale, const uint64_t *in,
                        uint64_t *out)
{
  sdpb_host::set_precision(prec);
  const int ew = elem_words();
  BigFloat x, sc(scale);
  for(long e = 0; e < count; ++e)
    {
      sdpb_host::unpack(x, in + e * ew);
      x *= sc;
      sdpb_host::pack(x, out + e * ew);
    }
  return 0;
}
}

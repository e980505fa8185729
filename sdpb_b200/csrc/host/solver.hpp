// Host side of the solver: SDP_Solver::run / SDP_Solver::step with the
// reference's structure (reference src/sdp_solve/SDP_Solver/run/run.cxx:184-470,
// run/step/step.cxx:51-229), single process, GMP mpf scalars.  The hot path —
// cholesky_decomposition, compute_bilinear_pairings,
// initialize_schur_complement_solver — is NOT implemented here: it is reached
// through the Hot_Path interface below, whose product implementation
// (hot_path_b200.cpp) calls the sm_100a kernels through the C-ABI of
// include/sdpb_b200.h.  Everything else in an iteration (objectives, residues,
// search direction, step length, termination, iterations.json / out.txt) is
// cheap O(sum s_p^3 + P N) host work and stays on the CPU as in the reference.
#pragma once
#include "direction.hpp"
#include "step_length.hpp"
#include "sdp.hpp"
#include "checkpoint.hpp"

#include <chrono>
#include <cmath>
#include <csignal>
#include <cstdio>
#include <functional>
#include <sys/stat.h>

namespace sdpb_host
{
// The three free functions SDP_Solver::run/step call (run.cxx:14-17,37-45;
// step.cxx:12-24) plus the upload of the immutable SDP data.  Errors are
// std::runtime_error with the reference's texts.
struct Hot_Path
{
  virtual ~Hot_Path() {}
  // cholesky_decomposition.cxx:5-28; which = 0: X, 1: Y.  L[b] = lower factor.
  virtual void cholesky_decomposition(int which, const std::vector<Matrix> &A, std::vector<Matrix> &L) = 0;
  // compute_bilinear_pairings.cxx:17-31.  A_X_inv stays with the implementation
  // (only the Schur assembly reads it); A_Y[b] = V^T Y V, (m n) x (m n).
  virtual void compute_bilinear_pairings(const std::vector<Matrix> &Y, std::vector<Matrix> &A_Y) = 0;
  // initialize_schur_complement_solver.cxx:62-104
  // schur_off_diagonal may come back empty: it is read only by the Schur solve below, which
  // the implementation runs on its own resident copy.
  virtual void initialize_schur_complement_solver(std::vector<Matrix> &schur_complement_cholesky,
                                                  std::vector<Matrix> &schur_off_diagonal, Matrix &Q)
    = 0;
  // solve_schur_complement_equation.cxx:16-79 with the L_j, L_j^-1 B_j and
  // chol(Q) of the preceding initialize_schur_complement_solver, which stay with the
  // implementation: dx[j] (P_j x 1) and dy (N x 1) hold r_x, r_y on entry, the solution on exit.
  virtual void solve_schur_complement_equation(std::vector<Matrix> &dx, Matrix &dy) = 0;
  // scale_multiply_add.cxx:4-16: C_b = alpha A_b B_b + beta C_b on the PSD-shaped
  // blocks; alpha is 1 or -1, beta 0 or 1 at every call site (step.cxx:137,
  // compute_search_direction.cxx:28,60).
  virtual void scale_multiply_add(int alpha, const std::vector<Matrix> &A, const std::vector<Matrix> &B, int beta,
                                  std::vector<Matrix> &C)
    = 0;
  // ---- rows N2: the search direction on the implementation's own (device-resident) copies of
  // X, Y, their factors and the step's temporaries (include/sdpb_b200.h, sdpb_b200_direction_*).
  // An implementation without them leaves resident_direction() false and the solver runs the host
  // restatement of direction.hpp -- the reference's own CPU code path.
  virtual bool resident_direction() const { return false; }
  virtual void direction_begin(std::vector<BigFloat> &) { throw std::logic_error("no resident direction"); }
  virtual void direction_R_errors(const BigFloat &, std::vector<BigFloat> &) { throw std::logic_error("no resident direction"); }
  virtual void direction_set_residues(const std::vector<Matrix> &, const std::vector<Matrix> &, const Matrix &)
  {
    throw std::logic_error("no resident direction");
  }
  virtual void compute_search_direction(const BigFloat &, bool) { throw std::logic_error("no resident direction"); }
  virtual void direction_frobenius(std::vector<BigFloat> &) { throw std::logic_error("no resident direction"); }
  virtual void direction_get(std::vector<Matrix> &, std::vector<Matrix> &, Matrix &, std::vector<Matrix> &)
  {
    throw std::logic_error("no resident direction");
  }
  // row N3 (step_length.cxx:27-46): mins[b] = smallest eigenvalue of L_b^-1 dM_b L_b^-T for every
  // non-empty block-parity b, with M = X, dM = dX (which = 0) or Y, dY (which = 1) of the resident
  // direction; false: not available, the caller runs step_length.hpp on the host
  virtual bool step_length_min_eigenvalues(int, std::vector<BigFloat> &) { return false; }
  virtual std::string name() const = 0;
};

// Solver_Parameters (reference src/sdp_solve/Solver_Parameters/Solver_Parameters.cxx:20-157);
// defaults are parsed from the same decimal strings at working precision (:13-18).
struct Solver_Parameters
{
  int precision = 400;
  long max_iterations = 500;
  long max_runtime = 1L << 60;
  bool find_primal_feasible = false, find_dual_feasible = false;
  bool detect_primal_feasible_jump = false, detect_dual_feasible_jump = false;
  std::string duality_gap_threshold = "1e-30", primal_error_threshold = "1e-30",
              dual_error_threshold = "1e-30", initial_matrix_scale_primal = "1e20",
              initial_matrix_scale_dual = "1e20", feasible_centering_parameter = "0.1",
              infeasible_centering_parameter = "0.3", step_length_reduction = "0.7",
              max_complementarity = "1e100", min_primal_step = "0", min_dual_step = "0";
  std::string write_solution = "x,y"; // Write_Solution.cxx
  // checkpoints (Solver_Parameters.cxx:140-156, sdpb/SDPB_Parameters.cxx:41-47,176-193)
  std::string checkpoint_out, checkpoint_in;
  bool checkpoint_out_set = false, checkpoint_in_set = false, no_final_checkpoint = false;
  long checkpoint_interval = 3600;
  bool set(const std::string &key, const std::string &value)
  {
    auto flag = [&](bool &b) { b = value.empty() || value == "1" || value == "true"; };
    if(key == "precision") precision = std::stoi(value);
    else if(key == "maxIterations") max_iterations = std::stol(value);
    else if(key == "maxRuntime") max_runtime = std::stol(value);
    else if(key == "findPrimalFeasible") flag(find_primal_feasible);
    else if(key == "findDualFeasible") flag(find_dual_feasible);
    else if(key == "detectPrimalFeasibleJump") flag(detect_primal_feasible_jump);
    else if(key == "detectDualFeasibleJump") flag(detect_dual_feasible_jump);
    else if(key == "dualityGapThreshold") duality_gap_threshold = value;
    else if(key == "primalErrorThreshold") primal_error_threshold = value;
    else if(key == "dualErrorThreshold") dual_error_threshold = value;
    else if(key == "initialMatrixScalePrimal") initial_matrix_scale_primal = value;
    else if(key == "initialMatrixScaleDual") initial_matrix_scale_dual = value;
    else if(key == "feasibleCenteringParameter") feasible_centering_parameter = value;
    else if(key == "infeasibleCenteringParameter") infeasible_centering_parameter = value;
    else if(key == "stepLengthReduction") step_length_reduction = value;
    else if(key == "maxComplementarity") max_complementarity = value;
    else if(key == "minPrimalStep") min_primal_step = value;
    else if(key == "minDualStep") min_dual_step = value;
    else if(key == "writeSolution") write_solution = value;
    else if(key == "checkpointDir" || key == "c")
      {
        checkpoint_out = value;
        checkpoint_out_set = true;
      }
    else if(key == "initialCheckpointDir" || key == "i")
      {
        checkpoint_in = value;
        checkpoint_in_set = true;
      }
    else if(key == "checkpointInterval") checkpoint_interval = std::stol(value);
    else if(key == "noFinalCheckpoint") flag(no_final_checkpoint);
    else if(key == "verbosity" || key == "procGranularity" || key == "maxSharedMemory")
      ; // accepted for command-line compatibility; no effect in this host
    else
      return false;
    return true;
  }
};

enum class Terminate_Reason
{
  PrimalDualOptimal,
  PrimalFeasible,
  DualFeasible,
  PrimalFeasibleJumpDetected,
  DualFeasibleJumpDetected,
  MaxComplementarityExceeded,
  MaxIterationsExceeded,
  MaxRuntimeExceeded,
  PrimalStepTooSmall,
  DualStepTooSmall,
  SIGTERM_Received
};
// SDP_Solver_Terminate_Reason.cxx:5-45
inline const char *to_string(Terminate_Reason r)
{
  switch(r)
    {
    case Terminate_Reason::PrimalDualOptimal: return "found primal-dual optimal solution";
    case Terminate_Reason::PrimalFeasible: return "found primal feasible solution";
    case Terminate_Reason::DualFeasible: return "found dual feasible solution";
    case Terminate_Reason::PrimalFeasibleJumpDetected: return "primal feasible jump detected";
    case Terminate_Reason::DualFeasibleJumpDetected: return "dual feasible jump detected";
    case Terminate_Reason::MaxIterationsExceeded: return "maxIterations exceeded";
    case Terminate_Reason::MaxRuntimeExceeded: return "maxRuntime exceeded";
    case Terminate_Reason::MaxComplementarityExceeded: return "maxComplementarity exceeded";
    case Terminate_Reason::PrimalStepTooSmall: return "primal step too small";
    case Terminate_Reason::DualStepTooSmall: return "dual step too small";
    case Terminate_Reason::SIGTERM_Received: return "SIGTERM signal received";
    }
  return "?";
}

// operator<< of a BigFloat in the reference's streams: default float format
// with ceil(prec log10 2) + 1 significant digits (sdpb_util/ostream/set_stream_precision.hxx:7-11),
// i.e. printf's %g: exponent form when exp10 < -4 or >= digits, trailing zeros dropped.
inline std::string format_bigfloat(const BigFloat &x)
{
  const int digits = (int)std::ceil(working_precision_bits() * std::log10(2.0)) + 1;
  if(x.sgn() == 0)
    return "0";
  mp_exp_t e;
  char *s = mpf_get_str(nullptr, &e, 10, (size_t)digits, x.v);
  std::string m(s);
  free(s);
  std::string out;
  if(m[0] == '-')
    {
      out = "-";
      m = m.substr(1);
    }
  const long x10 = (long)e - 1; // value = 0.m * 10^e = m[0].m[1..] * 10^(e-1)
  if(x10 < -4 || x10 >= digits)
    {
      out += m.substr(0, 1);
      if(m.size() > 1)
        out += "." + m.substr(1);
      out += (x10 < 0 ? "e-" : "e+");
      const long a = x10 < 0 ? -x10 : x10;
      if(a < 10)
        out += "0";
      out += std::to_string(a);
    }
  else if(x10 < 0)
    out += "0." + std::string((size_t)(-x10 - 1), '0') + m;
  else
    {
      if((long)m.size() <= x10 + 1)
        out += m + std::string((size_t)(x10 + 1 - (long)m.size()), '0');
      else
        out += m.substr(0, (size_t)x10 + 1) + "." + m.substr((size_t)x10 + 1);
    }
  return out;
}

// ---- small dense helpers (the reference calls El::Gemm / Trsm / Dotu here;
// Elemental's operation order is not observable, results are pinned to 2^-99) ----
// C = alpha A B + beta C
inline void gemm_nn(const BigFloat &alpha, const Matrix &A, const Matrix &B, const BigFloat &beta, Matrix &C)
{
  BigFloat acc, t;
  const bool beta_zero = beta.sgn() == 0;
  for(int j = 0; j < B.w; ++j)
    for(int i = 0; i < A.h; ++i)
      {
        acc.zero();
        for(int l = 0; l < A.w; ++l)
          {
            t = A(i, l);
            t *= B(l, j);
            acc += t;
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
// cholesky::SolveAfter(UPPER) with Q = U^T U: x <- U^{-1} U^{-T} x
inline void cholesky_upper_solve(const Matrix &U, Matrix &x)
{
  BigFloat t;
  const int n = U.h;
  for(int c = 0; c < x.w; ++c)
    {
      for(int i = 0; i < n; ++i)
        {
          for(int k = 0; k < i; ++k)
            {
              t = U(k, i);
              t *= x(k, c);
              x(i, c) -= t;
            }
          x(i, c) /= U(i, i);
        }
      for(int i = n - 1; i >= 0; --i)
        {
          for(int k = i + 1; k < n; ++k)
            {
              t = U(i, k);
              t *= x(k, c);
              x(i, c) -= t;
            }
          x(i, c) /= U(i, i);
        }
    }
}
inline BigFloat max_abs(const Matrix &A)
{
  BigFloat m;
  for(const auto &x : A.a)
    {
      const BigFloat a = Abs(x);
      if(a > m)
        m = a;
    }
  return m;
}
inline BigFloat dotu(const Matrix &A, const Matrix &B)
{
  BigFloat acc, t;
  for(size_t i = 0; i < A.a.size(); ++i)
    {
      t = A.a[i];
      t *= B.a[i];
      acc += t;
    }
  return acc;
}
inline void axpy(const BigFloat &alpha, const Matrix &X, Matrix &Y)
{
  BigFloat t;
  for(size_t i = 0; i < X.a.size(); ++i)
    {
      t = X.a[i];
      t *= alpha;
      Y.a[i] += t;
    }
}
// cholesky_condition_number (sdpb_util/cholesky_condition_number.hxx:8-36)
inline BigFloat cholesky_condition_number(const Matrix &L)
{
  if(L.h == 0)
    return BigFloat(0);
  BigFloat mx = L(0, 0), mn = L(0, 0);
  for(int i = 1; i < L.h; ++i)
    {
      if(L(i, i) > mx)
        mx = L(i, i);
      if(L(i, i) < mn)
        mn = L(i, i);
    }
  const BigFloat ratio = mx / mn;
  return ratio * ratio;
}

struct Iteration_Record
{
  long iteration;
  double total_time, iter_time;
  BigFloat mu, primal_objective, dual_objective, duality_gap, primal_error_P, primal_error_p,
    dual_error, R_error, primal_step_length, dual_step_length, beta_corrector, Q_cond_number,
    max_block_cond_number;
  std::string block_name;
};

class SDP_Solver
{
public:
  const Block_Info &block_info;
  const SDP &sdp;
  Hot_Path &hot;
  // SDP_Solver.hxx:27-60
  std::vector<Matrix> x;        // J, P_j x 1
  std::vector<Matrix> X, Y;     // 2J
  Matrix y;                     // N x 1 (the reference keeps one copy per block)
  std::vector<Matrix> primal_residues; // 2J
  std::vector<Matrix> dual_residues;   // J
  BigFloat primal_objective, dual_objective, duality_gap, primal_error_P, primal_error_p, dual_error,
    R_error;
  std::vector<Iteration_Record> iterations;
  // save_checkpoint.cxx / load_checkpoint.cxx (checkpoint.hpp): generations of the binary checkpoint
  Checkpoint_State checkpoint;
  bool loaded_checkpoint = false;
  std::string options_json = "{}\n"; // the "options" object of checkpoint.json
  double hot_path_seconds = 0, host_seconds = 0, direction_seconds = 0;
  std::function<void(const Iteration_Record &)> on_iteration;

  BigFloat primal_error() const { return Max(primal_error_P, primal_error_p); }

  // SDP_Solver::SDP_Solver (SDP_Solver/SDP_Solver.cxx:3-38): X = Omega_p I, Y = Omega_d I, x = y = 0
  SDP_Solver(const Solver_Parameters &parameters, const Block_Info &bi, const SDP &s, Hot_Path &h)
      : block_info(bi), sdp(s), hot(h)
  {
    const int J = bi.num_blocks();
    x.resize(J);
    dual_residues.resize(J);
    X.resize(2 * J);
    Y.resize(2 * J);
    primal_residues.resize(2 * J);
    const BigFloat op(parameters.initial_matrix_scale_primal), od(parameters.initial_matrix_scale_dual);
    for(int j = 0; j < J; ++j)
      {
        x[j].resize(bi.schur_block_size(j), 1);
        dual_residues[j].resize(bi.schur_block_size(j), 1);
        for(int p = 0; p < 2; ++p)
          {
            const int s_p = bi.psd_matrix_block_size(j, p);
            X[2 * j + p].resize(s_p, s_p);
            Y[2 * j + p].resize(s_p, s_p);
            primal_residues[2 * j + p].resize(s_p, s_p);
            for(int i = 0; i < s_p; ++i)
              {
                X[2 * j + p](i, i) = op;
                Y[2 * j + p](i, i) = od;
              }
          }
      }
    y.resize(s.N(), 1);
    // SDP_Solver.cxx:20-38: the initial point above unless a checkpoint (binary, else text) loads
    checkpoint.x = &x;
    checkpoint.X = &X;
    checkpoint.y = &y;
    checkpoint.Y = &Y;
    loaded_checkpoint = load_checkpoint(parameters.checkpoint_in, checkpoint,
                                        parameters.checkpoint_in_set && !parameters.checkpoint_in.empty());
  }
  void save_checkpoint(const Solver_Parameters &parameters)
  {
    sdpb_host::save_checkpoint(parameters.checkpoint_out, checkpoint, options_json, "sdpb-b200");
  }

  // constraint_matrix_weighted_sum.cxx:14-66 (direction.hpp)
  void constraint_matrix_weighted_sum(const std::vector<Matrix> &a, std::vector<Matrix> &result) const
  {
    sdpb_host::constraint_matrix_weighted_sum(block_info, sdp.bilinear_bases, a, result);
  }

  // compute_objectives.cxx
  void compute_objectives()
  {
    BigFloat s;
    for(size_t j = 0; j < x.size(); ++j)
      s += dotu(sdp.primal_objective_c[j], x[j]);
    primal_objective = sdp.objective_const + s;
    dual_objective = sdp.objective_const + dotu(sdp.dual_objective_b, y);
    duality_gap = Abs(primal_objective - dual_objective)
                  / Max(Abs(primal_objective) + Abs(dual_objective), BigFloat(1));
  }

  // compute_dual_residues_and_error.cxx:9-70.  A_Y[b] is the full (m n) x (m n)
  // matrix; the reference's tile [cb][rb](row,col) = A_Y[b](cb n + col, rb n + row).
  void compute_dual_residues_and_error(const std::vector<Matrix> &A_Y)
  {
    const int J = block_info.num_blocks();
    std::vector<BigFloat> local_max(J);
#pragma omp parallel for schedule(dynamic)
    for(int j = 0; j < J; ++j)
      {
        const int n = block_info.num_points[j], m = block_info.dimensions[j];
        Matrix &res = dual_residues[j];
        res.zero();
        for(int parity = 0; parity < 2; ++parity)
          {
            const Matrix &A = A_Y[2 * j + parity];
            if(A.h == 0)
              continue;
            for(int cb = 0; cb < m; ++cb)
              for(int rb = 0; rb <= cb; ++rb)
                {
                  const int off = (cb * (cb + 1) / 2 + rb) * n;
                  for(int k = 0; k < n; ++k)
                    res(off + k, 0) -= A(cb * n + k, rb * n + k);
                }
          }
        // dualResidues -= B y ; += c
        gemm_nn(BigFloat(-1), sdp.free_var_matrix[j], y, BigFloat(1), res);
        axpy(BigFloat(1), sdp.primal_objective_c[j], res);
        local_max[j] = max_abs(res);
      }
    dual_error.zero();
    for(int j = 0; j < J; ++j)
      dual_error = Max(dual_error, local_max[j]);
  }

  // compute_primal_residues_and_error_P_Ax_X.cxx
  void compute_primal_residues_and_error_P_Ax_X()
  {
    constraint_matrix_weighted_sum(x, primal_residues);
    primal_error_P.zero();
    for(size_t b = 0; b < X.size(); ++b)
      {
        for(size_t i = 0; i < X[b].a.size(); ++i)
          primal_residues[b].a[i] -= X[b].a[i];
        primal_error_P = Max(primal_error_P, max_abs(primal_residues[b]));
      }
  }
  // compute_primal_residues_and_error_p_b_Bx.cxx: p = b - B^T x (summed over blocks)
  void compute_primal_residues_and_error_p_b_Bx(Matrix &primal_residue_p)
  {
    const int N = sdp.N(), J = block_info.num_blocks();
    primal_residue_p.resize(N, 1);
    std::vector<Matrix> part(J);
#pragma omp parallel for schedule(dynamic)
    for(int j = 0; j < J; ++j)
      {
        part[j].resize(N, 1);
        BigFloat acc, t;
        const Matrix &B = sdp.free_var_matrix[j];
        for(int c = 0; c < N; ++c)
          {
            acc.zero();
            for(int r = 0; r < B.h; ++r)
              {
                t = B(r, c);
                t *= x[j](r, 0);
                acc += t;
              }
            part[j](c, 0) = -acc;
          }
      }
    for(int c = 0; c < N; ++c)
      primal_residue_p(c, 0) = sdp.dual_objective_b(c, 0);
    for(int j = 0; j < J; ++j)
      for(int c = 0; c < N; ++c)
        primal_residue_p(c, 0) += part[j](c, 0);
    primal_error_p = max_abs(primal_residue_p);
  }

  // compute_feasible_and_termination.cxx:5-75
  void compute_feasible_and_termination(const Solver_Parameters &parameters, const BigFloat &primal_step_length,
                                        const BigFloat &dual_step_length, long iteration,
                                        const std::chrono::steady_clock::time_point &start,
                                        bool &is_primal_and_dual_feasible, Terminate_Reason &reason,
                                        bool &terminate_now) const
  {
    const bool is_dual_feasible = dual_error < BigFloat(parameters.dual_error_threshold),
               is_primal_feasible = primal_error() < BigFloat(parameters.primal_error_threshold);
    is_primal_and_dual_feasible = is_primal_feasible && is_dual_feasible;
    const bool is_optimal = duality_gap < BigFloat(parameters.duality_gap_threshold);
    terminate_now = true;
    if(is_primal_and_dual_feasible && is_optimal)
      reason = Terminate_Reason::PrimalDualOptimal;
    else if(is_dual_feasible && parameters.find_dual_feasible)
      reason = Terminate_Reason::DualFeasible;
    else if(is_primal_feasible && parameters.find_primal_feasible)
      reason = Terminate_Reason::PrimalFeasible;
    else if(dual_step_length == BigFloat(1) && parameters.detect_dual_feasible_jump)
      reason = Terminate_Reason::DualFeasibleJumpDetected;
    else if(primal_step_length == BigFloat(1) && parameters.detect_primal_feasible_jump)
      reason = Terminate_Reason::PrimalFeasibleJumpDetected;
    else if(iteration > parameters.max_iterations)
      reason = Terminate_Reason::MaxIterationsExceeded;
    else if(std::chrono::duration_cast<std::chrono::seconds>(std::chrono::steady_clock::now() - start).count()
            >= parameters.max_runtime)
      reason = Terminate_Reason::MaxRuntimeExceeded;
    else if(iteration > 1 && primal_step_length < BigFloat(parameters.min_primal_step))
      reason = Terminate_Reason::PrimalStepTooSmall;
    else if(iteration > 1 && dual_step_length < BigFloat(parameters.min_dual_step))
      reason = Terminate_Reason::DualStepTooSmall;
    else
      terminate_now = false;
  }

  // C = alpha A B + beta C per block (scale_multiply_add.cxx): through the hot-path seam
  void scale_multiply_add(int alpha, const std::vector<Matrix> &A, const std::vector<Matrix> &B, int beta,
                          std::vector<Matrix> &C) const
  {
    hot.scale_multiply_add(alpha, A, B, beta, C);
  }

  // step_length.cxx:27-46 (+ lower_triangular_inverse_congruence.cxx, min_eigenvalue.cxx); the
  // per-block eigenvalues come from the hot path when it keeps the direction resident (row N3)
  BigFloat step_length(int which, const std::vector<Matrix> &MCholesky, const std::vector<Matrix> &dM,
                       const BigFloat &gamma)
  {
    std::vector<BigFloat> mins;
    if(!(hot.resident_direction() && hot.step_length_min_eigenvalues(which, mins)))
      {
        mins.assign(dM.size(), BigFloat());
#pragma omp parallel for schedule(dynamic)
        for(size_t b = 0; b < dM.size(); ++b)
          if(dM[b].h)
            mins[b] = block_min_eigenvalue(MCholesky[b], dM[b]);
      }
    // El::Min over the blocks, AllReduce MIN over the ranks (min_eigenvalue.cxx:31-32): exact
    bool first = true;
    BigFloat lambda;
    for(size_t b = 0; b < dM.size(); ++b)
      if(dM[b].h && (first || mins[b] < lambda))
        {
          lambda = mins[b];
          first = false;
        }
    if(first || lambda > -gamma)
      return BigFloat(1);
    return -gamma / lambda;
  }

  // SDP_Solver::step (step/step.cxx:51-229)
  void step(const Solver_Parameters &parameters, size_t total_psd_rows, bool is_primal_and_dual_feasible,
            const std::vector<Matrix> &X_cholesky, const std::vector<Matrix> &Y_cholesky,
            const Matrix &primal_residue_p, BigFloat &mu, BigFloat &beta_corrector,
            BigFloat &primal_step_length, BigFloat &dual_step_length, bool &terminate_now,
            BigFloat &Q_cond_number, BigFloat &max_block_cond_number, std::string &max_block_cond_number_name)
  {
    const int J = block_info.num_blocks();
    std::vector<Matrix> dx(x), dX(X), dY(Y);
    Matrix dy(y);
    {
      std::vector<Matrix> schur_complement_cholesky, schur_off_diagonal;
      Matrix Q;
      const auto t0 = std::chrono::steady_clock::now();
      hot.initialize_schur_complement_solver(schur_complement_cholesky, schur_off_diagonal, Q);
      hot_path_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

      // With a resident direction (rows N2) X, Y, their factors, -XY, R, Z, dX, dY stay with the
      // implementation (HBM); the host sees per-block scalars and, at the end, the direction.
      // Otherwise the same sequence runs on the host (direction.hpp), as in the reference.
      const bool resident = hot.resident_direction();
      const auto td = std::chrono::steady_clock::now();
      std::vector<Matrix> minus_XY;
      std::vector<BigFloat> per_block;
      if(resident)
        hot.direction_begin(per_block);
      else
        {
          minus_XY = X;
          scale_multiply_add(-1, X, Y, 0, minus_XY);
          block_traces(minus_XY, per_block);
        }
      mu = -ordered_sum(per_block) / BigFloat((long)total_psd_rows);
      if(mu > BigFloat(parameters.max_complementarity))
        {
          terminate_now = true;
          return;
        }
      // compute_R_error.hxx
      if(resident)
        hot.direction_R_errors(mu, per_block);
      else
        block_R_errors(minus_XY, mu, per_block);
      R_error.zero();
      for(const auto &v : per_block)
        R_error = Max(R_error, v);
      if(resident)
        hot.direction_set_residues(primal_residues, dual_residues, primal_residue_p);
      auto search_direction = [&](const BigFloat &beta, bool is_corrector_phase) {
        const BigFloat beta_mu = beta * mu;
        if(resident)
          hot.compute_search_direction(beta_mu, is_corrector_phase);
        else
          compute_search_direction(
            block_info, sdp.bilinear_bases, X, Y, X_cholesky, minus_XY, primal_residues, dual_residues,
            primal_residue_p, beta_mu, is_corrector_phase,
            [&](int alpha, const std::vector<Matrix> &A, const std::vector<Matrix> &B, int beta_, std::vector<Matrix> &C) {
              scale_multiply_add(alpha, A, B, beta_, C);
            },
            [&](std::vector<Matrix> &rx, Matrix &ry) { hot.solve_schur_complement_equation(rx, ry); }, dx, dX, dy,
            dY);
      };
      // predictor_centering_parameter.cxx
      const BigFloat beta_predictor
        = is_primal_and_dual_feasible ? BigFloat(0) : BigFloat(parameters.infeasible_centering_parameter);
      search_direction(beta_predictor, false);
      // corrector_centering_parameter.cxx (+ frobenius_product_of_sums.cxx)
      {
        if(resident)
          hot.direction_frobenius(per_block);
        else
          block_frobenius_products(X, dX, Y, dY, per_block);
        const BigFloat fp = ordered_sum(per_block);
        const BigFloat r = fp / (mu * BigFloat((long)total_psd_rows));
        const BigFloat beta = r < BigFloat(1) ? r * r : r;
        if(is_primal_and_dual_feasible)
          beta_corrector = Min(Max(BigFloat(parameters.feasible_centering_parameter), beta), BigFloat(1));
        else
          beta_corrector = Max(BigFloat(parameters.infeasible_centering_parameter), beta);
      }
      search_direction(beta_corrector, true);
      if(resident)
        hot.direction_get(dx, dX, dy, dY);
      direction_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - td).count();
      // update_cond_numbers.hxx
      Q_cond_number = cholesky_condition_number(Q);
      max_block_cond_number.zero();
      max_block_cond_number_name = "";
      for(int j = 0; j < J; ++j)
        {
          const BigFloat cs = cholesky_condition_number(schur_complement_cholesky[j]);
          if(max_block_cond_number < cs)
            {
              max_block_cond_number = cs;
              max_block_cond_number_name = "schur_complement_cholesky.block_" + std::to_string(j);
            }
          for(int parity = 0; parity < 2; ++parity)
            {
              const BigFloat cx = cholesky_condition_number(X_cholesky[2 * j + parity]);
              if(max_block_cond_number < cx)
                {
                  max_block_cond_number = cx;
                  max_block_cond_number_name
                    = "X_cholesky.block_" + std::to_string(j) + "_" + std::to_string(parity);
                }
              const BigFloat cy = cholesky_condition_number(Y_cholesky[2 * j + parity]);
              if(max_block_cond_number < cy)
                {
                  max_block_cond_number = cy;
                  max_block_cond_number_name
                    = "Y_cholesky.block_" + std::to_string(j) + "_" + std::to_string(parity);
                }
            }
        }
    }
    const BigFloat gamma(parameters.step_length_reduction);
    primal_step_length = step_length(0, X_cholesky, dX, gamma);
    dual_step_length = step_length(1, Y_cholesky, dY, gamma);
    if(is_primal_and_dual_feasible)
      {
        primal_step_length = Min(primal_step_length, dual_step_length);
        dual_step_length = primal_step_length;
      }
    for(size_t j = 0; j < x.size(); ++j)
      axpy(primal_step_length, dx[j], x[j]);
    for(size_t b = 0; b < X.size(); ++b)
      axpy(primal_step_length, dX[b], X[b]);
    axpy(dual_step_length, dy, y);
    for(size_t b = 0; b < Y.size(); ++b)
      axpy(dual_step_length, dY[b], Y[b]);
  }

  // SIGTERM is latched by a handler installed for the duration of run() (Environment::sigterm_received
  // in the reference, sdpb_util/Environment.hxx:26) and acted on at the top of the next iteration
  static volatile std::sig_atomic_t &sigterm_flag()
  {
    static volatile std::sig_atomic_t flag = 0;
    return flag;
  }
  static void on_sigterm(int) { sigterm_flag() = 1; }

  // SDP_Solver::run (run/run.cxx:184-470)
  Terminate_Reason run(const Solver_Parameters &parameters)
  {
    Terminate_Reason reason = Terminate_Reason::MaxIterationsExceeded;
    const auto start = std::chrono::steady_clock::now();
    auto last_checkpoint_time = start;
    BigFloat primal_step_length(0), dual_step_length(0);
    std::vector<Matrix> X_cholesky, Y_cholesky, A_Y;
    const size_t total_psd_rows = block_info.total_psd_rows();
    sigterm_flag() = 0;
    struct Handler_Guard
    {
      void (*old)(int);
      Handler_Guard() : old(std::signal(SIGTERM, &SDP_Solver::on_sigterm)) {}
      ~Handler_Guard()
      {
        if(old != SIG_ERR)
          std::signal(SIGTERM, old);
      }
    } handler_guard;
    for(long iteration = 1;; ++iteration)
      {
        const auto iter_start = std::chrono::steady_clock::now();
        // run.cxx:332-355: graceful exit on SIGTERM (the caller writes the checkpoint)
        if(sigterm_flag())
          {
            reason = Terminate_Reason::SIGTERM_Received;
            break;
          }
        // run.cxx:357-370: a checkpoint every checkpointInterval seconds
        if(std::chrono::duration_cast<std::chrono::seconds>(iter_start - last_checkpoint_time).count()
           >= parameters.checkpoint_interval)
          {
            save_checkpoint(parameters);
            last_checkpoint_time = std::chrono::steady_clock::now();
          }
        compute_objectives();
        {
          const auto t0 = std::chrono::steady_clock::now();
          hot.cholesky_decomposition(0, X, X_cholesky);
          hot.cholesky_decomposition(1, Y, Y_cholesky);
          hot.compute_bilinear_pairings(Y, A_Y);
          hot_path_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        }
        compute_dual_residues_and_error(A_Y);
        compute_primal_residues_and_error_P_Ax_X();
        Matrix primal_residue_p;
        compute_primal_residues_and_error_p_b_Bx(primal_residue_p);
        bool terminate_now, is_primal_and_dual_feasible;
        compute_feasible_and_termination(parameters, primal_step_length, dual_step_length, iteration, start,
                                         is_primal_and_dual_feasible, reason, terminate_now);
        if(terminate_now)
          break;
        Iteration_Record rec;
        terminate_now = false;
        step(parameters, total_psd_rows, is_primal_and_dual_feasible, X_cholesky, Y_cholesky, primal_residue_p,
             rec.mu, rec.beta_corrector, primal_step_length, dual_step_length, terminate_now, rec.Q_cond_number,
             rec.max_block_cond_number, rec.block_name);
        if(terminate_now)
          {
            reason = Terminate_Reason::MaxComplementarityExceeded;
            break;
          }
        // print_iteration.cxx:77-108
        const auto now = std::chrono::steady_clock::now();
        rec.iteration = iteration;
        rec.total_time = std::chrono::duration<double>(now - start).count();
        rec.iter_time = std::chrono::duration<double>(now - iter_start).count();
        rec.primal_objective = primal_objective;
        rec.dual_objective = dual_objective;
        rec.duality_gap = duality_gap;
        rec.primal_error_P = primal_error_P;
        rec.primal_error_p = primal_error_p;
        rec.dual_error = dual_error;
        rec.R_error = R_error;
        rec.primal_step_length = primal_step_length;
        rec.dual_step_length = dual_step_length;
        if(on_iteration)
          on_iteration(rec);
        iterations.push_back(std::move(rec));
      }
    host_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count()
                   - hot_path_seconds;
    return reason;
  }
};

// ---- output files (src/sdpb/save_solution.cxx:21-165, print_iteration.cxx:77-108,
// save_c_minus_By.hxx, sdpb_util/write_distmatrix.hxx:5-18) ----
inline void write_vector_file(const std::string &path, const Matrix &v)
{
  std::ofstream f(path);
  f << v.h << " " << v.w << "\n";
  for(int j = 0; j < v.w; ++j)
    for(int i = 0; i < v.h; ++i)
      f << format_bigfloat(v(i, j)) << "\n";
  if(!f.good())
    throw std::runtime_error("Error when writing to: " + path);
}
// iterations.json, appended to as the iterations complete (run.cxx:304-316, print_iteration.cxx:77-108): a run
// that dies at iteration k leaves the k-1 records it finished
class Iterations_Json
{
  std::string path;
  size_t written = 0;
  bool open = false;

public:
  explicit Iterations_Json(const std::string &p) : path(p)
  {
    if(path.empty())
      return;
    std::ofstream f(path);
    f << "[";
    open = f.good();
    if(!open)
      fprintf(stderr, "Warning: cannot write to %s\n", path.c_str());
  }
  void append(const Iteration_Record &r)
  {
    if(!open)
      return;
    std::ofstream f(path, std::ios::app);
    char tm[96];
    snprintf(tm, sizeof tm, ", \"total_time\": %.3f, \"iter_time\": %.3f", r.total_time, r.iter_time);
    f << (written ? "," : "") << "\n{ \"iteration\":" << r.iteration << tm << ", \"mu\": \""
      << format_bigfloat(r.mu) << "\", \"P-obj\": \"" << format_bigfloat(r.primal_objective)
      << "\", \"D-obj\": \"" << format_bigfloat(r.dual_objective) << "\", \"gap\": \""
      << format_bigfloat(r.duality_gap) << "\", \"P-err\": \"" << format_bigfloat(r.primal_error_P)
      << "\", \"p-err\": \"" << format_bigfloat(r.primal_error_p) << "\", \"D-err\": \""
      << format_bigfloat(r.dual_error) << "\", \"R-err\": \"" << format_bigfloat(r.R_error)
      << "\", \"P-step\": \"" << format_bigfloat(r.primal_step_length) << "\", \"D-step\": \""
      << format_bigfloat(r.dual_step_length) << "\", \"beta\": \"" << format_bigfloat(r.beta_corrector)
      << "\", \"Q_cond_number\": \"" << format_bigfloat(r.Q_cond_number)
      << "\", \"max_block_cond_number\": \"" << format_bigfloat(r.max_block_cond_number)
      << "\", \"block_name\": \"" << r.block_name << "\" }";
    ++written;
  }
  // closes the array; idempotent
  void close()
  {
    if(!open)
      return;
    std::ofstream f(path, std::ios::app);
    f << "\n]";
    open = false;
    if(!f.good())
      throw std::runtime_error("Error when writing to: " + path);
  }
  ~Iterations_Json()
  {
    try
      {
        close();
      }
    catch(...)
      {}
  }
};
inline void save_solution(const SDP_Solver &solver, Terminate_Reason reason, long runtime_seconds,
                          const std::string &out_dir, const std::string &write_solution)
{
  mkdir(out_dir.c_str(), 0777);
  {
    std::ofstream f(out_dir + "/out.txt");
    f << "terminateReason = \"" << to_string(reason) << "\";\n"
      << "primalObjective = " << format_bigfloat(solver.primal_objective) << ";\n"
      << "dualObjective   = " << format_bigfloat(solver.dual_objective) << ";\n"
      << "dualityGap      = " << format_bigfloat(solver.duality_gap) << ";\n"
      << "primalError     = " << format_bigfloat(solver.primal_error()) << ";\n"
      << "dualError       = " << format_bigfloat(solver.dual_error) << ";\n"
      << "Solver runtime  = " << runtime_seconds << ";\n";
    if(!f.good())
      throw std::runtime_error("Error when writing to: " + out_dir + "/out.txt");
  }
  auto wants = [&](const std::string &what) {
    std::stringstream ss(write_solution);
    std::string item;
    while(std::getline(ss, item, ','))
      if(item == what)
        return true;
    return false;
  };
  if(wants("y"))
    write_vector_file(out_dir + "/y.txt", solver.y);
  if(wants("z") && !solver.sdp.normalization.empty())
    {
      // save_solution.cxx:75-118: insert z[max_index] so that n.z == 1
      const auto &nrm = solver.sdp.normalization;
      size_t max_index = 0;
      for(size_t i = 1; i < nrm.size(); ++i)
        if(Abs(nrm[i]) > Abs(nrm[max_index]))
          max_index = i;
      const int Ny = solver.y.h;
      Matrix z(Ny + 1, 1);
      for(int i = 0; i < (int)max_index; ++i)
        z(i, 0) = solver.y(i, 0);
      for(int i = (int)max_index; i < Ny; ++i)
        z(i + 1, 0) = solver.y(i, 0);
      BigFloat nz;
      for(int i = 0; i <= Ny; ++i)
        nz += nrm[i] * z(i, 0);
      z((int)max_index, 0) = (BigFloat(1) - nz) / nrm[max_index];
      write_vector_file(out_dir + "/z.txt", z);
    }
  for(size_t j = 0; j < solver.x.size(); ++j)
    {
      if(wants("x"))
        write_vector_file(out_dir + "/x_" + std::to_string(j) + ".txt", solver.x[j]);
      for(int parity = 0; parity < 2; ++parity)
        {
          const std::string suffix = std::to_string(2 * j + parity) + ".txt";
          if(wants("X") && solver.X[2 * j + parity].h)
            write_vector_file(out_dir + "/X_matrix_" + suffix, solver.X[2 * j + parity]);
          if(wants("Y") && solver.Y[2 * j + parity].h)
            write_vector_file(out_dir + "/Y_matrix_" + suffix, solver.Y[2 * j + parity]);
        }
    }
  // c - B y (save_c_minus_By.hxx)
  {
    mkdir((out_dir + "/c_minus_By").c_str(), 0777);
    std::ofstream f(out_dir + "/c_minus_By/c_minus_By.json");
    f << "{\"c_minus_By\":[";
    for(size_t j = 0; j < solver.x.size(); ++j)
      {
        Matrix v(solver.sdp.primal_objective_c[j]);
        gemm_nn(BigFloat(-1), solver.sdp.free_var_matrix[j], solver.y, BigFloat(1), v);
        f << (j ? "," : "") << "[";
        for(int i = 0; i < v.h; ++i)
          f << (i ? "," : "") << "\"" << format_bigfloat(v(i, 0)) << "\"";
        f << "]";
      }
    f << "]}";
  }
}

// Reads the SDP, runs the solver on `hot`, writes out_dir/{out.txt, iterations.json, x_j, y, z, c_minus_By}.
// `make_hot_path` is called after the SDP is read (it needs the shapes).
inline Terminate_Reason
solve(const std::string &sdp_dir, const std::string &out_dir, const Solver_Parameters &parameters,
      const std::function<std::unique_ptr<Hot_Path>(const Block_Info &, const SDP &)> &make_hot_path,
      bool verbose, std::string *summary = nullptr)
{
  set_precision(parameters.precision);
  Block_Info block_info;
  SDP sdp;
  read_sdp(sdp_dir, block_info, sdp);
  std::unique_ptr<Hot_Path> hot = make_hot_path(block_info, sdp);
  // sdpb/SDPB_Parameters.cxx:176-193: checkpointDir defaults to <sdpDir>.ck, initialCheckpointDir to
  // checkpointDir; an explicit initialCheckpointDir must hold a checkpoint
  Solver_Parameters par = parameters;
  if(!par.checkpoint_out_set)
    {
      std::string base = sdp_dir;
      while(base.size() > 1 && base.back() == '/')
        base.pop_back();
      par.checkpoint_out = base + ".ck";
    }
  if(!par.checkpoint_in_set)
    par.checkpoint_in = par.checkpoint_out;
  SDP_Solver solver(par, block_info, sdp, *hot);
  {
    std::ostringstream o; // the "options" object of checkpoint.json (Solver_Parameters.cxx:160-188)
    o << "{\n    \"precision\": \"" << par.precision << "\",\n    \"maxIterations\": \"" << par.max_iterations
      << "\",\n    \"checkpointInterval\": \"" << par.checkpoint_interval << "\",\n    \"sdpDir\": \"" << sdp_dir
      << "\",\n    \"outDir\": \"" << out_dir << "\",\n    \"checkpointDir\": \"" << par.checkpoint_out
      << "\",\n    \"initialCheckpointDir\": \"" << par.checkpoint_in << "\"\n}\n";
    solver.options_json = o.str();
  }
  if(!out_dir.empty())
    create_directories(out_dir);
  Iterations_Json iterations_json(out_dir.empty() ? std::string() : out_dir + "/iterations.json");
  solver.on_iteration = [&](const Iteration_Record &r) {
    iterations_json.append(r);
    if(verbose)
      {
        printf("%4ld %8.2f  mu %.3e  P-obj %.10e  D-obj %.10e  gap %.2e  P-err %.2e  D-err %.2e  steps %.3g %.3g\n",
               r.iteration, r.total_time, r.mu.to_double(), r.primal_objective.to_double(),
               r.dual_objective.to_double(), r.duality_gap.to_double(),
               Max(r.primal_error_P, r.primal_error_p).to_double(), r.dual_error.to_double(),
               r.primal_step_length.to_double(), r.dual_step_length.to_double());
        fflush(stdout);
      }
  };
  const auto t0 = std::chrono::steady_clock::now();
  const Terminate_Reason reason = solver.run(par);
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  iterations_json.close();
  // sdpb/solve.cxx:81-88: a final checkpoint unless --noFinalCheckpoint; always after SIGTERM
  if(reason == Terminate_Reason::SIGTERM_Received || !par.no_final_checkpoint)
    solver.save_checkpoint(par);
  if(!out_dir.empty())
    save_solution(solver, reason, (long)secs, out_dir, par.write_solution);
  if(summary)
    {
      char buf[512];
      snprintf(buf, sizeof buf,
               "{\"terminateReason\": \"%s\", \"iterations\": %zu, \"seconds\": %.3f, \"hot_path_seconds\": %.3f, "
               "\"host_seconds\": %.3f, \"hot_path\": \"%s\", \"checkpoint_loaded\": %s, \"checkpoint_generation\": %ld}",
               to_string(reason), solver.iterations.size(), secs, solver.hot_path_seconds, solver.host_seconds,
               hot->name().c_str(), solver.loaded_checkpoint ? "true" : "false", solver.checkpoint.current_generation);
      *summary = buf;
    }
  return reason;
}
} // namespace sdpb_host

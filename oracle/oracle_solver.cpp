// ORACLE — TEST INFRASTRUCTURE ONLY.  The host solver of
// sdpb_b200/csrc/host/solver.hpp (SDP_Solver::run/step, reference
// src/sdp_solve/SDP_Solver/run/run.cxx:184-470) bound to the CPU oracle's hot
// path (oracle_capi.cpp), so that the oracle can be replayed against the
// reference's own end-to-end goldens (test/data/end-to-end_tests/*/output) on
// a machine without a GPU.  tests/test_golden_trajectory.py is the caller.
#include "../sdpb_b200/csrc/host/cli.hpp"
#include "../sdpb_b200/csrc/host/hot_path_c.hpp"

#include <cstdlib>
#include <cstring>

struct oracle_ctx;
extern "C" {
int oracle_create(oracle_ctx **out, int prec_bits, int num_blocks, const int *dims, const int *num_points,
                  int N);
void oracle_destroy(oracle_ctx *c);
const char *oracle_last_error(const oracle_ctx *c);
int oracle_set_block(oracle_ctx *c, int j, const uint64_t *B, const uint64_t *bases_even,
                     const uint64_t *bases_odd);
int oracle_cholesky_decomposition(oracle_ctx *c, int which, const uint64_t *const *A, uint64_t *const *L);
int oracle_compute_bilinear_pairings(oracle_ctx *c, const uint64_t *const *Y, uint64_t *const *A_X_inv,
                                     uint64_t *const *A_Y);
int oracle_initialize_schur_complement_solver(oracle_ctx *c, uint64_t *const *schur_complement_cholesky,
                                              uint64_t *const *schur_off_diagonal, uint64_t *Q,
                                              int32_t *block_timings_ms);
int oracle_solve_schur_complement_equation(oracle_ctx *c, uint64_t *const *dx, uint64_t *dy);
int oracle_scale_multiply_add(oracle_ctx *c, int alpha, const uint64_t *const *A, const uint64_t *const *B, int beta,
                              uint64_t *const *C);
int oracle_direction_begin(oracle_ctx *c, uint64_t *block_traces);
int oracle_direction_R_errors(oracle_ctx *c, const uint64_t *mu, uint64_t *block_maxima);
int oracle_direction_set_residues(oracle_ctx *c, const uint64_t *const *primal_residues,
                                  const uint64_t *const *dual_residues, const uint64_t *primal_residue_p);
int oracle_compute_search_direction(oracle_ctx *c, const uint64_t *beta_mu, int is_corrector);
int oracle_direction_frobenius(oracle_ctx *c, uint64_t *block_products);
int oracle_direction_get(oracle_ctx *c, uint64_t *const *dx, uint64_t *const *dX, uint64_t *dy, uint64_t *const *dY);
int oracle_step_length(oracle_ctx *c, int which, uint64_t *block_min_eigenvalues);
}

using namespace sdpb_host;

static Hot_Path_Table oracle_table(const Block_Info &bi, const SDP &sdp, int prec)
{
  oracle_ctx *c = nullptr;
  if(oracle_create(&c, prec, bi.num_blocks(), bi.dimensions.data(), bi.num_points.data(), sdp.N()))
    throw std::runtime_error("oracle_create failed");
  Hot_Path_Table t;
  t.ctx = c;
  t.set_block = [](void *x, int j, const uint64_t *B, const uint64_t *e, const uint64_t *o) {
    return oracle_set_block((oracle_ctx *)x, j, B, e, o);
  };
  t.cholesky_decomposition = [](void *x, int which, const uint64_t *const *A, uint64_t *const *L) {
    return oracle_cholesky_decomposition((oracle_ctx *)x, which, A, L);
  };
  t.compute_bilinear_pairings
    = [](void *x, const uint64_t *const *Y, uint64_t *const *AX, uint64_t *const *AY) {
        return oracle_compute_bilinear_pairings((oracle_ctx *)x, Y, AX, AY);
      };
  t.initialize_schur_complement_solver
    = [](void *x, uint64_t *const *L, uint64_t *const *P, uint64_t *Q, int32_t *ms) {
        return oracle_initialize_schur_complement_solver((oracle_ctx *)x, L, P, Q, ms);
      };
  t.solve_schur_complement_equation = [](void *x, uint64_t *const *dx, uint64_t *dy) {
    return oracle_solve_schur_complement_equation((oracle_ctx *)x, dx, dy);
  };
  t.scale_multiply_add = [](void *x, int al, const uint64_t *const *A, const uint64_t *const *B, int be,
                            uint64_t *const *C) { return oracle_scale_multiply_add((oracle_ctx *)x, al, A, B, be, C); };
  // SDPB_ORACLE_HOST_DIRECTION=1: leave the direction to the solver's own host code path
  // (direction.hpp called directly) instead of the oracle_direction_* entry points -- the two must
  // give the same trajectory (tests/test_golden_trajectory.py)
  if(!getenv("SDPB_ORACLE_HOST_DIRECTION"))
    {
      t.direction_begin = [](void *x, uint64_t *tr) { return oracle_direction_begin((oracle_ctx *)x, tr); };
      t.direction_R_errors = [](void *x, const uint64_t *mu, uint64_t *mx) {
        return oracle_direction_R_errors((oracle_ctx *)x, mu, mx);
      };
      t.direction_set_residues = [](void *x, const uint64_t *const *pr, const uint64_t *const *dr, const uint64_t *p) {
        return oracle_direction_set_residues((oracle_ctx *)x, pr, dr, p);
      };
      t.compute_search_direction = [](void *x, const uint64_t *bm, int corr) {
        return oracle_compute_search_direction((oracle_ctx *)x, bm, corr);
      };
      t.direction_frobenius = [](void *x, uint64_t *fp) { return oracle_direction_frobenius((oracle_ctx *)x, fp); };
      t.direction_get = [](void *x, uint64_t *const *dx, uint64_t *const *dX, uint64_t *dy, uint64_t *const *dY) {
        return oracle_direction_get((oracle_ctx *)x, dx, dX, dy, dY);
      };
      t.step_length = [](void *x, int which, uint64_t *mins) { return oracle_step_length((oracle_ctx *)x, which, mins); };
    }
  t.last_error = [](const void *x) { return oracle_last_error((const oracle_ctx *)x); };
  t.destroy = [](void *x) { oracle_destroy((oracle_ctx *)x); };
  t.name = "cpu-oracle(libgmp)";
  return t;
}

// test hook: rewrite a JSON sdp directory with binary block data (row N4)
extern "C" int oracle_sdp_to_binary(const char *in_dir, const char *out_dir, int precision, char *err, size_t errlen)
{
  try
    {
      set_precision(precision);
      convert_sdp_to_binary(in_dir, out_dir);
      return 0;
    }
  catch(std::exception &e)
    {
      if(err && errlen)
        {
          strncpy(err, e.what(), errlen - 1);
          err[errlen - 1] = 0;
        }
      return 1;
    }
}

// argv: the reference's sdpb options (--sdpDir, --outDir, --precision, ...).
// Returns 0 and writes a one-line JSON summary, or 1 with the error text.
extern "C" int oracle_solve(int argc, const char *const *argv, char *summary, size_t summary_len)
{
  try
    {
      const Solve_Options o = parse_options(argc, argv);
      std::string s;
      solve(
        o.sdp_dir, o.out_dir, o.parameters,
        [&](const Block_Info &bi, const SDP &sdp) {
          return std::unique_ptr<Hot_Path>(
            new Hot_Path_C(oracle_table(bi, sdp, o.parameters.precision), bi, sdp));
        },
        o.verbose, &s);
      if(summary && summary_len)
        {
          strncpy(summary, s.c_str(), summary_len - 1);
          summary[summary_len - 1] = 0;
        }
      return 0;
    }
  catch(std::exception &e)
    {
      if(summary && summary_len)
        {
          strncpy(summary, e.what(), summary_len - 1);
          summary[summary_len - 1] = 0;
        }
      return 1;
    }
}

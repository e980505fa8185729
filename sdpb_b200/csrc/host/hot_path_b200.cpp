// The product hot path: Hot_Path bound to the sm_100a kernels through the
// C-ABI of include/sdpb_b200.h (libsdpb_b200.so), and the sdpb_b200_solve entry
// point of include/sdpb_b200_solver.h.  No CPU implementation of the hot path
// is linked here; sdpb_b200_create fails loudly without a CUDA device.
#include "../../../include/sdpb_b200.h"
#include "../../../include/sdpb_b200_solver.h"
#include "cli.hpp"
#include "hot_path_c.hpp"

#include <cstring>

using namespace sdpb_host;

static Hot_Path_Table b200_table(const Block_Info &bi, const SDP &sdp, int prec, int device)
{
  sdpb_b200_ctx *c = nullptr;
  char err[512] = {0};
  if(sdpb_b200_create(&c, prec, device, bi.num_blocks(), bi.dimensions.data(), bi.num_points.data(),
                      sdp.N(), err, sizeof err))
    throw std::runtime_error(std::string("sdpb_b200_create: ") + err);
  Hot_Path_Table t;
  t.ctx = c;
  t.set_block = [](void *x, int j, const uint64_t *B, const uint64_t *e, const uint64_t *o) {
    return sdpb_b200_set_block((sdpb_b200_ctx *)x, j, B, e, o);
  };
  t.cholesky_decomposition = [](void *x, int which, const uint64_t *const *A, uint64_t *const *L) {
    return sdpb_b200_cholesky_decomposition((sdpb_b200_ctx *)x, which, A, L);
  };
  t.compute_bilinear_pairings
    = [](void *x, const uint64_t *const *Y, uint64_t *const *AX, uint64_t *const *AY) {
        return sdpb_b200_compute_bilinear_pairings((sdpb_b200_ctx *)x, Y, AX, AY);
      };
  t.initialize_schur_complement_solver
    = [](void *x, uint64_t *const *L, uint64_t *const *P, uint64_t *Q, int32_t *ms) {
        return sdpb_b200_initialize_schur_complement_solver((sdpb_b200_ctx *)x, L, P, Q, ms);
      };
  t.solve_schur_complement_equation = [](void *x, uint64_t *const *dx, uint64_t *dy) {
    return sdpb_b200_solve_schur_complement_equation((sdpb_b200_ctx *)x, dx, dy);
  };
  t.scale_multiply_add = [](void *x, int al, const uint64_t *const *A, const uint64_t *const *B, int be,
                            uint64_t *const *C) {
    return sdpb_b200_scale_multiply_add((sdpb_b200_ctx *)x, al, A, B, be, C);
  };
  t.direction_begin = [](void *x, uint64_t *tr) { return sdpb_b200_direction_begin((sdpb_b200_ctx *)x, tr); };
  t.direction_R_errors = [](void *x, const uint64_t *mu, uint64_t *mx) {
    return sdpb_b200_direction_R_errors((sdpb_b200_ctx *)x, mu, mx);
  };
  t.direction_set_residues = [](void *x, const uint64_t *const *pr, const uint64_t *const *dr, const uint64_t *p) {
    return sdpb_b200_direction_set_residues((sdpb_b200_ctx *)x, pr, dr, p);
  };
  t.compute_search_direction = [](void *x, const uint64_t *bm, int corr) {
    return sdpb_b200_compute_search_direction((sdpb_b200_ctx *)x, bm, corr);
  };
  t.direction_frobenius = [](void *x, uint64_t *fp) { return sdpb_b200_direction_frobenius((sdpb_b200_ctx *)x, fp); };
  t.direction_get = [](void *x, uint64_t *const *dx, uint64_t *const *dX, uint64_t *dy, uint64_t *const *dY) {
    return sdpb_b200_direction_get((sdpb_b200_ctx *)x, dx, dX, dy, dY);
  };
  t.step_length = [](void *x, int which, uint64_t *mins) { return sdpb_b200_step_length((sdpb_b200_ctx *)x, which, mins); };
  t.last_error = [](const void *x) { return sdpb_b200_last_error((const sdpb_b200_ctx *)x); };
  t.destroy = [](void *x) { sdpb_b200_destroy((sdpb_b200_ctx *)x); };
  t.name = "sm_100a(libsdpb_b200.so)";
  return t;
}

// an sdp directory with JSON block data rewritten with binary block data (block_data_bin.hpp)
extern "C" int sdpb_b200_sdp_to_binary(const char *in_dir, const char *out_dir, int precision, char *err, size_t errlen)
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

extern "C" int sdpb_b200_solve(int argc, const char *const *argv, char *summary, size_t summary_len)
{
  auto put = [&](const std::string &s) {
    if(summary && summary_len)
      {
        strncpy(summary, s.c_str(), summary_len - 1);
        summary[summary_len - 1] = 0;
      }
  };
  try
    {
      const Solve_Options o = parse_options(argc, argv);
      std::string s;
      solve(
        o.sdp_dir, o.out_dir, o.parameters,
        [&](const Block_Info &bi, const SDP &sdp) {
          return std::unique_ptr<Hot_Path>(
            new Hot_Path_C(b200_table(bi, sdp, o.parameters.precision, o.device), bi, sdp));
        },
        o.verbose, &s);
      put(s);
      return 0;
    }
  catch(std::exception &e)
    {
      put(e.what());
      return 1;
    }
}

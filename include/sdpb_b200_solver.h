/* sdpb_b200 — host solver entry point.
 *
 * SDPB's solver surface is the `sdpb` executable: main() -> solve() ->
 * SDP_Solver::run / SDP_Solver::step (reference src/sdpb/main.cxx:31,
 * src/sdpb/solve.cxx:23-110, src/sdp_solve/SDP_Solver.hxx:79-121).  This
 * library holds that host side — the reference's iteration structure on GMP
 * mpf scalars — with the hot path (cholesky_decomposition,
 * compute_bilinear_pairings, initialize_schur_complement_solver) executed on
 * the GPU through include/sdpb_b200.h.  There is no CPU implementation of the
 * hot path in this library: without a CUDA device the call fails.
 */
#ifndef SDPB_B200_SOLVER_H
#define SDPB_B200_SOLVER_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Runs the solver.  argv holds the reference's sdpb options
 * (src/sdpb/SDPB_Parameters.cxx:21-94, src/sdp_solve/Solver_Parameters/Solver_Parameters.cxx:20-157)
 * as "--key=value", "--key value" or bare flags:
 *   --sdpDir DIR (required; a directory as pmp2sdp writes it: control.json, objectives.json,
 *                 block_info_<j>.json and block_data_<j>.bin (Boost binary, the default of
 *                 pmp2sdp) or block_data_<j>.json; src/sdp_solve/SDP/read_block_data/SDP_Block_Data.cxx:32-48),
 *                 or the stored zip archive `pmp2sdp --zip` packs that directory into)
 *   --outDir DIR (out.txt, iterations.json, x_j.txt, y.txt, z.txt, c_minus_By/c_minus_By.json;
 *                 src/sdpb/save_solution.cxx:21-165, SDP_Solver/run/print_iteration.cxx:77-108)
 *   --precision BITS, --maxIterations, --dualityGapThreshold, --primalErrorThreshold,
 *   --dualErrorThreshold, --initialMatrixScalePrimal/Dual, --feasibleCenteringParameter,
 *   --infeasibleCenteringParameter, --stepLengthReduction, --maxComplementarity,
 *   --minPrimalStep, --minDualStep, --findPrimalFeasible, --findDualFeasible,
 *   --detectPrimalFeasibleJump, --detectDualFeasibleJump, --writeSolution x,y,z,X,Y, --maxRuntime
 *   --checkpointDir DIR (default <sdpDir>.ck; "" = no checkpoints), --initialCheckpointDir DIR
 *   (default checkpointDir; given explicitly it must hold a checkpoint), --checkpointInterval SECONDS,
 *   --noFinalCheckpoint: binary checkpoints checkpoint_<generation>_0 + checkpoint.json, or a text
 *   checkpoint (x_j.txt, y.txt, X_matrix_b.txt, Y_matrix_b.txt); SIGTERM ends the run gracefully
 *   with a checkpoint (src/sdp_solve/SDP_Solver/save_checkpoint.cxx, load_checkpoint/,
 *   run/run.cxx:332-370, src/sdpb/solve.cxx:81-88)
 *   --device N (CUDA ordinal, default 0), --verbose
 * Returns 0 and writes a one-line JSON summary (terminateReason, iterations,
 * seconds, hot_path_seconds, host_seconds) into `summary`; on failure returns
 * non-zero and `summary` holds the error text (the reference's RUNTIME_ERROR
 * wording for numerical failures). */
int sdpb_b200_solve(int argc, const char *const *argv, char *summary, size_t summary_len);

/* Rewrites an sdp directory with JSON block data as one with binary block data
 * (pmp2sdp --outputFormat=bin; src/pmp2sdp/write_block_data.cxx, src/sdpb_util/boost_serialization.hxx:17-97)
 * at the given --precision.  Returns 0, or non-zero with the error text in `err`. */
int sdpb_b200_sdp_to_binary(const char *in_dir, const char *out_dir, int precision, char *err, size_t errlen);

#ifdef __cplusplus
}
#endif
#endif

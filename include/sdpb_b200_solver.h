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
 *   --sdpDir DIR (required; the JSON form written by pmp2sdp --outputFormat=json)
 *   --outDir DIR (out.txt, iterations.json, x_j.txt, y.txt, z.txt, c_minus_By/c_minus_By.json;
 *                 src/sdpb/save_solution.cxx:21-165, SDP_Solver/run/print_iteration.cxx:77-108)
 *   --precision BITS, --maxIterations, --dualityGapThreshold, --primalErrorThreshold,
 *   --dualErrorThreshold, --initialMatrixScalePrimal/Dual, --feasibleCenteringParameter,
 *   --infeasibleCenteringParameter, --stepLengthReduction, --maxComplementarity,
 *   --minPrimalStep, --minDualStep, --findPrimalFeasible, --findDualFeasible,
 *   --detectPrimalFeasibleJump, --detectDualFeasibleJump, --writeSolution x,y,z,X,Y, --maxRuntime
 *   --device N (CUDA ordinal, default 0), --verbose
 * Returns 0 and writes a one-line JSON summary (terminateReason, iterations,
 * seconds, hot_path_seconds, host_seconds) into `summary`; on failure returns
 * non-zero and `summary` holds the error text (the reference's RUNTIME_ERROR
 * wording for numerical failures). */
int sdpb_b200_solve(int argc, const char *const *argv, char *summary, size_t summary_len);

#ifdef __cplusplus
}
#endif
#endif

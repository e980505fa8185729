/* sdpb_b200 — C-ABI of the B200-native Schur-complement step of SDPB.
 *
 * SDPB has no plugin/FFI layer.  The seam this library replaces is the three
 * free functions SDP_Solver::run / SDP_Solver::step call once per Newton
 * iteration (paths relative to the reference checkout):
 *
 *   cholesky_decomposition              src/sdp_solve/SDP_Solver/run/run.cxx:14-17    (called :386-387)
 *   compute_bilinear_pairings           src/sdp_solve/SDP_Solver/run/run.cxx:37-45    (called :390)
 *   initialize_schur_complement_solver  src/sdp_solve/SDP_Solver/run/step/step.cxx:12-24 (called :123)
 *
 * plus the immutable SDP data they read (SDP::bases_blocks, SDP::free_var_matrix,
 * src/sdp_solve/SDP.hxx:84-97) and the persistent syrk context created once in
 * run.cxx:250-255.  INTEGRATION.md shows the shim a maintainer adds on the
 * reference side.
 *
 * Scalars.  El::BigFloat is GMP's mpf_t.  Across this ABI every number is a
 * "packed element" of sdpb_b200_elem_words(prec) 64-bit little-endian words:
 *
 *   word 0        low 32 bits: exponent in 64-bit limbs (int32, == _mp_exp)
 *                 high 32 bits: sign (int32: -1, 0, +1)
 *   word 1..NL    NL = (prec+63)/64 + 2 mantissa limbs, least significant
 *                 first, TOP-ALIGNED: the mpf's |_mp_size| limbs occupy the
 *                 highest positions, lower positions are zero
 *   (one pad word when NL+1 is odd, so elements are 16-byte aligned)
 *
 * i.e. exactly the fields of __mpf_struct; sdpb_b200_pack_mpf below is the
 * whole conversion.  Matrices are column-major arrays of packed elements with
 * leading dimension = height.
 *
 * Threading: one caller thread per context; calls are synchronous.  The
 * library owns all device memory and streams; host buffers are caller-owned
 * and only touched during a call.  Every function returns 0 on success; on
 * failure it returns non-zero and sdpb_b200_last_error(ctx) names the stage,
 * block index and parity the way the reference's RUNTIME_ERROR texts do
 * (cholesky_decomposition.cxx:20-25, compute_Q.cxx:33-38,
 * initialize_schur_complement_solver.cxx:100-103).  There is no CPU fallback:
 * without a CUDA device sdpb_b200_create fails.
 */
#ifndef SDPB_B200_H
#define SDPB_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdpb_b200_ctx sdpb_b200_ctx;

/* error codes */
#define SDPB_B200_OK 0
#define SDPB_B200_ERR_ARG 1      /* bad argument / unsupported precision */
#define SDPB_B200_ERR_CUDA 2     /* CUDA runtime failure (incl. no device) */
#define SDPB_B200_ERR_NOT_HPD 3  /* non-positive pivot in a Cholesky */
#define SDPB_B200_ERR_Q_DIAG 4   /* check_normalized_Q_diagonal failed */
#define SDPB_B200_ERR_STATE 5    /* calls out of order */

/* 64-bit words per packed element at this --precision; 0 if unsupported. */
int sdpb_b200_elem_words(int prec_bits);
/* mantissa limbs NL stored per element. */
int sdpb_b200_stored_limbs(int prec_bits);

/* Create the per-process solver context (replaces
 * initialize_bigint_syrk_context, run.cxx:250-255, and the allocations of
 * step.cxx:92-115).  dims[j], num_points[j] as in Block_Info
 * (Block_Info.hxx:23-24) for the blocks owned by THIS process; N =
 * dual_objective_b.Height().  `device` is the CUDA ordinal. */
int sdpb_b200_create(sdpb_b200_ctx **ctx, int prec_bits, int device,
                     int num_blocks, const int *dims, const int *num_points,
                     int N, char *err, size_t errlen);
void sdpb_b200_destroy(sdpb_b200_ctx *ctx);
const char *sdpb_b200_last_error(const sdpb_b200_ctx *ctx);

/* Upload the immutable SDP data of local block j (SDP.hxx:84-97):
 *   B            free_var_matrix block, P_j x N, P_j = n*m*(m+1)/2
 *   bases_even   bilinear_bases[2j],   (floor(d/2)+1) x n, d = n-1
 *   bases_odd    bilinear_bases[2j+1], floor((d+1)/2) x n   (may be empty)
 * The library forms bases_blocks = I_m (x) basis itself
 * (SDP/set_bases_blocks.cxx:24-47). */
int sdpb_b200_set_block(sdpb_b200_ctx *ctx, int j, const uint64_t *B,
                        const uint64_t *bases_even, const uint64_t *bases_odd);

/* cholesky_decomposition (run/cholesky_decomposition.cxx:5-28).
 * which = 0: X -> X_cholesky (kept on the device for the pairings),
 * which = 1: Y -> Y_cholesky.  A[b], L[b] for b = 2*j + parity are host
 * pointers to s x s column-major matrices, s = psd_matrix_block_size
 * (Block_Info.hxx:83-95); blocks with s == 0 may be NULL.  L may be NULL
 * (results stay on the device). */
int sdpb_b200_cholesky_decomposition(sdpb_b200_ctx *ctx, int which,
                                     const uint64_t *const *A,
                                     uint64_t *const *L);

/* compute_bilinear_pairings (run/compute_bilinear_pairings/).  Uses the
 * device-resident X_cholesky of the preceding call and Y from the host.
 * Outputs (optional, may be NULL): for b = 2*j + parity the full symmetric
 * (m*n) x (m*n) matrices  V^T X^-1 V  and  V^T Y V ; the reference's tiles are
 *   A_X_inv[parity][j][cb][rb](row,col) = AX[b][(cb*n+row) + (rb*n+col)*m*n]
 *   A_Y    [parity][j][cb][rb](row,col) = AY[b][(cb*n+col) + (rb*n+row)*m*n]
 * (compute_A_X_inv.cxx:45-55, compute_A_Y.cxx:51-63). */
int sdpb_b200_compute_bilinear_pairings(sdpb_b200_ctx *ctx,
                                        const uint64_t *const *Y,
                                        uint64_t *const *A_X_inv,
                                        uint64_t *const *A_Y);

/* initialize_schur_complement_solver
 * (run/step/initialize_schur_complement_solver/): S assembly, per-block
 * Cholesky + L^-1 B, column normalisation, exact integer syrk, restore,
 * Cholesky(UPPER, Q).  Outputs (each may be NULL = keep on device only):
 *   schur_complement_cholesky[j]  P_j x P_j lower factor L_j
 *   schur_off_diagonal[j]         P_j x N  = L_j^-1 B_j (after the
 *                                 normalise/restore round trip, compute_Q.cxx:117,130)
 *   Q                             N x N, upper Cholesky factor, lower part 0
 *   block_timings_ms[j]           += ms spent in cholesky_j + solve_j
 *                                 (compute_Q.cxx:40,52) */
int sdpb_b200_initialize_schur_complement_solver(
  sdpb_b200_ctx *ctx, uint64_t *const *schur_complement_cholesky,
  uint64_t *const *schur_off_diagonal, uint64_t *Q, int32_t *block_timings_ms);

/* solve_schur_complement_equation
 * (run/step/compute_search_direction/solve_schur_complement_equation.cxx:16-79;
 * called twice per iteration, compute_search_direction.cxx:62, for the predictor
 * and the corrector): with the device-resident L_j, L_j^-1 B_j and chol(Q) of the
 * preceding initialize_schur_complement_solver / schur_step,
 *     dx_j <- L_j^-1 dx_j ;  dy <- dy - sum_j (L_j^-1 B_j)^T dx_j ;
 *     dy <- Q^-1 dy ;  dx_j <- dx_j + (L_j^-1 B_j) dy ;  dx_j <- L_j^-T dx_j .
 * dx[j]: P_j packed elements of local block j (in: r_x, out: dx); dy: N packed
 * elements (in: r_y, out: dy; the same on every rank).  On entry dy must be the FULLY
 * REDUCED right-hand side: the reference stores one share per block (dy.blocks[j] =
 * -B_j^T x_j, plus b on global block 0, compute_primal_residues_and_error_p_b_Bx.cxx) and
 * lets solve_schur_complement_equation.cxx:26-60 sum the shares over blocks and ranks; a
 * caller of this ABI forms r_y = sum_j dy.blocks[j] itself and all-reduces it over its ranks
 * (INTEGRATION.md shows the shim) -- every rank passes the same N values.  With a communicator the
 * per-block partial sums of the dy update are exchanged over NCCL and added in
 * GLOBAL block order, so the result does not depend on the sharding.  Keeping
 * these solves next to the factors removes the largest device->host copy of a
 * step (schur_off_diagonal, P x N elements). */
int sdpb_b200_solve_schur_complement_equation(sdpb_b200_ctx *ctx,
                                              uint64_t *const *dx, uint64_t *dy);

/* scale_multiply_add
 * (run/step/compute_search_direction/scale_multiply_add.cxx:4-16, forward-declared at
 * step.cxx:7 and compute_search_direction.cxx:19): per block-parity b = 2*j + parity
 *     C_b = alpha * A_b * B_b + beta * C_b
 * on s x s blocks, s = psd_matrix_block_size (the shape of X, Y, dX, dY, R, Z).
 * alpha is 1 or -1 and beta 0 or 1 -- the reference's call sites: -X Y (step.cxx:137),
 * (1, 0) (compute_search_direction.cxx:28), (-1, 1) (:60).  A, B, C: host pointers as for
 * cholesky_decomposition; C is read only when beta != 0. */
int sdpb_b200_scale_multiply_add(sdpb_b200_ctx *ctx, int alpha,
                                 const uint64_t *const *A,
                                 const uint64_t *const *B, int beta,
                                 uint64_t *const *C);

/* The diagonals of the Cholesky factors the last step left in HBM -- all that
 * update_cond_numbers (run/step/step.cxx:187-189, sdpb_util/cholesky_condition_number.hxx:8-36)
 * reads of them.  With solve_schur_complement_equation on the device, L_j (sum P_j^2 elements)
 * and L_j^-1 B_j (P x N) never have to cross PCIe: a step's device->host traffic drops from
 * gigabytes to these vectors.  X_diag / Y_diag: stacked over b = 2j + parity; S_diag: stacked
 * over local blocks j (P elements in all); Q_diag: N elements.  Any of them may be NULL. */
int sdpb_b200_cholesky_diagonals(sdpb_b200_ctx *ctx, uint64_t *X_diag, uint64_t *Y_diag,
                                 uint64_t *S_diag, uint64_t *Q_diag);

/* ---- The search direction on the device (compute_search_direction.cxx:44-90) ----------------
 * step() forms R, Z, dx, dX, dy, dY from X, Y, their factors and the residues
 * (run/step/step.cxx:131-176).  These calls keep all of that in HBM: the block products
 * (scale_multiply_add.cxx:4-16), cholesky_solve (cholesky_solve.cxx:4-13), symmetrize
 * (Block_Diagonal_Matrix.hxx:95-109), compute_schur_RHS (compute_schur_RHS.cxx:21-86), the Schur
 * solves and constraint_matrix_weighted_sum (constraint_matrix_weighted_sum.cxx:14-66) run on the
 * X, Y of the preceding step (sdpb_b200_schur_step, or cholesky_decomposition(X) +
 * compute_bilinear_pairings(Y)) and its factors; only the residues go up and per-block scalars
 * come down, until the caller asks for the direction itself.  Per iteration:
 *
 *   direction_begin            minus_XY = -X Y (step.cxx:137); block_traces[b], b = 2j + parity, is
 *                              the trace of block b: mu = -(sum_b traces) / total_psd_rows (:138-146)
 *   direction_R_errors         block_maxima[b] = max |(-XY + mu I)_b| (compute_R_error.hxx)
 *   direction_set_residues     primal_residues (2J blocks, shape of X), dual_residues (J vectors of
 *                              P_j elements), primal_residue_p (N elements, fully reduced -- see
 *                              solve_schur_complement_equation above)
 *   compute_search_direction   beta_mu = beta * mu (one packed element).  is_corrector = 0:
 *                              R = beta mu I - XY; 1: R = beta mu I - XY - dX dY with the dX, dY of
 *                              the preceding (predictor) call (compute_search_direction.cxx:50-56)
 *   direction_frobenius        block_products[b] = sum_ij (X+dX)_ij (Y+dY)_ij of block b
 *                              (corrector_centering_parameter.cxx, frobenius_product_of_sums.cxx)
 *   direction_get              copies dx (J vectors), dX, dY (2J blocks), dy (N) back; any may be NULL
 *
 * Sums over blocks are left to the caller, in the canonical two-level order (groups of 64
 * consecutive blocks) that csrc/host/direction.hpp::ordered_sum spells out; inside a block the
 * order is fixed by the library.  Results are bit-identical to that host restatement. */
int sdpb_b200_direction_begin(sdpb_b200_ctx *ctx, uint64_t *block_traces);
int sdpb_b200_direction_R_errors(sdpb_b200_ctx *ctx, const uint64_t *mu, uint64_t *block_maxima);
int sdpb_b200_direction_set_residues(sdpb_b200_ctx *ctx, const uint64_t *const *primal_residues,
                                     const uint64_t *const *dual_residues,
                                     const uint64_t *primal_residue_p);
int sdpb_b200_compute_search_direction(sdpb_b200_ctx *ctx, const uint64_t *beta_mu, int is_corrector);
int sdpb_b200_direction_frobenius(sdpb_b200_ctx *ctx, uint64_t *block_products);
int sdpb_b200_direction_get(sdpb_b200_ctx *ctx, uint64_t *const *dx, uint64_t *const *dX, uint64_t *dy,
                            uint64_t *const *dY);
/* dX, dY (2J blocks each, the shape of X) from the host into the resident direction -- for a
 * caller that forms the direction itself and only wants step_length on the device.  Needs the
 * factors of a step; either list may be NULL (that object keeps its contents). */
int sdpb_b200_direction_put(sdpb_b200_ctx *ctx, const uint64_t *const *dX, const uint64_t *const *dY);
/* step_length (run/step/step_length/step_length.cxx:27-46, with
 * lower_triangular_inverse_congruence.cxx:5-18 and min_eigenvalue.cxx:8-33) on the resident
 * Cholesky factors and the resident direction: block_min_eigenvalues[b], b = 2j + parity, is the
 * smallest eigenvalue of L_b^-1 dM_b L_b^-T with (L, dM) = (chol X, dX) for which = 0 and
 * (chol Y, dY) for which = 1 (2J packed elements; an empty block gives 0 and is skipped by the
 * caller).  The caller finishes as the reference does: lambda = min over its blocks, MIN-reduced
 * over its ranks (min_eigenvalue.cxx:31-32, exact), step = lambda > -gamma ? 1 : -gamma / lambda.
 * El::HermitianEig is the un-vendored Elemental fork's; the operation order is the one
 * csrc/host/step_length.hpp spells out (Householder tridiagonalisation, then Laguerre's iteration
 * from the Gershgorin bound), and the result is bit-identical to that host restatement. */
int sdpb_b200_step_length(sdpb_b200_ctx *ctx, int which, uint64_t *block_min_eigenvalues);
/* Device time of the last sdpb_b200_step_length, ms (CUDA events). */
float sdpb_b200_last_step_length_ms(const sdpb_b200_ctx *ctx);
/* Laguerre steps per block-parity of the last sdpb_b200_step_length (2J ints); diagnostics. */
int sdpb_b200_step_length_iterations(sdpb_b200_ctx *ctx, int *iterations);
/* Device time of the last direction_begin / compute_search_direction, ms (CUDA events). */
float sdpb_b200_last_direction_ms(const sdpb_b200_ctx *ctx);

/* Device time of the last sdpb_b200_solve_schur_complement_equation, ms (CUDA events). */
float sdpb_b200_last_solve_ms(const sdpb_b200_ctx *ctx);

/* The whole hot path of one Newton iteration in one call:
 * cholesky_decomposition(X), cholesky_decomposition(Y),
 * compute_bilinear_pairings, initialize_schur_complement_solver, with a
 * single host synchronisation at the end.  Pointer arguments as above. */
int sdpb_b200_schur_step(sdpb_b200_ctx *ctx, const uint64_t *const *X,
                         const uint64_t *const *Y, uint64_t *const *X_cholesky,
                         uint64_t *const *Y_cholesky, uint64_t *const *A_X_inv,
                         uint64_t *const *A_Y,
                         uint64_t *const *schur_complement_cholesky,
                         uint64_t *const *schur_off_diagonal, uint64_t *Q,
                         int32_t *block_timings_ms);

/* The same step split at the host<->device boundary, for callers that keep
 * X and Y in HBM between iterations (and for measuring the kernels alone):
 * upload_XY copies X and Y to the device; schur_step_resident runs the whole
 * hot path on the resident copies (it does not modify them, so it can be
 * repeated); download copies back whichever outputs are non-NULL. */
int sdpb_b200_upload_XY(sdpb_b200_ctx *ctx, const uint64_t *const *X,
                        const uint64_t *const *Y);
int sdpb_b200_schur_step_resident(sdpb_b200_ctx *ctx);
int sdpb_b200_download(sdpb_b200_ctx *ctx, uint64_t *const *X_cholesky,
                       uint64_t *const *Y_cholesky, uint64_t *const *A_X_inv,
                       uint64_t *const *A_Y,
                       uint64_t *const *schur_complement_cholesky,
                       uint64_t *const *schur_off_diagonal, uint64_t *Q);

/* Multi-GPU: the J SDP blocks are sharded over `world` processes, one GPU each
 * (the reference's block parallelism: Block_Info::block_indices,
 * sdpb_util/block_mapping/compute_block_grid_mapping.hxx:58-183).  Each process
 * creates its context with ITS blocks only and then joins the communicator:
 * rank 0 calls sdpb_b200_comm_get_unique_id (128 bytes) and the host distributes
 * them (MPI_Bcast in the reference's environment, torch.distributed in bench.py);
 * every rank calls sdpb_b200_comm_init with the global block count and the
 * global index of each of its local blocks.  From then on
 * initialize_schur_complement_solver / schur_step are collective: the per-block
 * column-norm partials (Matrix_Normalizer.cxx:116-139) and the exact integer
 * Q' partial sums (bigint_syrk/restore_and_reduce.cxx:137-212) are combined
 * with NCCL over NVLink; every rank ends up with the same Q factor, and with
 * L_j, L_j^-1 B_j for its own blocks.  The column norms are summed in GLOBAL
 * block order, so results are bit-identical for every sharding. */
#define SDPB_B200_COMM_ID_BYTES 128
int sdpb_b200_comm_get_unique_id(void *id);
int sdpb_b200_comm_init(sdpb_b200_ctx *ctx, int rank, int world, const void *id,
                        int num_blocks_global, const int *global_block_index);

/* The same sharding with every rank inside ONE process on ONE device: ctxs[r] (r < world) are
 * contexts created on the same device, each holding rank r's blocks and driven by its own host
 * thread; the exchanges go through device memory instead of NCCL.  This is how the sharded code
 * path (global-order sums, exact residue sums, panel-distributed Cholesky(Q), sharded Schur solve)
 * is exercised on a single GPU.  Collective calls block until all `world` contexts have made
 * them.  Call once, from one thread, before the first step. */
int sdpb_b200_comm_init_local(sdpb_b200_ctx *const *ctxs, int world, int num_blocks_global,
                              const int *const *global_block_index);

/* Scheduling of a step on the device.  level 1 (default): the independent chains
 * of the step -- chol(X) -> L_X^-1 V -> A_X_inv, Y V -> A_Y, chol(Y), and per block
 * S_j -> chol(S_j) -> L_j^-1 B_j -> norm partials in interleaved groups of blocks --
 * run on side streams, so the latency-bound tails of one chain are filled by the
 * others.  level 0: every kernel on one stream in program order; this is the mode
 * in which sdpb_b200_kernel_timings gives non-overlapping per-kernel durations
 * (and the per-stage times below are exact).  Results are bit-identical. */
int sdpb_b200_set_concurrency(sdpb_b200_ctx *ctx, int level);

/* Device-side timing of the last step, milliseconds per stage (CUDA events):
 * [0] cholesky X+Y  [1] bilinear pairings  [2] S assembly  [3] cholesky S_j +
 * L^-1 B  [4] norms+normalise  [5] exact syrk  [6] restore  [7] Cholesky(Q)
 * [8] whole step on device.  Fills min(n, 9) entries. */
int sdpb_b200_last_timings_ms(const sdpb_b200_ctx *ctx, float *ms, int n);

/* Per-launch timeline of the last sdpb_b200_schur_step_resident: every kernel
 * launch is bracketed by two CUDA events on the launching stream.  Fills up to
 * `max` (name, milliseconds) pairs in launch order; returns the count.  The
 * names are static strings owned by the library. */
int sdpb_b200_kernel_timings(const sdpb_b200_ctx *ctx, int max,
                             const char **names, float *ms);

/* Page-locked host memory for the caller's staging buffers (the reference-side
 * shim packs El::BigFloat into such a buffer anyway; pinning it makes the
 * H2D/D2H copies of a step asynchronous DMA). */
int sdpb_b200_host_alloc(void **p, size_t bytes);
void sdpb_b200_host_free(void *p);

/* Number of CUDA kernels this context has launched since creation. */
long sdpb_b200_kernel_launches(const sdpb_b200_ctx *ctx);

/* Test hook: element-wise mpf-exact scalar operations executed on the device
 * (op: 0 mul, 1 add, 2 sub, 3 div, 4 sqrt(a), 5 a<<k, 6 a>>k, 7 a/4) on
 * `count` packed elements; used to check the device arithmetic against libgmp. */
int sdpb_b200_scalar_op(sdpb_b200_ctx *ctx, int op, int k, long count,
                        const uint64_t *a, const uint64_t *b, uint64_t *r);

/* Convert between a GMP __mpf_struct's fields and a packed element.  Pure host
 * helpers (no device work); this is all a reference-side shim needs. */
void sdpb_b200_pack_mpf(int prec_bits, int mp_size, long mp_exp,
                        const uint64_t *mp_d, uint64_t *out);
/* Writes up to NL limbs to mp_d (caller-provided, >= NL limbs); returns the
 * signed _mp_size and stores _mp_exp. */
int sdpb_b200_unpack_mpf(int prec_bits, const uint64_t *in, uint64_t *mp_d,
                         long *mp_exp);

#ifdef __cplusplus
}
#endif
#endif

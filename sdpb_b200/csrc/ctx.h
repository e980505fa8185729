// Shared between capi.cu (host logic) and launch_nl.cu (one translation unit
// per supported precision, so the heavy templated kernels compile in parallel).
#pragma once
#include "../../include/sdpb_b200.h"
#include "kernels.cuh"
#include "solve.cuh"
#include "direction.cuh"
#include "eig.cuh"

#include <map>
#include <memory>
#include <tuple>
#include <string>
#include <vector>

using namespace sdpb_b200;
struct LocalGroup; // in-process communicator (capi.cu): several contexts on one device

#define CUDA_TRY(c, expr)                                                     \
  do                                                                          \
    {                                                                         \
      cudaError_t e_ = (expr);                                                \
      if(e_ != cudaSuccess)                                                   \
        {                                                                     \
          (c)->error = std::string("CUDA: ") + cudaGetErrorString(e_)         \
                       + " at " + __FILE__ + ":" + std::to_string(__LINE__);  \
          return SDPB_B200_ERR_CUDA;                                          \
        }                                                                     \
    }                                                                         \
  while(0)

// precisions this build instantiates kernels for (stored limbs NL)
#define SDPB_FOR_EACH_NL(F)                                                   \
  F(4) F(6) F(8) F(9) F(10) F(12) F(13) F(14) F(17) F(18) F(26)

struct BlockGeom
{
  int m, n, P, mn;
  int s[2], h[2];
  long row0;
};

struct sdpb_b200_ctx
{
  int prec = 0, nl = 0, es = 0, device = 0, J = 0, N = 0;
  long K = 0; // stacked rows of P
  std::vector<BlockGeom> g;
  std::string error;
  cudaStream_t stream = nullptr;

  // one arena; offsets in 64-bit words
  limb_t *arena = nullptr;
  size_t arena_words = 0;
  // contiguous regions (element offsets handled per block)
  limb_t *B = nullptr, *Pband = nullptr, *S = nullptr;
  limb_t *V = nullptr, *T = nullptr, *X = nullptr, *Y = nullptr, *LY = nullptr,
         *YV = nullptr, *AX = nullptr, *AY = nullptr;
  limb_t *Xin = nullptr, *Yin = nullptr; // pristine inputs of the resident step
  uint64_t *pinned = nullptr;            // host staging for X/Y uploads
  size_t pinned_words = 0;
  limb_t *part = nullptr, *norms = nullptr, *Q = nullptr;
  std::vector<size_t> oB, oS, oV, oXY, oA; // per block / block-parity offsets (words)
  size_t wB = 0, wS = 0, wV = 0, wXY = 0, wA = 0;

  uint32_t *R = nullptr, *Qres = nullptr;
  uint32_t *d_primes = nullptr, *d_pow28 = nullptr, *d_ginv = nullptr,
           *d_M = nullptr, *d_Mhalf = nullptr, *d_pow28p = nullptr;
  uint64_t *d_inv64 = nullptr;
  int NS = 0; // row stride of the residue planes R (N rounded up to 16)
  long KR = 0; // rows of a residue plane: K rounded up to the 64-row stages of syrk_imma_kernel, pad rows zero
  int syrk_imma = 1; // exact syrk on the integer tensor path (syrk_imma.cuh); 0: IMAD.WIDE (syrk_mod_kernel)
  CrtTables crt{};

  // tile-kernel descriptors (tile.cuh); matrices sorted by cost, largest first
  PotrfDesc *d_potrfX = nullptr, *d_potrfY = nullptr, *d_potrfS = nullptr;
  PotrfDesc *d_potrfQ = nullptr;
  TrsmTileDesc *d_trsmT = nullptr, *d_trsmP = nullptr;
  GemmTileDesc *d_gemmAX = nullptr, *d_gemmYV = nullptr, *d_gemmAY = nullptr;
  // sizes of the sorted batches (host copies, for the per-level grids)
  std::vector<int> szXY, szS, szT, szP, szQ;
  int n_gemm = 0, tiles_AX = 0, tiles_YV = 0, tiles_AY = 0;
  // pivot reciprocals (TileGeom::RS 32-bit words each)
  uint32_t *recipX = nullptr, *recipY = nullptr, *recipS = nullptr, *recipQ = nullptr,
           *recipN = nullptr;
  SchurDesc *d_schur = nullptr;
  BandDesc *d_bands = nullptr;
  std::vector<BandDesc> h_bands;
  int *d_status = nullptr; // [2J X | 2J Y | J S | 1 Q]
  int *d_flags = nullptr;  // [0] overflow, [1] first bad Q diagonal
  int max_s = 0, max_mn = 0, max_P = 0;

  // multi-GPU (sdpb_b200_comm_init): blocks sharded over `world` ranks; `part`
  // then holds one row of column-norm partials per GLOBAL block
  void *comm = nullptr; // ncclComm_t
  int rank = 0, world = 1, J_global = 0;
  limb_t *part_global = nullptr;
  int (*allreduce)(sdpb_b200_ctx *, void *buf, size_t count, int is_u64, const char *label) = nullptr;
  // Cholesky(Q) by panels over the ranks (block-cyclic tile columns) from this N on: the owner of
  // a block column factors it and broadcasts it (NCCL over NVLink), every rank applies it to the
  // tile columns it owns.  Below the threshold the N pivots' serial chain dominates and every
  // rank factors its own copy of Q.
  int (*bcast)(sdpb_b200_ctx *, void *buf, size_t bytes, int root, const char *label) = nullptr;
  limb_t *qpanel = nullptr; // one packed block column of Q + its status word
  int qdist_min_N = 512;
  std::shared_ptr<LocalGroup> local; // sdpb_b200_comm_init_local: the exchanges stay inside the process
  std::vector<int> gidx;             // global index of each local block (error texts name global blocks)
  int *d_fail = nullptr;             // [0] this rank failed, summed over the ranks at the end of a step
  bool sharded() const { return part_global != nullptr; }
  // sdpb_b200_cholesky_diagonals: descriptors [2J X | 2J Y | J S | Q] and the gathered diagonals
  DiagDesc *d_diag = nullptr;
  limb_t *diag_buf = nullptr;
  long diag_off[4] = {0, 0, 0, 0}, diag_total = 0; // element offsets of the four groups in diag_buf

  // Concurrent schedule.  The step is a small dependency graph -- chol(X) ->
  // L_X^-1 V -> A_X_inv | Y V -> A_Y | chol(Y) | per block: S_j -> chol(S_j) ->
  // L_j^-1 B_j -> norm partials -- and the
  // batched factorisations are chains of short level kernels with latency-bound
  // tails, so independent chains run on side streams and the S chain is cut
  // into G interleaved groups of blocks.  concurrency == 0 puts everything back
  // on `stream` in program order (what the per-kernel timeline is measured in).
  static constexpr int MAXG = 4;
  cudaStream_t cur = nullptr; // the stream the launch helpers enqueue on
  cudaStream_t copy = nullptr; // D2H of finished outputs while the step is still running (schur_step)
  cudaEvent_t evd[2] = {};     // [0] chol(Y) done
  cudaStream_t aux[MAXG] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t evf[20] = {};
  int concurrency = 1, G = 1;
  PotrfDesc *d_potrfS_g[MAXG] = {};
  TrsmTileDesc *d_trsmP_g[MAXG] = {};
  SchurDesc *d_schur_g[MAXG] = {};
  BandDesc *d_bands_g[MAXG] = {};
  std::vector<int> szS_g[MAXG], szP_g[MAXG], blocks_g[MAXG]; // blocks_g: local block indices of a group
  int nblk_g[MAXG] = {}, maxP_g[MAXG] = {};
  cudaStream_t side(int k) const { return concurrency ? aux[k] : stream; }
  // two streams of the greatest priority for the latency-bound chains (the pivot chains of chol(X)
  // and chol(Y) are a few warps per SM for milliseconds): their CTAs are placed ahead of the pending
  // CTAs of the throughput kernels running beside them, instead of queueing behind them
  cudaStream_t prio[2] = {nullptr, nullptr};
  // block-diagonal solves on the tile kernels (launch_nl.cu bdm_trsm_tiles): block-parities sorted
  // by size, their sizes, cumulative columns, and the descriptor arrays per (L, B, mode)
  std::vector<int> bdm_sorted, bdm_sizes, bdm_cum;
  std::map<std::tuple<const void *, const void *, int>, TrsmTileDesc *> bdm_trsm_descs;
  bool split_by_size = false; // the S chain runs as two size classes (group 0 = the large blocks)
  cudaStream_t urgent(int k) const { return concurrency && prio[k] ? prio[k] : stream; }
  // `to` waits for everything enqueued on `from` so far
  cudaError_t after(cudaStream_t from, cudaStream_t to, int e)
  {
    if(from == to)
      return cudaSuccess;
    cudaError_t r = cudaEventRecord(evf[e], from);
    return r != cudaSuccess ? r : cudaStreamWaitEvent(to, evf[e], 0);
  }

  bool have_X_cholesky = false, have_pairings = false;
  // solve_schur_complement_equation on the resident factors (solve.cuh)
  bool have_factors = false; // L_j, L_j^-1 B_j and chol(Q) of a successful step are in place
  SolveTriDesc *d_solveS = nullptr, *d_solveQ = nullptr; // S blocks largest first; the one Q system
  limb_t *sol_x = nullptr, *sol_y = nullptr;             // stacked dx (K elements), dy (N elements)
  uint64_t *sol_pinned = nullptr;                        // host staging, (K + N) elements
  float solve_ms = 0;                                    // device time of the last solve
  // scale_multiply_add (row N2): operands, product and result as block-diagonal objects
  limb_t *smaA = nullptr, *smaB = nullptr, *smaT = nullptr, *smaC = nullptr; // wXY words each
  GemmTileDesc *d_gemmSMA = nullptr;
  int tiles_SMA = 0;
  // search direction (row N2, direction.cuh): block-diagonal objects of the shape of X and the
  // vectors of compute_search_direction, resident between the calls of one iteration
  limb_t *dirMXY = nullptr, *dirR = nullptr, *dirZ = nullptr, *dirDX = nullptr, *dirDY = nullptr,
         *dirPR = nullptr;          // -XY, R, Z, dX, dY, primal residues: wXY words each
  limb_t *dir_dual = nullptr;       // dual residues, K elements (stacked like dx)
  limb_t *dir_prp = nullptr;        // primal_residue_p, N elements
  limb_t *dir_scal = nullptr;       // [0] beta mu  [1] 0.5 as mpf_set_d gives it  [2] mu  [3] 2^-(prec-16)
  limb_t *dir_part = nullptr;       // per block-parity scalars (traces, maxima, Frobenius products)
  limb_t *dir_colsum = nullptr;     // per column scratch of the Frobenius product
  BdmDesc *d_bdm = nullptr;         // block-parity b = 2j + parity, in that order
  int *d_row_block = nullptr;       // SDP block of every stacked row
  int bdm_cols = 0;                 // sum of the block sizes s
  GemmTileDesc *d_gemmXY = nullptr, *d_gemmDXDY = nullptr, *d_gemmPRY = nullptr, *d_gemmDXY = nullptr;
  int tiles_dir = 0;
  uint64_t *dir_pinned = nullptr;   // host staging: vectors up, per-block scalars down
  bool have_XY = false, have_minus_XY = false, have_residues = false, have_direction = false;
  float direction_ms = 0;
  // step_length (row N3, eig.cuh): the congruence and the tridiagonalisation work in dirZ (dead
  // after compute_search_direction); d, e, e^2 of the tridiagonal matrices stacked like the columns
  limb_t *eig_d = nullptr, *eig_e = nullptr, *eig_e2 = nullptr;
  int *eig_iter = nullptr; // Laguerre steps per block-parity of the last call
  float step_length_ms = 0;
  long launches = 0; // kernels launched since creation
  cudaEvent_t ev[12] = {}; // 0,1 pairings; 2..8 Schur stages; 9,10,11 resident step
  float stage_ms[9] = {0};

  // per-launch timeline of the last step: every kernel launch is bracketed by
  // two events on the launching stream (kt_begin / kt_end)
  struct KernelSpan
  {
    const char *name;
    cudaEvent_t e0, e1;
  };
  std::vector<KernelSpan> kt; // event pool, reused every step
  int kt_used = 0;
  int kt_begin(const char *name)
  {
    if(kt_used == (int)kt.size())
      {
        KernelSpan s{name, nullptr, nullptr};
        if(cudaEventCreate(&s.e0) != cudaSuccess
           || cudaEventCreate(&s.e1) != cudaSuccess)
          return -1;
        kt.push_back(s);
      }
    kt[kt_used].name = name;
    cudaEventRecord(kt[kt_used].e0, cur);
    return kt_used;
  }
  void kt_end()
  {
    cudaEventRecord(kt[kt_used].e1, cur);
    ++kt_used;
    ++launches;
  }
};


// per-precision kernel drivers, one table per stored-limb count NL
struct LaunchTable
{
  // a precision module built from other sources than the library that loads it would read the
  // context at wrong offsets: module_table (capi.cu) compares these two sizes before anything else
  size_t ctx_bytes, table_bytes;
  int (*cholesky)(sdpb_b200_ctx *, int which);
  int (*pairings)(sdpb_b200_ctx *, int part); // 0: X chain (L_X^-1 V, A_X_inv); 1: Y chain (Y V, A_Y)
  int (*schur_and_Q)(sdpb_b200_ctx *);
  int (*schur_solve)(sdpb_b200_ctx *); // solve.cuh, on sol_x / sol_y
  int (*scale_multiply_add)(sdpb_b200_ctx *, int alpha, int beta); // smaC = alpha smaA smaB + beta smaC
  int (*scalar)(sdpb_b200_ctx *, int op, int k, long count, const limb_t *a,
                const limb_t *b, limb_t *r);
  // direction.cuh on the resident objects.  op 0: -XY and its per-block traces; 1: per-block
  // max |-XY + mu I|; 2: compute_search_direction (arg: corrector phase); 3: per-block Frobenius
  // products of (X + dX, Y + dY)
  int (*direction)(sdpb_b200_ctx *, int op, int arg);
  // eig.cuh: per block-parity min eigenvalue of L^-1 dM L^-T into dir_part; which 0: X, dX; 1: Y, dY
  int (*step_length)(sdpb_b200_ctx *, int which);
};
// weak: a development build may compile only some precisions (make NLS="14")
#define F(n) extern "C" const LaunchTable sdpb_b200_launch_nl##n __attribute__((weak));
SDPB_FOR_EACH_NL(F)
#undef F

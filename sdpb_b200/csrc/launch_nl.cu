// Kernel drivers for ONE precision: compile with -DSDPB_NL=<stored limbs>.
// (sdpb_b200/csrc/Makefile builds one object per entry of SDPB_FOR_EACH_NL.)
#include "ctx.h"

#include <algorithm>
#include <climits>
#include <cstdlib>
#include <mutex>
#include <set>
#include <string>

#ifndef SDPB_NL
#error "compile with -DSDPB_NL=<n>"
#endif

// ---------------------------------------------------------- kernel drivers
template <int NL> struct Launch
{
  static constexpr size_t TILE_SMEM = sizeof(TileSmem<NL>);
  // Dynamic shared memory: every kernel is opted in to the full 227 KB of an sm_100 CTA, not to the
  // size of the launch at hand -- the attribute belongs to the (device, function) pair, and two
  // contexts of one process (one host thread per GPU, or the in-process communicator) setting it
  // to their own, different sizes race: "invalid argument" at the launch of the larger one.
  static constexpr int SMEM_OPT_IN = 227 * 1024;
  // opt in to > 48 KB of dynamic shared memory, once per kernel and device
  template <typename K> static int smem_opt_in(sdpb_b200_ctx *c, K kernel)
  {
    CUDA_TRY(c, cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN));
    return 0;
  }
  static constexpr size_t DIAG_SMEM = sizeof(DiagSmem<NL>);
  // "<label>/<part>": one timeline name per kind of launch inside a batched factorisation
  static const char *sub(const char *label, const char *part)
  {
    // shared by every context of the process (one host thread per GPU may be in here at once)
    static std::set<std::string> pool;
    static std::mutex guard;
    std::lock_guard<std::mutex> lock(guard);
    return pool.insert(std::string(label) + "/" + part).first->c_str();
  }
  // matrices of a batch (sorted, largest first) that still have tile index `t`
  static int alive(const std::vector<int> &sizes, int t)
  {
    int n = 0;
    while(n < (int)sizes.size() && sizes[n] > t * TS)
      ++n;
    return n;
  }
  // batched Cholesky, level-synchronous (tile.cuh); sizes sorted descending
  static int potrf(sdpb_b200_ctx *c, const char *label, const PotrfDesc *d,
                   const std::vector<int> &sizes, int *status, int nstatus, bool reset = true)
  {
    if(sizes.empty() || nstatus == 0)
      return 0;
    if(int rc = smem_opt_in(c, potrf_gemm_level<NL>))
      return rc;
    constexpr size_t WARP_SMEM = DIAG_WARPS * sizeof(WarpTileSmem<NL>);
    CUDA_TRY(c, cudaFuncSetAttribute(potrf_diag_warp<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN));
    if(int rc = smem_opt_in(c, potrf_panel_rl<NL>))
      return rc;
    if(reset)
      CUDA_TRY(c, cudaMemsetAsync(status, 0xFF, (size_t)nstatus * sizeof(int), c->cur));
    const int T = (sizes[0] + TS - 1) / TS;
    const char *l_gemm = sub(label, "gemm"), *l_diag = sub(label, "diag"), *l_panel = sub(label, "panel");
    for(int Jt = 0; Jt < T; ++Jt)
      {
        const int n = alive(sizes, Jt), nbelow = alive(sizes, Jt + 1);
        const int rows_from = sizes[0] - Jt * TS; // rows of the largest matrix in block column Jt
        if(Jt > 0) // a_ij -= sum_{k < J0} l_ik l_jk for every tile of the block column
          {
            dim3 g(n, (rows_from + TS - 1) / TS);
            c->kt_begin(l_gemm);
            potrf_gemm_level<NL><<<g, 256, TILE_SMEM, c->cur>>>(d, Jt, status);
            c->kt_end();
          }
        c->kt_begin(l_diag);
        potrf_diag_warp<NL><<<(n + DIAG_WARPS - 1) / DIAG_WARPS, 32 * DIAG_WARPS, WARP_SMEM, c->cur>>>(
          d, n, Jt, status);
        c->kt_end();
        if(nbelow == 0)
          continue;
        // tiles below the diagonal one: X = A_tile L_JJ^{-T}, one CTA per tile, the 16 steps of a
        // row shared by 16 threads (0.1 ms of latency instead of 0.45 for one thread per row)
        const int rows_below = rows_from - TS;
        dim3 g2(nbelow, (rows_below + TS - 1) / TS);
        c->kt_begin(l_panel);
        potrf_panel_rl<NL><<<g2, 256, TILE_SMEM, c->cur>>>(d, Jt, status);
        c->kt_end();
      }
    CUDA_TRY(c, cudaGetLastError());
    return 0;
  }
  // right-looking schedule (tile.cuh): the choice for a single large matrix
  static int potrf_rl(sdpb_b200_ctx *c, const char *label, const PotrfDesc *d,
                      const std::vector<int> &sizes, int *status, int nstatus)
  {
    if(sizes.empty() || nstatus == 0)
      return 0;
    if(int rc = smem_opt_in(c, potrf_diag_rl<NL>))
      return rc;
    if(int rc = smem_opt_in(c, potrf_panel_rl<NL>))
      return rc;
    if(int rc = smem_opt_in(c, potrf_trail_rl<NL>))
      return rc;
    CUDA_TRY(c, cudaMemsetAsync(status, 0xFF, (size_t)nstatus * sizeof(int), c->cur));
    const int T = (sizes[0] + TS - 1) / TS;
    // one matrix, several ranks, N large enough: block columns dealt round-robin over the ranks
    const bool dist = c->world > 1 && sizes.size() == 1 && sizes[0] >= c->qdist_min_N && c->bcast;
    const int mod = dist ? c->world : 1, me = dist ? c->rank : 0;
    const char *l_diag = sub(label, "diag"), *l_panel = sub(label, "panel"), *l_trail = sub(label, "trail");
    const char *l_fused = sub(label, "diag+panel");
    // one matrix on one rank: the pipelined diagonal + panel kernel, as long as its grid is
    // co-resident (the panel CTAs wait for CTA 0's columns).  c->d_flags[3] is its progress word,
    // zero since the start of the step (init_flags).  SDPB_B200_POTRF_FUSED=0: the two kernels.
    bool fused = !dist && sizes.size() == 1 && c->d_flags;
    if(const char *env = getenv("SDPB_B200_POTRF_FUSED"))
      fused = fused && atoi(env) != 0;
    int resident = 0;
    if(fused)
      {
        if(int rc = smem_opt_in(c, potrf_diag_panel_rl<NL>))
          return rc;
        int per_sm = 0, sms = 0;
        CUDA_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, potrf_diag_panel_rl<NL>, 256, TILE_SMEM));
        CUDA_TRY(c, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
        resident = per_sm * sms;
        CUDA_TRY(c, cudaMemsetAsync(c->d_flags + 3, 0, sizeof(int), c->cur));
      }
    for(int Jt = 0; Jt < T; ++Jt)
      {
        const int n = alive(sizes, Jt), nbelow = alive(sizes, Jt + 1);
        const int tb = (sizes[0] - (Jt + 1) * TS + TS - 1) / TS; // tiles below / right of Jt
        const bool mine = Jt % mod == me;
        if(mine && fused && nbelow && 1 + tb <= resident)
          {
            // diagonal tile and panel in one launch, the panel one column behind (tile.cuh)
            c->kt_begin(l_fused);
            potrf_diag_panel_rl<NL><<<1 + tb, 256, TILE_SMEM, c->cur>>>(d, Jt, status, c->d_flags + 3);
            c->kt_end();
          }
        else if(mine)
          {
            c->kt_begin(l_diag);
            potrf_diag_rl<NL><<<n, 256, TILE_SMEM, c->cur>>>(d, Jt, status);
            c->kt_end();
            if(nbelow)
              {
                c->kt_begin(l_panel);
                potrf_panel_rl<NL><<<dim3(nbelow, tb), 256, TILE_SMEM, c->cur>>>(d, Jt, status);
                c->kt_end();
              }
          }
        if(dist)
          {
            // the finished block column (and the owner's status) to every rank
            const int rows = sizes[0] - Jt * TS;
            const size_t bytes = ((size_t)sizes[0] * TS * Fmt<NL>::ES + 2) * 8 + (size_t)TS * TileGeom<NL>::RS * 4;
            const int pg = std::min(592, (rows * TS + 127) / 128);
            if(mine)
              {
                c->kt_begin("panel_pack");
                panel_pack<NL><<<pg, 128, 0, c->cur>>>(d, Jt, c->qpanel, status, 0);
                c->kt_end();
              }
            if(int rc = c->bcast(c, c->qpanel, bytes, Jt % mod, "nccl_broadcast_Q_panel"))
              return rc;
            if(!mine)
              {
                c->kt_begin("panel_pack");
                panel_pack<NL><<<pg, 128, 0, c->cur>>>(d, Jt, c->qpanel, status, 1);
                c->kt_end();
              }
          }
        if(nbelow == 0)
          continue;
        c->kt_begin(l_trail);
        potrf_trail_rl<NL><<<dim3(nbelow, tb * (tb + 1) / 2), 256, TILE_SMEM, c->cur>>>(d, Jt, status, mod, me);
        c->kt_end();
      }
    CUDA_TRY(c, cudaGetLastError());
    return 0;
  }
  // batched X <- L^{-1} B; sizes sorted descending.  Two schedules (same arithmetic):
  //   levels  per 16-row tile one launch of trsm_gemm_level2 (register tiles, (8 x 32) on a ragged
  //           last tile) and one of trsm_diag_level (one thread per column) -- the default;
  //   walk    trsm_walk_kernel, the whole solve of a group of columns in one CTA and one launch.
  //           Measured at c3 (profiles/r02_trsm_walk.md): slower.  A CTA that walks a 120-row block
  //           alone is a chain of 7000 dependent multiply-accumulates (32 ms against 29 ms for the
  //           whole batch), and even the 40-row blocks lose: 900 equal CTAs on 444 slots are 2.03
  //           waves, and the diagonal phase waits for its operand loads with four warps per
  //           scheduler.  SDPB_B200_TRSM=walk runs it everywhere, =hybrid on matrices of up to
  //           WALK_MAX_TILES row tiles beside the level kernels for the taller ones, =levels1 the
  //           round-1 update kernel (16 x 16 tiles only).
  static constexpr int WALK_MAX_TILES = 3;
  static int trsm(sdpb_b200_ctx *c, const char *label, const TrsmTileDesc *d,
                  const std::vector<int> &sizes, int maxcols, cudaStream_t side, int ev0,
                  bool levels_only = false)
  {
    if(sizes.empty() || sizes[0] == 0 || maxcols == 0)
      return 0;
    static const int env_mode = [] {
      const char *env = getenv("SDPB_B200_TRSM");
      const std::string m = env ? env : "levels";
      return m == "walk" ? 2 : m == "hybrid" ? 0 : m == "levels1" ? 3 : 1;
    }();
    // (descriptors with general strides -- bdm_trsm_tiles -- are for the level kernels only)
    const int mode = levels_only && env_mode != 3 ? 1 : env_mode;
    int nheavy = 0; // prefix of the (descending) batch that keeps the level kernels
    while(nheavy < (int)sizes.size()
          && (mode == 1 || mode == 3 || (mode == 0 && sizes[nheavy] > WALK_MAX_TILES * TS)))
      ++nheavy;
    const int nlight = (int)sizes.size() - nheavy;
    cudaStream_t main_stream = c->cur;
    if(nlight)
      {
        // block = 16 TC threads owning ~16 TC columns: the largest TC whose padding of the
        // column count stays within 3 % of the best one
        int best_waste = INT_MAX;
        for(int tc = 4; tc <= 16; ++tc)
          best_waste = std::min(best_waste, ((maxcols + 16 * tc - 1) / (16 * tc)) * 16 * tc - maxcols);
        int TC = 4;
        for(int tc = 4; tc <= 16; ++tc)
          if(((maxcols + 16 * tc - 1) / (16 * tc)) * 16 * tc - maxcols <= best_waste + maxcols * 3 / 100)
            TC = tc;
        while(TC > 4 && WalkGeom<NL>::bytes(TC) > 227 * 1024)
          --TC;
        const int ncg = (maxcols + 16 * TC - 1) / (16 * TC);
        const int Wc = (maxcols + ncg - 1) / ncg; // columns per CTA, balanced
        const size_t smem = WalkGeom<NL>::bytes(TC);
        CUDA_TRY(c, cudaFuncSetAttribute(trsm_walk_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN));
        if(nheavy)
          {
            CUDA_TRY(c, c->after(main_stream, side, ev0));
            c->cur = side;
          }
        c->kt_begin(sub(label, "walk"));
        trsm_walk_kernel<NL><<<(unsigned)nlight * ncg, 16 * TC, smem, c->cur>>>(d + nheavy, ncg, Wc);
        c->kt_end();
        c->cur = main_stream;
        CUDA_TRY(c, cudaGetLastError());
      }
    if(nheavy == 0)
      return 0;
    if(int rc = smem_opt_in(c, trsm_gemm_level<NL>))
      return rc;
    const size_t smem2 = WalkGeom<NL>::bytes(TS);
    CUDA_TRY(c, cudaFuncSetAttribute(trsm_gemm_level2<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN));
    CUDA_TRY(c, cudaFuncSetAttribute(trsm_diag_level<NL>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN));
    CUDA_TRY(c, cudaFuncSetAttribute(trsm_diag_tile<NL>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN));
    // diagonal solves of a level: one thread per column from this many columns (all matrices of the
    // level together) on, the 16 x 16 form (trsm_diag_tile) below it
    const char *tile_env = getenv("SDPB_B200_TRSM_TILE_BELOW"); // (read per call: the tests switch it)
    const long tile_below = tile_env ? atol(tile_env) : 32768L;
    const std::vector<int> heavy(sizes.begin(), sizes.begin() + nheavy);
    const int T = (heavy[0] + TS - 1) / TS;
    const char *l_gemm = sub(label, "gemm"), *l_diag = sub(label, "diag");
    for(int It = 0; It < T; ++It)
      {
        const int n = alive(heavy, It);
        if(It > 0)
          {
            dim3 g(n, (maxcols + TS - 1) / TS);
            c->kt_begin(l_gemm);
            if(mode == 3)
              trsm_gemm_level<NL><<<g, 256, TILE_SMEM, c->cur>>>(d, It);
            else
              trsm_gemm_level2<NL><<<g, 256, smem2, c->cur>>>(d, It);
            c->kt_end();
          }
        c->kt_begin(l_diag);
        if((long)n * maxcols <= tile_below)
          {
            // too few columns to fill the SMs with one thread each: 16 x 16 threads per tile
            dim3 g2(n, (maxcols + TS - 1) / TS);
            trsm_diag_tile<NL><<<g2, 256, sizeof(DiagTileSmem<NL>), c->cur>>>(d, It);
          }
        else
          {
            dim3 g2(n, (maxcols + ROWS_PER_CTA - 1) / ROWS_PER_CTA);
            trsm_diag_level<NL><<<g2, ROWS_PER_CTA, DIAG_SMEM, c->cur>>>(d, It);
          }
        c->kt_end();
      }
    CUDA_TRY(c, cudaGetLastError());
    if(nlight)
      CUDA_TRY(c, c->after(side, main_stream, ev0 + 1));
    return 0;
  }
  static int gemm(sdpb_b200_ctx *c, const char *label, const GemmTileDesc *d,
                  int count, int tiles)
  {
    if(count == 0 || tiles == 0)
      return 0;
    if(int rc = smem_opt_in(c, gemm_tile_kernel<NL>))
      return rc;
    c->kt_begin(label);
    gemm_tile_kernel<NL><<<tiles, 256, TILE_SMEM, c->cur>>>(d, count);
    c->kt_end();
    CUDA_TRY(c, cudaGetLastError());
    return 0;
  }
  static int cholesky(sdpb_b200_ctx *c, int which)
  {
    return potrf(c, which == 0 ? "potrf_X" : "potrf_Y", which == 0 ? c->d_potrfX : c->d_potrfY,
                 c->szXY, c->d_status + which * 2 * c->J, 2 * c->J);
  }
  static int pairings(sdpb_b200_ctx *c, int part)
  {
    const int nb = 2 * c->J;
    if(nb == 0)
      return 0;
    if(part == 0)
      {
        // T = V ; T <- L_X^{-1} T ; AX = T^T T
        CUDA_TRY(c, cudaMemcpyAsync(c->T, c->V, c->wV * 8, cudaMemcpyDeviceToDevice, c->cur));
        int rc = trsm(c, "trsm_LXinv_V", c->d_trsmT, c->szT, c->max_mn, c->side(2), 14);
        if(rc)
          return rc;
        return gemm(c, "gemm_A_X_inv", c->d_gemmAX, c->n_gemm, c->tiles_AX);
      }
    int rc = gemm(c, "gemm_YV", c->d_gemmYV, c->n_gemm, c->tiles_YV);
    if(rc)
      return rc;
    return gemm(c, "gemm_A_Y", c->d_gemmAY, c->n_gemm, c->tiles_AY);
  }
  static int schur_and_Q(sdpb_b200_ctx *c)
  {
    const int J = c->J, N = c->N;
    cudaStream_t st = c->stream;
    c->cur = st;
    CUDA_TRY(c, cudaEventRecord(c->ev[2], st));
    if(J == 0)
      CUDA_TRY(c, cudaEventRecord(c->ev[3], st));
    const bool sharded = c->sharded();
    limb_t *part = sharded ? c->part_global : c->part;
    const int Jsum = sharded ? c->J_global : J;
    if(J)
      {
        CUDA_TRY(c, cudaMemsetAsync(c->d_status + 4 * J, 0xFF, (size_t)J * sizeof(int), st));
        CUDA_TRY(c, cudaMemcpyAsync(c->Pband, c->B, c->wB * 8, cudaMemcpyDeviceToDevice, st));
      }
    if(sharded) // rows of blocks owned elsewhere: exact zeros, so the sum below is a gather
      CUDA_TRY(c, cudaMemsetAsync(part, 0, (size_t)Jsum * N * Fmt<NL>::ES * 8, st));
    // per group of blocks: S_j -> Cholesky(S_j) -> P_j = L_j^{-1} B_j -> column-norm partials
    for(int g = 0; g < c->G && J; ++g)
      {
        // two size classes: the large blocks' chain (long pivot chains on few matrices) goes on a
        // stream of the greatest priority, the small blocks' kernels fill the SMs under it
        cudaStream_t sg = g == 0 ? (c->split_by_size ? c->urgent(0) : st) : c->side(g);
        CUDA_TRY(c, c->after(st, sg, 4 + g));
        c->cur = sg;
        const int nb = c->nblk_g[g], mp = c->maxP_g[g];
        dim3 grid(nb, (unsigned)std::min<long>(((long)mp * mp + 127) / 128, 65535));
        c->kt_begin("schur_kernel");
        schur_kernel<NL><<<grid, 128, 0, sg>>>(c->d_schur_g[g]);
        c->kt_end();
        CUDA_TRY(c, cudaGetLastError());
        if(g == 0) // stage boundary "S assembled" (of the first group when the chain is cut into groups)
          CUDA_TRY(c, cudaEventRecord(c->ev[3], sg));
        int rc = potrf(c, "potrf_S", c->d_potrfS_g[g], c->szS_g[g], c->d_status + 4 * J, J, false);
        if(rc)
          return rc;
        rc = trsm(c, "trsm_Linv_B", c->d_trsmP_g[g], c->szP_g[g], N, c->G == 1 ? c->side(1) : sg, 12);
        if(rc)
          return rc;
        dim3 g1(nb, (N + 63) / 64);
        c->kt_begin("norm_partial_kernel");
        norm_partial_kernel<NL><<<g1, 64, 0, sg>>>(c->d_bands_g[g], N, part);
        c->kt_end();
        CUDA_TRY(c, cudaGetLastError());
      }
    c->cur = st;
    if(c->split_by_size && J)
      CUDA_TRY(c, c->after(c->urgent(0), st, 8));
    for(int g = 1; g < c->G && J; ++g)
      CUDA_TRY(c, c->after(c->side(g), st, 8 + g));
    CUDA_TRY(c, cudaEventRecord(c->ev[4], st));
    // norms, normalise, residues
    const int init_flags[4] = {0, INT_MAX, 0, 0};
    CUDA_TRY(c, cudaMemcpyAsync(c->d_flags, init_flags, sizeof(init_flags),
                                cudaMemcpyHostToDevice, st));
    if(sharded)
      if(int rc2 = c->allreduce(c, part, (size_t)Jsum * N * Fmt<NL>::ES, 1, "nccl_allreduce_norm_partials"))
        return rc2;
    // every rank adds the per-block partials in GLOBAL block order: the norms do
    // not depend on how the blocks are sharded
    c->kt_begin("norm_final_kernel");
    {
      const size_t sm = ((sizeof(coop::Work<NL>) + 15) & ~(size_t)15) + 33 * TileGeom<NL>::SW * 4;
      CUDA_TRY(c, cudaFuncSetAttribute(norm_final_kernel<NL>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN));
      norm_final_kernel<NL><<<N, 32, sm, st>>>(part, Jsum, N, c->norms, c->recipN);
    }
    c->kt_end();
    CUDA_TRY(c, cudaGetLastError());
    if(J)
      {
        dim3 g2(J, (unsigned)std::min<long>(((long)c->max_P * N + 127) / 128, 65535));
        c->kt_begin("normalize_kernel");
        const size_t nsmem = ((size_t)c->crt.np * NormGeom<NL>::NDP + ((c->crt.np + 1) & ~1)) * 4
                             + (size_t)c->crt.np * 8;
        CUDA_TRY(c, cudaFuncSetAttribute(normalize_kernel<NL>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN));
        normalize_kernel<NL><<<g2, 128, nsmem, st>>>(c->d_bands, N, c->NS, c->KR, c->norms,
                                                     c->recipN, c->prec, c->crt, c->R, c->d_flags);
        c->kt_end();
    CUDA_TRY(c, cudaGetLastError());
      }
    CUDA_TRY(c, cudaEventRecord(c->ev[5], st));
    if(c->syrk_imma && c->KR)
      {
        // byte-slice products on the integer tensor path (syrk_imma.cuh): re-pack the planes in
        // place, then one CTA per 64 x 64 tile pair and prime
        const long groups = (long)c->crt.np * c->KR / 4;
        c->kt_begin("syrk_pack_kernel");
        syrk_pack_kernel<4><<<(unsigned)std::min<long>((groups * c->NS + 255) / 256, 148L * 32), 256, 0, st>>>(
          c->R, groups, c->K, c->KR, c->NS);
        c->kt_end();
        CUDA_TRY(c, cudaGetLastError());
        CUDA_TRY(c, cudaFuncSetAttribute(syrk_imma_kernel<SI_STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         SMEM_OPT_IN));
        const int nt = (N + SI_TILE - 1) / SI_TILE;
        dim3 g3(nt * (nt + 1) / 2, c->crt.np);
        c->kt_begin("syrk_imma_kernel");
        syrk_imma_kernel<SI_STAGES><<<g3, 256, SI_SMEM, st>>>(c->R, c->KR, N, c->NS, c->d_primes, c->d_inv64, c->Qres);
        c->kt_end();
        CUDA_TRY(c, cudaGetLastError());
      }
    else
      {
        const int nt = (N + 15) / 16;
        dim3 g3(nt * (nt + 1) / 2, c->crt.np);
        c->kt_begin("syrk_mod_kernel");
        syrk_mod_kernel<4><<<g3, 256, 0, st>>>(c->R, c->KR, N, c->NS, c->d_primes, c->Qres);
        c->kt_end();
        CUDA_TRY(c, cudaGetLastError());
      }
    // exact integers: the cross-GPU sum is order-free.  Residues are < 2^28, so a
    // plain u32 sum over <= 16 ranks cannot overflow; the CRT kernel reduces mod p.
    if(sharded)
      if(int rc2 = c->allreduce(c, c->Qres, (size_t)c->crt.np * N * N, 0, "nccl_allreduce_Q_residues"))
        return rc2;
    CUDA_TRY(c, cudaEventRecord(c->ev[6], st));
    {
      const long tot = (long)N * N;
      c->kt_begin("crt_restore_kernel");
      crt_restore_kernel<NL><<<(unsigned)((tot + 63) / 64), 64, 0, st>>>(
        c->Qres, N, c->prec, c->crt, c->norms, c->Q, c->d_flags);
      c->kt_end();
    CUDA_TRY(c, cudaGetLastError());
    }
    CUDA_TRY(c, cudaEventRecord(c->ev[7], st));
    int rc = potrf_rl(c, "potrf_Q", c->d_potrfQ, c->szQ, c->d_status + 5 * J, 1);
    if(rc)
      return rc;
    CUDA_TRY(c, cudaEventRecord(c->ev[8], st));
    if(sharded)
      {
        // a rank whose blocks failed (non-positive pivot, overflow, bad Q diagonal) has fed garbage
        // into the exchanges: every rank must report the step as failed, not only that one
        c->kt_begin("fail_flag_kernel");
        fail_flag_kernel<<<1, 128, 0, st>>>(c->d_status, 5 * J + 1, c->d_flags, c->d_fail);
        c->kt_end();
        CUDA_TRY(c, cudaGetLastError());
        if(int rc2 = c->allreduce(c, c->d_fail, 1, 0, "nccl_allreduce_failure_flag"))
          return rc2;
      }
    return 0;
  }
  // one CTA per triangular system; unknowns in shared memory when the largest system fits
  template <bool BACK>
  static int solve_tri(sdpb_b200_ctx *c, const char *label, const SolveTriDesc *d, int count, int maxp,
                       limb_t *x)
  {
    if(count == 0 || maxp == 0)
      return 0;
    const size_t need = (size_t)maxp * TileGeom<NL>::SW * 4;
    // the unknowns stay in shared memory up to 200 KB (1422 rows at 768 bits); larger systems
    // substitute on the global vector itself.  SDPB_B200_SOLVE_SMEM_MAX (bytes) moves the limit
    // (tests force the global path with 0).
    size_t smem_max = 200 * 1024;
    if(const char *env = getenv("SDPB_B200_SOLVE_SMEM_MAX"))
      smem_max = std::min<size_t>(smem_max, (size_t)atol(env));
    const int use_smem = need <= smem_max;
    const size_t smem = use_smem ? need : 0;
    CUDA_TRY(c, cudaFuncSetAttribute(solve_tri_kernel<NL, BACK>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN));
    const int threads = std::min(SOLVE_MAX_THREADS, std::max(32, (maxp + 31) & ~31));
    c->kt_begin(label);
    solve_tri_kernel<NL, BACK><<<count, threads, smem, c->cur>>>(d, x, use_smem);
    c->kt_end();
    CUDA_TRY(c, cudaGetLastError());
    return 0;
  }
  // solve_schur_complement_equation.cxx:16-79 on sol_x (stacked dx) and sol_y (dy)
  static int schur_solve(sdpb_b200_ctx *c)
  {
    const int J = c->J, N = c->N;
    cudaStream_t st = c->stream;
    c->cur = st;
    const bool sharded = c->sharded();
    limb_t *part = sharded ? c->part_global : c->part;
    const int Jsum = sharded ? c->J_global : J;
    // dx_j <- L_j^-1 dx_j
    if(int rc = solve_tri<false>(c, "solve_Linv_dx", c->d_solveS, J, c->max_P, c->sol_x))
      return rc;
    // dy -= sum_j P_j^T dx_j: per-block partial rows, added in GLOBAL block order
    if(sharded)
      CUDA_TRY(c, cudaMemsetAsync(part, 0, (size_t)Jsum * N * Fmt<NL>::ES * 8, st));
    if(J)
      {
        dim3 g(J, (N + 63) / 64);
        c->kt_begin("solve_gemvT_kernel");
        solve_gemvT_kernel<NL><<<g, 64, 0, st>>>(c->d_bands, N, c->sol_x, part);
        c->kt_end();
        CUDA_TRY(c, cudaGetLastError());
      }
    if(sharded)
      if(int rc = c->allreduce(c, part, (size_t)Jsum * N * Fmt<NL>::ES, 1, "nccl_allreduce_dy_partials"))
        return rc;
    c->kt_begin("solve_dysum_kernel");
    solve_dysum_kernel<NL><<<N, 32, 32 * TileGeom<NL>::SW * 4, st>>>(part, Jsum, N, c->sol_y);
    c->kt_end();
    CUDA_TRY(c, cudaGetLastError());
    // dy <- U^-1 U^-T dy
    if(int rc = solve_tri<false>(c, "solve_Q_forward", c->d_solveQ, 1, N, c->sol_y))
      return rc;
    if(int rc = solve_tri<true>(c, "solve_Q_backward", c->d_solveQ, 1, N, c->sol_y))
      return rc;
    // dx_j += P_j dy ; dx_j <- L_j^-T dx_j
    if(J)
      {
        dim3 g(J, (c->max_P + 63) / 64);
        c->kt_begin("solve_gemv_kernel");
        solve_gemv_kernel<NL><<<g, 64, 0, st>>>(c->d_bands, N, c->sol_y, c->sol_x);
        c->kt_end();
        CUDA_TRY(c, cudaGetLastError());
      }
    return solve_tri<true>(c, "solve_LinvT_dx", c->d_solveS, J, c->max_P, c->sol_x);
  }
  // scale_multiply_add.cxx:4-16 on the resident operands smaA, smaB (-> smaT) and smaC
  static int scale_multiply_add(sdpb_b200_ctx *c, int alpha, int beta)
  {
    c->cur = c->stream;
    if(int rc = gemm(c, "gemm_scale_multiply_add", c->d_gemmSMA, c->n_gemm, c->tiles_SMA))
      return rc;
    const long count = (long)(c->wXY / Fmt<NL>::ES);
    if(count == 0)
      return 0;
    c->kt_begin("sma_epilogue_kernel");
    sma_epilogue_kernel<NL><<<(unsigned)std::min<long>((count + 127) / 128, 148 * 16), 128, 0, c->cur>>>(
      c->smaT, c->smaC, count, alpha, beta);
    c->kt_end();
    CUDA_TRY(c, cudaGetLastError());
    return 0;
  }
  // C = alpha A B + beta C on block-diagonal objects: gemm_tile_kernel into smaT, then the epilogue
  static int bdm_product(sdpb_b200_ctx *c, const char *label, const GemmTileDesc *d, int alpha, int beta, limb_t *C)
  {
    if(int rc = gemm(c, label, d, c->n_gemm, c->tiles_dir))
      return rc;
    const long count = (long)(c->wXY / Fmt<NL>::ES);
    if(count == 0)
      return 0;
    c->kt_begin("sma_epilogue_kernel");
    sma_epilogue_kernel<NL><<<(unsigned)std::min<long>((count + 127) / 128, 148 * 16), 128, 0, c->cur>>>(
      c->smaT, C, count, alpha, beta);
    c->kt_end();
    CUDA_TRY(c, cudaGetLastError());
    return 0;
  }
  static int bdm_elementwise(sdpb_b200_ctx *c, int op, const limb_t *A, limb_t *C)
  {
    const long count = (long)(c->wXY / Fmt<NL>::ES);
    if(count == 0)
      return 0;
    c->kt_begin("bdm_elementwise_kernel");
    bdm_elementwise_kernel<NL><<<(unsigned)std::min<long>((count + 127) / 128, 148 * 16), 128, 0, c->cur>>>(op, A, C, count);
    c->kt_end();
    CUDA_TRY(c, cudaGetLastError());
    return 0;
  }
  // B <- L^-1 B (mode 0), L^-T B (1), B L^-T (2) on every block-parity: right-looking, one CTA per
  // block (direction.cuh); blocks too large for the shared-memory line buffer take the
  // one-thread-per-line kernels
  // The same three solves on the tile kernels of the Schur-complement step (trsm_gemm_level2 +
  // trsm_diag_level / trsm_diag_tile) through descriptors with general strides:
  //   mode 0  B <- L^-1 B   the standard view
  //   mode 1  B <- L^-T B   both index ranges reversed, i' = s-1-i: M(i', k') = L(k, i) is lower
  //                         triangular, the substitution runs forward in i', i.e. k descending
  //   mode 2  B <- B L^-T   B read by rows: unknown u of line l at B(l, u)
  // Every element receives the same operations in the same order as in bdm_trsm_rl_kernel (updates
  // in the canonical order of the unknowns, then the division), so the bits are the same; the work
  // runs as 16 x 16 register tiles with TMA-staged operands instead of one element per thread and
  // step through global memory.  Descriptor arrays are built on first use per (L, B, mode).
  static int bdm_trsm_tiles(sdpb_b200_ctx *c, int mode, const limb_t *L, const uint32_t *recip, limb_t *B)
  {
    typedef TileGeom<NL> G;
    constexpr int ES = Fmt<NL>::ES;
    const int nb = 2 * c->J;
    if(c->bdm_cols == 0)
      return 0;
    if(c->bdm_sorted.empty())
      {
        std::vector<int> cum(nb + 1, 0);
        for(int q = 0; q < nb; ++q)
          cum[q + 1] = cum[q] + c->g[q / 2].s[q % 2];
        c->bdm_sorted.resize(nb);
        for(int q = 0; q < nb; ++q)
          c->bdm_sorted[q] = q;
        std::stable_sort(c->bdm_sorted.begin(), c->bdm_sorted.end(),
                         [&](int a, int b) { return c->g[a / 2].s[a % 2] > c->g[b / 2].s[b % 2]; });
        c->bdm_sizes.clear();
        for(int q : c->bdm_sorted)
          if(c->g[q / 2].s[q % 2] > 0)
            c->bdm_sizes.push_back(c->g[q / 2].s[q % 2]);
        c->bdm_cum = cum;
      }
    if(c->bdm_sizes.empty())
      return 0;
    const auto key = std::make_tuple((const void *)L, (const void *)B, mode);
    auto it = c->bdm_trsm_descs.find(key);
    if(it == c->bdm_trsm_descs.end())
      {
        std::vector<TrsmTileDesc> v;
        for(int q : c->bdm_sorted)
          {
            const long s = c->g[q / 2].s[q % 2];
            if(s == 0)
              continue;
            const uint64_t *Lb = L + c->oXY[q];
            uint64_t *Bb = B + c->oXY[q];
            const uint32_t *rc = recip + (long)c->bdm_cum[q] * G::RS;
            TrsmTileDesc d{Lb, rc, Bb, (int)s, (int)s, 0, 0};
            if(mode == 1)
              {
                d.L = Lb + ((s - 1) * s + (s - 1)) * ES;
                d.lsi = -s;
                d.lsk = -1;
                d.B = Bb + (s - 1) * ES;
                d.bsi = -1;
                d.bsc = s;
                d.recip = rc + (s - 1) * G::RS;
                d.rstep = -1;
              }
            else if(mode == 2)
              {
                d.bsi = s;
                d.bsc = 1;
              }
            v.push_back(d);
          }
        TrsmTileDesc *dev = nullptr;
        CUDA_TRY(c, cudaMalloc(&dev, v.size() * sizeof(TrsmTileDesc)));
        CUDA_TRY(c, cudaMemcpy(dev, v.data(), v.size() * sizeof(TrsmTileDesc), cudaMemcpyHostToDevice));
        it = c->bdm_trsm_descs.emplace(key, dev).first;
      }
    static const char *labels[3] = {"bdm_trsm_Linv", "bdm_trsm_LinvT", "bdm_trsm_right_LinvT"};
    return trsm(c, labels[mode], it->second, c->bdm_sizes, c->max_s, c->cur, 0, true);
  }
  static int bdm_trsm_rl(sdpb_b200_ctx *c, int mode, const limb_t *L, const uint32_t *recip, limb_t *B)
  {
    const int nb = 2 * c->J;
    if(c->bdm_cols == 0)
      return 0;
    {
      // "tiles": the level kernels of the step through strided descriptors (bdm_trsm_tiles).  Bit
      // for bit the same, but measured slower at c3 (2.9 against 2.1 ms per solve: blocks of 20 and
      // 40 rows pad to 32 and 48 in 16 x 16 tiles and fill 20 of 32 lanes in the diagonal solves)
      const char *env = getenv("SDPB_B200_BDM_TRSM");
      if(env && std::string(env) == "tiles")
        return bdm_trsm_tiles(c, mode, L, recip, B);
    }
    const size_t smem = (size_t)c->max_s * TileGeom<NL>::SW * 4;
    if(smem > (size_t)SMEM_OPT_IN)
      {
        const unsigned g = (unsigned)((c->bdm_cols + 63) / 64);
        c->kt_begin("bdm_trsm_kernel");
        if(mode == 0)
          bdm_trsm_kernel<NL, false><<<g, 64, 0, c->cur>>>(c->d_bdm, nb, c->bdm_cols, L, recip, B);
        else if(mode == 1)
          bdm_trsm_kernel<NL, true><<<g, 64, 0, c->cur>>>(c->d_bdm, nb, c->bdm_cols, L, recip, B);
        else
          eig_trsm_kernel<NL, true><<<g, 64, 0, c->cur>>>(c->d_bdm, nb, c->bdm_cols, L, recip, B);
        c->kt_end();
        CUDA_TRY(c, cudaGetLastError());
        return 0;
      }
    c->kt_begin("bdm_trsm_rl_kernel");
    if(mode == 0)
      {
        CUDA_TRY(c, cudaFuncSetAttribute(bdm_trsm_rl_kernel<NL, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN));
        bdm_trsm_rl_kernel<NL, 0><<<nb, BDM_RL_THREADS, smem, c->cur>>>(c->d_bdm, L, recip, B);
      }
    else if(mode == 1)
      {
        CUDA_TRY(c, cudaFuncSetAttribute(bdm_trsm_rl_kernel<NL, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN));
        bdm_trsm_rl_kernel<NL, 1><<<nb, BDM_RL_THREADS, smem, c->cur>>>(c->d_bdm, L, recip, B);
      }
    else
      {
        CUDA_TRY(c, cudaFuncSetAttribute(bdm_trsm_rl_kernel<NL, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_OPT_IN));
        bdm_trsm_rl_kernel<NL, 2><<<nb, BDM_RL_THREADS, smem, c->cur>>>(c->d_bdm, L, recip, B);
      }
    c->kt_end();
    CUDA_TRY(c, cudaGetLastError());
    return 0;
  }
  // cholesky_solve.cxx:4-13 with the resident factor of X, then symmetrize (and negate)
  static int bdm_cholesky_solve_symmetrize(sdpb_b200_ctx *c, limb_t *Z, int negate)
  {
    const int nb = 2 * c->J;
    if(c->bdm_cols == 0)
      return 0;
    if(int rc = bdm_trsm_rl(c, 0, c->X, c->recipX, Z)) // L^-1 Z
      return rc;
    if(int rc = bdm_trsm_rl(c, 1, c->X, c->recipX, Z)) // L^-T (.)
      return rc;
    const int ms = c->max_s;
    dim3 gs(nb, (unsigned)std::min<long>(((long)ms * (ms + 1) / 2 + 127) / 128, 65535));
    c->kt_begin("bdm_symmetrize_kernel");
    bdm_symmetrize_kernel<NL><<<gs, 128, 0, c->cur>>>(c->d_bdm, nb, Z, c->dir_scal + Fmt<NL>::ES, negate);
    c->kt_end();
    CUDA_TRY(c, cudaGetLastError());
    return 0;
  }
  static int direction(sdpb_b200_ctx *c, int op, int arg)
  {
    const int J = c->J, nb = 2 * J, N = c->N;
    cudaStream_t st = c->stream;
    c->cur = st;
    constexpr int ES = Fmt<NL>::ES;
    if(op == 0) // minus_XY = -X Y (step.cxx:137) and the trace of every block
      {
        if(int rc = bdm_product(c, "gemm_minus_XY", c->d_gemmXY, -1, 0, c->dirMXY))
          return rc;
        if(nb)
          {
            c->kt_begin("bdm_trace_kernel");
            bdm_trace_kernel<NL><<<(nb + 63) / 64, 64, 0, st>>>(c->d_bdm, nb, c->dirMXY, c->dir_part);
            c->kt_end();
            CUDA_TRY(c, cudaGetLastError());
          }
        return 0;
      }
    if(op == 1) // compute_R_error.hxx, per block
      {
        if(nb)
          {
            c->kt_begin("bdm_max_abs_kernel");
            bdm_max_abs_kernel<NL><<<nb, 32, 0, st>>>(c->d_bdm, nb, c->dirMXY, c->dir_scal + 2 * ES, c->dir_part);
            c->kt_end();
            CUDA_TRY(c, cudaGetLastError());
          }
        return 0;
      }
    if(op == 3) // frobenius_product_of_sums.cxx, per block
      {
        if(nb)
          {
            c->kt_begin("bdm_frobenius_kernel");
            bdm_frobenius_kernel<NL><<<nb, 32, 0, st>>>(c->d_bdm, nb, c->Xin, c->dirDX, c->Yin, c->dirDY,
                                                        c->dir_colsum, c->dir_part);
            c->kt_end();
            CUDA_TRY(c, cudaGetLastError());
          }
        return 0;
      }
    // ---- compute_search_direction.cxx:44-90 ----
    const bool corrector = arg != 0;
    const size_t bytes = c->wXY * 8;
    // R = beta mu I - X Y (- dX dY in the corrector phase)
    if(bytes)
      CUDA_TRY(c, cudaMemcpyAsync(c->dirR, c->dirMXY, bytes, cudaMemcpyDeviceToDevice, st));
    if(corrector)
      if(int rc = bdm_product(c, "gemm_dX_dY", c->d_gemmDXDY, -1, 1, c->dirR))
        return rc;
    if(c->bdm_cols)
      {
        c->kt_begin("bdm_add_diagonal_kernel");
        bdm_add_diagonal_kernel<NL><<<(c->bdm_cols + 127) / 128, 128, 0, st>>>(c->d_bdm, nb, c->bdm_cols, c->dirR,
                                                                               c->dir_scal);
        c->kt_end();
        CUDA_TRY(c, cudaGetLastError());
      }
    // Z = Symmetrize(X^-1 (PrimalResidues Y - R))
    if(int rc = bdm_product(c, "gemm_PR_Y", c->d_gemmPRY, 1, 0, c->dirZ))
      return rc;
    if(int rc = bdm_elementwise(c, 0, c->dirR, c->dirZ))
      return rc;
    if(int rc = bdm_cholesky_solve_symmetrize(c, c->dirZ, 0))
      return rc;
    // dx = -dual_residues - Tr(A_p Z)
    if(c->K)
      {
        c->kt_begin("schur_rhs_kernel");
        schur_rhs_kernel<NL><<<(unsigned)((c->K + 63) / 64), 64, 0, st>>>(c->d_bdm, J, c->d_row_block, c->K, c->V,
                                                                          c->dirZ, c->dir_dual, c->sol_x);
        c->kt_end();
        CUDA_TRY(c, cudaGetLastError());
      }
    // dy = primal_residue_p; (dx, dy) <- solve_schur_complement_equation
    CUDA_TRY(c, cudaMemcpyAsync(c->sol_y, c->dir_prp, (size_t)N * ES * 8, cudaMemcpyDeviceToDevice, st));
    if(int rc = schur_solve(c))
      return rc;
    // dX = PrimalResidues + sum_p A_p dx_p
    if(nb)
      {
        const int ms = c->max_s;
        dim3 g(nb, (unsigned)std::min<long>(((long)ms * ms + 63) / 64, 65535));
        c->kt_begin("weighted_sum_kernel");
        weighted_sum_kernel<NL><<<g, 64, 0, st>>>(c->d_bdm, nb, c->V, c->sol_x, c->dir_scal + ES, c->dirDX);
        c->kt_end();
        CUDA_TRY(c, cudaGetLastError());
      }
    if(int rc = bdm_elementwise(c, 1, c->dirPR, c->dirDX))
      return rc;
    // dY = -Symmetrize(X^-1 (dX Y - R))
    if(int rc = bdm_product(c, "gemm_dX_Y", c->d_gemmDXY, 1, 0, c->dirDY))
      return rc;
    if(int rc = bdm_elementwise(c, 0, c->dirR, c->dirDY))
      return rc;
    return bdm_cholesky_solve_symmetrize(c, c->dirDY, 1);
  }
  // step_length.cxx:27-46 up to the reduction over the blocks (row N3)
  static int step_length(sdpb_b200_ctx *c, int which)
  {
    const int nb = 2 * c->J;
    cudaStream_t st = c->stream;
    c->cur = st;
    constexpr int ES = Fmt<NL>::ES;
    if(nb == 0 || c->bdm_cols == 0)
      return 0;
    const size_t smem = tridiag_smem_bytes<NL>(c->max_s);
    if(smem > (size_t)SMEM_OPT_IN)
      {
        c->error = "step_length: a block of " + std::to_string(c->max_s) + " rows does not fit the tridiagonalisation kernel";
        return SDPB_B200_ERR_ARG;
      }
    limb_t *A = c->dirZ;
    const limb_t *L = which == 0 ? c->X : c->LY;
    const uint32_t *recip = which == 0 ? c->recipX : c->recipY;
    CUDA_TRY(c, cudaMemcpyAsync(A, which == 0 ? c->dirDX : c->dirDY, c->wXY * 8, cudaMemcpyDeviceToDevice, st));
    // A := L^-1 A L^-T
    if(int rc = bdm_trsm_rl(c, 2, L, recip, A)) // A L^-T
      return rc;
    if(int rc = bdm_trsm_rl(c, 0, L, recip, A)) // L^-1 (.)
      return rc;
    // tridiagonal form, then the smallest eigenvalue
    if(int rc = smem_opt_in(c, eig_tridiag_kernel<NL>))
      return rc;
    c->kt_begin("eig_tridiag_kernel");
    eig_tridiag_kernel<NL><<<nb, TRIDIAG_THREADS, smem, st>>>(c->d_bdm, A, c->eig_d, c->eig_e, c->dir_scal + ES);
    c->kt_end();
    CUDA_TRY(c, cudaGetLastError());
    if(int rc = smem_opt_in(c, eig_laguerre_kernel<NL>))
      return rc;
    c->kt_begin("eig_laguerre_kernel");
    eig_laguerre_kernel<NL><<<nb, 32, sizeof(LaguerreSmem<NL>), st>>>(c->d_bdm, c->eig_d, c->eig_e, c->eig_e2,
                                                                      c->dir_scal + 3 * ES, c->dir_part, c->eig_iter);
    c->kt_end();
    CUDA_TRY(c, cudaGetLastError());
    return 0;
  }
  static int scalar(sdpb_b200_ctx *c, int op, int k, long count,
                    const limb_t *a, const limb_t *b, limb_t *r)
  {
    if(op == 8 || op == 9)
      {
        const size_t sm = ((sizeof(coop::Work<NL>) + 15) & ~(size_t)15)
                          + (TileGeom<NL>::SW + TileGeom<NL>::RS) * 4;
        c->kt_begin("coop_test_kernel");
        coop_test_kernel<NL><<<(unsigned)count, 32, sm, c->cur>>>(op, count, a, r);
        c->kt_end();
        CUDA_TRY(c, cudaGetLastError());
        return 0;
      }
    c->kt_begin("scalar_op_kernel");
    scalar_op_kernel<NL><<<(unsigned)((count + 127) / 128), 128, 0, c->cur>>>(
      op, k, count, a, b, r);
    c->kt_end();
    CUDA_TRY(c, cudaGetLastError());
    return 0;
  }
};


#define SDPB_CAT2(a, b) a##b
#define SDPB_CAT(a, b) SDPB_CAT2(a, b)
extern "C" __attribute__((visibility("default"))) const LaunchTable
  SDPB_CAT(sdpb_b200_launch_nl, SDPB_NL) = {sizeof(sdpb_b200_ctx), sizeof(LaunchTable), &Launch<SDPB_NL>::cholesky, &Launch<SDPB_NL>::pairings,
     &Launch<SDPB_NL>::schur_and_Q, &Launch<SDPB_NL>::schur_solve, &Launch<SDPB_NL>::scale_multiply_add,
     &Launch<SDPB_NL>::scalar, &Launch<SDPB_NL>::direction, &Launch<SDPB_NL>::step_length};

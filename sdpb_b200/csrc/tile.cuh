// Tiled sm_100a kernels for the GEMM-class stages of the Schur step: batched
// Cholesky, triangular solve and matrix product on mpf numbers.
//
// One scheme serves all three.  A CTA of 256 threads owns a 16x16 tile of
// outputs, one mpf accumulator per thread, held in REGISTERS for the whole
// k-loop (mpfw::Reg, 32-bit words).  Operand tiles (16 x 8 elements of A and
// 8 x 16 of B per step) are staged into shared memory by TMA bulk copies
// (cp.async.bulk + mbarrier complete_tx, one 128-byte element per copy so that
// any stride / transposition is free), double buffered, at a padded stride
// that makes the 128-bit shared loads bank-conflict free.  Every output
// element receives its updates in ascending k, one mpf_mul and one
// mpf_add/sub per update: the canonical order of DESIGN.md §3, so the
// left-looking schedule used here is bit-identical to the oracle's.
//
// Divisions by a Cholesky pivot use the exact reciprocal scheme of mpfw.h; the
// reciprocals are produced once per pivot by the factorisation and kept in
// HBM next to the factor (they are reused by the triangular solves).
#pragma once
#include "mpfw.h"
#include "coop.cuh"

#include <cuda_runtime.h>
#include <stdint.h>

namespace sdpb_b200
{
using mpfw::Reg;

constexpr int TS = 16; // tile side
constexpr int KC = 4;  // k-chunk

// Thread -> output element of the 16x16 tile.  A warp covers 8 ROWS x 4 COLUMNS (even warps
// rows 0-7, odd warps rows 8-15; warp pair w/2 columns 4(w/2)..4(w/2)+3), so that a ragged
// tile -- 8 valid rows of a 40-row block, 12 valid columns of N = 300 -- leaves whole warps
// without work instead of half-filling every warp: their issue slots go to the other CTA.
__device__ __forceinline__ int tile_ti() { return (threadIdx.x & 7) + ((threadIdx.x >> 2) & 8); }
__device__ __forceinline__ int tile_tj() { return ((threadIdx.x >> 3) & 3) + ((threadIdx.x >> 4) & 12); }

// Resident CTAs per SM the 256-thread tile kernels are compiled for.  Two (128 registers per
// thread) up to ~1100 bits; beyond that an element no longer fits next to the accumulator and the
// product in 128 registers (NL = 26: 3 KB of spill stack per thread), so one CTA with 255.
#ifndef SDPB_TILE_OCC_ONE_FROM
#define SDPB_TILE_OCC_ONE_FROM 20
#endif
template <int NL> struct TileOcc
{
  static constexpr int value = NL >= SDPB_TILE_OCC_ONE_FROM ? 1 : 2;
};
template <int NL> struct TileGeom
{
  static constexpr int ES = (NL + 2) & ~1;                  // 64-bit words / element
  static constexpr int EW = 2 * ES;                         // 32-bit words / element
  static constexpr int EB = 8 * ES;                         // bytes / element
  static constexpr int SW = EW + ((ES % 4 == 0) ? 4 : 0);   // shared stride, 32-bit words
  static constexpr int RW = 2 * NL + 4;                     // reciprocal words
  static constexpr int RS = (RW + 3) & ~3;                  // reciprocal stride (16-byte multiple)
};

// The big arithmetic bodies are kept out of line (one copy per precision): the
// accumulator travels in registers through the call, the ~70 register moves
// are a few percent of a multiply-accumulate, and compile time and code size
// stay bounded at 1536 bits.
template <int NL>
__device__ __noinline__ Reg<NL> mac_nl(Reg<NL> acc, const uint32_t *a, const uint32_t *b,
                                       bool negate)
{
  mpfw::mac<NL>(acc, a, b, negate);
  return acc;
}
// the same for operands that both live in shared memory (the k-loop of the tile kernels): telling
// the compiler so turns the generic LD.E of the operand words into LDS
template <int NL>
__device__ __noinline__ Reg<NL> mac_ss_nl(Reg<NL> acc, const uint32_t *a, const uint32_t *b,
                                          bool negate)
{
  __builtin_assume(__isShared(a));
  __builtin_assume(__isShared(b));
  mpfw::mac<NL>(acc, a, b, negate);
  return acc;
}
// x / pivot, pivot given as a packed element, R its reciprocal
template <int NL>
__device__ __noinline__ Reg<NL> div_nl(Reg<NL> x, const uint32_t *piv, const uint32_t *R)
{
  mpfw::div_recip<NL>(x, (int32_t)piv[1], (int32_t)piv[0], piv + 2, R);
  return x;
}
// ------------------------------------------------------------ TMA / mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity)
{
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok)
               : "r"(smem_u32(bar)), "r"(parity)
               : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                 smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// make this thread's earlier generic-proxy global writes visible to later TMA reads
__device__ __forceinline__ void fence_async_proxy()
{
  asm volatile("fence.proxy.async;" ::: "memory");
}

// One operand of the k-loop: element (x, k) lives at base + (x*sx + k*sk) elements.
struct Operand
{
  const uint64_t *base;
  long sx, sk;
  int nx; // valid x in this tile (<= TS)
};

template <int NL> struct TileSmem
{
  typedef TileGeom<NL> G;
  uint32_t a[2][KC * TS * G::SW];
  uint32_t b[2][KC * TS * G::SW];
  uint32_t diag[TS * TS * G::SW]; // a factored diagonal tile, [k][x]
  uint32_t vec[2][TS * G::SW];    // a finished column / row of the current tile
  uint32_t recip[TS * G::RS];     // reciprocals of the 16 pivots of `diag`
  uint64_t bar[2];
  int bad;
  coop::Work<NL> work;            // workspace of the warp-cooperative pivot (coop.cuh)
};

// acc -+= sum_k A(ti,k) B(k,tj), k ascending in [0, K); `it` counts the chunks
// this CTA has staged so far (selects buffer and mbarrier phase).  All 256
// threads must call this together.
template <int NL>
__device__ __forceinline__ void tile_k_loop(Reg<NL> &acc, bool negate, const Operand &A,
                                            const Operand &B, int K, TileSmem<NL> &sm,
                                            uint32_t &it, bool active)
{
  typedef TileGeom<NL> G;
  const int t = threadIdx.x;
  const int ti = tile_ti(), tj = tile_tj();
  const int nchunks = (K + KC - 1) / KC;
  if(nchunks == 0)
    return;
  // this thread's copy duty: t < KC*16 -> A element (x = t&15, kk = t>>4), the
  // next KC*16 threads -> B element likewise; one TMA bulk copy each per chunk
  const bool dutyA = t < KC * TS, duty = t < 2 * KC * TS;
  const int dx = t & (TS - 1), dk = (t >> 4) & (KC - 1);
  const Operand &O = dutyA ? A : B;
  auto issue = [&](int c) {
    const uint32_t s = (it + c) & 1;
    const int k0 = c * KC;
    const int kcnt = min(KC, K - k0);
    if(t == 0)
      mbar_expect_tx(&sm.bar[s], (uint32_t)(G::EB * kcnt * (A.nx + B.nx)));
    if(duty && dx < O.nx && dk < kcnt)
      {
        uint32_t *dst = (dutyA ? sm.a[s] : sm.b[s]) + (dk * TS + dx) * G::SW;
        bulk_g2s(dst, O.base + (dx * O.sx + (k0 + dk) * O.sk) * G::ES, G::EB, &sm.bar[s]);
      }
  };
  issue(0);
  for(int c = 0; c < nchunks; ++c)
    {
      if(c + 1 < nchunks)
        issue(c + 1);
      const uint32_t s = (it + c) & 1, parity = ((it + c) >> 1) & 1;
      while(!mbar_try_wait(&sm.bar[s], parity))
        {
        }
      const int kcnt = min(KC, K - c * KC);
      if(active)
        {
          const uint32_t *pa = sm.a[s] + ti * G::SW, *pb = sm.b[s] + tj * G::SW;
          for(int kk = 0; kk < kcnt; ++kk)
            acc = mac_ss_nl<NL>(acc, pa + kk * TS * G::SW, pb + kk * TS * G::SW, negate);
        }
      __syncthreads();
    }
  it += nchunks;
}

template <int NL> __device__ __forceinline__ void tile_smem_init(TileSmem<NL> &sm)
{
  if(threadIdx.x == 0)
    {
      mbar_init(&sm.bar[0], 1);
      mbar_init(&sm.bar[1], 1);
      sm.bad = -1;
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  __syncthreads();
}

template <int NL> __device__ __forceinline__ void ldg_reg(Reg<NL> &r, const uint64_t *p)
{
  mpfw::load<NL>(r, reinterpret_cast<const uint32_t *>(p));
}
template <int NL> __device__ __forceinline__ void stg_reg(uint64_t *p, const Reg<NL> &r)
{
  mpfw::store<NL>(reinterpret_cast<uint32_t *>(p), r);
}

// ------------------------------------------------------------------- GEMM
struct GemmTileDesc // C(i,j) = sum_l A(i,l) B(l,j); strides in elements
{
  const uint64_t *A, *B;
  uint64_t *C;
  long sa_i, sa_l, sb_l, sb_j;
  int M, N, K;
  int sym;   // 1: only tiles/elements with i >= j, result mirrored
  int tile0; // first linear tile index of this matrix in the launch
  // Block structure of bases_blocks = I_m (x) v (SDP/set_bases_blocks.cxx:3-22): with hb rows
  // and nb columns per diagonal block, an operand entry (l, x) is an exact zero unless
  // l / hb == x / nb (band 1: the B operand, 2: the A operand), or -- for T = L_X^-1 V, whose
  // columns start with the zeros of V -- unless l >= (x / nb) hb (band 3: both operands).
  // A multiply-accumulate with an exact zero leaves the accumulator untouched
  // (mpf: 0 * x = 0, c + 0 = c), so only the k-range that can hold non-zeros is visited.
  int band, hb, nb;
};
// k-range [klo, khi) of the products that can be non-zero for an output tile
__device__ __forceinline__ void band_range(const GemmTileDesc &d, int i0, int i1, int j0, int j1,
                                           int &klo, int &khi)
{
  klo = 0;
  khi = d.K;
  if(d.band == 1)
    {
      klo = (j0 / d.nb) * d.hb;
      khi = min(d.K, ((j1 - 1) / d.nb + 1) * d.hb);
    }
  else if(d.band == 2)
    {
      klo = (i0 / d.nb) * d.hb;
      khi = min(d.K, ((i1 - 1) / d.nb + 1) * d.hb);
    }
  else if(d.band == 3)
    klo = max(i0 / d.nb, j0 / d.nb) * d.hb;
}

// grid.x = total number of 16x16 tiles over all matrices (host prefix sums in
// `tile0`); the matrix of a tile is found by binary search.
template <int NL>
__global__ void __launch_bounds__(256, TileOcc<NL>::value)
gemm_tile_kernel(const GemmTileDesc *descs, int count)
{
  typedef TileGeom<NL> G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem<NL> &sm = *reinterpret_cast<TileSmem<NL> *>(smem_raw);
  int lo = 0, hi = count - 1;
  while(lo < hi)
    {
      const int mid = (lo + hi + 1) >> 1;
      if(descs[mid].tile0 <= (int)blockIdx.x)
        lo = mid;
      else
        hi = mid - 1;
    }
  const GemmTileDesc d = descs[lo];
  const int tiles_m = (d.M + TS - 1) / TS;
  const int tile = blockIdx.x - d.tile0;
  const int It = tile % tiles_m, Jt = tile / tiles_m;
  if(Jt * TS >= d.N || (d.sym && It < Jt))
    return;
  tile_smem_init(sm);
  const int ti = tile_ti(), tj = tile_tj();
  const int i = It * TS + ti, j = Jt * TS + tj;
  const bool active = i < d.M && j < d.N && (!d.sym || i >= j);
  Reg<NL> acc;
  mpfw::set_zero(acc);
  uint32_t it = 0;
  int klo, khi;
  band_range(d, It * TS, min(d.M, It * TS + TS), Jt * TS, min(d.N, Jt * TS + TS), klo, khi);
  Operand A{d.A + ((long)It * TS * d.sa_i + (long)klo * d.sa_l) * G::ES, d.sa_i, d.sa_l,
            min(TS, d.M - It * TS)};
  Operand B{d.B + ((long)Jt * TS * d.sb_j + (long)klo * d.sb_l) * G::ES, d.sb_j, d.sb_l,
            min(TS, d.N - Jt * TS)};
  tile_k_loop<NL>(acc, false, A, B, max(0, khi - klo), sm, it, active);
  if(active)
    {
      stg_reg<NL>(d.C + ((long)j * d.M + i) * G::ES, acc);
      if(d.sym && i != j)
        stg_reg<NL>(d.C + ((long)i * d.M + j) * G::ES, acc);
    }
}

// ---------------------------------------------------------------- Cholesky
struct PotrfDesc
{
  uint64_t *A;     // s x s, in place
  uint32_t *recip; // s reciprocals of the pivots, stride TileGeom::RS words
  int s;
  long si, sj; // logical (i,j), i >= j, lives at A + (i*si + j*sj) elements:
               // lower: si = 1, sj = s; upper (A = U^T U): si = s, sj = 1
  int id;
};

// factor the diagonal tile Jt: acc(ti,tj) already holds a_ij - sum_{k<J0} l_ik l_jk.
// Leaves the factored tile in sm.diag ([k][x] = L(J0+x, J0+k)) and the pivots'
// reciprocals in sm.recip, and writes both to HBM.  Returns false (uniformly)
// on a non-positive pivot, with its index in sm.bad.
template <int NL>
__device__ __forceinline__ bool potrf_diag_tile(Reg<NL> &acc, const PotrfDesc &d, int Jt,
                                                TileSmem<NL> &sm)
{
  typedef TileGeom<NL> G;
  const int ti = tile_ti(), tj = tile_tj();
  const int J0 = Jt * TS;
  const int nd = min(TS, d.s - J0);
  for(int kk = 0; kk < nd; ++kk)
    {
      uint32_t *pslot = sm.diag + (kk * TS + kk) * G::SW;
      if(ti == kk && tj == kk)
        {
          if(acc.sign <= 0)
            sm.bad = J0 + kk;
          else
            mpfw::store<NL>(pslot, acc);
        }
      __syncthreads();
      if(sm.bad >= 0)
        return false;
      // l_kk = sqrt(a_kk) and its reciprocal, by the 32 lanes of warp 0 together
      if(threadIdx.x < 32)
        coop::pivot<NL>(sm.work, pslot, sm.recip + kk * G::RS, d.recip + (long)(J0 + kk) * G::RS);
      __syncthreads();
      if(ti == kk && tj == kk)
        mpfw::load<NL>(acc, pslot);
      if(tj == kk && ti > kk && ti < nd)
        {
          const uint32_t *piv = sm.diag + (kk * TS + kk) * G::SW;
          acc = div_nl<NL>(acc, piv, sm.recip + kk * G::RS);
          mpfw::store<NL>(sm.diag + (kk * TS + ti) * G::SW, acc);
        }
      __syncthreads();
      if(ti > kk && tj > kk && ti >= tj && ti < nd)
        acc = mac_nl<NL>(acc, sm.diag + (kk * TS + ti) * G::SW, sm.diag + (kk * TS + tj) * G::SW, true);
    }
  // write the tile: factor below/on the diagonal, exact zeros above
  if(ti < nd && tj < nd)
    {
      if(ti >= tj)
        stg_reg<NL>(d.A + ((long)(J0 + ti) * d.si + (long)(J0 + tj) * d.sj) * G::ES, acc);
      else
        {
          Reg<NL> z;
          mpfw::set_zero(z);
          stg_reg<NL>(d.A + ((long)(J0 + ti) * d.si + (long)(J0 + tj) * d.sj) * G::ES, z);
        }
    }
  return true;
}

// tile (It, Jt), It > Jt:  X = (A_tile - sum_k ...) L_JJ^{-T}; acc(ti,tj) holds
// the updated a_ij; sm.diag / sm.recip hold the factored diagonal tile Jt.
template <int NL>
__device__ __forceinline__ void potrf_row_tile_solve(Reg<NL> &acc, const PotrfDesc &d, int It,
                                                     int Jt, TileSmem<NL> &sm)
{
  typedef TileGeom<NL> G;
  const int ti = tile_ti(), tj = tile_tj();
  const int J0 = Jt * TS, I0 = It * TS;
  const int nd = min(TS, d.s - J0), ni = min(TS, d.s - I0);
  for(int kk = 0; kk < nd; ++kk)
    {
      uint32_t *xs = sm.vec[kk & 1];
      if(tj == kk && ti < ni)
        {
          const uint32_t *piv = sm.diag + (kk * TS + kk) * G::SW;
          acc = div_nl<NL>(acc, piv, sm.recip + kk * G::RS);
          mpfw::store<NL>(xs + ti * G::SW, acc);
        }
      __syncthreads();
      if(tj > kk && tj < nd && ti < ni)
        acc = mac_nl<NL>(acc, xs + ti * G::SW, sm.diag + (kk * TS + tj) * G::SW, true);
    }
  if(ti < ni && tj < nd)
    {
      stg_reg<NL>(d.A + ((long)(I0 + ti) * d.si + (long)(J0 + tj) * d.sj) * G::ES, acc);
      Reg<NL> z;
      mpfw::set_zero(z);
      stg_reg<NL>(d.A + ((long)(J0 + tj) * d.si + (long)(I0 + ti) * d.sj) * G::ES, z);
    }
}

// load the factored diagonal tile Jt and its reciprocals from HBM into shared
template <int NL>
__device__ __forceinline__ void load_diag_tile(const uint64_t *A, long si, long sj,
                                               const uint32_t *recip, int s, int Jt,
                                               TileSmem<NL> &sm)
{
  typedef TileGeom<NL> G;
  const int J0 = Jt * TS, nd = min(TS, s - J0);
  const int x = threadIdx.x & (TS - 1), k = threadIdx.x >> 4;
  if(x < nd && k < nd && x >= k)
    {
      const uint4 *src
        = reinterpret_cast<const uint4 *>(A + ((long)(J0 + x) * si + (long)(J0 + k) * sj) * G::ES);
      uint4 *dst = reinterpret_cast<uint4 *>(sm.diag + (k * TS + x) * G::SW);
#pragma unroll
      for(int w = 0; w < G::EB / 16; ++w)
        dst[w] = src[w];
    }
  for(int w = threadIdx.x; w < nd * G::RS; w += blockDim.x)
    sm.recip[w] = recip[(long)J0 * G::RS + w];
  __syncthreads();
}

// acc(ti,tj) = A(I0+ti, J0+tj) - sum_{k < J0} L(I0+ti,k) L(J0+tj,k)
template <int NL>
__device__ __forceinline__ void potrf_tile_update(Reg<NL> &acc, const PotrfDesc &d, int It, int Jt,
                                                  TileSmem<NL> &sm, uint32_t &it)
{
  typedef TileGeom<NL> G;
  const int ti = tile_ti(), tj = tile_tj();
  const int I0 = It * TS, J0 = Jt * TS;
  const int ni = min(TS, d.s - I0), nj = min(TS, d.s - J0);
  const bool active = ti < ni && tj < nj && (It > Jt || ti >= tj);
  if(active)
    ldg_reg<NL>(acc, d.A + ((long)(I0 + ti) * d.si + (long)(J0 + tj) * d.sj) * G::ES);
  else
    mpfw::set_zero(acc);
  Operand A{d.A + (long)I0 * d.si * G::ES, d.si, d.sj, ni};
  Operand B{d.A + (long)J0 * d.si * G::ES, d.si, d.sj, nj};
  tile_k_loop<NL>(acc, true, A, B, J0, sm, it, active);
}

// ---- level-synchronous batched Cholesky --------------------------------
// Block column Jt of every matrix of the batch is finished by three launches:
//   potrf_gemm_level   one CTA per tile of the block column, the diagonal one
//                      included: a_ij -= sum_{k<J0} l_ik l_jk
//   potrf_diag_warp    one WARP per matrix: factor the diagonal tile (below)
//   potrf_panel_rl     one CTA per tile below it: X = A_tile L_JJ^{-T}, the 16
//                      unknowns of a row shared by 16 threads
// `descs` is sorted by size (largest first); grid.x covers the prefix of
// matrices that still have a block column Jt.
template <int NL>
__global__ void __launch_bounds__(256, TileOcc<NL>::value)
potrf_gemm_level(const PotrfDesc *descs, int Jt, const int *status)
{
  typedef TileGeom<NL> G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem<NL> &sm = *reinterpret_cast<TileSmem<NL> *>(smem_raw);
  const PotrfDesc d = descs[blockIdx.x];
  const int It = Jt + blockIdx.y; // the diagonal tile included (factored by potrf_diag_warp)
  if(It * TS >= d.s || status[d.id] >= 0)
    return;
  tile_smem_init(sm);
  uint32_t it = 0;
  Reg<NL> acc;
  potrf_tile_update<NL>(acc, d, It, Jt, sm, it);
  const int ti = tile_ti(), tj = tile_tj();
  if(It * TS + ti < d.s && Jt * TS + tj < d.s)
    stg_reg<NL>(d.A + ((long)(It * TS + ti) * d.si + (long)(Jt * TS + tj) * d.sj) * G::ES, acc);
}

// ---- diagonal tiles of a BATCH: one warp per matrix ------------------------
// Factoring a 16x16 diagonal tile is a chain of 16 pivots (sqrt, reciprocal,
// column division, rank-1 update), ~0.5 ms of mostly serial work.  Giving it a
// whole 256-thread CTA (potrf_diag_level / potrf_diag_rl) leaves two tiles in
// flight per SM; here a tile lives in shared memory (packed lower triangle),
// one warp owns it -- the pivot by all 32 lanes (coop.cuh), the column and the
// trailing update one element per lane -- and eight tiles share an SM.
template <int NL> struct alignas(128) WarpTileSmem
{
  typedef TileGeom<NL> G;
  static constexpr int NE = TS * (TS + 1) / 2;
  uint32_t el[NE * G::SW];
  uint32_t recip[TS * G::RS];
  coop::Work<NL> work;
};
constexpr int DIAG_WARPS = 4;
__device__ __forceinline__ int tri_index(int i, int j) // i >= j, column-packed lower triangle
{
  return j * TS - j * (j - 1) / 2 + (i - j);
}
template <int NL>
__global__ void __launch_bounds__(32 * DIAG_WARPS, 2)
potrf_diag_warp(const PotrfDesc *descs, int count, int Jt, int *status)
{
  typedef TileGeom<NL> G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m = blockIdx.x * DIAG_WARPS + warp;
  if(m >= count)
    return;
  WarpTileSmem<NL> &sm = reinterpret_cast<WarpTileSmem<NL> *>(smem_raw)[warp];
  const PotrfDesc d = descs[m];
  const int J0 = Jt * TS;
  if(J0 >= d.s || status[d.id] >= 0)
    return;
  const int nd = min(TS, d.s - J0);
  if(lane == 0)
    sm.work.flag = 0;
  for(int e = lane; e < TS * TS; e += 32)
    {
      const int i = e & (TS - 1), j = e >> 4;
      if(i >= j && i < nd)
        {
          const uint4 *src = reinterpret_cast<const uint4 *>(
            d.A + ((long)(J0 + i) * d.si + (long)(J0 + j) * d.sj) * G::ES);
          uint4 *dst = reinterpret_cast<uint4 *>(sm.el + tri_index(i, j) * G::SW);
#pragma unroll
          for(int w = 0; w < G::EB / 16; ++w)
            dst[w] = src[w];
        }
    }
  __syncwarp();
  for(int kk = 0; kk < nd; ++kk)
    {
      uint32_t *pslot = sm.el + tri_index(kk, kk) * G::SW;
      if((int32_t)pslot[1] <= 0)
        {
          if(lane == 0)
            status[d.id] = J0 + kk;
          return;
        }
      coop::pivot<NL>(sm.work, pslot, sm.recip + kk * G::RS, d.recip + (long)(J0 + kk) * G::RS);
      {
        const int i = kk + 1 + lane;
        if(i < nd)
          {
            uint32_t *x = sm.el + tri_index(i, kk) * G::SW;
            Reg<NL> acc;
            mpfw::load<NL>(acc, x);
            acc = div_nl<NL>(acc, pslot, sm.recip + kk * G::RS);
            mpfw::store<NL>(x, acc);
          }
      }
      __syncwarp();
      const int r = nd - 1 - kk, cnt = r * (r + 1) / 2;
      for(int e = lane; e < cnt; e += 32)
        {
          int jj = 0, rem = e; // e -> (ii >= jj) in the r x r lower triangle, column-packed
          while(rem >= r - jj)
            {
              rem -= r - jj;
              ++jj;
            }
          const int i = kk + 1 + jj + rem, j = kk + 1 + jj;
          uint32_t *x = sm.el + tri_index(i, j) * G::SW;
          Reg<NL> acc;
          mpfw::load<NL>(acc, x);
          acc = mac_nl<NL>(acc, sm.el + tri_index(i, kk) * G::SW, sm.el + tri_index(j, kk) * G::SW, true);
          mpfw::store<NL>(x, acc);
        }
      __syncwarp();
    }
  // the factor below / on the diagonal, exact zeros above
  for(int e = lane; e < TS * TS; e += 32)
    {
      const int i = e & (TS - 1), j = e >> 4;
      if(i < nd && j < nd)
        {
          uint4 *dst = reinterpret_cast<uint4 *>(
            d.A + ((long)(J0 + i) * d.si + (long)(J0 + j) * d.sj) * G::ES);
          if(i >= j)
            {
              const uint4 *src = reinterpret_cast<const uint4 *>(sm.el + tri_index(i, j) * G::SW);
#pragma unroll
              for(int w = 0; w < G::EB / 16; ++w)
                dst[w] = src[w];
            }
          else
            {
#pragma unroll
              for(int w = 0; w < G::EB / 16; ++w)
                dst[w] = make_uint4(0, 0, 0, 0);
            }
        }
    }
}

// ---- right-looking variant ------------------------------------------------
// Same arithmetic (every element still receives its updates in ascending k),
// different schedule: after block column Jt is finished, ALL tiles to its right
// get its 16 updates at once.  The diagonal tile of the next level is then
// ready the moment the level starts, so the serial chain of a level is only
// factor(16 pivots) -> panel solve -> 16 updates, instead of also carrying the
// J0-long update loop of the left-looking form.  That chain bounds the time of
// a single large matrix (Cholesky of Q: N/16 levels).
//   potrf_diag_rl   one CTA per matrix: factor the (already updated) diagonal tile
//   potrf_panel_rl  one CTA per tile below it: X = A_tile L_JJ^{-T}
//   potrf_trail_rl  one CTA per tile (It >= Kt > Jt): a_ij -= sum_{k in column Jt} l_ik l_jk
// (one CTA per matrix and a serial chain of pivots: compiled for one CTA per SM, so that the
// pivot code keeps its operands in registers instead of the spill stack)
template <int NL>
__global__ void __launch_bounds__(256, 1)
potrf_diag_rl(const PotrfDesc *descs, int Jt, int *status)
{
  typedef TileGeom<NL> G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem<NL> &sm = *reinterpret_cast<TileSmem<NL> *>(smem_raw);
  const PotrfDesc d = descs[blockIdx.x];
  if(Jt * TS >= d.s || status[d.id] >= 0)
    return;
  tile_smem_init(sm);
  const int ti = tile_ti(), tj = tile_tj();
  const int J0 = Jt * TS, nd = min(TS, d.s - J0);
  Reg<NL> acc;
  if(ti < nd && tj < nd && ti >= tj)
    ldg_reg<NL>(acc, d.A + ((long)(J0 + ti) * d.si + (long)(J0 + tj) * d.sj) * G::ES);
  else
    mpfw::set_zero(acc);
  if(!potrf_diag_tile<NL>(acc, d, Jt, sm))
    {
      if(threadIdx.x == 0)
        status[d.id] = sm.bad;
    }
}
template <int NL>
__global__ void __launch_bounds__(256, TileOcc<NL>::value)
potrf_panel_rl(const PotrfDesc *descs, int Jt, const int *status)
{
  typedef TileGeom<NL> G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem<NL> &sm = *reinterpret_cast<TileSmem<NL> *>(smem_raw);
  const PotrfDesc d = descs[blockIdx.x];
  const int It = Jt + 1 + blockIdx.y;
  if(It * TS >= d.s || status[d.id] >= 0)
    return;
  load_diag_tile<NL>(d.A, d.si, d.sj, d.recip, d.s, Jt, sm);
  const int ti = tile_ti(), tj = tile_tj();
  Reg<NL> acc;
  if(It * TS + ti < d.s && Jt * TS + tj < d.s)
    ldg_reg<NL>(acc, d.A + ((long)(It * TS + ti) * d.si + (long)(Jt * TS + tj) * d.sj) * G::ES);
  else
    mpfw::set_zero(acc);
  potrf_row_tile_solve<NL>(acc, d, It, Jt, sm);
}
// ---- diagonal tile and panel in ONE launch, pipelined column by column ------------------------
// The panel tiles below a diagonal tile need column k of its factor (and pivot k's reciprocal)
// only at THEIR step k, so they do not have to wait for the whole tile: CTA 0 factors the diagonal
// tile and publishes every column as it is finished (the factor entries go straight to the matrix,
// a progress word counts the columns); CTA b >= 1 owns panel tile Jt + b and runs one column behind.
// A level of a single large matrix then costs the 16 pivots of the diagonal tile plus one panel
// step instead of the pivots plus a whole panel solve (94 of 490 us per level of Cholesky(Q) at
// c3).  Same arithmetic, same order per element as potrf_diag_rl + potrf_panel_rl.
// Forward progress: CTA 0 never waits; the grid (1 + tiles below <= 2 x 148) is co-resident.
constexpr int POTRF_FUSED_FAIL = 0x7fffffff;
__device__ __forceinline__ int ld_acquire_gpu(const int *p)
{
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int *p, int v)
{
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
template <int NL>
__global__ void __launch_bounds__(256, TileOcc<NL>::value)
potrf_diag_panel_rl(const PotrfDesc *descs, int Jt, int *status, int *progress)
{
  typedef TileGeom<NL> G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem<NL> &sm = *reinterpret_cast<TileSmem<NL> *>(smem_raw);
  const PotrfDesc d = descs[0];
  const int J0 = Jt * TS;
  if(J0 >= d.s || status[d.id] >= 0)
    return;
  tile_smem_init(sm);
  const int ti = tile_ti(), tj = tile_tj(), t = threadIdx.x;
  const int nd = min(TS, d.s - J0);
  Reg<NL> acc;
  if(blockIdx.x == 0)
    {
      // ---- the diagonal tile (potrf_diag_tile), every column published as soon as it is final
      if(ti < nd && tj < nd && ti >= tj)
        ldg_reg<NL>(acc, d.A + ((long)(J0 + ti) * d.si + (long)(J0 + tj) * d.sj) * G::ES);
      else
        mpfw::set_zero(acc);
      for(int kk = 0; kk < nd; ++kk)
        {
          uint32_t *pslot = sm.diag + (kk * TS + kk) * G::SW;
          if(ti == kk && tj == kk)
            {
              if(acc.sign <= 0)
                sm.bad = J0 + kk;
              else
                mpfw::store<NL>(pslot, acc);
            }
          __syncthreads();
          if(sm.bad >= 0)
            {
              if(t == 0)
                {
                  status[d.id] = sm.bad;
                  __threadfence();
                  st_release_gpu(progress, POTRF_FUSED_FAIL); // wake the panel CTAs: they leave
                }
              return;
            }
          if(t < 32)
            {
              coop::pivot<NL>(sm.work, pslot, sm.recip + kk * G::RS, d.recip + (long)(J0 + kk) * G::RS);
              __threadfence(); // the reciprocal words (global) before the progress word
            }
          __syncthreads();
          if(ti == kk && tj == kk)
            {
              mpfw::load<NL>(acc, pslot);
              stg_reg<NL>(d.A + ((long)(J0 + ti) * d.si + (long)(J0 + tj) * d.sj) * G::ES, acc);
              __threadfence();
            }
          if(tj == kk && ti > kk && ti < nd)
            {
              const uint32_t *piv = sm.diag + (kk * TS + kk) * G::SW;
              acc = div_nl<NL>(acc, piv, sm.recip + kk * G::RS);
              mpfw::store<NL>(sm.diag + (kk * TS + ti) * G::SW, acc);
              stg_reg<NL>(d.A + ((long)(J0 + ti) * d.si + (long)(J0 + tj) * d.sj) * G::ES, acc);
              __threadfence();
            }
          __syncthreads();
          if(t == 0)
            st_release_gpu(progress, J0 + kk + 1); // columns J0 .. J0+kk of the factor are in place
          if(ti > kk && tj > kk && ti >= tj && ti < nd)
            acc = mac_nl<NL>(acc, sm.diag + (kk * TS + ti) * G::SW, sm.diag + (kk * TS + tj) * G::SW, true);
        }
      // exact zeros above the diagonal (the factor itself went out column by column)
      if(ti < nd && tj < nd && ti < tj)
        {
          Reg<NL> z;
          mpfw::set_zero(z);
          stg_reg<NL>(d.A + ((long)(J0 + ti) * d.si + (long)(J0 + tj) * d.sj) * G::ES, z);
        }
      return;
    }
  // ---- panel tile It = Jt + blockIdx.x (potrf_row_tile_solve), one column behind the diagonal tile
  const int It = Jt + (int)blockIdx.x, I0 = It * TS;
  if(I0 >= d.s)
    return;
  const int ni = min(TS, d.s - I0);
  if(ti < ni && tj < nd)
    ldg_reg<NL>(acc, d.A + ((long)(I0 + ti) * d.si + (long)(J0 + tj) * d.sj) * G::ES);
  else
    mpfw::set_zero(acc);
  for(int kk = 0; kk < nd; ++kk)
    {
      if(t == 0)
        {
          int seen;
          while((seen = ld_acquire_gpu(progress)) < J0 + kk + 1)
            __nanosleep(200);
          sm.bad = seen == POTRF_FUSED_FAIL ? 1 : -1;
        }
      __syncthreads();
      if(sm.bad >= 0)
        return;
      // column kk of the factored diagonal tile (rows kk .. nd-1) and pivot kk's reciprocal, past L1
      for(int e = t; e < (nd - kk) * (G::EB / 16); e += 256)
        {
          const int x = kk + e / (G::EB / 16), w = e % (G::EB / 16);
          const uint4 *src = reinterpret_cast<const uint4 *>(
            d.A + ((long)(J0 + x) * d.si + (long)(J0 + kk) * d.sj) * G::ES);
          reinterpret_cast<uint4 *>(sm.diag + (kk * TS + x) * G::SW)[w] = __ldcg(src + w);
        }
      for(int w = t; w < G::RS; w += 256)
        sm.recip[kk * G::RS + w] = __ldcg(d.recip + (long)(J0 + kk) * G::RS + w);
      __syncthreads();
      uint32_t *xs = sm.vec[kk & 1];
      if(tj == kk && ti < ni)
        {
          const uint32_t *piv = sm.diag + (kk * TS + kk) * G::SW;
          acc = div_nl<NL>(acc, piv, sm.recip + kk * G::RS);
          mpfw::store<NL>(xs + ti * G::SW, acc);
        }
      __syncthreads();
      if(tj > kk && tj < nd && ti < ni)
        acc = mac_nl<NL>(acc, xs + ti * G::SW, sm.diag + (kk * TS + tj) * G::SW, true);
    }
  if(ti < ni && tj < nd)
    {
      stg_reg<NL>(d.A + ((long)(I0 + ti) * d.si + (long)(J0 + tj) * d.sj) * G::ES, acc);
      Reg<NL> z;
      mpfw::set_zero(z);
      stg_reg<NL>(d.A + ((long)(J0 + tj) * d.si + (long)(I0 + ti) * d.sj) * G::ES, z);
    }
}
// block column Jt (rows J0.., its 16 columns) <-> a contiguous buffer, for the broadcast of the
// panel-distributed Cholesky; the last word carries the matrix's status (a failed pivot on the
// owner must stop every rank)
template <int NL>
__global__ void panel_pack(const PotrfDesc *descs, int Jt, uint64_t *buf, int *status, int unpack)
{
  typedef TileGeom<NL> G;
  const PotrfDesc d = descs[0];
  const int J0 = Jt * TS, nd = min(TS, d.s - J0), rows = d.s - J0;
  const long total = (long)rows * nd;
  for(long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x)
    {
      const int j = (int)(e % nd), i = (int)(e / nd);
      uint4 *g = reinterpret_cast<uint4 *>(d.A + ((long)(J0 + i) * d.si + (long)(J0 + j) * d.sj) * G::ES);
      uint4 *b = reinterpret_cast<uint4 *>(buf + e * G::ES);
#pragma unroll
      for(int w = 0; w < G::EB / 16; ++w)
        {
          if(unpack)
            g[w] = b[w];
          else
            b[w] = g[w];
        }
    }
  // the pivots' reciprocals travel with the panel (the Schur solves divide by them on every rank)
  if(blockIdx.x == 0)
    {
      uint32_t *rb = reinterpret_cast<uint32_t *>(buf + (long)d.s * TS * G::ES + 2);
      uint32_t *rg = d.recip + (long)J0 * G::RS;
      for(int w = threadIdx.x; w < nd * G::RS; w += blockDim.x)
        {
          if(unpack)
            rg[w] = rb[w];
          else
            rb[w] = rg[w];
        }
    }
  if(blockIdx.x == 0 && threadIdx.x == 0)
    {
      uint64_t *sw = buf + (long)d.s * TS * G::ES; // fixed slot past the largest panel
      if(unpack)
        {
          if((int)(int64_t)*sw >= 0)
            status[d.id] = (int)(int64_t)*sw;
        }
      else
        *sw = (uint64_t)(int64_t)status[d.id];
    }
}

// cyc_mod > 1: only the tile columns Kt with Kt % cyc_mod == cyc_rem (this rank's share)
template <int NL>
__global__ void __launch_bounds__(256, TileOcc<NL>::value)
potrf_trail_rl(const PotrfDesc *descs, int Jt, const int *status, int cyc_mod, int cyc_rem)
{
  typedef TileGeom<NL> G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem<NL> &sm = *reinterpret_cast<TileSmem<NL> *>(smem_raw);
  const PotrfDesc d = descs[blockIdx.x];
  // blockIdx.y -> (a >= b) among the tiles right of Jt
  int t = blockIdx.y, a = 0;
  while(t > a)
    {
      t -= a + 1;
      ++a;
    }
  const int It = Jt + 1 + a, Kt = Jt + 1 + t;
  if(It * TS >= d.s || status[d.id] >= 0 || (cyc_mod > 1 && Kt % cyc_mod != cyc_rem))
    return;
  tile_smem_init(sm);
  const int ti = tile_ti(), tj = tile_tj();
  const int I0 = It * TS, K0 = Kt * TS, J0 = Jt * TS;
  const int ni = min(TS, d.s - I0), nk = min(TS, d.s - K0);
  const bool active = ti < ni && tj < nk && (It > Kt || ti >= tj);
  Reg<NL> acc;
  uint64_t *mine = d.A + ((long)(I0 + ti) * d.si + (long)(K0 + tj) * d.sj) * G::ES;
  if(active)
    ldg_reg<NL>(acc, mine);
  else
    mpfw::set_zero(acc);
  uint32_t it = 0;
  Operand A{d.A + ((long)I0 * d.si + (long)J0 * d.sj) * G::ES, d.si, d.sj, ni};
  Operand B{d.A + ((long)K0 * d.si + (long)J0 * d.sj) * G::ES, d.si, d.sj, nk};
  tile_k_loop<NL>(acc, true, A, B, TS, sm, it, active); // a full block column: rows exist below it
  if(active)
    stg_reg<NL>(mine, acc);
}

// a factored diagonal tile as the packed lower triangle ([tri_index(x, k)] = L(J0+x, J0+k), x >= k:
// 136 of the 256 slots, so that eight 64-column CTAs fit an SM) and the reciprocals of its pivots
template <int NL> struct DiagSmem
{
  typedef TileGeom<NL> G;
  uint32_t diag[TS * (TS + 1) / 2 * G::SW];
  uint32_t recip[TS * G::RS];
};
// cooperative load of a factored diagonal tile and its reciprocals
template <int NL>
__device__ __forceinline__ void load_diag(DiagSmem<NL> &sm, const uint64_t *A, long si, long sj,
                                          const uint32_t *recip, int s, int Jt, long rstep = 1)
{
  typedef TileGeom<NL> G;
  const int J0 = Jt * TS, nd = min(TS, s - J0);
  for(int e = threadIdx.x; e < TS * TS; e += blockDim.x)
    {
      const int x = e & (TS - 1), k = e >> 4;
      if(x < nd && k < nd && x >= k)
        {
          const uint4 *src = reinterpret_cast<const uint4 *>(
            A + ((long)(J0 + x) * si + (long)(J0 + k) * sj) * G::ES);
          uint4 *dst = reinterpret_cast<uint4 *>(sm.diag + tri_index(x, k) * G::SW);
#pragma unroll
          for(int w = 0; w < G::EB / 16; ++w)
            dst[w] = src[w];
        }
    }
  for(int w = threadIdx.x; w < nd * G::RS; w += blockDim.x)
    sm.recip[w] = recip[(long)(J0 + w / G::RS) * rstep * G::RS + w % G::RS];
  __syncthreads();
}

// columns per CTA of trsm_diag_level: 64 (two warps) keeps the padding of a 300-column band at
// 6 % (3 x 128 would be 28 %) and lets eight CTAs share an SM
constexpr int ROWS_PER_CTA = 64;

// -------------------------------------------------------- triangular solve
struct TrsmTileDesc // X <- L^{-1} B in place, L lower p x p (column-major)
{
  const uint64_t *L;
  const uint32_t *recip; // reciprocals of diag(L)
  uint64_t *B;           // p x ncols, column-major, ld = p
  int p, ncols;
  // right-hand sides that are bases_blocks (see GemmTileDesc): column c is zero above row
  // (c / nb) hb, and so is the solution; hb == 0: dense
  int hb, nb;
  // General views, in elements (0 = the standard column-major one): L(i, k) at L + i lsi + k lsk,
  // B(i, c) at B + i bsi + c bsc, the reciprocal of pivot i at recip + i rstep RS.  The level
  // kernels then also run L^-T B (both index ranges reversed: a forward substitution again, the
  // unknowns in descending order) and B L^-T (B read by rows) of the block-diagonal solves.
  long lsi = 0, lsk = 0, bsi = 0, bsc = 0;
  int rstep = 0;
};
struct TrsmView
{
  long lsi, lsk, bsi, bsc, rstep;
  __device__ __forceinline__ explicit TrsmView(const TrsmTileDesc &d)
    : lsi(d.lsi ? d.lsi : 1), lsk(d.lsk ? d.lsk : d.p), bsi(d.bsi ? d.bsi : 1), bsc(d.bsc ? d.bsc : d.p),
      rstep(d.rstep ? d.rstep : 1)
  {
  }
};

// Row tile It of every solve of the batch (sorted by p, largest first):
//   trsm_gemm_level  16x16 tiles: b_ic -= sum_{k<I0} l_ik x_kc
//   trsm_diag_level  one THREAD per column: the 16 rows of the tile top to bottom
template <int NL>
__global__ void __launch_bounds__(256, TileOcc<NL>::value) trsm_gemm_level(const TrsmTileDesc *descs, int It)
{
  typedef TileGeom<NL> G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  TileSmem<NL> &sm = *reinterpret_cast<TileSmem<NL> *>(smem_raw);
  const TrsmTileDesc d = descs[blockIdx.x];
  const int c0 = blockIdx.y * TS, I0 = It * TS;
  if(I0 >= d.p || c0 >= d.ncols)
    return;
  // leading zeros of these columns (block structure): rows above klo hold exact zeros
  const int klo = d.hb ? ((c0 / d.nb) * d.hb) & ~(KC - 1) : 0;
  if(I0 <= klo)
    return; // nothing above this tile can be non-zero: no update
  tile_smem_init(sm);
  const int ti = tile_ti(), tj = tile_tj();
  const int nc = min(TS, d.ncols - c0), ni = min(TS, d.p - I0);
  const bool active = ti < ni && tj < nc;
  Reg<NL> acc;
  const TrsmView v(d);
  uint64_t *mine = d.B + ((long)(c0 + tj) * v.bsc + (long)(I0 + ti) * v.bsi) * G::ES;
  if(active)
    ldg_reg<NL>(acc, mine);
  else
    mpfw::set_zero(acc);
  uint32_t it = 0;
  Operand A{d.L + ((long)I0 * v.lsi + (long)klo * v.lsk) * G::ES, v.lsi, v.lsk, ni};
  Operand B{d.B + ((long)c0 * v.bsc + (long)klo * v.bsi) * G::ES, v.bsc, v.bsi, nc};
  tile_k_loop<NL>(acc, true, A, B, I0 - klo, sm, it, active);
  if(active)
    stg_reg<NL>(mine, acc);
}
template <int NL>
__global__ void __launch_bounds__(ROWS_PER_CTA, 4 * TileOcc<NL>::value)
trsm_diag_level(const TrsmTileDesc *descs, int It)
{
  typedef TileGeom<NL> G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  DiagSmem<NL> &sm = *reinterpret_cast<DiagSmem<NL> *>(smem_raw);
  const TrsmTileDesc d = descs[blockIdx.x];
  const int I0 = It * TS, col0 = blockIdx.y * ROWS_PER_CTA;
  if(I0 >= d.p || col0 >= d.ncols)
    return;
  const TrsmView v(d);
  load_diag<NL>(sm, d.L, v.lsi, v.lsk, d.recip, d.p, It, v.rstep);
  const int col = col0 + threadIdx.x;
  if(col >= d.ncols)
    return;
  const int ni = min(TS, d.p - I0);
  uint64_t *colp = d.B + ((long)col * v.bsc + (long)I0 * v.bsi) * G::ES;
  const long rs = v.bsi * G::ES; // words between consecutive rows of the column
  for(int ii = 0; ii < ni; ++ii)
    {
      Reg<NL> acc;
      ldg_reg<NL>(acc, colp + ii * rs);
      for(int kk = 0; kk < ii; ++kk)
        acc = mac_nl<NL>(acc, sm.diag + tri_index(ii, kk) * G::SW,
                         reinterpret_cast<const uint32_t *>(colp + kk * rs), true);
      acc = div_nl<NL>(acc, sm.diag + tri_index(ii, ii) * G::SW, sm.recip + ii * G::RS);
      stg_reg<NL>(colp + ii * rs, acc);
    }
}

// The same diagonal solve with the 16 steps of a tile shared by 16 x 16 threads (element (i, c) of
// the tile per thread, as potrf_row_tile_solve does for the panel of a factorisation): step k is
// one division of row k and one update of the rows below it, so the chain of a tile is
// 16 (div + mac) instead of the 136 operations one thread per column runs through.  It costs
// ~2.5x the pipe time (most lanes of a step idle), so the host takes it only where the batch is too
// small to fill the SMs with one thread per column (BASELINE configs 0 and 1: N = 20 / 60).
// Same per-element order: updates in ascending k, then the division.
template <int NL> struct DiagTileSmem
{
  typedef TileGeom<NL> G;
  DiagSmem<NL> d;
  uint32_t x[2][TS * G::SW]; // the solved row k, one element per column of the tile
};
template <int NL>
__global__ void __launch_bounds__(256, TileOcc<NL>::value)
trsm_diag_tile(const TrsmTileDesc *descs, int It)
{
  typedef TileGeom<NL> G;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  DiagTileSmem<NL> &sm = *reinterpret_cast<DiagTileSmem<NL> *>(smem_raw);
  const TrsmTileDesc d = descs[blockIdx.x];
  const int I0 = It * TS, c0 = blockIdx.y * TS;
  if(I0 >= d.p || c0 >= d.ncols)
    return;
  const TrsmView v(d);
  load_diag<NL>(sm.d, d.L, v.lsi, v.lsk, d.recip, d.p, It, v.rstep);
  const int ti = tile_ti(), tj = tile_tj();
  const int ni = min(TS, d.p - I0), nc = min(TS, d.ncols - c0);
  const bool active = ti < ni && tj < nc;
  uint64_t *mine = d.B + ((long)(c0 + tj) * v.bsc + (long)(I0 + ti) * v.bsi) * G::ES;
  Reg<NL> acc;
  if(active)
    ldg_reg<NL>(acc, mine);
  else
    mpfw::set_zero(acc);
  for(int kk = 0; kk < ni; ++kk)
    {
      uint32_t *xs = sm.x[kk & 1];
      if(active && ti == kk)
        {
          acc = div_nl<NL>(acc, sm.d.diag + tri_index(kk, kk) * G::SW, sm.d.recip + kk * G::RS);
          mpfw::store<NL>(xs + tj * G::SW, acc);
        }
      __syncthreads();
      if(active && ti > kk)
        acc = mac_ss_nl<NL>(acc, sm.d.diag + tri_index(ti, kk) * G::SW, xs + tj * G::SW, true);
    }
  if(active)
    stg_reg<NL>(mine, acc);
}

// ---- the whole batched solve in ONE launch ---------------------------------
// Columns of X = L^-1 B are independent, so a CTA that owns a group of columns of one matrix
// can walk down ALL row tiles by itself: no level-synchronous launches (and none of their
// tails), no kernel boundary between the update and the diagonal solve.  Per row tile:
//   update  b_ic -= sum_{k < I0} l_ik x_kc   as (16 x TC) register tiles, TC = blockDim / 16,
//           one pass per TC columns; a LAST tile with <= 8 rows switches to (8 x 2 TC) tiles, so
//           a 40-row block (2.5 tiles) keeps every thread busy instead of half of them;
//   solve   the 16 rows of the tile top to bottom, ONE THREAD PER COLUMN (every lane busy; the
//           16-threads-per-row form spends most of its time with one row in sixteen active).
// The block size is chosen by the host so that (columns per CTA) ~ blockDim: both phases then
// use (nearly) all threads.  Same per-element operation order as the level kernels.
template <int NL> struct WalkGeom
{
  typedef TileGeom<NL> G;
  static constexpr int NTRI = TS * (TS + 1) / 2;
  static constexpr size_t OFF_DIAG = 16;                                  // after the two mbarriers
  static constexpr size_t OFF_RECIP = OFF_DIAG + (size_t)NTRI * G::SW * 4;
  static constexpr size_t OFF_A = OFF_RECIP + (size_t)TS * G::RS * 4;
  static constexpr size_t OFF_B = OFF_A + (size_t)2 * KC * TS * G::SW * 4;
  static size_t bytes(int tc) { return OFF_B + (size_t)2 * KC * (2 * tc) * G::SW * 4; }
};
// One update pass of the triangular solve: rows [I0, I0 + ni) of columns [c0, c1) receive
// b_ic -= sum_{klo <= k < I0} l_ik x_kc.  Tile shape (16 x TC) or, `wide`, (8 x 2 TC) with
// TC = blockDim / 16; operands staged by TMA in chunks of KC, double buffered (s_a: KC x 16
// elements per buffer, s_b: KC x 2 TC).  `it` counts the chunks staged so far by this CTA.
template <int NL>
__device__ __forceinline__ void
trsm_update_pass(const TrsmTileDesc &d, int I0, int ni, bool wide, int c0, int c1, int klo, uint64_t *bar,
                 uint32_t *s_a, uint32_t *s_b, uint32_t &it)
{
  typedef TileGeom<NL> G;
  const int nth = blockDim.x, TC = nth >> 4, BC = 2 * TC, t = threadIdx.x;
  const int rows = wide ? TS / 2 : TS, cols = wide ? BC : TC;
  const int ti = wide ? (t & 7) : (t & 15), tj = wide ? (t >> 3) : (t >> 4);
  const int nc = min(cols, c1 - c0), K = I0 - klo;
  const bool active = ti < ni && tj < nc;
  Reg<NL> acc;
  const TrsmView v(d);
  uint64_t *mine = d.B + ((long)(c0 + tj) * v.bsc + (long)(I0 + ti) * v.bsi) * G::ES;
  if(active)
    ldg_reg<NL>(acc, mine);
  else
    mpfw::set_zero(acc);
  const uint64_t *Abase = d.L + ((long)I0 * v.lsi + (long)klo * v.lsk) * G::ES; // (x, k) -> x lsi + k lsk
  const uint64_t *Bbase = d.B + ((long)c0 * v.bsc + (long)klo * v.bsi) * G::ES; // (x, k) -> x bsc + k bsi
  const int nchunks = (K + KC - 1) / KC;
  auto issue = [&](int c) {
    const uint32_t s = (it + c) & 1;
    const int k0 = c * KC, kcnt = min(KC, K - k0);
    if(t == 0)
      mbar_expect_tx(&bar[s], (uint32_t)(G::EB * kcnt * (ni + nc)));
    // copies of this chunk: kcnt x ni elements of L, kcnt x nc elements of X
    for(int e = t; e < KC * (rows + cols); e += nth)
      {
        const bool isA = e < KC * rows;
        const int q = isA ? e : e - KC * rows;
        const int dk = isA ? q / rows : q / cols, dx = isA ? q % rows : q % cols;
        if(dk >= kcnt || dx >= (isA ? ni : nc))
          continue;
        if(isA)
          bulk_g2s(s_a + ((size_t)s * KC * TS + dk * TS + dx) * G::SW,
                   Abase + ((long)dx * v.lsi + (long)(k0 + dk) * v.lsk) * G::ES, G::EB, &bar[s]);
        else
          bulk_g2s(s_b + ((size_t)s * KC * BC + dk * BC + dx) * G::SW,
                   Bbase + ((long)dx * v.bsc + (long)(k0 + dk) * v.bsi) * G::ES, G::EB, &bar[s]);
      }
  };
  issue(0);
  for(int c = 0; c < nchunks; ++c)
    {
      if(c + 1 < nchunks)
        issue(c + 1);
      const uint32_t s = (it + c) & 1, parity = ((it + c) >> 1) & 1;
      while(!mbar_try_wait(&bar[s], parity))
        {
        }
      const int kcnt = min(KC, K - c * KC);
      if(active)
        {
          const uint32_t *pa = s_a + ((size_t)s * KC * TS + ti) * G::SW;
          const uint32_t *pb = s_b + ((size_t)s * KC * BC + tj) * G::SW;
          for(int kk = 0; kk < kcnt; ++kk)
            acc = mac_ss_nl<NL>(acc, pa + (size_t)kk * TS * G::SW, pb + (size_t)kk * BC * G::SW, true);
        }
      __syncthreads();
    }
  it += nchunks;
  if(active)
    stg_reg<NL>(mine, acc);
}

// Level kernel with the two tile shapes of the walk kernel: CTAs whose row tile has more than 8
// rows update a (16 x 16) tile, CTAs on a last tile of <= 8 rows an (8 x 32) one (half as many
// CTAs, all 256 threads busy: a 40-row block is 2.5 tiles).  grid.y counts 16-column tiles.
template <int NL>
__global__ void __launch_bounds__(256, TileOcc<NL>::value) trsm_gemm_level2(const TrsmTileDesc *descs, int It)
{
  typedef TileGeom<NL> G;
  typedef WalkGeom<NL> WG;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
  uint32_t *s_a = reinterpret_cast<uint32_t *>(smem_raw + WG::OFF_A);
  uint32_t *s_b = reinterpret_cast<uint32_t *>(smem_raw + WG::OFF_B);
  const TrsmTileDesc d = descs[blockIdx.x];
  const int I0 = It * TS;
  if(I0 >= d.p)
    return;
  const int ni = min(TS, d.p - I0);
  const bool wide = ni <= TS / 2;
  const int cols = wide ? 2 * TS : TS;
  const int c0 = blockIdx.y * cols;
  if(c0 >= d.ncols)
    return;
  const int klo = d.hb ? ((c0 / d.nb) * d.hb) & ~(KC - 1) : 0;
  if(I0 <= klo)
    return; // nothing above this tile can be non-zero: no update
  if(threadIdx.x == 0)
    {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  __syncthreads();
  uint32_t it = 0;
  trsm_update_pass<NL>(d, I0, ni, wide, c0, min(d.ncols, c0 + cols), klo, bar, s_a, s_b, it);
}

template <int NL>
__global__ void __launch_bounds__(256, TileOcc<NL>::value)
trsm_walk_kernel(const TrsmTileDesc *descs, int ncg, int Wc)
{
  typedef TileGeom<NL> G;
  typedef WalkGeom<NL> WG;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(smem_raw);
  uint32_t *s_diag = reinterpret_cast<uint32_t *>(smem_raw + WG::OFF_DIAG);   // packed lower triangle
  uint32_t *s_recip = reinterpret_cast<uint32_t *>(smem_raw + WG::OFF_RECIP);
  uint32_t *s_a = reinterpret_cast<uint32_t *>(smem_raw + WG::OFF_A);
  uint32_t *s_b = reinterpret_cast<uint32_t *>(smem_raw + WG::OFF_B);
  const int nth = blockDim.x, TC = nth >> 4, BC = 2 * TC, t = threadIdx.x;
  const TrsmTileDesc d = descs[blockIdx.x / ncg];
  const int cg = blockIdx.x % ncg;
  const int c_lo = cg * Wc, c_hi = min(d.ncols, c_lo + Wc);
  if(c_lo >= c_hi || d.p == 0)
    return;
  if(t == 0)
    {
      mbar_init(&bar[0], 1);
      mbar_init(&bar[1], 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
  __syncthreads();
  uint32_t it = 0; // chunks staged so far: buffer and mbarrier phase
  const int T = (d.p + TS - 1) / TS;
  for(int It = 0; It < T; ++It)
    {
      const int I0 = It * TS, ni = min(TS, d.p - I0);
      // ---- update: rows of this tile against everything solved above it
      const bool wide = ni <= TS / 2;               // (8 x 2TC) tiles
      const int cols = wide ? BC : TC;
      for(int c0 = c_lo; c0 < c_hi && I0 > 0; c0 += cols)
        {
          // leading zeros of these columns (bases_blocks structure): rows above klo are exact zeros
          const int klo = d.hb ? ((c0 / d.nb) * d.hb) & ~(KC - 1) : 0;
          if(I0 <= klo)
            continue;
          trsm_update_pass<NL>(d, I0, ni, wide, c0, c_hi, klo, bar, s_a, s_b, it);
        }
      // ---- the factored diagonal tile (packed lower triangle) and its reciprocals
      for(int e = t; e < TS * TS; e += nth)
        {
          const int x = e & (TS - 1), k = e >> 4;
          if(x < ni && k <= x)
            {
              const uint4 *src = reinterpret_cast<const uint4 *>(d.L + ((long)(I0 + x) + (long)(I0 + k) * d.p) * G::ES);
              uint4 *dst = reinterpret_cast<uint4 *>(s_diag + (size_t)tri_index(x, k) * G::SW);
#pragma unroll
              for(int w = 0; w < G::EB / 16; ++w)
                dst[w] = src[w];
            }
        }
      for(int w = t; w < ni * G::RS; w += nth)
        s_recip[w] = d.recip[(long)I0 * G::RS + w];
      __syncthreads(); // also: the updates stored above are visible to the column owners below
      // ---- solve: one thread per column, rows top to bottom
      for(int col = c_lo + t; col < c_hi; col += nth)
        {
          uint64_t *colp = d.B + ((long)col * d.p + I0) * G::ES;
          for(int ii = 0; ii < ni; ++ii)
            {
              Reg<NL> acc;
              ldg_reg<NL>(acc, colp + (long)ii * G::ES);
              for(int kk = 0; kk < ii; ++kk)
                acc = mac_nl<NL>(acc, s_diag + (size_t)tri_index(ii, kk) * G::SW,
                                 reinterpret_cast<const uint32_t *>(colp + (long)kk * G::ES), true);
              acc = div_nl<NL>(acc, s_diag + (size_t)tri_index(ii, ii) * G::SW, s_recip + ii * G::RS);
              stg_reg<NL>(colp + (long)ii * G::ES, acc);
            }
        }
      // the rows just solved are TMA sources of every later update of this CTA
      fence_async_proxy();
      __syncthreads();
    }
}
} // namespace sdpb_b200

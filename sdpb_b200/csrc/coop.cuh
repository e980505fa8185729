// Warp-cooperative big-integer arithmetic for the Cholesky pivots.
//
// Every column of a Cholesky factorisation starts with l_kk = sqrt(a_kk) and
// the reciprocal of l_kk that the divisions of that column use.  Both are
// Newton iterations on ~30-word integers; done by one thread (mpfw::sqrt_fast,
// mpfw::reciprocal_fast) they are ~12 500 dependent instructions of fully
// unrolled code (200 KB: it does not even fit the instruction cache), ~37 us
// per pivot, and the pivots of a matrix form a serial chain (profiles/
// r01_v5_summary.md: 36 ms of a 146 ms step).  Here the 32 lanes of one warp
// share each multi-word operation: numbers live in shared memory as arrays of
// 32-bit words, a product is formed one COLUMN per lane (a 96-bit column sum,
// then a carry resolution across lanes), additions and shifts one word per
// lane.  The code is loops of runtime length -- a few hundred instructions --
// and the dependent chain per lane is ~5x shorter.
//
// The algorithms are those of mpfw.h (same Newton ladders, same exact final
// corrections), so the results are the exact integer square root and the exact
// floor reciprocal, i.e. identical to mpfx::sqrt / mpfw::reciprocal and to
// GMP's mpf_sqrt / mpf_div.  If a correction does not close (it never has:
// the fuzz tests count the fallbacks) one lane runs the slow exact routine.
#pragma once
#include "mpfw.h"

namespace coop
{
constexpr unsigned FULL = 0xFFFFFFFFu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// ---- word-parallel helpers (n <= 96 words; every lane takes words lane, lane+32, ...)
__device__ __forceinline__ void wcopy(uint32_t *dst, const uint32_t *src, int n)
{
  for(int i = lane_id(); i < n; i += 32)
    dst[i] = src[i];
  __syncwarp();
}
__device__ __forceinline__ void wzero(uint32_t *dst, int n)
{
  for(int i = lane_id(); i < n; i += 32)
    dst[i] = 0;
  __syncwarp();
}

// Carries of a 32-word chunk in one step.  g: lanes whose word overflowed, p: lanes whose
// word is all ones (a carry coming in passes through), cin: carry into lane 0.  A carry is
// injected above every g lane; adding the inject mask to the propagate mask lets the integer
// adder ripple it through each run of p lanes.  Bit i of the result = carry into lane i,
// bit 32 = carry out of the chunk.  (g and p exclude each other: x + y = 2^32 + 0xFFFFFFFF
// is impossible.)
__device__ __forceinline__ uint64_t carry_lookahead(uint32_t g, uint32_t p, uint32_t cin)
{
  const uint64_t inject = ((uint64_t)g << 1) | cin, prop = p;
  return inject | ((prop + inject) ^ prop ^ inject);
}

// r = a + b + cin (n words, r may alias a or b); returns the carry out (0/1), uniform.
__device__ __forceinline__ uint32_t wadd(uint32_t *r, const uint32_t *a, const uint32_t *b, int n,
                                         uint32_t cin = 0)
{
  const int l = lane_id();
  uint32_t carry_in = cin;
  for(int base = 0; base < n; base += 32)
    {
      const int i = base + l;
      const bool on = i < n;
      const uint32_t x = on ? a[i] : 0u, y = on ? b[i] : 0u;
      uint32_t s = x + y;
      const uint32_t g = __ballot_sync(FULL, s < x), p = __ballot_sync(FULL, s == 0xFFFFFFFFu);
      const uint64_t c = carry_lookahead(g, p, carry_in);
      s += (uint32_t)(c >> l) & 1u;
      __syncwarp();
      if(on)
        r[i] = s;
      // lanes beyond n hold zeros: the carry leaving word n-1 is the one entering lane n - base
      const int last = n - base < 32 ? n - base : 32;
      carry_in = (uint32_t)(c >> last) & 1u;
    }
  __syncwarp();
  return carry_in;
}
// r = a - b (n words); returns the borrow out (0/1).  a - b = a + ~b + 1.
__device__ __forceinline__ uint32_t wsub(uint32_t *r, const uint32_t *a, const uint32_t *b, int n,
                                         uint32_t *scratch)
{
  for(int i = lane_id(); i < n; i += 32)
    scratch[i] = ~b[i];
  __syncwarp();
  return 1u - wadd(r, a, scratch, n, 1u);
}
// r = -r (two's complement over n words)
__device__ __forceinline__ void wneg(uint32_t *r, int n, uint32_t *scratch, uint32_t *zero)
{
  wsub(r, zero, r, n, scratch);
}
// r -= k (small constant), n words
__device__ __forceinline__ void wsub_small(uint32_t *r, int n, uint32_t k, uint32_t *scratch,
                                           uint32_t *scratch2)
{
  __syncwarp();
  if(r[0] >= k) // no borrow leaves word 0 (all but 2^-30 of the time)
    {
      __syncwarp();
      if(lane_id() == 0)
        r[0] -= k;
      __syncwarp();
      return;
    }
  for(int i = lane_id(); i < n; i += 32)
    scratch2[i] = i == 0 ? k : 0u;
  __syncwarp();
  wsub(r, r, scratch2, n, scratch);
}
// dst[0..nd) = (src >> sh) (src has ns words, zeros beyond); dst must not overlap src
__device__ __forceinline__ void wshr(uint32_t *dst, int nd, const uint32_t *src, int ns, int sh)
{
  const int ws = sh >> 5, bs = sh & 31;
  for(int i = lane_id(); i < nd; i += 32)
    {
      const int a = i + ws;
      const uint32_t lo = a < ns ? src[a] : 0u, hi = a + 1 < ns ? src[a + 1] : 0u;
      dst[i] = bs ? (lo >> bs) | (hi << (32 - bs)) : lo;
    }
  __syncwarp();
}
// dst[0..nd) = (src << sh), src has ns words; bits shifted beyond nd words are dropped
__device__ __forceinline__ void wshl(uint32_t *dst, int nd, const uint32_t *src, int ns, int sh)
{
  const int ws = sh >> 5, bs = sh & 31;
  for(int i = lane_id(); i < nd; i += 32)
    {
      const int a = i - ws;
      const uint32_t hi = (a >= 0 && a < ns) ? src[a] : 0u;
      const uint32_t lo = (a - 1 >= 0 && a - 1 < ns) ? src[a - 1] : 0u;
      dst[i] = bs ? (hi << bs) | (lo >> (32 - bs)) : hi;
    }
  __syncwarp();
}
// number of leading zero bits of the n-word integer (32 n for zero)
__device__ __forceinline__ int wclz(const uint32_t *a, int n)
{
  int best = -1; // highest non-zero word
  for(int base = 0; base < n; base += 32)
    {
      const int i = base + lane_id();
      const unsigned m = __ballot_sync(FULL, i < n && a[i] != 0);
      if(m)
        best = base + 31 - __clz(m);
    }
  if(best < 0)
    return 32 * n;
  return 32 * (n - 1 - best) + __clz(a[best]);
}

// out[c - FROM] = word c, c in [FROM, TO), of  sum_{c' = C0}^{TO-1} (sum_{i+j=c'} a_i b_j) beta^c'
// (columns below C0 are not formed, the carry out of column TO-1 is dropped):
// the one primitive behind mpfw's mul_low / mul_mid / mul_high.
// t: scratch, 3 * (TO - C0) words.  out must not alias a or b.
// (inlined at its dozen call sites on purpose: out of line -- measured, profiles/r02_summary.md --
// the kernel shrinks from 32 000 to 16 000 instructions and instruction-fetch stalls halve, but the
// call sites lose their compile-time column ranges and the pivot gets 10 % slower)
__device__ __forceinline__ void wmul(uint32_t *out, const uint32_t *a, int KA, const uint32_t *b,
                                     int KB, int C0, int FROM, int TO, uint32_t *t)
{
  const int l = lane_id();
  const int ncols = TO - C0;
  uint32_t *t0 = t, *t1 = t + ncols, *t2 = t + 2 * ncols;
  for(int base = 0; base < ncols; base += 32)
    {
      const int k = base + l, c = C0 + k;
      uint32_t s0 = 0, s1 = 0, s2 = 0;
      // One trip count for the whole warp -- the rows i that ANY column of this round needs -- so
      // that a[i] is a broadcast load at a uniform address and the loop carries no per-lane
      // bounds; a lane whose column does not reach row i adds a zero product.
      const int cmin = C0 + base, cmax = C0 + (base + 32 < ncols ? base + 32 : ncols) - 1;
      const int ulo = cmin - (KB - 1) > 0 ? cmin - (KB - 1) : 0;
      const int uhi = cmax < KA - 1 ? cmax : KA - 1;
      const bool on = k < ncols;
#pragma unroll 4
      for(int i = ulo; i <= uhi; ++i)
        {
          const int j = c - i;
          const uint32_t bv = (on && j >= 0 && j < KB) ? b[j] : 0u;
          mpfw::mac3(s0, s1, s2, a[i], bv);
        }
      if(on)
        {
          t0[k] = s0;
          t1[k] = s1;
          t2[k] = s2;
        }
    }
  __syncwarp();
  // word c = t0[c] + t1[c-1] + t2[c-2] + carries
  uint32_t carry_in = 0;
  for(int base = 0; base < ncols; base += 32)
    {
      const int k = base + l;
      const bool on = k < ncols;
      uint64_t s = 0;
      if(on)
        {
          s = t0[k];
          if(k >= 1)
            s += t1[k - 1];
          if(k >= 2)
            s += t2[k - 2];
        }
      if(l == 0)
        s += carry_in;
      uint32_t w = (uint32_t)s;
      const uint32_t cy = (uint32_t)(s >> 32); // <= 3: goes to the next lane
      uint32_t in = __shfl_up_sync(FULL, cy, 1);
      if(l == 0)
        in = 0;
      const uint32_t w2 = w + in;
      const uint32_t g = __ballot_sync(FULL, w2 < w), p = __ballot_sync(FULL, w2 == 0xFFFFFFFFu);
      const uint64_t c = carry_lookahead(g, p, 0u);
      w = w2 + ((uint32_t)(c >> l) & 1u);
      const uint32_t top = __shfl_sync(FULL, cy, 31) + ((uint32_t)(c >> 32) & 1u);
      if(on && C0 + k >= FROM)
        out[C0 + k - FROM] = w;
      carry_in = top;
    }
  __syncwarp();
}

// ------------------------------------------------------------------ workspace
template <int NL> struct alignas(16) Work
{
  static constexpr int P = NL - 1;
  static constexpr int NT = 4 * P;       // words of the sqrt radicand frame
  static constexpr int NR = 2 * P;       // words of the root
  static constexpr int WFS = 2 * P + 3;  // rsqrt iterate, words
  static constexpr int N2 = 2 * NL;      // divisor words
  static constexpr int WFR = 2 * NL + 5; // reciprocal iterate, words
  static constexpr int MAXW = (NT > WFR + 4 ? NT : WFR + 4) + 4;
  uint32_t T[NT], Tn[NT];
  uint32_t A[2 * MAXW], B[2 * MAXW], C[2 * MAXW], D[2 * MAXW], E[2 * MAXW];
  uint32_t V[MAXW], Vp[MAXW];
  uint32_t S[MAXW], zero[2 * MAXW], scratch[2 * MAXW], scratch2[2 * MAXW];
  uint32_t t[3 * 2 * MAXW];
  uint32_t flag;
};

// first ladder entry <= maxw (the ladder is W -> (W <= 2) ? 1 : (W + 2) / 2)
template <int W, int MAXW, bool DONE = (W <= MAXW)> struct LadderEntry
{
  static constexpr int value = LadderEntry<((W <= 2) ? 1 : (W + 2) / 2), MAXW>::value;
};
template <int W, int MAXW> struct LadderEntry<W, MAXW, true>
{
  static constexpr int value = W;
};
// The low rungs of a ladder are a handful of words: one lane climbs them with the
// register code of mpfw.h (a few hundred instructions), the warp takes over from there.
constexpr int SOLO_MAX = 9;

// ladder sizes: W -> WP = (W <= 2) ? 1 : (W + 2) / 2, down to 1
__device__ __forceinline__ int ladder(int WF, int (&lv)[12])
{
  int n = 0, w = WF;
  lv[n++] = w;
  while(w > 1)
    {
      w = (w <= 2) ? 1 : (w + 2) / 2;
      lv[n++] = w;
    }
  return n; // lv[n-1] == 1
}

// S (NR words, in ws.S) = floor(sqrt(T)), T = u's mantissa in the 2P-limb frame of
// mpf_sqrt; returns false (uniformly) if the fast path failed to close.
// Leaves V ~ beta^WFS / (2 sqrt(Tn / beta^NT)) in ws.V and the shift in *shift_out.
template <int NL>
__device__ __forceinline__ bool sqrt_words(Work<NL> &ws, const uint32_t *uw, int expodd)
{
  typedef Work<NL> G;
  const int l = lane_id();
  // T: u's 2NL words top-aligned (expodd: one limb lower) in NT words
  for(int i = l; i < G::NT; i += 32)
    {
      const int a = i - (G::NT - 2 * NL) + (expodd ? 2 : 0);
      ws.T[i] = (a >= 0 && a < 2 * NL) ? uw[a] : 0u;
    }
  wzero(ws.zero, 2 * G::MAXW);
  const int s = wclz(ws.T, G::NT) & ~1;
  wshl(ws.Tn, G::NT, ws.T, G::NT, s);
  int lv[12];
  const int nlv = ladder(G::WFS, lv);
  // rungs up to WS words: lane 0 alone (mpfw::RsqrtLevel on the top WS words)
  constexpr int WS = LadderEntry<G::WFS, SOLO_MAX>::value;
  if(l == 0)
    {
      uint32_t top[WS], v[WS];
#pragma unroll
      for(int i = 0; i < WS; ++i)
        top[i] = ws.Tn[G::NT - WS + i];
      mpfw::RsqrtLevel<WS, WS>::run(v, top);
#pragma unroll
      for(int i = 0; i < WS; ++i)
        ws.V[i] = v[i];
    }
  __syncwarp();
  for(int q = nlv - 2; q >= 0; --q)
    {
      if(lv[q] <= WS)
        continue;
      const int W = lv[q], WP = lv[q + 1];
      wcopy(ws.Vp, ws.V, WP);
      const uint32_t *Tt = ws.Tn + (G::NT - W); // W <= NT always here
      // V2 = Vp^2 (exact, 2 WP words)
      wmul(ws.A, ws.Vp, WP, ws.Vp, WP, 0, 0, 2 * WP, ws.t);
      const int GL = 2 * WP - 3 > 0 ? 2 * WP - 3 : 0;
      const int NG = WP + W + 1 - GL;
      // G = words [GL, WP+W+1) of V2 * Tt, columns from GL-2
      wmul(ws.B, ws.A, 2 * WP, Tt, W, GL - 2 > 0 ? GL - 2 : 0, GL, WP + W + 1, ws.t);
      wshl(ws.C, NG, ws.B, NG, 2);
      wsub(ws.C, ws.zero, ws.C, NG, ws.scratch); // G = -4F mod beta^NG
      const int QF = 3 * WP - GL - 1, NQ = WP + NG - QF;
      wmul(ws.D, ws.Vp, WP, ws.C, NG, QF - 2 > 0 ? QF - 2 : 0, QF, WP + NG, ws.t);
      wshr(ws.E, NQ, ws.D, NQ, 1);
      // V = Vp beta^(W-WP) + Q[1..]
      for(int i = l; i <= W; i += 32)
        {
          ws.A[i] = (i >= W - WP && i < W) ? ws.Vp[i - (W - WP)] : 0u;
          ws.B[i] = (i + 1 < NQ) ? ws.E[i + 1] : 0u;
        }
      __syncwarp();
      wadd(ws.S, ws.A, ws.B, W + 1);
      const bool sat = ws.S[W] != 0;
      for(int i = l; i < W; i += 32)
        ws.V[i] = sat ? 0xFFFFFFFFu : ws.S[i];
      __syncwarp();
      wsub_small(ws.V, W, 4u, ws.scratch, ws.scratch2);
    }
  // root: Sx = words [WFS+2, NR+3+WFS) of Tt * V, Tt = top NR+3 words of Tn
  const uint32_t *Tt = ws.Tn + (G::NT - (G::NR + 3));
  wmul(ws.A, Tt, G::NR + 3, ws.V, G::WFS, G::WFS, G::WFS + 2, G::NR + 3 + G::WFS, ws.t);
  wshr(ws.S, G::NR, ws.A, G::NR + 1, 31 + (s >> 1));
  wsub_small(ws.S, G::NR, 2u, ws.scratch, ws.scratch2);
  // rem = T - S^2 (low NR+2 words)
  wmul(ws.A, ws.S, G::NR, ws.S, G::NR, 0, 0, G::NR + 2, ws.t);
  wsub(ws.B, ws.T, ws.A, G::NR + 2, ws.scratch); // rem in B
  bool ok = false;
  for(int round = 0; round < 8 && !ok; ++round)
    {
      // step = 2 S + 1 (NR+2 words)
      for(int i = l; i < G::NR + 2; i += 32)
        {
          const uint32_t lo = i < G::NR ? ws.S[i] : 0u;
          const uint32_t below = (i > 0 && i - 1 < G::NR) ? ws.S[i - 1] : 0u;
          ws.C[i] = (lo << 1) | (below >> 31) | (i == 0 ? 1u : 0u);
        }
      __syncwarp();
      const uint32_t bw = wsub(ws.D, ws.B, ws.C, G::NR + 2, ws.scratch);
      if(bw)
        ok = true;
      else
        {
          wcopy(ws.B, ws.D, G::NR + 2);
          for(int i = l; i < G::NR; i += 32)
            ws.C[i] = i == 0 ? 1u : 0u;
          __syncwarp();
          wadd(ws.S, ws.S, ws.C, G::NR);
        }
    }
  return ok && ws.B[G::NR + 1] == 0;
}

// R (N2+4 words, into Rout, shared or global) = floor(beta^(4 NL + 1) / D), D = dw[0..N2)
// seeded: ws.V holds the rsqrt iterate of the square root D was just taken of (sqrt_words): with
// d = Dn / beta^N2 = sqrt(t) up to the root's floor, beta^W / (2 d) IS beta^W / (2 sqrt(t)), so the
// top words of V seed the LAST rung directly and the lower rungs (a single lane's register code up
// to 9 words, then the 17-word rung) are skipped.  The exact final correction decides as before.
template <int NL>
__device__ __forceinline__ bool recip_words(Work<NL> &ws, const uint32_t *dw, uint32_t *Rout, bool seeded = false)
{
  typedef Work<NL> G;
  const int l = lane_id();
  constexpr int n2 = G::N2, WF = G::WFR;
  wzero(ws.zero, 2 * G::MAXW);
  const int s = wclz(dw, n2); // < 64
  wshl(ws.Tn, n2, dw, n2, s); // Dn
  int lv[12];
  const int nlv = ladder(WF, lv);
  constexpr int WS = LadderEntry<WF, SOLO_MAX>::value;
  const int WSEED = lv[1]; // the last rung's WP
  if(seeded)
    {
      // V (WFS words) -> its top WSEED words, less a few units (the rung wants an iterate from below)
      wcopy(ws.Vp, ws.V + (G::WFS - WSEED), WSEED);
      wcopy(ws.V, ws.Vp, WSEED);
      wsub_small(ws.V, WSEED, 8u, ws.scratch, ws.scratch2);
    }
  else if(l == 0)
    {
      uint32_t top[WS], v[WS];
#pragma unroll
      for(int i = 0; i < WS; ++i)
        top[i] = ws.Tn[n2 - WS + i];
      mpfw::RecipLevel<WS, WS>::run(v, top);
#pragma unroll
      for(int i = 0; i < WS; ++i)
        ws.V[i] = v[i];
    }
  __syncwarp();
  for(int q = seeded ? 0 : nlv - 2; q >= 0; --q)
    {
      if(lv[q] <= WS && !seeded)
        continue;
      const int W = lv[q], WP = lv[q + 1];
      wcopy(ws.Vp, ws.V, WP);
      // Dt = top W words of Dn (zero-extended below when W > n2)
      for(int i = l; i < W; i += 32)
        ws.E[i] = (n2 - W + i >= 0) ? ws.Tn[n2 - W + i] : 0u;
      __syncwarp();
      // G = -(2 Dt Yp) mod beta^(W+1)
      wmul(ws.A, ws.E, W, ws.Vp, WP, 0, 0, W + 1, ws.t);
      wshl(ws.B, W + 1, ws.A, W + 1, 1);
      wsub(ws.B, ws.zero, ws.B, W + 1, ws.scratch);
      // Q = words [2WP, WP+W+1) of Yp * G
      wmul(ws.C, ws.Vp, WP, ws.B, W + 1, 2 * WP - 2 > 0 ? 2 * WP - 2 : 0, 2 * WP, WP + W + 1, ws.t);
      for(int i = l; i <= W; i += 32)
        {
          ws.A[i] = (i >= W - WP && i < W) ? ws.Vp[i - (W - WP)] : 0u;
          ws.D[i] = (i < W - WP + 1) ? ws.C[i] : 0u;
        }
      __syncwarp();
      wadd(ws.S, ws.A, ws.D, W + 1);
      const bool sat = ws.S[W] != 0;
      for(int i = l; i < W; i += 32)
        ws.V[i] = sat ? 0xFFFFFFFFu : ws.S[i];
      __syncwarp();
      wsub_small(ws.V, W, 4u, ws.scratch, ws.scratch2);
    }
  // R~ = Y >> (127 - s); Rt = its low n2+4 words
  wshr(ws.S, n2 + 4, ws.V, WF, 127 - s);
  // rem = -(Rt * D) mod beta^(n2+2)
  wmul(ws.A, ws.S, n2 + 4, dw, n2, 0, 0, n2 + 2, ws.t);
  wsub(ws.B, ws.zero, ws.A, n2 + 2, ws.scratch);
  for(int i = l; i < n2 + 2; i += 32)
    ws.C[i] = i < n2 ? dw[i] : 0u; // dext
  __syncwarp();
  bool ok = false;
  for(int round = 0; round < 4 && !ok; ++round)
    {
      const uint32_t bw = wsub(ws.D, ws.B, ws.C, n2 + 2, ws.scratch);
      if(bw)
        ok = true;
      else
        {
          wcopy(ws.B, ws.D, n2 + 2);
          for(int i = l; i < n2 + 4; i += 32)
            ws.E[i] = i == 0 ? 1u : 0u;
          __syncwarp();
          wadd(ws.S, ws.S, ws.E, n2 + 4);
        }
    }
  for(int i = l; i < n2 + 4; i += 32)
    Rout[i] = ws.S[i];
  __syncwarp();
  return ok;
}

// a <- mpf_sqrt(a) in place, by one whole warp: a > 0 is a packed element in shared memory
// (header + 2NL words)
// returns true (uniformly) if the fast path closed, i.e. ws.V holds the rsqrt iterate of a
template <int NL> __device__ __forceinline__ bool sqrt_elem(Work<NL> &ws, uint32_t *a)
{
  typedef Work<NL> G;
  const int l = lane_id();
  const int32_t aexp = (int32_t)a[0];
  const int expodd = aexp & 1;
  __syncwarp();
  const bool ok_sqrt = sqrt_words<NL>(ws, a + 2, expodd);
  if(ok_sqrt)
    {
      // l = [0, 0, S] with exponent (aexp + expodd) / 2
      for(int i = l; i < 2 * NL; i += 32)
        a[2 + i] = i < 2 ? 0u : ws.S[i - 2];
      if(l == 0)
        {
          a[0] = (uint32_t)((aexp + expodd) / 2);
          a[1] = 1u;
        }
    }
  else if(l == 0) // never taken so far: the generic exact routine
    {
      ws.flag += 1;
      mpfw::Reg<NL> u;
      mpfw::load<NL>(u, a);
      mpfx::Num<NL> x, y;
      mpfw::to_num(x, u);
      mpfx::sqrt(y, x);
      mpfw::from_num(u, y);
      mpfw::store<NL>(a, u);
    }
  __syncwarp();
  return ok_sqrt;
}
// Rs / Rg (shared / global, RW words; either may be null) <- the reciprocal words of the packed
// element a != 0 (shared memory) that mpfw::div_recip takes, by one whole warp
template <int NL>
__device__ __forceinline__ void recip_elem(Work<NL> &ws, const uint32_t *a, uint32_t *Rs, uint32_t *Rg,
                                           bool seeded = false)
{
  typedef Work<NL> G;
  const int l = lane_id();
  uint32_t *R = ws.E + G::MAXW;
  __syncwarp();
  bool ok_recip = recip_words<NL>(ws, a + 2, R, seeded);
  if(!ok_recip && seeded) // uniform; the seed was off (it never is for d = sqrt(t)): the full ladder
    {
      if(l == 0)
        ws.flag += 1; // the tests count this like any other fallback
      __syncwarp();
      ok_recip = recip_words<NL>(ws, a + 2, R, false);
    }
  if(!ok_recip && l == 0)
    {
      ws.flag += 1;
      mpfw::Reg<NL> u;
      mpfw::load<NL>(u, a);
      mpfx::Num<NL> x;
      mpfw::to_num(x, u);
      uint32_t Rw[2 * NL + 4];
      mpfw::reciprocal<NL>(Rw, x);
      for(int i = 0; i < 2 * NL + 4; ++i)
        R[i] = Rw[i];
    }
  __syncwarp();
  for(int i = l; i < 2 * NL + 4; i += 32)
    {
      const uint32_t w = R[i];
      if(Rs)
        Rs[i] = w;
      if(Rg)
        Rg[i] = w;
    }
  __syncwarp();
}

// The pivot of one Cholesky column, by one whole warp: a > 0 is a packed element
// in shared memory at `a` (header + 2NL words).  On return the same slot holds
// l = mpf_sqrt(a) and Rs / Rg (shared / global, RW words; either may be null)
// hold the reciprocal words of l that mpfw::div_recip takes.
template <int NL>
__device__ __forceinline__ void pivot(Work<NL> &ws, uint32_t *a, uint32_t *Rs, uint32_t *Rg)
{
  // The seed is as good as d = sqrt(t) holds: to the NR = 2 NL - 2 words of the root.  The last rung
  // wants NL + 3 words, so the short precisions (NL < 7, below 320 bits) climb the whole ladder.
  constexpr bool SEED_OK = 2 * NL - 2 >= NL + 3 + 2;
  const bool seed = sqrt_elem<NL>(ws, a) && SEED_OK;
  recip_elem<NL>(ws, a, Rs, Rg, seed);
}
} // namespace coop

// solve_schur_complement_equation on the device (reference
// run/step/compute_search_direction/solve_schur_complement_equation.cxx:16-79).
//
// The factors the Schur step leaves in HBM -- L_j (arena S), the bands
// L_j^-1 B_j (arena Pband) and the upper factor of Q -- are consumed where they
// lie; only the right-hand side (P + N elements) crosses PCIe.  All of it is
// matrix-VECTOR work: 2 P N + sum P_j^2 + N^2 multiply-accumulates (2.7e7 at C3,
// 3 % of the step) but two chains of P_j and two of N dependent
// (multiply-subtract, divide) pairs, so the kernels are organised around the
// chains, not around throughput:
//
//   solve_tri_kernel   one CTA per triangular system (every L_j at once, or the
//                      one N x N factor of Q), the unknowns in shared memory;
//                      step k: the owner of row k divides by the pivot (exact
//                      reciprocal of the factorisation, mpfw::div_recip), one
//                      barrier, every other row subtracts its product -- the
//                      column-oriented substitution, in which an unknown
//                      receives its updates in the order the solved ones become
//                      available (forward: k ascending; backward: k descending)
//   solve_gemvT_kernel part_j[c] = -(sum_r P_j(r,c) x_j(r)), one thread per
//                      (block, column), rows ascending from an exact zero
//   solve_dysum_kernel dy[c] += sum_j part_j[c], the sum in the canonical two-level
//                      order over the GLOBAL blocks (the order is part of the result)
//   solve_gemv_kernel  x_j(r) += sum_c P_j(r,c) dy(c), one thread per stacked row
//
// Same canonical order as oracle/hotpath_core.hpp (schur_solve_forward / _Q /
// _backward); tests/test_parity_gpu.py compares byte for byte.
#pragma once
#include "kernels.cuh"

namespace sdpb_b200
{
struct SolveTriDesc // one triangular system  T x = b,  T(i,k) (i >= k) at A + (i*si + k*sj) elements
{
  const uint64_t *A;
  const uint32_t *recip; // reciprocals of the pivots T(k,k), stride TileGeom::RS words
  long si, sj;           // lower column-major: (1, p); upper factor U read as U^T: (p, 1)
  int p;
  long x0; // first element of this system's unknowns in the stacked vector
};

constexpr int SOLVE_MAX_THREADS = 512;

// BACK == false: T x = b (forward, k ascending); BACK == true: T^T x = b (k descending).
// xs: shared-memory copy of the unknowns when they fit (xs_stride words per element),
// else the substitution runs on the global vector itself.
template <int NL, bool BACK>
__global__ void __launch_bounds__(SOLVE_MAX_THREADS, 1)
solve_tri_kernel(const SolveTriDesc *descs, uint64_t *xg, int use_smem)
{
  typedef TileGeom<NL> G;
  extern __shared__ __align__(16) uint32_t solve_xs[];
  const SolveTriDesc d = descs[blockIdx.x];
  const int p = d.p, T = blockDim.x, tid = threadIdx.x;
  if(p == 0)
    return;
  uint32_t *xglob = reinterpret_cast<uint32_t *>(xg + d.x0 * G::ES);
  uint32_t *x = use_smem ? solve_xs : xglob;
  const int xs = use_smem ? G::SW : G::EW;
  if(use_smem)
    {
      for(int w = tid; w < p * G::EW; w += T)
        solve_xs[(w / G::EW) * G::SW + (w % G::EW)] = xglob[w];
      __syncthreads();
    }
  for(int step = 0; step < p; ++step)
    {
      const int k = BACK ? p - 1 - step : step;
      // rows this thread updates in this step: i == tid (mod T), i > k (forward) / i < k (backward)
      // their T(i,k) come from HBM: ask for the lines before the pivot's division hides the latency
      for(int i = tid; i < p; i += T)
        if(BACK ? i < k : i > k)
          {
            const uint64_t *t = BACK ? d.A + ((long)k * d.si + (long)i * d.sj) * G::ES
                                     : d.A + ((long)i * d.si + (long)k * d.sj) * G::ES;
            asm volatile("prefetch.global.L1 [%0];" ::"l"(t));
          }
      if(tid == k % T)
        {
          Reg<NL> v;
          mpfw::load<NL>(v, x + (long)k * xs);
          v = div_nl<NL>(v, reinterpret_cast<const uint32_t *>(d.A + ((long)k * (d.si + d.sj)) * G::ES),
                         d.recip + (long)k * G::RS);
          mpfw::store<NL>(x + (long)k * xs, v);
        }
      __syncthreads();
      for(int i = tid; i < p; i += T)
        if(BACK ? i < k : i > k)
          {
            const uint64_t *t = BACK ? d.A + ((long)k * d.si + (long)i * d.sj) * G::ES
                                     : d.A + ((long)i * d.si + (long)k * d.sj) * G::ES;
            Reg<NL> v;
            mpfw::load<NL>(v, x + (long)i * xs);
            v = mac_nl<NL>(v, x + (long)k * xs, reinterpret_cast<const uint32_t *>(t), true);
            mpfw::store<NL>(x + (long)i * xs, v);
          }
      // no second barrier: row k+1 (k-1) is divided by the thread that has just updated it, and
      // nobody else reads it before the next barrier; x_k itself is never written again
    }
  if(use_smem)
    {
      __syncthreads();
      for(int w = tid; w < p * G::EW; w += T)
        xglob[w] = solve_xs[(w / G::EW) * G::SW + (w % G::EW)];
    }
}

// part[gidx*N + c] = -(sum_r P(r,c) x(row0 + r)), rows ascending from an exact zero
template <int NL>
__global__ void __launch_bounds__(64, 8)
solve_gemvT_kernel(const BandDesc *bands, int N, const limb_t *x, limb_t *part)
{
  const BandDesc b = bands[blockIdx.x];
  const uint32_t *xb = reinterpret_cast<const uint32_t *>(x + (size_t)b.row0 * Fmt<NL>::ES);
  for(int c = blockIdx.y * blockDim.x + threadIdx.x; c < N; c += gridDim.y * blockDim.x)
    {
      Reg<NL> acc;
      mpfw::set_zero(acc);
      const uint32_t *col = reinterpret_cast<const uint32_t *>(b.P + (size_t)c * b.rows * Fmt<NL>::ES);
      for(int r = 0; r < b.rows; ++r)
        acc = mac_nl<NL>(acc, xb + (size_t)r * 2 * Fmt<NL>::ES, col + (size_t)r * 2 * Fmt<NL>::ES, false);
      acc.sign = -acc.sign;
      stg_reg<NL>(part + ((size_t)b.gidx * N + c) * Fmt<NL>::ES, acc);
    }
}

// dy[c] += sum_j part[j*N + c], the sum in the canonical two-level order over the GLOBAL
// blocks (ordered_block_sum, kernels.cuh); one warp per column
template <int NL>
__global__ void __launch_bounds__(32) solve_dysum_kernel(const limb_t *part, int J, int N, limb_t *dy)
{
  extern __shared__ __align__(16) uint32_t dysum_scratch[];
  const int c = blockIdx.x;
  if(c >= N)
    return;
  const Reg<NL> total = ordered_block_sum<NL>(part, J, N, c, dysum_scratch);
  if(threadIdx.x == 0)
    {
      Reg<NL> acc;
      ldg_reg<NL>(acc, dy + (size_t)c * Fmt<NL>::ES);
      acc = add_nl<NL>(acc, total);
      stg_reg<NL>(dy + (size_t)c * Fmt<NL>::ES, acc);
    }
}

// x(row0 + r) += sum_c P(r,c) dy(c), columns ascending from an exact zero
template <int NL>
__global__ void __launch_bounds__(64, 8)
solve_gemv_kernel(const BandDesc *bands, int N, const limb_t *dy, limb_t *x)
{
  const BandDesc b = bands[blockIdx.x];
  const uint32_t *y = reinterpret_cast<const uint32_t *>(dy);
  for(int r = blockIdx.y * blockDim.x + threadIdx.x; r < b.rows; r += gridDim.y * blockDim.x)
    {
      Reg<NL> acc, v;
      mpfw::set_zero(acc);
      const uint32_t *row = reinterpret_cast<const uint32_t *>(b.P + (size_t)r * Fmt<NL>::ES);
      for(int c = 0; c < N; ++c)
        acc = mac_nl<NL>(acc, y + (size_t)c * 2 * Fmt<NL>::ES, row + (size_t)c * b.rows * 2 * Fmt<NL>::ES, false);
      limb_t *xe = x + ((size_t)b.row0 + r) * Fmt<NL>::ES;
      ldg_reg<NL>(v, xe);
      v = add_nl<NL>(v, acc);
      stg_reg<NL>(xe, v);
    }
}

// epilogue of scale_multiply_add (scale_multiply_add.cxx:4-16): T holds sum_l A(i,l) B(l,j);
// C = alpha T (beta == 0) or C = (C * 1) + alpha T.  Multiplying by +-1 is NOT the identity in
// mpf: mpf_mul reads only the top `prec` limbs of its operands (mpf/mul.c), so a full
// (prec+1)-limb value loses its lowest limb -- in the top-aligned element format: the lowest
// stored limb becomes zero (nothing else changes: 1 has one limb, the product's top limb is
// zero and is stripped again).
template <int NL>
__global__ void __launch_bounds__(128)
sma_epilogue_kernel(const limb_t *T, limb_t *C, long count, int alpha, int beta)
{
  for(long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x)
    {
      Reg<NL> t;
      ldg_reg<NL>(t, T + e * Fmt<NL>::ES);
      t.w[0] = t.w[1] = 0; // acc *= alpha
      if(alpha < 0)
        t.sign = -t.sign;
      if(beta)
        {
          Reg<NL> c;
          ldg_reg<NL>(c, C + e * Fmt<NL>::ES);
          c.w[0] = c.w[1] = 0; // C *= beta
          c = add_nl<NL>(c, t);
          stg_reg<NL>(C + e * Fmt<NL>::ES, c);
        }
      else
        stg_reg<NL>(C + e * Fmt<NL>::ES, t);
    }
}
} // namespace sdpb_b200

// sm_100a kernels of the search direction, on block-diagonal matrices and
// vectors that stay in HBM between the calls of one Newton iteration:
//   compute_search_direction       run/step/compute_search_direction.cxx:44-90
//   cholesky_solve                 .../compute_search_direction/cholesky_solve.cxx:4-13
//   compute_schur_RHS              .../compute_search_direction/compute_schur_RHS.cxx:21-86
//   constraint_matrix_weighted_sum run/constraint_matrix_weighted_sum.cxx:14-66
//   symmetrize                     Block_Diagonal_Matrix.hxx:95-109
// plus the per-block pieces of mu, the R error and the corrector's Frobenius product
// (step.cxx:137-160).  The products C = alpha A B + beta C go through gemm_tile_kernel
// (tile.cuh); everything here is one thread (or one warp) per output, in the operation order
// csrc/host/direction.hpp spells out -- tests compare the two byte for byte.
//
// These stages are O(sum s^3) on s x s blocks with s = 20 ... 128: a few per cent of the Schur
// step.  They are written for residency (X, Y, dX, dY, R, Z never cross PCIe), not for the last
// per cent of the multiply pipe.
#pragma once
#include "kernels.cuh"

namespace sdpb_b200
{
struct BdmDesc // one block-parity of a block-diagonal matrix (the shape of X, Y, dX, dY, R, Z)
{
  long off;   // first element of the s x s block inside a block-diagonal object
  long voff;  // first element of bases_blocks[b] = I_m (x) basis (s rows): basis(a, k) = V[voff + k s + a]
  long x0;    // first stacked row of SDP block j in the K-vectors (x, dx, dual residues)
  int s, h, m, n;
  int cum_cols; // columns of all earlier block-parities (one thread per column launches)
};

__device__ __forceinline__ int bdm_find(const BdmDesc *d, int count, int col)
{
  int lo = 0, hi = count - 1;
  while(lo < hi)
    {
      const int mid = (lo + hi + 1) >> 1;
      if(d[mid].cum_cols <= col)
        lo = mid;
      else
        hi = mid - 1;
    }
  return lo;
}
template <int NL> __device__ __noinline__ Reg<NL> sub_nl(Reg<NL> acc, Reg<NL> v)
{
  mpfw::add_signed<NL>(acc, v, -v.sign);
  return acc;
}
template <int NL> __device__ __forceinline__ const uint32_t *elem32(const limb_t *base, long e)
{
  return reinterpret_cast<const uint32_t *>(base + e * Fmt<NL>::ES);
}
// |a| > |b| (both non-zero or zero), exact
template <int NL> __device__ __forceinline__ bool abs_greater(const Reg<NL> &a, const Reg<NL> &b)
{
  if(a.sign == 0)
    return false;
  if(b.sign == 0)
    return true;
  if(a.exp != b.exp)
    return a.exp > b.exp;
#pragma unroll
  for(int i = 2 * NL - 1; i >= 0; --i)
    if(a.w[i] != b.w[i])
      return a.w[i] > b.w[i];
  return false;
}

// op 0: C -= A   1: C += A   2: C = -C      (element-wise over a whole block-diagonal object)
template <int NL>
__global__ void __launch_bounds__(128) bdm_elementwise_kernel(int op, const limb_t *A, limb_t *C, long count)
{
  for(long e = (long)blockIdx.x * blockDim.x + threadIdx.x; e < count; e += (long)gridDim.x * blockDim.x)
    {
      Reg<NL> c;
      ldg_reg<NL>(c, C + e * Fmt<NL>::ES);
      if(op == 2)
        c.sign = -c.sign;
      else
        {
          Reg<NL> a;
          ldg_reg<NL>(a, A + e * Fmt<NL>::ES);
          c = op == 0 ? sub_nl<NL>(c, a) : add_nl<NL>(c, a);
        }
      stg_reg<NL>(C + e * Fmt<NL>::ES, c);
    }
}
// C(i,i) += v on every block (R = beta mu I - X Y)
template <int NL>
__global__ void __launch_bounds__(128)
bdm_add_diagonal_kernel(const BdmDesc *d, int count, int total_cols, limb_t *C, const limb_t *v)
{
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if(col >= total_cols)
    return;
  const BdmDesc b = d[bdm_find(d, count, col)];
  const int i = col - b.cum_cols;
  limb_t *p = C + (b.off + (long)i * b.s + i) * Fmt<NL>::ES;
  Reg<NL> c, x;
  ldg_reg<NL>(c, p);
  ldg_reg<NL>(x, v);
  c = add_nl<NL>(c, x);
  stg_reg<NL>(p, c);
}
// symmetrize: A *= 0.5 (an mpf_mul by the two-limb value mpf_set_d(0.5) gives), then
// off-diagonal A_ij = A_ji = A_ij + A_ji, diagonal A_ii += A_ii.  One thread per (i <= j).
template <int NL>
__global__ void __launch_bounds__(128)
bdm_symmetrize_kernel(const BdmDesc *d, int count, limb_t *A, const limb_t *half, int negate)
{
  const BdmDesc b = d[blockIdx.x];
  const long pairs = (long)b.s * (b.s + 1) / 2;
  for(long e = (long)blockIdx.y * blockDim.x + threadIdx.x; e < pairs; e += (long)gridDim.y * blockDim.x)
    {
      int j = 0; // e -> (i <= j), column-packed upper triangle: column j holds j + 1 entries
      long rem = e;
      {
        // j = floor((sqrt(8 e + 1) - 1) / 2), fixed up exactly
        j = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
        while((long)(j + 1) * (j + 2) / 2 <= e)
          ++j;
        while((long)j * (j + 1) / 2 > e)
          --j;
        rem = e - (long)j * (j + 1) / 2;
      }
      const int i = (int)rem;
      limb_t *pij = A + (b.off + (long)j * b.s + i) * Fmt<NL>::ES;
      limb_t *pji = A + (b.off + (long)i * b.s + j) * Fmt<NL>::ES;
      Reg<NL> x;
      ldg_reg<NL>(x, pij);
      x = mul_nl<NL>(x, reinterpret_cast<const uint32_t *>(half));
      Reg<NL> y = x;
      if(i != j)
        {
          ldg_reg<NL>(y, pji);
          y = mul_nl<NL>(y, reinterpret_cast<const uint32_t *>(half));
        }
      x = add_nl<NL>(x, y);
      if(negate)
        x.sign = -x.sign;
      stg_reg<NL>(pij, x);
      if(i != j)
        stg_reg<NL>(pji, x);
    }
}
// cholesky_solve: B <- L^{-1} B (TRANS == false, k ascending) or B <- L^{-T} B (k descending) on
// every s x s block; one thread per column of B.
template <int NL, bool TRANS>
__global__ void __launch_bounds__(64)
bdm_trsm_kernel(const BdmDesc *d, int count, int total_cols, const limb_t *L, const uint32_t *recip, limb_t *B)
{
  typedef TileGeom<NL> G;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if(col >= total_cols)
    return;
  const BdmDesc b = d[bdm_find(d, count, col)];
  const int c = col - b.cum_cols, s = b.s;
  limb_t *x = B + (b.off + (long)c * s) * Fmt<NL>::ES;
  const limb_t *Lb = L + b.off * Fmt<NL>::ES;
  const uint32_t *rc = recip + (long)b.cum_cols * G::RS; // reciprocals are stacked like the columns
  if(!TRANS)
    for(int i = 0; i < s; ++i)
      {
        Reg<NL> acc;
        ldg_reg<NL>(acc, x + (long)i * Fmt<NL>::ES);
        for(int k = 0; k < i; ++k)
          acc = mac_nl<NL>(acc, elem32<NL>(Lb, (long)k * s + i), elem32<NL>(x, k), true);
        acc = div_nl<NL>(acc, elem32<NL>(Lb, (long)i * s + i), rc + (long)i * G::RS);
        stg_reg<NL>(x + (long)i * Fmt<NL>::ES, acc);
      }
  else
    for(int i = s - 1; i >= 0; --i)
      {
        Reg<NL> acc;
        ldg_reg<NL>(acc, x + (long)i * Fmt<NL>::ES);
        for(int k = s - 1; k > i; --k)
          acc = mac_nl<NL>(acc, elem32<NL>(Lb, (long)i * s + k), elem32<NL>(x, k), true);
        acc = div_nl<NL>(acc, elem32<NL>(Lb, (long)i * s + i), rc + (long)i * G::RS);
        stg_reg<NL>(x + (long)i * Fmt<NL>::ES, acc);
      }
}
// The same three substitutions, right-looking, one CTA per block-parity: step k divides the k-th
// unknown of every line by the pivot (one thread per line), then ALL threads of the CTA subtract its
// multiple from the unknowns still open -- element by element, so every element still receives its
// updates in the canonical order (k ascending; MODE 1: k descending) and the result is bit-identical
// to the one-thread-per-line kernels above, but the dependent chain is s (division + a few
// multiply-accumulates) long instead of s^2 / 2.
//   MODE 0: B <- L^-1 B   (lines = columns of B, unknowns down the column)
//   MODE 1: B <- L^-T B   (lines = columns, unknowns up the column)
//   MODE 2: B <- B L^-T   (lines = rows of B, unknowns along the row)
constexpr int BDM_RL_THREADS = 128;
template <int NL, int MODE>
__global__ void __launch_bounds__(BDM_RL_THREADS)
bdm_trsm_rl_kernel(const BdmDesc *d, const limb_t *L, const uint32_t *recip, limb_t *B)
{
  typedef TileGeom<NL> G;
  constexpr int ES = Fmt<NL>::ES;
  extern __shared__ __align__(16) uint32_t bdm_xk[]; // the unknowns solved in this step, one per line
  const BdmDesc b = d[blockIdx.x];
  const int s = b.s, tid = threadIdx.x, T = blockDim.x;
  if(s == 0)
    return;
  limb_t *Bb = B + b.off * ES;
  const limb_t *Lb = L + b.off * ES;
  const uint32_t *rc = recip + (long)b.cum_cols * G::RS;
  // element (unknown u, line l) of B, in elements
  auto at = [&](int u, int l) -> long { return MODE == 2 ? (long)u * s + l : (long)l * s + u; };
  for(int step = 0; step < s; ++step)
    {
      const int k = MODE == 1 ? s - 1 - step : step;
      for(int l = tid; l < s; l += T)
        {
          Reg<NL> x;
          limb_t *e = Bb + at(k, l) * ES;
          ldg_reg<NL>(x, e);
          x = div_nl<NL>(x, elem32<NL>(Lb, (long)k * s + k), rc + (long)k * G::RS);
          stg_reg<NL>(e, x);
          mpfw::store<NL>(bdm_xk + (size_t)l * G::SW, x);
        }
      __syncthreads();
      const int m = MODE == 1 ? k : s - 1 - k; // unknowns still open
      for(int q = tid; q < m * s; q += T)
        {
          // consecutive threads on consecutive memory: MODE 2 runs down the lines (rows), the others
          // down the open unknowns of one line
          const int u = MODE == 2 ? k + 1 + q / s : (MODE == 1 ? q % m : k + 1 + q % m);
          const int l = MODE == 2 ? q % s : q / m;
          // forward: L(u, k); transposed: L(k, u)
          const long le = MODE == 1 ? (long)u * s + k : (long)k * s + u;
          limb_t *e = Bb + at(u, l) * ES;
          Reg<NL> acc;
          ldg_reg<NL>(acc, e);
          acc = mac_nl<NL>(acc, elem32<NL>(Lb, le), bdm_xk + (size_t)l * G::SW, true);
          stg_reg<NL>(e, acc);
        }
      __syncthreads();
    }
}
// compute_schur_RHS: dx(off + k) = -dual_residues(off + k) - sum_parity sum_a bases(a,k) (sum_b Z(rb h + a, cb h + b) bases(b,k)),
// off = (cb (cb+1)/2 + rb) n.  One thread per stacked row; d: the EVEN-parity descriptor of every
// SDP block followed by its odd one (index 2j, 2j+1).
template <int NL>
__global__ void __launch_bounds__(64)
schur_rhs_kernel(const BdmDesc *d, int J, const int *row_block, long K, const limb_t *V, const limb_t *Z,
                 const limb_t *dual_residues, limb_t *dx)
{
  const long row = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if(row >= K)
    return;
  const int j = row_block[row];
  const BdmDesc b0 = d[2 * j];
  const int n = b0.n, local = (int)(row - b0.x0);
  const int pair = local / n, k = local % n;
  int cb = 0;
  while((cb + 1) * (cb + 2) / 2 <= pair)
    ++cb;
  const int rb = pair - cb * (cb + 1) / 2;
  Reg<NL> out;
  ldg_reg<NL>(out, dual_residues + row * Fmt<NL>::ES);
  out.sign = -out.sign;
  for(int parity = 0; parity < 2; ++parity)
    {
      const BdmDesc b = d[2 * j + parity];
      const int h = b.h, s = b.s;
      if(s == 0)
        continue;
      Reg<NL> acc;
      mpfw::set_zero(acc);
      for(int a = 0; a < h; ++a)
        {
          Reg<NL> zq;
          mpfw::set_zero(zq);
          for(int q = 0; q < h; ++q)
            zq = mac_nl<NL>(zq, elem32<NL>(Z, b.off + (long)(cb * h + q) * s + rb * h + a),
                            elem32<NL>(V, b.voff + (long)k * s + q), false);
          zq = mul_nl<NL>(zq, elem32<NL>(V, b.voff + (long)k * s + a));
          acc = add_nl<NL>(acc, zq);
        }
      out = sub_nl<NL>(out, acc);
    }
  stg_reg<NL>(dx + row * Fmt<NL>::ES, out);
}
// constraint_matrix_weighted_sum: R(rb h + r, cb h + c) = sum_k bases(c,k) a(off + k) bases(r,k)
// (x 0.5 off the block diagonal), then the strictly lower part mirrors the upper one.
// One thread per element (row, col) with row-block <= column-block; the mirror is written by the
// thread that owns the upper element.
template <int NL>
__global__ void __launch_bounds__(64)
weighted_sum_kernel(const BdmDesc *d, int count, const limb_t *V, const limb_t *a, const limb_t *half, limb_t *R)
{
  const BdmDesc b = d[blockIdx.x];
  const int s = b.s, h = b.h, n = b.n;
  for(long e = (long)blockIdx.y * blockDim.x + threadIdx.x; e < (long)s * s; e += (long)gridDim.y * blockDim.x)
    {
      const int row = (int)(e % s), col = (int)(e / s);
      const int rb = row / h, r = row % h, cb = col / h, c = col % h;
      if(rb > cb || (b.m > 1 && row > col))
        continue; // m > 1: the strictly lower triangle is the mirror of the upper one (below)
      Reg<NL> acc;
      mpfw::set_zero(acc);
      const long off = b.x0 + (long)(cb * (cb + 1) / 2 + rb) * n;
      for(int k = 0; k < n; ++k)
        {
          Reg<NL> t;
          ldg_reg<NL>(t, V + (b.voff + (long)k * s + c) * Fmt<NL>::ES);
          t = mul_nl<NL>(t, elem32<NL>(a, off + k));
          t = mul_nl<NL>(t, elem32<NL>(V, b.voff + (long)k * s + r));
          acc = add_nl<NL>(acc, t);
        }
      if(cb != rb)
        acc = mul_nl<NL>(acc, reinterpret_cast<const uint32_t *>(half));
      stg_reg<NL>(R + (b.off + (long)col * s + row) * Fmt<NL>::ES, acc);
      // MakeSymmetric(UPPER) for m > 1: the strictly lower triangle mirrors the upper one.  Inside a
      // diagonal block (rb == cb) both (r, c) and (c, r) are computed by their own threads for
      // m == 1 (the reference leaves them as computed); for m > 1 the upper one wins.
      if(b.m > 1 && row < col)
        stg_reg<NL>(R + (b.off + (long)row * s + col) * Fmt<NL>::ES, acc);
    }
}
// per-block trace (from an exact zero, i ascending); one thread per block-parity
template <int NL>
__global__ void __launch_bounds__(64) bdm_trace_kernel(const BdmDesc *d, int count, const limb_t *M, limb_t *out)
{
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if(q >= count)
    return;
  const BdmDesc b = d[q];
  Reg<NL> acc, v;
  mpfw::set_zero(acc);
  for(int i = 0; i < b.s; ++i)
    {
      ldg_reg<NL>(v, M + (b.off + (long)i * b.s + i) * Fmt<NL>::ES);
      acc = add_nl<NL>(acc, v);
    }
  stg_reg<NL>(out + (long)q * Fmt<NL>::ES, acc);
}
// per-block max |M + mu I|; one warp per block-parity, lanes over the columns
template <int NL>
__global__ void __launch_bounds__(32)
bdm_max_abs_kernel(const BdmDesc *d, int count, const limb_t *M, const limb_t *mu, limb_t *out)
{
  typedef TileGeom<NL> G;
  __shared__ __align__(16) uint32_t best[32 * G::SW];
  const BdmDesc b = d[blockIdx.x];
  const int lane = threadIdx.x;
  Reg<NL> m, v, shift;
  mpfw::set_zero(m);
  ldg_reg<NL>(shift, mu);
  for(int j = lane; j < b.s; j += 32)
    for(int i = 0; i < b.s; ++i)
      {
        ldg_reg<NL>(v, M + (b.off + (long)j * b.s + i) * Fmt<NL>::ES);
        if(i == j)
          v = add_nl<NL>(v, shift);
        if(abs_greater<NL>(v, m))
          m = v;
      }
  if(m.sign < 0)
    m.sign = 1;
  mpfw::store<NL>(best + lane * G::SW, m);
  __syncwarp();
  if(lane == 0)
    {
      for(int q = 1; q < 32; ++q)
        {
          mpfw::load<NL>(v, best + q * G::SW);
          if(abs_greater<NL>(v, m))
            m = v;
        }
      stg_reg<NL>(out + (long)blockIdx.x * Fmt<NL>::ES, m);
    }
}
// per-block sum_ij (X + dX)_ij (Y + dY)_ij: by columns (rows ascending from an exact zero), then
// the column sums ascending.  One warp per block-parity, lanes over the columns.
template <int NL>
__global__ void __launch_bounds__(32)
bdm_frobenius_kernel(const BdmDesc *d, int count, const limb_t *X, const limb_t *dX, const limb_t *Y,
                     const limb_t *dY, limb_t *colsum, limb_t *out)
{
  const BdmDesc b = d[blockIdx.x];
  const int lane = threadIdx.x;
  limb_t *cs = colsum + (long)b.cum_cols * Fmt<NL>::ES;
  for(int j = lane; j < b.s; j += 32)
    {
      Reg<NL> col, t, u, w;
      mpfw::set_zero(col);
      for(int i = 0; i < b.s; ++i)
        {
          const long e = b.off + (long)j * b.s + i;
          ldg_reg<NL>(t, X + e * Fmt<NL>::ES);
          ldg_reg<NL>(w, dX + e * Fmt<NL>::ES);
          t = add_nl<NL>(t, w);
          ldg_reg<NL>(u, Y + e * Fmt<NL>::ES);
          ldg_reg<NL>(w, dY + e * Fmt<NL>::ES);
          u = add_nl<NL>(u, w);
          Reg<NL> p;
          mpfw::set_zero(p);
          if(t.sign != 0 && u.sign != 0)
            {
              const uint32_t *tw = t.w;
              mpfw::mul<NL>(p, t.sign, t.exp, tw, u.sign, u.exp, u.w);
            }
          col = add_nl<NL>(col, p);
        }
      stg_reg<NL>(cs + (long)j * Fmt<NL>::ES, col);
    }
  __syncwarp();
  __threadfence_block();
  if(lane == 0)
    {
      Reg<NL> acc, v;
      mpfw::set_zero(acc);
      for(int j = 0; j < b.s; ++j)
        {
          ldg_reg<NL>(v, cs + (long)j * Fmt<NL>::ES);
          acc = add_nl<NL>(acc, v);
        }
      stg_reg<NL>(out + (long)blockIdx.x * Fmt<NL>::ES, acc);
    }
}
} // namespace sdpb_b200

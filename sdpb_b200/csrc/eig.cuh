// sm_100a kernels of step_length:
//   step_length                          run/step/step_length/step_length.cxx:27-46
//   lower_triangular_inverse_congruence  run/step/step_length/lower_triangular_inverse_congruence.cxx:5-18
//   min_eigenvalue                       run/step/step_length/min_eigenvalue.cxx:8-33
// on the resident Cholesky factors of X / Y and the resident dX / dY: per block-parity the smallest
// eigenvalue of L^-1 dM L^-T.  The operation order is the one csrc/host/step_length.hpp spells out
// (El::HermitianEig lives in the un-vendored Elemental fork, so the order is this repo's):
// congruence by two forward substitutions, Householder tridiagonalisation with one thread per matrix
// row, Laguerre's iteration on the characteristic polynomial of the tridiagonal matrix.  Tests
// compare the eigenvalues with the host restatement byte for byte.
//
// Shapes: 2J matrices of s = 20 ... 128 rows; the work is O(sum s^3) with chains of s dependent
// steps, a few per cent of a Schur-complement step.  What matters is that dX, dY, X, Y and their
// factors never leave HBM: 2J eigenvalues come down.
#pragma once
#include "coop.cuh"
#include "direction.cuh"

namespace sdpb_b200
{
// exact three-way comparison of two mpf values (limb-normalised: a non-zero value has a non-zero top limb)
template <int NL> __device__ __forceinline__ int cmp_reg(const Reg<NL> &a, const Reg<NL> &b)
{
  if(a.sign != b.sign)
    return a.sign > b.sign ? 1 : -1;
  if(a.sign == 0)
    return 0;
  int m = 0; // comparison of the magnitudes
  if(a.exp != b.exp)
    m = a.exp > b.exp ? 1 : -1;
  else
    {
#pragma unroll
      for(int i = 2 * NL - 1; i >= 0; --i)
        if(m == 0 && a.w[i] != b.w[i])
          m = a.w[i] > b.w[i] ? 1 : -1;
    }
  return a.sign > 0 ? m : -m;
}
// the mpf value of a small positive integer (what mpf_set_si gives): one limb
template <int NL> __device__ __forceinline__ void set_small(Reg<NL> &r, uint32_t v)
{
  mpfw::set_zero(r);
  if(v)
    {
      r.sign = 1;
      r.exp = 1;
      r.w[2 * NL - 2] = v;
    }
}
template <int NL> __device__ __forceinline__ void lds_reg(Reg<NL> &r, const uint32_t *p) { mpfw::load<NL>(r, p); }
template <int NL> __device__ __forceinline__ void sts_reg(uint32_t *p, const Reg<NL> &r) { mpfw::store<NL>(p, r); }

// Forward substitution with the lower factor L of every block, k ascending, one division by the
// pivot per unknown (host: trsm_lower_transpose_right / trsm_lower_left_columns).
// ROWS: A <- A L^-T, one thread per ROW of A (the unknowns of a thread are strided by s: the
// threads of a warp read consecutive elements); !ROWS: A <- L^-1 A, one thread per column.
template <int NL, bool ROWS>
__global__ void __launch_bounds__(64)
eig_trsm_kernel(const BdmDesc *d, int count, int total_cols, const limb_t *L, const uint32_t *recip, limb_t *A)
{
  typedef TileGeom<NL> G;
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if(col >= total_cols)
    return;
  const BdmDesc b = d[bdm_find(d, count, col)];
  const int c = col - b.cum_cols, s = b.s;
  const long xs = ROWS ? s : 1;
  limb_t *x = A + (b.off + (ROWS ? (long)c : (long)c * s)) * Fmt<NL>::ES;
  const limb_t *Lb = L + b.off * Fmt<NL>::ES;
  const uint32_t *rc = recip + (long)b.cum_cols * G::RS;
  for(int i = 0; i < s; ++i)
    {
      Reg<NL> acc;
      ldg_reg<NL>(acc, x + (long)i * xs * Fmt<NL>::ES);
      for(int k = 0; k < i; ++k)
        acc = mac_nl<NL>(acc, elem32<NL>(Lb, (long)k * s + i), elem32<NL>(x, (long)k * xs), true);
      acc = div_nl<NL>(acc, elem32<NL>(Lb, (long)i * s + i), rc + (long)i * G::RS);
      stg_reg<NL>(x + (long)i * xs * Fmt<NL>::ES, acc);
    }
}

// ---- Householder tridiagonalisation, one CTA per block-parity, one thread per matrix row ----
template <int NL> struct TridiagSmem
{
  typedef TileGeom<NL> G;
  coop::Work<NL> work;
  uint32_t slot[G::SW];   // norm2 -> sigma
  uint32_t slotH[G::SW];  // H
  uint32_t slotK[G::SW];  // K
  uint32_t recipH[G::RS];
  int flag;               // 1: column already tridiagonal
  // followed by v[s], p[s], w[s] (w doubles as the buffer of the products that are summed), SW words each
};
template <int NL> __host__ __device__ constexpr size_t tridiag_smem_bytes(int s)
{
  return ((sizeof(TridiagSmem<NL>) + 15) & ~(size_t)15) + (size_t)3 * s * TileGeom<NL>::SW * 4;
}
// A: the symmetric matrices, given by their lower triangles (destroyed); dd / ee: diagonal and sub-diagonal, stacked like the
// columns of the block-diagonal object (cum_cols); half: 0.5 as mpf_set_d gives it
constexpr int TRIDIAG_THREADS = 128; // rows beyond that wrap around; the rank-2 update uses all of them
template <int NL>
__global__ void __launch_bounds__(TRIDIAG_THREADS)
eig_tridiag_kernel(const BdmDesc *d, limb_t *A, limb_t *dd, limb_t *ee, const limb_t *half)
{
  typedef TileGeom<NL> G;
  constexpr int ES = Fmt<NL>::ES;
  extern __shared__ __align__(16) unsigned char eig_raw[];
  TridiagSmem<NL> &sm = *reinterpret_cast<TridiagSmem<NL> *>(eig_raw);
  const BdmDesc b = d[blockIdx.x];
  const int s = b.s, tid = threadIdx.x, T = blockDim.x;
  if(s == 0)
    return;
  uint32_t *v = reinterpret_cast<uint32_t *>(eig_raw + ((sizeof(TridiagSmem<NL>) + 15) & ~(size_t)15));
  uint32_t *p = v + (size_t)s * G::SW, *w = p + (size_t)s * G::SW;
  limb_t *Ab = A + b.off * ES, *db = dd + (long)b.cum_cols * ES, *eb = ee + (long)b.cum_cols * ES;
  const uint32_t *half32 = reinterpret_cast<const uint32_t *>(half);
  if(tid == 0)
    sm.work.flag = 0;
  // the matrix is given by its lower triangle (El::HermitianEig(LOWER, ...)): mirror it
  for(int i = tid; i < s; i += T)
    for(int j = i + 1; j < s; ++j)
      {
        Reg<NL> x;
        ldg_reg<NL>(x, Ab + ((long)i * s + j) * ES);
        stg_reg<NL>(Ab + ((long)j * s + i) * ES, x);
      }
  __syncthreads();
  for(int k = 0; k + 2 < s; ++k)
    {
      // x = A(k+1.., k): v <- x, w <- x_i^2 (i > k+1)
      for(int i = k + 1 + tid; i < s; i += T)
        {
          Reg<NL> x;
          ldg_reg<NL>(x, Ab + ((long)k * s + i) * ES);
          sts_reg<NL>(v + (size_t)i * G::SW, x);
          if(i > k + 1)
            {
              x = mul_nl<NL>(x, v + (size_t)i * G::SW);
              sts_reg<NL>(w + (size_t)i * G::SW, x);
            }
        }
      __syncthreads();
      if(tid == 0)
        {
          Reg<NL> tail2, t;
          mpfw::set_zero(tail2);
          for(int i = k + 2; i < s; ++i)
            {
              lds_reg<NL>(t, w + (size_t)i * G::SW);
              tail2 = add_nl<NL>(tail2, t);
            }
          ldg_reg<NL>(t, Ab + ((long)k * s + k) * ES);
          stg_reg<NL>(db + (long)k * ES, t);
          sm.flag = tail2.sign == 0;
          if(tail2.sign == 0)
            {
              lds_reg<NL>(t, v + (size_t)(k + 1) * G::SW);
              stg_reg<NL>(eb + (long)k * ES, t);
            }
          else
            {
              lds_reg<NL>(t, v + (size_t)(k + 1) * G::SW);
              t = mul_nl<NL>(t, v + (size_t)(k + 1) * G::SW);
              tail2 = add_nl<NL>(tail2, t); // norm2 = tail2 + x1^2
              sts_reg<NL>(sm.slot, tail2);
              sts_reg<NL>(sm.slotH, tail2);
            }
        }
      __syncthreads();
      if(sm.flag)
        continue; // uniform: the column is tridiagonal already
      if(tid < 32)
        {
          coop::sqrt_elem<NL>(sm.work, sm.slot); // sigma
          if(tid == 0)
            {
              Reg<NL> alpha, x1, t, H;
              lds_reg<NL>(alpha, sm.slot);
              lds_reg<NL>(x1, v + (size_t)(k + 1) * G::SW);
              if(x1.sign > 0)
                alpha.sign = -1;
              stg_reg<NL>(eb + (long)k * ES, alpha);
              sts_reg<NL>(sm.slot, alpha);
              t = mul_nl<NL>(alpha, v + (size_t)(k + 1) * G::SW); // alpha x1
              lds_reg<NL>(H, sm.slotH);                           // norm2
              H = sub_nl<NL>(H, t);
              sts_reg<NL>(sm.slotH, H);
              x1 = sub_nl<NL>(x1, alpha); // v_{k+1} = x1 - alpha
              sts_reg<NL>(v + (size_t)(k + 1) * G::SW, x1);
            }
          __syncwarp();
          coop::recip_elem<NL>(sm.work, sm.slotH, sm.recipH, nullptr);
        }
      __syncthreads();
      // p_i = (sum_j A(i,j) v_j) / H ; w_i <- p_i v_i
      for(int i = k + 1 + tid; i < s; i += T)
        {
          Reg<NL> acc;
          mpfw::set_zero(acc);
          for(int j = k + 1; j < s; ++j)
            acc = mac_nl<NL>(acc, elem32<NL>(Ab, (long)j * s + i), v + (size_t)j * G::SW, false);
          acc = div_nl<NL>(acc, sm.slotH, sm.recipH);
          sts_reg<NL>(p + (size_t)i * G::SW, acc);
          acc = mul_nl<NL>(acc, v + (size_t)i * G::SW);
          sts_reg<NL>(w + (size_t)i * G::SW, acc);
        }
      __syncthreads();
      if(tid == 0)
        {
          // K = ((sum_i p_i v_i) * 0.5) / H
          Reg<NL> K, t;
          mpfw::set_zero(K);
          for(int i = k + 1; i < s; ++i)
            {
              lds_reg<NL>(t, w + (size_t)i * G::SW);
              K = add_nl<NL>(K, t);
            }
          K = mul_nl<NL>(K, half32);
          K = div_nl<NL>(K, sm.slotH, sm.recipH);
          sts_reg<NL>(sm.slotK, K);
        }
      __syncthreads();
      // w_i = p_i - K v_i
      for(int i = k + 1 + tid; i < s; i += T)
        {
          Reg<NL> acc;
          lds_reg<NL>(acc, p + (size_t)i * G::SW);
          acc = mac_nl<NL>(acc, sm.slotK, v + (size_t)i * G::SW, true);
          sts_reg<NL>(w + (size_t)i * G::SW, acc);
        }
      __syncthreads();
      // A(i,j) -= v_hi w_lo ; A(i,j) -= w_hi v_lo   (hi = max(i,j), lo = min(i,j)); the elements of
      // the trailing matrix are dealt out to all threads, consecutive threads down a column
      const int m = s - k - 1;
      for(int q = tid; q < m * m; q += T)
          {
            const int i = k + 1 + q % m, j = k + 1 + q / m;
            const int hi = i > j ? i : j, lo = i > j ? j : i;
            limb_t *e = Ab + ((long)j * s + i) * ES;
            Reg<NL> acc;
            ldg_reg<NL>(acc, e);
            acc = mac_nl<NL>(acc, v + (size_t)hi * G::SW, w + (size_t)lo * G::SW, true);
            acc = mac_nl<NL>(acc, w + (size_t)hi * G::SW, v + (size_t)lo * G::SW, true);
            stg_reg<NL>(e, acc);
          }
      __syncthreads();
    }
  if(tid == 0)
    {
      Reg<NL> t;
      if(s >= 2)
        {
          ldg_reg<NL>(t, Ab + ((long)(s - 2) * s + (s - 2)) * ES);
          stg_reg<NL>(db + (long)(s - 2) * ES, t);
          ldg_reg<NL>(t, Ab + ((long)(s - 2) * s + (s - 1)) * ES);
          stg_reg<NL>(eb + (long)(s - 2) * ES, t);
        }
      ldg_reg<NL>(t, Ab + ((long)(s - 1) * s + (s - 1)) * ES);
      stg_reg<NL>(db + (long)(s - 1) * ES, t);
    }
}

// ---- smallest eigenvalue of the tridiagonal matrices: Laguerre's iteration, one warp per matrix ----
// Lanes 0, 1, 2 carry the three-term recurrences of p = det(T_k - x), p' and p'' in lock step
// (the same instruction stream: new = (d_k - x) cur - m1 - m2 - e_{k-1}^2 prev with
// (m1, m2) = (0, 0), (p, 0), (p', p')); the Laguerre step itself is lane 0's, its square root and
// reciprocal are taken by the whole warp (coop.cuh).
constexpr int LAGUERRE_MAX_ITERATIONS = 4096; // == host/step_length.hpp
template <int NL> struct LaguerreSmem
{
  typedef TileGeom<NL> G;
  coop::Work<NL> work;
  uint32_t cur[3][G::SW], prev[3][G::SW];
  uint32_t zero[G::SW], x[G::SW], slot[G::SW], den[G::SW];
  uint32_t R[G::RS];
  int state; // 0: iterate, 1: done
};
template <int NL>
__global__ void __launch_bounds__(32)
eig_laguerre_kernel(const BdmDesc *d, const limb_t *dd, const limb_t *ee, limb_t *e2, const limb_t *eps, limb_t *out,
                    int *iterations)
{
  typedef TileGeom<NL> G;
  constexpr int ES = Fmt<NL>::ES;
  extern __shared__ __align__(16) unsigned char eig_raw[];
  LaguerreSmem<NL> &sm = *reinterpret_cast<LaguerreSmem<NL> *>(eig_raw);
  const BdmDesc b = d[blockIdx.x];
  const int n = b.s, lane = threadIdx.x;
  limb_t *res = out + (long)blockIdx.x * ES;
  Reg<NL> x, tol;
  mpfw::set_zero(x);
  if(iterations && lane == 0)
    iterations[blockIdx.x] = 0;
  if(n <= 1)
    {
      if(lane == 0)
        {
          if(n == 1)
            ldg_reg<NL>(x, dd + (long)b.cum_cols * ES);
          stg_reg<NL>(res, x);
        }
      return;
    }
  const limb_t *db = dd + (long)b.cum_cols * ES, *eb = ee + (long)b.cum_cols * ES;
  limb_t *e2b = e2 + (long)b.cum_cols * ES;
  // e2[k] = e[k]^2, one lane per k
  for(int k = lane; k + 1 < n; k += 32)
    {
      Reg<NL> t;
      ldg_reg<NL>(t, eb + (long)k * ES);
      t = mul_nl<NL>(t, elem32<NL>(eb, k));
      stg_reg<NL>(e2b + (long)k * ES, t);
    }
  if(lane == 0)
    {
      // Gershgorin: lo <= lambda_min, scale >= |lambda|
      Reg<NL> lo, scale, r, g, t;
      mpfw::set_zero(lo);
      mpfw::set_zero(scale);
      for(int k = 0; k < n; ++k)
        {
          mpfw::set_zero(r);
          if(k > 0)
            {
              ldg_reg<NL>(t, eb + (long)(k - 1) * ES);
              t.sign = t.sign != 0;
              r = add_nl<NL>(r, t);
            }
          if(k + 1 < n)
            {
              ldg_reg<NL>(t, eb + (long)k * ES);
              t.sign = t.sign != 0;
              r = add_nl<NL>(r, t);
            }
          ldg_reg<NL>(g, db + (long)k * ES);
          t = g;
          g = sub_nl<NL>(g, r);
          if(k == 0 || cmp_reg<NL>(g, lo) < 0)
            lo = g;
          t.sign = t.sign != 0;
          t = add_nl<NL>(t, r);
          if(cmp_reg<NL>(t, scale) > 0)
            scale = t;
        }
      sm.work.flag = 0;
      sm.state = scale.sign == 0;
      mpfw::set_zero(t);
      sts_reg<NL>(sm.zero, t);
      if(scale.sign != 0)
        {
          tol = mul_nl<NL>(scale, reinterpret_cast<const uint32_t *>(eps));
          x = sub_nl<NL>(lo, tol);
        }
      sts_reg<NL>(sm.x, x);
    }
  __syncwarp();
  __threadfence_block();
  Reg<NL> big_n, big_n1;
  set_small<NL>(big_n, (uint32_t)n);
  set_small<NL>(big_n1, (uint32_t)(n - 1));
  for(int it = 0; it < LAGUERRE_MAX_ITERATIONS && sm.state == 0; ++it)
    {
      __syncwarp();
      lds_reg<NL>(x, sm.x);
      if(lane < 3)
        {
          // k = 0: (p, p', p'') = (d_0 - x, -1, 0), previous (1, 0, 0)
          Reg<NL> c, pv;
          mpfw::set_zero(c);
          mpfw::set_zero(pv);
          if(lane == 0)
            {
              ldg_reg<NL>(c, db);
              c = sub_nl<NL>(c, x);
              set_small<NL>(pv, 1u);
            }
          else if(lane == 1)
            {
              set_small<NL>(c, 1u);
              c.sign = -1;
            }
          sts_reg<NL>(sm.cur[lane], c);
          sts_reg<NL>(sm.prev[lane], pv);
        }
      __syncwarp();
      for(int k = 1; k < n; ++k)
        {
          Reg<NL> nw;
          if(lane < 3)
            {
              Reg<NL> t;
              ldg_reg<NL>(nw, db + (long)k * ES);
              nw = sub_nl<NL>(nw, x);                 // d_k - x
              nw = mul_nl<NL>(nw, sm.cur[lane]);      // (d_k - x) cur
              lds_reg<NL>(t, lane == 0 ? sm.zero : sm.cur[lane - 1]);
              nw = sub_nl<NL>(nw, t);
              lds_reg<NL>(t, lane == 2 ? sm.cur[1] : sm.zero);
              nw = sub_nl<NL>(nw, t);
              nw = mac_nl<NL>(nw, elem32<NL>(e2b, k - 1), sm.prev[lane], true);
            }
          __syncwarp();
          if(lane < 3)
            {
              Reg<NL> t;
              lds_reg<NL>(t, sm.cur[lane]);
              sts_reg<NL>(sm.prev[lane], t);
              sts_reg<NL>(sm.cur[lane], nw);
            }
          __syncwarp();
        }
      // left of the spectrum p > 0 > p'; anything else: x has reached the smallest root to within
      // the rounding of the recurrence
      Reg<NL> pq;
      if(lane == 0)
        {
          const int ps = (int32_t)sm.cur[0][1], qs = (int32_t)sm.cur[1][1];
          if(ps <= 0 || qs >= 0)
            sm.state = 1;
          else
            {
              if(iterations)
                iterations[blockIdx.x] += 1;
              // disc = (n (q^2 - p s) - q^2) (n - 1)
              Reg<NL> q, qq, t;
              lds_reg<NL>(q, sm.cur[1]);
              qq = mul_nl<NL>(q, sm.cur[1]);
              lds_reg<NL>(t, sm.cur[0]);
              t = mul_nl<NL>(t, sm.cur[2]); // p s
              Reg<NL> disc = sub_nl<NL>(qq, t);
              sts_reg<NL>(sm.den, big_n);
              disc = mul_nl<NL>(disc, sm.den);
              disc = sub_nl<NL>(disc, qq);
              sts_reg<NL>(sm.den, big_n1);
              disc = mul_nl<NL>(disc, sm.den);
              if(disc.sign < 0)
                mpfw::set_zero(disc);
              sts_reg<NL>(sm.slot, disc);
            }
        }
      __syncwarp();
      if(sm.state)
        break;
      if((int32_t)sm.slot[1] > 0)
        coop::sqrt_elem<NL>(sm.work, sm.slot);
      if(lane == 0)
        {
          Reg<NL> den, q;
          lds_reg<NL>(den, sm.slot);
          lds_reg<NL>(q, sm.cur[1]);
          den = sub_nl<NL>(den, q);
          sts_reg<NL>(sm.den, den);
        }
      __syncwarp();
      coop::recip_elem<NL>(sm.work, sm.den, sm.R, nullptr);
      if(lane == 0)
        {
          // a = n p / den ; x += a ; stop once a <= tol
          Reg<NL> a;
          sts_reg<NL>(sm.slot, big_n);
          lds_reg<NL>(a, sm.slot);
          a = mul_nl<NL>(a, sm.cur[0]);
          a = div_nl<NL>(a, sm.den, sm.R);
          x = add_nl<NL>(x, a);
          sts_reg<NL>(sm.x, x);
          if(cmp_reg<NL>(a, tol) <= 0)
            sm.state = 1;
        }
      __syncwarp();
    }
  __syncwarp();
  if(lane == 0)
    {
      lds_reg<NL>(x, sm.x);
      stg_reg<NL>(res, x);
    }
}
} // namespace sdpb_b200

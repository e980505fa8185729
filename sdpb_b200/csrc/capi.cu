// C-ABI of the B200 Schur-complement step (include/sdpb_b200.h): context,
// HBM layout, kernel sequencing.  No CPU fallback: every entry point that
// computes requires a CUDA device.
#include "ctx.h"
#include "nccl_dyn.h"

#include <algorithm>
#include <chrono>
#include <climits>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

static const LaunchTable *table_for(int nl);
static bool nl_supported(int nl) { return table_for(nl) != nullptr; }
static std::string g_module_error; // why the kernels of a precision could not be loaded (module_table)

// ------------------------------------------------------------ CRT tables
static uint64_t powmod(uint64_t a, uint64_t e, uint64_t m)
{
  uint64_t r = 1;
  a %= m;
  while(e)
    {
      if(e & 1)
        r = (unsigned __int128)r * a % m;
      a = (unsigned __int128)a * a % m;
      e >>= 1;
    }
  return r;
}
static bool is_prime32(uint32_t n)
{
  if(n < 2)
    return false;
  for(uint32_t p : {2u, 3u, 5u, 7u, 11u, 13u, 17u, 19u, 23u, 29u, 31u, 37u})
    {
      if(n % p == 0)
        return n == p;
    }
  uint32_t d = n - 1;
  int r = 0;
  while((d & 1) == 0)
    {
      d >>= 1;
      ++r;
    }
  for(uint64_t a : {2ull, 7ull, 61ull}) // deterministic for n < 4,759,123,141
    {
      uint64_t x = powmod(a, d, n);
      if(x == 1 || x == n - 1)
        continue;
      bool comp = true;
      for(int i = 1; i < r; ++i)
        {
          x = (unsigned __int128)x * x % n;
          if(x == n - 1)
            {
              comp = false;
              break;
            }
        }
      if(comp)
        return false;
    }
  return true;
}

static int build_crt(sdpb_b200_ctx *c)
{
  // |Q'_ij| <= K * (2^prec + small)^2 ; leave 40 bits for K and one for sign
  const int bits_needed = 2 * c->prec + 2 + 40 + 1;
  std::vector<uint32_t> primes;
  double bits = 0;
  for(uint32_t q = (1u << 28) - 1; bits < bits_needed; q -= 2)
    if(is_prime32(q))
      {
        primes.push_back(q);
        bits += log2((double)q);
      }
  const int np = (int)primes.size();
  const int nd = (c->prec + 2 + 27) / 28;
  std::vector<uint32_t> pow28((size_t)np * nd), ginv((size_t)np * np, 0);
  for(int i = 0; i < np; ++i)
    {
      uint64_t x = 1;
      for(int k = 0; k < nd; ++k)
        {
          pow28[(size_t)i * nd + k] = (uint32_t)x;
          x = (x << 28) % primes[i];
        }
      for(int j = 0; j < i; ++j)
        ginv[(size_t)i * np + j]
          = (uint32_t)powmod(primes[j] % primes[i], primes[i] - 2, primes[i]);
    }
  const int mw = (np * 28 + 31) / 32 + 1;
  std::vector<uint32_t> M(mw, 0), Mh(mw, 0);
  M[0] = 1;
  for(int i = 0; i < np; ++i)
    {
      uint64_t carry = 0;
      for(int k = 0; k < mw; ++k)
        {
          const uint64_t z = (uint64_t)M[k] * primes[i] + carry;
          M[k] = (uint32_t)z;
          carry = z >> 32;
        }
    }
  // Mhalf = (M + 1) / 2
  {
    std::vector<uint32_t> t = M;
    uint64_t carry = 1;
    for(int k = 0; k < mw && carry; ++k)
      {
        const uint64_t z = (uint64_t)t[k] + carry;
        t[k] = (uint32_t)z;
        carry = z >> 32;
      }
    for(int k = 0; k < mw; ++k)
      Mh[k] = (t[k] >> 1) | (k + 1 < mw ? (t[k + 1] << 31) : 0);
  }
  auto up = [&](uint32_t **d, const std::vector<uint32_t> &h) -> cudaError_t {
    cudaError_t e = cudaMalloc(d, h.size() * 4);
    if(e != cudaSuccess)
      return e;
    return cudaMemcpy(*d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  };
  // the same powers with rows zero-padded to a multiple of 4 digits (normalize_kernel's
  // shared-memory table, NormGeom<NL>::NDP) and the Barrett constants floor((2^64-1)/p)
  const int ndp = ((64 * (c->nl - 2) + 2 + 27) / 28 + 3) & ~3;
  std::vector<uint32_t> pow28p((size_t)np * ndp, 0);
  std::vector<uint64_t> inv64(np);
  for(int i = 0; i < np; ++i)
    {
      for(int k = 0; k < nd && k < ndp; ++k)
        pow28p[(size_t)i * ndp + k] = pow28[(size_t)i * nd + k];
      inv64[i] = ~0ull / primes[i];
    }
  CUDA_TRY(c, up(&c->d_pow28p, pow28p));
  CUDA_TRY(c, cudaMalloc(&c->d_inv64, inv64.size() * 8));
  CUDA_TRY(c, cudaMemcpy(c->d_inv64, inv64.data(), inv64.size() * 8, cudaMemcpyHostToDevice));
  CUDA_TRY(c, up(&c->d_primes, primes));
  CUDA_TRY(c, up(&c->d_pow28, pow28));
  CUDA_TRY(c, up(&c->d_ginv, ginv));
  CUDA_TRY(c, up(&c->d_M, M));
  CUDA_TRY(c, up(&c->d_Mhalf, Mh));
  c->crt = CrtTables{c->d_primes, c->d_pow28, c->d_ginv, c->d_M, c->d_Mhalf,
                     np,          nd,         mw,        c->d_pow28p, c->d_inv64, ndp};
  return 0;
}

// ----------------------------------------------------------------- create
template <typename T>
static cudaError_t upload(T **d, const std::vector<T> &h)
{
  cudaError_t e = cudaMalloc(d, std::max<size_t>(1, h.size()) * sizeof(T));
  if(e != cudaSuccess)
    return e;
  if(h.empty())
    return cudaSuccess;
  return cudaMemcpy(*d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice);
}

extern "C" int sdpb_b200_elem_words(int prec_bits)
{
  if(prec_bits < 64)
    return 0;
  return (mpfx::stored_limbs(prec_bits) + 2) & ~1;
}
extern "C" int sdpb_b200_stored_limbs(int prec_bits)
{
  return mpfx::stored_limbs(prec_bits);
}

// batched GEMM descriptors, largest k first; returns the number of 16 x 16 output tiles
static int sort_gemm(std::vector<GemmTileDesc> &v)
{
  std::stable_sort(v.begin(), v.end(), [](const GemmTileDesc &a, const GemmTileDesc &b) { return a.K > b.K; });
  int tiles = 0;
  for(auto &d : v)
    {
      d.tile0 = tiles;
      tiles += ((d.M + TS - 1) / TS) * ((d.N + TS - 1) / TS);
    }
  return tiles;
}
// The large operand / result regions of scale_multiply_add and of the resident search direction are
// allocated on first use: a caller that only runs the Schur-complement step (and the largest C5
// corner, J = 1024 blocks of 256 rows at N = 512) does not pay 10 x the size of X for them.
static int ensure_sma_temp(sdpb_b200_ctx *c)
{
  if(c->smaT)
    return 0;
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaMalloc(&c->smaT, std::max<size_t>(16, c->wXY * 8)));
  return 0;
}
static int ensure_sma(sdpb_b200_ctx *c)
{
  if(c->d_gemmSMA)
    return 0;
  if(int rc = ensure_sma_temp(c))
    return rc;
  {
    // scale_multiply_add: C_b = A_b B_b on s x s blocks, column-major
    const size_t bytes = std::max<size_t>(16, c->wXY * 8);
    CUDA_TRY(c, cudaMalloc(&c->smaA, bytes));
    CUDA_TRY(c, cudaMalloc(&c->smaB, bytes));
    CUDA_TRY(c, cudaMalloc(&c->smaC, bytes));
    std::vector<GemmTileDesc> gS;
    for(int q = 0; q < 2 * c->J; ++q)
      {
        const int s = c->g[q / 2].s[q % 2];
        gS.push_back(GemmTileDesc{c->smaA + c->oXY[q], c->smaB + c->oXY[q], c->smaT + c->oXY[q], 1, (long)s,
                                  1, (long)s, s, s, s, 0, 0, 0, 0, 0});
      }
    c->tiles_SMA = sort_gemm(gS);
    CUDA_TRY(c, upload(&c->d_gemmSMA, gS));
  }
  return 0;
}
static int ensure_direction(sdpb_b200_ctx *c)
{
  if(c->d_gemmXY)
    return 0;
  if(int rc = ensure_sma_temp(c))
    return rc;
  const int es = c->es, nl = c->nl, N = c->N;
  {
    // search direction (row N2): resident block-diagonal objects, vectors and descriptors
    const size_t bytes = std::max<size_t>(16, c->wXY * 8);
    for(limb_t **p : {&c->dirMXY, &c->dirR, &c->dirZ, &c->dirDX, &c->dirDY, &c->dirPR})
      {
        CUDA_TRY(c, cudaMalloc(p, bytes));
        CUDA_TRY(c, cudaMemsetAsync(*p, 0, bytes, c->stream));
      }
    std::vector<BdmDesc> bd2;
    std::vector<int> row_block((size_t)c->K);
    int cols = 0;
    for(int j = 0; j < c->J; ++j)
      {
        const BlockGeom &b = c->g[j];
        for(long r = 0; r < b.P; ++r)
          row_block[(size_t)(b.row0 + r)] = j;
        for(int p = 0; p < 2; ++p)
          {
            const int q = 2 * j + p;
            bd2.push_back(BdmDesc{(long)(c->oXY[q] / es), (long)(c->oV[q] / es), b.row0, b.s[p], b.h[p], b.m, b.n, cols});
            cols += b.s[p];
          }
      }
    c->bdm_cols = cols;
    CUDA_TRY(c, upload(&c->d_bdm, bd2));
    CUDA_TRY(c, upload(&c->d_row_block, row_block));
    CUDA_TRY(c, cudaMalloc(&c->dir_dual, std::max<size_t>(16, (size_t)c->K * es * 8)));
    CUDA_TRY(c, cudaMalloc(&c->dir_prp, (size_t)N * es * 8));
    CUDA_TRY(c, cudaMalloc(&c->dir_scal, 4 * es * 8));
    CUDA_TRY(c, cudaMalloc(&c->dir_part, std::max<size_t>(16, (size_t)2 * c->J * es * 8)));
    CUDA_TRY(c, cudaMalloc(&c->dir_colsum, std::max<size_t>(16, (size_t)cols * es * 8)));
    for(limb_t **p : {&c->eig_d, &c->eig_e, &c->eig_e2})
      CUDA_TRY(c, cudaMalloc(p, std::max<size_t>(16, (size_t)cols * es * 8)));
    CUDA_TRY(c, cudaMalloc(&c->eig_iter, std::max<size_t>(16, (size_t)2 * c->J * sizeof(int))));
    const size_t pin = std::max<size_t>((size_t)c->K + N + 8, (size_t)2 * c->J + 8) * es * 8;
    CUDA_TRY(c, cudaMallocHost(&c->dir_pinned, pin));
    {
      // 0.5 as mpf_set_d(0.5) stores it (two limbs, the low one zero), top-aligned
      std::vector<uint64_t> half(4 * es, 0);
      half[es + 0] = (uint64_t)(uint32_t)0 | ((uint64_t)(uint32_t)1 << 32);
      half[es + nl] = 0x8000000000000000ull;
      // 2^-(prec - 16) = 2^r B^-q (one limb): where Laguerre's iteration stops (host/step_length.hpp)
      const int kbits = c->prec - 16, q = (kbits + 63) / 64, r = 64 * q - kbits;
      half[3 * es + 0] = (uint64_t)(uint32_t)(1 - q) | ((uint64_t)(uint32_t)1 << 32);
      half[3 * es + nl] = (uint64_t)1 << r;
      CUDA_TRY(c, cudaMemcpy(c->dir_scal, half.data(), half.size() * 8, cudaMemcpyHostToDevice));
    }
    auto products = [&](const limb_t *A, const limb_t *B, GemmTileDesc **out) -> cudaError_t {
      std::vector<GemmTileDesc> gd;
      for(int q = 0; q < 2 * c->J; ++q)
        {
          const int s = c->g[q / 2].s[q % 2];
          gd.push_back(GemmTileDesc{A + c->oXY[q], B + c->oXY[q], c->smaT + c->oXY[q], 1, (long)s, 1, (long)s, s, s,
                                    s, 0, 0, 0, 0, 0});
        }
      c->tiles_dir = sort_gemm(gd);
      return upload(out, gd);
    };
    CUDA_TRY(c, products(c->Xin, c->Yin, &c->d_gemmXY));
    CUDA_TRY(c, products(c->dirDX, c->dirDY, &c->d_gemmDXDY));
    CUDA_TRY(c, products(c->dirPR, c->Yin, &c->d_gemmPRY));
    CUDA_TRY(c, products(c->dirDX, c->Yin, &c->d_gemmDXY));
  }
  return 0;
}
extern "C" int sdpb_b200_create(sdpb_b200_ctx **out, int prec_bits, int device,
                                int num_blocks, const int *dims,
                                const int *num_points, int N, char *err,
                                size_t errlen)
{
  auto fail = [&](int rc, const std::string &msg) {
    if(err && errlen)
      snprintf(err, errlen, "%s", msg.c_str());
    return rc;
  };
  if(!out || num_blocks < 0 || N <= 0 || prec_bits < 64)
    return fail(SDPB_B200_ERR_ARG, "sdpb_b200_create: bad argument");
  const int nl = mpfx::stored_limbs(prec_bits);
  if(nl < 3 || nl > 34)
    return fail(SDPB_B200_ERR_ARG, "sdpb_b200_create: precision " + std::to_string(prec_bits)
                                     + " is outside the supported range (64 ... 2048 bits)");
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if(ce != cudaSuccess || ndev == 0)
    return fail(SDPB_B200_ERR_CUDA,
                std::string("sdpb_b200_create: no CUDA device (")
                  + cudaGetErrorString(ce)
                  + "); this library has no CPU fallback");
  if(device < 0 || device >= ndev)
    return fail(SDPB_B200_ERR_ARG, "sdpb_b200_create: bad device ordinal");
  // only now: a precision whose kernels are not linked in is built on demand (module_table),
  // which needs nvcc and minutes -- pointless on a host that cannot run them
  if(!nl_supported(nl))
    return fail(SDPB_B200_ERR_ARG,
                "sdpb_b200_create: precision " + std::to_string(prec_bits)
                  + " (NL=" + std::to_string(nl)
                  + ") has no compiled kernels: " + g_module_error);
  auto *c = new sdpb_b200_ctx();
  c->prec = prec_bits;
  c->nl = nl;
  c->es = (nl + 2) & ~1;
  c->device = device;
  c->J = num_blocks;
  c->N = N;
  auto bail = [&](int rc) {
    const std::string msg = c->error;
    sdpb_b200_destroy(c);
    return fail(rc, msg);
  };
  if(cudaSetDevice(device) != cudaSuccess)
    {
      c->error = "cudaSetDevice failed";
      return bail(SDPB_B200_ERR_CUDA);
    }
  long row0 = 0;
  const size_t es = c->es;
  for(int j = 0; j < num_blocks; ++j)
    {
      BlockGeom b;
      b.m = dims[j];
      b.n = num_points[j];
      if(b.m <= 0 || b.n <= 0)
        {
          c->error = "sdpb_b200_create: bad block shape";
          return bail(SDPB_B200_ERR_ARG);
        }
      b.P = b.n * b.m * (b.m + 1) / 2;
      b.mn = b.m * b.n;
      b.s[0] = b.m * ((b.n + 1) / 2);
      b.s[1] = b.mn - b.s[0];
      const int deg = b.n - 1;
      b.h[0] = deg / 2 + 1;
      b.h[1] = (deg + 1) / 2;
      b.row0 = row0;
      row0 += b.P;
      c->g.push_back(b);
      c->oB.push_back(c->wB);
      c->wB += (size_t)b.P * N * es;
      c->oS.push_back(c->wS);
      c->wS += (size_t)b.P * b.P * es;
      for(int p = 0; p < 2; ++p)
        {
          c->oV.push_back(c->wV);
          c->wV += (size_t)b.s[p] * b.mn * es;
          c->oXY.push_back(c->wXY);
          c->wXY += (size_t)b.s[p] * b.s[p] * es;
          c->oA.push_back(c->wA);
          c->wA += (size_t)b.mn * b.mn * es;
          c->max_s = std::max(c->max_s, b.s[p]);
        }
      c->max_mn = std::max(c->max_mn, b.mn);
      c->max_P = std::max(c->max_P, b.P);
    }
  c->K = row0;
  const size_t wPart = (size_t)std::max(1, num_blocks) * N * es;
  const size_t wNorm = (size_t)N * es, wQ = (size_t)N * N * es;
  c->arena_words = 2 * c->wB + c->wS + 3 * c->wV + 5 * c->wXY + 2 * c->wA
                   + wPart + wNorm + wQ + 64;
#define TRY_C(expr)                                                           \
  do                                                                          \
    {                                                                         \
      cudaError_t e_ = (expr);                                                \
      if(e_ != cudaSuccess)                                                   \
        {                                                                     \
          c->error = std::string("CUDA: ") + cudaGetErrorString(e_) + " at "  \
                     + __FILE__ + ":" + std::to_string(__LINE__);             \
          return bail(SDPB_B200_ERR_CUDA);                                    \
        }                                                                     \
    }                                                                         \
  while(0)
  TRY_C(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  c->cur = c->stream;
  for(auto &a : c->aux)
    TRY_C(cudaStreamCreateWithFlags(&a, cudaStreamNonBlocking));
  for(auto &e : c->evf)
    TRY_C(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  TRY_C(cudaStreamCreateWithFlags(&c->copy, cudaStreamNonBlocking));
  {
    int least = 0, greatest = 0;
    TRY_C(cudaDeviceGetStreamPriorityRange(&least, &greatest));
    const char *env = getenv("SDPB_B200_PRIORITY");
    if(!env || atoi(env) != 0)
      for(auto &p : c->prio)
        TRY_C(cudaStreamCreateWithPriority(&p, cudaStreamNonBlocking, greatest));
  }
  for(auto &e : c->evd)
    TRY_C(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  if(const char *env = getenv("SDPB_B200_CONCURRENCY"))
    c->concurrency = atoi(env) != 0;
  TRY_C(cudaMalloc(&c->arena, c->arena_words * sizeof(limb_t)));
  TRY_C(cudaMemsetAsync(c->arena, 0, c->arena_words * sizeof(limb_t), c->stream));
  {
    limb_t *p = c->arena;
    c->B = p;      p += c->wB;
    c->Pband = p;  p += c->wB;
    c->S = p;      p += c->wS;
    c->V = p;      p += c->wV;
    c->T = p;      p += c->wV;
    c->YV = p;     p += c->wV;
    c->X = p;      p += c->wXY;
    c->Y = p;      p += c->wXY;
    c->LY = p;     p += c->wXY;
    c->Xin = p;    p += c->wXY;
    c->Yin = p;    p += c->wXY;
    c->AX = p;     p += c->wA;
    c->AY = p;     p += c->wA;
    c->part = p;   p += wPart;
    c->norms = p;  p += wNorm;
    c->Q = p;      p += wQ;
  }
  {
    const int rc = build_crt(c);
    if(rc)
      return bail(rc);
  }
  c->NS = (N + 15) & ~15; // residue row stride: whole 16-column tiles, pad columns stay zero
  c->KR = (c->K + SI_KS - 1) / SI_KS * SI_KS; // whole pipeline stages of the tensor-path syrk; pad rows stay zero
  if(const char *e = getenv("SDPB_B200_SYRK"))
    c->syrk_imma = std::string(e) != "imad";
  TRY_C(cudaMalloc(&c->R, std::max<size_t>(4, (size_t)c->crt.np * c->KR * c->NS * 4)));
  TRY_C(cudaMemset(c->R, 0, std::max<size_t>(4, (size_t)c->crt.np * c->KR * c->NS * 4)));
  TRY_C(cudaMalloc(&c->Qres, (size_t)c->crt.np * N * N * 4));
  TRY_C(cudaMemset(c->Qres, 0, (size_t)c->crt.np * N * N * 4));
  // descriptors
  const int rs = ((2 * nl + 4) + 3) & ~3; // TileGeom::RS
  {
    size_t nXY = 0;
    for(const BlockGeom &b : c->g)
      nXY += b.s[0] + b.s[1];
    TRY_C(cudaMalloc(&c->recipX, std::max<size_t>(16, nXY * rs * 4)));
    TRY_C(cudaMalloc(&c->recipY, std::max<size_t>(16, nXY * rs * 4)));
    TRY_C(cudaMalloc(&c->recipS, std::max<size_t>(16, (size_t)c->K * rs * 4)));
    TRY_C(cudaMalloc(&c->recipQ, (size_t)N * rs * 4));
    TRY_C(cudaMalloc(&c->recipN, (size_t)N * rs * 4));
  }
  std::vector<PotrfDesc> pX, pY, pS;
  std::vector<TrsmTileDesc> tT, tP;
  std::vector<GemmTileDesc> gAX, gYV, gAY;
  std::vector<SchurDesc> sd;
  std::vector<BandDesc> bd;
  {
    size_t rXY = 0;
    for(int j = 0; j < num_blocks; ++j)
      {
        const BlockGeom &b = c->g[j];
        for(int p = 0; p < 2; ++p)
          {
            const int q = 2 * j + p;
            const int s = b.s[p];
            pX.push_back(PotrfDesc{c->X + c->oXY[q], c->recipX + rXY * rs, s, 1, (long)s, q});
            pY.push_back(PotrfDesc{c->LY + c->oXY[q], c->recipY + rXY * rs, s, 1, (long)s, q});
            // bases_blocks[q] = I_m (x) v: diagonal blocks of h rows x n columns
            const int hb = b.m > 1 ? s / b.m : 0, nbk = b.n;
            tT.push_back(TrsmTileDesc{c->X + c->oXY[q], c->recipX + rXY * rs, c->T + c->oV[q],
                                      s, b.mn, hb, nbk});
            rXY += s;
            // AX = T^T T : A(i,l) = T(l,i), B(l,j) = T(l,j)
            gAX.push_back(GemmTileDesc{c->T + c->oV[q], c->T + c->oV[q], c->AX + c->oA[q],
                                       (long)s, 1, 1, (long)s, b.mn, b.mn, s, 1, 0,
                                       hb ? 3 : 0, hb, nbk});
            // YV = Y V
            gYV.push_back(GemmTileDesc{c->Y + c->oXY[q], c->V + c->oV[q], c->YV + c->oV[q], 1,
                                       (long)s, 1, (long)s, s, b.mn, s, 0, 0,
                                       hb ? 1 : 0, hb, nbk});
            // AY = V^T (YV)
            gAY.push_back(GemmTileDesc{c->V + c->oV[q], c->YV + c->oV[q], c->AY + c->oA[q],
                                       (long)s, 1, 1, (long)s, b.mn, b.mn, s, 1, 0,
                                       hb ? 2 : 0, hb, nbk});
          }
        pS.push_back(PotrfDesc{c->S + c->oS[j], c->recipS + (size_t)b.row0 * rs, b.P, 1,
                               (long)b.P, j});
        tP.push_back(TrsmTileDesc{c->S + c->oS[j], c->recipS + (size_t)b.row0 * rs,
                                  c->Pband + c->oB[j], b.P, N, 0, 0});
        sd.push_back(SchurDesc{{c->AX + c->oA[2 * j], c->AX + c->oA[2 * j + 1]},
                               {c->AY + c->oA[2 * j], c->AY + c->oA[2 * j + 1]},
                               c->S + c->oS[j], b.m, b.n});
        bd.push_back(BandDesc{c->Pband + c->oB[j], b.P, (int)b.row0, j});
      }
  }
  // largest matrices first: the long CTAs start early, the short ones fill in
  auto by_size = [](const PotrfDesc &a, const PotrfDesc &b) { return a.s > b.s; };
  std::stable_sort(pX.begin(), pX.end(), by_size);
  std::stable_sort(pY.begin(), pY.end(), by_size);
  std::stable_sort(pS.begin(), pS.end(), by_size);
  auto sort_trsm = [](std::vector<TrsmTileDesc> &v, std::vector<int> &sizes) {
    std::stable_sort(v.begin(), v.end(),
                     [](const TrsmTileDesc &a, const TrsmTileDesc &b) { return a.p > b.p; });
    sizes.clear();
    for(auto &d : v)
      sizes.push_back(d.p);
  };
  sort_trsm(tT, c->szT);
  sort_trsm(tP, c->szP);
  for(auto &d : pX)
    c->szXY.push_back(d.s);
  for(auto &d : pS)
    c->szS.push_back(d.s);
  c->tiles_AX = sort_gemm(gAX);
  c->tiles_YV = sort_gemm(gYV);
  c->tiles_AY = sort_gemm(gAY);
  c->n_gemm = (int)gAX.size();
  // the S chain in G interleaved groups (largest blocks dealt round-robin), each with its
  // own sorted descriptor arrays, so that the groups can run on separate streams
  {
    // measured at c3 (profiles/r01_v5_summary.md): every level kernel already fills the
    // CTA slots, so splitting the batch only multiplies the level count (G = 1: 145 ms,
    // 2: 156, 4: 153); the mechanism stays for shapes with few, large blocks
    int G = 1;
    // default: the split by size class where the batch has two classes (note below), else ONE
    // group.  Measured at c3: one group 98.2 ms, split by size with the large blocks' chain on a
    // stream of the greatest priority 94.3 ms (without the priority 99.3: the round-2 v5 finding),
    // interleaved groups 113.6 (2) / 111.5 (3)
    bool by_size = c->prio[0] != nullptr, forced = false;
    if(const char *env = getenv("SDPB_B200_GROUPS"))
      {
        by_size = forced = std::string(env) == "size"; // "size": split whatever the block count
        G = by_size ? 1 : atoi(env);
      }
    std::vector<int> order(num_blocks);
    for(int j = 0; j < num_blocks; ++j)
      order[j] = j;
    std::stable_sort(order.begin(), order.end(),
                     [&](int a, int b) { return c->g[a].P > c->g[b].P; });
    // Split by size: the batched Cholesky of the S_j is level-synchronous over 16-row tiles, and
    // once the small blocks have dropped out its late levels are short chains of pivots on a few
    // large blocks that leave most of the machine idle.  With the blocks in two size classes
    // (c3: 150 of 120 rows, 450 of 40) the small class gets its own stream: its factorisation is
    // over after a few levels and its triangular solves fill the SMs under the large blocks'
    // pivot chains.  The split is where the tile count drops the most; none if it never halves.
    // (On streams of equal priority it does not pay -- the two chains' kernels mostly queue behind
    // one another instead of sharing the SMs; with the large blocks' chain on a stream of the
    // greatest priority its short kernels are placed at once and the small blocks' CTAs fill the
    // rest: 98.2 -> 94.3 ms at c3.  SDPB_B200_GROUPS=1 keeps one group.)
    std::vector<int> group_of(num_blocks, 0);
    if(by_size && (forced ? num_blocks >= 2 : num_blocks >= 2 * 148))
      {
        auto tiles = [&](int k) { return (c->g[order[k]].P + TS - 1) / TS; };
        int cut = 0, best = 0;
        for(int k = 1; k < num_blocks; ++k)
          if(tiles(k - 1) - tiles(k) > best)
            {
              best = tiles(k - 1) - tiles(k);
              cut = k;
            }
        if(best > 0 && (forced || (cut >= 148 / 2 && num_blocks - cut >= 148 && 2 * tiles(cut) <= tiles(0))))
          {
            G = 2;
            for(int k = cut; k < num_blocks; ++k)
              group_of[k] = 1;
          }
        else
          by_size = false;
      }
    else
      by_size = false;
    c->G = std::max(1, std::min(std::min(G, (int)sdpb_b200_ctx::MAXG), std::max(1, num_blocks)));
    c->split_by_size = by_size && c->G == 2;
    if(!by_size)
      for(int k = 0; k < num_blocks; ++k)
        group_of[k] = k % c->G; // interleaved: the largest blocks dealt round-robin
    for(int g = 0; g < c->G; ++g)
      {
        std::vector<PotrfDesc> gp;
        std::vector<TrsmTileDesc> gt;
        std::vector<SchurDesc> gs;
        std::vector<BandDesc> gb;
        for(int k = 0; k < num_blocks; ++k)
          {
            if(group_of[k] != g)
              continue;
            const int j = order[k];
            const BlockGeom &b = c->g[j];
            gp.push_back(PotrfDesc{c->S + c->oS[j], c->recipS + (size_t)b.row0 * rs, b.P, 1,
                                   (long)b.P, j});
            gt.push_back(TrsmTileDesc{c->S + c->oS[j], c->recipS + (size_t)b.row0 * rs,
                                      c->Pband + c->oB[j], b.P, N, 0, 0});
            gs.push_back(sd[j]);
            gb.push_back(bd[j]);
            c->blocks_g[g].push_back(j);
            c->szS_g[g].push_back(b.P);
            c->szP_g[g].push_back(b.P);
            c->maxP_g[g] = std::max(c->maxP_g[g], b.P);
          }
        c->nblk_g[g] = (int)gp.size();
        TRY_C(upload(&c->d_potrfS_g[g], gp));
        TRY_C(upload(&c->d_trsmP_g[g], gt));
        TRY_C(upload(&c->d_schur_g[g], gs));
        TRY_C(upload(&c->d_bands_g[g], gb));
      }
  }
  {
    // triangular systems of the Schur solves: every L_j (largest first), and Q = U^T U read as U^T
    std::vector<SolveTriDesc> sS, sQ{SolveTriDesc{c->Q, c->recipQ, (long)N, 1, N, 0}};
    for(const PotrfDesc &d : pS)
      sS.push_back(SolveTriDesc{d.A, d.recip, 1, (long)d.s, d.s, c->g[d.id].row0});
    TRY_C(upload(&c->d_solveS, sS));
    TRY_C(upload(&c->d_solveQ, sQ));
    TRY_C(cudaMalloc(&c->sol_x, std::max<size_t>(16, (size_t)c->K * es * 8)));
    TRY_C(cudaMalloc(&c->sol_y, std::max<size_t>(16, (size_t)N * es * 8)));
    TRY_C(cudaMallocHost(&c->sol_pinned, std::max<size_t>(16, (size_t)(c->K + N) * es * 8)));
  }
  std::vector<PotrfDesc> pQ{PotrfDesc{c->Q, c->recipQ, N, (long)N, 1, 0}}; // upper: A = U^T U
  c->szQ.assign(1, N);
  TRY_C(upload(&c->d_potrfQ, pQ));
  TRY_C(upload(&c->d_potrfX, pX));
  TRY_C(upload(&c->d_potrfY, pY));
  TRY_C(upload(&c->d_potrfS, pS));
  TRY_C(upload(&c->d_trsmT, tT));
  TRY_C(upload(&c->d_trsmP, tP));
  TRY_C(upload(&c->d_gemmAX, gAX));
  TRY_C(upload(&c->d_gemmYV, gYV));
  TRY_C(upload(&c->d_gemmAY, gAY));
  TRY_C(upload(&c->d_schur, sd));
  c->h_bands = bd;
  TRY_C(upload(&c->d_bands, bd));
  TRY_C(cudaMalloc(&c->d_status, (size_t)(5 * num_blocks + 8) * sizeof(int)));
  TRY_C(cudaMalloc(&c->d_flags, 4 * sizeof(int)));
  TRY_C(cudaMalloc(&c->d_fail, 4 * sizeof(int)));
  TRY_C(cudaMemset(c->d_fail, 0, 4 * sizeof(int)));
  {
    // sdpb_b200_cholesky_diagonals: [X factors | Y factors | L_j | chol(Q)]
    std::vector<DiagDesc> dd;
    long off = 0;
    c->diag_off[0] = off;
    for(int q = 0; q < 2 * num_blocks; ++q)
      {
        const int s = c->g[q / 2].s[q % 2];
        dd.push_back(DiagDesc{c->X + c->oXY[q], s, (long)s, off});
        off += s;
      }
    c->diag_off[1] = off;
    for(int q = 0; q < 2 * num_blocks; ++q)
      {
        const int s = c->g[q / 2].s[q % 2];
        dd.push_back(DiagDesc{c->LY + c->oXY[q], s, (long)s, off});
        off += s;
      }
    c->diag_off[2] = off;
    for(int j = 0; j < num_blocks; ++j)
      {
        dd.push_back(DiagDesc{c->S + c->oS[j], c->g[j].P, (long)c->g[j].P, off});
        off += c->g[j].P;
      }
    c->diag_off[3] = off;
    dd.push_back(DiagDesc{c->Q, N, (long)N, off});
    off += N;
    c->diag_total = off;
    TRY_C(upload(&c->d_diag, dd));
    TRY_C(cudaMalloc(&c->diag_buf, (size_t)off * es * 8));
  }
  c->gidx.resize(num_blocks);
  for(int j = 0; j < num_blocks; ++j)
    c->gidx[j] = j;
  for(auto &e : c->ev)
    TRY_C(cudaEventCreate(&e));
  TRY_C(cudaStreamSynchronize(c->stream));
#undef TRY_C
  *out = c;
  return 0;
}

// ------------------------------------------------------------- multi-GPU
// The SDP blocks are sharded over `world` processes, one GPU each (the
// reference's block parallelism, block_mapping/compute_block_grid_mapping.hxx:58-183).
// Two exchanges remain inside initialize_schur_complement_solver: the per-block
// column-norm partials (Matrix_Normalizer.cxx:131 is an MPI AllReduce) and the
// exact integer Q' partial sums (restore_and_reduce.cxx:137-212 is a ring of
// SendRecv); both are one ncclAllReduce here.  Cholesky(Q): replicated for small N, by
// broadcast panels from qdist_min_N on (launch_nl.cu potrf_rl).
static int nccl_allreduce(sdpb_b200_ctx *c, void *buf, size_t count, int is_u64, const char *label)
{
  NcclApi &api = nccl_api();
  c->kt_begin(label);
  const ncclResult_t r = api.AllReduce(buf, buf, count, is_u64 ? ncclUint64 : ncclUint32, ncclSum,
                                       (ncclComm_t)c->comm, c->stream);
  c->kt_end();
  --c->launches; // NCCL's kernel, not one of ours
  if(r != ncclSuccess)
    {
      c->error = std::string("NCCL: ") + api.GetErrorString(r) + " in " + label;
      return SDPB_B200_ERR_CUDA;
    }
  return 0;
}

static int nccl_bcast(sdpb_b200_ctx *c, void *buf, size_t bytes, int root, const char *label)
{
  NcclApi &api = nccl_api();
  c->kt_begin(label);
  const ncclResult_t r = api.Broadcast(buf, buf, bytes, ncclUint8, root, (ncclComm_t)c->comm, c->stream);
  c->kt_end();
  --c->launches; // NCCL's kernel, not one of ours
  if(r != ncclSuccess)
    {
      c->error = std::string("NCCL: ") + api.GetErrorString(r) + " in " + label;
      return SDPB_B200_ERR_CUDA;
    }
  return 0;
}

extern "C" int sdpb_b200_comm_get_unique_id(void *id)
{
  NcclApi &api = nccl_api();
  if(!id || !api.load())
    return SDPB_B200_ERR_CUDA;
  ncclUniqueId u;
  if(api.GetUniqueId(&u) != ncclSuccess)
    return SDPB_B200_ERR_CUDA;
  memcpy(id, u.internal, NCCL_UNIQUE_ID_BYTES);
  return 0;
}

// what both communicators share: the global block indices, the buffer of per-GLOBAL-block
// partial rows, the panel buffer of the distributed Cholesky(Q)
static int comm_common(sdpb_b200_ctx *c, int rank, int world, int num_blocks_global,
                       const int *global_block_index)
{
  if(world < 1 || world > 16 || rank < 0 || rank >= world || num_blocks_global < c->J
     || (c->J && !global_block_index))
    {
      c->error = "sdpb_b200_comm_init: bad argument (1 <= world <= 16: u32 residue sums must not overflow)";
      return SDPB_B200_ERR_ARG;
    }
  if(c->comm || c->local)
    {
      c->error = "sdpb_b200_comm_init: communicator already initialised";
      return SDPB_B200_ERR_STATE;
    }
  for(int j = 0; j < c->J; ++j)
    {
      if(global_block_index[j] < 0 || global_block_index[j] >= num_blocks_global)
        {
          c->error = "sdpb_b200_comm_init: global block index out of range";
          return SDPB_B200_ERR_ARG;
        }
      c->h_bands[j].gidx = global_block_index[j];
      c->gidx[j] = global_block_index[j];
    }
  CUDA_TRY(c, cudaSetDevice(c->device));
  if(c->J)
    CUDA_TRY(c, cudaMemcpy(c->d_bands, c->h_bands.data(), c->h_bands.size() * sizeof(BandDesc),
                           cudaMemcpyHostToDevice));
  for(int g = 0; g < c->G; ++g) // the per-group copies of the band descriptors carry gidx too
    {
      std::vector<BandDesc> gb;
      for(int j : c->blocks_g[g])
        gb.push_back(c->h_bands[j]);
      if(!gb.empty())
        CUDA_TRY(c, cudaMemcpy(c->d_bands_g[g], gb.data(), gb.size() * sizeof(BandDesc),
                               cudaMemcpyHostToDevice));
    }
  // one row of partial sums per GLOBAL block: used whenever a communicator exists (also with
  // world == 1, where the global indices need not be 0..J-1)
  CUDA_TRY(c, cudaMalloc(&c->part_global, (size_t)std::max(1, num_blocks_global) * c->N * c->es * 8));
  CUDA_TRY(c, cudaMalloc(&c->qpanel, ((size_t)c->N * TS * c->es + 2) * 8 + (size_t)TS * (2 * c->nl + 8) * 4));
  c->rank = rank;
  c->world = world;
  c->J_global = num_blocks_global;
  if(const char *env = getenv("SDPB_B200_QDIST_MIN_N"))
    c->qdist_min_N = atoi(env);
  return 0;
}

extern "C" int sdpb_b200_comm_init(sdpb_b200_ctx *c, int rank, int world, const void *id,
                                   int num_blocks_global, const int *global_block_index)
{
  if(!c)
    return SDPB_B200_ERR_ARG;
  if(!id)
    {
      c->error = "sdpb_b200_comm_init: null communicator id";
      return SDPB_B200_ERR_ARG;
    }
  NcclApi &api = nccl_api();
  if(!api.load())
    {
      c->error = api.error;
      return SDPB_B200_ERR_CUDA;
    }
  if(int rc = comm_common(c, rank, world, num_blocks_global, global_block_index))
    return rc;
  ncclUniqueId u;
  memcpy(u.internal, id, NCCL_UNIQUE_ID_BYTES);
  ncclComm_t comm;
  const ncclResult_t r = api.CommInitRank(&comm, world, u, rank);
  if(r != ncclSuccess)
    {
      c->error = std::string("NCCL: ncclCommInitRank: ") + api.GetErrorString(r);
      return SDPB_B200_ERR_CUDA;
    }
  c->comm = comm;
  c->allreduce = &nccl_allreduce;
  c->bcast = &nccl_bcast;
  return 0;
}

// ---- in-process communicator ------------------------------------------------
// Several contexts of ONE process on ONE device, each driven by its own host thread, exchange
// through device memory: the sharded code path (global-order sums, exact residue sums, the
// panel-distributed Cholesky(Q), the sharded Schur solve) then runs -- and can be checked against
// the unsharded oracle -- on a single GPU.  Same hooks as the NCCL communicator.
struct LocalGroup
{
  int world = 0;
  std::mutex m;
  std::condition_variable cv;
  int arrived = 0;
  long generation = 0;
  bool broken = false;
  std::vector<void *> ptr;
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  ~LocalGroup() { cudaFree(tmp); }
  // false: a peer did not arrive within the time limit (it failed before the exchange)
  bool barrier()
  {
    std::unique_lock<std::mutex> lock(m);
    if(broken)
      return false;
    const long gen = generation;
    if(++arrived == world)
      {
        arrived = 0;
        ++generation;
        cv.notify_all();
        return true;
      }
    if(!cv.wait_for(lock, std::chrono::seconds(120), [&] { return generation != gen || broken; }))
      {
        broken = true;
        cv.notify_all();
        return false;
      }
    return !broken;
  }
};
struct LocalPtrs
{
  const void *p[16];
};
template <typename T> __global__ void local_sum_kernel(LocalPtrs src, int world, size_t count, T *out)
{
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
    {
      T s = 0;
      for(int r = 0; r < world; ++r)
        s += static_cast<const T *>(src.p[r])[i];
      out[i] = s;
    }
}
static int local_fail(sdpb_b200_ctx *c, const char *label)
{
  c->error = std::string("in-process communicator: a peer context did not reach ") + label;
  return SDPB_B200_ERR_STATE;
}
static int local_allreduce(sdpb_b200_ctx *c, void *buf, size_t count, int is_u64, const char *label)
{
  LocalGroup &g = *c->local;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  {
    std::lock_guard<std::mutex> lock(g.m);
    g.ptr[c->rank] = buf;
  }
  if(!g.barrier())
    return local_fail(c, label);
  if(c->rank == 0)
    {
      const size_t bytes = count * (is_u64 ? 8 : 4);
      if(g.tmp_bytes < bytes)
        {
          cudaFree(g.tmp);
          g.tmp = nullptr;
          g.tmp_bytes = 0;
          CUDA_TRY(c, cudaMalloc(&g.tmp, bytes));
          g.tmp_bytes = bytes;
        }
      LocalPtrs src{};
      for(int r = 0; r < g.world; ++r)
        src.p[r] = g.ptr[r];
      const unsigned blocks = (unsigned)std::min<size_t>((count + 255) / 256, 148 * 8);
      c->cur = c->stream;
      c->kt_begin(label);
      if(is_u64)
        local_sum_kernel<uint64_t><<<blocks, 256, 0, c->stream>>>(src, g.world, count, (uint64_t *)g.tmp);
      else
        local_sum_kernel<uint32_t><<<blocks, 256, 0, c->stream>>>(src, g.world, count, (uint32_t *)g.tmp);
      c->kt_end();
      CUDA_TRY(c, cudaGetLastError());
      for(int r = 0; r < g.world; ++r)
        CUDA_TRY(c, cudaMemcpyAsync(g.ptr[r], g.tmp, bytes, cudaMemcpyDeviceToDevice, c->stream));
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
  if(!g.barrier())
    return local_fail(c, label);
  return 0;
}
static int local_bcast(sdpb_b200_ctx *c, void *buf, size_t bytes, int root, const char *label)
{
  LocalGroup &g = *c->local;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  {
    std::lock_guard<std::mutex> lock(g.m);
    g.ptr[c->rank] = buf;
  }
  if(!g.barrier())
    return local_fail(c, label);
  if(c->rank != root)
    {
      CUDA_TRY(c, cudaMemcpyAsync(buf, g.ptr[root], bytes, cudaMemcpyDeviceToDevice, c->stream));
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
  if(!g.barrier())
    return local_fail(c, label);
  return 0;
}

extern "C" int sdpb_b200_comm_init_local(sdpb_b200_ctx *const *ctxs, int world, int num_blocks_global,
                                         const int *const *global_block_index)
{
  if(!ctxs || world < 1 || world > 16)
    return SDPB_B200_ERR_ARG;
  for(int r = 0; r < world; ++r)
    if(!ctxs[r])
      return SDPB_B200_ERR_ARG;
  for(int r = 1; r < world; ++r)
    if(ctxs[r]->device != ctxs[0]->device || ctxs[r]->prec != ctxs[0]->prec || ctxs[r]->N != ctxs[0]->N)
      {
        ctxs[0]->error = ctxs[r]->error
          = "sdpb_b200_comm_init_local: the contexts must share device, precision and N";
        return SDPB_B200_ERR_ARG;
      }
  auto group = std::make_shared<LocalGroup>();
  group->world = world;
  group->ptr.assign(world, nullptr);
  for(int r = 0; r < world; ++r)
    {
      sdpb_b200_ctx *c = ctxs[r];
      if(int rc = comm_common(c, r, world, num_blocks_global, global_block_index ? global_block_index[r] : nullptr))
        return rc;
      c->local = group;
      c->allreduce = &local_allreduce;
      c->bcast = &local_bcast;
    }
  return 0;
}

extern "C" void sdpb_b200_destroy(sdpb_b200_ctx *c)
{
  if(!c)
    return;
  cudaSetDevice(c->device);
  if(c->comm)
    {
      cudaStreamSynchronize(c->stream);
      nccl_api().CommDestroy((ncclComm_t)c->comm);
    }
  cudaFree(c->part_global);
  cudaFree(c->qpanel);
  cudaFree(c->arena);
  cudaFree(c->R);
  cudaFree(c->Qres);
  cudaFree(c->d_primes);
  cudaFree(c->d_pow28);
  cudaFree(c->d_pow28p);
  cudaFree(c->d_inv64);
  cudaFree(c->d_ginv);
  cudaFree(c->d_M);
  cudaFree(c->d_Mhalf);
  cudaFree(c->d_potrfX);
  cudaFree(c->d_potrfY);
  cudaFree(c->d_potrfS);
  cudaFree(c->d_solveS);
  cudaFree(c->smaA);
  cudaFree(c->smaB);
  cudaFree(c->smaT);
  cudaFree(c->smaC);
  cudaFree(c->d_gemmSMA);
  cudaFree(c->d_solveQ);
  cudaFree(c->sol_x);
  cudaFree(c->sol_y);
  cudaFreeHost(c->sol_pinned);
  cudaFree(c->d_potrfQ);
  cudaFree(c->recipX);
  cudaFree(c->recipY);
  cudaFree(c->recipS);
  cudaFree(c->recipQ);
  cudaFree(c->recipN);
  cudaFree(c->d_trsmT);
  cudaFree(c->d_trsmP);
  cudaFree(c->d_gemmAX);
  cudaFree(c->d_gemmYV);
  cudaFree(c->d_gemmAY);
  cudaFree(c->d_schur);
  cudaFree(c->d_bands);
  cudaFree(c->d_status);
  cudaFree(c->d_flags);
  cudaFree(c->d_fail);
  for(limb_t *p : {c->eig_d, c->eig_e, c->eig_e2})
    cudaFree(p);
  cudaFree(c->eig_iter);
  for(limb_t *p : {c->dirMXY, c->dirR, c->dirZ, c->dirDX, c->dirDY, c->dirPR, c->dir_dual, c->dir_prp, c->dir_scal,
                   c->dir_part, c->dir_colsum})
    cudaFree(p);
  cudaFree(c->d_bdm);
  for(auto &kv : c->bdm_trsm_descs)
    cudaFree(kv.second);
  cudaFree(c->d_row_block);
  cudaFree(c->d_gemmXY);
  cudaFree(c->d_gemmDXDY);
  cudaFree(c->d_gemmPRY);
  cudaFree(c->d_gemmDXY);
  if(c->dir_pinned)
    cudaFreeHost(c->dir_pinned);
  cudaFree(c->d_diag);
  cudaFree(c->diag_buf);
  for(int g = 0; g < sdpb_b200_ctx::MAXG; ++g)
    {
      cudaFree(c->d_potrfS_g[g]);
      cudaFree(c->d_trsmP_g[g]);
      cudaFree(c->d_schur_g[g]);
      cudaFree(c->d_bands_g[g]);
      if(c->aux[g])
        cudaStreamDestroy(c->aux[g]);
    }
  for(auto &e : c->evf)
    if(e)
      cudaEventDestroy(e);
  for(auto &e : c->evd)
    if(e)
      cudaEventDestroy(e);
  if(c->copy)
    cudaStreamDestroy(c->copy);
  for(auto &p : c->prio)
    if(p)
      cudaStreamDestroy(p);
  if(c->pinned)
    cudaFreeHost(c->pinned);
  for(auto &k : c->kt)
    {
      cudaEventDestroy(k.e0);
      cudaEventDestroy(k.e1);
    }
  if(c->stream)
    {
      for(auto &e : c->ev)
        if(e)
          cudaEventDestroy(e);
      cudaStreamDestroy(c->stream);
    }
  delete c;
}

extern "C" const char *sdpb_b200_last_error(const sdpb_b200_ctx *c)
{
  return c ? c->error.c_str() : "null context";
}

// --------------------------------------------------------------- set_block
extern "C" int sdpb_b200_set_block(sdpb_b200_ctx *c, int j, const uint64_t *B,
                                   const uint64_t *bases_even,
                                   const uint64_t *bases_odd)
{
  if(!c || j < 0 || j >= c->J || !B)
    return SDPB_B200_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  const BlockGeom &b = c->g[j];
  const size_t es = c->es;
  CUDA_TRY(c, cudaMemcpyAsync(c->B + c->oB[j], B, (size_t)b.P * c->N * es * 8,
                              cudaMemcpyHostToDevice, c->stream));
  // bases_blocks = I_m (x) basis  (SDP/set_bases_blocks.cxx:24-47), dense
  for(int p = 0; p < 2; ++p)
    {
      const uint64_t *basis = p == 0 ? bases_even : bases_odd;
      const int h = b.h[p], s = b.s[p];
      if(s == 0)
        continue;
      if(!basis)
        {
          c->error = "sdpb_b200_set_block: missing bilinear basis";
          return SDPB_B200_ERR_ARG;
        }
      std::vector<uint64_t> V((size_t)s * b.mn * es, 0);
      for(int col = 0; col < b.mn; ++col)
        for(int row = 0; row < s; ++row)
          if(row / h == col / b.n)
            memcpy(&V[((size_t)col * s + row) * es],
                   basis + ((size_t)(col % b.n) * h + (row % h)) * es, es * 8);
      CUDA_TRY(c, cudaMemcpyAsync(c->V + c->oV[2 * j + p], V.data(),
                                  V.size() * 8, cudaMemcpyHostToDevice,
                                  c->stream));
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));
    }
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return 0;
}

static const LaunchTable *builtin_table(int nl)
{
  switch(nl)
    {
#define F(n) case n: return &sdpb_b200_launch_nl##n; // null if this build left the precision out (weak)
      SDPB_FOR_EACH_NL(F)
#undef F
    default: return nullptr;
    }
}
// Any --precision (Environment.cxx:29-36 accepts any): the kernels of a stored-limb count that is
// not linked into libsdpb_b200.so live in libsdpb_b200_nl<NL>.so next to it; when that file does
// not exist yet it is built on the spot from csrc/launch_nl.cu (`make nl NL=<n>`, nvcc for
// sm_100a, one to ten minutes depending on the precision) and kept for the next run.
// SDPB_B200_JIT=0 forbids the build.
static std::map<int, const LaunchTable *> g_modules;
static std::mutex g_modules_guard;
static const LaunchTable *module_table(int nl)
{
  std::lock_guard<std::mutex> lock(g_modules_guard);
  auto it = g_modules.find(nl);
  if(it != g_modules.end())
    return it->second;
  g_module_error.clear();
  if(nl < 3 || nl > 34)
    {
      g_module_error = "stored limbs out of range (precisions from 64 to 2048 bits)";
      return nullptr;
    }
  Dl_info info;
  if(!dladdr((void *)&sdpb_b200_elem_words, &info) || !info.dli_fname)
    {
      g_module_error = "cannot locate libsdpb_b200.so on disk";
      return nullptr;
    }
  std::string dir(info.dli_fname);
  dir = dir.substr(0, dir.find_last_of('/') == std::string::npos ? 0 : dir.find_last_of('/'));
  if(dir.empty())
    dir = ".";
  const std::string so = dir + "/libsdpb_b200_nl" + std::to_string(nl) + ".so";
  void *h = dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL);
  if(!h)
    {
      const char *jit = getenv("SDPB_B200_JIT");
      if(jit && atoi(jit) == 0)
        {
          g_module_error = so + " is missing and SDPB_B200_JIT=0 forbids building it";
          return nullptr;
        }
      const std::string log = so + ".log";
      const std::string cmd = "make -C '" + dir + "/csrc' nl NL=" + std::to_string(nl) + " > '" + log + "' 2>&1";
      fprintf(stderr, "sdpb_b200: building the kernels for %d stored limbs (%s)\n", nl, cmd.c_str());
      const int rc = system(cmd.c_str());
      h = rc == 0 ? dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL) : nullptr;
      if(!h)
        {
          g_module_error = "building " + so + " failed (see " + log + ")";
          return nullptr;
        }
    }
  const std::string sym = "sdpb_b200_launch_nl" + std::to_string(nl);
  const LaunchTable *t = static_cast<const LaunchTable *>(dlsym(h, sym.c_str()));
  if(!t)
    g_module_error = so + " lacks " + sym;
  else if(t->ctx_bytes != sizeof(sdpb_b200_ctx) || t->table_bytes != sizeof(LaunchTable))
    {
      g_module_error = so + " was built from other sources than this library (delete it to have it rebuilt)";
      t = nullptr;
    }
  else
    g_modules[nl] = t;
  return t;
}
static const LaunchTable *table_for(int nl)
{
  if(const LaunchTable *t = builtin_table(nl))
    return t;
  return module_table(nl);
}
static int dispatch_cholesky(sdpb_b200_ctx *c, int which)
{
  return table_for(c->nl)->cholesky(c, which);
}
// A_X_inv chain on `stream`, A_Y chain beside it; `from_event` (an index into evf, recorded by
// the caller on `stream` when Y became ready) lets the side chain start before chol(X) ends
static int dispatch_pairings(sdpb_b200_ctx *c, int y_ready_event = -1)
{
  cudaStream_t st = c->stream, sy = c->side(0);
  if(sy != st)
    {
      if(y_ready_event < 0)
        {
          y_ready_event = 0;
          CUDA_TRY(c, cudaEventRecord(c->evf[0], st));
        }
      CUDA_TRY(c, cudaStreamWaitEvent(sy, c->evf[y_ready_event], 0));
    }
  c->cur = sy;
  int rc = table_for(c->nl)->pairings(c, 1);
  c->cur = st;
  if(rc)
    return rc;
  rc = table_for(c->nl)->pairings(c, 0);
  if(rc)
    return rc;
  CUDA_TRY(c, c->after(sy, st, 1));
  return 0;
}
static int dispatch_schur_and_Q(sdpb_b200_ctx *c)
{
  return table_for(c->nl)->schur_and_Q(c);
}
static int dispatch_schur_solve(sdpb_b200_ctx *c)
{
  return table_for(c->nl)->schur_solve(c);
}
static int dispatch_scalar(sdpb_b200_ctx *c, int op, int k, long count,
                           const limb_t *a, const limb_t *b, limb_t *r)
{
  return table_for(c->nl)->scalar(c, op, k, count, a, b, r);
}

// block_timings (compute_Q.cxx:40,52 feed the reference's block_mapping on restart): the time of
// cholesky_j + solve_j is one batched stage here, split over the blocks by the reference's own
// cost model P^3/3 + P^2 N/2 (bigint_syrk/Readme.md:327-346); never less than 1 ms per block, so
// that a later block_mapping does not see zero-cost blocks
static void add_block_timings(const sdpb_b200_ctx *c, int32_t *block_timings_ms)
{
  if(!block_timings_ms)
    return;
  double tot = 0;
  for(int j = 0; j < c->J; ++j)
    {
      const double P = c->g[j].P;
      tot += P * P * P / 3 + P * P * c->N / 2;
    }
  for(int j = 0; j < c->J; ++j)
    {
      const double P = c->g[j].P;
      const double share = c->stage_ms[3] * (P * P * P / 3 + P * P * c->N / 2) / (tot > 0 ? tot : 1);
      block_timings_ms[j] += (int32_t)std::max(1L, std::lround(share));
    }
}

// read status words; returns index of first failing matrix or -1
static int first_bad(sdpb_b200_ctx *c, const int *d_status, int count, int *pivot)
{
  std::vector<int> h(count);
  if(cudaMemcpyAsync(h.data(), d_status, count * sizeof(int),
                     cudaMemcpyDeviceToHost, c->stream)
       != cudaSuccess
     || cudaStreamSynchronize(c->stream) != cudaSuccess)
    return -2;
  for(int i = 0; i < count; ++i)
    if(h[i] >= 0)
      {
        if(pivot)
          *pivot = h[i];
        return i;
      }
  return -1;
}

static int copy_blocks_in(sdpb_b200_ctx *c, const uint64_t *const *A, limb_t *dst)
{
  const int n = 2 * c->J;
  auto words_of = [&](int q) { return (size_t)c->g[q / 2].s[q % 2] * c->g[q / 2].s[q % 2] * c->es; };
  for(int q = 0; q < n; ++q) // checked before anything is enqueued
    if(words_of(q) && (!A || !A[q]))
      {
        c->error = "null input block " + std::to_string(q);
        return SDPB_B200_ERR_ARG;
      }
  for(int q = 0; q < n;)
    {
      size_t words = words_of(q);
      if(words == 0)
        {
          ++q;
          continue;
        }
      // blocks the caller packed back to back (as the arena is): one DMA for the run
      int r = q + 1;
      while(r < n && (words_of(r) == 0 || (A[r] == A[q] + words && c->oXY[r] == c->oXY[q] + words)))
        words += words_of(r++);
      CUDA_TRY(c, cudaMemcpyAsync(dst + c->oXY[q], A[q], words * 8, cudaMemcpyHostToDevice, c->stream));
      q = r;
    }
  return 0;
}
static int copy_blocks_out(sdpb_b200_ctx *c, const limb_t *src,
                           const std::vector<size_t> &off,
                           uint64_t *const *out, int count,
                           const std::vector<size_t> &elems)
{
  if(!out)
    return 0;
  // blocks that are adjacent on both sides (the arena is packed in block order; a caller
  // that packs its staging buffer the same way gets one DMA instead of one per block)
  for(int q = 0; q < count;)
    {
      if(!out[q] || !elems[q])
        {
          ++q;
          continue;
        }
      size_t words = elems[q] * c->es;
      int r = q + 1;
      while(r < count && out[r] && elems[r] && out[r] == out[q] + words
            && off[r] == off[q] + words)
        words += elems[r++] * c->es;
      CUDA_TRY(c, cudaMemcpyAsync(out[q], src + off[q], words * 8, cudaMemcpyDeviceToHost, c->stream));
      q = r;
    }
  return 0;
}

extern "C" int sdpb_b200_cholesky_decomposition(sdpb_b200_ctx *c, int which,
                                                const uint64_t *const *A,
                                                uint64_t *const *L)
{
  if(!c || (which != 0 && which != 1))
    return SDPB_B200_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  limb_t *dst = which == 0 ? c->X : c->LY;
  c->cur = c->stream;
  if(which == 0)
    c->kt_used = 0;
  int rc = copy_blocks_in(c, A, dst);
  if(rc)
    return rc;
  // the pristine matrix stays resident too: -XY and the corrector's Frobenius product read it
  if(c->wXY)
    CUDA_TRY(c, cudaMemcpyAsync(which == 0 ? c->Xin : c->Yin, dst, c->wXY * 8, cudaMemcpyDeviceToDevice, c->stream));
  rc = dispatch_cholesky(c, which);
  if(rc)
    {
      cudaStreamSynchronize(c->stream);
      return rc;
    }
  int pivot = 0;
  const int bad = first_bad(c, c->d_status + which * 2 * c->J, 2 * c->J, &pivot);
  if(bad == -2)
    {
      c->error = "CUDA failure while reading Cholesky status";
      return SDPB_B200_ERR_CUDA;
    }
  if(bad >= 0)
    {
      c->error = std::string("Error when computing Cholesky decomposition of "
                             "Block_Diagonal_Matrix ")
                 + (which == 0 ? "X" : "Y")
                 + ", block index = " + std::to_string(c->gidx[bad / 2])
                 + ", parity = " + std::to_string(bad % 2)
                 + ": non-positive pivot " + std::to_string(pivot);
      return SDPB_B200_ERR_NOT_HPD;
    }
  std::vector<size_t> elems(2 * c->J);
  for(int q = 0; q < 2 * c->J; ++q)
    elems[q] = (size_t)c->g[q / 2].s[q % 2] * c->g[q / 2].s[q % 2];
  rc = copy_blocks_out(c, dst, c->oXY, L, 2 * c->J, elems);
  if(rc)
    return rc;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if(which == 0)
    c->have_X_cholesky = true;
  return 0;
}

extern "C" int sdpb_b200_compute_bilinear_pairings(sdpb_b200_ctx *c,
                                                   const uint64_t *const *Y,
                                                   uint64_t *const *A_X_inv,
                                                   uint64_t *const *A_Y)
{
  if(!c)
    return SDPB_B200_ERR_ARG;
  if(!c->have_X_cholesky)
    {
      c->error = "compute_bilinear_pairings called before "
                 "cholesky_decomposition(X)";
      return SDPB_B200_ERR_STATE;
    }
  CUDA_TRY(c, cudaSetDevice(c->device));
  int rc = copy_blocks_in(c, Y, c->Y);
  if(rc)
    return rc;
  if(c->wXY)
    CUDA_TRY(c, cudaMemcpyAsync(c->Yin, c->Y, c->wXY * 8, cudaMemcpyDeviceToDevice, c->stream));
  CUDA_TRY(c, cudaEventRecord(c->ev[0], c->stream));
  rc = dispatch_pairings(c);
  if(rc)
    {
      cudaStreamSynchronize(c->stream);
      return rc;
    }
  CUDA_TRY(c, cudaEventRecord(c->ev[1], c->stream));
  std::vector<size_t> elems(2 * c->J);
  for(int q = 0; q < 2 * c->J; ++q)
    elems[q] = (size_t)c->g[q / 2].mn * c->g[q / 2].mn;
  rc = copy_blocks_out(c, c->AX, c->oA, A_X_inv, 2 * c->J, elems);
  if(rc)
    return rc;
  rc = copy_blocks_out(c, c->AY, c->oA, A_Y, 2 * c->J, elems);
  if(rc)
    return rc;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  cudaEventElapsedTime(&c->stage_ms[1], c->ev[0], c->ev[1]);
  c->have_pairings = true;
  return 0;
}

extern "C" int sdpb_b200_initialize_schur_complement_solver(
  sdpb_b200_ctx *c, uint64_t *const *schur_complement_cholesky,
  uint64_t *const *schur_off_diagonal, uint64_t *Q, int32_t *block_timings_ms)
{
  if(!c)
    return SDPB_B200_ERR_ARG;
  if(!c->have_pairings)
    {
      c->error = "initialize_schur_complement_solver called before "
                 "compute_bilinear_pairings";
      return SDPB_B200_ERR_STATE;
    }
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->have_factors = false;
  int rc = dispatch_schur_and_Q(c);
  if(rc)
    return rc;
  int pivot = 0;
  const int bad = first_bad(c, c->d_status + 4 * c->J, c->J, &pivot);
  if(bad == -2)
    {
      c->error = "CUDA failure while reading Cholesky status";
      return SDPB_B200_ERR_CUDA;
    }
  if(bad >= 0)
    {
      c->error = "Error when computing Cholesky decomposition of block_"
                 + std::to_string(c->gidx[bad]) + ": non-positive pivot "
                 + std::to_string(pivot);
      return SDPB_B200_ERR_NOT_HPD;
    }
  int flags[4];
  CUDA_TRY(c, cudaMemcpyAsync(flags, c->d_flags, sizeof(flags),
                              cudaMemcpyDeviceToHost, c->stream));
  int qstat = 0;
  CUDA_TRY(c, cudaMemcpyAsync(&qstat, c->d_status + 5 * c->J, sizeof(int),
                              cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if(flags[0])
    {
      c->error = "normalised P entry does not fit the integer syrk format";
      return SDPB_B200_ERR_Q_DIAG;
    }
  if(flags[1] != INT_MAX)
    {
      c->error = "Normalized Q should have ones on diagonal. For i = "
                 + std::to_string(flags[1]);
      return SDPB_B200_ERR_Q_DIAG;
    }
  if(qstat >= 0)
    {
      c->error = "Error when computing Cholesky(Q): non-positive pivot "
                 + std::to_string(qstat);
      return SDPB_B200_ERR_NOT_HPD;
    }
  if(c->sharded())
    {
      int fail0 = 0;
      CUDA_TRY(c, cudaMemcpyAsync(&fail0, c->d_fail, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      if(fail0 > 0)
        {
          c->error = "Schur-complement step failed on " + std::to_string(fail0)
                     + " peer rank(s) (their last_error names the block); Q is not usable";
          return SDPB_B200_ERR_NOT_HPD;
        }
    }
  c->have_factors = true;
  {
    std::vector<size_t> eS(c->J), eP(c->J);
    for(int j = 0; j < c->J; ++j)
      {
        eS[j] = (size_t)c->g[j].P * c->g[j].P;
        eP[j] = (size_t)c->g[j].P * c->N;
      }
    rc = copy_blocks_out(c, c->S, c->oS, schur_complement_cholesky, c->J, eS);
    if(rc)
      return rc;
    rc = copy_blocks_out(c, c->Pband, c->oB, schur_off_diagonal, c->J, eP);
    if(rc)
      return rc;
  }
  if(Q)
    CUDA_TRY(c, cudaMemcpyAsync(Q, c->Q, (size_t)c->N * c->N * c->es * 8,
                                cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  for(int k = 2; k < 8; ++k)
    cudaEventElapsedTime(&c->stage_ms[k], c->ev[k], c->ev[k + 1]);
  cudaEventElapsedTime(&c->stage_ms[8], c->ev[2], c->ev[8]);
  add_block_timings(c, block_timings_ms);
  return 0;
}

// ------------------------------------------- solve_schur_complement_equation
// (solve_schur_complement_equation.cxx:16-79) on the resident factors.  The
// right-hand sides travel through one pinned staging buffer: one H2D and one
// D2H of (P + N) elements per solve.
extern "C" int sdpb_b200_solve_schur_complement_equation(sdpb_b200_ctx *c, uint64_t *const *dx,
                                                         uint64_t *dy)
{
  if(!c)
    return SDPB_B200_ERR_ARG;
  if(!c->have_factors)
    {
      c->error = "solve_schur_complement_equation called before a successful "
                 "initialize_schur_complement_solver";
      return SDPB_B200_ERR_STATE;
    }
  if(!dy || (c->J && !dx))
    {
      c->error = "solve_schur_complement_equation: null argument";
      return SDPB_B200_ERR_ARG;
    }
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const size_t es = (size_t)c->es;
  for(int j = 0; j < c->J; ++j)
    {
      if(c->g[j].P && !dx[j])
        {
          c->error = "solve_schur_complement_equation: dx[" + std::to_string(j) + "] is null";
          return SDPB_B200_ERR_ARG;
        }
      memcpy(c->sol_pinned + (size_t)c->g[j].row0 * es, dx[j], (size_t)c->g[j].P * es * 8);
    }
  memcpy(c->sol_pinned + (size_t)c->K * es, dy, (size_t)c->N * es * 8);
  if(c->K)
    CUDA_TRY(c, cudaMemcpyAsync(c->sol_x, c->sol_pinned, (size_t)c->K * es * 8, cudaMemcpyHostToDevice, st));
  CUDA_TRY(c, cudaMemcpyAsync(c->sol_y, c->sol_pinned + (size_t)c->K * es, (size_t)c->N * es * 8,
                              cudaMemcpyHostToDevice, st));
  c->kt_used = 0;
  CUDA_TRY(c, cudaEventRecord(c->ev[9], st));
  if(int rc = dispatch_schur_solve(c))
    return rc;
  CUDA_TRY(c, cudaEventRecord(c->ev[10], st));
  if(c->K)
    CUDA_TRY(c, cudaMemcpyAsync(c->sol_pinned, c->sol_x, (size_t)c->K * es * 8, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaMemcpyAsync(c->sol_pinned + (size_t)c->K * es, c->sol_y, (size_t)c->N * es * 8,
                              cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  cudaEventElapsedTime(&c->solve_ms, c->ev[9], c->ev[10]);
  for(int j = 0; j < c->J; ++j)
    memcpy(dx[j], c->sol_pinned + (size_t)c->g[j].row0 * es, (size_t)c->g[j].P * es * 8);
  memcpy(dy, c->sol_pinned + (size_t)c->K * es, (size_t)c->N * es * 8);
  return 0;
}
extern "C" float sdpb_b200_last_solve_ms(const sdpb_b200_ctx *c) { return c ? c->solve_ms : 0.f; }

// ---------------------------------------------------------- scale_multiply_add
extern "C" int sdpb_b200_scale_multiply_add(sdpb_b200_ctx *c, int alpha, const uint64_t *const *A,
                                            const uint64_t *const *B, int beta, uint64_t *const *C)
{
  if(!c)
    return SDPB_B200_ERR_ARG;
  if((alpha != 1 && alpha != -1) || (beta != 0 && beta != 1) || !C)
    {
      c->error = "scale_multiply_add: alpha must be 1 or -1, beta 0 or 1, C non-null";
      return SDPB_B200_ERR_ARG;
    }
  CUDA_TRY(c, cudaSetDevice(c->device));
  if(int rc0 = ensure_sma(c))
    return rc0;
  int rc = copy_blocks_in(c, A, c->smaA);
  if(!rc)
    rc = copy_blocks_in(c, B, c->smaB);
  if(!rc && beta)
    rc = copy_blocks_in(c, C, c->smaC);
  c->kt_used = 0;
  if(!rc)
    rc = table_for(c->nl)->scale_multiply_add(c, alpha, beta);
  if(rc)
    {
      cudaStreamSynchronize(c->stream); // earlier uploads may still read the caller's buffers
      return rc;
    }
  std::vector<size_t> elems(2 * c->J);
  for(int q = 0; q < 2 * c->J; ++q)
    elems[q] = (size_t)c->g[q / 2].s[q % 2] * c->g[q / 2].s[q % 2];
  rc = copy_blocks_out(c, c->smaC, c->oXY, C, 2 * c->J, elems);
  if(rc)
    return rc;
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// ------------------------------------------------------- search direction
// Row N2: compute_search_direction.cxx:44-90 and the reductions of step.cxx:137-160
// on objects that stay in HBM.  The factors (X_cholesky, L_j, L_j^-1 B_j, chol(Q)) and the
// pristine X, Y are those of the preceding step / cholesky_decomposition + pairings calls.
static int direction_ready(sdpb_b200_ctx *c, const char *what, bool need_minus_XY, bool need_direction)
{
  if(!c->have_factors || (need_minus_XY && !c->have_minus_XY) || (need_direction && !c->have_direction))
    {
      c->error = std::string(what) + " called out of order (needs a successful Schur-complement step"
                 + (need_minus_XY ? ", direction_begin" : "") + (need_direction ? ", compute_search_direction" : "") + ")";
      return SDPB_B200_ERR_STATE;
    }
  return 0;
}
// per block-parity scalars back to the host (2J packed elements)
static int direction_scalars_out(sdpb_b200_ctx *c, uint64_t *out)
{
  const size_t bytes = (size_t)2 * c->J * c->es * 8;
  if(bytes && out)
    {
      CUDA_TRY(c, cudaMemcpyAsync(c->dir_pinned, c->dir_part, bytes, cudaMemcpyDeviceToHost, c->stream));
      CUDA_TRY(c, cudaStreamSynchronize(c->stream));
      memcpy(out, c->dir_pinned, bytes);
    }
  else
    CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return 0;
}
extern "C" int sdpb_b200_direction_begin(sdpb_b200_ctx *c, uint64_t *block_traces)
{
  if(!c)
    return SDPB_B200_ERR_ARG;
  if(int rc = direction_ready(c, "direction_begin", false, false))
    return rc;
  if(int rc = ensure_direction(c))
    return rc;
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->kt_used = 0;
  c->have_direction = false;
  CUDA_TRY(c, cudaEventRecord(c->ev[9], c->stream));
  if(int rc = table_for(c->nl)->direction(c, 0, 0))
    {
      cudaStreamSynchronize(c->stream);
      return rc;
    }
  CUDA_TRY(c, cudaEventRecord(c->ev[10], c->stream));
  if(int rc = direction_scalars_out(c, block_traces))
    return rc;
  cudaEventElapsedTime(&c->direction_ms, c->ev[9], c->ev[10]);
  c->have_minus_XY = true;
  return 0;
}
extern "C" int sdpb_b200_direction_R_errors(sdpb_b200_ctx *c, const uint64_t *mu, uint64_t *block_maxima)
{
  if(!c || !mu)
    return SDPB_B200_ERR_ARG;
  if(int rc = direction_ready(c, "direction_R_errors", true, false))
    return rc;
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->kt_used = 0;
  memcpy(c->dir_pinned, mu, (size_t)c->es * 8);
  CUDA_TRY(c, cudaMemcpyAsync(c->dir_scal + 2 * c->es, c->dir_pinned, (size_t)c->es * 8, cudaMemcpyHostToDevice,
                              c->stream));
  if(int rc = table_for(c->nl)->direction(c, 1, 0))
    {
      cudaStreamSynchronize(c->stream);
      return rc;
    }
  return direction_scalars_out(c, block_maxima);
}
extern "C" int sdpb_b200_direction_set_residues(sdpb_b200_ctx *c, const uint64_t *const *primal_residues,
                                                const uint64_t *const *dual_residues,
                                                const uint64_t *primal_residue_p)
{
  if(!c || !primal_residue_p || (c->J && (!primal_residues || !dual_residues)))
    return SDPB_B200_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  if(int rc0 = ensure_direction(c))
    return rc0;
  const size_t es = (size_t)c->es;
  for(int j = 0; j < c->J; ++j)
    if(c->g[j].P && !dual_residues[j])
      {
        c->error = "direction_set_residues: dual_residues[" + std::to_string(j) + "] is null";
        return SDPB_B200_ERR_ARG;
      }
  cudaStream_t main_stream = c->stream;
  int rc = copy_blocks_in(c, primal_residues, c->dirPR);
  if(rc)
    return rc;
  for(int j = 0; j < c->J; ++j)
    memcpy(c->dir_pinned + (size_t)c->g[j].row0 * es, dual_residues[j], (size_t)c->g[j].P * es * 8);
  memcpy(c->dir_pinned + (size_t)c->K * es, primal_residue_p, (size_t)c->N * es * 8);
  if(c->K)
    CUDA_TRY(c, cudaMemcpyAsync(c->dir_dual, c->dir_pinned, (size_t)c->K * es * 8, cudaMemcpyHostToDevice, main_stream));
  CUDA_TRY(c, cudaMemcpyAsync(c->dir_prp, c->dir_pinned + (size_t)c->K * es, (size_t)c->N * es * 8,
                              cudaMemcpyHostToDevice, main_stream));
  CUDA_TRY(c, cudaStreamSynchronize(main_stream));
  c->have_residues = true;
  return 0;
}
extern "C" int sdpb_b200_compute_search_direction(sdpb_b200_ctx *c, const uint64_t *beta_mu, int is_corrector)
{
  if(!c || !beta_mu)
    return SDPB_B200_ERR_ARG;
  if(int rc = direction_ready(c, "compute_search_direction", true, is_corrector != 0))
    return rc;
  if(!c->have_residues)
    {
      c->error = "compute_search_direction called before direction_set_residues";
      return SDPB_B200_ERR_STATE;
    }
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->kt_used = 0;
  memcpy(c->dir_pinned, beta_mu, (size_t)c->es * 8);
  CUDA_TRY(c, cudaMemcpyAsync(c->dir_scal, c->dir_pinned, (size_t)c->es * 8, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaEventRecord(c->ev[9], c->stream));
  if(int rc = table_for(c->nl)->direction(c, 2, is_corrector))
    {
      cudaStreamSynchronize(c->stream);
      return rc;
    }
  CUDA_TRY(c, cudaEventRecord(c->ev[10], c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  cudaEventElapsedTime(&c->direction_ms, c->ev[9], c->ev[10]);
  c->have_direction = true;
  return 0;
}
extern "C" int sdpb_b200_direction_frobenius(sdpb_b200_ctx *c, uint64_t *block_products)
{
  if(!c)
    return SDPB_B200_ERR_ARG;
  if(int rc = direction_ready(c, "direction_frobenius", true, true))
    return rc;
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->kt_used = 0;
  if(int rc = table_for(c->nl)->direction(c, 3, 0))
    {
      cudaStreamSynchronize(c->stream);
      return rc;
    }
  return direction_scalars_out(c, block_products);
}
// Row N3: step_length.cxx:27-46 on the resident factors and direction.
// block_min_eigenvalues[b], b = 2j + parity: smallest eigenvalue of L_b^-1 dM_b L_b^-T with
// (L, dM) = (chol X, dX) for which = 0 and (chol Y, dY) for which = 1; empty blocks give 0.
extern "C" int sdpb_b200_step_length(sdpb_b200_ctx *c, int which, uint64_t *block_min_eigenvalues)
{
  if(!c || (which != 0 && which != 1))
    return SDPB_B200_ERR_ARG;
  if(int rc = direction_ready(c, "step_length", false, true))
    return rc;
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->kt_used = 0;
  CUDA_TRY(c, cudaEventRecord(c->ev[9], c->stream));
  if(int rc = table_for(c->nl)->step_length(c, which))
    {
      cudaStreamSynchronize(c->stream);
      return rc;
    }
  CUDA_TRY(c, cudaEventRecord(c->ev[10], c->stream));
  if(int rc = direction_scalars_out(c, block_min_eigenvalues))
    return rc;
  cudaEventElapsedTime(&c->step_length_ms, c->ev[9], c->ev[10]);
  return 0;
}
extern "C" float sdpb_b200_last_step_length_ms(const sdpb_b200_ctx *c) { return c ? c->step_length_ms : 0.f; }
// Laguerre steps the last sdpb_b200_step_length took per block-parity (2J ints); diagnostics
extern "C" int sdpb_b200_step_length_iterations(sdpb_b200_ctx *c, int *iterations)
{
  if(!c || !iterations)
    return SDPB_B200_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  if(c->J)
    CUDA_TRY(c, cudaMemcpy(iterations, c->eig_iter, (size_t)2 * c->J * sizeof(int), cudaMemcpyDeviceToHost));
  return 0;
}
extern "C" int sdpb_b200_direction_get(sdpb_b200_ctx *c, uint64_t *const *dx, uint64_t *const *dX, uint64_t *dy,
                                       uint64_t *const *dY)
{
  if(!c)
    return SDPB_B200_ERR_ARG;
  if(int rc = direction_ready(c, "direction_get", true, true))
    return rc;
  CUDA_TRY(c, cudaSetDevice(c->device));
  const size_t es = (size_t)c->es;
  std::vector<size_t> eXY(2 * c->J);
  for(int q = 0; q < 2 * c->J; ++q)
    eXY[q] = (size_t)c->g[q / 2].s[q % 2] * c->g[q / 2].s[q % 2];
  int rc = copy_blocks_out(c, c->dirDX, c->oXY, dX, 2 * c->J, eXY);
  if(!rc)
    rc = copy_blocks_out(c, c->dirDY, c->oXY, dY, 2 * c->J, eXY);
  if(rc)
    {
      cudaStreamSynchronize(c->stream);
      return rc;
    }
  if(dx && c->K)
    CUDA_TRY(c, cudaMemcpyAsync(c->dir_pinned, c->sol_x, (size_t)c->K * es * 8, cudaMemcpyDeviceToHost, c->stream));
  if(dy)
    CUDA_TRY(c, cudaMemcpyAsync(c->dir_pinned + (size_t)c->K * es, c->sol_y, (size_t)c->N * es * 8,
                                cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  if(dx)
    for(int j = 0; j < c->J; ++j)
      if(dx[j])
        memcpy(dx[j], c->dir_pinned + (size_t)c->g[j].row0 * es, (size_t)c->g[j].P * es * 8);
  if(dy)
    memcpy(dy, c->dir_pinned + (size_t)c->K * es, (size_t)c->N * es * 8);
  return 0;
}
// dX, dY (2J blocks each, the shape of X) from the host into the resident direction: for a caller
// that forms the direction itself and wants step_length on the device, and for tests.  Needs the
// factors of a step; a block list may be NULL (that object is left as it is).
extern "C" int sdpb_b200_direction_put(sdpb_b200_ctx *c, const uint64_t *const *dX, const uint64_t *const *dY)
{
  if(!c)
    return SDPB_B200_ERR_ARG;
  if(int rc = direction_ready(c, "direction_put", false, false))
    return rc;
  if(int rc = ensure_direction(c))
    return rc;
  CUDA_TRY(c, cudaSetDevice(c->device));
  int rc = dX ? copy_blocks_in(c, dX, c->dirDX) : 0;
  if(!rc && dY)
    rc = copy_blocks_in(c, dY, c->dirDY);
  cudaError_t e = cudaStreamSynchronize(c->stream);
  if(rc)
    return rc;
  CUDA_TRY(c, e);
  if(dX && dY)
    c->have_direction = true; // -XY and the residues are not touched
  return 0;
}
extern "C" float sdpb_b200_last_direction_ms(const sdpb_b200_ctx *c) { return c ? c->direction_ms : 0.f; }

// ----------------------------------------------------- resident step
// H2D of X and Y through one pinned staging buffer (two large copies instead
// of 4J small pageable ones).
extern "C" int sdpb_b200_upload_XY(sdpb_b200_ctx *c, const uint64_t *const *X,
                                   const uint64_t *const *Y)
{
  if(!c || !X || !Y)
    return SDPB_B200_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  if(c->pinned_words < 2 * c->wXY)
    {
      if(c->pinned)
        cudaFreeHost(c->pinned);
      c->pinned = nullptr;
      c->pinned_words = 0;
      CUDA_TRY(c, cudaMallocHost(&c->pinned, std::max<size_t>(8, 2 * c->wXY * 8)));
      c->pinned_words = 2 * c->wXY;
    }
  for(int which = 0; which < 2; ++which)
    {
      const uint64_t *const *A = which == 0 ? X : Y;
      uint64_t *dst = c->pinned + which * c->wXY;
      for(int q = 0; q < 2 * c->J; ++q)
        {
          const int s = c->g[q / 2].s[q % 2];
          if(s == 0)
            continue;
          if(!A[q])
            {
              c->error = "null input block " + std::to_string(q);
              return SDPB_B200_ERR_ARG;
            }
          memcpy(dst + c->oXY[q], A[q], (size_t)s * s * c->es * 8);
        }
    }
  if(c->wXY)
    {
      CUDA_TRY(c, cudaMemcpyAsync(c->Xin, c->pinned, c->wXY * 8,
                                  cudaMemcpyHostToDevice, c->stream));
      CUDA_TRY(c, cudaMemcpyAsync(c->Yin, c->pinned + c->wXY, c->wXY * 8,
                                  cudaMemcpyHostToDevice, c->stream));
    }
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// The whole hot path from device-resident X, Y (sdpb_b200_upload_XY): every
// kernel is enqueued back to back, one host synchronisation at the end.
// y_uploaded: Yin is still arriving on another stream (sdpb_b200_schur_step uploads Y on the copy
// stream while the X chain already runs); the event marks its arrival.  null: Yin is in place.
static int enqueue_step(sdpb_b200_ctx *c, cudaEvent_t y_uploaded = nullptr)
{
  cudaStream_t st = c->stream;
  c->kt_used = 0;
  c->have_factors = false;
  CUDA_TRY(c, cudaEventRecord(c->ev[9], st));
  // chol(Y) on a side stream, the A_Y chain on another, chol(X) -> A_X_inv here.  Y and LY are set
  // up on chol(Y)'s stream: nothing of the X chain reads them, so the X chain does not wait for Y.
  cudaStream_t sl = c->prio[1] ? c->urgent(1) : c->side(3), sx = c->urgent(0);
  if(c->wXY)
    CUDA_TRY(c, cudaMemcpyAsync(c->X, c->Xin, c->wXY * 8, cudaMemcpyDeviceToDevice, st));
  if(sl != st)
    {
      // (evf[1] is recorded again by dispatch_pairings later: a wait refers to the record that
      // precedes it, so sharing the event is safe)
      CUDA_TRY(c, cudaEventRecord(c->evf[1], st)); // everything enqueued before this step (uploads of X)
      CUDA_TRY(c, cudaStreamWaitEvent(sl, c->evf[1], 0));
    }
  if(y_uploaded)
    CUDA_TRY(c, cudaStreamWaitEvent(sl, y_uploaded, 0));
  if(c->wXY)
    {
      CUDA_TRY(c, cudaMemcpyAsync(c->Y, c->Yin, c->wXY * 8, cudaMemcpyDeviceToDevice, sl));
      CUDA_TRY(c, cudaMemcpyAsync(c->LY, c->Yin, c->wXY * 8, cudaMemcpyDeviceToDevice, sl));
    }
  CUDA_TRY(c, cudaEventRecord(c->evf[2], sl)); // Y, LY in place (what the A_Y chain waits for)
  c->cur = sl;
  int rc = dispatch_cholesky(c, 1);
  c->cur = st;
  if(rc)
    return rc;
  CUDA_TRY(c, cudaEventRecord(c->evd[0], sl));
  // chol(X) on its own urgent stream: evf[1] (recorded above on st) covers the copy of X
  if(sx != st)
    {
      if(sl == st) // (not reached: urgent streams exist together) evf[1] not recorded yet
        CUDA_TRY(c, cudaEventRecord(c->evf[1], st));
      CUDA_TRY(c, cudaStreamWaitEvent(sx, c->evf[1], 0));
    }
  c->cur = sx;
  rc = dispatch_cholesky(c, 0);
  c->cur = st;
  if(rc)
    return rc;
  CUDA_TRY(c, c->after(sx, st, 16));
  CUDA_TRY(c, cudaEventRecord(c->ev[0], st));
  rc = dispatch_pairings(c, 2);
  if(rc)
    return rc;
  CUDA_TRY(c, c->after(sl, st, 3));
  CUDA_TRY(c, cudaEventRecord(c->ev[1], st));
  rc = dispatch_schur_and_Q(c);
  if(rc)
    return rc;
  // (the P bands are final after normalize_kernel, which restores them in the same pass: ev[5])
  CUDA_TRY(c, cudaEventRecord(c->ev[10], st));
  return 0;
}

// wait for the step, read the status words back and turn them into the reference's errors
static int finish_step(sdpb_b200_ctx *c)
{
  cudaStream_t st = c->stream;
  const int J = c->J;
  std::vector<int> status(5 * J + 1);
  int flags[4], fail[4] = {0, 0, 0, 0};
  CUDA_TRY(c, cudaMemcpyAsync(status.data(), c->d_status, status.size() * sizeof(int),
                              cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaMemcpyAsync(flags, c->d_flags, sizeof(flags), cudaMemcpyDeviceToHost, st));
  if(c->sharded())
    CUDA_TRY(c, cudaMemcpyAsync(fail, c->d_fail, sizeof(int), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(c, cudaStreamSynchronize(st));
  cudaEventElapsedTime(&c->stage_ms[0], c->ev[9], c->ev[0]);
  cudaEventElapsedTime(&c->stage_ms[1], c->ev[0], c->ev[1]);
  for(int k = 2; k < 8; ++k)
    cudaEventElapsedTime(&c->stage_ms[k], c->ev[k], c->ev[k + 1]);
  cudaEventElapsedTime(&c->stage_ms[8], c->ev[9], c->ev[10]);
  for(int which = 0; which < 2; ++which)
    for(int q = 0; q < 2 * J; ++q)
      if(status[which * 2 * J + q] >= 0)
        {
          c->error = std::string("Error when computing Cholesky decomposition of "
                                 "Block_Diagonal_Matrix ")
                     + (which == 0 ? "X" : "Y")
                     + ", block index = " + std::to_string(c->gidx[q / 2])
                     + ", parity = " + std::to_string(q % 2)
                     + ": non-positive pivot "
                     + std::to_string(status[which * 2 * J + q]);
          return SDPB_B200_ERR_NOT_HPD;
        }
  for(int j = 0; j < J; ++j)
    if(status[4 * J + j] >= 0)
      {
        c->error = "Error when computing Cholesky decomposition of block_"
                   + std::to_string(c->gidx[j]) + ": non-positive pivot "
                   + std::to_string(status[4 * J + j]);
        return SDPB_B200_ERR_NOT_HPD;
      }
  if(flags[0])
    {
      c->error = "normalised P entry does not fit the integer syrk format";
      return SDPB_B200_ERR_Q_DIAG;
    }
  if(flags[1] != INT_MAX)
    {
      c->error = "Normalized Q should have ones on diagonal. For i = "
                 + std::to_string(flags[1]);
      return SDPB_B200_ERR_Q_DIAG;
    }
  if(status[5 * J] >= 0)
    {
      c->error = "Error when computing Cholesky(Q): non-positive pivot "
                 + std::to_string(status[5 * J]);
      return SDPB_B200_ERR_NOT_HPD;
    }
  if(fail[0] > 0)
    {
      // this rank's blocks are fine, but a peer's failed: its contributions to the column norms
      // and to Q are garbage, so Q is on every rank
      c->error = "Schur-complement step failed on " + std::to_string(fail[0])
                 + " peer rank(s) (their last_error names the block); Q is not usable";
      return SDPB_B200_ERR_NOT_HPD;
    }
  c->have_X_cholesky = c->have_pairings = c->have_factors = true;
  return 0;
}

extern "C" int sdpb_b200_schur_step_resident(sdpb_b200_ctx *c)
{
  if(!c)
    return SDPB_B200_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  const int rc = enqueue_step(c);
  return rc ? rc : finish_step(c);
}

// D2H of whichever outputs of the last step the caller wants (NULL = skip).
extern "C" int sdpb_b200_download(
  sdpb_b200_ctx *c, uint64_t *const *X_cholesky, uint64_t *const *Y_cholesky,
  uint64_t *const *A_X_inv, uint64_t *const *A_Y,
  uint64_t *const *schur_complement_cholesky,
  uint64_t *const *schur_off_diagonal, uint64_t *Q)
{
  if(!c)
    return SDPB_B200_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  const int J = c->J;
  std::vector<size_t> eXY(2 * J), eA(2 * J), eS(J), eP(J);
  for(int q = 0; q < 2 * J; ++q)
    {
      eXY[q] = (size_t)c->g[q / 2].s[q % 2] * c->g[q / 2].s[q % 2];
      eA[q] = (size_t)c->g[q / 2].mn * c->g[q / 2].mn;
    }
  for(int j = 0; j < J; ++j)
    {
      eS[j] = (size_t)c->g[j].P * c->g[j].P;
      eP[j] = (size_t)c->g[j].P * c->N;
    }
  int rc = copy_blocks_out(c, c->X, c->oXY, X_cholesky, 2 * J, eXY);
  if(!rc)
    rc = copy_blocks_out(c, c->LY, c->oXY, Y_cholesky, 2 * J, eXY);
  if(!rc)
    rc = copy_blocks_out(c, c->AX, c->oA, A_X_inv, 2 * J, eA);
  if(!rc)
    rc = copy_blocks_out(c, c->AY, c->oA, A_Y, 2 * J, eA);
  if(!rc)
    rc = copy_blocks_out(c, c->S, c->oS, schur_complement_cholesky, J, eS);
  if(!rc)
    rc = copy_blocks_out(c, c->Pband, c->oB, schur_off_diagonal, J, eP);
  if(rc)
    return rc;
  if(Q)
    CUDA_TRY(c, cudaMemcpyAsync(Q, c->Q, (size_t)c->N * c->N * c->es * 8,
                                cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int sdpb_b200_schur_step(
  sdpb_b200_ctx *c, const uint64_t *const *X, const uint64_t *const *Y,
  uint64_t *const *X_cholesky, uint64_t *const *Y_cholesky,
  uint64_t *const *A_X_inv, uint64_t *const *A_Y,
  uint64_t *const *schur_complement_cholesky,
  uint64_t *const *schur_off_diagonal, uint64_t *Q, int32_t *block_timings_ms)
{
  // One pipeline: X, Y go up block by block straight from the caller's buffers, the whole
  // step is enqueued without a host synchronisation, and every output is copied back on a
  // separate stream as soon as the kernel that finishes it has run -- L_j while L_j^-1 B_j
  // is still being solved, P (the largest, P x N) beside the exact syrk and Cholesky(Q).
  if(!c || !X || !Y)
    return SDPB_B200_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  const int J = c->J;
  // every pointer is checked before the first copy is enqueued: an error return must not leave
  // DMA from the caller's buffers in flight
  for(int which = 0; which < 2; ++which)
    for(int q = 0; q < 2 * J; ++q)
      if(c->g[q / 2].s[q % 2] && !(which == 0 ? X : Y)[q])
        {
          c->error = "null input block " + std::to_string(q);
          return SDPB_B200_ERR_ARG;
        }
  // (both streams are idle here: every call of this library ends with their synchronisation)
  for(int which = 0; which < 2; ++which)
    {
      const uint64_t *const *A = which == 0 ? X : Y;
      limb_t *dst = which == 0 ? c->Xin : c->Yin;
      for(int q = 0; q < 2 * J;)
        {
          const int s = c->g[q / 2].s[q % 2];
          if(s == 0)
            {
              ++q;
              continue;
            }
          size_t words = (size_t)s * s * c->es;
          int r = q + 1; // merge blocks the caller stored back to back
          while(r < 2 * J)
            {
              const size_t sr = c->g[r / 2].s[r % 2];
              if(sr == 0 || A[r] != A[q] + words || c->oXY[r] != c->oXY[q] + words)
                break;
              words += sr * sr * c->es;
              ++r;
            }
          // X on the main stream, Y on the copy stream: the X chain (chol X, L_X^-1 V, A_X_inv) starts
          // as soon as X is up and runs under Y's upload
          CUDA_TRY(c, cudaMemcpyAsync(dst + c->oXY[q], A[q], words * 8, cudaMemcpyHostToDevice,
                                      which == 0 ? c->stream : c->copy));
          q = r;
        }
    }
  CUDA_TRY(c, cudaEventRecord(c->evd[1], c->copy)); // Y uploaded
  int rc = enqueue_step(c, c->evd[1]);
  if(rc)
    {
      cudaStreamSynchronize(c->copy);
      cudaStreamSynchronize(c->stream); // the uploads still read the caller's buffers
      return rc;
    }
  {
    std::vector<size_t> eXY(2 * J), eA(2 * J), eS(J), eP(J);
    for(int q = 0; q < 2 * J; ++q)
      {
        eXY[q] = (size_t)c->g[q / 2].s[q % 2] * c->g[q / 2].s[q % 2];
        eA[q] = (size_t)c->g[q / 2].mn * c->g[q / 2].mn;
      }
    for(int j = 0; j < J; ++j)
      {
        eS[j] = (size_t)c->g[j].P * c->g[j].P;
        eP[j] = (size_t)c->g[j].P * c->N;
      }
    cudaStream_t main_stream = c->stream;
    c->stream = c->copy; // copy_blocks_out enqueues on c->stream
    auto out = [&](cudaEvent_t ready, const limb_t *src, const std::vector<size_t> &off,
                   uint64_t *const *dst, int count, const std::vector<size_t> &elems) -> int {
      if(!dst)
        return 0;
      if(cudaStreamWaitEvent(c->copy, ready, 0) != cudaSuccess)
        return SDPB_B200_ERR_CUDA;
      return copy_blocks_out(c, src, off, dst, count, elems);
    };
    rc = out(c->ev[0], c->X, c->oXY, X_cholesky, 2 * J, eXY);
    if(!rc)
      rc = out(c->evd[0], c->LY, c->oXY, Y_cholesky, 2 * J, eXY);
    if(!rc)
      rc = out(c->ev[1], c->AX, c->oA, A_X_inv, 2 * J, eA);
    if(!rc)
      rc = out(c->ev[1], c->AY, c->oA, A_Y, 2 * J, eA);
    if(!rc)
      rc = out(c->ev[4], c->S, c->oS, schur_complement_cholesky, J, eS);
    if(!rc)
      rc = out(c->ev[5], c->Pband, c->oB, schur_off_diagonal, J, eP);
    if(!rc && Q)
      {
        if(cudaStreamWaitEvent(c->copy, c->ev[10], 0) != cudaSuccess
           || cudaMemcpyAsync(Q, c->Q, (size_t)c->N * c->N * c->es * 8, cudaMemcpyDeviceToHost, c->copy)
                != cudaSuccess)
          rc = SDPB_B200_ERR_CUDA;
      }
    c->stream = main_stream;
    if(rc)
      {
        cudaStreamSynchronize(c->copy);
        cudaStreamSynchronize(c->stream);
        if(c->error.empty())
          c->error = "CUDA failure while enqueueing the output copies";
        return rc;
      }
  }
  rc = finish_step(c);
  CUDA_TRY(c, cudaStreamSynchronize(c->copy));
  if(rc)
    return rc;
  add_block_timings(c, block_timings_ms);
  return 0;
}

// The diagonals of the Cholesky factors of the last step, which is all that
// update_cond_numbers (step.cxx:187-189, sdpb_util/cholesky_condition_number.hxx:8-36) reads of
// L_j and chol(Q): with the Schur solves on the device (solve_schur_complement_equation above)
// neither L_j (sum P_j^2 elements) nor L_j^-1 B_j (P x N) has to cross PCIe any more.
// X_diag / Y_diag: stacked over b = 2j + parity (sum of psd sizes); S_diag: stacked over j (P
// elements); Q_diag: N elements.  Each may be NULL.
extern "C" int sdpb_b200_cholesky_diagonals(sdpb_b200_ctx *c, uint64_t *X_diag, uint64_t *Y_diag,
                                            uint64_t *S_diag, uint64_t *Q_diag)
{
  if(!c)
    return SDPB_B200_ERR_ARG;
  if(!c->have_factors)
    {
      c->error = "cholesky_diagonals called before a successful Schur-complement step";
      return SDPB_B200_ERR_STATE;
    }
  CUDA_TRY(c, cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int count = 5 * c->J + 1; // 2J X, 2J Y, J S, Q
  diag_gather_kernel<<<std::min(count, 148 * 8), 128, 0, st>>>(c->d_diag, count, c->es, c->diag_buf);
  ++c->launches;
  CUDA_TRY(c, cudaGetLastError());
  uint64_t *dst[4] = {X_diag, Y_diag, S_diag, Q_diag};
  for(int k = 0; k < 4; ++k)
    {
      const long n = (k < 3 ? c->diag_off[k + 1] : c->diag_total) - c->diag_off[k];
      if(dst[k] && n)
        CUDA_TRY(c, cudaMemcpyAsync(dst[k], c->diag_buf + (size_t)c->diag_off[k] * c->es, (size_t)n * c->es * 8,
                                    cudaMemcpyDeviceToHost, st));
    }
  CUDA_TRY(c, cudaStreamSynchronize(st));
  return 0;
}

// 0: every kernel on one stream in program order (the mode per-kernel timings are taken in);
// 1 (default): independent chains on side streams, the S chain in interleaved groups
extern "C" int sdpb_b200_set_concurrency(sdpb_b200_ctx *c, int level)
{
  if(!c || level < 0 || level > 1)
    return SDPB_B200_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  CUDA_TRY(c, cudaDeviceSynchronize());
  c->concurrency = level;
  return 0;
}

extern "C" int sdpb_b200_last_timings_ms(const sdpb_b200_ctx *c, float *ms, int n)
{
  if(!c || !ms)
    return SDPB_B200_ERR_ARG;
  for(int i = 0; i < n && i < 9; ++i)
    ms[i] = c->stage_ms[i];
  return 0;
}

// element-wise scalar ops on the device, for parity tests of mpfx itself
extern "C" int sdpb_b200_scalar_op(sdpb_b200_ctx *c, int op, int k, long count,
                                   const uint64_t *a, const uint64_t *b,
                                   uint64_t *r)
{
  if(!c || count < 0)
    return SDPB_B200_ERR_ARG;
  CUDA_TRY(c, cudaSetDevice(c->device));
  c->kt_used = 0;
  c->cur = c->stream;
  limb_t *da = nullptr, *db = nullptr, *dr = nullptr;
  const size_t bytes = std::max<size_t>(16, (size_t)count * c->es * 8);
  struct Guard // frees the temporaries on every return path
  {
    limb_t *&a, *&b, *&r;
    cudaStream_t st;
    ~Guard()
    {
      cudaStreamSynchronize(st);
      cudaFree(a);
      cudaFree(b);
      cudaFree(r);
    }
  } guard{da, db, dr, c->stream};
  CUDA_TRY(c, cudaMalloc(&da, bytes));
  CUDA_TRY(c, cudaMalloc(&db, bytes));
  CUDA_TRY(c, cudaMalloc(&dr, bytes));
  CUDA_TRY(c, cudaMemcpyAsync(da, a, (size_t)count * c->es * 8, cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(c, cudaMemcpyAsync(db, b, (size_t)count * c->es * 8, cudaMemcpyHostToDevice, c->stream));
  int rc = dispatch_scalar(c, op, k, count, da, db, dr);
  if(rc)
    return rc;
  CUDA_TRY(c, cudaMemcpyAsync(r, dr, (size_t)count * c->es * 8, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(c, cudaStreamSynchronize(c->stream));
  return 0;
}

// ------------------------------------------------ mpf <-> packed (host only)
extern "C" void sdpb_b200_pack_mpf(int prec_bits, int mp_size, long mp_exp,
                                   const uint64_t *mp_d, uint64_t *out)
{
  const int nl = mpfx::stored_limbs(prec_bits), ew = (nl + 2) & ~1;
  for(int i = 0; i < ew; ++i)
    out[i] = 0;
  int asz = mp_size < 0 ? -mp_size : mp_size;
  if(asz == 0)
    return;
  const uint64_t *src = mp_d;
  if(asz > nl) // cannot happen for an mpf of this precision; keep the top
    {
      src += asz - nl;
      asz = nl;
    }
  const int32_t sign = mp_size < 0 ? -1 : 1;
  out[0] = (uint64_t)(uint32_t)(int32_t)mp_exp | ((uint64_t)(uint32_t)sign << 32);
  for(int i = 0; i < asz; ++i)
    out[1 + nl - asz + i] = src[i];
}
extern "C" int sdpb_b200_unpack_mpf(int prec_bits, const uint64_t *in,
                                    uint64_t *mp_d, long *mp_exp)
{
  const int nl = mpfx::stored_limbs(prec_bits);
  const int32_t e = (int32_t)(uint32_t)in[0];
  const int32_t sign = (int32_t)(uint32_t)(in[0] >> 32);
  if(sign == 0)
    {
      *mp_exp = 0;
      return 0;
    }
  int lo = 0;
  while(lo < nl - 1 && in[1 + lo] == 0)
    lo++;
  const int asz = nl - lo;
  for(int i = 0; i < asz; ++i)
    mp_d[i] = in[1 + lo + i];
  *mp_exp = e;
  return sign < 0 ? -asz : asz;
}

extern "C" int sdpb_b200_host_alloc(void **p, size_t bytes)
{
  return cudaMallocHost(p, bytes) == cudaSuccess ? 0 : SDPB_B200_ERR_CUDA;
}
extern "C" void sdpb_b200_host_free(void *p) { cudaFreeHost(p); }

extern "C" int sdpb_b200_kernel_timings(const sdpb_b200_ctx *c, int max,
                                        const char **names, float *ms)
{
  if(!c)
    return 0;
  int n = 0;
  for(int i = 0; i < c->kt_used && n < max; ++i, ++n)
    {
      names[n] = c->kt[i].name;
      ms[n] = 0;
      cudaEventElapsedTime(&ms[n], c->kt[i].e0, c->kt[i].e1);
    }
  return n;
}

extern "C" long sdpb_b200_kernel_launches(const sdpb_b200_ctx *c)
{
  return c ? c->launches : 0;
}

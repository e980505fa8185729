// Hot_Path over a table of C entry points with the signatures of
// include/sdpb_b200.h.  The product binds the table to the sdpb_b200_* symbols
// of libsdpb_b200.so (hot_path_b200.cpp); the tests bind the same adaptor to the
// CPU oracle's oracle_* symbols, so both run through identical host code.
// This is the shim INTEGRATION.md describes for the reference side: pack
// El::BigFloat -> packed elements, call, unpack, turn rc != 0 into RUNTIME_ERROR.
#pragma once
#include "solver.hpp"

namespace sdpb_host
{
struct Hot_Path_Table
{
  void *ctx = nullptr;
  int (*set_block)(void *, int, const uint64_t *, const uint64_t *, const uint64_t *) = nullptr;
  int (*cholesky_decomposition)(void *, int, const uint64_t *const *, uint64_t *const *) = nullptr;
  int (*compute_bilinear_pairings)(void *, const uint64_t *const *, uint64_t *const *, uint64_t *const *)
    = nullptr;
  int (*initialize_schur_complement_solver)(void *, uint64_t *const *, uint64_t *const *, uint64_t *, int32_t *)
    = nullptr;
  int (*solve_schur_complement_equation)(void *, uint64_t *const *, uint64_t *) = nullptr;
  int (*scale_multiply_add)(void *, int, const uint64_t *const *, const uint64_t *const *, int, uint64_t *const *)
    = nullptr;
  // rows N2 (sdpb_b200_direction_* / oracle_direction_*); all six or none
  int (*direction_begin)(void *, uint64_t *) = nullptr;
  int (*direction_R_errors)(void *, const uint64_t *, uint64_t *) = nullptr;
  int (*direction_set_residues)(void *, const uint64_t *const *, const uint64_t *const *, const uint64_t *) = nullptr;
  int (*compute_search_direction)(void *, const uint64_t *, int) = nullptr;
  int (*direction_frobenius)(void *, uint64_t *) = nullptr;
  int (*direction_get)(void *, uint64_t *const *, uint64_t *const *, uint64_t *, uint64_t *const *) = nullptr;
  // row N3 (sdpb_b200_step_length / oracle_step_length); optional
  int (*step_length)(void *, int, uint64_t *) = nullptr;
  const char *(*last_error)(const void *) = nullptr;
  void (*destroy)(void *) = nullptr;
  std::string name;
};

class Hot_Path_C : public Hot_Path
{
  Hot_Path_Table t;
  const Block_Info &bi;
  int N;
  typedef std::vector<uint64_t> Buf;
  std::vector<Buf> in2J, inB2J, out2J, outJ_L, ioJ_dx;
  Buf outQ, io_dy;
  std::vector<int32_t> block_timings_ms;

  static std::vector<const uint64_t *> cptrs(const std::vector<Buf> &v)
  {
    std::vector<const uint64_t *> p(v.size());
    for(size_t i = 0; i < v.size(); ++i)
      p[i] = v[i].empty() ? nullptr : v[i].data();
    return p;
  }
  static std::vector<uint64_t *> ptrs(std::vector<Buf> &v)
  {
    std::vector<uint64_t *> p(v.size());
    for(size_t i = 0; i < v.size(); ++i)
      p[i] = v[i].empty() ? nullptr : v[i].data();
    return p;
  }
  void check(int rc) const
  {
    if(rc)
      throw std::runtime_error(t.last_error(t.ctx));
  }
  void pack_psd(const std::vector<Matrix> &A)
  {
    const size_t ew = (size_t)elem_words();
#pragma omp parallel for schedule(dynamic)
    for(size_t b = 0; b < A.size(); ++b)
      {
        in2J[b].resize(A[b].a.size() * ew);
        if(!A[b].a.empty())
          pack_matrix(A[b], in2J[b].data());
      }
  }

public:
  Hot_Path_C(const Hot_Path_Table &table, const Block_Info &block_info, const SDP &sdp)
      : t(table), bi(block_info), N(sdp.N())
  {
    const int J = bi.num_blocks();
    const size_t ew = (size_t)elem_words();
    in2J.resize(2 * J);
    inB2J.resize(2 * J);
    out2J.resize(2 * J);
    outJ_L.resize(J);
    ioJ_dx.resize(J);
    block_timings_ms.assign(J, 0);
    for(int j = 0; j < J; ++j)
      {
        Buf B(sdp.free_var_matrix[j].a.size() * ew), e(sdp.bilinear_bases[2 * j].a.size() * ew),
          o(sdp.bilinear_bases[2 * j + 1].a.size() * ew);
        if(!B.empty())
          pack_matrix(sdp.free_var_matrix[j], B.data());
        if(!e.empty())
          pack_matrix(sdp.bilinear_bases[2 * j], e.data());
        if(!o.empty())
          pack_matrix(sdp.bilinear_bases[2 * j + 1], o.data());
        // never hand out NULL for an empty matrix: the callee may form base + 0
        B.resize(B.size() + 1);
        e.resize(e.size() + 1);
        o.resize(o.size() + 1);
        check(t.set_block(t.ctx, j, B.data(), e.data(), o.data()));
      }
    outQ.resize((size_t)N * N * ew);
  }
  ~Hot_Path_C() override
  {
    if(t.destroy && t.ctx)
      t.destroy(t.ctx);
  }
  std::string name() const override { return t.name; }

  void cholesky_decomposition(int which, const std::vector<Matrix> &A, std::vector<Matrix> &L) override
  {
    const size_t ew = (size_t)elem_words();
    pack_psd(A);
    for(size_t b = 0; b < A.size(); ++b)
      out2J[b].resize(A[b].a.size() * ew);
    const auto in = cptrs(in2J);
    const auto out = ptrs(out2J);
    check(t.cholesky_decomposition(t.ctx, which, in.data(), out.data()));
    L.resize(A.size());
#pragma omp parallel for schedule(dynamic)
    for(size_t b = 0; b < A.size(); ++b)
      {
        if(A[b].h)
          unpack_matrix(L[b], A[b].h, A[b].w, out2J[b].data());
        else
          L[b].resize(0, 0);
      }
  }

  void compute_bilinear_pairings(const std::vector<Matrix> &Y, std::vector<Matrix> &A_Y) override
  {
    const size_t ew = (size_t)elem_words();
    pack_psd(Y);
    const int J = bi.num_blocks();
    for(int b = 0; b < 2 * J; ++b)
      {
        const size_t mn = (size_t)bi.bilinear_pairing_block_size(b / 2);
        out2J[b].resize(mn * mn * ew);
      }
    const auto in = cptrs(in2J);
    const auto out = ptrs(out2J);
    check(t.compute_bilinear_pairings(t.ctx, in.data(), nullptr, out.data()));
    A_Y.resize(2 * J);
#pragma omp parallel for schedule(dynamic)
    for(int b = 0; b < 2 * J; ++b)
      {
        const int mn = bi.bilinear_pairing_block_size(b / 2);
        unpack_matrix(A_Y[b], mn, mn, out2J[b].data());
      }
  }

  void initialize_schur_complement_solver(std::vector<Matrix> &L, std::vector<Matrix> &P, Matrix &Q) override
  {
    const size_t ew = (size_t)elem_words();
    const int J = bi.num_blocks();
    for(int j = 0; j < J; ++j)
      {
        const size_t Pj = (size_t)bi.schur_block_size(j);
        outJ_L[j].resize(Pj * Pj * ew);
      }
    const auto pl = ptrs(outJ_L);
    // L_j and chol(Q) come back for the condition numbers (update_cond_numbers, step.cxx:187-189); L_j^-1 B_j is
    // read only by the Schur solve and stays with the implementation
    check(t.initialize_schur_complement_solver(t.ctx, pl.data(), nullptr, outQ.data(),
                                               block_timings_ms.data()));
    L.resize(J);
    P.clear();
#pragma omp parallel for schedule(dynamic)
    for(int j = 0; j < J; ++j)
      {
        const int Pj = bi.schur_block_size(j);
        unpack_matrix(L[j], Pj, Pj, outJ_L[j].data());
      }
    unpack_matrix(Q, N, N, outQ.data());
  }

  void scale_multiply_add(int alpha, const std::vector<Matrix> &A, const std::vector<Matrix> &B, int beta,
                          std::vector<Matrix> &C) override
  {
    const size_t ew = (size_t)elem_words();
    pack_psd(A);
#pragma omp parallel for schedule(dynamic)
    for(size_t b = 0; b < B.size(); ++b)
      {
        inB2J[b].resize(B[b].a.size() * ew);
        if(!B[b].a.empty())
          pack_matrix(B[b], inB2J[b].data());
        out2J[b].resize(A[b].a.size() * ew);
        if(beta && !C[b].a.empty())
          pack_matrix(C[b], out2J[b].data());
      }
    const auto pa = cptrs(in2J), pb = cptrs(inB2J);
    const auto pc = ptrs(out2J);
    check(t.scale_multiply_add(t.ctx, alpha, pa.data(), pb.data(), beta, pc.data()));
    C.resize(A.size());
#pragma omp parallel for schedule(dynamic)
    for(size_t b = 0; b < A.size(); ++b)
      {
        if(A[b].h)
          unpack_matrix(C[b], A[b].h, A[b].w, out2J[b].data());
        else
          C[b].resize(0, 0);
      }
  }

  // ---- rows N2: the search direction stays with the implementation ----
  bool resident_direction() const override { return t.compute_search_direction != nullptr; }
  void scalars_in(std::vector<BigFloat> &out, const Buf &buf) const
  {
    const size_t ew = (size_t)elem_words();
    out.assign(2 * (size_t)bi.num_blocks(), BigFloat());
    for(size_t b = 0; b < out.size(); ++b)
      unpack(out[b], buf.data() + b * ew);
  }
  void direction_begin(std::vector<BigFloat> &block_traces) override
  {
    Buf buf(2 * (size_t)bi.num_blocks() * elem_words() + 1);
    check(t.direction_begin(t.ctx, buf.data()));
    scalars_in(block_traces, buf);
  }
  void direction_R_errors(const BigFloat &mu, std::vector<BigFloat> &maxima) override
  {
    Buf buf(2 * (size_t)bi.num_blocks() * elem_words() + 1), m((size_t)elem_words());
    pack(mu, m.data());
    check(t.direction_R_errors(t.ctx, m.data(), buf.data()));
    scalars_in(maxima, buf);
  }
  void direction_set_residues(const std::vector<Matrix> &primal_residues, const std::vector<Matrix> &dual_residues,
                              const Matrix &primal_residue_p) override
  {
    const size_t ew = (size_t)elem_words();
    const int J = bi.num_blocks();
    pack_psd(primal_residues);
    for(int j = 0; j < J; ++j)
      {
        ioJ_dx[j].resize(dual_residues[j].a.size() * ew + 1);
        if(!dual_residues[j].a.empty())
          pack_matrix(dual_residues[j], ioJ_dx[j].data());
      }
    io_dy.resize((size_t)N * ew + 1);
    pack_matrix(primal_residue_p, io_dy.data());
    const auto pr = cptrs(in2J);
    std::vector<const uint64_t *> pd(J);
    for(int j = 0; j < J; ++j)
      pd[j] = ioJ_dx[j].data();
    check(t.direction_set_residues(t.ctx, pr.data(), pd.data(), io_dy.data()));
  }
  void compute_search_direction(const BigFloat &beta_mu, bool is_corrector) override
  {
    Buf bm((size_t)elem_words());
    pack(beta_mu, bm.data());
    check(t.compute_search_direction(t.ctx, bm.data(), is_corrector ? 1 : 0));
  }
  void direction_frobenius(std::vector<BigFloat> &products) override
  {
    Buf buf(2 * (size_t)bi.num_blocks() * elem_words() + 1);
    check(t.direction_frobenius(t.ctx, buf.data()));
    scalars_in(products, buf);
  }
  void direction_get(std::vector<Matrix> &dx, std::vector<Matrix> &dX, Matrix &dy, std::vector<Matrix> &dY) override
  {
    const size_t ew = (size_t)elem_words();
    const int J = bi.num_blocks();
    for(int j = 0; j < J; ++j)
      ioJ_dx[j].resize((size_t)bi.schur_block_size(j) * ew + 1);
    io_dy.resize((size_t)N * ew + 1);
    for(int b = 0; b < 2 * J; ++b)
      {
        const size_t s = (size_t)bi.psd_matrix_block_size(b / 2, b % 2);
        out2J[b].resize(s * s * ew);
        inB2J[b].resize(s * s * ew);
      }
    const auto px = ptrs(ioJ_dx), pX = ptrs(out2J), pY = ptrs(inB2J);
    check(t.direction_get(t.ctx, px.data(), pX.data(), io_dy.data(), pY.data()));
    dx.resize(J);
    dX.resize(2 * J);
    dY.resize(2 * J);
    for(int j = 0; j < J; ++j)
      unpack_matrix(dx[j], bi.schur_block_size(j), 1, ioJ_dx[j].data());
#pragma omp parallel for schedule(dynamic)
    for(int b = 0; b < 2 * J; ++b)
      {
        const int s = bi.psd_matrix_block_size(b / 2, b % 2);
        if(s)
          {
            unpack_matrix(dX[b], s, s, out2J[b].data());
            unpack_matrix(dY[b], s, s, inB2J[b].data());
          }
        else
          {
            dX[b].resize(0, 0);
            dY[b].resize(0, 0);
          }
      }
    unpack_matrix(dy, N, 1, io_dy.data());
  }

  bool step_length_min_eigenvalues(int which, std::vector<BigFloat> &mins) override
  {
    if(!t.step_length)
      return false;
    Buf buf(2 * (size_t)bi.num_blocks() * elem_words() + 1);
    check(t.step_length(t.ctx, which, buf.data()));
    scalars_in(mins, buf);
    return true;
  }

  void solve_schur_complement_equation(std::vector<Matrix> &dx, Matrix &dy) override
  {
    const size_t ew = (size_t)elem_words();
    const int J = bi.num_blocks();
    for(int j = 0; j < J; ++j)
      {
        ioJ_dx[j].resize(dx[j].a.size() * ew + 1);
        if(!dx[j].a.empty())
          pack_matrix(dx[j], ioJ_dx[j].data());
      }
    io_dy.resize((size_t)N * ew + 1);
    pack_matrix(dy, io_dy.data());
    const auto px = ptrs(ioJ_dx);
    check(t.solve_schur_complement_equation(t.ctx, px.data(), io_dy.data()));
    for(int j = 0; j < J; ++j)
      if(!dx[j].a.empty())
        unpack_matrix(dx[j], dx[j].h, 1, ioJ_dx[j].data());
    unpack_matrix(dy, N, 1, io_dy.data());
  }
};
} // namespace sdpb_host

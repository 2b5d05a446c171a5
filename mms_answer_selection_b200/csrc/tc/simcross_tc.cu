// SimCross mode 2 (S[n,k] = Q_n M_k A_n^T + B_k) on tcgen05 tensor cores (TF32 multiply, fp32
// accumulate in TMEM).  Reference semantics: src/caffe/layers/sim_cross_layer.cpp:140-161
// (forward) and :251-307 (backward); the reference runs 2 (fwd) / 6 (bwd) cblas_sgemm calls
// per (pair, measure) on the host.
//
// Here the batch is flattened so that every contraction with a D x D operand is ONE GEMM
// over all N*L token rows, and the per-pair contractions are one batched launch each:
//
//   forward   T[k]   = Qall M_k                (N*Lq x D x D,  batch k)
//             S[n,k] = T[k][n] A_n^T + B_k     (Lq x La x D,   batch (k, n), bias fused)
//   backward  U[k][n] = dS[n,k] A_n            (Lq x D x La,   batch (k, n))
//             dM_k   += Qall^T U[k]            (D x D x N*Lq,  batch k, split-K, atomics)
//             dQall   = sum_k U[k] M_k^T       (N*Lq x D x mc*D, k as reduction segments)
//             T[k]    = Qall M_k               (recomputed unless MMS_OPT_REUSE_FORWARD lets the backward
//                                               read what the last forward left in the workspace)
//             dA_n    = sum_k dS[n,k]^T T[k][n] (La x D x mc*Lq, batch n, k as segments)
//
// No operand is ever transposed in memory: the UMMA descriptors take K-major and MN-major
// tiles alike (tc_gemm.cu).  T / U live in the handle's scratch buffer; N is processed in
// chunks when mc*N*L*D floats exceed MMS_OPT_SCRATCH_BYTES.
#include "../mms_common.cuh"
#include "tc_gemm.cuh"

// Every operand of the contractions below is a TF32-rounded copy in the handle's scratch buffer
// (external inputs q, a, M, dS pass through one rounding/repacking launch; T and U are rounded by
// the epilogue that produces them), with leading dimensions padded to 4 floats, so that all of
// them are fetched by TMA whatever D / La are.
namespace {

struct Plan {
  int Dp, Lap;               // padded leading dimensions of the scratch copies
  size_t per_pair, fixed;    // scratch floats per QA pair / independent of N
  int nc_max;
};

Plan make_plan(mms_context* ctx, int N, int Lq, int La, int D, int mc, bool backward) {
  Plan p;
  // rows of the scratch copies start on 128-byte lines: every 32-float row segment a TMA box fetches is then
  // one L2 line instead of straddling two (measured: the per-SM TMA ingest rate is what bounds these kernels)
  p.Dp = (int)((D + 31) & ~31); p.Lap = (int)tc_pad4(La);
  const int Lmax = mms_max(Lq, La);
  p.per_pair = (size_t)(Lq + La) * p.Dp + (size_t)mc * (backward ? Lmax : Lq) * p.Dp +
               (backward ? (size_t)mc * Lq * p.Lap : 0);
  p.fixed = (size_t)mc * D * p.Dp;
  long long c = ((long long)(ctx->scratch_cap / sizeof(float)) - (long long)p.fixed) / (long long)p.per_pair;
  p.nc_max = (int)mms_max<long long>(1, mms_min<long long>(c, N));
  return p;
}

// T[k][row][:] = Qall[row][:] M_k      (row = chunk-local token row; result rounded for reuse)
int gemm_T(mms_context* ctx, const float* qr, const float* Mr, float* Tk, int rows, int D, int Dp, int mc) {
  TcGemmArgs g = tc_gemm_args(qr, Dp, 0, Mr, Dp, 1, Tk, Dp, rows, D, D);
  g.nb1 = mc; g.sB1 = (long long)D * Dp; g.sC1 = (long long)rows * Dp;
  g.operands_tf32 = 1; g.round_out = 1;
  return mms_tc_gemm(ctx, g);
}

}  // namespace

// Rounds M into the place the next forward on this handle will look for it (the head of the scratch buffer): M does not
// depend on the step's inputs, so a net can do this on a side stream beside the Embed gathers instead of in front of the
// forward contraction (8 us of a 95 us step at 50 pairs).  Does nothing when the buffer does not exist yet.
int mms_tc_simcross2_prepare(mms_context* ctx, const float* Mw, int D, int mc) {
  const int Dp = (D + 31) & ~31;
  ctx->m_prepared.valid = false;
  if (ctx->math != MMS_MATH_TF32 || !ctx->scratch || ctx->scratch_bytes < sizeof(float) * (size_t)mc * D * Dp) return 0;
  RoundJob job{Mw, static_cast<float*>(ctx->scratch), (long long)mc * D, D, D, Dp, nullptr};
  MMS_TRY(mms_tf32_round(ctx, &job, 1));
  mms_context::MPrepared& mp = ctx->m_prepared;
  mp.M = Mw; mp.at = ctx->scratch; mp.D = D; mp.mc = mc; mp.clock = mms_write_clock(); mp.valid = true;
  return 0;
}

int mms_tc_simcross2_forward(mms_context* ctx, const float* q, const float* a, const float* Mw,
                             const float* B, float* S, int N, int Lq, int La, int D, int mc) {
  const Plan p = make_plan(ctx, N, Lq, La, D, mc, false);
  const int Dp = p.Dp;
  const mms_context::MPrepared mp = ctx->m_prepared;       // (mms_scratch clears the flag: any other user of the buffer does)
  void* sp = nullptr;
  // (+ room for the blocked U export of a later backward: up to 31 padding rows per measure)
  MMS_TRY(mms_scratch(ctx, sizeof(float) * (p.fixed + p.per_pair * p.nc_max + (size_t)mc * 32 * Dp), &sp));
  float* Mr = static_cast<float*>(sp);
  float* qr_ws = Mr + p.fixed;
  float* ar_ws = qr_ws + (size_t)p.nc_max * Lq * Dp;
  float* Tk = ar_ws + (size_t)p.nc_max * La * Dp;
  // operands a producer already staged (MMS_OPT_STAGE_TF32 on the Embed handle): read in place, no rounding pass
  const float* qs = p.nc_max == N ? mms_stage_lookup(q, (long long)N * Lq, D, Dp) : nullptr;
  const float* as = p.nc_max == N ? mms_stage_lookup(a, (long long)N * La, D, Dp) : nullptr;
  MMS_TRY(mms_stage_require_real(q, qs != nullptr, "bottom q"));
  MMS_TRY(mms_stage_require_real(a, as != nullptr, "bottom a"));
  const float* qr = qs ? qs : qr_ws;
  const float* ar = as ? as : ar_ws;
  // the rounded copies stay in the scratch buffer: a backward on the same handle may reuse them (MMS_OPT_REUSE_FORWARD)
  struct CacheMark {
    mms_context* c; bool ok;
    ~CacheMark() { c->fwd_cache.valid = ok; }
  } mark{ctx, false};
  for (int n0 = 0; n0 < N; n0 += p.nc_max) {
    const int nc = mms_min(p.nc_max, N - n0);
    float* Sc = S + (size_t)n0 * mc * Lq * La;
    RoundJob jobs[3];
    int nj = 0;
    if (!qs) jobs[nj++] = RoundJob{q + (size_t)n0 * Lq * D, qr_ws, (long long)nc * Lq, D, D, Dp, nullptr};
    if (!as) jobs[nj++] = RoundJob{a + (size_t)n0 * La * D, ar_ws, (long long)nc * La, D, D, Dp, nullptr};
    const bool m_ready = mp.valid && mp.M == Mw && mp.at == sp && mp.D == D && mp.mc == mc &&
                         mms_unchanged_since(mp.clock, Mw, sizeof(float) * (size_t)mc * D * D);
    if (n0 == 0 && !m_ready) jobs[nj++] = RoundJob{Mw, Mr, (long long)mc * D, D, D, Dp, nullptr};
    if (nj) MMS_TRY(mms_tf32_round(ctx, jobs, nj));
    {  // one kernel for both contractions, T stays in tensor memory (tc/simcross_fused.cu)
      const int rc = mms_tc_simcross2_forward_fused(ctx, qr, ar, Mr, B, Sc, nc, Lq, La, D, mc, Dp);
      if (rc == 0) {
        if (nc == N) {
          mms_context::FwdCache& fc = ctx->fwd_cache;
          fc.q = q; fc.a = a; fc.M = Mw; fc.N = N; fc.Lq = Lq; fc.La = La; fc.D = D; fc.mc = mc;
          fc.generation = mms_write_clock();
          fc.qr = qr; fc.ar = ar;
          mark.ok = true;
        }
        continue;
      }
      if (rc != MMS_E_UNSUPPORTED) return rc;
    }
    MMS_TRY(gemm_T(ctx, qr, Mr, Tk, nc * Lq, D, Dp, mc));                  // sim_cross_layer.cpp:148-149
    TcGemmArgs g = tc_gemm_args(Tk, Dp, 0, ar, Dp, 0, Sc, La, Lq, La, D);  // :151-153
    g.nb1 = mc; g.nb2 = nc;
    g.sA1 = (long long)nc * Lq * Dp; g.sA2 = (long long)Lq * Dp;
    g.sB2 = (long long)La * Dp;
    g.sC1 = (long long)Lq * La; g.sC2 = (long long)mc * Lq * La;
    if (B) { g.c_add = B; g.ld_add = La; g.s_add1 = (long long)Lq * La; }  // :155-159
    g.operands_tf32 = 1;
    MMS_TRY(mms_tc_gemm(ctx, g));
  }
  return 0;
}

namespace {

// dM_k += Qall^T U[k]   (= sum_n Q_n^T G_nk A_n, sim_cross_layer.cpp:286-289): reduction over all token rows,
// split over CTAs so that one full wave runs (a 149th tile would double the kernel's time)
int gemm_dM(mms_context* ctx, const float* qr, const float* U, float* dM, int rows, int D, int Dp, int mc) {
  TcGemmArgs g = tc_gemm_args(qr, Dp, 1, U, Dp, 1, dM, D, D, D, rows, TC_ATOMIC);
  g.nb1 = mc; g.sB1 = (long long)rows * Dp; g.sC1 = (long long)D * D;
  const int tiles = mc * mms_ceil_div(D, 128) * mms_ceil_div(D, 256);
  g.ksplit = mms_max(1, mms_min(ctx->sm_count / mms_max(tiles, 1), mms_ceil_div(rows, 128)));
  g.operands_tf32 = 1;
  return mms_tc_gemm(ctx, g);
}

// phase 0: the whole backward; 1: the bottom gradients only (dq, da; U = dS A stays in the workspace and
// ctx->dm_pending remembers where); 2: the weight gradient from that workspace
int backward_fused(mms_context* ctx, const float* q, const float* a, const float* Mw, const float* dS, float* dq,
                   float* da, float* dM, int N, int Lq, int La, int D, int mc, int phase = 0) {
  const int Dp = (D + 31) & ~31;
  // scratch per pair: rounded q and a rows + the exported U (mc x Lq x Dp)
  const size_t per_pair = (size_t)(Lq + La) * Dp + (size_t)mc * Lq * Dp;
  const size_t fixed = (size_t)mc * D * Dp;
  const long long cap = ((long long)(ctx->scratch_cap / sizeof(float)) - (long long)fixed) / (long long)per_pair;
  const int nc_max = (int)mms_max<long long>(1, mms_min<long long>(cap, N));
  // MMS_OPT_REUSE_FORWARD: the last forward on this handle left the rounded q, a and M at the head of the scratch
  // buffer (same layout: Mr | qr | ar | per-measure intermediate) and nothing has touched the buffer since
  const mms_context::FwdCache& fc = ctx->fwd_cache;
  const size_t need = sizeof(float) * (fixed + per_pair * nc_max + (size_t)mc * 32 * Dp);   // blocked U: padded to 32-row groups
  const bool unchanged = fc.valid && mms_unchanged_since(fc.generation, q, sizeof(float) * (size_t)N * Lq * D) &&
                         mms_unchanged_since(fc.generation, a, sizeof(float) * (size_t)N * La * D) &&
                         mms_unchanged_since(fc.generation, Mw, sizeof(float) * (size_t)mc * D * D);
  const bool reuse = ctx->reuse_forward && unchanged && fc.q == q && fc.a == a && fc.M == Mw && fc.N == N &&
                     fc.Lq == Lq && fc.La == La && fc.D == D && fc.mc == mc && nc_max == N && need <= ctx->scratch_bytes;
  void* sp = ctx->scratch;
  if (phase != 0 && nc_max != N) return MMS_E_UNSUPPORTED;      // the split form keeps U of the whole batch
  if (phase == 2) {
    const mms_context::DmPending& dp = ctx->dm_pending;
    MMS_REQUIRE(dp.valid && dp.N == N && dp.Lq == Lq && dp.La == La && dp.D == D && dp.mc == mc, MMS_E_INVALID,
                "mms_simcross_backward_params must directly follow the matching mms_simcross_backward_bottoms on this handle");
  } else if (!reuse) {
    MMS_TRY(mms_scratch(ctx, need, &sp));
  }
  float* Mr = static_cast<float*>(sp);
  float* qr_ws = Mr + fixed;
  float* ar_ws = qr_ws + (size_t)nc_max * Lq * Dp;
  float* U = ar_ws + (size_t)nc_max * La * Dp;
  // where the rounded operands are: what the forward used (reuse), a producer's staged copy, or the workspace
  const float* qs = nullptr; const float* as = nullptr;
  if (phase == 2) { qs = ctx->dm_pending.qr; }
  else if (reuse) { qs = fc.qr; as = fc.ar; }
  else if (nc_max == N) { qs = mms_stage_lookup(q, (long long)N * Lq, D, Dp); as = mms_stage_lookup(a, (long long)N * La, D, Dp); }
  if (phase != 2) {
    MMS_TRY(mms_stage_require_real(q, qs != nullptr, "bottom q"));
    MMS_TRY(mms_stage_require_real(a, as != nullptr, "bottom a"));
  }
  const float* qr = qs ? qs : qr_ws;
  const float* ar = as ? as : ar_ws;
  if (phase == 2) {
    MMS_CUDA(cudaMemsetAsync(dM, 0, sizeof(float) * (size_t)mc * D * D, ctx->stream));   // :256
    const int use_dm = mms_tc_simcross2_dm_plan(D) == 0 && (long long)N * Lq >= 16384;
    if (use_dm) MMS_TRY(mms_tc_simcross2_dm(ctx, qr, U, dM, (long long)N * Lq, D, Dp, mc, 1));        // :286-289
    else MMS_TRY(gemm_dM(ctx, qr, U, dM, N * Lq, D, Dp, mc));
    ctx->dm_pending.valid = false;
    return 0;
  }
  if (phase == 0) MMS_CUDA(cudaMemsetAsync(dM, 0, sizeof(float) * (size_t)mc * D * D, ctx->stream));   // :256
  // dq (+ U, dM) and da are independent: with MMS_OPT_CONCURRENCY the da kernel runs on a private stream and each
  // kernel is sized for half of the SMs when the batch is too small to fill them
  const bool want_conc = ctx->concurrency != 0;
  for (int n0 = 0; n0 < N; n0 += nc_max) {
    const int nc = mms_min(nc_max, N - n0);
    float* dqc = dq + (size_t)n0 * Lq * D;
    float* dac = da + (size_t)n0 * La * D;
    const float* dSc = dS + (size_t)n0 * mc * Lq * La;
    if (!reuse) {
      RoundJob jobs[3];
      int nj = 0;
      if (!qs) jobs[nj++] = RoundJob{q + (size_t)n0 * Lq * D, qr_ws, (long long)nc * Lq, D, D, Dp, nullptr};
      if (!as) jobs[nj++] = RoundJob{a + (size_t)n0 * La * D, ar_ws, (long long)nc * La, D, D, Dp, nullptr};
      if (n0 == 0) jobs[nj++] = RoundJob{Mw, Mr, (long long)mc * D, D, D, Dp, nullptr};
      if (nj) MMS_TRY(mms_tf32_round(ctx, jobs, nj));
    }
    // two kernels side by side only pay when one of them cannot fill the GPU: then each gets half of the SMs
    // (a handle with MMS_OPT_EMBED_DETERMINISTIC never spreads a pair group's measures over CTAs: one writer per row)
    const int allow_split = ctx->embed_deterministic ? 0 : 1;
    int ksplit = 1, split_from = nc;
    MMS_TRY(mms_tc_simcross2_backward_fused_plan(0, nc, Lq, La, D, mc, ctx->sm_count, allow_split, &ksplit, &split_from));
    const bool conc = want_conc && ksplit > 1 && split_from == 0;
    if (conc) MMS_TRY(mms_tc_simcross2_backward_fused_plan(0, nc, Lq, La, D, mc, mms_max(1, ctx->sm_count / 2), allow_split, &ksplit, &split_from));
    if (ksplit > 1 && split_from < nc) {   // measures of these pair groups are spread over CTAs that add into the output
      MMS_CUDA(cudaMemsetAsync(dqc + (size_t)split_from * Lq * D, 0, sizeof(float) * (size_t)(nc - split_from) * Lq * D, ctx->stream));
      MMS_CUDA(cudaMemsetAsync(dac + (size_t)split_from * La * D, 0, sizeof(float) * (size_t)(nc - split_from) * La * D, ctx->stream));
    }
    if (conc) {
      MMS_TRY(mms_fork(ctx, 0));
      MmsStreamSwitch sw(ctx, 0);
      MMS_TRY(mms_tc_simcross2_backward_fused(ctx, 1, qr, Mr, dSc, dac, nullptr, nc, Lq, La, D, mc, Dp, ksplit, split_from));  // :296-299
    }
    // dedicated dM kernel (tc/simcross_dm.cu) over the blocked U export; small batches are latency-bound and do
    // better with the 32-byte row-major export and the many small tiles of the generic engine (measured at 50 pairs)
    const int use_dm = mms_tc_simcross2_dm_plan(D) == 0 && (long long)nc * Lq >= 16384;
    const int blocked = use_dm;                               // U in the blocked layout that kernel reads best
    MMS_TRY(mms_tc_simcross2_backward_fused(ctx, 0, ar, Mr, dSc, dqc, U, nc, Lq, La, D, mc, Dp, ksplit, split_from, blocked));   // :291-294
    if (phase == 0) {
      if (use_dm) MMS_TRY(mms_tc_simcross2_dm(ctx, qr, U, dM, (long long)nc * Lq, D, Dp, mc, blocked));        // :286-289
      else MMS_TRY(gemm_dM(ctx, qr, U, dM, nc * Lq, D, Dp, mc));
    }
    if (conc) MMS_TRY(mms_join(ctx, 0));
    else MMS_TRY(mms_tc_simcross2_backward_fused(ctx, 1, qr, Mr, dSc, dac, nullptr, nc, Lq, La, D, mc, Dp, ksplit, split_from));
  }
  if (phase == 1) {
    mms_context::DmPending& dp = ctx->dm_pending;
    dp.valid = true; dp.N = N; dp.Lq = Lq; dp.La = La; dp.D = D; dp.mc = mc; dp.qr = qr;
  }
  return 0;
}

}  // namespace

int mms_tc_simcross2_backward_bottoms(mms_context* ctx, const float* q, const float* a, const float* Mw, const float* dS,
                                      float* dq, float* da, int N, int Lq, int La, int D, int mc) {
  int ksplit = 1;
  if (mms_tc_simcross2_backward_fused_plan(0, N, Lq, La, D, mc, ctx->sm_count, 0, &ksplit, nullptr) != 0) return MMS_E_UNSUPPORTED;
  return backward_fused(ctx, q, a, Mw, dS, dq, da, nullptr, N, Lq, La, D, mc, 1);
}

int mms_tc_simcross2_backward_params(mms_context* ctx, float* dM, int N, int Lq, int La, int D, int mc) {
  return backward_fused(ctx, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, dM, N, Lq, La, D, mc, 2);
}

// Computes dq, da (overwritten) and dM (overwritten: zeroed here, :256).  dB is left to the caller.
int mms_tc_simcross2_backward(mms_context* ctx, const float* q, const float* a, const float* Mw,
                              const float* dS, float* dq, float* da, float* dM, int N, int Lq, int La,
                              int D, int mc) {
  {  // fused kernels for dq / da (tc/simcross_fused_bwd.cu) + one GEMM for dM, when the shape allows
    int ksplit = 1;
    if (mms_tc_simcross2_backward_fused_plan(0, N, Lq, La, D, mc, ctx->sm_count, 0, &ksplit, nullptr) == 0) {
      const int rc = backward_fused(ctx, q, a, Mw, dS, dq, da, dM, N, Lq, La, D, mc);
      if (rc != MMS_E_UNSUPPORTED) return rc;
    }
  }
  const Plan p = make_plan(ctx, N, Lq, La, D, mc, true);
  const int Dp = p.Dp, Lap = p.Lap, Lmax = mms_max(Lq, La);
  void* sp = nullptr;
  MMS_TRY(mms_scratch(ctx, sizeof(float) * (p.fixed + p.per_pair * p.nc_max), &sp));
  float* Mr = static_cast<float*>(sp);
  float* qr = Mr + p.fixed;
  float* ar = qr + (size_t)p.nc_max * Lq * Dp;
  float* Gr = ar + (size_t)p.nc_max * La * Dp;
  float* buf = Gr + (size_t)p.nc_max * mc * Lq * Lap;
  (void)Lmax;
  MMS_CUDA(cudaMemsetAsync(dM, 0, sizeof(float) * (size_t)mc * D * D, ctx->stream));
  for (int n0 = 0; n0 < N; n0 += p.nc_max) {
    const int nc = mms_min(p.nc_max, N - n0);
    float* dqc = dq + (size_t)n0 * Lq * D;
    float* dac = da + (size_t)n0 * La * D;
    const RoundJob jobs[4] = {
        {q + (size_t)n0 * Lq * D, qr, (long long)nc * Lq, D, D, Dp, nullptr},
        {a + (size_t)n0 * La * D, ar, (long long)nc * La, D, D, Dp, nullptr},
        {dS + (size_t)n0 * mc * Lq * La, Gr, (long long)nc * mc * Lq, La, La, Lap, nullptr},
        {Mw, Mr, (long long)mc * D, D, D, Dp, nullptr}};
    MMS_TRY(mms_tf32_round(ctx, jobs, n0 == 0 ? 4 : 3));
    const long long sU1 = (long long)nc * Lq * Dp, sU2 = (long long)Lq * Dp;
    const long long sG1 = (long long)Lq * Lap, sG2 = (long long)mc * Lq * Lap;
    {  // U[k][n] = G_nk A_n
      TcGemmArgs g = tc_gemm_args(Gr, Lap, 0, ar, Dp, 1, buf, Dp, Lq, D, La);
      g.nb1 = mc; g.nb2 = nc;
      g.sA1 = sG1; g.sA2 = sG2;
      g.sB2 = (long long)La * Dp;
      g.sC1 = sU1; g.sC2 = sU2;
      g.operands_tf32 = 1; g.round_out = 1;
      MMS_TRY(mms_tc_gemm(ctx, g));
    }
    {  // dM_k += Qall^T U[k]            (= sum_n Q_n^T G_nk A_n, :286-289)
      const int kdim = nc * Lq;
      TcGemmArgs g = tc_gemm_args(qr, Dp, 1, buf, Dp, 1, dM, D, D, D, kdim, TC_ATOMIC);
      g.nb1 = mc; g.sB1 = sU1; g.sC1 = (long long)D * D;
      const int tiles = mc * mms_ceil_div(D, 128) * mms_ceil_div(D, 256);
      // one full wave of CTAs: tiles * ksplit <= SM count (a 149th tile would double the kernel's time)
      g.ksplit = mms_max(1, mms_min(ctx->sm_count / mms_max(tiles, 1), mms_ceil_div(kdim, 128)));
      g.operands_tf32 = 1;
      MMS_TRY(mms_tc_gemm(ctx, g));
    }
    {  // dQall = sum_k U[k] M_k^T       (= G (M_k A^T)^T summed over k, :291-294)
      TcGemmArgs g = tc_gemm_args(buf, Dp, 0, Mr, Dp, 0, dqc, D, nc * Lq, D, D);
      g.nseg = mc; g.segA = sU1; g.segB = (long long)D * Dp;
      g.operands_tf32 = 1;
      // a small batch has too few output tiles to fill the GPU: split the reduction over the measures
      // (atomic accumulation into a zeroed dq) until there is about one tile per SM
      const int tiles = mms_ceil_div(nc * Lq, 128) * mms_ceil_div(D, 256);
      const int split = mms_min(mc, ctx->sm_count / mms_max(tiles, 1));
      if (split > 1) {
        MMS_CUDA(cudaMemsetAsync(dqc, 0, sizeof(float) * (size_t)nc * Lq * D, ctx->stream));
        g.ksplit = split; g.mode = TC_ATOMIC;
      }
      MMS_TRY(mms_tc_gemm(ctx, g));
    }
    MMS_TRY(gemm_T(ctx, qr, Mr, buf, nc * Lq, D, Dp, mc));
    {  // dA_n = sum_k G_nk^T T[k][n]    (:296-299)
      TcGemmArgs g = tc_gemm_args(Gr, Lap, 1, buf, Dp, 1, dac, D, La, D, Lq);
      g.nb2 = nc;
      g.sA2 = sG2; g.sB2 = sU2; g.sC2 = (long long)La * D;
      g.nseg = mc; g.segA = sG1; g.segB = sU1;
      g.operands_tf32 = 1;
      MMS_TRY(mms_tc_gemm(ctx, g));
    }
  }
  return 0;
}

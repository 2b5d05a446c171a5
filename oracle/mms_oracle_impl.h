/* TEST INFRASTRUCTURE ONLY -- body of the plain-C oracle, included twice by
 * mms_oracle.c with REAL = float (suffix _f32) and REAL = double (suffix _f64).
 *
 * Every function restates one reference routine (cited file:line under
 * /root/reference).  Loops that the reference writes out by hand (FM, PairRankLoss,
 * SimCross modes 0/1, Embed) keep the reference's statement order and accumulate in
 * REAL exactly as the reference accumulates in Dtype, so they are bit-comparable
 * with oracle/_ref.  Contractions that the reference delegates to BLAS
 * (cblas_?gemm/ger/gemv/dot) are restated as plain dot-product loops accumulating
 * in REAL: same mathematics, BLAS-unspecified summation order, hence compared
 * within tolerance (tests/test_oracle.py).
 */

/* ---------------------------------------------------------------- Embed ---- */
/* embed_layer.cpp:135-153  top[n,:] = W[int(idx[n]),:] (+ bias via rank-1 gemm) */
int FN(mmso_embed_forward)(const REAL* idx, const REAL* W, const REAL* bias,
                           REAL* top, int M, int D, int V) {
  for (int n = 0; n < M; ++n) {
    const int index = (int)idx[n];            /* static_cast<int>, :142 */
    if (index < 0 || index >= V) return 1;    /* DCHECK_GE / DCHECK_LT, :143-144 */
    memcpy(top + (size_t)n * D, W + (size_t)index * D, sizeof(REAL) * D);  /* :146 */
  }
  if (bias) {                                 /* gemm(M,D,1, 1, ones, bias, 1, top) :150 */
    for (int n = 0; n < M; ++n)
      for (int d = 0; d < D; ++d) top[(size_t)n * D + d] += (REAL)1 * bias[d];
  }
  return 0;
}

/* embed_layer.cpp:156-180  dW[idx[n],:] += dtop[n,:] (axpy, in n order);
 * db += dtop^T * ones (gemv, beta = 1).  Both ACCUMULATE. */
int FN(mmso_embed_backward)(const REAL* idx, const REAL* dtop, REAL* dW, REAL* dbias,
                            int M, int D, int V) {
  if (dW) {
    for (int n = 0; n < M; ++n) {
      const int index = (int)idx[n];
      if (index < 0 || index >= V) return 1;
      for (int d = 0; d < D; ++d) dW[(size_t)index * D + d] += dtop[(size_t)n * D + d];
    }
  }
  if (dbias) {
    for (int d = 0; d < D; ++d) {
      REAL acc = 0;
      for (int n = 0; n < M; ++n) acc += dtop[(size_t)n * D + d];
      dbias[d] += acc;
    }
  }
  return 0;
}

/* ------------------------------------------------------------- SimCross ---- */
/* sim_cross_layer.cpp:83-161.  q (N,Lq,D), a (N,La,D), Mw (mc,D,D), B (mc,Lq,La).
 * S is (N,mc,Lq,La) for mode 2 and (N,1,Lq,La) for modes 0/1.
 * norm0 (N,Lq) / norm1 (N,La) are the mode-0 caches data0_norm_/data1_norm_. */
int FN(mmso_simcross_forward)(int mode, const REAL* q, const REAL* a, const REAL* Mw,
                              const REAL* B, REAL* S, REAL* norm0, REAL* norm1,
                              int N, int Lq, int La, int D, int mc) {
  if (mode == 1) {                                           /* :96-111 */
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < Lq; ++j)
        for (int k = 0; k < La; ++k) {
          REAL dist = 0;
          for (int dd = 0; dd < D; ++dd) {
            REAL diff = q[((size_t)i * Lq + j) * D + dd] - a[((size_t)i * La + k) * D + dd];
            dist += diff * diff;
          }
          dist = (REAL)sqrt(dist);
          S[((size_t)i * Lq + j) * La + k] = 1 / (1 + dist);
        }
  } else if (mode == 0) {                                    /* :112-139 */
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < Lq; ++j) {
        REAL acc = 0;
        const REAL* x = q + ((size_t)i * Lq + j) * D;
        for (int dd = 0; dd < D; ++dd) acc += x[dd] * x[dd];
        norm0[(size_t)i * Lq + j] = (REAL)sqrt(acc);
      }
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < La; ++j) {
        REAL acc = 0;
        const REAL* x = a + ((size_t)i * La + j) * D;
        for (int dd = 0; dd < D; ++dd) acc += x[dd] * x[dd];
        norm1[(size_t)i * La + j] = (REAL)sqrt(acc);
      }
    for (int i = 0; i < N; ++i)
      for (int j = 0; j < Lq; ++j)
        for (int k = 0; k < La; ++k) {
          REAL acc = 0;
          const REAL* x = q + ((size_t)i * Lq + j) * D;
          const REAL* y = a + ((size_t)i * La + k) * D;
          for (int dd = 0; dd < D; ++dd) acc += x[dd] * y[dd];
          S[((size_t)i * Lq + j) * La + k] =
              acc / norm0[(size_t)i * Lq + j] / norm1[(size_t)i * La + k];
        }
  } else if (mode == 2) {                                    /* :140-161 */
    REAL* T = (REAL*)malloc(sizeof(REAL) * (size_t)Lq * D);  /* measure_temp0_ */
    if (!T) return 2;
    for (int i = 0; i < N; ++i) {
      for (int j = 0; j < mc; ++j) {
        const REAL* Q = q + (size_t)i * Lq * D;
        const REAL* A = a + (size_t)i * La * D;
        const REAL* Mk = Mw + (size_t)j * D * D;
        /* T = Q * M_k          (gemm NoTrans,NoTrans  Lq x D x D)  :148-149 */
        for (int r = 0; r < Lq; ++r)
          for (int c = 0; c < D; ++c) {
            REAL acc = 0;
            for (int t = 0; t < D; ++t) acc += Q[(size_t)r * D + t] * Mk[(size_t)t * D + c];
            T[(size_t)r * D + c] = acc;
          }
        /* S[i,j] = T * A^T     (gemm NoTrans,Trans    Lq x La x D) :151-153 */
        REAL* Sij = S + (((size_t)i * mc + j) * Lq) * La;
        for (int r = 0; r < Lq; ++r)
          for (int c = 0; c < La; ++c) {
            REAL acc = 0;
            for (int t = 0; t < D; ++t) acc += T[(size_t)r * D + t] * A[(size_t)c * D + t];
            Sij[(size_t)r * La + c] = acc;
          }
      }
      if (B) {                                               /* caffe_add :155-159 */
        REAL* Si = S + (size_t)i * mc * Lq * La;
        for (size_t e = 0; e < (size_t)mc * Lq * La; ++e) Si[e] = B[e] + Si[e];
      }
    }
    free(T);
  } else {
    return 3;
  }
  return 0;
}

/* sim_cross_layer.cpp:166-307.  Always zeroes dq and da (:176-177); the rest only
 * runs if prop0 || prop1 (:201).  Mode 2 zeroes dM (:256) but ACCUMULATES into dB
 * (:301-304).  S is the forward output (modes 0/1 read it back). */
int FN(mmso_simcross_backward)(int mode, const REAL* q, const REAL* a, const REAL* Mw,
                               const REAL* S, const REAL* dS, const REAL* norm0,
                               const REAL* norm1, REAL* dq, REAL* da, REAL* dM, REAL* dB,
                               int N, int Lq, int La, int D, int mc, int prop0, int prop1) {
  memset(dq, 0, sizeof(REAL) * (size_t)N * Lq * D);
  memset(da, 0, sizeof(REAL) * (size_t)N * La * D);
  if (!(prop0 || prop1)) return 0;
  if (mode == 1) {                                           /* :208-226 */
    for (int dd = 0; dd < D; ++dd)
      for (int j = 0; j < N; ++j)
        for (int k = 0; k < Lq; ++k)
          for (int m = 0; m < La; ++m) {
            const size_t t = ((size_t)j * Lq + k) * La + m;
            const size_t b0 = ((size_t)j * Lq + k) * D + dd;
            const size_t b1 = ((size_t)j * La + m) * D + dd;
            /* the 1e-9 literal is a double: the division is carried out in double */
            REAL tt = (REAL)(dS[t] * S[t] * S[t] * S[t] * (q[b0] - a[b1]) / (S[t] - 1 + 1e-9));
            dq[b0] += tt;
            da[b1] += -tt;
          }
  } else if (mode == 0) {                                    /* :227-250 */
    for (int dd = 0; dd < D; ++dd)
      for (int j = 0; j < N; ++j)
        for (int k = 0; k < Lq; ++k)
          for (int m = 0; m < La; ++m) {
            const size_t t = ((size_t)j * Lq + k) * La + m;
            const size_t b0 = ((size_t)j * Lq + k) * D + dd;
            const size_t b1 = ((size_t)j * La + m) * D + dd;
            const REAL n0 = norm0[(size_t)j * Lq + k];
            const REAL n1 = norm1[(size_t)j * La + m];
            REAL tt = dS[t] * (a[b1] / n0 / n1 - q[b0] * S[t] / (n0 * n0));
            dq[b0] += tt;
            tt = dS[t] * (q[b0] / n0 / n1 - a[b1] * S[t] / (n1 * n1));
            da[b1] += tt;
          }
  } else if (mode == 2) {                                    /* :251-307 */
    REAL* t0 = (REAL*)malloc(sizeof(REAL) * (size_t)Lq * D);  /* measure_temp0_ */
    REAL* t1 = (REAL*)malloc(sizeof(REAL) * (size_t)La * D);  /* measure_temp1_ (used as D x La) */
    if (!t0 || !t1) { free(t0); free(t1); return 2; }
    memset(dM, 0, sizeof(REAL) * (size_t)mc * D * D);        /* :256 */
    for (int i = 0; i < N; ++i) {
      const REAL* Q = q + (size_t)i * Lq * D;
      const REAL* A = a + (size_t)i * La * D;
      REAL* dQ = dq + (size_t)i * Lq * D;
      REAL* dA = da + (size_t)i * La * D;
      for (int j = 0; j < mc; ++j) {
        const REAL* G = dS + (((size_t)i * mc + j) * Lq) * La;
        const REAL* Mk = Mw + (size_t)j * D * D;
        REAL* dMk = dM + (size_t)j * D * D;
        /* temp1 (D x La) = Q^T * G                       :286-287 */
        for (int r = 0; r < D; ++r)
          for (int c = 0; c < La; ++c) {
            REAL acc = 0;
            for (int t = 0; t < Lq; ++t) acc += Q[(size_t)t * D + r] * G[(size_t)t * La + c];
            t1[(size_t)r * La + c] = acc;
          }
        /* dM_k += temp1 * A   (D x D, K = La)            :288-289 */
        for (int r = 0; r < D; ++r)
          for (int c = 0; c < D; ++c) {
            REAL acc = 0;
            for (int t = 0; t < La; ++t) acc += t1[(size_t)r * La + t] * A[(size_t)t * D + c];
            dMk[(size_t)r * D + c] += acc;
          }
        /* temp1 (D x La) = M_k * A^T                     :291-292 */
        for (int r = 0; r < D; ++r)
          for (int c = 0; c < La; ++c) {
            REAL acc = 0;
            for (int t = 0; t < D; ++t) acc += Mk[(size_t)r * D + t] * A[(size_t)c * D + t];
            t1[(size_t)r * La + c] = acc;
          }
        /* dQ += G * temp1^T   (Lq x D, K = La)           :293-294 */
        for (int r = 0; r < Lq; ++r)
          for (int c = 0; c < D; ++c) {
            REAL acc = 0;
            for (int t = 0; t < La; ++t) acc += G[(size_t)r * La + t] * t1[(size_t)c * La + t];
            dQ[(size_t)r * D + c] += acc;
          }
        /* temp0 (Lq x D) = Q * M_k                       :296-297 */
        for (int r = 0; r < Lq; ++r)
          for (int c = 0; c < D; ++c) {
            REAL acc = 0;
            for (int t = 0; t < D; ++t) acc += Q[(size_t)r * D + t] * Mk[(size_t)t * D + c];
            t0[(size_t)r * D + c] = acc;
          }
        /* dA += G^T * temp0   (La x D, K = Lq)           :298-299 */
        for (int r = 0; r < La; ++r)
          for (int c = 0; c < D; ++c) {
            REAL acc = 0;
            for (int t = 0; t < Lq; ++t) acc += G[(size_t)t * La + r] * t0[(size_t)t * D + c];
            dA[(size_t)r * D + c] += acc;
          }
      }
      if (dB) {                                              /* caffe_add :301-304 */
        const REAL* Gi = dS + (size_t)i * mc * Lq * La;
        for (size_t e = 0; e < (size_t)mc * Lq * La; ++e) dB[e] = Gi[e] + dB[e];
      }
    }
    free(t0);
    free(t1);
  } else {
    return 3;
  }
  return 0;
}

/* ------------------------------------------------------------ SimMatrix ---- */
/* sim_matrix_layer.cpp:53-65.  T (N,K2) is the scratch the reference keeps in
 * bottom[1]'s diff buffer (:58); s_n = <a_n, T_n>. */
int FN(mmso_simmatrix_forward)(const REAL* q, const REAL* a, const REAL* W, REAL* s,
                               REAL* T, int N, int K1, int K2) {
  for (int n = 0; n < N; ++n)
    for (int c = 0; c < K2; ++c) {
      REAL acc = 0;
      for (int t = 0; t < K1; ++t) acc += q[(size_t)n * K1 + t] * W[(size_t)t * K2 + c];
      T[(size_t)n * K2 + c] = acc;
    }
  for (int n = 0; n < N; ++n) {
    REAL acc = 0;
    for (int c = 0; c < K2; ++c) acc += a[(size_t)n * K2 + c] * T[(size_t)n * K2 + c];
    s[n] = acc;
  }
  return 0;
}

/* sim_matrix_layer.cpp:68-95.  dW += sum_n ds_n q_n a_n^T (ger, ACCUMULATES, in n
 * order); dq_n = ds_n W a_n, da_n = ds_n W^T q_n (gemv beta = 0: OVERWRITE). */
int FN(mmso_simmatrix_backward)(const REAL* q, const REAL* a, const REAL* W, const REAL* ds,
                                REAL* dW, REAL* dq, REAL* da, int N, int K1, int K2,
                                int prop_w, int prop0, int prop1) {
  if (prop_w) {
    for (int n = 0; n < N; ++n)
      for (int r = 0; r < K1; ++r) {
        const REAL f = ds[n] * q[(size_t)n * K1 + r];
        for (int c = 0; c < K2; ++c) dW[(size_t)r * K2 + c] += f * a[(size_t)n * K2 + c];
      }
  }
  if (prop0) {
    for (int n = 0; n < N; ++n)
      for (int r = 0; r < K1; ++r) {
        REAL acc = 0;
        for (int c = 0; c < K2; ++c) acc += W[(size_t)r * K2 + c] * a[(size_t)n * K2 + c];
        dq[(size_t)n * K1 + r] = ds[n] * acc;
      }
  }
  if (prop1) {
    for (int n = 0; n < N; ++n)
      for (int c = 0; c < K2; ++c) {
        REAL acc = 0;
        for (int r = 0; r < K1; ++r) acc += W[(size_t)r * K2 + c] * q[(size_t)n * K1 + r];
        da[(size_t)n * K2 + c] = ds[n] * acc;
      }
  }
  return 0;
}

/* --------------------------------------------------------- PairRankLoss ---- */
/* pair_rank_loss_layer.cpp:26-52 with the author's MKL axpby semantics (= the GPU
 * path pair_rank_loss_layer.cu:17-41): ordered = margin - y*(a-b), similar = a-b,
 * loss = (1/count) * sum[max(0,ordered) + |(1-y)*similar|], summed in index order. */
int FN(mmso_pairrankloss_forward)(const REAL* a, const REAL* b, const REAL* y, REAL margin,
                                  int count, REAL* loss_out, REAL* ordered, REAL* similar) {
  REAL loss = 0;
  for (int i = 0; i < count; ++i) {
    const REAL d = a[i] - b[i];
    similar[i] = d;
    REAL o = d * y[i];         /* caffe_mul :36 */
    o = (REAL)(-1) * o;        /* axpby(-1, X, 0, Y=X) element-wise (MKL) :37 */
    o += margin;               /* caffe_add_scalar :38 */
    ordered[i] = o;
  }
  for (int i = 0; i < count; ++i) {
    const REAL h = ordered[i] > 0 ? ordered[i] : (REAL)0;       /* std::max(0, .) :44 */
    const REAL s = (REAL)fabs((1 - y[i]) * similar[i]);         /* std::abs :45 */
    loss += h + s;
  }
  loss /= (REAL)count;                                          /* :50 */
  *loss_out = loss;
  return 0;
}

/* pair_rank_loss_layer.cpp:55-84 (ge = 0, `ordered > 0`) and
 * pair_rank_loss_layer.cu:46-55 (ge = 1, `ordered >= 0`).
 * sign_i * (top_diff/count) * ( [ordered>0]*y - sgn((1-y)*similar)*(1-y) ),
 * sgn(0) := -1, sign = -1 for bottom 0 and +1 for bottom 1; OVERWRITES. */
int FN(mmso_pairrankloss_backward)(const REAL* y, const REAL* ordered, const REAL* similar,
                                   REAL top_diff, int count, int ge, REAL* da, REAL* db) {
  for (int side = 0; side < 2; ++side) {
    REAL* out = side == 0 ? da : db;
    if (!out) continue;
    REAL sign = side == 0 ? (REAL)-1 : (REAL)1;
    sign *= top_diff / count;                                   /* :67 */
    for (int i = 0; i < count; ++i) {
      const REAL ot = (ge ? ordered[i] >= 0 : ordered[i] > 0) ? (REAL)1 : (REAL)0;
      const REAL st = (1 - y[i]) * similar[i] > 0 ? (REAL)1 : (REAL)-1;
      out[i] = sign * (ot * y[i] - st * (1 - y[i]));            /* :79 */
    }
  }
  return 0;
}

/* ------------------------------------------------------------------- FM ---- */
/* fm_layer.cpp:33-62.  x (N,C,Dm): column 0 linear, columns 1.. factors. */
int FN(mmso_fm_forward)(const REAL* x, const REAL* bias, REAL* y, int N, int C, int Dm) {
  for (int i = 0; i < N; ++i) {
    REAL t1 = 0;
    for (int j = 1; j < Dm; ++j) {
      REAL t2 = 0;
      for (int k = 0; k < C; ++k) {
        const size_t ind = ((size_t)i * C + k) * Dm + j;
        t2 += x[ind];
        t1 -= x[ind] * x[ind];
      }
      t1 += t2 * t2;
    }
    t1 /= 2;
    for (int k = 0; k < C; ++k) t1 += x[((size_t)i * C + k) * Dm];
    if (bias) t1 += bias[0];
    y[i] = t1;
  }
  return 0;
}

/* fm_layer.cpp:65-99.  db is OVERWRITTEN (:77); dx overwritten. */
int FN(mmso_fm_backward)(const REAL* x, const REAL* dy, REAL* dx, REAL* dbias,
                         int N, int C, int Dm, int prop0) {
  if (dbias) {
    dbias[0] = 0;
    for (int i = 0; i < N; ++i) dbias[0] += dy[i];
  }
  if (prop0) {
    for (int i = 0; i < N; ++i) {
      for (int k = 0; k < C; ++k) dx[((size_t)i * C + k) * Dm] = dy[i];
      for (int j = 1; j < Dm; ++j) {
        REAL tt = 0;
        for (int k = 0; k < C; ++k) tt += x[((size_t)i * C + k) * Dm + j];
        for (int k = 0; k < C; ++k) {
          const size_t ind = ((size_t)i * C + k) * Dm + j;
          dx[ind] = dy[i] * (tt - x[ind]);
        }
      }
    }
  }
  return 0;
}

/* One AdaDelta step of one learnable blob, in the order SGDSolver::ApplyUpdate makes its passes
 * (src/caffe/solvers/sgd_solver.cpp:102-116):
 *   Normalize      :118-141 (caffe_scal by 1/iter_size) -- here the general gradient scale, which also carries the
 *                  1/solver_count of P2PSync::on_gradients_ready (src/caffe/parallel.cpp:377)
 *   Regularize     :143-161, L2 only: diff += local_decay * data   (local_decay = weight_decay * decay_mult)
 *   ComputeUpdateValue  src/caffe/solvers/adadelta_solver.cpp:33-94 (CPU branch), pass by pass: powx 2, axpby into the
 *                  gradient history, (delta + update history) / (delta + gradient history), powx 0.5, mul, powx 2,
 *                  axpby into the update history, scale by local_rate (= base_lr * lr_mult)
 *   Net::Update    -> Blob::Update (src/caffe/blob.cpp): data -= diff
 * diff is left holding the applied update, as in the reference.  data == NULL skips Net::Update (the bare
 * adadelta_update of adadelta_solver.cu:6-26).  caffe_powx is pow() per element (mkl_alternate.hpp vsPowx). */
int FN(mmso_adadelta_step)(REAL* data, REAL* diff, REAL* hist_g, REAL* hist_u, long long n, REAL grad_scale,
                           REAL local_decay, REAL momentum, REAL delta, REAL local_rate) {
  for (long long i = 0; i < n; ++i) {
    REAL g = diff[i];
    if (grad_scale != (REAL)1) g *= grad_scale;
    if (local_decay != (REAL)0 && data) g += local_decay * data[i];
    REAL upd = MMSO_POW(g, (REAL)2);
    hist_g[i] = ((REAL)1 - momentum) * upd + momentum * hist_g[i];
    REAL tmp = delta;
    upd = tmp + hist_u[i];
    tmp = tmp + hist_g[i];
    upd = upd / tmp;
    upd = MMSO_POW(upd, (REAL)0.5);
    g = g * upd;
    upd = MMSO_POW(g, (REAL)2);
    hist_u[i] = ((REAL)1 - momentum) * upd + momentum * hist_u[i];
    g = local_rate * g;
    diff[i] = g;
    if (data) data[i] = data[i] - g;
  }
  return 0;
}

/* ---- ranking metrics (evaluation layers) -------------------------------------------------
 * Shared by MAP and MRR: samples are bucketed by int(group) into a std::map (ascending key), each bucket holds
 * (float score, int label) pairs in input order and is sorted by score, descending, with std::sort
 * (map_layer.cpp:47-51,72; mrr_layer.cpp:47-50,57).  std::sort is not stable: the order of EQUAL scores is
 * unspecified in the reference; this restatement (and the GPU implementation) keeps input order among equals. */
#ifndef MMSO_RANK_HELPERS
#define MMSO_RANK_HELPERS
typedef struct { int group; float score; int label; int index; } mmso_rank_item;
static int mmso_rank_cmp(const void* pa, const void* pb) {
  const mmso_rank_item* a = (const mmso_rank_item*)pa; const mmso_rank_item* b = (const mmso_rank_item*)pb;
  if (a->group != b->group) return a->group < b->group ? -1 : 1;
  if (a->score != b->score) return a->score > b->score ? -1 : 1;
  return a->index < b->index ? -1 : (a->index > b->index ? 1 : 0);
}
#endif

/* MAPLayer::Forward_cpu, map_layer.cpp:41-100 and MRRLayer::Forward_cpu, mrr_layer.cpp:38-79.  The score of sample i
 * is data[i * (fixed_axis + 1) + fixed_axis] (:50 / :49).  MAP counts label == 1 as positive and anything else as a
 * negative (:78-85); MRR needs a label == 0 for its negative (:64-66).  Groups without a positive or without a
 * negative are skipped; the mean is over the remaining groups (0/0 if there is none, as in the reference). */
int FN(mmso_map_mrr)(const REAL* data, const REAL* label, const REAL* group, int n, int fixed_axis, REAL* map_out,
                     REAL* mrr_out) {
  mmso_rank_item* it = (mmso_rank_item*)malloc(sizeof(mmso_rank_item) * (size_t)(n > 0 ? n : 1));
  if (!it) return 1;
  for (int i = 0; i < n; ++i) {
    it[i].group = (int)group[i];
    it[i].score = (float)data[(size_t)i * (fixed_axis + 1) + fixed_axis];
    it[i].label = (int)label[i];
    it[i].index = i;
  }
  qsort(it, (size_t)n, sizeof(mmso_rank_item), mmso_rank_cmp);
  REAL map_ = 0, mrr = 0;
  int eff_map = 0, eff_mrr = 0;
  for (int s = 0; s < n;) {
    int e = s;
    while (e < n && it[e].group == it[s].group) ++e;
    REAL ap = 0;
    int map_rank = 0, neg_any = 0, mrr_rank = -1, neg_zero = 0;
    for (int i = s; i < e; ++i) {
      if (it[i].label == 1) { ap += (REAL)(++map_rank) / (REAL)(i - s + 1); if (mrr_rank < 0) mrr_rank = i - s; }
      else neg_any = 1;
      if (it[i].label == 0) neg_zero = 1;
    }
    if (map_rank >= 1 && neg_any) { ++eff_map; map_ += ap / map_rank; }
    if (mrr_rank >= 0 && neg_zero) { ++eff_mrr; mrr = (REAL)((double)mrr + 1.0 / (mrr_rank + 1)); }   /* :75: Dtype += double */
    s = e;
  }
  if (map_out) *map_out = map_ / (REAL)eff_map;
  if (mrr_out) *mrr_out = mrr / (REAL)eff_mrr;
  free(it);
  return 0;
}

/* AUCLayer::Forward_cpu, auc_layer.cpp:47-136, for the 2-D (N, C) bottom the reference net feeds it
 * (label axis 1, inner_num 1): score of sample i = data[i * dim + fixed_axis]; sorted by score, descending (compared
 * as float, :43-45); auc = sum over samples of high * (1 - label), high = positives so far; / high / (count - high). */
int FN(mmso_auc)(const REAL* data, const REAL* label, int n, int dim, int fixed_axis, int has_ignore, int ignore_label,
                 REAL* out) {
  mmso_rank_item* it = (mmso_rank_item*)malloc(sizeof(mmso_rank_item) * (size_t)(n > 0 ? n : 1));
  if (!it) return 1;
  int count = 0;
  for (int i = 0; i < n; ++i) {
    const int lv = (int)label[i];
    if (has_ignore && lv == ignore_label) continue;
    it[count].group = 0;
    it[count].score = (float)data[(size_t)i * dim + fixed_axis];
    it[count].label = lv;
    it[count].index = count;
    ++count;
  }
  qsort(it, (size_t)count, sizeof(mmso_rank_item), mmso_rank_cmp);
  int high = 0;
  REAL auc = 0;
  for (int i = 0; i < count; ++i) { high += it[i].label; auc += (REAL)(high * (1 - it[i].label)); }
  *out = high > 0 ? auc / high / (count - high) : (REAL)0;
  free(it);
  return 0;
}

/* RankAccuracyLayer::Forward_cpu, rank_accuracy_layer.cpp:36-50. */
int FN(mmso_rank_accuracy)(const REAL* a, const REAL* b, const REAL* label, int n, REAL* out) {
  REAL acc = 0;
  for (int i = 0; i < n; ++i) acc += (label[i] * (a[i] - b[i])) > 0 ? 1 : 0;
  *out = acc / n;
  return 0;
}

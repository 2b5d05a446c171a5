/* TEST INFRASTRUCTURE ONLY -- prototypes of the plain-C oracle (mms_oracle.c).
 * Generated layout: every routine exists as <name>_f32 (REAL = float) and
 * <name>_f64 (REAL = double).  All return 0 on success. */
#ifndef MMS_ORACLE_H_
#define MMS_ORACLE_H_
#ifdef __cplusplus
extern "C" {
#endif

int mmso_embed_forward_f32(const float* idx, const float* W, const float* bias, float* top, int M, int D, int V);
int mmso_embed_backward_f32(const float* idx, const float* dtop, float* dW, float* dbias, int M, int D, int V);
int mmso_simcross_forward_f32(int mode, const float* q, const float* a, const float* Mw, const float* B, float* S, float* norm0, float* norm1, int N, int Lq, int La, int D, int mc);
int mmso_simcross_backward_f32(int mode, const float* q, const float* a, const float* Mw, const float* S, const float* dS, const float* norm0, const float* norm1, float* dq, float* da, float* dM, float* dB, int N, int Lq, int La, int D, int mc, int prop0, int prop1);
int mmso_simmatrix_forward_f32(const float* q, const float* a, const float* W, float* s, float* T, int N, int K1, int K2);
int mmso_simmatrix_backward_f32(const float* q, const float* a, const float* W, const float* ds, float* dW, float* dq, float* da, int N, int K1, int K2, int prop_w, int prop0, int prop1);
int mmso_pairrankloss_forward_f32(const float* a, const float* b, const float* y, float margin, int count, float* loss_out, float* ordered, float* similar);
int mmso_pairrankloss_backward_f32(const float* y, const float* ordered, const float* similar, float top_diff, int count, int ge, float* da, float* db);
int mmso_fm_forward_f32(const float* x, const float* bias, float* y, int N, int C, int Dm);
int mmso_fm_backward_f32(const float* x, const float* dy, float* dx, float* dbias, int N, int C, int Dm, int prop0);
int mmso_map_mrr_f32(const float* data, const float* label, const float* group, int n, int fixed_axis, float* map_out, float* mrr_out);
int mmso_auc_f32(const float* data, const float* label, int n, int dim, int fixed_axis, int has_ignore, int ignore_label, float* out);
int mmso_rank_accuracy_f32(const float* a, const float* b, const float* label, int n, float* out);
int mmso_adadelta_step_f32(float* data, float* diff, float* hist_g, float* hist_u, long long n, float grad_scale, float local_decay, float momentum, float delta, float local_rate);

int mmso_embed_forward_f64(const double* idx, const double* W, const double* bias, double* top, int M, int D, int V);
int mmso_embed_backward_f64(const double* idx, const double* dtop, double* dW, double* dbias, int M, int D, int V);
int mmso_simcross_forward_f64(int mode, const double* q, const double* a, const double* Mw, const double* B, double* S, double* norm0, double* norm1, int N, int Lq, int La, int D, int mc);
int mmso_simcross_backward_f64(int mode, const double* q, const double* a, const double* Mw, const double* S, const double* dS, const double* norm0, const double* norm1, double* dq, double* da, double* dM, double* dB, int N, int Lq, int La, int D, int mc, int prop0, int prop1);
int mmso_simmatrix_forward_f64(const double* q, const double* a, const double* W, double* s, double* T, int N, int K1, int K2);
int mmso_simmatrix_backward_f64(const double* q, const double* a, const double* W, const double* ds, double* dW, double* dq, double* da, int N, int K1, int K2, int prop_w, int prop0, int prop1);
int mmso_pairrankloss_forward_f64(const double* a, const double* b, const double* y, double margin, int count, double* loss_out, double* ordered, double* similar);
int mmso_pairrankloss_backward_f64(const double* y, const double* ordered, const double* similar, double top_diff, int count, int ge, double* da, double* db);
int mmso_fm_forward_f64(const double* x, const double* bias, double* y, int N, int C, int Dm);
int mmso_fm_backward_f64(const double* x, const double* dy, double* dx, double* dbias, int N, int C, int Dm, int prop0);
int mmso_map_mrr_f64(const double* data, const double* label, const double* group, int n, int fixed_axis, double* map_out, double* mrr_out);
int mmso_auc_f64(const double* data, const double* label, int n, int dim, int fixed_axis, int has_ignore, int ignore_label, double* out);
int mmso_rank_accuracy_f64(const double* a, const double* b, const double* label, int n, double* out);
int mmso_adadelta_step_f64(double* data, double* diff, double* hist_g, double* hist_u, long long n, double grad_scale, double local_decay, double momentum, double delta, double local_rate);

#ifdef __cplusplus
}
#endif
#endif  /* MMS_ORACLE_H_ */

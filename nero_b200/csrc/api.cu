// extern "C" entry points of libnero_b200 (declared in include/nero_b200.h).
#include "common.cuh"
#include "../../include/nero_b200.h"

namespace nero {
struct LinearParams {
  const float* A; int lda; int k_valid;
  const uint8_t* wimg; int n_pad; int k_chunks;
  const float* bias;
  float* out; int ldo; int ncol_out; float oscale;
  int mode; int act; float act_param;
  const float* H; int ldh; float hscale; int dact;
  const float* V; int ldv;
  float* out2; int ldo2;
  const float* addend; int ldadd;
  int ncol_main; float* tail; int ldt;
  const int* m_ptr; int m_cap;
};
struct WgradParams {
  const float* dY; int ldy; const float* X; int ldx;
  const float* dY2; int ldy2; const float* X2; int ldx2;
  int n0; int n_valid; int k0; int k_valid;
  float* partial; int ld_partial; int rows_partial;
  float* bias_partial;
  const int* m_ptr; int m_cap;
};
int linear_dispatch(const LinearParams& p, cudaStream_t stream);
int wgrad_dispatch(WgradParams p, int n_rows_pad, int k_pad, int P, cudaStream_t stream);
int prep_weight(const float* v, const float* g, int K, int row0, int nrows, const int* kmap, float in_scale,
                uint8_t* img_f, int rows_pad_f, uint8_t* img_t, int rows_pad_t, int t_c0, int t_ncols, float* w_eff, int ld_weff,
                cudaStream_t stream);
int wgrad_finish(const float* partial, int P, int rows_partial, int ld_partial, const float* bias_partial, int K,
                 int row0, int nrows, const int* kmap, float in_scale, const float* v, const float* g, float* grad_w,
                 float* grad_g, float* grad_b, const float* extra_row, float extra_scale, cudaStream_t stream);
int colsum(const float* X, int ldx, int ncol, const float* w, int ldw, const int* m_ptr, int m_cap, float* out, cudaStream_t stream);
}  // namespace nero

using namespace nero;

extern "C" {

int nero_version(void) { return 100; }

int nero_prep_weight(const float* v, const float* g, int K, int row0, int nrows, const int* kmap, float in_scale,
                     void* img_f, int rows_pad_f, void* img_t, int rows_pad_t, int t_c0, int t_ncols, float* w_eff, int ld_weff,
                     void* stream) {
  return prep_weight(v, g, K, row0, nrows, kmap, in_scale, (uint8_t*)img_f, rows_pad_f, (uint8_t*)img_t, rows_pad_t, t_c0, t_ncols,
                     w_eff, ld_weff, (cudaStream_t)stream);
}

int nero_linear(const float* A, int lda, int k_valid, const void* wimg, int n_pad, int k_chunks, const float* bias,
                float* out, int ldo, int ncol_out, float oscale, int mode, int act, float act_param,
                const float* H, int ldh, float hscale, int dact, const float* V, int ldv, float* out2, int ldo2,
                const float* addend, int ldadd, int ncol_main, float* tail, int ldt,
                const int* m_ptr, int m_cap, void* stream) {
  LinearParams p{A, lda, k_valid, (const uint8_t*)wimg, n_pad, k_chunks, bias, out, ldo, ncol_out, oscale, mode, act, act_param,
                 H, ldh, hscale, dact, V, ldv, out2, ldo2, addend, ldadd, ncol_main, tail, ldt, m_ptr, m_cap};
  if (!A || !wimg || !out || ncol_out > n_pad) return NERO_ERR_ARG;
  return linear_dispatch(p, (cudaStream_t)stream);
}

int nero_wgrad(const float* dY, int ldy, int n_valid, const float* X, int ldx, int k_valid,
               const float* dY2, int ldy2, const float* X2, int ldx2,
               float* partial, int ld_partial, int rows_partial, float* bias_partial,
               int n_rows_pad, int k_pad, int P, const int* m_ptr, int m_cap, void* stream) {
  WgradParams p{dY, ldy, X, ldx, dY2, ldy2, X2, ldx2, 0, n_valid, 0, k_valid, partial, ld_partial, rows_partial, bias_partial,
                m_ptr, m_cap};
  if (!dY || !X || !partial || (dY2 && !X2)) return NERO_ERR_ARG;
  return wgrad_dispatch(p, n_rows_pad, k_pad, P, (cudaStream_t)stream);
}

int nero_wgrad_finish(const float* partial, int P, int rows_partial, int ld_partial, const float* bias_partial, int K,
                      int row0, int nrows, const int* kmap, float in_scale, const float* v, const float* g,
                      float* grad_w, float* grad_g, float* grad_b, const float* extra_row, float extra_scale, void* stream) {
  return wgrad_finish(partial, P, rows_partial, ld_partial, bias_partial, K, row0, nrows, kmap, in_scale, v, g, grad_w, grad_g,
                      grad_b, extra_row, extra_scale, (cudaStream_t)stream);
}

int nero_colsum(const float* X, int ldx, int ncol, const float* w, int ldw, const int* m_ptr, int m_cap, float* out,
                void* stream) {
  return colsum(X, ldx, ncol, w, ldw, m_ptr, m_cap, out, (cudaStream_t)stream);
}

}  // extern "C"

/* libnero_b200 -- C ABI of the B200-native NeRO stage-I volume-rendering hot path.
 *
 * There is no FFI in the reference (it is pure PyTorch); the boundary this library sits behind is the Python
 * nn.Module contract of network/renderer.py (NeROShapeRenderer.render / render_core) -- see INTEGRATION.md.
 * Each entry point below replaces a group of PyTorch op sites of the reference; the file:line it replaces is
 * cited per function.  Conventions: all pointers are DEVICE pointers unless marked host; fp32 row-major
 * matrices with an explicit leading dimension (in floats); `stream` is a cudaStream_t passed as void*;
 * no allocation and no host synchronisation inside; data-dependent row counts are read from device memory
 * (`m_ptr`, clamped to the host-known capacity `m_cap`); the return value is 0 on success, non-zero on a bad
 * argument (1) or a CUDA launch error (2).
 */
#ifndef NERO_B200_H_
#define NERO_B200_H_
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* activation codes */
#define NERO_ACT_NONE 0
#define NERO_ACT_SOFTPLUS100 1 /* nn.Softplus(beta=100), network/field.py:124 */
#define NERO_ACT_RELU 2
#define NERO_ACT_SIGMOID 3
#define NERO_ACT_EXPCLAMP 4    /* exp(min(x, p)), network/field.py:302-308 */
/* epilogue modes of nero_linear */
#define NERO_EPI_BIAS_ACT 0
#define NERO_EPI_MUL_DACT 1
#define NERO_EPI_TANGENT 2

int nero_version(void);

/* ---- weights ------------------------------------------------------------------------------------------
 * Fold weight_norm (network/field.py:118-119, 324-331: W = g * v/||v||) and build the tensor-core operand images
 * (split-bf16, K-major SWIZZLE_128B, one 64-wide K chunk = [hi plane | lo plane]).  g == NULL: plain weight
 * (NeRFNetwork, field.py:239-256).  kmap (device, may be NULL) maps reference input column -> layout column.
 * img_t (for input-gradient GEMMs) only covers layout columns [t_c0, t_c0+t_ncols). */
int nero_prep_weight(const float* v, const float* g, int K, int row0, int nrows, const int* kmap, float in_scale,
                     void* img_f, int rows_pad_f, void* img_t, int rows_pad_t, int t_c0, int t_ncols, float* w_eff, int ld_weff,
                     void* stream);

/* ---- fused linear layer on tcgen05 tensor cores ----------------------------------------------------------
 * out = epilogue(A[M,K] * W^T).  Replaces nn.Linear + activation of SDFNetwork.forward (field.py:130-147),
 * NeRFNetwork.forward (field.py:258-283), make_predictor (field.py:310-346) and the autograd input-gradient /
 * double-backward GEMMs of SDFNetwork.gradient (field.py:155-167).
 *   mode BIAS_ACT : out = oscale * act(acc + bias)
 *   mode MUL_DACT : out[:, :ncol_main] = oscale * act'(H*hscale) * acc (+ addend);  tail[:, :] = oscale * acc[:, ncol_main:]
 *   mode TANGENT  : as MUL_DACT, plus out2 = 100 * (1 - act'(H*hscale)) * V * acc   (softplus'' term)       */
int nero_linear(const float* A, int lda, int k_valid, const void* wimg, int n_pad, int k_chunks, const float* bias,
                float* out, int ldo, int ncol_out, float oscale, int mode, int act, float act_param,
                const float* H, int ldh, float hscale, int dact, const float* V, int ldv, float* out2, int ldo2,
                const float* addend, int ldadd, int ncol_main, float* tail, int ldt,
                const int* m_ptr, int m_cap, void* stream);

/* ---- weight gradients -------------------------------------------------------------------------------------
 * partial[p] = dY^T X (+ dY2^T X2) over the p-th slice of rows; replaces the autograd weight-gradient GEMMs. */
int nero_wgrad(const float* dY, int ldy, int n_valid, const float* X, int ldx, int k_valid,
               const float* dY2, int ldy2, const float* X2, int ldx2,
               float* partial, int ld_partial, int rows_partial, float* bias_partial,
               int n_rows_pad, int k_pad, int P, const int* m_ptr, int m_cap, void* stream);
/* reduce partials, undo kmap, weight-norm chain rule, accumulate into grads (g == NULL: plain weight) */
int nero_wgrad_finish(const float* partial, int P, int rows_partial, int ld_partial, const float* bias_partial, int K,
                      int row0, int nrows, const int* kmap, float in_scale, const float* v, const float* g,
                      float* grad_w, float* grad_g, float* grad_b, const float* extra_row, float extra_scale, void* stream);
/* out[j] += sum_m w[m*ldw] * X[m, j]  (w NULL = 1) */
int nero_colsum(const float* X, int ldx, int ncol, const float* w, int ldw, const int* m_ptr, int m_cap, float* out,
                void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NERO_B200_H_ */

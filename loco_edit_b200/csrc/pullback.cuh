// Kernels around the U-Net on the editing-direction hot path: posterior-mean-predictor (PMP)
// epilogue and its transpose, basis orthonormalisation (the reference's per-iteration SVD),
// masked null-space projection, DDIM update, mask index/gather.  All HBM-bound.
#pragma once
#include "common.cuh"

namespace loco {

// reference: EditUncondDiffusion.get_x0 / get_et (modules/edit.py:2369-2403) differentiated along
// k directions.  v, eps_dot: [k, d] tangents of x_t and of eps; mask: d bytes (0/1) or null.
//   noise == 0:  u = mask * (v - eps_dot*sqrt(1-at)) / sqrt(at)
//   noise != 0:  u = mask * eps_dot
// Also emits the seeds of the transposed pass: g_eps = d<u, out>/d eps, gx_direct = d<u, out>/d x.
// Rows >= k_invert use the complement of the mask (pass k_invert = k for a single mask).
int pmp_jvp_epilogue(const float* v, const float* eps_dot, const unsigned char* mask, float at,
                     int noise, int k, int k_invert, long long d, float* u, float* g_eps,
                     float* gx_direct, cudaStream_t s);
// out = wa*a + wb*b + wc*c; b and c may be null
int combine3(const float* a, float wa, const float* b, float wb, const float* c, float wc, long long n,
             float* out, cudaStream_t s);
// PMP value for primal rows: P = (x - eps*sqrt(1-at))/sqrt(at)   (bit-compatible op order)
int pmp_forward(const float* x, const float* eps, float at, long long n, float* out, cudaStream_t s);

// G[i][j] = sum_d A[i][d] * B[j][d]   (double accumulation; G must be zeroed by the caller)
int gram(const float* A, int ka, const float* B, int kb, long long d, double* G, cudaStream_t s);

// Orthonormalise the k rows of W (k x d): G = W W^T = Q L Q^T, V = L^-1/2 Q^T W (rows sorted by
// descending eigenvalue) == Vh of svd(W) up to row signs; s_out[i] = L_i^(1/4) == sqrt of the
// singular values, which is what the reference returns (modules/edit.py:2482, 2499-2502).
// If v_prev != null the sign of every row is chosen so that <V_i, v_prev_i> >= 0.
// scratch: >= (3*k*k + 2*k) doubles.
int orthonormalise(const float* W, int k, long long d, const float* v_prev, float* V, float* s_out,
                   double* scratch, cudaStream_t s);

// reference: modules/edit.py:2317-2323.  project != 0:
//   vT = vT_mod - (Vn^T (Vn vT_mod^T))^T ; vT /= ||vT||_row     else: normalise only.
// scratch: >= (k_null*k + k) doubles.
int nullspace_project(const float* vT_mod, int k, const float* Vn, int k_null, long long d,
                      int project, float* out, double* scratch, cudaStream_t s);

// reference: YHCustomScheduler.step (utils/utils.py:342-374), eta = 0 and eta > 0 (noise given).
int ddim_step(const float* xt, const float* et, const float* noise, float at, float at_next,
              float eta, long long n, float* xt_next, float* x0_pred, cudaStream_t s);

// xt_edit = xt + scale * v   (reference: x_space_guidance_direct, modules/edit.py:2618-2625)
int axpy(const float* x, const float* v, float scale, long long n, float* out, cudaStream_t s);

// Ascending flat indices of the set bits of mask[0..d) (== row-major boolean selection order of
// P_xt[:, mask]).  count_out is a device int; idx must hold d entries.  Single-pass, exact.
int mask_indices(const unsigned char* mask, long long d, int* idx, int* count_out, cudaStream_t s);
// out[r][j] = src[r][idx[j]]
int gather_rows(const float* src, int rows, long long d, const int* idx, int count, float* out,
                cudaStream_t s);
// out[r][idx[j]] = src[r][j], zero elsewhere
int scatter_rows(const float* src, int rows, long long d, const int* idx, int count, float* out,
                 cudaStream_t s);

}  // namespace loco

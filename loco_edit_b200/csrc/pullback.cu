// PMP epilogue, basis orthonormalisation, null-space projection, DDIM update: see pullback.cuh.
#include "pullback.cuh"
#include <math.h>

namespace loco {

namespace {

inline int grid_for(long long total, int block, int cap = 148 * 8) {
  long long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

// ------------------------------------------------------------------------------------------------
// PMP
// ------------------------------------------------------------------------------------------------
__global__ void pmp_jvp_kernel(const float* __restrict__ v, const float* __restrict__ ed,
                               const unsigned char* __restrict__ mask, float at, int noise, int k,
                               int k_invert, long long d, float* __restrict__ u,
                               float* __restrict__ g_eps, float* __restrict__ gx) {
  const float s1 = __fsqrt_rn(__fsub_rn(1.0f, at));
  const float sa = __fsqrt_rn(at);
  const long long total = (long long)k * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long j = i % d;
    // rows >= k_invert use the complement mask (edit basis and null basis probed in one batch)
    const bool m = mask ? ((mask[j] != 0) != (i / d >= k_invert)) : true;
    float uu, ge, gd;
    if (noise) {
      uu = m ? ed[i] : 0.f;
      ge = uu;
      gd = 0.f;
    } else {
      uu = m ? __fdiv_rn(__fsub_rn(v[i], __fmul_rn(ed[i], s1)), sa) : 0.f;
      const float w = __fdiv_rn(uu, sa);
      ge = -__fmul_rn(w, s1);
      gd = w;
    }
    u[i] = uu;
    g_eps[i] = ge;
    gx[i] = gd;
  }
}

__global__ void pmp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ e, float at,
                               long long n, float* __restrict__ out) {
  const float s1 = __fsqrt_rn(__fsub_rn(1.0f, at));
  const float sa = __fsqrt_rn(at);
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = __fdiv_rn(__fsub_rn(x[i], __fmul_rn(e[i], s1)), sa);
}

// ------------------------------------------------------------------------------------------------
// Gram matrices
// ------------------------------------------------------------------------------------------------
// Small ranks (<= 8 rows each side): every thread keeps the whole ka x kb accumulator in registers
// and streams d with coalesced loads; one double atomic per entry per block.
// Products and sums are fp64 throughout (the product of two floats is exact in a double), so the
// Gram matrix carries the full fp32 precision of W: eigenvalues down to ~1e-14 of the largest are
// resolved, where fp32 partial sums would square the condition number into the 1e-7 rounding floor.
__global__ void __launch_bounds__(256)
gram_small_kernel(const float* __restrict__ A, int ka, const float* __restrict__ B, int kb,
                  long long d, double* __restrict__ G) {
  double acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0;
  // each block owns a contiguous slab
  const long long per_block = (d + gridDim.x - 1) / gridDim.x;
  const long long c0 = blockIdx.x * per_block;
  const long long c1 = min(d, c0 + per_block);
  for (long long c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
    double a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = i < ka ? (double)A[i * d + c] : 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) b[j] = j < kb ? (double)B[j * d + c] : 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
  }
  __shared__ double red[8][64];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const double v = warp_sum(acc[i][j]);
      if (lane == 0) red[warp][i * 8 + j] = v;
    }
  __syncthreads();
  if (threadIdx.x < 64) {
    const int i = threadIdx.x / 8, j = threadIdx.x % 8;
    if (i < ka && j < kb) {
      double v = 0;
      for (int w = 0; w < 8; ++w) v += red[w][threadIdx.x];
      atomicAdd(&G[i * kb + j], v);
    }
  }
}

// General ranks (<= 64): slab of 64 columns staged in shared memory, 16x16 threads x 4x4 entries.
constexpr int kGramSlab = 64;
__global__ void __launch_bounds__(256)
gram_big_kernel(const float* __restrict__ A, int ka, const float* __restrict__ B, int kb,
                long long d, double* __restrict__ G, int slabs_per_block) {
  __shared__ float As[64][kGramSlab + 1];
  __shared__ float Bs[64][kGramSlab + 1];
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  double acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
  for (int sl = 0; sl < slabs_per_block; ++sl) {
    const long long c0 = ((long long)blockIdx.x * slabs_per_block + sl) * kGramSlab;
    if (c0 >= d) break;
    for (int e = threadIdx.x; e < 64 * kGramSlab; e += 256) {
      const int r = e / kGramSlab, c = e % kGramSlab;
      const bool ok = c0 + c < d;
      As[r][c] = (r < ka && ok) ? A[r * d + c0 + c] : 0.f;
      Bs[r][c] = (r < kb && ok) ? B[r * d + c0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int c = 0; c < kGramSlab; ++c) {
      double a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = (double)As[ty * 4 + i][c];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = (double)Bs[tx * 4 + j][c];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fma(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = ty * 4 + i, c = tx * 4 + j;
      if (r < ka && c < kb) atomicAdd(&G[r * kb + c], acc[i][j]);
    }
}

// ------------------------------------------------------------------------------------------------
// k x k symmetric eigen-decomposition (cyclic Jacobi, fp64) -> transform T = L^-1/2 Q^T
// scratch layout: G[k*k] | C2[k*k] | T[k*k] | lambda[k] | (unused k)
// mode 0 / 1: T = diag(sign) L^-1/2 Q^T of G = W W^T, rows by descending eigenvalue (1: signs from
// C2 = W V_prev^T), s_out = L^(1/4).
// mode 2 (re-orthonormalisation pass, "CholeskyQR2" in its symmetric / Loewdin form): G = V1 V1^T
// of the first-pass result V1 = T W; T <- G^-1/2 T, which keeps row order and signs (G ~ I) and
// brings ||V V^T - I|| from cond(W)^2 * 1e-7 down to fp32 round-off.  s_out is not touched.
// ------------------------------------------------------------------------------------------------
constexpr int kEigThreads = 256;
__global__ void __launch_bounds__(kEigThreads)
eig_transform_kernel(double* __restrict__ scratch, int k, int mode,
                                     float* __restrict__ s_out) {
  const int has_prev = mode == 1;
  extern __shared__ double sm[];
  double* A = sm;                 // k*k
  double* Q = sm + k * k;         // k*k
  double* lam = Q + k * k;        // k
  int* order = reinterpret_cast<int*>(lam + k);
  const double* G = scratch;
  const double* C2 = scratch + k * k;
  double* T = scratch + 2 * k * k;
  const int t = threadIdx.x;
  for (int e = t; e < k * k; e += blockDim.x) {
    const int i = e / k, j = e % k;
    A[e] = 0.5 * (G[i * k + j] + G[j * k + i]);
    Q[e] = (i == j) ? 1.0 : 0.0;
  }
  __syncthreads();
  // Parallel-order cyclic Jacobi: a sweep is n - 1 rounds of n / 2 DISJOINT pairs (round-robin tournament
  // schedule, n = k rounded up to even with a dummy index), so the rotations of a round commute: all of their
  // parameters come from the same A, then all column updates (A J, Q J) and all row updates (J^T A) run in
  // parallel over (row, pair).  Same rotations as the serial cyclic order up to their sequence; 3 barriers per
  // ROUND instead of per rotation (k = 64: 63 x 3 per sweep instead of 2016 x 3 -- the serial version took 11 ms
  // of a 23 ms rank-64 orthonormalisation).
  const int n = (k + 1) & ~1, half = n / 2;
  double* cs = reinterpret_cast<double*>(order + ((k + 2) & ~1));   // [half][2] = (c, s) of the round's pairs
  int* pq = reinterpret_cast<int*>(cs + 2 * half);                   // [half][2] = (p, q), p < q, p = -1: idle pair
  __shared__ double red[2][32];
  for (int sweep = 0; sweep < 30; ++sweep) {
    double off = 0, diag = 0;
    for (int e = t; e < k * k; e += blockDim.x) {
      const double a = A[e];
      if (e / k == e % k) diag += a * a; else off += a * a;
    }
    off = warp_sum(off); diag = warp_sum(diag);
    if ((t & 31) == 0) { red[0][t >> 5] = off; red[1][t >> 5] = diag; }
    __syncthreads();
    off = 0; diag = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { off += red[0][w]; diag += red[1][w]; }   // fixed order, uniform
    __syncthreads();
    if (off <= 1e-30 * diag) break;
    for (int r = 0; r < n - 1; ++r) {
      if (t < half) {
        int p, q;
        if (t == 0) { p = n - 1; q = r; }
        else { p = (r + t) % (n - 1); q = (r - t + (n - 1)) % (n - 1); }
        if (p > q) { const int tmp = p; p = q; q = tmp; }
        double c = 1.0, sn = 0.0;
        if (q >= k) {
          p = -1;                                   // pair with the dummy index of an odd k
        } else {
          const double apq = A[p * k + q];
          if (fabs(apq) > 1e-300) {
            const double tau = (A[q * k + q] - A[p * k + p]) / (2.0 * apq);
            const double tt = (tau >= 0 ? 1.0 : -1.0) / (fabs(tau) + sqrt(1.0 + tau * tau));
            c = 1.0 / sqrt(1.0 + tt * tt);
            sn = tt * c;
          }
        }
        cs[2 * t] = c; cs[2 * t + 1] = sn;
        pq[2 * t] = p; pq[2 * t + 1] = q;
      }
      __syncthreads();
      for (int e = t; e < k * half; e += blockDim.x) {      // columns p, q of A and Q, row = e / half
        const int row = e / half, pr = e % half;
        const int p = pq[2 * pr], q = pq[2 * pr + 1];
        if (p < 0) continue;
        const double c = cs[2 * pr], sn = cs[2 * pr + 1];
        const double ap = A[row * k + p], aq = A[row * k + q];
        A[row * k + p] = c * ap - sn * aq;
        A[row * k + q] = sn * ap + c * aq;
        const double qp = Q[row * k + p], qq = Q[row * k + q];
        Q[row * k + p] = c * qp - sn * qq;
        Q[row * k + q] = sn * qp + c * qq;
      }
      __syncthreads();
      for (int e = t; e < k * half; e += blockDim.x) {      // rows p, q of A, column = e % k
        const int pr = e / k, col = e % k;
        const int p = pq[2 * pr], q = pq[2 * pr + 1];
        if (p < 0) continue;
        const double c = cs[2 * pr], sn = cs[2 * pr + 1];
        const double ap = A[p * k + col], aq = A[q * k + col];
        A[p * k + col] = c * ap - sn * aq;
        A[q * k + col] = sn * ap + c * aq;
      }
      __syncthreads();
    }
  }
  __syncthreads();
  if (mode == 2) {
    // A (k*k) is free once the eigenvalues are read: build S = Q L^-1/2 Q^T there, then T <- S T
    if (t < k) { double l = A[t * k + t]; lam[t] = 1.0 / sqrt(l < 1e-300 ? 1e-300 : l); }
    __syncthreads();
    for (int e = t; e < k * k; e += blockDim.x) {
      const int i = e / k, j = e % k;
      double acc = 0.0;
      for (int m = 0; m < k; ++m) acc += Q[i * k + m] * lam[m] * Q[j * k + m];
      A[e] = acc;
    }
    __syncthreads();
    for (int e = t; e < k * k; e += blockDim.x) {     // Q <- S T (Q is free now)
      const int i = e / k, j = e % k;
      double acc = 0.0;
      for (int m = 0; m < k; ++m) acc += A[i * k + m] * T[m * k + j];
      Q[e] = acc;
    }
    __syncthreads();
    for (int e = t; e < k * k; e += blockDim.x) T[e] = Q[e];
    return;
  }
  if (t == 0) {
    for (int i = 0; i < k; ++i) { lam[i] = A[i * k + i]; order[i] = i; }
    for (int i = 0; i < k; ++i) {          // selection sort, descending
      int best = i;
      for (int j = i + 1; j < k; ++j)
        if (lam[order[j]] > lam[order[best]]) best = j;
      const int tmp = order[i]; order[i] = order[best]; order[best] = tmp;
    }
  }
  __syncthreads();
  if (t < k) {
    const int col = order[t];
    double l = lam[col];
    if (l < 1e-300) l = 1e-300;
    const double inv = 1.0 / sqrt(l);
    double sign_acc = 0.0, big = 0.0;
    for (int j = 0; j < k; ++j) {
      const double tij = Q[j * k + col] * inv;
      if (has_prev) sign_acc += tij * C2[j * k + t];
      else if (fabs(tij) > fabs(big)) big = tij;
    }
    const double sg = has_prev ? (sign_acc < 0 ? -1.0 : 1.0) : (big < 0 ? -1.0 : 1.0);
    for (int j = 0; j < k; ++j) T[t * k + j] = sg * Q[j * k + col] * inv;
    s_out[t] = (float)sqrt(sqrt(l));
  }
}

// V[i][c] = sum_j T[i][j] W[j][c], accumulated in fp64: rows of T that belong to small singular
// values have entries ~ 1/sigma and cancel against nearly dependent rows of W, so fp32 products
// would lose cond(W) * 6e-8 of the result.
__global__ void apply_transform_kernel(const float* __restrict__ W, const double* __restrict__ T,
                                       int k, long long d, float* __restrict__ V) {
  extern __shared__ double Ts[];   // k*k
  for (int e = threadIdx.x; e < k * k; e += blockDim.x) Ts[e] = T[e];
  __syncthreads();
  for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < d;
       c += (long long)gridDim.x * blockDim.x) {
    for (int i0 = 0; i0 < k; i0 += 8) {
      double acc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = 0.0;
      for (int j = 0; j < k; ++j) {
        const double w = (double)W[j * d + c];
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (i0 + i < k) acc[i] = fma(Ts[(i0 + i) * k + j], w, acc[i]);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (i0 + i < k) V[(i0 + i) * d + c] = (float)acc[i];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Null-space projection
// ------------------------------------------------------------------------------------------------
// out[i][c] = Vm[i][c] - sum_j C[j][i] Vn[j][c]; norms[i] += out^2
__global__ void nullproj_kernel(const float* __restrict__ Vm, int k, const float* __restrict__ Vn,
                                int kn, long long d, const double* __restrict__ C, int project,
                                float* __restrict__ out, double* __restrict__ norms) {
  extern __shared__ float Cs[];   // kn*k
  for (int e = threadIdx.x; e < kn * k; e += blockDim.x) Cs[e] = project ? (float)C[e] : 0.f;
  __syncthreads();
  for (int i = 0; i < k; ++i) {
    float nrm = 0.f;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < d;
         c += (long long)gridDim.x * blockDim.x) {
      float v = Vm[i * d + c];
      if (project) {
        float p = 0.f;
        for (int j = 0; j < kn; ++j) p = fmaf(Cs[j * k + i], Vn[j * d + c], p);
        v -= p;
      }
      out[i * d + c] = v;
      nrm = fmaf(v, v, nrm);
    }
    const double w = warp_sum((double)nrm);
    if ((threadIdx.x & 31) == 0) atomicAdd(&norms[i], w);
  }
}
__global__ void scale_rows_kernel(float* __restrict__ x, int k, long long d,
                                  const double* __restrict__ norms) {
  const long long total = (long long)k * d;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const float nrm = (float)sqrt(norms[i / d]);
    x[i] = __fdiv_rn(x[i], nrm);
  }
}

// ------------------------------------------------------------------------------------------------
// DDIM step
// ------------------------------------------------------------------------------------------------
__global__ void ddim_step_kernel(const float* __restrict__ xt, const float* __restrict__ et,
                                 const float* __restrict__ noise, float at, float atn, float eta,
                                 long long n, float* __restrict__ xn, float* __restrict__ x0) {
  // scalar coefficients, evaluated with the reference's fp32 operation order
  const float s1 = __fsqrt_rn(__fsub_rn(1.0f, at));
  const float sa = __fsqrt_rn(at);
  const float san = __fsqrt_rn(atn);
  float dcoef, ncoef = 0.f;
  if (eta == 0.f) {
    dcoef = __fsqrt_rn(__fsub_rn(1.0f, atn));
  } else {
    const float sig = __fsqrt_rn(__fdiv_rn(
        __fmul_rn(__fsub_rn(1.0f, __fdiv_rn(at, atn)), __fsub_rn(1.0f, atn)), __fsub_rn(1.0f, at)));
    dcoef = __fsqrt_rn(__fsub_rn(__fsub_rn(1.0f, atn), __fmul_rn(eta, __fmul_rn(sig, sig))));
    ncoef = __fmul_rn(eta, sig);
  }
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    const float e = et[i];
    const float P = __fdiv_rn(__fsub_rn(xt[i], __fmul_rn(e, s1)), sa);
    float r = __fadd_rn(__fmul_rn(san, P), __fmul_rn(dcoef, e));
    if (eta != 0.f) r = __fadd_rn(r, __fmul_rn(ncoef, noise[i]));
    xn[i] = r;
    if (x0) x0[i] = P;
  }
}

__global__ void axpy_kernel(const float* __restrict__ x, const float* __restrict__ v, float scale,
                            long long n, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = __fadd_rn(x[i], __fmul_rn(scale, v[i]));
}

// out = wa*a + wb*b + wc*c (b, c optional): classifier-free-guidance combination of noise
// predictions (reference: modules/edit.py:660-673, 1326-1373) and of their Jacobian products
__global__ void combine3_kernel(const float* __restrict__ a, float wa, const float* __restrict__ b, float wb,
                                const float* __restrict__ c, float wc, long long n, float* __restrict__ out) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    float v = wa * a[i];
    if (b) v = fmaf(wb, b[i], v);
    if (c) v = fmaf(wc, c[i], v);
    out[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// Mask compaction / gather / scatter
// ------------------------------------------------------------------------------------------------
__global__ void mask_indices_kernel(const unsigned char* __restrict__ mask, long long d,
                                    int* __restrict__ idx, int* __restrict__ count_out) {
  __shared__ int counts[1024];
  const int t = threadIdx.x;
  const long long per = (d + blockDim.x - 1) / blockDim.x;
  const long long c0 = t * per, c1 = min(d, c0 + per);
  int cnt = 0;
  for (long long c = c0; c < c1; ++c) cnt += mask[c] != 0;
  counts[t] = cnt;
  __syncthreads();
  // inclusive Hillis-Steele scan
  for (int off = 1; off < blockDim.x; off <<= 1) {
    int v = 0;
    if (t >= off) v = counts[t - off];
    __syncthreads();
    counts[t] += v;
    __syncthreads();
  }
  int pos = counts[t] - cnt;
  for (long long c = c0; c < c1; ++c)
    if (mask[c] != 0) idx[pos++] = (int)c;
  if (t == blockDim.x - 1) *count_out = counts[t];
}
__global__ void gather_rows_kernel(const float* __restrict__ src, int rows, long long d,
                                   const int* __restrict__ idx, int count, float* __restrict__ out) {
  const long long total = (long long)rows * count;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = src[(i / count) * d + idx[i % count]];
}
__global__ void scatter_rows_kernel(const float* __restrict__ src, int rows, long long d,
                                    const int* __restrict__ idx, int count,
                                    float* __restrict__ out) {
  const long long total = (long long)rows * count;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    out[(i / count) * d + idx[i % count]] = src[i];
}

}  // namespace

int pmp_jvp_epilogue(const float* v, const float* eps_dot, const unsigned char* mask, float at,
                     int noise, int k, int k_invert, long long d, float* u, float* g_eps,
                     float* gx_direct, cudaStream_t s) {
  pmp_jvp_kernel<<<grid_for((long long)k * d, 256), 256, 0, s>>>(v, eps_dot, mask, at, noise, k,
                                                                k_invert, d, u, g_eps, gx_direct);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int combine3(const float* a, float wa, const float* b, float wb, const float* c, float wc, long long n,
             float* out, cudaStream_t s) {
  combine3_kernel<<<grid_for(n, 256), 256, 0, s>>>(a, wa, b, wb, c, wc, n, out);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int pmp_forward(const float* x, const float* eps, float at, long long n, float* out,
                cudaStream_t s) {
  pmp_fwd_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, eps, at, n, out);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int gram(const float* A, int ka, const float* B, int kb, long long d, double* G, cudaStream_t s) {
  LOCO_REQUIRE(ka >= 1 && kb >= 1 && ka <= 64 && kb <= 64, "gram: ranks (%d,%d) out of range", ka, kb);
  if (ka <= 8 && kb <= 8) {
    int blocks = (int)((d + 2047) / 2048);
    if (blocks > 148 * 4) blocks = 148 * 4;
    gram_small_kernel<<<blocks, 256, 0, s>>>(A, ka, B, kb, d, G);
  } else {
    const long long slabs = (d + kGramSlab - 1) / kGramSlab;
    int spb = (int)((slabs + 148 * 4 - 1) / (148 * 4));
    if (spb < 1) spb = 1;
    const int blocks = (int)((slabs + spb - 1) / spb);
    gram_big_kernel<<<blocks, 256, 0, s>>>(A, ka, B, kb, d, G, spb);
  }
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int orthonormalise(const float* W, int k, long long d, const float* v_prev, float* V, float* s_out,
                   double* scratch, cudaStream_t s) {
  LOCO_REQUIRE(k >= 1 && k <= 64, "orthonormalise: rank %d out of range [1,64]", k);
  LOCO_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * (size_t)(3 * k * k + 2 * k), s));
  LOCO_TRY(gram(W, k, W, k, d, scratch, s));
  if (v_prev) LOCO_TRY(gram(W, k, v_prev, k, d, scratch + k * k, s));
  const size_t smem = sizeof(double) * (size_t)(2 * k * k + k) + sizeof(int) * (size_t)((k + 2) & ~1) +
                      (sizeof(double) + sizeof(int)) * (size_t)(k + 2);   // A | Q | lambda | order | (c, s) | (p, q)
  static bool attr_set[kMaxDevices] = {false};
  if (first_time_on_device(attr_set)) {
    LOCO_CHECK_CUDA(cudaFuncSetAttribute(eig_transform_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024));
  }
  eig_transform_kernel<<<1, kEigThreads, smem, s>>>(scratch, k, v_prev ? 1 : 0, s_out);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  apply_transform_kernel<<<grid_for(d, 256), 256, sizeof(double) * (size_t)(k * k), s>>>(
      W, scratch + 2 * k * k, k, d, V);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  // second pass: V1 V1^T = I + E with |E| ~ cond(W)^2 * 1e-7 (the Gram route squares the condition
  // number); fold G2^-1/2 into the transform and re-apply it to W (no [k,d] temporary needed)
  LOCO_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * (size_t)(k * k), s));
  LOCO_TRY(gram(V, k, V, k, d, scratch, s));
  eig_transform_kernel<<<1, kEigThreads, smem, s>>>(scratch, k, 2, s_out);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  apply_transform_kernel<<<grid_for(d, 256), 256, sizeof(double) * (size_t)(k * k), s>>>(
      W, scratch + 2 * k * k, k, d, V);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int nullspace_project(const float* vT_mod, int k, const float* Vn, int k_null, long long d,
                      int project, float* out, double* scratch, cudaStream_t s) {
  LOCO_REQUIRE(k >= 1 && k <= 64 && k_null >= 0 && k_null <= 64, "nullspace_project: bad ranks");
  double* C = scratch;
  double* norms = scratch + (size_t)k_null * k;
  LOCO_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double) * (size_t)(k_null * k + k), s));
  if (project && k_null > 0) LOCO_TRY(gram(Vn, k_null, vT_mod, k, d, C, s));
  nullproj_kernel<<<grid_for(d, 256), 256, sizeof(float) * (size_t)(k_null * k + 1), s>>>(
      vT_mod, k, Vn, k_null, d, C, (project && k_null > 0) ? 1 : 0, out, norms);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  scale_rows_kernel<<<grid_for((long long)k * d, 256), 256, 0, s>>>(out, k, d, norms);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int ddim_step(const float* xt, const float* et, const float* noise, float at, float at_next,
              float eta, long long n, float* xt_next, float* x0_pred, cudaStream_t s) {
  LOCO_REQUIRE(eta == 0.f || noise != nullptr, "ddim_step: eta > 0 needs a noise tensor");
  ddim_step_kernel<<<grid_for(n, 256), 256, 0, s>>>(xt, et, noise, at, at_next, eta, n, xt_next,
                                                   x0_pred);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int axpy(const float* x, const float* v, float scale, long long n, float* out, cudaStream_t s) {
  axpy_kernel<<<grid_for(n, 256), 256, 0, s>>>(x, v, scale, n, out);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int mask_indices(const unsigned char* mask, long long d, int* idx, int* count_out, cudaStream_t s) {
  LOCO_REQUIRE(d < (1LL << 31), "mask_indices: d too large");
  mask_indices_kernel<<<1, 1024, 0, s>>>(mask, d, idx, count_out);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int gather_rows(const float* src, int rows, long long d, const int* idx, int count, float* out,
                cudaStream_t s) {
  if (count == 0 || rows == 0) return 0;
  gather_rows_kernel<<<grid_for((long long)rows * count, 256), 256, 0, s>>>(src, rows, d, idx, count,
                                                                           out);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int scatter_rows(const float* src, int rows, long long d, const int* idx, int count, float* out,
                 cudaStream_t s) {
  LOCO_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)rows * (size_t)d, s));
  if (count == 0 || rows == 0) return 0;
  scatter_rows_kernel<<<grid_for((long long)rows * count, 256), 256, 0, s>>>(src, rows, d, idx,
                                                                            count, out);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace loco

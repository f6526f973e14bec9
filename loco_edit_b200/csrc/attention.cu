// Attention core (scores, softmax, value mixing) with JVP and VJP: see attention.cuh.
#include "attention.cuh"
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

namespace loco {

namespace {

constexpr int BM = 32, BN = 64, BK = 16;
constexpr int kGemmGroup = 128;                 // threads that own one K slab at a time
constexpr int kGemmThreads = 2 * kGemmGroup;

// 32x64 output tile per block.  The attention operands are tiny (256 or 64 tokens), so a launch has
// only about one block per SM and is bound by the latency of its K loop, not by FMA throughput:
// the block therefore runs TWO 128-thread groups (8x16 threads, 4x4 outputs each) that walk
// alternate K slabs with their own shared-memory tiles and named barriers, which halves the
// length of the dependent chain; the next slab of a group is prefetched into registers while the
// current one is multiplied, and group 1 hands its partial sums to group 0 through shared memory
// (fixed order: bit-reproducible).
__global__ void __launch_bounds__(kGemmThreads)
batched_gemm_kernel(GemmOperand A, GemmOperand B, float* __restrict__ C, long long sCm,
                    long long sCn, long long sCb, long long sCh, int heads, int M, int N, int K,
                    float alpha, float beta, int round_out, int a_kcontig, int b_ncontig) {
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  __shared__ float red[kGemmGroup][17];
  const int b = blockIdx.z / heads;       // batch row
  const int hd = blockIdx.z % heads;      // attention head
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const float* Ab = A.ptr + b * A.sb + hd * A.sh;
  const float* Bb = B.ptr + b * B.sb + hd * B.sh;
  const int grp = threadIdx.x / kGemmGroup;
  const int tid = threadIdx.x % kGemmGroup;
  const int tx = tid % 16, ty = tid / 16;   // 8x16 threads, 4x4 outputs each
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  float ra[4], rb[8];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + kGemmGroup * r;
      int m, k;
      if (a_kcontig) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      ra[r] = (m0 + m < M && k0 + k < K) ? Ab[(long long)(m0 + m) * A.s0 + (long long)(k0 + k) * A.s1] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int e = tid + kGemmGroup * r;
      int n, k;
      if (b_ncontig) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
      rb[r] = (n0 + n < N && k0 + k < K) ? Bb[(long long)(k0 + k) * B.s0 + (long long)(n0 + n) * B.s1] : 0.f;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int e = tid + kGemmGroup * r;
      int m, k;
      if (a_kcontig) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      As[grp][k][m] = ra[r];
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      const int e = tid + kGemmGroup * r;
      int n, k;
      if (b_ncontig) { n = e % BN; k = e / BN; } else { k = e % BK; n = e / BK; }
      Bs[grp][k][n] = rb[r];
    }
  };
  auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(1 + grp), "n"(kGemmGroup) : "memory"); };
  const int kstep = 2 * BK;
  if (grp * BK < K) fetch(grp * BK);
  for (int k0 = grp * BK; k0 < K; k0 += kstep) {
    stash();
    group_sync();
    if (k0 + kstep < K) fetch(k0 + kstep);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[4], bb[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[grp][k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) bb[j] = Bs[grp][k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
    }
    group_sync();
  }
  if (grp == 1) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) red[tid][i * 4 + j] = acc[i][j];
  }
  __syncthreads();
  if (grp == 1) return;
  float* Cb = C + b * sCb + hd * sCh;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float* cp = Cb + (long long)m * sCm + (long long)n * sCn;
      float v = alpha * (acc[i][j] + red[tid][i * 4 + j]);
      if (beta != 0.f) v += beta * (*cp);
      if (round_out) v = round_tf32(v);
      *cp = v;
    }
  }
}


// ---- tensor-core version for the long-sequence case -------------------------------------------------
// The VAE decoder's mid-block attention runs over 64 x 64 = 4096 tokens (512 channels, one head): 17 GFLOP per
// product and row, 40 products per Jacobian product of the latent-space twin -- half of a decoder pass on the
// fp32 CUDA-core kernel above (25 TFLOP/s).  Same contract (strided operands, alpha / beta / rounding), products on
// warp-level mma.sync.m16n8k8 with tf32 operands (the 10-bit mantissa every other GEMM of the path uses) and fp32
// accumulation: 128 x BN output tile per block, 8 warps of 64 x BN/4, K slabs of 32 staged in shared memory in the
// orientation of the operand's contiguous index (pitches 36 / 8 mod 32: conflict-free fragment loads), next slab
// prefetched into registers with 16-byte loads.  (tcgen05 needs TMA-describable, K-major 16-bit or tf32 tiles; half
// of these products have an M- or N-major operand in fp32 and would need a transposing copy of a 64 MB matrix first.)
__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
constexpr int kTM = 128, kTK = 32;
template <int BN_, bool AKC, bool BNC>
__global__ void __launch_bounds__(256)
batched_gemm_mma_kernel(GemmOperand A, GemmOperand B, float* __restrict__ C, long long sCm, long long sCn,
                        long long sCb, long long sCh, int heads, int M, int N, int K, float alpha, float beta,
                        int round_out) {
  constexpr int WN = BN_ / 4, NT = WN / 8;
  constexpr int AP = AKC ? kTK + 4 : kTM + 8;               // A tile: [m][k] (K contiguous) or [k][m]
  constexpr int BP = BNC ? BN_ + 8 : kTK + 4;               // B tile: [k][n] (N contiguous) or [n][k]
  constexpr int ASZ = AKC ? kTM * AP : kTK * AP;
  constexpr int BSZ = BNC ? kTK * BP : BN_ * BP;
  constexpr int AV = kTM * kTK / 4 / 256;                   // 16-byte vectors per thread and slab
  constexpr int BV = BN_ * kTK / 4 / 256;
  __shared__ __align__(16) float As[ASZ];
  __shared__ __align__(16) float Bs[BSZ];
  const int b = blockIdx.z / heads, hd = blockIdx.z % heads;
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * BN_;
  const float* __restrict__ Ab = A.ptr + b * A.sb + hd * A.sh;
  const float* __restrict__ Bb = B.ptr + b * B.sb + hd * B.sh;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wm = (warp >> 2) * 64, wn = (warp & 3) * WN;
  float acc[4][NT][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int q = 0; q < 4; ++q) acc[i][j][q] = 0.f;
  float4 ra[AV], rb[BV];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int r = 0; r < AV; ++r) {
      const int e = tid + 256 * r;
      if (AKC) { const int m = e >> 3, kq = e & 7; ra[r] = *reinterpret_cast<const float4*>(Ab + (long long)(m0 + m) * A.s0 + k0 + 4 * kq); }
      else { const int k = e >> 5, mq = e & 31; ra[r] = *reinterpret_cast<const float4*>(Ab + (long long)(k0 + k) * A.s1 + m0 + 4 * mq); }
    }
#pragma unroll
    for (int r = 0; r < BV; ++r) {
      const int e = tid + 256 * r;
      if (BNC) { const int k = e / (BN_ / 4), nq = e % (BN_ / 4); rb[r] = *reinterpret_cast<const float4*>(Bb + (long long)(k0 + k) * B.s0 + n0 + 4 * nq); }
      else { const int n = e >> 3, kq = e & 7; rb[r] = *reinterpret_cast<const float4*>(Bb + (long long)(n0 + n) * B.s1 + k0 + 4 * kq); }
    }
  };
  auto rnd4 = [](float4 v) { return make_float4(round_tf32(v.x), round_tf32(v.y), round_tf32(v.z), round_tf32(v.w)); };
  auto stash = [&]() {
#pragma unroll
    for (int r = 0; r < AV; ++r) {
      const int e = tid + 256 * r;
      if (AKC) { const int m = e >> 3, kq = e & 7; *reinterpret_cast<float4*>(&As[m * AP + 4 * kq]) = rnd4(ra[r]); }
      else { const int k = e >> 5, mq = e & 31; *reinterpret_cast<float4*>(&As[k * AP + 4 * mq]) = rnd4(ra[r]); }
    }
#pragma unroll
    for (int r = 0; r < BV; ++r) {
      const int e = tid + 256 * r;
      if (BNC) { const int k = e / (BN_ / 4), nq = e % (BN_ / 4); *reinterpret_cast<float4*>(&Bs[k * BP + 4 * nq]) = rnd4(rb[r]); }
      else { const int n = e >> 3, kq = e & 7; *reinterpret_cast<float4*>(&Bs[n * BP + 4 * kq]) = rnd4(rb[r]); }
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += kTK) {
    __syncthreads();                      // previous slab consumed
    stash();
    __syncthreads();
    if (k0 + kTK < K) fetch(k0 + kTK);
#pragma unroll
    for (int ks = 0; ks < kTK / 8; ++ks) {
      const int c0 = ks * 8 + t, c1 = c0 + 4;
      uint32_t bf[NT][2];
#pragma unroll
      for (int j = 0; j < NT; ++j) {
        const int n = wn + j * 8 + g;
        bf[j][0] = __float_as_uint(BNC ? Bs[c0 * BP + n] : Bs[n * BP + c0]);
        bf[j][1] = __float_as_uint(BNC ? Bs[c1 * BP + n] : Bs[n * BP + c1]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r0 = wm + i * 16 + g, r1 = r0 + 8;
        uint32_t af[4];
        af[0] = __float_as_uint(AKC ? As[r0 * AP + c0] : As[c0 * AP + r0]);
        af[1] = __float_as_uint(AKC ? As[r1 * AP + c0] : As[c0 * AP + r1]);
        af[2] = __float_as_uint(AKC ? As[r0 * AP + c1] : As[c1 * AP + r0]);
        af[3] = __float_as_uint(AKC ? As[r1 * AP + c1] : As[c1 * AP + r1]);
#pragma unroll
        for (int j = 0; j < NT; ++j) mma_tf32(acc[i][j], af, bf[j][0], bf[j][1]);
      }
    }
  }
  float* Cb = C + b * sCb + hd * sCh;
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < NT; ++j)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int m = m0 + wm + i * 16 + g + 8 * h;
        const int n = n0 + wn + j * 8 + 2 * t;
        float* cp = Cb + (long long)m * sCm + (long long)n * sCn;
        float v0 = alpha * acc[i][j][2 * h], v1 = alpha * acc[i][j][2 * h + 1];
        if (beta != 0.f) { v0 += beta * cp[0]; v1 += beta * cp[sCn]; }
        if (round_out) { v0 = round_tf32(v0); v1 = round_tf32(v1); }
        cp[0] = v0; cp[sCn] = v1;
      }
}

// One warp per row.
__global__ void softmax_rows_kernel(float* __restrict__ S, int T, long long rows, float scale) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  float* p = S + row * T;
  float mx = -INFINITY;
  for (int j = lane; j < T; j += 32) mx = fmaxf(mx, p[j] * scale);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float sum = 0.f;
  for (int j = lane; j < T; j += 32) {
    const float e = expf(p[j] * scale - mx);
    p[j] = e;
    sum += e;
  }
  sum = warp_sum(sum);
  const float inv = 1.0f / sum;
  for (int j = lane; j < T; j += 32) p[j] *= inv;
}

// rows of X are [batch][heads * T]; P0 holds the primal probabilities [heads * T][T]
__global__ void softmax_lin_rows_kernel(const float* __restrict__ P0, float* __restrict__ X, int T,
                                        long long rows, long long prow_mod, float scale) {
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* p = P0 + (row % prow_mod) * T;
  float* x = X + row * T;
  float dot = 0.f;
  for (int j = lane; j < T; j += 32) dot += p[j] * x[j];
  dot = warp_sum(dot);
  for (int j = lane; j < T; j += 32) x[j] = scale * p[j] * (x[j] - dot);
}

int check_tokens(const View& v, const char* what) {
  LOCO_REQUIRE(v.sH == (long long)v.W * v.sW, "%s: tokens must be uniformly strided", what);
  return 0;
}

}  // namespace

int attention_init() {
  static bool done[kMaxDevices] = {false};
  if (!first_time_on_device(done)) return 0;
  const int co = cudaSharedmemCarveoutMaxShared;
  LOCO_CHECK_CUDA(cudaFuncSetAttribute(batched_gemm_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co));
  LOCO_CHECK_CUDA(cudaFuncSetAttribute(softmax_rows_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co));
  LOCO_CHECK_CUDA(cudaFuncSetAttribute(softmax_lin_rows_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, co));
  return attention_tc_init();
}

static bool gemm_mma_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LOCO_ATTN_MMA");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int batched_gemm(GemmOperand A, GemmOperand B, float* C, long long sCm, long long sCn, long long sCb,
                 long long sCh, int M, int N, int K, int batch, int heads, float alpha, float beta,
                 int round_out, cudaStream_t s) {
  if (batch <= 0) return 0;
  // long sequences: the mma.sync kernel (whole tiles, one unit stride per operand, 16-byte aligned vectors)
  {
    const bool akc = A.s1 == 1, amc = A.s0 == 1, bnc = B.s1 == 1, bkc = B.s0 == 1;
    const int bn = (N % 128 == 0) ? 128 : 64;
    auto al4 = [](long long v) { return v % 4 == 0; };
    const bool vec = (akc ? al4(A.s0) : al4(A.s1)) && (bnc ? al4(B.s0) : al4(B.s1)) && al4(A.sb) && al4(A.sh) &&
                     al4(B.sb) && al4(B.sh) && (((uintptr_t)A.ptr | (uintptr_t)B.ptr) & 15) == 0;
    if (gemm_mma_enabled() && (long long)M * N * K >= (1LL << 27) && M % kTM == 0 && N % 64 == 0 && K % kTK == 0 &&
        (akc || amc) && (bnc || bkc) && vec) {
      dim3 grid(N / bn, M / kTM, batch * heads);
#define LOCO_GEMM_MMA(BN_, AKC, BNC)                                                                             \
  batched_gemm_mma_kernel<BN_, AKC, BNC><<<grid, 256, 0, s>>>(A, B, C, sCm, sCn, sCb, sCh, heads, M, N, K, alpha, \
                                                              beta, round_out)
      if (bn == 128) {
        if (akc) { if (bnc) LOCO_GEMM_MMA(128, true, true); else LOCO_GEMM_MMA(128, true, false); }
        else { if (bnc) LOCO_GEMM_MMA(128, false, true); else LOCO_GEMM_MMA(128, false, false); }
      } else {
        if (akc) { if (bnc) LOCO_GEMM_MMA(64, true, true); else LOCO_GEMM_MMA(64, true, false); }
        else { if (bnc) LOCO_GEMM_MMA(64, false, true); else LOCO_GEMM_MMA(64, false, false); }
      }
#undef LOCO_GEMM_MMA
      count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
      return 0;
    }
  }
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch * heads);
  batched_gemm_kernel<<<grid, kGemmThreads, 0, s>>>(A, B, C, sCm, sCn, sCb, sCh, heads, M, N, K, alpha,
                                                    beta, round_out, A.s1 == 1 ? 1 : 0, B.s1 == 1 ? 1 : 0);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int softmax_rows(float* S, int T, int batch, float scale, cudaStream_t s) {
  if (batch <= 0) return 0;
  const long long rows = (long long)batch * T;
  softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(S, T, rows, scale);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int softmax_lin_rows(const float* P0, float* X, int T, int batch, int heads, float scale,
                     cudaStream_t s) {
  if (batch <= 0) return 0;
  const long long rows = (long long)batch * heads * T;
  softmax_lin_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, s>>>(P0, X, T, rows,
                                                                     (long long)heads * T, scale);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

namespace {
// Channel geometry of the fused q|k|v tensor.
//   head_ch == 0: DDPM AttnBlock, one head over all C channels, q | k | v blocks of C channels,
//                 scores scaled by C^-1/2 (ddpm/diffusion.py:941-966);
//   head_ch  > 0: guided-diffusion QKVAttentionLegacy, heads split BEFORE q/k/v: head h owns the
//                 channels [3 h dh, 3 (h+1) dh) = q | k | v of dh each, scores scaled by dh^-1/2
//                 (q and k each by dh^-1/4; guided_diffusion/unet.py:339-356).
struct HeadGeom { int heads, dh; long long qo, ko, vo, hs; float scale; };
HeadGeom head_geom(int C, int head_ch) {
  HeadGeom g;
  if (head_ch <= 0) { g.heads = 1; g.dh = C; g.qo = 0; g.ko = C; g.vo = 2 * C; g.hs = 0; }
  else { g.heads = C / head_ch; g.dh = head_ch; g.qo = 0; g.ko = head_ch; g.vo = 2 * head_ch; g.hs = 3 * head_ch; }
  g.scale = 1.0f / sqrtf((float)g.dh);
  return g;
}
}  // namespace

int attention_forward(View qkv, int n_primal, int head_ch, float* S, View o, cudaStream_t s) {
  LOCO_TRY(check_tokens(qkv, "attention_forward(qkv)"));
  LOCO_TRY(check_tokens(o, "attention_forward(o)"));
  const int C = qkv.C / 3;
  const int T = qkv.H * qkv.W;
  const int N = qkv.N;
  const int nt = N - n_primal;
  LOCO_REQUIRE(o.C == C && o.N == N, "attention_forward: shape mismatch");
  LOCO_REQUIRE(nt == 0 || n_primal == 1, "attention_forward: tangents need exactly one primal row");
  LOCO_REQUIRE(head_ch <= 0 || C % head_ch == 0, "attention_forward: %d channels, heads of %d", C, head_ch);
  if (attention_tc_eligible(T, C, head_ch) && !qkv.half && !o.half)
    return attention_forward_tc(qkv, n_primal, head_ch, S, o, s);
  const HeadGeom G = head_geom(C, head_ch);
  const int Hh = G.heads, D = G.dh;
  const float scale = G.scale;
  const long long ts = qkv.sW;                  // token stride
  const long long TT = (long long)T * T;        // one head's score matrix
  const long long HTT = TT * Hh;                // one batch row's score matrices
  const float* q = qkv.ptr + G.qo;
  const float* k = qkv.ptr + G.ko;
  const float* v = qkv.ptr + G.vo;
  // primal rows: S = q k^T ; P = softmax(scale S) ; o = P v
  LOCO_TRY(batched_gemm({q, ts, 1, qkv.sN, G.hs}, {k, 1, ts, qkv.sN, G.hs}, S, T, 1, HTT, TT, T, T, D,
                        n_primal, Hh, 1.f, 0.f, 0, s));
  LOCO_TRY(softmax_rows(S, T, n_primal * Hh, scale, s));
  LOCO_TRY(batched_gemm({S, T, 1, HTT, TT}, {v, ts, 1, qkv.sN, G.hs}, o.ptr, o.sW, 1, o.sN, D, T, D, T,
                        n_primal, Hh, 1.f, 0.f, 1, s));
  if (nt > 0) {
    const float* qd = q + qkv.sN;
    const float* kd = k + qkv.sN;
    const float* vd = v + qkv.sN;
    float* Sd = S + HTT;
    float* od = o.ptr + o.sN;
    // dS = dq k0^T + q0 dk^T
    LOCO_TRY(batched_gemm({qd, ts, 1, qkv.sN, G.hs}, {k, 1, ts, 0, G.hs}, Sd, T, 1, HTT, TT, T, T, D, nt,
                          Hh, 1.f, 0.f, 0, s));
    LOCO_TRY(batched_gemm({q, ts, 1, 0, G.hs}, {kd, 1, ts, qkv.sN, G.hs}, Sd, T, 1, HTT, TT, T, T, D, nt,
                          Hh, 1.f, 1.f, 0, s));
    // dP = scale * P0 o (dS - rowsum(P0 o dS))
    LOCO_TRY(softmax_lin_rows(S, Sd, T, nt, Hh, scale, s));
    // do = P0 dv + dP v0
    LOCO_TRY(batched_gemm({S, T, 1, 0, TT}, {vd, ts, 1, qkv.sN, G.hs}, od, o.sW, 1, o.sN, D, T, D, T, nt,
                          Hh, 1.f, 0.f, 0, s));
    LOCO_TRY(batched_gemm({Sd, T, 1, HTT, TT}, {v, ts, 1, 0, G.hs}, od, o.sW, 1, o.sN, D, T, D, T, nt, Hh,
                          1.f, 1.f, 1, s));
  }
  return 0;
}

int attention_vjp(View go, View qkv0, int head_ch, const float* P0, float* gP, View gqkv,
                  cudaStream_t s) {
  LOCO_TRY(check_tokens(go, "attention_vjp(go)"));
  LOCO_TRY(check_tokens(qkv0, "attention_vjp(qkv0)"));
  LOCO_TRY(check_tokens(gqkv, "attention_vjp(gqkv)"));
  const int C = qkv0.C / 3;
  const int T = qkv0.H * qkv0.W;
  const int K = go.N;
  LOCO_REQUIRE(go.C == C && gqkv.C == 3 * C && gqkv.N == K, "attention_vjp: shape mismatch");
  if (attention_tc_eligible(T, C, head_ch) && !go.half && !qkv0.half && !gqkv.half)
    return attention_vjp_tc(go, qkv0, head_ch, P0, gP, gqkv, s);
  const HeadGeom G = head_geom(C, head_ch);
  const int Hh = G.heads, D = G.dh;
  const float scale = G.scale;
  const long long ts = qkv0.sW, gts = gqkv.sW;
  const long long TT = (long long)T * T;
  const long long HTT = TT * Hh;
  const float* q0 = qkv0.ptr + G.qo;
  const float* k0 = qkv0.ptr + G.ko;
  const float* v0 = qkv0.ptr + G.vo;
  float* gq = gqkv.ptr + G.qo;
  float* gk = gqkv.ptr + G.ko;
  float* gv = gqkv.ptr + G.vo;
  // gv = P0^T go      : (j, c) = sum_i P0[i][j] go[i][c]
  LOCO_TRY(batched_gemm({P0, 1, T, 0, TT}, {go.ptr, go.sW, 1, go.sN, D}, gv, gts, 1, gqkv.sN, G.hs, T, D,
                        T, K, Hh, 1.f, 0.f, 1, s));
  // gP = go v0^T      : (i, j) = sum_c go[i][c] v0[j][c]
  LOCO_TRY(batched_gemm({go.ptr, go.sW, 1, go.sN, D}, {v0, 1, ts, 0, G.hs}, gP, T, 1, HTT, TT, T, T, D, K,
                        Hh, 1.f, 0.f, 0, s));
  // gS = scale * P0 o (gP - rowsum(P0 o gP))   (gradient w.r.t. the unscaled scores)
  LOCO_TRY(softmax_lin_rows(P0, gP, T, K, Hh, scale, s));
  // gq = gS k0        : (i, c) = sum_j gS[i][j] k0[j][c]
  LOCO_TRY(batched_gemm({gP, T, 1, HTT, TT}, {k0, ts, 1, 0, G.hs}, gq, gts, 1, gqkv.sN, G.hs, T, D, T, K,
                        Hh, 1.f, 0.f, 1, s));
  // gk = gS^T q0      : (j, c) = sum_i gS[i][j] q0[i][c]
  LOCO_TRY(batched_gemm({gP, 1, T, HTT, TT}, {q0, ts, 1, 0, G.hs}, gk, gts, 1, gqkv.sN, G.hs, T, D, T, K,
                        Hh, 1.f, 0.f, 1, s));
  return 0;
}

}  // namespace loco

// Fused attention core on the Blackwell tensor cores (tcgen05.mma kind::tf32, accumulators in TMEM,
// operands staged by TMA): scores, softmax and value mixing of the DDPM AttnBlock
// (reference: ddpm/diffusion.py:941-966) and of guided-diffusion's QKVAttentionLegacy heads
// (guided_diffusion/unet.py:339-356), with the JVP (tangent rows) and the VJP (cotangent rows).
// The score matrix of a 128-query block lives in TMEM, the probabilities in shared memory; the only
// HBM traffic besides q|k|v and o is the primal probability matrix P0, which the tangent and the
// cotangent kernels need (softmax linearisation) and which is written once by the primal kernel.
//
//   primal   (row n, head h, 128 queries):  S = q k^T -> TMEM;  P = softmax(scale S) -> smem (+ HBM);
//                                           o = P v -> TMEM -> HBM
//   tangent  (row r):   dS = dq k0^T + q0 dk^T;  dP = scale P0 o (dS - rowsum(P0 o dS));
//                       do = dP v0 + P0 dv
//   cotangent, kernel 1 (row r, 128 queries):  gP = go v0^T;  gS = scale P0 o (gP - rowsum(P0 o gP))
//                       -> smem (+ HBM);  gq = gS k0
//   cotangent, kernel 2 (row r, 128 keys):     gv = P0^T go;  gk = gS^T q0   (contraction over ALL
//                       queries inside one CTA: no cross-CTA reduction, bit-reproducible)
//
// Operand layouts.  q, k, go rows are K-major operands (the contraction runs over channels, which
// are contiguous in a token row): TMA boxes of 32 channels x 128 tokens land in 128B-swizzled
// shared memory exactly like the conv kernels' operands.  v, k0, q0, go as the SECOND factor of a
// product over tokens are MN-major operands (tokens = K rows, channels contiguous along N): a
// 4-D tensor map (32 channels, tokens, channel slab, batch row) delivers [slab][16 tokens][128 B]
// chunks, the canonical MN-major SWIZZLE_128B atoms (8 K-rows x 128 B, LBO = slab pitch); P0^T / gS^T
// as the FIRST factor read the [token rows][32 keys] probability slabs MN-major the same way.
//
// CTA = 6 warps: 0-3 own one TMEM lane quarter each (softmax, linearisation, epilogues: a thread
// owns one query / key row, so row reductions need no shuffles), 4 = TMA producer, 5 = MMA issuer.
#include "attention.cuh"
#include <math.h>
#include <stdlib.h>
#include <string.h>

namespace loco {

namespace {

constexpr int kTcThreads = 192;
constexpr int kTcStageBytes = 48 * 1024;   // K-major stage: A slab 16 KB + B slab(s) 32 KB; MN chunk: <= 32 KB
constexpr int kTcStages = 2;
constexpr int kTcSlab = 16 * 1024;         // 128 rows x 128 B
constexpr int kTcPBytes = 8 * kTcSlab;     // 128 rows x 256 keys fp32 as 8 K-major slabs
constexpr int kTcSmem = kTcPBytes + kTcStages * kTcStageBytes + 1024 + 256;
constexpr float kLog2e = 1.4426950408889634f;

struct AttnTcParams {
  CUtensorMap qkv_k;     // q|k|v rows, K-major boxes: dims (Cq, T, N), box (32, 128, 1)
  CUtensorMap qkv_mn;    // q|k|v rows, MN-major chunks: dims (32, T, Cq / 32, N), box (32, 16, D / 32, 1)
  CUtensorMap p_map;     // primal probabilities P0: dims (T, T, heads * rows), box (32, 128, 1)
  CUtensorMap go_k;      // VJP: cotangent of o, K-major boxes: dims (C, T, K), box (32, 128, 1)
  CUtensorMap go_mn;     // VJP: cotangent of o, MN-major chunks
  CUtensorMap gs_map;    // VJP kernel 2: gS scratch read MN-major, dims (T, T, heads * K), box (32, 128, 1), 32B-atom swizzle
  CUtensorMap p_map_mn;  // VJP kernel 2: P0 read MN-major (32B-atom swizzle)
  int T, D, heads;
  int qo, ko, vo, hs;    // channel offsets of q / k / v inside a token row, channel stride of a head
  int n_primal;          // forward kernel: batch rows; tangent kernel: tangent row r is batch row n_primal + r
  float scale;
  int debug;
  float* S;              // probabilities [rows][heads][T][T] (forward: written; tangent / VJP: P0 = row 0)
  float* gS;             // VJP scratch [K][heads][T][T]
  float* o;              // forward / tangent: o rows; VJP: gqkv rows
  long long o_sN, o_sT;  // batch-row and token strides of `o` (floats)
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// UMMA shared-memory descriptors (cute::UMMA::SmemDescriptor bit layout, version 1, SWIZZLE_128B).
// K-major: rows of 128 B, 8-row groups 1024 B apart (SBO).
__device__ __forceinline__ uint64_t desc_k(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// MN-major tf32 operands have ONE legal shared-memory layout, SWIZZLE_128B with a 32-byte base
// (descriptor layout type 1; TMA: CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): an atom is 4 K-rows x 128 B
// (32 fp32 along M / N), the 32-byte chunks of a row XOR-ed with (row % 4).  LBO = pitch between
// 32-element groups along M / N, SBO = pitch between 4-row groups along K (512 B for 128-byte rows).
__device__ __forceinline__ uint64_t desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)1 << 61;
  return d;
}
// instruction descriptor: tf32 x tf32 -> fp32, optional MN-major A (bit 15) / B (bit 16)
__device__ __forceinline__ uint32_t idesc_tf32(int m, int n, int a_mn, int b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;
  d |= 2u << 7;
  d |= 2u << 10;
  d |= (uint32_t)(a_mn ? 1 : 0) << 15;
  d |= (uint32_t)(b_mn ? 1 : 0) << 16;
  d |= (uint32_t)(n >> 3) << 17;
  d |= (uint32_t)(m >> 4) << 24;
  return d;
}

struct Smem {
  uint8_t* P;          // probability slabs
  uint8_t* stage;      // operand ring
  uint64_t* full;      // [2] TMA -> MMA
  uint64_t* empty;     // [2] MMA -> TMA
  uint64_t* s_ready;   // MMA -> row warps: score tile complete in TMEM
  uint64_t* p_ready;   // row warps -> MMA: probability slabs written (count 128)
  uint64_t* o_ready;   // MMA -> row warps: output tile complete (phase flips per output tile)
  uint64_t* t_free;    // row warps -> MMA: TMEM drained (count 128)
  uint64_t* p0_full;   // TMA -> row warps / MMA: P0 (or gS) slabs landed
  uint64_t* aux;       // MMA -> TMA: products reading the probability slabs are complete
  uint32_t* tmem_slot;
};

__device__ __forceinline__ Smem carve(uint8_t* raw) {
  Smem s;
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~(uintptr_t)1023);
  s.P = base;
  s.stage = base + kTcPBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s.stage + kTcStages * kTcStageBytes);
  s.full = bars; s.empty = bars + 2; s.s_ready = bars + 4; s.p_ready = bars + 5; s.o_ready = bars + 6;
  s.t_free = bars + 7; s.p0_full = bars + 8; s.aux = bars + 9;
  s.tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  return s;
}

__device__ __forceinline__ uint32_t setup(const Smem& sm, int warp) {
  if (threadIdx.x == 0) {
    mbar_init(&sm.full[0], 1); mbar_init(&sm.full[1], 1);
    mbar_init(&sm.empty[0], 1); mbar_init(&sm.empty[1], 1);
    mbar_init(sm.s_ready, 1); mbar_init(sm.p_ready, 128); mbar_init(sm.o_ready, 1);
    mbar_init(sm.t_free, 128); mbar_init(sm.p0_full, 1); mbar_init(sm.aux, 1);
    fence_mbar_init();
  }
  if (warp == 5) { tmem_alloc(sm.tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *sm.tmem_slot;
}
__device__ __forceinline__ void teardown(uint32_t tmem, int warp) {
  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem, 512);
}

// ---- producer / MMA building blocks (one elected thread each; `it` counts ring stages) ----
// K-major product: D[128, nB] (+)= A[128 rows, Dk] * B[nB rows, Dk]^T over Dk / 32 channel slabs.
__device__ __forceinline__ void produce_k(const Smem& sm, int& it, const CUtensorMap* ma, int ca, int ra, int na,
                                          const CUtensorMap* mb, int cb, int nb, int rowsB, int nslab) {
  for (int c = 0; c < nslab; ++c, ++it) {
    const int s = it & 1;
    mbar_wait(&sm.empty[s], ((it >> 1) & 1) ^ 1);
    uint8_t* st = sm.stage + s * kTcStageBytes;
    mbar_arrive_expect_tx(&sm.full[s], kTcSlab * (rowsB > 128 ? 3 : 2));
    tma_load_3d(st, ma, &sm.full[s], ca + c * 32, ra, na);
    tma_load_3d(st + kTcSlab, mb, &sm.full[s], cb + c * 32, 0, nb);
    if (rowsB > 128) tma_load_3d(st + 2 * kTcSlab, mb, &sm.full[s], cb + c * 32, 128, nb);
  }
}
__device__ __forceinline__ void mma_k(const Smem& sm, int& it, uint32_t tmem, int rowsB, int nslab, bool first_acc) {
  const uint32_t idesc = idesc_tf32(128, rowsB, 0, 0);
  for (int c = 0; c < nslab; ++c, ++it) {
    const int s = it & 1;
    mbar_wait(&sm.full[s], (it >> 1) & 1);
    tc_fence_after();
    const uint32_t a = smem_u32(sm.stage + s * kTcStageBytes), b = a + kTcSlab;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      umma_tf32(tmem, desc_k(a + k * 32), desc_k(b + k * 32), idesc, (first_acc || c > 0 || k > 0) ? 1u : 0u);
    umma_commit(&sm.empty[s]);
  }
}
// MN-major second factor: chunks of 16 tokens x D channels ([slab][16 tokens][128 B]) of map `m`,
// channel slab cs0, tokens [t0, t0 + 16 nchunk), batch row n.
__device__ __forceinline__ void produce_mn(const Smem& sm, int& it, const CUtensorMap* m, int cs0, int t0, int n,
                                           int nchunk, int D) {
  for (int j = 0; j < nchunk; ++j, ++it) {
    const int s = it & 1;
    mbar_wait(&sm.empty[s], ((it >> 1) & 1) ^ 1);
    mbar_arrive_expect_tx(&sm.full[s], 16 * D * 4);
    tma_load_4d(sm.stage + s * kTcStageBytes, m, &sm.full[s], 0, t0 + j * 16, cs0, n);
  }
}
// D[128, D] (+)= A * B with B = the MN-major chunks above (K = 16 tokens per chunk) and A either
//   a_mn = 0: K-major probability slabs (rows = queries, K = keys [16 j, 16 j + 16) of the chunk), or
//   a_mn = 1: the same slabs read MN-major (M = 128 keys starting at slab m_slab0, K = token rows
//             [16 j, 16 j + 16) of the slabs).
__device__ __forceinline__ void mma_mn(const Smem& sm, int& it, uint32_t tmem, int nchunk, int D, int a_mn,
                                       int m_slab0, bool first_acc, int debug = 0) {
  const uint32_t pbase = smem_u32(sm.P);
  for (int j = 0; j < nchunk; ++j, ++it) {
    const int s = it & 1;
    mbar_wait(&sm.full[s], (it >> 1) & 1);
    tc_fence_after();
    const uint32_t b0 = smem_u32(sm.stage + s * kTcStageBytes);
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) {
      const int krow = j * 16 + k2 * 8;       // first of the 8 contraction rows of this instruction
      uint64_t adesc;
      if (a_mn) adesc = desc_mn(pbase + m_slab0 * kTcSlab + krow * 128, kTcSlab, 512);
      else adesc = desc_k(pbase + (krow >> 5) * kTcSlab + (krow & 31) * 4);
      for (int h = 0; h * 256 < D; ++h) {
        const int nh = (D - h * 256) < 256 ? (D - h * 256) : 256;
        uint64_t bdesc = desc_mn(b0 + h * (8 * 2048) + k2 * 1024, 2048, 512);
        int b_mn = 1;
        if (debug == 1) { bdesc = desc_k(b0 + h * (8 * 2048) + k2 * 1024); b_mn = 0; }
        umma_tf32(tmem + h * 256, adesc, bdesc, idesc_tf32(128, nh, a_mn, b_mn),
                  (first_acc || j > 0 || k2 > 0) ? 1u : 0u);
      }
    }
    umma_commit(&sm.empty[s]);
  }
}
// load `nsl` probability slabs (32 keys x 128 rows each) of matrix `mat` of map `m`: rows [r0, r0 + 128),
// key slabs [ks0, ks0 + nsl)
__device__ __forceinline__ void load_slabs(const Smem& sm, const CUtensorMap* m, int ks0, int nsl, int r0, int mat) {
  mbar_arrive_expect_tx(sm.p0_full, nsl * kTcSlab);
  for (int s = 0; s < nsl; ++s) tma_load_3d(sm.P + s * kTcSlab, m, sm.p0_full, (ks0 + s) * 32, r0, mat);
}

// ---- row-warp helpers (thread = TMEM lane = row of the tile) ----
// byte offset of the 16-byte group c4 (4 keys) of `row` inside a 128B-swizzled slab
__device__ __forceinline__ int swz(int row, int c4) { return row * 128 + ((c4 ^ (row & 7)) << 4); }

// out[row][0..D) = round_tf32(TMEM row), 32 columns at a time
__device__ __forceinline__ void store_rows(uint32_t trow, int D, float* dst, bool valid) {
  for (int ch = 0; ch < D / 32; ++ch) {
    uint32_t r[32];
    tmem_ld_32x32(trow + ch * 32, r);
    tmem_ld_wait();
    if (valid) {
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4)
        *reinterpret_cast<float4*>(dst + ch * 32 + c4 * 4) =
            make_float4(round_tf32(__uint_as_float(r[4 * c4])), round_tf32(__uint_as_float(r[4 * c4 + 1])),
                        round_tf32(__uint_as_float(r[4 * c4 + 2])), round_tf32(__uint_as_float(r[4 * c4 + 3])));
    }
  }
}

// X (TMEM row, T columns) -> scale * P0 o (X - sum_j P0 X) written over P0 in the slabs (tf32-rounded),
// optionally also to a global row
__device__ __forceinline__ void linearise_row(const Smem& sm, uint32_t trow, int T, int row, float scale,
                                              float* grow) {
  float dot = 0.f;
  for (int ch = 0; ch < T / 32; ++ch) {
    uint32_t r[32];
    tmem_ld_32x32(trow + ch * 32, r);
    tmem_ld_wait();
    const uint8_t* slab = sm.P + ch * kTcSlab;
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const float4 p = *reinterpret_cast<const float4*>(slab + swz(row, c4));
      dot += p.x * __uint_as_float(r[4 * c4]) + p.y * __uint_as_float(r[4 * c4 + 1]) +
             p.z * __uint_as_float(r[4 * c4 + 2]) + p.w * __uint_as_float(r[4 * c4 + 3]);
    }
  }
  for (int ch = 0; ch < T / 32; ++ch) {
    uint32_t r[32];
    tmem_ld_32x32(trow + ch * 32, r);
    tmem_ld_wait();
    uint8_t* slab = sm.P + ch * kTcSlab;
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      float4* pp = reinterpret_cast<float4*>(slab + swz(row, c4));
      const float4 p = *pp;
      float4 v;
      v.x = round_tf32(scale * p.x * (__uint_as_float(r[4 * c4]) - dot));
      v.y = round_tf32(scale * p.y * (__uint_as_float(r[4 * c4 + 1]) - dot));
      v.z = round_tf32(scale * p.z * (__uint_as_float(r[4 * c4 + 2]) - dot));
      v.w = round_tf32(scale * p.w * (__uint_as_float(r[4 * c4 + 3]) - dot));
      *pp = v;
      if (grow) *reinterpret_cast<float4*>(grow + ch * 32 + c4 * 4) = v;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// primal rows: grid (query blocks, heads, rows)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const Smem sm = carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, hd = blockIdx.y, n = blockIdx.z;
  const int T = p.T, D = p.D;
  const uint32_t tmem = setup(sm, warp);
  const int cq = hd * p.hs + p.qo, ck = hd * p.hs + p.ko, cv = hd * p.hs + p.vo;
  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      produce_k(sm, it, &p.qkv_k, cq, q0, n, &p.qkv_k, ck, n, T, D / 32);
      produce_mn(sm, it, &p.qkv_mn, cv / 32, 0, n, T / 16, D);
    }
  } else if (warp == 5) {
    if (lane == 0) {
      int it = 0;
      mma_k(sm, it, tmem, T, D / 32, false);
      umma_commit(sm.s_ready);
      mbar_wait(sm.p_ready, 0);
      tc_fence_after();
      mma_mn(sm, it, tmem, T / 16, D, 0, 0, false, p.debug);
      umma_commit(sm.o_ready);
    }
  } else {
    const int row = threadIdx.x;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const bool valid = q0 + row < T;
    mbar_wait(sm.s_ready, 0);
    tc_fence_after();
    const float sc = p.scale * kLog2e;
    float mx = -INFINITY;
    for (int ch = 0; ch < T / 32; ++ch) {
      uint32_t r[32];
      tmem_ld_32x32(trow + ch * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(r[i]));
    }
    float sum = 0.f;
    for (int ch = 0; ch < T / 32; ++ch) {
      uint32_t r[32];
      tmem_ld_32x32(trow + ch * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) sum += exp2f((__uint_as_float(r[i]) - mx) * sc);
    }
    const float inv = 1.0f / sum;
    float* Srow = p.S + (((long long)n * p.heads + hd) * T + q0 + row) * T;
    for (int ch = 0; ch < T / 32; ++ch) {
      uint32_t r[32];
      tmem_ld_32x32(trow + ch * 32, r);
      tmem_ld_wait();
      uint8_t* slab = sm.P + ch * kTcSlab;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        float4 v;
        v.x = round_tf32(exp2f((__uint_as_float(r[4 * c4]) - mx) * sc) * inv);
        v.y = round_tf32(exp2f((__uint_as_float(r[4 * c4 + 1]) - mx) * sc) * inv);
        v.z = round_tf32(exp2f((__uint_as_float(r[4 * c4 + 2]) - mx) * sc) * inv);
        v.w = round_tf32(exp2f((__uint_as_float(r[4 * c4 + 3]) - mx) * sc) * inv);
        *reinterpret_cast<float4*>(slab + swz(row, c4)) = v;
        if (valid) *reinterpret_cast<float4*>(Srow + ch * 32 + c4 * 4) = v;
      }
    }
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(sm.p_ready);
    mbar_wait(sm.o_ready, 0);
    tc_fence_after();
    store_rows(trow, D, p.o + (long long)n * p.o_sN + (long long)(q0 + row) * p.o_sT + hd * D, valid);
  }
  teardown(tmem, warp);
}

// ------------------------------------------------------------------------------------------------
// tangent rows: grid (query blocks, heads, tangent rows); primal = batch row 0
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
attn_jvp_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const Smem sm = carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, hd = blockIdx.y, n = p.n_primal + blockIdx.z;
  const int T = p.T, D = p.D;
  const uint32_t tmem = setup(sm, warp);
  const int cq = hd * p.hs + p.qo, ck = hd * p.hs + p.ko, cv = hd * p.hs + p.vo;
  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      load_slabs(sm, &p.p_map, 0, T / 32, q0, hd);                               // P0 (primal row 0)
      produce_k(sm, it, &p.qkv_k, cq, q0, n, &p.qkv_k, ck, 0, T, D / 32);        // dq k0^T
      produce_k(sm, it, &p.qkv_k, cq, q0, 0, &p.qkv_k, ck, n, T, D / 32);        // q0 dk^T
      produce_mn(sm, it, &p.qkv_mn, cv / 32, 0, 0, T / 16, D);                   // dP v0
      mbar_wait(sm.aux, 0);                                                      // dP consumed
      load_slabs(sm, &p.p_map, 0, T / 32, q0, hd);                               // P0 again
      produce_mn(sm, it, &p.qkv_mn, cv / 32, 0, n, T / 16, D);                   // P0 dv
    }
  } else if (warp == 5) {
    if (lane == 0) {
      int it = 0;
      mma_k(sm, it, tmem, T, D / 32, false);
      mma_k(sm, it, tmem, T, D / 32, true);
      umma_commit(sm.s_ready);
      mbar_wait(sm.p_ready, 0);
      tc_fence_after();
      mma_mn(sm, it, tmem, T / 16, D, 0, 0, false);
      umma_commit(sm.aux);
      mbar_wait(sm.p0_full, 1);
      tc_fence_after();
      mma_mn(sm, it, tmem, T / 16, D, 0, 0, true);
      umma_commit(sm.o_ready);
    }
  } else {
    const int row = threadIdx.x;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const bool valid = q0 + row < T;
    mbar_wait(sm.p0_full, 0);
    mbar_wait(sm.s_ready, 0);
    tc_fence_after();
    linearise_row(sm, trow, T, row, p.scale, nullptr);
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(sm.p_ready);
    mbar_wait(sm.o_ready, 0);
    tc_fence_after();
    store_rows(trow, D, p.o + (long long)n * p.o_sN + (long long)(q0 + row) * p.o_sT + hd * D, valid);
  }
  teardown(tmem, warp);
}

// ------------------------------------------------------------------------------------------------
// cotangent rows, kernel 1: grid (query blocks, heads, K).  gS -> scratch, gq -> gqkv
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
attn_vjp1_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const Smem sm = carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, hd = blockIdx.y, r = blockIdx.z;
  const int T = p.T, D = p.D;
  const uint32_t tmem = setup(sm, warp);
  const int cq = hd * p.hs + p.qo, ck = hd * p.hs + p.ko, cv = hd * p.hs + p.vo;
  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      load_slabs(sm, &p.p_map, 0, T / 32, q0, hd);
      produce_k(sm, it, &p.go_k, hd * D, q0, r, &p.qkv_k, cv, 0, T, D / 32);     // gP = go v0^T
      produce_mn(sm, it, &p.qkv_mn, ck / 32, 0, 0, T / 16, D);                   // gq = gS k0
    }
  } else if (warp == 5) {
    if (lane == 0) {
      int it = 0;
      mma_k(sm, it, tmem, T, D / 32, false);
      umma_commit(sm.s_ready);
      mbar_wait(sm.p_ready, 0);
      tc_fence_after();
      mma_mn(sm, it, tmem, T / 16, D, 0, 0, false);
      umma_commit(sm.o_ready);
    }
  } else {
    const int row = threadIdx.x;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const bool valid = q0 + row < T;
    mbar_wait(sm.p0_full, 0);
    mbar_wait(sm.s_ready, 0);
    tc_fence_after();
    float* grow = valid ? p.gS + (((long long)r * p.heads + hd) * T + q0 + row) * T : nullptr;
    linearise_row(sm, trow, T, row, p.scale, grow);
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(sm.p_ready);
    mbar_wait(sm.o_ready, 0);
    tc_fence_after();
    store_rows(trow, D, p.o + (long long)r * p.o_sN + (long long)(q0 + row) * p.o_sT + cq, valid);
  }
  teardown(tmem, warp);
}

// ------------------------------------------------------------------------------------------------
// cotangent rows, kernel 2: grid (key blocks, heads, K).  gv = P0^T go, gk = gS^T q0 for 128 keys,
// the contraction over all T queries in blocks of 128 (slabs hold [128 queries][32 keys] x 4).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
attn_vjp2_tc_kernel(const __grid_constant__ AttnTcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const Smem sm = carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int k0 = blockIdx.x * 128, hd = blockIdx.y, r = blockIdx.z;
  const int T = p.T, D = p.D;
  const uint32_t tmem = setup(sm, warp);
  const int cq = hd * p.hs + p.qo, ck = hd * p.hs + p.ko, cv = hd * p.hs + p.vo;
  const int nqb = (T + 127) / 128;             // query blocks of the contraction
  const int nks = (T - k0) < 128 ? (T - k0) / 32 : 4;   // key slabs of this block that exist
  const int qrows = T < 128 ? T : 128;         // token rows per query block
  // the probability buffer holds one query block at a time: 4 slabs [128 queries][32 keys]
  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      for (int pass = 0; pass < 2; ++pass) {   // 0: gv (P0, go), 1: gk (gS, q0)
        for (int qb = 0; qb < nqb; ++qb) {
          const int u = pass * nqb + qb;       // use index of the probability buffer
          if (u > 0) mbar_wait(sm.aux, (u - 1) & 1);   // previous user's products complete
          if (pass == 0) load_slabs(sm, &p.p_map_mn, k0 / 32, nks, qb * 128, hd);
          else load_slabs(sm, &p.gs_map, k0 / 32, nks, qb * 128, r * p.heads + hd);
          if (pass == 0) produce_mn(sm, it, &p.go_mn, (hd * D) / 32, qb * 128, r, qrows / 16, D);
          else produce_mn(sm, it, &p.qkv_mn, cq / 32, qb * 128, 0, qrows / 16, D);
        }
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      int it = 0;
      for (int pass = 0; pass < 2; ++pass) {
        if (pass == 1) { mbar_wait(sm.t_free, 0); tc_fence_after(); }   // gv drained from TMEM
        for (int qb = 0; qb < nqb; ++qb) {
          const int u = pass * nqb + qb;
          mbar_wait(sm.p0_full, u & 1);
          tc_fence_after();
          mma_mn(sm, it, tmem, qrows / 16, D, 1, 0, qb > 0);
          umma_commit(sm.aux);
        }
        umma_commit(sm.o_ready);
      }
    }
  } else {
    const int row = threadIdx.x;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    const bool valid = k0 + row < T;
    float* base = p.o + (long long)r * p.o_sN + (long long)(k0 + row) * p.o_sT;
    mbar_wait(sm.o_ready, 0);
    tc_fence_after();
    store_rows(trow, D, base + cv, valid);
    tc_fence_before();
    mbar_arrive(sm.t_free);
    mbar_wait(sm.o_ready, 1);
    tc_fence_after();
    store_rows(trow, D, base + ck, valid);
  }
  teardown(tmem, warp);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) {
      set_error("cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
      return nullptr;
    }
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}
// K-major rows of a [N][T][C] fp32 tensor (token pitch sT, row pitch sN floats): dims (C, T, N), box (32, 128, 1)
int encode_rows_k(CUtensorMap* m, const float* base, int C, int T, int N, long long sT, long long sN,
                  bool mn_swizzle = false) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return 3;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)sT * 4, (cuuint64_t)sN * 4};
  cuuint32_t box[3] = {32, 128, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  mn_swizzle ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LOCO_REQUIRE(r == CUDA_SUCCESS, "attention: cuTensorMapEncodeTiled(rows) failed: %d (C=%d T=%d N=%d)", (int)r, C, T, N);
  return 0;
}
// MN-major chunks of the same tensor: dims (32, T, C / 32, N), box (32, 16, D / 32, 1)
int encode_rows_mn(CUtensorMap* m, const float* base, int C, int T, int N, long long sT, long long sN, int D) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return 3;
  cuuint64_t dims[4] = {32, (cuuint64_t)T, (cuuint64_t)(C / 32), (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)sT * 4, 128, (cuuint64_t)sN * 4};
  cuuint32_t box[4] = {32, 16, (cuuint32_t)(D / 32), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  LOCO_REQUIRE(r == CUDA_SUCCESS, "attention: cuTensorMapEncodeTiled(chunks) failed: %d (C=%d T=%d N=%d D=%d)", (int)r,
               C, T, N, D);
  return 0;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace

bool attention_tc_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("LOCO_ATTN_TC");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

bool attention_tc_eligible(int T, int C, int head_ch) {
  const int D = head_ch > 0 ? head_ch : C;
  return attention_tc_enabled() && T % 32 == 0 && T >= 32 && T <= 256 && D % 32 == 0 && D >= 32 && D <= 512 &&
         C % 32 == 0;
}

int attention_tc_init() {
  static bool done[kMaxDevices] = {false};
  if (!first_time_on_device(done)) return 0;
  LOCO_CHECK_CUDA(cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
  LOCO_CHECK_CUDA(cudaFuncSetAttribute(attn_jvp_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
  LOCO_CHECK_CUDA(cudaFuncSetAttribute(attn_vjp1_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
  LOCO_CHECK_CUDA(cudaFuncSetAttribute(attn_vjp2_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmem));
  return 0;
}

namespace {
void fill_geom(AttnTcParams& P, int T, int C, int head_ch) {
  P.T = T;
  if (head_ch <= 0) { P.heads = 1; P.D = C; P.qo = 0; P.ko = C; P.vo = 2 * C; P.hs = 0; }
  else { P.heads = C / head_ch; P.D = head_ch; P.qo = 0; P.ko = head_ch; P.vo = 2 * head_ch; P.hs = 3 * head_ch; }
  P.scale = 1.0f / sqrtf((float)P.D);
}
}  // namespace

int attention_forward_tc(View qkv, int n_primal, int head_ch, float* S, View o, cudaStream_t s) {
  const int C = qkv.C / 3, T = qkv.H * qkv.W, N = qkv.N, nt = N - n_primal;
  LOCO_REQUIRE(!qkv.half && !o.half, "attention (tcgen05): fp32 tensors expected");
  LOCO_REQUIRE(aligned16(qkv.ptr) && aligned16(o.ptr) && aligned16(S) && qkv.sW % 4 == 0 && qkv.sN % 4 == 0 &&
                   o.sW % 4 == 0 && o.sN % 4 == 0,
               "attention (tcgen05): tensors must be 16-byte aligned");
  LOCO_TRY(attention_tc_init());
  AttnTcParams P;
  memset(&P, 0, sizeof(P));
  fill_geom(P, T, C, head_ch);
  P.n_primal = n_primal;
  { const char* e = getenv("LOCO_ATTN_DEBUG"); P.debug = e ? atoi(e) : 0; }
  P.S = S; P.o = o.ptr; P.o_sN = o.sN; P.o_sT = o.sW;
  LOCO_TRY(encode_rows_k(&P.qkv_k, qkv.ptr, qkv.C, T, N, qkv.sW, qkv.sN));
  LOCO_TRY(encode_rows_mn(&P.qkv_mn, qkv.ptr, qkv.C, T, N, qkv.sW, qkv.sN, P.D));
  LOCO_TRY(encode_rows_k(&P.p_map, S, T, T, P.heads * N, T, (long long)T * T));
  const int qblocks = (T + 127) / 128;
  attn_fwd_tc_kernel<<<dim3(qblocks, P.heads, n_primal), kTcThreads, kTcSmem, s>>>(P);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  if (nt > 0) {
    attn_jvp_tc_kernel<<<dim3(qblocks, P.heads, nt), kTcThreads, kTcSmem, s>>>(P);
    count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

int attention_vjp_tc(View go, View qkv0, int head_ch, const float* P0, float* gP, View gqkv, cudaStream_t s) {
  const int C = qkv0.C / 3, T = qkv0.H * qkv0.W, K = go.N;
  LOCO_REQUIRE(!go.half && !qkv0.half && !gqkv.half, "attention (tcgen05): fp32 tensors expected");
  LOCO_REQUIRE(aligned16(go.ptr) && aligned16(qkv0.ptr) && aligned16(gqkv.ptr) && aligned16(P0) && aligned16(gP) &&
                   go.sW % 4 == 0 && go.sN % 4 == 0 && qkv0.sW % 4 == 0 && gqkv.sW % 4 == 0 && gqkv.sN % 4 == 0,
               "attention (tcgen05): tensors must be 16-byte aligned");
  LOCO_TRY(attention_tc_init());
  AttnTcParams P;
  memset(&P, 0, sizeof(P));
  fill_geom(P, T, C, head_ch);
  P.S = const_cast<float*>(P0); P.gS = gP; P.o = gqkv.ptr; P.o_sN = gqkv.sN; P.o_sT = gqkv.sW;
  LOCO_TRY(encode_rows_k(&P.qkv_k, qkv0.ptr, qkv0.C, T, 1, qkv0.sW, qkv0.sN));
  LOCO_TRY(encode_rows_mn(&P.qkv_mn, qkv0.ptr, qkv0.C, T, 1, qkv0.sW, qkv0.sN, P.D));
  LOCO_TRY(encode_rows_k(&P.p_map, P0, T, T, P.heads, T, (long long)T * T));
  LOCO_TRY(encode_rows_k(&P.go_k, go.ptr, go.C, T, K, go.sW, go.sN));
  LOCO_TRY(encode_rows_mn(&P.go_mn, go.ptr, go.C, T, K, go.sW, go.sN, P.D));
  LOCO_TRY(encode_rows_k(&P.gs_map, gP, T, T, P.heads * K, T, (long long)T * T, true));
  LOCO_TRY(encode_rows_k(&P.p_map_mn, P0, T, T, P.heads, T, (long long)T * T, true));
  const int qblocks = (T + 127) / 128;
  attn_vjp1_tc_kernel<<<dim3(qblocks, P.heads, K), kTcThreads, kTcSmem, s>>>(P);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  attn_vjp2_tc_kernel<<<dim3(qblocks, P.heads, K), kTcThreads, kTcSmem, s>>>(P);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace loco

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
#include <stdio.h>
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
  CUtensorMap o_st;      // store map of the output rows (o, or gqkv in the VJP): dims (C_out, T, rows), box (32, 32, 1)
  CUtensorMap p_st;      // store map of the probabilities (forward: S; VJP kernel 1: gS scratch), box (32, 32, 1)
  CUtensorMap kv_k;      // key / value source, K-major boxes (self-attention: = qkv_k; cross-attention: the context projections)
  CUtensorMap kv_mn;     // key / value source, MN-major chunks
  int T, D, heads;       // T = query tokens
  int Tk, Tk_valid;      // key tokens (padded to a multiple of 64) and how many of them exist
  int cross;             // cross-attention: keys / values are constants (no dk, dv terms; VJP returns gq only)
  int qo, ko, vo, hs;    // channel offsets of q / k / v inside a token row, channel stride of a head
  int n_primal;          // forward kernel: batch rows; tangent kernel: tangent row r is batch row n_primal + r
  float scale;
  int debug;
  unsigned long long* stamps;   // LOCO_ATTN_DEBUG=9: %globaltimer stamps of CTA 0 (profiling aid)
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define LOCO_STAMP(i) do { if (p.stamps && threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) p.stamps[i] = gtime(); } while (0)

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                            int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// TMA stores (shared -> global, bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 32 lanes x 64 consecutive fp32 columns -> 64 registers per thread: a row warp spends its time on
// TMEM round trips, so it fetches two probability slabs per trip
__device__ __forceinline__ void tmem_ld_32x32_x64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]), "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]), "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]), "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]), "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63])
      : "r"(taddr)
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

// The 32 rows of this warp of a [128 x D] TMEM tile -> global through `map` (channel c0, row r0, matrix n):
// 64 columns per TMEM round trip, tf32-rounded, written into the warp's 32-row window of staging slabs
// [slab0, slab0 + nslab) of the probability buffer (128B-swizzled like any other box) and pushed out
// by 32 x 32 TMA stores (per-thread scattered stores of 2 KB rows cost 11 of the kernel's 33 us).
__device__ __forceinline__ void store_rows(const Smem& sm, uint32_t trow, int D, const CUtensorMap* map, int c0,
                                           int r0, int n, int warp, int lane, int slab0, int nslab) {
  const int row = warp * 32 + lane;
  const int ngroups = nslab / 2;                 // bulk groups (of two slabs) that may be in flight
  for (int c = 0; c < D / 64; ++c) {
    uint32_t r[64];
    tmem_ld_32x32_x64(trow + c * 64, r);
    if (c >= ngroups) {                          // the slabs about to be overwritten have been read
      if (lane == 0) { if (ngroups >= 4) tma_store_wait_read<3>(); else tma_store_wait_read<1>(); }
      __syncwarp();
    }
    tmem_ld_wait();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint8_t* slab = sm.P + (slab0 + (2 * c + h) % nslab) * kTcSlab;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4)
        *reinterpret_cast<float4*>(slab + swz(row, c4)) =
            make_float4(round_tf32(__uint_as_float(r[32 * h + 4 * c4])), round_tf32(__uint_as_float(r[32 * h + 4 * c4 + 1])),
                        round_tf32(__uint_as_float(r[32 * h + 4 * c4 + 2])), round_tf32(__uint_as_float(r[32 * h + 4 * c4 + 3])));
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
        tma_store_3d(map, sm.P + (slab0 + (2 * c + h) % nslab) * kTcSlab + warp * 4096, c0 + (2 * c + h) * 32, r0 + warp * 32, n);
      tma_store_commit();
    }
  }
  if (lane == 0) tma_store_wait_all();
}
// the warp's 32 rows of the first T / 32 probability slabs -> matrix `mat` of `map` (rows r0 + 32 warp ..)
__device__ __forceinline__ void store_slabs(const Smem& sm, const CUtensorMap* map, int T, int r0, int mat, int warp,
                                            int lane) {
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    for (int sl = 0; sl < T / 32; ++sl)
      tma_store_3d(map, sm.P + sl * kTcSlab + warp * 4096, sl * 32, r0 + warp * 32, mat);
    tma_store_commit();
  }
}

// X (TMEM row, T columns) -> scale * P0 o (X - sum_j P0 X) written over P0 in the slabs (tf32-rounded)
__device__ __forceinline__ void linearise_row(const Smem& sm, uint32_t trow, int T, int row, float scale) {
  float dot = 0.f;
  for (int c = 0; c < T / 64; ++c) {
    uint32_t r[64];
    tmem_ld_32x32_x64(trow + c * 64, r);
    tmem_ld_wait();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const uint8_t* slab = sm.P + (2 * c + h) * kTcSlab;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 p = *reinterpret_cast<const float4*>(slab + swz(row, c4));
        dot += p.x * __uint_as_float(r[32 * h + 4 * c4]) + p.y * __uint_as_float(r[32 * h + 4 * c4 + 1]) +
               p.z * __uint_as_float(r[32 * h + 4 * c4 + 2]) + p.w * __uint_as_float(r[32 * h + 4 * c4 + 3]);
      }
    }
  }
  for (int c = 0; c < T / 64; ++c) {
    uint32_t r[64];
    tmem_ld_32x32_x64(trow + c * 64, r);
    tmem_ld_wait();
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      uint8_t* slab = sm.P + (2 * c + h) * kTcSlab;
      float4 pv[8];
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) pv[c4] = *reinterpret_cast<const float4*>(slab + swz(row, c4));
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 p = pv[c4];
        float4 v;
        v.x = round_tf32(scale * p.x * (__uint_as_float(r[32 * h + 4 * c4]) - dot));
        v.y = round_tf32(scale * p.y * (__uint_as_float(r[32 * h + 4 * c4 + 1]) - dot));
        v.z = round_tf32(scale * p.z * (__uint_as_float(r[32 * h + 4 * c4 + 2]) - dot));
        v.w = round_tf32(scale * p.w * (__uint_as_float(r[32 * h + 4 * c4 + 3]) - dot));
        *reinterpret_cast<float4*>(slab + swz(row, c4)) = v;
      }
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
  const int T = p.Tk, D = p.D;                 // everything below is indexed by keys
  const int nkv = p.cross ? 0 : n;             // batch row of the key / value source
  LOCO_STAMP(0);
  const uint32_t tmem = setup(sm, warp);
  LOCO_STAMP(1);
  const int cq = hd * p.hs + p.qo, ck = hd * p.hs + p.ko, cv = hd * p.hs + p.vo;
  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      produce_k(sm, it, &p.qkv_k, cq, q0, n, &p.kv_k, ck, nkv, T, D / 32);
      produce_mn(sm, it, &p.kv_mn, cv / 32, 0, nkv, T / 16, D);
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
    mbar_wait(sm.s_ready, 0);
    tc_fence_after();
    LOCO_STAMP(2);
    const float sc = p.scale * kLog2e;
    float mx = -INFINITY;
    for (int c = 0; c < T / 64; ++c) {
      uint32_t r[64];
      tmem_ld_32x32_x64(trow + c * 64, r);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 64; ++i)
        if (c * 64 + i < p.Tk_valid) mx = fmaxf(mx, __uint_as_float(r[i]));     // padded keys do not exist
    }
    LOCO_STAMP(7);
    // unnormalised exponentials go to the slabs, the row is rescaled in place once its sum is known
    float sum = 0.f;
    const float off = mx * sc;
    for (int c = 0; c < T / 64; ++c) {
      uint32_t r[64];
      tmem_ld_32x32_x64(trow + c * 64, r);
      tmem_ld_wait();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint8_t* slab = sm.P + (2 * c + h) * kTcSlab;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          float4 v;
          v.x = ex2_approx(fmaf(__uint_as_float(r[32 * h + 4 * c4]), sc, -off));
          v.y = ex2_approx(fmaf(__uint_as_float(r[32 * h + 4 * c4 + 1]), sc, -off));
          v.z = ex2_approx(fmaf(__uint_as_float(r[32 * h + 4 * c4 + 2]), sc, -off));
          v.w = ex2_approx(fmaf(__uint_as_float(r[32 * h + 4 * c4 + 3]), sc, -off));
          const int col = c * 64 + 32 * h + 4 * c4;
          if (col + 3 >= p.Tk_valid) {
            if (col >= p.Tk_valid) v.x = 0.f;
            if (col + 1 >= p.Tk_valid) v.y = 0.f;
            if (col + 2 >= p.Tk_valid) v.z = 0.f;
            v.w = 0.f;
          }
          sum += (v.x + v.y) + (v.z + v.w);
          *reinterpret_cast<float4*>(slab + swz(row, c4)) = v;
        }
      }
    }
    const float inv = 1.0f / sum;
    LOCO_STAMP(8);
    for (int sl = 0; sl < T / 32; ++sl) {
      uint8_t* slab = sm.P + sl * kTcSlab;
      float4 v[8];                 // the loads of a slab row first: their latencies overlap
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) v[c4] = *reinterpret_cast<const float4*>(slab + swz(row, c4));
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        v[c4].x = round_tf32(v[c4].x * inv); v[c4].y = round_tf32(v[c4].y * inv);
        v[c4].z = round_tf32(v[c4].z * inv); v[c4].w = round_tf32(v[c4].w * inv);
        *reinterpret_cast<float4*>(slab + swz(row, c4)) = v[c4];
      }
    }
    LOCO_STAMP(9);
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(sm.p_ready);
    store_slabs(sm, &p.p_st, T, q0, n * p.heads + hd, warp, lane);     // P0 for the tangent / cotangent kernels
    LOCO_STAMP(3);
    mbar_wait(sm.o_ready, 0);
    tc_fence_after();
    LOCO_STAMP(4);
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
    store_rows(sm, trow, D, &p.o_st, hd * D, q0, n, warp, lane, 0, 8);
    LOCO_STAMP(5);
  }
  teardown(tmem, warp);
  LOCO_STAMP(6);
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
  const int T = p.Tk, D = p.D;
  const uint32_t tmem = setup(sm, warp);
  const int cq = hd * p.hs + p.qo, ck = hd * p.hs + p.ko, cv = hd * p.hs + p.vo;
  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      load_slabs(sm, &p.p_map, 0, T / 32, q0, hd);                               // P0 (primal row 0)
      produce_k(sm, it, &p.qkv_k, cq, q0, n, &p.kv_k, ck, 0, T, D / 32);         // dq k0^T
      if (!p.cross) produce_k(sm, it, &p.qkv_k, cq, q0, 0, &p.kv_k, ck, n, T, D / 32);   // q0 dk^T
      produce_mn(sm, it, &p.kv_mn, cv / 32, 0, 0, T / 16, D);                    // dP v0
      if (!p.cross) {
        mbar_wait(sm.aux, 0);                                                    // dP consumed
        load_slabs(sm, &p.p_map, 0, T / 32, q0, hd);                             // P0 again
        produce_mn(sm, it, &p.kv_mn, cv / 32, 0, n, T / 16, D);                  // P0 dv
      }
    }
  } else if (warp == 5) {
    if (lane == 0) {
      int it = 0;
      mma_k(sm, it, tmem, T, D / 32, false);
      if (!p.cross) mma_k(sm, it, tmem, T, D / 32, true);
      umma_commit(sm.s_ready);
      mbar_wait(sm.p_ready, 0);
      tc_fence_after();
      mma_mn(sm, it, tmem, T / 16, D, 0, 0, false);
      if (!p.cross) {
        umma_commit(sm.aux);
        mbar_wait(sm.p0_full, 1);
        tc_fence_after();
        mma_mn(sm, it, tmem, T / 16, D, 0, 0, true);
      }
      umma_commit(sm.o_ready);
    }
  } else {
    const int row = threadIdx.x;
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    mbar_wait(sm.p0_full, 0);
    mbar_wait(sm.s_ready, 0);
    tc_fence_after();
    linearise_row(sm, trow, T, row, p.scale);
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(sm.p_ready);
    mbar_wait(sm.o_ready, 0);
    tc_fence_after();
    store_rows(sm, trow, D, &p.o_st, hd * D, q0, n, warp, lane, 0, 8);
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
  const int T = p.Tk, D = p.D;
  const uint32_t tmem = setup(sm, warp);
  const int cq = hd * p.hs + p.qo, ck = hd * p.hs + p.ko, cv = hd * p.hs + p.vo;
  if (warp == 4) {
    if (lane == 0) {
      int it = 0;
      load_slabs(sm, &p.p_map, 0, T / 32, q0, hd);
      produce_k(sm, it, &p.go_k, hd * D, q0, r, &p.kv_k, cv, 0, T, D / 32);      // gP = go v0^T
      produce_mn(sm, it, &p.kv_mn, ck / 32, 0, 0, T / 16, D);                    // gq = gS k0
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
    mbar_wait(sm.p0_full, 0);
    mbar_wait(sm.s_ready, 0);
    tc_fence_after();
    linearise_row(sm, trow, T, row, p.scale);
    fence_proxy_async();
    tc_fence_before();
    mbar_arrive(sm.p_ready);
    if (!p.cross) store_slabs(sm, &p.p_st, T, q0, r * p.heads + hd, warp, lane);     // gS for kernel 2
    mbar_wait(sm.o_ready, 0);
    tc_fence_after();
    if (lane == 0) tma_store_wait_read<0>();
    __syncwarp();
    store_rows(sm, trow, D, &p.o_st, cq, q0, r, warp, lane, 0, 8);
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
    const uint32_t trow = tmem + ((uint32_t)(warp * 32) << 16);
    mbar_wait(sm.o_ready, 0);
    tc_fence_after();
    // staging in slabs 4..7: slabs 0..3 are being refilled for the second pass
    store_rows(sm, trow, D, &p.o_st, cv, k0, r, warp, lane, 4, 4);
    tc_fence_before();
    mbar_arrive(sm.t_free);
    mbar_wait(sm.o_ready, 1);
    tc_fence_after();
    store_rows(sm, trow, D, &p.o_st, ck, k0, r, warp, lane, 4, 4);
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
                  bool mn_swizzle = false, int box_rows = 128) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return 3;
  cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)T, (cuuint64_t)N};
  cuuint64_t strides[2] = {(cuuint64_t)sT * 4, (cuuint64_t)sN * 4};
  cuuint32_t box[3] = {32, (cuuint32_t)box_rows, 1};
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
  return attention_tc_enabled() && T % 64 == 0 && T >= 64 && T <= 256 && D % 64 == 0 && D >= 64 && D <= 512 &&
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
  LOCO_TRY(encode_rows_k(&P.qkv_k, qkv.ptr, qkv.C, T, N, qkv.sW, qkv.sN));
  LOCO_TRY(encode_rows_mn(&P.qkv_mn, qkv.ptr, qkv.C, T, N, qkv.sW, qkv.sN, P.D));
  P.kv_k = P.qkv_k; P.kv_mn = P.qkv_mn; P.Tk = T; P.Tk_valid = T;
  LOCO_TRY(encode_rows_k(&P.p_map, S, T, T, P.heads * N, T, (long long)T * T));
  LOCO_TRY(encode_rows_k(&P.p_st, S, T, T, P.heads * N, T, (long long)T * T, false, 32));
  LOCO_TRY(encode_rows_k(&P.o_st, o.ptr, o.C, T, N, o.sW, o.sN, false, 32));
  const int qblocks = (T + 127) / 128;
  static unsigned long long* stamps = nullptr;
  if (P.debug == 9) {
    if (!stamps) { LOCO_CHECK_CUDA(cudaMalloc(&stamps, 64 * sizeof(unsigned long long))); }
    P.stamps = stamps;
  }
  attn_fwd_tc_kernel<<<dim3(qblocks, P.heads, n_primal), kTcThreads, kTcSmem, s>>>(P);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  if (P.debug == 9) {
    unsigned long long h[10];
    LOCO_CHECK_CUDA(cudaStreamSynchronize(s));
    LOCO_CHECK_CUDA(cudaMemcpy(h, stamps, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "attn_fwd stamps (ns from entry): setup %llu s_ready %llu [max %llu exp %llu norm %llu] softmax %llu o_ready %llu stored %llu end %llu\n",
            h[1] - h[0], h[2] - h[0], h[7] - h[0], h[8] - h[0], h[9] - h[0], h[3] - h[0], h[4] - h[0], h[5] - h[0], h[6] - h[0]);
  }
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
  LOCO_TRY(encode_rows_k(&P.qkv_k, qkv0.ptr, qkv0.C, T, 1, qkv0.sW, qkv0.sN));
  LOCO_TRY(encode_rows_mn(&P.qkv_mn, qkv0.ptr, qkv0.C, T, 1, qkv0.sW, qkv0.sN, P.D));
  P.kv_k = P.qkv_k; P.kv_mn = P.qkv_mn; P.Tk = T; P.Tk_valid = T;
  LOCO_TRY(encode_rows_k(&P.p_map, P0, T, T, P.heads, T, (long long)T * T));
  LOCO_TRY(encode_rows_k(&P.go_k, go.ptr, go.C, T, K, go.sW, go.sN));
  LOCO_TRY(encode_rows_mn(&P.go_mn, go.ptr, go.C, T, K, go.sW, go.sN, P.D));
  LOCO_TRY(encode_rows_k(&P.gs_map, gP, T, T, P.heads * K, T, (long long)T * T, true));
  LOCO_TRY(encode_rows_k(&P.p_map_mn, P0, T, T, P.heads, T, (long long)T * T, true));
  LOCO_TRY(encode_rows_k(&P.p_st, gP, T, T, P.heads * K, T, (long long)T * T, false, 32));
  LOCO_TRY(encode_rows_k(&P.o_st, gqkv.ptr, gqkv.C, T, K, gqkv.sW, gqkv.sN, false, 32));
  const int qblocks = (T + 127) / 128;
  attn_vjp1_tc_kernel<<<dim3(qblocks, P.heads, K), kTcThreads, kTcSmem, s>>>(P);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  attn_vjp2_tc_kernel<<<dim3(qblocks, P.heads, K), kTcThreads, kTcSmem, s>>>(P);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Cross-attention of image tokens to a fixed context (text-conditioned U-Nets):
//   o = softmax(q K_c^T / sqrt(D)) V_c,   K_c | V_c = projections of the prompt embedding, constants
// of the Jacobian (d/dx only): the tangent rule is do = dP V_c with dS = dq K_c^T, the VJP is
// gq = gS K_c.  Same kernels as above with the key / value source decoupled from the query tensor:
// any number of query tokens, up to 256 (padded) keys of which Tk_valid exist.
//   q: [N, Tq, C] (heads of D = C / heads channels), kv: [1, Tk, 2C] (K_c | V_c, rows >= Tk_valid zero),
//   S: [N, heads, Tq, Tk] probabilities (P0 = row 0 for the tangent / cotangent rules), o: [N, Tq, C].
// ------------------------------------------------------------------------------------------------
bool attention_cross_eligible(int Tk, int C, int heads) {
  if (heads < 1 || C % heads != 0) return false;
  const int D = C / heads;
  return Tk % 64 == 0 && Tk >= 64 && Tk <= 256 && D % 64 == 0 && D <= 512;
}

namespace {
int fill_cross(AttnTcParams& P, View q, View kv, int heads, int Tk_valid) {
  const int C = q.C, Tq = q.H * q.W, Tk = kv.H * kv.W;
  LOCO_REQUIRE(attention_cross_eligible(Tk, C, heads), "cross-attention: unsupported shape (Tk=%d C=%d heads=%d)", Tk, C, heads);
  LOCO_REQUIRE(kv.C == 2 * C && kv.N == 1 && Tk_valid >= 1 && Tk_valid <= Tk, "cross-attention: kv must be [1, Tk, 2C]");
  LOCO_REQUIRE(!q.half && !kv.half && aligned16(q.ptr) && aligned16(kv.ptr) && q.sW % 4 == 0 && q.sN % 4 == 0 && kv.sW % 4 == 0,
               "cross-attention: fp32, 16-byte aligned tensors expected");
  memset(&P, 0, sizeof(P));
  P.T = Tq; P.Tk = Tk; P.Tk_valid = Tk_valid; P.cross = 1;
  P.heads = heads; P.D = C / heads; P.qo = 0; P.ko = 0; P.vo = C; P.hs = P.D;
  P.scale = 1.0f / sqrtf((float)P.D);
  LOCO_TRY(encode_rows_k(&P.kv_k, kv.ptr, kv.C, Tk, 1, kv.sW, kv.sN));
  LOCO_TRY(encode_rows_mn(&P.kv_mn, kv.ptr, kv.C, Tk, 1, kv.sW, kv.sN, P.D));
  return 0;
}
}  // namespace

int attention_cross_forward_tc(View q, View kv, int n_primal, int heads, int Tk_valid, float* S, View o, cudaStream_t s) {
  const int N = q.N, nt = N - n_primal, Tq = q.H * q.W, Tk = kv.H * kv.W;
  LOCO_REQUIRE(nt == 0 || n_primal == 1, "cross-attention: tangents need exactly one primal row");
  LOCO_REQUIRE(o.C == q.C && o.N == N && !o.half && aligned16(o.ptr) && aligned16(S) && o.sW % 4 == 0 && o.sN % 4 == 0,
               "cross-attention: output shape / alignment");
  LOCO_TRY(attention_tc_init());
  AttnTcParams P;
  LOCO_TRY(fill_cross(P, q, kv, heads, Tk_valid));
  P.n_primal = n_primal;
  LOCO_TRY(encode_rows_k(&P.qkv_k, q.ptr, q.C, Tq, N, q.sW, q.sN));
  LOCO_TRY(encode_rows_k(&P.p_map, S, Tk, Tq, heads * N, Tk, (long long)Tq * Tk));
  LOCO_TRY(encode_rows_k(&P.p_st, S, Tk, Tq, heads * N, Tk, (long long)Tq * Tk, false, 32));
  LOCO_TRY(encode_rows_k(&P.o_st, o.ptr, o.C, Tq, N, o.sW, o.sN, false, 32));
  const int qblocks = (Tq + 127) / 128;
  attn_fwd_tc_kernel<<<dim3(qblocks, heads, n_primal), kTcThreads, kTcSmem, s>>>(P);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  if (nt > 0) {
    attn_jvp_tc_kernel<<<dim3(qblocks, heads, nt), kTcThreads, kTcSmem, s>>>(P);
    count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

int attention_cross_vjp_tc(View go, View kv, int heads, int Tk_valid, const float* P0, View gq, cudaStream_t s) {
  const int K = go.N, Tq = go.H * go.W, Tk = kv.H * kv.W;
  LOCO_REQUIRE(gq.C == go.C && gq.N == K && !gq.half && !go.half && aligned16(go.ptr) && aligned16(gq.ptr) && aligned16(P0) &&
                   go.sW % 4 == 0 && go.sN % 4 == 0 && gq.sW % 4 == 0 && gq.sN % 4 == 0,
               "cross-attention VJP: shape / alignment");
  LOCO_TRY(attention_tc_init());
  AttnTcParams P;
  LOCO_TRY(fill_cross(P, go, kv, heads, Tk_valid));
  LOCO_TRY(encode_rows_k(&P.go_k, go.ptr, go.C, Tq, K, go.sW, go.sN));
  LOCO_TRY(encode_rows_k(&P.p_map, P0, Tk, Tq, heads, Tk, (long long)Tq * Tk));
  LOCO_TRY(encode_rows_k(&P.o_st, gq.ptr, gq.C, Tq, K, gq.sW, gq.sN, false, 32));
  attn_vjp1_tc_kernel<<<dim3((Tq + 127) / 128, heads, K), kTcThreads, kTcSmem, s>>>(P);
  count_launch(); LOCO_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace loco

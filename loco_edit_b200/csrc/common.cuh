// Shared device/host helpers for the LOCO-Edit B200 hot path (sm_100a only).
// PTX wrappers for mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (MMA / TMEM) and
// the small error-reporting layer used behind the C-ABI in include/loco_b200.h.
#pragma once
#include <cuda_runtime.h>
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>

namespace loco {

// ---------------------------------------------------------------------------------------------
// Error reporting: every C-ABI entry point returns 0 on success, non-zero on failure and never
// throws; the message is retrievable with loco_last_error().
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* get_error();

#define LOCO_CHECK_CUDA(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::loco::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                            \
    }                                                                                      \
  } while (0)

#define LOCO_REQUIRE(cond, ...)                                                            \
  do {                                                                                     \
    if (!(cond)) {                                                                         \
      ::loco::set_error(__VA_ARGS__);                                                      \
      return 2;                                                                            \
    }                                                                                      \
  } while (0)

#define LOCO_TRY(expr)                                                                     \
  do {                                                                                     \
    int _r = (expr);                                                                       \
    if (_r != 0) return _r;                                                                \
  } while (0)

// A 4-D activation view in HBM. Layout is channels-last: element (n,y,x,c) lives at
// ptr[n*sN + y*sH + x*sW + c] (strides in ELEMENTS). Channel slices of a wider buffer
// (skip-connection concat buffers) are expressed by offsetting ptr and keeping sW = the full buffer
// width.  Elements are fp32 (half == 0) or fp16 (half == 1: `ptr` is then only a typed handle to
// the first element, all address arithmetic goes through elem_ptr / the ld4t / st4t helpers).
// fp16 storage carries the same 10-bit mantissa as the tf32-rounded fp32 storage (every producer
// rounds to tf32 before the tensor core would truncate), at half the HBM bytes and twice the
// tcgen05 rate (kind::f16); its 5-bit exponent is why it is used for primal rows only.
struct View {
  float* ptr;
  int N, H, W, C;
  long long sN, sH, sW;
  int half;
};

inline float* elem_ptr(const View& v, long long elem_off) {
  return v.half ? reinterpret_cast<float*>(reinterpret_cast<char*>(v.ptr) + elem_off * 2)
                : v.ptr + elem_off;
}
inline View make_view(float* p, int N, int H, int W, int C, int half = 0) {
  View v; v.ptr = p; v.N = N; v.H = H; v.W = W; v.C = C; v.half = half;
  v.sW = C; v.sH = (long long)W * C; v.sN = (long long)H * W * C;
  return v;
}
inline View slice_c(const View& v, int c0, int C) {
  View r = v; r.ptr = elem_ptr(v, c0); r.C = C; return r;
}
inline View slice_n(const View& v, int n0, int N) {
  View r = v; r.ptr = elem_ptr(v, (long long)n0 * v.sN); r.N = N; return r;
}

// ---- programmatic dependent launch (PDL) -----------------------------------------------------
// A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start while its
// predecessor on the stream is still running: pdl_wait() blocks until every predecessor grid has
// completed and its memory operations are visible (a no-op for a plain launch); pdl_trigger() lets the
// successor's CTAs be scheduled as soon as every CTA of this grid has reached it.  The conv kernels
// run their prologue (tensor-map prefetch, mbarrier init, TMEM allocation) before pdl_wait(), i.e.
// under the tail of the GroupNorm / resampling kernel that produces their input.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

int num_sms();

// ---- multi-device support -----------------------------------------------------------------
// Every C-ABI entry point runs on the device that OWNS the caller's pointers (not on whatever the
// host thread's current device happens to be), so `--device cuda:1` works without a prior
// cudaSetDevice by the caller.  One-time kernel attribute setup, the graph-capture stream and the
// SM count are kept per device.
constexpr int kMaxDevices = 64;
int current_device();                       // cudaGetDevice, 0 on failure
int pointer_device(const void* dev_ptr);    // device owning a device pointer, -1 if unknown
struct DeviceGuard {                        // switch to `dev` (>= 0), restore on destruction
  explicit DeviceGuard(int dev);
  explicit DeviceGuard(const void* dev_ptr) : DeviceGuard(pointer_device(dev_ptr)) {}
  ~DeviceGuard();
  int prev = -1;
  bool switched = false;
};
// true exactly once per device for a given flag array (callers run their per-device setup then)
inline bool first_time_on_device(bool (&flags)[kMaxDevices]) {
  const int d = current_device();
  if (d < 0 || d >= kMaxDevices || flags[d]) return false;
  flags[d] = true;
  return true;
}

// Launch accounting (bench.py's `gpu_launches`) and optional per-kernel-family event timing
// (bench.py's roofline leg).  Families: 0 = tcgen05 conv GEMM, 1 = GroupNorm, 2 = other.
void count_launch(int n = 1);
long long launch_count();
struct ProfScope {   // records a start/stop event pair around a launch when profiling is on
  ProfScope(int family, double work, cudaStream_t s);
  ~ProfScope();
  int family; cudaStream_t stream; int slot;
};
void profile_enable(bool on);
bool profile_is_on();
// sums elapsed ms / work / launches per family since profile_enable(true); synchronises the device
int profile_collect(double* ms, double* work, long long* launches, int nfam);

#ifdef __CUDACC__
// ---------------------------------------------------------------------------------------------
// Device-side PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (-> launch error) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > (1u << 24)) {
      printf("loco: mbarrier timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

// ---- thread-block clusters / CTA pairs (cta_group::2) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared-memory offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(const void* local_smem, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem)), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads of a CTA pair: the box lands in THIS CTA's shared memory, the bytes are completed on
// an mbarrier given by its shared::cluster address (the leader CTA's barrier).
__device__ __forceinline__ void tma_load_4d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                 int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint32_t bar_cluster_addr,
                                                 int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish_pair() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
// D[tmem of both CTAs] (+)= A * B^T over the CTA pair: M = 256 (128 rows from each CTA's A tile),
// B = N/2 rows from each CTA's shared memory.  Issued by the leader CTA only.
__device__ __forceinline__ void umma_tf32_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                               uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive (once all previously issued MMAs of the pair completed) on the mbarrier at this offset in
// the CTAs selected by `mask`.
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
          smem_u32(bar)),
      "h"(mask)
      : "memory");
}

// ---- tcgen05 / TMEM ----
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                   smem_u32(smem_dst)),
               "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, TF32 inputs (fp32 words in smem), fp32 accumulate.
__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with fp16 operands (kind::f16: 16 elements = 32 bytes of K per instruction, twice the rate).
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                              uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Arrive on an mbarrier once all previously issued tcgen05.mma of this thread have completed.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread t = lane t).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// Round-to-nearest fp32 -> tf32 (kept in an fp32 container). The tensor core ignores the low
// 13 mantissa bits, so producers round instead of letting the MMA truncate.
__device__ __forceinline__ float round_tf32(float x) {
  uint32_t u;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
  return __uint_as_float(u);
}

// ---- typed 4-element vector access of an activation tensor (H16: fp16 storage, else fp32) ----
template <bool H16>
__device__ __forceinline__ float4 ld4t(const float* base, long long elem_off) {
  if (H16) {
    const uint2 u = *reinterpret_cast<const uint2*>(reinterpret_cast<const __half*>(base) + elem_off);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
    return make_float4(a.x, a.y, b.x, b.y);
  }
  return *reinterpret_cast<const float4*>(base + elem_off);
}
template <bool H16>
__device__ __forceinline__ void st4t(float* base, long long elem_off, float4 v) {
  if (H16) {
    const __half2 a = __floats2half2_rn(v.x, v.y), b = __floats2half2_rn(v.z, v.w);
    uint2 u;
    u.x = *reinterpret_cast<const uint32_t*>(&a);
    u.y = *reinterpret_cast<const uint32_t*>(&b);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(base) + elem_off) = u;
  } else {
    *reinterpret_cast<float4*>(base + elem_off) = v;
  }
}
// value as it will be read back from fp16 storage
__device__ __forceinline__ float round_f16(float x) { return __half2float(__float2half_rn(x)); }

// 256-bit global accesses (sm_100: LDG.256 / STG.256): one full 32-byte sector per thread
struct __align__(32) U32x8 { uint32_t v[8]; };
__device__ __forceinline__ U32x8 ldg256(const void* p) {
  U32x8 r;
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                 "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
__device__ __forceinline__ U32x8 ld256(const void* p) {       // coherent (read-modify-write targets)
  U32x8 r;
  asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                 "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ void st256(void* p, const uint32_t* v) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(v[0]), "r"(v[1]), "r"(v[2]),
               "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}

// sum 16 per-lane values over the 32 lanes with a halving butterfly (31 shuffles instead of 80):
// afterwards lane l holds the total of value (l >> 1).
template <int OFF, int HALF>
__device__ __forceinline__ void butterfly_step(float (&v)[16], int lane) {
  const bool hi = (lane & OFF) != 0;
#pragma unroll
  for (int i = 0; i < HALF; ++i) {
    const float send = hi ? v[i] : v[i + HALF];
    const float keep = hi ? v[i + HALF] : v[i];
    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, OFF);
  }
}
__device__ __forceinline__ float butterfly16(float (&v)[16], int lane) {
  butterfly_step<16, 8>(v, lane);
  butterfly_step<8, 4>(v, lane);
  butterfly_step<4, 2>(v, lane);
  butterfly_step<2, 1>(v, lane);
  return v[0] + __shfl_xor_sync(0xffffffffu, v[0], 1);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
#endif  // __CUDACC__

}  // namespace loco

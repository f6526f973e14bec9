// extern "C" boundary of libloco_b200.so -- see include/loco_b200.h for the contract.
#include "../../include/loco_b200.h"
#include <algorithm>
#include "attention.cuh"
#include "conv_gemm.cuh"
#include "layers.cuh"
#include "pullback.cuh"
#include "unet.cuh"

using namespace loco;

struct loco_unet { Model* m; };
struct loco_plan { Plan* p; };

#define ST(s) reinterpret_cast<cudaStream_t>(s)
// run on the device that owns the caller's buffers (see DeviceGuard in common.cuh)
#define ON_DEVICE_OF(ptr) DeviceGuard _dev_guard(static_cast<const void*>(ptr))
#define GUARD_BEGIN try {
#define GUARD_END                                            \
  }                                                          \
  catch (const std::exception& e) {                          \
    set_error("exception: %s", e.what());                    \
    return 99;                                               \
  }                                                          \
  catch (...) {                                              \
    set_error("unknown exception");                          \
    return 99;                                               \
  }

static int require_device() {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    cudaGetLastError();
    set_error("no CUDA device: this library has no CPU fallback (%s)", cudaGetErrorString(e));
    return 10;
  }
  return 0;
}

extern "C" {

int loco_abi_version(void) { return 3; }
long long loco_launch_count(void) { return launch_count(); }
int loco_profile_enable(int on) { profile_enable(on != 0); return 0; }
int loco_profile_collect(double* ms, double* work, long long* launches, int nfam) {
  return profile_collect(ms, work, launches, nfam);
}
const char* loco_last_error(void) { return get_error(); }

// ------------------------------------------------------------------------------------------------
int loco_unet_create(const loco_arch_t* a, loco_unet_t** out) {
  GUARD_BEGIN
  LOCO_REQUIRE(a && out, "loco_unet_create: null argument");
  LOCO_REQUIRE(a->n_levels >= 1 && a->n_levels <= 8 && a->n_attn >= 0 && a->n_attn <= 4,
               "loco_unet_create: bad level/attention counts");
  LOCO_REQUIRE(a->in_ch >= 1 && a->in_ch <= 4 && a->out_ch >= 1 && a->out_ch <= 4,
               "loco_unet_create: 1..4 input / output channels (got %d / %d)", a->in_ch, a->out_ch);
  LOCO_REQUIRE(a->kind != 1 || (a->in_ch == 3 && a->out_ch == 3), "loco_unet_create: the guided-diffusion U-Net takes 3-channel images");
  LOCO_REQUIRE(a->ch % 128 == 0 && a->ch >= 128, "loco_unet_create: ch must be a multiple of 128");
  LOCO_REQUIRE((a->kind == 2 ? a->resolution : (a->resolution >> (a->n_levels - 1))) >= 4 &&
                   (a->resolution & (a->resolution - 1)) == 0,
               "loco_unet_create: resolution must be a power of two, >= 4 at the coarsest level");
  Arch A;
  A.ch = a->ch; A.n_levels = a->n_levels;
  for (int i = 0; i < 8; ++i) A.ch_mult[i] = a->ch_mult[i];
  A.num_res_blocks = a->num_res_blocks; A.n_attn = a->n_attn;
  for (int i = 0; i < 4; ++i) A.attn_resolutions[i] = a->attn_resolutions[i];
  A.resolution = a->resolution; A.in_ch = a->in_ch; A.out_ch = a->out_ch; A.gn_eps = a->gn_eps;
  LOCO_REQUIRE(a->kind >= 0 && a->kind <= 2, "loco_unet_create: unknown architecture kind %d", a->kind);
  LOCO_REQUIRE(a->kind != 1 || (a->head_ch > 0 && a->head_ch % 4 == 0),
               "loco_unet_create: guided-diffusion U-Net needs head_ch > 0");
  A.kind = a->kind; A.head_ch = a->kind == 1 ? a->head_ch : 0;
  LOCO_REQUIRE(a->ctx_dim >= 0 && (a->ctx_dim == 0 || (a->kind == 0 && a->ctx_dim % 4 == 0 && a->ctx_heads >= 1)),
               "loco_unet_create: cross-attention (ctx_dim %d, heads %d) needs kind 0, ctx_dim %% 4 == 0", a->ctx_dim,
               a->ctx_heads);
  A.ctx_dim = a->ctx_dim; A.ctx_heads = a->ctx_dim > 0 ? a->ctx_heads : 1;
  *out = new loco_unet{new Model(A)};
  return 0;
  GUARD_END
}
void loco_unet_destroy(loco_unet_t* m) {
  if (m) { delete m->m; delete m; }
}
long long loco_unet_weight_floats(const loco_unet_t* m) { return m ? (long long)m->m->arena_floats : -1; }
int loco_unet_bind_weights(loco_unet_t* m, float* arena) {
  LOCO_REQUIRE(m && arena, "loco_unet_bind_weights: null argument");
  LOCO_REQUIRE(((uintptr_t)arena & 255) == 0, "loco_unet_bind_weights: arena must be 256B aligned");
  m->m->arena = arena;
  return 0;
}
int loco_unet_num_params(const loco_unet_t* m) { return m ? (int)m->m->slots.size() : -1; }
int loco_unet_param_info(const loco_unet_t* m, int i, char* name, int name_cap, int* shape, int* ndim) {
  LOCO_REQUIRE(m && i >= 0 && i < (int)m->m->slots.size(), "loco_unet_param_info: bad index");
  const ParamSlot& s = m->m->slots[i];
  if (name && name_cap > 0) {
    strncpy(name, s.name.c_str(), name_cap - 1);
    name[name_cap - 1] = 0;
  }
  if (ndim) *ndim = (int)s.shape.size();
  if (shape)
    for (size_t j = 0; j < s.shape.size() && j < 4; ++j) shape[j] = s.shape[j];
  return 0;
}
int loco_unet_load_param(loco_unet_t* m, const char* name, const float* src, long long numel,
                         void* stream) {
  ON_DEVICE_OF(src);
  GUARD_BEGIN
  LOCO_REQUIRE(m && name && src, "loco_unet_load_param: null argument");
  LOCO_TRY(require_device());
  return m->m->load_param(name, src, numel, ST(stream));
  GUARD_END
}

// ------------------------------------------------------------------------------------------------
int loco_plan_create(const loco_unet_t* m, int n_primal, int n_tangent, int n_cotangent,
                     loco_plan_t** out) {
  GUARD_BEGIN
  LOCO_REQUIRE(m && out, "loco_plan_create: null argument");
  Plan* p = new Plan(m->m, n_primal, n_tangent, n_cotangent);
  const int r = p->build(nullptr);
  if (r != 0) { delete p; return r; }
  *out = new loco_plan{p};
  return 0;
  GUARD_END
}
int loco_plan_create_ex(const loco_unet_t* m, int n_primal, int n_tangent, int n_cotangent, int flags,
                        loco_plan_t** out) {
  GUARD_BEGIN
  LOCO_REQUIRE(m && out, "loco_plan_create_ex: null argument");
  LOCO_REQUIRE((flags & ~LOCO_PLAN_FP16) == 0, "loco_plan_create_ex: unknown flags 0x%x", flags);
  Plan* p = new Plan(m->m, n_primal, n_tangent, n_cotangent, flags);
  const int r = p->build(nullptr);
  if (r != 0) { delete p; return r; }
  *out = new loco_plan{p};
  return 0;
  GUARD_END
}
void loco_plan_destroy(loco_plan_t* p) {
  if (p) { delete p->p; delete p; }
}
long long loco_plan_workspace_bytes(const loco_plan_t* p) {
  return p ? (long long)p->p->workspace_floats * 4 : -1;
}
int loco_plan_bind(loco_plan_t* p, void* workspace) {
  GUARD_BEGIN
  LOCO_REQUIRE(p && workspace, "loco_plan_bind: null argument");
  LOCO_REQUIRE(((uintptr_t)workspace & 255) == 0, "loco_plan_bind: workspace must be 256B aligned");
  LOCO_TRY(require_device());
  LOCO_TRY(p->p->model->check_loaded());
  return p->p->build(reinterpret_cast<float*>(workspace));
  GUARD_END
}
int loco_plan_info(const loco_plan_t* p, double* fwd_flops, double* vjp_flops, int* fwd_ops,
                   int* vjp_ops) {
  LOCO_REQUIRE(p, "loco_plan_info: null plan");
  if (fwd_flops) *fwd_flops = p->p->fwd_flops;
  if (vjp_flops) *vjp_flops = p->p->vjp_flops;
  if (fwd_ops) *fwd_ops = p->p->fwd_launches;
  if (vjp_ops) *vjp_ops = p->p->vjp_launches;
  return 0;
}
int loco_unet_forward(loco_plan_t* p, const float* x, float t, float* eps, void* stream) {
  GUARD_BEGIN
  LOCO_REQUIRE(p && x && eps, "loco_unet_forward: null argument");
  return p->p->forward(x, t, eps, ST(stream));
  GUARD_END
}
int loco_plan_set_condition(loco_plan_t* p, const float* cond, void* stream) {
  GUARD_BEGIN
  LOCO_REQUIRE(p, "loco_plan_set_condition: null plan");
  return p->p->set_condition(cond, ST(stream));
  GUARD_END
}
int loco_plan_set_context(loco_plan_t* p, const float* ctx, int n_tokens, void* stream) {
  GUARD_BEGIN
  LOCO_REQUIRE(p, "loco_plan_set_context: null plan");
  return p->p->set_context(ctx, n_tokens, ST(stream));
  GUARD_END
}
int loco_unet_vjp(loco_plan_t* p, const float* g_eps, float* gx, void* stream) {
  GUARD_BEGIN
  LOCO_REQUIRE(p && g_eps && gx, "loco_unet_vjp: null argument");
  return p->p->vjp(g_eps, gx, ST(stream));
  GUARD_END
}

// ------------------------------------------------------------------------------------------------
static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

long long loco_orthonormalise_scratch_bytes(int k) { return (long long)sizeof(double) * (3 * k * k + 2 * k); }

long long loco_pullback_scratch_bytes(int k, long long d) {
  const size_t kd = align_up((size_t)k * d * 4, 256);
  const size_t k1d = align_up((size_t)(k + 1) * d * 4, 256);
  return (long long)(2 * k1d + 4 * kd + align_up((size_t)loco_orthonormalise_scratch_bytes(k), 256));
}

static int pullback_probe_impl(loco_plan_t* p, const float* xt, float t, float at,
                               const unsigned char* mask, int noise, const float* V, int k, int k_invert,
                               long long d, float* u_full, float* w_out, void* scratch, void* stream);

int loco_pullback_probe(loco_plan_t* p, const float* xt, float t, float at, const unsigned char* mask,
                        int noise, const float* V, int k, long long d, float* u_full, float* w_out,
                        void* scratch, void* stream) {
  return pullback_probe_impl(p, xt, t, at, mask, noise, V, k, k, d, u_full, w_out, scratch, stream);
}

// Probe rows of a mixed batch: rows [0, k_invert) see the mask, rows [k_invert, k) its complement.
// A rank's shard of the joint {edit, null} probe set of run_edit_null_space_projection may straddle
// the boundary between the two bases ("shard the 10 probes of {edit, null} jointly").
int loco_pullback_probe_pair(loco_plan_t* p, const float* xt, float t, float at, const unsigned char* mask,
                             int noise, const float* V, int k, int k_invert, long long d, float* u_full,
                             float* w_out, void* scratch, void* stream) {
  LOCO_REQUIRE(mask != nullptr && k_invert >= 0 && k_invert <= k, "loco_pullback_probe_pair: bad mask / split");
  return pullback_probe_impl(p, xt, t, at, mask, noise, V, k, k_invert, d, u_full, w_out, scratch, stream);
}

int loco_pullback_pair_iteration(loco_plan_t* p, const float* xt, float t, float at,
                                 const unsigned char* mask, int noise, const float* V, int k1, int k2,
                                 long long d, int align_sign, float* u_full, float* w_out, float* V_out,
                                 float* s_out, void* scratch, void* stream) {
  ON_DEVICE_OF(xt);
  GUARD_BEGIN
  LOCO_REQUIRE(V_out && s_out && scratch && w_out && mask, "loco_pullback_pair_iteration: null argument");
  const int k = k1 + k2;
  LOCO_TRY(pullback_probe_impl(p, xt, t, at, mask, noise, V, k, k1, d, u_full, w_out, scratch, stream));
  const size_t kd = align_up((size_t)k * d * 4, 256);
  const size_t k1d = align_up((size_t)(k + 1) * d * 4, 256);
  double* oscr = reinterpret_cast<double*>(reinterpret_cast<char*>(scratch) + 2 * k1d + 4 * kd);
  // the two bases are orthonormalised independently (rows [0,k1) and [k1,k1+k2))
  LOCO_TRY(orthonormalise(w_out, k1, d, align_sign ? V : nullptr, V_out, s_out, oscr, ST(stream)));
  LOCO_TRY(orthonormalise(w_out + (size_t)k1 * d, k2, d, align_sign ? V + (size_t)k1 * d : nullptr,
                          V_out + (size_t)k1 * d, s_out + k1, oscr, ST(stream)));
  return 0;
  GUARD_END
}

static int pullback_probe_impl(loco_plan_t* p, const float* xt, float t, float at,
                               const unsigned char* mask, int noise, const float* V, int k, int k_invert,
                               long long d, float* u_full, float* w_out, void* scratch, void* stream) {
  ON_DEVICE_OF(xt);
  GUARD_BEGIN
  LOCO_REQUIRE(p && xt && V && u_full && w_out && scratch, "loco_pullback_probe: null argument");
  Plan& P = *p->p;
  LOCO_REQUIRE(P.NP == 1 && P.NT == k && P.NC == k, "loco_pullback_probe: plan is (%d,%d,%d), need (1,%d,%d)",
               P.NP, P.NT, P.NC, k, k);
  LOCO_REQUIRE(P.in_elems() == P.out_elems(), "loco_pullback_probe: needs a network with equal input and output shapes");
  LOCO_REQUIRE(d == P.in_elems(), "loco_pullback_probe: d=%lld does not match the model", d);
  cudaStream_t s = ST(stream);
  const size_t kd = align_up((size_t)k * d * 4, 256);
  const size_t k1d = align_up((size_t)(k + 1) * d * 4, 256);
  char* sp = reinterpret_cast<char*>(scratch);
  float* xin = reinterpret_cast<float*>(sp); sp += k1d;
  float* eps = reinterpret_cast<float*>(sp); sp += k1d;
  float* g_eps = reinterpret_cast<float*>(sp); sp += kd;
  float* gx_direct = reinterpret_cast<float*>(sp); sp += kd;
  float* gx_unet = reinterpret_cast<float*>(sp); sp += kd;
  // batch = [x_t ; v_1 .. v_k]: one fused primal + k-tangent pass
  LOCO_CHECK_CUDA(cudaMemcpyAsync(xin, xt, sizeof(float) * d, cudaMemcpyDeviceToDevice, s));
  LOCO_CHECK_CUDA(cudaMemcpyAsync(xin + d, V, sizeof(float) * k * d, cudaMemcpyDeviceToDevice, s));
  LOCO_TRY(P.forward(xin, t, eps, s));
  LOCO_TRY(pmp_jvp_epilogue(V, eps + d, mask, at, noise, k, k_invert, d, u_full, g_eps, gx_direct, s));
  LOCO_TRY(P.vjp(g_eps, gx_unet, s));
  LOCO_TRY(axpy(gx_direct, gx_unet, 1.0f, (long long)k * d, w_out, s));
  return 0;
  GUARD_END
}

int loco_pullback_iteration(loco_plan_t* p, const float* xt, float t, float at,
                            const unsigned char* mask, int noise, const float* V, int k, long long d,
                            int align_sign, float* u_full, float* w_out, float* V_out, float* s_out,
                            void* scratch, void* stream) {
  ON_DEVICE_OF(xt);
  GUARD_BEGIN
  LOCO_REQUIRE(V_out && s_out && scratch, "loco_pullback_iteration: null argument");
  const size_t kd = align_up((size_t)k * d * 4, 256);
  const size_t k1d = align_up((size_t)(k + 1) * d * 4, 256);
  char* sp = reinterpret_cast<char*>(scratch) + 2 * k1d + 3 * kd;
  float* w_tmp = reinterpret_cast<float*>(sp); sp += kd;
  double* oscr = reinterpret_cast<double*>(sp);
  float* w = w_out ? w_out : w_tmp;
  LOCO_TRY(loco_pullback_probe(p, xt, t, at, mask, noise, V, k, d, u_full, w, scratch, stream));
  LOCO_TRY(orthonormalise(w, k, d, align_sign ? V : nullptr, V_out, s_out, oscr, ST(stream)));
  return 0;
  GUARD_END
}

// ------------------------------------------------------------------------------------------------
int loco_pmp_forward(const float* x, const float* eps, float at, long long n, float* out, void* stream) {
  ON_DEVICE_OF(x);
  LOCO_TRY(require_device());
  return pmp_forward(x, eps, at, n, out, ST(stream));
}
int loco_combine3(const float* a, float wa, const float* b, float wb, const float* c, float wc, long long n,
                  float* out, void* stream) {
  ON_DEVICE_OF(a);
  LOCO_TRY(require_device());
  LOCO_REQUIRE(a && out, "loco_combine3: null argument");
  return combine3(a, wa, b, wb, c, wc, n, out, ST(stream));
}
int loco_pmp_jvp_epilogue(const float* V, const float* eps_dot, const unsigned char* mask, float at, int noise,
                          int k, int k_invert, long long d, float* u, float* g_eps, float* gx_direct,
                          void* stream) {
  ON_DEVICE_OF(V);
  LOCO_TRY(require_device());
  LOCO_REQUIRE(V && eps_dot && u && g_eps && gx_direct, "loco_pmp_jvp_epilogue: null argument");
  return pmp_jvp_epilogue(V, eps_dot, mask, at, noise, k, k_invert, d, u, g_eps, gx_direct, ST(stream));
}
int loco_orthonormalise(const float* W, int k, long long d, const float* v_prev, float* V,
                        float* s_out, void* scratch, void* stream) {
  ON_DEVICE_OF(W);
  LOCO_TRY(require_device());
  return orthonormalise(W, k, d, v_prev, V, s_out, reinterpret_cast<double*>(scratch), ST(stream));
}
int loco_nullspace_project(const float* vT_mod, int k, const float* Vn, int k_null, long long d,
                           int project, float* out, void* scratch, void* stream) {
  ON_DEVICE_OF(vT_mod);
  LOCO_TRY(require_device());
  return nullspace_project(vT_mod, k, Vn, k_null, d, project, out, reinterpret_cast<double*>(scratch),
                           ST(stream));
}
int loco_ddim_step(const float* xt, const float* et, const float* noise, float at, float at_next,
                   float eta, long long n, float* xt_next, float* x0_pred, void* stream) {
  ON_DEVICE_OF(xt);
  LOCO_TRY(require_device());
  return ddim_step(xt, et, noise, at, at_next, eta, n, xt_next, x0_pred, ST(stream));
}
int loco_axpy(const float* x, const float* v, float scale, long long n, float* out, void* stream) {
  ON_DEVICE_OF(x);
  LOCO_TRY(require_device());
  return axpy(x, v, scale, n, out, ST(stream));
}
int loco_mask_indices(const unsigned char* mask, long long d, int* idx, int* count, void* stream) {
  ON_DEVICE_OF(mask);
  LOCO_TRY(require_device());
  return mask_indices(mask, d, idx, count, ST(stream));
}
int loco_gather_rows(const float* src, int rows, long long d, const int* idx, int count, float* out,
                     void* stream) {
  ON_DEVICE_OF(src);
  LOCO_TRY(require_device());
  return gather_rows(src, rows, d, idx, count, out, ST(stream));
}
int loco_scatter_rows(const float* src, int rows, long long d, const int* idx, int count, float* out,
                      void* stream) {
  ON_DEVICE_OF(src);
  LOCO_TRY(require_device());
  return scatter_rows(src, rows, d, idx, count, out, ST(stream));
}
int loco_gram(const float* A, int ka, const float* B, int kb, long long d, double* G, void* stream) {
  ON_DEVICE_OF(A);
  LOCO_TRY(require_device());
  return gram(A, ka, B, kb, d, G, ST(stream));
}

// ------------------------------------------------------------------------------------------------
int loco_conv2d_nhwc(int kind, const float* x, int N, int H, int W, int Cx, const float* w, int Cout,
                     int Cin, float* wpack, const float* bias, int bias_rows, const float* addend,
                     int accumulate, float* y, void* splitk_scratch, long long splitk_bytes,
                     void* stream) {
  return loco_conv2d_nhwc_ex(kind, x, N, H, W, Cx, w, Cout, Cin, wpack, bias, bias_rows, addend, accumulate, y,
                             splitk_scratch, splitk_bytes, 0, 0, stream);
}

int loco_conv2d_nhwc_ex(int kind, const void* x_, int N, int H, int W, int Cx, const float* w, int Cout,
                        int Cin, void* wpack_, const float* bias, int bias_rows, const void* addend_,
                        int accumulate, void* y_, void* splitk_scratch, long long splitk_bytes,
                        int in16, int out16, void* stream) {
  const float* x = reinterpret_cast<const float*>(x_);
  float* wpack = reinterpret_cast<float*>(wpack_);
  const float* addend = reinterpret_cast<const float*>(addend_);
  float* y = reinterpret_cast<float*>(y_);
  ON_DEVICE_OF(x);
  GUARD_BEGIN
  LOCO_TRY(require_device());
  cudaStream_t s = ST(stream);
  ConvProblem p;
  const int ksz = kind == CONV_1x1 ? 1 : 3;
  int Ho = H, Wo = W, Cy = Cout;
  if (kind == CONV_3x3 || kind == CONV_1x1 || kind == CONV_3x3_S2) {
    LOCO_REQUIRE(Cx == Cin, "loco_conv2d_nhwc: x has %d channels, weight expects %d", Cx, Cin);
    if (in16) LOCO_TRY(pack_conv_fprop16(w, wpack, Cout, Cin, ksz, ksz, s));
    else LOCO_TRY(pack_conv_fprop(w, wpack, Cout, Cin, ksz, ksz, s));
    if (kind == CONV_3x3_S2) { Ho = H / 2; Wo = W / 2; }
    p.Kc = Cin; p.Ngemm = Cout;
  } else {
    LOCO_REQUIRE(Cx == Cout, "loco_conv2d_nhwc: dy has %d channels, weight has Cout=%d", Cx, Cout);
    if (in16) LOCO_TRY(pack_conv_dgrad16(w, wpack, Cout, Cin, ksz, ksz, Cout, 0, s));
    else LOCO_TRY(pack_conv_dgrad(w, wpack, Cout, Cin, ksz, ksz, Cout, 0, s));
    if (kind == CONV_3x3_S2_DGRAD) { Ho = 2 * H; Wo = 2 * W; }
    p.Kc = Cout; p.Ngemm = Cin; Cy = Cin;
  }
  p.kind = kind;
  p.in = make_view(const_cast<float*>(x), N, H, W, Cx, in16 ? 1 : 0);
  p.out = make_view(y, N, Ho, Wo, Cy, out16 ? 1 : 0);
  p.wpack = wpack;
  p.bias = bias; p.bias_rows = bias_rows;
  View add;
  if (addend) { add = make_view(const_cast<float*>(addend), N, Ho, Wo, Cy, out16 ? 1 : 0); p.addend = &add; }
  p.accumulate = accumulate; p.round_out = 0;
  if (splitk_scratch && splitk_bytes >= (1 << 20)) {
    // layout: [4096 int counters (zeroed by the caller)] [partial tiles]
    p.splitk_counters = reinterpret_cast<int*>(splitk_scratch);
    p.splitk_max_tiles = 2048;   // arrive [0,2048) + done [2048,4096)
    p.splitk_partial = reinterpret_cast<float*>(reinterpret_cast<char*>(splitk_scratch) + 4096 * 4);
    p.splitk_partial_floats = (splitk_bytes - 4096 * 4) / 4;
  }
  ConvLaunch L;
  LOCO_TRY(conv_prepare(p, &L));
  return conv_run(L, s);
  GUARD_END
}

// y = conv3x3(x; w) + conv1x1(x2; w2) + bias (first bias_rows rows) in ONE launch (the ResnetBlock's
// conv2 + nin_shortcut, halo variants only), optionally accumulating the GroupNorm statistics
// (sum, sum of squares per row and group of `stat_cg` channels) of the stored result.
int loco_conv2d_fused_nhwc(const float* x, int N, int H, int W, int Cin, const float* w, int Cout,
                           const float* x2, int C2, const float* w2, float* wpack, float* wpack2,
                           const float* bias, int bias_rows, float* y, double* stats, int stat_cg,
                           void* stream) {
  ON_DEVICE_OF(x);
  GUARD_BEGIN
  LOCO_TRY(require_device());
  LOCO_REQUIRE(x && w && y && wpack, "loco_conv2d_fused_nhwc: null argument");
  cudaStream_t s = ST(stream);
  LOCO_TRY(conv_init());
  LOCO_REQUIRE(conv_halo_eligible(CONV_3x3, N, H, W, Cout),
               "loco_conv2d_fused_nhwc: shape %d x %dx%d x %d is not served by the halo conv variants", N, H,
               W, Cout);
  ConvProblem p;
  p.kind = CONV_3x3;
  LOCO_TRY(pack_conv_fprop(w, wpack, Cout, Cin, 3, 3, s));
  p.Kc = Cin; p.Ngemm = Cout;
  p.in = make_view(const_cast<float*>(x), N, H, W, Cin);
  p.out = make_view(y, N, H, W, Cout);
  p.wpack = wpack;
  p.bias = bias; p.bias_rows = bias_rows;
  View v2;
  if (x2) {
    LOCO_REQUIRE(w2 && wpack2, "loco_conv2d_fused_nhwc: shortcut weights missing");
    LOCO_TRY(pack_conv_fprop(w2, wpack2, Cout, C2, 1, 1, s));
    v2 = make_view(const_cast<float*>(x2), N, H, W, C2);
    p.in2 = &v2; p.wpack2 = wpack2; p.Kc2 = C2;
  }
  if (stats) {
    LOCO_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 64 * N, s));
    p.st_ptr[0] = stats; p.st_cg[0] = stat_cg; p.st_choff[0] = 0;
  }
  p.round_out = 0;
  ConvLaunch L;
  LOCO_TRY(conv_prepare(p, &L));
  return conv_run(L, s);
  GUARD_END
}

// Host-side variant decision for a stride-1 3x3 (fprop or dgrad) of this shape: 1 = served by the
// halo / CTA-pair kernels (and eligible for the fused 1x1 shortcut), 0 = one-/two-tile kernel.
int loco_conv_halo_eligible(int kind, int N, int H, int W, int Cout) {
  return conv_halo_eligible(kind, N, H, W, Cout) ? 1 : 0;
}

// Micro-benchmark of one prepared conv launch: `reps` back-to-back launches between two events.
int loco_conv_bench(int kind, float* x, int N, int H, int W, int Cx, float* wpack, int Cout, int Cin,
                    float* y, void* splitk_scratch, long long splitk_bytes, int max_ksplit, int reps,
                    float* ms_out, int* ksplit_out, int* grid_out, void* stream) {
  return loco_conv_bench_ex(kind, x, N, H, W, Cx, wpack, Cout, Cin, y, splitk_scratch, splitk_bytes, max_ksplit,
                            reps, 0, 0, nullptr, nullptr, ms_out, ksplit_out, grid_out, stream);
}

int loco_conv_bench_ex(int kind, void* x_, int N, int H, int W, int Cx, void* wpack_, int Cout, int Cin,
                       void* y_, void* splitk_scratch, long long splitk_bytes, int max_ksplit, int reps,
                       int in16, int out16, const void* addend_, double* stats, float* ms_out,
                       int* ksplit_out, int* grid_out, void* stream) {
  float* x = reinterpret_cast<float*>(x_);
  float* wpack = reinterpret_cast<float*>(wpack_);
  float* y = reinterpret_cast<float*>(y_);
  ON_DEVICE_OF(x);
  GUARD_BEGIN
  LOCO_TRY(require_device());
  cudaStream_t s = ST(stream);
  ConvProblem p;
  int Ho = H, Wo = W, Cy = Cout;
  if (kind == CONV_3x3 || kind == CONV_1x1 || kind == CONV_3x3_S2) {
    if (kind == CONV_3x3_S2) { Ho = H / 2; Wo = W / 2; }
    p.Kc = Cin; p.Ngemm = Cout;
  } else {
    if (kind == CONV_3x3_S2_DGRAD) { Ho = 2 * H; Wo = 2 * W; }
    p.Kc = Cout; p.Ngemm = Cin; Cy = Cin;
  }
  p.kind = kind;
  p.in = make_view(x, N, H, W, Cx, in16 ? 1 : 0);
  p.out = make_view(y, N, Ho, Wo, Cy, out16 ? 1 : 0);
  p.wpack = wpack;
  p.round_out = 1;
  View addv;
  if (addend_) {
    addv = make_view(reinterpret_cast<float*>(const_cast<void*>(addend_)), N, Ho, Wo, Cy, out16 ? 1 : 0);
    p.addend = &addv;
  }
  if (stats) { p.st_ptr[0] = stats; p.st_cg[0] = Cy / 32; p.st_choff[0] = 0; }
  if (splitk_scratch && max_ksplit > 1) {
    p.splitk_counters = reinterpret_cast<int*>(splitk_scratch);
    p.splitk_max_tiles = 2048;   // arrive [0,2048) + done [2048,4096)
    p.splitk_partial = reinterpret_cast<float*>(reinterpret_cast<char*>(splitk_scratch) + 4096 * 4);
    p.splitk_partial_floats = (splitk_bytes - 4096 * 4) / 4;
  }
  ConvLaunch L;
  LOCO_TRY(conv_prepare(p, &L));
  for (int i = 0; i < L.nlaunch; ++i)
    if (L.p[i].ksplit > max_ksplit) {
      L.p[i].ksplit = max_ksplit;
      const int items = L.p[i].tiles_x * L.p[i].tiles_y * L.p[i].tiles_n * L.p[i].tiles_co * max_ksplit;
      L.grid[i] = items < num_sms() ? items : num_sms();
    }
  if (ksplit_out) *ksplit_out = L.p[0].ksplit;
  if (grid_out) *grid_out = L.grid[0];
  cudaEvent_t a, b;
  LOCO_CHECK_CUDA(cudaEventCreate(&a));
  LOCO_CHECK_CUDA(cudaEventCreate(&b));
  for (int i = 0; i < 3; ++i) LOCO_TRY(conv_run(L, s));
  LOCO_CHECK_CUDA(cudaEventRecord(a, s));
  for (int i = 0; i < reps; ++i) LOCO_TRY(conv_run(L, s));
  LOCO_CHECK_CUDA(cudaEventRecord(b, s));
  LOCO_CHECK_CUDA(cudaEventSynchronize(b));
  float ms = 0.f;
  LOCO_CHECK_CUDA(cudaEventElapsedTime(&ms, a, b));
  *ms_out = ms / reps;
  cudaEventDestroy(a); cudaEventDestroy(b);
  return 0;
  GUARD_END
}

int loco_groupnorm_silu_fwd(const float* x, int N, int H, int W, int C, int n_primal,
                            const float* gamma, const float* beta, float eps, int silu, float* y,
                            void* stats, void* stream) {
  ON_DEVICE_OF(x);
  LOCO_TRY(require_device());
  cudaStream_t s = ST(stream);
  View xv = make_view(const_cast<float*>(x), N, H, W, C);
  View yv = make_view(y, N, H, W, C);
  LOCO_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 64 * N, s));
  LOCO_TRY(gn_stats_fwd(xv, n_primal, reinterpret_cast<double*>(stats), s));
  return gn_apply_fwd(xv, n_primal, reinterpret_cast<double*>(stats), gamma, beta, eps, silu, 0, yv, s);
}
int loco_groupnorm_silu_vjp(const float* xp, int H, int W, int C, const float* gy, int K,
                            const float* gamma, const float* beta, float eps, int silu, float* gx,
                            void* stats, void* stream) {
  ON_DEVICE_OF(gy);
  LOCO_TRY(require_device());
  cudaStream_t s = ST(stream);
  View xv = make_view(const_cast<float*>(xp), 1, H, W, C);
  View gyv = make_view(const_cast<float*>(gy), K, H, W, C);
  View gxv = make_view(gx, K, H, W, C);
  double* ps = reinterpret_cast<double*>(stats);
  double* bs = ps + 64;
  LOCO_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 64 * (K + 1), s));
  LOCO_TRY(gn_stats_fwd(xv, 1, ps, s));
  LOCO_TRY(gn_stats_vjp(xv, ps, gyv, gamma, beta, eps, silu, bs, s));
  return gn_apply_vjp(xv, ps, gyv, bs, gamma, beta, eps, silu, nullptr, 0, 0, gxv, s);
}
// Typed variants (half = 1: fp16 tensors, the storage type of the default programs) with the two
// passes selectable (stages bit 0: statistics (+ their memset), bit 1: apply), so that each
// bandwidth-bound kernel can be checked and timed on its own.
int loco_groupnorm_silu_fwd_ex(const void* x, int half, int N, int H, int W, int C, int n_primal,
                               const float* gamma, const float* beta, float eps, int silu, void* y,
                               void* stats, int stages, void* stream) {
  ON_DEVICE_OF(x);
  LOCO_TRY(require_device());
  cudaStream_t s = ST(stream);
  View xv = make_view(reinterpret_cast<float*>(const_cast<void*>(x)), N, H, W, C, half);
  View yv = make_view(reinterpret_cast<float*>(y), N, H, W, C, half);
  if (stages & 4)     // the one-launch kernel of the small sites (statistics + apply, a block per (row, group))
    return gn_small_fwd(xv, n_primal, reinterpret_cast<double*>(stats), gamma, beta, eps, silu, 0, yv, s);
  if (stages & 1) {
    LOCO_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 64 * N, s));
    LOCO_TRY(gn_stats_fwd(xv, n_primal, reinterpret_cast<double*>(stats), s));
  }
  if (stages & 2)
    LOCO_TRY(gn_apply_fwd(xv, n_primal, reinterpret_cast<double*>(stats), gamma, beta, eps, silu, 0, yv, s));
  return 0;
}
int loco_groupnorm_silu_vjp_ex(const void* xp, int half, int H, int W, int C, const void* gy, int K,
                               const float* gamma, const float* beta, float eps, int silu,
                               const void* addend, int accumulate, void* gx, void* stats, int stages,
                               void* stream) {
  ON_DEVICE_OF(gy);
  LOCO_TRY(require_device());
  cudaStream_t s = ST(stream);
  View xv = make_view(reinterpret_cast<float*>(const_cast<void*>(xp)), 1, H, W, C, half);
  View gyv = make_view(reinterpret_cast<float*>(const_cast<void*>(gy)), K, H, W, C, half);
  View gxv = make_view(reinterpret_cast<float*>(gx), K, H, W, C, half);
  View addv = make_view(reinterpret_cast<float*>(const_cast<void*>(addend)), K, H, W, C, half);
  double* ps = reinterpret_cast<double*>(stats);
  double* bs = ps + 64;
  if (stages & 4) {   // one launch for the primal statistics, one for statistics + apply of the cotangent rows
    LOCO_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 64, s));
    LOCO_TRY(gn_stats_fwd(xv, 1, ps, s));
    return gn_small_vjp(xv, ps, gyv, gamma, beta, eps, silu, addend ? &addv : nullptr, accumulate, 0, gxv, s);
  }
  if (stages & 1) {
    LOCO_CHECK_CUDA(cudaMemsetAsync(stats, 0, sizeof(double) * 64 * (K + 1), s));
    LOCO_TRY(gn_stats_fwd(xv, 1, ps, s));
    LOCO_TRY(gn_stats_vjp(xv, ps, gyv, gamma, beta, eps, silu, bs, s));
  }
  if (stages & 2)
    LOCO_TRY(gn_apply_vjp(xv, ps, gyv, bs, gamma, beta, eps, silu, addend ? &addv : nullptr, accumulate, 0,
                          gxv, s));
  return 0;
}
int loco_attention_fwd(const float* qkv, int N, int T, int C, int n_primal, int head_ch, float* S,
                       float* o, void* stream) {
  ON_DEVICE_OF(qkv);
  LOCO_TRY(require_device());
  View q = make_view(const_cast<float*>(qkv), N, 1, T, 3 * C);
  View ov = make_view(o, N, 1, T, C);
  return attention_forward(q, n_primal, head_ch, S, ov, ST(stream));
}
int loco_cross_attention_fwd(const float* q, int N, int Tq, int C, int n_primal, const float* kv, int Tk,
                             int Tk_valid, int heads, float* S, float* o, void* stream) {
  ON_DEVICE_OF(q);
  LOCO_TRY(require_device());
  View qv = make_view(const_cast<float*>(q), N, 1, Tq, C);
  View kvv = make_view(const_cast<float*>(kv), 1, 1, Tk, 2 * C);
  View ov = make_view(o, N, 1, Tq, C);
  return attention_cross_forward_tc(qv, kvv, n_primal, heads, Tk_valid, S, ov, ST(stream));
}
int loco_cross_attention_vjp(const float* go, int K, int Tq, int C, const float* kv, int Tk, int Tk_valid,
                             int heads, const float* P0, float* gq, void* stream) {
  ON_DEVICE_OF(go);
  LOCO_TRY(require_device());
  View g = make_view(const_cast<float*>(go), K, 1, Tq, C);
  View kvv = make_view(const_cast<float*>(kv), 1, 1, Tk, 2 * C);
  View gqv = make_view(gq, K, 1, Tq, C);
  return attention_cross_vjp_tc(g, kvv, heads, Tk_valid, P0, gqv, ST(stream));
}
int loco_attention_vjp(const float* go, int K, int T, int C, int head_ch, const float* qkv0,
                       const float* P0, float* gP, float* gqkv, void* stream) {
  ON_DEVICE_OF(go);
  LOCO_TRY(require_device());
  View g = make_view(const_cast<float*>(go), K, 1, T, C);
  View q0 = make_view(const_cast<float*>(qkv0), 1, 1, T, 3 * C);
  View gq = make_view(gqkv, K, 1, T, 3 * C);
  return attention_vjp(g, q0, head_ch, P0, gP, gq, ST(stream));
}

}  // extern "C"

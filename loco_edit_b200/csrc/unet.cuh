// U-Net executor for the two unconditional model families of the reference: the DDPM U-Net
// (models/ddpm/diffusion.py:22-200, i.e. google/ddpm-ema-celebahq-256) and the P2 / guided-diffusion
// U-Net (models/guided_diffusion/unet.py:398-684 with P2_DICT).  Forward, fused primal+k-tangent
// forward (JVP) and k-cotangent backward (VJP) as static launch programs over a caller-provided
// workspace.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>
#include "attention.cuh"
#include "common.cuh"
#include "conv_gemm.cuh"
#include "layers.cuh"

namespace loco {

struct Arch {
  int ch = 128;
  int n_levels = 6;
  int ch_mult[8] = {1, 1, 2, 2, 4, 4, 0, 0};
  int num_res_blocks = 2;
  int n_attn = 1;
  int attn_resolutions[4] = {16, 0, 0, 0};
  int resolution = 256;
  int in_ch = 3;
  int out_ch = 3;
  float gn_eps = 1e-6f;
  int kind = 0;       // 0 = DDPM (ddpm/diffusion.py), 1 = P2 / guided diffusion (guided_diffusion/unet.py),
                      // 2 = VAE decoder (latent [in_ch, R, R] -> image [out_ch, R << (n_levels-1), ...]; no timestep)
  int head_ch = 0;    // kind 1: channels per attention head
  // kind 0 only: > 0 adds a cross-attention sub-block (GroupNorm, q projection, softmax(q K_c^T) V_c,
  // output projection, residual) to every AttnBlock; K_c | V_c are linear maps of a prompt embedding
  // [n_tok <= 128, ctx_dim] (text-conditioned twins of the path)
  int ctx_dim = 0;
  int ctx_heads = 1;
};
constexpr int kCtxPad = 128;   // context rows of the key / value buffers (tokens are zero-padded)

// One named parameter of the reference state_dict and where its packed forms live in the arena.
struct ParamSlot {
  std::string name;
  std::vector<int> shape;
  long long numel;
  // CONV_THIN: a 3x3 conv with <= 4 channels on one side, packed like CONV_GEMM with that side
  // zero-padded to kThinPad channels (pad_cin / pad_cout are the padded extents)
  enum Kind { RAW, CONV_GEMM, CONV_QKV, CONV_EDGE_IN, CONV_EDGE_OUT, BIAS_QKV, CONV_THIN } kind;
  int pad_cin = 0, pad_cout = 0;
  size_t off_a = 0;   // RAW: copy; CONV_GEMM/QKV: fprop pack; EDGE: [9][3][C] pack
  size_t off_b = 0;   // CONV_GEMM/QKV: dgrad pack
  size_t off_a16 = 0, off_b16 = 0;   // CONV_GEMM/QKV: fp16 copies of the two packs
  int cout = 0, cin = 0, ksz = 0;
  int row_off = 0, rows_total = 0;   // QKV fusion: rows [row_off, row_off+cout) of rows_total
  bool loaded = false;
};

// wf / wd: fprop / dgrad GEMM operands (tf32-rounded fp32); wf16 / wd16: the same in fp16 (offsets
// in floats, half the length) for plans that keep their activations in fp16
struct ConvRef { int cin = 0, cout = 0, ksz = 0; size_t wf = 0, wd = 0, bias = 0, wf16 = 0, wd16 = 0; };
struct NormRef { int C = 0; size_t gamma = 0, beta = 0; };
struct ResRef {
  int cin = 0, cout = 0;
  NormRef n1, n2;
  ConvRef c1, c2, nin;
  bool has_nin = false;
  int temb_off = 0;    // offset into the packed timestep-projection vector (-1: no timestep input, VAE decoder)
  // guided-diffusion ResBlock (unet.py:161-258): timestep projection is (scale | shift) of the
  // second GroupNorm instead of a bias; resample 1 = avg-pool /2, 2 = nearest x2 on both branches
  bool scale_shift = false;
  int resample = 0;
};
struct AttnRef {
  int C = 0; NormRef n; ConvRef qkv, proj;
  bool cross = false;            // cross-attention sub-block after the self-attention one
  NormRef n2; ConvRef q2, proj2;
  size_t kv_w = 0, kv_b = 0;     // Linear(ctx_dim -> 2C): [2C][ctx_dim] fp32, [2C]
};

class Model {
 public:
  explicit Model(const Arch& a);
  Arch arch;
  std::vector<ParamSlot> slots;
  std::map<std::string, int> slot_index;
  size_t arena_floats = 0;
  float* arena = nullptr;

  // layer references (offsets into the arena)
  size_t temb_w0 = 0, temb_b0 = 0, temb_w1 = 0, temb_b1 = 0;
  size_t tproj_w = 0, tproj_b = 0;   // all temb_proj layers stacked: [tproj_rows][temb_ch]
  int tproj_rows = 0;
  size_t conv_in_w = 0, conv_in_b = 0, conv_out_w = 0, conv_out_b = 0;
  // thin edges (arch.in_ch / out_ch != 3, and the VAE decoder): tensor-core convs over a padded side
  bool thin = false;
  ConvRef thin_in, thin_out;
  size_t pq_mix = 0;                 // kind 2: post_quant_conv [c*c weight | c bias]
  bool has_pq = false;
  NormRef norm_out;
  std::vector<std::vector<ResRef>> down_res, up_res;
  std::vector<std::vector<AttnRef>> down_attn, up_attn;
  std::vector<ConvRef> down_sample, up_sample;   // per level (cin == 0 when absent), kind 0
  std::vector<ResRef> down_rb, up_rb;            // per level: resampling ResBlocks, kind 1
  ResRef mid1, mid2;
  AttnRef mid_attn;

  int load_param(const char* name, const float* dev_src, long long numel, cudaStream_t s);
  int check_loaded() const;
  float* w(size_t off) const { return arena + off; }

 private:
  size_t alloc(size_t n);
  int add_slot(ParamSlot s);
  void build_ddpm();
  void build_p2();
  void build_decoder();
  ConvRef add_thin_conv(const std::string& prefix, int cin, int cout);
  void finish_temb();
  ConvRef add_conv(const std::string& prefix, int cin, int cout, int ksz, bool conv1d = false);
  ResRef add_res_p2(const std::string& prefix, int cin, int cout, int resample);
  AttnRef add_attn_p2(const std::string& prefix, int C);
  NormRef add_norm(const std::string& prefix, int C);
  ResRef add_res(const std::string& prefix, int cin, int cout, bool temb = true);
  AttnRef add_attn(const std::string& prefix, int C);
};

class Plan {
 public:
  // flags bit 0: activations stored in fp16 and GEMM layers on tcgen05 kind::f16 (primal-only
  // programs: the DDIM inversion / denoising loops); 0: fp32 storage, kind::tf32
  Plan(const Model* m, int n_primal, int n_tangent, int n_cot, int flags = 0)
      : model(m), NP(n_primal), NT(n_tangent), NC(n_cot), act16(flags & 1) {}
  const Model* model;
  int NP, NT, NC;
  int act16;
  size_t workspace_floats = 0;
  float* base = nullptr;
  int device = -1;    // device owning the bound workspace (every launch of the plan runs there)
  double fwd_flops = 0, vjp_flops = 0;
  int fwd_launches = 0, vjp_launches = 0;

  int build(float* workspace);   // workspace == nullptr: size query only
  // x: [NP+NT, in_ch, R, R] (row 0.. primal samples, then tangents); eps: [NP+NT, out_ch, Ro, Ro]
  // (Ro = R for the U-Nets, R << (n_levels-1) for the VAE decoder).
  long long in_elems() const;    // per row
  long long out_elems() const;
  int forward(const float* x_nchw, float t, float* eps_nchw, cudaStream_t s);
  // conditioning embedding [4 ch] (device, or null = none) added to the timestep embedding of every
  // following forward(); x -> eps(x, t, c) stays a function of x only, so JVP / VJP are unchanged
  int set_condition(const float* cond, cudaStream_t s);
  // prompt embedding [n_tok, ctx_dim] (device) of the cross-attention layers: computes every layer's
  // K_c | V_c once; like the conditioning embedding it is a constant of the Jacobian passes
  int set_context(const float* ctx, int n_tok, cudaStream_t s);
  // g_eps: [NC, 3, R, R] cotangents of eps; gx: [NC, 3, R, R] = J_eps^T g_eps.
  int vjp(const float* g_eps_nchw, float* gx_nchw, cudaStream_t s);

  struct Impl;
  std::shared_ptr<Impl> impl;
};

}  // namespace loco

// DDPM U-Net executor: builds static forward / JVP / VJP launch programs (see unet.cuh).
//
// Data layout in HBM
//   * every activation is channels-last fp32 [N, H, W, C]; N stacks the primal rows and the k probe
//     tangents (forward/JVP program) or the k cotangents (VJP program);
//   * the skip-connection concat of the decoder is never materialised: each up ResnetBlock owns one
//     buffer [N, H, W, C_dec + C_skip]; the encoder writes its skip tensor straight into the
//     channel slice, the decoder writes the other slice; cotangents use the same geometry;
//   * all buffers live in one caller-provided workspace; every op output has its own region (the
//     primal rows are what the VJP program re-reads, 180 GB of HBM make recycling unnecessary).
#include "unet.cuh"
#include <deque>
#include <functional>
#include <stdarg.h>

namespace loco {

// ------------------------------------------------------------------------------------------------
// error plumbing / device info
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }

static long long g_launches = 0;
void count_launch(int n) { g_launches += n; }
long long launch_count() { return g_launches; }

namespace {
struct ProfRec { cudaEvent_t a, b; int family; double work; };
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
}  // namespace
bool profile_is_on() { return g_prof_on; }
void profile_enable(bool on) {
  for (auto& r : g_prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  g_prof.clear();
  g_prof_on = on;
}
ProfScope::ProfScope(int fam, double work, cudaStream_t s) : family(fam), stream(s), slot(-1) {
  if (!g_prof_on) return;
  ProfRec r; r.family = fam; r.work = work;
  if (cudaEventCreate(&r.a) != cudaSuccess || cudaEventCreate(&r.b) != cudaSuccess) return;
  cudaEventRecord(r.a, s);
  g_prof.push_back(r);
  slot = (int)g_prof.size() - 1;
}
ProfScope::~ProfScope() {
  if (slot >= 0) cudaEventRecord(g_prof[slot].b, stream);
}
int profile_collect(double* ms, double* work, long long* launches, int nfam) {
  LOCO_CHECK_CUDA(cudaDeviceSynchronize());
  for (int i = 0; i < nfam; ++i) { ms[i] = 0; work[i] = 0; launches[i] = 0; }
  for (auto& r : g_prof) {
    float t = 0.f;
    if (cudaEventElapsedTime(&t, r.a, r.b) != cudaSuccess) continue;
    if (r.family < nfam) { ms[r.family] += t; work[r.family] += r.work; launches[r.family] += 1; }
  }
  return 0;
}

int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); return 0; }
  return dev;
}
int pointer_device(const void* p) {
  if (p == nullptr) return -1;
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return -1; }
  if (a.type != cudaMemoryTypeDevice && a.type != cudaMemoryTypeManaged) return -1;
  return a.device;
}
DeviceGuard::DeviceGuard(int dev) {
  if (dev < 0) return;
  if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; return; }
  if (prev != dev && cudaSetDevice(dev) == cudaSuccess) switched = true;
}
DeviceGuard::~DeviceGuard() {
  if (switched) cudaSetDevice(prev);
}

int num_sms() {
  static int n[kMaxDevices] = {0};
  const int dev = current_device();
  int& v = n[dev < kMaxDevices ? dev : 0];
  if (v == 0) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
  }
  return v;
}

// ------------------------------------------------------------------------------------------------
// Model: parameter registry mirroring the reference module tree (ddpm/diffusion.py:24-126)
// ------------------------------------------------------------------------------------------------
size_t Model::alloc(size_t n) {
  const size_t off = arena_floats;
  arena_floats += (n + 63) & ~(size_t)63;
  return off;
}
int Model::add_slot(ParamSlot s) {
  s.numel = 1;
  for (int d : s.shape) s.numel *= d;
  slot_index[s.name] = (int)slots.size();
  slots.push_back(s);
  return (int)slots.size() - 1;
}
ConvRef Model::add_conv(const std::string& prefix, int cin, int cout, int ksz, bool conv1d) {
  ConvRef c;
  c.cin = cin; c.cout = cout; c.ksz = ksz;
  c.wf = alloc((size_t)cout * cin * ksz * ksz);
  c.wd = alloc((size_t)cout * cin * ksz * ksz);
  c.wf16 = alloc(((size_t)cout * cin * ksz * ksz + 1) / 2);
  c.wd16 = alloc(((size_t)cout * cin * ksz * ksz + 1) / 2);
  c.bias = alloc(cout);
  ParamSlot w;
  w.name = prefix + ".weight"; w.kind = ParamSlot::CONV_GEMM;
  if (conv1d) w.shape = {cout, cin, ksz}; else w.shape = {cout, cin, ksz, ksz};
  w.off_a = c.wf; w.off_b = c.wd; w.off_a16 = c.wf16; w.off_b16 = c.wd16;
  w.cout = cout; w.cin = cin; w.ksz = ksz;
  w.row_off = 0; w.rows_total = cout;
  add_slot(w);
  ParamSlot b;
  b.name = prefix + ".bias"; b.shape = {cout}; b.kind = ParamSlot::RAW; b.off_a = c.bias;
  add_slot(b);
  return c;
}
NormRef Model::add_norm(const std::string& prefix, int C) {
  NormRef n;
  n.C = C; n.gamma = alloc(C); n.beta = alloc(C);
  ParamSlot g; g.name = prefix + ".weight"; g.shape = {C}; g.kind = ParamSlot::RAW; g.off_a = n.gamma;
  add_slot(g);
  ParamSlot b; b.name = prefix + ".bias"; b.shape = {C}; b.kind = ParamSlot::RAW; b.off_a = n.beta;
  add_slot(b);
  return n;
}
ResRef Model::add_res(const std::string& prefix, int cin, int cout, bool temb) {
  ResRef r;
  r.cin = cin; r.cout = cout;
  r.n1 = add_norm(prefix + ".norm1", cin);
  r.c1 = add_conv(prefix + ".conv1", cin, cout, 3);
  r.temb_off = temb ? tproj_rows : -1;
  if (temb) {
    tproj_rows += cout;
    // rows of the stacked projection matrix; arena offsets are fixed up after construction
    ParamSlot w; w.name = prefix + ".temb_proj.weight"; w.shape = {cout, arch.ch * 4};
    w.kind = ParamSlot::RAW; w.off_a = (size_t)r.temb_off; w.cout = -1;   // marker: temb matrix row
    add_slot(w);
    ParamSlot b; b.name = prefix + ".temb_proj.bias"; b.shape = {cout}; b.kind = ParamSlot::RAW;
    b.off_a = (size_t)r.temb_off; b.cout = -2;                            // marker: temb bias row
    add_slot(b);
  }
  r.n2 = add_norm(prefix + ".norm2", cout);
  r.c2 = add_conv(prefix + ".conv2", cout, cout, 3);
  r.has_nin = cin != cout;
  if (r.has_nin) r.nin = add_conv(prefix + ".nin_shortcut", cin, cout, 1);
  return r;
}
AttnRef Model::add_attn(const std::string& prefix, int C) {
  AttnRef a;
  a.C = C;
  a.n = add_norm(prefix + ".norm", C);
  a.qkv.cin = C; a.qkv.cout = 3 * C; a.qkv.ksz = 1;
  a.qkv.wf = alloc((size_t)3 * C * C);
  a.qkv.wd = alloc((size_t)3 * C * C);
  a.qkv.wf16 = alloc(((size_t)3 * C * C + 1) / 2);
  a.qkv.wd16 = alloc(((size_t)3 * C * C + 1) / 2);
  a.qkv.bias = alloc(3 * C);
  const char* names[3] = {"q", "k", "v"};
  for (int i = 0; i < 3; ++i) {
    ParamSlot w;
    w.name = prefix + "." + names[i] + ".weight"; w.shape = {C, C, 1, 1}; w.kind = ParamSlot::CONV_QKV;
    w.off_a = a.qkv.wf; w.off_b = a.qkv.wd; w.off_a16 = a.qkv.wf16; w.off_b16 = a.qkv.wd16;
    w.cout = C; w.cin = C; w.ksz = 1;
    w.row_off = i * C; w.rows_total = 3 * C;
    add_slot(w);
    ParamSlot b;
    b.name = prefix + "." + names[i] + ".bias"; b.shape = {C}; b.kind = ParamSlot::BIAS_QKV;
    b.off_a = a.qkv.bias; b.row_off = i * C;
    add_slot(b);
  }
  a.proj = add_conv(prefix + ".proj_out", C, C, 1);
  if (arch.ctx_dim > 0) {
    a.cross = true;
    a.n2 = add_norm(prefix + ".norm2", C);
    a.q2 = add_conv(prefix + ".q2", C, C, 1);
    a.kv_w = alloc((size_t)2 * C * arch.ctx_dim);
    a.kv_b = alloc(2 * C);
    ParamSlot w; w.name = prefix + ".kv2.weight"; w.shape = {2 * C, arch.ctx_dim}; w.kind = ParamSlot::RAW; w.off_a = a.kv_w;
    add_slot(w);
    ParamSlot b; b.name = prefix + ".kv2.bias"; b.shape = {2 * C}; b.kind = ParamSlot::RAW; b.off_a = a.kv_b;
    add_slot(b);
    a.proj2 = add_conv(prefix + ".proj_out2", C, C, 1);
  }
  return a;
}

// 3x3 conv with a thin (<= 4 channel) side: GEMM packs over that side zero-padded to kThinPad
ConvRef Model::add_thin_conv(const std::string& prefix, int cin, int cout) {
  const int pcin = cin <= 4 ? kThinPad : cin, pcout = cout <= 4 ? kThinPad : cout;
  ConvRef c;
  c.cin = pcin; c.cout = pcout; c.ksz = 3;
  const size_t n = (size_t)pcout * pcin * 9;
  c.wf = alloc(n); c.wd = alloc(n); c.wf16 = alloc((n + 1) / 2); c.wd16 = alloc((n + 1) / 2);
  c.bias = alloc(pcout);
  ParamSlot w;
  w.name = prefix + ".weight"; w.kind = ParamSlot::CONV_THIN; w.shape = {cout, cin, 3, 3};
  w.off_a = c.wf; w.off_b = c.wd; w.off_a16 = c.wf16; w.off_b16 = c.wd16;
  w.cout = cout; w.cin = cin; w.ksz = 3; w.pad_cin = pcin; w.pad_cout = pcout;
  add_slot(w);
  ParamSlot b;
  b.name = prefix + ".bias"; b.shape = {cout}; b.kind = ParamSlot::RAW; b.off_a = c.bias;
  add_slot(b);
  return c;
}

// VAE decoder of a latent-diffusion model: diffusers' AutoencoderKL.decode (the `self.vae.decode` of
// src/modules/edit.py:770) = Decoder(post_quant_conv(z)).  Module tree and parameter names of the
// CompVis `Decoder` that AutoencoderKL restates (same ResnetBlock / AttnBlock / Upsample code as
// src/models/ddpm/diffusion.py:816-966, without timestep input): conv_in, mid.{block_1, attn_1, block_2},
// up.{l}.block.{0..nrb} (+ up.{l}.upsample.conv for l > 0), norm_out, conv_out.
void Model::build_decoder() {
  const Arch& a = arch;
  const int ch = a.ch, L = a.n_levels;
  thin = true;
  has_pq = true;
  pq_mix = alloc((size_t)a.in_ch * a.in_ch + a.in_ch);
  {
    ParamSlot s; s.kind = ParamSlot::RAW;
    s.name = "post_quant_conv.weight"; s.shape = {a.in_ch, a.in_ch, 1, 1}; s.off_a = pq_mix; add_slot(s);
    s.name = "post_quant_conv.bias"; s.shape = {a.in_ch}; s.off_a = pq_mix + (size_t)a.in_ch * a.in_ch; add_slot(s);
  }
  int block_in = ch * a.ch_mult[L - 1];
  thin_in = add_thin_conv("decoder.conv_in", a.in_ch, block_in);
  mid1 = add_res("decoder.mid.block_1", block_in, block_in, false);
  mid_attn = add_attn("decoder.mid.attn_1", block_in);
  mid2 = add_res("decoder.mid.block_2", block_in, block_in, false);
  up_res.resize(L); up_attn.resize(L); up_sample.resize(L);
  down_res.resize(L); down_attn.resize(L); down_sample.resize(L);
  for (int l = L - 1; l >= 0; --l) {
    const int block_out = ch * a.ch_mult[l];
    for (int b = 0; b < a.num_res_blocks + 1; ++b) {
      up_res[l].push_back(add_res("decoder.up." + std::to_string(l) + ".block." + std::to_string(b), block_in, block_out, false));
      block_in = block_out;
    }
    if (l != 0) up_sample[l] = add_conv("decoder.up." + std::to_string(l) + ".upsample.conv", block_in, block_in, 3);
  }
  norm_out = add_norm("decoder.norm_out", block_in);
  thin_out = add_thin_conv("decoder.conv_out", block_in, a.out_ch);
}

static bool in_list(const int* lst, int n, int v) {
  for (int i = 0; i < n; ++i)
    if (lst[i] == v) return true;
  return false;
}

// guided-diffusion ResBlock (unet.py:161-236): in_layers = GN, SiLU, conv3x3; emb_layers = SiLU,
// Linear(4ch -> 2 Cout); out_layers = GN, SiLU, Dropout, conv3x3; skip_connection = 1x1 iff Cin != Cout
ResRef Model::add_res_p2(const std::string& prefix, int cin, int cout, int resample) {
  ResRef r;
  r.cin = cin; r.cout = cout; r.scale_shift = true; r.resample = resample;
  r.n1 = add_norm(prefix + ".in_layers.0", cin);
  r.c1 = add_conv(prefix + ".in_layers.2", cin, cout, 3);
  r.temb_off = tproj_rows;
  tproj_rows += 2 * cout;
  {
    ParamSlot w; w.name = prefix + ".emb_layers.1.weight"; w.shape = {2 * cout, arch.ch * 4};
    w.kind = ParamSlot::RAW; w.off_a = (size_t)r.temb_off; w.cout = -1;
    add_slot(w);
    ParamSlot b; b.name = prefix + ".emb_layers.1.bias"; b.shape = {2 * cout}; b.kind = ParamSlot::RAW;
    b.off_a = (size_t)r.temb_off; b.cout = -2;
    add_slot(b);
  }
  r.n2 = add_norm(prefix + ".out_layers.0", cout);
  r.c2 = add_conv(prefix + ".out_layers.3", cout, cout, 3);
  r.has_nin = cin != cout;
  if (r.has_nin) r.nin = add_conv(prefix + ".skip_connection", cin, cout, 1);
  return r;
}
// guided-diffusion AttentionBlock (unet.py:261-308): norm, qkv = conv1d(C, 3C, 1) in the per-head
// q|k|v ("legacy") channel order, proj_out = conv1d(C, C, 1)
AttnRef Model::add_attn_p2(const std::string& prefix, int C) {
  AttnRef a;
  a.C = C;
  a.n = add_norm(prefix + ".norm", C);
  a.qkv = add_conv(prefix + ".qkv", C, 3 * C, 1, true);
  a.proj = add_conv(prefix + ".proj_out", C, C, 1, true);
  return a;
}

void Model::finish_temb() {
  // stacked timestep projections
  const int temb_ch = 4 * arch.ch;
  tproj_w = alloc((size_t)tproj_rows * temb_ch);
  tproj_b = alloc(tproj_rows);
  for (auto& s : slots) {
    if (s.kind == ParamSlot::RAW && s.cout == -1) s.off_a = tproj_w + s.off_a * (size_t)temb_ch;
    if (s.kind == ParamSlot::RAW && s.cout == -2) s.off_a = tproj_b + s.off_a;
  }
}

Model::Model(const Arch& a) : arch(a) {
  if (a.kind == 2) build_decoder(); else if (a.kind == 1) build_p2(); else build_ddpm();
  finish_temb();
}

// Module tree of UNetModel.__init__ (guided_diffusion/unet.py:470-618) with resblock_updown = True,
// use_scale_shift_norm = True, learn_sigma = True.  Same skeleton as the DDPM U-Net (one skip tensor
// per block, nrb+1 decoder blocks per level), so the two families share the plan builder.
void Model::build_p2() {
  const Arch& a = arch;
  const int ch = a.ch, temb_ch = 4 * a.ch, L = a.n_levels;
  temb_w0 = alloc((size_t)temb_ch * ch); temb_b0 = alloc(temb_ch);
  temb_w1 = alloc((size_t)temb_ch * temb_ch); temb_b1 = alloc(temb_ch);
  {
    ParamSlot s;
    s.kind = ParamSlot::RAW;
    s.name = "time_embed.0.weight"; s.shape = {temb_ch, ch}; s.off_a = temb_w0; add_slot(s);
    s.name = "time_embed.0.bias"; s.shape = {temb_ch}; s.off_a = temb_b0; add_slot(s);
    s.name = "time_embed.2.weight"; s.shape = {temb_ch, temb_ch}; s.off_a = temb_w1; add_slot(s);
    s.name = "time_embed.2.bias"; s.shape = {temb_ch}; s.off_a = temb_b1; add_slot(s);
  }
  conv_in_w = alloc(27 * ch); conv_in_b = alloc(ch);
  {
    ParamSlot s;
    s.name = "input_blocks.0.0.weight"; s.shape = {ch, a.in_ch, 3, 3}; s.kind = ParamSlot::CONV_EDGE_IN;
    s.off_a = conv_in_w; s.cout = ch; add_slot(s);
    ParamSlot b; b.name = "input_blocks.0.0.bias"; b.shape = {ch}; b.kind = ParamSlot::RAW;
    b.off_a = conv_in_b; add_slot(b);
  }
  down_res.resize(L); down_attn.resize(L); down_rb.resize(L);
  up_res.resize(L); up_attn.resize(L); up_rb.resize(L);
  down_sample.resize(L); up_sample.resize(L);
  int curr_res = a.resolution;
  int block_in = ch * a.ch_mult[0];
  int ib = 1;
  for (int l = 0; l < L; ++l) {
    const int block_out = ch * a.ch_mult[l];
    for (int b = 0; b < a.num_res_blocks; ++b, ++ib) {
      const std::string p = "input_blocks." + std::to_string(ib);
      down_res[l].push_back(add_res_p2(p + ".0", block_in, block_out, 0));
      block_in = block_out;
      if (in_list(a.attn_resolutions, a.n_attn, curr_res))
        down_attn[l].push_back(add_attn_p2(p + ".1", block_in));
    }
    if (l != L - 1) {
      down_rb[l] = add_res_p2("input_blocks." + std::to_string(ib) + ".0", block_in, block_in, 1);
      ++ib;
      curr_res /= 2;
    }
  }
  mid1 = add_res_p2("middle_block.0", block_in, block_in, 0);
  mid_attn = add_attn_p2("middle_block.1", block_in);
  mid2 = add_res_p2("middle_block.2", block_in, block_in, 0);
  int ob = 0;
  for (int l = L - 1; l >= 0; --l) {
    const int block_out = ch * a.ch_mult[l];
    int skip_in = ch * a.ch_mult[l];
    for (int b = 0; b < a.num_res_blocks + 1; ++b, ++ob) {
      if (b == a.num_res_blocks) skip_in = ch * (l == 0 ? a.ch_mult[0] : a.ch_mult[l - 1]);
      const std::string p = "output_blocks." + std::to_string(ob);
      up_res[l].push_back(add_res_p2(p + ".0", block_in + skip_in, block_out, 0));
      block_in = block_out;
      int j = 1;
      if (in_list(a.attn_resolutions, a.n_attn, curr_res)) {
        up_attn[l].push_back(add_attn_p2(p + "." + std::to_string(j), block_in));
        ++j;
      }
      if (l != 0 && b == a.num_res_blocks) {
        up_rb[l] = add_res_p2(p + "." + std::to_string(j), block_in, block_in, 2);
        curr_res *= 2;
      }
    }
  }
  norm_out = add_norm("out.0", block_in);
  conv_out_w = alloc(27 * block_in); conv_out_b = alloc(64);
  {
    // learn_sigma: 6 output channels, eps = the first 3 (unet.py:680-684); only those are packed
    ParamSlot s;
    s.name = "out.2.weight"; s.shape = {2 * a.out_ch, block_in, 3, 3}; s.kind = ParamSlot::CONV_EDGE_OUT;
    s.off_a = conv_out_w; s.cin = block_in; add_slot(s);
    ParamSlot b; b.name = "out.2.bias"; b.shape = {2 * a.out_ch}; b.kind = ParamSlot::RAW;
    b.off_a = conv_out_b; add_slot(b);
  }
}

void Model::build_ddpm() {
  const Arch& a = arch;
  const int ch = a.ch, temb_ch = 4 * a.ch, L = a.n_levels;
  temb_w0 = alloc((size_t)temb_ch * ch); temb_b0 = alloc(temb_ch);
  temb_w1 = alloc((size_t)temb_ch * temb_ch); temb_b1 = alloc(temb_ch);
  {
    ParamSlot s;
    s.kind = ParamSlot::RAW;
    s.name = "temb.dense.0.weight"; s.shape = {temb_ch, ch}; s.off_a = temb_w0; add_slot(s);
    s.name = "temb.dense.0.bias"; s.shape = {temb_ch}; s.off_a = temb_b0; add_slot(s);
    s.name = "temb.dense.1.weight"; s.shape = {temb_ch, temb_ch}; s.off_a = temb_w1; add_slot(s);
    s.name = "temb.dense.1.bias"; s.shape = {temb_ch}; s.off_a = temb_b1; add_slot(s);
  }
  thin = a.in_ch != 3 || a.out_ch != 3;     // latent-space U-Net (4 channels): padded tensor-core edges
  if (thin) {
    thin_in = add_thin_conv("conv_in", a.in_ch, ch);
  } else {
    conv_in_w = alloc(27 * ch); conv_in_b = alloc(ch);
    ParamSlot s;
    s.name = "conv_in.weight"; s.shape = {ch, a.in_ch, 3, 3}; s.kind = ParamSlot::CONV_EDGE_IN;
    s.off_a = conv_in_w; s.cout = ch; add_slot(s);
    ParamSlot b; b.name = "conv_in.bias"; b.shape = {ch}; b.kind = ParamSlot::RAW; b.off_a = conv_in_b;
    add_slot(b);
  }
  int curr_res = a.resolution;
  int block_in = ch;
  down_res.resize(L); down_attn.resize(L); down_sample.resize(L);
  up_res.resize(L); up_attn.resize(L); up_sample.resize(L);
  for (int l = 0; l < L; ++l) {
    block_in = ch * (l == 0 ? 1 : a.ch_mult[l - 1]);
    const int block_out = ch * a.ch_mult[l];
    for (int b = 0; b < a.num_res_blocks; ++b) {
      const std::string p = "down." + std::to_string(l) + ".block." + std::to_string(b);
      down_res[l].push_back(add_res(p, block_in, block_out));
      block_in = block_out;
      if (in_list(a.attn_resolutions, a.n_attn, curr_res))
        down_attn[l].push_back(
            add_attn("down." + std::to_string(l) + ".attn." + std::to_string(b), block_in));
    }
    if (l != L - 1) {
      down_sample[l] = add_conv("down." + std::to_string(l) + ".downsample.conv", block_in, block_in, 3);
      curr_res /= 2;
    }
  }
  mid1 = add_res("mid.block_1", block_in, block_in);
  mid_attn = add_attn("mid.attn_1", block_in);
  mid2 = add_res("mid.block_2", block_in, block_in);
  for (int l = L - 1; l >= 0; --l) {
    const int block_out = ch * a.ch_mult[l];
    int skip_in = ch * a.ch_mult[l];
    for (int b = 0; b < a.num_res_blocks + 1; ++b) {
      if (b == a.num_res_blocks) skip_in = ch * (l == 0 ? 1 : a.ch_mult[l - 1]);
      const std::string p = "up." + std::to_string(l) + ".block." + std::to_string(b);
      up_res[l].push_back(add_res(p, block_in + skip_in, block_out));
      block_in = block_out;
      if (in_list(a.attn_resolutions, a.n_attn, curr_res))
        up_attn[l].push_back(
            add_attn("up." + std::to_string(l) + ".attn." + std::to_string(b), block_in));
    }
    if (l != 0) {
      up_sample[l] = add_conv("up." + std::to_string(l) + ".upsample.conv", block_in, block_in, 3);
      curr_res *= 2;
    }
  }
  norm_out = add_norm("norm_out", block_in);
  if (thin) {
    thin_out = add_thin_conv("conv_out", block_in, a.out_ch);
  } else {
    conv_out_w = alloc(27 * block_in); conv_out_b = alloc(64);
    ParamSlot s;
    s.name = "conv_out.weight"; s.shape = {a.out_ch, block_in, 3, 3}; s.kind = ParamSlot::CONV_EDGE_OUT;
    s.off_a = conv_out_w; s.cin = block_in; add_slot(s);
    ParamSlot b; b.name = "conv_out.bias"; b.shape = {a.out_ch}; b.kind = ParamSlot::RAW;
    b.off_a = conv_out_b; add_slot(b);
  }
}

int Model::load_param(const char* name, const float* src, long long numel, cudaStream_t s) {
  LOCO_REQUIRE(arena != nullptr, "load_param: weight arena not bound");
  auto it = slot_index.find(name);
  LOCO_REQUIRE(it != slot_index.end(), "load_param: unknown parameter '%s'", name);
  ParamSlot& sl = slots[it->second];
  LOCO_REQUIRE(sl.numel == numel, "load_param: '%s' expects %lld elements, got %lld", name, sl.numel,
               numel);
  switch (sl.kind) {
    case ParamSlot::RAW:
      LOCO_CHECK_CUDA(cudaMemcpyAsync(w(sl.off_a), src, sizeof(float) * numel,
                                      cudaMemcpyDeviceToDevice, s));
      break;
    case ParamSlot::BIAS_QKV:
      LOCO_CHECK_CUDA(cudaMemcpyAsync(w(sl.off_a) + sl.row_off, src, sizeof(float) * numel,
                                      cudaMemcpyDeviceToDevice, s));
      break;
    case ParamSlot::CONV_GEMM:
    case ParamSlot::CONV_QKV:
      LOCO_TRY(pack_conv_fprop(src, w(sl.off_a) + (size_t)sl.row_off * sl.cin * sl.ksz * sl.ksz,
                               sl.cout, sl.cin, sl.ksz, sl.ksz, s));
      LOCO_TRY(pack_conv_dgrad(src, w(sl.off_b), sl.cout, sl.cin, sl.ksz, sl.ksz, sl.rows_total,
                               sl.row_off, s));
      // fp16 copies (half-sized: the row offset of a fused q|k|v block counts fp16 elements)
      LOCO_TRY(pack_conv_fprop16(src, reinterpret_cast<char*>(w(sl.off_a16)) +
                                          2 * (size_t)sl.row_off * sl.cin * sl.ksz * sl.ksz,
                                 sl.cout, sl.cin, sl.ksz, sl.ksz, s));
      LOCO_TRY(pack_conv_dgrad16(src, w(sl.off_b16), sl.cout, sl.cin, sl.ksz, sl.ksz, sl.rows_total,
                                 sl.row_off, s));
      break;
    case ParamSlot::CONV_THIN: {
      // zero-pad the thin side to kThinPad channels in a scratch copy, then pack like any conv
      const size_t row_real = (size_t)sl.cin * 9, row_pad = (size_t)sl.pad_cin * 9;
      const size_t n = (size_t)sl.pad_cout * row_pad;
      float* tmp = nullptr;
      LOCO_CHECK_CUDA(cudaMalloc(&tmp, sizeof(float) * n));
      int r = 0;
      cudaError_t e = cudaMemsetAsync(tmp, 0, sizeof(float) * n, s);
      if (e == cudaSuccess)
        e = cudaMemcpy2DAsync(tmp, sizeof(float) * row_pad, src, sizeof(float) * row_real, sizeof(float) * row_real,
                              (size_t)sl.cout, cudaMemcpyDeviceToDevice, s);
      if (e == cudaSuccess) {
        r = pack_conv_fprop(tmp, w(sl.off_a), sl.pad_cout, sl.pad_cin, 3, 3, s);
        if (r == 0) r = pack_conv_dgrad(tmp, w(sl.off_b), sl.pad_cout, sl.pad_cin, 3, 3, sl.pad_cout, 0, s);
        if (r == 0) r = pack_conv_fprop16(tmp, w(sl.off_a16), sl.pad_cout, sl.pad_cin, 3, 3, s);
        if (r == 0) r = pack_conv_dgrad16(tmp, w(sl.off_b16), sl.pad_cout, sl.pad_cin, 3, 3, sl.pad_cout, 0, s);
      }
      cudaStreamSynchronize(s);
      cudaFree(tmp);
      LOCO_CHECK_CUDA(e);
      if (r != 0) return r;
      break;
    }
    case ParamSlot::CONV_EDGE_IN:
      LOCO_TRY(pack_conv_edge(src, w(sl.off_a), sl.cout, 1, s));
      break;
    case ParamSlot::CONV_EDGE_OUT:
      LOCO_TRY(pack_conv_edge(src, w(sl.off_a), sl.cin, 0, s));
      break;
  }
  sl.loaded = true;
  return 0;
}

int Model::check_loaded() const {
  for (const auto& s : slots) LOCO_REQUIRE(s.loaded, "parameter '%s' was never loaded", s.name.c_str());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Plan
// ------------------------------------------------------------------------------------------------
typedef std::function<int(cudaStream_t)> Fn;

struct StatTarget {         // where a producer may accumulate GroupNorm statistics of its output
  double* st = nullptr;     // [rows][32][2] doubles of the consumer GroupNorm
  int cg = 0;               // consumer's channels per group
  int choff = 0;            // channel offset of the producer's tensor inside the consumer's input
};
struct TH {                 // a forward tensor and the cotangent buffer that mirrors it
  View v;                   // [NP+NT, H, W, C]
  View g;                   // [NC, H, W, C]
  std::vector<int> ids;     // cotangent region ids (two for a full concat buffer)
  StatTarget st_self;       // GroupNorm over this tensor's own channels
  StatTarget st_cb;         // GroupNorm over the decoder concat buffer this tensor is a slice of
  int partner = -1;         // id of the other slice of that concat buffer
};

struct Plan::Impl {
  // call-time arguments read by the closures (plan-owned staging buffers, so that the launch
  // programs can be captured once into CUDA graphs and replayed for any caller pointers / t)
  const float* cur_x = nullptr;
  float* t_dev = nullptr;
  float* cond_dev = nullptr;      // [4 ch] conditioning embedding added to the timestep embedding (zeros: none)
  // fp16 JVP / VJP programs: power-of-two range scales {s, 1/s} of the tangent rows and of the
  // cotangent rows (applied by the edge convolutions, removed again at the other end), + scratch
  float* tscale_dev = nullptr;
  float* cscale_dev = nullptr;
  unsigned* scale_tmp = nullptr;
  float* cur_eps = nullptr;
  const float* cur_geps = nullptr;
  float* cur_gx = nullptr;
  float *in_buf = nullptr, *out_buf = nullptr, *gin_buf = nullptr, *gout_buf = nullptr;
  cudaGraphExec_t fwd_graph = nullptr, bwd_graph = nullptr;
  long long fwd_graph_launches = 0, bwd_graph_launches = 0;
  // cross-attention layers: projection weights and the K_c | V_c buffer [kCtxPad][2C] of each
  struct CrossLayer { const float* w; const float* b; int C; float* kv; };
  std::vector<CrossLayer> cross_layers;
  int ctx_tokens = 0;             // tokens of the current context (0: none set yet)
  ~Impl() {
    if (fwd_graph) cudaGraphExecDestroy(fwd_graph);
    if (bwd_graph) cudaGraphExecDestroy(bwd_graph);
  }

  std::vector<Fn> fwd;
  std::vector<std::vector<Fn>> bwd_groups;
  std::vector<Fn> bwd;      // flattened in execution order
  std::deque<ConvLaunch> launches;

  // region sizes (floats) from the dry run
  size_t act_floats = 0, fstat_floats = 0, bstat_floats = 0;
  bool sized = false;
  // cotangent writer bookkeeping
  struct Writer { std::vector<int> ids; int group, pos; };
  std::vector<Writer> writers;
  std::vector<int> acc_flags;
  int n_ids = 0;
  double* fstats = nullptr; size_t fstat_bytes = 0;
  double* bstats = nullptr; size_t bstat_bytes = 0;
};

long long Plan::in_elems() const {
  const Arch& A = model->arch;
  return (long long)A.in_ch * A.resolution * A.resolution;
}
long long Plan::out_elems() const {
  const Arch& A = model->arch;
  const long long Ro = A.kind == 2 ? ((long long)A.resolution << (A.n_levels - 1)) : A.resolution;
  return (long long)A.out_ch * Ro * Ro;
}

int Plan::build(float* workspace) {
  if (!impl) impl = std::make_shared<Impl>();
  Impl& I = *impl;
  const bool dry = workspace == nullptr;
  LOCO_REQUIRE(dry || I.sized, "plan: size query must precede binding");
  if (!dry) device = pointer_device(workspace);
  DeviceGuard guard(device);
  if (!dry && model->arena != nullptr) {
    const int wdev = pointer_device(model->arena);
    LOCO_REQUIRE(wdev < 0 || device < 0 || wdev == device,
                 "plan: workspace lives on device %d but the model weights on device %d", device, wdev);
  }
  LOCO_REQUIRE(NP >= 1 && NT >= 0 && NC >= 0, "plan: bad batch configuration");
  LOCO_REQUIRE((NT == 0 && NC == 0) || NP == 1, "plan: tangents/cotangents need exactly one primal row");
  const Model& M = *model;
  const Arch& A = M.arch;
  const int NB = NP + NT;
  const int L = A.n_levels;
  LOCO_REQUIRE(A.ch % 128 == 0, "plan: base channel count %d must be a multiple of 128", A.ch);
  if (!dry) LOCO_REQUIRE(M.arena != nullptr, "plan: model weights not bound");
  if (!dry) { LOCO_TRY(conv_init()); LOCO_TRY(layers_init()); LOCO_TRY(attention_init()); }

  I.fwd.clear(); I.bwd_groups.clear(); I.bwd.clear(); I.launches.clear(); I.writers.clear(); I.cross_layers.clear();
  I.n_ids = 0;
  fwd_flops = vjp_flops = 0;
  base = workspace;

  // ---- bump allocators over three regions: activations | forward stats | backward stats ----
  size_t act_off = 0, fs_off = 0, bs_off = 0;
  // the size query uses a fake non-null base so that "pointer == nullptr" keeps meaning "absent"
  uintptr_t act_base = dry ? (uintptr_t)4096 : (uintptr_t)workspace;
  uintptr_t fs_base = act_base + I.act_floats * 4;
  uintptr_t bs_base = fs_base + I.fstat_floats * 4;
  auto alloc_act = [&](size_t n) -> float* {
    float* p = (float*)(act_base + act_off * 4);
    act_off += (n + 63) & ~(size_t)63;
    return p;
  };
  auto alloc_fstat = [&](int rows) -> double* {
    double* p = (double*)(fs_base + fs_off * 4);
    fs_off += (size_t)rows * 32 * 2 * 2;   // doubles -> float units
    return p;
  };
  auto alloc_bstat = [&](int rows) -> double* {
    double* p = (double*)(bs_base + bs_off * 4);
    bs_off += (size_t)rows * 32 * 2 * 2;
    return p;
  };
  // activations of a plan share one storage type (act16), except the attention core's q|k|v, o and
  // score tensors, which stay fp32 (Tf32 / Tg32): the conv kernels convert at those boundaries
  const int A16 = act16;
  auto act_floats_of = [&](size_t elems, int half) { return half ? (elems + 1) / 2 : elems; };
  auto Tfx = [&](int H, int W, int C, int half) {
    return make_view(alloc_act(act_floats_of((size_t)NB * H * W * C, half)), NB, H, W, C, half);
  };
  auto Tgx = [&](int H, int W, int C, int half) {
    if (NC == 0) return make_view(nullptr, 0, H, W, C, half);
    return make_view(alloc_act(act_floats_of((size_t)NC * H * W * C, half)), NC, H, W, C, half);
  };
  auto Tf = [&](int H, int W, int C) { return Tfx(H, W, C, A16); };
  auto Tg = [&](int H, int W, int C) { return Tgx(H, W, C, A16); };
  auto Tf32 = [&](int H, int W, int C) { return Tfx(H, W, C, 0); };
  auto Tg32 = [&](int H, int W, int C) { return Tgx(H, W, C, 0); };
  // GroupNorm statistics are fused into the producing conv epilogue in forward-only programs
  // (tangent rows need sum(x0*dx), which a tile of one batch row cannot form)
  const bool fuse_stats = (NT == 0);
  std::vector<char> fused_self, fused_cb;
  auto new_id = [&]() { fused_self.push_back(0); fused_cb.push_back(0); return I.n_ids++; };
  auto new_tensor = [&](int H, int W, int C) {
    TH t; t.v = Tf(H, W, C); t.g = Tg(H, W, C); t.ids = {new_id()};
    t.st_self.st = alloc_fstat(NB); t.st_self.cg = C / 32; t.st_self.choff = 0;
    return t;
  };
  auto row0 = [](const View& v) { return slice_n(v, 0, 1); };

  // ---- cotangent writer flags (resolved by the dry run) ----
  int writer_counter = 0;
  int cur_group = -1;
  auto begin_group = [&]() { I.bwd_groups.emplace_back(); cur_group = (int)I.bwd_groups.size() - 1; };
  auto push_b = [&](Fn f) { I.bwd_groups[cur_group].push_back(std::move(f)); };
  auto writer_flag = [&](const std::vector<int>& ids) -> int {
    const int idx = writer_counter++;
    if (dry) {
      Impl::Writer w; w.ids = ids; w.group = cur_group; w.pos = (int)I.bwd_groups[cur_group].size();
      I.writers.push_back(w);
      return 0;
    }
    return I.acc_flags[idx];
  };

  // ---- split-K scratch shared by every conv launch of this plan (same stream => serialised) ----
  const long long splitk_floats = 4LL << 20;          // 16 MB of partial tiles
  const int splitk_tiles = 4096;
  float* splitk_partial = alloc_act((size_t)splitk_floats);
  int* splitk_counters = reinterpret_cast<int*>(alloc_act(2 * splitk_tiles));   // arrive + done
  if (!dry) {
    const cudaError_t ce = cudaMemset(splitk_counters, 0, sizeof(int) * 2 * splitk_tiles);
    LOCO_REQUIRE(ce == cudaSuccess, "plan: cudaMemset(counters) failed: %s", cudaGetErrorString(ce));
  }
  int err = 0;
  auto mk_conv = [&](ConvProblem prob, bool is_bwd) -> ConvLaunch* {
    I.launches.emplace_back();
    ConvLaunch* L_ = &I.launches.back();
    const double fl = 2.0 * prob.out.N * prob.out.H * prob.out.W * (double)prob.Ngemm *
                      ((double)prob.Kc * (prob.kind == CONV_1x1 ? 1 : 9) / (prob.kind == CONV_3x3_S2_DGRAD ? 4 : 1) +
                       prob.Kc2);
    (is_bwd ? vjp_flops : fwd_flops) += fl;
    prob.splitk_partial = splitk_partial; prob.splitk_partial_floats = splitk_floats;
    prob.splitk_counters = splitk_counters; prob.splitk_max_tiles = splitk_tiles;
    if (!dry) {
      const int r = conv_prepare(prob, L_);
      if (r != 0 && err == 0) err = r;
    }
    return L_;
  };
  // in2 / c2: optional fused 1x1 shortcut (out = conv(in) + conv1x1_{c2}(in2)), halo variant only
  auto conv_fwd = [&](int kind, View in, View out, const ConvRef& c, const float* bias2,
                      const View* addend, const TH* out_th = nullptr, StatTarget extra = StatTarget(),
                      const View* in2 = nullptr, const ConvRef* c2 = nullptr) {
    ConvProblem p;
    p.kind = kind; p.in = in; p.out = out; p.wpack = dry ? nullptr : M.w(in.half ? c.wf16 : c.wf);
    p.Kc = c.cin; p.Ngemm = c.cout;
    if (in2 != nullptr) { p.in2 = in2; p.wpack2 = dry ? nullptr : M.w(in.half ? c2->wf16 : c2->wf); p.Kc2 = c2->cin; }
    p.bias = dry ? nullptr : M.w(c.bias); p.bias2 = bias2; p.bias_rows = NP;
    p.addend = addend; p.accumulate = 0; p.round_out = 1;
    if (fuse_stats) {
      int nt = 0;
      auto add_target = [&](const StatTarget& t) {
        p.st_ptr[nt] = t.st; p.st_cg[nt] = t.cg; p.st_choff[nt] = t.choff; ++nt;
      };
      if (extra.st) add_target(extra);
      if (out_th) {
        const int id = out_th->ids[0];
        if (out_th->st_self.st) { add_target(out_th->st_self); fused_self[id] = 1; }
        // a concat buffer's statistics are fused only if both of its slices are produced by convs
        if (out_th->st_cb.st && (out_th->partner < 0 || fused_cb[out_th->partner])) {
          add_target(out_th->st_cb); fused_cb[id] = 1;
        }
      }
    }
    ConvLaunch* l = mk_conv(p, false);
    I.fwd.push_back([l](cudaStream_t s) { return conv_run(*l, s); });
  };
  // data-gradient convolution: gy [NC,..,cout] -> gx [NC,..,cin]
  auto conv_bwd = [&](int fwd_kind, View gy, View gx, const ConvRef& c, int accumulate) {
    ConvProblem p;
    p.kind = fwd_kind == CONV_3x3 ? CONV_3x3_DGRAD
                                  : (fwd_kind == CONV_3x3_S2 ? CONV_3x3_S2_DGRAD : CONV_1x1);
    p.in = gy; p.out = gx; p.wpack = dry ? nullptr : M.w(gy.half ? c.wd16 : c.wd);
    p.Kc = c.cout; p.Ngemm = c.cin;
    p.accumulate = accumulate; p.round_out = 1;
    ConvLaunch* l = mk_conv(p, true);
    push_b([l](cudaStream_t s) { return conv_run(*l, s); });
  };
  const float eps = A.gn_eps;
  // `aff` (optional): effective affine [gamma' | beta'] of a scale-shift norm site (kind 1)
  auto gn_fwd = [&](View x, const NormRef& n, int silu, int round_out, View y, double* st = nullptr,
                    bool fused = false, const float* aff = nullptr) -> double* {
    if (!st) st = alloc_fstat(NB);
    const float* ga = aff ? aff : (dry ? nullptr : M.w(n.gamma));
    const float* be = aff ? aff + n.C : (dry ? nullptr : M.w(n.beta));
    const int np = NP;
    // small sites whose statistics no conv epilogue produced: statistics + apply in one launch
    const bool small = !fused && gn_small_eligible(x) && y.sH == (long long)y.W * y.sW;
    I.fwd.push_back([=](cudaStream_t s) {
      if (small) return gn_small_fwd(x, np, st, ga, be, eps, silu, round_out, y, s);
      if (!fused) LOCO_TRY(gn_stats_fwd(x, np, st, s));   // else: accumulated by the producer's epilogue
      return gn_apply_fwd(x, np, st, ga, be, eps, silu, round_out, y, s);
    });
    return st;
  };
  auto th_fused = [&](const TH& t) -> bool {
    if (!fuse_stats) return false;
    if (t.ids.size() == 2) return fused_cb[t.ids[0]] && fused_cb[t.ids[1]];
    return fused_self[t.ids[0]] != 0;
  };
  auto gn_bwd = [&](View xp, const double* pstats, View gy, const NormRef& n, int silu,
                    const View* addend, int accumulate, int round_out, View gx,
                    const float* aff = nullptr) {
    double* st = alloc_bstat(NC);
    const float* ga = aff ? aff : (dry ? nullptr : M.w(n.gamma));
    const float* be = aff ? aff + n.C : (dry ? nullptr : M.w(n.beta));
    const bool has_add = addend != nullptr;
    const View add = has_add ? *addend : View();
    const bool small = gn_small_eligible(xp) && gn_small_eligible(gy) && gx.sH == (long long)gx.W * gx.sW &&
                       (!has_add || add.sH == (long long)add.W * add.sW);
    push_b([=](cudaStream_t s) {
      if (small) return gn_small_vjp(xp, pstats, gy, ga, be, eps, silu, has_add ? &add : nullptr, accumulate, round_out, gx, s);
      LOCO_TRY(gn_stats_vjp(xp, pstats, gy, ga, be, eps, silu, st, s));
      return gn_apply_vjp(xp, pstats, gy, st, ga, be, eps, silu, has_add ? &add : nullptr, accumulate,
                          round_out, gx, s);
    });
  };

  // ---- staging buffers (NCHW images at the ABI) ----
  {
    const size_t img_in = (size_t)in_elems(), img_out = (size_t)out_elems();
    I.in_buf = alloc_act((size_t)NB * img_in);
    I.out_buf = alloc_act((size_t)NB * img_out);
    I.gin_buf = NC ? alloc_act((size_t)NC * img_out) : nullptr;
    I.gout_buf = NC ? alloc_act((size_t)NC * img_in) : nullptr;
    I.t_dev = alloc_act(64);
    I.cond_dev = alloc_act(4 * (size_t)A.ch);
    I.tscale_dev = alloc_act(64);
    I.cscale_dev = I.tscale_dev + 8;
    I.scale_tmp = reinterpret_cast<unsigned*>(I.tscale_dev + 16);
    if (!dry) {
      const cudaError_t ce = cudaMemset(I.cond_dev, 0, sizeof(float) * 4 * (size_t)A.ch);
      LOCO_REQUIRE(ce == cudaSuccess, "plan: cudaMemset(condition) failed: %s", cudaGetErrorString(ce));
    }
    if (I.fwd_graph) { cudaGraphExecDestroy(I.fwd_graph); I.fwd_graph = nullptr; }
    if (I.bwd_graph) { cudaGraphExecDestroy(I.bwd_graph); I.bwd_graph = nullptr; }
  }
  // ---- timestep embedding ----
  const int temb_ch = 4 * A.ch;
  float* temb_scratch = alloc_act(2 * temb_ch);
  float* tproj = alloc_act(M.tproj_rows);
  // kind 1: every ResBlock's second GroupNorm takes its affine from (scale | shift) = tproj slice;
  // the effective [gamma' | beta'] vectors live in `aff_buf`, filled by one launch per forward
  std::vector<AffineSite> aff_sites;
  float* aff_buf = alloc_act(A.kind == 1 ? (size_t)M.tproj_rows : 0);
  AffineSite* aff_dev = reinterpret_cast<AffineSite*>(alloc_act(A.kind == 1 ? 4096 : 0));
  const int aff_cap = (int)(4096 * sizeof(float) / sizeof(AffineSite));
  size_t aff_off = 0;
  auto affine_site = [&](const ResRef& R) -> const float* {
    if (!R.scale_shift) return nullptr;
    AffineSite st;
    st.gamma_off = (long long)R.n2.gamma; st.beta_off = (long long)R.n2.beta;
    st.tproj_off = R.temb_off; st.C = R.cout; st.out_off = (long long)aff_off;
    aff_sites.push_back(st);
    const float* p = aff_buf + aff_off;
    aff_off += 2 * (size_t)R.cout;
    return p;
  };
  const size_t temb_op_index = I.fwd.size();
  I.fwd.push_back([](cudaStream_t) { return 0; });   // replaced below once the sites are known

  // ---- blocks ----
  auto resblock = [&](const ResRef& R, const TH& x, const TH& out) {
    const int H = x.v.H, W = x.v.W;
    const int Ho = R.resample == 1 ? H / 2 : (R.resample == 2 ? 2 * H : H);
    const int Wo = R.resample == 1 ? W / 2 : (R.resample == 2 ? 2 * W : W);
    const float* aff = affine_site(R);
    View a1 = Tf(H, W, R.cin);
    const int x_fused = th_fused(x);
    double* st1 = gn_fwd(x.v, R.n1, 1, 1, a1, x.st_self.st, x_fused);
    // Forward-only programs: a skip tensor no conv epilogue produced (the 3 -> C conv_in output) just had its own
    // statistics computed; when the decoder concat buffer it also belongs to has twice its channels per group, that
    // buffer's statistics over this slice are pair sums of them -- no second pass over the tensor, and the decoder
    // slice's producer may then fuse its half (fused_cb) so the concat GroupNorm needs no statistics pass at all.
    if (fuse_stats && x_fused == 0 && x.ids.size() == 1 && x.partner < 0 && x.st_cb.st != nullptr && x.st_self.st != nullptr &&
        !fused_cb[x.ids[0]] && x.st_cb.cg == 2 * x.st_self.cg && x.st_cb.choff % x.st_cb.cg == 0 &&
        x.st_cb.choff / x.st_cb.cg + 16 <= 32) {
      const double* src = x.st_self.st;
      double* dst = x.st_cb.st;
      const int g0 = x.st_cb.choff / x.st_cb.cg, rows = NB;
      I.fwd.push_back([=](cudaStream_t s) { return gn_stats_fold_pairs(src, dst, rows, g0, s); });
      fused_cb[x.ids[0]] = 1;
    }
    // guided-diffusion up/down ResBlock: both branches are resampled after GN+SiLU
    // (unet.py:238-244): avg-pool 2x2 (Downsample, use_conv = False) or nearest x2 (Upsample)
    View a1r = a1, xr = x.v;
    if (R.resample != 0) {
      a1r = Tf(Ho, Wo, R.cin);
      xr = Tf(Ho, Wo, R.cin);
      const View xv = x.v;
      if (R.resample == 1)
        I.fwd.push_back([=](cudaStream_t s) {
          LOCO_TRY(sumpool2x(a1, a1r, 0.25f, 0, 1, s));
          return sumpool2x(xv, xr, 0.25f, 0, 0, s);
        });
      else
        I.fwd.push_back([=](cudaStream_t s) {
          LOCO_TRY(upsample2x(a1, a1r, 1.f, 0, 0, s));
          return upsample2x(xv, xr, 1.f, 0, 0, s);
        });
    }
    View h1 = Tf(Ho, Wo, R.cout);
    StatTarget t2; t2.st = alloc_fstat(NB); t2.cg = R.cout / 32; t2.choff = 0;
    conv_fwd(CONV_3x3, a1r, h1, R.c1, (R.scale_shift || R.temb_off < 0) ? nullptr : tproj + R.temb_off, nullptr, nullptr, t2);
    View a2 = Tf(Ho, Wo, R.cout);
    double* st2 = gn_fwd(h1, R.n2, 1, 1, a2, t2.st, fuse_stats, aff);
    View sc;
    if (R.has_nin && conv_halo_eligible(CONV_3x3, NB, Ho, Wo, R.cout)) {
      // the 1x1 shortcut rides the K loop of conv2 (its bias goes in the second bias slot)
      conv_fwd(CONV_3x3, a2, out.v, R.c2, dry ? nullptr : M.w(R.nin.bias), nullptr, &out, StatTarget(),
               &xr, &R.nin);
    } else if (R.has_nin) {
      sc = Tf(Ho, Wo, R.cout);
      conv_fwd(CONV_1x1, xr, sc, R.nin, nullptr, nullptr);
      conv_fwd(CONV_3x3, a2, out.v, R.c2, nullptr, &sc, &out);
    } else {
      conv_fwd(CONV_3x3, a2, out.v, R.c2, nullptr, &xr, &out);
    }
    if (NC > 0) {
      begin_group();
      View ga2 = Tg(Ho, Wo, R.cout);
      conv_bwd(CONV_3x3, out.g, ga2, R.c2, 0);
      View gh1 = Tg(Ho, Wo, R.cout);
      gn_bwd(row0(h1), st2, ga2, R.n2, 1, nullptr, 0, 1, gh1, aff);
      View ga1r = Tg(Ho, Wo, R.cin);
      conv_bwd(CONV_3x3, gh1, ga1r, R.c1, 0);
      View ga1 = ga1r;
      if (R.resample != 0) {
        ga1 = Tg(H, W, R.cin);
        if (R.resample == 1) push_b([=](cudaStream_t s) { return upsample2x(ga1r, ga1, 0.25f, 0, 0, s); });
        else push_b([=](cudaStream_t s) { return sumpool2x(ga1r, ga1, 1.f, 0, 0, s); });
      }
      // the 1x1 shortcut's data gradient goes FIRST (a plain store), the GroupNorm rule accumulates onto it: the
      // read-modify-write of the (up to 384-channel) input gradient then happens in the HBM-rate GroupNorm kernel
      // instead of the K = cout conv epilogue (one epilogue warpgroup per CTA: 369 us = 2.3 TB/s for the 838 MB of the
      // 256^2 decoder sites at rank 10, profiles/r2_launches_step_k10.csv)
      if (R.has_nin) {
        if (R.resample != 0 && err == 0) {   // (a void lambda: report through `err`)
          set_error("plan: resampling ResBlock with a 1x1 shortcut is not supported");
          err = 3;
        }
        const int f2 = writer_flag(x.ids);
        conv_bwd(CONV_1x1, out.g, x.g, R.nin, f2);
      }
      const int f1 = writer_flag(x.ids);
      const bool fuse_identity = !R.has_nin && R.resample == 0;
      gn_bwd(row0(x.v), st1, ga1, R.n1, 1, fuse_identity ? &out.g : nullptr, f1, 1, x.g);
      if (R.has_nin) {
      } else if (R.resample != 0) {
        // identity shortcut through the resampling: x.g += resample^T(out.g)
        const int f2 = writer_flag(x.ids);
        const View og = out.g, xg = x.g;
        if (R.resample == 1) push_b([=](cudaStream_t s) { return upsample2x(og, xg, 0.25f, f2, 1, s); });
        else push_b([=](cudaStream_t s) { return sumpool2x(og, xg, 1.f, f2, 1, s); });
      }
    }
  };
  auto self_attn = [&](const AttnRef& R, const TH& x, const TH& out) {
    const int H = x.v.H, W = x.v.W, C = R.C, T = H * W;
    View hn = Tf(H, W, C);
    double* st = gn_fwd(x.v, R.n, 0, 1, hn, x.st_self.st, th_fused(x));
    View qkv = Tf32(H, W, 3 * C);
    conv_fwd(CONV_1x1, hn, qkv, R.qkv, nullptr, nullptr);
    const int hc = A.kind == 1 ? A.head_ch : 0;
    const int heads = hc > 0 ? C / hc : 1;
    float* S = alloc_act((size_t)NB * heads * T * T);
    View o = Tf32(H, W, C);
    const int np = NP;
    I.fwd.push_back([=](cudaStream_t s) { return attention_forward(qkv, np, hc, S, o, s); });
    conv_fwd(CONV_1x1, o, out.v, R.proj, nullptr, &x.v, &out);
    if (NC > 0) {
      begin_group();
      View go = Tg32(H, W, C);
      conv_bwd(CONV_1x1, out.g, go, R.proj, 0);
      View gqkv = Tg32(H, W, 3 * C);
      float* gP = alloc_act((size_t)NC * heads * T * T);
      View qkv0 = row0(qkv);
      push_b([=](cudaStream_t s) { return attention_vjp(go, qkv0, hc, S, gP, gqkv, s); });
      View ghn = Tg(H, W, C);
      conv_bwd(CONV_1x1, gqkv, ghn, R.qkv, 0);
      const int f = writer_flag(x.ids);
      gn_bwd(row0(x.v), st, ghn, R.n, 0, &out.g, f, 1, x.g);
    }
  };
  // out = x + proj2(softmax(q2(GN(x)) K_c^T / sqrt(D)) V_c): cross-attention to the prompt embedding.
  // K_c | V_c do not depend on x, so the tangent / cotangent rules only involve the query side.
  auto cross_attn = [&](const AttnRef& R, const TH& x, const TH& out) {
    const int H = x.v.H, W = x.v.W, C = R.C, T = H * W;
    const int heads = A.ctx_heads;
    if (err == 0 && !attention_cross_eligible(kCtxPad, C, heads)) {
      set_error("plan: cross-attention with %d channels in %d heads is not supported", C, heads);
      err = 3;
    }
    float* kvb = alloc_act((size_t)kCtxPad * 2 * C);
    if (!dry) {
      Impl::CrossLayer cl; cl.w = M.w(R.kv_w); cl.b = M.w(R.kv_b); cl.C = C; cl.kv = kvb;
      I.cross_layers.push_back(cl);
    }
    const View kv = make_view(kvb, 1, 1, kCtxPad, 2 * C);
    View hn = Tf(H, W, C);
    double* st = gn_fwd(x.v, R.n2, 0, 1, hn, x.st_self.st, th_fused(x));
    View q = Tf32(H, W, C);
    conv_fwd(CONV_1x1, hn, q, R.q2, nullptr, nullptr);
    float* S = alloc_act((size_t)NB * heads * T * kCtxPad);
    View o = Tf32(H, W, C);
    const int np = NP;
    Impl* Ip = impl.get();
    I.fwd.push_back([=](cudaStream_t s) { return attention_cross_forward_tc(q, kv, np, heads, Ip->ctx_tokens, S, o, s); });
    conv_fwd(CONV_1x1, o, out.v, R.proj2, nullptr, &x.v, &out);
    if (NC > 0) {
      begin_group();
      View go = Tg32(H, W, C);
      conv_bwd(CONV_1x1, out.g, go, R.proj2, 0);
      View gq = Tg32(H, W, C);
      push_b([=](cudaStream_t s) { return attention_cross_vjp_tc(go, kv, heads, Ip->ctx_tokens, S, gq, s); });
      View ghn = Tg(H, W, C);
      conv_bwd(CONV_1x1, gq, ghn, R.q2, 0);
      const int f = writer_flag(x.ids);
      gn_bwd(row0(x.v), st, ghn, R.n2, 0, &out.g, f, 1, x.g);
    }
  };
  auto attnblock = [&](const AttnRef& R, const TH& x, const TH& out) {
    if (!R.cross) { self_attn(R, x, out); return; }
    TH mid = new_tensor(x.v.H, x.v.W, R.C);
    self_attn(R, x, mid);
    cross_attn(R, mid, out);
  };

  // ---- network input / output ends ----
  // 3-channel images: CUDA-core / mma.sync edge convolutions; thin (4-channel latent) ends: zero-padded
  // 64-channel tensors through the tcgen05 conv kernels (layers.cuh: thin_pad / thin_extract)
  const bool sc16 = A16 != 0;     // range scaling of the tangent / cotangent rows (fp16 storage only)
  auto emit_input = [&](TH& h0) {
    Impl* Ip = impl.get();
    const int np = NP;
    if (M.thin) {
      const int cin = A.in_ch, R0 = A.resolution;
      View zin = Tf(R0, R0, kThinPad);
      const float* mix = (M.has_pq && !dry) ? M.w(M.pq_mix) : nullptr;
      I.fwd.push_back([=](cudaStream_t s) {
        return thin_pad(Ip->cur_x, cin, zin, mix, np, sc16 ? Ip->tscale_dev : nullptr, np, 1, s);
      });
      conv_fwd(CONV_3x3, zin, h0.v, M.thin_in, nullptr, nullptr, &h0);
      if (NC > 0) {
        begin_group();
        View gzin = Tg(R0, R0, kThinPad);
        conv_bwd(CONV_3x3, h0.g, gzin, M.thin_in, 0);
        push_b([=](cudaStream_t s) { return thin_extract(gzin, cin, Ip->cur_gx, mix, sc16 ? Ip->cscale_dev : nullptr, 0, s); });
      }
      return;
    }
    const float* wi = dry ? nullptr : M.w(M.conv_in_w);
    const float* bi = dry ? nullptr : M.w(M.conv_in_b);
    const View hv = h0.v, hg = h0.g;
    I.fwd.push_back([=](cudaStream_t s) {
      return edge_conv_expand(Ip->cur_x, wi, bi, np, hv, 0, 1, s, sc16 ? Ip->tscale_dev : nullptr, np);
    });
    if (NC > 0) {
      begin_group();
      push_b([=](cudaStream_t s) {
        return edge_conv_reduce(hg, wi, nullptr, 0, Ip->cur_gx, 1, s, sc16 ? Ip->cscale_dev : nullptr, 0);
      });
    }
  };
  TH hfin;
  if (A.kind == 2) {
    // ---- VAE decoder (build_decoder): conv_in, mid, up levels L-1 .. 0, head ----
    int res = A.resolution;
    TH cur = new_tensor(res, res, M.mid1.cin);
    emit_input(cur);
    {
      TH m1 = new_tensor(res, res, M.mid1.cout);
      resblock(M.mid1, cur, m1);
      TH m2 = new_tensor(res, res, M.mid1.cout);
      attnblock(M.mid_attn, m1, m2);
      TH m3 = new_tensor(res, res, M.mid2.cout);
      resblock(M.mid2, m2, m3);
      cur = m3;
    }
    for (int l = L - 1; l >= 0; --l) {
      for (int b = 0; b < A.num_res_blocks + 1; ++b) {
        const ResRef& R = M.up_res[l][b];
        TH nxt = new_tensor(res, res, R.cout);
        resblock(R, cur, nxt);
        cur = nxt;
      }
      if (l != 0) {
        const ConvRef& c = M.up_sample[l];
        const int C = cur.v.C;
        View hu = Tf(2 * res, 2 * res, C);
        const View tv = cur.v;
        I.fwd.push_back([=](cudaStream_t s) { return upsample2x(tv, hu, 1.f, 0, 0, s); });
        TH nxt = new_tensor(2 * res, 2 * res, C);
        conv_fwd(CONV_3x3, hu, nxt.v, c, nullptr, nullptr, &nxt);
        if (NC > 0) {
          begin_group();
          View ghu = Tg(2 * res, 2 * res, C);
          conv_bwd(CONV_3x3, nxt.g, ghu, c, 0);
          const int f = writer_flag(cur.ids);
          const View tg = cur.g;
          push_b([=](cudaStream_t s) { return sumpool2x(ghu, tg, 1.f, f, 0, s); });
        }
        cur = nxt;
        res *= 2;
      }
    }
    hfin = cur;
  } else {
  // ---- topology (reference: PullBackDDPM.forward, ddpm/diffusion.py:145-200) ----
  // simulate the encoder to learn the skip stack, then size the decoder concat buffers
  struct HsInfo { int C, res; };
  std::vector<HsInfo> hs_info;
  {
    int res = A.resolution;
    hs_info.push_back({A.ch, res});
    for (int l = 0; l < L; ++l) {
      const int bo = A.ch * A.ch_mult[l];
      for (int b = 0; b < A.num_res_blocks; ++b) hs_info.push_back({bo, res});
      if (l != L - 1) { res /= 2; hs_info.push_back({bo, res}); }
    }
  }
  const int n_hs = (int)hs_info.size();
  LOCO_REQUIRE(n_hs == L * (A.num_res_blocks + 1), "plan: skip stack size mismatch");
  struct CB { TH full, dec, skip; };
  std::vector<CB> cbs(n_hs);
  {
    int block_in = A.ch * A.ch_mult[L - 1];
    int u = 0;
    for (int l = L - 1; l >= 0; --l) {
      const int bo = A.ch * A.ch_mult[l];
      for (int b = 0; b < A.num_res_blocks + 1; ++b, ++u) {
        const HsInfo& h = hs_info[n_hs - 1 - u];
        const int C0 = block_in, C1 = h.C;
        CB& cb = cbs[u];
        cb.full.v = Tf(h.res, h.res, C0 + C1);
        cb.full.g = Tg(h.res, h.res, C0 + C1);
        const int id_d = new_id(), id_s = new_id();
        cb.full.ids = {id_d, id_s};
        cb.full.st_self.st = alloc_fstat(NB); cb.full.st_self.cg = (C0 + C1) / 32; cb.full.st_self.choff = 0;
        cb.dec.v = slice_c(cb.full.v, 0, C0); cb.dec.g = slice_c(cb.full.g, 0, C0); cb.dec.ids = {id_d};
        cb.skip.v = slice_c(cb.full.v, C0, C1); cb.skip.g = slice_c(cb.full.g, C0, C1); cb.skip.ids = {id_s};
        // the skip slice feeds the next encoder GroupNorm on its own and the decoder GroupNorm as a
        // part of the concat buffer; the decoder slice only the latter
        cb.skip.st_self.st = alloc_fstat(NB); cb.skip.st_self.cg = C1 / 32; cb.skip.st_self.choff = 0;
        cb.skip.st_cb = cb.full.st_self; cb.skip.st_cb.choff = C0;
        cb.dec.st_cb = cb.full.st_self; cb.dec.st_cb.choff = 0; cb.dec.partner = id_s;
        block_in = bo;
      }
    }
  }
  auto hs_slot = [&](int i) -> TH& { return cbs[n_hs - 1 - i].skip; };
  // conv_in
  emit_input(hs_slot(0));
  // encoder
  int hs_top = 0;   // index of the last pushed skip tensor
  {
    int res = A.resolution;
    for (int l = 0; l < L; ++l) {
      for (int b = 0; b < A.num_res_blocks; ++b) {
        const ResRef& R = M.down_res[l][b];
        TH& x = hs_slot(hs_top);
        TH& out = hs_slot(hs_top + 1);
        if (!M.down_attn[l].empty()) {
          TH tmp = new_tensor(res, res, R.cout);
          resblock(R, x, tmp);
          attnblock(M.down_attn[l][b], tmp, out);
        } else {
          resblock(R, x, out);
        }
        ++hs_top;
      }
      if (l != L - 1) {
        TH& x = hs_slot(hs_top);
        TH& out = hs_slot(hs_top + 1);
        if (A.kind == 1) {
          resblock(M.down_rb[l], x, out);
        } else {
          const ConvRef& c = M.down_sample[l];
          conv_fwd(CONV_3x3_S2, x.v, out.v, c, nullptr, nullptr, &out);
          if (NC > 0) {
            begin_group();
            const int f = writer_flag(x.ids);
            conv_bwd(CONV_3x3_S2, out.g, x.g, c, f);
          }
        }
        ++hs_top;
        res /= 2;
      }
    }
  }
  // middle
  {
    TH& x = hs_slot(hs_top);
    const int res = x.v.H;
    TH m1 = new_tensor(res, res, M.mid1.cout);
    resblock(M.mid1, x, m1);
    TH m2 = new_tensor(res, res, M.mid1.cout);
    attnblock(M.mid_attn, m1, m2);
    resblock(M.mid2, m2, cbs[0].dec);
  }
  // decoder
  {
    int u = 0;
    for (int l = L - 1; l >= 0; --l) {
      for (int b = 0; b < A.num_res_blocks + 1; ++b, ++u) {
        const ResRef& R = M.up_res[l][b];
        const int res = cbs[u].full.v.H;
        const bool has_attn = !M.up_attn[l].empty();
        const bool last_in_level = b == A.num_res_blocks;
        const bool needs_up = last_in_level && l != 0;
        const bool very_last = last_in_level && l == 0;
        TH target;
        if (needs_up || very_last) target = new_tensor(res, res, R.cout);
        else target = cbs[u + 1].dec;
        if (has_attn) {
          TH tmp = new_tensor(res, res, R.cout);
          resblock(R, cbs[u].full, tmp);
          attnblock(M.up_attn[l][b], tmp, target);
        } else {
          resblock(R, cbs[u].full, target);
        }
        if (needs_up && A.kind == 1) {
          resblock(M.up_rb[l], target, cbs[u + 1].dec);
        } else if (needs_up) {
          const ConvRef& c = M.up_sample[l];
          View hu = Tf(2 * res, 2 * res, R.cout);
          const View tv = target.v;
          I.fwd.push_back([=](cudaStream_t s) { return upsample2x(tv, hu, 1.f, 0, 0, s); });
          TH& nxt = cbs[u + 1].dec;
          conv_fwd(CONV_3x3, hu, nxt.v, c, nullptr, nullptr, &nxt);
          if (NC > 0) {
            begin_group();
            View ghu = Tg(2 * res, 2 * res, R.cout);
            conv_bwd(CONV_3x3, nxt.g, ghu, c, 0);
            const int f = writer_flag(target.ids);
            const View tg = target.g;
            push_b([=](cudaStream_t s) { return sumpool2x(ghu, tg, 1.f, f, 0, s); });
          }
        }
        if (very_last) hfin = target;
      }
    }
  }
  }   // U-Net topology
  // head
  if (M.thin) {
    const int res = hfin.v.H, C = hfin.v.C, cout = A.out_ch;
    View a = Tf(res, res, C);
    double* st = gn_fwd(hfin.v, M.norm_out, 1, 1, a, hfin.st_self.st, th_fused(hfin));
    Impl* Ip = impl.get();
    const int np = NP;
    View o64 = Tf(res, res, kThinPad);
    conv_fwd(CONV_3x3, a, o64, M.thin_out, nullptr, nullptr);
    I.fwd.push_back([=](cudaStream_t s) { return thin_extract(o64, cout, Ip->cur_eps, nullptr, sc16 ? Ip->tscale_dev : nullptr, np, s); });
    if (NC > 0) {
      begin_group();
      View g64 = Tg(res, res, kThinPad);
      push_b([=](cudaStream_t s) { return thin_pad(Ip->cur_geps, cout, g64, nullptr, 0, sc16 ? Ip->cscale_dev : nullptr, 0, 1, s); });
      View ga = Tg(res, res, C);
      conv_bwd(CONV_3x3, g64, ga, M.thin_out, 0);
      const int f = writer_flag(hfin.ids);
      gn_bwd(row0(hfin.v), st, ga, M.norm_out, 1, nullptr, f, 1, hfin.g);
    }
  } else {
    const int res = hfin.v.H, C = hfin.v.C;
    View a = Tf(res, res, C);
    double* st = gn_fwd(hfin.v, M.norm_out, 1, 0, a, hfin.st_self.st, th_fused(hfin));
    Impl* Ip = impl.get();
    const float* wo = dry ? nullptr : M.w(M.conv_out_w);
    const float* bo = dry ? nullptr : M.w(M.conv_out_b);
    const int np = NP;
    I.fwd.push_back([=](cudaStream_t s) {
      return edge_conv_reduce(a, wo, bo, np, Ip->cur_eps, 0, s, sc16 ? Ip->tscale_dev : nullptr, np);
    });
    if (NC > 0) {
      begin_group();
      View ga = Tg(res, res, C);
      push_b([=](cudaStream_t s) {
        return edge_conv_expand(Ip->cur_geps, wo, nullptr, 0, ga, 1, 0, s, sc16 ? Ip->cscale_dev : nullptr, 0);
      });
      const int f = writer_flag(hfin.ids);
      gn_bwd(row0(hfin.v), st, ga, M.norm_out, 1, nullptr, f, 1, hfin.g);
    }
  }
  if (err != 0) return err;

  // ---- timestep embedding op (first op of the forward program) ----
  LOCO_REQUIRE((int)aff_sites.size() <= aff_cap, "plan: %d scale-shift sites exceed the table", (int)aff_sites.size());
  if (!dry) {
    const Model* Mp = model;
    Impl* Ip = impl.get();
    const int n_aff = (int)aff_sites.size();
    if (n_aff > 0) {
      const cudaError_t ce = cudaMemcpy(aff_dev, aff_sites.data(), sizeof(AffineSite) * n_aff,
                                        cudaMemcpyHostToDevice);
      LOCO_REQUIRE(ce == cudaSuccess, "plan: cudaMemcpy(affine sites) failed: %s", cudaGetErrorString(ce));
    }
    if (A.kind != 2) I.fwd[temb_op_index] = [=](cudaStream_t s) {
      LOCO_TRY(temb_forward(Ip->t_dev, Mp->arch.ch, Mp->w(Mp->temb_w0), Mp->w(Mp->temb_b0),
                            Mp->w(Mp->temb_w1), Mp->w(Mp->temb_b1), temb_scratch, Mp->arch.kind, Ip->cond_dev, s));
      LOCO_TRY(temb_project(temb_scratch + temb_ch, temb_ch, Mp->w(Mp->tproj_w), Mp->w(Mp->tproj_b),
                            Mp->tproj_rows, tproj, s));
      return scale_shift_affine(aff_dev, n_aff, Mp->arena, tproj, aff_buf, s);
    };
  }

  if (dry) {
    I.act_floats = act_off; I.fstat_floats = fs_off; I.bstat_floats = bs_off;
    workspace_floats = act_off + fs_off + bs_off + 64;
    // resolve accumulate flags in backward execution order (groups reversed, ops in order)
    std::vector<int> order(I.writers.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = (int)i;
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
      const auto& wa = I.writers[a]; const auto& wb = I.writers[b];
      if (wa.group != wb.group) return wa.group > wb.group;
      return wa.pos < wb.pos;
    });
    std::vector<char> init(I.n_ids, 0);
    I.acc_flags.assign(I.writers.size(), 0);
    for (int wi : order) {
      const auto& w = I.writers[wi];
      int n_init = 0;
      for (int id : w.ids) n_init += init[id];
      LOCO_REQUIRE(n_init == 0 || n_init == (int)w.ids.size(),
                   "plan: inconsistent cotangent initialisation (writer %d)", wi);
      I.acc_flags[wi] = n_init ? 1 : 0;
      for (int id : w.ids) init[id] = 1;
    }
    I.sized = true;
  } else {
    LOCO_REQUIRE(act_off == I.act_floats && fs_off == I.fstat_floats && bs_off == I.bstat_floats,
                 "plan: non-deterministic layout between size query and binding");
    I.fstats = (double*)fs_base; I.fstat_bytes = fs_off * 4;
    I.bstats = (double*)bs_base; I.bstat_bytes = bs_off * 4;
    for (int g = (int)I.bwd_groups.size() - 1; g >= 0; --g)
      for (auto& f : I.bwd_groups[g]) I.bwd.push_back(f);
  }
  fwd_launches = (int)I.fwd.size();
  vjp_launches = (int)I.bwd.size();
  return 0;
}

static bool graphs_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("LOCO_NO_GRAPH");
    v = (e && e[0] == '1') ? 0 : 1;
  }
  return v == 1 && !profile_is_on();
}

// Run `ops` on `s`, through a CUDA graph captured on first use (the program is static: same
// kernels, same buffers every call; only the staging buffers' contents and *t_dev change).
static int run_program(std::vector<Fn>& ops, double* stats, size_t stat_bytes, cudaGraphExec_t* exec,
                       long long* graph_launches, cudaStream_t s) {
  auto direct = [&](cudaStream_t st) -> int {
    if (stat_bytes) LOCO_CHECK_CUDA(cudaMemsetAsync(stats, 0, stat_bytes, st));
    for (auto& f : ops) LOCO_TRY(f(st));
    return 0;
  };
  if (!graphs_enabled()) return direct(s);
  if (*exec == nullptr) {
    // capture on a private stream (the caller's may be the legacy default stream, which cannot
    // be captured); nothing executes during capture, the graph is then launched on `s`
    static cudaStream_t caps[kMaxDevices] = {nullptr};
    const int cdev = current_device();
    cudaStream_t& cap = caps[cdev < kMaxDevices ? cdev : 0];
    if (!cap) LOCO_CHECK_CUDA(cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    const long long l0 = launch_count();
    LOCO_CHECK_CUDA(cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal));
    const int r = direct(cap);
    const cudaError_t e = cudaStreamEndCapture(cap, &graph);
    if (r != 0) { if (graph) cudaGraphDestroy(graph); return r; }
    LOCO_CHECK_CUDA(e);
    *graph_launches = launch_count() - l0;
    count_launch((int)-(*graph_launches));   // capture enqueued nothing
    const cudaError_t ie = cudaGraphInstantiate(exec, graph, 0);
    cudaGraphDestroy(graph);
    LOCO_CHECK_CUDA(ie);
  }
  LOCO_CHECK_CUDA(cudaGraphLaunch(*exec, s));
  count_launch((int)*graph_launches);
  return 0;
}

int Plan::forward(const float* x, float t, float* eps_out, cudaStream_t s) {
  LOCO_REQUIRE(impl && base, "plan: not bound to a workspace");
  DeviceGuard guard(device);
  Impl& I = *impl;
  LOCO_REQUIRE(model->arch.ctx_dim == 0 || I.ctx_tokens > 0,
               "plan: this U-Net has cross-attention layers, call set_context() before forward()");
  const size_t bytes = sizeof(float) * (size_t)(NP + NT) * in_elems();
  LOCO_CHECK_CUDA(cudaMemcpyAsync(I.in_buf, x, bytes, cudaMemcpyDeviceToDevice, s));
  LOCO_TRY(set_scalar(I.t_dev, t, s));
  if (act16 && NT > 0) {
    const long long img = in_elems();
    LOCO_TRY(pow2_scale(I.in_buf + (size_t)NP * img, (long long)NT * img, 64.f, I.tscale_dev, I.scale_tmp, s));
  }
  I.cur_x = I.in_buf; I.cur_eps = I.out_buf;
  LOCO_TRY(run_program(I.fwd, I.fstats, I.fstat_bytes, &I.fwd_graph, &I.fwd_graph_launches, s));
  LOCO_CHECK_CUDA(cudaMemcpyAsync(eps_out, I.out_buf, sizeof(float) * (size_t)(NP + NT) * out_elems(), cudaMemcpyDeviceToDevice, s));
  return 0;
}

int Plan::set_context(const float* ctx, int n_tok, cudaStream_t s) {
  LOCO_REQUIRE(impl && base, "plan: not bound to a workspace");
  const Arch& A = model->arch;
  LOCO_REQUIRE(A.ctx_dim > 0, "plan: this architecture has no cross-attention layers");
  LOCO_REQUIRE(ctx != nullptr && n_tok >= 1 && n_tok <= kCtxPad, "plan: context must have 1..%d tokens (got %d)", kCtxPad, n_tok);
  DeviceGuard guard(device);
  Impl& I = *impl;
  if (n_tok != I.ctx_tokens) {
    // the token count is a launch parameter of the attention kernels: re-capture the programs
    if (I.fwd_graph) { cudaGraphExecDestroy(I.fwd_graph); I.fwd_graph = nullptr; }
    if (I.bwd_graph) { cudaGraphExecDestroy(I.bwd_graph); I.bwd_graph = nullptr; }
    I.ctx_tokens = n_tok;
  }
  for (const Impl::CrossLayer& cl : I.cross_layers)
    LOCO_TRY(context_kv(ctx, n_tok, A.ctx_dim, cl.w, cl.b, 2 * cl.C, kCtxPad, cl.kv, s));
  return 0;
}

int Plan::set_condition(const float* cond, cudaStream_t s) {
  LOCO_REQUIRE(impl && base, "plan: not bound to a workspace");
  DeviceGuard guard(device);
  const size_t bytes = sizeof(float) * 4 * (size_t)model->arch.ch;
  if (cond) LOCO_CHECK_CUDA(cudaMemcpyAsync(impl->cond_dev, cond, bytes, cudaMemcpyDeviceToDevice, s));
  else LOCO_CHECK_CUDA(cudaMemsetAsync(impl->cond_dev, 0, bytes, s));
  return 0;
}

int Plan::vjp(const float* g_eps, float* gx, cudaStream_t s) {
  LOCO_REQUIRE(impl && base, "plan: not bound to a workspace");
  LOCO_REQUIRE(NC > 0, "plan: built without cotangent rows");
  DeviceGuard guard(device);
  Impl& I = *impl;
  const size_t bytes = sizeof(float) * (size_t)NC * out_elems();
  LOCO_CHECK_CUDA(cudaMemcpyAsync(I.gin_buf, g_eps, bytes, cudaMemcpyDeviceToDevice, s));
  if (act16) {
    const long long img = out_elems();
    LOCO_TRY(pow2_scale(I.gin_buf, (long long)NC * img, 64.f, I.cscale_dev, I.scale_tmp, s));
  }
  I.cur_geps = I.gin_buf; I.cur_gx = I.gout_buf;
  LOCO_TRY(run_program(I.bwd, I.bstats, I.bstat_bytes, &I.bwd_graph, &I.bwd_graph_launches, s));
  LOCO_CHECK_CUDA(cudaMemcpyAsync(gx, I.gout_buf, sizeof(float) * (size_t)NC * in_elems(), cudaMemcpyDeviceToDevice, s));
  return 0;
}

}  // namespace loco

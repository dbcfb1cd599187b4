// SEAN/SPADE generator forward as a fixed schedule of conv_igemm launches.
//
// Follows sean_codes/models/networks/generator.py:72-109 (SPADEGenerator.forward, 'normal' upsampling),
// architecture.py:69-96 (SPADEResnetBlock) and normalization.py:108-189 (ACE) in the exactly equivalent
// region-factored form: the 512-channel piecewise-constant style map is never materialised; instead
//   mu[b,j]   = relu(fc_mu_j(code[b,j]))                                  (normalization.py:131,148)
//   Weff[b]   = alpha * conv_{gamma,beta}.weight contracted with mu[b]    -> a per-image 19-channel 3x3 kernel
//   [g | b]   = conv3x3([one_hot | actv], [Weff[b] | (1-alpha) * mlp_{gamma,beta}.weight]) + folded bias
// The nearest 2x upsample between blocks (generator.py:85-100) is folded into the readers (y>>1, x>>1).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/ctrlhair_b200.h"
#include "conv_igemm.cuh"

namespace chb {

struct TensorInfo {
  std::string name;
  int64_t offset, nbytes;
  int dtype;
};

struct AceInfo {
  int C;          // norm_nc
  int styled;
  int actv_off;   // channel offset of this ACE's actv inside the block's actv tensor
  int style_idx;  // index among styled ACEs (-1 if unstyled)
  int64_t weff_row0;  // first row of this ACE inside the per-image Weff table
  int t_gbw, t_gbb, t_chan, t_stylew;  // blob tensor ids
  int64_t noise_pix0;  // offset (in pixels per image) of this ACE's noise plane
};

struct BlockInfo {
  std::string name;
  int fin, fout, fmid, r, level, in_shift, shortcut, styled;
  AceInfo ace[3];  // order: s, 0, 1 (s unused when !shortcut)
  int n_ace;
  int t_shw, t_shb, t_c0w, t_c0b, t_c1w, t_c1b, t_csw, t_cswlo, t_c0wlo, t_c1wlo;
  int split_hs, split_h0, split_h1;  // fp16 hi+lo split of the conv_s / conv_0 / conv_1 input (chb_gen_config.precision)
  int split_w;                       // conv_0 / conv_1 weights as hi + lo: one more K-segment a_hi * w_lo (CHB_PREC_W)
  int64_t ws_xout;  // workspace offset of the block output
  int64_t ws_actv;  // this block's mlp_shared output (own buffer: the mlp_shared launches of all blocks are independent)
};

struct Step {
  ConvPlan plan;
  int kind = 0;  // 0: conv_igemm launch, 1: image from the per-tap partial sums (img_from_taps)
  std::string name;
  int noise_ace_block = -1, noise_ace = -1;  // modulate steps: which noise plane
  bool final_image = false;
};

}  // namespace chb

using namespace chb;

struct chb_generator {
  chb_gen_config cfg;
  int sw;  // latent size = crop / 32
  std::vector<TensorInfo> tensors;
  int64_t blob_bytes = 0;
  std::vector<BlockInfo> blocks;
  int n_styled = 0;
  int64_t weff_rows = 0;      // rows per image of the Weff table (each row = 32 fp16)
  int64_t noise_pix = 0;      // noise floats per image
  int t_fcw, t_fcb, t_imgw, t_imgb, t_fcmuw, t_fcmub, t_imgwy = -1, t_imgwylo = -1;
  int64_t ws_y = 0;  // CHB_PREC_IMG: per-tap partial sums of conv_img, fp32 [B,S,S,32]
  int64_t ws_ks = 0; // split-K workspace (tile counters + partial accumulators) shared by the launches that split
  // workspace layout (byte offsets, sized for max_batch)
  int64_t ws_bytes = 0;
  int64_t ws_labels, ws_codes32, ws_out, ws_codes16, ws_noise, ws_mu, ws_weff, ws_x0;
  int64_t ws_labels_b, ws_codes32_b, ws_out_b;  // second set of host-I/O buffers (forward_host_async double buffering)
  // forward_host_async state: H2D and D2H run on their own streams so they overlap the neighbouring batches' compute
  cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  uint64_t async_calls = 0;
  int64_t ws_onehot[6];
  int64_t ws_actv, ws_hs, ws_h0, ws_h1, ws_dx0;
  const uint8_t* blob = nullptr;
  uint8_t* ws = nullptr;
  std::map<int, std::vector<Step>> plans;  // keyed by batch size
  // chb_generator_forward_graph: the whole schedule of a batch size captured once (device-drawn noise; its seed node
  // is re-parameterised per call), inputs / image staged through the first set of I/O buffers of the workspace
  struct GraphEntry {
    cudaGraph_t graph = nullptr;  // kept alive: node handles (noise_node) belong to it
    cudaGraphExec_t exec = nullptr;
    cudaGraphNode_t noise_node = nullptr;
    cudaKernelNodeParams noise_params;
    float* nz = nullptr;
    long long nz_n = 0;
    uint64_t seed = 0, offset = 0;
    void* args[4];
  };
  std::map<int, GraphEntry*> graphs;
  std::vector<cudaStream_t> side;  // capture-time side streams of chb_generator_forward_graph
  std::map<std::string, std::pair<int64_t, int>> debug;  // name -> (ws offset, dtype)
  int step_limit = -1;  // debug: run only the first n conv steps
};

namespace chb {

static int add_tensor(chb_generator* g, const std::string& name, int64_t nbytes, int dtype) {
  TensorInfo t;
  t.name = name;
  t.offset = g->blob_bytes;
  t.nbytes = nbytes;
  t.dtype = dtype;
  g->tensors.push_back(t);
  g->blob_bytes += (nbytes + 255) / 256 * 256;
  return (int)g->tensors.size() - 1;
}

static int64_t ws_alloc(chb_generator* g, int64_t nbytes) {
  const int64_t off = g->ws_bytes;
  g->ws_bytes += (nbytes + 1023) / 1024 * 1024;
  return off;
}

// Interactive batches (B <= 4) are bound by launch latency, not by MMA work: under the default policy (every h_1 split)
// they also store h_0 as a hi+lo pair — conv_0's K doubles for ~3 % of a one-image forward, and the worst max-norm over
// the test inputs drops from 1.0e-3 to 7.6e-4 (the large-batch schedule keeps the policy as configured).
constexpr int kSmallBatch = 4;
static bool small_batch_h0_split(const chb_gen_config& c, int B) {
  unsigned all_h1 = 0;
  for (int i = 0; i < 7; ++i) all_h1 |= CHB_PREC_H1(i);
  return B <= kSmallBatch && (c.precision & all_h1) == all_h1;
}

static void build_layout(chb_generator* g) {
  const chb_gen_config& c = g->cfg;
  const int nf = c.ngf, L = c.style_len, B = c.max_batch;
  g->sw = c.crop / 32;
  struct Spec { const char* name; int fin, fout, rmul, styled; };
  const Spec specs[7] = {{"head_0", 16, 16, 1, 1},     {"G_middle_0", 16, 16, 2, 1}, {"G_middle_1", 16, 16, 2, 1},
                         {"up_0", 16, 8, 4, 1},        {"up_1", 8, 4, 8, 1},         {"up_2", 4, 2, 16, 1},
                         {"up_3", 2, 1, 32, 0}};
  g->t_fcw = add_tensor(g, "fc.w", (int64_t)16 * nf * 9 * 32 * 2, CHB_F16);
  g->t_fcb = add_tensor(g, "fc.b", (int64_t)16 * nf * 4, CHB_F32);
  int prev_r = g->sw;
  int64_t noise_pix = 0;
  for (int i = 0; i < 7; ++i) {
    BlockInfo b;
    b.name = specs[i].name;
    b.fin = specs[i].fin * nf;
    b.fout = specs[i].fout * nf;
    b.fmid = b.fin < b.fout ? b.fin : b.fout;
    b.r = g->sw * specs[i].rmul;
    b.level = 0;
    while ((g->sw << b.level) < b.r) ++b.level;
    b.in_shift = (b.r == prev_r) ? 0 : 1;
    prev_r = b.r;
    b.shortcut = b.fin != b.fout;
    b.styled = specs[i].styled;
    b.split_hs = (b.shortcut && (c.precision & CHB_PREC_SHORTCUT)) ? 1 : 0;
    b.split_h0 = (c.precision & CHB_PREC_H0(i)) ? 1 : 0;
    b.split_h1 = (c.precision & CHB_PREC_H1(i)) ? 1 : 0;
    b.split_w = (c.precision & CHB_PREC_W(i)) ? 1 : 0;
    b.n_ace = b.shortcut ? 3 : 2;
    const char* an[3] = {"ace_s", "ace_0", "ace_1"};
    int actv_off = 0;
    b.t_shw = add_tensor(g, b.name + ".sh.w", (int64_t)128 * b.n_ace * 9 * 32 * 2, CHB_F16);
    b.t_shb = add_tensor(g, b.name + ".sh.b", (int64_t)128 * b.n_ace * 4, CHB_F32);
    for (int a = 0; a < 3; ++a) {
      AceInfo& A = b.ace[a];
      memset(&A, 0, sizeof A);
      A.style_idx = -1;
      if (a == 0 && !b.shortcut) continue;
      A.C = (a == 2) ? b.fmid : b.fin;
      A.styled = b.styled;
      A.actv_off = actv_off;
      actv_off += 128;
      A.noise_pix0 = noise_pix;
      noise_pix += (int64_t)b.r * b.r;
      const std::string p = b.name + "." + an[a];
      A.t_gbw = add_tensor(g, p + ".gb.w", (int64_t)2 * A.C * 9 * 128 * 2, CHB_F16);
      A.t_gbb = add_tensor(g, p + ".gb.b", (int64_t)2 * A.C * 4, CHB_F32);
      A.t_chan = add_tensor(g, p + ".chan", (int64_t)A.C * 12, CHB_F32);
      A.t_stylew = -1;
      if (A.styled) {
        A.style_idx = g->n_styled++;
        A.weff_row0 = g->weff_rows;
        g->weff_rows += (int64_t)2 * A.C * 9;
      }
    }
    b.t_c0w = add_tensor(g, b.name + ".conv_0.w", (int64_t)b.fmid * 9 * b.fin * 2, CHB_F16);
    b.t_c0b = add_tensor(g, b.name + ".conv_0.b", (int64_t)b.fmid * 4, CHB_F32);
    b.t_c1w = add_tensor(g, b.name + ".conv_1.w", (int64_t)b.fout * 9 * b.fmid * 2, CHB_F16);
    b.t_c1b = add_tensor(g, b.name + ".conv_1.b", (int64_t)b.fout * 4, CHB_F32);
    b.t_csw = b.shortcut ? add_tensor(g, b.name + ".conv_s.w", (int64_t)b.fout * b.fin * 2, CHB_F16) : -1;
    b.t_cswlo = b.split_hs ? add_tensor(g, b.name + ".conv_s.wlo", (int64_t)b.fout * b.fin * 2, CHB_F16) : -1;
    b.t_c0wlo = b.split_w ? add_tensor(g, b.name + ".conv_0.wlo", (int64_t)b.fmid * 9 * b.fin * 2, CHB_F16) : -1;
    b.t_c1wlo = b.split_w ? add_tensor(g, b.name + ".conv_1.wlo", (int64_t)b.fout * 9 * b.fmid * 2, CHB_F16) : -1;
    g->blocks.push_back(b);
  }
  g->noise_pix = noise_pix;
  // style weights: one tensor per styled ACE, laid out back to back
  for (auto& b : g->blocks)
    for (int a = 0; a < 3; ++a) {
      AceInfo& A = b.ace[a];
      if (A.C && A.styled) {
        const char* an[3] = {"ace_s", "ace_0", "ace_1"};
        A.t_stylew = add_tensor(g, b.name + "." + an[a] + ".style.w", (int64_t)2 * A.C * 9 * L * 2, CHB_F16);
      }
    }
  g->t_fcmuw = add_tensor(g, "fcmu.w", (int64_t)c.label_nc * g->n_styled * L * L * 2, CHB_F16);
  g->t_fcmub = add_tensor(g, "fcmu.b", (int64_t)c.label_nc * g->n_styled * L * 4, CHB_F32);
  g->t_imgw = add_tensor(g, "conv_img.w", (int64_t)16 * 9 * nf * 2, CHB_F16);
  g->t_imgb = add_tensor(g, "conv_img.b", (int64_t)16 * 4, CHB_F32);
  if (c.precision & CHB_PREC_IMG) {
    // conv_img as a 1x1 GEMM onto 27 (tap, out channel) partial sums + a 9-neighbour gather: rows tap*3 + co
    g->t_imgwy = add_tensor(g, "conv_img.wy", (int64_t)32 * nf * 2, CHB_F16);
    g->t_imgwylo = add_tensor(g, "conv_img.wylo", (int64_t)32 * nf * 2, CHB_F16);
  }

  // ---------------- workspace
  const int64_t S = c.crop;
  g->ws_labels = ws_alloc(g, (int64_t)B * S * S);
  g->ws_codes32 = ws_alloc(g, (int64_t)B * c.label_nc * L * 4);
  g->ws_out = ws_alloc(g, (int64_t)B * 3 * S * S * 4);
  g->ws_labels_b = ws_alloc(g, (int64_t)B * S * S);
  g->ws_codes32_b = ws_alloc(g, (int64_t)B * c.label_nc * L * 4);
  g->ws_out_b = ws_alloc(g, (int64_t)B * 3 * S * S * 4);
  g->ws_codes16 = ws_alloc(g, (int64_t)B * c.label_nc * L * 2);
  g->ws_noise = ws_alloc(g, (int64_t)B * g->noise_pix * 4);
  g->ws_mu = ws_alloc(g, (int64_t)g->n_styled * B * 32 * L * 2);
  g->ws_weff = ws_alloc(g, (int64_t)B * g->weff_rows * 32 * 2);
  for (int l = 0; l < 6; ++l) {
    const int64_t r = (int64_t)g->sw << l;
    g->ws_onehot[l] = ws_alloc(g, (int64_t)B * r * r * 32 * 2);
  }
  g->ws_x0 = ws_alloc(g, (int64_t)B * g->sw * g->sw * 16 * nf * 4);
  g->debug["x_fc"] = {g->ws_x0, CHB_F32};
  int64_t m_hs = 0, m_hin = 0, m_h1 = 0, m_dx0 = 0;
  for (auto& b : g->blocks) {
    const int64_t px = (int64_t)B * b.r * b.r;
    b.ws_actv = ws_alloc(g, px * 128 * b.n_ace * 2);
    m_hs = std::max<int64_t>(m_hs, px * b.fin * 2 * (b.split_hs ? 2 : 1));
    // (interactive batches split h_0 whatever the policy says, see small_batch_h0_split: room for [hi | lo] at B <= 4)
    m_hin = std::max<int64_t>(m_hin, px * b.fin * 2 * (b.split_h0 ? 2 : 1));
    m_hin = std::max<int64_t>(m_hin, (int64_t)std::min(B, kSmallBatch) * b.r * b.r * b.fin * 2 * 2);
    m_h1 = std::max<int64_t>(m_h1, px * b.fmid * 2 * (b.split_h1 ? 2 : 1));
    m_dx0 = std::max<int64_t>(m_dx0, px * b.fmid * 4);
    const bool last = (&b == &g->blocks.back());
    // the last block's output feeds conv_img as an fp16 operand ([hi | lo] with CHB_PREC_IMG)
    b.ws_xout = ws_alloc(g, px * b.fout * (last ? ((c.precision & CHB_PREC_IMG) ? 4 : 2) : 4));
    g->debug["x_" + b.name] = {b.ws_xout, last ? CHB_F16 : CHB_F32};
  }
  g->ws_actv = g->blocks.back().ws_actv;
  g->ws_ks = ws_alloc(g, (int64_t)ksplit_workspace_bytes(device_sm_count()));
  g->ws_hs = ws_alloc(g, m_hs);
  g->ws_h0 = ws_alloc(g, m_hin);
  if (c.precision & CHB_PREC_IMG) {
    g->ws_y = ws_alloc(g, (int64_t)B * S * S * 32 * 4);
    g->debug["y_img"] = {g->ws_y, CHB_F32};
  }
  g->ws_h1 = ws_alloc(g, m_h1);
  g->ws_dx0 = ws_alloc(g, m_dx0);
  g->debug["actv"] = {g->ws_actv, CHB_F16};
  g->debug["h_s"] = {g->ws_hs, CHB_F16};
  g->debug["h_0"] = {g->ws_h0, CHB_F16};
  g->debug["h_1"] = {g->ws_h1, CHB_F16};
  g->debug["dx0"] = {g->ws_dx0, CHB_F32};
  g->debug["mu"] = {g->ws_mu, CHB_F16};
  g->debug["weff"] = {g->ws_weff, CHB_F16};
  g->debug["noise"] = {g->ws_noise, CHB_F32};
  g->debug["out"] = {g->ws_out, CHB_F32};
  for (int l = 0; l < 6; ++l) g->debug["onehot" + std::to_string(l)] = {g->ws_onehot[l], CHB_F16};
}

static const void* blobp(const chb_generator* g, int t) { return g->blob + g->tensors[t].offset; }

static void tile_for(int r, int* TW, int* TH) {
  *TW = r < 8 ? r : 8;
  *TH = r < 16 ? r : 16;
  if (*TW * *TH > 128) *TH = 128 / *TW;
}

static chb_conv_seg make_seg(const void* a, int r, int Ca, int ch_off, int C, int taps, const void* w) {
  chb_conv_seg s;
  memset(&s, 0, sizeof s);
  s.a = a;
  s.a_sx = Ca;
  s.a_sy = (int64_t)r * Ca;
  s.a_sb = (int64_t)r * r * Ca;
  s.Ca = Ca;
  s.ch_off = ch_off;
  s.C = C;
  s.taps = taps;
  s.w = w;
  return s;
}

// N tile of a PLAIN conv.  256 (or the whole width) at production batch sizes; for a handful of images (interactive
// B = 1: hair_editor.py:159-179) the 8x8 .. 32x32 layers have 1-8 pixel tiles, so the N tile narrows (down to 64) until
// ~64 CTAs share the streaming of the layer's weight matrix (19 MB for a 1024 -> 1024 conv).  The K order inside a tile
// does not depend on BN, so results stay bitwise identical across batch sizes.
static int pick_bn(int N, int B, int r) {
  int bn = N < 256 ? N : 256;
  int tw, th;
  tw = r < 8 ? r : 8; th = r < 16 ? r : 16;
  const long long m_tiles = (long long)B * ((r + tw - 1) / tw) * ((r + th - 1) / th);
  while (bn > 64 && bn % 128 == 0 && m_tiles * (N / bn) < 64) bn /= 2;
  return bn;
}

static chb_conv_desc base_desc(int B, int r) {
  chb_conv_desc d;
  memset(&d, 0, sizeof d);
  d.B = B; d.H = r; d.W = r;
  tile_for(r, &d.TW, &d.TH);
  d.TB = 1;
  return d;
}

// Images smaller than a 128-pixel tile (8x8 at crop 256): a PLAIN conv with shared weights batches several images into
// one 128-row MMA tile instead of issuing half-empty MMAs.  Rows of an MMA are independent, so an image's result does
// not depend on its tile-mates (the batch-invariance tests stay bitwise).
static void batch_small_tiles(chb_conv_desc* d) {
  const int px = d->TW * d->TH;
  if (px >= 128) return;
  int tb = 128 / px;
  if (tb > d->B) tb = d->B;
  d->TB = tb < 1 ? 1 : tb;
}

// A handful of images (interactive B = 1 .. 4): the 8x8 .. 32x32 PLAIN convs have fewer output tiles than the GPU has
// SMs, and one CTA walks the whole K of its tile: 144-288 K steps of 64 at ~0.19 us = 28-55 us per launch, which is
// what the latency of a one-image forward consisted of (tools/gpu_chain_floor.py, gpu_trace_small.py).  Split-K lets
// 4-8 CTAs share an output tile.  The combine (partial tile through L2, two fences, a counter, the in-order sum) costs
// ~7 us, so a launch only splits when it has >= 96 K steps and at least four ways to go; an N tile that pick_bn
// narrowed to get more CTAs is widened again when that buys the four splits (wider MMAs, fewer reads of the A tile).
// The partial sums are added in a fixed order (deterministic), but the fp32 association differs from an unsplit launch:
// a one-image call and the same image inside a large batch agree to fp16 storage rounding of the activations, no longer
// bitwise.  At production batch sizes nothing splits.
static void split_k_few_tiles(const chb_generator* g, chb_conv_desc* d) {
  const long long m_tiles = (long long)((d->W + d->TW - 1) / d->TW) * ((d->H + d->TH - 1) / d->TH) *
                            ((d->B + d->TB - 1) / d->TB);
  int min_chunks = 1 << 30, steps = 0;
  for (int i = 0; i < d->nseg; ++i) {
    const int nch = d->seg[i].C / (d->seg[i].C == 32 ? 32 : 64);
    if (nch < min_chunks) min_chunks = nch;
    steps += d->seg[i].taps * nch;
  }
  if (steps < 96) return;
  const int sms = device_sm_count();
  auto ways = [&](int bn) {
    long long s = sms / (m_tiles * (d->Nrows / bn));
    if (s > 8) s = 8;
    if (s > min_chunks / 2) s = min_chunks / 2;
    return (int)s;
  };
  int bn = d->BN, s = ways(bn);
  if (s < 4 && bn < 256 && d->Nrows % (2 * bn) == 0 && ways(2 * bn) >= 4) {
    bn *= 2;
    s = ways(bn);
  }
  if (s < 4) return;
  d->BN = bn;
  d->ksplit = s;
  d->ks_ws = g->ws + g->ws_ks;
}

static void nhwc_out(chb_conv_desc* d, void* out, int dtype, int r, int C) {
  d->out = out; d->out_dtype = dtype;
  d->o_sn = 1; d->o_sx = C; d->o_sy = (int64_t)r * C; d->o_sb = (int64_t)r * r * C;
}

static int push_step(std::vector<Step>& steps, const chb_conv_desc& d, const std::string& name, int nb = -1,
                     int na = -1, bool fin = false) {
  Step s;
  s.name = name;
  int rc = build_conv_plan(d, &s.plan);
  if (rc != CHB_OK) return rc;
  s.noise_ace_block = nb;
  s.noise_ace = na;
  s.final_image = fin;
  steps.push_back(s);
  return CHB_OK;
}

static int build_steps(chb_generator* g, int B, std::vector<Step>& steps) {
  const chb_gen_config& c = g->cfg;
  const int nf = c.ngf, L = c.style_len, NC = c.label_nc;
  uint8_t* ws = g->ws;
  int rc;
  // ---- style path 1: mu[a][b][j][:] = relu(fc_mu_j(code[b][j]))  — one grouped launch, image == class j
  if (g->n_styled > 0) {
    chb_conv_desc d;
    memset(&d, 0, sizeof d);
    d.B = NC; d.H = 1; d.W = B;
    d.TW = B >= 128 ? 128 : (B + 7) / 8 * 8; d.TH = 1; d.TB = 1;
    d.nseg = 1;
    chb_conv_seg& s = d.seg[0];
    memset(&s, 0, sizeof s);
    s.a = ws + g->ws_codes16;
    s.a_sx = L; s.a_sy = (int64_t)B * L; s.a_sb = (int64_t)B * L;  // codes16 is [class][image][L]
    s.Ca = L; s.ch_off = 0; s.C = L; s.taps = 1;
    s.w = blobp(g, g->t_fcmuw);
    s.per_image = 1;
    d.N = d.Nrows = g->n_styled * L;
    d.BN = 256;
    d.epi = CHB_EPI_PLAIN; d.act = CHB_ACT_RELU;
    d.bias = reinterpret_cast<const float*>(blobp(g, g->t_fcmub));
    d.bias_per_image = 1;
    d.out = ws + g->ws_mu; d.out_dtype = CHB_F16;
    // mu is stored [ace][image][32 classes (19 real, rest stay zero)][L]
    d.o_sb = L; d.o_sy = 0; d.o_sx = (int64_t)32 * L; d.o_sn = 1;
    d.o_ngroup = L; d.o_sgroup = (int64_t)B * 32 * L;
    if ((rc = push_step(steps, d, "fc_mu")) != CHB_OK) return rc;
  }
  // ---- style path 2: Weff[b][n*9+tap][j] = sum_ci Wstyle[n*9+tap][ci] * mu[b][j][ci]  (per styled ACE).
  // GEMM roles are swapped (M = weight rows n*9+tap, N = (image, class)) so that every thread of the epilogue
  // owns one Weff row and writes its 32 class columns as one contiguous 64-byte run.
  int gimg = 1;
  while (gimg < 8 && B % (gimg * 2) == 0) gimg *= 2;
  // Consecutive styled ACEs of equal width share ONE launch: their style weights lie back to back in the blob, their mu
  // tables back to back in the workspace and their Weff rows back to back in the per-image table, so the ACE index is the
  // operator's "image" dimension (15 launches -> 4: 8 x C=1024, 3 x 512, 3 x 256, 1 x 128).
  {
    std::vector<const AceInfo*> styled;
    for (auto& b : g->blocks)
      for (int a = 0; a < 3; ++a)
        if (b.ace[a].C && b.ace[a].styled) styled.push_back(&b.ace[a]);
    for (size_t i = 0; i < styled.size();) {
      size_t j = i;
      while (j < styled.size() && styled[j]->C == styled[i]->C) ++j;
      const AceInfo& A = *styled[i];
      const int n_grp = (int)(j - i);
      const int rows = 2 * A.C * 9;
      chb_conv_desc d;
      memset(&d, 0, sizeof d);
      d.B = n_grp; d.H = 1; d.W = rows;
      d.TW = 128; d.TH = 1; d.TB = 1;
      d.nseg = 1;
      chb_conv_seg& s = d.seg[0];
      memset(&s, 0, sizeof s);
      s.a = blobp(g, A.t_stylew);
      s.a_sx = L; s.a_sy = (int64_t)rows * L; s.a_sb = (int64_t)rows * L;
      s.Ca = L; s.C = L; s.taps = 1;
      s.w = ws + g->ws_mu + (int64_t)A.style_idx * B * 32 * L * 2;
      s.per_image = 1;
      s.w_sb = (int64_t)B * 32 * L;
      d.N = d.Nrows = B * 32;
      d.BN = 32 * gimg;
      d.epi = CHB_EPI_PLAIN; d.act = CHB_ACT_NONE;
      d.out = ws + g->ws_weff + A.weff_row0 * 32 * 2; d.out_dtype = CHB_F16;
      d.o_sb = (int64_t)rows * 32; d.o_sy = 0; d.o_sx = 32; d.o_sn = 1;
      d.o_ngroup = 32; d.o_sgroup = g->weff_rows * 32;
      for (size_t k = i; k + 1 < j; ++k) {  // the layout facts this grouping relies on
        if (styled[k + 1]->style_idx != styled[k]->style_idx + 1 || styled[k + 1]->weff_row0 != styled[k]->weff_row0 + rows ||
            g->tensors[styled[k + 1]->t_stylew].offset != g->tensors[styled[k]->t_stylew].offset + (int64_t)rows * L * 2) {
          set_error("generator: styled ACE tables are not contiguous");
          return CHB_ERR_ARG;
        }
      }
      char nm[64];
      snprintf(nm, sizeof nm, "weff[%d ACEs x C=%d]", n_grp, A.C);
      if ((rc = push_step(steps, d, nm)) != CHB_OK) return rc;
      i = j;
    }
  }
  // ---- x = fc(one_hot @ sw)   (generator.py:75-76)
  {
    chb_conv_desc d = base_desc(B, g->sw);
    d.nseg = 1;
    d.seg[0] = make_seg(ws + g->ws_onehot[0], g->sw, 32, 0, 32, 9, blobp(g, g->t_fcw));
    d.N = d.Nrows = 16 * nf; d.BN = pick_bn(16 * nf, B, g->sw);
    d.epi = CHB_EPI_PLAIN; d.act = CHB_ACT_NONE;
    d.bias = nullptr;  // carried by the constant one-hot channels (fc.w centre tap); fc.b stays in the blob for checkers
    nhwc_out(&d, ws + g->ws_x0, CHB_F32, g->sw, 16 * nf);
    if ((rc = push_step(steps, d, "fc")) != CHB_OK) return rc;
  }
  const float* xin = reinterpret_cast<const float*>(ws + g->ws_x0);
  int xin_r = g->sw;
  for (size_t bi = 0; bi < g->blocks.size(); ++bi) {
    const BlockInfo& b = g->blocks[bi];
    const bool last = bi + 1 == g->blocks.size();
    const int r = b.r, actvC = 128 * b.n_ace;
    const void* onehot = ws + g->ws_onehot[b.level];
    // actv = relu(mlp_shared(seg)) for every ACE of the block at once (normalization.py:253)
    {
      chb_conv_desc d = base_desc(B, r);
      d.nseg = 1;
      d.seg[0] = make_seg(onehot, r, 32, 0, 32, 9, blobp(g, b.t_shw));
      d.N = d.Nrows = actvC; d.BN = actvC == 384 ? 192 : 256;  // one-hot A tile is re-read per N tile: keep N tiles few
      d.epi = CHB_EPI_PLAIN; d.act = CHB_ACT_RELU;
      d.bias = nullptr;  // carried by the constant one-hot channels (sh.w centre tap)
      nhwc_out(&d, ws + b.ws_actv, CHB_F16, r, actvC);
      if ((rc = push_step(steps, d, b.name + ".mlp_shared")) != CHB_OK) return rc;
    }
    auto modulate = [&](int a, const float* x, int x_r, int x_shift, int xC, void* hout, int act, int split) -> int {
      const AceInfo& A = b.ace[a];
      chb_conv_desc d = base_desc(B, r);
      int ns = 0;
      if (A.styled) {
        d.seg[ns] = make_seg(onehot, r, 32, 0, 32, 9, ws + g->ws_weff + A.weff_row0 * 32 * 2);
        d.seg[ns].per_image = 1;
        d.seg[ns].w_sb = g->weff_rows * 32;
        ++ns;
      }
      d.seg[ns++] = make_seg(ws + b.ws_actv, r, actvC, A.actv_off, 128, 9, blobp(g, A.t_gbw));
      d.nseg = ns;
      d.N = d.Nrows = 2 * A.C;
      d.BN = d.N < 256 ? d.N : 256;
      d.epi = CHB_EPI_MODULATE; d.act = act;
      d.bias = reinterpret_cast<const float*>(blobp(g, A.t_gbb));
      d.chan = reinterpret_cast<const float*>(blobp(g, A.t_chan));
      d.x = x; d.x_shift = x_shift;
      d.x_sx = xC; d.x_sy = (int64_t)x_r * xC; d.x_sb = (int64_t)x_r * x_r * xC;
      d.noise = reinterpret_cast<const float*>(ws + g->ws_noise) + (int64_t)B * A.noise_pix0;
      nhwc_out(&d, hout, CHB_F16, r, A.C * (split ? 2 : 1));  // split: pixel row = [hi(C) | lo(C)]
      d.o_split = split; d.o_lo_off = split ? A.C : 0;
      const char* an3[3] = {"ace_s", "ace_0", "ace_1"};
      return push_step(steps, d, b.name + "." + an3[a] + ".gamma_beta_mod", (int)bi, a);
    };
    if (b.shortcut) {
      if ((rc = modulate(0, xin, xin_r, b.in_shift, b.fin, ws + g->ws_hs, CHB_ACT_NONE, b.split_hs)) != CHB_OK) return rc;
    }
    const int sh0 = (b.split_h0 || small_batch_h0_split(c, B)) ? 1 : 0;
    if ((rc = modulate(1, xin, xin_r, b.in_shift, b.fin, ws + g->ws_h0, CHB_ACT_LRELU, sh0)) != CHB_OK) return rc;
    // dx = conv_0(lrelu(ace_0(x)))   (architecture.py:73-75)
    {
      chb_conv_desc d = base_desc(B, r);
      batch_small_tiles(&d);
      d.nseg = 1;
      const int m0 = sh0 ? 2 : 1;  // [hi | lo] halves share conv_0's weights
      d.seg[0] = make_seg(ws + g->ws_h0, r, b.fin * m0, 0, b.fin * m0, 9, blobp(g, b.t_c0w));
      d.seg[0].w_dup = m0;
      if (b.split_w) d.seg[d.nseg++] = make_seg(ws + g->ws_h0, r, b.fin * m0, 0, b.fin, 9, blobp(g, b.t_c0wlo));  // h_hi * w_lo
      d.N = d.Nrows = b.fmid; d.BN = pick_bn(b.fmid, B, r);
      d.epi = CHB_EPI_PLAIN; d.act = CHB_ACT_NONE;
      d.bias = reinterpret_cast<const float*>(blobp(g, b.t_c0b));
      nhwc_out(&d, ws + g->ws_dx0, CHB_F32, r, b.fmid);
      split_k_few_tiles(g, &d);
      if ((rc = push_step(steps, d, b.name + ".conv_0")) != CHB_OK) return rc;
    }
    if ((rc = modulate(2, reinterpret_cast<const float*>(ws + g->ws_dx0), r, 0, b.fmid, ws + g->ws_h1,
                       CHB_ACT_LRELU, b.split_h1)) != CHB_OK)
      return rc;
    // out = x_s + conv_1(lrelu(ace_1(dx)))   (architecture.py:77-84); conv_s rides along as a 1x1 K-segment
    {
      chb_conv_desc d = base_desc(B, r);
      batch_small_tiles(&d);
      int ns = 0;
      const int m1 = b.split_h1 ? 2 : 1;
      d.seg[ns] = make_seg(ws + g->ws_h1, r, b.fmid * m1, 0, b.fmid * m1, 9, blobp(g, b.t_c1w));
      d.seg[ns++].w_dup = m1;
      if (b.shortcut && b.split_hs) {
        // x_s = conv_s(h_s) to ~2^-22: (hs_hi + hs_lo) * ws_hi + hs_hi * ws_lo, three 1x1 K-segments' worth
        d.seg[ns] = make_seg(ws + g->ws_hs, r, b.fin * 2, 0, b.fin * 2, 1, blobp(g, b.t_csw));
        d.seg[ns++].w_dup = 2;
        d.seg[ns++] = make_seg(ws + g->ws_hs, r, b.fin * 2, 0, b.fin, 1, blobp(g, b.t_cswlo));
      } else if (b.shortcut) {
        d.seg[ns++] = make_seg(ws + g->ws_hs, r, b.fin, 0, b.fin, 1, blobp(g, b.t_csw));
      } else {
        d.res = xin; d.r_shift = b.in_shift;
        d.r_sx = b.fin; d.r_sy = (int64_t)xin_r * b.fin; d.r_sb = (int64_t)xin_r * xin_r * b.fin;
      }
      if (b.split_w) d.seg[ns++] = make_seg(ws + g->ws_h1, r, b.fmid * m1, 0, b.fmid, 9, blobp(g, b.t_c1wlo));  // h_hi * w_lo
      d.nseg = ns;
      d.N = d.Nrows = b.fout; d.BN = pick_bn(b.fout, B, r);
      d.epi = CHB_EPI_PLAIN;
      d.act = last ? CHB_ACT_LRELU : CHB_ACT_NONE;  // generator.py:107 leaky_relu before conv_img
      d.bias = reinterpret_cast<const float*>(blobp(g, b.t_c1b));
      const int lsplit = (last && (c.precision & CHB_PREC_IMG)) ? 1 : 0;
      nhwc_out(&d, ws + b.ws_xout, last ? CHB_F16 : CHB_F32, r, b.fout * (lsplit ? 2 : 1));
      d.o_split = lsplit; d.o_lo_off = lsplit ? b.fout : 0;
      split_k_few_tiles(g, &d);
      if ((rc = push_step(steps, d, b.name + (b.shortcut ? ".conv_1+conv_s" : ".conv_1+res"))) != CHB_OK) return rc;
    }
    xin = reinterpret_cast<const float*>(ws + b.ws_xout);
    xin_r = r;
  }
  // ---- image = tanh(conv_img(lrelu(x)))   (generator.py:107-108), fp32 NCHW
  if (c.precision & CHB_PREC_IMG) {
    // conv_img has 3 output channels: as an implicit GEMM with N = 16 every MMA still reads a full 128-row A tile per
    // K step (A-bandwidth bound).  Instead ONE 1x1 GEMM computes the 27 per-(tap, channel) partial sums
    // Y[p][tap*3+co] = sum_ci x[p][ci] * W[co][ci][tap] (K = 64 instead of 576), on the hi+lo split input and weights
    // (x_hi + x_lo) * w_hi + x_hi * w_lo, and a small gather kernel adds the nine neighbours' sums, the bias and tanh.
    const BlockInfo& b = g->blocks.back();
    const int r = b.r;
    chb_conv_desc d = base_desc(B, r);
    d.nseg = 2;
    d.seg[0] = make_seg(ws + b.ws_xout, r, 2 * b.fout, 0, 2 * b.fout, 1, blobp(g, g->t_imgwy));
    d.seg[0].w_dup = 2;
    d.seg[1] = make_seg(ws + b.ws_xout, r, 2 * b.fout, 0, b.fout, 1, blobp(g, g->t_imgwylo));
    d.N = d.Nrows = 32; d.BN = 32;
    d.epi = CHB_EPI_PLAIN; d.act = CHB_ACT_NONE;
    nhwc_out(&d, ws + g->ws_y, CHB_F32, r, 32);
    if ((rc = push_step(steps, d, "conv_img.taps")) != CHB_OK) return rc;
    Step s;
    s.name = "conv_img.gather+tanh";
    s.kind = 1;
    s.final_image = true;
    memset(&s.plan.kp, 0, sizeof s.plan.kp);
    memset(&s.plan.desc, 0, sizeof s.plan.desc);
    s.plan.grid = 0; s.plan.smem_bytes = 0; s.plan.flops = 0;
    steps.push_back(s);
  } else {
    const BlockInfo& b = g->blocks.back();
    const int r = b.r;
    chb_conv_desc d = base_desc(B, r);
    d.nseg = 1;
    d.seg[0] = make_seg(ws + b.ws_xout, r, b.fout, 0, b.fout, 9, blobp(g, g->t_imgw));
    d.N = 3; d.Nrows = 16; d.BN = 16;
    d.epi = CHB_EPI_PLAIN; d.act = CHB_ACT_TANH;
    d.bias = reinterpret_cast<const float*>(blobp(g, g->t_imgb));
    d.out = ws + g->ws_out; d.out_dtype = CHB_F32;
    d.o_sn = (int64_t)r * r; d.o_sx = 1; d.o_sy = r; d.o_sb = (int64_t)3 * r * r;
    if ((rc = push_step(steps, d, "conv_img", -1, -1, true)) != CHB_OK) return rc;
  }
  return CHB_OK;
}

}  // namespace chb

extern "C" {

int chb_generator_create(const chb_gen_config* cfg, chb_generator** out) {
  if (!cfg || !out) {
    set_error("chb_generator_create: NULL argument");
    return CHB_ERR_ARG;
  }
  if (cfg->ngf <= 0 || cfg->ngf % 64 != 0 || cfg->label_nc <= 0 || cfg->label_nc > 30 || cfg->crop < 32 ||
      cfg->crop % 32 != 0 || (cfg->crop & (cfg->crop - 1)) != 0 || cfg->style_len <= 0 || cfg->style_len % 64 != 0 ||
      cfg->max_batch <= 0) {
    set_error(
        "chb_generator_create: need ngf % 64 == 0, label_nc in 1..30, crop a power of two >= 32, style_len % 64 == 0, "
        "max_batch > 0");
    return CHB_ERR_ARG;
  }
  chb_generator* g = new chb_generator();
  g->cfg = *cfg;
  build_layout(g);
  *out = g;
  return CHB_OK;
}

void chb_generator_destroy(chb_generator* g) {
  if (!g) return;
  for (int i = 0; i < 2; ++i) {
    if (g->ev_in[i]) cudaEventDestroy(g->ev_in[i]);
    if (g->ev_done[i]) cudaEventDestroy(g->ev_done[i]);
    if (g->ev_copied[i]) cudaEventDestroy(g->ev_copied[i]);
  }
  if (g->h2d_stream) cudaStreamDestroy(g->h2d_stream);
  if (g->d2h_stream) cudaStreamDestroy(g->d2h_stream);
  for (auto& kv : g->graphs) {
    if (kv.second->exec) cudaGraphExecDestroy(kv.second->exec);
    if (kv.second->graph) cudaGraphDestroy(kv.second->graph);
    delete kv.second;
  }
  for (cudaStream_t st : g->side) cudaStreamDestroy(st);
  delete g;
}

int chb_generator_num_tensors(const chb_generator* g) { return g ? (int)g->tensors.size() : 0; }

int chb_generator_tensor_info(const chb_generator* g, int i, char* name, int name_cap, int64_t* offset,
                              int64_t* nbytes, int* dtype) {
  if (!g || i < 0 || i >= (int)g->tensors.size()) {
    set_error("chb_generator_tensor_info: index out of range");
    return CHB_ERR_ARG;
  }
  const TensorInfo& t = g->tensors[i];
  if (name && name_cap > 0) snprintf(name, name_cap, "%s", t.name.c_str());
  if (offset) *offset = t.offset;
  if (nbytes) *nbytes = t.nbytes;
  if (dtype) *dtype = t.dtype;
  return CHB_OK;
}

int64_t chb_generator_blob_bytes(const chb_generator* g) { return g ? g->blob_bytes : 0; }
int64_t chb_generator_workspace_bytes(const chb_generator* g) { return g ? g->ws_bytes : 0; }
int64_t chb_generator_noise_floats(const chb_generator* g, int B) { return g ? g->noise_pix * B : 0; }

int chb_generator_bind(chb_generator* g, const void* blob, void* workspace) {
  if (!g || !blob || !workspace) {
    set_error("chb_generator_bind: NULL argument");
    return CHB_ERR_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(blob) & 255) || (reinterpret_cast<uintptr_t>(workspace) & 1023)) {
    set_error("chb_generator_bind: blob must be 256-byte and workspace 1024-byte aligned");
    return CHB_ERR_ARG;
  }
  int rc = chb_check_device();
  if (rc != CHB_OK) return rc;
  g->blob = reinterpret_cast<const uint8_t*>(blob);
  g->ws = reinterpret_cast<uint8_t*>(workspace);
  g->plans.clear();
  for (auto& kv : g->graphs) {
    if (kv.second->exec) cudaGraphExecDestroy(kv.second->exec);
    if (kv.second->graph) cudaGraphDestroy(kv.second->graph);
    delete kv.second;
  }
  g->graphs.clear();
  // mu rows 19..31 (class padding) are never written by the fc_mu GEMM and must read as zero, so that the
  // Weff columns they produce are exact zeros.
  cudaError_t err = cudaMemset(g->ws + g->ws_mu, 0, (size_t)g->n_styled * g->cfg.max_batch * 32 * g->cfg.style_len * 2);
  if (err == cudaSuccess) err = cudaMemset(g->ws + g->ws_ks, 0, kKsCounterBytes);   // split-K tile counters start at zero
  if (err != cudaSuccess) {
    set_error(std::string("chb_generator_bind: cudaMemset failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

static int get_steps(chb_generator* g, int B, std::vector<Step>** out) {
  if (B <= 0 || B > g->cfg.max_batch) {  // plans address the workspace, which is sized for max_batch
    set_error("chb_generator: batch must be in 1..max_batch");
    return CHB_ERR_ARG;
  }
  auto it = g->plans.find(B);
  if (it == g->plans.end()) {
    std::vector<Step> steps;
    int rc = build_steps(g, B, steps);
    if (rc != CHB_OK) return rc;
    it = g->plans.emplace(B, std::move(steps)).first;
  }
  *out = &it->second;
  return CHB_OK;
}

static int forward_impl(chb_generator* g, const uint8_t* labels, const float* codes, const float* noise,
                        uint64_t seed, float* out, int B, int impl, void* stream_, cudaEvent_t* evs, int nev,
                        bool dag = false);
static int capture_steps_dag(chb_generator* g, std::vector<Step>& steps, int B, float* out, cudaStream_t cap);

int chb_generator_forward(chb_generator* g, const uint8_t* labels, const float* codes, const float* noise,
                          uint64_t seed, float* out, int B, int impl, void* stream_) {
  return forward_impl(g, labels, codes, noise, seed, out, B, impl, stream_, nullptr, 0);
}

int chb_generator_forward_timed(chb_generator* g, const uint8_t* labels, const float* codes, const float* noise,
                                uint64_t seed, float* out, int B, void* stream_, float* ms, double* flops,
                                int cap) {
  if (!g || !ms || !flops || !g->ws) {
    set_error("chb_generator_forward_timed: NULL argument or unbound generator");
    return CHB_ERR_ARG;
  }
  std::vector<Step>* steps = nullptr;
  int rc = get_steps(g, B, &steps);
  if (rc != CHB_OK) return rc;
  const int n = (int)steps->size();
  if (cap < n) {
    set_error("chb_generator_forward_timed: output arrays too small");
    return CHB_ERR_ARG;
  }
  std::vector<cudaEvent_t> evs(n + 1);
  for (auto& e : evs) cudaEventCreate(&e);
  rc = forward_impl(g, labels, codes, noise, seed, out, B, CHB_IMPL_TCGEN05, stream_, evs.data(), n + 1);
  if (rc == CHB_OK) {
    if (cudaEventSynchronize(evs[n]) != cudaSuccess) {
      set_error(std::string("forward_timed: ") + cudaGetErrorString(cudaGetLastError()));
      rc = CHB_ERR_CUDA;
    } else {
      for (int i = 0; i < n; ++i) {
        cudaEventElapsedTime(&ms[i], evs[i], evs[i + 1]);
        flops[i] = (*steps)[i].plan.flops;
      }
      rc = n;
    }
  }
  for (auto& e : evs) cudaEventDestroy(e);
  return rc;
}

static int forward_impl(chb_generator* g, const uint8_t* labels, const float* codes, const float* noise,
                        uint64_t seed, float* out, int B, int impl, void* stream_, cudaEvent_t* evs, int nev, bool dag) {
  if (!g || !labels || !codes || !out) {
    set_error("chb_generator_forward: NULL argument");
    return CHB_ERR_ARG;
  }
  if (!g->blob || !g->ws) {
    set_error("chb_generator_forward: generator is not bound to a weight blob / workspace");
    return CHB_ERR_ARG;
  }
  if (B <= 0 || B > g->cfg.max_batch) {
    set_error("chb_generator_forward: batch exceeds max_batch");
    return CHB_ERR_ARG;
  }
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  std::vector<Step>* steps = nullptr;
  int rc = get_steps(g, B, &steps);
  if (rc != CHB_OK) return rc;
  const chb_gen_config& c = g->cfg;
  uint8_t* ws = g->ws;
  rc = codes_cast_transpose(codes, ws + g->ws_codes16, B, c.label_nc, c.style_len, stream);
  if (rc != CHB_OK) return rc;
  int shifts[6];
  void* outs[6];
  for (int l = 0; l < 6; ++l) {
    shifts[l] = 5 - l;
    outs[l] = ws + g->ws_onehot[l];
  }
  // channels label_nc, label_nc + 1 are constant 1: they carry the fc / mlp_shared biases through the GEMM (packer.py)
  rc = onehot_pyramid_ones(labels, B, c.crop, 6, shifts, outs, c.label_nc, c.label_nc, 2, stream);
  if (rc != CHB_OK) return rc;
  float* nz = reinterpret_cast<float*>(ws + g->ws_noise);
  if (noise) {
    if (noise != nz) {
      cudaError_t err =
          cudaMemcpyAsync(nz, noise, (size_t)B * g->noise_pix * 4, cudaMemcpyDeviceToDevice, stream);
      if (err != cudaSuccess) {
        set_error(std::string("noise copy failed: ") + cudaGetErrorString(err));
        return CHB_ERR_CUDA;
      }
    }
  } else {
    rc = chb_noise_fill(nz, (int64_t)B * g->noise_pix, seed, 0, stream);
    if (rc != CHB_OK) return rc;
  }
  if (dag) return capture_steps_dag(g, *steps, B, out, stream);   // under capture: steps with their real dependencies
  int nrun = 0, iev = 0;
  if (evs && iev < nev) cudaEventRecord(evs[iev++], stream);
  for (const Step& s : *steps) {
    if (g->step_limit >= 0 && nrun++ >= g->step_limit) break;
    if (s.kind == 1) {
      rc = img_from_taps(reinterpret_cast<const float*>(ws + g->ws_y),
                         reinterpret_cast<const float*>(blobp(g, g->t_imgb)), out, B, c.crop, stream);
    } else if (s.final_image && out != reinterpret_cast<float*>(ws + g->ws_out)) {
      ConvPlan p = s.plan;
      p.kp.e.out = out;
      p.desc.out = out;
      rc = launch_conv_plan(p, impl, stream);
    } else {
      rc = launch_conv_plan(s.plan, impl, stream);
    }
    if (rc != CHB_OK) return rc;
    if (evs && iev < nev) cudaEventRecord(evs[iev++], stream);
  }
  return CHB_OK;
}

// ---- dependency-driven capture -------------------------------------------------------------------------------------
// Byte ranges (hulls of the strided accesses) a step reads / writes inside the workspace.  Weights in the blob are
// read-only and left out; per-image style tables (Weff) live in the workspace and are included.
struct ByteRange { const char* lo; const char* hi; };
static void add_br(std::vector<ByteRange>& v, const void* p, long long bytes) {
  if (p && bytes > 0) v.push_back({static_cast<const char*>(p), static_cast<const char*>(p) + bytes});
}
static bool br_overlap(const std::vector<ByteRange>& a, const std::vector<ByteRange>& b) {
  for (const ByteRange& x : a)
    for (const ByteRange& y : b)
      if (x.lo < y.hi && y.lo < x.hi) return true;
  return false;
}
static void step_ranges(const chb_generator* g, const Step& s, int B, const float* out, std::vector<ByteRange>& r,
                        std::vector<ByteRange>& w) {
  const chb_gen_config& c = g->cfg;
  if (s.kind == 1) {
    add_br(r, g->ws + g->ws_y, (long long)B * c.crop * c.crop * 32 * 4);
    add_br(w, out, (long long)B * 3 * c.crop * c.crop * 4);
    return;
  }
  const chb_conv_desc& d = s.plan.desc;
  for (int i = 0; i < d.nseg; ++i) {
    const chb_conv_seg& q = d.seg[i];
    add_br(r, q.a, ((long long)(d.B - 1) * q.a_sb + (long long)(d.H + 2 * q.a_pad - 1) * q.a_sy +
                    (long long)(d.W + 2 * q.a_pad - 1) * q.a_sx + q.Ca) * 2);
    if (q.per_image) {
      const long long K = (long long)q.taps * (q.w_dup == 2 ? q.C / 2 : q.C);
      const long long per = q.w_sb > 0 ? q.w_sb : (long long)d.Nrows * K;
      add_br(r, q.w, ((long long)(d.B - 1) * per + (long long)d.Nrows * K) * 2);
    }
  }
  const long long ob = d.out_dtype == CHB_F16 ? 2 : 4;
  long long last = (long long)(d.B - 1) * d.o_sb + (long long)(d.H - 1) * d.o_sy + (long long)(d.W - 1) * d.o_sx;
  if (d.o_ngroup > 0) last += (long long)((d.N - 1) / d.o_ngroup) * d.o_sgroup + (long long)(d.o_ngroup - 1) * d.o_sn;
  else last += (long long)(d.N - 1) * d.o_sn;
  if (d.o_split) last += d.o_lo_off;
  add_br(w, s.final_image ? static_cast<const void*>(out) : d.out, (last + 1) * ob);
  if (d.x)
    add_br(r, d.x, ((long long)(d.B - 1) * d.x_sb + (long long)((d.H - 1) >> d.x_shift) * d.x_sy +
                    (long long)((d.W - 1) >> d.x_shift) * d.x_sx + d.N / 2) * 4);
  if (d.res)
    add_br(r, d.res, ((long long)(d.B - 1) * d.r_sb + (long long)((d.H - 1) >> d.r_shift) * d.r_sy +
                      (long long)((d.W - 1) >> d.r_shift) * d.r_sx + d.N) * 4);
  if (d.noise) add_br(r, d.noise, (long long)d.B * d.W * d.H * 4);
  if (d.ksplit > 1) add_br(w, d.ks_ws, (long long)ksplit_workspace_bytes(device_sm_count()));
}

// Captures the conv schedule into the graph being recorded on `cap` with its REAL dependencies: every step goes to a
// side stream and waits for the events of the earlier steps it conflicts with (RAW / WAR / WAW on workspace ranges).  A
// linear capture makes the 48 launches one chain; with dependencies the style path (fc_mu, weff), the mlp_shared launches
// of all seven blocks and the two gamma/beta launches that read the same block input run beside the chain, which is
// what the latency of a one-image forward consists of.
static int capture_steps_dag(chb_generator* g, std::vector<Step>& steps, int B, float* out, cudaStream_t cap) {
  const int kSide = 6;
  cudaError_t err = cudaSuccess;
  while ((int)g->side.size() < kSide && err == cudaSuccess) {
    cudaStream_t st = nullptr;
    err = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (err == cudaSuccess) g->side.push_back(st);
  }
  const int n = (int)steps.size();
  std::vector<std::vector<ByteRange>> R(n), W(n);
  for (int i = 0; i < n; ++i) step_ranges(g, steps[i], B, out, R[i], W[i]);
  std::vector<cudaEvent_t> ev(n + 1, nullptr);
  for (auto& e : ev)
    if (err == cudaSuccess) err = cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  std::vector<int> stream_of(n, -1), tail(kSide, -1);
  std::vector<char> joined(kSide, 0);
  if (err == cudaSuccess) err = cudaEventRecord(ev[n], cap);   // root: the helper kernels recorded on `cap` so far
  int rc = CHB_OK, rr = 0;
  for (int i = 0; i < n && err == cudaSuccess && rc == CHB_OK; ++i) {
    std::vector<int> deps;
    for (int j = 0; j < i; ++j)
      if (br_overlap(W[j], R[i]) || br_overlap(W[j], W[i]) || br_overlap(R[j], W[i])) deps.push_back(j);
    // continue the chain of the latest dependency when it is still the tail of its stream, else take an idle stream
    int k = -1;
    for (int q = (int)deps.size() - 1; q >= 0 && k < 0; --q)
      if (tail[stream_of[deps[q]]] == deps[q]) k = stream_of[deps[q]];
    if (k < 0) {
      for (int q = 0; q < kSide && k < 0; ++q)
        if (tail[(rr + q) % kSide] < 0) k = (rr + q) % kSide;
      if (k < 0) k = rr % kSide;
      ++rr;
    }
    cudaStream_t st = g->side[k];
    if (!joined[k]) {
      err = cudaStreamWaitEvent(st, ev[n], 0);
      joined[k] = 1;
    }
    for (int j : deps)
      if (err == cudaSuccess && !(stream_of[j] == k)) err = cudaStreamWaitEvent(st, ev[j], 0);
    if (err != cudaSuccess) break;
    const Step& s = steps[i];
    if (s.kind == 1) {
      rc = img_from_taps(reinterpret_cast<const float*>(g->ws + g->ws_y),
                         reinterpret_cast<const float*>(blobp(g, g->t_imgb)), out, B, g->cfg.crop, st);
    } else if (s.final_image && out != reinterpret_cast<float*>(g->ws + g->ws_out)) {
      ConvPlan p = s.plan;
      p.kp.e.out = out;
      p.desc.out = out;
      rc = launch_conv_plan(p, CHB_IMPL_TCGEN05, st);
    } else {
      rc = launch_conv_plan(s.plan, CHB_IMPL_TCGEN05, st);
    }
    if (rc == CHB_OK) err = cudaEventRecord(ev[i], st);
    stream_of[i] = k;
    tail[k] = i;
  }
  for (int k = 0; k < kSide && err == cudaSuccess; ++k)   // join every side stream back into the capturing stream
    if (tail[k] >= 0) err = cudaStreamWaitEvent(cap, ev[tail[k]], 0);
  for (auto& e : ev)
    if (e) cudaEventDestroy(e);
  if (rc != CHB_OK) return rc;
  if (err != cudaSuccess) {
    set_error(std::string("forward_graph: dependency capture failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

// One cudaGraphLaunch per forward: at B = 1 the ~64 launches of the schedule are a few microseconds of work each, so
// launch gaps, not kernels, set the latency of the reference's only real call pattern (gen_img, one image at a time).
int chb_generator_forward_graph(chb_generator* g, const uint8_t* labels, const float* codes, uint64_t seed, float* out,
                                int B, void* stream_) {
  if (!g || !labels || !codes || !out || !g->ws || !g->blob) {
    set_error("chb_generator_forward_graph: NULL argument or unbound generator");
    return CHB_ERR_ARG;
  }
  if (B <= 0 || B > g->cfg.max_batch) {
    set_error("chb_generator_forward_graph: batch exceeds max_batch");
    return CHB_ERR_ARG;
  }
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const chb_gen_config& c = g->cfg;
  uint8_t* ws = g->ws;
  const size_t S2 = (size_t)c.crop * c.crop;
  uint8_t* d_labels = ws + g->ws_labels;
  float* d_codes = reinterpret_cast<float*>(ws + g->ws_codes32);
  float* d_out = reinterpret_cast<float*>(ws + g->ws_out);
  cudaError_t err = cudaSuccess;
  if (g->d2h_stream && g->async_calls > 0) {
    // the streamed host entry point shares these staging buffers: let its copies drain first
    err = cudaStreamSynchronize(g->h2d_stream);
    if (err == cudaSuccess) err = cudaStreamSynchronize(g->d2h_stream);
  }
  auto it = g->graphs.find(B);
  if (it == g->graphs.end()) {
    std::vector<Step>* steps = nullptr;
    int rc = get_steps(g, B, &steps);  // plans (tensor maps) are built outside the capture
    if (rc != CHB_OK) return rc;
    rc = ensure_conv_kernels_ready();
    if (rc != CHB_OK) return rc;
    chb_generator::GraphEntry* ge = new chb_generator::GraphEntry();
    cudaStream_t cap = nullptr;
    cudaGraph_t graph = nullptr;
    err = cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
    if (err == cudaSuccess) {
      rc = forward_impl(g, d_labels, d_codes, nullptr, seed, d_out, B, CHB_IMPL_TCGEN05, cap, nullptr, 0, /*dag=*/true);
      err = cudaStreamEndCapture(cap, &graph);
      if (rc != CHB_OK) {
        if (graph) cudaGraphDestroy(graph);
        cudaStreamDestroy(cap);
        delete ge;
        return rc;
      }
    }
    if (err == cudaSuccess) {
      size_t n = 0;
      err = cudaGraphGetNodes(graph, nullptr, &n);
      std::vector<cudaGraphNode_t> nodes(n);
      if (err == cudaSuccess) err = cudaGraphGetNodes(graph, nodes.data(), &n);
      for (size_t i = 0; i < n && err == cudaSuccess; ++i) {
        cudaGraphNodeType ty;
        if (cudaGraphNodeGetType(nodes[i], &ty) != cudaSuccess || ty != cudaGraphNodeTypeKernel) continue;
        cudaKernelNodeParams kp;
        if (cudaGraphKernelNodeGetParams(nodes[i], &kp) != cudaSuccess) continue;
        if (kp.func == chb_noise_fill_kernel_address()) {
          ge->noise_node = nodes[i];
          ge->noise_params = kp;
        }
      }
      if (err == cudaSuccess && !ge->noise_node) {
        set_error("forward_graph: noise node not found in the captured graph");
        cudaGraphDestroy(graph);
        cudaStreamDestroy(cap);
        delete ge;
        return CHB_ERR_CUDA;
      }
    }
    if (err == cudaSuccess) err = cudaGraphInstantiate(&ge->exec, graph, 0);
    if (cap) cudaStreamDestroy(cap);
    ge->graph = graph;
    if (err != cudaSuccess) {
      set_error(std::string("forward_graph: capture / instantiate failed: ") + cudaGetErrorString(err));
      if (graph) cudaGraphDestroy(graph);
      delete ge;
      cudaGetLastError();
      return CHB_ERR_CUDA;
    }
    ge->nz = reinterpret_cast<float*>(ws + g->ws_noise);
    ge->nz_n = (long long)B * g->noise_pix;
    ge->offset = 0;
    ge->args[0] = &ge->nz; ge->args[1] = &ge->nz_n; ge->args[2] = &ge->seed; ge->args[3] = &ge->offset;
    ge->noise_params.kernelParams = ge->args;
    ge->noise_params.extra = nullptr;
    it = g->graphs.emplace(B, ge).first;
  }
  chb_generator::GraphEntry* ge = it->second;
  if (err == cudaSuccess && labels != d_labels)
    err = cudaMemcpyAsync(d_labels, labels, (size_t)B * S2, cudaMemcpyDeviceToDevice, stream);
  if (err == cudaSuccess && codes != d_codes)
    err = cudaMemcpyAsync(d_codes, codes, (size_t)B * c.label_nc * c.style_len * 4, cudaMemcpyDeviceToDevice, stream);
  if (err == cudaSuccess) {
    ge->seed = seed;
    err = cudaGraphExecKernelNodeSetParams(ge->exec, ge->noise_node, &ge->noise_params);
  }
  if (err == cudaSuccess) err = cudaGraphLaunch(ge->exec, stream);
  if (err == cudaSuccess && out != d_out)
    err = cudaMemcpyAsync(out, d_out, (size_t)B * 3 * S2 * 4, cudaMemcpyDeviceToDevice, stream);
  if (err != cudaSuccess) {
    set_error(std::string("forward_graph: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

int chb_generator_forward_host(chb_generator* g, const uint8_t* labels_host, const float* codes_host,
                               const float* noise_host, uint64_t seed, float* out_host, int B, int impl,
                               void* stream_) {
  if (!g || !labels_host || !codes_host || !out_host || !g->ws) {
    set_error("chb_generator_forward_host: NULL argument or unbound generator");
    return CHB_ERR_ARG;
  }
  if (B <= 0 || B > g->cfg.max_batch) {
    set_error("chb_generator_forward_host: batch exceeds max_batch");
    return CHB_ERR_ARG;
  }
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const chb_gen_config& c = g->cfg;
  uint8_t* ws = g->ws;
  const size_t S2 = (size_t)c.crop * c.crop;
  cudaError_t err = cudaMemcpyAsync(ws + g->ws_labels, labels_host, (size_t)B * S2, cudaMemcpyHostToDevice, stream);
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(ws + g->ws_codes32, codes_host, (size_t)B * c.label_nc * c.style_len * 4,
                          cudaMemcpyHostToDevice, stream);
  if (err == cudaSuccess && noise_host)
    err = cudaMemcpyAsync(ws + g->ws_noise, noise_host, (size_t)B * g->noise_pix * 4, cudaMemcpyHostToDevice, stream);
  if (err != cudaSuccess) {
    set_error(std::string("H2D copy failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  float* dev_out = reinterpret_cast<float*>(ws + g->ws_out);
  int rc = chb_generator_forward(g, ws + g->ws_labels, reinterpret_cast<const float*>(ws + g->ws_codes32),
                                 noise_host ? reinterpret_cast<const float*>(ws + g->ws_noise) : nullptr, seed,
                                 dev_out, B, impl, stream_);
  if (rc != CHB_OK) return rc;
  err = cudaMemcpyAsync(out_host, dev_out, (size_t)B * 3 * S2 * 4, cudaMemcpyDeviceToHost, stream);
  if (err == cudaSuccess) err = cudaStreamSynchronize(stream);
  if (err != cudaSuccess) {
    set_error(std::string("D2H copy / sync failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

// Streamed form of forward_host: batch n's H2D (own stream) overlaps batch n-1's compute, its D2H (own stream) overlaps
// batch n+1's compute; inputs and the output image are double buffered in the workspace.
int chb_generator_forward_host_async(chb_generator* g, const uint8_t* labels_host, const float* codes_host,
                                     uint64_t seed, float* out_host, int B, void* stream_) {
  if (!g || !labels_host || !codes_host || !out_host || !g->ws) {
    set_error("chb_generator_forward_host_async: NULL argument or unbound generator");
    return CHB_ERR_ARG;
  }
  if (B <= 0 || B > g->cfg.max_batch) {
    set_error("chb_generator_forward_host_async: batch exceeds max_batch");
    return CHB_ERR_ARG;
  }
  cudaError_t err = cudaSuccess;
  if (!g->h2d_stream) {
    err = cudaStreamCreateWithFlags(&g->h2d_stream, cudaStreamNonBlocking);
    if (err == cudaSuccess) err = cudaStreamCreateWithFlags(&g->d2h_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 2 && err == cudaSuccess; ++i) {
      err = cudaEventCreateWithFlags(&g->ev_in[i], cudaEventDisableTiming);
      if (err == cudaSuccess) err = cudaEventCreateWithFlags(&g->ev_done[i], cudaEventDisableTiming);
      if (err == cudaSuccess) err = cudaEventCreateWithFlags(&g->ev_copied[i], cudaEventDisableTiming);
    }
    if (err != cudaSuccess) {
      set_error(std::string("forward_host_async: stream/event creation failed: ") + cudaGetErrorString(err));
      return CHB_ERR_CUDA;
    }
  }
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const chb_gen_config& c = g->cfg;
  uint8_t* ws = g->ws;
  const size_t S2 = (size_t)c.crop * c.crop;
  const int p = (int)(g->async_calls & 1);
  const bool reuse = g->async_calls >= 2;  // buffer set p was last used by call n-2
  uint8_t* d_labels = ws + (p ? g->ws_labels_b : g->ws_labels);
  float* d_codes = reinterpret_cast<float*>(ws + (p ? g->ws_codes32_b : g->ws_codes32));
  float* d_out = reinterpret_cast<float*>(ws + (p ? g->ws_out_b : g->ws_out));
  // H2D of this batch: may start as soon as call n-2 (same buffers) has finished computing
  if (reuse) err = cudaStreamWaitEvent(g->h2d_stream, g->ev_done[p], 0);
  if (err == cudaSuccess) err = cudaMemcpyAsync(d_labels, labels_host, (size_t)B * S2, cudaMemcpyHostToDevice, g->h2d_stream);
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(d_codes, codes_host, (size_t)B * c.label_nc * c.style_len * 4, cudaMemcpyHostToDevice,
                          g->h2d_stream);
  if (err == cudaSuccess) err = cudaEventRecord(g->ev_in[p], g->h2d_stream);
  // compute: needs the inputs, and the D2H of call n-2 must have drained the output buffer
  if (err == cudaSuccess) err = cudaStreamWaitEvent(stream, g->ev_in[p], 0);
  if (err == cudaSuccess && reuse) err = cudaStreamWaitEvent(stream, g->ev_copied[p], 0);
  if (err != cudaSuccess) {
    set_error(std::string("forward_host_async: H2D stage failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  int rc = chb_generator_forward(g, d_labels, d_codes, nullptr, seed, d_out, B, CHB_IMPL_TCGEN05, stream_);
  if (rc != CHB_OK) return rc;
  err = cudaEventRecord(g->ev_done[p], stream);
  if (err == cudaSuccess) err = cudaStreamWaitEvent(g->d2h_stream, g->ev_done[p], 0);
  if (err == cudaSuccess)
    err = cudaMemcpyAsync(out_host, d_out, (size_t)B * 3 * S2 * 4, cudaMemcpyDeviceToHost, g->d2h_stream);
  if (err == cudaSuccess) err = cudaEventRecord(g->ev_copied[p], g->d2h_stream);
  if (err != cudaSuccess) {
    set_error(std::string("forward_host_async: D2H stage failed: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  g->async_calls++;
  return CHB_OK;
}

// Waits until every image enqueued by chb_generator_forward_host_async has landed in its host buffer.
int chb_generator_host_sync(chb_generator* g) {
  if (!g) {
    set_error("chb_generator_host_sync: NULL generator");
    return CHB_ERR_ARG;
  }
  if (!g->d2h_stream) return CHB_OK;
  cudaError_t err = cudaStreamSynchronize(g->d2h_stream);
  if (err != cudaSuccess) {
    set_error(std::string("chb_generator_host_sync: ") + cudaGetErrorString(err));
    return CHB_ERR_CUDA;
  }
  return CHB_OK;
}

int chb_generator_launches(const chb_generator* g) {
  if (!g) return 0;
  int n = 3;  // codes cast, one-hot pyramid, noise
  if (g->n_styled > 0) {
    n += 1;  // fc_mu
    int prevC = -1;  // one weff launch per run of equal-width styled ACEs
    for (auto& b : g->blocks)
      for (int a = 0; a < 3; ++a)
        if (b.ace[a].C && b.ace[a].styled) {
          if (b.ace[a].C != prevC) ++n;
          prevC = b.ace[a].C;
        }
  }
  n += 1;  // fc
  for (auto& b : g->blocks) n += 1 + b.n_ace + 2;
  return n + ((g->cfg.precision & CHB_PREC_IMG) ? 2 : 1);  // conv_img (taps GEMM + gather, or one conv)
}

double chb_generator_flops(const chb_generator* g_, int B) {
  chb_generator* g = const_cast<chb_generator*>(g_);
  if (!g || !g->ws) return 0.0;
  std::vector<Step>* steps = nullptr;
  if (get_steps(g, B, &steps) != CHB_OK) return 0.0;
  double f = 0;
  for (const Step& s : *steps) f += s.plan.flops;
  return f;
}

int chb_generator_step_name(chb_generator* g, int B, int i, char* name, int cap) {
  if (!g || !name || cap <= 0 || !g->ws) return CHB_ERR_ARG;
  std::vector<Step>* steps = nullptr;
  int rc = get_steps(g, B, &steps);
  if (rc != CHB_OK) return rc;
  if (i < 0 || i >= (int)steps->size()) return CHB_ERR_ARG;
  snprintf(name, cap, "%s", (*steps)[i].name.c_str());
  return CHB_OK;
}

int chb_generator_set_step_limit(chb_generator* g, int n) {
  if (!g) return CHB_ERR_ARG;
  g->step_limit = n;
  return CHB_OK;
}

int64_t chb_generator_debug_tensor(const chb_generator* g, const char* name, int B, void** dev_ptr, int* dtype) {
  if (!g || !name || !g->ws) return -1;
  auto it = g->debug.find(name);
  if (it == g->debug.end()) {
    set_error(std::string("unknown debug tensor ") + name);
    return -1;
  }
  if (dev_ptr) *dev_ptr = g->ws + it->second.first;
  if (dtype) *dtype = it->second.second;
  (void)B;
  return 0;
}

}  // extern "C"

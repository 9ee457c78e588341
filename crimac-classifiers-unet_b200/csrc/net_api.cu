// Network-level C-ABI: the whole UNet_Baseline forward (eval / train) and backward as a fixed sequence of kernel
// launches on the caller's stream.  Mirrors reference models/unet.py:200-343 (topology), pipeline.py:135-138,176-177
// (loss, backward).  The context owns only sub-allocations of the caller's workspace plus host-side launch
// descriptors (TMA tensor maps are encoded once here, not per call).
//
// Data layout in HBM (all activations NHWC bf16, [patch][range row][ping][channel]):
//   level l has C_l = 64*2^l channels at (H>>l) x (W>>l).
//   Cat_j  (decoder block j, level l = depth-2-j): ONE buffer of 2*C_l channels; ConvTranspose2d scatters into
//          channels [0,C_l), the encoder's second conv of level l writes its activation into [C_l,2*C_l) — torch.cat
//          (unet.py:132) never happens as a copy.
//   train mode additionally keeps every conv's raw (pre-BN) output for the BN/ReLU backward.
//
// Streams: everything runs on the caller's stream except the weight-gradient GEMMs and the ConvTranspose bias sums of
// crimac_backward, which go to a context-owned side stream (fork/join with events recorded inside the same call, so a
// whole train step can be captured in a CUDA graph: trainer.py does).  While crimac_profile_enable(1) is on, the side
// stream is not used and every launch is bracketed by events on the caller's stream.
#include "host_util.h"
#include "../../include/crimac_b200.h"
#include <vector>
#include <new>
#include <algorithm>

namespace {

struct Bump {
  uint8_t* base;
  size_t off = 0;
  explicit Bump(void* b) : base(static_cast<uint8_t*>(b)) {}
  void* take(size_t bytes) {
    off = (off + 1023) & ~static_cast<size_t>(1023);
    void* p = base ? base + off : nullptr;
    off += bytes;
    return p;
  }
  template <typename T>
  T* arr(size_t n) {
    return static_cast<T*>(take(n * sizeof(T)));
  }
};

struct Conv3 {
  int cin = 0, cout = 0, level = 0;
  bool first = false;
  int s_w = 0, s_b = 0, s_g = 0, s_beta = 0, s_rm = 0, s_rv = 0, s_nbt = 0;
  int g_w = 0, g_b = 0, g_g = 0, g_beta = 0;
  View in{}, raw{}, act{}, pool{}, gin{};
  bf16* gin2 = nullptr;  // decoder first convs ("concat"): input-gradient channels >= gin_split go to this dense tensor
  int gin_split = 0;
  uint16_t* pool_arg = nullptr;  // train: 2-bit arg-max map of the fused max-pool (one uint16 per 8 channels)
  bf16* w_fwd = nullptr;  // backward-data reads the same matrix as an MN-major operand
  float *scale = nullptr, *shift = nullptr, *mean = nullptr, *invstd = nullptr;
  int bn_fwd = 0, bn_bwd = 0, bn_wg = 0;
  ConvParams fwd{}, dgrad{};
  WgradHaloParams wg{};   // all-taps halo kernel (wide, shallow layers)
  WgradParams wg_tap{};   // one-tap-per-CTA kernel (deep layers: few pixels, many channels)
  bool wg_use_halo = true;
  float* wg_scratch = nullptr;  // [9][cout][cin] fp32, zero between steps
  float* wg_slabs = nullptr;    // deterministic mode: wg_max_splits copies of the scratch, one per split-K slice
  int wg_max_splits = 0, wg_active = 0;  // allocated / written by the last launch
  int gr_idx = 0;               // which of the two dRaw buffers this layer's BatchNorm backward writes
};
struct ConvT {
  int cin = 0, cout = 0, level_in = 0;
  int s_w = 0, s_b = 0, g_w = 0, g_b = 0;
  View in{}, out{}, gout{}, gin{};
  View tmp{}, dtmp{};     // up_mode "upsample": 1x1-conv output at the LOW resolution (before the bilinear 2x) and its gradient
  bf16* w_fwd = nullptr;  // backward-data reads the same matrix as an MN-major operand
  int bn_fwd = 0, bn_bwd = 0, bn_wg = 0;
  ConvParams fwd{}, dgrad{};
  WgradParams wg{};
  float* wg_scratch = nullptr;  // [4][cin][cout] fp32, zero between steps
  float* wg_slabs = nullptr;
  int wg_max_splits = 0, wg_active = 0;
};

int pick_bn(int n) { return n % 256 == 0 ? 256 : (n % 128 == 0 ? 128 : 64); }

}  // namespace

struct crimac_ctx {
  crimac_config cfg{};
  int device = 0;
  int D = 0;
  std::vector<Conv3> conv;  // enc c1,c2 per level, then dec c1,c2 per block (forward order is built separately)
  std::vector<ConvT> up;
  std::vector<int> enc1, enc2, dec1, dec2;  // indices into conv
  int s_head_w = 0, s_head_b = 0, g_head_w = 0, g_head_b = 0;
  // shared scratch
  bf16 *GA = nullptr, *GP = nullptr;
  bf16* GR[2] = {nullptr, nullptr};  // dRaw, double-buffered: layer k's weight gradient overlaps layer k+1's BN backward
  cudaStream_t side = nullptr;       // weight-gradient GEMMs run here, concurrently with the HBM-bound backward kernels
  cudaEvent_t ev_draw[2] = {nullptr, nullptr}, ev_wg[2] = {nullptr, nullptr}, ev_cat = nullptr, ev_join = nullptr;
  bool overlap = true;
  std::vector<bf16*> dcat;   // gradient of decoder block j's merged input: "concat": the up-sampled half (dense, C channels);
  std::vector<bf16*> dskip;  //   the skip half lives in dskip[j] (dense, C channels).  "add": dcat[j] is both
  float* stats = nullptr;
  float* red_partials = nullptr;
  float* colsum_partials = nullptr;  // ConvTranspose bias-gradient partials (side stream: must not share red_partials)
  float* c1c2 = nullptr;
  float* wg_arena = nullptr;   // all layers' weight-gradient scratch, contiguous
  size_t wg_arena_bytes = 0;
  bool wg_dirty = true;        // scratch may hold partial sums (first use, or a failed backward)
  float* head_partials = nullptr;
  float* fc_partials = nullptr;
  bf16* xs = nullptr;          // first conv: the input echogram as bf16 hi/lo pairs, NHWC (first_conv_tc.cu)
  double* ce_partials = nullptr;
  float* logits = nullptr;   // internal logits / dlogits for the fused train step
  float* dlogits = nullptr;
  float* loss3 = nullptr;
  ConvParams head_fused{};   // last conv with the 1x1 head in its epilogue (eval)
  int prepared_mode = -1;
  int staged_nb = 0;         // patches crimac_preprocess_staged left in xs (consumed by the next forward with x == NULL)
  // data-parallel gradient exchange over peer memory (crimac_set_comm): buckets are reduced on `comm` while backward runs
  crimac_comm_config comm{};
  bool comm_on = false;
  crimac_optimizer_config opt{};   // fused SGD(momentum) per gradient bucket (crimac_set_optimizer)
  bool opt_on = false;
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_bk_main[3] = {nullptr, nullptr, nullptr}, ev_bk_side[3] = {nullptr, nullptr, nullptr}, ev_comm = nullptr;
};

namespace {

int level_h(const crimac_ctx* c, int l) { return c->cfg.height >> l; }
int level_w(const crimac_ctx* c, int l) { return c->cfg.width >> l; }

int validate(const crimac_config* cfg) {
  CRIMAC_REQUIRE(cfg != nullptr, "cfg is NULL");
  CRIMAC_REQUIRE(cfg->in_channels >= 1 && cfg->in_channels <= 12, "in_channels must be 1..12 (frequencies + metadata channels)");
  CRIMAC_REQUIRE(cfg->n_classes >= 1 && cfg->n_classes <= CRIMAC_MAX_CLASSES, "n_classes must be 1..8");
  CRIMAC_REQUIRE(cfg->depth >= 2 && cfg->depth <= 5, "depth must be 2..5");
  CRIMAC_REQUIRE(cfg->start_filts == 64, "start_filts must be 64 (tensor-core tiles are 64 channels wide)");
  CRIMAC_REQUIRE(cfg->max_batch >= 1, "max_batch");
  CRIMAC_REQUIRE((cfg->up_mode == 0 || cfg->up_mode == 1) && (cfg->merge_mode == 0 || cfg->merge_mode == 1), "up_mode / merge_mode must be 0 or 1");
  CRIMAC_REQUIRE(!(cfg->up_mode == 1 && cfg->merge_mode == 1), "up_mode upsample is incompatible with merge_mode add (as in the reference, unet.py:243-251)");
  const int m = 1 << (cfg->depth - 1);
  CRIMAC_REQUIRE(cfg->height > 0 && cfg->width > 0 && cfg->height % m == 0 && cfg->width % m == 0,
                 "height/width must be multiples of 2^(depth-1) (the reference has the same constraint, unet.py:132)");
  return 0;
}

// Builds the layer graph.  With ws == nullptr only sizes are computed (crimac_workspace_bytes).
int build(crimac_ctx* c, void* ws, size_t* bytes_out, bool encode_maps) {
  const crimac_config& cfg = c->cfg;
  const int D = c->D = cfg.depth;
  const int B = cfg.max_batch;
  const bool train = cfg.train != 0;
  const bool add = cfg.merge_mode == 1;    // merge_mode "add": from_up + from_down instead of torch.cat (unet.py:131-134)
  const bool ups = cfg.up_mode == 1;       // up_mode "upsample": bilinear 2x + conv1x1 instead of ConvTranspose2d (unet.py:50-56)
  Bump bump(ws);
  auto chan = [&](int l) { return cfg.start_filts << l; };
  auto dense = [&](int l, int C) {
    View v{bump.arr<bf16>(static_cast<size_t>(B) * level_h(c, l) * level_w(c, l) * C), B, level_h(c, l), level_w(c, l), C,
           C};
    return v;
  };
  auto slice = [&](const View& v, int c0, int C) {
    View s = v;
    s.ptr = v.ptr ? v.ptr + c0 : nullptr;
    s.C = C;
    return s;
  };

  // ---- activations
  std::vector<View> cat(D - 1), pooled(D - 1);
  for (int j = 0; j < D - 1; ++j) cat[j] = dense(D - 2 - j, (add ? 1 : 2) * chan(D - 2 - j));
  for (int l = 0; l < D - 1; ++l) pooled[l] = dense(l + 1, chan(l));

  c->conv.clear();
  c->up.clear();
  c->enc1.clear();
  c->enc2.clear();
  c->dec1.clear();
  c->dec2.clear();
  int s_idx = 0, g_idx = 0;
  auto new_conv = [&](int cin, int cout, int level) {
    Conv3 L;
    L.cin = cin;
    L.cout = cout;
    L.level = level;
    c->conv.push_back(L);
    return static_cast<int>(c->conv.size()) - 1;
  };
  // state/grad indices follow state_dict()/parameters() order
  for (int l = 0; l < D; ++l) {
    const int cin = l == 0 ? cfg.in_channels : chan(l - 1), co = chan(l);
    const int i1 = new_conv(cin, co, l);
    const int i2 = new_conv(co, co, l);
    c->enc1.push_back(i1);
    c->enc2.push_back(i2);
    for (int which = 0; which < 2; ++which) {
      Conv3& L = c->conv[which == 0 ? i1 : i2];
      L.s_w = s_idx++;
      L.s_b = s_idx++;
      L.s_g = s_idx++;
      L.s_beta = s_idx++;
      L.s_rm = s_idx++;
      L.s_rv = s_idx++;
      L.s_nbt = s_idx++;
      L.g_w = g_idx++;
      L.g_b = g_idx++;
      L.g_g = g_idx++;
      L.g_beta = g_idx++;
    }
    c->conv[i1].first = (l == 0);
  }
  for (int j = 0; j < D - 1; ++j) {
    const int l = D - 2 - j, co = chan(l);
    ConvT U;
    U.cin = 2 * co;
    U.cout = co;
    U.level_in = l + 1;
    U.s_w = s_idx++;
    U.s_b = s_idx++;
    U.g_w = g_idx++;
    U.g_b = g_idx++;
    c->up.push_back(U);
    const int i1 = new_conv(add ? co : 2 * co, co, l);
    const int i2 = new_conv(co, co, l);
    c->dec1.push_back(i1);
    c->dec2.push_back(i2);
    Conv3& L1 = c->conv[i1];
    Conv3& L2 = c->conv[i2];
    L1.s_w = s_idx++;
    L1.s_b = s_idx++;
    L2.s_w = s_idx++;
    L2.s_b = s_idx++;
    for (Conv3* L : {&L1, &L2}) {
      L->s_g = s_idx++;
      L->s_beta = s_idx++;
      L->s_rm = s_idx++;
      L->s_rv = s_idx++;
      L->s_nbt = s_idx++;
    }
    L1.g_w = g_idx++;
    L1.g_b = g_idx++;
    L2.g_w = g_idx++;
    L2.g_b = g_idx++;
    L1.g_g = g_idx++;
    L1.g_beta = g_idx++;
    L2.g_g = g_idx++;
    L2.g_beta = g_idx++;
  }
  c->s_head_w = s_idx++;
  c->s_head_b = s_idx++;
  c->g_head_w = g_idx++;
  c->g_head_b = g_idx++;

  // ---- wire views
  for (int l = 0; l < D; ++l) {
    Conv3& L1 = c->conv[c->enc1[l]];
    Conv3& L2 = c->conv[c->enc2[l]];
    if (l > 0) L1.in = pooled[l - 1];
    L1.act = dense(l, chan(l));
    L2.in = L1.act;
    if (l < D - 1) {
      L2.act = slice(cat[D - 2 - l], add ? 0 : chan(l), chan(l));   // "add": the up-sampled tensor is added in place later
      L2.pool = pooled[l];
    } else {
      L2.act = dense(l, chan(l));
    }
  }
  for (int j = 0; j < D - 1; ++j) {
    const int l = D - 2 - j;
    ConvT& U = c->up[j];
    U.in = (j == 0) ? c->conv[c->enc2[D - 1]].act : c->conv[c->dec2[j - 1]].act;
    U.out = slice(cat[j], 0, chan(l));
    if (ups) U.tmp = dense(l + 1, chan(l));
    Conv3& L1 = c->conv[c->dec1[j]];
    Conv3& L2 = c->conv[c->dec2[j]];
    L1.in = cat[j];
    L1.act = dense(l, chan(l));
    L2.in = L1.act;
    L2.act = dense(l, chan(l));
  }
  // ---- per-layer parameters-derived buffers
  size_t max_stats = 0;
  for (Conv3& L : c->conv) {
    if (!L.first) {
      L.w_fwd = bump.arr<bf16>(static_cast<size_t>(L.cout) * 9 * L.cin);
    }
    L.scale = bump.arr<float>(L.cout);
    L.shift = bump.arr<float>(L.cout);
    L.mean = bump.arr<float>(L.cout);
    L.invstd = bump.arr<float>(L.cout);
    if (train) L.raw = dense(L.level, L.cout);
    if (train && L.pool.C > 0)  // encoder second convs below the deepest level (same test in the sizing pass)
      L.pool_arg = bump.arr<uint16_t>(static_cast<size_t>(B) * (level_h(c, L.level) / 2) * (level_w(c, L.level) / 2) * (L.cout / 8));
    const size_t rows = L.first ? 148 * 8 : 256;  // >= first_conv_grid() / >= number of SMs (one partial row per CTA)
    if (rows * 2 * L.cout > max_stats) max_stats = rows * 2 * L.cout;
    L.bn_fwd = pick_bn(L.cout);
    L.bn_bwd = pick_bn(L.cin);
    L.bn_wg = pick_bn(L.cin);
  }
  for (ConvT& U : c->up) {
    U.w_fwd = bump.arr<bf16>(static_cast<size_t>(U.cin) * U.cout * (ups ? 1 : 4));
    U.bn_fwd = pick_bn(ups ? U.cout : 4 * U.cout);
    U.bn_bwd = pick_bn(U.cin);
    U.bn_wg = pick_bn(U.cout);
  }
  c->xs = bump.arr<bf16>(first_conv_split_elems(B, cfg.in_channels, cfg.height, cfg.width));
  // ---- training scratch
  c->dcat.assign(D - 1, nullptr);
  c->dskip.assign(D - 1, nullptr);
  if (train) {
    const size_t lvl0 = static_cast<size_t>(B) * cfg.height * cfg.width * cfg.start_filts;
    c->GA = bump.arr<bf16>(lvl0);
    c->GR[0] = bump.arr<bf16>(lvl0);
    c->GR[1] = bump.arr<bf16>(lvl0);
    {
      // backward visits the convs in this order; consecutive layers alternate between the two dRaw buffers
      int pos = 0;
      for (int j = D - 2; j >= 0; --j) {
        c->conv[c->dec2[j]].gr_idx = (pos++) & 1;
        c->conv[c->dec1[j]].gr_idx = (pos++) & 1;
      }
      for (int l = D - 1; l >= 0; --l) {
        c->conv[c->enc2[l]].gr_idx = (pos++) & 1;
        c->conv[c->enc1[l]].gr_idx = (pos++) & 1;
      }
    }
    c->GP = bump.arr<bf16>(lvl0 / 4);
    for (int j = 0; j < D - 1; ++j) {
      const int l = D - 2 - j;
      // two DENSE halves instead of one concat-shaped gradient: the ConvTranspose backward reads only the first, the
      // encoder's skip branch only the second - strided half-reads of a 2C-wide buffer cost DRAM row bandwidth
      c->dcat[j] = bump.arr<bf16>(static_cast<size_t>(B) * level_h(c, l) * level_w(c, l) * chan(l));
      c->dskip[j] = add ? c->dcat[j] : bump.arr<bf16>(static_cast<size_t>(B) * level_h(c, l) * level_w(c, l) * chan(l));
      if (ups) c->up[j].dtmp = dense(l + 1, chan(l));
    }
    const int cmax = chan(D - 1);
    c->stats = bump.arr<float>(max_stats);
    c->red_partials = bump.arr<float>(static_cast<size_t>(reduce_blocks()) * 2 * cmax);
    c->colsum_partials = bump.arr<float>(static_cast<size_t>(reduce_blocks()) * cmax);
    c->c1c2 = bump.arr<float>(2 * cmax);
    {
      // one contiguous arena, one region per layer: the unpack kernel re-zeroes behind itself, so no per-layer memset
      size_t total = 0;
      for (Conv3& L : c->conv)
        if (!L.first) total += static_cast<size_t>(9) * L.cout * L.cin;
      for (ConvT& U : c->up) total += static_cast<size_t>(ups ? 1 : 4) * U.cin * U.cout;
      c->wg_arena = bump.arr<float>(total);
      c->wg_arena_bytes = total * sizeof(float);
      float* q = c->wg_arena;
      for (Conv3& L : c->conv)
        if (!L.first) {
          L.wg_scratch = q;
          if (q) q += static_cast<size_t>(9) * L.cout * L.cin;
        }
      for (ConvT& U : c->up) {
        U.wg_scratch = q;
        if (q) q += static_cast<size_t>(ups ? 1 : 4) * U.cin * U.cout;
      }
    }
    if (cfg.deterministic) {
      // fixed-order split-K: every split stores its partial tile into its own slab (<= 2 waves of 148 CTAs per layer)
      for (Conv3& L : c->conv)
        if (!L.first) {
          const bool halo = static_cast<long>(L.cout) * L.cin <= 256L * 128L;
          const int tiles = halo ? (L.cout / 64) * (L.cin / 64) : 9 * ((L.cout + 127) / 128) * (L.cin / pick_bn(L.cin));
          L.wg_max_splits = std::max(1, (2 * 148) / tiles);
          L.wg_slabs = bump.arr<float>(static_cast<size_t>(L.wg_max_splits) * 9 * L.cout * L.cin);
        }
      for (ConvT& U : c->up) {
        const int tiles = ups ? ((U.cout + 127) / 128) * (U.cin / pick_bn(U.cin))
                              : 4 * ((U.cin + 127) / 128) * (U.cout / pick_bn(U.cout));
        U.wg_max_splits = std::max(1, (2 * 148) / tiles);
        U.wg_slabs = bump.arr<float>(static_cast<size_t>(U.wg_max_splits) * (ups ? 1 : 4) * U.cin * U.cout);
      }
    }
    c->head_partials = bump.arr<float>(static_cast<size_t>(head_bwd_blocks()) * 4 * (cfg.n_classes * 64 + cfg.n_classes));
    c->fc_partials = bump.arr<float>(first_conv_wgrad_partial_floats(cfg.in_channels));
    c->ce_partials = bump.arr<double>(static_cast<size_t>(ce_blocks()) * 2);
    const size_t lg = static_cast<size_t>(B) * cfg.n_classes * cfg.height * cfg.width;
    c->logits = bump.arr<float>(lg);
    c->dlogits = bump.arr<float>(lg);
    c->loss3 = bump.arr<float>(4);
    // gradient destinations of every dgrad
    for (int l = 0; l < D; ++l) {
      Conv3& L1 = c->conv[c->enc1[l]];
      Conv3& L2 = c->conv[c->enc2[l]];
      if (l > 0) L1.gin = View{c->GP, B, level_h(c, l), level_w(c, l), L1.cin, L1.cin};
      L2.gin = View{c->GA, B, level_h(c, l), level_w(c, l), L2.cin, L2.cin};
    }
    for (int j = 0; j < D - 1; ++j) {
      const int l = D - 2 - j;
      Conv3& L1 = c->conv[c->dec1[j]];
      Conv3& L2 = c->conv[c->dec2[j]];
      L1.gin = View{c->dcat[j], B, level_h(c, l), level_w(c, l), chan(l), chan(l)};
      if (!add) {   // channels [C, 2C) of the concat gradient = the skip half
        L1.gin2 = c->dskip[j];
        L1.gin_split = chan(l);
      }
      L2.gin = View{c->GA, B, level_h(c, l), level_w(c, l), L2.cin, L2.cin};
      ConvT& U = c->up[j];
      U.gout = View{c->dcat[j], B, level_h(c, l), level_w(c, l), U.cout, U.cout};
      U.gin = View{c->GA, B, level_h(c, l + 1), level_w(c, l + 1), U.cin, U.cin};
    }
  } else {
    c->stats = nullptr;
  }
  if (bytes_out) *bytes_out = bump.off + 1024;
  if (!encode_maps) return 0;

  // ---- launch descriptors (tensor maps encoded once)
  // every 3x3 conv (forward and backward-data) runs the halo main loop; the plain box-per-tap loop serves ConvTranspose
  auto geom = [&](ConvParams& p, int H, int W, int n_total, int bn) {
    p.H = H;
    p.W = W;
    p.halo = (p.taps == 9) ? 1 : 0;
    p.tiles_x = p.halo ? (W + 7) / 8 : (W + TILE_W - 1) / TILE_W;
    p.tiles_y = p.halo ? (H + 15) / 16 : (H + TILE_H - 1) / TILE_H;
    p.n_tiles = n_total / bn;
  };
  auto conv_map = [&](ConvParams& p, const View& v) {
    return p.halo ? make_act_map(&p.a_map[0], v, 18, 0, 0, 0, 10) : make_act_map(&p.a_map[0], v, TILE_H);
  };
  auto wgeom = [&](WgradParams& p, int H, int W, int m_total, int n_total, int bn, int taps, int tap_mode) {
    p.taps = taps;
    p.tap_mode = tap_mode;
    p.M_total = m_total;
    p.N_total = n_total;
    p.H = H;
    p.W = W;
    p.tiles_x = (W + 15) / 16;
    p.tiles_y = (H + 3) / 4;
    p.m_tiles = (m_total + 127) / 128;
    p.n_tiles = n_total / bn;
  };
  int rc;
  for (Conv3& L : c->conv) {
    const int H = level_h(c, L.level), W = level_w(c, L.level);
    if (!L.first) {
      ConvParams& p = L.fwd;
      p.taps = 9;
      p.tap_mode = 0;
      p.cin = L.cin;
      geom(p, H, W, L.cout, L.bn_fwd);
      if ((rc = conv_map(p, L.in))) return rc;
      if ((rc = make_weight_map(&p.b_map, L.w_fwd, L.cout, 9 * L.cin, L.bn_fwd))) return rc;
    }
    if (train) {
      View gr{c->GR[L.gr_idx], B, H, W, L.cout, L.cout};
      if (!L.first) {
        ConvParams& p = L.dgrad;
        p.taps = 9;
        p.tap_mode = 0;
        p.cin = L.cout;
        p.b_mn = 1;
        p.b_tap_cols = L.cin;
        geom(p, H, W, L.cin, L.bn_bwd);
        if ((rc = conv_map(p, gr))) return rc;
        if ((rc = make_weight_map(&p.b_map, L.w_fwd, L.cout, 9 * L.cin, 64))) return rc;
        p.out = L.gin.ptr;
        p.out_pitch = L.gin.pitch;
        p.out2 = L.gin2;
        p.out2_pitch = L.gin_split;
        p.out_split = L.gin_split;
        WgradHaloParams& w = L.wg;
        w.Cs = L.cout;
        w.Cf = L.cin;
        w.H = H;
        w.W = W;
        w.tiles_x = (W + 15) / 16;
        w.tiles_y = (H + 3) / 4;
        w.s_tiles = L.cout / 64;
        w.f_tiles = L.cin / 64;
        w.dw = L.wg_scratch;
        if ((rc = make_act_map(&w.s_map, gr, 6, 0, 0, 0, 18))) return rc;
        if ((rc = make_act_map(&w.f_map, L.in, 4))) return rc;
        // measured on B200 (profiles/): the 64x64-channel halo tiles win up to Cout*Cin = 256*128, beyond that the
        // 128 x 256 one-tap tiles re-read fewer operand bytes per FLOP
        L.wg_use_halo = static_cast<long>(L.cout) * L.cin <= 256L * 128L;
        // 128-wide X tiles where Cin allows: 2/3 of the shared-memory traffic per FLOP (the N = 64 tiles are port-bound)
        w.nf = (L.cin % 128 == 0) ? 128 : 64;
        WgradParams& wt = L.wg_tap;
        wgeom(wt, H, W, L.cout, L.cin, L.bn_wg, 9, 0);
        wt.dw = L.wg_scratch;
        if ((rc = make_act_map(&wt.a_map, gr, 4))) return rc;
        if ((rc = make_act_map(&wt.b_map[0], L.in, 4))) return rc;
      }
    }
  }
  for (size_t j = 0; j < c->up.size(); ++j) {
    ConvT& U = c->up[j];
    const int H = level_h(c, U.level_in), W = level_w(c, U.level_in);
    ConvParams& p = U.fwd;
    p.taps = 1;
    p.tap_mode = 0;
    p.cin = U.cin;
    if (ups) {
      // conv1x1 and bilinear up-sampling commute (both linear, the interpolation weights sum to 1): the 1x1 conv runs as
      // a one-tap GEMM at the LOW resolution (a quarter of the pixels), the bilinear 2x follows as an HBM-bound kernel
      geom(p, H, W, U.cout, U.bn_fwd);
      if ((rc = make_act_map(&p.a_map[0], U.in, TILE_H))) return rc;
      if ((rc = make_weight_map(&p.b_map, U.w_fwd, U.cout, U.cin, U.bn_fwd))) return rc;
      p.out = U.tmp.ptr;
      p.out_pitch = U.tmp.pitch;
      if (train) {
        ConvParams& d = U.dgrad;
        d.taps = 1;
        d.tap_mode = 0;
        d.cin = U.cout;
        geom(d, H, W, U.cin, U.bn_bwd);
        if ((rc = make_act_map(&d.a_map[0], U.dtmp, TILE_H))) return rc;
        d.b_mn = 1;
        if ((rc = make_weight_map(&d.b_map, U.w_fwd, U.cout, U.cin, 64))) return rc;
        d.out = U.gin.ptr;
        d.out_pitch = U.gin.pitch;
        WgradParams& w = U.wg;   // dW[co][ci] = sum_p dTmp[p][co] * X[p][ci]
        wgeom(w, H, W, U.cout, U.cin, pick_bn(U.cin), 1, 0);
        w.dw = U.wg_scratch;
        if ((rc = make_act_map(&w.a_map, U.dtmp, 4))) return rc;
        if ((rc = make_act_map(&w.b_map[0], U.in, 4))) return rc;
      }
      continue;
    }
    geom(p, H, W, 4 * U.cout, U.bn_fwd);
    if ((rc = make_act_map(&p.a_map[0], U.in, TILE_H))) return rc;
    if ((rc = make_weight_map(&p.b_map, U.w_fwd, 4 * U.cout, U.cin, U.bn_fwd))) return rc;
    p.out = U.out.ptr;
    p.out_pitch = U.out.pitch;
    p.convt_cout = U.cout;
    p.convt_add = add ? 1 : 0;
    if (train) {
      ConvParams& d = U.dgrad;
      d.taps = 4;
      d.tap_mode = 1;
      d.cin = U.cout;
      geom(d, H, W, U.cin, U.bn_bwd);
      for (int kk = 0; kk < 4; ++kk)
        if ((rc = make_act_map(&d.a_map[kk], U.gout, TILE_H, 1, kk >> 1, kk & 1))) return rc;
      d.b_mn = 1;
      if ((rc = make_weight_map(&d.b_map, U.w_fwd, 4 * U.cout, U.cin, 64))) return rc;
      d.out = U.gin.ptr;
      d.out_pitch = U.gin.pitch;
      WgradParams& w = U.wg;
      wgeom(w, H, W, U.cin, U.cout, U.bn_wg, 4, 1);
      w.dw = U.wg_scratch;
      if ((rc = make_act_map(&w.a_map, U.in, 4))) return rc;
      for (int kk = 0; kk < 4; ++kk)
        if ((rc = make_act_map(&w.b_map[kk], U.gout, 4, 1, kk >> 1, kk & 1))) return rc;
    }
  }
  return 0;
}

void set_batch(ConvParams& p, int nb) {
  p.NB = nb;
  p.total_tiles = nb * p.tiles_x * p.tiles_y * p.n_tiles;
}
void set_batch(WgradParams& p, int nb) {
  p.NB = nb;
  p.k_tiles_total = nb * p.tiles_x * p.tiles_y;
  const int tiles = p.taps * p.m_tiles * p.n_tiles;
  // one CTA per SM is resident: aim for (at most) two FULL waves - rounding the split factor UP would leave a third,
  // almost empty wave (e.g. 18 tiles x 17 splits = 306 CTAs on 148 SMs)
  int splits = (2 * device_num_sms()) / tiles;
  if (splits > p.k_tiles_total / 8) splits = p.k_tiles_total / 8;
  if (splits < 1) splits = 1;
  p.splits = splits;
}

void set_batch(WgradHaloParams& p, int nb) {
  p.NB = nb;
  p.k_tiles_total = nb * p.tiles_x * p.tiles_y;
  const int tiles = p.s_tiles * p.f_tiles;   // nf = 128: Cf/128 tiles x 2 CTA kinds = the same count
  int splits = (2 * device_num_sms()) / tiles;
  if (splits > p.k_tiles_total / 8) splits = p.k_tiles_total / 8;
  if (splits < 1) splits = 1;
  p.splits = splits;
}

template <typename T>
const T* S(const void* const* state, int i) {
  return static_cast<const T*>(state[i]);
}
template <typename T>
T* SM(const void* const* state, int i) {
  return static_cast<T*>(const_cast<void*>(state[i]));
}

int check_call(crimac_ctx* c, const void* const* state, int nb) {
  CRIMAC_REQUIRE(c != nullptr, "ctx is NULL");
  CRIMAC_REQUIRE(state != nullptr, "state table is NULL");
  CRIMAC_REQUIRE(nb >= 1 && nb <= c->cfg.max_batch, "nb must be in 1..max_batch");
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) {
    crimac_set_error(std::string("cudaSetDevice failed: ") + cudaGetErrorString(e));
    return 2;
  }
  return 0;
}

double igemm_flops_n(const ConvParams& p, int n_total) {
  return 2.0 * p.NB * static_cast<double>(p.H) * p.W * n_total * p.taps * p.cin;
}

// The weight-gradient GEMMs accumulate (red.add) into their layer's zeroed scratch; crimac_backward turns them into
// PyTorch-layout gradients with one unpack launch per gradient bucket (decoder, deep encoder, shallow encoder).
// active = number of split-K slices that have work (= slabs written in deterministic mode)
int active_splits(int k_tiles_total, int splits) {
  const int per = (k_tiles_total + splits - 1) / splits;
  return (k_tiles_total + per - 1) / per;
}

int wgrad_run(crimac_ctx* c, WgradParams& w, int bn, int nb, cudaStream_t st, float* slabs = nullptr, int max_splits = 0,
              int* active = nullptr) {
  set_batch(w, nb);
  w.slabs = slabs;
  if (slabs) {
    if (w.splits > max_splits) w.splits = max_splits;
    w.slab_stride = static_cast<long>(w.taps) * w.M_total * w.N_total;
    *active = active_splits(w.k_tiles_total, w.splits);
  }
  ProfScope ps("wgrad_gemm", 2.0 * nb * static_cast<double>(w.H) * w.W * w.M_total * w.N_total * w.taps, 0, st);
  CRIMAC_CHECK_CUDA(launch_wgrad_gemm(w, bn, st));
  return 0;
}

int wgrad_halo_run(crimac_ctx* c, WgradHaloParams& w, int nb, cudaStream_t st, float* slabs = nullptr, int max_splits = 0,
                   int* active = nullptr) {
  set_batch(w, nb);
  w.slabs = slabs;
  if (slabs) {
    if (w.splits > max_splits) w.splits = max_splits;
    w.slab_stride = static_cast<long>(9) * w.Cs * w.Cf;
    *active = active_splits(w.k_tiles_total, w.splits);
  }
  ProfScope ps("wgrad_gemm", 2.0 * nb * static_cast<double>(w.H) * w.W * w.Cs * w.Cf * 9, 0, st);
  CRIMAC_CHECK_CUDA(launch_wgrad_halo(w, st));
  return 0;
}

View with_batch(View v, int nb) {
  v.N = nb;
  return v;
}

}  // namespace

extern "C" int crimac_state_count(const crimac_config* cfg) {
  if (!cfg) return 0;
  return cfg->depth * 14 + (cfg->depth - 1) * 16 + 2;
}
extern "C" int crimac_grad_count(const crimac_config* cfg) {
  if (!cfg) return 0;
  return cfg->depth * 8 + (cfg->depth - 1) * 10 + 2;
}

extern "C" int crimac_workspace_bytes(const crimac_config* cfg, size_t* bytes) {
  int rc = validate(cfg);
  if (rc) return rc;
  CRIMAC_REQUIRE(bytes != nullptr, "bytes is NULL");
  crimac_ctx tmp;
  tmp.cfg = *cfg;
  return build(&tmp, nullptr, bytes, false);
}

extern "C" int crimac_create(crimac_ctx** out, const crimac_config* cfg, void* workspace_dev, size_t workspace_bytes,
                             int device) {
  int rc = validate(cfg);
  if (rc) return rc;
  CRIMAC_REQUIRE(out != nullptr && workspace_dev != nullptr, "NULL argument");
  CRIMAC_REQUIRE((reinterpret_cast<uintptr_t>(workspace_dev) & 1023) == 0, "workspace must be 1024-byte aligned");
  CRIMAC_CHECK_CUDA(cudaSetDevice(device));
  int major = 0, minor = 0;
  CRIMAC_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
  CRIMAC_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device));
  // the library carries ONE cubin, sm_100a: architecture-specific code does not load on sm_103 (B300) or any other 10.x part
  CRIMAC_REQUIRE(major == 10 && minor == 0, "libcrimac_b200 runs on sm_100a (B200, compute capability 10.0) only; there is no fallback path");
  crimac_ctx* c = new (std::nothrow) crimac_ctx;
  CRIMAC_REQUIRE(c != nullptr, "out of host memory");
  c->cfg = *cfg;
  c->device = device;
  size_t need = 0;
  rc = build(c, nullptr, &need, false);
  if (rc == 0 && need > workspace_bytes) {
    crimac_set_error("workspace too small: need " + std::to_string(need) + " bytes");
    rc = 1;
  }
  if (rc == 0) rc = build(c, workspace_dev, nullptr, true);
  if (rc == 0 && cfg->train) {
    c->overlap = true;  // switched off only while crimac_profile_enable(1) times every kernel on its own
    cudaError_t e = cudaStreamCreateWithFlags(&c->side, cudaStreamNonBlocking);
    for (cudaEvent_t* ev : {&c->ev_draw[0], &c->ev_draw[1], &c->ev_wg[0], &c->ev_wg[1], &c->ev_cat, &c->ev_join,
                            &c->ev_bk_main[0], &c->ev_bk_main[1], &c->ev_bk_main[2], &c->ev_bk_side[0], &c->ev_bk_side[1],
                            &c->ev_bk_side[2], &c->ev_comm})
      if (e == cudaSuccess) e = cudaEventCreateWithFlags(ev, cudaEventDisableTiming);
    if (e == cudaSuccess) {
      // the exchange kernels are tiny and latency-critical: highest priority, so that their CTAs take the first SM slots
      // that free up next to the resident backward kernels
      int lo_p = 0, hi_p = 0;
      cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
      e = cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, hi_p);
    }
    if (e != cudaSuccess) {
      crimac_set_error(std::string("side stream / event creation failed: ") + cudaGetErrorString(e));
      rc = 2;
    }
  }
  if (rc) {
    crimac_destroy(c);
    return rc;
  }
  *out = c;
  return 0;
}

extern "C" int crimac_destroy(crimac_ctx* c) {
  if (c) {
    for (cudaEvent_t ev : {c->ev_draw[0], c->ev_draw[1], c->ev_wg[0], c->ev_wg[1], c->ev_cat, c->ev_join})
      if (ev) cudaEventDestroy(ev);
    if (c->side) cudaStreamDestroy(c->side);
    for (cudaEvent_t ev : {c->ev_bk_main[0], c->ev_bk_main[1], c->ev_bk_main[2], c->ev_bk_side[0], c->ev_bk_side[1],
                           c->ev_bk_side[2], c->ev_comm})
      if (ev) cudaEventDestroy(ev);
    if (c->comm_stream) cudaStreamDestroy(c->comm_stream);
  }
  delete c;
  return 0;
}

extern "C" int crimac_prepare(crimac_ctx* c, const void* const* state, int train, void* stream) {
  int rc = check_call(c, state, 1);
  if (rc) return rc;
  CRIMAC_REQUIRE(!train || c->cfg.train, "context was created without train=1");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  {
    ProfScope ps("pack_weights", 0, 0, st, 2);
    PackTable t3{}, tt{};
    for (Conv3& L : c->conv)
      if (!L.first) t3.e[t3.n++] = PackEntry{S<float>(state, L.s_w), L.w_fwd, L.cout, L.cin, 0};
    CRIMAC_CHECK_CUDA(launch_pack_conv3x3_all(t3, st));
    if (c->cfg.up_mode == 1) {
      for (ConvT& U : c->up)   // (Cout,Cin,1,1): K contiguous already
        CRIMAC_CHECK_CUDA(launch_pack_cast(S<float>(state, U.s_w), U.w_fwd, static_cast<long>(U.cout) * U.cin, st));
    } else {
      for (ConvT& U : c->up) tt.e[tt.n++] = PackEntry{S<float>(state, U.s_w), U.w_fwd, U.cout, U.cin, 0};
      CRIMAC_CHECK_CUDA(launch_pack_convt_all(tt, st));
    }
  }
  if (!train)
    for (Conv3& L : c->conv)
      CRIMAC_CHECK_CUDA(launch_bn_fold_eval(S<float>(state, L.s_g), S<float>(state, L.s_beta), S<float>(state, L.s_rm),
                                            S<float>(state, L.s_rv), S<float>(state, L.s_b), 1e-5f, L.cout, L.scale,
                                            L.shift, st));
  c->prepared_mode = train ? 1 : 0;
  return 0;
}

struct StitchArgs {
  const int* centres;
  const uint8_t* nan_mask;
  const short* labels;
  const int* seabed;
  int seabed_pad, overlap, ping_start, Pc, R, K;
  int cls[4];
  void* out;
};

static int forward_impl(crimac_ctx* c, const void* const* state, const float* x, int nb, float* out, int softmax,
                        bool train, cudaStream_t st, bool skip_head = false, const StitchArgs* stitch = nullptr) {
  const int sms = device_num_sms();
  const int D = c->D;
  const int last = c->dec2[D - 2];
  // nn.BatchNorm2d raises "Expected more than 1 value per channel when training" (torch/nn/functional.py) when the
  // deepest level has a single value per channel; a silent var = 0 would produce huge gradients instead
  CRIMAC_REQUIRE(!train || static_cast<long>(nb) * level_h(c, D - 1) * level_w(c, D - 1) > 1,
                 "train-mode BatchNorm needs more than 1 value per channel at the deepest level (nb*H*W/4^(depth-1) > 1)");
  auto run_conv = [&](int idx) -> int {
    Conv3& L = c->conv[idx];
    const int H = level_h(c, L.level), W = level_w(c, L.level);
    int stat_rows = 0;  // partial rows written by the conv kernel (one per CTA)
    if (train) {
      View raw = with_batch(L.raw, nb);
      const double px = static_cast<double>(nb) * H * W;
      if (L.first) {
        ProfScope ps("first_conv", 2.0 * px * 64 * 9 * L.cin, px * (4.0 * L.cin + 2.0 * 64), st);
        stat_rows = first_conv_grid(nb, L.cin, H, W);
        CRIMAC_CHECK_CUDA(launch_first_conv(x, c->xs, S<float>(state, L.s_w), nullptr, S<float>(state, L.s_b), 0, nb, L.cin, H,
                                            W, raw.ptr, raw.pitch, c->stats, st));
      } else {
        ConvParams p = L.fwd;
        set_batch(p, nb);
        p.out = raw.ptr;
        p.out_pitch = raw.pitch;
        p.scale = nullptr;
        p.shift = S<float>(state, L.s_b);
        p.relu = 0;
        p.stats = c->stats;
        stat_rows = conv_grid(p.total_tiles, sms);
        ProfScope ps("conv3x3_fwd", igemm_flops_n(p, L.cout), px * 2.0 * (L.cin + L.cout), st);
        CRIMAC_CHECK_CUDA(launch_conv_igemm(p, L.bn_fwd, EPI_STATS, sms, st));
      }
      {
      ProfScope ps("bn_finalize", 0, 0, st);
      CRIMAC_CHECK_CUDA(launch_bn_finalize(c->stats, stat_rows, L.cout, static_cast<double>(nb) * H * W,
                                           S<float>(state, L.s_g), S<float>(state, L.s_beta), SM<float>(state, L.s_rm),
                                           SM<float>(state, L.s_rv), SM<long long>(state, L.s_nbt), 0.1f, 1e-5f, L.scale,
                                           L.shift, L.mean, L.invstd, st));
      }
      View pool = L.pool;
      if (pool.ptr) pool.N = nb;
      ProfScope ps("bn_apply", 0, px * L.cout * (pool.ptr ? 4.5 : 4.0), st);
      CRIMAC_CHECK_CUDA(launch_bn_apply(raw, L.scale, L.shift, with_batch(L.act, nb), pool, L.pool_arg, st));
    } else {
      const double px = static_cast<double>(nb) * H * W;
      if (L.first) {
        ProfScope ps("first_conv", 2.0 * px * 64 * 9 * L.cin, px * (4.0 * L.cin + 2.0 * 64), st);
        CRIMAC_CHECK_CUDA(launch_first_conv(x, c->xs, S<float>(state, L.s_w), L.scale, L.shift, 1, nb, L.cin, H, W, L.act.ptr,
                                            L.act.pitch, nullptr, st));
      } else {
        ConvParams p = L.fwd;
        set_batch(p, nb);
        p.scale = L.scale;
        p.shift = L.shift;
        p.relu = 1;
        int epi = EPI_STORE;
        if (idx == last) {
          // fused 1x1 head (+softmax): the 64-channel activation of the last conv never goes to HBM
          p.out = nullptr;
          p.head_w = S<float>(state, c->s_head_w);
          p.head_b = S<float>(state, c->s_head_b);
          p.head_out = out;
          p.n_classes = c->cfg.n_classes;
          p.head_softmax = softmax;
          if (stitch != nullptr) {
            p.head_stitch = 1;
            p.head_softmax = 1;
            p.st_centres = stitch->centres;
            p.st_nan = stitch->nan_mask;
            p.st_labels = stitch->labels;
            p.st_seabed = stitch->seabed;
            p.st_seabed_pad = stitch->seabed_pad;
            p.st_overlap = stitch->overlap;
            p.st_ping_start = stitch->ping_start;
            p.st_pc = stitch->Pc;
            p.st_r = stitch->R;
            p.st_k = stitch->K;
            for (int k = 0; k < 4; ++k) p.st_cls[k] = stitch->cls[k];
            p.st_out = stitch->out;
          }
          epi = EPI_HEAD;
        } else {
          p.out = L.act.ptr;
          p.out_pitch = L.act.pitch;
          p.pool_out = L.pool.ptr;
          p.pool_pitch = L.pool.pitch;
        }
        ProfScope ps(epi == EPI_HEAD ? "conv3x3_fwd_head" : "conv3x3_fwd", igemm_flops_n(p, L.cout),
                     px * 2.0 * (L.cin + L.cout), st);
        CRIMAC_CHECK_CUDA(launch_conv_igemm(p, L.bn_fwd, epi, sms, st));
      }
    }
    return 0;
  };
  int rc;
  for (int l = 0; l < D; ++l) {
    if ((rc = run_conv(c->enc1[l]))) return rc;
    if ((rc = run_conv(c->enc2[l]))) return rc;
  }
  for (int j = 0; j < D - 1; ++j) {
    ConvT& U = c->up[j];
    ConvParams p = U.fwd;
    set_batch(p, nb);
    p.scale = nullptr;
    p.shift = S<float>(state, U.s_b);
    p.relu = 0;
    if (c->cfg.up_mode == 1) {
      {
        ProfScope ps("up1x1_fwd", igemm_flops_n(p, U.cout), 2.0 * nb * p.H * p.W * (U.cin + U.cout), st);
        CRIMAC_CHECK_CUDA(launch_conv_igemm(p, U.bn_fwd, EPI_STORE, sms, st));
      }
      ProfScope ps("upsample2x", 0, 2.0 * nb * p.H * p.W * U.cout * 5.0, st);
      CRIMAC_CHECK_CUDA(launch_upsample2x(with_batch(U.tmp, nb), with_batch(U.out, nb), st));
    } else {
      ProfScope ps("convT_fwd", igemm_flops_n(p, 4 * U.cout), 2.0 * nb * p.H * p.W * (U.cin + 4.0 * U.cout), st);
      CRIMAC_CHECK_CUDA(launch_conv_igemm(p, U.bn_fwd, EPI_STORE, sms, st));
    }
    if ((rc = run_conv(c->dec1[j]))) return rc;
    if ((rc = run_conv(c->dec2[j]))) return rc;
  }
  if (train && !skip_head) {
    ProfScope ps("head_fwd", 0, static_cast<double>(nb) * c->cfg.height * c->cfg.width * (128.0 + 4.0 * c->cfg.n_classes), st);
    CRIMAC_CHECK_CUDA(launch_head_fwd(with_batch(c->conv[last].act, nb), S<float>(state, c->s_head_w),
                                      S<float>(state, c->s_head_b), c->cfg.n_classes, out, st));
  }
  return 0;
}

extern "C" int crimac_forward_infer(crimac_ctx* c, const void* const* state, const float* x, int nb, float* out,
                                    int softmax, void* stream) {
  int rc = check_call(c, state, nb);
  if (rc) return rc;
  CRIMAC_REQUIRE(out != nullptr, "NULL tensor");
  CRIMAC_REQUIRE(c->prepared_mode == 0, "call crimac_prepare(ctx, state, train=0) first");
  if (x == nullptr) {
    // the first conv's operand was staged by crimac_preprocess_staged: tensor-core first conv only (<= 8 frequencies)
    CRIMAC_REQUIRE(c->staged_nb == nb, "x is NULL but crimac_preprocess_staged has not staged exactly nb patches");
    CRIMAC_REQUIRE(first_conv_grid(nb, c->cfg.in_channels, c->cfg.height, c->cfg.width) ==
                       first_conv_tc_grid(nb, c->cfg.height, c->cfg.width),
                   "staged input needs the tensor-core first conv (CRIMAC_FC_CUDACORE is set)");
  }
  c->staged_nb = 0;
  return forward_impl(c, state, x, nb, out, softmax, false, static_cast<cudaStream_t>(stream));
}

extern "C" int crimac_forward_train(crimac_ctx* c, const void* const* state, const float* x, int nb, float* logits,
                                    void* stream) {
  int rc = check_call(c, state, nb);
  if (rc) return rc;
  CRIMAC_REQUIRE(x != nullptr && logits != nullptr, "NULL tensor");
  CRIMAC_REQUIRE(c->cfg.train, "context was created without train=1");
  CRIMAC_REQUIRE(c->prepared_mode == 1, "call crimac_prepare(ctx, state, train=1) first");
  return forward_impl(c, state, x, nb, logits, 0, true, static_cast<cudaStream_t>(stream));
}

extern "C" int crimac_loss(crimac_ctx* c, const float* logits, const int64_t* labels, const float* class_w,
                           int64_t ignore_index, int nb, float* out3, float* dlogits, void* stream) {
  CRIMAC_REQUIRE(c != nullptr && c->cfg.train, "train context required");
  CRIMAC_REQUIRE(nb >= 1 && nb <= c->cfg.max_batch, "nb");
  CRIMAC_REQUIRE(logits && labels && class_w && out3, "NULL tensor");
  CRIMAC_CHECK_CUDA(cudaSetDevice(c->device));
  ProfScope ps("ce_loss", 0, static_cast<double>(nb) * c->cfg.height * c->cfg.width * (8.0 * c->cfg.n_classes + 8.0),
               static_cast<cudaStream_t>(stream), 2);
  CRIMAC_CHECK_CUDA(launch_ce(logits, reinterpret_cast<const long long*>(labels), class_w, c->cfg.n_classes, nb,
                              static_cast<long>(c->cfg.height) * c->cfg.width, ignore_index, dlogits, c->ce_partials,
                              out3, static_cast<cudaStream_t>(stream)));
  return 0;
}

// head_done: the fused head/CE kernel has already written the UNNORMALISED gradient of the last activation into GA and
// the head's own gradients; gscale (1/sum_w) is then folded into the last layer's BatchNorm backward.
static int backward_impl(crimac_ctx* c, const void* const* state, const float* x, const float* dlogits,
                         const float* gscale, int nb, float* const* grads, cudaStream_t st, bool head_done) {
  int rc;
  const int sms = device_num_sms();
  const int D = c->D;
  const bool saved_overlap = c->overlap;
  struct Restore {
    crimac_ctx* c;
    bool v;
    ~Restore() { c->overlap = v; }
  } restore{c, saved_overlap};
  if (crimac_profiling()) c->overlap = false;  // isolated per-kernel timings
  if (c->wg_dirty) {
    ProfScope ps("wgrad_zero", 0, static_cast<double>(c->wg_arena_bytes), st);
    CRIMAC_CHECK_CUDA(cudaMemsetAsync(c->wg_arena, 0, c->wg_arena_bytes, st));
  }
  c->wg_dirty = true;  // cleared by the unpack launch at the end

  // BN/ReLU backward of layer idx (dA in GA) -> dRaw in GR; then backward-data (into L.gin) and the weight gradient
  bool wg_recorded[2] = {false, false};
  // pool_arg != nullptr (encoder second convs below the deepest level): the incoming gradient dA = dSkip + unpool(dPool)
  // is formed inside the BatchNorm-backward kernels instead of being materialised by a separate pass
  auto conv_bwd = [&](int idx, const uint16_t* pool_arg = nullptr, View dpool = View{}, View dskip = View{}) -> int {
    Conv3& L = c->conv[idx];
    const int H = level_h(c, L.level), W = level_w(c, L.level);
    View da{c->GA, nb, H, W, L.cout, L.cout};
    View dr{c->GR[L.gr_idx], nb, H, W, L.cout, L.cout};
    const double px = static_cast<double>(nb) * H * W;
    // the weight gradient that last read this dRaw buffer (two layers ago, in THIS call) must have finished.  Events of
    // an earlier call are never waited on: the call ends with a join, and a stream capture must not depend on them.
    if (c->overlap && wg_recorded[L.gr_idx]) CRIMAC_CHECK_CUDA(cudaStreamWaitEvent(st, c->ev_wg[L.gr_idx], 0));
    {
      ProfScope ps(pool_arg ? "bn_relu_bwd_pool" : "bn_relu_bwd", 0, px * L.cout * (pool_arg ? 11.2 : 10.0), st, 3);
      CRIMAC_CHECK_CUDA(launch_bn_bwd(da, with_batch(L.raw, nb), L.scale, L.shift, L.mean, L.invstd, dr, grads[L.g_g],
                                      grads[L.g_beta], grads[L.g_b], 0, c->red_partials, c->c1c2,
                                      (head_done && idx == c->dec2[c->D - 2]) ? gscale : nullptr, st, pool_arg, dpool,
                                      dskip));
    }
    if (c->overlap) CRIMAC_CHECK_CUDA(cudaEventRecord(c->ev_draw[L.gr_idx], st));
    // backward-data first (critical path, issued first so that it is scheduled first) ...
    if (!L.first && L.gin.ptr != nullptr) {
      ConvParams p = L.dgrad;
      set_batch(p, nb);
      ProfScope ps("conv3x3_dgrad", igemm_flops_n(p, L.cin), px * 2.0 * (L.cin + L.cout), st);
      CRIMAC_CHECK_CUDA(launch_conv_igemm(p, L.bn_bwd, EPI_STORE, sms, st));
    }
    // ... then the weight gradient: tensor-bound and off the critical path -> side stream, where it overlaps the
    // HBM-bound BatchNorm / pooling backward kernels of the following layers (dRaw is double-buffered for this)
    cudaStream_t ws = st;
    if (c->overlap) {
      CRIMAC_CHECK_CUDA(cudaStreamWaitEvent(c->side, c->ev_draw[L.gr_idx], 0));
      ws = c->side;
    }
    if (L.first) {
      ProfScope ps("first_conv_wgrad", 2.0 * px * 64 * 9 * L.cin, px * (4.0 * L.cin + 2.0 * 64), ws, 2);
      CRIMAC_CHECK_CUDA(launch_first_conv_wgrad(x, c->xs, dr, L.cin, c->fc_partials, grads[L.g_w], 0, ws));
    } else {
      int r = L.wg_use_halo ? wgrad_halo_run(c, L.wg, nb, ws, L.wg_slabs, L.wg_max_splits, &L.wg_active)
                            : wgrad_run(c, L.wg_tap, L.bn_wg, nb, ws, L.wg_slabs, L.wg_max_splits, &L.wg_active);
      if (r) return r;
    }
    if (c->overlap) {
      CRIMAC_CHECK_CUDA(cudaEventRecord(c->ev_wg[L.gr_idx], c->side));
      wg_recorded[L.gr_idx] = true;
    }
    return 0;
  };

  // Gradient buckets in backward order: 0 = decoder + head (the tail of parameters()), 1 = the two deepest encoder
  // blocks, 2 = the remaining encoder blocks.  When a bucket's last kernels have been issued, its weight-gradient scratch
  // is un-packed ON THE SIDE STREAM (in order behind the GEMMs that filled it) and - data parallel - its slice of the
  // flat gradient arena is all-reduced over peer memory on the communication stream while backward continues.
  const int enc_split = D >= 3 ? D - 2 : 0;  // encoder blocks >= enc_split form bucket 1
  auto close_bucket = [&](int b) -> int {
    UnpackTable t{};
    auto add3 = [&](const Conv3& L) {
      if (L.first) return;
      UnpackEntry e{L.wg_scratch, grads[L.g_w], static_cast<long>(L.cout) * L.cin, 9, 0};
      if (L.wg_slabs) {
        e.slabs = L.wg_slabs;
        e.splits = L.wg_active;
        e.slab_stride = 9L * L.cout * L.cin;
      }
      t.e[t.n++] = e;
    };
    if (b == 0) {
      for (int j = 0; j < D - 1; ++j) {
        add3(c->conv[c->dec1[j]]);
        add3(c->conv[c->dec2[j]]);
        const ConvT& U = c->up[j];
        const int utaps = c->cfg.up_mode == 1 ? 1 : 4;
        UnpackEntry e{U.wg_scratch, grads[U.g_w], static_cast<long>(U.cin) * U.cout, utaps, 0};
        if (U.wg_slabs) {
          e.slabs = U.wg_slabs;
          e.splits = U.wg_active;
          e.slab_stride = static_cast<long>(utaps) * U.cin * U.cout;
        }
        t.e[t.n++] = e;
      }
    } else {
      const int l0 = b == 1 ? enc_split : 0, l1 = b == 1 ? D : enc_split;
      for (int l = l0; l < l1; ++l) {
        add3(c->conv[c->enc1[l]]);
        add3(c->conv[c->enc2[l]]);
      }
    }
    cudaStream_t us = c->overlap ? c->side : st;
    if (t.n > 0) {
      ProfScope ps("wgrad_unpack", 0, 0, us);
      CRIMAC_CHECK_CUDA(launch_wgrad_unpack_all(t, us));
    }
    const bool opt_on = c->opt_on && head_done;   // the optimizer rides only on the fused train step, never on autograd's backward
    if (!c->comm_on && !opt_on) return 0;
    // arena slice of the bucket, from the gradient table (parameters() order: encoder blocks, decoder blocks, head)
    const float* base = c->comm_on ? c->comm.peer_arenas[c->comm.rank] : c->opt.grads;
    const size_t arena_floats = c->comm_on ? c->comm.arena_floats : (c->opt.n + 3) & ~static_cast<size_t>(3);
    const float* lo = b == 0 ? grads[c->up[0].g_w] : grads[c->conv[c->enc1[b == 1 ? enc_split : 0]].g_w];
    const float* hi = b == 0 ? base + arena_floats
                             : (b == 1 ? grads[c->up[0].g_w] : grads[c->conv[c->enc1[enc_split]].g_w]);
    if (hi <= lo) return 0;
    CRIMAC_REQUIRE(lo >= base && hi <= base + arena_floats && (lo - base) % 4 == 0,
                   "gradient tensors are not views of the flat arena given to crimac_set_comm / crimac_set_optimizer (parameters() order, 16-byte aligned)");
    CRIMAC_REQUIRE(!(c->comm_on && opt_on) || c->opt.grads == base, "crimac_set_optimizer: grads must be the arena of crimac_set_comm");
    size_t count = static_cast<size_t>(hi - lo);
    count = (count + 3) & ~static_cast<size_t>(3);   // the arena is padded to a multiple of 4 floats
    const size_t off = static_cast<size_t>(lo - base);
    cudaStream_t cs = st;
    if (c->overlap) {
      // the bucket's gradients are final once the main stream (BatchNorm / bias gradients) and the side stream (weight
      // gradients, un-pack) have reached this point; exchange and optimizer then run beside the rest of backward
      CRIMAC_CHECK_CUDA(cudaEventRecord(c->ev_bk_main[b], st));
      CRIMAC_CHECK_CUDA(cudaEventRecord(c->ev_bk_side[b], c->side));
      CRIMAC_CHECK_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_bk_main[b], 0));
      CRIMAC_CHECK_CUDA(cudaStreamWaitEvent(c->comm_stream, c->ev_bk_side[b], 0));
      cs = c->comm_stream;
    }
    if (c->comm_on) {
      int r = crimac_peer_allreduce(c->comm.peer_arenas, c->comm.peer_pads, c->comm.multicast_arena, c->comm.local_state,
                                    c->comm.rank, c->comm.world, b, off, count, c->comm.ctas, cs);
      if (r) return r;
    }
    if (opt_on && off < c->opt.n) {
      // SGD(momentum) on this bucket's slice right behind its exchange: nothing later in this backward reads the fp32
      // parameters of a closed bucket (the GEMMs use the bf16 operands packed at the start of the step, BatchNorm
      // backward the scale / shift saved in forward), and the next step's re-pack is ordered behind the final join
      const size_t n = std::min(count, c->opt.n - off);
      ProfScope ps("sgd", 0, 16.0 * n, cs);
      int r = crimac_sgd_step(c->opt.params + off, c->opt.momentum + off, c->opt.grads + off, n, c->opt.lr,
                              c->opt.momentum_coef, c->opt.gscale, cs);
      if (r) return r;
    }
    return 0;
  };

  // head
  const int last = c->dec2[D - 2];
  if (!head_done) {
    Conv3& L = c->conv[last];
    View dact{c->GA, nb, c->cfg.height, c->cfg.width, 64, 64};
    ProfScope ps("head_bwd", 0, static_cast<double>(nb) * c->cfg.height * c->cfg.width * (256.0 + 4.0 * c->cfg.n_classes), st, 2);
    CRIMAC_CHECK_CUDA(launch_head_bwd(dlogits, gscale, with_batch(L.act, nb), S<float>(state, c->s_head_w),
                                      c->cfg.n_classes, dact, c->head_partials, grads[c->g_head_w],
                                      grads[c->g_head_b], 0, st));
  }
  // decoder, last block first
  for (int j = D - 2; j >= 0; --j) {
    if ((rc = conv_bwd(c->dec2[j]))) return rc;   // dgrad -> dA of dec1[j]
    if ((rc = conv_bwd(c->dec1[j]))) return rc;   // dgrad wrote dCat_j
    ConvT& U = c->up[j];
    const bool ups = c->cfg.up_mode == 1;
    if (ups) {
      // adjoint of the bilinear 2x: dCat_j[:, :C] (full resolution) -> dTmp (low resolution); the 1x1 conv's backward
      // then reads dTmp
      ProfScope ps("upsample2x_bwd", 0, 2.0 * nb * U.gout.H * U.gout.W * U.cout * 1.25, st);
      CRIMAC_CHECK_CUDA(launch_upsample2x_bwd(with_batch(U.gout, nb), with_batch(U.dtmp, nb), st));
    }
    // everything that only READS dCat_j (dTmp) and is off the critical path goes to the side stream: the bias gradient
    // (HBM-bound column sum, own partial buffer) and the weight gradient; the main stream continues with the
    // backward-data GEMM
    cudaStream_t ss = c->overlap ? c->side : st;
    if (c->overlap) {
      CRIMAC_CHECK_CUDA(cudaEventRecord(c->ev_cat, st));
      CRIMAC_CHECK_CUDA(cudaStreamWaitEvent(c->side, c->ev_cat, 0));
    }
    {
      const View gsrc = with_batch(ups ? U.dtmp : U.gout, nb);
      ProfScope ps("convT_bias_grad", 0, 2.0 * nb * gsrc.H * gsrc.W * U.cout, ss, 2);
      CRIMAC_CHECK_CUDA(launch_view_colsum(gsrc, c->colsum_partials, grads[U.g_b], 0, ss));
    }
    {
      ConvParams p = U.dgrad;
      set_batch(p, nb);
      ProfScope ps("convT_dgrad", igemm_flops_n(p, U.cin), 2.0 * nb * p.H * p.W * (U.cin + 4.0 * U.cout), st);
      CRIMAC_CHECK_CUDA(launch_conv_igemm(p, U.bn_bwd, EPI_STORE, sms, st));
    }
    if ((rc = wgrad_run(c, U.wg, ups ? pick_bn(U.cin) : U.bn_wg, nb, ss, U.wg_slabs, U.wg_max_splits, &U.wg_active))) return rc;
  }
  if ((rc = close_bucket(0))) return rc;
  // encoder, deepest level first
  for (int l = D - 1; l >= 0; --l) {
    Conv3& L2 = c->conv[c->enc2[l]];
    if (l < D - 1) {
      // max-pool backward + skip-gradient add (autograd of unet.py:86,92,132) fused into the BatchNorm backward of this layer
      const int j = D - 2 - l;
      View dpool{c->GP, nb, level_h(c, l + 1), level_w(c, l + 1), L2.cout, L2.cout};
      View dskip{c->dskip[j], nb, level_h(c, l), level_w(c, l), L2.cout, L2.cout};
      if ((rc = conv_bwd(c->enc2[l], L2.pool_arg, dpool, dskip))) return rc;  // dgrad -> dA of enc1[l]
    } else {
      if ((rc = conv_bwd(c->enc2[l]))) return rc;
    }
    if ((rc = conv_bwd(c->enc1[l]))) return rc;  // dgrad wrote GP (grad of the pooled input), none for l == 0
    if (l == enc_split && enc_split > 0 && (rc = close_bucket(1))) return rc;
  }
  if ((rc = close_bucket(enc_split > 0 ? 2 : 1))) return rc;
  if (c->overlap) {
    CRIMAC_CHECK_CUDA(cudaEventRecord(c->ev_join, c->side));
    CRIMAC_CHECK_CUDA(cudaStreamWaitEvent(st, c->ev_join, 0));
    if (c->comm_on || (c->opt_on && head_done)) {
      CRIMAC_CHECK_CUDA(cudaEventRecord(c->ev_comm, c->comm_stream));
      CRIMAC_CHECK_CUDA(cudaStreamWaitEvent(st, c->ev_comm, 0));
    }
  }
  c->wg_dirty = false;  // every layer's scratch has been un-packed (and re-zeroed) by the three bucket launches
  return 0;
}

// Data-parallel gradient exchange: from now on every crimac_backward / crimac_train_step on this context all-reduces (sum)
// the gradient arena over the replicas with crimac_peer_allreduce, bucket by bucket, overlapped with backward.  The
// `grads` table of those calls must then point INTO the symmetric arena (parameters() order).  cfg == NULL switches the
// exchange off.  Every replica must issue the same sequence of calls.
extern "C" int crimac_set_comm(crimac_ctx* c, const crimac_comm_config* cfg) {
  CRIMAC_REQUIRE(c != nullptr && c->cfg.train, "train context required");
  if (cfg == nullptr) {
    c->comm_on = false;
    return 0;
  }
  CRIMAC_REQUIRE(cfg->world >= 1 && cfg->world <= 8 && cfg->rank >= 0 && cfg->rank < cfg->world, "rank / world");
  CRIMAC_REQUIRE(cfg->local_state != nullptr && cfg->arena_floats % 4 == 0, "local_state / arena_floats (multiple of 4)");
  for (int r = 0; r < cfg->world; ++r)
    CRIMAC_REQUIRE(cfg->peer_arenas[r] != nullptr && cfg->peer_pads[r] != nullptr, "NULL peer pointer");
  c->comm = *cfg;
  c->comm_on = cfg->world > 1;
  return 0;
}

extern "C" int crimac_backward(crimac_ctx* c, const void* const* state, const float* x, const float* dlogits,
                               const float* gscale, int nb, float* const* grads, void* stream) {
  int rc = check_call(c, state, nb);
  if (rc) return rc;
  CRIMAC_REQUIRE(c->cfg.train && c->prepared_mode == 1, "backward needs a train context after forward_train");
  CRIMAC_REQUIRE(x && dlogits && grads, "NULL tensor");
  return backward_impl(c, state, x, dlogits, gscale, nb, grads, static_cast<cudaStream_t>(stream), false);
}

extern "C" int crimac_train_step(crimac_ctx* c, const void* const* state, const float* x, const int64_t* labels,
                                 const float* class_w, int64_t ignore_index, int nb, float* const* grads, float* loss3,
                                 void* stream) {
  int rc = crimac_prepare(c, state, 1, stream);
  if (rc) return rc;
  CRIMAC_REQUIRE(x && labels && class_w && grads, "NULL tensor");
  CRIMAC_REQUIRE(nb >= 1 && nb <= c->cfg.max_batch, "nb must be in 1..max_batch");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // forward without the head: logits are never materialised on the fused path
  if ((rc = forward_impl(c, state, x, nb, nullptr, 0, true, st, /*skip_head=*/true))) return rc;
  float* l3 = loss3 ? loss3 : c->loss3;
  {
    const int last = c->dec2[c->D - 2];
    const double px = static_cast<double>(nb) * c->cfg.height * c->cfg.width;
    View dact{c->GA, nb, c->cfg.height, c->cfg.width, 64, 64};
    ProfScope ps("head_ce_fused", 0, px * (256.0 + 8.0), st, 2);
    CRIMAC_CHECK_CUDA(launch_head_ce_fused(with_batch(c->conv[last].act, nb), S<float>(state, c->s_head_w),
                                           S<float>(state, c->s_head_b), c->cfg.n_classes,
                                           reinterpret_cast<const long long*>(labels), class_w, ignore_index, dact,
                                           c->head_partials, c->ce_partials, grads[c->g_head_w], grads[c->g_head_b], l3,
                                           st));
  }
  return backward_impl(c, state, x, nullptr, l3 + 1, nb, grads, st, /*head_done=*/true);
}

// Test hook: copies one tensor the last crimac_forward_train / crimac_train_step saved for backward out of the
// workspace into a dense NHWC bf16 buffer (nb, H, W, C).  which: 0 = raw conv output (pre-BatchNorm), 1 = activation
// (post BN + ReLU), 2 = 2x2 max-pooled activation (encoder second convs), 3 = ConvTranspose2d output (index = decoder
// block).  index (which < 3): encoder block l -> 2l (first conv), 2l+1 (second); decoder block j -> 2*depth + 2j, +1.
// dims_out (host, optional): {H, W, C}.  The gradient parity test uses it to evaluate torch autograd AT the native
// forward state (tests/test_gpu_unet.py::test_backward_at_the_native_forward_state).
extern "C" int crimac_dbg_saved(crimac_ctx* c, int index, int which, int nb, void* dst_dev, int* dims_out, void* stream) {
  CRIMAC_REQUIRE(c != nullptr && c->cfg.train, "train context required");
  CRIMAC_REQUIRE(nb >= 1 && nb <= c->cfg.max_batch, "nb");
  View v{};
  if (which == 3) {
    CRIMAC_REQUIRE(index >= 0 && index < static_cast<int>(c->up.size()), "decoder block index");
    v = c->up[index].out;
  } else {
    CRIMAC_REQUIRE(which >= 0 && which <= 2, "which");
    const int D = c->D;
    int idx = -1;
    if (index >= 0 && index < 2 * D) idx = (index & 1) ? c->enc2[index >> 1] : c->enc1[index >> 1];
    else if (index >= 2 * D && index < 2 * D + 2 * (D - 1)) idx = ((index - 2 * D) & 1) ? c->dec2[(index - 2 * D) >> 1] : c->dec1[(index - 2 * D) >> 1];
    CRIMAC_REQUIRE(idx >= 0, "layer index");
    const Conv3& L = c->conv[idx];
    v = which == 0 ? L.raw : (which == 1 ? L.act : L.pool);
    CRIMAC_REQUIRE(v.ptr != nullptr, "this layer does not keep that tensor");
  }
  if (dims_out) {
    dims_out[0] = v.H;
    dims_out[1] = v.W;
    dims_out[2] = v.C;
  }
  if (dst_dev != nullptr)
    CRIMAC_CHECK_CUDA(cudaMemcpy2DAsync(dst_dev, static_cast<size_t>(v.C) * 2, v.ptr, static_cast<size_t>(v.pitch) * 2,
                                        static_cast<size_t>(v.C) * 2, static_cast<size_t>(nb) * v.H * v.W,
                                        cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return 0;
}

// Preprocessing that FEEDS THE FIRST CONV DIRECTLY (north_star): the patch gather + NaN fill + dB transform
// of crimac_preprocess, written as the bf16 hi/lo NHWC operand the tensor-core first conv reads by TMA - into the
// context's own staging buffer.  The next crimac_forward_infer(ctx, state, x_dev = NULL, nb = n, ...) consumes it; the
// fp32 NCHW patch tensor of the reference (batch/dataset.py:192-205 -> pipeline.py:208) is never written.  Patch size =
// the context's (height, width).
extern "C" int crimac_preprocess_staged(crimac_ctx* c, const float* sv, int F, int R, int P, int data_ping0,
                                        const int32_t* centres, int n, uint8_t* nan_mask, void* stream) {
  CRIMAC_REQUIRE(c != nullptr && sv != nullptr && centres != nullptr, "NULL argument");
  CRIMAC_REQUIRE(F == c->cfg.in_channels, "F must equal the context's in_channels");
  CRIMAC_REQUIRE(n >= 1 && n <= c->cfg.max_batch, "n must be in 1..max_batch");
  CRIMAC_REQUIRE(R >= 1 && P >= 1, "empty input");
  CRIMAC_CHECK_CUDA(cudaSetDevice(c->device));
  ProfScope ps("preprocess_staged", 0, static_cast<double>(n) * c->cfg.height * c->cfg.width * (4.0 * F + 16.0 * (F <= 4 ? 1 : 2)),
               static_cast<cudaStream_t>(stream));
  int rc = launch_preprocess_split(sv, F, R, P, data_ping0, centres, n, c->cfg.height, c->cfg.width, c->xs, nan_mask,
                                   static_cast<cudaStream_t>(stream));
  if (rc) {
    crimac_set_error("preprocess_split_kernel launch failed");
    return rc;
  }
  c->staged_nb = n;
  return 0;
}

// Fused optimizer: from now on every crimac_backward / crimac_train_step on this context also applies
// v = momentum*v + g*gscale ; p -= lr*v (torch.optim.SGD(momentum), pipeline.py:156,178) to each gradient bucket as soon
// as it is final (and, data parallel, all-reduced) - on the communication stream, beside the rest of backward.
// cfg->params / momentum / grads are flat fp32 arrays in parameters() order; the `grads` table of those calls must point
// into cfg->grads.  gscale = 1/world.  cfg == NULL (or params == NULL) switches it off.  lr is baked into the launch: call
// again when the schedule changes it (and re-capture a CUDA graph of the step).
extern "C" int crimac_set_optimizer(crimac_ctx* c, const crimac_optimizer_config* cfg) {
  CRIMAC_REQUIRE(c != nullptr && c->cfg.train, "train context required");
  if (cfg == nullptr || cfg->params == nullptr) {
    c->opt_on = false;
    return 0;
  }
  CRIMAC_REQUIRE(cfg->momentum != nullptr && cfg->grads != nullptr && cfg->n > 0, "NULL array");
  CRIMAC_REQUIRE(((reinterpret_cast<uintptr_t>(cfg->params) | reinterpret_cast<uintptr_t>(cfg->momentum) |
                   reinterpret_cast<uintptr_t>(cfg->grads)) & 15) == 0, "arrays must be 16-byte aligned");
  c->opt = *cfg;
  c->opt_on = true;
  return 0;
}

// Sliding-window inference with the overlap stitching fused into the last conv's epilogue: eval forward + softmax, and
// every kept pixel's classes cls[0..K) go straight into the chunk's (K, R, Pc) fp16 output (fill_out_array,
// save_predict.py:41-65, with the label masks of crimac_stitch) - the (nb, n_classes, H, W) probability tensor is never
// written.  x_dev may be NULL after crimac_preprocess_staged (whose nan_dev output is this call's nan_dev input).
extern "C" int crimac_forward_infer_stitch(crimac_ctx* c, const void* const* state, const float* x, int nb,
                                           const int32_t* centres, const uint8_t* nan_mask, const int16_t* labels,
                                           const int32_t* seabed, int seabed_pad, int overlap, int ping_start, int Pc,
                                           int R, const int32_t* cls, int K, void* out, void* stream) {
  int rc = check_call(c, state, nb);
  if (rc) return rc;
  CRIMAC_REQUIRE(centres != nullptr && out != nullptr && cls != nullptr, "NULL tensor");
  CRIMAC_REQUIRE(c->prepared_mode == 0, "call crimac_prepare(ctx, state, train=0) first");
  CRIMAC_REQUIRE(K >= 1 && K <= 4, "K must be 1..4");
  for (int k = 0; k < K; ++k) CRIMAC_REQUIRE(cls[k] >= 0 && cls[k] < c->cfg.n_classes, "class index out of range");
  if (x == nullptr) CRIMAC_REQUIRE(c->staged_nb == nb, "x is NULL but crimac_preprocess_staged has not staged exactly nb patches");
  c->staged_nb = 0;
  StitchArgs sa{centres, nan_mask, labels, seabed, seabed_pad, overlap, ping_start, Pc, R, K, {0, 0, 0, 0}, out};
  for (int k = 0; k < K; ++k) sa.cls[k] = cls[k];
  return forward_impl(c, state, x, nb, nullptr, 1, false, static_cast<cudaStream_t>(stream), false, &sa);
}

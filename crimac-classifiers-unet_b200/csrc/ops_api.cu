// Op-level C-ABI entry points (crimac_op_*): one kernel each, tensor maps encoded per call.  The network-level
// entry points in net_api.cu drive the same launchers with maps cached in the context; these exist so the parity
// tests can pin every kernel against the oracle in isolation.
#include "host_util.h"
#include "../../include/crimac_b200.h"

static int pick_block_n(int n_total, int requested) {
  if (requested == 64 || requested == 128 || requested == 256) return (n_total % requested == 0) ? requested : 0;
  if (n_total % 256 == 0) return 256;
  if (n_total % 128 == 0) return 128;
  if (n_total % 64 == 0) return 64;
  return 0;
}

// Generic implicit GEMM.  mode: 0 = 3x3 conv (taps 9, halo main loop), 1 = 1x1 / ConvTranspose forward (taps 1; convt_cout > 0 turns on
// the 2x upsampling scatter), 2 = ConvTranspose backward-data (taps 4: x is the (2H x 2W) gradient, sub-sampled).
// x: NHWC bf16 view (NB,H,W,cin) with pixel pitch x_pitch (for mode 2: dims of x are 2H x 2W).
// w: packed bf16 [n_total][taps*cin].   out: NHWC bf16 with pitch out_pitch.
// stats != NULL selects the train-mode epilogue (raw + per-tile channel sums); head_w != NULL the fused head.
extern "C" int crimac_op_igemm(int mode, const void* x, int NB, int H, int W, int cin, int x_pitch, const void* w,
                               int n_total, const float* scale, const float* shift, int relu, void* out,
                               int out_pitch, int convt_cout, void* pool_out, int pool_pitch, float* stats,
                               const float* head_w, const float* head_b, float* head_out, int n_classes,
                               int head_softmax, int block_n, void* stream) {
  CRIMAC_REQUIRE(mode >= 0 && mode <= 5 && mode != 3, "mode");
  // modes 4 / 5: backward-data of a 3x3 conv / ConvTranspose reading the FORWARD-packed weights as an MN-major operand
  // (w = [Cout][9*Cin] resp. [(kk,co)][Cin] of the forward layer; n_total = the forward layer's Cin)
  const bool b_mn = (mode == 4 || mode == 5);
  if (mode == 4) mode = 0;
  if (mode == 5) mode = 2;
  const bool halo = (mode == 0);
  CRIMAC_REQUIRE(cin % 64 == 0, "cin must be a multiple of 64");
  const int bn = pick_block_n(n_total, head_w ? 64 : block_n);
  CRIMAC_REQUIRE(bn != 0, "n_total must be a multiple of 64 (and of block_n when given)");
  ConvParams p{};
  p.taps = (mode == 0) ? 9 : (mode == 1 ? 1 : 4);
  p.tap_mode = (mode == 2) ? 1 : 0;
  p.cin = cin;
  p.NB = NB;
  p.H = H;
  p.W = W;
  p.halo = halo ? 1 : 0;
  p.tiles_x = halo ? (W + 7) / 8 : (W + TILE_W - 1) / TILE_W;
  p.tiles_y = halo ? (H + 15) / 16 : (H + TILE_H - 1) / TILE_H;
  p.n_tiles = n_total / bn;
  p.total_tiles = NB * p.tiles_x * p.tiles_y * p.n_tiles;
  if (mode == 2) {
    View v{static_cast<bf16*>(const_cast<void*>(x)), NB, 2 * H, 2 * W, cin, x_pitch};
    for (int kk = 0; kk < 4; ++kk) {
      int rc = make_act_map(&p.a_map[kk], v, TILE_H, 1, kk >> 1, kk & 1);
      if (rc) return rc;
    }
  } else {
    View v{static_cast<bf16*>(const_cast<void*>(x)), NB, H, W, cin, x_pitch};
    int rc = halo ? make_act_map(&p.a_map[0], v, 18, 0, 0, 0, 10) : make_act_map(&p.a_map[0], v, TILE_H);
    if (rc) return rc;
  }
  int rc;
  if (b_mn) {
    CRIMAC_REQUIRE(stats == nullptr && head_w == nullptr, "backward-data has the plain store epilogue");
    p.b_mn = 1;
    p.b_tap_cols = n_total;
    rc = (mode == 0) ? make_weight_map(&p.b_map, static_cast<const bf16*>(w), cin, 9 * n_total, 64)
                     : make_weight_map(&p.b_map, static_cast<const bf16*>(w), 4 * cin, n_total, 64);
  } else {
    rc = make_weight_map(&p.b_map, static_cast<const bf16*>(w), n_total, p.taps * cin, bn);
  }
  if (rc) return rc;
  p.out = static_cast<bf16*>(out);
  p.out_pitch = out_pitch;
  p.relu = relu;
  p.convt_cout = convt_cout;
  p.scale = scale;
  p.shift = shift;
  p.pool_out = static_cast<bf16*>(pool_out);
  p.pool_pitch = pool_pitch;
  p.stats = stats;
  p.head_w = head_w;
  p.head_b = head_b;
  p.head_out = head_out;
  p.n_classes = n_classes;
  p.head_softmax = head_softmax;
  int epi = EPI_STORE;
  if (stats) epi = EPI_STATS;
  if (head_w) {
    CRIMAC_REQUIRE(n_total == 64 && n_classes >= 1 && n_classes <= CRIMAC_MAX_CLASSES, "fused head needs Cout == 64");
    epi = EPI_HEAD;
  }
  CRIMAC_CHECK_CUDA(launch_conv_igemm(p, bn, epi, device_num_sms(), static_cast<cudaStream_t>(stream)));
  return 0;
}

// Weight gradient.  mode 0: 3x3 conv  (f = dY (NB,H,W,m_total), t = X (NB,H,W,n_total), 9 taps)
//                   mode 1: ConvTranspose (f = X (NB,H,W,m_total), t = dY (NB,2H,2W,n_total), 4 taps)
// scratch: fp32 [taps][m_total][n_total] (zeroed here when splits > 1); dw: fp32 PyTorch layout [m][n][taps].
extern "C" int crimac_op_wgrad(int mode, const void* f, int f_pitch, int m_total, const void* t, int t_pitch,
                               int n_total, int NB, int H, int W, float* scratch, float* dw, int splits, int block_n,
                               void* stream) {
  CRIMAC_REQUIRE(mode == 0 || mode == 1, "mode");
  CRIMAC_REQUIRE(m_total % 64 == 0 && n_total % 64 == 0, "channel counts must be multiples of 64");
  const int bn = pick_block_n(n_total, block_n);
  CRIMAC_REQUIRE(bn != 0, "n_total/block_n");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WgradParams p{};
  p.taps = mode == 0 ? 9 : 4;
  p.tap_mode = mode;
  p.M_total = m_total;
  p.N_total = n_total;
  p.NB = NB;
  p.H = H;
  p.W = W;
  p.tiles_x = (W + 15) / 16;
  p.tiles_y = (H + 3) / 4;
  p.k_tiles_total = NB * p.tiles_x * p.tiles_y;
  p.m_tiles = (m_total + 127) / 128;
  p.n_tiles = n_total / bn;
  if (splits <= 0) {
    const int tiles = p.taps * p.m_tiles * p.n_tiles;
    splits = (2 * device_num_sms() + tiles - 1) / tiles;
    if (splits > p.k_tiles_total / 8) splits = p.k_tiles_total / 8;
    if (splits < 1) splits = 1;
  }
  p.splits = splits;
  p.dw = scratch;
  View vf{static_cast<bf16*>(const_cast<void*>(f)), NB, H, W, m_total, f_pitch};
  int rc = make_act_map(&p.a_map, vf, 4);
  if (rc) return rc;
  if (mode == 0) {
    View vt{static_cast<bf16*>(const_cast<void*>(t)), NB, H, W, n_total, t_pitch};
    rc = make_act_map(&p.b_map[0], vt, 4);
    if (rc) return rc;
  } else {
    View vt{static_cast<bf16*>(const_cast<void*>(t)), NB, 2 * H, 2 * W, n_total, t_pitch};
    for (int kk = 0; kk < 4; ++kk) {
      rc = make_act_map(&p.b_map[kk], vt, 4, 1, kk >> 1, kk & 1);
      if (rc) return rc;
    }
  }
  if (splits > 1)
    CRIMAC_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * p.taps * static_cast<size_t>(m_total) * n_total, st));
  CRIMAC_CHECK_CUDA(launch_wgrad_gemm(p, bn, st));
  CRIMAC_CHECK_CUDA(launch_wgrad_unpack(scratch, dw, m_total, n_total, p.taps, 0, st));
  return 0;
}

// 3x3-conv weight gradient with all nine taps per CTA (halo tile of dY, taps paired along M).
// dy: NHWC bf16 (NB,H,W,cout) pitch dy_pitch; x: NHWC bf16 (NB,H,W,cin) pitch x_pitch; scratch fp32 [9][cout][cin];
// dw: fp32 (cout,cin,3,3).
extern "C" int crimac_op_wgrad_halo(const void* dy, int dy_pitch, int cout, const void* x, int x_pitch, int cin, int NB,
                                    int H, int W, float* scratch, float* dw, int splits, void* stream) {
  CRIMAC_REQUIRE(cout % 64 == 0 && cin % 64 == 0, "channel counts must be multiples of 64");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  WgradHaloParams p{};
  p.Cs = cout;
  p.Cf = cin;
  p.NB = NB;
  p.H = H;
  p.W = W;
  p.tiles_x = (W + 15) / 16;
  p.tiles_y = (H + 3) / 4;
  p.k_tiles_total = NB * p.tiles_x * p.tiles_y;
  p.s_tiles = cout / 64;
  p.f_tiles = cin / 64;
  p.nf = (cin % 128 == 0) ? 128 : 64;   // as the network does (net_api.cu)
  if (splits <= 0) {
    const int tiles = p.s_tiles * p.f_tiles;
    splits = (2 * device_num_sms() + tiles - 1) / tiles;
    if (splits > p.k_tiles_total / 8) splits = p.k_tiles_total / 8;
    if (splits < 1) splits = 1;
  }
  p.splits = splits;
  p.dw = scratch;
  View vs{static_cast<bf16*>(const_cast<void*>(dy)), NB, H, W, cout, dy_pitch};
  View vf{static_cast<bf16*>(const_cast<void*>(x)), NB, H, W, cin, x_pitch};
  int rc = make_act_map(&p.s_map, vs, 6, 0, 0, 0, 18);
  if (rc) return rc;
  rc = make_act_map(&p.f_map, vf, 4);
  if (rc) return rc;
  if (splits > 1) CRIMAC_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float) * 9 * static_cast<size_t>(cout) * cin, st));
  CRIMAC_CHECK_CUDA(launch_wgrad_halo(p, st));
  CRIMAC_CHECK_CUDA(launch_wgrad_unpack(scratch, dw, cout, cin, 9, 0, st));
  return 0;
}

// ------------------------------------------------------------------------------------------------------------------
// Op-level entry points of the HBM-bound kernels (elementwise.cu, wgrad_gemm.cu's unpack, the weight packers): the
// parity tests pin each of them against torch fp32 autograd of the same operator on identical bf16-rounded inputs.
// Scratch: every call that needs per-block partial sums takes `scratch` of at least crimac_op_scratch_bytes() bytes.
extern "C" size_t crimac_op_scratch_bytes() {
  // BN backward: reduce_blocks() x 2 x C floats + 2 x C (c1, c2), C <= 1024; head: head_bwd_blocks() x (8*64+8) floats
  // + head_bwd_blocks() x 2 doubles; statistics rows of the conv kernels are NOT part of this
  return static_cast<size_t>(16) << 20;
}

static View mk_view(const void* p, int N, int H, int W, int C, int pitch) {
  return View{static_cast<bf16*>(const_cast<void*>(p)), N, H, W, C, pitch};
}

// Train-mode BatchNorm, forward half 1 (unet.py:78,81,121-122): partial rows [rows][2][C] (sum, sum of squares, as the
// conv kernels' EPI_STATS epilogue writes them) -> scale/shift for the apply pass, batch mean / inv-std for backward,
// running statistics (momentum, unbiased variance) and num_batches_tracked.
extern "C" int crimac_op_bn_finalize(const float* partials, int rows, int C, double count, const float* gamma,
                                     const float* beta, float* running_mean, float* running_var,
                                     int64_t* num_batches_tracked, float momentum, float eps, float* scale, float* shift,
                                     float* save_mean, float* save_invstd, void* stream) {
  CRIMAC_REQUIRE(partials && gamma && beta && scale && shift && save_mean && save_invstd, "NULL tensor");
  CRIMAC_REQUIRE(rows >= 1 && C >= 1 && count > 1.0, "rows, C >= 1 and more than one value per channel");
  CRIMAC_CHECK_CUDA(launch_bn_finalize(partials, rows, C, count, gamma, beta, running_mean, running_var,
                                       reinterpret_cast<long long*>(num_batches_tracked), momentum, eps, scale, shift,
                                       save_mean, save_invstd, static_cast<cudaStream_t>(stream)));
  return 0;
}

// Forward half 2: act = relu(raw*scale + shift) (+ 2x2 max-pool copy and its 2-bit-per-channel arg-max map).
// raw/act: NHWC bf16 (N,H,W,C) views with pixel pitches; pool (optional): (N,H/2,W/2,C); pool_arg (optional with pool):
// uint16 per (pooled pixel, 8-channel group).
extern "C" int crimac_op_bn_apply(const void* raw, int raw_pitch, int N, int H, int W, int C, const float* scale,
                                  const float* shift, void* act, int act_pitch, void* pool, int pool_pitch,
                                  uint16_t* pool_arg, void* stream) {
  CRIMAC_REQUIRE(raw && act && scale && shift, "NULL tensor");
  CRIMAC_REQUIRE(C % 8 == 0 && raw_pitch % 8 == 0 && act_pitch % 8 == 0, "channels / pitches must be multiples of 8");
  CRIMAC_REQUIRE(pool == nullptr || (H % 2 == 0 && W % 2 == 0 && pool_pitch % 8 == 0), "pooling needs even H, W");
  View vp = mk_view(pool, N, H / 2, W / 2, C, pool_pitch);
  CRIMAC_CHECK_CUDA(launch_bn_apply(mk_view(raw, N, H, W, C, raw_pitch), scale, shift, mk_view(act, N, H, W, C, act_pitch),
                                    vp, pool ? pool_arg : nullptr, static_cast<cudaStream_t>(stream)));
  return 0;
}

// BatchNorm + ReLU backward (autograd of unet.py:76-83,135-136): dact = gradient w.r.t. the post-ReLU activation,
// raw = the conv output saved in forward; writes draw (gradient w.r.t. the conv output, bf16), dgamma, dbeta and the
// conv-bias gradient (exactly zero under train-mode BatchNorm).  gscale (optional): one device float multiplied into dact.
extern "C" int crimac_op_bn_bwd(const void* dact, int dact_pitch, const void* raw, int raw_pitch, int N, int H, int W,
                                int C, const float* scale, const float* shift, const float* mean, const float* invstd,
                                void* draw, int draw_pitch, float* dgamma, float* dbeta, float* dbias,
                                const float* gscale, void* scratch, void* stream) {
  CRIMAC_REQUIRE(dact && raw && draw && scale && shift && mean && invstd && dgamma && dbeta && scratch, "NULL tensor");
  CRIMAC_REQUIRE(C % 8 == 0 && C <= 1024, "C must be a multiple of 8, at most 1024");
  float* partials = static_cast<float*>(scratch);
  float* c1c2 = partials + static_cast<size_t>(reduce_blocks()) * 2 * C;
  CRIMAC_CHECK_CUDA(launch_bn_bwd(mk_view(dact, N, H, W, C, dact_pitch), mk_view(raw, N, H, W, C, raw_pitch), scale, shift,
                                  mean, invstd, mk_view(draw, N, H, W, C, draw_pitch), dgamma, dbeta, dbias, 0, partials,
                                  c1c2, gscale, static_cast<cudaStream_t>(stream)));
  return 0;
}

// The same BatchNorm + ReLU backward with the incoming gradient formed on the fly as dskip + unpool(dpool) through the
// arg-max map of crimac_op_bn_apply (max-pool backward + skip add fused in; what the encoder's second convs run).
extern "C" int crimac_op_bn_bwd_pool(const uint16_t* pool_arg, const void* dpool, int dpool_pitch, const void* dskip,
                                     int dskip_pitch, const void* raw, int raw_pitch, int N, int H, int W, int C,
                                     const float* scale, const float* shift, const float* mean, const float* invstd,
                                     void* draw, int draw_pitch, float* dgamma, float* dbeta, float* dbias, void* scratch,
                                     void* stream) {
  CRIMAC_REQUIRE(pool_arg && dpool && dskip && raw && draw && scale && shift && mean && invstd && dgamma && dbeta && scratch, "NULL tensor");
  CRIMAC_REQUIRE(C % 8 == 0 && C <= 1024 && H % 2 == 0 && W % 2 == 0, "C multiple of 8 (<= 1024), even H and W");
  float* partials = static_cast<float*>(scratch);
  float* c1c2 = partials + static_cast<size_t>(reduce_blocks()) * 2 * C;
  CRIMAC_CHECK_CUDA(launch_bn_bwd(View{}, mk_view(raw, N, H, W, C, raw_pitch), scale, shift, mean, invstd,
                                  mk_view(draw, N, H, W, C, draw_pitch), dgamma, dbeta, dbias, 0, partials, c1c2, nullptr,
                                  static_cast<cudaStream_t>(stream), pool_arg, mk_view(dpool, N, H / 2, W / 2, C, dpool_pitch),
                                  mk_view(dskip, N, H, W, C, dskip_pitch)));
  return 0;
}

// Max-pool backward + skip-gradient add (autograd of unet.py:86,92,132): dact[2x2 window] = dskip[window] + (first
// maximal element ? dpool : 0), arg-max from the map crimac_op_bn_apply wrote.  dskip may be NULL (no skip branch).
extern "C" int crimac_op_pool_bwd_add(const uint16_t* pool_arg, const void* dpool, int dpool_pitch, const void* dskip,
                                      int dskip_pitch, void* dact, int dact_pitch, int N, int H, int W, int C,
                                      void* stream) {
  CRIMAC_REQUIRE(pool_arg && dpool && dact, "NULL tensor");
  CRIMAC_REQUIRE(C % 8 == 0 && H % 2 == 0 && W % 2 == 0, "C multiple of 8, even H and W");
  CRIMAC_CHECK_CUDA(launch_pool_bwd_add(pool_arg, mk_view(dpool, N, H / 2, W / 2, C, dpool_pitch),
                                        mk_view(dskip, N, H, W, C, dskip_pitch), mk_view(dact, N, H, W, C, dact_pitch),
                                        static_cast<cudaStream_t>(stream)));
  return 0;
}

// 1x1 head + class-weighted cross-entropy + head backward in ONE pass (unet.py:342, pipeline.py:135-138,176-177).
// act: NHWC bf16 (N,H,W,64).  Outputs: dact (N,H,W,64) bf16 = UNNORMALISED gradient (multiply by out3[1]), dw (ncls,64),
// db (ncls) normalised, out3 = {loss, 1/sum_w, sum_w}.
extern "C" int crimac_op_head_ce(const void* act, int act_pitch, int N, int H, int W, const float* hw, const float* hb,
                                 int ncls, const int64_t* labels, const float* cw, int64_t ignore_index, void* dact,
                                 int dact_pitch, float* dw, float* db, float* out3, void* scratch, void* stream) {
  CRIMAC_REQUIRE(act && hw && hb && labels && cw && dact && dw && db && out3 && scratch, "NULL tensor");
  CRIMAC_REQUIRE(ncls >= 1 && ncls <= CRIMAC_MAX_CLASSES, "n_classes must be 1..8");
  float* partials = static_cast<float*>(scratch);  // head_bwd_blocks() x (ncls*64 + ncls) floats < 8 MB
  double* lp = reinterpret_cast<double*>(static_cast<uint8_t*>(scratch) + (static_cast<size_t>(8) << 20));
  CRIMAC_CHECK_CUDA(launch_head_ce_fused(mk_view(act, N, H, W, 64, act_pitch), hw, hb, ncls,
                                         reinterpret_cast<const long long*>(labels), cw, ignore_index,
                                         mk_view(dact, N, H, W, 64, dact_pitch), partials, lp, dw, db, out3,
                                         static_cast<cudaStream_t>(stream)));
  return 0;
}

// The same three steps as separate kernels (the autograd path with a user-supplied loss): head forward -> logits NCHW
// fp32; head backward from dlogits (x optional *gscale) -> dact, dw, db.
extern "C" int crimac_op_head_fwd(const void* act, int act_pitch, int N, int H, int W, const float* hw, const float* hb,
                                  int ncls, float* logits, void* stream) {
  CRIMAC_REQUIRE(act && hw && hb && logits && ncls >= 1 && ncls <= CRIMAC_MAX_CLASSES, "bad argument");
  CRIMAC_CHECK_CUDA(launch_head_fwd(mk_view(act, N, H, W, 64, act_pitch), hw, hb, ncls, logits,
                                    static_cast<cudaStream_t>(stream)));
  return 0;
}
extern "C" int crimac_op_head_bwd(const float* dlogits, const float* gscale, const void* act, int act_pitch, int N, int H,
                                  int W, const float* hw, int ncls, void* dact, int dact_pitch, float* dw, float* db,
                                  void* scratch, void* stream) {
  CRIMAC_REQUIRE(dlogits && act && hw && dact && dw && db && scratch && ncls >= 1 && ncls <= CRIMAC_MAX_CLASSES, "bad argument");
  CRIMAC_CHECK_CUDA(launch_head_bwd(dlogits, gscale, mk_view(act, N, H, W, 64, act_pitch), hw, ncls,
                                    mk_view(dact, N, H, W, 64, dact_pitch), static_cast<float*>(scratch), dw, db, 0,
                                    static_cast<cudaStream_t>(stream)));
  return 0;
}

// Per-channel sum over pixels of an NHWC bf16 view (ConvTranspose2d bias gradient).
extern "C" int crimac_op_colsum(const void* v, int pitch, int N, int H, int W, int C, float* out, void* scratch,
                                void* stream) {
  CRIMAC_REQUIRE(v && out && scratch && C % 8 == 0 && C <= 1024, "bad argument");
  CRIMAC_CHECK_CUDA(launch_view_colsum(mk_view(v, N, H, W, C, pitch), static_cast<float*>(scratch), out, 0,
                                       static_cast<cudaStream_t>(stream)));
  return 0;
}

// fp32 parameters -> bf16 GEMM operands.  kind 0: conv (Cout,Cin,3,3) -> [Cout][tap][Cin]; kind 1: ConvTranspose
// (Cin,Cout,2,2) -> [(ky*2+kx)*Cout + co][Cin].
extern "C" int crimac_op_pack(int kind, const float* w, int cout, int cin, void* out, void* stream) {
  CRIMAC_REQUIRE(w && out && (kind == 0 || kind == 1), "bad argument");
  CRIMAC_REQUIRE(cin % 64 == 0 && cout % 8 == 0, "Cin must be a multiple of 64, Cout of 8");
  PackTable t{};
  t.n = 1;
  t.e[0] = PackEntry{w, static_cast<bf16*>(out), cout, cin, 0};
  CRIMAC_CHECK_CUDA(kind == 0 ? launch_pack_conv3x3_all(t, static_cast<cudaStream_t>(stream))
                              : launch_pack_convt_all(t, static_cast<cudaStream_t>(stream)));
  return 0;
}

// The end-of-backward un-pack of up to 24 layers in one launch: scratch_i [taps_i][mn_i] -> dw_i [mn_i][taps_i], the
// scratch is left zeroed.  Arrays are HOST arrays of n entries.
extern "C" int crimac_op_wgrad_unpack_all(int n, float* const* scratch, float* const* dw, const int64_t* mn,
                                          const int* taps, void* stream) {
  CRIMAC_REQUIRE(n >= 1 && n <= 24 && scratch && dw && mn && taps, "bad argument");
  UnpackTable t{};
  t.n = n;
  for (int i = 0; i < n; ++i) {
    CRIMAC_REQUIRE(taps[i] >= 1 && taps[i] <= 9 && mn[i] >= 1, "taps must be 1..9");
    t.e[i] = UnpackEntry{scratch[i], dw[i], static_cast<long>(mn[i]), taps[i], 0};
  }
  CRIMAC_CHECK_CUDA(launch_wgrad_unpack_all(t, static_cast<cudaStream_t>(stream)));
  return 0;
}

// Bilinear 2x up-sampling of up_mode "upsample" (unet.py:50-56; align_corners=False) on NHWC bf16 views.  backward == 0:
// lo (N,H,W,C) -> hi (N,2H,2W,C); backward != 0: the adjoint, hi = gradient at full resolution -> lo.
extern "C" int crimac_op_upsample2x(void* lo, int lo_pitch, void* hi, int hi_pitch, int N, int H, int W, int C,
                                    int backward, void* stream) {
  CRIMAC_REQUIRE(lo && hi && C % 8 == 0 && lo_pitch % 8 == 0 && hi_pitch % 8 == 0, "bad argument");
  View vl = mk_view(lo, N, H, W, C, lo_pitch), vh = mk_view(hi, N, 2 * H, 2 * W, C, hi_pitch);
  CRIMAC_CHECK_CUDA(backward ? launch_upsample2x_bwd(vh, vl, static_cast<cudaStream_t>(stream))
                             : launch_upsample2x(vl, vh, static_cast<cudaStream_t>(stream)));
  return 0;
}

// Op-level C-ABI entry points (crimac_op_*): one kernel each, tensor maps encoded per call.  The network-level
// entry points in net_api.cu drive the same launchers with maps cached in the context; these exist so the parity
// tests can pin every kernel against the oracle in isolation.
#include "host_util.h"

static int pick_block_n(int n_total, int requested) {
  if (requested == 64 || requested == 128 || requested == 256) return (n_total % requested == 0) ? requested : 0;
  if (n_total % 256 == 0) return 256;
  if (n_total % 128 == 0) return 128;
  if (n_total % 64 == 0) return 64;
  return 0;
}

// Generic implicit GEMM.  mode: 0 = 3x3 conv (taps 9, halo main loop; 3 = same through the 9-box main loop), 1 = 1x1 / ConvTranspose forward (taps 1; convt_cout > 0 turns on
// the 2x upsampling scatter), 2 = ConvTranspose backward-data (taps 4: x is the (2H x 2W) gradient, sub-sampled).
// x: NHWC bf16 view (NB,H,W,cin) with pixel pitch x_pitch (for mode 2: dims of x are 2H x 2W).
// w: packed bf16 [n_total][taps*cin].   out: NHWC bf16 with pitch out_pitch.
// stats != NULL selects the train-mode epilogue (raw + per-tile channel sums); head_w != NULL the fused head.
extern "C" int crimac_op_igemm(int mode, const void* x, int NB, int H, int W, int cin, int x_pitch, const void* w,
                               int n_total, const float* scale, const float* shift, int relu, void* out,
                               int out_pitch, int convt_cout, void* pool_out, int pool_pitch, float* stats,
                               const float* head_w, const float* head_b, float* head_out, int n_classes,
                               int head_softmax, int block_n, void* stream) {
  CRIMAC_REQUIRE(mode >= 0 && mode <= 5, "mode");
  // modes 4 / 5: backward-data of a 3x3 conv / ConvTranspose reading the FORWARD-packed weights as an MN-major operand
  // (w = [Cout][9*Cin] resp. [(kk,co)][Cin] of the forward layer; n_total = the forward layer's Cin)
  const bool b_mn = (mode == 4 || mode == 5);
  if (mode == 4) mode = 0;
  if (mode == 5) mode = 2;
  const bool halo = (mode == 0);  // mode 3 = 3x3 conv through the plain 9-box main loop (kept for A/B measurements)
  if (mode == 3) mode = 0;
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

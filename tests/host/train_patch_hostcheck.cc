// Host self-check of the training-sample kernels (TEST INFRASTRUCTURE, g++ only, never linked into the product):
// runs the SAME per-thread bodies the CUDA kernels in csrc/train_patches.cu run (csrc/train_patch_core.h), block by
// block and phase by phase, so that the index arithmetic, the bit-row morphology and the label logic are checked
// against the oracle and the reference-made fixtures without a GPU.  Only the warp- and cluster-collective parts of
// the label kernel (ballot, shuffle reduction, the exchange of the bounding-box partials) are restated with loops.
#include "../../crimac-classifiers-unet_b200/csrc/train_patch_core.h"
#include <vector>

extern "C" int tp_host_train_patches(const float* sv, const float* labels, int F, int P, int R, const int32_t* centres,
                                     const uint8_t* flags, const float* noise_mult, uint64_t noise_seed, int n, int ph,
                                     int pw, int thr_freq, double thr_lo, double thr_hi, int scaled, int border_zero,
                                     float* x_out, int64_t* labels_out) {
  if (ph % 32 || pw % 32 || ph != pw) return 1;
  tp::GatherParams p;
  p.sv = sv; p.labels = labels; p.centres = centres; p.flags = flags; p.noise = noise_mult;
  p.seed = noise_seed;
  p.F = F; p.P = P; p.R = R; p.n = n; p.ph = ph; p.pw = pw;
  p.thr_freq = thr_freq; p.thr_lo = thr_lo; p.thr_hi = thr_hi; p.scaled = scaled;
  p.x = x_out; p.lab = reinterpret_cast<long long*>(labels_out);
  // train_gather_kernel<<<((ph/32)*(pw/32), F+1, n), (32, 8)>>>
  float tile[32][33], tile_lab[32][33];
  for (int z = 0; z < n; ++z)
    for (int y = 0; y <= F; ++y)
      for (int x = 0; x < (ph / 32) * (pw / 32); ++x) {
        tp::TileCoord blk;
        blk.tile = x; blk.chan = y; blk.crop = z;
        for (int ty = 0; ty < 8; ++ty)
          for (int tx = 0; tx < 32; ++tx) tp::gather_load(p, blk, tx, ty, tile, tile_lab);
        for (int ty = 0; ty < 8; ++ty)
          for (int tx = 0; tx < 32; ++tx) tp::gather_store(p, blk, tx, ty, flags[z], tile, tile_lab);
      }
  // train_labels_kernel<<<n * 8 (clusters of 8), 512>>>: the three cluster-barrier-separated stages, band by band
  const int wpr = pw >> 5, npx = ph * pw, rows = ph / tp::kBands;
  for (int b = 0; b < n; ++b) {
    long long* L = p.lab + static_cast<long>(b) * npx;
    std::vector<std::vector<uint32_t>> T(tp::kBands), D(tp::kBands), E(tp::kBands);
    int box[tp::kBands][4];
    for (int band = 0; band < tp::kBands; ++band) {  // stage 1: threshold mask with halo rows + partial boxes
      const int r0 = band * rows;
      T[band].assign((rows + 12) * wpr, 0u);
      D[band].assign((rows + 6) * wpr, 0u);
      E[band].assign(rows * wpr, 0u);
      int* bx = box[band];
      bx[0] = ph; bx[1] = -1; bx[2] = pw; bx[3] = -1;
      for (int wi = 0; wi < (rows + 12) * wpr; ++wi)
        for (int lane = 0; lane < 32; ++lane) {
          const int ly = wi / wpr, w = wi - ly * wpr;
          const int y = r0 - 6 + ly;
          int code = tp::L_BOUNDARY;
          if (y >= 0 && y < ph) code = static_cast<int>(L[static_cast<long>(y) * pw + 32 * w + lane]);
          if (code & tp::kThresholdBit) T[band][wi] |= 1u << lane;
          if (y >= r0 && y < r0 + rows && (code & 7) != tp::L_BOUNDARY) {
            const int xx = 32 * w + lane;
            if (y < bx[0]) bx[0] = y;
            if (y > bx[1]) bx[1] = y;
            if (xx < bx[2]) bx[2] = xx;
            if (xx > bx[3]) bx[3] = xx;
          }
        }
    }
    tp::BBox bb;
    bb.y0 = ph; bb.y1 = -1; bb.x0 = pw; bb.x1 = -1;
    for (int r = 0; r < tp::kBands; ++r) {
      if (box[r][0] < bb.y0) bb.y0 = box[r][0];
      if (box[r][1] > bb.y1) bb.y1 = box[r][1];
      if (box[r][2] < bb.x0) bb.x0 = box[r][2];
      if (box[r][3] > bb.x1) bb.x1 = box[r][3];
    }
    bb.y1 += 1;
    bb.x1 += 1;
    for (int band = 0; band < tp::kBands; ++band) {  // stage 2
      const int r0 = band * rows;
      for (int i = 0; i < (rows + 6) * wpr; ++i) tp::labels_dilate_band(T[band].data(), D[band].data(), bb, wpr, r0, rows, i);
      for (int i = 0; i < rows * wpr; ++i) tp::labels_erode_band(D[band].data(), E[band].data(), wpr, rows, i);
    }
    float* x_crop = x_out + static_cast<long>(b) * F * npx;
    for (int band = 0; band < tp::kBands; ++band)    // stage 3
      for (int i = 0; i < rows * pw; ++i)
        tp::labels_finish_band(L, x_crop, E[band].data(), F, npx, band * rows * pw, border_zero, i);
  }
  return 0;
}

extern "C" void tp_host_noise_field(uint64_t seed, int chan, int ph, int pw, float* out) {
  for (int py = 0; py < ph; ++py)
    for (int px = 0; px < pw; ++px) out[py * pw + px] = tp::noise_multiplier(seed, chan, py, px);
}

extern "C" void tp_host_closing(const uint8_t* mask, int H, int W, int y0, int y1, int x0, int x1, uint8_t* out) {
  const int wpr = W >> 5, rows = H / tp::kBands;
  tp::BBox bb;
  bb.y0 = y0; bb.y1 = y1; bb.x0 = x0; bb.x1 = x1;
  for (int band = 0; band < tp::kBands; ++band) {
    const int r0 = band * rows;
    std::vector<uint32_t> T((rows + 12) * wpr, 0u), D((rows + 6) * wpr), E(rows * wpr);
    for (int ly = 0; ly < rows + 12; ++ly) {
      const int y = r0 - 6 + ly;
      if (y < 0 || y >= H) continue;
      for (int xx = 0; xx < W; ++xx)
        if (mask[y * W + xx]) T[ly * wpr + (xx >> 5)] |= 1u << (xx & 31);
    }
    for (int i = 0; i < (rows + 6) * wpr; ++i) tp::labels_dilate_band(T.data(), D.data(), bb, wpr, r0, rows, i);
    for (int i = 0; i < rows * wpr; ++i) tp::labels_erode_band(D.data(), E.data(), wpr, rows, i);
    for (int i = 0; i < rows * W; ++i) out[r0 * W + i] = (E[i >> 5] >> (i & 31)) & 1u;
  }
}

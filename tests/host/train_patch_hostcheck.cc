// Host self-check of the training-sample kernels (TEST INFRASTRUCTURE, g++ only, never linked into the product):
// runs the SAME per-thread bodies the CUDA kernels in csrc/train_patches.cu run (csrc/train_patch_core.h), block by
// block and phase by phase, so that the index arithmetic, the bit-row morphology and the label logic are checked
// against the oracle and the reference-made fixtures without a GPU.  Only the warp-collective part of the label
// kernel's first phase (ballot / shuffle reduction) is restated with plain loops.
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
          for (int tx = 0; tx < 32; ++tx) tp::gather_store(p, blk, tx, ty, tile, tile_lab);
      }
  // train_labels_kernel<<<n, 256, 3 * ph * pw / 8>>>
  const int wpr = pw >> 5, nwords = ph * wpr, npx = ph * pw;
  std::vector<uint32_t> T(nwords), D(nwords), E(nwords);
  for (int b = 0; b < n; ++b) {
    long long* L = p.lab + static_cast<long>(b) * npx;
    int ymin = ph, ymax = -1, xmin = pw, xmax = -1;
    for (int i = 0; i < nwords; ++i) T[i] = 0u;
    for (int i = 0; i < npx; ++i) {
      const int code = static_cast<int>(L[i]);
      if (code & tp::kThresholdBit) T[i >> 5] |= 1u << (i & 31);
      if ((code & 7) != tp::L_BOUNDARY) {
        const int yy = i / pw, xx = i - yy * pw;
        if (yy < ymin) ymin = yy;
        if (yy > ymax) ymax = yy;
        if (xx < xmin) xmin = xx;
        if (xx > xmax) xmax = xx;
      }
    }
    tp::BBox bb;
    bb.y0 = ymin; bb.y1 = ymax + 1; bb.x0 = xmin; bb.x1 = xmax + 1;
    for (int i = 0; i < nwords; ++i) tp::labels_dilate(T.data(), D.data(), bb, ph, wpr, i);
    for (int i = 0; i < nwords; ++i) tp::labels_erode(D.data(), E.data(), ph, wpr, i);
    float* x_crop = x_out + static_cast<long>(b) * F * npx;
    for (int i = 0; i < npx; ++i) tp::labels_finish(L, x_crop, E.data(), F, npx, border_zero, i);
  }
  return 0;
}

extern "C" double tp_host_noise_multiplier(uint64_t seed, uint64_t index) { return tp::noise_multiplier(seed, index); }

extern "C" void tp_host_closing(const uint8_t* mask, int H, int W, int y0, int y1, int x0, int x1, uint8_t* out) {
  const int wpr = W >> 5, nwords = H * wpr;
  std::vector<uint32_t> T(nwords, 0u), D(nwords), E(nwords);
  for (int i = 0; i < H * W; ++i)
    if (mask[i]) T[i >> 5] |= 1u << (i & 31);
  tp::BBox bb;
  bb.y0 = y0; bb.y1 = y1; bb.x0 = x0; bb.x1 = x1;
  for (int i = 0; i < nwords; ++i) tp::labels_dilate(T.data(), D.data(), bb, H, wpr, i);
  for (int i = 0; i < nwords; ++i) tp::labels_erode(D.data(), E.data(), H, wpr, i);
  for (int i = 0; i < H * W; ++i) out[i] = (E[i >> 5] >> (i & 31)) & 1u;
}

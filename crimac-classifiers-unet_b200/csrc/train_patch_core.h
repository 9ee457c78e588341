// Arithmetic of the training-sample path shared by the CUDA kernels in train_patches.cu and by
// the host-side self-check library (train_patch_hostcheck.cc, g++ only): the same inline functions run on both sides,
// so the label logic, the bit-mask morphology and the counter-based noise generator are testable without a GPU.
//
// Reference steps restated here (paths relative to the reference's crimac_unet/):
//   batch/dataset.py:358-407                     get_crop_zarr: out-of-data -> 0 / LABEL_BOUNDARY_VAL, nan_to_num
//   batch/data_augmentation/add_noise.py:22-40   5 % of the samples scaled by U(1,10) or U(0,1)
//   batch/label_transforms/refine_label_boundary.py:38-104   7x7 disc closing of the thresholded school mask
//   batch/label_transforms/convert_label_indexing.py:24-35   0 -> 0, 27 -> 1, 1 -> 2, everything else -> -100
//   batch/data_transforms/db_with_limits.py:20-38            10 log10(v + 1e-10) clipped to [-75, 0] (optionally scaled)
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define TP_HD __host__ __device__ __forceinline__
#else
#define TP_HD inline
#endif

namespace tp {

// ---- label codes ---------------------------------------------------------------------------------------------
// get_crop_zarr hands refine_label_boundary a float label image; everything downstream only distinguishes:
enum LabelCode : int {
  L_BACKGROUND = 0,   // raw 0                      -> BACKGROUND (0)
  L_OTHER = 1,        // raw 1                      -> OTHER (2) unless the closing rejects the pixel
  L_SANDEEL = 2,      // raw 27                     -> SANDEEL (1) unless the closing rejects the pixel
  L_POSITIVE = 3,     // any other raw > 0          -> -100, but takes part in the school mask (labels > 0)
  L_NONPOSITIVE = 4,  // raw < 0 and != -100, or fractional non-positive: -> -100, inside the bounding box
  L_BOUNDARY = 5      // raw -100 / NaN / outside the data: -> -100, outside the bounding box
};
constexpr int kThresholdBit = 8;  // OR-ed into the code by the gather kernel: sample passes lo < v < hi on thr_freq

TP_HD int label_code(float raw) {
  if (isnan(raw) || raw == -100.f) return L_BOUNDARY;  // np.nan_to_num(labels.T, nan=LABEL_BOUNDARY_VAL), dataset.py:405
  if (raw == 0.f) return L_BACKGROUND;
  if (raw == 1.f) return L_OTHER;
  if (raw == 27.f) return L_SANDEEL;
  return raw > 0.f ? L_POSITIVE : L_NONPOSITIVE;
}

// refine_label_boundary.py:96-101 followed by convert_label_indexing.py:29-33
TP_HD long final_label(int code, bool closed) {
  if (code == L_BACKGROUND) return 0;
  if (code == L_OTHER) return closed ? 2 : -100;    // LABEL_REFINE_BOUNDARY_VAL (-30) is not in {0, 1, 27} -> -100
  if (code == L_SANDEEL) return closed ? 1 : -100;
  return -100;
}

// np.nan_to_num on the fp32 slice (dataset.py:404): NaN -> 0, +-inf -> +-FLT_MAX
TP_HD float nan_to_num_f32(float s) {
  if (isnan(s)) return 0.f;
  if (isinf(s)) return s > 0.f ? 3.402823466e+38f : -3.402823466e+38f;
  return s;
}

// db_with_limits / db_with_limits_scaled in double like the reference (out_data is float64, dataset.py:361); the
// explicit comparisons keep a NaN (negative sv) a NaN exactly as numpy's masked assignments do.
TP_HD float sv_to_db(double v, int scaled) {
  double d = 10.0 * log10(v + 1e-10);
  if (d > 0.0) d = 0.0;
  if (d < -75.0) d = -75.0;
  if (scaled) d = 1.0 + d / 75.0;
  return static_cast<float>(d);
}

// ---- counter-based noise (Philox4x32-10) --------------------------------------------------------------------
// The reference draws from numpy's global generator (add_noise.py:25-38); a device kernel cannot replay that stream,
// so the product draws the same distribution from a counter-based generator keyed by (seed, sample index): every
// sample's multiplier is a pure function of its index, which lets the label pass recompute the threshold channel.
TP_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4]) {
  for (int r = 0; r < 10; ++r) {
    const uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0;
    const uint64_t p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    const uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n1 = static_cast<uint32_t>(p1);
    const uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    const uint32_t n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

TP_HD double u01(uint32_t r) { return (static_cast<double>(r) + 0.5) * (1.0 / 4294967296.0); }

// add_noise.py:28-38: (1 - change) + change * (increase * U(1,10) + (1 - increase) * U(0,1)), P(change) = 0.05,
// P(increase) = 0.5, the three draws independent.
TP_HD double noise_multiplier(uint64_t seed, uint64_t index) {
  uint32_t r[4];
  philox4x32_10(static_cast<uint32_t>(index), static_cast<uint32_t>(index >> 32), 0x43524d43u, 0u,
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  if (u01(r[0]) >= 0.05) return 1.0;
  const double u = u01(r[2]);
  return u01(r[1]) < 0.5 ? 1.0 + 9.0 * u : u;
}

// ---- 7x7 disc closing on bit rows ----------------------------------------------------------------------------
// Masks are stored one bit per sample: word (y, w) holds pings 32w .. 32w+31 of range row y (bit i = ping 32w+i).
// The structuring element of refine_label_boundary.py:53-61 is symmetric; row dy of it spans |dx| <= kHalfWidth[dy+3].
struct BBox {
  int y0, y1, x0, x1;  // half-open [y0,y1) x [x0,x1); empty when y1 <= y0
};

TP_HD int disc_half_width(int dy) {
  const int a = dy < 0 ? -dy : dy;
  return a == 3 ? 1 : (a == 2 ? 2 : 3);
}

TP_HD uint32_t mask_word(const uint32_t* m, int wpr, int H, int y, int w) {
  return (y < 0 || y >= H || w < 0 || w >= wpr) ? 0u : m[y * wpr + w];
}

// bit i of the result = sample (32w + i + dx) of row y, 0 outside the image; |dx| <= 3
TP_HD uint32_t shifted_word(const uint32_t* m, int wpr, int H, int y, int w, int dx) {
  const uint32_t cur = mask_word(m, wpr, H, y, w);
  if (dx == 0) return cur;
  if (dx > 0) return (cur >> dx) | (mask_word(m, wpr, H, y, w + 1) << (32 - dx));
  const int s = -dx;
  return (cur << s) | (mask_word(m, wpr, H, y, w - 1) >> (32 - s));
}

// scipy.ndimage.binary_dilation(structure=disc, border_value=0): OR over the structuring element
TP_HD uint32_t dilate_word(const uint32_t* m, int wpr, int H, int y, int w) {
  uint32_t acc = 0u;
  for (int dy = -3; dy <= 3; ++dy) {
    const int hw = disc_half_width(dy);
    for (int dx = -hw; dx <= hw; ++dx) acc |= shifted_word(m, wpr, H, y + dy, w, dx);
  }
  return acc;
}

// scipy.ndimage.binary_erosion(structure=disc, border_value=0): AND over the structuring element, outside = 0
TP_HD uint32_t erode_word(const uint32_t* m, int wpr, int H, int y, int w) {
  uint32_t acc = 0xFFFFFFFFu;
  for (int dy = -3; dy <= 3; ++dy) {
    const int hw = disc_half_width(dy);
    for (int dx = -hw; dx <= hw; ++dx) acc &= shifted_word(m, wpr, H, y + dy, w, dx);
  }
  return acc;
}

// bits of word (y, w) that lie inside the bounding box
TP_HD uint32_t bbox_word(const BBox& b, int y, int w) {
  if (y < b.y0 || y >= b.y1) return 0u;
  const int lo = b.x0 - 32 * w, hi = b.x1 - 32 * w;  // bit range [lo, hi)
  if (hi <= 0 || lo >= 32) return 0u;
  const uint32_t m_lo = lo <= 0 ? 0xFFFFFFFFu : (0xFFFFFFFFu << lo);
  const uint32_t m_hi = hi >= 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u);
  return m_lo & m_hi;
}

// ---- the gather kernel's two phases (one 32 x 32 tile of one crop and one channel per thread block) -----------
#if defined(__CUDA_ARCH__)
#define TP_LDG(ptr) __ldg(ptr)
#else
#define TP_LDG(ptr) (*(ptr))
#endif

struct GatherParams {
  const float* sv;        // (F, P, R)
  const float* labels;    // (P, R)
  const int* centres;     // (n, 2) = (range, ping)
  const uint8_t* flags;   // (n) bit0 = add noise, bit1 = flip the ping axis
  const float* noise;     // optional (n, F, ph, pw) explicit multipliers in un-flipped crop coordinates
  unsigned long long seed;
  int F, P, R, n, ph, pw;
  int thr_freq;
  double thr_lo, thr_hi;
  int scaled;
  float* x;               // (n, F, ph, pw)
  long long* lab;         // (n, ph, pw)
};

struct TileCoord {  // blockIdx of the gather kernel: x = tile of the crop, y = channel (F = labels), z = crop
  int tile, chan, crop;
};

TP_HD double sample_multiplier(const GatherParams& p, int b, int f, int py, int px, bool noisy) {
  if (!noisy) return 1.0;
  const long idx = ((static_cast<long>(b) * p.F + f) * p.ph + py) * p.pw + px;
  return p.noise != nullptr ? static_cast<double>(TP_LDG(p.noise + idx))
                            : noise_multiplier(p.seed, static_cast<uint64_t>(idx));
}

// Load phase: thread (tx, ty) of the (32, 8) block; tx runs along RANGE, which is contiguous in the store.
// tile[i][j]: i = ping offset inside the tile, j = range offset.
TP_HD void gather_load(const GatherParams& p, const TileCoord& blk, int tx, int ty, float (*tile)[33],
                       float (*tile_lab)[33]) {
  const bool is_label = blk.chan == p.F;
  const int tiles_x = p.pw >> 5;
  const int ty0 = (blk.tile / tiles_x) << 5, tx0 = (blk.tile % tiles_x) << 5;
  const int cy = p.centres[2 * blk.crop], cx = p.centres[2 * blk.crop + 1];
  const int Y = cy - p.ph / 2 + 1 + ty0 + tx;  // utils/np.py:378-380
  const int src_f = is_label ? p.thr_freq : blk.chan;
  for (int k = 0; k < 4; ++k) {
    const int i = ty + 8 * k;
    const int X = cx - p.pw / 2 + 1 + tx0 + i;
    const bool in = Y >= 0 && Y < p.R && X >= 0 && X < p.P;
    float s = 0.f;  // boundary_val_data, dataset.py:360
    if (in) s = nan_to_num_f32(TP_LDG(p.sv + (static_cast<long>(src_f) * p.P + X) * p.R + Y));
    tile[i][tx] = s;
    if (is_label) tile_lab[i][tx] = in ? TP_LDG(p.labels + static_cast<long>(X) * p.R + Y) : -100.f;
  }
}

// Store phase: tx runs along PING, which is contiguous in the network input.
TP_HD void gather_store(const GatherParams& p, const TileCoord& blk, int tx, int ty, const float (*tile)[33],
                        const float (*tile_lab)[33]) {
  const bool is_label = blk.chan == p.F;
  const int tiles_x = p.pw >> 5;
  const int ty0 = (blk.tile / tiles_x) << 5, tx0 = (blk.tile % tiles_x) << 5;
  const int b = blk.crop;
  const uint8_t fl = p.flags[b];
  const bool noisy = (fl & 1) != 0, flip = (fl & 2) != 0;
  const int src_f = is_label ? p.thr_freq : blk.chan;
  const int px = tx0 + tx;
  const int ox = flip ? p.pw - 1 - px : px;  // flip_x_axis.py:22-25 (after the noise)
  for (int k = 0; k < 4; ++k) {
    const int j = ty + 8 * k;
    const int py = ty0 + j;
    const double v = static_cast<double>(tile[tx][j]) * sample_multiplier(p, b, src_f, py, px, noisy);
    if (!is_label) {
      p.x[((static_cast<long>(b) * p.F + blk.chan) * p.ph + py) * p.pw + ox] = sv_to_db(v, p.scaled);
    } else {
      int code = label_code(tile_lab[tx][j]);
      // refine_label_boundary.py:88-89: (labels > 0) & (data > lo) & (data < hi) on the threshold frequency
      const bool positive = code == L_OTHER || code == L_SANDEEL || code == L_POSITIVE;
      if (positive && v > p.thr_lo && v < p.thr_hi) code |= kThresholdBit;
      p.lab[(static_cast<long>(b) * p.ph + py) * p.pw + ox] = code;
    }
  }
}

// ---- the label kernel's word / sample phases (one crop per thread block; T, D, E = bit masks of ph*pw/32 words) --
TP_HD void labels_dilate(const uint32_t* T, uint32_t* D, const BBox& bb, int ph, int wpr, int i) {
  const int y = i / wpr, w = i - y * wpr;
  // binary_closing on the bounding-box crop (refine_label_boundary.py:91) = dilation, cut to the box, erosion with
  // everything outside the box counting as 0 (border_value = 0 in both passes)
  D[i] = dilate_word(T, wpr, ph, y, w) & bbox_word(bb, y, w);
}

TP_HD void labels_erode(const uint32_t* D, uint32_t* E, int ph, int wpr, int i) {
  const int y = i / wpr, w = i - y * wpr;
  E[i] = erode_word(D, wpr, ph, y, w);
}

TP_HD void labels_finish(long long* L, float* x_crop, const uint32_t* E, int F, int npx, int border_zero, int i) {
  const int code = static_cast<int>(L[i]) & 7;
  const bool closed = ((E[i >> 5] >> (i & 31)) & 1u) != 0;
  const long out = final_label(code, closed);
  L[i] = out;
  if (border_zero && out == -100) {  // set_data_border_value.py:21-24 on this path's final labels
    for (int f = 0; f < F; ++f) x_crop[static_cast<long>(f) * npx + i] = 0.f;
  }
}

}  // namespace tp

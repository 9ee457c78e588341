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

// The data channels use fp32 arithmetic (the network input is fp32 and log10f is within 2 ulp: |error| < 2e-5 dB, the
// same bound as crimac_preprocess); double is kept for the label threshold decision only, where a rounding difference
// could flip a label.
TP_HD float sv_to_db_f32(float v, int scaled) {
  float d = 10.f * log10f(v + 1e-10f);
  if (d > 0.f) d = 0.f;
  if (d < -75.f) d = -75.f;
  if (scaled) d = 1.f + d / 75.f;
  return d;
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

// add_noise.py:28-38: (1 - change) + change * (increase * U(1,10) + (1 - increase) * U(0,1)), P(change) = 0.05,
// P(increase) = 0.5, the three draws independent.  One sample consumes 64 random bits: word `lo` decides change
// (31 bits against 0.05) and increase (1 bit), word `hi` is the uniform magnitude (24 bits, fp32 like the network input).
TP_HD float noise_from_bits(uint32_t lo, uint32_t hi) {
  if ((lo >> 1) >= 107374182u) return 1.f;  // 0.05 * 2^31
  const float u = (static_cast<float>(hi >> 8) + 0.5f) * (1.f / 16777216.f);
  return (lo & 1u) ? 1.f + 9.f * u : u;
}

// One Philox call serves the two samples (py, px) and (py + 8, px) with ((py >> 3) & 1) == 0 — the two range rows a
// thread of the gather kernel handles back to back — so the generator costs half a call per sample.  chan = crop * F +
// frequency.  The multiplier stays a pure function of (seed, chan, py, px).
TP_HD void noise_pair(uint64_t seed, int chan, int base_row, int px, float out[2]) {
  uint32_t r[4];
  philox4x32_10(static_cast<uint32_t>(px), static_cast<uint32_t>(base_row), static_cast<uint32_t>(chan), 0x43524d43u,
                static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32), r);
  out[0] = noise_from_bits(r[0], r[1]);
  out[1] = noise_from_bits(r[2], r[3]);
}

TP_HD float noise_multiplier(uint64_t seed, int chan, int py, int px) {
  float m[2];
  const int q = (py >> 3) & 1;
  noise_pair(seed, chan, py - 8 * q, px, m);
  return m[q];
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
  // the loads are issued unconditionally from clamped (always valid) addresses so that all of a thread's loads are in
  // flight together; a sample outside the survey is replaced afterwards
  const int Yc = Y < 0 ? 0 : (Y >= p.R ? p.R - 1 : Y);
  float s[4], l[4];
  bool in[4];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; ++k) {
    const int X = cx - p.pw / 2 + 1 + tx0 + ty + 8 * k;
    const int Xc = X < 0 ? 0 : (X >= p.P ? p.P - 1 : X);
    in[k] = Y >= 0 && Y < p.R && X >= 0 && X < p.P;
    s[k] = TP_LDG(p.sv + (static_cast<long>(src_f) * p.P + Xc) * p.R + Yc);
    l[k] = is_label ? TP_LDG(p.labels + static_cast<long>(Xc) * p.R + Yc) : 0.f;
  }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int k = 0; k < 4; ++k) {
    const int i = ty + 8 * k;
    tile[i][tx] = in[k] ? nan_to_num_f32(s[k]) : 0.f;  // boundary_val_data, dataset.py:360
    if (is_label) tile_lab[i][tx] = in[k] ? l[k] : -100.f;
  }
}

// Store phase: tx runs along PING, which is contiguous in the network input.
// fl = p.flags[crop], read by the caller before the load phase (one less dependent load after the barrier).
TP_HD void gather_store(const GatherParams& p, const TileCoord& blk, int tx, int ty, uint8_t fl,
                        const float (*tile)[33], const float (*tile_lab)[33]) {
  const bool is_label = blk.chan == p.F;
  const int tiles_x = p.pw >> 5;
  const int ty0 = (blk.tile / tiles_x) << 5, tx0 = (blk.tile % tiles_x) << 5;
  const int b = blk.crop;
  const bool noisy = (fl & 1) != 0, flip = (fl & 2) != 0;
  const int src_f = is_label ? p.thr_freq : blk.chan;
  const int px = tx0 + tx;
  const int ox = flip ? p.pw - 1 - px : px;  // flip_x_axis.py:22-25 (after the noise)
  for (int kp = 0; kp < 2; ++kp) {  // rows (ty + 16 kp) and (ty + 16 kp + 8) of the tile share one generator call
    float m[2] = {1.f, 1.f};
    if (noisy && p.noise == nullptr) noise_pair(p.seed, b * p.F + src_f, ty0 + ty + 16 * kp, px, m);
    for (int q = 0; q < 2; ++q) {
      const int j = ty + 16 * kp + 8 * q;
      const int py = ty0 + j;
      if (noisy && p.noise != nullptr)
        m[q] = TP_LDG(p.noise + ((static_cast<long>(b) * p.F + src_f) * p.ph + py) * p.pw + px);
      if (!is_label) {
        const float v = tile[tx][j] * m[q];
        p.x[((static_cast<long>(b) * p.F + blk.chan) * p.ph + py) * p.pw + ox] = sv_to_db_f32(v, p.scaled);
      } else {
        // float64 like the reference's out_data (dataset.py:361): the product of two fp32 numbers is exact in double
        const double v = static_cast<double>(tile[tx][j]) * static_cast<double>(m[q]);
        int code = label_code(tile_lab[tx][j]);
        // refine_label_boundary.py:88-89: (labels > 0) & (data > lo) & (data < hi) on the threshold frequency
        const bool positive = code == L_OTHER || code == L_SANDEEL || code == L_POSITIVE;
        if (positive && v > p.thr_lo && v < p.thr_hi) code |= kThresholdBit;
        p.lab[(static_cast<long>(b) * p.ph + py) * p.pw + ox] = code;
      }
    }
  }
}

// ---- the label kernel's phases.  A crop is split into kBands horizontal bands of ph/kBands range rows, one thread
// block each (the blocks of a crop form a cluster); a band keeps three bit masks in shared memory:
//   T  threshold mask,  rows [r0 - 6, r0 + rows + 6)   (rows outside the crop are 0)
//   D  its dilation cut to the bounding box, rows [r0 - 3, r0 + rows + 3)
//   E  the closing, rows [r0, r0 + rows)
constexpr int kBands = 8;

TP_HD void labels_dilate_band(const uint32_t* T, uint32_t* D, const BBox& bb, int wpr, int r0, int rows, int i) {
  const int ly = i / wpr, w = i - ly * wpr;  // D row ly = crop row r0 - 3 + ly = T row ly + 3
  // binary_closing on the bounding-box crop (refine_label_boundary.py:91) = dilation, cut to the box, erosion with
  // everything outside the box counting as 0 (border_value = 0 in both passes)
  D[i] = dilate_word(T, wpr, rows + 12, ly + 3, w) & bbox_word(bb, r0 - 3 + ly, w);
}

TP_HD void labels_erode_band(const uint32_t* D, uint32_t* E, int wpr, int rows, int i) {
  const int ly = i / wpr, w = i - ly * wpr;  // E row ly = crop row r0 + ly = D row ly + 3
  E[i] = erode_word(D, wpr, rows + 6, ly + 3, w);
}

// i = sample index inside the band; L / x_crop point at the crop, npx = ph * pw
TP_HD void labels_finish_band(long long* L, float* x_crop, const uint32_t* E, int F, int npx, int band_px0,
                              int border_zero, int i) {
  const int gi = band_px0 + i;
  const int code = static_cast<int>(L[gi]) & 7;
  const bool closed = ((E[i >> 5] >> (i & 31)) & 1u) != 0;
  const long out = final_label(code, closed);
  L[gi] = out;
  if (border_zero && out == -100) {  // set_data_border_value.py:21-24 on this path's final labels
    for (int f = 0; f < F; ++f) x_crop[static_cast<long>(f) * npx + gi] = 0.f;
  }
}

}  // namespace tp

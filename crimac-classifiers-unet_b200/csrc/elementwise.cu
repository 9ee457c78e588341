// CUDA-core kernels of the echogram U-Net path: everything that is HBM-bound rather than a dense contraction.
//   first conv (Cin = #frequencies, K = 36/54: fp32 in, too thin for the tensor pipe)   unet.py:76 (down_convs.0.main.0)
//   train-mode BatchNorm statistics finalize / apply (+ReLU, +2x2 max-pool, +concat write) unet.py:78-92
//   1x1 head forward / backward, class-weighted cross-entropy                              unet.py:342, pipeline.py:135-138,176
//   BatchNorm/ReLU backward (two-phase), max-pool backward + skip-gradient add             autograd of unet.py:76-93,130-136
// All activation tensors are NHWC bf16 "views" (pointer, pitch) so that concat buffers are written/read in place.
#include "host_util.h"
#include "devfn.cuh"
#include <cstdlib>

namespace {

constexpr int kMaxBlocks = 148 * 8;

// ============================================================================ first conv (fp32 NCHW in -> 64 ch)
// One thread = TWO horizontally adjacent pixels of an 8x16 tile (64 threads per tile): every broadcast LDS.128 of four
// weights feeds 8 FFMAs, which keeps the kernel FFMA-bound instead of shared-memory-bound.  Train mode accumulates the
// channel statistics of the stored (bf16-rounded) values per block: stats[blockIdx.x][2][64].
template <int CIN>
__global__ void __launch_bounds__(64) first_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                        const float* __restrict__ scale,
                                                        const float* __restrict__ shift, int relu, int NB, int H, int W,
                                                        bf16* __restrict__ out, int out_pitch, float* stats) {
  constexpr int K = CIN * 9;
  __shared__ __align__(16) float ws[K * 64];  // [k][co]
  __shared__ float s_sc[64], s_sh[64];
  __shared__ float s_red[2][2][64];
  __shared__ float s_acc[2][64];
  for (int i = threadIdx.x; i < K * 64; i += 64) {
    const int co = i & 63, k = i >> 6;
    ws[i] = w[co * K + k];
  }
  s_sc[threadIdx.x] = scale ? scale[threadIdx.x] : 1.f;
  s_sh[threadIdx.x] = shift ? shift[threadIdx.x] : 0.f;
  s_acc[0][threadIdx.x] = 0.f;
  s_acc[1][threadIdx.x] = 0.f;
  __syncthreads();
  const int tiles_x = (W + TILE_W - 1) / TILE_W, tiles_y = (H + TILE_H - 1) / TILE_H;
  const int total = NB * tiles_x * tiles_y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int py = threadIdx.x >> 3, px = (threadIdx.x & 7) * 2;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    int t = tile;
    const int tx = t % tiles_x;
    t /= tiles_x;
    const int ty = t % tiles_y;
    const int img = t / tiles_y;
    const int y = ty * TILE_H + py, xx = tx * TILE_W + px;
    const bool valid0 = y < H && xx < W, valid1 = y < H && xx + 1 < W;
    // 3 rows x 4 columns of inputs per frequency cover both pixels' 3x3 windows
    float in[CIN][3][4];
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const int yy = y + r - 1, xq = xx + c - 1;
          const bool ok = yy >= 0 && yy < H && xq >= 0 && xq < W;
          in[ci][r][c] = ok ? __ldg(&x[((static_cast<long>(img) * CIN + ci) * H + yy) * W + xq]) : 0.f;
        }
    bf16* dst = out + ((static_cast<long>(img) * H + y) * W + xx) * out_pitch;
#pragma unroll 1
    for (int half = 0; half < 2; ++half) {
      float a0[32], a1[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) a0[j] = a1[j] = 0.f;
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
        for (int tap = 0; tap < 9; ++tap) {
          const float v0 = in[ci][tap / 3][tap % 3], v1 = in[ci][tap / 3][tap % 3 + 1];
          const float4* wr = reinterpret_cast<const float4*>(&ws[(ci * 9 + tap) * 64 + half * 32]);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 wv = wr[j4];
            a0[4 * j4 + 0] = fmaf(v0, wv.x, a0[4 * j4 + 0]);
            a0[4 * j4 + 1] = fmaf(v0, wv.y, a0[4 * j4 + 1]);
            a0[4 * j4 + 2] = fmaf(v0, wv.z, a0[4 * j4 + 2]);
            a0[4 * j4 + 3] = fmaf(v0, wv.w, a0[4 * j4 + 3]);
            a1[4 * j4 + 0] = fmaf(v1, wv.x, a1[4 * j4 + 0]);
            a1[4 * j4 + 1] = fmaf(v1, wv.y, a1[4 * j4 + 1]);
            a1[4 * j4 + 2] = fmaf(v1, wv.z, a1[4 * j4 + 2]);
            a1[4 * j4 + 3] = fmaf(v1, wv.w, a1[4 * j4 + 3]);
          }
        }
      uint32_t pk0[16], pk1[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float sa = s_sc[half * 32 + 2 * j], sb = s_sc[half * 32 + 2 * j + 1];
        const float ha = s_sh[half * 32 + 2 * j], hb = s_sh[half * 32 + 2 * j + 1];
        float p0 = a0[2 * j] * sa + ha, p1 = a0[2 * j + 1] * sb + hb;
        float q0 = a1[2 * j] * sa + ha, q1 = a1[2 * j + 1] * sb + hb;
        if (relu) {
          p0 = fmaxf(p0, 0.f);
          p1 = fmaxf(p1, 0.f);
          q0 = fmaxf(q0, 0.f);
          q1 = fmaxf(q1, 0.f);
        }
        pk0[j] = pack_bf16x2(p0, p1);
        pk1[j] = pack_bf16x2(q0, q1);
      }
      if (valid0) {
        store16(dst + half * 32, pk0);
        store16(dst + half * 32 + 8, pk0 + 4);
        store16(dst + half * 32 + 16, pk0 + 8);
        store16(dst + half * 32 + 24, pk0 + 12);
      }
      if (valid1) {
        store16(dst + out_pitch + half * 32, pk1);
        store16(dst + out_pitch + half * 32 + 8, pk1 + 4);
        store16(dst + out_pitch + half * 32 + 16, pk1 + 8);
        store16(dst + out_pitch + half * 32 + 24, pk1 + 12);
      }
      if (stats != nullptr) {
        float s1[32], s2[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 u = unpack_bf16x2(pk0[j]), v = unpack_bf16x2(pk1[j]);
          const float ua = valid0 ? u.x : 0.f, ub = valid0 ? u.y : 0.f;
          const float va = valid1 ? v.x : 0.f, vb = valid1 ? v.y : 0.f;
          s1[2 * j] = ua + va;
          s1[2 * j + 1] = ub + vb;
          s2[2 * j] = ua * ua + va * va;
          s2[2 * j + 1] = ub * ub + vb * vb;
        }
        xpose_reduce(s1, lane);
        xpose_reduce(s2, lane);
        s_red[warp][0][half * 32 + lane] = s1[0];
        s_red[warp][1][half * 32 + lane] = s2[0];
      }
    }
    if (stats != nullptr) {
      __syncthreads();
      const int c = threadIdx.x;
      s_acc[0][c] += s_red[0][0][c] + s_red[1][0][c];
      s_acc[1][c] += s_red[0][1][c] + s_red[1][1][c];
      __syncthreads();
    }
  }
  if (stats != nullptr) {
    stats[static_cast<long>(blockIdx.x) * 128 + threadIdx.x] = s_acc[0][threadIdx.x];
    stats[static_cast<long>(blockIdx.x) * 128 + 64 + threadIdx.x] = s_acc[1][threadIdx.x];
  }
}

// ============================================================================ BN statistics -> scale/shift
// partials [m_tiles][2][C] (sum, sum of squares) -> batch mean / biased var; running stats (momentum, unbiased var).
// Row-lane `tl` (of 32) sums rows tl, tl+32, ... of partials[row][NQ][C] for channel c: eight rows per round, all loads
// of a round in flight before the first add (a round is one L2 round trip instead of one per row).
template <int NQ>
__device__ __forceinline__ void column_partial_sums(const float* __restrict__ partials, int nparts, int C, int c, int tl,
                                                    double& q0, double& q1) {
  float f[NQ][2];
#pragma unroll
  for (int q = 0; q < NQ; ++q) f[q][0] = f[q][1] = 0.f;
  for (int i0 = tl; i0 < nparts; i0 += 8 * 32) {
    float v[8][NQ];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = i0 + u * 32;
#pragma unroll
      for (int q = 0; q < NQ; ++q) v[u][q] = i < nparts ? partials[(static_cast<long>(i) * NQ + q) * C + c] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u)
#pragma unroll
      for (int q = 0; q < NQ; ++q) f[q][u & 1] += v[u][q];
  }
  q0 = static_cast<double>(f[0][0]) + static_cast<double>(f[0][1]);
  q1 = NQ > 1 ? static_cast<double>(f[NQ - 1][0]) + static_cast<double>(f[NQ - 1][1]) : 0.0;
}
// Sum over the 32 row-lanes of a (8 channels x 32 row-lanes) block; threads with tl == 0 receive the totals.
// A warp holds 4 row-lanes x 8 channels: two shuffles, then 8 per-warp values per channel through shared memory.
__device__ __forceinline__ void block_lane_sum8x32(double (&sh)[2][8][8], double& a, double& b) {
  a += __shfl_xor_sync(0xffffffffu, a, 8);
  b += __shfl_xor_sync(0xffffffffu, b, 8);
  a += __shfl_xor_sync(0xffffffffu, a, 16);
  b += __shfl_xor_sync(0xffffffffu, b, 16);
  const int cl = threadIdx.x & 7, w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) < 8) {
    sh[0][w][cl] = a;
    sh[1][w][cl] = b;
  }
  __syncthreads();
  if ((threadIdx.x >> 3) == 0) {
    a = sh[0][0][cl];
    b = sh[1][0][cl];
#pragma unroll
    for (int t = 1; t < 8; ++t) {
      a += sh[0][t][cl];
      b += sh[1][t][cl];
    }
  }
}

__global__ void __launch_bounds__(256) bn_finalize_kernel(const float* __restrict__ partials, int m_tiles, int C,
                                                          double count, const float* __restrict__ gamma,
                                                          const float* __restrict__ beta, float* running_mean,
                                                          float* running_var, long long* num_batches_tracked,
                                                          float momentum, float eps, float* scale, float* shift,
                                                          float* save_mean, float* save_invstd) {
  // block = 8 channels x 32 row-lanes.  The kernel is pure latency: every load a thread needs is issued before the
  // first use (eight partial rows per round; gamma / beta / running statistics up front by the threads that finish).
  __shared__ double sh[2][8][8];
  const int cl = threadIdx.x & 7, tl = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl;
  const bool fin = tl == 0 && c < C;
  float g = 0.f, b = 0.f, rm = 0.f, rv = 0.f;
  if (fin) {
    g = gamma[c];
    b = beta[c];
    if (running_mean != nullptr) {
      rm = running_mean[c];
      rv = running_var[c];
    }
  }
  double s1 = 0.0, s2 = 0.0;
  if (c < C) column_partial_sums<2>(partials, m_tiles, C, c, tl, s1, s2);
  block_lane_sum8x32(sh, s1, s2);
  if (fin) {
    const double mean = s1 / count;
    double var = s2 / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const float invstd = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    const float sc = g * invstd;
    scale[c] = sc;
    shift[c] = b - static_cast<float>(mean) * sc;
    save_mean[c] = static_cast<float>(mean);
    save_invstd[c] = invstd;
    if (running_mean != nullptr) {
      const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
      running_mean[c] = (1.f - momentum) * rm + momentum * static_cast<float>(mean);
      running_var[c] = (1.f - momentum) * rv + momentum * static_cast<float>(unbiased);
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
}

// eval-mode folding: y = conv*scale + shift with scale = g/sqrt(rv+eps), shift = (bias - rm)*scale + beta
__global__ void bn_fold_eval_kernel(const float* gamma, const float* beta, const float* rm, const float* rv,
                                    const float* conv_bias, float eps, int C, float* scale, float* shift) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c < C) {
    const float sc = gamma[c] * rsqrtf(rv[c] + eps);
    scale[c] = sc;
    shift[c] = (conv_bias[c] - rm[c]) * sc + beta[c];
  }
}

// ============================================================================ BN apply + ReLU (+pool)
// A thread owns one 8-channel group (its scale/shift stay in registers) and strides over pixels: no per-element index
// division, several independent 16-byte loads in flight.
template <bool POOL>
__global__ void __launch_bounds__(256) bn_apply_kernel(View raw, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, View act, View pool,
                                                       uint16_t* __restrict__ pool_arg) {
  const int groups = raw.C >> 3, ppb = blockDim.x / groups;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  float sc[8], sh[8];
  ldg8f(scale + g * 8, sc);
  ldg8f(shift + g * 8, sh);
  const unsigned stride = gridDim.x * ppb;
  if (!POOL) {
    const unsigned npix = static_cast<unsigned>(raw.N) * raw.H * raw.W;
    constexpr int U = 4;
    for (unsigned p0 = blockIdx.x * ppb + pl; p0 < npix; p0 += U * stride) {
      uint4 rv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned p = p0 + u * stride;
        rv[u] = p < npix ? __ldcs(reinterpret_cast<const uint4*>(raw.ptr + static_cast<size_t>(p) * raw.pitch + g * 8))
                         : make_uint4(0, 0, 0, 0);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned p = p0 + u * stride;
        if (p < npix) {
          const uint32_t rw[4] = {rv[u].x, rv[u].y, rv[u].z, rv[u].w};
          uint32_t pk[4];
#pragma unroll
          for (int h = 0; h < 4; ++h) {
            const float2 r = unpack_bf16x2(rw[h]);
            pk[h] = pack_bf16x2(fmaxf(fmaf(r.x, sc[2 * h], sh[2 * h]), 0.f), fmaxf(fmaf(r.y, sc[2 * h + 1], sh[2 * h + 1]), 0.f));
          }
          store16(act.ptr + static_cast<size_t>(p) * act.pitch + g * 8, pk);
        }
      }
    }
  } else {
    const unsigned Ho = raw.H >> 1, Wo = raw.W >> 1;
    const unsigned npool = static_cast<unsigned>(raw.N) * Ho * Wo;
    for (unsigned pp = blockIdx.x * ppb + pl; pp < npool; pp += stride) {
      const unsigned q = pp / Wo, xo = pp - q * Wo, n = q / Ho, yo = q - n * Ho;
      const size_t p00 = (static_cast<size_t>(n) * raw.H + 2 * yo) * raw.W + 2 * xo;
      uint4 rv[4];
#pragma unroll
      for (int w = 0; w < 4; ++w)
        rv[w] = __ldcs(reinterpret_cast<const uint4*>(raw.ptr + (p00 + (w >> 1) * raw.W + (w & 1)) * raw.pitch + g * 8));
      uint32_t mx[4] = {0, 0, 0, 0};
      float best[8];
      uint32_t arg = 0;  // 2 bits per channel: index (row*2 + col) of the FIRST maximal window element (PyTorch's tie-break)
#pragma unroll
      for (int w = 0; w < 4; ++w) {
        const uint32_t rw[4] = {rv[w].x, rv[w].y, rv[w].z, rv[w].w};
        uint32_t pk[4];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float2 r = unpack_bf16x2(rw[h]);
          pk[h] = pack_bf16x2(fmaxf(fmaf(r.x, sc[2 * h], sh[2 * h]), 0.f), fmaxf(fmaf(r.y, sc[2 * h + 1], sh[2 * h + 1]), 0.f));
          mx[h] = w == 0 ? pk[h] : bf16x2_max(mx[h], pk[h]);
          const float2 v = unpack_bf16x2(pk[h]);  // compare the stored (rounded) activations, as the backward used to
          if (w == 0) {
            best[2 * h] = v.x;
            best[2 * h + 1] = v.y;
          } else {
            if (v.x > best[2 * h]) {
              best[2 * h] = v.x;
              arg = (arg & ~(3u << (4 * h))) | (static_cast<uint32_t>(w) << (4 * h));
            }
            if (v.y > best[2 * h + 1]) {
              best[2 * h + 1] = v.y;
              arg = (arg & ~(3u << (4 * h + 2))) | (static_cast<uint32_t>(w) << (4 * h + 2));
            }
          }
        }
        store16(act.ptr + (p00 + (w >> 1) * raw.W + (w & 1)) * act.pitch + g * 8, pk);
      }
      if (pool_arg != nullptr) pool_arg[static_cast<size_t>(pp) * groups + g] = static_cast<uint16_t>(arg);
      store16(pool.ptr + static_cast<size_t>(pp) * pool.pitch + g * 8, mx);
    }
  }
}

// ============================================================================ 1x1 head forward (train path)
// act (N,H,W,64) bf16 -> logits (N,ncls,H,W) fp32
__global__ void __launch_bounds__(128) head_fwd_kernel(View act, const float* __restrict__ hw,
                                                       const float* __restrict__ hb, int ncls, float* logits) {
  __shared__ float s_w[CRIMAC_MAX_CLASSES * 64 + CRIMAC_MAX_CLASSES];
  for (int i = threadIdx.x; i < ncls * 64; i += blockDim.x) s_w[i] = hw[i];
  if (threadIdx.x < ncls) s_w[CRIMAC_MAX_CLASSES * 64 + threadIdx.x] = hb[threadIdx.x];
  __syncthreads();
  const long HW = static_cast<long>(act.H) * act.W;
  const long total = act.N * HW;
  for (long p = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long>(gridDim.x) * blockDim.x) {
    float lg[CRIMAC_MAX_CLASSES];
#pragma unroll
    for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k) lg[k] = k < ncls ? s_w[CRIMAC_MAX_CLASSES * 64 + k] : 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      float v[8];
      load8(act.ptr + p * act.pitch + g * 8, v);
#pragma unroll
      for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k)
        if (k < ncls) {
#pragma unroll
          for (int j = 0; j < 8; ++j) lg[k] = fmaf(v[j], s_w[k * 64 + g * 8 + j], lg[k]);
        }
    }
    const long n = p / HW, r = p - n * HW;
#pragma unroll
    for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k)
      if (k < ncls) logits[(n * ncls + k) * HW + r] = lg[k];
  }
}

// ============================================================================ class-weighted cross-entropy
// nn.CrossEntropyLoss(weight, ignore_index=-100, reduction='mean') (pipeline.py:135-138): per-pixel
// w[y]*(lse - z[y]); writes the UNNORMALISED gradient w[y]*(softmax - onehot) and per-block partial sums.
__global__ void __launch_bounds__(256) ce_fwd_bwd_kernel(const float* __restrict__ logits,
                                                         const long long* __restrict__ labels,
                                                         const float* __restrict__ cw, int ncls, long HW, long total,
                                                         long long ignore_index, float* dlogits, double* partials) {
  __shared__ double s_a[8], s_b[8];
  double la = 0.0, lb = 0.0;
  for (long p = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long>(gridDim.x) * blockDim.x) {
    const long n = p / HW, r = p - n * HW;
    float z[CRIMAC_MAX_CLASSES];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k)
      if (k < ncls) {
        z[k] = logits[(n * ncls + k) * HW + r];
        mx = fmaxf(mx, z[k]);
      }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k)
      if (k < ncls) sum += expf(z[k] - mx);
    const float lse = mx + logf(sum);
    const long long y = labels[p];
    const bool use = (y != ignore_index) && y >= 0 && y < ncls;
    // a label outside [0, n_classes) that is not ignore_index trips a device assert in PyTorch; here it poisons the
    // loss (NaN) so that the mistake is loud without a host synchronisation
    if (y != ignore_index && !use) la = static_cast<double>(NAN);
    const float wy = use ? cw[y] : 0.f;
#pragma unroll
    for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k)
      if (k < ncls) {
        const float pk = expf(z[k] - lse);
        if (dlogits) dlogits[(n * ncls + k) * HW + r] = wy * (pk - ((use && k == y) ? 1.f : 0.f));
        if (use && k == y) la += static_cast<double>(wy) * static_cast<double>(lse - z[k]);
      }
    lb += wy;
  }
  for (int o = 16; o > 0; o >>= 1) {
    la += __shfl_xor_sync(0xffffffffu, la, o);
    lb += __shfl_xor_sync(0xffffffffu, lb, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_a[threadIdx.x >> 5] = la;
    s_b[threadIdx.x >> 5] = lb;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) {
      la += s_a[i];
      lb += s_b[i];
    }
    partials[2 * blockIdx.x] = la;
    partials[2 * blockIdx.x + 1] = lb;
  }
}
// out[0] = loss, out[1] = 1/sum_w (gradient normaliser), out[2] = sum_w
__global__ void ce_finalize_kernel(const double* partials, int nblocks, float* out) {
  double a = 0.0, b = 0.0;
  for (int i = 0; i < nblocks; ++i) {
    a += partials[2 * i];
    b += partials[2 * i + 1];
  }
  out[0] = static_cast<float>(a / b);  // all-ignored batch -> NaN, as the reference
  out[1] = static_cast<float>(1.0 / b);
  out[2] = static_cast<float>(b);
}

// ============================================================================ validation-path loss
// pipeline.py:222-239 (set_label_ignore_val: -70 overlap, -30 refined boundary, -100 outside data, -10 unused species
// -> ignore; -50 below seabed -> background 0), :264 (the same weighted CE on the remapped labels) and :269-270
// (softmax, SANDEEL channel) in one pass over the eval-mode logits.  LT = int16 (what the dataset emits) or int64.
template <typename LT>
__global__ void __launch_bounds__(256) eval_loss_kernel(const float* __restrict__ logits, const LT* __restrict__ labels,
                                                        const float* __restrict__ cw, int ncls, long HW, long total,
                                                        int prob_class, float* __restrict__ prob_out,
                                                        long long* __restrict__ labels_out, double* partials) {
  __shared__ double s_a[8], s_b[8];
  double la = 0.0, lb = 0.0;
  for (long p = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long>(gridDim.x) * blockDim.x) {
    const long n = p / HW, r = p - n * HW;
    float z[CRIMAC_MAX_CLASSES];
    float mx = -INFINITY;
#pragma unroll
    for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k)
      if (k < ncls) {
        z[k] = logits[(n * ncls + k) * HW + r];
        mx = fmaxf(mx, z[k]);
      }
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k)
      if (k < ncls) sum += expf(z[k] - mx);
    const float lse = mx + logf(sum);
    long long y = static_cast<long long>(labels[p]);
    if (y == -70 || y == -30 || y == -100 || y == -10) y = -100;
    else if (y == -50) y = 0;
    if (labels_out != nullptr) labels_out[p] = y;
    const bool use = (y != -100) && y >= 0 && y < ncls;
    if (y != -100 && !use) la = static_cast<double>(NAN);  // invalid label: PyTorch asserts, we poison the loss
    const float wy = use ? cw[y] : 0.f;
#pragma unroll
    for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k)
      if (k < ncls) {
        if (use && k == y) la += static_cast<double>(wy) * static_cast<double>(lse - z[k]);
        if (k == prob_class && prob_out != nullptr) prob_out[p] = expf(z[k] - lse);
      }
    lb += wy;
  }
  for (int o = 16; o > 0; o >>= 1) {
    la += __shfl_xor_sync(0xffffffffu, la, o);
    lb += __shfl_xor_sync(0xffffffffu, lb, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_a[threadIdx.x >> 5] = la;
    s_b[threadIdx.x >> 5] = lb;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < 8; ++i) {
      la += s_a[i];
      lb += s_b[i];
    }
    partials[2 * blockIdx.x] = la;
    partials[2 * blockIdx.x + 1] = lb;
  }
}

// ============================================================================ 1x1 head backward
// dAct[p][c] = s * sum_k dl[k][p] W[k][c];  dW[k][c] = s * sum_p dl[k][p] act[p][c];  db[k] = s * sum_p dl[k][p]
// (s = *gscale or 1).  A thread owns one 8-channel group for its lifetime (8 threads per pixel: every warp access is
// 512 contiguous bytes) and keeps its dW partial sums in registers; per-block partials are summed by
// head_bwd_finalize.
template <int NCLS>
__global__ void __launch_bounds__(256) head_bwd_kernel(const float* __restrict__ dlogits, const float* gscale,
                                                       View act, const float* __restrict__ hw, View dact,
                                                       float* partials) {
  __shared__ float s_part[8][NCLS * 64 + NCLS];
  const int g = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float s = gscale ? *gscale : 1.f;
  const long HW = static_cast<long>(act.H) * act.W;
  const long total = act.N * HW;
  float w[NCLS][8], accw[NCLS][8], accb[NCLS];
#pragma unroll
  for (int k = 0; k < NCLS; ++k) {
    ldg8f(hw + k * 64 + g * 8, w[k]);
    accb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) accw[k][j] = 0.f;
  }
  const long stride = static_cast<long>(gridDim.x) * 32;
  for (long p0 = static_cast<long>(blockIdx.x) * 32 + pl; p0 < total; p0 += 2 * stride) {
    float dl[2][NCLS], a[2][8];
    bool ok[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long p = p0 + u * stride;
      ok[u] = p < total;
      const long n = ok[u] ? p / HW : 0, r = ok[u] ? p - n * HW : 0;
#pragma unroll
      for (int k = 0; k < NCLS; ++k) dl[u][k] = ok[u] ? s * __ldg(&dlogits[(n * NCLS + k) * HW + r]) : 0.f;
      if (ok[u]) {
        load8(act.ptr + p * act.pitch + g * 8, a[u]);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[u][j] = 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < NCLS; ++k) {
          t = fmaf(dl[u][k], w[k][j], t);
          accw[k][j] = fmaf(dl[u][k], a[u][j], accw[k][j]);
        }
        o[j] = t;
      }
#pragma unroll
      for (int k = 0; k < NCLS; ++k) accb[k] += dl[u][k];
      if (ok[u]) store8(dact.ptr + (p0 + u * stride) * dact.pitch + g * 8, o);
    }
  }
  // lanes g, g+8, g+16, g+24 of a warp hold the same channel group
#pragma unroll
  for (int k = 0; k < NCLS; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      accw[k][j] += __shfl_xor_sync(0xffffffffu, accw[k][j], 8);
      accw[k][j] += __shfl_xor_sync(0xffffffffu, accw[k][j], 16);
    }
    accb[k] += __shfl_xor_sync(0xffffffffu, accb[k], 8);
    accb[k] += __shfl_xor_sync(0xffffffffu, accb[k], 16);
  }
  if (lane < 8) {
#pragma unroll
    for (int k = 0; k < NCLS; ++k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s_part[warp][k * 64 + g * 8 + j] = accw[k][j];
      if (g == 0) s_part[warp][NCLS * 64 + k] = accb[k];
    }
  }
  __syncthreads();
  constexpr int PER = NCLS * 64 + NCLS;
  for (int i = threadIdx.x; i < PER; i += 256) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += s_part[wv][i];
    partials[static_cast<long>(blockIdx.x) * PER + i] = t;
  }
}
// ============================================================================ fused head + cross-entropy + head backward
// The fused train step (crimac_train_step) never materialises logits: one pass over the last activation computes the
// 1x1 head (unet.py:342), the class-weighted CE (pipeline.py:135-138,176) and the head's backward.  The loss
// normaliser 1/sum_w is only known after the whole batch has been reduced, so dAct, dW, db are produced UNNORMALISED;
// head_ce_finalize scales dW/db and publishes 1/sum_w, which the last layer's BatchNorm backward folds in (it is
// linear in its incoming gradient).  Thread mapping as head_bwd_kernel: 8 threads per pixel, 8 channels each.
template <int NCLS>
__global__ void __launch_bounds__(256) head_ce_fused_kernel(View act, const float* __restrict__ hw,
                                                            const float* __restrict__ hb,
                                                            const long long* __restrict__ labels,
                                                            const float* __restrict__ cw, long long ignore_index,
                                                            View dact, float* partials, double* loss_partials) {
  constexpr int PER = NCLS * 64 + NCLS;
  __shared__ float s_part[8][PER];
  __shared__ double s_loss[8][2];
  const int g = threadIdx.x & 7, pl = threadIdx.x >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long total = static_cast<long>(act.N) * act.H * act.W;
  float w[NCLS][8], accw[NCLS][8], accb[NCLS], bias[NCLS], wcls[NCLS];
#pragma unroll
  for (int k = 0; k < NCLS; ++k) {
    ldg8f(hw + k * 64 + g * 8, w[k]);
    bias[k] = hb[k];
    wcls[k] = cw[k];
    accb[k] = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) accw[k][j] = 0.f;
  }
  double la = 0.0, lb = 0.0;
  const long stride = static_cast<long>(gridDim.x) * 32;
  constexpr int U = 4;  // pixels per thread in flight (the kernel is bound by load latency, not by bandwidth or math)
  for (long p0 = static_cast<long>(blockIdx.x) * 32 + pl; p0 < total; p0 += U * stride) {
    uint4 av[U];
    long long y[U];
    bool ok[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long p = p0 + u * stride;
      ok[u] = p < total;  // uniform over the 8 threads of a pixel
      y[u] = ok[u] ? labels[p] : ignore_index;
      av[u] = ok[u] ? *reinterpret_cast<const uint4*>(act.ptr + p * act.pitch + g * 8) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      float a[8];
      {
        const float2 a0 = unpack_bf16x2(av[u].x), a1 = unpack_bf16x2(av[u].y), a2 = unpack_bf16x2(av[u].z),
                     a3 = unpack_bf16x2(av[u].w);
        a[0] = a0.x; a[1] = a0.y; a[2] = a1.x; a[3] = a1.y; a[4] = a2.x; a[5] = a2.y; a[6] = a3.x; a[7] = a3.y;
      }
      float z[NCLS];
#pragma unroll
      for (int k = 0; k < NCLS; ++k) {
        float t = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) t = fmaf(a[j], w[k][j], t);
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        t += __shfl_xor_sync(0xffffffffu, t, 4);
        z[k] = t + bias[k];
      }
      // the softmax / loss of a pixel is computed ONCE, by the first of its 8 threads, and the NCLS logit gradients are
      // broadcast to the others (7/8 of the expf / logf work of the kernel was redundant)
      const bool use = ok[u] && (y[u] != ignore_index) && y[u] >= 0 && y[u] < NCLS;
      float dl[NCLS] = {};
      if (g == 0) {
        float mx = z[0];
#pragma unroll
        for (int k = 1; k < NCLS; ++k) mx = fmaxf(mx, z[k]);
        // fast-math exp / log (2^-21 relative): the kernel is instruction-issue bound, and the softmax of one pixel in 8
        // was most of its instructions; the loss is a mean over ~2 M pixels and the tests hold it to 1e-5 relative
        float e[NCLS], sum = 0.f;
#pragma unroll
        for (int k = 0; k < NCLS; ++k) {
          e[k] = __expf(z[k] - mx);
          sum += e[k];
        }
        const float lse = mx + __logf(sum), inv = __fdividef(1.f, sum);
        if (ok[u] && y[u] != ignore_index && !use) la = static_cast<double>(NAN);  // invalid label: see ce_fwd_bwd_kernel
        float wy = 0.f, zy = 0.f;
#pragma unroll
        for (int k = 0; k < NCLS; ++k)
          if (use && y[u] == k) {
            wy = wcls[k];
            zy = z[k];
          }
#pragma unroll
        for (int k = 0; k < NCLS; ++k) dl[k] = wy * (e[k] * inv - ((use && y[u] == k) ? 1.f : 0.f));
        if (use) la += static_cast<double>(wy) * static_cast<double>(lse - zy);
        lb += wy;
      }
#pragma unroll
      for (int k = 0; k < NCLS; ++k) dl[k] = __shfl_sync(0xffffffffu, dl[k], lane & ~7);
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = 0.f;
#pragma unroll
        for (int k = 0; k < NCLS; ++k) {
          t = fmaf(dl[k], w[k][j], t);
          accw[k][j] = fmaf(dl[k], a[j], accw[k][j]);
        }
        o[j] = t;
      }
#pragma unroll
      for (int k = 0; k < NCLS; ++k) accb[k] += dl[k];
      if (ok[u]) store8(dact.ptr + (p0 + u * stride) * dact.pitch + g * 8, o);
    }
  }
#pragma unroll
  for (int k = 0; k < NCLS; ++k) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      accw[k][j] += __shfl_xor_sync(0xffffffffu, accw[k][j], 8);
      accw[k][j] += __shfl_xor_sync(0xffffffffu, accw[k][j], 16);
    }
    accb[k] += __shfl_xor_sync(0xffffffffu, accb[k], 8);
    accb[k] += __shfl_xor_sync(0xffffffffu, accb[k], 16);
  }
  la += __shfl_xor_sync(0xffffffffu, la, 8);
  la += __shfl_xor_sync(0xffffffffu, la, 16);
  lb += __shfl_xor_sync(0xffffffffu, lb, 8);
  lb += __shfl_xor_sync(0xffffffffu, lb, 16);
  if (lane < 8) {
#pragma unroll
    for (int k = 0; k < NCLS; ++k) {
#pragma unroll
      for (int j = 0; j < 8; ++j) s_part[warp][k * 64 + g * 8 + j] = accw[k][j];
      if (g == 0) s_part[warp][NCLS * 64 + k] = accb[k];
    }
    if (g == 0) {
      s_loss[warp][0] = la;
      s_loss[warp][1] = lb;
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < PER; i += 256) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += s_part[wv][i];
    partials[static_cast<long>(blockIdx.x) * PER + i] = t;
  }
  if (threadIdx.x < 2) {
    double t = 0.0;
    for (int wv = 0; wv < 8; ++wv) t += s_loss[wv][threadIdx.x];
    loss_partials[2 * blockIdx.x + threadIdx.x] = t;
  }
}
// out3 = {loss, 1/sum_w, sum_w}; dw, db = (1/sum_w) * sum over blocks of the unnormalised partials.
// Block = 8 gradient elements x 32 block-lanes; every block re-derives sum_w (cheap) so that one launch suffices.
__global__ void __launch_bounds__(256) head_ce_finalize_kernel(const float* __restrict__ partials,
                                                               const double* __restrict__ loss_partials, int nblocks,
                                                               int ncls, float* dw, float* db, float* out3) {
  __shared__ double s_a[256], s_b[256];
  __shared__ float s_g[32][8];
  double a = 0.0, b = 0.0;
  for (int i = threadIdx.x; i < nblocks; i += 256) {
    a += loss_partials[2 * i];
    b += loss_partials[2 * i + 1];
  }
  s_a[threadIdx.x] = a;
  s_b[threadIdx.x] = b;
  const int per = ncls * 64 + ncls;
  const int el = threadIdx.x & 7, bl = threadIdx.x >> 3;
  const int i = blockIdx.x * 8 + el;
  float t = 0.f;
  if (i < per)
    for (int blk = bl; blk < nblocks; blk += 32) t += partials[static_cast<long>(blk) * per + i];
  s_g[bl][el] = t;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (static_cast<int>(threadIdx.x) < o) {
      s_a[threadIdx.x] += s_a[threadIdx.x + o];
      s_b[threadIdx.x] += s_b[threadIdx.x + o];
    }
    __syncthreads();
  }
  a = s_a[0];
  b = s_b[0];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    out3[0] = static_cast<float>(a / b);  // all-ignored batch -> NaN, as the reference
    out3[1] = static_cast<float>(1.0 / b);
    out3[2] = static_cast<float>(b);
  }
  if (bl == 0 && i < per) {
    double g = 0.0;
    for (int l = 0; l < 32; ++l) g += s_g[l][el];
    const float v = static_cast<float>(g / b);
    if (i < ncls * 64) dw[i] = v; else db[i - ncls * 64] = v;
  }
}

__global__ void head_bwd_finalize_kernel(const float* partials, int nparts, int ncls, float* dw, float* db,
                                         int accumulate) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = ncls * 64 + ncls;
  if (i >= per) return;
  double a = 0.0;
  for (int b = 0; b < nparts; ++b) a += partials[static_cast<long>(b) * per + i];
  float* d = i < ncls * 64 ? dw + i : db + (i - ncls * 64);
  *d = accumulate ? *d + static_cast<float>(a) : static_cast<float>(a);
}

// ============================================================================ BN + ReLU backward
// Thread = fixed 8-channel group (its per-channel constants stay in registers), grid-stride over pixels; the block's
// partial sums go to partials[blockIdx.x][NQ][C] and are combined by a small parallel finalize kernel (deterministic).
template <int NQ>
__device__ __forceinline__ void block_channel_sums(const float (&acc)[NQ][8], int C, int g, int pl, int ppb,
                                                   float* partials) {
  extern __shared__ float s_cr[];  // [ppb][NQ][C]
#pragma unroll
  for (int q = 0; q < NQ; ++q)
#pragma unroll
    for (int j = 0; j < 8; ++j) s_cr[(pl * NQ + q) * C + g * 8 + j] = acc[q][j];
  __syncthreads();
  for (int i = threadIdx.x; i < NQ * C; i += blockDim.x) {
    float a = 0.f;
    for (int l = 0; l < ppb; ++l) a += s_cr[l * NQ * C + i];
    partials[static_cast<long>(blockIdx.x) * NQ * C + i] = a;
  }
}

// phase 1: g = dA * (bn(raw) > 0);  partial 0 = sum g, partial 1 = sum g * raw.  (sum g * xhat is derived from the two
// by the finalize kernel: invstd * (sum g*raw - mean * sum g); the kernel then needs only two per-channel constants
// in registers, which doubles its occupancy - it is bound by the bytes it keeps in flight.)
__global__ void __launch_bounds__(256, 3) bn_bwd_reduce_kernel(View dact, View raw, const float* __restrict__ scale,
                                                               const float* __restrict__ shift, float* partials) {
  const int C = raw.C, groups = C >> 3, ppb = blockDim.x / groups;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  const long npix = static_cast<long>(raw.N) * raw.H * raw.W;
  float sc[8], sh[8];
  ldg8f(scale + g * 8, sc);
  ldg8f(shift + g * 8, sh);
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  const long stride = static_cast<long>(gridDim.x) * ppb;
  constexpr int U = 4;  // pixels per iteration: 2*U independent 16-byte loads in flight per thread
  for (long p0 = static_cast<long>(blockIdx.x) * ppb + pl; p0 < npix; p0 += U * stride) {
    uint4 dv[U], rv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long p = p0 + u * stride;
      const bool ok = p < npix;
      dv[u] = ok ? __ldcs(reinterpret_cast<const uint4*>(dact.ptr + p * dact.pitch + g * 8)) : make_uint4(0, 0, 0, 0);
      rv[u] = ok ? __ldcs(reinterpret_cast<const uint4*>(raw.ptr + p * raw.pitch + g * 8)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t dw[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w}, rw[4] = {rv[u].x, rv[u].y, rv[u].z, rv[u].w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float2 d = unpack_bf16x2(dw[h]), r = unpack_bf16x2(rw[h]);
        const float g0 = fmaf(r.x, sc[2 * h], sh[2 * h]) > 0.f ? d.x : 0.f;
        const float g1 = fmaf(r.y, sc[2 * h + 1], sh[2 * h + 1]) > 0.f ? d.y : 0.f;
        acc[0][2 * h] += g0;
        acc[0][2 * h + 1] += g1;
        acc[1][2 * h] = fmaf(g0, r.x, acc[1][2 * h]);
        acc[1][2 * h + 1] = fmaf(g1, r.y, acc[1][2 * h + 1]);
      }
    }
  }
  block_channel_sums<2>(acc, C, g, pl, ppb, partials);
}

// The encoder's second convs: the incoming gradient dA = dSkip + unpool(dPool) (autograd of unet.py:86,92,132) is not
// read from memory but formed on the fly - the separate pool_bwd_add pass (4.6 bytes per element, written and read back)
// disappears.  A thread owns one 2x2 pooling WINDOW x 8 channels (the mapping of bn_apply's pooling variant): one load of
// the pooled gradient and of the 2-bit-per-channel arg-max word serves four pixels, and the window position each
// channel's gradient routes to is a compile-time constant of the unrolled loop.
struct PoolWindow {
  long p00;       // linear pixel index of the window's upper-left pixel
  uint4 dp;       // pooled gradient, 8 channels
  uint32_t arg;   // 2 bits per channel: window position (row*2 + col) of the forward's first maximum
};
__device__ __forceinline__ PoolWindow load_window(const uint16_t* __restrict__ pool_arg, const View& dpool, int H, int W,
                                                  unsigned pp, int g, int groups) {
  const unsigned Wo = W >> 1, Ho = H >> 1;
  const unsigned q = pp / Wo, xo = pp - q * Wo, n = q / Ho, yo = q - n * Ho;
  PoolWindow w;
  w.p00 = (static_cast<long>(n) * H + 2 * yo) * W + 2 * xo;
  w.dp = __ldg(reinterpret_cast<const uint4*>(dpool.ptr + static_cast<long>(pp) * dpool.pitch + g * 8));
  w.arg = __ldg(pool_arg + static_cast<long>(pp) * groups + g);
  return w;
}
// d[8] = dSkip values of window pixel `pos` + the pooled gradient of the channels whose maximum was at `pos`
template <int POS>
__device__ __forceinline__ void window_grad(const uint4& ds, const PoolWindow& w, float (&d)[8]) {
  const uint32_t sw[4] = {ds.x, ds.y, ds.z, ds.w}, pw[4] = {w.dp.x, w.dp.y, w.dp.z, w.dp.w};
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float2 s = unpack_bf16x2(sw[h]), t = unpack_bf16x2(pw[h]);
    d[2 * h] = s.x + ((((w.arg >> (4 * h)) & 3u) == POS) ? t.x : 0.f);
    d[2 * h + 1] = s.y + ((((w.arg >> (4 * h + 2)) & 3u) == POS) ? t.y : 0.f);
  }
}

__global__ void __launch_bounds__(256, 2) bn_bwd_reduce_pool_kernel(const uint16_t* __restrict__ pool_arg, View dpool,
                                                                    View dskip, View raw,
                                                                    const float* __restrict__ scale,
                                                                    const float* __restrict__ shift, float* partials) {
  const int C = raw.C, groups = C >> 3, ppb = blockDim.x / groups;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  const unsigned npool = static_cast<unsigned>(raw.N) * (raw.H >> 1) * (raw.W >> 1);
  float sc[8], sh[8];
  ldg8f(scale + g * 8, sc);
  ldg8f(shift + g * 8, sh);
  float acc[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = acc[1][j] = 0.f;
  const unsigned stride = gridDim.x * ppb;
  for (unsigned pp = blockIdx.x * ppb + pl; pp < npool; pp += stride) {
    const PoolWindow w = load_window(pool_arg, dpool, raw.H, raw.W, pp, g, groups);
    uint4 ds[4], rv[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long p = w.p00 + (k >> 1) * raw.W + (k & 1);
      ds[k] = __ldcs(reinterpret_cast<const uint4*>(dskip.ptr + p * dskip.pitch + g * 8));
      rv[k] = __ldcs(reinterpret_cast<const uint4*>(raw.ptr + p * raw.pitch + g * 8));
    }
    auto one = [&](const float (&d)[8], const uint4& r4) {
      const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
      for (int h = 0; h < 4; ++h) {
        const float2 r = unpack_bf16x2(rw[h]);
        const float g0 = fmaf(r.x, sc[2 * h], sh[2 * h]) > 0.f ? d[2 * h] : 0.f;
        const float g1 = fmaf(r.y, sc[2 * h + 1], sh[2 * h + 1]) > 0.f ? d[2 * h + 1] : 0.f;
        acc[0][2 * h] += g0;
        acc[0][2 * h + 1] += g1;
        acc[1][2 * h] = fmaf(g0, r.x, acc[1][2 * h]);
        acc[1][2 * h + 1] = fmaf(g1, r.y, acc[1][2 * h + 1]);
      }
    };
    float d[8];
    window_grad<0>(ds[0], w, d); one(d, rv[0]);
    window_grad<1>(ds[1], w, d); one(d, rv[1]);
    window_grad<2>(ds[2], w, d); one(d, rv[2]);
    window_grad<3>(ds[3], w, d); one(d, rv[3]);
  }
  block_channel_sums<2>(acc, C, g, pl, ppb, partials);
}

// partials [nparts][NQ][C] -> out_q[c] = sum over parts (block = 32 channels x 8 part-lanes)
// NQ == 2: BN backward: dbeta = q0, dgamma = q1, c1 = q0/count, c2 = q1/count.  NQ == 1: plain column sum.
template <int NQ>
__global__ void __launch_bounds__(256) partial_sum_finalize_kernel(const float* __restrict__ partials, int nparts, int C,
                                                                   double count, float* out0, float* out1,
                                                                   int accumulate, float* c1, float* c2,
                                                                   const float* gscale, const float* mean,
                                                                   const float* invstd, float* zero_out) {
  // block = 8 channels x 32 part-lanes; see column_partial_sums / block_lane_sum8x32
  __shared__ double sh[2][8][8];
  const int cl = threadIdx.x & 7, tl = threadIdx.x >> 3;
  const int c = blockIdx.x * 8 + cl;
  const bool fin = tl == 0 && c < C;
  double mu = 0.0, is = 0.0, gs = 1.0;
  float prev0 = 0.f, prev1 = 0.f;
  if (fin) {   // everything the finishing threads need, issued before the partial rows
    if (gscale) gs = static_cast<double>(*gscale);  // incoming gradient was produced unnormalised
    if (NQ == 2 && mean != nullptr) {
      mu = static_cast<double>(mean[c]);
      is = static_cast<double>(invstd[c]);
    }
    if (accumulate) {
      prev0 = out0[c];
      if (NQ == 2) prev1 = out1[c];
    }
  }
  double a[2] = {0.0, 0.0};
  if (c < C) column_partial_sums<NQ>(partials, nparts, C, c, tl, a[0], a[1]);
  block_lane_sum8x32(sh, a[0], a[1]);
  if (fin) {
    a[0] *= gs;
    a[1] *= gs;
    // BN backward: the second partial is sum g*raw; sum g*xhat = invstd * (sum g*raw - mean * sum g)
    if (NQ == 2 && mean != nullptr) a[1] = is * (a[1] - mu * a[0]);
    out0[c] = prev0 + static_cast<float>(a[0]);
    if (zero_out != nullptr && !accumulate) zero_out[c] = 0.f;
    if (NQ == 2) {
      out1[c] = prev1 + static_cast<float>(a[1]);
      c1[c] = static_cast<float>(a[0] / count);
      c2[c] = static_cast<float>(a[1] / count);
    }
  }
}

// per-channel constants of phase 2: dRaw = scale * (g - c1 - xhat*c2) = A*g + B*raw + K
struct BnBwdCoef {
  float sc[8], sh[8], ca[8], cb[8], ck[8];
};
__device__ __forceinline__ void load_bn_bwd_coef(BnBwdCoef& k, int g, const float* scale, const float* shift,
                                                 const float* mean, const float* invstd, const float* c1,
                                                 const float* c2, const float* gscale) {
  float mu[8], is[8], k1[8], k2[8];
  ldg8f(scale + g * 8, k.sc);
  ldg8f(shift + g * 8, k.sh);
  ldg8f(mean + g * 8, mu);
  ldg8f(invstd + g * 8, is);
  ldg8f(c1 + g * 8, k1);
  ldg8f(c2 + g * 8, k2);
  const float gs = gscale ? *gscale : 1.f;  // the incoming gradient was produced unnormalised (fused head/CE)
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    k.ca[j] = k.sc[j] * gs;
    k.cb[j] = -k.sc[j] * k2[j] * is[j];
    k.ck[j] = -k.sc[j] * k1[j] - k.cb[j] * mu[j];
  }
}
__device__ __forceinline__ void bn_bwd_pixel(const BnBwdCoef& k, const float (&d)[8], const uint4& r4, float (&o)[8]) {
  const uint32_t rw[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
  for (int h = 0; h < 4; ++h) {
    const float2 r = unpack_bf16x2(rw[h]);
    const float g0 = fmaf(r.x, k.sc[2 * h], k.sh[2 * h]) > 0.f ? d[2 * h] : 0.f;
    const float g1 = fmaf(r.y, k.sc[2 * h + 1], k.sh[2 * h + 1]) > 0.f ? d[2 * h + 1] : 0.f;
    o[2 * h] = fmaf(k.ca[2 * h], g0, fmaf(k.cb[2 * h], r.x, k.ck[2 * h]));
    o[2 * h + 1] = fmaf(k.ca[2 * h + 1], g1, fmaf(k.cb[2 * h + 1], r.y, k.ck[2 * h + 1]));
  }
}

// phase 2 (apply): dRaw = A*g + B*raw + K per channel -> bf16.
// (The conv-bias gradient sum(dRaw) is EXACTLY zero in exact arithmetic - train-mode BatchNorm removes any bias - and the
// reference's value is pure cancellation noise ~1e-9; it is written as 0 by the finalize kernel instead of being summed.)
__global__ void __launch_bounds__(256, 2) bn_bwd_apply_kernel(View dact, View raw, const float* __restrict__ scale,
                                                              const float* __restrict__ shift,
                                                              const float* __restrict__ mean,
                                                              const float* __restrict__ invstd,
                                                              const float* __restrict__ c1, const float* __restrict__ c2,
                                                              const float* gscale, View draw) {
  const int C = raw.C, groups = C >> 3, ppb = blockDim.x / groups;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  const long npix = static_cast<long>(raw.N) * raw.H * raw.W;
  BnBwdCoef k;
  load_bn_bwd_coef(k, g, scale, shift, mean, invstd, c1, c2, gscale);
  const long stride = static_cast<long>(gridDim.x) * ppb;
  constexpr int U = 4;
  for (long p0 = static_cast<long>(blockIdx.x) * ppb + pl; p0 < npix; p0 += U * stride) {
    uint4 dv[U], rv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long p = p0 + u * stride;
      const bool ok = p < npix;
      dv[u] = ok ? __ldcs(reinterpret_cast<const uint4*>(dact.ptr + p * dact.pitch + g * 8)) : make_uint4(0, 0, 0, 0);
      rv[u] = ok ? __ldcs(reinterpret_cast<const uint4*>(raw.ptr + p * raw.pitch + g * 8)) : make_uint4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long p = p0 + u * stride;
      if (p < npix) {
        const uint32_t dw[4] = {dv[u].x, dv[u].y, dv[u].z, dv[u].w};
        float d[8], o[8];
#pragma unroll
        for (int h = 0; h < 4; ++h) {
          const float2 t = unpack_bf16x2(dw[h]);
          d[2 * h] = t.x;
          d[2 * h + 1] = t.y;
        }
        bn_bwd_pixel(k, d, rv[u], o);
        store8(draw.ptr + p * draw.pitch + g * 8, o);
      }
    }
  }
}

__global__ void __launch_bounds__(256, 2) bn_bwd_apply_pool_kernel(const uint16_t* __restrict__ pool_arg, View dpool,
                                                                   View dskip, View raw, const float* __restrict__ scale,
                                                                   const float* __restrict__ shift,
                                                                   const float* __restrict__ mean,
                                                                   const float* __restrict__ invstd,
                                                                   const float* __restrict__ c1,
                                                                   const float* __restrict__ c2, const float* gscale,
                                                                   View draw) {
  const int C = raw.C, groups = C >> 3, ppb = blockDim.x / groups;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  const unsigned npool = static_cast<unsigned>(raw.N) * (raw.H >> 1) * (raw.W >> 1);
  BnBwdCoef k;
  load_bn_bwd_coef(k, g, scale, shift, mean, invstd, c1, c2, gscale);
  const unsigned stride = gridDim.x * ppb;
  for (unsigned pp = blockIdx.x * ppb + pl; pp < npool; pp += stride) {
    const PoolWindow w = load_window(pool_arg, dpool, raw.H, raw.W, pp, g, groups);
    uint4 ds[4], rv[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long p = w.p00 + (q >> 1) * raw.W + (q & 1);
      ds[q] = __ldcs(reinterpret_cast<const uint4*>(dskip.ptr + p * dskip.pitch + g * 8));
      rv[q] = __ldcs(reinterpret_cast<const uint4*>(raw.ptr + p * raw.pitch + g * 8));
    }
    float d[8], o[8];
    window_grad<0>(ds[0], w, d); bn_bwd_pixel(k, d, rv[0], o);
    store8(draw.ptr + w.p00 * draw.pitch + g * 8, o);
    window_grad<1>(ds[1], w, d); bn_bwd_pixel(k, d, rv[1], o);
    store8(draw.ptr + (w.p00 + 1) * draw.pitch + g * 8, o);
    window_grad<2>(ds[2], w, d); bn_bwd_pixel(k, d, rv[2], o);
    store8(draw.ptr + (w.p00 + raw.W) * draw.pitch + g * 8, o);
    window_grad<3>(ds[3], w, d); bn_bwd_pixel(k, d, rv[3], o);
    store8(draw.ptr + (w.p00 + raw.W + 1) * draw.pitch + g * 8, o);
  }
}

// plain per-channel sum of a view (ConvTranspose bias gradient)
__global__ void __launch_bounds__(256) view_colsum_kernel(View v, float* partials) {
  const int C = v.C, groups = C >> 3, ppb = blockDim.x / groups;
  const int g = threadIdx.x % groups, pl = threadIdx.x / groups;
  const long npix = static_cast<long>(v.N) * v.H * v.W;
  float acc[1][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[0][j] = 0.f;
  for (long p = static_cast<long>(blockIdx.x) * ppb + pl; p < npix; p += static_cast<long>(gridDim.x) * ppb) {
    float d[8];
    load8(v.ptr + p * v.pitch + g * 8, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[0][j] += d[j];
  }
  block_channel_sums<1>(acc, C, g, pl, ppb, partials);
}

// ============================================================================ max-pool backward + skip gradient
// dA[2x2 window] = dSkip[window] + (first maximal element of the window ? dP : 0)     (autograd of unet.py:86,92,132)
// The arg-max comes from the 2-bit-per-channel map the forward BN-apply pass wrote (16 bits per 8 channels), so the
// full-resolution activation is not read again.
__global__ void __launch_bounds__(256) pool_bwd_add_kernel(const uint16_t* __restrict__ pool_arg, View dpool, View dskip,
                                                           View dact) {
  const int groups = dact.C >> 3;
  const int Ho = dact.H >> 1, Wo = dact.W >> 1;
  const long total = static_cast<long>(dact.N) * Ho * Wo * groups;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    long pix = i / groups;
    const long pp = pix;
    const int xo = static_cast<int>(pix % Wo);
    pix /= Wo;
    const int yo = static_cast<int>(pix % Ho);
    const int n = static_cast<int>(pix / Ho);
    float dp[8];
    load8(dpool.ptr + pp * dpool.pitch + g * 8, dp);
    const uint32_t arg = pool_arg[pp * groups + g];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const long p = (static_cast<long>(n) * dact.H + 2 * yo + (q >> 1)) * dact.W + 2 * xo + (q & 1);
      float o[8];
      if (dskip.ptr != nullptr)
        load8(dskip.ptr + p * dskip.pitch + g * 8, o);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += (((arg >> (2 * j)) & 3u) == static_cast<uint32_t>(q)) ? dp[j] : 0.f;
      store8(dact.ptr + p * dact.pitch + g * 8, o);
    }
  }
}

// ============================================================================ bilinear 2x up-sampling (up_mode "upsample")
// nn.Upsample(mode="bilinear", scale_factor=2) = F.interpolate(align_corners=False) (reference unet.py:50-56): output row
// Y reads source rows (Y>>1) - 1 + (Y&1) and that + 1, clamped to the image, with weights 0.25 / 0.75 (even Y) resp.
// 0.75 / 0.25 (odd Y); the same along pings.  One thread = one output pixel x 8 channels.
__global__ void __launch_bounds__(256) upsample2x_kernel(View in, View out) {
  const int groups = in.C >> 3;
  const int Ho = in.H * 2, Wo = in.W * 2;
  const long total = static_cast<long>(in.N) * Ho * Wo * groups;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    long pix = i / groups;
    const int X = static_cast<int>(pix % Wo);
    pix /= Wo;
    const int Y = static_cast<int>(pix % Ho);
    const int n = static_cast<int>(pix / Ho);
    const int y0 = max((Y >> 1) - 1 + (Y & 1), 0), y1 = min((Y >> 1) + (Y & 1), in.H - 1);
    const int x0 = max((X >> 1) - 1 + (X & 1), 0), x1 = min((X >> 1) + (X & 1), in.W - 1);
    const float wy0 = (Y & 1) ? 0.75f : 0.25f, wx0 = (X & 1) ? 0.75f : 0.25f;
    float a[8], b[8], c[8], d[8], o[8];
    const long base = static_cast<long>(n) * in.H;
    load8(in.ptr + ((base + y0) * in.W + x0) * in.pitch + g * 8, a);
    load8(in.ptr + ((base + y0) * in.W + x1) * in.pitch + g * 8, b);
    load8(in.ptr + ((base + y1) * in.W + x0) * in.pitch + g * 8, c);
    load8(in.ptr + ((base + y1) * in.W + x1) * in.pitch + g * 8, d);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = wy0 * (wx0 * a[j] + (1.f - wx0) * b[j]) + (1.f - wy0) * (wx0 * c[j] + (1.f - wx0) * d[j]);
    store8(out.ptr + ((static_cast<long>(n) * Ho + Y) * Wo + X) * out.pitch + g * 8, o);
  }
}
// Adjoint of the above: source pixel (y, x) collects dOut of the up to 4 x 4 output pixels that read it.
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(View dout, View din) {
  const int groups = din.C >> 3;
  const int Ho = din.H * 2, Wo = din.W * 2;
  const long total = static_cast<long>(din.N) * din.H * din.W * groups;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int g = static_cast<int>(i % groups);
    long pix = i / groups;
    const int x = static_cast<int>(pix % din.W);
    pix /= din.W;
    const int y = static_cast<int>(pix % din.H);
    const int n = static_cast<int>(pix / din.H);
    // weights of output rows 2y-1 .. 2y+2 on source row y (edge rows also take the clamped neighbour's share)
    float wy[4] = {y >= 1 ? 0.25f : 0.f, y == 0 ? 1.f : 0.75f, y == din.H - 1 ? 1.f : 0.75f, y <= din.H - 2 ? 0.25f : 0.f};
    float wx[4] = {x >= 1 ? 0.25f : 0.f, x == 0 ? 1.f : 0.75f, x == din.W - 1 ? 1.f : 0.75f, x <= din.W - 2 ? 0.25f : 0.f};
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
#pragma unroll
    for (int dy = 0; dy < 4; ++dy) {
      const int Y = 2 * y - 1 + dy;
      if (wy[dy] == 0.f) continue;
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) {
        const int X = 2 * x - 1 + dx;
        if (wx[dx] == 0.f) continue;
        float v[8];
        load8(dout.ptr + ((static_cast<long>(n) * Ho + Y) * Wo + X) * dout.pitch + g * 8, v);
        const float w = wy[dy] * wx[dx];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] = fmaf(w, v[j], acc[j]);
      }
    }
    store8(din.ptr + ((static_cast<long>(n) * din.H + y) * din.W + x) * din.pitch + g * 8, acc);
  }
}

// (Cout,Cin,1,1) fp32 -> [Cout][Cin] bf16 (the 1x1 conv of up_mode "upsample": a plain cast, K contiguous already)
__global__ void __launch_bounds__(256) pack_cast_kernel(const float* __restrict__ w, bf16* __restrict__ out, long n) {
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < n; i += static_cast<long>(gridDim.x) * blockDim.x)
    out[i] = __float2bfloat16(w[i]);
}

// ============================================================================ first conv weight gradient
// dW[co][ci][tap] = sum_p dRaw[p][co] * x[p + tap][ci]   (fp32 NCHW input, Cin = #frequencies, K = 9*Cin <= 72)
// Register-tiled: a thread owns 4 output channels x KPT taps (36-72 accumulators) for one quarter of the tile's
// pixels: per pixel 1 LDS.64 (4 bf16 gradients) + KPT broadcast LDS feed 4*KPT FFMAs.  Partial rows:
// partials[(block*4 + pixel_lane)][64*K].
template <int CIN>
__global__ void __launch_bounds__(256) first_conv_wgrad_kernel(const float* __restrict__ x, View draw, int NB, int H,
                                                               int W, float* partials) {
  constexpr int K = CIN * 9;
  constexpr int KPT = (K + 3) / 4;
  constexpr int XW = TILE_W + 2, XH = TILE_H + 2;
  __shared__ float s_x[CIN * XH * XW];
  __shared__ __align__(16) bf16 s_d[TILE_M][72];
  const int pl = threadIdx.x >> 6, u = threadIdx.x & 63, cg = u & 15, kg = u >> 4;
  int off[KPT];
#pragma unroll
  for (int j = 0; j < KPT; ++j) {
    const int k = kg * KPT + j;
    const int kk = k < K ? k : 0;
    const int ci = kk / 9, tap = kk % 9;
    off[j] = (ci * XH + tap / 3) * XW + tap % 3;
  }
  float acc[4][KPT];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < KPT; ++j) acc[c][j] = 0.f;
  const int tiles_x = (W + TILE_W - 1) / TILE_W, tiles_y = (H + TILE_H - 1) / TILE_H;
  const int total = NB * tiles_x * tiles_y;
  for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
    int t = tile;
    const int tx = t % tiles_x;
    t /= tiles_x;
    const int ty = t % tiles_y;
    const int img = t / tiles_y;
    const int y0 = ty * TILE_H, x0 = tx * TILE_W;
    for (int i = threadIdx.x; i < CIN * XH * XW; i += 256) {
      const int xx = i % XW, yy = (i / XW) % XH, ci = i / (XW * XH);
      const int gy = y0 + yy - 1, gx = x0 + xx - 1;
      s_x[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(&x[((static_cast<long>(img) * CIN + ci) * H + gy) * W + gx]) : 0.f;
    }
    for (int i = threadIdx.x; i < TILE_M * 8; i += 256) {
      const int pix = i >> 3, g = i & 7;
      const int gy = y0 + (pix >> 4), gx = x0 + (pix & 15);
      uint4 v = make_uint4(0, 0, 0, 0);
      if (gy < H && gx < W)
        v = *reinterpret_cast<const uint4*>(draw.ptr + ((static_cast<long>(img) * H + gy) * W + gx) * draw.pitch + g * 8);
      *reinterpret_cast<uint4*>(&s_d[pix][g * 8]) = v;
    }
    __syncthreads();
#pragma unroll 2
    for (int i = 0; i < TILE_M / 4; ++i) {
      const int pix = pl * (TILE_M / 4) + i;
      const uint2 dv = *reinterpret_cast<const uint2*>(&s_d[pix][cg * 4]);
      const float2 d01 = unpack_bf16x2(dv.x), d23 = unpack_bf16x2(dv.y);
      const int pb = (pix >> 4) * XW + (pix & 15);
#pragma unroll
      for (int j = 0; j < KPT; ++j) {
        const float xv = s_x[pb + off[j]];
        acc[0][j] = fmaf(d01.x, xv, acc[0][j]);
        acc[1][j] = fmaf(d01.y, xv, acc[1][j]);
        acc[2][j] = fmaf(d23.x, xv, acc[2][j]);
        acc[3][j] = fmaf(d23.y, xv, acc[3][j]);
      }
    }
    __syncthreads();
  }
  float* dst = partials + (static_cast<long>(blockIdx.x) * 4 + pl) * 64 * K;
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < KPT; ++j)
      if (kg * KPT + j < K) dst[(cg * 4 + c) * K + kg * KPT + j] = acc[c][j];
}

// ============================================================================ weight packing (fp32 params -> bf16 GEMM operands)
// ONE launch per layer kind for the whole network (the parameters change every optimisation step, so packing is on the
// training hot path).  Both kernels go through shared memory so that the fp32 reads and the bf16 writes are coalesced.
// Backward-data needs no second copy: it reads these forward-packed matrices as an MN-major operand (conv_igemm.cu).
//
// conv (Cout,Cin,3,3) -> [Cout][tap][Cin].  Work item = one output channel x up to 256 input channels.
__global__ void __launch_bounds__(256) pack_conv3x3_all_kernel(const __grid_constant__ PackTable t) {
  __shared__ float s[256 * 9];
  int li = 0;
  while (li + 1 < t.n && static_cast<int>(blockIdx.x) >= t.e[li + 1].item0) ++li;
  const PackEntry& L = t.e[li];
  const int chunks = (L.cin + 255) / 256;
  const int item = blockIdx.x - L.item0;
  const int co = item / chunks, ci0 = (item % chunks) * 256;
  const int n = min(256, L.cin - ci0);
  const float* src = L.w + (static_cast<long>(co) * L.cin + ci0) * 9;
  for (int i = threadIdx.x; i < n * 9; i += 256) s[i] = src[i];
  __syncthreads();
  bf16* dst = L.out + static_cast<long>(co) * 9 * L.cin + ci0;
  const int half = n >> 1;  // Cin is a multiple of 64
  for (int j = threadIdx.x; j < 9 * half; j += 256) {
    const int tap = j / half, c = (j - tap * half) * 2;
    *reinterpret_cast<uint32_t*>(dst + static_cast<long>(tap) * L.cin + c) = pack_bf16x2(s[c * 9 + tap], s[(c + 1) * 9 + tap]);
  }
}
// convT (Cin,Cout,2,2) -> [(kk*Cout+co)][Cin].  Work item = 32 input channels x 8 output channels (x 4 taps).
__global__ void __launch_bounds__(256) pack_convt_all_kernel(const __grid_constant__ PackTable t) {
  __shared__ float s[32][33];
  int li = 0;
  while (li + 1 < t.n && static_cast<int>(blockIdx.x) >= t.e[li + 1].item0) ++li;
  const PackEntry& L = t.e[li];  // cin = Cin, cout = Cout
  const int item = blockIdx.x - L.item0;
  const int cgroups = L.cout / 8;
  const int ci0 = (item / cgroups) * 32, co0 = (item % cgroups) * 8;
  const int lane = threadIdx.x & 31, row = threadIdx.x >> 5;
  for (int r = row; r < 32; r += 8)  // 32 consecutive floats (co0..co0+7, kk 0..3) of input channel ci0 + r
    s[r][lane] = L.w[(static_cast<long>(ci0 + r) * L.cout + co0) * 4 + lane];
  __syncthreads();
  for (int j = row; j < 32; j += 8) {  // j = co_local*4 + kk
    const int co = co0 + (j >> 2), kk = j & 3;
    L.out[(static_cast<long>(kk) * L.cout + co) * L.cin + ci0 + lane] = __float2bfloat16(s[lane][j]);
  }
}

inline int grid_for(long work_items, int threads) {
  long b = (work_items + threads - 1) / threads;
  if (b > kMaxBlocks) b = kMaxBlocks;
  if (b < 1) b = 1;
  return static_cast<int>(b);
}

}  // namespace

// ---------------------------------------------------------------------------- launchers (used by net_api.cu / ops_api2.cu)
// The first conv runs on the tensor cores (first_conv_tc.cu) for up to 8 input channels; the fp32 CUDA-core kernels
// below serve 9..12 channels (4 frequencies + up to 7 metadata channels, pipeline.py:392,413-425) and remain the A/B
// reference of the tensor-core path (CRIMAC_FC_CUDACORE=1).
static bool first_conv_use_tc(int cin) {
  static const bool off = getenv("CRIMAC_FC_CUDACORE") != nullptr;
  return !off && cin <= 8;
}
static int first_conv_cc_grid(int NB, int H, int W) {
  const int tiles = NB * ((H + TILE_H - 1) / TILE_H) * ((W + TILE_W - 1) / TILE_W);
  return tiles < 148 * 8 ? tiles : 148 * 8;
}
int first_conv_grid(int NB, int cin, int H, int W) {
  return first_conv_use_tc(cin) ? first_conv_tc_grid(NB, H, W) : first_conv_cc_grid(NB, H, W);
}
size_t first_conv_wgrad_partial_floats(int cin) {
  const size_t cc = static_cast<size_t>(first_conv_wgrad_blocks()) * 4 * 64 * cin * 9;
  const size_t tc = static_cast<size_t>(256) * 2 * 128 * 64;  // >= number of SMs x 2 accumulator groups of [128][64]
  return cc > tc ? cc : tc;
}
cudaError_t launch_first_conv(const float* x, bf16* xs, const float* w, const float* scale, const float* shift, int relu,
                              int NB, int cin, int H, int W, bf16* out, int out_pitch, float* stats, cudaStream_t st) {
  if (first_conv_use_tc(cin))
    return launch_first_conv_tc(x, xs, w, scale, shift, relu, NB, cin, H, W, out, out_pitch, stats, st);
  const int grid = first_conv_cc_grid(NB, H, W);
#define FC(C)                                                                                                   \
  if (cin == C) {                                                                                               \
    first_conv_kernel<C><<<grid, 64, 0, st>>>(x, w, scale, shift, relu, NB, H, W, out, out_pitch, stats);       \
    return cudaGetLastError();                                                                                  \
  }
  FC(1) FC(2) FC(3) FC(4) FC(5) FC(6) FC(7) FC(8) FC(9) FC(10) FC(11) FC(12)
#undef FC
  return cudaErrorInvalidValue;
}

int first_conv_wgrad_blocks() { return 148 * 2; }
cudaError_t launch_first_conv_wgrad(const float* x, const bf16* xs, View draw, int cin, float* partials, float* dw,
                                    int accumulate, cudaStream_t st) {
  if (first_conv_use_tc(cin)) return launch_first_conv_wgrad_tc(xs, draw, cin, partials, dw, accumulate, st);
  const int tiles = draw.N * ((draw.H + TILE_H - 1) / TILE_H) * ((draw.W + TILE_W - 1) / TILE_W);
  const int grid = tiles < first_conv_wgrad_blocks() ? tiles : first_conv_wgrad_blocks();
#define FW(C)                                                                                       \
  if (cin == C) {                                                                                   \
    first_conv_wgrad_kernel<C><<<grid, 256, 0, st>>>(x, draw, draw.N, draw.H, draw.W, partials);    \
    partial_sum_finalize_kernel<1><<<(64 * C * 9 + 7) / 8, 256, 0, st>>>(partials, grid * 4, 64 * C * 9, 1.0, dw, nullptr, accumulate, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr); \
    return cudaGetLastError();                                                                      \
  }
  FW(1) FW(2) FW(3) FW(4) FW(5) FW(6) FW(7) FW(8) FW(9) FW(10) FW(11) FW(12)
#undef FW
  return cudaErrorInvalidValue;
}

cudaError_t launch_bn_finalize(const float* partials, int m_tiles, int C, double count, const float* gamma,
                               const float* beta, float* rm, float* rv, long long* nbt, float momentum, float eps,
                               float* scale, float* shift, float* save_mean, float* save_invstd, cudaStream_t st) {
  bn_finalize_kernel<<<(C + 7) / 8, 256, 0, st>>>(partials, m_tiles, C, count, gamma, beta, rm, rv, nbt, momentum,
                                                    eps, scale, shift, save_mean, save_invstd);
  return cudaGetLastError();
}
cudaError_t launch_bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv,
                                const float* conv_bias, float eps, int C, float* scale, float* shift,
                                cudaStream_t st) {
  bn_fold_eval_kernel<<<(C + 127) / 128, 128, 0, st>>>(gamma, beta, rm, rv, conv_bias, eps, C, scale, shift);
  return cudaGetLastError();
}
cudaError_t launch_bn_apply(View raw, const float* scale, const float* shift, View act, View pool, uint16_t* pool_arg,
                            cudaStream_t st) {
  if (raw.C % 8 != 0 || raw.C / 8 > 256) return cudaErrorInvalidValue;
  if (static_cast<long>(raw.N) * raw.H * raw.W > 0x7fffffffL) return cudaErrorInvalidValue;  // 32-bit pixel index
  const int ppb = 256 / (raw.C / 8);
  if (pool.ptr != nullptr) {
    const long items = static_cast<long>(raw.N) * (raw.H / 2) * (raw.W / 2);
    bn_apply_kernel<true><<<grid_for(items, ppb), 256, 0, st>>>(raw, scale, shift, act, pool, pool_arg);
  } else {
    const long items = static_cast<long>(raw.N) * raw.H * raw.W;
    bn_apply_kernel<false><<<grid_for((items + 3) / 4, ppb), 256, 0, st>>>(raw, scale, shift, act, pool, nullptr);
  }
  return cudaGetLastError();
}
cudaError_t launch_head_fwd(View act, const float* hw, const float* hb, int ncls, float* logits, cudaStream_t st) {
  const long px = static_cast<long>(act.N) * act.H * act.W;
  head_fwd_kernel<<<grid_for(px, 128), 128, 0, st>>>(act, hw, hb, ncls, logits);
  return cudaGetLastError();
}
int ce_blocks() { return 148 * 4; }
cudaError_t launch_ce(const float* logits, const long long* labels, const float* cw, int ncls, int NB, long HW,
                      long long ignore_index, float* dlogits, double* partials, float* out3, cudaStream_t st) {
  const long total = NB * HW;
  int blocks = grid_for(total, 256);
  if (blocks > ce_blocks()) blocks = ce_blocks();
  ce_fwd_bwd_kernel<<<blocks, 256, 0, st>>>(logits, labels, cw, ncls, HW, total, ignore_index, dlogits, partials);
  ce_finalize_kernel<<<1, 1, 0, st>>>(partials, blocks, out3);
  return cudaGetLastError();
}
cudaError_t launch_eval_loss(const float* logits, const void* labels, int label_bits, const float* cw, int ncls, int NB,
                             long HW, int prob_class, float* prob_out, long long* labels_out, double* partials,
                             float* out3, cudaStream_t st) {
  const long total = NB * HW;
  int blocks = grid_for(total, 256);
  if (blocks > ce_blocks()) blocks = ce_blocks();
  if (label_bits == 16)
    eval_loss_kernel<short><<<blocks, 256, 0, st>>>(logits, static_cast<const short*>(labels), cw, ncls, HW, total,
                                                    prob_class, prob_out, labels_out, partials);
  else if (label_bits == 64)
    eval_loss_kernel<long long><<<blocks, 256, 0, st>>>(logits, static_cast<const long long*>(labels), cw, ncls, HW,
                                                        total, prob_class, prob_out, labels_out, partials);
  else
    return cudaErrorInvalidValue;
  ce_finalize_kernel<<<1, 1, 0, st>>>(partials, blocks, out3);
  return cudaGetLastError();
}
int head_bwd_blocks() { return 148 * 4; }
cudaError_t launch_head_bwd(const float* dlogits, const float* gscale, View act, const float* hw, int ncls, View dact,
                            float* partials, float* dw, float* db, int accumulate, cudaStream_t st) {
  const long px = static_cast<long>(act.N) * act.H * act.W;
  if (act.C != 64) return cudaErrorInvalidValue;
  int blocks = grid_for(px, 64);
  if (blocks > head_bwd_blocks()) blocks = head_bwd_blocks();
#define HB(K)                                                                                        \
  if (ncls == K) head_bwd_kernel<K><<<blocks, 256, 0, st>>>(dlogits, gscale, act, hw, dact, partials);
  HB(1) HB(2) HB(3) HB(4) HB(5) HB(6) HB(7) HB(8)
#undef HB
  const int per = ncls * 64 + ncls;
  head_bwd_finalize_kernel<<<(per + 127) / 128, 128, 0, st>>>(partials, blocks, ncls, dw, db, accumulate);
  return cudaGetLastError();
}

cudaError_t launch_head_ce_fused(View act, const float* hw, const float* hb, int ncls, const long long* labels,
                                 const float* cw, long long ignore_index, View dact, float* partials,
                                 double* loss_partials, float* dw, float* db, float* out3, cudaStream_t st) {
  const long px = static_cast<long>(act.N) * act.H * act.W;
  if (act.C != 64) return cudaErrorInvalidValue;
  int blocks = grid_for(px, 64);
  if (blocks > head_bwd_blocks()) blocks = head_bwd_blocks();
#define HC(K)                                                                                                       \
  if (ncls == K)                                                                                                    \
    head_ce_fused_kernel<K><<<blocks, 256, 0, st>>>(act, hw, hb, labels, cw, ignore_index, dact, partials, loss_partials);
  HC(1) HC(2) HC(3) HC(4) HC(5) HC(6) HC(7) HC(8)
#undef HC
  head_ce_finalize_kernel<<<(ncls * 64 + ncls + 7) / 8, 256, 0, st>>>(partials, loss_partials, blocks, ncls, dw, db, out3);
  return cudaGetLastError();
}

int reduce_blocks() { return 148 * 4; }  // 4 blocks of 256 threads per SM
static int reduce_grid(const View& v) {
  const int ppb = 256 / (v.C / 8);
  const long px = static_cast<long>(v.N) * v.H * v.W;
  long b = (px + ppb - 1) / ppb;
  if (b > reduce_blocks()) b = reduce_blocks();
  return static_cast<int>(b < 1 ? 1 : b);
}
// BN+ReLU backward: dact (grad wrt post-ReLU activation) -> draw (grad wrt conv output), dgamma/dbeta/dbias.
// scratch: partials, at least reduce_blocks()*2*C floats; c1c2: 2*C floats.
cudaError_t launch_bn_bwd(View dact, View raw, const float* scale, const float* shift, const float* mean,
                          const float* invstd, View draw, float* dgamma, float* dbeta, float* dbias, int accumulate,
                          float* partials, float* c1c2, const float* gscale, cudaStream_t st, const uint16_t* pool_arg,
                          View dpool, View dskip) {
  const int C = raw.C;
  if (C % 8 != 0 || C / 8 > 256) return cudaErrorInvalidValue;
  const int grid = reduce_grid(raw);
  // the reduce kernel fits three blocks per SM: one full wave (a 592-block grid would leave a 1/3-occupancy tail wave)
  int grid_r = grid < 148 * 3 ? grid : 148 * 3;
  const int ppb = 256 / (C / 8);
  const double count = static_cast<double>(raw.N) * raw.H * raw.W;
  if (pool_arg != nullptr) {
    // dA = dSkip + unpool(dPool) formed on the fly, one thread per pooling window
    if (count > 4.0e9 || dskip.C != C || dpool.C != C || (raw.H & 1) || (raw.W & 1)) return cudaErrorInvalidValue;
    const long windows = static_cast<long>(raw.N) * (raw.H / 2) * (raw.W / 2);
    long gw = (windows + ppb - 1) / ppb;
    if (gw > 148 * 2) gw = 148 * 2;   // two resident blocks per SM: one full wave
    if (gw < 1) gw = 1;
    grid_r = static_cast<int>(gw);
    bn_bwd_reduce_pool_kernel<<<grid_r, 256, ppb * 2 * C * sizeof(float), st>>>(pool_arg, dpool, dskip, raw, scale, shift, partials);
  } else {
    bn_bwd_reduce_kernel<<<grid_r, 256, ppb * 2 * C * sizeof(float), st>>>(dact, raw, scale, shift, partials);
  }
  partial_sum_finalize_kernel<2><<<(C + 7) / 8, 256, 0, st>>>(partials, grid_r, C, count, dbeta,
                                                                dgamma, accumulate, c1c2, c1c2 + C, gscale, mean, invstd,
                                                                dbias);
  if (pool_arg != nullptr) {
    const long windows = static_cast<long>(raw.N) * (raw.H / 2) * (raw.W / 2);
    bn_bwd_apply_pool_kernel<<<grid_for(windows, ppb), 256, 0, st>>>(pool_arg, dpool, dskip, raw, scale, shift, mean, invstd,
                                                                     c1c2, c1c2 + C, gscale, draw);
  } else {
    bn_bwd_apply_kernel<<<grid, 256, 0, st>>>(dact, raw, scale, shift, mean, invstd, c1c2, c1c2 + C, gscale, draw);
  }
  return cudaGetLastError();
}
cudaError_t launch_view_colsum(View v, float* partials, float* out, int accumulate, cudaStream_t st) {
  const int C = v.C;
  if (C % 8 != 0 || C / 8 > 256) return cudaErrorInvalidValue;
  const int grid = reduce_grid(v);
  const int ppb = 256 / (C / 8);
  view_colsum_kernel<<<grid, 256, ppb * C * sizeof(float), st>>>(v, partials);
  partial_sum_finalize_kernel<1><<<(C + 7) / 8, 256, 0, st>>>(partials, grid, C, 1.0, out, nullptr, accumulate, nullptr,
                                                                nullptr, nullptr, nullptr, nullptr, nullptr);
  return cudaGetLastError();
}
cudaError_t launch_pool_bwd_add(const uint16_t* pool_arg, View dpool, View dskip, View dact, cudaStream_t st) {
  const long items = static_cast<long>(dact.N) * (dact.H / 2) * (dact.W / 2) * (dact.C / 8);
  pool_bwd_add_kernel<<<grid_for(items, 256), 256, 0, st>>>(pool_arg, dpool, dskip, dact);
  return cudaGetLastError();
}
cudaError_t launch_upsample2x(View in, View out, cudaStream_t st) {
  if (in.C % 8 != 0 || out.C != in.C || out.H != 2 * in.H || out.W != 2 * in.W) return cudaErrorInvalidValue;
  const long items = static_cast<long>(in.N) * out.H * out.W * (in.C / 8);
  upsample2x_kernel<<<grid_for(items, 256), 256, 0, st>>>(in, out);
  return cudaGetLastError();
}
cudaError_t launch_upsample2x_bwd(View dout, View din, cudaStream_t st) {
  if (din.C % 8 != 0 || dout.C != din.C || dout.H != 2 * din.H || dout.W != 2 * din.W) return cudaErrorInvalidValue;
  const long items = static_cast<long>(din.N) * din.H * din.W * (din.C / 8);
  upsample2x_bwd_kernel<<<grid_for(items, 256), 256, 0, st>>>(dout, din);
  return cudaGetLastError();
}
cudaError_t launch_pack_cast(const float* w, bf16* out, long n, cudaStream_t st) {
  pack_cast_kernel<<<grid_for(n, 256), 256, 0, st>>>(w, out, n);
  return cudaGetLastError();
}
cudaError_t launch_pack_conv3x3_all(PackTable& t, cudaStream_t st) {
  int items = 0;
  for (int i = 0; i < t.n; ++i) {
    t.e[i].item0 = items;
    items += t.e[i].cout * ((t.e[i].cin + 255) / 256);
  }
  if (items > 0) pack_conv3x3_all_kernel<<<items, 256, 0, st>>>(t);
  return cudaGetLastError();
}
cudaError_t launch_pack_convt_all(PackTable& t, cudaStream_t st) {
  int items = 0;
  for (int i = 0; i < t.n; ++i) {
    if (t.e[i].cin % 32 != 0 || t.e[i].cout % 8 != 0) return cudaErrorInvalidValue;
    t.e[i].item0 = items;
    items += (t.e[i].cin / 32) * (t.e[i].cout / 8);
  }
  if (items > 0) pack_convt_all_kernel<<<items, 256, 0, st>>>(t);
  return cudaGetLastError();
}

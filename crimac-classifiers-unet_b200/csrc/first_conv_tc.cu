// First conv of the echogram U-Net (down_convs.0.main.0, reference models/unet.py:76: nn.Conv2d(#frequencies, 64, 3,
// padding=1)) and its weight gradient on the tensor cores, with the im2col done by TMA.
//
// The input is the fp32 NCHW dB echogram (values in [-75, 0]: bf16 alone would lose 0.25-0.5 dB).  A tiny pre-pass
// (split_input_kernel) rewrites it ONCE per step as NHWC bf16 pairs  hi = bf16(x), lo = bf16(x - hi)  (16 mantissa bits
// of x) in 16-byte chunks per pixel: one plane [pixel][hi c0..3 | lo c0..3] for <= 4 frequencies (P = 4), else two planes
// [pixel][hi c0..7], [pixel][lo c0..7] (P = 8).
//
// One TMA box {16 pings x 8 elements (256 contiguous bytes), 8 rows} of a plane at a tap-shifted coordinate is then exactly 16 UN-swizzled
// UMMA core matrices (8 pixels x 16 B, 128 B apart): the nine taps are nine (P = 8: eighteen) boxes, 2 KB apart, and a
// 16-wide k-slice of the GEMM is two consecutive boxes (descriptor LBO = 2048, SBO = 128; probed on B200 with
// tools/gpu_probe_noswizzle.py).  Out-of-image taps are zero-filled by TMA = the conv's zero padding.  No thread touches
// the operand: the builder warps of the previous version (bound by load latency + conversion ALU) are gone.
//
//   forward : M = 128 pixels, N = 64, k' = (tap, hi|lo, c).  D = A*B1 + A*B2 with B1 = W_hi at the hi AND lo positions
//             (x_hi*W_hi + x_lo*W_hi) and B2 = W_lo at the hi positions only (x_hi*W_lo); the dropped x_lo*W_lo term is
//             2^-16 relative.  Epilogue as conv_igemm: +bias | scale/shift+ReLU -> bf16 NHWC, BatchNorm statistics in
//             registers per CTA.
//   wgrad   : D[k'][co] = sum_p A[p][k'] * dRaw[p][co]: the same boxes read MN-major (LBO = 128, SBO = 2048), B = the dRaw
//             tile (SW128 TMA box); the fp32 accumulator stays in TMEM for the CTA's lifetime; a fold kernel adds the hi
//             and lo rows and the per-CTA partials.
#include "host_util.h"
#include "ptx.cuh"
#include "devfn.cuh"

namespace {

constexpr int BOX_BYTES = 128 * 16;  // one tap box: 128 pixels x 8 bf16
constexpr int FC_THREADS = 384;      // w0 TMA, w1 MMA, w2 TMEM alloc, w4-11 epilogue

template <int CIN>
struct FcCfg {
  static constexpr int P = (CIN <= 4) ? 4 : 8;       // padded channels per half
  static constexpr int CT = P / 4;                    // 16-byte chunks per tap (hi|lo in one chunk, or hi chunk + lo chunk)
  static constexpr int NCH = 9 * CT;                  // chunks (= TMA boxes) per tile
  static constexpr int KS = (NCH + 1) / 2;            // 16-wide k-slices
  static constexpr int A_STAGE = 2 * KS * BOX_BYTES;  // forward: NCH boxes + a zero pad box when NCH is odd
  static constexpr int WG_GROUPS = (NCH + 15) / 16;   // wgrad accumulators (128 k' rows = 16 boxes each)
  static constexpr int A_STAGE_WG = WG_GROUPS * 16 * BOX_BYTES;
  static constexpr int B_BYTES = (2 * KS) * 1024;     // weights: 2*KS chunks x [64 co][16 B]
};

__device__ __forceinline__ void epi_bar256() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// fp32 NCHW -> [pixel][hi c0..P-1 | lo c0..P-1] bf16 (channels >= CIN are zero)
template <int CIN>
__global__ void __launch_bounds__(256) split_input_kernel(const float* __restrict__ x, long npix_per_img, int NB,
                                                          bf16* __restrict__ xs) {
  constexpr int P = FcCfg<CIN>::P;
  const long total = static_cast<long>(NB) * npix_per_img;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const long n = i / npix_per_img, r = i - n * npix_per_img;
    float hi[P], lo[P];
#pragma unroll
    for (int c = 0; c < P; ++c) {
      const float v = c < CIN ? __ldg(x + (n * CIN + c) * npix_per_img + r) : 0.f;
      const float h = __bfloat162float(__float2bfloat16(v));
      hi[c] = h;
      lo[c] = v - h;
    }
    uint32_t w[P];  // P/2 words of hi then P/2 words of lo
#pragma unroll
    for (int c = 0; c < P / 2; ++c) {
      w[c] = pack_bf16x2(hi[2 * c], hi[2 * c + 1]);
      w[P / 2 + c] = pack_bf16x2(lo[2 * c], lo[2 * c + 1]);
    }
    // planes [part][pixel][8]: P = 4: one plane (hi4 | lo4); P = 8: plane 0 = hi8, plane 1 = lo8
    *reinterpret_cast<uint4*>(xs + i * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    if (P == 8) *reinterpret_cast<uint4*>(xs + (total + i) * 8) = make_uint4(w[P - 4], w[P - 3], w[P - 2], w[P - 1]);
  }
}

// ---------------------------------------------------------------------------------------------------------------- forward
template <int CIN>
__global__ void __launch_bounds__(FC_THREADS, 1)
first_conv_tc_kernel(const __grid_constant__ CUtensorMap x_map, const float* __restrict__ w,
                     const float* __restrict__ scale, const float* __restrict__ shift, int relu, int NB, int H, int W,
                     bf16* __restrict__ out, int out_pitch, float* stats) {
  using Cfg = FcCfg<CIN>;
  constexpr int P = Cfg::P, CT = Cfg::CT, NCH = Cfg::NCH, KS = Cfg::KS;
  constexpr int STAGES = 4;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* a_ring = smem;                                   // [STAGES][2*KS boxes]
  uint8_t* b1 = a_ring + STAGES * Cfg::A_STAGE;             // W_hi at hi and lo positions
  uint8_t* b2 = b1 + Cfg::B_BYTES;                          // W_lo at hi positions
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(b2 + Cfg::B_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_sc = reinterpret_cast<float*>(tmem_ptr + 4);     // [64]
  float* s_sh = s_sc + 64;                                  // [64]
  float* s_red = s_sh + 64;                                 // [4 quarters][2][64]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&x_map);
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 256);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr, 128);
    ptx::tmem_relinquish();
  }
  // operand ring zeroed once: the pad box behind an odd number of chunks must read as zeros
  for (int i = threadIdx.x; i < (STAGES * Cfg::A_STAGE + 2 * Cfg::B_BYTES) / 16; i += FC_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x < 64) {
    s_sc[threadIdx.x] = scale ? scale[threadIdx.x] : 1.f;
    s_sh[threadIdx.x] = shift ? shift[threadIdx.x] : 0.f;
  }
  __syncthreads();
  // weights (co, c, tap) fp32 -> un-swizzled K-major core matrices: chunk j at j*1024, co group g at g*128, row co%8
  for (int i = threadIdx.x; i < 64 * NCH; i += FC_THREADS) {
    const int co = i & 63, j = i >> 6;
    const int tap = j / CT, part = j % CT;  // P = 8: part 0 = hi chunk, part 1 = lo chunk
    uint32_t h1[4], h2[4];
#pragma unroll
    for (int e2 = 0; e2 < 4; ++e2) {
      float wv[2], lv[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int e = 2 * e2 + u;
        const int c = (P == 4) ? (e & 3) : e;
        const bool is_lo_pos = (P == 4) ? (e >= 4) : (part == 1);
        const float v = c < CIN ? w[(co * CIN + c) * 9 + tap] : 0.f;
        const float hv = __bfloat162float(__float2bfloat16(v));
        wv[u] = hv;                           // W_hi multiplies x_hi and x_lo
        lv[u] = is_lo_pos ? 0.f : (v - hv);   // W_lo multiplies x_hi only
      }
      h1[e2] = pack_bf16x2(wv[0], wv[1]);
      h2[e2] = pack_bf16x2(lv[0], lv[1]);
    }
    const int off = j * 1024 + (co >> 3) * 128 + (co & 7) * 16;
    *reinterpret_cast<uint4*>(b1 + off) = make_uint4(h1[0], h1[1], h1[2], h1[3]);
    *reinterpret_cast<uint4*>(b2 + off) = make_uint4(h2[0], h2[1], h2[2], h2[3]);
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int tiles_x = (W + TILE_W - 1) / TILE_W, tiles_y = (H + TILE_H - 1) / TILE_H;
  const int total = NB * tiles_x * tiles_y;

  if (warp == 0) {
    if (ptx::elect_one()) {
      // ===================== TMA producer: NCH tap boxes per tile =====================
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int t = tile;
        const int tx = t % tiles_x;
        t /= tiles_x;
        const int ty = t % tiles_y;
        const int img = t / tiles_y;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
        ptx::mbar_arrive_expect_tx(&full_bar[stage], NCH * BOX_BYTES);
        uint8_t* sa = a_ring + stage * Cfg::A_STAGE;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const int tap = j / CT, part = j % CT;
          ptx::tma_load_4d(sa + j * BOX_BYTES, &x_map, &full_bar[stage], (tx * TILE_W + tap % 3 - 1) * 8,
                           ty * TILE_H + tap / 3 - 1, img, part);
        }
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = ptx::make_idesc_bf16(128, 64, 0, 0);
      // un-swizzled K-major: LBO = K-direction stride between 16-byte chunks, SBO = stride between 8-row groups
      const uint64_t adesc0 = ptx::make_smem_desc(0, BOX_BYTES, 128, 0, 0);
      const uint64_t b1d = ptx::make_smem_desc(ptx::smem_u32(b1), 1024, 128, 0, 0);
      const uint64_t b2d = ptx::make_smem_desc(ptx::smem_u32(b2), 1024, 128, 0, 0);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        const int as = it & 1;
        ptx::mbar_wait(&tmem_empty[as], ((it >> 1) & 1) ^ 1u);
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * 64;
        const uint64_t adesc = adesc0 + (ptx::smem_u32(a_ring + stage * Cfg::A_STAGE) >> 4);
#pragma unroll
        for (int k = 0; k < KS; ++k)  // one k-slice = two chunks: +4096 B in A, +2048 B in B
          ptx::umma_bf16(d_tmem, adesc + k * (2 * BOX_BYTES >> 4), b1d + k * (2048 >> 4), idesc, k != 0);
#pragma unroll
        for (int k = 0; k < KS; ++k)
          ptx::umma_bf16(d_tmem, adesc + k * (2 * BOX_BYTES >> 4), b2d + k * (2048 >> 4), idesc, 1);
        ptx::umma_commit(&empty_bar[stage]);
        ptx::umma_commit(&tmem_full[as]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int q = warp & 3, hc = (warp - 4) >> 2;  // TMEM lane quarter / 32-column half of this warp
    const int r = q * 32 + lane;
    const int e = threadIdx.x - 128;
    float r1[32], r2[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r1[j] = r2[j] = 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int as = it & 1;
      int t = tile;
      const int tx = t % tiles_x;
      t /= tiles_x;
      const int ty = t % tiles_y;
      const int img = t / tiles_y;
      const int y = ty * TILE_H + (r >> 4), xx = tx * TILE_W + (r & 15);
      const bool valid = y < H && xx < W;
      ptx::mbar_wait(&tmem_full[as], (it >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t v[32];
      ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * 64 + hc * 32, v);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[as]);  // accumulator is in registers: the stage may be overwritten
      const float4* sc4 = reinterpret_cast<const float4*>(s_sc + hc * 32);
      const float4* sh4 = reinterpret_cast<const float4*>(s_sh + hc * 32);
      uint32_t pk[16];
#pragma unroll
      for (int j4 = 0; j4 < 8; ++j4) {
        const float4 a = sc4[j4], b = sh4[j4];
        float f0 = fmaf(__uint_as_float(v[4 * j4 + 0]), a.x, b.x), f1 = fmaf(__uint_as_float(v[4 * j4 + 1]), a.y, b.y);
        float f2 = fmaf(__uint_as_float(v[4 * j4 + 2]), a.z, b.z), f3 = fmaf(__uint_as_float(v[4 * j4 + 3]), a.w, b.w);
        if (relu) {
          f0 = fmaxf(f0, 0.f);
          f1 = fmaxf(f1, 0.f);
          f2 = fmaxf(f2, 0.f);
          f3 = fmaxf(f3, 0.f);
        }
        pk[2 * j4] = pack_bf16x2(f0, f1);
        pk[2 * j4 + 1] = pack_bf16x2(f2, f3);
      }
      if (valid) {
        bf16* dst = out + ((static_cast<long>(img) * H + y) * W + xx) * out_pitch + hc * 32;
        store32(dst, pk);
        store32(dst + 16, pk + 8);
      }
      if (stats != nullptr) {
        // statistics of the bf16-rounded values the BN-apply pass reads back; per-thread until the CTA is done
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float2 u = unpack_bf16x2(valid ? pk[j] : 0u);
          r1[2 * j] += u.x;
          r1[2 * j + 1] += u.y;
          r2[2 * j] = fmaf(u.x, u.x, r2[2 * j]);
          r2[2 * j + 1] = fmaf(u.y, u.y, r2[2 * j + 1]);
        }
      }
    }
    if (stats != nullptr) {
      xpose_reduce(r1, lane);
      xpose_reduce(r2, lane);
      s_red[(q * 2 + 0) * 64 + hc * 32 + lane] = r1[0];
      s_red[(q * 2 + 1) * 64 + hc * 32 + lane] = r2[0];
      epi_bar256();
      if (e < 128) {
        const int which = e >> 6, c = e & 63;
        float a = 0.f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) a += s_red[(qq * 2 + which) * 64 + c];
        stats[static_cast<long>(blockIdx.x) * 128 + which * 64 + c] = a;
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 128);
  }
}
template <int CIN>
constexpr int fc_fwd_smem() {
  return 1024 + 4 * FcCfg<CIN>::A_STAGE + 2 * FcCfg<CIN>::B_BYTES + 256 + (2 * 64 + 8 * 64) * 4;
}

// ---------------------------------------------------------------------------------------------------------------- wgrad
// w0 TMA, w1 MMA, w2 TMEM alloc, w4-7 read the accumulator out once at the end.
constexpr int FC_WG_THREADS = 256;
template <int CIN>
__global__ void __launch_bounds__(FC_WG_THREADS, 1)
first_conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap x_map, const __grid_constant__ CUtensorMap d_map, int NB,
                           int H, int W, float* partials) {
  using Cfg = FcCfg<CIN>;
  constexpr int CT = Cfg::CT, NCH = Cfg::NCH, G = Cfg::WG_GROUPS;
  constexpr int STAGES = (G == 1) ? 3 : 2;
  constexpr int D_TILE = 128 * 128;  // dRaw tile: 128 pixels x 64 channels bf16, SW128
  constexpr int STAGE = Cfg::A_STAGE_WG + D_TILE;
  constexpr int TMEM_COLS = (G == 1) ? 64 : 128;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* acc_bar = empty_bar + STAGES;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_bar + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&x_map);
    ptx::prefetch_tmap(&d_map);
    for (int i = 0; i < STAGES; ++i) {
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);
    }
    ptx::mbar_init(acc_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr, TMEM_COLS);
    ptx::tmem_relinquish();
  }
  // the MMA reads 16 boxes per accumulator group, TMA only writes NCH of them: the rest must be finite (zeros)
  for (int i = threadIdx.x; i < STAGES * STAGE / 16; i += FC_WG_THREADS)
    reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int tiles_x = (W + TILE_W - 1) / TILE_W, tiles_y = (H + TILE_H - 1) / TILE_H;
  const int total = NB * tiles_x * tiles_y;

  if (warp == 0) {
    if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        int t = tile;
        const int tx = t % tiles_x;
        t /= tiles_x;
        const int ty = t % tiles_y;
        const int img = t / tiles_y;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
        ptx::mbar_arrive_expect_tx(&full_bar[stage], NCH * BOX_BYTES + D_TILE);
        uint8_t* sa = smem + stage * STAGE;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
          const int tap = j / CT, part = j % CT;
          ptx::tma_load_4d(sa + j * BOX_BYTES, &x_map, &full_bar[stage], (tx * TILE_W + tap % 3 - 1) * 8,
                           ty * TILE_H + tap / 3 - 1, img, part);
        }
        ptx::tma_load_4d(sa + Cfg::A_STAGE_WG, &d_map, &full_bar[stage], 0, tx * TILE_W, ty * TILE_H, img);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    if (ptx::elect_one()) {
      // A: un-swizzled MN-major (M = k' rows: 8 per box, boxes 2048 B apart = SBO; K = pixels: 8-pixel groups 128 B apart
      // = LBO; 16 pixels per MMA = +256 B).  B: dRaw tile, SW128 MN-major as in wgrad_gemm (+2048 B per 16 pixels).
      const uint32_t idesc = ptx::make_idesc_bf16(128, 64, 1, 1);
      const uint64_t adesc0 = ptx::make_smem_desc(0, 128, BOX_BYTES, 0, 0);
      const uint64_t bdesc0 = ptx::make_smem_desc(0, 8192, 1024);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sa16 = ptx::smem_u32(smem + stage * STAGE) >> 4;
        const uint64_t bdesc = bdesc0 + (sa16 + (Cfg::A_STAGE_WG >> 4));
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const uint64_t adesc = adesc0 + (sa16 + g * (16 * BOX_BYTES >> 4));
#pragma unroll
          for (int k = 0; k < 8; ++k)
            ptx::umma_bf16(tmem_base + g * 64, adesc + k * (256 >> 4), bdesc + k * (2048 >> 4), idesc, (it | k) != 0);
        }
        ptx::umma_commit(&empty_bar[stage]);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      ptx::umma_commit(acc_bar);
    }
  } else if (warp >= 4) {
    const int q = warp & 3;
    ptx::mbar_wait(acc_bar, 0);
    ptx::tc_fence_after();
#pragma unroll 1
    for (int c = 0; c < 2 * G; ++c) {  // 32-column chunks: group c/2, half c%2
      uint32_t v[32];
      ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + c * 32, v);
      ptx::tmem_ld_wait();
      float* dst = partials + ((static_cast<long>(blockIdx.x) * G + c / 2) * 128 + q * 32 + lane) * 64 + (c & 1) * 32;
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                          __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, TMEM_COLS);
  }
}
template <int CIN>
constexpr int fc_wg_smem() {
  return 1024 + (FcCfg<CIN>::WG_GROUPS == 1 ? 3 : 2) * (FcCfg<CIN>::A_STAGE_WG + 128 * 128) + 256;
}

// partials [ncta][G*128 k' rows][64 co] -> dw[co][c][tap] = hi row + lo row.  Block = one (c, tap): 64 co x 4 CTA-lanes.
template <int CIN>
__global__ void __launch_bounds__(256) first_conv_wgrad_fold_kernel(const float* __restrict__ partials, int ncta,
                                                                    float* dw, int accumulate) {
  using Cfg = FcCfg<CIN>;
  constexpr int P = Cfg::P, G = Cfg::WG_GROUPS;
  __shared__ float s[4][64];
  const int c = blockIdx.x / 9, tap = blockIdx.x % 9;
  const int co = threadIdx.x & 63, ln = threadIdx.x >> 6;
  // k' row of (tap, hi|lo, c): P = 4: tap*8 + {0,4} + c ; P = 8: (tap*2 + {0,1})*8 + c
  const int row_hi = (P == 4) ? tap * 8 + c : (tap * 2) * 8 + c;
  const int row_lo = (P == 4) ? tap * 8 + 4 + c : (tap * 2 + 1) * 8 + c;
  float a = 0.f;
  for (int b = ln; b < ncta; b += 4) {
    const float* base = partials + static_cast<long>(b) * G * 128 * 64;
    a += base[row_hi * 64 + co] + base[row_lo * 64 + co];
  }
  s[ln][co] = a;
  __syncthreads();
  if (ln == 0) {
    const float g = (s[0][co] + s[1][co]) + (s[2][co] + s[3][co]);
    float* d = dw + (co * CIN + c) * 9 + tap;
    *d = accumulate ? *d + g : g;
  }
}

int x_map_for(CUtensorMap* map, bf16* xs, int NB, int H, int W, int P) {
  return make_split_input_map(map, xs, P / 4, NB, H, W);
}

}  // namespace

int first_conv_tc_grid(int NB, int H, int W) {
  const int tiles = NB * ((H + TILE_H - 1) / TILE_H) * ((W + TILE_W - 1) / TILE_W);
  const int sms = device_num_sms();
  return tiles < sms ? tiles : sms;
}
size_t first_conv_split_elems(int NB, int cin, int H, int W) {
  return static_cast<size_t>(NB) * H * W * 2 * (cin <= 4 ? 4 : 8);
}

// cin <= 8.  xs: scratch of first_conv_split_elems() bf16 (written here, reused by the weight gradient of the same
// input); stats: [grid][2][64] partial rows
cudaError_t launch_first_conv_tc(const float* x, bf16* xs, const float* w, const float* scale, const float* shift,
                                 int relu, int NB, int cin, int H, int W, bf16* out, int out_pitch, float* stats,
                                 cudaStream_t st) {
  const int grid = first_conv_tc_grid(NB, H, W);
  const long npix = static_cast<long>(H) * W;
#define FC(C)                                                                                                      \
  if (cin == C) {                                                                                                  \
    if (cudaError_t e = ensure_dynamic_smem(first_conv_tc_kernel<C>, fc_fwd_smem<C>()); e != cudaSuccess) return e; \
    long blocks = (NB * npix + 255) / 256;                                                                         \
    if (blocks > 148 * 8) blocks = 148 * 8;                                                                        \
    if (x != nullptr) /* else: xs was written by crimac_preprocess_staged */                                        \
      split_input_kernel<C><<<static_cast<int>(blocks), 256, 0, st>>>(x, npix, NB, xs);                            \
    CUtensorMap map;                                                                                               \
    if (x_map_for(&map, xs, NB, H, W, FcCfg<C>::P) != 0) return cudaErrorInvalidValue;                             \
    first_conv_tc_kernel<C><<<grid, FC_THREADS, fc_fwd_smem<C>(), st>>>(map, w, scale, shift, relu, NB, H, W, out,  \
                                                                        out_pitch, stats);                         \
    return cudaGetLastError();                                                                                     \
  }
  FC(1) FC(2) FC(3) FC(4) FC(5) FC(6) FC(7) FC(8)
#undef FC
  return cudaErrorInvalidValue;
}

// partials: at least first_conv_tc_grid() * 2 * 128 * 64 floats.  xs: the split input written by the matching forward.
cudaError_t launch_first_conv_wgrad_tc(const bf16* xs, View draw, int cin, float* partials, float* dw, int accumulate,
                                       cudaStream_t st) {
  if (draw.C != 64) return cudaErrorInvalidValue;
  CUtensorMap dmap;
  if (make_act_map(&dmap, draw, TILE_H) != 0) return cudaErrorInvalidValue;
  const int grid = first_conv_tc_grid(draw.N, draw.H, draw.W);
#define FW(C)                                                                                                      \
  if (cin == C) {                                                                                                  \
    if (cudaError_t e = ensure_dynamic_smem(first_conv_wgrad_tc_kernel<C>, fc_wg_smem<C>()); e != cudaSuccess)     \
      return e;                                                                                                    \
    CUtensorMap xmap;                                                                                              \
    if (x_map_for(&xmap, const_cast<bf16*>(xs), draw.N, draw.H, draw.W, FcCfg<C>::P) != 0)                         \
      return cudaErrorInvalidValue;                                                                                \
    first_conv_wgrad_tc_kernel<C><<<grid, FC_WG_THREADS, fc_wg_smem<C>(), st>>>(xmap, dmap, draw.N, draw.H, draw.W, \
                                                                                partials);                         \
    first_conv_wgrad_fold_kernel<C><<<C * 9, 256, 0, st>>>(partials, grid, dw, accumulate);                        \
    return cudaGetLastError();                                                                                     \
  }
  FW(1) FW(2) FW(3) FW(4) FW(5) FW(6) FW(7) FW(8)
#undef FW
  return cudaErrorInvalidValue;
}

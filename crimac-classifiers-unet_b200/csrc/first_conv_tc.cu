// First conv of the echogram U-Net (down_convs.0.main.0, reference models/unet.py:76: nn.Conv2d(#frequencies, 64, 3,
// padding=1)) and its weight gradient on the tensor cores.
//
// The input is the fp32 NCHW dB echogram (values in [-75, 0]: bf16 alone would lose 0.25-0.5 dB), K = 9*Cin <= 63 is one
// 64-wide k-block.  Both kernels build the im2col tile [128 pixels][64 k] in shared memory themselves (no TMA: the
// source is fp32 NCHW) in the 128B-swizzled layout tcgen05.mma reads, as TWO bf16 tiles: hi = bf16(x), lo = bf16(x - hi),
// i.e. 16 mantissa bits of x.  Forward also splits the weights, D = x_hi*W_hi + x_lo*W_hi + x_hi*W_lo (the dropped
// x_lo*W_lo term is 2^-16 relative), so the result matches the fp32 CUDA-core kernel it replaces to ~1e-5 relative
// while running ~4x faster (that kernel was FFMA/LSU bound at 0.5 ms per batch of 32).
//
//   forward : M = 128 pixels (8 rows x 16 pings), N = 64 channels, K = 3 x 16*ceil(9*Cin/16); epilogue as conv_igemm:
//             +bias | scale/shift+ReLU -> bf16 NHWC, train-mode BatchNorm statistics kept in registers per CTA.
//   wgrad   : dW[co][k] = sum_p dRaw[p][co] * im2col(x)[p][k]: M = 128 = (k of x_hi | k of x_lo), N = 64 channels,
//             K = pixels; A = the same two im2col tiles read MN-major, B = the dRaw tile (TMA box, MN-major); the fp32
//             accumulator stays in TMEM for the CTA's lifetime, one partial [128][64] per CTA, folded by a finalize.
#include "host_util.h"
#include "ptx.cuh"
#include "devfn.cuh"

namespace {

constexpr int FC_A_TILE = 128 * 128;  // 128 pixels x 64 k bf16
constexpr int FC_B_TILE = 64 * 128;   // 64 channels x 64 k bf16

// byte offset of 16-byte chunk c (8 bf16) of 128-byte row r in a SWIZZLE_128B tile whose base is 1024-byte aligned
__device__ __forceinline__ uint32_t sw128(int r, int c) { return static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void split_store8(uint8_t* hi_tile, uint8_t* lo_tile, uint32_t off, const float (&v)[8]) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const __nv_bfloat162 hh = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
    const float2 hf = __bfloat1622float2(hh);
    h[j] = *reinterpret_cast<const uint32_t*>(&hh);
    l[j] = pack_bf16x2(v[2 * j] - hf.x, v[2 * j + 1] - hf.y);
  }
  *reinterpret_cast<uint4*>(hi_tile + off) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(lo_tile + off) = make_uint4(l[0], l[1], l[2], l[3]);
}

// im2col of one pixel row of an 8x16-pixel tile: row r = pixel (r>>4, r&15), k = ci*9 + ky*3 + kx.  One thread builds
// the whole 128-byte row (all its loads are independent and in flight together).  Out-of-image taps are zero (= the
// conv's zero padding).
template <int CIN>
__device__ __forceinline__ void build_im2col_row(const float* __restrict__ x, int img, int y0, int x0, int H, int W,
                                                 int r, uint8_t* a_hi, uint8_t* a_lo) {
  constexpr int K = CIN * 9;
  constexpr int NCHUNK = 2 * ((K + 15) / 16);
  const int y = y0 + (r >> 4), xx = x0 + (r & 15);
  const float* xi = x + static_cast<long>(img) * CIN * H * W;
  float v[NCHUNK][8];
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = c * 8 + j;
      v[c][j] = 0.f;
      if (k < K) {
        const int ci = k / 9, tap = k % 9;
        const int yy = y + tap / 3 - 1, xq = xx + tap % 3 - 1;
        const bool ok = yy >= 0 && yy < H && xq >= 0 && xq < W;
        const float* src = xi + (static_cast<long>(ci) * H + (ok ? yy : 0)) * W + (ok ? xq : 0);
        const float t = __ldg(src);  // unconditional load of a clamped address: no branch, loads stay independent
        v[c][j] = ok ? t : 0.f;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < NCHUNK; ++c) split_store8(a_hi, a_lo, sw128(r, c), v[c]);
}

__device__ __forceinline__ void named_bar(int id, int threads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

constexpr int FC_A_BUF = 2 * FC_A_TILE;  // hi | lo
// shared memory: [A buf 0 | A buf 1 | B region 32 KB | barriers | floats]
//   forward: B region = W_hi (8 KB) | W_lo (8 KB);  wgrad: B region = dRaw tile 0 (16 KB) | dRaw tile 1 (16 KB)
constexpr int FC_SMEM_BYTES = 1024 + 2 * FC_A_BUF + 2 * FC_A_TILE + 128 + (2 * 64 + 8 * 64) * 4;
struct FcSmem {
  uint8_t* a[2];
  uint8_t* b;
  uint64_t* bars;
  uint32_t* tmem_ptr;
  float* fl;
};
__device__ __forceinline__ FcSmem carve(uint8_t* raw) {
  uint8_t* base = raw + ((1024u - (ptx::smem_u32(raw) & 1023u)) & 1023u);
  FcSmem s;
  s.a[0] = base;
  s.a[1] = base + FC_A_BUF;
  s.b = base + 2 * FC_A_BUF;
  s.bars = reinterpret_cast<uint64_t*>(s.b + 2 * FC_A_TILE);
  s.tmem_ptr = reinterpret_cast<uint32_t*>(s.bars + 8);
  s.fl = reinterpret_cast<float*>(s.bars + 16);
  return s;
}

// ---------------------------------------------------------------------------------------------------------------- forward
// 12 warps: w0-3 build the im2col tiles (one pixel row per thread, double-buffered) and w0's elected lane issues the
// MMAs; w4-11 are the epilogue (TMEM lane quarter w%4, 32-column half (w-4)/4) on two TMEM accumulator stages.
constexpr int FC_FWD_THREADS = 384;
template <int CIN>
__global__ void __launch_bounds__(FC_FWD_THREADS, 1)
first_conv_tc_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ scale,
                     const float* __restrict__ shift, int relu, int NB, int H, int W, bf16* __restrict__ out,
                     int out_pitch, float* stats) {
  constexpr int K = CIN * 9;
  constexpr int KS = (K + 15) / 16;  // 16-wide k-slices
  extern __shared__ uint8_t smem_raw[];
  const FcSmem s = carve(smem_raw);
  uint64_t* tmem_full = s.bars;       // [2] MMAs of the tile done: accumulator ready, operand buffer free
  uint64_t* tmem_empty = s.bars + 2;  // [2] accumulator drained by the 256 epilogue threads
  float* s_sc = s.fl;                 // [64]
  float* s_sh = s.fl + 64;            // [64]
  float* s_red = s.fl + 128;          // [4 quarters][2][64]
  uint8_t* w_hi = s.b;
  uint8_t* w_lo = s.b + FC_B_TILE;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&tmem_full[i], 1);
      ptx::mbar_init(&tmem_empty[i], 256);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(s.tmem_ptr, 128);
    ptx::tmem_relinquish();
  }
  // k columns >= 16*KS of the operand tiles are never written and never read; zero everything once for hygiene
  for (int i = threadIdx.x; i < (2 * FC_A_BUF + 2 * FC_A_TILE) / 16; i += FC_FWD_THREADS)
    reinterpret_cast<uint4*>(s.a[0])[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  if (threadIdx.x < 64) {
    s_sc[threadIdx.x] = scale ? scale[threadIdx.x] : 1.f;
    s_sh[threadIdx.x] = shift ? shift[threadIdx.x] : 0.f;
  }
  if (threadIdx.x < 256) {  // weights [co][k] fp32 -> W_hi / W_lo tiles (row = co)
    const int co = threadIdx.x & 63, c0 = threadIdx.x >> 6;
    for (int c = c0; c < 2 * KS; c += 4) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int k = c * 8 + j;
        v[j] = k < K ? w[co * K + k] : 0.f;
      }
      split_store8(w_hi, w_lo, sw128(co, c), v);
    }
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s.tmem_ptr;

  const int tiles_x = (W + TILE_W - 1) / TILE_W, tiles_y = (H + TILE_H - 1) / TILE_H;
  const int total = NB * tiles_x * tiles_y;

  if (warp < 4) {
    // ===================== builders + MMA issue =====================
    const uint32_t idesc = ptx::make_idesc_bf16(128, 64, 0, 0);
    const uint64_t b_hi_d = ptx::make_smem_desc(ptx::smem_u32(w_hi), 16, 1024);
    const uint64_t b_lo_d = ptx::make_smem_desc(ptx::smem_u32(w_lo), 16, 1024);
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      int t = tile;
      const int tx = t % tiles_x;
      t /= tiles_x;
      const int ty = t % tiles_y;
      const int img = t / tiles_y;
      // the MMAs that read this operand buffer two tiles ago have completed
      if (it >= 2) ptx::mbar_wait(&tmem_full[buf], ((it >> 1) & 1) ^ 1u);
      build_im2col_row<CIN>(x, img, ty * TILE_H, tx * TILE_W, H, W, threadIdx.x, s.a[buf], s.a[buf] + FC_A_TILE);
      ptx::fence_proxy_async_smem();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
      named_bar(2, 128);
      if (warp == 0) {
        if (ptx::elect_one()) {
          ptx::mbar_wait(&tmem_empty[buf], ((it >> 1) & 1) ^ 1u);
          ptx::tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * 64;
          const uint64_t a_hi_d = ptx::make_smem_desc(ptx::smem_u32(s.a[buf]), 16, 1024);
          const uint64_t a_lo_d = ptx::make_smem_desc(ptx::smem_u32(s.a[buf] + FC_A_TILE), 16, 1024);
#pragma unroll
          for (int k = 0; k < KS; ++k) ptx::umma_bf16(d_tmem, a_hi_d + 2 * k, b_hi_d + 2 * k, idesc, k != 0);
#pragma unroll
          for (int k = 0; k < KS; ++k) ptx::umma_bf16(d_tmem, a_lo_d + 2 * k, b_hi_d + 2 * k, idesc, 1);
#pragma unroll
          for (int k = 0; k < KS; ++k) ptx::umma_bf16(d_tmem, a_hi_d + 2 * k, b_lo_d + 2 * k, idesc, 1);
          ptx::umma_commit(&tmem_full[buf]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue =====================
    const int q = warp & 3, hc = (warp - 4) >> 2;  // TMEM lane quarter / 32-column half of this warp
    const int r = q * 32 + lane;
    const int e = threadIdx.x - 128;
    float r1[32], r2[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r1[j] = r2[j] = 0.f;
    int it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++it) {
      const int buf = it & 1;
      int t = tile;
      const int tx = t % tiles_x;
      t /= tiles_x;
      const int ty = t % tiles_y;
      const int img = t / tiles_y;
      const int y = ty * TILE_H + (r >> 4), xx = tx * TILE_W + (r & 15);
      const bool valid = y < H && xx < W;
      ptx::mbar_wait(&tmem_full[buf], (it >> 1) & 1);
      ptx::tc_fence_after();
      uint32_t v[32];
      ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * 64 + hc * 32, v);
      ptx::tmem_ld_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[buf]);  // accumulator is in registers: the stage may be overwritten
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
        store16(dst, pk);
        store16(dst + 8, pk + 4);
        store16(dst + 16, pk + 8);
        store16(dst + 24, pk + 12);
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
      named_bar(1, 256);
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
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 128);
  }
}

// ---------------------------------------------------------------------------------------------------------------- wgrad
// 8 warps = two builder groups (w0-3 even tiles, w4-7 odd tiles); the elected lane of each group's first warp fetches the
// tile's dRaw box by TMA and issues its 8 MMAs.  The accumulator never leaves TMEM until the CTA has done all tiles.
constexpr int FC_WG_THREADS = 256;
template <int CIN>
__global__ void __launch_bounds__(FC_WG_THREADS, 1)
first_conv_wgrad_tc_kernel(const float* __restrict__ x, const __grid_constant__ CUtensorMap d_map, int NB, int H, int W,
                           float* partials) {
  extern __shared__ uint8_t smem_raw[];
  const FcSmem s = carve(smem_raw);
  uint64_t* d_full = s.bars;      // [2] dRaw tile landed
  uint64_t* mma_done = s.bars + 2;  // [2] MMAs of the tile done: its operand buffers are free
  uint64_t* first_done = s.bars + 4;  // one-shot: group 0's first tile has started the accumulator
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&d_map);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&d_full[i], 1);
      ptx::mbar_init(&mma_done[i], 1);
    }
    ptx::mbar_init(first_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(s.tmem_ptr, 64);
    ptx::tmem_relinquish();
  }
  for (int i = threadIdx.x; i < 2 * FC_A_BUF / 16; i += FC_WG_THREADS)  // k columns >= 16*KS are never written
    reinterpret_cast<uint4*>(s.a[0])[i] = make_uint4(0, 0, 0, 0);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *s.tmem_ptr;
  const int tiles_x = (W + TILE_W - 1) / TILE_W, tiles_y = (H + TILE_H - 1) / TILE_H;
  const int total = NB * tiles_x * tiles_y;
  // both operands MN-major: A = (x_hi | x_lo) im2col tiles, two 64-wide boxes 16 KB apart; B = dRaw [128 px][64 co]
  const uint32_t idesc = ptx::make_idesc_bf16(128, 64, 1, 1);
  const int grp = warp >> 2;                 // builder group = operand buffer
  const int r = threadIdx.x & 127;
  uint8_t* a_buf = s.a[grp];
  uint8_t* d_buf = s.b + grp * FC_A_TILE;
  const uint64_t adesc = ptx::make_smem_desc(ptx::smem_u32(a_buf), FC_A_TILE, 1024);
  const uint64_t bdesc = ptx::make_smem_desc(ptx::smem_u32(d_buf), FC_A_TILE, 1024);
  const bool leader = (warp & 3) == 0;

  // group g handles the CTA's tiles number g, g+2, ...; the MMAs of the two groups interleave on one accumulator, which
  // is fine: accumulation order inside a CTA is free (fp32 sums), only the very first MMA must not accumulate.
  int n = 0;  // tiles this group has done
  for (int tile = blockIdx.x + grp * gridDim.x; tile < total; tile += 2 * gridDim.x, ++n) {
    int t = tile;
    const int tx = t % tiles_x;
    t /= tiles_x;
    const int ty = t % tiles_y;
    const int img = t / tiles_y;
    const int y0 = ty * TILE_H, x0 = tx * TILE_W;
    if (n >= 1) ptx::mbar_wait(&mma_done[grp], (n - 1) & 1);  // this group's previous MMAs have consumed the buffers
    if (leader) {
      if (ptx::elect_one()) {
        ptx::mbar_arrive_expect_tx(&d_full[grp], FC_A_TILE);
        ptx::tma_load_4d(d_buf, &d_map, &d_full[grp], 0, x0, y0, img);  // out-of-image pixels arrive as zeros
      }
      __syncwarp();
    }
    build_im2col_row<CIN>(x, img, y0, x0, H, W, r, a_buf, a_buf + FC_A_TILE);
    ptx::fence_proxy_async_smem();
    named_bar(2 + grp, 128);
    if (leader) {
      if (ptx::elect_one()) {
        ptx::mbar_wait(&d_full[grp], n & 1);
        ptx::tc_fence_after();
        // group 1 must not start the accumulator: its first MMA accumulates onto group 0's first tile, so it waits for it
        if (grp == 1 && n == 0) ptx::mbar_wait(first_done, 0);
#pragma unroll
        for (int k = 0; k < 8; ++k)  // 16 pixels per MMA: +2048 B in both tiles
          ptx::umma_bf16(tmem_base, adesc + 128 * k, bdesc + 128 * k, idesc, (grp | n | k) != 0);
        ptx::umma_commit(&mma_done[grp]);
        if (grp == 0 && n == 0) ptx::umma_commit(first_done);
      }
      __syncwarp();
    }
  }
  // all MMAs of both groups complete -> read the accumulator
  if (n >= 1) ptx::mbar_wait(&mma_done[grp], (n - 1) & 1);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  {
    const int q = warp & 3, hc = warp >> 2;
    uint32_t v[32];
    ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + hc * 32, v);
    ptx::tmem_ld_wait();
    float* dst = partials + (static_cast<long>(blockIdx.x) * 128 + q * 32 + lane) * 64 + hc * 32;
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      *reinterpret_cast<float4*>(dst + j) = make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                                                        __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, 64);
  }
}

// partials [ncta][128 = (k_hi | k_lo)][64 co] -> dw[co][k], k < K.  Block = one k: 64 channels x 4 CTA-lanes.
__global__ void __launch_bounds__(256) first_conv_wgrad_fold_kernel(const float* __restrict__ partials, int ncta, int K,
                                                                    float* dw, int accumulate) {
  __shared__ float s[4][64];
  const int k = blockIdx.x, co = threadIdx.x & 63, ln = threadIdx.x >> 6;
  float a = 0.f;
  for (int c = ln; c < ncta; c += 4)
    a += partials[(static_cast<long>(c) * 128 + k) * 64 + co] + partials[(static_cast<long>(c) * 128 + 64 + k) * 64 + co];
  s[ln][co] = a;
  __syncthreads();
  if (ln == 0) {
    const float g = (s[0][co] + s[1][co]) + (s[2][co] + s[3][co]);
    float* d = dw + co * K + k;
    *d = accumulate ? *d + g : g;
  }
}

template <typename Kern>
cudaError_t set_smem(Kern kern) {
  return cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, FC_SMEM_BYTES);
}

}  // namespace

int first_conv_tc_grid(int NB, int H, int W) {
  const int tiles = NB * ((H + TILE_H - 1) / TILE_H) * ((W + TILE_W - 1) / TILE_W);
  const int sms = device_num_sms();
  return tiles < sms ? tiles : sms;
}

// cin <= 7 (K = 9*cin <= 63 fits one 64-wide k-block); stats: [grid][2][64] partial rows
cudaError_t launch_first_conv_tc(const float* x, const float* w, const float* scale, const float* shift, int relu,
                                 int NB, int cin, int H, int W, bf16* out, int out_pitch, float* stats,
                                 cudaStream_t st) {
  const int grid = first_conv_tc_grid(NB, H, W);
#define FC(C)                                                                                                     \
  if (cin == C) {                                                                                                 \
    static bool done = false;                                                                                     \
    if (!done) {                                                                                                  \
      cudaError_t e = set_smem(first_conv_tc_kernel<C>);                                                          \
      if (e != cudaSuccess) return e;                                                                             \
      done = true;                                                                                                \
    }                                                                                                             \
    first_conv_tc_kernel<C><<<grid, FC_FWD_THREADS, FC_SMEM_BYTES, st>>>(x, w, scale, shift, relu, NB, H, W, out,     \
                                                                     out_pitch, stats);                           \
    return cudaGetLastError();                                                                                    \
  }
  FC(1) FC(2) FC(3) FC(4) FC(5) FC(6) FC(7)
#undef FC
  return cudaErrorInvalidValue;
}

// partials: at least first_conv_tc_grid() * 128 * 64 floats
cudaError_t launch_first_conv_wgrad_tc(const float* x, View draw, int cin, float* partials, float* dw, int accumulate,
                                       cudaStream_t st) {
  if (draw.C != 64) return cudaErrorInvalidValue;
  CUtensorMap map;
  if (make_act_map(&map, draw, TILE_H) != 0) return cudaErrorInvalidValue;
  const int grid = first_conv_tc_grid(draw.N, draw.H, draw.W);
#define FW(C)                                                                                                     \
  if (cin == C) {                                                                                                 \
    static bool done = false;                                                                                     \
    if (!done) {                                                                                                  \
      cudaError_t e = set_smem(first_conv_wgrad_tc_kernel<C>);                                                    \
      if (e != cudaSuccess) return e;                                                                             \
      done = true;                                                                                                \
    }                                                                                                             \
    first_conv_wgrad_tc_kernel<C><<<grid, FC_WG_THREADS, FC_SMEM_BYTES, st>>>(x, map, draw.N, draw.H, draw.W,        \
                                                                           partials);                             \
    first_conv_wgrad_fold_kernel<<<C * 9, 256, 0, st>>>(partials, grid, C * 9, dw, accumulate);                   \
    return cudaGetLastError();                                                                                    \
  }
  FW(1) FW(2) FW(3) FW(4) FW(5) FW(6) FW(7)
#undef FW
  return cudaErrorInvalidValue;
}

// Weight-gradient GEMM for the echogram U-Net on sm_100a (autograd of reference models/unet.py:35-49,
// i.e. what loss.backward() at pipeline_train_predict/pipeline.py:177 computes for every conv weight):
//
//     dW[m][n][tap] = sum over pixels p of  F[p][m] * T_tap[p][n]
//
//   3x3 conv      : F = dY (m = Cout),  T_tap = X shifted by (ky-1,kx-1) with zero fill (n = Cin), 9 taps
//   ConvTranspose : F = X  (m = Cin),   T_tap = dY sub-sampled at (2y+ky, 2x+kx)        (n = Cout), 4 taps
//
// The reduction dimension (pixels) is the slow dimension of both NHWC operands, so both are fed to
// tcgen05.mma as MN-major SWIZZLE_128B tiles: a k-step is a 4x16-pixel TMA box {64 ch, 16, 4, 1}
// per 64-channel block, 64 k-rows of 128 bytes.  One CTA owns one (m-tile, n-tile, tap) output tile over one
// split of the pixel range; partial tiles are accumulated with vectorised red.global.add.f32 into a
// [tap][m][n] fp32 scratch which a small kernel then permutes into PyTorch's (m, n, kh, kw) layout.
#include "host_util.h"
#include "ptx.cuh"

namespace {

constexpr int WG_KPIX = 64;                  // pixels per k-step
constexpr int WG_BOX_BYTES = WG_KPIX * 128;  // one {64ch x 64px} box

template <int BLOCK_N>
struct WgCfg {
  static constexpr int A_BYTES = 2 * WG_BOX_BYTES;  // 128 m-channels
  static constexpr int B_BYTES = (BLOCK_N / 64) * WG_BOX_BYTES;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  static constexpr int TMEM_COLS = BLOCK_N;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256 + 1024;
};

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <int BLOCK_N>
__global__ void __launch_bounds__(256, 1) wgrad_gemm_kernel(const __grid_constant__ WgradParams p) {
  using Cfg = WgCfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* aux = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* acc_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  // tile decode: blockIdx.x -> (tap, m_tile, n_tile); blockIdx.y -> split of the pixel range
  int t = blockIdx.x;
  const int n_tile = t % p.n_tiles;
  t /= p.n_tiles;
  const int m_tile = t % p.m_tiles;
  const int tap = t / p.m_tiles;
  const int split = blockIdx.y;
  const int kt_per = (p.k_tiles_total + p.splits - 1) / p.splits;
  const int kt_begin = split * kt_per;
  const int kt_end = min(p.k_tiles_total, kt_begin + kt_per);
  const int ksteps = kt_end - kt_begin;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.a_map);
    ptx::prefetch_tmap(&p.b_map[p.tap_mode ? tap : 0]);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(acc_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (ksteps > 0) {
    // single-thread roles are chosen with elect.sync so that the compiler emits the uniform-datapath TMA / MMA
    // instructions without per-instruction active-thread loops
    if (warp == 0) {
      if (ptx::elect_one()) {
      // ===================== TMA producer =====================
      int dy = 0, dx = 0, mi = 0;
      if (p.tap_mode == 0) {
        if (p.taps == 9) {
          dy = tap / 3 - 1;
          dx = tap % 3 - 1;
        }
      } else {
        mi = tap;
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = kt_begin; kt < kt_end; ++kt) {
        int u = kt;
        const int tx = u % p.tiles_x;
        u /= p.tiles_x;
        const int ty = u % p.tiles_y;
        const int img = u / p.tiles_y;
        const int x0 = tx * 16, y0 = ty * 4;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
#pragma unroll
        for (int b = 0; b < 2; ++b)
          ptx::tma_load_4d(sa + b * WG_BOX_BYTES, &p.a_map, &full_bar[stage], m_tile * 128 + b * 64, x0, y0, img);
#pragma unroll
        for (int b = 0; b < BLOCK_N / 64; ++b)
          ptx::tma_load_4d(sb + b * WG_BOX_BYTES, &p.b_map[mi], &full_bar[stage], n_tile * BLOCK_N + b * 64, x0 + dx,
                           y0 + dy, img);
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      }
    } else if (warp == 1) {
      if (ptx::elect_one()) {
      // ===================== MMA issuer =====================
      const uint32_t idesc = ptx::make_idesc_bf16(128, BLOCK_N, 1, 1);  // both operands MN-major
      int stage = 0;
      uint32_t phase = 0;
      for (int ks = 0; ks < ksteps; ++ks) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
        // MN-major SW128: LBO = byte distance between 64-channel boxes, SBO = 1024 (8 k-rows of 128 B)
        const uint64_t adesc = ptx::make_smem_desc(sa, WG_BOX_BYTES, 1024);
        const uint64_t bdesc = ptx::make_smem_desc(sa + Cfg::A_BYTES, WG_BOX_BYTES, 1024);
#pragma unroll
        for (int k = 0; k < WG_KPIX / 16; ++k) {
          // 16 pixels further along K = 16 rows * 128 B = 2048 B: start-address field += 128
          ptx::umma_bf16(tmem_base, adesc + 128 * k, bdesc + 128 * k, idesc, (ks | k) != 0);
        }
        ptx::umma_commit(&empty_bar[stage]);
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      ptx::umma_commit(acc_bar);
      }
    } else if (warp >= 4) {
      // ===================== epilogue: TMEM -> red.add into [tap][m][n] scratch =====================
      const int q = warp & 3;
      const int m = m_tile * 128 + q * 32 + lane;
      ptx::mbar_wait(acc_bar, 0);
      ptx::tc_fence_after();
      const bool direct = (p.splits == 1) || (p.slabs != nullptr);   // plain stores: sole writer of this element
      float* row = (p.slabs ? p.slabs + split * p.slab_stride : p.dw) +
                   (static_cast<long>(tap) * p.M_total + m) * p.N_total + n_tile * BLOCK_N;
#pragma unroll 1
      for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + chunk * 32, v);
        ptx::tmem_ld_wait();
        if (m < p.M_total) {
          if (direct) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<float4*>(row + chunk * 32 + j) =
                  make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                              __uint_as_float(v[j + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              red_add_v4(row + chunk * 32 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                         __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// scratch [taps][M][N] fp32 -> PyTorch layout [M][N][taps] (conv: (Cout,Cin,3,3); convT: (Cin,Cout,2,2))
__global__ void wgrad_unpack_kernel(const float* __restrict__ scratch, float* __restrict__ dw, int M, int N, int taps,
                                    int accumulate) {
  const long total = static_cast<long>(M) * N;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    for (int tp = 0; tp < taps; ++tp) {
      const float g = scratch[tp * total + i];
      float* d = dw + i * taps + tp;
      *d = accumulate ? (*d + g) : g;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// 3x3-conv weight gradient, all nine taps per CTA.
//   dW[co][ci][ky][kx] = sum_q dY[q - (ky-1, kx-1)][co] * X[q][ci]
// k-step = a 4x16-pixel tile q of X (fixed operand, GEMM-N, 64 ci) plus ONE 6x18-pixel halo tile of dY (GEMM-M, 64 co).
// Tap (ky,kx) is the view of the halo tile starting at row (2-ky)*18 + (2-kx): the SWIZZLE_128B pattern is a pure
// function of the shared-memory address, so an MN-major descriptor may start at any 128-byte row.  Two taps are
// stacked along M (rows 0-63 / 64-127 of the MMA = two "64-channel boxes" LBO bytes apart), which fills the 128-row
// tensor-core tile although only 64 output channels are live: 5 MMA groups (the 5th repeats tap 1 in its lower half,
// discarded) cover 9 taps.  TMEM: 5 accumulators x 64 fp32 columns.  Operand traffic per k-step: 13.5 KB + 8 KB for
// 20 MMAs (640 cycles) instead of 9 x 24 KB.
constexpr int WH_HALO_W = 18, WH_HALO_H = 6;
constexpr int WH_HALO_BYTES = WH_HALO_W * WH_HALO_H * 128;  // 13824
constexpr int WH_HALO_SLOT = 14336;                          // padded to a multiple of 1024
constexpr int WH_FIXED_BYTES = 64 * 128;                     // 8192 per 64-channel box of the fixed operand
constexpr int WH_TMEM_COLS = 512;

// NF = channels of the fixed operand (X) per CTA = MMA N.
//   NF = 64 : five accumulators (tap pairs (8,7) (6,5) (4,3) (2,1) (1,0); the lower half of the last repeats tap 1 and is
//             discarded) - 20 MMAs of 128x64x16 per k-step; bound by shared-memory bandwidth (6 KB per 32-cycle MMA).
//   NF = 128: (Cin a multiple of 128) 128-wide X tiles: 8 KB per 64-cycle MMA, i.e. 2/3 of the shared-memory traffic per
//             FLOP.  Five 128-column accumulators do not fit TMEM (512 columns), so the nine taps are split over TWO CTAs
//             (blockIdx.z): kind 0 owns the pairs (8,7) (6,5) (4,3), kind 1 the pairs (2,1) (1,0) (lower half of the
//             last discarded, as above).  Both kinds load the same operand tiles; the long kind is launched first.
template <int NF>
struct WhCfg {
  static constexpr int GROUPS = (NF == 64) ? 5 : 3;   // NF = 128: at most three per CTA kind
  static constexpr int STAGE = WH_HALO_SLOT + (NF / 64) * WH_FIXED_BYTES;
  static constexpr int STAGES = (NF == 64) ? 8 : 6;
  static constexpr int SMEM = STAGES * STAGE + 256 + 1024;
};

template <int NF>
__global__ void __launch_bounds__(256, 1) wgrad_halo_kernel(const __grid_constant__ WgradHaloParams p) {
  using Cfg = WhCfg<NF>;
  constexpr int G = Cfg::GROUPS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* aux = smem + Cfg::STAGES * Cfg::STAGE;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);
  uint64_t* empty_bar = full_bar + Cfg::STAGES;
  uint64_t* acc_bar = empty_bar + Cfg::STAGES;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(acc_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int f_tiles = p.Cf / NF;
  const int f_tile = blockIdx.x % f_tiles;
  const int s_tile = blockIdx.x / f_tiles;
  const int split = blockIdx.y;
  const int kind = (NF == 64) ? 0 : blockIdx.z;        // NF = 128: which subset of the tap pairs this CTA owns
  const int g_run = (NF == 64) ? 5 : (kind == 0 ? 3 : 2);
  const int kt_per = (p.k_tiles_total + p.splits - 1) / p.splits;
  const int kt_begin = split * kt_per;
  const int kt_end = min(p.k_tiles_total, kt_begin + kt_per);
  const int ksteps = kt_end - kt_begin;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.s_map);
    ptx::prefetch_tmap(&p.f_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    ptx::mbar_init(acc_bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_smem, WH_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // tap pairs (first -> MMA rows 0..63, second -> rows 64..127); halo row offsets grow from first to second
  constexpr int PA5[5] = {8, 6, 4, 2, 1}, PB5[5] = {7, 5, 3, 1, 0};
  constexpr int PA3[2][3] = {{8, 6, 4}, {2, 1, 1}}, PB3[2][3] = {{7, 5, 3}, {1, 0, 0}};   // NF = 128, per CTA kind
  auto pair_a = [&](int g) { return (NF == 64) ? PA5[g] : PA3[kind][g]; };
  auto pair_b = [&](int g) { return (NF == 64) ? PB5[g] : PB3[kind][g]; };

  if (ksteps > 0) {
    if (warp == 0) {
      if (ptx::elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int kt = kt_begin; kt < kt_end; ++kt) {
        int u = kt;
        const int tx = u % p.tiles_x;
        u /= p.tiles_x;
        const int ty = u % p.tiles_y;
        const int img = u / p.tiles_y;
        const int x0 = tx * 16, y0 = ty * 4;
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
        uint8_t* sh = smem + stage * Cfg::STAGE;
        ptx::mbar_arrive_expect_tx(&full_bar[stage], WH_HALO_BYTES + (NF / 64) * WH_FIXED_BYTES);
        ptx::tma_load_4d(sh, &p.s_map, &full_bar[stage], s_tile * 64, x0 - 1, y0 - 1, img);
#pragma unroll
        for (int b = 0; b < NF / 64; ++b)
          ptx::tma_load_4d(sh + WH_HALO_SLOT + b * WH_FIXED_BYTES, &p.f_map, &full_bar[stage], f_tile * NF + b * 64, x0,
                           y0, img);
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      }
    } else if (warp == 1) {
      if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, NF, 1, 1);
      // per pair: descriptor (without start address) and the byte offset of the first tap's view
      uint64_t dtmpl[G];
      uint32_t off16[G];
#pragma unroll
      for (int g = 0; g < G; ++g) {
        const int ta = pair_a(g), tb = pair_b(g);
        const int oa = ((2 - ta / 3) * WH_HALO_W + (2 - ta % 3)) * 128;
        const int ob = ((2 - tb / 3) * WH_HALO_W + (2 - tb % 3)) * 128;
        dtmpl[g] = ptx::make_smem_desc(0, ob - oa, 1024);
        off16[g] = oa >> 4;
      }
      int stage = 0;
      uint32_t phase = 0;
      for (int ks = 0; ks < ksteps; ++ks) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint32_t sh = ptx::smem_u32(smem + stage * Cfg::STAGE);
        const uint32_t sh16 = sh >> 4;
        const uint64_t bdesc0 = ptx::make_smem_desc(sh + WH_HALO_SLOT, WH_FIXED_BYTES, 1024);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
          const uint64_t bdesc = bdesc0 + static_cast<uint64_t>(r * 128);          // 16 pixels * 128 B = 2048 B
          const uint32_t row16 = sh16 + static_cast<uint32_t>(r * WH_HALO_W * 8);  // halo row pitch 18 * 128 B
#pragma unroll
          for (int g = 0; g < G; ++g)
            if (g < g_run) ptx::umma_bf16(tmem_base + g * NF, dtmpl[g] + row16 + off16[g], bdesc, idesc, (ks | r) != 0);
        }
        ptx::umma_commit(&empty_bar[stage]);
        if (++stage == Cfg::STAGES) {
          stage = 0;
          phase ^= 1u;
        }
      }
      ptx::umma_commit(acc_bar);
      }
    } else if (warp >= 4) {
      const int q = warp & 3;
      const int m = q * 32 + lane;
      const int half = m >> 6;
      const int co = s_tile * 64 + (m & 63);
      ptx::mbar_wait(acc_bar, 0);
      ptx::tc_fence_after();
#pragma unroll 1
      for (int g = 0; g < g_run; ++g) {
        const int tap = half ? pair_b(g) : pair_a(g);
        // the lower half of the last (2,1)/(1,0) chain repeats tap 1: discarded
        const bool live = !(half == 0 && ((NF == 64 && g == 4) || (NF == 128 && kind == 1 && g == 1)));
        const bool direct = (p.splits == 1) || (p.slabs != nullptr);
        float* row = (p.slabs ? p.slabs + split * p.slab_stride : p.dw) + (static_cast<long>(tap) * p.Cs + co) * p.Cf +
                     f_tile * NF;
#pragma unroll 1
        for (int chunk = 0; chunk < NF / 32; ++chunk) {
          uint32_t v[32];
          ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + g * NF + chunk * 32, v);
          ptx::tmem_ld_wait();
          if (live) {
            if (direct) {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(row + chunk * 32 + j) =
                    make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]),
                                __uint_as_float(v[j + 3]));
            } else {
#pragma unroll
              for (int j = 0; j < 32; j += 4)
                red_add_v4(row + chunk * 32 + j, __uint_as_float(v[j]), __uint_as_float(v[j + 1]),
                           __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
            }
          }
        }
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, WH_TMEM_COLS);
  }
}

// All layers in one launch at the end of backward: scratch [taps][M*N] -> dw [M*N][taps], and the scratch is left
// ZEROED for the next step's red.add accumulation (no memset launches).  Work item = 1024 (m,n) pairs of one layer.
__global__ void __launch_bounds__(256) wgrad_unpack_all_kernel(const __grid_constant__ UnpackTable t) {
  __shared__ float s[256 * 9];
  int li = 0;
  while (li + 1 < t.n && static_cast<int>(blockIdx.x) >= t.e[li + 1].item0) ++li;
  const UnpackEntry& L = t.e[li];
  const long base = static_cast<long>(blockIdx.x - L.item0) * 1024;
  for (int u = 0; u < 4; ++u) {
    const long i0 = base + u * 256;
    if (i0 >= L.mn) break;
    const int n = static_cast<int>(min(256L, L.mn - i0));
    // coalesced read (and zero) of each tap plane, transposed through shared memory into contiguous [pair][tap] rows
    if (static_cast<int>(threadIdx.x) < n) {
      // all loads first (independent, in flight together), then the zero stores: the compiler must not be forced to
      // order each load behind the previous tap's store to the same array
      float* src = L.scratch + i0 + threadIdx.x;
      float g[9];
      if (L.slabs != nullptr) {
        // deterministic mode: the split-K partial tiles were stored, not added: sum them here in split order
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) g[tp] = 0.f;
        for (int sp = 0; sp < L.splits; ++sp) {
          const float* q = L.slabs + sp * L.slab_stride + i0 + threadIdx.x;
#pragma unroll
          for (int tp = 0; tp < 9; ++tp)
            if (tp < L.taps) g[tp] += __ldcs(q + tp * L.mn);
        }
#pragma unroll
        for (int tp = 0; tp < 9; ++tp)
          if (tp < L.taps) s[threadIdx.x * L.taps + tp] = g[tp];
      } else {
#pragma unroll
        for (int tp = 0; tp < 9; ++tp) g[tp] = (tp < L.taps) ? __ldcs(src + tp * L.mn) : 0.f;
#pragma unroll
        for (int tp = 0; tp < 9; ++tp)
          if (tp < L.taps) {
            src[tp * L.mn] = 0.f;
            s[threadIdx.x * L.taps + tp] = g[tp];
          }
      }
    }
    __syncthreads();
    float* dst = L.dw + i0 * L.taps;
    for (int j = threadIdx.x; j < n * L.taps; j += 256) dst[j] = s[j];
    __syncthreads();
  }
}

template <int BLOCK_N>
cudaError_t launch_wg(const WgradParams& p, cudaStream_t stream) {
  using Cfg = WgCfg<BLOCK_N>;
  auto kern = wgrad_gemm_kernel<BLOCK_N>;
  if (cudaError_t e = ensure_dynamic_smem(kern, Cfg::SMEM_BYTES); e != cudaSuccess) return e;
  dim3 grid(p.taps * p.m_tiles * p.n_tiles, p.splits);
  kern<<<grid, 256, Cfg::SMEM_BYTES, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_wgrad_gemm(const WgradParams& p, int block_n, cudaStream_t stream) {
  if (block_n == 64) return launch_wg<64>(p, stream);
  if (block_n == 128) return launch_wg<128>(p, stream);
  if (block_n == 256) return launch_wg<256>(p, stream);
  return cudaErrorInvalidValue;
}

namespace {
template <int NF>
cudaError_t launch_wh(const WgradHaloParams& p, cudaStream_t stream) {
  if (cudaError_t e = ensure_dynamic_smem(wgrad_halo_kernel<NF>, WhCfg<NF>::SMEM); e != cudaSuccess) return e;
  dim3 grid(p.s_tiles * (p.Cf / NF), p.splits, NF == 64 ? 1 : 2);
  wgrad_halo_kernel<NF><<<grid, 256, WhCfg<NF>::SMEM, stream>>>(p);
  return cudaGetLastError();
}
}  // namespace

// p.nf = 64: all nine taps per CTA, 64 x 64 channel tiles.  p.nf = 128 (Cf % 128 == 0): 64 x 128 channel tiles, the nine
// taps split over two CTA kinds (see WhCfg).
cudaError_t launch_wgrad_halo(const WgradHaloParams& p, cudaStream_t stream) {
  if (p.nf == 128) return (p.Cf % 128 == 0) ? launch_wh<128>(p, stream) : cudaErrorInvalidValue;
  return launch_wh<64>(p, stream);
}

cudaError_t launch_wgrad_unpack(const float* scratch, float* dw, int M, int N, int taps, int accumulate,
                                cudaStream_t stream) {
  const long total = static_cast<long>(M) * N;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 148 * 16) blocks = 148 * 16;
  wgrad_unpack_kernel<<<blocks, 256, 0, stream>>>(scratch, dw, M, N, taps, accumulate);
  return cudaGetLastError();
}

cudaError_t launch_wgrad_unpack_all(UnpackTable& t, cudaStream_t stream) {
  int items = 0;
  for (int i = 0; i < t.n; ++i) {
    t.e[i].item0 = items;
    items += static_cast<int>((t.e[i].mn + 1023) / 1024);
  }
  if (items > 0) wgrad_unpack_all_kernel<<<items, 256, 0, stream>>>(t);
  return cudaGetLastError();
}

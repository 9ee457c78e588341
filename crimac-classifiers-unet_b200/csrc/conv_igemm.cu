// Implicit-GEMM convolution for the echogram U-Net on sm_100a: tcgen05.mma (bf16 x bf16 -> fp32 in TMEM),
// operands staged by TMA into 128B-swizzled shared memory, persistent warp-specialised CTAs.
//
//   GEMM view:  M = output pixels (8x16 spatial tile = 128 TMEM lanes),
//                                  N = output channels, K = taps * Cin.
//   A k-step = (tap, 64-channel block): ONE 4-D TMA box {64 ch, 16 pings, 8 rows, 1 patch} of the NHWC
//   activation tensor at the tap-shifted coordinate; out-of-bounds rows/pings are zero-filled by TMA,
//   which IS nn.Conv2d's zero padding (reference models/unet.py:35-44).  The same kernel serves
//     - 3x3 conv forward + BN/ReLU epilogue            (unet.py:76-83, :135-136)
//     - 3x3 conv backward-data (weights packed rotated, Cin<->Cout swapped)
//     - ConvTranspose2d(k2,s2) forward as a 1-tap GEMM with a scatter epilogue into the concat buffer
//       (unet.py:47-49, :130-132)
//     - ConvTranspose2d backward-data as a 4-"tap" GEMM, one sub-sampled tensor map per (ky,kx)
//     - the fused 1x1 head + softmax epilogue           (unet.py:284,342; pipeline.py:218)
//
// Warp roles (384 threads): w0 = TMA producer, w1 = MMA issuer (one thread), w2 = TMEM allocator,
// w4..w11 = epilogue: warp w reads TMEM lane quarter w%4 (lane <-> pixel of the tile) and owns every second
// 32-column chunk of the accumulator, so two warps per SM sub-partition hide each other's TMEM / shuffle latency.
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.  With one chunk per
// warp (Cout = 64 layers) the train-mode BatchNorm statistics stay in registers for the CTA's lifetime.
#include "host_util.h"
#include "ptx.cuh"
#include "devfn.cuh"
#include <cuda_fp16.h>

namespace {

// HALO variant (3x3 convs): the M tile is 16 rows x 8 pings and the A operand of one 64-channel block is ONE
// 18 x 10 pixel TMA box; the nine taps are nine UMMA descriptors into that tile (start = the tap's first halo row,
// stride between 8-pixel row groups SBO = 10*128 B).  Valid because SWIZZLE_128B is a pure function of the shared
// memory address (probed on B200, profiles/r01_probe_first_run.log).  A-operand L2->SMEM traffic drops 6.25x; the
// weights keep their own, deeper ring.
constexpr int HALO_W = 10, HALO_H = 18;
constexpr int HALO_BYTES = HALO_W * HALO_H * 128;  // 23040
constexpr int HALO_SLOT = 23552;                   // padded to a multiple of 1024

template <int BLOCK_N, bool HALO = false>
struct ConvCfg {
  static constexpr int A_BYTES = TILE_M * KBLK * 2;
  static constexpr int B_BYTES = BLOCK_N * KBLK * 2;   // weights of one tap x 64 input channels
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  // halo rings: NA activation halo tiles; NB weight stages of TPS taps each (one mbarrier round trip and one
  // tcgen05.commit per TPS*4 MMAs: the single issuing thread, not the tensor pipe, is the scarce resource for narrow N)
  static constexpr int NA = (BLOCK_N == 256) ? 2 : 3;
  static constexpr int TPS = (BLOCK_N == 256) ? 1 : 3;
  static constexpr int NB = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 3 : 4);
  static constexpr int BSTAGE_BYTES = TPS * B_BYTES;
  // resident-weights variant (narrow layers whose whole weight matrix is <= 144 KB): [weights | 2-3 halo tiles]
  static constexpr int RES_MAX_B = 147456;
  static constexpr int RES_BYTES = (BLOCK_N == 256) ? 0 : (RES_MAX_B + 2 * HALO_SLOT);
  static constexpr int RING_BYTES = NA * HALO_SLOT + NB * BSTAGE_BYTES;
  static constexpr int OPERAND_BYTES = HALO ? (RING_BYTES > RES_BYTES ? RING_BYTES : RES_BYTES) : (STAGES * STAGE_BYTES);
  static constexpr int TMEM_COLS = 2 * BLOCK_N;
  static constexpr int MAX_STAT_CH = 4 * BLOCK_N;  // per-CTA running channel sums (EPI_STATS) over all n-tiles
  static constexpr int HEAD_BYTES = (BLOCK_N == 64) ? (CRIMAC_MAX_CLASSES * 64 + CRIMAC_MAX_CLASSES) * 4 : 0;
  static constexpr int AUX_BYTES = 512 /*barriers*/ + 4 * BLOCK_N * 4 /*scale/shift x2*/ + 8 * BLOCK_N * 4 /*stats*/ +
                                   2 * MAX_STAT_CH * 4 + HEAD_BYTES;
  static constexpr int SMEM_BYTES = OPERAND_BYTES + AUX_BYTES + 1024;
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory a CTA can have");
  static_assert(9 % TPS == 0, "taps per weight stage must divide 9");
};

// epilogue warps: 4 TMEM lane quarters x EPI_COLGROUPS column groups
constexpr int EPI_COLGROUPS = 2;
constexpr int EPI_WARPS = 4 * EPI_COLGROUPS;
constexpr int EPI_THREADS = 32 * EPI_WARPS;
constexpr int CONV_THREADS = 128 + EPI_THREADS;
__device__ __forceinline__ void epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

template <int BLOCK_N, int EPI, bool HALO, bool BMN>
__global__ void __launch_bounds__(CONV_THREADS, 1) conv_igemm_kernel(const __grid_constant__ ConvParams p) {
  using Cfg = ConvCfg<BLOCK_N, HALO>;
  constexpr int TW = HALO ? 8 : TILE_W;    // tile width (pings)
  constexpr int TH = HALO ? 16 : TILE_H;   // tile height (range rows)
  constexpr int NBAR_A = HALO ? Cfg::NA : Cfg::STAGES;   // non-halo: full/empty per stage live in the "A" arrays
  constexpr int NBAR_B = HALO ? Cfg::NB : 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* aux = smem + Cfg::OPERAND_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(aux);   // [NBAR_A]  (halo: A tiles)
  uint64_t* empty_bar = full_bar + NBAR_A;
  uint64_t* bfull_bar = empty_bar + NBAR_A;                // [NBAR_B]  (halo: weight tiles)
  uint64_t* bempty_bar = bfull_bar + NBAR_B;
  uint64_t* tmem_full = bempty_bar + NBAR_B;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_affine = reinterpret_cast<float*>(aux + 512);  // [2 acc stages][scale BLOCK_N | shift BLOCK_N]
  float* s_red = s_affine + 4 * BLOCK_N;                  // [4 warps][2][BLOCK_N]
  float* s_acc = s_red + 8 * BLOCK_N;                     // [2][n_total] CTA-lifetime channel sums
  float* s_head = s_acc + 2 * Cfg::MAX_STAT_CH;           // [ncls][64] + [ncls]           (BLOCK_N == 64 only)
  constexpr bool HAS_STATS = (EPI == EPI_STATS);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&p.a_map[0]);
    ptx::prefetch_tmap(&p.b_map);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NBAR_A; ++s) {
      ptx::mbar_init(&full_bar[s], 1);
      ptx::mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < NBAR_B; ++s) {
      ptx::mbar_init(&bfull_bar[s], 1);
      ptx::mbar_init(&bempty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&tmem_full[s], 1);
      // EPI_HEAD: a tile is read out by ONE of the two warps per lane quarter (all 64 columns), see the epilogue
      ptx::mbar_init(&tmem_empty[s], EPI == EPI_HEAD ? EPI_THREADS / EPI_COLGROUPS : EPI_THREADS);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) {
    ptx::tmem_alloc(tmem_ptr_smem, Cfg::TMEM_COLS);
    ptx::tmem_relinquish();
  }
  if (HAS_STATS && warp >= 4) {
    for (int i = threadIdx.x - 128; i < 2 * p.n_tiles * BLOCK_N; i += EPI_THREADS) s_acc[i] = 0.f;
  }
  if (EPI == EPI_HEAD && warp >= 4) {
    const int e = threadIdx.x - 128;
    for (int i = e; i < p.n_classes * 64; i += EPI_THREADS) s_head[i] = p.head_w[i];
    if (e < p.n_classes) s_head[CRIMAC_MAX_CLASSES * 64 + e] = p.head_b[e];
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  // register re-allocation between the warpgroups: the TMA / MMA / allocator warps need few registers, the epilogue
  // warps keep up to 128 BatchNorm accumulators per thread (128 x 56 + 256 x 224 = 64512 <= 65536 registers)
  // (setmaxnreg sits at the top of each role branch below, with no control-flow merge in between)

  const int cblocks = p.cin / KBLK;
  const int ksteps = p.taps * cblocks;

  // Only the variant that keeps 128 BatchNorm accumulators per epilogue thread re-allocates registers between the
  // warpgroups (128 x 72 + 256 x 216 = 64512 <= 65536); measured: the others run faster with the static 168.
  constexpr bool REALLOC = HAS_STATS && (BLOCK_N == 128);
  if (warp < 4) {
  if constexpr (REALLOC) asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
  if (warp == 0) {
    // ===================== TMA producer =====================
    // elect.sync (not lane == 0): the compiler then knows a single thread runs the loop and emits the uniform-datapath
    // TMA / MMA / commit instructions straight, without a per-instruction active-thread serialisation loop
    if (ptx::elect_one()) {
      int stage = 0, bstage = 0;
      uint32_t phase = 0, bphase = 0;
      (void)bstage;
      (void)bphase;
      if (HALO && p.resident) {
        // the whole weight matrix is loaded ONCE per CTA (n_tiles == 1) and stays in shared memory: block (tap, cb)
        // at (tap*cblocks + cb) * B_BYTES.  Per tile only the activation halo tiles move.
        ptx::mbar_arrive_expect_tx(&bfull_bar[0], 9 * cblocks * Cfg::B_BYTES);
        for (int tap = 0; tap < 9; ++tap)
          for (int cb = 0; cb < cblocks; ++cb) {
            uint8_t* sb = smem + (tap * cblocks + cb) * Cfg::B_BYTES;
            if constexpr (BMN) {
#pragma unroll
              for (int b = 0; b < BLOCK_N / 64; ++b)
                ptx::tma_load_2d(sb + b * 8192, &p.b_map, &bfull_bar[0], (8 - tap) * p.b_tap_cols + b * 64, cb * KBLK);
            } else {
              ptx::tma_load_2d(sb, &p.b_map, &bfull_bar[0], tap * p.cin + cb * KBLK, 0);
            }
          }
      }
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int n_tile = tile % p.n_tiles;
        int m_tile = tile / p.n_tiles;
        const int tx = m_tile % p.tiles_x;
        m_tile /= p.tiles_x;
        const int ty = m_tile % p.tiles_y;
        const int img = m_tile / p.tiles_y;
        const int x0 = tx * TW, y0 = ty * TH, n0 = n_tile * BLOCK_N;
        if (HALO && p.resident) {
          uint8_t* a_base = smem + 9 * cblocks * Cfg::B_BYTES;
          for (int cb = 0; cb < cblocks; ++cb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            ptx::mbar_arrive_expect_tx(&full_bar[stage], HALO_BYTES);
            ptx::tma_load_4d(a_base + stage * HALO_SLOT, &p.a_map[0], &full_bar[stage], cb * KBLK, x0 - 1, y0 - 1, img);
            if (++stage == p.resident) {  // p.resident = number of halo slots (2 or 3)
              stage = 0;
              phase ^= 1u;
            }
          }
          continue;
        }
        if constexpr (HALO) {
          for (int cb = 0; cb < cblocks; ++cb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            ptx::mbar_arrive_expect_tx(&full_bar[stage], HALO_BYTES);
            ptx::tma_load_4d(smem + stage * HALO_SLOT, &p.a_map[0], &full_bar[stage], cb * KBLK, x0 - 1, y0 - 1, img);
            if (++stage == Cfg::NA) {
              stage = 0;
              phase ^= 1u;
            }
            for (int tg = 0; tg < 9 / Cfg::TPS; ++tg) {
              ptx::mbar_wait(&bempty_bar[bstage], bphase ^ 1u);
              ptx::mbar_arrive_expect_tx(&bfull_bar[bstage], Cfg::BSTAGE_BYTES);
              uint8_t* sb = smem + Cfg::NA * HALO_SLOT + bstage * Cfg::BSTAGE_BYTES;
#pragma unroll
              for (int t = 0; t < Cfg::TPS; ++t) {
                const int tap = tg * Cfg::TPS + t;
                if constexpr (BMN) {
                  // backward-data reads the FORWARD-packed weights [Cout][tap][Cin] as an MN-major operand: N = Cin is
                  // contiguous, the reduction rows (Cout) are 9*Cin elements apart; the kernel is rotated: tap -> 8-tap
#pragma unroll
                  for (int b = 0; b < BLOCK_N / 64; ++b)
                    ptx::tma_load_2d(sb + t * Cfg::B_BYTES + b * 8192, &p.b_map, &bfull_bar[bstage],
                                     (8 - tap) * p.b_tap_cols + n0 + b * 64, cb * KBLK);
                } else {
                  ptx::tma_load_2d(sb + t * Cfg::B_BYTES, &p.b_map, &bfull_bar[bstage], tap * p.cin + cb * KBLK, n0);
                }
              }
              if (++bstage == Cfg::NB) {
                bstage = 0;
                bphase ^= 1u;
              }
            }
          }
        } else {
        for (int tap = 0; tap < p.taps; ++tap) {
          int dy = 0, dx = 0, mi = 0;
          if (p.tap_mode == 0) {
            if (p.taps == 9) {
              dy = tap / 3 - 1;
              dx = tap % 3 - 1;
            }
          } else {
            mi = tap;
          }
          for (int cb = 0; cb < cblocks; ++cb) {
            ptx::mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
            uint8_t* sb = sa + Cfg::A_BYTES;
            ptx::mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            ptx::tma_load_4d(sa, &p.a_map[mi], &full_bar[stage], cb * KBLK, x0 + dx, y0 + dy, img);
            if constexpr (BMN) {
              // ConvTranspose backward-data on the forward-packed weights [(kk,co)][Cin]: N = Cin contiguous,
              // reduction rows = (tap, co)
#pragma unroll
              for (int b = 0; b < BLOCK_N / 64; ++b)
                ptx::tma_load_2d(sb + b * 8192, &p.b_map, &full_bar[stage], n0 + b * 64, tap * p.cin + cb * KBLK);
            } else {
              ptx::tma_load_2d(sb, &p.b_map, &full_bar[stage], tap * p.cin + cb * KBLK, n0);
            }
            if (++stage == Cfg::STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_bf16(TILE_M, BLOCK_N, 0, BMN ? 1 : 0);
      // K-major B: +32 B per 16-element k-slice inside the 128-byte swizzle row.  MN-major B (BMN): 64-channel boxes
      // LBO = 8192 B apart, 16 reduction rows further = +2048 B
      constexpr uint32_t B_KSTEP16 = BMN ? 128 : 2;
      int stage = 0, bstage = 0;
      uint32_t phase = 0, bphase = 0;
      (void)bstage;
      (void)bphase;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (it >> 1) & 1;
        ptx::mbar_wait(&tmem_empty[as], aphase ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * BLOCK_N;
        if (HALO && p.resident) {
          const uint64_t adesc0 = ptx::make_smem_desc(0, 16, HALO_W * 128);
          const uint64_t bdesc0 = BMN ? ptx::make_smem_desc(0, 8192, 1024) : ptx::make_smem_desc(0, 16, 1024);
          const uint32_t sb16 = ptx::smem_u32(smem) >> 4;
          const uint32_t sa16 = sb16 + ((9 * cblocks * Cfg::B_BYTES) >> 4);
          if (it == 0) ptx::mbar_wait(&bfull_bar[0], 0);  // weights resident from here on
          for (int cb = 0; cb < cblocks; ++cb) {
            ptx::mbar_wait(&full_bar[stage], phase);
            ptx::tc_fence_after();
            const uint64_t adesc = adesc0 + (sa16 + stage * (HALO_SLOT >> 4));
            const uint64_t bdesc = bdesc0 + (sb16 + cb * (Cfg::B_BYTES >> 4));
            const uint32_t btap16 = static_cast<uint32_t>(cblocks) * (Cfg::B_BYTES >> 4);
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
              const uint32_t aoff = (((tap / 3) * HALO_W + tap % 3) * 128) >> 4;
#pragma unroll
              for (int k = 0; k < KBLK / 16; ++k)
                ptx::umma_bf16(d_tmem, adesc + (aoff + 2 * k), bdesc + (tap * btap16 + B_KSTEP16 * k), idesc,
                               (tap | k) != 0 ? 1u : static_cast<uint32_t>(cb != 0));
            }
            ptx::umma_commit(&empty_bar[stage]);
            if (++stage == p.resident) {
              stage = 0;
              phase ^= 1u;
            }
          }
          ptx::umma_commit(&tmem_full[as]);
          continue;
        }
        if constexpr (HALO) {
          // descriptor templates: everything but the 14-bit start-address field (shared memory addresses are < 256 KB,
          // so adding (address >> 4) never carries out of the field)
          const uint64_t adesc0 = ptx::make_smem_desc(0, 16, HALO_W * 128);
          const uint64_t bdesc0 = BMN ? ptx::make_smem_desc(0, 8192, 1024) : ptx::make_smem_desc(0, 16, 1024);
          const uint32_t sa16 = ptx::smem_u32(smem) >> 4;
          const uint32_t sb16 = (ptx::smem_u32(smem) + Cfg::NA * HALO_SLOT) >> 4;
          for (int cb = 0; cb < cblocks; ++cb) {
            ptx::mbar_wait(&full_bar[stage], phase);
            const uint64_t adesc = adesc0 + (sa16 + stage * (HALO_SLOT >> 4));
#pragma unroll
            for (int tg = 0; tg < 9 / Cfg::TPS; ++tg) {
              ptx::mbar_wait(&bfull_bar[bstage], bphase);
              ptx::tc_fence_after();
              const uint64_t bdesc = bdesc0 + (sb16 + bstage * (Cfg::BSTAGE_BYTES >> 4));
#pragma unroll
              for (int t = 0; t < Cfg::TPS; ++t) {
                const int tap = tg * Cfg::TPS + t;
                // output pixel (ty,tx) reads halo pixel (ty+ky, tx+kx): first row (ky*10+kx), row groups 10 pixels apart
                const uint32_t aoff = (((tap / 3) * HALO_W + tap % 3) * 128) >> 4;
                const uint32_t boff = (t * Cfg::B_BYTES) >> 4;
#pragma unroll
                for (int k = 0; k < KBLK / 16; ++k)
                  ptx::umma_bf16(d_tmem, adesc + (aoff + 2 * k), bdesc + (boff + B_KSTEP16 * k), idesc,
                                 (tap | k) != 0 ? 1u : static_cast<uint32_t>(cb != 0));
              }
              ptx::umma_commit(&bempty_bar[bstage]);
              if (++bstage == Cfg::NB) {
                bstage = 0;
                bphase ^= 1u;
              }
            }
            ptx::umma_commit(&empty_bar[stage]);
            if (++stage == Cfg::NA) {
              stage = 0;
              phase ^= 1u;
            }
          }
        } else {
        for (int ks = 0; ks < ksteps; ++ks) {
          ptx::mbar_wait(&full_bar[stage], phase);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t adesc = ptx::make_smem_desc(sa, 16, 1024);
          const uint64_t bdesc = BMN ? ptx::make_smem_desc(sa + Cfg::A_BYTES, 8192, 1024)
                                     : ptx::make_smem_desc(sa + Cfg::A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < KBLK / 16; ++k) {
            // +32 bytes (= 16 bf16) along K inside the 128-byte swizzle row: start-address field += 2
            ptx::umma_bf16(d_tmem, adesc + 2 * k, bdesc + B_KSTEP16 * k, idesc, (ks | k) != 0);
          }
          ptx::umma_commit(&empty_bar[stage]);
          if (++stage == Cfg::STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        }
        ptx::umma_commit(&tmem_full[as]);
      }
    }
  }
  } else {
    if constexpr (REALLOC) asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    // ===================== epilogue: 8 warps = 4 TMEM lane quarters x 2 column groups =====================
    const int e = threadIdx.x - 128;      // 0..255
    const int q = warp & 3;               // TMEM lane quarter this warp may read (hardware: warp id % 4)
    const int h = (warp - 4) >> 2;        // column group: this warp owns the 32-column chunks h, h+2, ...
    const int r = q * 32 + lane;          // tile row <-> pixel
    const int py = HALO ? (r >> 3) : (r >> 4), px = HALO ? (r & 7) : (r & 15);
    constexpr int POOL_Y_XOR = HALO ? 8 : 16;  // lane distance of the vertical 2x2-pool partner
    constexpr int NCHUNK = BLOCK_N / 32;
    // one n-tile for the whole kernel: scale/shift are loaded once, and with one chunk per warp the BN statistics
    // stay in per-thread registers until the CTA has finished all of its tiles (no shuffles in the tile loop)
    const bool fixed_n = (p.n_tiles == 1);
    const bool unit_scale = (p.scale == nullptr), identity = (p.scale == nullptr && p.shift == nullptr);
    // chunks per warp; with at most two the statistics of all of them fit in registers (the epilogue warps raise their
    // register allowance with setmaxnreg for this)
    constexpr int CPW = NCHUNK / EPI_COLGROUPS;
    constexpr int RUN_CPW = (HAS_STATS && CPW <= 2) ? CPW : 0;
    const bool run_stats = (RUN_CPW > 0) && fixed_n;
    float r1[RUN_CPW > 0 ? RUN_CPW : 1][32], r2[RUN_CPW > 0 ? RUN_CPW : 1][32];
    if (RUN_CPW > 0) {
#pragma unroll
      for (int ci = 0; ci < (RUN_CPW > 0 ? RUN_CPW : 1); ++ci)
#pragma unroll
        for (int j = 0; j < 32; ++j) r1[ci][j] = r2[ci][j] = 0.f;
    }
    auto load_affine = [&](int as, int n0) {
      float* sc = s_affine + as * 2 * BLOCK_N;
      float* sh = sc + BLOCK_N;
      for (int i = e; i < BLOCK_N; i += EPI_THREADS) {
        sc[i] = p.scale ? p.scale[n0 + i] : 1.0f;
        // ConvTranspose scatter: N index = (ky,kx,co) and the bias is per co
        sh[i] = p.shift ? p.shift[p.convt_cout > 0 ? (n0 + i) % p.convt_cout : n0 + i] : 0.0f;
      }
    };
    if (fixed_n) {
      load_affine(0, 0);
      epi_bar();
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (it >> 1) & 1;
      const int n_tile = tile % p.n_tiles;
      int m_tile = tile / p.n_tiles;
      const int tx = m_tile % p.tiles_x;
      m_tile /= p.tiles_x;
      const int ty = m_tile % p.tiles_y;
      const int img = m_tile / p.tiles_y;
      const int n0 = n_tile * BLOCK_N;
      const int y = ty * TH + py, x = tx * TW + px;
      const bool valid = (y < p.H) && (x < p.W);
      // fused head: the per-pixel dot product over all 64 channels must end up in ONE thread.  Instead of splitting the
      // columns between the two warps of a lane quarter and exchanging partial sums through shared memory (one block
      // barrier per tile: measured 70 us on the last layer), warp group h takes every second tile - accumulator stage
      // `as` == h - and reads both 32-column chunks itself.
      if (EPI == EPI_HEAD && as != h) continue;

      if (!fixed_n) {
        load_affine(as, n0);
        epi_bar();
      } else if (HAS_STATS && !run_stats) {
        epi_bar();  // the previous tile's s_red rows have been consumed
      }
      const float* sc = s_affine + (fixed_n ? 0 : as) * 2 * BLOCK_N;
      const float* sh = sc + BLOCK_N;

      ptx::mbar_wait(&tmem_full[as], aphase);
      ptx::tc_fence_after();

      float logit[CRIMAC_MAX_CLASSES];
      if (EPI == EPI_HEAD) {
#pragma unroll
        for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k) logit[k] = (k < p.n_classes) ? s_head[CRIMAC_MAX_CLASSES * 64 + k] : 0.f;
      }

      auto chunk_body = [&](const int chunk, float (&ra)[32], float (&rb)[32]) {
        uint32_t v[32];
        ptx::tmem_ld32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * BLOCK_N + chunk * 32, v);
        ptx::tmem_ld_wait();
        const int ng = n0 + chunk * 32;
        float f[32];
        if (EPI == EPI_STORE && identity) {
          // backward-data: no scale, no shift - the accumulator goes out as it is
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        } else if (EPI != EPI_HEAD && unit_scale) {
          // train-mode forward: conv bias only
          const float4* sh4 = reinterpret_cast<const float4*>(sh + chunk * 32);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 t = sh4[j4];
            f[4 * j4 + 0] = __uint_as_float(v[4 * j4 + 0]) + t.x;
            f[4 * j4 + 1] = __uint_as_float(v[4 * j4 + 1]) + t.y;
            f[4 * j4 + 2] = __uint_as_float(v[4 * j4 + 2]) + t.z;
            f[4 * j4 + 3] = __uint_as_float(v[4 * j4 + 3]) + t.w;
          }
        } else {
          const float4* sc4 = reinterpret_cast<const float4*>(sc + chunk * 32);
          const float4* sh4 = reinterpret_cast<const float4*>(sh + chunk * 32);
#pragma unroll
          for (int j4 = 0; j4 < 8; ++j4) {
            const float4 s = sc4[j4], t = sh4[j4];
            f[4 * j4 + 0] = fmaf(__uint_as_float(v[4 * j4 + 0]), s.x, t.x);
            f[4 * j4 + 1] = fmaf(__uint_as_float(v[4 * j4 + 1]), s.y, t.y);
            f[4 * j4 + 2] = fmaf(__uint_as_float(v[4 * j4 + 2]), s.z, t.z);
            f[4 * j4 + 3] = fmaf(__uint_as_float(v[4 * j4 + 3]), s.w, t.w);
          }
        }
        if (EPI != EPI_STATS && p.relu) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(f[2 * j], f[2 * j + 1]);

        if (EPI == EPI_HEAD) {
#pragma unroll
          for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k) {
            if (k < p.n_classes) {
              float acc = logit[k];
#pragma unroll
              for (int j = 0; j < 32; ++j) acc = fmaf(f[j], s_head[k * 64 + chunk * 32 + j], acc);
              logit[k] = acc;
            }
          }
        }

        if (p.out != nullptr && valid) {
          bf16* dst;
          if (p.convt_cout > 0) {
            const int kk = ng / p.convt_cout, co = ng - kk * p.convt_cout;
            const long opix = (static_cast<long>(img) * (2 * p.H) + (2 * y + (kk >> 1))) * (2 * p.W) + (2 * x + (kk & 1));
            dst = p.out + opix * p.out_pitch + co;
            if (EPI == EPI_STORE && p.convt_add) {
              // merge_mode "add" (unet.py:131-134): the destination already holds the encoder's skip activation; the
              // up-sampled value is added in fp32 and the sum rounded once
              const uint4* d4 = reinterpret_cast<const uint4*>(dst);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const uint4 o = d4[i];
                const uint32_t ow[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                  const float2 t = unpack_bf16x2(ow[w]);
                  pk[4 * i + w] = pack_bf16x2(f[8 * i + 2 * w] + t.x, f[8 * i + 2 * w + 1] + t.y);
                }
              }
            }
          } else {
            const long opix = (static_cast<long>(img) * p.H + y) * p.W + x;
            dst = (p.out2 != nullptr && ng >= p.out_split) ? p.out2 + opix * p.out2_pitch + (ng - p.out_split)
                                                           : p.out + opix * p.out_pitch + ng;
          }
          store32(dst, pk);
          store32(dst + 16, pk + 8);
        }

        if (EPI == EPI_STORE && p.pool_out != nullptr) {
          uint32_t pm[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            uint32_t m = bf16x2_max(pk[j], __shfl_xor_sync(0xffffffffu, pk[j], 1));
            pm[j] = bf16x2_max(m, __shfl_xor_sync(0xffffffffu, m, POOL_Y_XOR));
          }
          if (valid && !(px & 1) && !(py & 1)) {
            const long ppix = (static_cast<long>(img) * (p.H >> 1) + (y >> 1)) * (p.W >> 1) + (x >> 1);
            bf16* dst = p.pool_out + ppix * p.pool_pitch + ng;
            store32(dst, pm);
            store32(dst + 16, pm + 8);
          }
        }

        if (HAS_STATS) {
          // sum, sum of squares of the bf16-rounded values the BN-apply pass will read back
          if (RUN_CPW > 0 && run_stats) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 t = unpack_bf16x2(valid ? pk[j] : 0u);
              ra[2 * j] += t.x;
              ra[2 * j + 1] += t.y;
              rb[2 * j] = fmaf(t.x, t.x, rb[2 * j]);
              rb[2 * j + 1] = fmaf(t.y, t.y, rb[2 * j + 1]);
            }
          } else {
            float s1[32], s2[32];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 t = unpack_bf16x2(valid ? pk[j] : 0u);
              s1[2 * j] = t.x;
              s1[2 * j + 1] = t.y;
              s2[2 * j] = t.x * t.x;
              s2[2 * j + 1] = t.y * t.y;
            }
            xpose_reduce(s1, lane);
            xpose_reduce(s2, lane);
            s_red[(q * 2 + 0) * BLOCK_N + chunk * 32 + lane] = s1[0];
            s_red[(q * 2 + 1) * BLOCK_N + chunk * 32 + lane] = s2[0];
          }
        }
      };
      if constexpr (EPI == EPI_HEAD) {
#pragma unroll
        for (int chunk = 0; chunk < NCHUNK; ++chunk) chunk_body(chunk, r1[0], r2[0]);
      } else if constexpr (RUN_CPW > 0) {
#pragma unroll
        for (int ci = 0; ci < RUN_CPW; ++ci) chunk_body(h + ci * EPI_COLGROUPS, r1[ci], r2[ci]);
      } else {
#pragma unroll 1
        for (int chunk = h; chunk < NCHUNK; chunk += EPI_COLGROUPS) chunk_body(chunk, r1[0], r2[0]);
      }

      // accumulator fully read: hand the TMEM stage back to the MMA warp
      ptx::tc_fence_before();
      ptx::mbar_arrive(&tmem_empty[as]);

      if (HAS_STATS && !run_stats) {
        epi_bar();
        const int n_total = p.n_tiles * BLOCK_N;
        for (int c = e; c < BLOCK_N; c += EPI_THREADS) {
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            a += s_red[(w * 2 + 0) * BLOCK_N + c];
            b += s_red[(w * 2 + 1) * BLOCK_N + c];
          }
          // channel c of this n-tile always belongs to this thread: no race on the CTA-lifetime accumulators
          s_acc[n0 + c] += a;
          s_acc[n_total + n0 + c] += b;
        }
      }

      if (EPI == EPI_HEAD) {
        if (valid) {
          if (p.head_softmax) {
            float mx = logit[0];
#pragma unroll
            for (int k = 1; k < CRIMAC_MAX_CLASSES; ++k)
              if (k < p.n_classes) mx = fmaxf(mx, logit[k]);
            float sum = 0.f;
#pragma unroll
            for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k)
              if (k < p.n_classes) {
                logit[k] = __expf(logit[k] - mx);
                sum += logit[k];
              }
            const float inv = 1.f / sum;
#pragma unroll
            for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k) logit[k] *= inv;
          }
          if (p.head_stitch) {
            // fill_out_array + the label masks (overlap frame, outside the chunk, non-finite input, below seabed + pad on
            // background): the same decisions, in the same order, as stitch_kernel (pipeline_kernels.cu)
            const int ov = p.st_overlap;
            bool keep = y >= ov && y < p.H - ov && x >= ov && x < p.W - ov;
            const int y_upper = p.st_centres[2 * img] - p.H / 2 + 1;
            const int yd = y_upper + y;
            const int xl = p.st_centres[2 * img + 1] - p.W / 2 + 1 + x - p.st_ping_start;
            keep = keep && yd >= 0 && yd < p.st_r && xl >= 0 && xl < p.st_pc;
            if (keep && p.st_nan != nullptr) keep = p.st_nan[(static_cast<long>(img) * p.H + y) * p.W + x] == 0;
            if (keep) {
              const int l0 = p.st_labels ? p.st_labels[static_cast<long>(yd) * p.st_pc + xl] : 0;
              keep = !(l0 == -100 || l0 == -70 || l0 == -50);
              if (keep && p.st_seabed != nullptr && l0 == 0) {
                const int win0 = y_upper > 0 ? y_upper : 0;
                keep = !(yd - win0 >= p.st_seabed_pad && yd - p.st_seabed_pad >= p.st_seabed[xl]);
              }
            }
            if (keep) {
              __half* so = static_cast<__half*>(p.st_out);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                if (kk < p.st_k) {
                  float v = 0.f;
#pragma unroll
                  for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k)
                    if (k == p.st_cls[kk]) v = logit[k];
                  so[(static_cast<long>(kk) * p.st_r + yd) * p.st_pc + xl] = __float2half(v);
                }
            }
          } else {
#pragma unroll
            for (int k = 0; k < CRIMAC_MAX_CLASSES; ++k)
              if (k < p.n_classes)
                p.head_out[((static_cast<long>(img) * p.n_classes + k) * p.H + y) * p.W + x] = logit[k];
          }
        }
      }
    }

    if (HAS_STATS) {
      // one partial row per CTA: stats[blockIdx.x][2][n_total]; bn_finalize sums gridDim.x rows
      const int n_total = p.n_tiles * BLOCK_N;
      const int n2 = 2 * n_total;
      if (RUN_CPW > 0 && run_stats) {
#pragma unroll
        for (int ci = 0; ci < (RUN_CPW > 0 ? RUN_CPW : 1); ++ci) {
          xpose_reduce(r1[ci], lane);
          xpose_reduce(r2[ci], lane);
          s_red[(q * 2 + 0) * BLOCK_N + (h + ci * EPI_COLGROUPS) * 32 + lane] = r1[ci][0];
          s_red[(q * 2 + 1) * BLOCK_N + (h + ci * EPI_COLGROUPS) * 32 + lane] = r2[ci][0];
        }
        epi_bar();
        for (int c = e; c < BLOCK_N; c += EPI_THREADS) {
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int w = 0; w < 4; ++w) {
            a += s_red[(w * 2 + 0) * BLOCK_N + c];
            b += s_red[(w * 2 + 1) * BLOCK_N + c];
          }
          p.stats[static_cast<long>(blockIdx.x) * n2 + c] = a;
          p.stats[static_cast<long>(blockIdx.x) * n2 + n_total + c] = b;
        }
      } else {
        epi_bar();
        for (int i = e; i < n2; i += EPI_THREADS) p.stats[static_cast<long>(blockIdx.x) * n2 + i] = s_acc[i];
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

template <int BLOCK_N, int EPI, bool HALO, bool BMN>
cudaError_t launch_one(const ConvParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = ConvCfg<BLOCK_N, HALO>;
  auto kern = conv_igemm_kernel<BLOCK_N, EPI, HALO, BMN>;
  if (cudaError_t e = ensure_dynamic_smem(kern, Cfg::SMEM_BYTES); e != cudaSuccess) return e;
  if (EPI == EPI_STATS && p.n_tiles * BLOCK_N > Cfg::MAX_STAT_CH) return cudaErrorInvalidValue;
  const int grid = p.total_tiles < num_sms ? p.total_tiles : num_sms;
  if (HALO && BLOCK_N != 256) {
    // resident weights when the whole matrix fits beside two or three halo tiles
    const int wbytes = 9 * (p.cin / KBLK) * Cfg::B_BYTES;
    ConvParams q = p;
    q.resident = 0;
    if (p.n_tiles == 1 && wbytes <= Cfg::RES_MAX_B) q.resident = (wbytes + 3 * HALO_SLOT <= Cfg::OPERAND_BYTES) ? 3 : 2;
    kern<<<grid, CONV_THREADS, Cfg::SMEM_BYTES, stream>>>(q);
    return cudaGetLastError();
  }
  kern<<<grid, CONV_THREADS, Cfg::SMEM_BYTES, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace

// BLOCK_N in {64,128,256}; p.n_tiles*BLOCK_N == N_total must hold.
cudaError_t launch_conv_igemm(const ConvParams& p, int block_n, int epi, int num_sms, cudaStream_t stream) {
  if (p.halo && p.taps != 9) return cudaErrorInvalidValue;
  // the epilogue writes 32-byte (256-bit) vectors
  if (p.out && ((reinterpret_cast<uintptr_t>(p.out) & 31) || p.out_pitch % 16)) return cudaErrorInvalidValue;
  if (p.pool_out && ((reinterpret_cast<uintptr_t>(p.pool_out) & 31) || p.pool_pitch % 16)) return cudaErrorInvalidValue;
  if (p.out2 && ((reinterpret_cast<uintptr_t>(p.out2) & 31) || p.out2_pitch % 16 || p.out_split % 32 || p.convt_cout > 0 || epi != EPI_STORE))
    return cudaErrorInvalidValue;
  if (p.b_mn && (epi != EPI_STORE || (p.taps != 9 && p.taps != 4 && p.taps != 1))) return cudaErrorInvalidValue;
  if (p.b_mn && p.taps == 9 && !p.halo) return cudaErrorInvalidValue;  // 3x3 MN-major weights: halo main loop only
#define CASE(BN, EP)                                                               \
  if (block_n == BN && epi == EP && !p.b_mn)                                       \
    return p.halo ? launch_one<BN, EP, true, false>(p, num_sms, stream) : launch_one<BN, EP, false, false>(p, num_sms, stream);
  CASE(64, EPI_STORE) CASE(128, EPI_STORE) CASE(256, EPI_STORE)
  CASE(64, EPI_STATS) CASE(128, EPI_STATS) CASE(256, EPI_STATS)
  CASE(64, EPI_HEAD)
#undef CASE
#define CASE_MN(BN, EP)                                                            \
  if (block_n == BN && p.b_mn && epi == EP)                                        \
    return p.halo ? launch_one<BN, EP, true, true>(p, num_sms, stream)             \
                  : launch_one<BN, EP, false, true>(p, num_sms, stream);
  CASE_MN(64, EPI_STORE) CASE_MN(128, EPI_STORE) CASE_MN(256, EPI_STORE)
#undef CASE_MN
  return cudaErrorInvalidValue;
}

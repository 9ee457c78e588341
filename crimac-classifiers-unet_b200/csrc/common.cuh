// Shared host/device definitions for the echogram U-Net kernels (libcrimac_b200.so).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

typedef __nv_bfloat16 bf16;

// Spatial M-tile of every implicit-GEMM kernel: 8 range rows x 16 pings = 128 output pixels = the 128 TMEM lanes.
// A 2x2 max-pool window therefore always lives inside one warp's 32 lanes (rows 2w, 2w+1).
constexpr int TILE_W = 16;
constexpr int TILE_H = 8;
constexpr int TILE_M = 128;
constexpr int KBLK = 64;  // bf16 channels per k-step = one 128-byte swizzle row

// View of an NHWC bf16 activation tensor that may be a channel slice of a wider (concat) buffer.
struct View {
  bf16* ptr;    // first element of channel 0 of this view (slice offset already applied)
  int N, H, W;  // batch, range rows, pings
  int C;        // channels in the view
  int pitch;    // elements between consecutive pixels (>= C)
};

// ---- epilogue modes of conv_igemm
enum EpiMode : int {
  EPI_STORE = 0,  // y = acc*scale+shift, optional ReLU, optional 2x2 max-pool copy, optional convT scatter
  EPI_STATS = 1,  // raw = acc+shift stored as bf16 + per-tile per-channel sum / sum of squares (train-mode BN)
  EPI_HEAD = 2,   // folded BN + ReLU, then the 1x1 head (+softmax) per pixel; BLOCK_N must be all 64 channels
};

struct ConvParams {
  CUtensorMap a_map[4];  // activations; 3x3 / 1x1 use map 0, convT backward-data uses one map per (ky,kx)
  CUtensorMap b_map;     // packed weights [N_total][taps*cin] bf16, K contiguous (b_mn: the forward matrix, box {64,64})
  int b_mn;              // 1: backward-data on forward-packed weights (B operand MN-major); see conv_igemm.cu
  int b_tap_cols;        // b_mn, 3x3: columns per tap of the forward matrix (= its Cin = this GEMM's N_total)
  int taps;              // 9, 1 or 4
  int tap_mode;          // 0: tap -> (dy,dx) offsets on map 0 (3x3: tap=ky*3+kx; 1 tap: centre); 1: tap -> map index
  int halo;              // 1: 3x3 conv with the 16x8 tile / halo-box main loop (a_map[0] = box {64,10,18,1})
  int resident;          // set by the launcher: >0 = weights stay in shared memory, value = number of halo slots
  int cin;               // channels per tap, multiple of 64
  int NB, H, W;          // GEMM-M geometry: output pixels = NB*H*W
  int tiles_x, tiles_y, n_tiles, total_tiles;
  // epilogue
  bf16* out;
  int out_pitch;         // elements per output pixel
  bf16* out2;            // optional second destination: output channels >= out_split go to out2 (channel - out_split)
  int out2_pitch;        //   (backward-data of a conv that read a concat: the two halves of the concat gradient are
  int out_split;         //   kept as two DENSE tensors - every reader takes one half only)
  int relu;
  int convt_cout;        // >0: convT scatter, N index = (ky*2+kx)*convt_cout + co, output is 2H x 2W
  int convt_add;         // convT scatter ADDS to what the destination holds (merge_mode "add": the skip activation)
  const float* scale;    // [N_total]
  const float* shift;    // [N_total]
  bf16* pool_out;        // optional (EPI_STORE): 2x2 max-pooled copy, H/2 x W/2
  int pool_pitch;
  float* stats;          // EPI_STATS: [gridDim.x][2][N_total] per-CTA partial sums (sum, sum of squares)
  // EPI_HEAD
  const float* head_w;   // [n_classes][64]
  const float* head_b;   // [n_classes]
  float* head_out;       // NCHW fp32 (NB, n_classes, H, W): probabilities (softmax=1) or logits
  int n_classes;
  int head_softmax;
  // EPI_HEAD with on-device overlap stitching (sliding-window inference): instead of writing the probabilities of the
  // whole patch, every kept pixel's classes st_cls[0..st_k) go straight into the chunk's (K, R, Pc) fp16 output -
  // fill_out_array (save_predict.py:41-65) with the label masks of crimac_stitch, fused into the last conv's epilogue
  int head_stitch;
  const int* st_centres;       // (nb, 2) patch centres (y, x), survey coordinates
  const uint8_t* st_nan;       // (nb, H, W): 1 where frequency 0 was non-finite (or nullptr)
  const short* st_labels;      // (R, Pc) chunk labels or nullptr (= all background)
  const int* st_seabed;        // (Pc) or nullptr
  int st_seabed_pad, st_overlap, st_ping_start, st_pc, st_r, st_k;
  int st_cls[4];
  void* st_out;                // __half (K, R, Pc)
};

struct WgradParams {
  CUtensorMap a_map;     // "fixed" operand, rows of GEMM-M (conv: dY -> Cout; convT: X -> Cin)
  CUtensorMap b_map[4];  // "tap" operand, rows of GEMM-N (conv: X shifted by tap; convT: dY sub-sampled per tap)
  int taps, tap_mode;    // as ConvParams
  int M_total, N_total;  // channels of the two operands
  int NB, H, W;          // pixel geometry of the reduction dimension
  int tiles_x, tiles_y;  // 4x16-pixel k-tiles per image
  int k_tiles_total;     // NB*tiles_y*tiles_x
  int splits;            // split-K factor (gridDim.z)
  int m_tiles, n_tiles;
  float* dw;             // scratch [taps][M_total][N_total] fp32; split-K partial tiles accumulate with red.add ...
  float* slabs;          // ... or (deterministic mode, != nullptr) split s stores its partial tile into slabs + s*slab_stride
  long slab_stride;      //     (same [tap][m][n] layout per slab); the unpack kernel sums the slabs in fixed order
};

// Weight gradient of a 3x3 conv with ALL nine taps per CTA: the shifted operand (dY) is loaded once per k-step as a
// halo tile and read through nine row-shifted UMMA descriptors; taps are paired along the MMA M dimension.
struct WgradHaloParams {
  CUtensorMap s_map;     // shifted operand dY (NB,H,W,Cs): box {64, 18, 6, 1}
  CUtensorMap f_map;     // fixed operand X (NB,H,W,Cf):   box {64, 16, 4, 1}
  int Cs, Cf;            // Cout, Cin
  int NB, H, W;
  int tiles_x, tiles_y;  // 4x16-pixel k-tiles per image
  int k_tiles_total;
  int splits;
  int s_tiles, f_tiles;  // Cs/64, Cf/64
  int nf;                // 64: nine taps per CTA, 64-wide X tiles; 128: 128-wide X tiles, taps split over two CTA kinds
  float* dw;             // scratch [9][Cs][Cf] fp32
  float* slabs;          // deterministic mode: per-split slabs, see WgradParams
  long slab_stride;
};

#define CRIMAC_MAX_CLASSES 8

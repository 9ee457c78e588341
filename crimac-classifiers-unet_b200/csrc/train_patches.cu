// Training-sample path on the device: what Dataset.__getitem__ (batch/dataset.py:75-108) does
// per sample on the CPU DataLoader workers, for a whole batch in two launches, straight from the survey's sv / label
// arrays resident in HBM in the zarr store's own [frequency][ping][range] order:
//
//   get_crop_zarr (dataset.py:358-407)  ->  add_noise (data_augmentation/add_noise.py)  ->  flip_x_axis
//   ->  refine_label_boundary (label_transforms/refine_label_boundary.py)  ->  convert_label_indexing
//   ->  remove_nan_inf  ->  db_with_limits | db_with_limits_scaled   (batch/transforms.py:40-75)
//
//   train_gather_kernel   HBM-bound transpose-gather: 32 ping x 32 range tiles through shared memory (the store is
//                         range-major, the network wants ping-major), noise multiplier, flip, dB transform; the label
//                         pseudo-channel leaves a 4-bit code per sample (label class + "passes the school threshold")
//   train_labels_kernel   one cluster of 8 CTAs per crop (a band of range rows each): bounding box of the in-data
//                         samples exchanged through distributed shared memory, 7x7 disc closing on bit rows in
//                         shared memory, final int64 training labels written in place
#include "host_util.h"
#include "train_patch_core.h"
#include "../../include/crimac_b200.h"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace {

// grid.x = tiles of one (crop, channel) = (ph/32)*(pw/32); grid.y = F + 1 (channel F = labels); grid.z = crop.
// block (32, 8).  The per-thread bodies live in train_patch_core.h so that the host self-check runs the same code.
__global__ void __launch_bounds__(256) train_gather_kernel(const tp::GatherParams p) {
  __shared__ float tile[32][33];
  __shared__ float tile_lab[32][33];
  tp::TileCoord blk;
  blk.tile = blockIdx.x; blk.chan = blockIdx.y; blk.crop = blockIdx.z;
  const uint8_t fl = p.flags[blk.crop];
  tp::gather_load(p, blk, threadIdx.x, threadIdx.y, tile, tile_lab);
  __syncthreads();
  tp::gather_store(p, blk, threadIdx.x, threadIdx.y, fl, tile, tile_lab);
}

// One CLUSTER of tp::kBands thread blocks per crop, one band of ph/8 range rows per block.  The blocks exchange only
// their bounding-box partials (distributed shared memory); each builds its own halo rows of the threshold mask from
// the codes in global memory.  The second cluster barrier orders every read of the codes (own rows, halo rows) before
// any block overwrites them with the final labels, which is what makes the in-place update safe.
__global__ void __cluster_dims__(tp::kBands, 1, 1) __launch_bounds__(512)
train_labels_kernel(long long* lab, float* x, int F, int ph, int pw, int border_zero) {
  extern __shared__ uint32_t masks[];
  __shared__ int box[4];  // min y, max y, min x, max x over this band's samples that are not LABEL_BOUNDARY_VAL
  cg::cluster_group cluster = cg::this_cluster();
  const int band = static_cast<int>(cluster.block_rank());
  const int b = blockIdx.x / tp::kBands;
  const int rows = ph / tp::kBands, r0 = band * rows;
  const int wpr = pw >> 5;
  const int npx = ph * pw;
  uint32_t* T = masks;
  uint32_t* D = T + (rows + 12) * wpr;
  uint32_t* E = D + (rows + 6) * wpr;
  long long* L = lab + static_cast<long>(b) * npx;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;

  if (threadIdx.x == 0) {
    box[0] = ph; box[1] = -1; box[2] = pw; box[3] = -1;
  }
  __syncthreads();
  int ymin = ph, ymax = -1, xmin = pw, xmax = -1;
  const int n_t = (rows + 12) * wpr;
#pragma unroll 4
  for (int wi = warp; wi < n_t; wi += nwarps) {  // one warp = one mask word = 32 pings of one range row
    const int ly = wi / wpr, w = wi - ly * wpr;
    const int y = r0 - 6 + ly;
    int code = tp::L_BOUNDARY;
    if (y >= 0 && y < ph) code = static_cast<int>(L[static_cast<long>(y) * pw + 32 * w + lane]);
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, (code & tp::kThresholdBit) != 0);
    if (lane == 0) T[wi] = word;
    if (y >= r0 && y < r0 + rows && (code & 7) != tp::L_BOUNDARY) {  // refine_label_boundary.py:76-84
      const int xx = 32 * w + lane;
      ymin = min(ymin, y); ymax = max(ymax, y);
      xmin = min(xmin, xx); xmax = max(xmax, xx);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ymin = min(ymin, __shfl_xor_sync(0xFFFFFFFFu, ymin, o));
    ymax = max(ymax, __shfl_xor_sync(0xFFFFFFFFu, ymax, o));
    xmin = min(xmin, __shfl_xor_sync(0xFFFFFFFFu, xmin, o));
    xmax = max(xmax, __shfl_xor_sync(0xFFFFFFFFu, xmax, o));
  }
  if (lane == 0) {
    atomicMin(&box[0], ymin); atomicMax(&box[1], ymax);
    atomicMin(&box[2], xmin); atomicMax(&box[3], xmax);
  }
  __syncthreads();
  cluster.sync();  // every band's partial box is complete
  tp::BBox bb;
  bb.y0 = ph; bb.y1 = -1; bb.x0 = pw; bb.x1 = -1;
  for (int r = 0; r < tp::kBands; ++r) {
    const int* rb = cluster.map_shared_rank(box, r);
    bb.y0 = min(bb.y0, rb[0]); bb.y1 = max(bb.y1, rb[1]);
    bb.x0 = min(bb.x0, rb[2]); bb.x1 = max(bb.x1, rb[3]);
  }
  bb.y1 += 1;
  bb.x1 += 1;
  for (int i = threadIdx.x; i < (rows + 6) * wpr; i += blockDim.x) tp::labels_dilate_band(T, D, bb, wpr, r0, rows, i);
  __syncthreads();
  for (int i = threadIdx.x; i < rows * wpr; i += blockDim.x) tp::labels_erode_band(D, E, wpr, rows, i);
  __syncthreads();
  cluster.sync();  // all codes and all remote boxes have been read: the labels may be overwritten, blocks may exit
  float* x_crop = x + static_cast<long>(b) * F * npx;
  for (int i = threadIdx.x; i < rows * pw; i += blockDim.x)
    tp::labels_finish_band(L, x_crop, E, F, npx, r0 * pw, border_zero, i);
}

}  // namespace

extern "C" int crimac_train_patches(const float* sv, const float* labels, int F, int P, int R, const int32_t* centres,
                                    const uint8_t* flags, const float* noise_mult, uint64_t noise_seed, int n, int ph,
                                    int pw, int thr_freq, double thr_lo, double thr_hi, int scaled, int border_zero,
                                    float* x_out, int64_t* labels_out, void* stream) {
  CRIMAC_REQUIRE(sv && labels && centres && flags && x_out && labels_out, "NULL tensor");
  CRIMAC_REQUIRE(F >= 1 && P >= 1 && R >= 1, "empty survey");
  CRIMAC_REQUIRE(n >= 1 && n <= 65535, "patch count must be 1..65535");
  CRIMAC_REQUIRE(ph >= 32 && pw >= 32 && ph % 32 == 0 && pw % 32 == 0, "patch sides must be multiples of 32");
  CRIMAC_REQUIRE(ph == pw, "get_crop_zarr mixes the two window sides (dataset.py:397-398): square patches only");
  CRIMAC_REQUIRE(thr_freq >= 0 && thr_freq < F, "threshold frequency index out of range");
  const int smem = (3 * (ph / tp::kBands) + 18) * (pw / 32) * static_cast<int>(sizeof(uint32_t));
  constexpr int kMaxMaskBytes = 200 * 1024;
  CRIMAC_REQUIRE(smem <= kMaxMaskBytes, "patch too large for the on-chip label masks");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  tp::GatherParams p;
  p.sv = sv; p.labels = labels; p.centres = centres; p.flags = flags; p.noise = noise_mult;
  p.seed = noise_seed;
  p.F = F; p.P = P; p.R = R; p.n = n; p.ph = ph; p.pw = pw;
  p.thr_freq = thr_freq; p.thr_lo = thr_lo; p.thr_hi = thr_hi; p.scaled = scaled;
  p.x = x_out; p.lab = reinterpret_cast<long long*>(labels_out);
  const dim3 grid((ph / 32) * (pw / 32), F + 1, n);
  train_gather_kernel<<<grid, dim3(32, 8), 0, st>>>(p);
  CRIMAC_CHECK_CUDA(cudaGetLastError());
  CRIMAC_CHECK_CUDA(ensure_dynamic_smem(train_labels_kernel, kMaxMaskBytes));  // one opt-in covers every patch size
  train_labels_kernel<<<n * tp::kBands, 512, smem, st>>>(reinterpret_cast<long long*>(labels_out), x_out, F, ph, pw, border_zero);
  CRIMAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

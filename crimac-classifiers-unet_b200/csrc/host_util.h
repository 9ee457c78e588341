// Host-side helpers shared by the C-ABI translation units: error reporting, TMA tensor-map encoding,
// kernel launch prototypes.
#pragma once
#include "common.cuh"
#include <string>

// thread-local last-error string behind crimac_last_error()
void crimac_set_error(const std::string& msg);
#define CRIMAC_CHECK_CUDA(expr)                                                                           \
  do {                                                                                                    \
    cudaError_t _e = (expr);                                                                              \
    if (_e != cudaSuccess) {                                                                              \
      crimac_set_error(std::string(#expr) + " failed: " + cudaGetErrorString(_e) + " (" __FILE__ ":" +    \
                       std::to_string(__LINE__) + ")");                                                   \
      return 2;                                                                                           \
    }                                                                                                     \
  } while (0)
#define CRIMAC_REQUIRE(cond, msg)                                                                         \
  do {                                                                                                    \
    if (!(cond)) {                                                                                        \
      crimac_set_error(std::string("invalid argument: ") + (msg) + " [" #cond "]");                       \
      return 1;                                                                                           \
    }                                                                                                     \
  } while (0)

// 4-D map {C, W, H, N} over an NHWC bf16 view, box {64, 16, box_h, 1}, SWIZZLE_128B, zero OOB fill.
// sub = 0: plain view.  sub = 1: the (ky,kx) 2x2 sub-sampled view of a (2H x 2W) tensor, i.e. pixel (y,x) of the map
// is pixel (2y+ky, 2x+kx) of v (used for ConvTranspose2d backward).
int make_act_map(CUtensorMap* out, const View& v, int box_h, int sub = 0, int ky = 0, int kx = 0);
// 2-D map over packed weights [rows][cols] bf16 (cols contiguous), box {64, box_rows}.
int make_weight_map(CUtensorMap* out, const bf16* w, int rows, int cols, int box_rows);

cudaError_t launch_conv_igemm(const ConvParams& p, int block_n, int epi, int num_sms, cudaStream_t stream);
cudaError_t launch_wgrad_gemm(const WgradParams& p, int block_n, cudaStream_t stream);
cudaError_t launch_wgrad_unpack(const float* scratch, float* dw, int M, int N, int taps, int accumulate,
                                cudaStream_t stream);

int device_num_sms();

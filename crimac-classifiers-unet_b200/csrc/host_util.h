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
int make_act_map(CUtensorMap* out, const View& v, int box_h, int sub = 0, int ky = 0, int kx = 0, int box_w = 16,
                 int box_c = 64, int swizzle = 1);  // box_c = 8, swizzle = 0: un-swizzled 16-byte-per-pixel boxes (first conv)
int make_split_input_map(CUtensorMap* out, const bf16* xs, int parts, int NB, int H, int W);
// 2-D map over packed weights [rows][cols] bf16 (cols contiguous), box {64, box_rows}.
int make_weight_map(CUtensorMap* out, const bf16* w, int rows, int cols, int box_rows);

cudaError_t launch_conv_igemm(const ConvParams& p, int block_n, int epi, int num_sms, cudaStream_t stream);
inline int conv_grid(int total_tiles, int num_sms) { return total_tiles < num_sms ? total_tiles : num_sms; }
cudaError_t launch_wgrad_gemm(const WgradParams& p, int block_n, cudaStream_t stream);
cudaError_t launch_wgrad_halo(const WgradHaloParams& p, cudaStream_t stream);
cudaError_t launch_wgrad_unpack(const float* scratch, float* dw, int M, int N, int taps, int accumulate,
                                cudaStream_t stream);
// every layer's scratch -> PyTorch-layout gradient in one launch; the scratch is zeroed behind the read
struct UnpackEntry {
  float* scratch;  // [taps][mn]
  float* dw;       // [mn][taps]
  long mn;
  int taps;
  int item0;
  const float* slabs = nullptr;  // deterministic mode: `splits` partial copies of scratch, slab_stride floats apart,
  int splits = 0;                // summed in order instead of reading (and zeroing) scratch
  long slab_stride = 0;
};
struct UnpackTable {
  int n;
  UnpackEntry e[24];
};
cudaError_t launch_wgrad_unpack_all(UnpackTable& t, cudaStream_t stream);

int device_num_sms();
// Opt a kernel in to `bytes` of dynamic shared memory once per (kernel, device): the attribute is per device, and one
// process may drive several devices.  (Keyed by the function ADDRESS: kernels with the same signature share one type.)
cudaError_t ensure_dynamic_smem_impl(const void* kern, int bytes);
template <typename Kern>
inline cudaError_t ensure_dynamic_smem(Kern kern, int bytes) {
  return ensure_dynamic_smem_impl(reinterpret_cast<const void*>(kern), bytes);
}
bool crimac_profiling();  // per-launch event timing is on: kernels are then kept on ONE stream so that times are isolated

// ---- launch accounting + optional per-launch CUDA-event timing (crimac_profile_* in the C-ABI).
// Every kernel launch of the network-level entry points goes through a Scope: it always counts the launch and, when
// profiling is enabled, brackets it with cudaEvents on the launching stream (read back by crimac_profile_read).
struct ProfScope {
  ProfScope(const char* name, double flops, double bytes, cudaStream_t st, int launches = 1);
  ~ProfScope();
  cudaStream_t st_;
  int slot_;
};

// ---- CUDA-core kernels (elementwise.cu)
// xs: scratch of first_conv_split_elems() bf16 - the tensor-core path rewrites x there as bf16 hi/lo pairs (forward) and
// reads it back (weight gradient of the same input)
cudaError_t launch_first_conv(const float* x, bf16* xs, const float* w, const float* scale, const float* shift, int relu,
                              int NB, int cin, int H, int W, bf16* out, int out_pitch, float* stats, cudaStream_t st);
int first_conv_grid(int NB, int cin, int H, int W);  // = rows of the statistics partials the kernel writes
int first_conv_wgrad_blocks();
size_t first_conv_wgrad_partial_floats(int cin);  // scratch the weight-gradient launcher needs
size_t first_conv_split_elems(int NB, int cin, int H, int W);
// tensor-core variants (first_conv_tc.cu); the launchers above dispatch to them
int first_conv_tc_grid(int NB, int H, int W);
cudaError_t launch_first_conv_tc(const float* x, bf16* xs, const float* w, const float* scale, const float* shift,
                                 int relu, int NB, int cin, int H, int W, bf16* out, int out_pitch, float* stats,
                                 cudaStream_t st);
cudaError_t launch_first_conv_wgrad_tc(const bf16* xs, View draw, int cin, float* partials, float* dw, int accumulate,
                                       cudaStream_t st);
cudaError_t launch_first_conv_wgrad(const float* x, const bf16* xs, View draw, int cin, float* partials, float* dw,
                                    int accumulate, cudaStream_t st);
cudaError_t launch_bn_finalize(const float* partials, int m_tiles, int C, double count, const float* gamma,
                               const float* beta, float* rm, float* rv, long long* nbt, float momentum, float eps,
                               float* scale, float* shift, float* save_mean, float* save_invstd, cudaStream_t st);
cudaError_t launch_bn_fold_eval(const float* gamma, const float* beta, const float* rm, const float* rv,
                                const float* conv_bias, float eps, int C, float* scale, float* shift, cudaStream_t st);
// pool_arg (optional, with pool): uint16 per (pooled pixel, 8-channel group): 2-bit arg-max per channel for the backward
cudaError_t launch_bn_apply(View raw, const float* scale, const float* shift, View act, View pool, uint16_t* pool_arg,
                            cudaStream_t st);
cudaError_t launch_head_fwd(View act, const float* hw, const float* hb, int ncls, float* logits, cudaStream_t st);
int ce_blocks();
cudaError_t launch_ce(const float* logits, const long long* labels, const float* cw, int ncls, int NB, long HW,
                      long long ignore_index, float* dlogits, double* partials, float* out3, cudaStream_t st);
// validation path (pipeline.py:222-239,264,269-270): label remap + weighted CE + one class's softmax probability
cudaError_t launch_eval_loss(const float* logits, const void* labels, int label_bits, const float* cw, int ncls, int NB,
                             long HW, int prob_class, float* prob_out, long long* labels_out, double* partials,
                             float* out3, cudaStream_t st);
int head_bwd_blocks();
cudaError_t launch_head_bwd(const float* dlogits, const float* gscale, View act, const float* hw, int ncls, View dact,
                            float* partials, float* dw, float* db, int accumulate, cudaStream_t st);
// fused train step: 1x1 head + weighted CE + head backward in one pass; dact/dw/db relative to the UNNORMALISED loss
// gradient, out3 = {loss, 1/sum_w, sum_w}; loss_partials: head_bwd_blocks()*2 doubles
cudaError_t launch_head_ce_fused(View act, const float* hw, const float* hb, int ncls, const long long* labels,
                                 const float* cw, long long ignore_index, View dact, float* partials,
                                 double* loss_partials, float* dw, float* db, float* out3, cudaStream_t st);
// pipeline_kernels.cu: patch gather + dB transform written as the first conv's bf16 hi/lo operand (0 = ok, 2 = CUDA error)
int launch_preprocess_split(const float* sv, int F, int R, int P, int data_ping0, const int32_t* centres, int n, int ph,
                            int pw, bf16* xs, uint8_t* nan_mask, cudaStream_t st);
int reduce_blocks();
cudaError_t launch_bn_bwd(View dact, View raw, const float* scale, const float* shift, const float* mean,
                          const float* invstd, View draw, float* dgamma, float* dbeta, float* dbias, int accumulate,
                          float* partials, float* c1c2, const float* gscale, cudaStream_t st,
                          // optional: dact is not read; the incoming gradient is dskip + unpool(dpool) through the forward's
                          // arg-max map (max-pool backward + skip-gradient add fused into both passes)
                          const uint16_t* pool_arg = nullptr, View dpool = View{}, View dskip = View{});
cudaError_t launch_view_colsum(View v, float* partials, float* out, int accumulate, cudaStream_t st);
cudaError_t launch_pool_bwd_add(const uint16_t* pool_arg, View dpool, View dskip, View dact, cudaStream_t st);
// up_mode "upsample" (unet.py:50-56): bilinear 2x (align_corners=False) forward / adjoint on NHWC bf16 views, and the
// plain bf16 cast that packs a (Cout,Cin,1,1) weight
cudaError_t launch_upsample2x(View in, View out, cudaStream_t st);
cudaError_t launch_upsample2x_bwd(View dout, View din, cudaStream_t st);
cudaError_t launch_pack_cast(const float* w, bf16* out, long n, cudaStream_t st);
// fp32 parameters -> bf16 GEMM operands for every layer of one kind in ONE launch (item0 is filled by the launcher)
struct PackEntry {
  const float* w;
  bf16* out;
  int cout, cin;
  int item0;
};
struct PackTable {
  int n;
  PackEntry e[24];
};
cudaError_t launch_pack_conv3x3_all(PackTable& t, cudaStream_t st);  // (Cout,Cin,3,3) -> [Cout][tap][Cin]
cudaError_t launch_pack_convt_all(PackTable& t, cudaStream_t st);    // (Cin,Cout,2,2) -> [(kk*Cout+co)][Cin]

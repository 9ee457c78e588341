// HBM-bound kernels either side of the U-Net: patch gather + sv->dB transform (feeds the first conv), on-device
// overlap stitching of whole-echogram predictions, and the SGD-momentum step.
#include "host_util.h"
#include "../../include/crimac_b200.h"
#include <cuda_fp16.h>

namespace {

__device__ __forceinline__ uint32_t pack_bf16x2f(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// batch/dataset.py:192-205 (get_preload_data_labels) + utils/np.py:40-46,362-375 (getGrid, new_get_crop_3d) +
// remove_nan_inf.py:23-34 + db_with_limits.py:20-24,36-38.  One thread = 4 consecutive pings of one patch row.
__global__ void __launch_bounds__(256) preprocess_kernel(const float* __restrict__ sv, int F, int R, int P,
                                                         int data_ping0, const int* __restrict__ centres, int n,
                                                         int ph, int pw, float* __restrict__ out,
                                                         uint8_t* __restrict__ nan_mask) {
  const int pw4 = pw >> 2;
  const long total = static_cast<long>(n) * F * ph * pw4;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int x4 = static_cast<int>(i % pw4);
    long t = i / pw4;
    const int py = static_cast<int>(t % ph);
    t /= ph;
    const int f = static_cast<int>(t % F);
    const int b = static_cast<int>(t / F);
    const int cy = centres[2 * b], cx = centres[2 * b + 1];
    const int y = cy - ph / 2 + 1 + py;
    const int x0 = cx - pw / 2 + 1 + 4 * x4 - data_ping0;
    float v[4];
    uint8_t bad[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int x = x0 + j;
      float s = 0.f;  // out-of-data -> 0 BEFORE the dB transform (boundary_val_data, dataset.py:194)
      bool nf = false;
      if (y >= 0 && y < R && x >= 0 && x < P) {
        s = __ldg(&sv[(static_cast<long>(f) * R + y) * P + x]);
        nf = !isfinite(s);
        if (nf) s = 0.f;
      }
      float d = 10.f * log10f(s + 1e-10f);
      d = fminf(fmaxf(d, -75.f), 0.f);
      v[j] = d;
      bad[j] = nf ? 1 : 0;
    }
    const long o = ((static_cast<long>(b) * F + f) * ph + py) * pw + 4 * x4;
    *reinterpret_cast<float4*>(out + o) = make_float4(v[0], v[1], v[2], v[3]);
    if (f == 0 && nan_mask != nullptr)
      *reinterpret_cast<uchar4*>(nan_mask + (static_cast<long>(b) * ph + py) * pw + 4 * x4) =
          make_uchar4(bad[0], bad[1], bad[2], bad[3]);
  }
}

// The same gather + transform, but the result goes straight into the operand format of the first conv
// (first_conv_tc.cu): NHWC bf16 pairs hi = bf16(v), lo = bf16(v - hi) in 16-byte chunks per pixel - one plane
// [pixel][hi c0..3 | lo c0..3] for F <= 4, two planes [pixel][hi c0..7], [pixel][lo c0..7] for F <= 8.  The fp32 NCHW
// patch tensor (batch/dataset.py:192-205 -> pipeline.py:208) is never materialised.  One thread = one pixel, all
// frequencies: the reads of one frequency plane are coalesced along pings, the write is one 16-byte (two for P = 8) store.
template <int PCH>
__global__ void __launch_bounds__(256) preprocess_split_kernel(const float* __restrict__ sv, int F, int R, int P,
                                                               int data_ping0, const int* __restrict__ centres, int n,
                                                               int ph, int pw, bf16* __restrict__ xs,
                                                               uint8_t* __restrict__ nan_mask) {
  const long total = static_cast<long>(n) * ph * pw;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(i % pw);
    const int py = static_cast<int>((i / pw) % ph);
    const int b = static_cast<int>(i / (static_cast<long>(pw) * ph));
    const int y = centres[2 * b] - ph / 2 + 1 + py;
    const int x = centres[2 * b + 1] - pw / 2 + 1 + px - data_ping0;
    const bool inside = y >= 0 && y < R && x >= 0 && x < P;
    float hi[PCH], lo[PCH];
    bool bad0 = false;
#pragma unroll
    for (int f = 0; f < PCH; ++f) {
      float d = 0.f;
      if (f < F) {
        float s = 0.f;  // out-of-data -> 0 BEFORE the dB transform (boundary_val_data, dataset.py:194)
        if (inside) {
          s = __ldg(&sv[(static_cast<long>(f) * R + y) * P + x]);
          const bool nf = !isfinite(s);
          if (nf) s = 0.f;
          if (f == 0) bad0 = nf;
        }
        d = fminf(fmaxf(10.f * log10f(s + 1e-10f), -75.f), 0.f);
      }
      const float h = __bfloat162float(__float2bfloat16(d));
      hi[f] = h;
      lo[f] = d - h;
    }
    uint32_t w[PCH];
#pragma unroll
    for (int c = 0; c < PCH / 2; ++c) {
      w[c] = pack_bf16x2f(hi[2 * c], hi[2 * c + 1]);
      w[PCH / 2 + c] = pack_bf16x2f(lo[2 * c], lo[2 * c + 1]);
    }
    *reinterpret_cast<uint4*>(xs + i * 8) = make_uint4(w[0], w[1], w[2], w[3]);
    if (PCH == 8) *reinterpret_cast<uint4*>(xs + (total + i) * 8) = make_uint4(w[PCH - 4], w[PCH - 3], w[PCH - 2], w[PCH - 1]);
    if (nan_mask != nullptr) nan_mask[i] = bad0 ? 1 : 0;
  }
}

// save_predict.py:41-65 (fill_out_array) with the label semantics of mask_label_overlap.py:31-48,
// mask_label_seabed.py:32-68 (seabed_pad = 10, shifted inside the patch's clipped range window),
// new_get_crop_2d's LABEL_BOUNDARY_VAL outside [ping_start, ping_start+Pc) x [0,R), and remove_nan_inf.py:32.
__global__ void __launch_bounds__(256) stitch_kernel(const float* __restrict__ probs, int n, int ncls, int ph, int pw,
                                                     const int* __restrict__ centres,
                                                     const uint8_t* __restrict__ nan_mask,
                                                     const short* __restrict__ labels, const int* __restrict__ seabed,
                                                     int seabed_pad, int overlap, int ping_start, int Pc, int R,
                                                     int c0, int c1, int c2, int c3, int K, __half* __restrict__ out) {
  const int cls[4] = {c0, c1, c2, c3};
  const long total = static_cast<long>(n) * ph * pw;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(i % pw);
    const int py = static_cast<int>((i / pw) % ph);
    const int b = static_cast<int>(i / (static_cast<long>(pw) * ph));
    if (py < overlap || py >= ph - overlap || px < overlap || px >= pw - overlap) continue;  // -70 frame
    const int cy = centres[2 * b], cx = centres[2 * b + 1];
    const int y_upper = cy - ph / 2 + 1;
    const int y = y_upper + py;
    const int xl = cx - pw / 2 + 1 + px - ping_start;  // ping index inside the chunk
    if (y < 0 || y >= R || xl < 0 || xl >= Pc) continue;  // -100 outside the chunk's label slice
    if (nan_mask != nullptr && nan_mask[i]) continue;     // -100 where frequency 0 was non-finite
    const int l0 = labels ? labels[static_cast<long>(y) * Pc + xl] : 0;
    if (l0 == -100 || l0 == -70 || l0 == -50) continue;
    if (seabed != nullptr && l0 == 0) {
      const int win0 = y_upper > 0 ? y_upper : 0;  // the mask is padded inside the patch's clipped range window
      if (y - win0 >= seabed_pad && y - seabed_pad >= seabed[xl]) continue;  // -50 below seabed (+pad)
    }
    for (int k = 0; k < K; ++k)
      out[(static_cast<long>(k) * R + y) * Pc + xl] =
          __float2half(probs[((static_cast<long>(b) * ncls + cls[k]) * ph + py) * pw + px]);
  }
}

// v = momentum*v + g*gscale ; p -= lr*v  (torch.optim.SGD, dampening 0, pipeline.py:156,178).  16 bytes per parameter of
// HBM traffic: float4 loads / stores (n4 vectors), the last n % 4 elements by the first threads of block 0.
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, float* __restrict__ v,
                                                  const float* __restrict__ g, size_t n, float lr, float momentum,
                                                  float gscale) {
  const size_t n4 = n >> 2;
  float4* p4 = reinterpret_cast<float4*>(p);
  float4* v4 = reinterpret_cast<float4*>(v);
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 gv = __ldcs(g4 + i);
    float4 vv = v4[i], pv = p4[i];
    vv.x = momentum * vv.x + gv.x * gscale;
    vv.y = momentum * vv.y + gv.y * gscale;
    vv.z = momentum * vv.z + gv.z * gscale;
    vv.w = momentum * vv.w + gv.w * gscale;
    pv.x -= lr * vv.x;
    pv.y -= lr * vv.y;
    pv.z -= lr * vv.z;
    pv.w -= lr * vv.w;
    v4[i] = vv;
    p4[i] = pv;
  }
  if (blockIdx.x == 0) {
    const size_t i = (n4 << 2) + threadIdx.x;
    if (i < n) {
      const float nv = momentum * v[i] + g[i] * gscale;
      v[i] = nv;
      p[i] -= lr * nv;
    }
  }
}
// scalar form for arenas that are not 16-byte aligned
__global__ void __launch_bounds__(256) sgd_scalar_kernel(float* __restrict__ p, float* __restrict__ v,
                                                         const float* __restrict__ g, size_t n, float lr, float momentum,
                                                         float gscale) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float nv = momentum * v[i] + g[i] * gscale;
    v[i] = nv;
    p[i] -= lr * nv;
  }
}

}  // namespace

extern "C" int crimac_preprocess(const float* sv, int F, int R, int P, int data_ping0, const int32_t* centres, int n,
                                 int ph, int pw, float* out, uint8_t* nan_mask, void* stream) {
  CRIMAC_REQUIRE(sv && centres && out, "NULL tensor");
  CRIMAC_REQUIRE(F >= 1 && R >= 1 && P >= 1 && n >= 1, "empty input");
  CRIMAC_REQUIRE(ph % 2 == 0 && pw % 4 == 0, "patch height must be even and width a multiple of 4");
  const long total = static_cast<long>(n) * F * ph * (pw / 4);
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  preprocess_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      sv, F, R, P, data_ping0, centres, n, ph, pw, out, nan_mask);
  CRIMAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// crimac_preprocess with the first conv's operand as the destination (see preprocess_split_kernel); xs: the staging
// buffer of a context (crimac_preprocess_staged in net_api.cu passes it) holding n patches of (ph, pw).
int launch_preprocess_split(const float* sv, int F, int R, int P, int data_ping0, const int32_t* centres, int n, int ph,
                            int pw, bf16* xs, uint8_t* nan_mask, cudaStream_t st) {
  const long total = static_cast<long>(n) * ph * pw;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (F <= 4)
    preprocess_split_kernel<4><<<static_cast<int>(blocks), 256, 0, st>>>(sv, F, R, P, data_ping0, centres, n, ph, pw, xs, nan_mask);
  else
    preprocess_split_kernel<8><<<static_cast<int>(blocks), 256, 0, st>>>(sv, F, R, P, data_ping0, centres, n, ph, pw, xs, nan_mask);
  return cudaGetLastError() == cudaSuccess ? 0 : 2;
}

extern "C" int crimac_stitch(const float* probs, int n, int n_classes, int ph, int pw, const int32_t* centres,
                             const uint8_t* nan_mask, const int16_t* labels, const int32_t* seabed, int seabed_pad,
                             int overlap, int ping_start, int Pc, int R, const int32_t* cls, int K, void* out,
                             void* stream) {
  CRIMAC_REQUIRE(probs && centres && out && cls, "NULL tensor");
  CRIMAC_REQUIRE(K >= 1 && K <= 4, "K must be 1..4");
  for (int k = 0; k < K; ++k) CRIMAC_REQUIRE(cls[k] >= 0 && cls[k] < n_classes, "class index out of range");
  const long total = static_cast<long>(n) * ph * pw;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  stitch_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      probs, n, n_classes, ph, pw, centres, nan_mask, labels, seabed, seabed_pad, overlap, ping_start, Pc, R, cls[0],
      K > 1 ? cls[1] : 0, K > 2 ? cls[2] : 0, K > 3 ? cls[3] : 0, K, static_cast<__half*>(out));
  CRIMAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int crimac_sgd_step(float* params, float* mom, const float* grads, size_t n, float lr, float momentum,
                               float gscale, void* stream) {
  CRIMAC_REQUIRE(params && mom && grads, "NULL tensor");
  if (n == 0) return 0;
  const bool aligned = ((reinterpret_cast<uintptr_t>(params) | reinterpret_cast<uintptr_t>(mom) | reinterpret_cast<uintptr_t>(grads)) & 15) == 0;
  size_t blocks = ((aligned ? (n >> 2) + 1 : n) + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (aligned)
    sgd_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(params, mom, grads, n, lr, momentum, gscale);
  else
    sgd_scalar_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(params, mom, grads, n, lr, momentum, gscale);
  CRIMAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Validation step of the training loop (pipeline.py:249-270, get_predictions_dataloader): on the eval-mode logits,
// remap the label codes (set_label_ignore_val, :222-239), evaluate the same class-weighted CE (:264) and extract the
// softmax probability of one class (SANDEEL, :269-270) - one pass, no host work.  scratch: >= 16 KB.
extern "C" int crimac_eval_loss(const float* logits, int nb, int n_classes, int H, int W, const void* labels,
                                int label_bits, const float* class_w, int prob_class, float* prob_out,
                                int64_t* labels_out, float* out3, void* scratch, void* stream) {
  CRIMAC_REQUIRE(logits && labels && class_w && out3 && scratch, "NULL tensor");
  CRIMAC_REQUIRE(nb >= 1 && H >= 1 && W >= 1, "empty input");
  CRIMAC_REQUIRE(n_classes >= 1 && n_classes <= CRIMAC_MAX_CLASSES, "n_classes must be 1..8");
  CRIMAC_REQUIRE(label_bits == 16 || label_bits == 64, "labels must be int16 or int64");
  CRIMAC_REQUIRE(prob_out == nullptr || (prob_class >= 0 && prob_class < n_classes), "prob_class out of range");
  CRIMAC_CHECK_CUDA(launch_eval_loss(logits, labels, label_bits, class_w, n_classes, nb, static_cast<long>(H) * W,
                                     prob_class, prob_out, reinterpret_cast<long long*>(labels_out),
                                     static_cast<double*>(scratch), out3, static_cast<cudaStream_t>(stream)));
  return 0;
}

// Metadata input channels of get_crop_memmap (batch/dataset.py:296-349; selected by data.meta_channels in the config,
// pipeline.py:413-425), generated on the device straight into the network input tensor.  Channel order as the reference
// appends them: portion_year, sin / cos of portion_day, time_diff, depth_rel, depth_abs_surface, depth_abs_seabed.
// Index conventions are the reference's own (NOT the data crop's): rows cy - ph/2 .. cy + ph/2 - 1 (unclamped), pings
// cx - pw/2 .. cx + pw/2 - 1 with negative -> 0 and >= size -> the LAST element; arithmetic in double, stored as fp32.
namespace {
__global__ void __launch_bounds__(256) meta_channels_kernel(const int* __restrict__ centres, int n, int ph, int pw,
                                                            unsigned mask, double portion_year,
                                                            const double* __restrict__ pod, int n_pod,
                                                            const double* __restrict__ tvd, int n_tvd,
                                                            const double* __restrict__ seabed, int n_sb,
                                                            float* __restrict__ out, int c_total, int c_off) {
  const long total = static_cast<long>(n) * ph * pw;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int px = static_cast<int>(i % pw);
    const int py = static_cast<int>((i / pw) % ph);
    const int b = static_cast<int>(i / (static_cast<long>(pw) * ph));
    const int cy = centres[2 * b], cx = centres[2 * b + 1];
    const long plane = static_cast<long>(ph) * pw;
    float* o = out + (static_cast<long>(b) * c_total + c_off) * plane + static_cast<long>(py) * pw + px;
    auto clampi = [](int idx, int size) { return idx < 0 ? 0 : (idx >= size ? size - 1 : idx); };
    int ch = 0;
    if (mask & 1u) o[(ch++) * plane] = static_cast<float>(portion_year);
    if (mask & 2u) {
      const double t = pod[clampi(cx, n_pod)];
      o[(ch++) * plane] = static_cast<float>(sin(2.0 * 3.141592653589793 * t));
      o[(ch++) * plane] = static_cast<float>(cos(2.0 * 3.141592653589793 * t));
    }
    const int col = cx - pw / 2 + px;
    if (mask & 4u) o[(ch++) * plane] = static_cast<float>(tvd[clampi(col, n_tvd)]);
    if (mask & 56u) {
      const double row = static_cast<double>(cy - ph / 2 + py);
      const double sb = seabed[clampi(col, n_sb)];
      if (mask & 8u) o[(ch++) * plane] = static_cast<float>(row / sb);
      if (mask & 16u) o[(ch++) * plane] = static_cast<float>(row / static_cast<double>(ph));
      if (mask & 32u) o[(ch++) * plane] = static_cast<float>((sb - row) / static_cast<double>(ph));
    }
  }
}
}  // namespace

extern "C" int crimac_meta_channels(const int32_t* centres, int n, int ph, int pw, unsigned mask, double portion_year,
                                    const double* portion_of_day, int n_pod, const double* time_diff, int n_tvd,
                                    const double* seabed, int n_sb, float* x_out, int c_total, int c_off, void* stream) {
  CRIMAC_REQUIRE(centres && x_out && n >= 1 && ph >= 1 && pw >= 1, "bad argument");
  CRIMAC_REQUIRE(mask != 0 && mask < 64, "mask selects 1..6 channel kinds");
  CRIMAC_REQUIRE(!(mask & 2u) || (portion_of_day && n_pod >= 1), "portion_of_day vector missing");
  CRIMAC_REQUIRE(!(mask & 4u) || (time_diff && n_tvd >= 1), "time_diff vector missing");
  CRIMAC_REQUIRE(!(mask & 56u) || (seabed && n_sb >= 1), "seabed vector missing");
  int m = 0;
  for (unsigned bit : {1u, 4u, 8u, 16u, 32u}) m += (mask & bit) ? 1 : 0;
  m += (mask & 2u) ? 2 : 0;
  CRIMAC_REQUIRE(c_off >= 0 && c_off + m <= c_total, "channel offset + metadata channels exceed the tensor");
  const long total = static_cast<long>(n) * ph * pw;
  long blocks = (total + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  meta_channels_kernel<<<static_cast<int>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      centres, n, ph, pw, mask, portion_year, portion_of_day, n_pod, time_diff, n_tvd, seabed, n_sb, x_out, c_total, c_off);
  CRIMAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

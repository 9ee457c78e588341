// fp32 VALIDATION mode of the eval forward (north_star: "1e-4 in an fp32 validation mode").
//
// A deliberately independent second implementation of reference models/unet.py:327-343 + pipeline.py:218: plain fp32
// CUDA-core kernels on NCHW tensors, reading the module's fp32 parameters directly (no packing, no BatchNorm folding
// into weights, no tensor cores, no bf16 anywhere).  It exists to separate bf16 rounding from logic errors: the
// production path is compared with the oracle at 2e-2, this one at 1e-4.  It is ~50x slower than the tcgen05 path and
// is not used by any product entry point.
#include "host_util.h"
#include "../../include/crimac_b200.h"
#include <vector>

namespace {

// 3x3 conv (pad 1) + bias + eval BatchNorm + ReLU.  Block = 16x16 output pixels of one image x 8 output channels;
// input channels are staged 8 at a time through shared memory (18x18 halo tile) together with their 8x8x9 weights.
__global__ void __launch_bounds__(256) conv3x3_bn_relu_fp32_kernel(const float* __restrict__ x, long x_bstride, int Cin,
                                                                   const float* __restrict__ w,
                                                                   const float* __restrict__ bias,
                                                                   const float* __restrict__ gamma,
                                                                   const float* __restrict__ beta,
                                                                   const float* __restrict__ rm,
                                                                   const float* __restrict__ rv, float eps, int Cout,
                                                                   int H, int W, float* __restrict__ out,
                                                                   long out_bstride) {
  __shared__ float s_in[8][18][18];
  __shared__ float s_w[8][8][9];  // [co][ci][tap]
  const int tiles_x = (W + 15) / 16;
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int co0 = blockIdx.y * 8, n = blockIdx.z;
  const int lx = threadIdx.x & 15, ly = threadIdx.x >> 4;
  const int ox = tx * 16 + lx, oy = ty * 16 + ly;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
  for (int ci0 = 0; ci0 < Cin; ci0 += 8) {
    for (int i = threadIdx.x; i < 8 * 18 * 18; i += 256) {
      const int c = i / 324, r = (i % 324) / 18, q = i % 18;
      const int gy = ty * 16 + r - 1, gx = tx * 16 + q - 1;
      float v = 0.f;
      if (ci0 + c < Cin && gy >= 0 && gy < H && gx >= 0 && gx < W)
        v = x[n * x_bstride + (static_cast<long>(ci0 + c) * H + gy) * W + gx];
      s_in[c][r][q] = v;
    }
    for (int i = threadIdx.x; i < 8 * 8 * 9; i += 256) {
      const int k = i / 72, c = (i % 72) / 9, t = i % 9;
      s_w[k][c][t] = (co0 + k < Cout && ci0 + c < Cin) ? w[(static_cast<long>(co0 + k) * Cin + ci0 + c) * 9 + t] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float v[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) v[t] = s_in[c][ly + t / 3][lx + t % 3];
#pragma unroll
      for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int t = 0; t < 9; ++t) acc[k] = fmaf(v[t], s_w[k][c][t], acc[k]);
    }
    __syncthreads();
  }
  if (ox < W && oy < H) {
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int co = co0 + k;
      if (co < Cout) {
        const float inv = 1.0f / sqrtf(rv[co] + eps);
        const float y = (acc[k] + bias[co] - rm[co]) * inv * gamma[co] + beta[co];
        out[n * out_bstride + (static_cast<long>(co) * H + oy) * W + ox] = fmaxf(y, 0.f);
      }
    }
  }
}

__global__ void maxpool2_fp32_kernel(const float* __restrict__ x, long x_bstride, int C, int H, int W,
                                     float* __restrict__ out, int N) {
  const int Ho = H / 2, Wo = W / 2;
  const long total = static_cast<long>(N) * C * Ho * Wo;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int xo = i % Wo, yo = (i / Wo) % Ho, c = (i / (static_cast<long>(Wo) * Ho)) % C;
    const int n = i / (static_cast<long>(Wo) * Ho * C);
    const float* p = x + n * x_bstride + (static_cast<long>(c) * H + 2 * yo) * W + 2 * xo;
    out[i] = fmaxf(fmaxf(p[0], p[1]), fmaxf(p[W], p[W + 1]));
  }
}

// ConvTranspose2d(k=2, s=2): out[n][co][2y+ky][2x+kx] = b[co] + sum_ci x[n][ci][y][x] * w[ci][co][ky][kx]
__global__ void convt2x2_fp32_kernel(const float* __restrict__ x, int Cin, int h, int w_, const float* __restrict__ wt,
                                     const float* __restrict__ bias, int Cout, float* __restrict__ out,
                                     long out_bstride, int N) {
  const int Ho = 2 * h, Wo = 2 * w_;
  const long total = static_cast<long>(N) * Cout * Ho * Wo;
  for (long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<long>(gridDim.x) * blockDim.x) {
    const int xo = i % Wo, yo = (i / Wo) % Ho, co = (i / (static_cast<long>(Wo) * Ho)) % Cout;
    const int n = i / (static_cast<long>(Wo) * Ho * Cout);
    const int kk = (yo & 1) * 2 + (xo & 1);
    const float* xp = x + (static_cast<long>(n) * Cin * h + (yo >> 1)) * w_ + (xo >> 1);
    float a = bias[co];
    for (int ci = 0; ci < Cin; ++ci)
      a = fmaf(xp[static_cast<long>(ci) * h * w_], wt[(static_cast<long>(ci) * Cout + co) * 4 + kk], a);
    out[n * out_bstride + (static_cast<long>(co) * Ho + yo) * Wo + xo] = a;
  }
}

__global__ void head_softmax_fp32_kernel(const float* __restrict__ x, int C, long HW, const float* __restrict__ w,
                                         const float* __restrict__ b, int ncls, int softmax, float* __restrict__ out,
                                         int N) {
  const long total = static_cast<long>(N) * HW;
  for (long p = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; p < total;
       p += static_cast<long>(gridDim.x) * blockDim.x) {
    const long n = p / HW, r = p - n * HW;
    float z[CRIMAC_MAX_CLASSES];
    for (int k = 0; k < ncls; ++k) {
      float a = b[k];
      for (int c = 0; c < C; ++c) a = fmaf(x[(n * C + c) * HW + r], w[k * C + c], a);
      z[k] = a;
    }
    if (softmax) {
      float mx = z[0];
      for (int k = 1; k < ncls; ++k) mx = fmaxf(mx, z[k]);
      float s = 0.f;
      for (int k = 0; k < ncls; ++k) {
        z[k] = expf(z[k] - mx);
        s += z[k];
      }
      for (int k = 0; k < ncls; ++k) z[k] /= s;
    }
    for (int k = 0; k < ncls; ++k) out[(n * ncls + k) * HW + r] = z[k];
  }
}

template <typename T>
const T* S(const void* const* state, int i) {
  return static_cast<const T*>(state[i]);
}

struct Plan {
  size_t floats = 0;
  size_t take(size_t n) {
    const size_t o = floats;
    floats += (n + 63) & ~static_cast<size_t>(63);
    return o;
  }
};

int check_cfg(const crimac_config* cfg, int nb) {
  CRIMAC_REQUIRE(cfg != nullptr, "cfg is NULL");
  CRIMAC_REQUIRE(cfg->depth >= 2 && cfg->depth <= 5 && cfg->start_filts >= 1, "depth must be 2..5");
  CRIMAC_REQUIRE(cfg->n_classes >= 1 && cfg->n_classes <= CRIMAC_MAX_CLASSES, "n_classes must be 1..8");
  CRIMAC_REQUIRE(nb >= 1, "nb");
  const int m = 1 << (cfg->depth - 1);
  CRIMAC_REQUIRE(cfg->height % m == 0 && cfg->width % m == 0, "height/width must be multiples of 2^(depth-1)");
  return 0;
}

}  // namespace

extern "C" int crimac_fp32_workspace_bytes(const crimac_config* cfg, int nb, size_t* bytes) {
  int rc = check_cfg(cfg, nb);
  if (rc) return rc;
  CRIMAC_REQUIRE(bytes != nullptr, "bytes is NULL");
  const int D = cfg->depth;
  Plan pl;
  for (int l = 0; l < D; ++l) {
    const size_t px = static_cast<size_t>(nb) * (cfg->height >> l) * (cfg->width >> l), C = cfg->start_filts << l;
    pl.take(px * C);                      // enc conv1
    pl.take(px * (l < D - 1 ? 2 * C : C));  // concat buffer (skip half) / deepest conv2
    if (l < D - 1) {
      pl.take(px / 4 * C);  // pooled
      pl.take(px * C);      // dec conv1
      pl.take(px * C);      // dec conv2
    }
  }
  *bytes = pl.floats * sizeof(float);
  return 0;
}

extern "C" int crimac_forward_infer_fp32(const crimac_config* cfg, const void* const* state, const float* x, int nb,
                                         float* out, int softmax, void* workspace_dev, size_t workspace_bytes,
                                         void* stream) {
  int rc = check_cfg(cfg, nb);
  if (rc) return rc;
  CRIMAC_REQUIRE(state && x && out && workspace_dev, "NULL argument");
  size_t need = 0;
  crimac_fp32_workspace_bytes(cfg, nb, &need);
  CRIMAC_REQUIRE(workspace_bytes >= need, "fp32 validation workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int D = cfg->depth;
  float* ws = static_cast<float*>(workspace_dev);
  Plan pl;
  std::vector<float*> e1(D), cat(D), pooled(D), d1(D), d2(D);
  for (int l = 0; l < D; ++l) {
    const size_t px = static_cast<size_t>(nb) * (cfg->height >> l) * (cfg->width >> l), C = cfg->start_filts << l;
    e1[l] = ws + pl.take(px * C);
    cat[l] = ws + pl.take(px * (l < D - 1 ? 2 * C : C));
    if (l < D - 1) {
      pooled[l] = ws + pl.take(px / 4 * C);
      d1[l] = ws + pl.take(px * C);
      d2[l] = ws + pl.take(px * C);
    }
  }
  auto run_conv = [&](const float* in, long in_bs, int Cin, int s_w, int s_b, int s_bn, int Cout, int H, int W, float* o,
                      long o_bs) -> int {
    dim3 grid(((W + 15) / 16) * ((H + 15) / 16), (Cout + 7) / 8, nb);
    conv3x3_bn_relu_fp32_kernel<<<grid, 256, 0, st>>>(in, in_bs, Cin, S<float>(state, s_w), S<float>(state, s_b),
                                                      S<float>(state, s_bn), S<float>(state, s_bn + 1),
                                                      S<float>(state, s_bn + 2), S<float>(state, s_bn + 3), 1e-5f, Cout,
                                                      H, W, o, o_bs);
    CRIMAC_CHECK_CUDA(cudaGetLastError());
    return 0;
  };
  // state-table order: encoder block i: conv1 {w,b}, bn1 {w,b,rm,rv,nbt}, conv2 {w,b}, bn2 {...5}
  int si = 0;
  const float* cur = x;
  long cur_bs = static_cast<long>(cfg->in_channels) * cfg->height * cfg->width;
  int cur_c = cfg->in_channels;
  for (int l = 0; l < D; ++l) {
    const int H = cfg->height >> l, W = cfg->width >> l, C = cfg->start_filts << l;
    const long px = static_cast<long>(H) * W;
    if ((rc = run_conv(cur, cur_bs, cur_c, si, si + 1, si + 2, C, H, W, e1[l], C * px))) return rc;
    // conv2 writes the skip half [C,2C) of the level's concat buffer (deepest level: a plain buffer)
    float* o2 = (l < D - 1) ? cat[l] + C * px : cat[l];
    const long o2_bs = (l < D - 1) ? 2 * C * px : C * px;
    if ((rc = run_conv(e1[l], C * px, C, si + 7, si + 8, si + 9, C, H, W, o2, o2_bs))) return rc;
    si += 14;
    if (l < D - 1) {
      const long total = static_cast<long>(nb) * C * (px / 4);
      maxpool2_fp32_kernel<<<static_cast<int>((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256), 256, 0, st>>>(
          o2, o2_bs, C, H, W, pooled[l], nb);
      CRIMAC_CHECK_CUDA(cudaGetLastError());
      cur = pooled[l];
      cur_bs = C * (px / 4);
      cur_c = C;
    } else {
      cur = cat[l];
      cur_bs = C * px;
      cur_c = C;
    }
  }
  // decoder block j (level l = D-2-j): upconv {w,b}, conv1 {w,b}, conv2 {w,b}, bn1 {...5}, bn2 {...5}
  for (int j = 0; j < D - 1; ++j) {
    const int l = D - 2 - j;
    const int H = cfg->height >> l, W = cfg->width >> l, C = cfg->start_filts << l;
    const long px = static_cast<long>(H) * W;
    {
      const long total = static_cast<long>(nb) * C * px;
      convt2x2_fp32_kernel<<<static_cast<int>((total + 255) / 256 > 8192 ? 8192 : (total + 255) / 256), 256, 0, st>>>(
          cur, 2 * C, H / 2, W / 2, S<float>(state, si), S<float>(state, si + 1), C, cat[l], 2 * C * px, nb);
      CRIMAC_CHECK_CUDA(cudaGetLastError());
    }
    if ((rc = run_conv(cat[l], 2 * C * px, 2 * C, si + 2, si + 3, si + 6, C, H, W, d1[l], C * px))) return rc;
    if ((rc = run_conv(d1[l], C * px, C, si + 4, si + 5, si + 11, C, H, W, d2[l], C * px))) return rc;
    si += 16;
    cur = d2[l];
  }
  {
    const long HW = static_cast<long>(cfg->height) * cfg->width;
    const long total = nb * HW;
    head_softmax_fp32_kernel<<<static_cast<int>((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256), 256, 0, st>>>(
        cur, cfg->start_filts, HW, S<float>(state, si), S<float>(state, si + 1), cfg->n_classes, softmax, out, nb);
    CRIMAC_CHECK_CUDA(cudaGetLastError());
  }
  return 0;
}

// Small device helpers shared by the tensor-core and the CUDA-core kernels.
#pragma once
#include "common.cuh"

template <int STEP>
__device__ __forceinline__ void xpose_reduce_step(float (&s)[32], int lane) {
  const bool up = (lane & STEP) != 0;
#pragma unroll
  for (int i = 0; i < STEP; ++i) {
    const float send = up ? s[i] : s[i + STEP];
    const float keep = up ? s[i + STEP] : s[i];
    s[i] = keep + __shfl_xor_sync(0xffffffffu, send, STEP);
  }
}
// Transposing warp reduction: after the call lane L holds in s[0] the sum over all 32 lanes of the original s[L]
// (31 shuffles for 32 sums instead of 160).
__device__ __forceinline__ void xpose_reduce(float (&s)[32], int lane) {
  xpose_reduce_step<16>(s, lane);
  xpose_reduce_step<8>(s, lane);
  xpose_reduce_step<4>(s, lane);
  xpose_reduce_step<2>(s, lane);
  xpose_reduce_step<1>(s, lane);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
__device__ __forceinline__ void store16(bf16* dst, const uint32_t* pk) {
  *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
}
// 32 bytes (16 bf16) in one 256-bit store (STG.E.256 on sm_100); dst must be 32-byte aligned
__device__ __forceinline__ void store32(bf16* dst, const uint32_t* pk) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(dst), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]),
               "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
               : "memory");
}
// 8 consecutive bf16 channels <-> 8 floats
__device__ __forceinline__ void load8(const bf16* src, float (&f)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(src);
  float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y), c = unpack_bf16x2(u.z), d = unpack_bf16x2(u.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8(bf16* dst, const float (&f)[8]) {
  *reinterpret_cast<uint4*>(dst) = make_uint4(pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]),
                                              pack_bf16x2(f[4], f[5]), pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void ldg8f(const float* src, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(src));
  const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// Test-only probes (exported as crimac_dbg_*): they let the GPU parity tests pin down, on real hardware, the
// UMMA shared-memory descriptor semantics and the TMA swizzle/zero-fill behaviour the production kernels rely on.
#include "host_util.h"
#include "ptx.cuh"

namespace {

// Copies a raw byte image into 1024-aligned shared memory, issues n_mma tcgen05.mma with host-supplied descriptors
// (start-address fields are offsets into the image), and dumps the 128 x N fp32 accumulator.
__global__ void __launch_bounds__(128, 1)
dbg_umma_kernel(const uint8_t* image, int image_bytes, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, int n_mma,
                int a_step16, int b_step16, float* out, int N) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_ptr;
  for (int i = threadIdx.x * 16; i < image_bytes; i += blockDim.x * 16)
    *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(image + i);
  ptx::fence_proxy_async_smem();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&tmem_ptr, 256);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tbase = tmem_ptr;
  if (threadIdx.x == 0) {
    const uint32_t base16 = ptx::smem_u32(smem) >> 4;
    for (int k = 0; k < n_mma; ++k)
      ptx::umma_bf16(tbase, a_desc + base16 + static_cast<uint64_t>(k) * a_step16,
                     b_desc + base16 + static_cast<uint64_t>(k) * b_step16, idesc, k != 0);
    ptx::umma_commit(&bar);
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  for (int chunk = 0; chunk < N / 32; ++chunk) {
    uint32_t v[32];
    ptx::tmem_ld32(tbase + (static_cast<uint32_t>(warp * 32) << 16) + chunk * 32, v);
    ptx::tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * N + chunk * 32 + j] = __uint_as_float(v[j]);
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tbase, 256);
  }
}

// One TMA box {64, 16, box_h, 1} at (c0, x0, y0, n0) -> raw shared-memory bytes.
__global__ void dbg_tma_kernel(const __grid_constant__ CUtensorMap map, int c0, int x0, int y0, int n0, int bytes,
                               uint8_t* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (ptx::smem_u32(smem_raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bar, 1);
    ptx::fence_barrier_init();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    ptx::mbar_arrive_expect_tx(&bar, bytes);
    ptx::tma_load_4d(smem, &map, &bar, c0, x0, y0, n0);
  }
  ptx::mbar_wait(&bar, 0);
  for (int i = threadIdx.x; i < bytes; i += blockDim.x) out[i] = smem[i];
}

}  // namespace

extern "C" int crimac_dbg_umma(const void* image_dev, int image_bytes, uint64_t a_desc, uint64_t b_desc,
                               uint32_t idesc, int n_mma, int a_step_bytes, int b_step_bytes, float* out_dev, int N,
                               void* stream) {
  CRIMAC_REQUIRE(image_bytes % 16 == 0 && image_bytes <= 200 * 1024, "image must be <=200 KiB, multiple of 16");
  CRIMAC_REQUIRE(N % 32 == 0 && N <= 256, "N must be a multiple of 32, <= 256");
  const int smem = image_bytes + 1024;
  CRIMAC_CHECK_CUDA(cudaFuncSetAttribute(dbg_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024));
  dbg_umma_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint8_t*>(image_dev), image_bytes, a_desc, b_desc, idesc, n_mma, a_step_bytes >> 4,
      b_step_bytes >> 4, out_dev, N);
  CRIMAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int crimac_dbg_tma_box(const void* x_dev, int NB, int H, int W, int C, int pitch, int box_h, int sub,
                                  int ky, int kx, int c0, int x0, int y0, int n0, void* out_dev, void* stream) {
  View v{static_cast<bf16*>(const_cast<void*>(x_dev)), NB, H, W, C, pitch};
  CUtensorMap map;
  int rc = make_act_map(&map, v, box_h, sub, ky, kx);
  if (rc) return rc;
  const int bytes = 128 * 16 * box_h;
  dbg_tma_kernel<<<1, 128, bytes + 1024, static_cast<cudaStream_t>(stream)>>>(map, c0, x0, y0, n0, bytes,
                                                                               static_cast<uint8_t*>(out_dev));
  CRIMAC_CHECK_CUDA(cudaGetLastError());
  return 0;
}

#include "host_util.h"
#include <mutex>
#include <unordered_map>
#include <cudaTypedefs.h>
#include <mutex>

static thread_local std::string g_last_error;
void crimac_set_error(const std::string& msg) { g_last_error = msg; }
extern "C" const char* crimac_last_error() { return g_last_error.c_str(); }

// cuTensorMapEncodeTiled is a driver-API entry point; resolve it through the runtime so the library has no
// link-time dependency on libcuda.so (and loads, without computing, on a GPU-less host).
static PFN_cuTensorMapEncodeTiled_v12000 get_encode() {
  static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(sym);
  });
  return fn;
}

int make_act_map(CUtensorMap* out, const View& v, int box_h, int sub, int ky, int kx, int box_w, int box_c, int swizzle) {
  auto enc = get_encode();
  CRIMAC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  CRIMAC_REQUIRE(v.C % 8 == 0 && v.pitch % 8 == 0, "channel count / pitch must be multiples of 8");
  CRIMAC_REQUIRE((reinterpret_cast<uintptr_t>(v.ptr) & 15) == 0, "activation pointer must be 16-byte aligned");
  const cuuint64_t es = 2;
  bf16* base = v.ptr;
  cuuint64_t W = v.W, H = v.H;
  cuuint64_t sx = static_cast<cuuint64_t>(v.pitch) * es;
  cuuint64_t sy = static_cast<cuuint64_t>(v.W) * v.pitch * es;
  const cuuint64_t sn = static_cast<cuuint64_t>(v.H) * v.W * v.pitch * es;
  if (sub) {
    base = v.ptr + (static_cast<size_t>(ky) * v.W + kx) * v.pitch;
    W = v.W / 2;
    H = v.H / 2;
    sx *= 2;
    sy *= 2;
  }
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(v.C), W, H, static_cast<cuuint64_t>(v.N)};
  cuuint64_t strides[3] = {sx, sy, sn};
  cuuint32_t box[4] = {static_cast<cuuint32_t>(box_c), static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(box_h), 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  // L2 promotion: 256 B for dense views (the neighbouring 128 B are the next channel block / pixel, which the same kernel
  // reads next); 128 B when the view is a channel SLICE of a wider buffer (one half of a concat buffer) or every second
  // pixel (ConvTranspose backward): there a 256 B promotion drags the other half of the pixel - data this kernel never
  // uses - through DRAM (measured in round 1: 531 MB read for 268 MB of operand on the ConvTranspose backward-data and
  // weight-gradient launches)
  const bool dense_run = !sub && v.C == v.pitch;
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                   dense_run ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    crimac_set_error("cuTensorMapEncodeTiled(activation) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 2;
  }
  return 0;
}

// Un-swizzled map over the first conv's split input planes xs[part][n][y][x*8 + e] (first_conv_tc.cu): the ping and the
// 8-element chunk are ONE flattened inner dimension, so a box {16 pings x 8, 8 rows} is fetched as 8 rows of 256 B
// instead of 128 rows of 16 B while landing in shared memory in the same order (16 core matrices of 8 pixels x 16 B).
int make_split_input_map(CUtensorMap* out, const bf16* xs, int parts, int NB, int H, int W) {
  auto enc = get_encode();
  CRIMAC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  cuuint64_t dims[4] = {static_cast<cuuint64_t>(W) * 8, static_cast<cuuint64_t>(H), static_cast<cuuint64_t>(NB),
                        static_cast<cuuint64_t>(parts)};
  cuuint64_t strides[3] = {static_cast<cuuint64_t>(W) * 16, static_cast<cuuint64_t>(H) * W * 16,
                           static_cast<cuuint64_t>(NB) * H * W * 16};
  cuuint32_t box[4] = {128, 8, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(xs), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    crimac_set_error("cuTensorMapEncodeTiled(split input) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 2;
  }
  return 0;
}

int make_weight_map(CUtensorMap* out, const bf16* w, int rows, int cols, int box_rows) {
  auto enc = get_encode();
  CRIMAC_REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled unavailable (no CUDA driver?)");
  CRIMAC_REQUIRE(cols % 64 == 0, "packed weight K must be a multiple of 64");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(cols) * 2};
  cuuint32_t box[2] = {64, static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(w), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    crimac_set_error("cuTensorMapEncodeTiled(weights) failed with CUresult " + std::to_string(static_cast<int>(r)));
    return 2;
  }
  return 0;
}

cudaError_t ensure_dynamic_smem_impl(const void* kern, int bytes) {
  static std::mutex mu;
  static std::unordered_map<const void*, uint64_t> done;  // kernel -> bit mask of devices already configured
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  const uint64_t bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0;
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = done.find(kern);
    if (bit && it != done.end() && (it->second & bit)) return cudaSuccess;
  }
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && bit) {
    std::lock_guard<std::mutex> lk(mu);
    done[kern] |= bit;
  }
  return e;
}

int device_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

// ------------------------------------------------------------------------------------------------ profiler
#include <vector>
namespace {
struct ProfRec {
  const char* name;
  double flops, bytes;
  int launches;
  cudaEvent_t e0, e1;
};
bool g_prof_on = false;
std::vector<ProfRec> g_prof;
std::vector<cudaEvent_t> g_event_pool;
unsigned long long g_launches = 0;
cudaEvent_t take_event() {
  if (!g_event_pool.empty()) {
    cudaEvent_t e = g_event_pool.back();
    g_event_pool.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreate(&e);
  return e;
}
}  // namespace

ProfScope::ProfScope(const char* name, double flops, double bytes, cudaStream_t st, int launches) : st_(st), slot_(-1) {
  g_launches += launches;
  if (!g_prof_on) return;
  ProfRec r{name, flops, bytes, launches, take_event(), take_event()};
  cudaEventRecord(r.e0, st);
  g_prof.push_back(r);
  slot_ = static_cast<int>(g_prof.size()) - 1;
}
ProfScope::~ProfScope() {
  if (slot_ >= 0) cudaEventRecord(g_prof[slot_].e1, st_);
}

extern "C" unsigned long long crimac_launch_count() { return g_launches; }
bool crimac_profiling() { return g_prof_on; }

extern "C" int crimac_profile_enable(int on) {
  for (ProfRec& r : g_prof) {
    g_event_pool.push_back(r.e0);
    g_event_pool.push_back(r.e1);
  }
  g_prof.clear();
  g_prof_on = on != 0;
  return 0;
}
// Synchronises the device, then copies up to `cap` records out: names (pointers to static strings), milliseconds,
// algorithmic flops and bytes per record.  Returns the number of records available.
extern "C" int crimac_profile_read(const char** names, float* ms, double* flops, double* bytes, int* launches, int cap) {
  cudaDeviceSynchronize();
  const int n = static_cast<int>(g_prof.size());
  for (int i = 0; i < n && i < cap; ++i) {
    float t = 0.f;
    cudaEventElapsedTime(&t, g_prof[i].e0, g_prof[i].e1);
    if (names) names[i] = g_prof[i].name;
    if (ms) ms[i] = t;
    if (flops) flops[i] = g_prof[i].flops;
    if (bytes) bytes[i] = g_prof[i].bytes;
    if (launches) launches[i] = g_prof[i].launches;
  }
  return n;
}

// C-ABI implementation (see include/bbt_b200.h): errors, twiddle tables,
// launch counting / profiling and the measurement helper.
#include <stdlib.h>

#include <map>
#include <mutex>

#include "common.cuh"
#include "tma.cuh"

namespace bbt {

// ------------------------------------------------- launch count and profiling
thread_local const char* prof_next_name = nullptr;
namespace prof {
std::mutex g_prof_mu;
long long g_launches = 0;
bool g_prof_on = false;
#if !defined(BBT_EMULATE)
struct ProfRec {
  std::string name;
  cudaEvent_t e0, e1;
};
std::vector<ProfRec> g_prof_recs;
thread_local ProfRec* g_prof_open = nullptr;
thread_local cudaStream_t g_prof_stream = nullptr;
#endif
}  // namespace prof
using namespace prof;
void prof_count() {
  std::lock_guard<std::mutex> lock(g_prof_mu);
  ++g_launches;
}
#if !defined(BBT_EMULATE)
void prof_begin(const char* name, cudaStream_t stream) {
  if (!g_prof_on) return;
  ProfRec* r = new ProfRec();
  r->name = name;
  cudaEventCreate(&r->e0);
  cudaEventCreate(&r->e1);
  cudaEventRecord(r->e0, stream);
  g_prof_open = r;
  g_prof_stream = stream;
}
void prof_end() {
  if (!g_prof_open) return;
  cudaEventRecord(g_prof_open->e1, g_prof_stream);
  {
    std::lock_guard<std::mutex> lock(g_prof_mu);
    g_prof_recs.push_back(*g_prof_open);
  }
  delete g_prof_open;
  g_prof_open = nullptr;
}
#endif

// ------------------------------------------------------------------- errors
namespace {
thread_local std::string g_err;
}
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
const std::string& last_error() { return g_err; }
int check_launch(const char* what) {
  const char* e = launch_error();
  if (e) return fail(BBT_ECUDA, std::string(what) + ": " + e);
  return BBT_OK;
}

// ----------------------------------------------------------- twiddle tables
static int current_device() {
#if defined(BBT_EMULATE)
  return 0;
#else
  int d = 0;
  cudaGetDevice(&d);
  return d;
#endif
}

cf* make_roots(int64_t count, double denom) {
  std::vector<cf> host(count);
  for (int64_t m = 0; m < count; ++m) {
    const double ang = -2.0 * M_PI * (double)m / denom;
    host[m] = mk((float)cos(ang), (float)sin(ang));
  }
  void* dev = nullptr;
  if (dev_alloc(&dev, count * sizeof(cf))) return nullptr;
  bool ok = !h2d(dev, host.data(), count * sizeof(cf), 0);
#if !defined(BBT_EMULATE)
  // The host vector goes out of scope: wait for the copy.
  if (ok && cudaStreamSynchronize(0) != cudaSuccess) ok = false;
#endif
  if (!ok) {
    dev_free(dev);
    return nullptr;
  }
  return static_cast<cf*>(dev);
}

const cf* twiddle_table(int log2n) {
  static std::mutex mu;
  static std::map<std::pair<int, int>, cf*> tables;
  std::lock_guard<std::mutex> lock(mu);
  const std::pair<int, int> key(current_device(), log2n);
  auto it = tables.find(key);
  if (it != tables.end()) return it->second;
  const int64_t n = int64_t(1) << log2n;
  cf* t = make_roots(n, (double)n);
  if (t) tables[key] = t;  // a failed allocation is retried next time
  return t;
}

// ------------------------------------------------------------ tuning knobs
// Kernel variants are chosen by the library; for A/B measurements a variant
// can be forced with bbt_tune_set("key", value) or the environment variable
// BBT_TUNE="key=value,key=value".
namespace {
std::mutex g_tune_mu;
std::map<std::string, int> g_tune;
bool g_tune_env_read = false;
}  // namespace
int tune(const char* key, int dflt) {
  std::lock_guard<std::mutex> lock(g_tune_mu);
  if (!g_tune_env_read) {
    g_tune_env_read = true;
    const char* env = getenv("BBT_TUNE");
    if (env) {
      std::string text(env);
      size_t pos = 0;
      while (pos < text.size()) {
        size_t end = text.find(',', pos);
        if (end == std::string::npos) end = text.size();
        const std::string item = text.substr(pos, end - pos);
        const size_t eq = item.find('=');
        if (eq != std::string::npos && !g_tune.count(item.substr(0, eq)))
          g_tune[item.substr(0, eq)] = atoi(item.c_str() + eq + 1);
        pos = end + 1;
      }
    }
  }
  auto it = g_tune.find(key);
  return it == g_tune.end() ? dflt : it->second;
}
void tune_set(const char* key, int value) {
  std::lock_guard<std::mutex> lock(g_tune_mu);
  g_tune[key] = value;
}

// ------------------------------------------------------------- tensor maps
#if defined(BBT_EMULATE)
int make_tensor_map(TensorMap* map, void* base, int elem_bytes, int rank,
                    const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  if (rank < 1 || rank > 3) return -1;
  map->base = static_cast<char*>(base);
  map->rank = rank;
  for (int i = 0; i < 3; ++i) {
    map->dim[i] = i < rank ? dims[i] : 1;
    map->stride[i] = i == 0 ? (uint64_t)elem_bytes
                            : (i < rank ? strides_bytes[i] : 0);
    map->box[i] = i < rank ? box[i] : 1;
  }
  return 0;
}
#else
int make_tensor_map(TensorMap* map, void* base, int elem_bytes, int rank,
                    const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box) {
  // The driver entry point is looked up through the runtime, so the library
  // does not link against libcuda (and loads on machines without a driver).
  typedef CUresult (*Encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t,
                             void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*,
                             CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static Encode encode = nullptr;
  static std::mutex mu;
  {
    std::lock_guard<std::mutex> lock(mu);
    if (!encode) {
      void* fn = nullptr;
      cudaDriverEntryPointQueryResult q;
      if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn,
                                  cudaEnableDefault, &q) != cudaSuccess ||
          !fn)
        return -1;
      encode = reinterpret_cast<Encode>(fn);
    }
  }
  if (rank < 1 || rank > 3) return -1;
  cuuint64_t d[3], s[2];
  cuuint32_t b[3], es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) d[i] = dims[i], b[i] = box[i];
  for (int i = 1; i < rank; ++i) s[i - 1] = strides_bytes[i];
  const CUtensorMapDataType dt =
      elem_bytes == 8   ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64
      : elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                        : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  const CUresult r =
      encode(map, dt, rank, base, d, s, b, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}
#endif

#if !defined(BBT_EMULATE)
int sm_count() {
  static std::mutex mu;
  static std::map<int, int> counts;
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(mu);
  auto it = counts.find(dev);
  if (it != counts.end()) return it->second;
  int n = 148;
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  counts[dev] = n;
  return n;
}

// cudaFuncSetAttribute once per (device, kernel) and size increase instead of
// on every launch.
int set_max_smem(const void* kernel, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> done;
  const std::pair<int, const void*> key(current_device(), kernel);
  std::lock_guard<std::mutex> lock(mu);
  auto it = done.find(key);
  if (it != done.end() && it->second >= bytes) return 0;
  const int rc = (int)cudaFuncSetAttribute(
      kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (!rc) done[key] = bytes;
  return rc;
}
#endif

namespace {
BBT_GLOBAL void strided_copy_kernel(const float4* in, float4* out,
                                    long long rows, long long row_stride16,
                                    long long chunk16, long long n_tiles) {
  // One CTA per column tile; threads sweep (row, 16-byte word) pairs.
  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const long long total = rows * chunk16;
    for (long long i = threadIdx.x; i < total; i += blockDim.x) {
      const long long r = i / chunk16, c = i % chunk16;
      const long long idx = r * row_stride16 + tile * chunk16 + c;
      out[idx] = in[idx];
    }
  }
}
}  // namespace

}  // namespace bbt

using namespace bbt;

extern "C" {

int bbt_version(void) { return 200; }

int bbt_tune_set(const char* key, int value) {
  if (!key) return fail(BBT_EINVAL, "null key");
  tune_set(key, value);
  return BBT_OK;
}

const char* bbt_last_error(void) { return last_error().c_str(); }

int64_t bbt_launch_count(void) {
  std::lock_guard<std::mutex> lock(bbt::prof::g_prof_mu);
  return bbt::prof::g_launches;
}

int bbt_profile_enable(int on) {
  std::lock_guard<std::mutex> lock(bbt::prof::g_prof_mu);
  bbt::prof::g_prof_on = on != 0;
  return BBT_OK;
}

int bbt_profile_report(char* buf, int64_t size) {
  if (!buf || size < 1) return fail(BBT_EINVAL, "null buffer");
  std::string text;
#if !defined(BBT_EMULATE)
  std::map<std::string, std::pair<long long, double>> acc;
  std::vector<bbt::prof::ProfRec> recs;
  {
    std::lock_guard<std::mutex> lock(bbt::prof::g_prof_mu);
    recs.swap(bbt::prof::g_prof_recs);
  }
  for (auto& r : recs) {
    float ms = 0.f;
    if (cudaEventSynchronize(r.e1) == cudaSuccess &&
        cudaEventElapsedTime(&ms, r.e0, r.e1) == cudaSuccess) {
      auto& a = acc[r.name];
      a.first += 1;
      a.second += ms;
    }
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  for (auto& kv : acc) {
    char line[256];
    snprintf(line, sizeof line, "%s %lld %.6f\n", kv.first.c_str(),
             kv.second.first, kv.second.second);
    text += line;
  }
#endif
  if ((int64_t)text.size() + 1 > size) return fail(BBT_EINVAL, "buffer too small");
  memcpy(buf, text.c_str(), text.size() + 1);
  return BBT_OK;
}

int bbt_strided_copy_bench(const void* in, void* out, int64_t rows,
                           int64_t row_stride_bytes, int64_t chunk_bytes,
                           int64_t n_tiles, void* stream) {
  if (!in || !out) return fail(BBT_EINVAL, "null argument");
  if (chunk_bytes % 16 || row_stride_bytes % 16)
    return fail(BBT_EINVAL, "sizes must be multiples of 16 bytes");
  const unsigned blocks = (unsigned)std::min<int64_t>(n_tiles, (int64_t)sm_count() * 8);
  BBT_LAUNCH(strided_copy_kernel, dim3(blocks), dim3(256), 0, as_stream(stream),
             static_cast<const float4*>(in), static_cast<float4*>(out),
             (long long)rows, (long long)(row_stride_bytes / 16),
             (long long)(chunk_bytes / 16), (long long)n_tiles);
  return check_launch("strided copy kernel");
}

}  // extern "C"

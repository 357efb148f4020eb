// C-ABI implementation (see include/bbt_b200.h): errors, twiddle tables,
// launch counting / profiling and the measurement helper.
#include <map>
#include <mutex>

#include "common.cuh"

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

int bbt_version(void) { return 101; }

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

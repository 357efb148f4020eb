// C-ABI implementation (see include/bbt_b200.h) and plan objects.
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/bbt_b200.h"
#include "kernels_dedisperse.cuh"
#include "kernels_detect.cuh"
#include "kernels_fft.cuh"

using namespace bbt;

// ------------------------------------------------- launch count and profiling
namespace bbt {
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
  // Template arguments are not part of the stringified name; keep it short.
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
}  // namespace bbt

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

int check_launch(const char* what) {
  const char* e = launch_error();
  if (e) return fail(BBT_ECUDA, std::string(what) + ": " + e);
  return BBT_OK;
}

int ilog2(int64_t n) {
  int l = 0;
  while ((int64_t(1) << l) < n) ++l;
  return l;
}
bool is_pow2(int64_t n) { return n > 0 && (n & (n - 1)) == 0; }

int current_device() {
#if defined(BBT_EMULATE)
  return 0;
#else
  int d = 0;
  cudaGetDevice(&d);
  return d;
#endif
}

// exp(-2 pi i m / count * step) table on the device, computed in float64.
cf* make_roots(int64_t count, double denom) {
  std::vector<cf> host(count);
  for (int64_t m = 0; m < count; ++m) {
    const double ang = -2.0 * M_PI * (double)m / denom;
    host[m] = mk((float)cos(ang), (float)sin(ang));
  }
  void* dev = nullptr;
  if (dev_alloc(&dev, count * sizeof(cf))) return nullptr;
  if (h2d(dev, host.data(), count * sizeof(cf), 0)) return nullptr;
  return static_cast<cf*>(dev);
}

// Per-device table of the 8192nd roots of unity shared by all block FFTs.
const cf* twiddle_table() {
  static std::mutex mu;
  static std::map<int, cf*> tables;
  std::lock_guard<std::mutex> lock(mu);
  const int dev = current_device();
  auto it = tables.find(dev);
  if (it != tables.end()) return it->second;
  cf* t = make_roots(kTwiddleTable, (double)kTwiddleTable);
  tables[dev] = t;
  return t;
}

bbt_stream_t as_stream(void* s) { return static_cast<bbt_stream_t>(s); }

#define BBT_FOR_LOG2(L, F)                                                    \
  switch (L) {                                                                \
    case 1: F(1); break;                                                      \
    case 2: F(2); break;                                                      \
    case 3: F(3); break;                                                      \
    case 4: F(4); break;                                                      \
    case 5: F(5); break;                                                      \
    case 6: F(6); break;                                                      \
    case 7: F(7); break;                                                      \
    case 8: F(8); break;                                                      \
    case 9: F(9); break;                                                      \
    case 10: F(10); break;                                                    \
    case 11: F(11); break;                                                    \
    case 12: F(12); break;                                                    \
    case 13: F(13); break;                                                    \
    default: break;                                                           \
  }

int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// ------------------------------------------------------------------ FFT plan
template <int L, bool LANEFAST>
int launch_fft(int kind, const FftArgs& a, bbt_stream_t st) {
  using C = FftCfg<L>;
  const int64_t lanes = LANEFAST ? a.outer * a.inner : a.outer;
  const int64_t blocks = ceil_div(lanes, C::G);
  if (blocks <= 0) return BBT_OK;
  if (blocks > 2147483647LL) return fail(BBT_EUNSUPPORTED, "grid too large");
  const size_t smem = C::SMEM_BYTES;
  auto kern = kind == BBT_C2C   ? fft_c2c_kernel<L, LANEFAST>
              : kind == BBT_R2C ? fft_r2c_kernel<L, LANEFAST>
                                : fft_c2r_kernel<L, LANEFAST>;
  if (BBT_SET_SMEM(kern, smem))
    return fail(BBT_ECUDA, "cannot set shared memory size");
  prof_next_name = kind == BBT_C2C ? "fft_c2c" : kind == BBT_R2C ? "fft_r2c"
                                                                 : "fft_c2r";
  BBT_LAUNCH(kern, dim3((unsigned)blocks), dim3(C::THREADS), smem, st, a);
  return check_launch("fft kernel");
}

int run_fft(int log2n, int kind, const FftArgs& a, bbt_stream_t st) {
  int rc = BBT_EUNSUPPORTED;
  const bool lanefast = a.inner > 1;
#define F(L)                                                   \
  rc = lanefast ? launch_fft<L, true>(kind, a, st)             \
                : launch_fft<L, false>(kind, a, st)
  BBT_FOR_LOG2(log2n, F)
#undef F
  if (rc == BBT_EUNSUPPORTED && g_err.empty())
    fail(rc, "unsupported FFT length");
  return rc;
}

}  // namespace

struct bbt_fft_plan {
  int64_t n, outer, inner;
  int kind, direction;
  float scale;
  int log2n, log2n1, log2n2;  // n = n1*n2 when n > 8192
  const cf* tw;
  cf* big_lo;
  cf* big_hi;
};

struct bbt_dedisperse_plan {
  int64_t n, n_series, pad_start, n_valid, n_chirp;
  int log2n, log2n1, log2n2;
  const cf* tw;
  cf* big_lo;
  cf* big_hi;
  cf* chirp;        // [n_chirp][n1][n2]
  int* series_map;  // device
};

extern "C" {

int bbt_version(void) { return 100; }

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

const char* bbt_last_error(void) { return g_err.c_str(); }

int bbt_fft_plan_create(bbt_fft_plan** plan, int64_t n, int64_t outer,
                        int64_t inner, int kind, int direction, double scale) {
  if (!plan) return fail(BBT_EINVAL, "null plan pointer");
  *plan = nullptr;
  if (n < 1 || outer < 0 || inner < 1) return fail(BBT_EINVAL, "bad FFT shape");
  if (!is_pow2(n) || n < 2)
    return fail(BBT_EUNSUPPORTED,
                "FFT length must be a power of two >= 2 (use "
                "CudaFFTMaker.next_fast_len)");
  if (kind < BBT_C2C || kind > BBT_C2R) return fail(BBT_EINVAL, "bad FFT kind");
  const int l = ilog2(n);
  if (l > kLog2TwiddleTable) {
    if (kind != BBT_C2C || inner != 1)
      return fail(BBT_EUNSUPPORTED,
                  "FFT lengths above 8192 need complex data on a contiguous "
                  "axis (inner == 1)");
    if (l > 2 * kLog2TwiddleTable)
      return fail(BBT_EUNSUPPORTED, "FFT length above 2^26");
  }
  bbt_fft_plan* p = new bbt_fft_plan();
  p->n = n;
  p->outer = outer;
  p->inner = inner;
  p->kind = kind;
  p->direction = direction == BBT_BACKWARD ? BBT_BACKWARD : BBT_FORWARD;
  p->scale = (float)scale;
  p->log2n = l;
  p->log2n1 = p->log2n2 = 0;
  p->big_lo = p->big_hi = nullptr;
  p->tw = twiddle_table();
  if (!p->tw) {
    delete p;
    return fail(BBT_ENOMEM, "cannot allocate twiddle table");
  }
  if (l > kLog2TwiddleTable) {
    p->log2n2 = (l + 1) / 2;
    p->log2n1 = l - p->log2n2;
    p->big_lo = make_roots(kTwiddleTable, (double)n);
    p->big_hi = make_roots(n >> kLog2TwiddleTable,
                           (double)n / (double)kTwiddleTable);
    if (!p->big_lo || !p->big_hi) {
      bbt_fft_plan_destroy(p);
      return fail(BBT_ENOMEM, "cannot allocate twiddle tables");
    }
  }
  *plan = p;
  return BBT_OK;
}

int64_t bbt_fft_plan_work_bytes(const bbt_fft_plan* p) {
  if (!p || p->log2n <= kLog2TwiddleTable) return 0;
  return p->outer * p->n * (int64_t)sizeof(cf);
}

int bbt_fft_exec(const bbt_fft_plan* p, const void* in, void* out, void* work,
                 void* stream) {
  if (!p || !in || !out) return fail(BBT_EINVAL, "null argument");
  bbt_stream_t st = as_stream(stream);
  if (p->outer == 0) return BBT_OK;
  const int inverse = p->direction == BBT_BACKWARD;
  if (p->log2n <= kLog2TwiddleTable) {
    FftArgs a{in, out, p->tw, p->outer, p->inner, inverse, p->scale};
    return run_fft(p->log2n, p->kind, a, st);
  }
  // Four-step transform of a contiguous axis, X[k1 + n1 k2]:
  //   columns (n1) -> twiddle -> rows (n2) -> transpose to natural order.
  if (!work) return fail(BBT_EINVAL, "large FFT needs a work buffer");
  const int64_t n1 = int64_t(1) << p->log2n1, n2 = int64_t(1) << p->log2n2;
  BigTwiddle big{p->big_lo, p->big_hi};
  const int64_t total = p->outer * p->n;
  const unsigned tw_blocks =
      (unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 32);
  cf* w = static_cast<cf*>(work);
  int rc;
  // Forward: x[n1][n2] (n = n1 idx * n2 + n2 idx): FFT over the n1 index is
  // a strided transform with inner = n2.
  FftArgs col{in, w, p->tw, p->outer, n2, inverse, 1.f};
  if ((rc = run_fft(p->log2n1, BBT_C2C, col, st))) return rc;
  BBT_LAUNCH(twiddle_kernel, dim3(tw_blocks), dim3(256), 0, st, w, n1, n2,
             p->outer, big, inverse);
  if ((rc = check_launch("twiddle kernel"))) return rc;
  FftArgs row{w, w, p->tw, p->outer * n1, 1, inverse, p->scale};
  if ((rc = run_fft(p->log2n2, BBT_C2C, row, st))) return rc;
  // w[k1][k2] -> out[k2][k1]  (bin k = k1 + n1*k2)
  dim3 grid((unsigned)ceil_div(n2, 32), (unsigned)ceil_div(n1, 32),
            (unsigned)p->outer);
  BBT_LAUNCH(transpose_kernel, grid, dim3(32, 8), 32 * 33 * sizeof(cf), st, w,
             static_cast<cf*>(out), n1, n2);
  return check_launch("transpose kernel");
}

int bbt_fft_plan_destroy(bbt_fft_plan* p) {
  if (!p) return BBT_OK;
  if (p->big_lo) dev_free(p->big_lo);
  if (p->big_hi) dev_free(p->big_hi);
  delete p;
  return BBT_OK;
}

// ------------------------------------------------------------- dedispersion
int bbt_dedisperse_plan_create(bbt_dedisperse_plan** plan, int64_t n,
                               int64_t n_series, int64_t pad_start,
                               int64_t n_valid, int64_t n_chirp,
                               const int32_t* series_map,
                               const double* freq_mhz, const double* fref_mhz,
                               const int8_t* sideband, double dm,
                               double rate_mhz, double sample_offset,
                               int log2n1_hint) {
  if (!plan) return fail(BBT_EINVAL, "null plan pointer");
  *plan = nullptr;
  if (n < 2 || !is_pow2(n))
    return fail(BBT_EUNSUPPORTED, "frame length must be a power of two >= 2");
  const int l = ilog2(n);
  if (l > 2 * kLog2TwiddleTable)
    return fail(BBT_EUNSUPPORTED, "frame length above 2^26");
  if (n_series < 1 || n_chirp < 1 || pad_start < 0 || n_valid < 1 ||
      pad_start + n_valid > n)
    return fail(BBT_EINVAL, "bad dedispersion geometry");
  if (!series_map) return fail(BBT_EINVAL, "null series_map");
  for (int64_t s = 0; s < n_series; ++s)
    if (series_map[s] < 0 || series_map[s] >= n_chirp)
      return fail(BBT_EINVAL, "series_map entry out of range");
  bbt_dedisperse_plan* p = new bbt_dedisperse_plan();
  p->n = n;
  p->n_series = n_series;
  p->pad_start = pad_start;
  p->n_valid = n_valid;
  p->n_chirp = n_chirp;
  p->log2n = l;
  p->big_lo = p->big_hi = p->chirp = nullptr;
  p->series_map = nullptr;
  if (l <= kLog2TwiddleTable) {
    p->log2n1 = 0;
    p->log2n2 = l;
  } else {
    int l1 = log2n1_hint > 0 ? log2n1_hint : l - 12;
    if (l - l1 > kLog2TwiddleTable) l1 = l - kLog2TwiddleTable;
    if (l1 > kLog2TwiddleTable) l1 = kLog2TwiddleTable;
    if (l1 < 1) l1 = 1;
    p->log2n1 = l1;
    p->log2n2 = l - l1;
  }
  p->tw = twiddle_table();
  void* d = nullptr;
  int rc = BBT_OK;
  if (!p->tw) rc = BBT_ENOMEM;
  if (!rc && dev_alloc(&d, n_chirp * n * sizeof(cf))) rc = BBT_ENOMEM;
  p->chirp = static_cast<cf*>(d);
  d = nullptr;
  if (!rc && dev_alloc(&d, n_series * sizeof(int))) rc = BBT_ENOMEM;
  p->series_map = static_cast<int*>(d);
  if (!rc && h2d(p->series_map, series_map, n_series * sizeof(int), 0))
    rc = BBT_ECUDA;
  if (!rc && l > kLog2TwiddleTable) {
    p->big_lo = make_roots(kTwiddleTable, (double)n);
    p->big_hi = make_roots(n >> kLog2TwiddleTable,
                           (double)n / (double)kTwiddleTable);
    if (!p->big_lo || !p->big_hi) rc = BBT_ENOMEM;
  }
  if (rc) {
    bbt_dedisperse_plan_destroy(p);
    return fail(rc, "cannot allocate dedispersion plan tables");
  }
  if (freq_mhz && fref_mhz && sideband) {
    // Chirp parameters to the device, then one float64 kernel.
    void *dfreq = nullptr, *dref = nullptr, *dsb = nullptr;
    if (dev_alloc(&dfreq, n_chirp * sizeof(double)) ||
        dev_alloc(&dref, n_chirp * sizeof(double)) ||
        dev_alloc(&dsb, n_chirp) ||
        h2d(dfreq, freq_mhz, n_chirp * sizeof(double), 0) ||
        h2d(dref, fref_mhz, n_chirp * sizeof(double), 0) ||
        h2d(dsb, sideband, n_chirp, 0)) {
      bbt_dedisperse_plan_destroy(p);
      return fail(BBT_ENOMEM, "cannot stage chirp parameters");
    }
    ChirpArgs ca;
    ca.chirp = p->chirp;
    ca.freq_mhz = static_cast<const double*>(dfreq);
    ca.fref_mhz = static_cast<const double*>(dref);
    ca.sideband = static_cast<const signed char*>(dsb);
    ca.N = n;
    ca.n1 = int64_t(1) << p->log2n1;
    ca.n_chirp = n_chirp;
    ca.d = dm / 2.41e-4;  // dm.py:37
    ca.rate_mhz = rate_mhz;
    ca.sample_offset = sample_offset;
    const unsigned blocks =
        (unsigned)std::min<int64_t>(ceil_div(n_chirp * n, 256), 148 * 32);
    BBT_LAUNCH(chirp_kernel, dim3(blocks), dim3(256), 0, (bbt_stream_t)0, ca);
    rc = check_launch("chirp kernel");
#if !defined(BBT_EMULATE)
    if (!rc && cudaStreamSynchronize(0) != cudaSuccess)
      rc = fail(BBT_ECUDA, "chirp kernel failed");
#endif
    dev_free(dfreq);
    dev_free(dref);
    dev_free(dsb);
    if (rc) {
      bbt_dedisperse_plan_destroy(p);
      return rc;
    }
  }
  *plan = p;
  return BBT_OK;
}

int bbt_dedisperse_plan_set_response(bbt_dedisperse_plan* p,
                                     const void* host_response) {
  if (!p || !host_response) return fail(BBT_EINVAL, "null argument");
  const int64_t total = p->n_chirp * p->n;
  void* tmp = nullptr;
  if (dev_alloc(&tmp, total * sizeof(cf)))
    return fail(BBT_ENOMEM, "cannot stage response");
  int rc = BBT_OK;
  if (h2d(tmp, host_response, total * sizeof(cf), 0)) rc = BBT_ECUDA;
  if (!rc) {
    const unsigned blocks =
        (unsigned)std::min<int64_t>(ceil_div(total, 256), 148 * 32);
    BBT_LAUNCH(chirp_scatter_kernel, dim3(blocks), dim3(256), 0,
               (bbt_stream_t)0, p->chirp, static_cast<const cf*>(tmp), p->n,
               int64_t(1) << p->log2n1, p->n_chirp);
    rc = check_launch("response scatter kernel");
#if !defined(BBT_EMULATE)
    if (!rc && cudaStreamSynchronize(0) != cudaSuccess)
      rc = fail(BBT_ECUDA, "response scatter failed");
#endif
  }
  dev_free(tmp);
  return rc;
}

int bbt_dedisperse_plan_get_response(const bbt_dedisperse_plan* p,
                                     void* host_response) {
  if (!p || !host_response) return fail(BBT_EINVAL, "null argument");
  const int64_t n1 = int64_t(1) << p->log2n1, n2 = p->n / n1;
  std::vector<cf> tmp(p->n_chirp * p->n);
#if defined(BBT_EMULATE)
  memcpy(tmp.data(), p->chirp, tmp.size() * sizeof(cf));
#else
  if (cudaMemcpy(tmp.data(), p->chirp, tmp.size() * sizeof(cf),
                 cudaMemcpyDeviceToHost) != cudaSuccess)
    return fail(BBT_ECUDA, "cannot copy chirp to host");
#endif
  cf* out = static_cast<cf*>(host_response);
  for (int64_t c = 0; c < p->n_chirp; ++c)
    for (int64_t k1 = 0; k1 < n1; ++k1)
      for (int64_t k2 = 0; k2 < n2; ++k2)
        out[c * p->n + k1 + n1 * k2] = tmp[c * p->n + k1 * n2 + k2];
  return BBT_OK;
}

int64_t bbt_dedisperse_work_bytes(const bbt_dedisperse_plan* p,
                                  int64_t n_frames) {
  if (!p || p->log2n1 == 0) return 0;
  return n_frames * p->n * p->n_series * (int64_t)sizeof(cf);
}

}  // extern "C"
namespace {
template <int L1>
int launch_dd_col(bool inverse, const DdArgs& a, int64_t n_frames,
                  bbt_stream_t st) {
  using C = FftCfg<L1>;
  const int64_t cols = (a.N >> L1) * a.S;
  dim3 grid((unsigned)ceil_div(cols, C::G), (unsigned)n_frames);
  const size_t smem = C::SMEM_BYTES;
  auto kern = inverse ? dd_col_inv_kernel<L1> : dd_col_fwd_kernel<L1>;
  if (BBT_SET_SMEM(kern, smem))
    return fail(BBT_ECUDA, "cannot set shared memory size");
  prof_next_name = inverse ? "dd_col_inv" : "dd_col_fwd";
  BBT_LAUNCH(kern, grid, dim3(C::THREADS), smem, st, a);
  return check_launch("dedispersion column kernel");
}

template <int L2>
int launch_dd_row(const DdArgs& a, int64_t n_frames, bbt_stream_t st) {
  using C = FftCfg<L2>;
  const int64_t n1 = a.N >> L2;
  dim3 grid((unsigned)ceil_div(n1, C::G), (unsigned)a.S, (unsigned)n_frames);
  const size_t smem = C::SMEM_BYTES;
  auto kern = dd_row_kernel<L2>;
  if (BBT_SET_SMEM(kern, smem))
    return fail(BBT_ECUDA, "cannot set shared memory size");
  prof_next_name = "dd_row";
  BBT_LAUNCH(kern, grid, dim3(C::THREADS), smem, st, a);
  return check_launch("dedispersion row kernel");
}

template <int L, bool LANEFAST>
int launch_dd_small(const DdArgs& a, int64_t n_frames, bbt_stream_t st) {
  using C = FftCfg<L>;
  const int64_t blocks = ceil_div(n_frames * a.S, C::G);
  const size_t smem = C::SMEM_BYTES;
  auto kern = dd_small_kernel<L, LANEFAST>;
  if (BBT_SET_SMEM(kern, smem))
    return fail(BBT_ECUDA, "cannot set shared memory size");
  prof_next_name = "dd_small";
  BBT_LAUNCH(kern, dim3((unsigned)blocks), dim3(C::THREADS), smem, st, a,
             (long long)n_frames);
  return check_launch("dedispersion kernel");
}
}  // namespace
extern "C" {

int bbt_dedisperse_exec(const bbt_dedisperse_plan* p, const void* in,
                        int64_t in_frame_stride, int64_t n_frames,
                        int64_t skip, void* out, int64_t out_frame_stride,
                        void* work, void* stream) {
  if (!p || !in || !out) return fail(BBT_EINVAL, "null argument");
  if (n_frames <= 0) return BBT_OK;
  if (n_frames > 65535) return fail(BBT_EUNSUPPORTED, "too many frames per call");
  if (skip < 0 || skip >= p->n_valid) return fail(BBT_EINVAL, "bad skip");
  bbt_stream_t st = as_stream(stream);
  DdArgs a;
  a.in = static_cast<const cf*>(in);
  a.out = static_cast<cf*>(out);
  a.work = static_cast<cf*>(work);
  a.tw = p->tw;
  a.big = BigTwiddle{p->big_lo, p->big_hi};
  a.chirp = p->chirp;
  a.series_map = p->series_map;
  a.in_frame_stride = in_frame_stride;
  a.out_frame_stride = out_frame_stride;
  a.N = p->n;
  a.S = p->n_series;
  a.log2n1 = p->log2n1;
  a.log2n2 = p->log2n2;
  a.lo = (p->pad_start + skip) * p->n_series;
  a.hi = (p->pad_start + p->n_valid) * p->n_series;
  // Valid samples are stored from out + f*stride on, i.e. sample
  // pad_start + skip lands at offset 0.
  a.out_shift = (p->pad_start + skip) * p->n_series;
  a.scale = (float)(1.0 / (double)p->n);
  int rc = BBT_EUNSUPPORTED;
  if (p->log2n1 == 0) {
    const bool lanefast = p->n_series > 1;
#define F(L)                                                        \
  rc = lanefast ? launch_dd_small<L, true>(a, n_frames, st)         \
                : launch_dd_small<L, false>(a, n_frames, st)
    BBT_FOR_LOG2(p->log2n2, F)
#undef F
    return rc;
  }
  if (!work) return fail(BBT_EINVAL, "dedispersion needs a work buffer");
#define F(L) rc = launch_dd_col<L>(false, a, n_frames, st)
  BBT_FOR_LOG2(p->log2n1, F)
#undef F
  if (rc) return rc;
  rc = BBT_EUNSUPPORTED;
#define F(L) rc = launch_dd_row<L>(a, n_frames, st)
  BBT_FOR_LOG2(p->log2n2, F)
#undef F
  if (rc) return rc;
  rc = BBT_EUNSUPPORTED;
#define F(L) rc = launch_dd_col<L>(true, a, n_frames, st)
  BBT_FOR_LOG2(p->log2n1, F)
#undef F
  return rc;
}

int bbt_dedisperse_plan_destroy(bbt_dedisperse_plan* p) {
  if (!p) return BBT_OK;
  if (p->big_lo) dev_free(p->big_lo);
  if (p->big_hi) dev_free(p->big_hi);
  if (p->chirp) dev_free(p->chirp);
  if (p->series_map) dev_free(p->series_map);
  delete p;
  return BBT_OK;
}

// ----------------------------------------------------------------- detection
static unsigned grid_for(int64_t total, int threads) {
  return (unsigned)std::max<int64_t>(
      1, std::min<int64_t>(ceil_div(total, threads), (int64_t)sm_count() * 16));
}

int bbt_power_exec(const void* in, void* out, int64_t a, int64_t b,
                   void* stream) {
  if (!in || !out) return fail(BBT_EINVAL, "null argument");
  if (a <= 0 || b <= 0) return BBT_OK;
  BBT_LAUNCH(power_kernel, dim3(grid_for(a * b, 256)), dim3(256), 0,
             as_stream(stream), static_cast<const cf*>(in),
             static_cast<float*>(out), (long long)a, (long long)b);
  return check_launch("power kernel");
}

int bbt_square_exec(const void* in, void* out, int64_t n, int is_complex,
                    void* stream) {
  if (!in || !out) return fail(BBT_EINVAL, "null argument");
  if (n <= 0) return BBT_OK;
  BBT_LAUNCH(square_kernel, dim3(grid_for(n, 256)), dim3(256), 0,
             as_stream(stream), static_cast<const float*>(in),
             static_cast<float*>(out), (long long)n, is_complex);
  return check_launch("square kernel");
}

}  // extern "C"
namespace {
template <int L, bool LANEFAST, bool INTEGRATE>
int launch_chanpow(const ChanPowArgs& a, int64_t n_bins, bbt_stream_t st) {
  using C = FftCfg<L>;
  const int64_t blocks = ceil_div(a.msub * a.M, C::G);
  dim3 grid((unsigned)blocks, (unsigned)(INTEGRATE ? n_bins : 1));
  const size_t smem = 2 * C::SMEM_BYTES;
  auto kern = chanpow_kernel<L, LANEFAST, INTEGRATE>;
  if (BBT_SET_SMEM(kern, smem))
    return fail(BBT_ECUDA, "cannot set shared memory size");
  prof_next_name = INTEGRATE ? "chanpow_integrate" : "chanpow";
  BBT_LAUNCH(kern, grid, dim3(C::THREADS), smem, st, a);
  return check_launch("channelize-power kernel");
}

template <bool INTEGRATE>
int run_chanpow(int log2n, ChanPowArgs& a, int64_t n_bins, int64_t max_width,
                bbt_stream_t st) {
  int rc = BBT_EUNSUPPORTED;
  const bool lanefast = a.M > 1;
#define F(L)                                                                 \
  {                                                                          \
    const int64_t g = FftCfg<L>::G;                                          \
    const int64_t want = (int64_t)sm_count() * 8 * g;                        \
    int64_t msub = ceil_div(want, a.M * (INTEGRATE ? n_bins : 1));           \
    if (msub > max_width) msub = max_width;                                  \
    if (msub < 1) msub = 1;                                                  \
    a.msub = msub;                                                           \
    rc = lanefast ? launch_chanpow<L, true, INTEGRATE>(a, n_bins, st)        \
                  : launch_chanpow<L, false, INTEGRATE>(a, n_bins, st);      \
  }
  BBT_FOR_LOG2(log2n, F)
#undef F
  if (rc == BBT_EUNSUPPORTED)
    fail(rc, "channelizer length must be a power of two in [2, 8192]");
  return rc;
}
}  // namespace
extern "C" {

int bbt_channelize_power_exec(const void* in, void* out, int64_t n, int64_t m,
                              int64_t n_spec, void* stream) {
  if (!in || !out) return fail(BBT_EINVAL, "null argument");
  if (!is_pow2(n) || n < 2 || m < 1) return fail(BBT_EUNSUPPORTED, "bad channelizer shape");
  if (n_spec <= 0) return BBT_OK;
  ChanPowArgs a{};
  a.in = static_cast<const cf2*>(in);
  a.out = static_cast<float*>(out);
  a.tw = twiddle_table();
  a.M = m;
  a.n_spec = n_spec;
  return run_chanpow<false>(ilog2(n), a, 1, n_spec, as_stream(stream));
}

int bbt_channelize_power_integrate_exec(const void* in, int64_t n, int64_t m,
                                        int64_t n_spec, int64_t j_first,
                                        const int64_t* offsets,
                                        int64_t b_first, int64_t n_bins,
                                        void* sum, void* count, void* stream) {
  if (!in || !sum || !count || !offsets) return fail(BBT_EINVAL, "null argument");
  if (!is_pow2(n) || n < 2 || m < 1) return fail(BBT_EUNSUPPORTED, "bad channelizer shape");
  if (n_spec <= 0 || n_bins <= 0) return BBT_OK;
  if (n_bins > 65535) return fail(BBT_EUNSUPPORTED, "too many bins per call");
  ChanPowArgs a{};
  a.in = static_cast<const cf2*>(in);
  a.out = static_cast<float*>(sum);
  a.count = static_cast<unsigned long long*>(count);
  a.offsets = reinterpret_cast<const long long*>(offsets);
  a.tw = twiddle_table();
  a.M = m;
  a.n_spec = n_spec;
  a.j_first = j_first;
  a.b_first = b_first;
  return run_chanpow<true>(ilog2(n), a, n_bins,
                           std::max<int64_t>(1, ceil_div(n_spec, n_bins)),
                           as_stream(stream));
}

int bbt_integrate_exec(const void* in, int64_t n, int64_t inner,
                       int64_t i_first, const int64_t* offsets,
                       int64_t b_first, int64_t n_bins, void* sum, void* count,
                       void* stream) {
  if (!in || !sum || !count || !offsets) return fail(BBT_EINVAL, "null argument");
  if (n <= 0 || n_bins <= 0 || inner <= 0) return BBT_OK;
  if (n_bins > 65535) return fail(BBT_EUNSUPPORTED, "too many bins per call");
  IntegrateArgs a;
  a.in = static_cast<const float*>(in);
  a.sum = static_cast<float*>(sum);
  a.count = static_cast<unsigned long long*>(count);
  a.offsets = reinterpret_cast<const long long*>(offsets);
  a.inner = inner;
  a.n = n;
  a.i_first = i_first;
  a.b_first = b_first;
  const int64_t want = (int64_t)sm_count() * 2048;
  int64_t msub = ceil_div(want, inner * n_bins);
  msub = std::max<int64_t>(1, std::min<int64_t>(msub, ceil_div(n, n_bins)));
  a.msub = msub;
  dim3 grid((unsigned)ceil_div(msub * inner, 256), (unsigned)n_bins);
  BBT_LAUNCH(integrate_kernel, grid, dim3(256), 0, as_stream(stream), a);
  return check_launch("integrate kernel");
}

int bbt_fold_exec(const void* in, int power, int64_t n, int64_t inner,
                  int64_t i_first, const int64_t* lo, const int64_t* hi,
                  int64_t b_first, int64_t n_bins, const int32_t* pbin,
                  const double* coef, int ncoef, double i_ref, double rate,
                  int n_phase, void* sum, void* count, void* stream) {
  if (!in || !sum || !count || !lo || !hi) return fail(BBT_EINVAL, "null argument");
  if (!pbin && (!coef || ncoef < 1 || ncoef > 8))
    return fail(BBT_EINVAL, "need phase bins or 1..8 polynomial coefficients");
  if (n_phase < 1 || inner < 1 || (power && inner % 4))
    return fail(BBT_EINVAL, "bad fold shape");
  if (n <= 0 || n_bins <= 0) return BBT_OK;
  if (n_bins > 65535) return fail(BBT_EUNSUPPORTED, "too many bins per call");
  FoldArgs a{};
  a.in = in;
  a.sum = static_cast<float*>(sum);
  a.count = static_cast<unsigned long long*>(count);
  a.lo = reinterpret_cast<const long long*>(lo);
  a.hi = reinterpret_cast<const long long*>(hi);
  a.pbin = pbin;
  a.inner = inner;
  a.n = n;
  a.i_first = i_first;
  a.b_first = b_first;
  a.i_ref = i_ref;
  a.rate = rate;
  a.ncoef = pbin ? 1 : ncoef;
  for (int k = 0; k < 8; ++k) a.coef[k] = (!pbin && k < ncoef) ? coef[k] : 0.;
  a.n_phase = n_phase;
  const size_t smem = (size_t)n_phase * (inner + 1) * 4;
  a.use_smem = smem <= 40 * 1024;
  const int64_t chunks = std::max<int64_t>(
      1, std::min<int64_t>(ceil_div((int64_t)sm_count() * 8, n_bins),
                           ceil_div(n, n_bins * 1024)));
  dim3 grid((unsigned)chunks, (unsigned)n_bins);
  if (power)
    BBT_LAUNCH(fold_kernel<true>, grid, dim3(256), a.use_smem ? smem : 0,
               as_stream(stream), a);
  else
    BBT_LAUNCH(fold_kernel<false>, grid, dim3(256), a.use_smem ? smem : 0,
               as_stream(stream), a);
  return check_launch("fold kernel");
}

int bbt_average_exec(const void* sum, const void* count, void* out,
                     int64_t n_bins, int64_t inner, void* stream) {
  if (!sum || !count || !out) return fail(BBT_EINVAL, "null argument");
  if (n_bins <= 0 || inner <= 0) return BBT_OK;
  BBT_LAUNCH(average_kernel, dim3(grid_for(n_bins * inner, 256)), dim3(256), 0,
             as_stream(stream), static_cast<const float*>(sum),
             static_cast<const unsigned long long*>(count),
             static_cast<float*>(out), (long long)n_bins, (long long)inner);
  return check_launch("average kernel");
}

// ------------------------------------------------------- measurement helper
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

// C-ABI implementation (see include/bbt_b200.h): FFT plans.
#include "common.cuh"

using namespace bbt;

namespace {

template <int L, bool LANEFAST>
int launch_fft(int kind, const FftArgs& a, bbt_stream_t st) {
  using C = typename DefaultCfg<L>::type;
  const int64_t lanes = LANEFAST ? a.outer * a.inner : a.outer;
  const int64_t blocks = ceil_div(lanes, C::G);
  if (blocks <= 0) return BBT_OK;
  if (blocks > 2147483647LL) return fail(BBT_EUNSUPPORTED, "grid too large");
  const size_t smem = C::SMEM_BYTES;
  auto kern = kind == BBT_C2C   ? fft_c2c_kernel<C, LANEFAST>
              : kind == BBT_R2C ? fft_r2c_kernel<C, LANEFAST>
                                : fft_c2r_kernel<C, LANEFAST>;
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
  if (rc == BBT_EUNSUPPORTED && last_error().empty())
    fail(rc, "unsupported FFT length");
  return rc;
}

}  // namespace

struct bbt_fft_plan {
  int64_t n, outer, inner;
  int kind, direction;
  float scale;
  int log2n, log2n1, log2n2;  // n = n1*n2 when n > kTwiddleTable
  const cf* tw;   // roots of unity for n (or n2 of the four-step split)
  const cf* tw1;  // roots of unity for n1
  cf* big_lo;
  cf* big_hi;
};

extern "C" {

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
                  "FFT lengths above 16384 need complex data on a contiguous "
                  "axis (inner == 1)");
    if (l > 2 * kLog2TwiddleTable)
      return fail(BBT_EUNSUPPORTED, "FFT length above 2^28");
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
  p->tw = twiddle_table(l <= kLog2TwiddleTable ? l : (l + 1) / 2);
  p->tw1 = l <= kLog2TwiddleTable ? p->tw : twiddle_table(l - (l + 1) / 2);
  if (!p->tw || !p->tw1) {
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
  FftArgs col{in, w, p->tw1, p->outer, n2, inverse, 1.f};
  if ((rc = run_fft(p->log2n1, BBT_C2C, col, st))) return rc;
  BBT_LAUNCH(twiddle_kernel, dim3(tw_blocks), dim3(256), 0, st, w, n1, n2,
             p->outer, big, inverse);
  if ((rc = check_launch("twiddle kernel"))) return rc;
  FftArgs row{w, w, p->tw, p->outer * n1, 1, inverse, p->scale};
  if ((rc = run_fft(p->log2n2, BBT_C2C, row, st))) return rc;
  // w[k1][k2] -> out[k2][k1]  (bin k = k1 + n1*k2); the batch goes in grid.z,
  // at most 65535 at a time.
  for (int64_t b0 = 0; b0 < p->outer; b0 += 65535) {
    const int64_t nb = std::min<int64_t>(65535, p->outer - b0);
    dim3 grid((unsigned)ceil_div(n2, 32), (unsigned)ceil_div(n1, 32),
              (unsigned)nb);
    BBT_LAUNCH(transpose_kernel, grid, dim3(32, 8), 32 * 33 * sizeof(cf), st,
               w + b0 * p->n, static_cast<cf*>(out) + b0 * p->n, n1, n2);
    if ((rc = check_launch("transpose kernel"))) return rc;
  }
  return BBT_OK;
}

int bbt_fft_plan_destroy(bbt_fft_plan* p) {
  if (!p) return BBT_OK;
  if (p->big_lo) dev_free(p->big_lo);
  if (p->big_hi) dev_free(p->big_hi);
  delete p;
  return BBT_OK;
}

}  // extern "C"

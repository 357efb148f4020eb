// C-ABI implementation (see include/bbt_b200.h): FFT plans.
//
// A plan transforms the middle axis of [outer][n][inner] data:
//   * powers of two up to 16384: one block-FFT kernel (any inner);
//   * larger powers of two: four-step (columns, twiddle, rows, transpose);
//   * any other length: Bluestein's algorithm on a power-of-two length
//     m >= 2n - 1 (three transforms of length m and three pointwise passes);
//   * real transforms (np.fft.rfft / irfft, fourier/numpy.py:41-49) beyond the
//     single-kernel sizes go through the complex transform of the same length
//     (conversion in, n/2+1 bins out; Hermitian extension in, real part out).
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

unsigned blocks_for(int64_t total) {
  return (unsigned)std::max<int64_t>(
      1, std::min<int64_t>(ceil_div(total, 256), (int64_t)sm_count() * 32));
}

// Complex transform of a power-of-two length (tables per length).
struct Pow2 {
  int64_t n = 0;
  int log2n = 0, log2n1 = 0, log2n2 = 0;
  const cf* tw = nullptr;   // roots of unity of n (or n2 of the split)
  const cf* tw1 = nullptr;  // of n1
  cf* big_lo = nullptr;
  cf* big_hi = nullptr;

  int init(int64_t n_) {
    n = n_;
    log2n = ilog2(n);
    if (log2n > 2 * kLog2TwiddleTable)
      return fail(BBT_EUNSUPPORTED, "FFT length above 2^28");
    const bool big = log2n > kLog2TwiddleTable;
    tw = twiddle_table(big ? (log2n + 1) / 2 : log2n);
    tw1 = big ? twiddle_table(log2n - (log2n + 1) / 2) : tw;
    if (!tw || !tw1) return fail(BBT_ENOMEM, "cannot allocate twiddle table");
    if (big) {
      log2n2 = (log2n + 1) / 2;
      log2n1 = log2n - log2n2;
      big_lo = make_roots(kTwiddleTable, (double)n);
      big_hi = make_roots(n >> kLog2TwiddleTable,
                          (double)n / (double)kTwiddleTable);
      if (!big_lo || !big_hi)
        return fail(BBT_ENOMEM, "cannot allocate twiddle tables");
    }
    return BBT_OK;
  }
  void release() {
    if (big_lo) dev_free(big_lo);
    if (big_hi) dev_free(big_hi);
    big_lo = big_hi = nullptr;
  }
  bool large() const { return log2n > kLog2TwiddleTable; }
  int64_t scratch_elems(int64_t outer, int64_t inner) const {
    return large() ? outer * n * inner : 0;
  }
  // [outer][n][inner] complex, in -> out (in == out allowed); `scratch` holds
  // scratch_elems() values for the four-step transforms.
  int run(const cf* in, cf* out, cf* scratch, int64_t outer, int64_t inner,
          int inverse, float scale, bbt_stream_t st) const {
    if (!large()) {
      FftArgs a{in, out, tw, outer, inner, inverse, scale};
      return run_fft(log2n, BBT_C2C, a, st);
    }
    // Four-step, X[k1 + n1 k2]: columns (n1) -> twiddle -> rows (n2) ->
    // transpose to natural order.
    const int64_t n1 = int64_t(1) << log2n1, n2 = int64_t(1) << log2n2;
    BigTwiddle big{big_lo, big_hi};
    int rc;
    FftArgs col{in, scratch, tw1, outer, n2 * inner, inverse, 1.f};
    if ((rc = run_fft(log2n1, BBT_C2C, col, st))) return rc;
    BBT_LAUNCH(twiddle_kernel, dim3(blocks_for(outer * n * inner)), dim3(256),
               0, st, scratch, n1, n2, outer, inner, big, inverse);
    if ((rc = check_launch("twiddle kernel"))) return rc;
    FftArgs row{scratch, scratch, tw, outer * n1, inner, inverse, scale};
    if ((rc = run_fft(log2n2, BBT_C2C, row, st))) return rc;
    if (inner > 1) {
      BBT_LAUNCH(transpose_inner_kernel, dim3(blocks_for(outer * n * inner)),
                 dim3(256), 0, st, scratch, out, outer, n1, n2, inner);
      return check_launch("transpose kernel");
    }
    // w[k1][k2] -> out[k2][k1]; the batch goes in grid.z, 65535 at a time.
    for (int64_t b0 = 0; b0 < outer; b0 += 65535) {
      const int64_t nb = std::min<int64_t>(65535, outer - b0);
      dim3 grid((unsigned)ceil_div(n2, 32), (unsigned)ceil_div(n1, 32),
                (unsigned)nb);
      BBT_LAUNCH(transpose_kernel, grid, dim3(32, 8), 32 * 33 * sizeof(cf), st,
                 scratch + b0 * n, out + b0 * n, n1, n2);
      if ((rc = check_launch("transpose kernel"))) return rc;
    }
    return BBT_OK;
  }
};

}  // namespace

struct bbt_fft_plan {
  int64_t n, outer, inner;
  int kind, direction;
  float scale;
  bool direct;      // one block-FFT kernel does everything (also real kinds)
  bool bluestein;   // n is not a power of two
  Pow2 fft;         // of length n (power of two) or m (Bluestein)
  int64_t m;        // Bluestein length
  cf* chirp;        // a_j = exp(-i pi j^2 / n), j < n
  cf* filter;       // FFT_m of conj(a) wrapped around
};

namespace {

// The Bluestein tables: a_j in float64 with j^2 reduced mod 2n exactly, and
// the transform of the wrapped conjugate chirp (computed here on the device).
int make_bluestein(bbt_fft_plan* p) {
  const int64_t n = p->n, m = p->m;
  std::vector<cf> a(n), c(m, mk(0.f, 0.f));
  for (int64_t j = 0; j < n; ++j) {
    const int64_t r = (int64_t)(((unsigned __int128)j * j) % (2 * n));
    const double ang = -M_PI * (double)r / (double)n;
    a[j] = mk((float)cos(ang), (float)sin(ang));
    const cf conj_a = mk(a[j].x, -a[j].y);
    c[j] = conj_a;
    if (j) c[m - j] = conj_a;
  }
  void *da = nullptr, *db = nullptr, *scratch = nullptr;
  int rc = BBT_OK;
  if (dev_alloc(&da, n * sizeof(cf)) || dev_alloc(&db, m * sizeof(cf)))
    rc = BBT_ENOMEM;
  p->chirp = static_cast<cf*>(da);
  p->filter = static_cast<cf*>(db);
  const int64_t se = p->fft.scratch_elems(1, 1);
  if (!rc && se && dev_alloc(&scratch, se * sizeof(cf))) rc = BBT_ENOMEM;
  if (!rc && (h2d(da, a.data(), n * sizeof(cf), 0) ||
              h2d(db, c.data(), m * sizeof(cf), 0)))
    rc = BBT_ECUDA;
  if (!rc)
    rc = p->fft.run(p->filter, p->filter, static_cast<cf*>(scratch), 1, 1, 0,
                    1.f, (bbt_stream_t)0);
#if !defined(BBT_EMULATE)
  if (cudaStreamSynchronize(0) != cudaSuccess && !rc) rc = BBT_ECUDA;
#endif
  if (scratch) dev_free(scratch);
  return rc ? fail(rc, "cannot set up the chirp-z tables") : BBT_OK;
}

}  // namespace

extern "C" {

int bbt_fft_plan_create(bbt_fft_plan** plan, int64_t n, int64_t outer,
                        int64_t inner, int kind, int direction, double scale) {
  if (!plan) return fail(BBT_EINVAL, "null plan pointer");
  *plan = nullptr;
  if (n < 1 || outer < 0 || inner < 1) return fail(BBT_EINVAL, "bad FFT shape");
  if (n < 2) return fail(BBT_EUNSUPPORTED, "FFT length must be at least 2");
  if (kind < BBT_C2C || kind > BBT_C2R) return fail(BBT_EINVAL, "bad FFT kind");
  bbt_fft_plan* p = new bbt_fft_plan();
  p->n = n;
  p->outer = outer;
  p->inner = inner;
  p->kind = kind;
  p->direction = direction == BBT_BACKWARD ? BBT_BACKWARD : BBT_FORWARD;
  p->scale = (float)scale;
  p->bluestein = !is_pow2(n);
  p->chirp = p->filter = nullptr;
  p->m = n;
  if (p->bluestein) {
    p->m = 2;
    while (p->m < 2 * n - 1) p->m <<= 1;
  }
  p->direct = !p->bluestein && ilog2(n) <= kLog2TwiddleTable;
  int rc = p->fft.init(p->m);
  if (!rc && p->bluestein) rc = make_bluestein(p);
  if (rc) {
    bbt_fft_plan_destroy(p);
    return rc;
  }
  *plan = p;
  return BBT_OK;
}

int64_t bbt_fft_plan_work_bytes(const bbt_fft_plan* p) {
  if (!p || p->direct) return 0;
  // One complex array of the (padded) transform length, plus the scratch of a
  // four-step transform of that length.
  const int64_t staged = p->outer * p->m * p->inner;
  return (staged + p->fft.scratch_elems(p->outer, p->inner)) *
         (int64_t)sizeof(cf);
}

int bbt_fft_exec(const bbt_fft_plan* p, const void* in, void* out, void* work,
                 void* stream) {
  if (!p || !in || !out) return fail(BBT_EINVAL, "null argument");
  bbt_stream_t st = as_stream(stream);
  if (p->outer == 0) return BBT_OK;
  const int inverse = p->direction == BBT_BACKWARD;
  if (p->direct) {
    FftArgs a{in, out, p->fft.tw, p->outer, p->inner, inverse, p->scale};
    return run_fft(p->fft.log2n, p->kind, a, st);
  }
  if (!work) return fail(BBT_EINVAL, "this FFT needs a work buffer");
  const int64_t outer = p->outer, inner = p->inner, n = p->n, m = p->m;
  cf* staged = static_cast<cf*>(work);
  cf* scratch = staged + outer * m * inner;
  const int64_t half = n / 2 + 1;
  int rc;
  if (!p->bluestein && p->kind == BBT_C2C)
    return p->fft.run(static_cast<const cf*>(in), static_cast<cf*>(out),
                      staged, outer, inner, inverse, p->scale, st);
  // Everything else: into the staged array (with the chirp for Bluestein),
  // transform, out of it.  The inverse transform of a length that is not a
  // power of two is conj(forward(conj(x))).
  AxisArgs a{};
  a.in = in;
  a.out = staged;
  a.chirp = p->bluestein ? p->chirp : nullptr;
  a.outer = outer;
  a.inner = inner;
  a.n = n;
  a.n_in = p->kind == BBT_C2R ? half : n;
  a.n_out = m;
  a.in_real = p->kind == BBT_R2C;
  a.conj_in = p->bluestein && inverse;
  a.scale = 1.f;
  const unsigned blocks = blocks_for(outer * m * inner);
  if (p->kind == BBT_C2R)
    BBT_LAUNCH(axis_hermitian_kernel, dim3(blocks), dim3(256), 0, st, a);
  else
    BBT_LAUNCH(axis_pre_kernel, dim3(blocks), dim3(256), 0, st, a);
  if ((rc = check_launch("FFT staging kernel"))) return rc;
  AxisArgs post{};
  post.in = staged;
  post.out = out;
  post.outer = outer;
  post.inner = inner;
  post.n = n;
  post.n_in = m;
  post.n_out = p->kind == BBT_R2C ? half : n;
  post.out_real = p->kind == BBT_C2R;
  post.scale = p->scale;
  if (p->bluestein) {
    // x a -> FFT_m -> x filter -> inverse FFT_m -> x a.
    if ((rc = p->fft.run(staged, staged, scratch, outer, inner, 0, 1.f, st)))
      return rc;
    AxisArgs f{};
    f.out = staged;
    f.chirp = p->filter;
    f.outer = outer;
    f.inner = inner;
    f.n_out = m;
    BBT_LAUNCH(axis_filter_kernel, dim3(blocks), dim3(256), 0, st, f);
    if ((rc = check_launch("FFT filter kernel"))) return rc;
    if ((rc = p->fft.run(staged, staged, scratch, outer, inner, 1,
                         (float)(1.0 / (double)m), st)))
      return rc;
    post.chirp = p->chirp;
    post.conj_out = inverse;
  } else {
    if ((rc = p->fft.run(staged, staged, scratch, outer, inner, inverse, 1.f,
                         st)))
      return rc;
  }
  BBT_LAUNCH(axis_post_kernel, dim3(blocks_for(outer * post.n_out * inner)),
             dim3(256), 0, st, post);
  return check_launch("FFT output kernel");
}

int bbt_fft_plan_destroy(bbt_fft_plan* p) {
  if (!p) return BBT_OK;
  p->fft.release();
  if (p->chirp) dev_free(p->chirp);
  if (p->filter) dev_free(p->filter);
  delete p;
  return BBT_OK;
}

}  // extern "C"

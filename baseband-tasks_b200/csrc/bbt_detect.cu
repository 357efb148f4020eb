// C-ABI implementation (see include/bbt_b200.h): detection, integration, fold.
#include "common.cuh"
#include "kernels_detect.cuh"

using namespace bbt;

extern "C" {

// ----------------------------------------------------------------- detection
int bbt_power_exec(const void* in, void* out, int64_t a, int64_t b,
                   void* stream) {
  if (!in || !out) return fail(BBT_EINVAL, "null argument");
  if (a <= 0 || b <= 0) return BBT_OK;
  BBT_LAUNCH(power_kernel, dim3(grid_for(a * b, 256)), dim3(256), 0,
             as_stream(stream), static_cast<const cf*>(in),
             static_cast<float*>(out), (long long)a, (long long)b);
  return check_launch("power kernel");
}

int bbt_multiply_exec(const void* a, const void* b, void* out, int64_t n,
                      void* stream) {
  if (!a || !b || !out) return fail(BBT_EINVAL, "null argument");
  if (n <= 0) return BBT_OK;
  BBT_LAUNCH(multiply_kernel, dim3(grid_for(n, 256)), dim3(256), 0,
             as_stream(stream), static_cast<const cf*>(a),
             static_cast<const cf*>(b), static_cast<cf*>(out), (long long)n);
  return check_launch("multiply kernel");
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
// Shape of the channelizer tiles: 2^LOG2E values per thread; enough threads
// for 8 lanes (runs of 64 bytes per time sample) where the transform allows.
template <int L>
struct ChanCfg {
  static constexpr int LOG2E = L < 4 ? L : (L == 14 ? 5 : 4);
  static constexpr int T = (1 << L) >> LOG2E;
  static constexpr int THREADS = T * 8 > 256 ? (T * 8 > 1024 ? 1024 : T * 8)
                                             : 256;
  using type = FftCfg<L, LOG2E, THREADS>;
};

template <class C, bool INTEGRATE>
int launch_chanpow_cfg(ChanPowArgs& a, int64_t n_bins, int64_t max_width,
                       bbt_stream_t st);

template <int L, bool INTEGRATE>
int launch_chanpow(ChanPowArgs& a, int64_t n_bins, int64_t max_width,
                   bbt_stream_t st) {
  if constexpr (L == 10) {
    // Measured best for the 1024-channel case: 32 values per thread; 16
    // lanes when spectra are written out, 8 (with the accumulators in
    // registers) when they are integrated.
    if (!INTEGRATE)
      return launch_chanpow_cfg<FftCfg<10, 5, 512>, INTEGRATE>(a, n_bins,
                                                               max_width, st);
    if (tune("chanpow_e16", 0))   // 16 values per thread, twice the warps
      return launch_chanpow_cfg<FftCfg<10, 4, 512>, INTEGRATE>(a, n_bins,
                                                               max_width, st);
    // Narrow samples (one or two polarization pairs, the bulk-copy kernel):
    // 4 lanes per CTA, so that two CTAs share an SM and one runs its
    // butterflies while the other waits at a barrier or for its tile
    // (C4, per 32 frames: 8 lanes 2.11 ms, 4 lanes 1.64 ms).
    const int lanes = tune("chanpow_g", a.M <= 2 ? 4 : 8);
    if (lanes == 4)
      return launch_chanpow_cfg<FftCfg<10, 5, 128>, INTEGRATE>(a, n_bins,
                                                               max_width, st);
    if (lanes == 2)
      return launch_chanpow_cfg<FftCfg<10, 5, 64>, INTEGRATE>(a, n_bins,
                                                              max_width, st);
    return launch_chanpow_cfg<FftCfg<10, 5, 256>, INTEGRATE>(a, n_bins,
                                                             max_width, st);
  }
  return launch_chanpow_cfg<typename ChanCfg<L>::type, INTEGRATE>(
      a, n_bins, max_width, st);
}

template <class C, bool INTEGRATE>
int launch_chanpow_cfg(ChanPowArgs& a, int64_t n_bins, int64_t max_width,
                       bbt_stream_t st) {
  constexpr int64_t units = C::G / 2;  // (sub-stream, m) pairs per CTA
  // Sub-streams per bin so that the grid fills the GPU a few times over.
  // (Eight CTAs' worth per SM: with four, the last wave of a long launch
  // left a tenth of the time to a partly idle GPU.)
  const int64_t want = (int64_t)sm_count() * tune("chanpow_waves", 8) * units;
  int64_t msub = ceil_div(want, a.M * (INTEGRATE ? n_bins : 1));
  // Whole tiles of sub-streams: a CTA of the bulk-copy kernel takes
  // units / M adjacent spectra at a time.
  if (INTEGRATE && a.M <= units && units % a.M == 0) {
    const int64_t nj = units / a.M;
    msub = ceil_div(msub, nj) * nj;
  }
  if (msub > max_width) msub = max_width;
  // In-kernel averaging divides every partial sum: keep one per bin.
  if (msub < 1 || (INTEGRATE && a.average)) msub = 1;
  a.msub = msub;
  const int64_t blocks = ceil_div(msub * a.M, units);
  dim3 grid((unsigned)blocks, (unsigned)(INTEGRATE ? n_bins : 1));
  if constexpr (INTEGRATE && units >= 1 &&
                ChanPowTma<C>::smem_bytes(1) <= 220 * 1024) {
    // Narrow samples: whole tiles by bulk copies (chanpow_tma_kernel).
    if (a.M <= units && units % a.M == 0 &&
        !(reinterpret_cast<uintptr_t>(a.in) & 15) && tune("chanpow_tma", 1)) {
      const size_t smem = ChanPowTma<C>::smem_bytes(a.M);
      auto kern = chanpow_tma_kernel<C>;
      if (BBT_SET_SMEM(kern, smem))
        return fail(BBT_ECUDA, "cannot set shared memory size");
      prof_next_name = "chanpow_integrate";
      BBT_LAUNCH(kern, grid, dim3(C::THREADS), smem, st, a);
      return check_launch("channelize-power kernel");
    }
  }
  // A second tile for asynchronous staging where it fits (see the kernel).
  const bool stage =
      INTEGRATE && C::SMEM_BYTES + (size_t)C::G * C::N * sizeof(cf) <= 200 * 1024;
  const size_t smem =
      C::SMEM_BYTES + (stage ? (size_t)C::G * C::N * sizeof(cf) : 0);
  auto kern = chanpow_kernel<C, INTEGRATE>;
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
#define F(L) rc = launch_chanpow<L, INTEGRATE>(a, n_bins, max_width, st)
  BBT_FOR_LOG2(log2n, F)
#undef F
  if (rc == BBT_EUNSUPPORTED)
    fail(rc, "channelizer length must be a power of two in [2, 16384]");
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
  a.tw = twiddle_table(ilog2(n));
  a.M = m;
  a.n_spec = n_spec;
  return run_chanpow<false>(ilog2(n), a, 1, n_spec, as_stream(stream));
}

int bbt_channelize_power_integrate_exec(const void* in, int64_t n, int64_t m,
                                        int64_t n_spec, int64_t j_first,
                                        const int64_t* offsets,
                                        int64_t b_first, int64_t n_bins,
                                        void* sum, void* count, int average,
                                        void* stream) {
  if (!in || !sum || !count || !offsets) return fail(BBT_EINVAL, "null argument");
  if (!is_pow2(n) || n < 2 || m < 1) return fail(BBT_EUNSUPPORTED, "bad channelizer shape");
  if (n_spec <= 0 || n_bins <= 0) return BBT_OK;
  if (n_bins > 65535) return fail(BBT_EUNSUPPORTED, "too many bins per call");
  ChanPowArgs a{};
  a.in = static_cast<const cf2*>(in);
  a.out = static_cast<float*>(sum);
  a.count = static_cast<unsigned long long*>(count);
  a.offsets = reinterpret_cast<const long long*>(offsets);
  a.tw = twiddle_table(ilog2(n));
  a.M = m;
  a.n_spec = n_spec;
  a.j_first = j_first;
  a.b_first = b_first;
  a.average = average;
  return run_chanpow<true>(ilog2(n), a, n_bins,
                           std::max<int64_t>(1, ceil_div(n_spec, n_bins)),
                           as_stream(stream));
}

int bbt_integrate_exec(const void* in, int64_t n, int64_t inner,
                       int64_t i_first, const int64_t* offsets,
                       int64_t b_first, int64_t n_bins, void* sum, void* count,
                       int average, void* stream) {
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
  a.average = average;
  const int64_t want = (int64_t)sm_count() * 2048;
  int64_t msub = ceil_div(want, inner * n_bins);
  msub = std::max<int64_t>(1, std::min<int64_t>(msub, ceil_div(n, n_bins)));
  if (average) msub = 1;  // one partial sum per bin and call
  a.msub = msub;
  dim3 grid((unsigned)ceil_div(msub * inner, 256), (unsigned)n_bins);
  BBT_LAUNCH(integrate_kernel, grid, dim3(256), 0, as_stream(stream), a);
  return check_launch("integrate kernel");
}

int bbt_fold_exec(const void* in, int power, int64_t n, int64_t inner,
                  int64_t i_first, int64_t i_phase, const int64_t* lo,
                  const int64_t* hi,
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
  a.i_phase = i_phase;
  a.b_first = b_first;
  a.i_ref = i_ref;
  a.rate = rate;
  {
    // The reciprocal shortcut needs a normal rate whose significand is not
    // all ones (the one case outside Markstein's theorem).
    uint64_t bits;
    memcpy(&bits, &rate, 8);
    const bool all_ones = (bits & 0xfffffffffffffULL) == 0xfffffffffffffULL;
    a.inv_rate = (std::isnormal(rate) && rate > 0. && !all_ones) ? 1. / rate : 0.;
  }
  a.ncoef = pbin ? 1 : ncoef;
  for (int k = 0; k < 8; ++k) a.coef[k] = (!pbin && k < ncoef) ? coef[k] : 0.;
  a.n_phase = n_phase;
  const size_t smem = (size_t)n_phase * (inner + 1) * 4;
  a.use_smem = smem <= 40 * 1024;
  size_t smem_total = a.use_smem ? smem : 0;
  if (a.use_smem && inner == 4 && !(reinterpret_cast<uintptr_t>(in) & 15)) {
    a.use_smem = 2;
#if !defined(BBT_EMULATE) && BBT_FOLD_TMA
    if (!pbin) {
      // A ring of kFoldStages tiles filled by bulk copies, plus its barriers.
      a.use_smem = 3;
      a.ring_offset = (int)((smem + 127) / 128 * 128);
      smem_total = a.ring_offset + (size_t)kFoldStages * kFoldTile * 16 +
                   kFoldStages * (kFoldThreads / 32) * 8;
    }
#endif
  }
  const int64_t chunks = std::max<int64_t>(
      1, std::min<int64_t>(ceil_div((int64_t)sm_count() * 8, n_bins),
                           ceil_div(n, n_bins * 1024)));
  dim3 grid((unsigned)chunks, (unsigned)n_bins);
  prof_next_name = "fold";
  if (BBT_SET_SMEM(fold_kernel<true>, smem_total) ||
      BBT_SET_SMEM(fold_kernel<false>, smem_total))
    return fail(BBT_ECUDA, "cannot set shared memory size");
  if (power)
    BBT_LAUNCH(fold_kernel<true>, grid, dim3(kFoldThreads), smem_total,
               as_stream(stream), a);
  else
    BBT_LAUNCH(fold_kernel<false>, grid, dim3(kFoldThreads), smem_total,
               as_stream(stream), a);
  return check_launch("fold kernel");
}

}  // extern "C"
namespace {
// Real input, columns in pairs: one complex transform per pair
// (pfb_pair_kernel), in CTAs of THREADS threads.
template <int L, int THREADS>
int launch_pfb_pair(const PfbArgs& a, int kind, bbt_stream_t st) {
  using D = DefaultCfg<L>;
  using C = FftCfg<L, D::LOG2E, THREADS>;
  const int64_t blocks = ceil_div(a.n_spec * (a.inner / 2), C::G);
  if (blocks > 2147483647LL) return fail(BBT_EUNSUPPORTED, "grid too large");
  const size_t smem =
      std::max<size_t>(C::SMEM_BYTES, (size_t)C::G * C::NPAD * sizeof(cf));
  auto kern = kind == 2 ? pfb_pair_kernel<C, 2> : pfb_pair_kernel<C, 1>;
  if (BBT_SET_SMEM(kern, smem))
    return fail(BBT_ECUDA, "cannot set shared memory size");
  prof_next_name = "pfb";
  BBT_LAUNCH(kern, dim3((unsigned)blocks), dim3(C::THREADS), smem, st, a);
  return check_launch("polyphase filter bank kernel");
}

template <int L>
int launch_pfb(const PfbArgs& a, int kind, bbt_stream_t st) {
  using D = DefaultCfg<L>;
  using C = FftCfg<L, D::LOG2E, 256>;
  const int64_t blocks = ceil_div(a.n_spec * a.inner, C::G);
  if (blocks > 2147483647LL) return fail(BBT_EUNSUPPORTED, "grid too large");
  if (kind != 0 && a.inner <= C::G && C::G % a.inner == 0 &&
      (C::N * a.inner) % 4 == 0 &&
      !(reinterpret_cast<uintptr_t>(a.in) & (kind == 2 ? 3 : 15)) &&
      tune("pfb_vector", 0)) {
    // Real input: the FIR as a vectorised phase of its own.  Off by default:
    // measured slower than the fused per-thread FIR (C3: 2.59 against
    // 1.89 ms) -- the kernel's time goes into the 2048-point transforms of
    // real data as complex, not into the byte loads.
    const size_t y_bytes = (size_t)(C::G / a.inner) *
                           (C::N * a.inner + 16) * sizeof(float);
    const size_t smem = std::max<size_t>(C::SMEM_BYTES, y_bytes);
    auto kern = kind == 2 ? pfb_real_kernel<C, 2> : pfb_real_kernel<C, 1>;
    if (BBT_SET_SMEM(kern, smem))
      return fail(BBT_ECUDA, "cannot set shared memory size");
    prof_next_name = "pfb";
    BBT_LAUNCH(kern, dim3((unsigned)blocks), dim3(C::THREADS), smem, st, a);
    return check_launch("polyphase filter bank kernel");
  }
  if (kind != 0 && a.inner % 2 == 0 && C::N >= 4 &&
      !(reinterpret_cast<uintptr_t>(a.in) & (kind == 2 ? 1 : 7)) &&
      !(reinterpret_cast<uintptr_t>(a.out) & 15) && tune("pfb_pair", 1)) {
    // Real input, columns in pairs: one complex transform per pair.  Small
    // CTAs, so that several share an SM and their barriers do not line up.
    const int threads = tune("pfb_threads", 64);
    if (threads == 64) return launch_pfb_pair<L, 64>(a, kind, st);
    if (threads == 128) return launch_pfb_pair<L, 128>(a, kind, st);
    return launch_pfb_pair<L, 256>(a, kind, st);
  }
  const size_t smem = C::SMEM_BYTES;
  auto kern = kind == 2   ? pfb_kernel<C, 2>
              : kind == 1 ? pfb_kernel<C, 1>
                          : pfb_kernel<C, 0>;
  if (BBT_SET_SMEM(kern, smem))
    return fail(BBT_ECUDA, "cannot set shared memory size");
  prof_next_name = "pfb";
  BBT_LAUNCH(kern, dim3((unsigned)blocks), dim3(C::THREADS), smem, st, a);
  return check_launch("polyphase filter bank kernel");
}
}  // namespace
extern "C" {

int bbt_pfb_exec(const void* in, void* out, const void* response, int64_t n,
                 int64_t n_tap, int64_t inner, int64_t n_spec, int is_real,
                 void* stream) {
  if (!in || !out || !response) return fail(BBT_EINVAL, "null argument");
  if (!is_pow2(n) || n < 2 || n > 8192)
    return fail(BBT_EUNSUPPORTED,
                "polyphase filter bank needs a power-of-two number of "
                "samples per spectrum in [2, 8192]");
  if (n_tap < 1 || inner < 1 || is_real < 0 || is_real > 2)
    return fail(BBT_EINVAL, "bad filter bank shape");
  if (n_spec <= 0) return BBT_OK;
  PfbArgs a;
  a.in = in;
  a.out = static_cast<cf*>(out);
  a.h = static_cast<const float*>(response);
  a.tw = twiddle_table(ilog2(n));
  a.inner = inner;
  a.n_spec = n_spec;
  a.n_tap = (int)n_tap;
  int rc = BBT_EUNSUPPORTED;
#define F(L) rc = launch_pfb<L>(a, is_real, as_stream(stream))
  BBT_FOR_LOG2(ilog2(n), F)
#undef F
  return rc;
}

int bbt_shift_exec(const void* in, void* out, const int64_t* offset,
                   int64_t n_out, int64_t n_series, int item_bytes,
                   void* stream) {
  if (!in || !out || !offset) return fail(BBT_EINVAL, "null argument");
  if (item_bytes != 4 && item_bytes != 8)
    return fail(BBT_EUNSUPPORTED, "items must be 4 or 8 bytes");
  if (n_out <= 0 || n_series <= 0) return BBT_OK;
  const unsigned grid = grid_for(n_out * n_series, 256);
  const long long* off = reinterpret_cast<const long long*>(offset);
  if (item_bytes == 4)
    BBT_LAUNCH(shift_kernel<float>, dim3(grid), dim3(256), 0, as_stream(stream),
               static_cast<const float*>(in), static_cast<float*>(out), off,
               (long long)n_out, (long long)n_series);
  else
    BBT_LAUNCH(shift_kernel<cf>, dim3(grid), dim3(256), 0, as_stream(stream),
               static_cast<const cf*>(in), static_cast<cf*>(out), off,
               (long long)n_out, (long long)n_series);
  return check_launch("shift kernel");
}

int bbt_convert_exec(const void* in, void* out, int64_t n, int to_real,
                     void* stream) {
  if (!in || !out) return fail(BBT_EINVAL, "null argument");
  if (n <= 0) return BBT_OK;
  if (to_real)
    BBT_LAUNCH(complex_to_real_kernel, dim3(grid_for(n, 256)), dim3(256), 0,
               as_stream(stream), static_cast<const cf*>(in),
               static_cast<float*>(out), (long long)n);
  else
    BBT_LAUNCH(real_to_complex_kernel, dim3(grid_for(n, 256)), dim3(256), 0,
               as_stream(stream), static_cast<const float*>(in),
               static_cast<cf*>(out), (long long)n);
  return check_launch("conversion kernel");
}

int bbt_pair_frames_exec(const void* in, void* out, int64_t n_in,
                         int64_t samples_per_frame, int64_t n, int64_t n_series,
                         int64_t n_frames, void* stream) {
  if (!in || !out) return fail(BBT_EINVAL, "null argument");
  if (n_in <= 0 || samples_per_frame <= 0 || n < samples_per_frame ||
      n_series <= 0)
    return fail(BBT_EINVAL, "bad frame shape");
  if (n_frames <= 0) return BBT_OK;
  const int64_t n_pairs = (n_frames + 1) / 2;
  const int64_t per = n * n_series;
  // Two values per thread where both frames of a pair stay 8-byte aligned.
  const bool vec = !((samples_per_frame * n_series) & 1) && !(per & 1) &&
                   !(reinterpret_cast<uintptr_t>(in) & 7) &&
                   !(reinterpret_cast<uintptr_t>(out) & 15);
  const unsigned gy = (unsigned)std::min<int64_t>(n_pairs, 65535);
  const unsigned gx = (unsigned)std::max<int64_t>(
      1, std::min<int64_t>(ceil_div(per, 256 * (vec ? 2 : 1)),
                           ceil_div((int64_t)sm_count() * 16, gy)));
  if (vec)
    BBT_LAUNCH(pair_frames_kernel<true>, dim3(gx, gy), dim3(256), 0,
               as_stream(stream), static_cast<const float*>(in),
               static_cast<cf*>(out), (long long)n_in,
               (long long)samples_per_frame, (long long)n,
               (long long)n_series, (long long)n_frames);
  else
    BBT_LAUNCH(pair_frames_kernel<false>, dim3(gx, gy), dim3(256), 0,
               as_stream(stream), static_cast<const float*>(in),
               static_cast<cf*>(out), (long long)n_in,
               (long long)samples_per_frame, (long long)n,
               (long long)n_series, (long long)n_frames);
  return check_launch("frame pairing kernel");
}

int bbt_unpair_frames_exec(const void* in, void* out,
                           int64_t samples_per_frame, int64_t n_series,
                           int64_t n_frames, void* stream) {
  if (!in || !out) return fail(BBT_EINVAL, "null argument");
  if (samples_per_frame <= 0 || n_series <= 0)
    return fail(BBT_EINVAL, "bad frame shape");
  if (n_frames <= 0) return BBT_OK;
  const int64_t n_pairs = (n_frames + 1) / 2;
  const int64_t per = samples_per_frame * n_series;
  const bool vec = !(per & 1) && !(reinterpret_cast<uintptr_t>(in) & 15) &&
                   !(reinterpret_cast<uintptr_t>(out) & 7);
  const unsigned gy = (unsigned)std::min<int64_t>(n_pairs, 65535);
  const unsigned gx = (unsigned)std::max<int64_t>(
      1, std::min<int64_t>(ceil_div(per, 256 * (vec ? 2 : 1)),
                           ceil_div((int64_t)sm_count() * 16, gy)));
  if (vec)
    BBT_LAUNCH(unpair_frames_kernel<true>, dim3(gx, gy), dim3(256), 0,
               as_stream(stream), static_cast<const cf*>(in),
               static_cast<float*>(out), (long long)samples_per_frame,
               (long long)n_series, (long long)n_frames);
  else
    BBT_LAUNCH(unpair_frames_kernel<false>, dim3(gx, gy), dim3(256), 0,
               as_stream(stream), static_cast<const cf*>(in),
               static_cast<float*>(out), (long long)samples_per_frame,
               (long long)n_series, (long long)n_frames);
  return check_launch("frame unpairing kernel");
}

int bbt_decode_exec(const void* in, void* out, const float* levels, int64_t n,
                    int bps, void* stream) {
  if (!in || !out || !levels) return fail(BBT_EINVAL, "null argument");
  if (bps != 1 && bps != 2 && bps != 4 && bps != 8)
    return fail(BBT_EUNSUPPORTED, "bits per sample must be 1, 2, 4 or 8");
  if (reinterpret_cast<uintptr_t>(out) & 15)
    return fail(BBT_EINVAL, "decode output must be 16-byte aligned");
  if (n <= 0) return BBT_OK;
  BBT_LAUNCH(decode_kernel, dim3(grid_for((n + 3) / 4, 256)), dim3(256),
             256 * sizeof(float), as_stream(stream),
             static_cast<const unsigned char*>(in), static_cast<float*>(out),
             levels, (long long)n, bps);
  return check_launch("decode kernel");
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


}  // extern "C"

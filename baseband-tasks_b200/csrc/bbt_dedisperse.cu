// C-ABI implementation (see include/bbt_b200.h): coherent (de)dispersion.
#include "common.cuh"
#include "kernels_dedisperse.cuh"
#include "tma.cuh"

using namespace bbt;

struct bbt_dedisperse_plan {
  int64_t n, n_series, pad_start, n_valid, n_chirp;
  int log2n, log2n1, log2n2;
  int planar;      // work-buffer layout of the three-pass split
  int64_t work_pitch;  // series pitch of the work buffer (>= n_series)
  int half;        // 256-thread CTAs, half-size tiles
  int row2;        // row pass by dd_row2_kernel (chirp rows in its order)
  const cf* tw_sub;  // roots of unity for its sub-transforms (n2 / 32)
  const cf* tw1;   // roots of unity for the column FFTs (n1)
  const cf* tw2;   // for the row FFTs (n2), or the whole single-pass frame
  cf* big_lo;
  cf* big_hi;
  cf* chirp;        // [n_chirp][n1][n2]
  int* series_map;  // device
  // Chirp parameters kept on the device for the regenerating row pass.
  double* d_freq;
  double* d_fref;
  signed char* d_sb;
  double ch_d, ch_rate, ch_soff;
};

namespace {

constexpr int kColThreads = 512;
constexpr int64_t kMaxFramesPerLaunch = 65535;  // the frame goes in grid.y
constexpr int64_t kCounterBytes = 256;

// Elements per thread of the column FFTs.  HALF: 256-thread CTAs with half
// the lanes (64 KB tiles, two CTAs per SM).
template <int L1, bool HALF>
struct ColCfg {
  static constexpr int LOG2E = L1 <= 4 ? L1 : (L1 <= 8 ? 4 : 5);
  using type = FftCfg<L1, LOG2E, HALF ? kColThreads / 2 : kColThreads>;
};


int col_lanes(int l1) {
  const int log2e = l1 <= 4 ? l1 : (l1 <= 8 ? 4 : 5);
  return kColThreads >> (l1 - log2e);
}

// Column passes through tensor-map bulk copies; BBT_EUNSUPPORTED (without
// touching the error message) when the geometry does not fit a tensor map.
template <class C>
int launch_dd_col_tma(bool inverse, const DdArgs& a, int64_t n_frames,
                      bbt_stream_t st) {
  const int64_t N2 = a.N >> a.log2n1, S = a.S, n2s = N2 * S;
  const bool planar_src = inverse && a.planar;
  if (C::G > 256 || n2s < C::G || (n2s & 1)) return BBT_EUNSUPPORTED;
  DdColTma m;
  m.box_rows = C::N < 256 ? C::N : 256;
  m.n_boxes = C::N / m.box_rows;
  m.tn = 1;
  m.stagger_ns = C::THREADS <= 256 ? tune("col_stagger_ns", 0) : 0;
  TensorMap map;
  uint64_t dims[3], strides[3];
  uint32_t box[3];
  void* base;
  if (planar_src) {
    if (S > C::G || C::G % S || ((C::G / S) & 1) || N2 < C::G / S)
      return BBT_EUNSUPPORTED;
    m.tn = (int)(C::G / S);
    base = a.work;
    dims[0] = N2, dims[1] = S, dims[2] = n_frames * C::N;
    strides[0] = 8, strides[1] = N2 * 8, strides[2] = S * N2 * 8;
    box[0] = m.tn, box[1] = (uint32_t)S, box[2] = m.box_rows;
  } else {
    const int64_t frame_stride = inverse ? a.N * a.Sw : a.in_frame_stride;
    if (a.Sw != a.S) return BBT_EUNSUPPORTED;  // padded rows: per-thread loads
    if (frame_stride & 1) return BBT_EUNSUPPORTED;
    base = inverse ? static_cast<void*>(a.work)
                   : const_cast<void*>(static_cast<const void*>(a.in));
    dims[0] = n2s, dims[1] = C::N, dims[2] = n_frames;
    strides[0] = 8, strides[1] = n2s * 8, strides[2] = frame_stride * 8;
    box[0] = C::G, box[1] = m.box_rows, box[2] = 1;
  }
  if (reinterpret_cast<uintptr_t>(base) & 15) return BBT_EUNSUPPORTED;
  if (make_tensor_map(&map, base, 8, 3, dims, strides, box))
    return BBT_EUNSUPPORTED;
  constexpr size_t kTile = (size_t)C::N * C::G * sizeof(cf);
  // Behind the tile: the barrier and the twiddle tables of the kernel.
  constexpr int LE = C::LOG2E > 0 ? C::LOG2E : 1;
  const size_t smem = (C::SMEM_BYTES > kTile ? C::SMEM_BYTES : kTile) +
                      2 * sizeof(Mbar) +
                      (LE * C::G + 2 * C::T + 16 + 2) * sizeof(cf);
  // Tile counter of this pass, behind the frames in the work buffer.
  m.next_tile = reinterpret_cast<unsigned*>(
                    reinterpret_cast<char*>(a.work) +
                    n_frames * a.N * a.Sw * (int64_t)sizeof(cf)) +
                (inverse ? 16 : 0);
  if (dev_zero(m.next_tile, sizeof(unsigned), st))
    return fail(BBT_ECUDA, "cannot reset the tile counter");
  m.fast_tw = (C::G % S == 0 || planar_src) && tune("col_fast_tw", 1);
  const int64_t tiles = ceil_div(n2s, C::G) * n_frames;
  const int per_sm = C::THREADS <= 256 ? 2 : 1;
  const int64_t ctas = std::min<int64_t>(tiles, (int64_t)sm_count() * per_sm);
  int rc;
  if (inverse && a.detect) {
    auto kern = dd_col_tma_kernel<C, true, true>;
    if (BBT_SET_SMEM(kern, smem))
      return fail(BBT_ECUDA, "cannot set shared memory size");
    prof_next_name = "dd_col_inv";
    BBT_LAUNCH(kern, dim3((unsigned)ctas), dim3(C::THREADS), smem, st, a, m,
               map);
    rc = check_launch("dedispersion column kernel");
  } else if (inverse) {
    auto kern = dd_col_tma_kernel<C, true>;
    if (BBT_SET_SMEM(kern, smem))
      return fail(BBT_ECUDA, "cannot set shared memory size");
    prof_next_name = "dd_col_inv";
    BBT_LAUNCH(kern, dim3((unsigned)ctas), dim3(C::THREADS), smem, st, a, m,
               map);
    rc = check_launch("dedispersion column kernel");
  } else {
    auto kern = dd_col_tma_kernel<C, false>;
    if (BBT_SET_SMEM(kern, smem))
      return fail(BBT_ECUDA, "cannot set shared memory size");
    prof_next_name = "dd_col_fwd";
    BBT_LAUNCH(kern, dim3((unsigned)ctas), dim3(C::THREADS), smem, st, a, m,
               map);
    rc = check_launch("dedispersion column kernel");
  }
  return rc;
}

template <int L1, bool HALF = false>
int launch_dd_col(bool inverse, const DdArgs& a0, int64_t n_frames,
                  bbt_stream_t st) {
  using C = typename ColCfg<L1, HALF>::type;
  DdArgs a = a0;
  if (HALF) a.ahead *= 2;
  const int64_t cols = (a.N >> L1) * a.S;
  // (Fused power, planar source: the two polarizations of a sample must sit
  // in one warp, G / S <= 16 lanes apart; else the per-thread-load kernel,
  // whose lanes are flat columns.)
  const bool far_pairs = inverse && a.detect && a.planar && a.S <= C::G &&
                         C::G / a.S > 16;
  if (tune("col_tma", 1) && !far_pairs) {
    const int rc = launch_dd_col_tma<C>(inverse, a, n_frames, st);
    if (rc != BBT_EUNSUPPORTED) return rc;  // else: the per-thread-load kernels
  }
  dim3 grid((unsigned)ceil_div(cols, C::G), (unsigned)n_frames);
  const size_t smem = C::SMEM_BYTES;
  auto kern = inverse ? dd_col_inv_kernel<C> : dd_col_fwd_kernel<C>;
  if (BBT_SET_SMEM(kern, smem))
    return fail(BBT_ECUDA, "cannot set shared memory size");
  prof_next_name = inverse ? "dd_col_inv" : "dd_col_fwd";
  BBT_LAUNCH(kern, grid, dim3(C::THREADS), smem, st, a);
  return check_launch("dedispersion column kernel");
}

// HALF: 256-thread CTAs on half-size tiles, two per SM.  (Quarter-size tiles,
// four per SM, measured 3 % slower on C2.)
template <int L2, bool PLANAR, bool HALF = false>
int launch_dd_row(const DdArgs& a0, int64_t n_frames, bbt_stream_t st) {
  using C = FftCfg<L2, 5, HALF ? 256 : 512>;
  DdArgs a = a0;
  if (HALF) a.ahead *= 2;
  const int64_t n1 = a.N >> L2;
  int64_t blocks;
  a.row_sc = a.row_rpc = a.row_chunks = 1;
  if (PLANAR) {
    blocks = ceil_div(n1 * a.S, C::G);
  } else {
    const int sc = a.S < C::G ? (int)a.S : C::G;
    const int rpc = (C::G % sc == 0) ? C::G / sc : 1;
    a.row_sc = sc;
    a.row_rpc = rpc;
    a.row_chunks = (int)ceil_div(a.S, sc);
    blocks = ceil_div(n1, rpc) * ceil_div(a.S, sc);
  }
  if (blocks * n_frames > 2147483647LL)
    return fail(BBT_EUNSUPPORTED, "grid too large");
  // (HALF: rows of up to 8192 points in 256-thread CTAs, two per SM, which
  // do not wait for one another.)
  if constexpr (PLANAR && L2 >= 11 && (!HALF || L2 <= 13)) {
    if (a.tw_sub) {   // the plan stored its chirp for dd_row2_kernel
      if (reinterpret_cast<uintptr_t>(a.work) & 15)
        return fail(BBT_EINVAL, "work buffer must be 16-byte aligned");
      size_t smem = Row2Cfg<C>::kSmemBytes;
      const bool regen = a.ch_freq && tune("chirp_regen", 0);
      // row_landp 1: pieces of >= 2 KB landing at the pitch of the exchange
      // matrix; 2 (one row per tile): half of the next row lands in a side
      // buffer a whole row time ahead.
      constexpr int kCanLandP = L2 >= 13 ? 1 : 0;
      constexpr int kCanSide =
          (C::G == 1 && Row2Cfg<C>::kSmemBytesSide <= 232448) ? 2 : kCanLandP;
      const int landp = regen || !kCanLandP ? 0 : tune("row_landp", 0);
      auto kern = dd_row2_kernel<C, false>;
      if (regen) {
        kern = dd_row2_kernel<C, true>;
      } else if (landp == 2 && kCanSide == 2) {
        kern = dd_row2_kernel<C, false, kCanSide>;
        smem = Row2Cfg<C>::kSmemBytesSide;
      } else if (landp) {
        kern = dd_row2_kernel<C, false, kCanLandP>;
      }
      if (BBT_SET_SMEM(kern, smem))
        return fail(BBT_ECUDA, "cannot set shared memory size");
      const int64_t ctas = std::min<int64_t>(
          blocks * n_frames, (int64_t)sm_count() * (HALF ? 2 : 1));
      prof_next_name = "dd_row";
      BBT_LAUNCH(kern, dim3((unsigned)ctas), dim3(C::THREADS), smem, st, a);
      return check_launch("dedispersion row kernel");
    }
  }
  if constexpr (!PLANAR && !HALF) {
    // Interleaved rows whose series all fit one tile: whole rows by bulk
    // copies (dd_rowi_tma_kernel).
    if (a.S <= C::G && C::G % a.S == 0 && tune("rowi_tma", 0) &&
        !(reinterpret_cast<uintptr_t>(a.work) & 15)) {
      constexpr size_t kTile = (size_t)C::N * C::G * sizeof(cf);
      const size_t smem =
          (C::SMEM_BYTES > kTile ? C::SMEM_BYTES : kTile) + sizeof(Mbar);
      auto kern = dd_rowi_tma_kernel<C>;
      if (BBT_SET_SMEM(kern, smem))
        return fail(BBT_ECUDA, "cannot set shared memory size");
      const int64_t tiles = ceil_div(n1, C::G / a.S) * n_frames;
      const int64_t ctas = std::min<int64_t>(tiles, (int64_t)sm_count());
      prof_next_name = "dd_row";
      BBT_LAUNCH(kern, dim3((unsigned)ctas), dim3(C::THREADS), smem, st, a);
      return check_launch("dedispersion row kernel");
    }
  }
  if constexpr (PLANAR && !HALF) {
    // Persistent CTAs fed by bulk copies (one tile = one contiguous range).
    if constexpr (L2 == 14) {
      // Sixteen values per thread in 1024-thread CTAs: 8 warps per scheduler
      // instead of 4, at 64 registers and one more exchange per transform
      // (A/B knob row_e16; rows in natural order, i.e. plans made with row2=0).
      if (tune("row_e16", 0) && !(reinterpret_cast<uintptr_t>(a.work) & 15)) {
        using C16 = FftCfg<L2, 4, 1024>;
        const size_t smem = C16::SMEM_BYTES + sizeof(Mbar);
        auto kern = dd_row_tma_kernel<C16>;
        if (BBT_SET_SMEM(kern, smem))
          return fail(BBT_ECUDA, "cannot set shared memory size");
        const int64_t ctas =
            std::min<int64_t>(n1 * a.S * n_frames, (int64_t)sm_count());
        prof_next_name = "dd_row";
        BBT_LAUNCH(kern, dim3((unsigned)ctas), dim3(C16::THREADS), smem, st, a);
        return check_launch("dedispersion row kernel");
      }
    }
    if (tune("row_tma", 1) && !(reinterpret_cast<uintptr_t>(a.work) & 15)) {
      const size_t smem = C::SMEM_BYTES + sizeof(Mbar);
      auto kern = dd_row_tma_kernel<C>;
      if (BBT_SET_SMEM(kern, smem))
        return fail(BBT_ECUDA, "cannot set shared memory size");
      const int64_t ctas =
          std::min<int64_t>(blocks * n_frames, (int64_t)sm_count());
      prof_next_name = "dd_row";
      BBT_LAUNCH(kern, dim3((unsigned)ctas), dim3(C::THREADS), smem, st, a);
      return check_launch("dedispersion row kernel");
    }
  }
  dim3 grid((unsigned)(blocks * n_frames));
  const size_t smem = C::SMEM_BYTES;
  auto kern = dd_row_kernel<C, PLANAR>;
  if (BBT_SET_SMEM(kern, smem))
    return fail(BBT_ECUDA, "cannot set shared memory size");
  prof_next_name = "dd_row";
  BBT_LAUNCH(kern, grid, dim3(C::THREADS), smem, st, a);
  return check_launch("dedispersion row kernel");
}

template <int L, bool LANEFAST>
int launch_dd_small(const DdArgs& a, int64_t n_frames, bbt_stream_t st) {
  using D = DefaultCfg<L>;
  using C = FftCfg<L, D::LOG2E, LANEFAST ? 512 : D::THREADS>;
  const int64_t blocks = ceil_div(n_frames * a.S, C::G);
  const size_t smem = C::SMEM_BYTES;
  auto kern = dd_small_kernel<C, LANEFAST>;
  if (BBT_SET_SMEM(kern, smem))
    return fail(BBT_ECUDA, "cannot set shared memory size");
  prof_next_name = "dd_small";
  BBT_LAUNCH(kern, dim3((unsigned)blocks), dim3(C::THREADS), smem, st, a,
             (long long)n_frames);
  return check_launch("dedispersion kernel");
}

#define BBT_FOR_ROW(L, F)   \
  switch (L) {              \
    case 10: F(10); break;  \
    case 11: F(11); break;  \
    case 12: F(12); break;  \
    case 13: F(13); break;  \
    case 14: F(14); break;  \
    default: break;         \
  }

}  // namespace

extern "C" {

int bbt_dedisperse_plan_create(bbt_dedisperse_plan** plan, int64_t n,
                               int64_t n_series, int64_t pad_start,
                               int64_t n_valid, int64_t n_chirp,
                               const int32_t* series_map,
                               const double* freq_mhz, const double* fref_mhz,
                               const int8_t* sideband, double dm,
                               double rate_mhz, double sample_offset,
                               int hint) {
  if (!plan) return fail(BBT_EINVAL, "null plan pointer");
  *plan = nullptr;
  if (n < 2 || !is_pow2(n))
    return fail(BBT_EUNSUPPORTED, "frame length must be a power of two >= 2");
  const int l = ilog2(n);
  if (l > 26) return fail(BBT_EUNSUPPORTED, "frame length above 2^26");
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
  p->d_freq = p->d_fref = nullptr;
  p->d_sb = nullptr;
  p->ch_d = p->ch_rate = p->ch_soff = 0.;
  p->planar = 1;
  if (!hint) hint = tune("dd_hint", 0);
  p->half = (hint >> 12) & 3;  // bit 0: column passes, bit 1: row pass
  const int hint_l1 = hint & 0xff;
  const bool force_planar = (hint >> 8) & 1, force_inter = (hint >> 9) & 1;
  if (l <= kLog2TwiddleTable &&
      (n_series == 1 || l <= tune("dd_small_max", 10)) && !hint_l1) {
    p->log2n1 = 0;  // single pass
    p->log2n2 = l;
  } else {
    // Planar candidate: longest row the row kernel takes.
    int l2 = l - 1 < 14 ? l - 1 : 14;
    if (l2 < 10) l2 = 10;
    int l1 = l - l2;
    bool planar = true;
    if (n_series > 1 && !force_planar) {
      const int tn = col_lanes(l1) / (int)std::min<int64_t>(
                                                     n_series, 1 << 20);
      // Rows of >= 8 interleaved series move in 64-byte runs as they are.
      if (((tn < 8 || n_series >= 8) && l - 10 <= 12) || force_inter) {
        planar = false;
        l2 = 10;
        l1 = l - 10;
      }
    }
    if (hint_l1) {
      l1 = hint_l1;
      l2 = l - l1;
    }
    if (l1 < 1 || l1 > 14 || l2 < 10 || l2 > 14) {
      delete p;
      return fail(BBT_EUNSUPPORTED, "unsupported split of the frame length");
    }
    p->log2n1 = l1;
    p->log2n2 = l2;
    p->planar = planar ? 1 : 0;
    // Interleaved rows run as two 64 KB tiles per SM (measured: a little
    // faster than one 128 KB tile); bit 14 of the hint switches that off.
    // ... unless a whole row (all series) fits one 512-thread tile: those go
    // by bulk copies (dd_rowi_tma_kernel).
    // Off by default: measured slower than two half tiles per SM (C2: 3.01
    // against 2.26 ms per 32 frames).
    const bool whole_rows = n_series <= 16 && 16 % n_series == 0 &&
                            tune("rowi_tma", 0);
    if (!planar && !((hint >> 14) & 1) && !whole_rows) p->half |= 2;
  }
  // Long contiguous rows: the formulation with warp-local sub-transforms.
  // Interleaved rows of many series whose count is not a multiple of eight
  // (C3: 2050) would put the 64-byte runs of a tile at odd offsets, every one
  // straddling sectors in all three passes: the work buffer pads the series
  // to a multiple of eight.  (Only where the column passes are the
  // per-thread-load kernels, N1 <= 16: the tensor-map tiles walk flat columns.)
  p->work_pitch = n_series;
  if (p->log2n1 > 0 && !p->planar && n_series > 16 && (n_series & 7) &&
      p->log2n1 <= 4 && tune("dd_work_pad", 1))
    p->work_pitch = (n_series + 7) & ~int64_t(7);
  p->row2 = p->log2n1 > 0 && p->planar && p->log2n2 >= 11 &&
            (!(p->half & 2) || p->log2n2 <= 13) && tune("row2", 1);
  p->tw_sub = p->row2 ? twiddle_table(p->log2n2 - 5) : nullptr;
  if (p->row2 && !p->tw_sub) p->row2 = 0;
  p->tw2 = twiddle_table(p->log2n2);
  p->tw1 = p->log2n1 > 0 ? twiddle_table(p->log2n1) : p->tw2;
  void* d = nullptr;
  int rc = BBT_OK;
  if (!p->tw1 || !p->tw2) rc = BBT_ENOMEM;
  if (!rc && dev_alloc(&d, n_chirp * n * sizeof(cf))) rc = BBT_ENOMEM;
  p->chirp = static_cast<cf*>(d);
  d = nullptr;
  if (!rc && dev_alloc(&d, n_series * sizeof(int))) rc = BBT_ENOMEM;
  p->series_map = static_cast<int*>(d);
  if (!rc && h2d(p->series_map, series_map, n_series * sizeof(int), 0))
    rc = BBT_ECUDA;
  if (!rc && p->log2n1 > 0) {
    p->big_lo = make_roots(kTwiddleTable, (double)n);
    const int64_t nhi = std::max<int64_t>(1, n >> kLog2TwiddleTable);
    p->big_hi = make_roots(nhi, (double)n / (double)kTwiddleTable);
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
      if (dfreq) dev_free(dfreq);
      if (dref) dev_free(dref);
      if (dsb) dev_free(dsb);
      bbt_dedisperse_plan_destroy(p);
      return fail(BBT_ENOMEM, "cannot stage chirp parameters");
    }
    ChirpArgs ca;
    ca.row2_log2n2 = p->row2 ? p->log2n2 : 0;
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
    p->d_freq = static_cast<double*>(dfreq);
    p->d_fref = static_cast<double*>(dref);
    p->d_sb = static_cast<signed char*>(dsb);
    p->ch_d = ca.d, p->ch_rate = rate_mhz, p->ch_soff = sample_offset;
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
  // An arbitrary response cannot be regenerated from chirp parameters.
  if (p->d_freq) dev_free(p->d_freq);
  p->d_freq = nullptr;
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
               int64_t(1) << p->log2n1, p->n_chirp,
               p->row2 ? p->log2n2 : 0);
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
        out[c * p->n + k1 + n1 * k2] =
            tmp[c * p->n + k1 * n2 +
                (p->row2 ? row2_pos(k2, p->log2n2) : k2)];
  return BBT_OK;
}

int64_t bbt_dedisperse_work_bytes(const bbt_dedisperse_plan* p,
                                  int64_t n_frames) {
  if (!p || p->log2n1 == 0) return 0;
  // Frames x points, and behind them the tile counters of the persistent
  // column passes (one per pass).
  return std::min<int64_t>(n_frames, kMaxFramesPerLaunch) * p->n *
             p->work_pitch * (int64_t)sizeof(cf) + kCounterBytes;
}

}  // extern "C"


static int dedisperse_exec_run(const bbt_dedisperse_plan* p, const void* in,
                               int64_t in_frame_stride, int64_t n_frames,
                               int64_t skip, void* out,
                               int64_t out_frame_stride, void* work,
                               void* stream, int detect);

static int dedisperse_exec_all(const bbt_dedisperse_plan* p, const void* in,
                               int64_t in_frame_stride, int64_t n_frames,
                               int64_t skip, void* out,
                               int64_t out_frame_stride, void* work,
                               void* stream, int detect) {
  if (!p || !in || !out) return fail(BBT_EINVAL, "null argument");
  if (n_frames <= 0) return BBT_OK;
  if (skip < 0 || skip >= p->n_valid) return fail(BBT_EINVAL, "bad skip");
  // Long runs of short frames go in several launches (same stream, so the
  // work buffer of the first part is free again when the next one starts).
  for (int64_t f0 = 0; f0 < n_frames; f0 += kMaxFramesPerLaunch) {
    const int64_t nf = std::min(kMaxFramesPerLaunch, n_frames - f0);
    const int rc = dedisperse_exec_run(
        p, static_cast<const cf*>(in) + f0 * in_frame_stride, in_frame_stride,
        nf, skip, static_cast<cf*>(out) + f0 * out_frame_stride,
        out_frame_stride, work, stream, detect);
    if (rc) return rc;
  }
  return BBT_OK;
}

extern "C" {

int bbt_dedisperse_power_supported(const bbt_dedisperse_plan* p) {
  // The products are formed in the last of the three passes, between the two
  // lanes of a polarization pair.
  return p && p->log2n1 > 0 && !(p->n_series & 1) && tune("dd_power", 1);
}

int bbt_dedisperse_power_exec(const bbt_dedisperse_plan* p, const void* in,
                              int64_t in_frame_stride, int64_t n_frames,
                              int64_t skip, void* out,
                              int64_t out_frame_stride, void* work,
                              void* stream) {
  if (!bbt_dedisperse_power_supported(p))
    return fail(BBT_EUNSUPPORTED,
                "no fused power for this plan (single pass or an odd number "
                "of series)");
  return dedisperse_exec_all(p, in, in_frame_stride, n_frames, skip, out,
                             out_frame_stride, work, stream, 1);
}

int bbt_dedisperse_exec(const bbt_dedisperse_plan* p, const void* in,
                        int64_t in_frame_stride, int64_t n_frames,
                        int64_t skip, void* out, int64_t out_frame_stride,
                        void* work, void* stream) {
  return dedisperse_exec_all(p, in, in_frame_stride, n_frames, skip, out,
                             out_frame_stride, work, stream, 0);
}

}  // extern "C"

static int dedisperse_exec_run(const bbt_dedisperse_plan* p, const void* in,
                               int64_t in_frame_stride, int64_t n_frames,
                               int64_t skip, void* out,
                               int64_t out_frame_stride, void* work,
                               void* stream, int detect) {
  bbt_stream_t st = as_stream(stream);
  DdArgs a;
  a.detect = detect;
  a.in = static_cast<const cf*>(in);
  a.out = static_cast<cf*>(out);
  a.work = static_cast<cf*>(work);
  a.tw = p->tw2;
  a.tw1 = p->tw1;
  a.tw_sub = p->tw_sub;
  a.ch_freq = p->d_freq, a.ch_fref = p->d_fref, a.ch_sb = p->d_sb;
  a.ch_d = p->ch_d, a.ch_rate = p->ch_rate, a.ch_soff = p->ch_soff;
  a.big = BigTwiddle{p->big_lo, p->big_hi};
  a.chirp = p->chirp;
  a.series_map = p->series_map;
  a.in_frame_stride = in_frame_stride;
  a.out_frame_stride = out_frame_stride;
  a.N = p->n;
  a.S = p->n_series;
  a.Sw = p->work_pitch;
  a.log2n1 = p->log2n1;
  a.log2n2 = p->log2n2;
  a.planar = p->planar;
  a.lo = (p->pad_start + skip) * p->n_series;
  a.hi = (p->pad_start + p->n_valid) * p->n_series;
  // Valid samples are stored from out + f*stride on, i.e. sample
  // pad_start + skip lands at offset 0.
  a.out_shift = (p->pad_start + skip) * p->n_series;
  a.scale = (float)(1.0 / (double)p->n);
  a.ahead = sm_count();
  a.n_frames = (int)n_frames;
  int rc = BBT_EUNSUPPORTED;
  if (p->log2n1 == 0) {
    const bool lanefast = p->n_series > 1 && tune("dd_small_lanefast", 1);
#define F(L)                                                        \
  rc = lanefast ? launch_dd_small<L, true>(a, n_frames, st)         \
                : launch_dd_small<L, false>(a, n_frames, st)
    BBT_FOR_LOG2(p->log2n2, F)
#undef F
    return rc;
  }
  if (!work) return fail(BBT_EINVAL, "dedispersion needs a work buffer");
  const bool hcol = p->half & 1, hrow = p->half & 2;
#define F(L)                                                        \
  rc = hcol ? launch_dd_col<L, true>(false, a, n_frames, st)        \
            : launch_dd_col<L, false>(false, a, n_frames, st)
  BBT_FOR_LOG2(p->log2n1, F)
#undef F
  if (rc) return rc;
  rc = BBT_EUNSUPPORTED;
#define F(L)                                                             \
  rc = hrow ? (p->planar ? launch_dd_row<L, true, true>(a, n_frames, st)  \
                         : launch_dd_row<L, false, true>(a, n_frames, st)) \
            : (p->planar ? launch_dd_row<L, true>(a, n_frames, st)        \
                         : launch_dd_row<L, false>(a, n_frames, st))
  BBT_FOR_ROW(p->log2n2, F)
#undef F
  if (rc) return rc;
  rc = BBT_EUNSUPPORTED;
#define F(L)                                                        \
  rc = hcol ? launch_dd_col<L, true>(true, a, n_frames, st)         \
            : launch_dd_col<L, false>(true, a, n_frames, st)
  BBT_FOR_LOG2(p->log2n1, F)
#undef F
  return rc;
}

extern "C" {

int bbt_dedisperse_plan_destroy(bbt_dedisperse_plan* p) {
  if (!p) return BBT_OK;
  if (p->big_lo) dev_free(p->big_lo);
  if (p->big_hi) dev_free(p->big_hi);
  if (p->chirp) dev_free(p->chirp);
  if (p->series_map) dev_free(p->series_map);
  if (p->d_freq) dev_free(p->d_freq);
  if (p->d_fref) dev_free(p->d_fref);
  if (p->d_sb) dev_free(p->d_sb);
  delete p;
  return BBT_OK;
}

}  // extern "C"

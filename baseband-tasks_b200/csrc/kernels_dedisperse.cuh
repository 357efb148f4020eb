// Overlap-save coherent (de)dispersion: ifft(fft(x) * chirp)[pad_start:pad_start+spf]
// (reference: baseband_tasks/dispersion.py:135-139, framing base.py:775-795).
//
// A frame of N = N1*N2 samples of S interleaved series ([N][S], time-major as
// the Task API delivers it) is convolved in three HBM round trips:
//   pass 1  dd_col_fwd : N1-point FFTs down the columns of the [N1][N2*S] view
//   pass 2  dd_row     : per row k1: x W_N^{k1 n2}, N2-point FFT, x chirp,
//                        inverse N2-point FFT, x conj(W_N^{k1 n2}) / N
//   pass 3  dd_col_inv : inverse N1-point FFTs down the columns, storing only
//                        the valid samples [pad_start+skip, pad_start+spf)
// The spectrum is never brought into natural order: bin k = k1 + N1*k2 lives at
// row k1, position k2, and the chirp is stored in that layout.
//
// Layout of the work buffer between the passes (element (k1, n2, s)):
//   interleaved  (k1*N2 + n2)*S + s   the input's own layout; pass 2 takes the
//                                     S series of a row as the lanes of a tile
//   planar       (k1*S + s)*N2 + n2   pass 1 de-interleaves on its stores, so
//                                     pass 2 streams whole contiguous rows
// chosen per plan so that every global access of every pass moves runs of at
// least 64-128 contiguous bytes.
// For N <= 16384 a single kernel (dd_small) does everything in one round trip.
#pragma once
#include "kernels_fft.cuh"
#include "tma.cuh"

namespace bbt {

struct DdArgs {
  const cf* in;         // first frame; frame f starts in_frame_stride later
  cf* out;              // valid output of frame f at out + f*out_frame_stride
  cf* work;             // n_frames * N * S scratch
  const cf* tw;         // N2-th roots of unity (row FFTs, single-pass frames)
  const cf* tw1;        // N1-th roots of unity (column FFTs)
  const cf* tw_sub;     // (N2/32)-th roots of unity (dd_row2_kernel)
  // Chirp parameters, for regenerating the chirp in the row pass instead of
  // reading the cached table (A/B variant, tuning knob chirp_regen).
  const double* ch_freq;
  const double* ch_fref;
  const signed char* ch_sb;
  double ch_d, ch_rate, ch_soff;
  BigTwiddle big;       // W_N^m
  const cf* chirp;      // [n_chirp][N1][N2]
  const int* series_map;  // series -> chirp index
  long long in_frame_stride, out_frame_stride;  // in complex elements
  long long N, S;       // frame length, interleaved series
  long long Sw;         // series pitch of the work buffer: S, or (interleaved
                        // layout, many series) S rounded up so that every
                        // run of a tile starts on a 64-byte boundary
  int log2n1, log2n2;
  int planar;           // work-buffer layout
  long long lo, hi;     // valid flat range [(pad_start+skip)*S, (pad_start+spf)*S)
  long long out_shift;  // (pad_start+skip)*S
  float scale;          // 1/N
  int ahead;            // CTAs resident at a time: L2 prefetch distance
  int n_frames;         // frames in this launch
  // Interleaved row tiles (set by the launcher): series per tile, rows per
  // tile, tiles per row.
  int row_sc, row_rpc, row_chunks;
  // Power fused into the last pass (bbt_dedisperse_power_exec): series 2q and
  // 2q+1 are the two polarizations of a pair, and instead of their voltages
  // X, Y the output holds [|X|^2, |Y|^2, Re(X conj Y), Im(X conj Y)] as four
  // floats in the same 16 bytes (functions.py:138-142).
  int detect;
};

// Ask L2 for `rows` runs of `run_bytes` each, `stride_bytes` apart: the tile a
// CTA launched `ahead` blocks later will load, so that its DRAM fetch overlaps
// this CTA's arithmetic.
BBT_HD void prefetch_tile(const cf* base, int rows, long long stride,
                          int run_elems, int tid, int nthreads) {
  const int lines = (run_elems * 8 + 127) / 128;
  if (lines == 1) {
    for (int r = tid; r < rows; r += nthreads) prefetch_l2(base + r * stride);
  } else {
    const int total = rows * lines;
    for (int i = tid; i < total; i += nthreads) {
      const int r = i / lines, l = i - r * lines;
      prefetch_l2(base + r * stride + l * 16);
    }
  }
}

// The twiddle W_N^{k1 n2} between the column and the row transforms is applied
// in the column passes (they are bound by memory and have arithmetic to
// spare; the row pass is bound by arithmetic): a thread holding the elements
// k1 = t + T e of column n2 multiplies them by scale * W_N^{n2 k1}, generated
// as base * step^e from a few table look-ups.
#ifndef BBT_RAMP_SQUARE
// 0: every power of the ramp step comes from the table.  1: powers by
// squaring (one look-up; each squaring doubles the rounding error of the
// step).  Measured on a 2^24-point dedispersion frame (tests/accuracy.py), max /
// RMS error in units of the RMS: 1.2e-5 / 1.3e-6 with squaring here and in
// apply_twiddles (BBT_TW_SQUARE=1) -- outside the 1e-5 parity tolerance --,
// 5.9e-6 / 9.6e-7 with 0 here, 2.5e-6 / 4.2e-7 with 0 here and
// BBT_TW_SQUARE=2 (the defaults; 2-4 % slower than all squaring).  numpy's
// single-precision path: 8.6e-7 / 1.9e-7.
#define BBT_RAMP_SQUARE 0
#endif
template <class C, int MODE>
BBT_HD void col_twiddle(cf* v, const BigTwiddle& big, int n2, int t, float scale) {
  cf pw[C::LOG2E > 0 ? C::LOG2E : 1];
  const cf base = cscale(big.get((long long)n2 * t), scale);
#if BBT_RAMP_SQUARE
  pw[0] = big.get((long long)n2 * C::T);
#pragma unroll
  for (int b = 1; b < C::LOG2E; ++b) pw[b] = cmul(pw[b - 1], pw[b - 1]);
#else
#pragma unroll
  for (int b = 0; b < C::LOG2E; ++b)
    pw[b] = big.get(((long long)n2 * C::T) << b);
#endif
  Ramp<C::LOG2E, MODE>::run(v, base, pw);
}

// Pass 1: forward column FFTs, frame -> work.  Lanes are consecutive flat
// columns q = n2*S + s.
template <class C>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS, (C::THREADS <= 256 ? 2 : 1))
    dd_col_fwd_kernel(DdArgs a) {
  cf* smem = BBT_SMEM(cf);
  const long long n2s = (a.N >> a.log2n1) * a.S;  // N2*S columns
  const int tid = threadIdx.x;
  const int g = tid % C::G, t = tid / C::G;
  const long long col = (long long)blockIdx.x * C::G + g;
  const long long frame = blockIdx.y;
  const bool valid = col < n2s;
  const cf* src = a.in + frame * a.in_frame_stride + col;
  cf v[C::E];
  {
    // Walk down the column: one 64-bit add per element.
    const cf* pe = src + (long long)t * n2s;
    const long long pstep = (long long)C::T * n2s;
#pragma unroll
    for (int e = 0; e < C::E; ++e) {
      v[e] = valid ? ld_stream(pe) : mk(0.f, 0.f);
      pe += pstep;
    }
  }
  {
    long long nb = (long long)blockIdx.x + a.ahead, nf = frame;
    if (nb >= gridDim.x) {
      nb -= gridDim.x;
      ++nf;
    }
    if (nf < gridDim.y && nb < gridDim.x) {
      const long long c0 = nb * C::G;
      const long long left = n2s - c0;
      prefetch_tile(a.in + nf * a.in_frame_stride + c0, C::N, n2s,
                    left < C::G ? (int)left : C::G, tid, C::THREADS);
    }
  }
  SmemLaneFast<C::PADSHIFT> sm{smem, g, C::G};
  block_fft<C>(v, t, a.tw1, sm);
  if (valid) {
    const unsigned n2 = (unsigned)col / (unsigned)a.S;
    col_twiddle<C, 0>(v, a.big, (int)n2, t, 1.f);
    cf* dst = a.work + frame * a.N * a.Sw;
    long long step;
    const unsigned s = (unsigned)col - n2 * (unsigned)a.S;
    const long long N2 = a.N >> a.log2n1;
    if (a.planar) {
      dst += s * N2 + n2;
      step = a.S * N2;
    } else {
      dst += (long long)n2 * a.Sw + s;
      step = N2 * a.Sw;
    }
    dst += (long long)t * step;
    const long long pstep = (long long)C::T * step;
#pragma unroll
    for (int e = 0; e < C::E; ++e) {
      *dst = v[e];
      dst += pstep;
    }
  }
}

// What a lane stores when Power is fused into the last pass: `mine` is its
// voltage, the lane `d` away (lane ^ d) holds the other polarization of the
// same sample.  The first lane of a pair (X) returns [|X|^2, |Y|^2], the
// second (Y) [Re(X conj Y), Im(X conj Y)] (functions.py:138-142).  All lanes
// of a warp must call this together.
BBT_DEV cf detect_pair(cf mine, bool second, int d) {
  cf other;
  other.x = shfl_xor(mine.x, d);
  other.y = shfl_xor(mine.y, d);
  cf r;
  if (second) {   // X = other, Y = mine
    r.x = other.x * mine.x + other.y * mine.y;
    r.y = other.y * mine.x - other.x * mine.y;
  } else {
    r.x = mine.x * mine.x + mine.y * mine.y;
    r.y = other.x * other.x + other.y * other.y;
  }
  return r;
}

// Pass 3: inverse column FFTs, work -> valid part of the output stream.
template <class C>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS, (C::THREADS <= 256 ? 2 : 1))
    dd_col_inv_kernel(DdArgs a) {
  cf* smem = BBT_SMEM(cf);
  const long long n2s = (a.N >> a.log2n1) * a.S;
  const int tid = threadIdx.x;
  const int g = tid % C::G, t = tid / C::G;
  const long long col = (long long)blockIdx.x * C::G + g;
  const long long frame = blockIdx.y;
  const bool valid = col < n2s;
  const cf* src = a.work + frame * a.N * a.Sw;
  long long step;
  {
    const unsigned n2 = (unsigned)col / (unsigned)a.S;
    const unsigned s = (unsigned)col - n2 * (unsigned)a.S;
    const long long N2 = a.N >> a.log2n1;
    if (a.planar) {
      src += s * N2 + n2;
      step = a.S * N2;
    } else {
      src += (long long)n2 * a.Sw + s;
      step = N2 * a.Sw;
    }
  }
  cf v[C::E];
  {
    const cf* pe = src + (long long)t * step;
    const long long pstep = (long long)C::T * step;
#pragma unroll
    for (int e = 0; e < C::E; ++e) {
      v[e] = valid ? ld_stream(pe) : mk(0.f, 0.f);
      pe += pstep;
    }
  }
  {
    long long nb = (long long)blockIdx.x + a.ahead, nf = frame;
    if (nb >= gridDim.x) {
      nb -= gridDim.x;
      ++nf;
    }
    if (nf < gridDim.y && nb < gridDim.x) {
      const long long c0 = nb * C::G;
      const cf* base = a.work + nf * a.N * a.Sw;
      if (a.planar) {
        // Runs of the tile's n2 values, one per (k1, s).
        const long long N2 = a.N >> a.log2n1;
        const unsigned n20 = (unsigned)c0 / (unsigned)a.S;
        const unsigned s0 = (unsigned)c0 - n20 * (unsigned)a.S;
        const int ns = a.S < C::G ? (int)a.S : C::G;       // series in tile
        const int tn = a.S < C::G ? C::G / (int)a.S : 1;   // n2 per series
        for (int si = 0; si < ns; ++si)
          prefetch_tile(base + (s0 + si) * N2 + n20, C::N, a.S * N2, tn, tid,
                        C::THREADS);
      } else {
        const long long left = n2s - c0;
        const unsigned n20 = (unsigned)c0 / (unsigned)a.S;
        const unsigned s0 = (unsigned)c0 - n20 * (unsigned)a.S;
        prefetch_tile(base + (long long)n20 * a.Sw + s0, C::N,
                      (a.N >> a.log2n1) * a.Sw,
                      left < C::G ? (int)left : C::G, tid, C::THREADS);
      }
    }
  }
  // The row pass left fft(conj(Y)) = conj(ifft(Y)) N; times the twiddle and
  // 1/N this is the conjugate of what the inverse column transform takes.
  if (valid) col_twiddle<C, 0>(v, a.big, (int)((unsigned)col / (unsigned)a.S), t, a.scale);
  SmemLaneFast<C::PADSHIFT> sm{smem, g, C::G};
  block_fft<C>(v, t, a.tw1, sm);
  if (a.detect) {
    // Power fused in (see DdArgs::detect): columns 2q, 2q+1 -- neighbouring
    // lanes -- are the two polarizations of a sample.
    long long flat = (long long)t * n2s + col;
    const long long fstep = (long long)C::T * n2s;
    cf* dst = a.out + frame * a.out_frame_stride - a.out_shift + flat;
    const bool second = (col & 1) != 0;
#pragma unroll
    for (int e = 0; e < C::E; ++e) {
      const cf r = detect_pair(cconj(v[e]), second, 1);
      if (valid && flat >= a.lo && flat < a.hi) *dst = r;
      flat += fstep;
      dst += fstep;
    }
  } else if (valid) {
    long long flat = (long long)t * n2s + col;
    const long long fstep = (long long)C::T * n2s;
    cf* dst = a.out + frame * a.out_frame_stride - a.out_shift + flat;
#pragma unroll
    for (int e = 0; e < C::E; ++e) {
      if (flat >= a.lo && flat < a.hi) *dst = cconj(v[e]);
      flat += fstep;
      dst += fstep;
    }
  }
}

// Pass 2 on one row per lane.  PLANAR: lanes are G consecutive rows
// rho = k1*S + s of the planar work buffer, threads of a lane walk along the
// contiguous row.  Otherwise lanes are the series of one row (or of several
// rows when S < G) of the interleaved buffer, consecutive threads taking
// consecutive series.
template <class C, bool PLANAR>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS, (C::THREADS <= 256 ? 2 : 1))
    dd_row_kernel(DdArgs a) {
  cf* smem = BBT_SMEM(cf);
  const long long n1 = a.N >> a.log2n2;
  const int tid = threadIdx.x;
  // Frames vary fastest over the grid: CTAs that run together share the
  // chirp rows, which then come from L2 for all but the first frame.
  const unsigned xblk = blockIdx.x / (unsigned)a.n_frames;
  const long long frame = blockIdx.x - xblk * (unsigned)a.n_frames;
  int t, g;
  long long k1, s, stride;
  bool valid;
  cf* row = a.work + frame * a.N * a.Sw;
  if (PLANAR) {
    t = tid % C::T;
    g = tid / C::T;
    const unsigned rho = xblk * C::G + g;
    valid = rho < (unsigned)(n1 * a.S);
    k1 = rho / (unsigned)a.S;
    s = rho - (unsigned)k1 * (unsigned)a.S;
    row += (long long)rho * C::N;
    stride = 1;
  } else {
    g = tid % C::G;
    t = tid / C::G;
    // Series per tile, rows per tile, tiles per row (from the launcher).
    const unsigned sc = a.row_sc, rpc = a.row_rpc, chunks = a.row_chunks;
    const unsigned rblk = xblk / chunks, chunk = xblk - rblk * chunks;
    const unsigned kl = (unsigned)g / sc, sl = (unsigned)g - kl * sc;
    k1 = rblk * rpc + kl;
    s = chunk * sc + sl;
    valid = kl < rpc && k1 < n1 && s < a.S;
    row += k1 * C::N * a.Sw + s;
    stride = a.Sw;
  }
  const cf* chirp = a.chirp;
  if (valid) chirp += ((long long)a.series_map[s] * n1 + k1) * C::N;
  if (valid && (t & 15) == 0) {
#pragma unroll
    for (int e = 0; e < C::E; ++e) prefetch_l2(chirp + t + C::T * e);
  }
  cf v[C::E];
  // Offsets within a row tile fit in 32 bits: independent address per element.
  const unsigned ustride = (unsigned)stride;
  const unsigned off0 = (unsigned)t * ustride;
#pragma unroll
  for (int e = 0; e < C::E; ++e)
    v[e] = valid ? ld_stream(row + (off0 + (unsigned)(C::T * e) * ustride))
                 : mk(0.f, 0.f);
  {
    // The rows the CTA `ahead` blocks later will load.
    const unsigned nlin = blockIdx.x + (unsigned)a.ahead;
    const unsigned nb = nlin / (unsigned)a.n_frames;
    const unsigned nf = nlin - nb * (unsigned)a.n_frames;
    if (nlin < gridDim.x) {
      const cf* base2 = a.work + nf * a.N * a.Sw;
      long long elems = (long long)C::G * C::N;
      if (PLANAR) {
        base2 += (long long)nb * C::G * C::N;
      } else {
        // The tiles of one row block (its `row_chunks` groups of series) are
        // interleaved in memory; together they form one contiguous range, of
        // which each of their CTAs fetches an equal contiguous share, so that
        // DRAM sees whole lines rather than the 64-byte pieces of one tile.
        const unsigned chunks = a.row_chunks;
        const unsigned rb = nb / chunks, ck = nb - rb * chunks;
        if (chunks > 4) {
          // Many series (C3: 2050, 257 tiles per row block of 17 MB): the
          // CTAs of a row block are too far apart in time for shares
          // fetched by one to still be in L2 when the others come for them
          // (measured: 2.6 x the data read from DRAM).  Each tile's own
          // runs, then: rpc * N of them, S elements apart.
          prefetch_tile(base2 + ((long long)rb * a.row_rpc) * C::N * a.Sw +
                            (long long)ck * a.row_sc,
                        (int)a.row_rpc * C::N, a.Sw, (int)a.row_sc, tid,
                        C::THREADS);
          elems = 0;
        } else {
          elems = (long long)a.row_rpc * C::N * a.Sw / chunks;
          base2 += ((long long)rb * a.row_rpc) * C::N * a.Sw + ck * elems;
        }
      }
      const long long limit =
          a.N * a.Sw - (base2 - (a.work + nf * a.N * a.Sw));
      if (elems > limit) elems = limit;
      for (long long i = (long long)tid * 16; i < elems;
           i += (long long)C::THREADS * 16)
        prefetch_l2(base2 + i);
    }
  }
  if (PLANAR) {
    SmemLaneSlow<C::PADSHIFT> sm{smem + (size_t)g * C::NPAD};
    block_fft<C>(v, t, a.tw, sm);
  } else {
    SmemLaneFast<C::PADSHIFT> sm{smem, g, C::G};
    block_fft<C>(v, t, a.tw, sm);
  }
  if (valid) {
#pragma unroll
    for (int e = 0; e < C::E; ++e)
      // Default cache policy: the row is re-read for the other frames.
      v[e] = cconj(cmul(v[e], ldtw(chirp, t + C::T * e)));
  }
  if (PLANAR) {
    SmemLaneSlow<C::PADSHIFT> sm{smem + (size_t)g * C::NPAD};
    block_fft<C>(v, t, a.tw, sm);
  } else {
    SmemLaneFast<C::PADSHIFT> sm{smem, g, C::G};
    block_fft<C>(v, t, a.tw, sm);
  }
  if (valid) {
    // fft(conj(Y)); the inverse column pass applies the twiddle and 1/N.
#pragma unroll
    for (int e = 0; e < C::E; ++e)
      row[off0 + (unsigned)(C::T * e) * ustride] = v[e];
  }
}

// Pass 2 for the planar layout as a persistent kernel fed by bulk copies (TMA):
// one CTA per SM walks over the tiles (G consecutive rows = one contiguous
// 128 KB range).  The exchange buffer is idle from the last exchange of the
// inverse transform until the first one of the next tile; in that window one
// thread lets a bulk copy (cp.async.bulk, completion on an mbarrier) bring the
// next tile into it, out of L2, where a bulk prefetch issued a tile earlier
// has put it.  The per-thread global loads, their address arithmetic and the
// per-line L2 prefetches of dd_row_kernel are gone: a tile starts with
// conflict-free reads from shared memory.
#if defined(__CUDACC__)
#define BBT_DEV_NOINLINE __device__ __noinline__
#else
#define BBT_DEV_NOINLINE static __attribute__((noinline))
#endif
// One tile of the persistent row pass.  Kept out of line so that the compiler
// treats it like the body of a one-tile kernel: inlined into the tile loop,
// loop-invariant address arithmetic was hoisted into registers the butterflies
// need (300 bytes of spills at the 128-register cap).
template <class C>
BBT_DEV_NOINLINE void dd_row_tma_tile(
    cf* smem, Mbar* bar, cf* row, const cf* chirp, const cf* tw, bool valid,
    unsigned phase, const char* next_src, unsigned next_bytes) {
  const int tid = threadIdx.x;
  const int t = tid % C::T, g = tid / C::T;
  constexpr unsigned kChunk = 32 * 1024;            // bytes per bulk copy
  mbar_wait(bar, phase & 1u, phase);
  cf v[C::E];
  {
    const cf* land = smem + g * C::N + t;
#pragma unroll
    for (int e = 0; e < C::E; ++e)
      v[e] = valid ? land[C::T * e] : mk(0.f, 0.f);
  }
  BBT_SYNC();  // the landing zone becomes the exchange buffer
  SmemLaneSlow<C::PADSHIFT> sm{smem + (size_t)g * C::NPAD};
  block_fft<C>(v, t, tw, sm);
  if (valid) {
#pragma unroll
    for (int e = 0; e < C::E; ++e)
      v[e] = cconj(cmul(v[e], ldtw(chirp, t + C::T * e)));
  }
  block_fft_head<C>(v, t, tw, sm);
  // Every thread is past its last read of the exchange buffer: the next
  // tile may land there while the last butterflies and the stores run.
  if (tid == 0 && next_bytes) {
    fence_proxy_async();  // the exchanges wrote here through the generic proxy
    mbar_expect_tx(bar, next_bytes);
    char* dst = reinterpret_cast<char*>(smem);
    for (unsigned o = 0; o < next_bytes; o += kChunk)
      bulk_load(dst + o, next_src + o,
                next_bytes - o < kChunk ? next_bytes - o : kChunk, bar,
                o + kChunk >= next_bytes);
  }
  block_fft_tail<C>(v, t, tw, sm);
  if (valid) {
#pragma unroll
    for (int e = 0; e < C::E; ++e) row[t + C::T * e] = v[e];
  }
}

template <class C>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS, 1) dd_row_tma_kernel(DdArgs a) {
  cf* smem = BBT_SMEM(cf);
  Mbar* bar = reinterpret_cast<Mbar*>(smem + C::SMEM_BYTES / sizeof(cf));
  const unsigned n1 = (unsigned)(a.N >> a.log2n2);
  const unsigned S = (unsigned)a.S;
  const unsigned rows = n1 * S;                     // rows per frame
  const unsigned tiles_per_frame = (rows + C::G - 1) / C::G;
  const unsigned n_tiles = tiles_per_frame * (unsigned)a.n_frames;
  const int tid = threadIdx.x;
  const int t = tid % C::T, g = tid / C::T;
  constexpr unsigned kChunk = 32 * 1024;            // bytes per bulk copy
  // Tile `lin`: frames vary fastest, so CTAs that run together share chirp rows.
  auto tile_base = [&](unsigned lin, unsigned& rho0) -> cf* {
    const unsigned xblk = lin / (unsigned)a.n_frames;
    const unsigned frame = lin - xblk * (unsigned)a.n_frames;
    rho0 = xblk * C::G;
    return a.work + (long long)frame * a.N * a.S + (long long)rho0 * C::N;
  };
  auto tile_bytes = [&](unsigned rho0) -> unsigned {
    const unsigned left = rows - rho0;
    return (left < (unsigned)C::G ? left : (unsigned)C::G) * C::N *
           (unsigned)sizeof(cf);
  };
  // (A fixed tile-to-CTA map: rows are contiguous, so unlike the column
  // passes nothing is gained by handing the tiles out in order -- measured
  // 4 % slower, for the extra barrier.)
  if (tid == 0) mbar_init(bar, 1);
  BBT_SYNC();
  if (tid == 0 && blockIdx.x < n_tiles) {
    unsigned rho0;
    const char* src = reinterpret_cast<const char*>(tile_base(blockIdx.x, rho0));
    const unsigned bytes = tile_bytes(rho0);
    mbar_expect_tx(bar, bytes);
    char* dst = reinterpret_cast<char*>(smem);
    for (unsigned o = 0; o < bytes; o += kChunk)
      bulk_load(dst + o, src + o, bytes - o < kChunk ? bytes - o : kChunk, bar,
                o + kChunk >= bytes);
  }
  unsigned k = 0;
#pragma unroll 1
  for (unsigned lin = blockIdx.x; lin < n_tiles; lin += gridDim.x, ++k) {
    const unsigned next = lin + gridDim.x;
    unsigned rho0;
    cf* row = tile_base(lin, rho0) + (long long)g * C::N;
    const unsigned rho = rho0 + g;
    const bool valid = rho < rows;
    const unsigned k1 = rho / S, s = rho - k1 * S;
    const cf* chirp = a.chirp;
    if (valid) chirp += ((long long)a.series_map[s] * n1 + k1) * C::N;
    // Into L2 meanwhile: this tile's chirp rows and the next tile.
    if (valid && t == 0)
      for (unsigned o = 0; o < C::N * sizeof(cf); o += kChunk)
        bulk_prefetch_l2(reinterpret_cast<const char*>(chirp) + o,
                         C::N * sizeof(cf) - o < kChunk
                             ? (unsigned)(C::N * sizeof(cf)) - o : kChunk);
    const char* next_src = nullptr;
    unsigned next_bytes = 0;
    if (next < n_tiles) {
      unsigned r0;
      next_src = reinterpret_cast<const char*>(tile_base(next, r0));
      next_bytes = tile_bytes(r0);
      if (tid == 32)
        for (unsigned o = 0; o < next_bytes; o += kChunk)
          bulk_prefetch_l2(next_src + o,
                           next_bytes - o < kChunk ? next_bytes - o : kChunk);
    }
    dd_row_tma_tile<C>(smem, bar, row, chirp, a.tw, valid, k, next_src,
                       next_bytes);
  }
}

// Pass 2 for the interleaved layout, fed by bulk copies: a tile is `rpc`
// consecutive rows with all S series (rpc S = G lanes), one contiguous range
// of G N2 values; lanes are (row, series) pairs with the series fastest, as the
// data lie in memory.  Otherwise as dd_row_tma_kernel.
template <class C>
BBT_DEV_NOINLINE void dd_rowi_tma_tile(
    cf* smem, Mbar* bar, cf* row, const cf* chirp, const cf* tw, int S,
    bool valid, unsigned phase, const char* next_src, unsigned next_bytes) {
  const int tid = threadIdx.x;
  const int g = tid % C::G, t = tid / C::G;
  const int kl = g / S, sl = g - kl * S;
  constexpr unsigned kChunk = 32 * 1024;
  mbar_wait(bar, phase & 1u, phase);
  cf v[C::E];
  {
    const cf* land = smem + ((size_t)kl * C::N + t) * S + sl;
#pragma unroll
    for (int e = 0; e < C::E; ++e)
      v[e] = valid ? land[(size_t)C::T * e * S] : mk(0.f, 0.f);
  }
  BBT_SYNC();  // the landing zone becomes the exchange buffer
  SmemLaneFast<C::PADSHIFT> sm{smem, g, C::G};
  block_fft<C>(v, t, tw, sm);
  if (valid) {
#pragma unroll
    for (int e = 0; e < C::E; ++e)
      v[e] = cconj(cmul(v[e], ldtw(chirp, t + C::T * e)));
  }
  block_fft_head<C>(v, t, tw, sm);
  if (tid == 0 && next_bytes) {
    fence_proxy_async();
    mbar_expect_tx(bar, next_bytes);
    char* dst = reinterpret_cast<char*>(smem);
    for (unsigned o = 0; o < next_bytes; o += kChunk)
      bulk_load(dst + o, next_src + o,
                next_bytes - o < kChunk ? next_bytes - o : kChunk, bar,
                o + kChunk >= next_bytes);
  }
  block_fft_tail<C>(v, t, tw, sm);
  if (valid) {
    cf* out = row + (size_t)t * S;
#pragma unroll
    for (int e = 0; e < C::E; ++e) out[(size_t)C::T * e * S] = v[e];
  }
}

template <class C>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS, 1) dd_rowi_tma_kernel(DdArgs a) {
  cf* smem = BBT_SMEM(cf);
  constexpr size_t kTile = (size_t)C::N * C::G;
  constexpr size_t kBuf = C::SMEM_BYTES / sizeof(cf) > kTile
                              ? C::SMEM_BYTES / sizeof(cf) : kTile;
  Mbar* bar = reinterpret_cast<Mbar*>(smem + kBuf);
  const unsigned n1 = (unsigned)(a.N >> a.log2n2);
  const int S = (int)a.S;
  const unsigned rpc = C::G / S;                     // rows per tile
  const unsigned tiles_per_frame = (n1 + rpc - 1) / rpc;
  const unsigned n_tiles = tiles_per_frame * (unsigned)a.n_frames;
  const int tid = threadIdx.x;
  const int g = tid % C::G;
  const unsigned kl = g / S, sl = g - kl * S;
  constexpr unsigned kChunk = 32 * 1024;
  auto tile_base = [&](unsigned lin, unsigned& k10) -> cf* {
    const unsigned xblk = lin / (unsigned)a.n_frames;
    const unsigned frame = lin - xblk * (unsigned)a.n_frames;
    k10 = xblk * rpc;
    return a.work + (long long)frame * a.N * a.S +
           (long long)k10 * C::N * a.S;
  };
  auto tile_bytes = [&](unsigned k10) -> unsigned {
    const unsigned left = n1 - k10;
    return (left < rpc ? left : rpc) * C::N * (unsigned)S *
           (unsigned)sizeof(cf);
  };
  if (tid == 0) mbar_init(bar, 1);
  BBT_SYNC();
  if (tid == 0 && blockIdx.x < n_tiles) {
    unsigned k10;
    const char* src = reinterpret_cast<const char*>(tile_base(blockIdx.x, k10));
    const unsigned bytes = tile_bytes(k10);
    mbar_expect_tx(bar, bytes);
    char* dst = reinterpret_cast<char*>(smem);
    for (unsigned o = 0; o < bytes; o += kChunk)
      bulk_load(dst + o, src + o, bytes - o < kChunk ? bytes - o : kChunk, bar,
                o + kChunk >= bytes);
  }
  unsigned k = 0;
#pragma unroll 1
  for (unsigned lin = blockIdx.x; lin < n_tiles; lin += gridDim.x, ++k) {
    unsigned k10;
    cf* base = tile_base(lin, k10);
    const unsigned k1 = k10 + kl;
    const bool valid = k1 < n1;
    cf* row = base + (long long)kl * C::N * a.S + sl;
    const cf* chirp = a.chirp;
    if (valid) chirp += ((long long)a.series_map[sl] * n1 + k1) * C::N;
    const unsigned next = lin + gridDim.x;
    const char* next_src = nullptr;
    unsigned next_bytes = 0;
    if (next < n_tiles) {
      unsigned r0;
      next_src = reinterpret_cast<const char*>(tile_base(next, r0));
      next_bytes = tile_bytes(r0);
      if (tid == 32)
        for (unsigned o = 0; o < next_bytes; o += kChunk)
          bulk_prefetch_l2(next_src + o,
                           next_bytes - o < kChunk ? next_bytes - o : kChunk);
    }
    dd_rowi_tma_tile<C>(smem, bar, row, chirp, a.tw, S, valid, k, next_src,
                        next_bytes);
  }
}

// Pass 2, second formulation: the N2-point transforms are split as 32 x M
// (M = N2/32) so that only ONE exchange per transform needs the whole CTA.
//   forward:  radix-32 butterflies over e of x[u + M e] in the thread that
//             loaded them, x W_N2^{u k1}, transpose through shared memory,
//             then 32 independent M-point transforms over u, each done by
//             M/32 threads of ONE warp (exchange with __syncwarp only);
//             bin k1 + 32 k2 ends up in the sub-transform k1, position k2.
//   multiply by the chirp (stored in that order, see row2_pos), conjugate;
//   mirror:   the M-point transforms over k2, x W_N2^{k1 u}, transpose back,
//             radix-32 butterflies over k1 in the thread that stores
//             out[u + M e] -- natural order, coalesced.
// Between the two transposes every warp runs two M-point transforms, the
// chirp multiply and the twiddles without waiting for any other warp, so the
// butterflies of some warps overlap the exchanges and chirp loads of others
// (in dd_row_kernel every exchange is a CTA-wide barrier pair and the FP32
// pipe idles while shared memory is busy, and vice versa).  Data movement is
// that of dd_row_tma_kernel: persistent CTAs, bulk copies into the exchange
// buffer, results stored from registers.
//
// Position within a stored chirp row of the bin kk = q1 + 32 (tt + Ts e) of
// that row (Ts = M/32 threads per sub-transform): the thread with index
// q1 Ts + tt in its row reads its e-th value at e M + q1 Ts + tt.
BBT_HD long long row2_pos(long long kk, int log2n2) {
  const long long M = 1LL << (log2n2 - 5), Ts = M >> 5 ? M >> 5 : 1;
  const long long q1 = kk & 31, q2 = kk >> 5;
  if (M < 32) return kk;  // not used below 1024 points
  return (q2 / Ts) * M + q1 * Ts + (q2 % Ts);
}
BBT_HD long long row2_bin(long long pos, int log2n2) {  // inverse of row2_pos
  const long long M = 1LL << (log2n2 - 5), Ts = M >> 5 ? M >> 5 : 1;
  if (M < 32) return pos;
  const long long e = pos / M, rem = pos % M;
  return (rem / Ts) + 32 * ((rem % Ts) + Ts * e);
}

// One chirp value from float64 phase (the arithmetic of chirp_kernel with the
// final sincos in float32 of the phase reduced to [-1/2, 1/2] cycles).
BBT_DEV cf chirp_value(double freq, double fref, double sb, double d,
                       double rate_mhz, double soff, long long k, long long N) {
  const long long ks = (k < (N + 1) / 2) ? k : k - N;
  const double fftfreq = (double)ks * (rate_mhz / (double)N);
  const double f = freq + fftfreq * sb;
  const double u = 1. / fref - 1. / f;
  double phase = d * f * (u * u) * 1e6 * sb;
  if (soff != 0.) phase += soff / rate_mhz * fftfreq;
  phase -= rint(phase);
  float sn, cs;
#if defined(__CUDA_ARCH__)
  sincospif(2.f * (float)phase, &sn, &cs);
#else
  sn = sinf(6.283185307179586f * (float)phase);
  cs = cosf(6.283185307179586f * (float)phase);
#endif
  return mk(cs, sn);
}

template <class C>
struct Row2Cfg {
  static constexpr int M = C::T;                    // points per sub-transform
  static constexpr int Ts = M / 32;                 // its threads
  using CS = FftCfg<C::LOG2N - 5, 5, Ts>;           // the M-point transform
  static constexpr int P = CS::NPAD > M ? CS::NPAD : M;  // pitch of a row of X
  static constexpr size_t kElems =
      (size_t)C::G * 32 * P > C::SMEM_BYTES / sizeof(cf)
          ? (size_t)C::G * 32 * P : C::SMEM_BYTES / sizeof(cf);
  // Behind the matrix: the barrier and the gathered twiddle tables (outer
  // twiddles W_N2^{u {1, 8, 16}}, u < M, then those of the sub-transform).
  static constexpr int kTabOuter = 3 * M;
  static constexpr size_t kSmemBytes =
      kElems * sizeof(cf) + 2 * sizeof(Mbar) +
      (kTabOuter + CS::twtab_size()) * sizeof(cf);
  // Side buffer (dd_row2_kernel with LAND = 2, one row per tile): the first
  // kSide of the 32 M-point pieces of the NEXT row land here while the
  // current row is still being worked on in X.
  static constexpr int kSide = 16;
  static constexpr size_t kSideOffset = (kSmemBytes + 127) / 128 * 128;
  static constexpr size_t kSmemBytesSide =
      kSideOffset + (size_t)kSide * M * sizeof(cf);
};

struct ChirpRegen {          // per row: what chirp_value needs
  double freq, fref, sb, d, rate_mhz, soff;
  long long k1, n1, N;
  int log2n2;
};

#ifndef BBT_ROW_FLAGS
// dd_row2_kernel: bit 0, groups of warps take turns at shared memory after a
// barrier (stagger_in); bit 1, the barrier between the landing zone's reads
// and the first exchange comes after the butterflies; bit 2, only the warp
// that issues the next tile's copy waits for the last reads of a tile.
#define BBT_ROW_FLAGS 1
#endif
// After a CTA-wide barrier every warp wants shared memory (the reads of an
// exchange) and then the FP32 pipe (its butterflies): with fair scheduling the
// warps move through these phases together, and the pipe idles while shared
// memory is busy and vice versa.  Here the CTA's warps form four groups (warp
// w in group w / 4 of the n = THREADS / 128, so that every scheduler has one
// warp of each) which issue their loads one group after the other: the first
// group is at its butterflies while the others still load, and the skew stays
// for the phases that follow.  Barriers 1.. chain the groups.
BBT_DEV void stagger_in(int grp, int flags) {
  if ((flags & 1) && grp > 0) named_bar_sync(grp, 256);
}
BBT_DEV void stagger_out(int grp, int n_grp, int flags) {
  if ((flags & 1) && grp + 1 < n_grp) named_bar_arrive(grp + 1, 256);
}

// The copies that bring a tile of rows into shared memory.  LANDP: every
// M-point piece x[M e .. M e + M) of a row lands at the pitch P of the
// exchange matrix X, so that the thread which takes x[u + M e] out of the
// landing zone later writes its butterflies' results to the very same places
// (X[e P + u]) and no barrier is needed between the two; otherwise the tile
// lands as it lies in memory, in chunks of 32 KB.
// LAND = 2 (one row per tile): the first kSide pieces go to a side buffer --
// issued as soon as the previous row has been taken out of it, a whole row
// time before they are needed -- and only the rest to X once X is free.
template <class C>
BBT_DEV void row2_load_side(cf* smem, const char* src, Mbar* bar_side) {
  using R = Row2Cfg<C>;
  constexpr unsigned kPiece = R::M * sizeof(cf);
  char* dst = reinterpret_cast<char*>(smem) + R::kSideOffset;
  mbar_expect_tx(bar_side, R::kSide * kPiece);
  for (unsigned i = 0; i < (unsigned)R::kSide; ++i)
    bulk_load(dst + (size_t)i * kPiece, src + (size_t)i * kPiece, kPiece,
              bar_side, i + 1 == (unsigned)R::kSide);
}
template <class C>
BBT_DEV void row2_load_rest(cf* smem, const char* src, Mbar* bar) {
  using R = Row2Cfg<C>;
  constexpr unsigned kPiece = R::M * sizeof(cf);
  constexpr unsigned kPitch = R::P * sizeof(cf);
  char* dst = reinterpret_cast<char*>(smem);
  mbar_expect_tx(bar, (32 - R::kSide) * kPiece);
  for (unsigned i = R::kSide; i < 32; ++i)
    bulk_load(dst + (size_t)i * kPitch, src + (size_t)i * kPiece, kPiece, bar,
              i == 31);
}

template <class C, int LANDP>
BBT_DEV void row2_load(cf* smem, const char* src, unsigned bytes, Mbar* bar) {
  char* dst = reinterpret_cast<char*>(smem);
  if constexpr (LANDP != 0) {
    constexpr unsigned kPiece = Row2Cfg<C>::M * sizeof(cf);
    constexpr unsigned kPitch = Row2Cfg<C>::P * sizeof(cf);
    const unsigned n = bytes / kPiece;
    for (unsigned i = 0; i < n; ++i)
      bulk_load(dst + (size_t)i * kPitch, src + (size_t)i * kPiece, kPiece, bar,
                i + 1 == n);
  } else {
    constexpr unsigned kChunk = 32 * 1024;
    for (unsigned o = 0; o < bytes; o += kChunk)
      bulk_load(dst + o, src + o, bytes - o < kChunk ? bytes - o : kChunk, bar,
                o + kChunk >= bytes);
  }
}

template <class C, bool REGEN, int LANDP>
BBT_DEV_NOINLINE void dd_row2_tile(
    cf* smem, Mbar* bar, cf* row, const cf* chirp, const cf* tw,
    const cf* tw_sub, bool valid, unsigned phase, const char* next_src,
    unsigned next_bytes, const ChirpRegen* rg) {
  static_assert(C::LOG2E == 5 && C::LOG2N >= 10, "32 values per thread");
  constexpr int flags = BBT_ROW_FLAGS;
  static_assert(C::THREADS % 128 == 0, "groups of four warps");
  constexpr int kGroups = C::THREADS / 128;
  const int grp = threadIdx.x / 128;
  using R = Row2Cfg<C>;
  using CS = typename R::CS;
  constexpr int M = R::M, Ts = R::Ts, P = R::P;
  static_assert(CS::T == Ts && CS::G == 1 && CS::N == M && 32 % Ts == 0,
                "a sub-transform sits in one warp");
  const int tid = threadIdx.x;
  const int u = tid % C::T, g = tid / C::T;
  const int k1 = u / Ts, tt = u % Ts;        // sub-transform and place in it
  cf* X = smem + (size_t)g * 32 * P;         // this row's [32][P] matrix
  const cf* tab_outer = reinterpret_cast<const cf*>(bar + 2);
  const cf* tab_sub = tab_outer + R::kTabOuter;
  cf v[32];
  if constexpr (LANDP == 2) {
    static_assert(LANDP != 2 || C::G == 1, "side buffer: one row per tile");
    const cf* side = reinterpret_cast<const cf*>(
        reinterpret_cast<const char*>(smem) + R::kSideOffset) + u;
    mbar_wait(bar + 1, phase & 1u, phase);
    stagger_in(grp, flags);
#pragma unroll
    for (int e = 0; e < R::kSide; ++e) v[e] = side[M * e];
    mbar_wait(bar, phase & 1u, phase);
#pragma unroll
    for (int e = R::kSide; e < 32; ++e) v[e] = X[u + P * e];
    stagger_out(grp, kGroups, flags);
  } else {
    mbar_wait(bar, phase & 1u, phase);
    stagger_in(grp, flags);
    const cf* land = LANDP ? X + u : smem + g * C::N + u;
    constexpr int pitch = LANDP ? P : M;
#pragma unroll
    for (int e = 0; e < 32; ++e) v[e] = valid ? land[pitch * e] : mk(0.f, 0.f);
    stagger_out(grp, kGroups, flags);
  }
  // The landing zone becomes the exchange buffer once every thread has taken
  // its values (with LANDP a thread overwrites only what it took itself); the
  // butterflies in between need registers only.
  if (!LANDP && !(flags & 2)) BBT_SYNC();
  // Forward: over e in this thread, twiddle, transpose.
  Dft<32>::run(v);
  apply_twiddles_tab<32>(v, tab_outer, M, u);
  if (!LANDP && (flags & 2)) BBT_SYNC();
#pragma unroll
  for (int r = 0; r < 32; ++r) X[r * P + u] = v[r];
  BBT_SYNC();
  if constexpr (LANDP == 2) {
    // Every thread has taken this row out of the side buffer: the next
    // row's first pieces may land there.
    if (tid == C::THREADS - 32 && next_bytes) {
      fence_proxy_async();
      row2_load_side<C>(smem, next_src, bar + 1);
    }
  }
  // From here to the next barrier warps do not wait for one another.
  cf* mine = X + k1 * P;
  stagger_in(grp, flags);
#pragma unroll
  for (int e = 0; e < 32; ++e) v[e] = mine[tt + Ts * e];
  stagger_out(grp, kGroups, flags);
  BBT_SYNCWARP();
  SmemWarp<CS::PADSHIFT> sw{mine, tab_sub};
  block_fft<CS>(v, tt, tw_sub, sw);
  if (valid) {
    if constexpr (REGEN) {
      // The chirp from float64 phase instead of the cached table: bin
      // k1 + n1 kk of the frame, kk the bin of the row this thread holds.
#pragma unroll
      for (int e = 0; e < 32; ++e) {
        const long long kk = row2_bin(u + M * e, rg->log2n2);
        v[e] = cconj(cmul(v[e], chirp_value(rg->freq, rg->fref, rg->sb, rg->d,
                                            rg->rate_mhz, rg->soff,
                                            rg->k1 + rg->n1 * kk, rg->N)));
      }
    } else {
#pragma unroll
      for (int e = 0; e < 32; ++e)
        v[e] = cconj(cmul(v[e], ldtw(chirp, u + M * e)));
    }
  }
  block_fft<CS>(v, tt, tw_sub, sw);
  {
    // x W_N2^{k1 (tt + Ts e)}: a ramp in e from a few table look-ups.
    cf pw[5];
#pragma unroll
    for (int b = 0; b < 5; ++b) pw[b] = ldtw(tw, (k1 * Ts) << b);
    Ramp<5, 0>::run(v, ldtw(tw, k1 * tt), pw);
  }
  BBT_SYNCWARP();  // the sub-transform's last reads of this row of X
#pragma unroll
  for (int e = 0; e < 32; ++e) mine[tt + Ts * e] = v[e];
  BBT_SYNC();
  stagger_in(grp, flags);
#pragma unroll
  for (int r = 0; r < 32; ++r) v[r] = X[r * P + u];
  stagger_out(grp, kGroups, flags);
  // X is free once every thread has read: the next tile may land while we
  // finish.  Only the warp that issues the copy (the last one, whose group
  // reads last) has to wait for that; the others go on to their butterflies.
#if defined(BBT_EMULATE)
  BBT_SYNC();
#else
  if constexpr ((flags & 4) != 0) {
    if (tid >= C::THREADS - 32)
      named_bar_sync(8, C::THREADS);
    else
      named_bar_arrive(8, C::THREADS);
  } else {
    BBT_SYNC();
  }
#endif
  if (tid == ((flags & 4) ? C::THREADS - 32 : 0) && next_bytes) {
    fence_proxy_async();
    if constexpr (LANDP == 2) {
      row2_load_rest<C>(smem, next_src, bar);
    } else {
      mbar_expect_tx(bar, next_bytes);
      row2_load<C, LANDP>(smem, next_src, next_bytes, bar);
    }
  }
  Dft<32>::run(v);
  if (valid) {
#pragma unroll
    for (int e = 0; e < 32; ++e) row[u + M * e] = v[e];
  }
}

template <class C, bool REGEN, int LANDP = 0>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS, (C::THREADS <= 256 ? 2 : 1))
    dd_row2_kernel(DdArgs a) {
  cf* smem = BBT_SMEM(cf);
  Mbar* bar = reinterpret_cast<Mbar*>(smem + Row2Cfg<C>::kElems);
  {
    // Twiddle tables, gathered once per CTA (see apply_twiddles_tab).
    using R = Row2Cfg<C>;
    cf* tab_outer = reinterpret_cast<cf*>(bar + 2);
    for (int i = threadIdx.x; i < R::kTabOuter; i += C::THREADS) {
      const int which = i / R::M, u = i - which * R::M;
      tab_outer[i] = a.tw[u * (which == 0 ? 1 : (which == 1 ? 8 : 16))];
    }
    fill_twtab<typename R::CS>(tab_outer + R::kTabOuter, a.tw_sub,
                               threadIdx.x, C::THREADS);
  }
  const unsigned n1 = (unsigned)(a.N >> a.log2n2);
  const unsigned S = (unsigned)a.S;
  const unsigned rows = n1 * S;
  const unsigned tiles_per_frame = (rows + C::G - 1) / C::G;
  const unsigned n_tiles = tiles_per_frame * (unsigned)a.n_frames;
  const int tid = threadIdx.x;
  const int t = tid % C::T, g = tid / C::T;
  constexpr unsigned kChunk = 32 * 1024;
  auto tile_base = [&](unsigned lin, unsigned& rho0) -> cf* {
    const unsigned xblk = lin / (unsigned)a.n_frames;
    const unsigned frame = lin - xblk * (unsigned)a.n_frames;
    rho0 = xblk * C::G;
    return a.work + (long long)frame * a.N * a.S + (long long)rho0 * C::N;
  };
  auto tile_bytes = [&](unsigned rho0) -> unsigned {
    const unsigned left = rows - rho0;
    return (left < (unsigned)C::G ? left : (unsigned)C::G) * C::N *
           (unsigned)sizeof(cf);
  };
  // (A fixed tile-to-CTA map: rows are contiguous, so unlike the column
  // passes nothing is gained by handing the tiles out in order -- measured
  // 4 % slower, for the extra barrier.)
  if (tid == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
  }
  BBT_SYNC();
  if (tid == 0 && blockIdx.x < n_tiles) {
    unsigned rho0;
    const char* src = reinterpret_cast<const char*>(tile_base(blockIdx.x, rho0));
    const unsigned bytes = tile_bytes(rho0);
    if constexpr (LANDP == 2) {
      row2_load_side<C>(smem, src, bar + 1);
      row2_load_rest<C>(smem, src, bar);
    } else {
      mbar_expect_tx(bar, bytes);
      row2_load<C, LANDP>(smem, src, bytes, bar);
    }
  }
  unsigned k = 0;
#pragma unroll 1
  for (unsigned lin = blockIdx.x; lin < n_tiles; lin += gridDim.x, ++k) {
    const unsigned next = lin + gridDim.x;
    unsigned rho0;
    cf* row = tile_base(lin, rho0) + (long long)g * C::N;
    const unsigned rho = rho0 + g;
    const bool valid = rho < rows;
    const unsigned k1 = rho / S, s = rho - k1 * S;
    const cf* chirp = a.chirp;
    if (valid) chirp += ((long long)a.series_map[s] * n1 + k1) * C::N;
    if (valid && t == 0)
      for (unsigned o = 0; o < C::N * sizeof(cf); o += kChunk)
        bulk_prefetch_l2(reinterpret_cast<const char*>(chirp) + o,
                         C::N * sizeof(cf) - o < kChunk
                             ? (unsigned)(C::N * sizeof(cf)) - o : kChunk);
    const char* next_src = nullptr;
    unsigned next_bytes = 0;
    if (next < n_tiles) {
      unsigned r0;
      next_src = reinterpret_cast<const char*>(tile_base(next, r0));
      next_bytes = tile_bytes(r0);
      if (tid == 32)
        for (unsigned o = 0; o < next_bytes; o += kChunk)
          bulk_prefetch_l2(next_src + o,
                           next_bytes - o < kChunk ? next_bytes - o : kChunk);
    }
    ChirpRegen rg;
    if (REGEN && valid) {
      const int c = a.series_map[s];
      rg.freq = a.ch_freq[c], rg.fref = a.ch_fref[c];
      rg.sb = (double)a.ch_sb[c], rg.d = a.ch_d;
      rg.rate_mhz = a.ch_rate, rg.soff = a.ch_soff;
      rg.k1 = k1, rg.n1 = n1, rg.N = a.N, rg.log2n2 = a.log2n2;
    }
    dd_row2_tile<C, REGEN, LANDP>(smem, bar, row, chirp, a.tw, a.tw_sub, valid,
                                  k, next_src, next_bytes, &rg);
  }
}

// Passes 1 and 3 as persistent kernels fed by tensor-map bulk copies (TMA).
// A column tile is N1 rows of G lanes (128 bytes), rows N2*S*8 bytes apart: one
// cp.async.bulk.tensor box of [<=256 rows][G lanes] per quarter of the tile,
// which lands in shared memory already in the lane-fastest layout the block
// FFT exchanges in.  As in dd_row_tma_kernel the exchange buffer doubles as
// the landing zone: the copy of the next tile is issued right after the last
// exchange, and overlaps the last butterfly stage, the twiddle ramp and the
// stores.  For the planar work buffer (pass 3) the box is [rows][S][G/S]: the
// de-interleaving gather is done by the copy engine, lane g of the tile being
// series g / (G/S), column n2_0 + g % (G/S).
// exp(-2 pi i m / N) for an exact integer m, rounded once from float64.
BBT_DEV cf unit_root(long long m, long long N) {
  double sn, cs;
  sincospi_d(-2. * (double)(m & (N - 1)) / (double)N, &sn, &cs);
  return mk((float)cs, (float)sn);
}

struct DdColTma {
  unsigned* next_tile;  // device counter (zeroed before the launch): tiles are
                        // handed out in order, so the tiles in flight stay
                        // neighbours in memory however the CTAs drift apart
                        // (with a fixed tile-to-CTA map the column passes
                        // lost 15-25 % of their speed on long launches)
  int fast_tw;      // twiddle ramps from per-CTA tables (needs G % S == 0)
  int n_boxes;      // copies per tile
  int box_rows;     // rows per copy
  int tn;           // planar source: n2 values per series in a tile (G / S)
  int stagger_ns;   // delay of every other CTA, so that the CTAs sharing an SM
                    // run out of phase (one in its butterflies while the
                    // other exchanges and stores)
};

template <class C, bool INVERSE, bool DETECT = false>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS, (C::THREADS <= 256 ? 2 : 1))
    dd_col_tma_kernel(DdArgs a, DdColTma m, BBT_TMAP_PARAM map) {
  static_assert(INVERSE || !DETECT, "products are formed by the last pass");
  cf* smem = BBT_SMEM(cf);
  // The barrier sits behind the exchange buffer (or the landing zone, for
  // transforms without an exchange).
  constexpr size_t kTile = (size_t)C::N * C::G;          // elements
  constexpr size_t kBuf = C::SMEM_BYTES / sizeof(cf) > kTile
                              ? C::SMEM_BYTES / sizeof(cf) : kTile;
  Mbar* bar = reinterpret_cast<Mbar*>(smem + kBuf);
  // Twiddle tables behind the barrier: the ramp of column n2 = n2_0 + j is
  //   W_N^{n2 (t + T e)} = [W^{n2_0 t} W^{j t}] [W^{n2_0 T} W^{j T}]^e,
  // whose second factors do not depend on the tile (W^{j t}: one value per
  // thread, kept in registers; W^{j T 2^b}: tab_d[b][g]) and whose first
  // factors are the same for all lanes (tab_a[t], tab_c[b]: computed per tile
  // by a few threads, one tile ahead).  All are rounded once from float64,
  // so the threads do no divergent table look-ups at all.
  constexpr int LE = C::LOG2E > 0 ? C::LOG2E : 1;
  cf* tab_d = reinterpret_cast<cf*>(bar + 2);       // [LE][G]
  cf* tab_a = tab_d + LE * C::G;                    // [2][T]
  cf* tab_c = tab_a + 2 * C::T;                     // [2][8]
  unsigned* tile_slot = reinterpret_cast<unsigned*>(tab_c + 16);  // [2]
  const unsigned N2 = (unsigned)(a.N >> a.log2n1);
  const unsigned S = (unsigned)a.S;
  const unsigned n2s = N2 * S;
  const unsigned nblk = (n2s + C::G - 1) / C::G;
  const unsigned n_tiles = nblk * (unsigned)a.n_frames;
  const int tid = threadIdx.x;
  const int g = tid % C::G, t = tid / C::G;
  const bool planar_src = INVERSE && a.planar;
  // Column offset of lane g within its tile.
  auto lane_j = [&](int lane) -> long long {
    return planar_src ? lane % m.tn : lane / (int)S;
  };
  auto tile_tables = [&](unsigned lin, unsigned slot) {   // threads < T + LE
    const unsigned frame = lin / nblk;
    const long long n20 = (long long)((lin - frame * nblk) * C::G) / S;
    if (tid < C::T)
      tab_a[slot * C::T + tid] = unit_root(n20 * tid, a.N);
    else if (tid < C::T + LE)
      tab_c[slot * 8 + (tid - C::T)] =
          unit_root((n20 * C::T) << (tid - C::T), a.N);
  };
  cf w_jt = mk(1.f, 0.f);
  if (m.fast_tw) {
    w_jt = unit_root(lane_j(g) * t, a.N);
    for (int i = tid; i < LE * C::G; i += C::THREADS)
      tab_d[i] = unit_root((lane_j(i % C::G) * C::T) << (i / C::G), a.N);
  }
  auto issue = [&](unsigned lin) {                  // one thread
    const unsigned frame = lin / nblk, c0 = (lin - frame * nblk) * C::G;
    fence_proxy_async();
    mbar_expect_tx(bar, (unsigned)(kTile * sizeof(cf)));
    for (int q = 0; q < m.n_boxes; ++q) {
      cf* dst = smem + (size_t)q * m.box_rows * C::G;
      const bool last = q == m.n_boxes - 1;
      if (planar_src)
        tensor_load_3d(dst, &map, (int)(c0 / S), 0,
                       (int)(frame * C::N + q * m.box_rows), bar, last);
      else
        tensor_load_3d(dst, &map, (int)c0, q * m.box_rows, (int)frame, bar,
                       last);
    }
  };
  auto prefetch = [&](unsigned lin) {
    const unsigned frame = lin / nblk, c0 = (lin - frame * nblk) * C::G;
    for (int q = 0; q < m.n_boxes; ++q) {
      if (planar_src)
        tensor_prefetch_3d(&map, (int)(c0 / S), 0,
                           (int)(frame * C::N + q * m.box_rows));
      else
        tensor_prefetch_3d(&map, (int)c0, q * m.box_rows, (int)frame);
    }
  };
  if (tid == 0) {
    mbar_init(bar, 1);
    tile_slot[0] = atomic_add(m.next_tile, 1u);
  }
  BBT_SYNC();
  if (m.fast_tw && tile_slot[0] < n_tiles) tile_tables(tile_slot[0], 0);
  if (tid == 0 && tile_slot[0] < n_tiles) issue(tile_slot[0]);
  unsigned k = 0;
#pragma unroll 1
  for (;; ++k) {
    const unsigned lin = tile_slot[k & 1u];
    if (lin >= n_tiles) break;
    // Thread 0 draws the next tile now and passes it on through shared
    // memory (read after the next barrier).
    unsigned next = 0;
    if (tid == 0) {
      next = atomic_add(m.next_tile, 1u);
      tile_slot[(k + 1) & 1u] = next;
      if (next < n_tiles) prefetch(next);
    }
    const unsigned frame = lin / nblk, c0 = (lin - frame * nblk) * C::G;
    // Flat column n2*S + s of this lane.
    unsigned n2, sser;
    if (planar_src) {
      sser = (unsigned)g / (unsigned)m.tn;
      n2 = c0 / S + ((unsigned)g - sser * (unsigned)m.tn);
    } else {
      const unsigned col = c0 + g;
      n2 = col / S;
      sser = col - n2 * S;
    }
    const bool valid = n2 < N2 && sser < S;
    mbar_wait(bar, k & 1u, k);
    cf v[C::E];
    {
      const cf* land = smem + (size_t)t * C::G + g;
#pragma unroll
      for (int e = 0; e < C::E; ++e) v[e] = land[(size_t)e * C::T * C::G];
    }
    BBT_SYNC();  // the landing zone becomes the exchange buffer
    if (m.fast_tw && tile_slot[(k + 1) & 1u] < n_tiles)
      tile_tables(tile_slot[(k + 1) & 1u], (k + 1) & 1u);
    auto twiddle = [&](float scale) {
      if (m.fast_tw) {
        cf pw[LE];
#pragma unroll
        for (int b = 0; b < C::LOG2E; ++b)
          pw[b] = cmul(tab_c[(k & 1u) * 8 + b], tab_d[b * C::G + g]);
        const cf base =
            cscale(cmul(tab_a[(k & 1u) * C::T + t], w_jt), scale);
        Ramp<C::LOG2E, 0>::run(v, base, pw);
      } else {
        col_twiddle<C, 0>(v, a.big, (int)n2, t, scale);
      }
    };
    if (INVERSE && valid) twiddle(a.scale);
    SmemLaneFast<C::PADSHIFT> sm{smem, g, C::G};
    block_fft_head<C>(v, t, a.tw1, sm);
    if (tid == 0 && next < n_tiles) issue(next);
    block_fft_tail<C>(v, t, a.tw1, sm);
    // (With the products fused in, every lane takes part in the exchange
    // between the lanes of a polarization pair.)
    if (!valid && !DETECT) continue;
    if (!INVERSE) {
      twiddle(1.f);
      cf* dst = a.work + (long long)frame * a.N * a.S;
      long long step;
      if (a.planar) {
        dst += (long long)sser * N2 + n2;
        step = (long long)a.S * N2;
      } else {
        dst += (long long)n2 * S + sser;
        step = n2s;
      }
      dst += (long long)t * step;
      const long long pstep = (long long)C::T * step;
#pragma unroll
      for (int e = 0; e < C::E; ++e) {
        *dst = v[e];
        dst += pstep;
      }
    } else {
      const long long col = (long long)n2 * S + sser;
      long long flat = (long long)t * n2s + col;
      const long long fstep = (long long)C::T * n2s;
      cf* dst = a.out + (long long)frame * a.out_frame_stride - a.out_shift +
                flat;
      if constexpr (DETECT) {
        // The partner lane holds the other polarization of this sample: the
        // first lane of a pair stores the two powers, the second the cross
        // terms, each in the 8 bytes its voltage would have gone to.
        const int d = planar_src ? m.tn : 1;
        const bool second = (sser & 1u) != 0;
#pragma unroll
        for (int e = 0; e < C::E; ++e) {
          const cf r = detect_pair(cconj(v[e]), second, d);
          if (valid && flat >= a.lo && flat < a.hi) *dst = r;
          flat += fstep;
          dst += fstep;
        }
      } else {
#pragma unroll
        for (int e = 0; e < C::E; ++e) {
          if (flat >= a.lo && flat < a.hi) *dst = cconj(v[e]);
          flat += fstep;
          dst += fstep;
        }
      }
    }
  }
}

// Single pass for N <= 16384: lanes are (frame, series) pairs, series fastest.
// LANEFAST (S > 1): consecutive threads take consecutive series; otherwise
// consecutive threads walk along time.
template <class C, bool LANEFAST>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS, (C::THREADS <= 256 ? 2 : 1))
    dd_small_kernel(DdArgs a, long long n_frames) {
  cf* smem = BBT_SMEM(cf);
  const int tid = threadIdx.x;
  const int g = LANEFAST ? tid % C::G : tid / C::T;
  const int t = LANEFAST ? tid / C::G : tid % C::T;
  const long long lane = (long long)blockIdx.x * C::G + g;
  const long long frame = lane / a.S, s = lane % a.S;
  const bool valid = lane < n_frames * a.S;
  const cf* src = a.in + frame * a.in_frame_stride + s;
  cf* dst = a.out + frame * a.out_frame_stride - a.out_shift + s;
  const cf* chirp = a.chirp;
  if (valid) chirp += (long long)a.series_map[s] * C::N;
  cf v[C::E];
#pragma unroll
  for (int e = 0; e < C::E; ++e)
    v[e] = valid ? src[(long long)(t + C::T * e) * a.S] : mk(0.f, 0.f);
  lane_fft<C, LANEFAST>(v, t, g, a.tw, smem);
  if (valid) {
#pragma unroll
    for (int e = 0; e < C::E; ++e)
      v[e] = cconj(cmul(v[e], ldtw(chirp, t + C::T * e)));
  }
  lane_fft<C, LANEFAST>(v, t, g, a.tw, smem);
  if (valid) {
#pragma unroll
    for (int e = 0; e < C::E; ++e) {
      const long long flat = (long long)(t + C::T * e) * a.S + s;
      if (flat >= a.lo && flat < a.hi)
        dst[flat - s] = cscale(cconj(v[e]), a.scale);
    }
  }
}

// Chirp table, computed in float64 and rounded to complex64 exactly like
// Disperse.phase_factor (dispersion.py:115-129 with dm.py:103-105):
//   F     = f + fftfreq * sideband                      [MHz]
//   phase = K DM F (1/f_ref - 1/F)^2 1e6 * sideband     [cycles]
//         + sample_offset / rate * fftfreq
//   chirp = exp(2 pi i phase)
// stored at [c][k1][k2] for bin k = k1 + N1*k2.
struct ChirpArgs {
  int row2_log2n2;          // > 0: rows stored in dd_row2_kernel's order
  cf* chirp;
  const double* freq_mhz;   // [n_chirp]
  const double* fref_mhz;   // [n_chirp]
  const signed char* sideband;  // [n_chirp]
  long long N, n1, n_chirp;
  double d;                 // K * DM (signed), s MHz^2
  double rate_mhz;
  double sample_offset;
};

BBT_GLOBAL void chirp_kernel(ChirpArgs a) {
  const long long n2 = a.N / a.n1;
  const long long total = a.n_chirp * a.N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long c = i / a.N, r = i % a.N;
    const long long k1 = r / n2;
    long long k2 = r % n2;
    if (a.row2_log2n2 > 0) k2 = row2_bin(k2, a.row2_log2n2);
    const long long k = k1 + a.n1 * k2;
    const long long ks = (k < (a.N + 1) / 2) ? k : k - a.N;  // np.fft.fftfreq
    const double fftfreq = (double)ks * (a.rate_mhz / (double)a.N);
    const double sb = (double)a.sideband[c];
    const double f = a.freq_mhz[c] + fftfreq * sb;
    const double u = 1. / a.fref_mhz[c] - 1. / f;
    double phase = a.d * f * (u * u) * 1e6 * sb;
    if (a.sample_offset != 0.) phase += a.sample_offset / a.rate_mhz * fftfreq;
    phase -= rint(phase);
    double sn, cs;
    sincospi_d(2. * phase, &sn, &cs);
    a.chirp[i] = mk((float)cs, (float)sn);
  }
}

// Scatter a caller-supplied response (natural bin order, [n_chirp][N]) into
// the [c][k1][k2] layout; lets Convolve-style tasks reuse the plan.
BBT_GLOBAL void chirp_scatter_kernel(cf* dst, const cf* src, long long N,
                                     long long n1, long long n_chirp,
                                     int row2_log2n2) {
  const long long n2 = N / n1;
  const long long total = n_chirp * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long c = i / N, r = i % N;
    long long k2 = r % n2;
    if (row2_log2n2 > 0) k2 = row2_bin(k2, row2_log2n2);
    dst[i] = src[c * N + (r / n2) + n1 * k2];
  }
}

}  // namespace bbt

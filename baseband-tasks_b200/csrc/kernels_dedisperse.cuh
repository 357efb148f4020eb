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
// For N <= 8192 a single kernel (dd_small) does everything in one round trip.
#pragma once
#include "kernels_fft.cuh"

namespace bbt {

struct DdArgs {
  const cf* in;         // first frame; frame f starts in_frame_stride later
  cf* out;              // valid output of frame f at out + f*out_frame_stride
  cf* work;             // n_frames * N * S scratch
  const cf* tw;         // kTwiddleTable roots of unity
  BigTwiddle big;       // W_N^m
  const cf* chirp;      // [n_chirp][N1][N2]
  const int* series_map;  // series -> chirp index
  long long in_frame_stride, out_frame_stride;  // in complex elements
  long long N, S;       // frame length, interleaved series
  int log2n1, log2n2;
  long long lo, hi;     // valid flat range [(pad_start+skip)*S, (pad_start+spf)*S)
  long long out_shift;  // pad_start*S
  float scale;          // 1/N
};

// Pass 1: forward column FFTs, frame -> work.
template <int LOG2N1>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(FftCfg<LOG2N1>::THREADS, 1) dd_col_fwd_kernel(DdArgs a) {
  using C = FftCfg<LOG2N1>;
  cf* smem = BBT_SMEM(cf);
  const long long cols = a.N / C::N * a.S;  // N2*S columns
  const int tid = threadIdx.x;
  const int g = tid % C::G, t = tid / C::G;
  const long long col = (long long)blockIdx.x * C::G + g;
  const long long frame = blockIdx.y;
  const bool valid = col < cols;
  const cf* src = a.in + frame * a.in_frame_stride + col;
  cf* dst = a.work + frame * a.N * a.S + col;
  cf v[C::E];
#pragma unroll
  for (int e = 0; e < C::E; ++e)
    v[e] = valid ? src[(long long)(t + C::T * e) * cols] : mk(0.f, 0.f);
  SmemLaneFast sm{smem, g, C::G};
  block_fft<LOG2N1>(v, t, a.tw, sm);
  if (valid) {
#pragma unroll
    for (int e = 0; e < C::E; ++e) dst[(long long)(t + C::T * e) * cols] = v[e];
  }
}

// Pass 3: inverse column FFTs, work -> valid part of the output stream.
template <int LOG2N1>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(FftCfg<LOG2N1>::THREADS, 1) dd_col_inv_kernel(DdArgs a) {
  using C = FftCfg<LOG2N1>;
  cf* smem = BBT_SMEM(cf);
  const long long cols = a.N / C::N * a.S;
  const int tid = threadIdx.x;
  const int g = tid % C::G, t = tid / C::G;
  const long long col = (long long)blockIdx.x * C::G + g;
  const long long frame = blockIdx.y;
  const bool valid = col < cols;
  const cf* src = a.work + frame * a.N * a.S + col;
  cf* dst = a.out + frame * a.out_frame_stride - a.out_shift + col;
  cf v[C::E];
#pragma unroll
  for (int e = 0; e < C::E; ++e)
    v[e] = valid ? cconj(src[(long long)(t + C::T * e) * cols]) : mk(0.f, 0.f);
  SmemLaneFast sm{smem, g, C::G};
  block_fft<LOG2N1>(v, t, a.tw, sm);
  if (valid) {
#pragma unroll
    for (int e = 0; e < C::E; ++e) {
      const long long flat = (long long)(t + C::T * e) * cols + col;
      if (flat >= a.lo && flat < a.hi) dst[flat - col] = cconj(v[e]);
    }
  }
}

// Pass 2: one CTA per G rows of one series.  Lanes are rows (k1), threads of a
// lane walk along n2 with stride S.
template <int LOG2N2>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(FftCfg<LOG2N2>::THREADS, 1) dd_row_kernel(DdArgs a) {
  using C = FftCfg<LOG2N2>;
  cf* smem = BBT_SMEM(cf);
  const long long n1 = a.N >> LOG2N2;
  const int tid = threadIdx.x;
  const int t = tid % C::T, g = tid / C::T;
  const long long k1 = (long long)blockIdx.x * C::G + g;
  const long long s = blockIdx.y;
  const long long frame = blockIdx.z;
  const bool valid = k1 < n1;
  cf* row = a.work + frame * a.N * a.S + k1 * C::N * a.S + s;
  const cf* chirp = a.chirp + ((long long)a.series_map[s] * n1 + k1) * C::N;
  SmemLaneSlow<C::NPAD> sm{smem + (size_t)g * C::NPAD};
  cf v[C::E], w[C::E];
#pragma unroll
  for (int e = 0; e < C::E; ++e) {
    const int n2 = t + C::T * e;
    if (valid) {
      w[e] = a.big.get(k1 * n2);
      v[e] = cmul(row[(long long)n2 * a.S], w[e]);
    } else {
      w[e] = mk(1.f, 0.f);
      v[e] = mk(0.f, 0.f);
    }
  }
  block_fft<LOG2N2>(v, t, a.tw, sm);
  if (valid) {
#pragma unroll
    for (int e = 0; e < C::E; ++e)
      v[e] = cconj(cmul(v[e], ldtw(chirp, t + C::T * e)));
  }
  block_fft<LOG2N2>(v, t, a.tw, sm);
  if (valid) {
#pragma unroll
    for (int e = 0; e < C::E; ++e)
      // conj(fft(conj(Y))) * conj(w) / N = conj(fft(conj(Y)) * w) / N
      row[(long long)(t + C::T * e) * a.S] =
          cscale(cconj(cmul(v[e], w[e])), a.scale);
  }
}

// Single pass for N <= 8192: lanes are (frame, series) pairs, series fastest.
// LANEFAST (S > 1): consecutive threads take consecutive series; otherwise
// consecutive threads walk along time.
template <int LOG2N, bool LANEFAST>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(FftCfg<LOG2N>::THREADS, 1) dd_small_kernel(DdArgs a, long long n_frames) {
  using C = FftCfg<LOG2N>;
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
  LaneMap<LOG2N, LANEFAST> m(t, g);
  cf v[C::E];
#pragma unroll
  for (int e = 0; e < C::E; ++e)
    v[e] = valid ? src[(long long)(t + C::T * e) * a.S] : mk(0.f, 0.f);
  lane_fft<LOG2N, LANEFAST>(v, m, a.tw, smem);
  if (valid) {
#pragma unroll
    for (int e = 0; e < C::E; ++e)
      v[e] = cconj(cmul(v[e], ldtw(chirp, t + C::T * e)));
  }
  lane_fft<LOG2N, LANEFAST>(v, m, a.tw, smem);
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
    const long long k1 = r / n2, k2 = r % n2;
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
                                     long long n1, long long n_chirp) {
  const long long n2 = N / n1;
  const long long total = n_chirp * N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long c = i / N, r = i % N;
    dst[i] = src[c * N + (r / n2) + n1 * (r % n2)];
  }
}

}  // namespace bbt

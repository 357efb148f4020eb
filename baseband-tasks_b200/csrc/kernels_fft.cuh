// Batched strided FFT kernels (c2c, r2c, c2r) built on block_fft.
//
// Data are viewed as [outer][n][inner] (C order); the transform runs along the
// middle axis.  This one shape covers every FFT call site on the hot path:
//   Channelize          axis=1 of (spf, n)+sample_shape  (channelize.py:57-74)
//   Disperse small-N    axis=0 of (N,)+sample_shape      (dispersion.py:105-108)
//   FFTMaker plugin     any axis of any shape            (fourier/base.py:262-311)
// A CTA transforms G "lanes" (one lane = one (outer, inner) pair) at once.
//   LANEFAST: lanes are consecutive inner columns of one outer index, so a
//             warp touches G*8 contiguous bytes per FFT row (inner > 1).
//   !LANEFAST: lanes are consecutive outer indices (inner == 1); consecutive
//             threads of a lane touch consecutive elements.
#pragma once
#include "fft_core.cuh"
#include "rt.cuh"

namespace bbt {

struct FftArgs {
  const void* in;
  void* out;
  const cf* tw;     // exp(-2 pi i m / n), m < n
  long long outer;  // number of outer indices
  long long inner;  // number of inner columns
  int inverse;      // 0 forward, 1 backward
  float scale;      // applied to the output
};

template <class C, bool LANEFAST>
struct LaneMap {
  int t, g;
  long long b, c;  // outer index, inner column
  bool valid;
  BBT_HD LaneMap(int tid, long long blk, const FftArgs& a) {
    if (LANEFAST) {
      // Lanes enumerate (outer, inner) pairs with the inner column fastest.
      g = tid % C::G;
      t = tid / C::G;
      const long long lane = blk * C::G + g;
      b = lane / a.inner;
      c = lane % a.inner;
      valid = lane < a.outer * a.inner;
    } else {
      t = tid % C::T;
      g = tid / C::T;
      b = blk * C::G + g;
      c = 0;
      valid = b < a.outer;
    }
  }
};

template <class C, bool LANEFAST>
BBT_HD void lane_fft(cf* v, int t, int g, const cf* tw, cf* smem) {
  if (LANEFAST) {
    SmemLaneFast<C::PADSHIFT> sm{smem, g, C::G};
    block_fft<C>(v, t, tw, sm);
  } else {
    SmemLaneSlow<C::PADSHIFT> sm{smem + (size_t)g * C::NPAD};
    block_fft<C>(v, t, tw, sm);
  }
}

// Default shape of a block FFT of 2^LOG2N points: elements per thread and
// threads per CTA.
template <int LOG2N>
struct DefaultCfg {
  static constexpr int LOG2E = LOG2N <= 4 ? LOG2N : (LOG2N <= 8 ? 4 : 5);
  static constexpr int THREADS = LOG2N >= 14 ? 512 : 256;
  using type = FftCfg<LOG2N, LOG2E, THREADS>;
};

// Complex to complex.
template <class C, bool LANEFAST>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS, 1) fft_c2c_kernel(FftArgs a) {
  cf* smem = BBT_SMEM(cf);
  const long long blk = blockIdx.x;
  LaneMap<C, LANEFAST> m((int)threadIdx.x, blk, a);
  const cf* in = static_cast<const cf*>(a.in);
  cf* out = static_cast<cf*>(a.out);
  const long long base = m.b * C::N * a.inner + m.c;
  cf v[C::E];
#pragma unroll
  for (int e = 0; e < C::E; ++e) {
    cf x = mk(0.f, 0.f);
    if (m.valid) x = in[base + (long long)(m.t + C::T * e) * a.inner];
    v[e] = a.inverse ? cconj(x) : x;
  }
  lane_fft<C, LANEFAST>(v, m.t, m.g, a.tw, smem);
  if (m.valid) {
#pragma unroll
    for (int e = 0; e < C::E; ++e) {
      cf y = cscale(v[e], a.scale);
      out[base + (long long)(m.t + C::T * e) * a.inner] =
          a.inverse ? cconj(y) : y;
    }
  }
}

// Real to complex: n real samples -> n/2+1 bins (fourier/numpy.py:41-43).
// First version: transform the real data as complex with zero imaginary part.
template <class C, bool LANEFAST>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS, 1) fft_r2c_kernel(FftArgs a) {
  cf* smem = BBT_SMEM(cf);
  const long long blk = blockIdx.x;
  LaneMap<C, LANEFAST> m((int)threadIdx.x, blk, a);
  const float* in = static_cast<const float*>(a.in);
  cf* out = static_cast<cf*>(a.out);
  const long long ibase = m.b * C::N * a.inner + m.c;
  const long long obase = m.b * (C::N / 2 + 1) * a.inner + m.c;
  cf v[C::E];
#pragma unroll
  for (int e = 0; e < C::E; ++e) {
    float x = 0.f;
    if (m.valid) x = in[ibase + (long long)(m.t + C::T * e) * a.inner];
    v[e] = mk(x, 0.f);
  }
  lane_fft<C, LANEFAST>(v, m.t, m.g, a.tw, smem);
  if (m.valid) {
#pragma unroll
    for (int e = 0; e < C::E; ++e) {
      const int k = m.t + C::T * e;
      if (k <= C::N / 2) out[obase + (long long)k * a.inner] = cscale(v[e], a.scale);
    }
  }
}

// Complex to real: n/2+1 bins -> n real samples (fourier/numpy.py:46-49).
// Like numpy's irfft the imaginary parts of bins 0 and n/2 are ignored.
template <class C, bool LANEFAST>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS, 1) fft_c2r_kernel(FftArgs a) {
  cf* smem = BBT_SMEM(cf);
  const long long blk = blockIdx.x;
  LaneMap<C, LANEFAST> m((int)threadIdx.x, blk, a);
  const cf* in = static_cast<const cf*>(a.in);
  float* out = static_cast<float*>(a.out);
  const long long ibase = m.b * (C::N / 2 + 1) * a.inner + m.c;
  const long long obase = m.b * C::N * a.inner + m.c;
  cf v[C::E];
#pragma unroll
  for (int e = 0; e < C::E; ++e) {
    const int k = m.t + C::T * e;
    cf x = mk(0.f, 0.f);
    if (m.valid) {
      if (k <= C::N / 2) {
        x = in[ibase + (long long)k * a.inner];
        if (k == 0 || k == C::N / 2) x.y = 0.f;
        x = cconj(x);  // inverse transform = conj(fft(conj(X)))
      } else {
        x = in[ibase + (long long)(C::N - k) * a.inner];  // conj(conj(X[n-k]))
      }
    }
    v[e] = x;
  }
  lane_fft<C, LANEFAST>(v, m.t, m.g, a.tw, smem);
  if (m.valid) {
#pragma unroll
    for (int e = 0; e < C::E; ++e)
      out[obase + (long long)(m.t + C::T * e) * a.inner] = v[e].x * a.scale;
  }
}

// Out-of-place transpose of a [rows][cols] complex matrix per batch
// (natural-order output of the two-pass large FFT).
BBT_GLOBAL void transpose_kernel(const cf* BBT_RESTRICT in, cf* BBT_RESTRICT out,
                                 long long rows, long long cols) {
  cf* tile = BBT_SMEM(cf);  // 32 x 33
  const long long batch = blockIdx.z;
  const cf* src = in + batch * rows * cols;
  cf* dst = out + batch * rows * cols;
  const long long c0 = (long long)blockIdx.x * 32, r0 = (long long)blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  for (int i = ty; i < 32; i += blockDim.y) {
    const long long r = r0 + i, c = c0 + tx;
    if (r < rows && c < cols) tile[i * 33 + tx] = src[r * cols + c];
  }
  BBT_SYNC();
  for (int i = ty; i < 32; i += blockDim.y) {
    const long long c = c0 + i, r = r0 + tx;
    if (r < rows && c < cols) dst[c * rows + r] = tile[tx * 33 + i];
  }
}

// Pointwise twiddle of the four-step FFT: a[k1][n2] *= W_N^{k1 n2} (or its
// conjugate).  Only used by the generic large-N FFTMaker path; the
// dedispersion plan folds this into its row kernel.
struct BigTwiddle {
  const cf* lo;  // exp(-2 pi i m / N), m < kTwiddleTable
  const cf* hi;  // exp(-2 pi i m kTwiddleTable / N), m < N / kTwiddleTable
  BBT_HD cf get(long long m) const {
    cf w = ldtw(lo, (int)(m & (kTwiddleTable - 1)));
    const long long h = m >> kLog2TwiddleTable;
    if (h) w = cmul(w, ldtw(hi, (int)h));
    return w;
  }
};

// data is [batch][n1][n2][inner].
BBT_GLOBAL void twiddle_kernel(cf* data, long long n1, long long n2,
                               long long batch, long long inner, BigTwiddle tw,
                               int conj) {
  const long long total = batch * n1 * n2 * inner;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long r = (i / inner) % (n1 * n2);
    cf w = tw.get((r / n2) * (r % n2));
    data[i] = conj ? cmulc(data[i], w) : cmul(data[i], w);
  }
}

// [batch][rows][cols][inner] -> [batch][cols][rows][inner] for inner > 1
// (writes coalesced; runs of `inner` values on the reading side).
BBT_GLOBAL void transpose_inner_kernel(const cf* BBT_RESTRICT in,
                                       cf* BBT_RESTRICT out, long long batch,
                                       long long rows, long long cols,
                                       long long inner) {
  const long long total = batch * rows * cols * inner;
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (long long)gridDim.x * blockDim.x) {
    const long long c = o % inner, q = o / inner;
    const long long r = q % rows, q2 = q / rows;
    const long long col = q2 % cols, b = q2 / cols;
    out[o] = in[((b * rows + r) * cols + col) * inner + c];
  }
}

// ---------------------------------------------------------------------------
// Helpers of the general FFT plan (bbt_fft.cu): transforms of any length n
// through power-of-two transforms of length m >= 2n - 1 (Bluestein: with
// a_j = exp(-i pi j^2 / n), X_k = a_k sum_j (x_j a_j) conj(a)_{k-j}), and real
// transforms through complex ones.  Data are [outer][len][inner].
struct AxisArgs {
  const void* in;
  void* out;
  const cf* chirp;     // a_j, j < n (or the filter B_j, j < m)
  long long outer, inner;
  long long n_in;      // length of the axis in `in`
  long long n_out;     // length of the axis in `out`
  long long n;         // transform length (Hermitian extension)
  int in_real;         // input is float32 rather than complex64
  int out_real;        // output is float32
  int conj_in;         // conjugate the input (inverse transforms)
  int conj_out;        // conjugate the output
  float scale;
};

// out[b][j][c] = (j < n_in ? in[b][j][c] (conjugated?) * chirp[j] : 0) for
// j < n_out: multiply by the chirp and zero-pad.  With chirp == nullptr the
// values are only converted / copied (real -> complex, truncation).
BBT_GLOBAL void axis_pre_kernel(AxisArgs a) {
  const long long total = a.outer * a.n_out * a.inner;
  cf* out = static_cast<cf*>(a.out);
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (long long)gridDim.x * blockDim.x) {
    const long long c = o % a.inner, q = o / a.inner;
    const long long j = q % a.n_out, b = q / a.n_out;
    cf x = mk(0.f, 0.f);
    if (j < a.n_in) {
      const long long i = (b * a.n_in + j) * a.inner + c;
      x = a.in_real ? mk(static_cast<const float*>(a.in)[i], 0.f)
                    : static_cast<const cf*>(a.in)[i];
      if (a.conj_in) x = cconj(x);
      if (a.chirp) x = cmul(x, a.chirp[j]);
    }
    out[o] = x;
  }
}

// Hermitian extension of half spectra [outer][n/2+1][inner] to
// [outer][n][inner], as numpy's irfft reads them (imaginary parts of bins 0
// and n/2 ignored); optionally conjugated and times the chirp.
BBT_GLOBAL void axis_hermitian_kernel(AxisArgs a) {
  const long long total = a.outer * a.n_out * a.inner;
  const cf* in = static_cast<const cf*>(a.in);
  cf* out = static_cast<cf*>(a.out);
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (long long)gridDim.x * blockDim.x) {
    const long long c = o % a.inner, q = o / a.inner;
    const long long j = q % a.n_out, b = q / a.n_out;
    cf x = mk(0.f, 0.f);
    if (j < a.n) {
      const long long h = j <= a.n / 2 ? j : a.n - j;
      x = in[(b * a.n_in + h) * a.inner + c];
      if (j == 0 || 2 * j == a.n) x.y = 0.f;
      if (j > a.n / 2) x = cconj(x);
      if (a.conj_in) x = cconj(x);
      if (a.chirp) x = cmul(x, a.chirp[j]);
    }
    out[o] = x;
  }
}

// data[b][j][c] *= filter[j]  (j < n_out).
BBT_GLOBAL void axis_filter_kernel(AxisArgs a) {
  const long long total = a.outer * a.n_out * a.inner;
  cf* data = static_cast<cf*>(a.out);
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (long long)gridDim.x * blockDim.x) {
    const long long j = (o / a.inner) % a.n_out;
    data[o] = cmul(data[o], a.chirp[j]);
  }
}

// out[b][k][c] = in[b][k][c] * chirp[k] * scale (conjugated?) for k < n_out,
// taken from an axis of length n_in; as complex64 or its real part.
BBT_GLOBAL void axis_post_kernel(AxisArgs a) {
  const long long total = a.outer * a.n_out * a.inner;
  const cf* in = static_cast<const cf*>(a.in);
  for (long long o = (long long)blockIdx.x * blockDim.x + threadIdx.x; o < total;
       o += (long long)gridDim.x * blockDim.x) {
    const long long c = o % a.inner, q = o / a.inner;
    const long long k = q % a.n_out, b = q / a.n_out;
    cf x = in[(b * a.n_in + k) * a.inner + c];
    if (a.chirp) x = cmul(x, a.chirp[k]);
    x = cscale(x, a.scale);
    if (a.conj_out) x = cconj(x);
    if (a.out_real)
      static_cast<float*>(a.out)[o] = x.x;
    else
      static_cast<cf*>(a.out)[o] = x;
  }
}

}  // namespace bbt

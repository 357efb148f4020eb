// Detection and integration kernels.
//   power / square      functions.py:15-16,132-143
//   channelize+power(+integrate) fused   channelize.py:73-74 + functions.py:132-143
//                                        + integration.py:273-303
//   integrate           integration.py:290-303 (segmented sums + counts)
//   fold                integration.py:380-395 (phase-bin scatter add)
#pragma once
#include "kernels_fft.cuh"
#include "tma.cuh"

namespace bbt {

struct alignas(16) f4 {
  float x, y, z, w;
};

BBT_HD f4 stokes_like(cf a, cf b) {
  // [|X|^2, |Y|^2, Re(X conj Y), Im(X conj Y)]  (functions.py:138-142)
  f4 r;
  r.x = a.x * a.x + a.y * a.y;
  r.y = b.x * b.x + b.y * b.y;
  r.z = a.x * b.x + a.y * b.y;
  r.w = a.y * b.x - a.x * b.y;
  return r;
}

// The same four products accumulated for a polarization pair whose halves sit
// in two threads: each keeps one value (`keep`, its own polarization) and is
// sent the partner's (`other`).  Which of the two is X depends on the thread,
// but |keep|^2, |other|^2 and Re(keep conj other) do not care and
// Im(keep conj other) only changes sign, so the sums are formed the same way
// in both threads (no selects in the loop) and put in order once at the end
// (stokes_finish).  acc = [sum |keep|^2, sum |other|^2, sum Re, sum Im]; the
// two squares share packed multiply-adds.
BBT_HD void stokes_accumulate(f4& acc, cf keep, cf other) {
#if defined(BBT_PACKED)
  cf pw = mk(acc.x, acc.y);
  pw = u2(fma2(p2(keep.x, other.x), p2(keep.x, other.x), p2(pw)));
  pw = u2(fma2(p2(keep.y, other.y), p2(keep.y, other.y), p2(pw)));
  acc.x = pw.x;
  acc.y = pw.y;
#else
  acc.x = fmaf(keep.y, keep.y, fmaf(keep.x, keep.x, acc.x));
  acc.y = fmaf(other.y, other.y, fmaf(other.x, other.x, acc.y));
#endif
  acc.z = fmaf(keep.y, other.y, fmaf(keep.x, other.x, acc.z));
  acc.w = fmaf(-keep.x, other.y, fmaf(keep.y, other.x, acc.w));
}
// keep = Y, other = X (p = 1): swap the squares, Im(X conj Y) = -Im(Y conj X).
BBT_HD f4 stokes_finish(f4 acc, int p) {
  f4 r = acc;
  if (p) {
    r.x = acc.y;
    r.y = acc.x;
    r.w = -acc.w;
  }
  return r;
}

// (A, 2, B) complex -> (A, 4, B) float.
BBT_GLOBAL void power_kernel(const cf* BBT_RESTRICT in, float* BBT_RESTRICT out,
                             long long A, long long B) {
  const long long total = A * B;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long a = i / B, b = i % B;
    const f4 p = stokes_like(in[(a * 2) * B + b], in[(a * 2 + 1) * B + b]);
    float* o = out + (a * 4) * B + b;
    o[0] = p.x;
    o[B] = p.y;
    o[2 * B] = p.z;
    o[3 * B] = p.w;
  }
}

// out[i] = a[i] * b[i], complex64 (the frequency-domain multiply of
// dispersion.py:137 for frame lengths the fused plan does not take).
BBT_GLOBAL void multiply_kernel(const cf* BBT_RESTRICT a, const cf* BBT_RESTRICT b,
                                cf* BBT_RESTRICT out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = cmul(a[i], b[i]);
}

BBT_GLOBAL void square_kernel(const float* BBT_RESTRICT in, float* BBT_RESTRICT out,
                              long long n, int is_complex) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x) {
    if (is_complex) {
      const float re = in[2 * i], im = in[2 * i + 1];
      out[i] = re * re + im * im;
    } else {
      out[i] = in[i] * in[i];
    }
  }
}

// ---------------------------------------------------------------------------
// Fused Channelize(n) -> Power (-> Integrate).
// Input  x[(j*n + i)][m][p]  complex, j spectrum, i sample in block, m < M,
//        p polarization (2).
// Output (no integrate)  out[j][k][m][4] float.
// Output (integrate)     sum[b][k][m][4] += ..., count[b] += width, for bins
//        b with spectra [offsets[b], offsets[b+1]) clipped to the spectra
//        [j_first, j_first + n_spec) present in this call.
struct ChanPowArgs {
  const cf2* in;
  float* out;             // out or sum
  unsigned long long* count;
  const long long* offsets;  // bin edges in spectra (absolute)
  const cf* tw;
  long long M;
  long long n_spec;       // spectra in this call
  long long j_first;      // absolute index of the first spectrum in `in`
  long long b_first;      // first bin handled (blockIdx.y = 0)
  long long msub;         // concurrent spectrum sub-streams per bin
  int average;            // scale sums by 1 / (bin width) on the way out
};

// Lanes of a tile are (unit, polarization) pairs, polarization fastest, where
// a unit is one (spectrum sub-stream, m) pair with m fastest; consecutive
// threads take consecutive lanes, so a warp reads runs of G*8 contiguous
// bytes per time sample.  After the transform the two threads holding X and Y
// of a channel swap half of their values (one shuffle per value), and each
// forms all four products for every other channel.
template <class C, bool INTEGRATE>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS, 1) chanpow_kernel(ChanPowArgs a) {
  static_assert(C::G % 2 == 0 && C::E % 2 == 0, "pairs of lanes and values");
  cf* smem = BBT_SMEM(cf);
  const int tid = threadIdx.x;
  const int g = tid % C::G, t = tid / C::G;
  const int p = g & 1;
  const long long unit = (long long)blockIdx.x * (C::G / 2) + (g >> 1);
  const long long m = unit % a.M, jsub = unit / a.M;
  long long lo, hi;  // spectra (relative to j_first) this CTA walks through
  long long b = 0;
  if (INTEGRATE) {
    b = a.b_first + blockIdx.y;
    lo = a.offsets[b] - a.j_first;
    hi = a.offsets[b + 1] - a.j_first;
    if (lo < 0) lo = 0;
    if (hi > a.n_spec) hi = a.n_spec;
  } else {
    lo = 0;
    hi = a.n_spec;
  }
  const bool lane_ok = jsub < a.msub;
  const cf* in = reinterpret_cast<const cf*>(a.in);
  const long long row = a.M * 2;  // complex values per time sample
  f4 acc[C::E / 2];
  if (INTEGRATE) {
#pragma unroll
    for (int i = 0; i < C::E / 2; ++i)
      acc[i].x = acc[i].y = acc[i].z = acc[i].w = 0.f;
  }
  SmemLaneFast<C::PADSHIFT> sm{smem, g, C::G};
  // When shared memory has room for a second tile, the next spectrum is
  // copied in asynchronously (cp.async) while this one is transformed; each
  // thread stages exactly the values it will consume itself.
  constexpr bool STAGE =
      INTEGRATE && C::SMEM_BYTES + (size_t)C::G * C::N * sizeof(cf) <= 200 * 1024;
  cf* stage = smem + (C::SMEM_BYTES / sizeof(cf)) + (size_t)t * C::G + g;
  const long long pstep = (long long)C::T * row;
  // CTAs sharing one spectrum block (when the series divide evenly).
  long long share = 0, share_at = 0, jsub0 = 0;
  if (STAGE && a.M % (C::G / 2) == 0) {
    const long long nshare = a.M / (C::G / 2);
    jsub0 = blockIdx.x / nshare;
    share = (long long)C::N * row / nshare;
    share_at = (blockIdx.x - jsub0 * nshare) * share;
    if (jsub0 >= a.msub) share = 0;
  }
  if (STAGE && lo < hi) {
    const long long j = lo + jsub;
    const bool valid = lane_ok && j < hi;
    const cf* pe = in + (j * C::N) * row + m * 2 + p + (long long)t * row;
#pragma unroll
    for (int e = 0; e < C::E; ++e) {
      cp_async8(stage + e * (C::T * C::G), valid ? pe : in, valid);
      pe += pstep;
    }
  }
  for (long long j0 = lo; j0 < hi; j0 += a.msub) {
    const long long j = j0 + jsub;
    const bool valid = lane_ok && j < hi;
    const cf* src = in + (j * C::N) * row + m * 2 + p;
    cf v[C::E];
    if (STAGE) {
      cp_async_wait();
#pragma unroll
      for (int e = 0; e < C::E; ++e) v[e] = stage[e * (C::T * C::G)];
      if (j0 + a.msub < hi) {
        const long long jn = j + a.msub;
        const bool nvalid = lane_ok && jn < hi;
        const cf* pe = src + a.msub * C::N * row + (long long)t * row;
#pragma unroll
        for (int e = 0; e < C::E; ++e) {
          cp_async8(stage + e * (C::T * C::G), nvalid ? pe : in, nvalid);
          pe += pstep;
        }
      }
      if (share > 0 && j0 + 2 * a.msub < hi) {
        // The spectrum after next, into L2: the CTAs that split the series
        // of a spectrum each ask for a contiguous share of its block, so
        // DRAM sees whole lines rather than each CTA's 64-byte pieces.
        const cf* pb = in + ((j0 + 2 * a.msub + jsub0) * C::N) * row + share_at;
        for (long long i = (long long)tid * 16; i < share;
             i += (long long)C::THREADS * 16)
          prefetch_l2(pb + i);
      }
    } else {
      {
        const cf* pe = src + (long long)t * row;
#pragma unroll
        for (int e = 0; e < C::E; ++e) {
          v[e] = valid ? ld_stream(pe) : mk(0.f, 0.f);
          pe += pstep;
        }
      }
      if (lane_ok && j + a.msub < hi && (g & 15) == 0) {
        // Next spectrum of this lane group into L2 meanwhile.
        const cf* pe = src + a.msub * C::N * row + (long long)t * row;
#pragma unroll
        for (int e = 0; e < C::E; ++e) {
          prefetch_l2(pe);
          pe += pstep;
        }
      }
    }
    block_fft<C>(v, t, a.tw, sm);
#pragma unroll
    for (int i = 0; i < C::E / 2; ++i) {
      // Keep value 2i+p, swap value 2i+(1-p) for the partner's 2i+p.
      const cf mine = p ? v[2 * i + 1] : v[2 * i];
      const cf send = p ? v[2 * i] : v[2 * i + 1];
      cf other;
      other.x = shfl_xor1(send.x);
      other.y = shfl_xor1(send.y);
      if (INTEGRATE) {
        stokes_accumulate(acc[i], mine, other);
      } else if (valid) {
        const f4 q = p ? stokes_like(other, mine) : stokes_like(mine, other);
        const int k = t + C::T * (2 * i + p);
        reinterpret_cast<f4*>(a.out)[(j * C::N + k) * a.M + m] = q;
      }
    }
  }
  if (INTEGRATE) {
    if (lane_ok && hi > lo) {
      // Averages: every partial sum is divided by the full width of its bin
      // (known from the offsets table), so no separate division pass is needed.
      const float w =
          a.average ? (float)(a.offsets[b + 1] - a.offsets[b]) : 1.f;
#pragma unroll
      for (int i = 0; i < C::E / 2; ++i) {
        const int k = t + C::T * (2 * i + p);
        float* o = a.out + ((b * C::N + k) * a.M + m) * 4;
        const f4 q = stokes_finish(acc[i], p);
        atomic_add(o + 0, q.x / w);
        atomic_add(o + 1, q.y / w);
        atomic_add(o + 2, q.z / w);
        atomic_add(o + 3, q.w / w);
      }
    }
    if (tid == 0 && blockIdx.x == 0 && hi > lo)
      atomic_add(a.count + b, (unsigned long long)(hi - lo));
  }
}

// The same for narrow samples (M = 1, 2 or 4 polarization pairs per time
// sample), fed by bulk asynchronous copies: the G/2 units of a CTA are then
// G/2/M adjacent spectra of all M series, i.e. one contiguous 64 KB block of
// the input per tile, which one thread copies into a two-deep ring in shared
// memory (cp.async.bulk, completion counted on an mbarrier) while the other
// tile is transformed.  No per-thread loads, address arithmetic or L2
// prefetches are left; DRAM latency is covered by the two tiles in flight.
// Stage layout: [jl][i][m][p] with a few values of padding between the spectra
// jl so that what a half-warp reads falls in distinct banks.
template <class C>
struct ChanPowTma {
#ifndef BBT_CHANPOW_STAGES
// One stage: the copy of the next tile is issued as soon as every thread has
// taken its values of this one and lands during the transform; a second stage
// bought nothing and its 33 KB cost 13 % (C4, 4 lanes: 1.53 -> 1.34 ms per 32
// frames), with a register cap for three CTAs per SM on top 1.53 again.
#define BBT_CHANPOW_STAGES 1
#endif
  static constexpr int kStages = BBT_CHANPOW_STAGES;
  static BBT_HD constexpr long long spectrum_elems(long long M) {
    return (long long)C::N * M * 2;
  }
  // Padding between the spectra of a stage: a half-warp reads 16 / nj
  // consecutive values of each of the nj spectra; with this stride they fall
  // in distinct banks (multiples of two keep the copies 16-byte aligned).
  static BBT_HD constexpr long long pad_elems(long long M) {
    // (Eight values measured faster than the four that would spread a
    // half-warp's 64-bit reads over all banks: 2.17 against 2.30 ms per 32
    // C4 frames.)
    return C::G / 2 / M <= 1 ? 0 : 8;
  }
  static BBT_HD constexpr long long stride_elems(long long M) {
    return spectrum_elems(M) + pad_elems(M);
  }
  static BBT_HD constexpr long long stage_elems(long long M) {
    return (C::G / 2 / M) * stride_elems(M) + 16;
  }
  static BBT_HD constexpr size_t smem_bytes(long long M) {
    return C::SMEM_BYTES + kStages * stage_elems(M) * sizeof(cf) +
           kStages * sizeof(Mbar);
  }
};

#ifndef BBT_CHANPOW_CTAS
#define BBT_CHANPOW_CTAS 1
#endif
template <class C>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS,
                                  (C::THREADS <= 128 ? BBT_CHANPOW_CTAS : 1))
    chanpow_tma_kernel(ChanPowArgs a) {
  static_assert(C::G % 2 == 0 && C::E % 2 == 0, "pairs of lanes and values");
  using K = ChanPowTma<C>;
  cf* smem = BBT_SMEM(cf);
  const int tid = threadIdx.x;
  const int g = tid % C::G, t = tid / C::G;
  const int p = g & 1;
  const int M = (int)a.M;
  const int nj = C::G / 2 / M;                 // spectra per tile
  const long long jsub0 = (long long)blockIdx.x * nj;
  const int jl = (g >> 1) / M;                 // this lane's spectrum in tile
  const long long m = (g >> 1) - jl * M, jsub = jsub0 + jl;
  const long long b = a.b_first + blockIdx.y;
  long long lo = a.offsets[b] - a.j_first, hi = a.offsets[b + 1] - a.j_first;
  if (lo < 0) lo = 0;
  if (hi > a.n_spec) hi = a.n_spec;
  const bool lane_ok = jsub < a.msub;
  const long long spec = K::spectrum_elems(M);   // complex values per spectrum
  cf* ring = smem + C::SMEM_BYTES / sizeof(cf);
  Mbar* bars = reinterpret_cast<Mbar*>(ring + K::kStages * K::stage_elems(M));
  const cf* in = reinterpret_cast<const cf*>(a.in);
  f4 acc[C::E / 2];
#pragma unroll
  for (int i = 0; i < C::E / 2; ++i)
    acc[i].x = acc[i].y = acc[i].z = acc[i].w = 0.f;
  if (tid == 0)
    for (int s = 0; s < K::kStages; ++s) mbar_init(bars + s, 1);
  BBT_SYNC();
  const long long n_iter = hi > lo ? (hi - lo + a.msub - 1) / a.msub : 0;
  // Spectra of this CTA's tile that exist in iteration `it`.
  auto n_valid = [&](long long it) -> int {
    const long long j0 = lo + it * a.msub;
    long long end = j0 + a.msub;
    if (end > hi) end = hi;
    long long n = end - (j0 + jsub0);
    return n < 0 ? 0 : (n > nj ? nj : (int)n);
  };
  auto issue = [&](long long it, int s) {        // one thread
    const int nv = n_valid(it);
    Mbar* bar = bars + s;
    if (nv == 0) {
      mbar_arrive(bar);
      return;
    }
    const cf* src = in + (lo + it * a.msub + jsub0) * spec;
    cf* dst = ring + s * K::stage_elems(M);
    mbar_expect_tx(bar, (uint32_t)(nv * spec * sizeof(cf)));
    for (int q = 0; q < nv; ++q)
      bulk_load(dst + q * K::stride_elems(M), src + q * spec,
                (uint32_t)(spec * sizeof(cf)), bar, q == nv - 1);
  };
  if (tid == 0)
    for (int s = 0; s < K::kStages; ++s)
      if (s < n_iter) issue(s, s);
  SmemLaneFast<C::PADSHIFT> sm{smem, g, C::G};
  const cf* mine = ring + jl * K::stride_elems(M) + (long long)t * (2 * M) +
                   (m * 2 + p);
  for (long long it = 0; it < n_iter; ++it) {
    const int s = (int)(it % K::kStages);
    const bool valid = lane_ok && jl < n_valid(it);
    mbar_wait(bars + s, (uint32_t)(it / K::kStages) & 1u,
              (uint32_t)(it / K::kStages));
    cf v[C::E];
    const cf* src = mine + s * K::stage_elems(M);
#pragma unroll
    for (int e = 0; e < C::E; ++e)
      v[e] = valid ? src[(long long)e * C::T * (2 * M)] : mk(0.f, 0.f);
    BBT_SYNC();  // every thread has taken its values: the stage is free
    if (tid == 0 && it + K::kStages < n_iter) issue(it + K::kStages, s);
    block_fft<C>(v, t, a.tw, sm);
#pragma unroll
    for (int i = 0; i < C::E / 2; ++i) {
      const cf keep = p ? v[2 * i + 1] : v[2 * i];
      const cf send = p ? v[2 * i] : v[2 * i + 1];
      cf other;
      other.x = shfl_xor1(send.x);
      other.y = shfl_xor1(send.y);
      stokes_accumulate(acc[i], keep, other);
    }
  }
  if (lane_ok && hi > lo) {
    const float w =
        a.average ? (float)(a.offsets[b + 1] - a.offsets[b]) : 1.f;
#pragma unroll
    for (int i = 0; i < C::E / 2; ++i) {
      const int k = t + C::T * (2 * i + p);
      float* o = a.out + ((b * C::N + k) * a.M + m) * 4;
      const f4 q = stokes_finish(acc[i], p);
      atomic_add(o + 0, q.x / w);
      atomic_add(o + 1, q.y / w);
      atomic_add(o + 2, q.z / w);
      atomic_add(o + 3, q.w / w);
    }
  }
  if (tid == 0 && blockIdx.x == 0 && hi > lo)
    atomic_add(a.count + b, (unsigned long long)(hi - lo));
}

// ---------------------------------------------------------------------------
// Integrate: sum[b][c] += sum_{i in [offsets[b], offsets[b+1])} in[i][c],
// count[b] += width, with the bin clipped to the samples
// [i_first, i_first + n) present in this call (integration.py:290-303).
struct IntegrateArgs {
  const float* in;
  float* sum;
  unsigned long long* count;
  const long long* offsets;
  long long inner, n, i_first, b_first, msub;
  int average;
};

BBT_GLOBAL void integrate_kernel(IntegrateArgs a) {
  const long long lane = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long c = lane % a.inner, sub = lane / a.inner;
  const long long b = a.b_first + blockIdx.y;
  long long lo = a.offsets[b] - a.i_first, hi = a.offsets[b + 1] - a.i_first;
  if (lo < 0) lo = 0;
  if (hi > a.n) hi = a.n;
  if (hi <= lo) return;
  if (lane == 0) atomic_add(a.count + b, (unsigned long long)(hi - lo));
  if (sub >= a.msub) return;
  float acc = 0.f;
  for (long long i = lo + sub; i < hi; i += a.msub) acc += a.in[i * a.inner + c];
  if (a.average) acc /= (float)(a.offsets[b + 1] - a.offsets[b]);
  atomic_add(a.sum + b * a.inner + c, acc);
}

// ---------------------------------------------------------------------------
// Fold: for every sample i of time bin b, phase bin
//   p = int(((phase(i) mod 1) * n_phase))        (integration.py:389-391)
// and sum[b][p][c] += x[i][c], count[b][p] += 1  (integration.py:394-395).
// phase(i) = Horner(coef, dt), dt = (double(i_phase + i) - i_ref) / rate, all in float64
// with individually rounded operations (no FMA), so that the bin assignment
// is bit-identical to the oracle's numpy arithmetic.  Alternatively the phase
// bins may be supplied precomputed (pbin != nullptr) for arbitrary callables.
// With POWER the input is [n][M][2] complex and the four polarization
// products are formed on the fly (Power -> Fold without an HBM round trip).
struct FoldArgs {
  const void* in;
  float* sum;                 // [n_tbin][n_phase][inner]
  unsigned long long* count;  // [n_tbin][n_phase]
  const long long* lo;        // [n_tbin] first sample (absolute) of bin
  const long long* hi;        // [n_tbin] one past last sample of bin
  const int* pbin;            // optional precomputed phase bins [n]
  long long inner;            // floats per sample in the output
  long long n, i_first;       // samples in this call; absolute index of first
  long long i_phase;          // index of the first sample on the grid the
                              // phase polynomial counts on (i_ref refers to it)
  long long b_first;          // time bin of blockIdx.y = 0
  double i_ref;               // sample index (may be fractional) at which dt = 0
  double rate;
  double inv_rate;            // RN(1 / rate), or 0 to divide (see fold_phase_bin_fast)
  double coef[8];
  int ncoef;
  int n_phase;
  int use_smem;               // privatise the profile in shared memory
                              // (2: and take the four-values-per-sample path;
                              //  3: with TMA-staged tiles)
  int ring_offset;            // bytes from the profile to the tile ring
};

BBT_DEV int fold_phase_bin(const FoldArgs& a, long long i_abs) {
  const double dt = ddiv(dadd((double)i_abs, -a.i_ref), a.rate);
  double ph = a.coef[a.ncoef - 1];
  for (int k = a.ncoef - 2; k >= 0; --k) ph = dadd(dmul(ph, dt), a.coef[k]);
  // fmod(ph, 1) is exactly ph - trunc(ph) (the difference is representable).
  double r = dadd(ph, -trunc(ph));
  if (r < 0.) r = dadd(r, 1.0);  // numpy's floored modulo
  return (int)dmul(r, (double)a.n_phase);
}

// Consecutive samples mostly fall in the same phase bin (a bin lasts many
// samples), so a warp first checks whether all its 32 samples share one bin;
// if so it reduces them with shuffles and lane 0 adds the total to a running
// per-warp accumulator that is only flushed (one atomic per value) when the
// bin changes.  Mixed warps fall back to one shared-memory atomic per sample.
// 1: feed the fold kernel from TMA-staged tiles (cp.async.bulk + mbarrier rings
// per warp).  Parity-tested on the B200 and measured slower than direct
// streaming loads (0.52 against 0.34 ms per C5 launch: every sample is used
// once, so the detour through shared memory only adds latency and costs a
// fourth resident CTA), hence off.
#ifndef BBT_FOLD_TMA
#define BBT_FOLD_TMA 0
#endif
constexpr int kFoldFast = 8;  // most values per sample kept in registers
constexpr int kFoldUnroll = 4;  // samples in flight per thread (4 values each)
constexpr int kFoldThreads = 256;
#ifndef BBT_FOLD_STAGES
#define BBT_FOLD_STAGES 4
#endif
constexpr int kFoldStages = BBT_FOLD_STAGES;  // tiles in flight in the TMA-staged variant
constexpr int kFoldTile = kFoldThreads * kFoldUnroll;  // samples per tile

#if defined(__CUDA_ARCH__)
// Phase bin of absolute sample index xi (already a double; exact below 2^53),
// the same arithmetic as fold_phase_bin with two shortcuts that do not change
// a single bit: the division by the sample rate is done with the correctly
// rounded reciprocal y = RN(1 / rate) and two Newton corrections in FMA
// arithmetic (Markstein: with y correctly rounded and q1 faithful,
// RN(q1 + RN(x - q1 rate) y) is the correctly rounded quotient), and the
// Horner recurrence always runs its last four steps, starting from zero
// (0 * dt + c is exactly c, so missing high coefficients change nothing).
struct FoldPhase {
  double i_ref, rate, inv_rate, c0, c1, c2, c3, n_phase;
};
BBT_DEV int fold_phase_bin_fast(const FoldArgs& a, const FoldPhase& f, double xi) {
  const double x = dadd(xi, -f.i_ref);
  double dt;
  if (f.inv_rate != 0.) {
    const double q0 = dmul(x, f.inv_rate);
    const double q1 = __fma_rn(__fma_rn(-q0, f.rate, x), f.inv_rate, q0);
    dt = __fma_rn(__fma_rn(-q1, f.rate, x), f.inv_rate, q1);
  } else {
    dt = ddiv(x, f.rate);
  }
  double ph = 0.;
  for (int k = a.ncoef - 1; k >= 4; --k) ph = dadd(dmul(ph, dt), a.coef[k]);
  ph = dadd(dmul(ph, dt), f.c3);
  ph = dadd(dmul(ph, dt), f.c2);
  ph = dadd(dmul(ph, dt), f.c1);
  ph = dadd(dmul(ph, dt), f.c0);
  double r = dadd(ph, -trunc(ph));
  if (r < 0.) r = dadd(r, 1.0);
  int q = (int)dmul(r, f.n_phase);
  q = q < 0 ? 0 : q;
  return q >= a.n_phase ? a.n_phase - 1 : q;
}

// Four values per sample (the Stokes-like products of one polarization pair,
// or four floats).  Each thread has kFoldUnroll samples in flight so that DRAM
// latency is covered; a warp whose 32 samples share a phase bin reduces its
// 128 values with six shuffles (lanes trade halves, then quarters, of their
// values before the plain butterfly), leaving component c in the lanes with
// (lane >> 3) == c, where it is kept in a running sum until the bin changes.
template <bool POWER, bool CHECK, int USTRIDE>
BBT_DEV void fold_four_bins(const FoldArgs& a, const FoldPhase& f, float* hist,
                            unsigned* hcnt, const f4* x, long long ib,
                            long long i1, double xd, int lane, float& acc,
                            unsigned& acc_n, int& cur);

template <bool POWER, bool CHECK>
BBT_DEV void fold_four_tile(const FoldArgs& a, const FoldPhase& f, float* hist,
                            unsigned* hcnt, long long ib, long long i1, double xd,
                            int lane, float& acc, unsigned& acc_n, int& cur) {
  f4 x[kFoldUnroll];
#pragma unroll
  for (int u = 0; u < kFoldUnroll; ++u) {
    const long long i = ib + u * kFoldThreads + lane;
    x[u].x = x[u].y = x[u].z = x[u].w = 0.f;
    if (!CHECK || i < i1) {
      const float4 q = __ldcs(static_cast<const float4*>(a.in) + i);
      x[u].x = q.x, x[u].y = q.y, x[u].z = q.z, x[u].w = q.w;
    }
  }
  fold_four_bins<POWER, CHECK, kFoldThreads>(a, f, hist, hcnt, x, ib, i1, xd,
                                             lane, acc, acc_n, cur);
}

// x[u] is the sample ib + u * USTRIDE + lane.
template <bool POWER, bool CHECK, int USTRIDE>
BBT_DEV void fold_four_bins(const FoldArgs& a, const FoldPhase& f, float* hist,
                            unsigned* hcnt, const f4* x, long long ib,
                            long long i1, double xd, int lane, float& acc,
                            unsigned& acc_n, int& cur) {
  const unsigned full = 0xffffffffu;
  const int comp = lane >> 3;
  const bool flusher = (lane & 7) == 0;
  int p[kFoldUnroll];
#pragma unroll
  for (int u = 0; u < kFoldUnroll; ++u) {
    const long long i = ib + u * USTRIDE + lane;
    p[u] = -1;
    if (!CHECK || i < i1) {
      if (a.pbin) {
        int q = a.pbin[i];
        q = q < 0 ? 0 : q;
        p[u] = q >= a.n_phase ? a.n_phase - 1 : q;
      } else {
        p[u] = fold_phase_bin_fast(a, f, xd + (double)(u * USTRIDE));
      }
    }
  }
#pragma unroll
  for (int u = 0; u < kFoldUnroll; ++u) {
    if (CHECK && ib + u * USTRIDE >= i1) break;  // warp-uniform
    f4 v = x[u];
    if (POWER) {
      cf xa, xb;
      xa.x = v.x, xa.y = v.y, xb.x = v.z, xb.y = v.w;
      v = stokes_like(xa, xb);
    }
    const int p0 = __shfl_sync(full, p[u], 0);
    if (__all_sync(full, p[u] == p0)) {
      const bool up = lane & 16;
      float k0 = up ? v.z : v.x, k1 = up ? v.w : v.y;
      k0 += __shfl_xor_sync(full, up ? v.x : v.z, 16);
      k1 += __shfl_xor_sync(full, up ? v.y : v.w, 16);
      const bool up2 = lane & 8;
      float k = up2 ? k1 : k0;
      k += __shfl_xor_sync(full, up2 ? k0 : k1, 8);
      k += __shfl_xor_sync(full, k, 4);
      k += __shfl_xor_sync(full, k, 2);
      k += __shfl_xor_sync(full, k, 1);
      if (p0 != cur) {
        if (cur >= 0 && flusher) {
          atomicAdd(hist + cur * 4 + comp, acc);
          if (lane == 0) atomicAdd(hcnt + cur, acc_n);
        }
        cur = p0;
        acc = 0.f;
        acc_n = 0;
      }
      acc += k;
      acc_n += 32;
    } else if (p[u] >= 0) {
      atomicAdd(hist + p[u] * 4 + 0, v.x);
      atomicAdd(hist + p[u] * 4 + 1, v.y);
      atomicAdd(hist + p[u] * 4 + 2, v.z);
      atomicAdd(hist + p[u] * 4 + 3, v.w);
      atomicAdd(hcnt + p[u], 1u);
    }
  }
}

template <bool POWER>
BBT_DEV void fold_four(const FoldArgs& a, float* hist, unsigned* hcnt,
                       long long i0, long long i1) {
  const int lane = threadIdx.x & 31;
  FoldPhase f;
  f.i_ref = a.i_ref, f.rate = a.rate, f.inv_rate = a.inv_rate;
  f.c0 = a.coef[0], f.c1 = a.coef[1], f.c2 = a.coef[2], f.c3 = a.coef[3];
  f.n_phase = (double)a.n_phase;
  float acc = 0.f;
  unsigned acc_n = 0;
  int cur = -1;
  constexpr long long step = (long long)kFoldThreads * kFoldUnroll;
  long long ib = i0 + (threadIdx.x - lane);
  // Sample indices as doubles: integers below 2^53 add exactly.
  double xd = (double)(a.i_phase + ib + lane);
  for (; ib + step - kFoldThreads + 32 <= i1; ib += step, xd += (double)step)
    fold_four_tile<POWER, false>(a, f, hist, hcnt, ib, i1, xd, lane, acc, acc_n, cur);
  if (ib < i1)
    fold_four_tile<POWER, true>(a, f, hist, hcnt, ib, i1, xd, lane, acc, acc_n, cur);
  if (cur >= 0 && (lane & 7) == 0) {
    atomicAdd(hist + cur * 4 + (lane >> 3), acc);
    if (lane == 0) atomicAdd(hcnt + cur, acc_n);
  }
}

#if BBT_FOLD_TMA
// The same with TMA-staged tiles: lane 0 of every warp keeps kFoldStages bulk
// copies (cp.async.bulk, 2 KB = 128 samples each, completion counted on an
// mbarrier) in flight into the warp's own ring in shared memory, so the loads
// in flight are bounded neither by registers nor by the other warps; the
// warps of a CTA take consecutive 128-sample pieces in turn.
BBT_DEV unsigned smem_addr(const void* p) {
  return (unsigned)__cvta_generic_to_shared(p);
}
BBT_DEV void mbar_wait(unsigned bar, unsigned parity) {
  unsigned ok = 0;
  for (unsigned spins = 0; !ok; ++spins) {
    asm volatile(
        "{\n.reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (spins > (1u << 28)) __trap();  // a lost copy must not hang the GPU
  }
}

template <bool POWER>
BBT_DEV void fold_four_staged(const FoldArgs& a, float* hist, unsigned* hcnt,
                              long long i0, long long i1, float4* ring,
                              unsigned long long* bars) {
  constexpr int WT = 32 * kFoldUnroll;           // samples per warp tile
  constexpr int NW = kFoldThreads / 32;          // warps per CTA
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  FoldPhase f;
  f.i_ref = a.i_ref, f.rate = a.rate, f.inv_rate = a.inv_rate;
  f.c0 = a.coef[0], f.c1 = a.coef[1], f.c2 = a.coef[2], f.c3 = a.coef[3];
  f.n_phase = (double)a.n_phase;
  const long long n = i1 - i0;
  const long long nt = (n + WT - 1) / WT;        // warp tiles in this chunk
  const float4* src = static_cast<const float4*>(a.in) + i0;
  float4* wring = ring + (size_t)w * kFoldStages * WT;
  unsigned long long* wbars = bars + w * kFoldStages;
  if (lane == 0) {
    for (int s = 0; s < kFoldStages; ++s)
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(
          smem_addr(wbars + s)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  auto issue = [&](long long kt, int s) {  // lane 0: start the copy of tile kt
    const long long first = kt * WT;
    const long long left = n - first;
    const unsigned bytes = (unsigned)(left < WT ? left : WT) * 16u;
    const unsigned bar = smem_addr(wbars + s);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                 "r"(bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1], %2, [%3];" ::"r"(smem_addr(wring + (size_t)s * WT)),
        "l"(src + first), "r"(bytes), "r"(bar)
        : "memory");
  };
  if (lane == 0)
    for (int s = 0; s < kFoldStages; ++s)
      if (w + (long long)s * NW < nt) issue(w + (long long)s * NW, s);
  float acc = 0.f;
  unsigned acc_n = 0;
  int cur = -1;
  int j = 0;
  for (long long kt = w; kt < nt; kt += NW, ++j) {
    const int s = j % kFoldStages;
    mbar_wait(smem_addr(wbars + s), (unsigned)(j / kFoldStages) & 1u);
    const float4* tile = wring + (size_t)s * WT;
    const long long first = i0 + kt * WT;
    const bool whole = first + WT <= i1;
    f4 x[kFoldUnroll];
#pragma unroll
    for (int u = 0; u < kFoldUnroll; ++u) {
      const int idx = u * 32 + lane;
      x[u].x = x[u].y = x[u].z = x[u].w = 0.f;
      if (whole || first + idx < i1) {
        const float4 q = tile[idx];
        x[u].x = q.x, x[u].y = q.y, x[u].z = q.z, x[u].w = q.w;
      }
    }
    __syncwarp();  // the warp has taken its samples: the stage is free
    const long long knext = kt + (long long)kFoldStages * NW;
    if (lane == 0 && knext < nt) issue(knext, s);
    const double xd = (double)(a.i_phase + first + lane);
    if (whole)
      fold_four_bins<POWER, false, 32>(a, f, hist, hcnt, x, first, i1, xd, lane,
                                       acc, acc_n, cur);
    else
      fold_four_bins<POWER, true, 32>(a, f, hist, hcnt, x, first, i1, xd, lane,
                                      acc, acc_n, cur);
  }
  if (cur >= 0 && (lane & 7) == 0) {
    atomicAdd(hist + cur * 4 + (lane >> 3), acc);
    if (lane == 0) atomicAdd(hcnt + cur, acc_n);
  }
}
#endif  // BBT_FOLD_TMA
#endif

template <bool POWER>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(kFoldThreads, 4) fold_kernel(FoldArgs a) {
  float* hist = BBT_SMEM(float);
  unsigned* hcnt = reinterpret_cast<unsigned*>(hist + (size_t)a.n_phase * a.inner);
  const long long b = a.b_first + blockIdx.y;
  long long lo = a.lo[b] - a.i_first, hi = a.hi[b] - a.i_first;
  if (lo < 0) lo = 0;
  if (hi > a.n) hi = a.n;
  // This CTA's share of the bin.
  const long long per = (hi - lo + gridDim.x - 1) / (long long)gridDim.x;
  const long long i0 = lo + per * blockIdx.x;
  const long long i1 = (i0 + per < hi) ? i0 + per : hi;
  const int nh = a.n_phase * (int)a.inner;
  if (a.use_smem) {
    for (int q = threadIdx.x; q < nh; q += blockDim.x) hist[q] = 0.f;
    for (int q = threadIdx.x; q < a.n_phase; q += blockDim.x) hcnt[q] = 0u;
    BBT_SYNC();
  }
  float* gsum = a.sum + b * (long long)nh;
  unsigned long long* gcnt = a.count + b * a.n_phase;
  const long long width = POWER ? a.inner / 4 : a.inner;  // input items per sample
#if defined(__CUDA_ARCH__)
  if (a.use_smem >= 2) {
#if BBT_FOLD_TMA
    if (a.use_smem == 3) {
      // Ring and barriers follow the profile (16-byte aligned by the launcher).
      char* base = reinterpret_cast<char*>(hist) + a.ring_offset;
      fold_four_staged<POWER>(
          a, hist, hcnt, i0, i1, reinterpret_cast<float4*>(base),
          reinterpret_cast<unsigned long long*>(
              base + (size_t)kFoldStages * kFoldTile * 16));  // after the rings
    } else
#endif
    {
      fold_four<POWER>(a, hist, hcnt, i0, i1);
    }
  } else if (a.use_smem && a.inner <= kFoldFast) {
    const unsigned full = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    float acc[kFoldFast];
#pragma unroll
    for (int c = 0; c < kFoldFast; ++c) acc[c] = 0.f;
    unsigned acc_n = 0;
    int cur = -1;
    const int inner = (int)a.inner;
    // All threads of a warp run the same number of iterations.
    for (long long ib = i0 + (threadIdx.x - lane); ib < i1; ib += blockDim.x) {
      const long long i = ib + lane;
      const bool ok = i < i1;
      int p = -1;
      float x[kFoldFast];
#pragma unroll
      for (int c = 0; c < kFoldFast; ++c) x[c] = 0.f;
      if (ok) {
        p = a.pbin ? a.pbin[i] : fold_phase_bin(a, a.i_phase + i);
        if (p < 0) p = 0;
        if (p >= a.n_phase) p = a.n_phase - 1;
        if (POWER) {
          const cf2* src = static_cast<const cf2*>(a.in) + i * width;
#pragma unroll
          for (int m = 0; m < kFoldFast / 4; ++m) {
            if (m < width) {
              const f4 q = stokes_like(src[m].a, src[m].b);
              x[4 * m] = q.x;
              x[4 * m + 1] = q.y;
              x[4 * m + 2] = q.z;
              x[4 * m + 3] = q.w;
            }
          }
        } else {
          const float* src = static_cast<const float*>(a.in) + i * width;
#pragma unroll
          for (int c = 0; c < kFoldFast; ++c)
            if (c < inner) x[c] = src[c];
        }
      }
      const int p0 = __shfl_sync(full, p, 0);
      if (__all_sync(full, p == p0)) {
        // One bin for the whole warp (p0 >= 0 since lane 0 is in range).
#pragma unroll
        for (int c = 0; c < kFoldFast; ++c) {
          if (c < inner) {
            float r = x[c];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(full, r, o);
            x[c] = r;
          }
        }
        if (lane == 0) {
          if (p0 != cur) {
            if (cur >= 0) {
              for (int c = 0; c < inner; ++c)
                atomicAdd(hist + cur * inner + c, acc[c]);
              atomicAdd(hcnt + cur, acc_n);
            }
            cur = p0;
            acc_n = 0;
#pragma unroll
            for (int c = 0; c < kFoldFast; ++c) acc[c] = 0.f;
          }
#pragma unroll
          for (int c = 0; c < kFoldFast; ++c) acc[c] += x[c];
          acc_n += 32;
        }
      } else if (ok) {
#pragma unroll
        for (int c = 0; c < kFoldFast; ++c)
          if (c < inner) atomicAdd(hist + p * inner + c, x[c]);
        atomicAdd(hcnt + p, 1u);
      }
    }
    if (lane == 0 && cur >= 0) {
      for (int c = 0; c < inner; ++c) atomicAdd(hist + cur * inner + c, acc[c]);
      atomicAdd(hcnt + cur, acc_n);
    }
  } else
#endif
  {
    for (long long i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
      int p = a.pbin ? a.pbin[i] : fold_phase_bin(a, a.i_phase + i);
      if (p < 0) p = 0;
      if (p >= a.n_phase) p = a.n_phase - 1;
      float* dsum = (a.use_smem ? hist : gsum) + (long long)p * a.inner;
      if (POWER) {
        const cf2* x = static_cast<const cf2*>(a.in) + i * width;
        for (long long m = 0; m < width; ++m) {
          const f4 q = stokes_like(x[m].a, x[m].b);
          atomic_add(dsum + 4 * m + 0, q.x);
          atomic_add(dsum + 4 * m + 1, q.y);
          atomic_add(dsum + 4 * m + 2, q.z);
          atomic_add(dsum + 4 * m + 3, q.w);
        }
      } else {
        const float* x = static_cast<const float*>(a.in) + i * width;
        for (long long c = 0; c < width; ++c) atomic_add(dsum + c, x[c]);
      }
      if (a.use_smem) {
#if defined(BBT_EMULATE)
        __atomic_fetch_add(hcnt + p, 1u, __ATOMIC_RELAXED);
#else
        atomicAdd(hcnt + p, 1u);
#endif
      } else {
        atomic_add(gcnt + p, 1ull);
      }
    }
  }
  if (a.use_smem) {
    BBT_SYNC();
    for (int q = threadIdx.x; q < nh; q += blockDim.x)
      if (hist[q] != 0.f) atomic_add(gsum + q, hist[q]);
    for (int q = threadIdx.x; q < a.n_phase; q += blockDim.x)
      if (hcnt[q]) atomic_add(gcnt + q, (unsigned long long)hcnt[q]);
  }
}

// ---------------------------------------------------------------------------
// Polyphase filter bank, FIR and FFT fused (pfb.py:91-100 then
// channelize.py:73-74): spectrum j of column c is
//   FFT_n( sum_{tap < n_tap} h[tap][i] * x[((j + tap) n + i)][c] )_i .
// Lanes are (spectrum, column) pairs with the column fastest; REAL input keeps
// the n/2+1 non-negative frequencies (fourier/numpy.py:41-43).
struct PfbArgs {
  const void* in;     // [(n_spec + n_tap - 1) * n][inner] int8, float or complex
  cf* out;            // [n_spec][n_chan][inner]
  const float* h;     // [n_tap][n]
  const cf* tw;
  long long inner, n_spec;
  int n_tap;
};

// KIND: 0 complex64, 1 float32, 2 int8 (raw 8-bit samples, real).
template <class C, int KIND>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS, 1) pfb_kernel(PfbArgs a) {
  constexpr bool REAL = KIND != 0;
  cf* smem = BBT_SMEM(cf);
  const int tid = threadIdx.x;
  const int g = tid % C::G, t = tid / C::G;
  const long long lane = (long long)blockIdx.x * C::G + g;
  const long long j = lane / a.inner, c = lane % a.inner;
  const bool valid = j < a.n_spec;
  cf v[C::E];
#pragma unroll
  for (int e = 0; e < C::E; ++e) v[e] = mk(0.f, 0.f);
  if (valid) {
    for (int tap = 0; tap < a.n_tap; ++tap) {
      const float* h = a.h + (long long)tap * C::N;
      const long long base = ((j + tap) * C::N) * a.inner + c;
#pragma unroll
      for (int e = 0; e < C::E; ++e) {
        const int i = t + C::T * e;
        const float w = BBT_LDGF(h + i);
        if (KIND == 2) {
          v[e].x += w * (float)static_cast<const signed char*>(
                            a.in)[base + i * a.inner];
        } else if (KIND == 1) {
          v[e].x += w * static_cast<const float*>(a.in)[base + i * a.inner];
        } else {
          const cf x = static_cast<const cf*>(a.in)[base + i * a.inner];
          v[e].x += w * x.x;
          v[e].y += w * x.y;
        }
      }
    }
  }
  SmemLaneFast<C::PADSHIFT> sm{smem, g, C::G};
  block_fft<C>(v, t, a.tw, sm);
  if (valid) {
    const long long n_chan = REAL ? C::N / 2 + 1 : C::N;
    cf* dst = a.out + (j * n_chan) * a.inner + c;
#pragma unroll
    for (int e = 0; e < C::E; ++e) {
      const int k = t + C::T * e;
      if (!REAL || k <= C::N / 2) dst[k * a.inner] = v[e];
    }
  }
}

// Real input with an even number of columns: two adjacent columns (for a
// baseband stream, the two polarizations) are filtered and transformed as ONE
// complex series z = x_a + i x_b -- the pair of 8-bit samples or floats in
// memory IS that complex number -- and the two spectra are taken apart after
// the transform (A[k] = (Z[k] + conj Z[n-k]) / 2, B[k] = (Z[k] - conj Z[n-k])
// / 2i), which costs one more pass through shared memory but half the
// transforms, half the filter loads and half the exchanges of pfb_kernel.
// Lanes are (spectrum, column pair) with the pair fastest; the threads of a
// lane are consecutive (SmemLaneSlow), so a warp reads 32 adjacent sample
// pairs and stores 32 adjacent channels of both columns (512 bytes).
template <class C, int KIND>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS,
                                  (C::THREADS <= 512 ? 512 / C::THREADS : 1))
    pfb_pair_kernel(PfbArgs a) {
  static_assert(KIND == 1 || KIND == 2, "real input");
  cf* smem = BBT_SMEM(cf);
  const int tid = threadIdx.x;
  const int t = tid % C::T, g = tid / C::T;
  const long long pairs = a.inner / 2;
  const long long lane = (long long)blockIdx.x * C::G + g;
  const long long j = lane / pairs, cp = lane - j * pairs;
  const bool valid = j < a.n_spec;
  cf v[C::E];
#pragma unroll
  for (int e = 0; e < C::E; ++e) v[e] = mk(0.f, 0.f);
  if (valid) {
    for (int tap = 0; tap < a.n_tap; ++tap) {
      const float* h = a.h + (long long)tap * C::N + t;
      const long long base = ((j + tap) * C::N + t) * a.inner + 2 * cp;
      const long long step = (long long)C::T * a.inner;
#pragma unroll
      for (int e = 0; e < C::E; ++e) {
        const float w = BBT_LDGF(h + C::T * e);
        if (KIND == 2) {
          const short two = *reinterpret_cast<const short*>(
              static_cast<const signed char*>(a.in) + base + e * step);
          v[e].x = fmaf(w, (float)(signed char)(two & 0xff), v[e].x);
          v[e].y = fmaf(w, (float)(signed char)(two >> 8), v[e].y);
        } else {
          const cf x = *reinterpret_cast<const cf*>(
              static_cast<const float*>(a.in) + base + e * step);
          v[e].x = fmaf(w, x.x, v[e].x);
          v[e].y = fmaf(w, x.y, v[e].y);
        }
      }
    }
  }
  cf* z = smem + (size_t)g * C::NPAD;
  SmemLaneSlow<C::PADSHIFT> sm{z};
  block_fft<C>(v, t, a.tw, sm);
#pragma unroll
  for (int e = 0; e < C::E; ++e) z[t + C::T * e] = v[e];
  BBT_SYNC();
  if (valid) {
    const long long n_chan = C::N / 2 + 1;
    f4* dst = reinterpret_cast<f4*>(a.out + (j * n_chan) * a.inner + 2 * cp);
#pragma unroll
    for (int r = 0; r <= C::E / 2; ++r) {
      const int k = t + C::T * r;
      if (k <= C::N / 2) {
        const cf zk = z[k], zm = z[(C::N - k) & (C::N - 1)];
        f4 q;
        q.x = 0.5f * (zk.x + zm.x);
        q.y = 0.5f * (zk.y - zm.y);
        q.z = 0.5f * (zk.y + zm.y);
        q.w = 0.5f * (zm.x - zk.x);
        dst[k * pairs] = q;
      }
    }
  }
}

// The same for real input (float32 or raw int8) with the FIR done as a
// separate, vectorised phase: all threads of the CTA walk over the tile's
// G n values four at a time (one 32-bit load of four 8-bit samples, or one
// 128-bit load of four floats, per tap -- coalesced, where pfb_kernel issues
// one byte load per thread, tap and value), and leave the filtered block in
// shared memory, in the buffer the transform then exchanges through; the
// transform threads pick their values up from there.  Needs G % inner == 0
// (a tile is G / inner whole spectra of all columns) and n inner % 4 == 0.
template <class C, int KIND>
BBT_GLOBAL void BBT_LAUNCH_BOUNDS(C::THREADS, (C::THREADS <= 256 ? 2 : 1))
    pfb_real_kernel(PfbArgs a) {
  static_assert(KIND == 1 || KIND == 2, "real input");
  cf* smem = BBT_SMEM(cf);
  float* y = reinterpret_cast<float*>(smem);
  const int tid = threadIdx.x;
  const int g = tid % C::G, t = tid / C::G;
  const int inner = (int)a.inner;
  const int J = C::G / inner;                       // spectra per tile
  const long long j0 = (long long)blockIdx.x * J;
  const int block = C::N * inner;                   // values per spectrum
  const int ys = block + 16;                        // pitch in y (bank shift)
  // Phase 1: y[jl][q] = sum_tap h[tap][q / inner] x[(j0 + jl + tap) n inner + q].
  for (int w = tid; w < J * block / 4; w += C::THREADS) {
    const int q4 = w * 4;
    const int jl = q4 / block, q = q4 - jl * block;
    if (j0 + jl >= a.n_spec) continue;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    // Filter coefficient of each of the four values (same for every tap).
    int hi[4];
    if (inner == 2) {
      hi[0] = hi[1] = q >> 1;
      hi[2] = hi[3] = (q >> 1) + 1;
    } else if (inner == 1) {
      hi[0] = q, hi[1] = q + 1, hi[2] = q + 2, hi[3] = q + 3;
    } else {
#pragma unroll
      for (int r = 0; r < 4; ++r) hi[r] = (q + r) / inner;
    }
    for (int tap = 0; tap < a.n_tap; ++tap) {
      const long long at = (j0 + jl + tap) * (long long)block + q;
      float x[4];
      if (KIND == 2) {
        const unsigned word =
            *reinterpret_cast<const unsigned*>(
                static_cast<const signed char*>(a.in) + at);
#pragma unroll
        for (int r = 0; r < 4; ++r)
          x[r] = (float)(signed char)((word >> (8 * r)) & 0xffu);
      } else {
        const f4 v = *reinterpret_cast<const f4*>(
            static_cast<const float*>(a.in) + at);
        x[0] = v.x, x[1] = v.y, x[2] = v.z, x[3] = v.w;
      }
      const float* h = a.h + (long long)tap * C::N;
#pragma unroll
      for (int r = 0; r < 4; ++r)
        acc[r] = fmaf(BBT_LDGF(h + hi[r]), x[r], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) y[jl * ys + q + r] = acc[r];
  }
  BBT_SYNC();
  // Phase 2: lanes are (spectrum, column) pairs, column fastest.
  const int jl = g / inner, c = g - jl * inner;
  const long long j = j0 + jl;
  const bool valid = j < a.n_spec;
  cf v[C::E];
#pragma unroll
  for (int e = 0; e < C::E; ++e)
    v[e] = mk(valid ? y[jl * ys + (t + C::T * e) * inner + c] : 0.f, 0.f);
  BBT_SYNC();  // y becomes the exchange buffer
  SmemLaneFast<C::PADSHIFT> sm{smem, g, C::G};
  block_fft<C>(v, t, a.tw, sm);
  if (valid) {
    const long long n_chan = C::N / 2 + 1;
    cf* dst = a.out + (j * n_chan) * a.inner + c;
#pragma unroll
    for (int e = 0; e < C::E; ++e) {
      const int k = t + C::T * e;
      if (k <= C::N / 2) dst[k * a.inner] = v[e];
    }
  }
}

// Integer sample shifts per series (sampling.py:380-425):
// out[i][s] = in[i + offset[s]][s] for items of 4 or 8 bytes.
template <typename T>
BBT_GLOBAL void shift_kernel(const T* BBT_RESTRICT in, T* BBT_RESTRICT out,
                             const long long* BBT_RESTRICT offset,
                             long long n_out, long long S) {
  const long long total = n_out * S;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const long long s = i % S;
    out[i] = in[i + offset[s] * S];
  }
}

// Real samples to complex (zero imaginary part) and back (real part): lets
// real-valued streams use the complex dedispersion kernels.
BBT_GLOBAL void real_to_complex_kernel(const float* BBT_RESTRICT in,
                                       cf* BBT_RESTRICT out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = mk(in[i], 0.f);
}
BBT_GLOBAL void complex_to_real_kernel(const cf* BBT_RESTRICT in,
                                       float* BBT_RESTRICT out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = in[i].x;
}

// Two real overlap-save frames as one complex frame: a real response (the
// Hermitian phase factor of a real-valued stream, i.e. rfft -> x factor ->
// irfft) convolves the real and the imaginary part separately, so frames 2p
// and 2p+1 of a real stream ride through the complex kernels together and
// come out as the real and imaginary part of the result.
//   pack:   z[p][n][s] = x[(2p) spf + n][s] + i x[(2p+1) spf + n][s], n < N
//           (zero where the stream of n_in samples has ended)
//   unpack: y[(2p) spf + m][s] = Re w[p][m][s], y[(2p+1) spf + m][s] = Im,
//           m < spf, for the n_frames frames that exist.
// (grid.y walks over the pairs, grid.x over the frame: no division per
// element.  VEC: spf * S and N * S even, so that two neighbouring values of
// either frame are one aligned 8-byte load and their pair one 16-byte store.)
template <bool VEC>
BBT_GLOBAL void pair_frames_kernel(const float* BBT_RESTRICT in,
                                   cf* BBT_RESTRICT out, long long n_in,
                                   long long spf, long long N, long long S,
                                   long long n_frames) {
  const long long per = N * S, limit = n_in * S;
  const long long n_pairs = (n_frames + 1) / 2;
  constexpr int W = VEC ? 2 : 1;
  for (long long p = blockIdx.y; p < n_pairs; p += gridDim.y) {
    const long long a0 = 2 * p * spf * S;
    const bool second = 2 * p + 1 < n_frames;
    const long long b0 = a0 + spf * S;
    cf* o = out + p * per;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * W;
         r < per; r += (long long)gridDim.x * blockDim.x * W) {
      if constexpr (VEC) {
        const long long a = a0 + r, b = b0 + r;
        cf xa = mk(0.f, 0.f), xb = mk(0.f, 0.f);
        if (a + 1 < limit) {
          xa = *reinterpret_cast<const cf*>(in + a);
        } else if (a < limit) {
          xa.x = in[a];
        }
        if (second) {
          if (b + 1 < limit) {
            xb = *reinterpret_cast<const cf*>(in + b);
          } else if (b < limit) {
            xb.x = in[b];
          }
        }
        cf2 z;
        z.a = mk(xa.x, xb.x);
        z.b = mk(xa.y, xb.y);
        *reinterpret_cast<cf2*>(o + r) = z;
      } else {
        const long long a = a0 + r, b = b0 + r;
        o[r] = mk(a < limit ? in[a] : 0.f, second && b < limit ? in[b] : 0.f);
      }
    }
  }
}
template <bool VEC>
BBT_GLOBAL void unpair_frames_kernel(const cf* BBT_RESTRICT in,
                                     float* BBT_RESTRICT out, long long spf,
                                     long long S, long long n_frames) {
  const long long per = spf * S, n_pairs = (n_frames + 1) / 2;
  constexpr int W = VEC ? 2 : 1;
  for (long long p = blockIdx.y; p < n_pairs; p += gridDim.y) {
    const cf* w = in + p * per;
    float* re = out + (2 * p) * per;
    float* im = re + per;
    const bool second = 2 * p + 1 < n_frames;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * W;
         r < per; r += (long long)gridDim.x * blockDim.x * W) {
      if constexpr (VEC) {
        const cf2 z = *reinterpret_cast<const cf2*>(w + r);
        *reinterpret_cast<cf*>(re + r) = mk(z.a.x, z.b.x);
        if (second) *reinterpret_cast<cf*>(im + r) = mk(z.a.y, z.b.y);
      } else {
        const cf z = w[r];
        re[r] = z.x;
        if (second) im[r] = z.y;
      }
    }
  }
}

// Packed payload decode: value v occupies bits [v*bps, (v+1)*bps) of the byte
// stream (first value in the least significant bits) and maps to levels[code].
// One thread per four output values, so stores are full float4 lines.
BBT_GLOBAL void decode_kernel(const unsigned char* BBT_RESTRICT in,
                              float* BBT_RESTRICT out,
                              const float* BBT_RESTRICT levels, long long n,
                              int bps) {
  float* lut = BBT_SMEM(float);
  for (int i = threadIdx.x; i < (1 << bps); i += blockDim.x) lut[i] = levels[i];
  BBT_SYNC();
  const unsigned mask = (1u << bps) - 1u;
  const long long n_bytes = (n * bps + 7) >> 3;
  const long long n4 = (n + 3) >> 2;
  for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < n4;
       q += (long long)gridDim.x * blockDim.x) {
    const long long bit = q * 4 * bps;
    const long long b0 = bit >> 3;
    const int sh = (int)(bit & 7);
    const int nb = (sh + 4 * bps + 7) >> 3;
    unsigned long long w = 0;
    for (int k = 0; k < nb; ++k)
      if (b0 + k < n_bytes) w |= (unsigned long long)in[b0 + k] << (8 * k);
    w >>= sh;
    float v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = lut[(unsigned)(w >> (k * bps)) & mask];
    if (q * 4 + 4 <= n) {
      float4 o;
      o.x = v[0], o.y = v[1], o.z = v[2], o.w = v[3];
      *reinterpret_cast<float4*>(out + q * 4) = o;
    } else {
      for (int k = 0; q * 4 + k < n; ++k) out[q * 4 + k] = v[k];
    }
  }
}

// out[b][c] = sum[b][c] / count[b]; 0/0 gives NaN like numpy's division.
BBT_GLOBAL void average_kernel(const float* BBT_RESTRICT sum,
                               const unsigned long long* BBT_RESTRICT count,
                               float* BBT_RESTRICT out, long long n_bins,
                               long long inner) {
  const long long total = n_bins * inner;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x)
    out[i] = sum[i] / (float)count[i / inner];
}

}  // namespace bbt

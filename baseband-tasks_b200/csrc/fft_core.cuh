// Shared-memory Stockham FFT building block for sm_100a.
//
// Replaces the np.fft.{fft,ifft} calls the reference's NumpyFFTMaker makes
// (baseband_tasks/fourier/numpy.py:33-49).  One FFT of N = 2^LOG2N points is
// computed by T = N/E cooperating threads, each holding E = 2^LOG2E (<= 32)
// complex values in registers.  The transform is split into the fewest
// radix-R stages with R <= E, as equal as possible; every stage is a
// register-resident butterfly, and between stages the values are exchanged
// through a padded, bank-conflict-free shared-memory buffer.  Thread t owns
// elements t + T*e (e < E) both before and after the transform, so global
// loads and stores of consecutive threads are coalesced and no bit-reversal
// pass exists.
//
// The header is plain C++ so that the same code can be compiled by g++ in the
// kernel-emulation test harness (tests/emu) as well as by nvcc.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BBT_HD __host__ __device__ __forceinline__
#define BBT_D __device__ __forceinline__
#else
#define BBT_HD inline
#define BBT_D inline
#endif

namespace bbt {

struct alignas(8) cf {
  float x, y;
};
struct alignas(16) cf2 {  // two adjacent series (e.g. both polarizations)
  cf a, b;
};

BBT_HD cf mk(float x, float y) { cf r; r.x = x; r.y = y; return r; }

// Complex arithmetic.  On sm_100a a complex value sits in an aligned register
// pair and is worked on by the packed FP32 instructions (FADD2 / FMUL2 /
// FFMA2: two lanes per instruction, half the issue slots of scalar code);
// ptxas turns the swapped, broadcast and half-negated operands written out
// below into operand modifiers (.LO_HI, .F32, .NP), so they cost nothing.
#ifndef BBT_F32X2
#define BBT_F32X2 1
#endif
#if defined(__CUDA_ARCH__) && BBT_F32X2
#define BBT_PACKED 1
typedef unsigned long long u64x;
BBT_D u64x p2(float x, float y) {
  u64x r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(x), "f"(y));
  return r;
}
BBT_D u64x p2(cf a) { return p2(a.x, a.y); }
BBT_D cf u2(u64x v) {
  cf r;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(r.x), "=f"(r.y) : "l"(v));
  return r;
}
BBT_D u64x add2(u64x a, u64x b) {
  u64x r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
BBT_D u64x sub2(u64x a, u64x b) {
  u64x r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
BBT_D u64x mul2(u64x a, u64x b) {
  u64x r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
BBT_D u64x fma2(u64x a, u64x b, u64x c) {
  u64x r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
BBT_D cf operator+(cf a, cf b) { return u2(add2(p2(a), p2(b))); }
BBT_D cf operator-(cf a, cf b) { return u2(sub2(p2(a), p2(b))); }
BBT_D cf cmul(cf a, cf b) {
  return u2(fma2(p2(-a.y, a.x), p2(b.y, b.y), mul2(p2(a), p2(b.x, b.x))));
}
BBT_D cf cmulc(cf a, cf b) {  // a * conj(b)
  return u2(fma2(p2(a.y, -a.x), p2(b.y, b.y), mul2(p2(a), p2(b.x, b.x))));
}
BBT_D cf cscale(cf a, float s) { return u2(mul2(p2(a), p2(s, s))); }
// a + (wr, wi) b
BBT_D cf cfma(cf a, cf b, float wr, float wi) {
  return u2(fma2(p2(-b.y, b.x), p2(wi, wi), fma2(p2(b), p2(wr, wr), p2(a))));
}
// s = a + w b and d = a - w b = 2 a - s for w = (wr, wi).
BBT_D void fma_pm(cf a, cf b, float wr, float wi, cf& s, cf& d) {
  const u64x ps = fma2(p2(-b.y, b.x), p2(wi, wi), fma2(p2(b), p2(wr, wr), p2(a)));
  s = u2(ps);
  d = u2(fma2(p2(a), p2(2.f, 2.f), p2(-s.x, -s.y)));
}
#else
BBT_HD cf operator+(cf a, cf b) { return mk(a.x + b.x, a.y + b.y); }
BBT_HD cf operator-(cf a, cf b) { return mk(a.x - b.x, a.y - b.y); }
BBT_HD cf cmul(cf a, cf b) {
  return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
BBT_HD cf cmulc(cf a, cf b) {  // a * conj(b)
  return mk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y);
}
BBT_HD cf cscale(cf a, float s) { return mk(a.x * s, a.y * s); }
BBT_HD cf cfma(cf a, cf b, float wr, float wi) {
  return mk(fmaf(wr, b.x, fmaf(-wi, b.y, a.x)), fmaf(wr, b.y, fmaf(wi, b.x, a.y)));
}
// s = a + w b and d = a - w b = 2 a - s for a constant w = (wr, wi): six
// fused multiply-adds instead of a complex multiply and two additions.
BBT_HD void fma_pm(cf a, cf b, float wr, float wi, cf& s, cf& d) {
  s.x = fmaf(wr, b.x, fmaf(-wi, b.y, a.x));
  s.y = fmaf(wr, b.y, fmaf(wi, b.x, a.y));
  d.x = fmaf(2.f, a.x, -s.x);
  d.y = fmaf(2.f, a.y, -s.y);
}
#endif
// b * (wr, wi) for a constant w.
BBT_HD cf cmulk(cf b, float wr, float wi) { return cmul(b, mk(wr, wi)); }
BBT_HD cf cconj(cf a) { return mk(a.x, -a.y); }
BBT_HD cf mul_mi(cf a) { return mk(a.y, -a.x); }  // a * (-i)
BBT_HD cf mul_pi(cf a) { return mk(-a.y, a.x); }  // a * (+i)

// ---------------------------------------------------------------------------
// Register butterflies: forward DFT (exp(-2 pi i nk/R)), natural order in and
// out, fully unrolled so everything stays in registers.
template <int R>
struct Dft;

template <>
struct Dft<1> {
  static BBT_HD void run(cf*) {}
};

template <>
struct Dft<2> {
  static BBT_HD void run(cf* v) {
    cf a = v[0], b = v[1];
    v[0] = a + b;
    v[1] = a - b;
  }
};

template <>
struct Dft<4> {
  static BBT_HD void run(cf* v) {
    cf s02 = v[0] + v[2], d02 = v[0] - v[2];
    cf s13 = v[1] + v[3], d13 = mul_mi(v[1] - v[3]);
    v[0] = s02 + s13;
    v[1] = d02 + d13;
    v[2] = s02 - s13;
    v[3] = d02 - d13;
  }
};

// R = R1*R2 with n = R2*n1 + n2 and k = k1 + R1*k2:
// X[k1 + R1 k2] = sum_n2 W_R^{n2 k1} W_R2^{n2 k2} sum_n1 x[R2 n1 + n2] W_R1^{n1 k1}
template <>
struct Dft<8> {
  static BBT_HD void run(cf* v) {
    const float h = 0.70710678118654752440f;
    cf a[4] = {v[0], v[2], v[4], v[6]};
    cf b[4] = {v[1], v[3], v[5], v[7]};
    Dft<4>::run(a);
    Dft<4>::run(b);
    // b[k1] *= W8^{k1}
    b[1] = cmulk(b[1], h, -h);
    b[2] = mul_mi(b[2]);
    b[3] = cmulk(b[3], -h, -h);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      v[k] = a[k] + b[k];
      v[k + 4] = a[k] - b[k];
    }
  }
};

template <>
struct Dft<16> {
  static BBT_HD void run(cf* v) {
    // cos/sin of pi/8 multiples.
    const float c1 = 0.92387953251128675613f, s1 = 0.38268343236508977173f;
    const float h = 0.70710678118654752440f;
    cf a[4][4];
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
#pragma unroll
      for (int n1 = 0; n1 < 4; ++n1) a[n2][n1] = v[4 * n1 + n2];
      Dft<4>::run(a[n2]);
    }
    // Second layer: X[k1 + 4 k2] = DFT4_n2( W16^{n2 k1} a[n2][k1] ), with the
    // twiddles of n2 = 2 and n2 = 3 fused into the first additions.
    // W16^m = (cos(m pi/8), -sin(m pi/8)).
    {
      cf b[4] = {a[0][0], a[1][0], a[2][0], a[3][0]};
      Dft<4>::run(b);
#pragma unroll
      for (int k2 = 0; k2 < 4; ++k2) v[4 * k2] = b[k2];
    }
    {  // k1 = 1: W16^1, W16^2, W16^3
      cf s02, d02, s13, d13;
      fma_pm(a[0][1], a[2][1], h, -h, s02, d02);
      const cf b1 = cmul(a[1][1], mk(c1, -s1));
      fma_pm(b1, a[3][1], s1, -c1, s13, d13);
      d13 = mul_mi(d13);
      v[1] = s02 + s13;
      v[5] = d02 + d13;
      v[9] = s02 - s13;
      v[13] = d02 - d13;
    }
    {  // k1 = 2: W16^2, W16^4 = -i, W16^6
      const cf b2 = mul_mi(a[2][2]);
      const cf s02 = a[0][2] + b2, d02 = a[0][2] - b2;
      const cf b1 = cmulk(a[1][2], h, -h);
      cf s13, d13;
      fma_pm(b1, a[3][2], -h, -h, s13, d13);
      d13 = mul_mi(d13);
      v[2] = s02 + s13;
      v[6] = d02 + d13;
      v[10] = s02 - s13;
      v[14] = d02 - d13;
    }
    {  // k1 = 3: W16^3, W16^6, W16^9
      cf s02, d02, s13, d13;
      fma_pm(a[0][3], a[2][3], -h, -h, s02, d02);
      const cf b1 = cmul(a[1][3], mk(s1, -c1));
      fma_pm(b1, a[3][3], -c1, s1, s13, d13);
      d13 = mul_mi(d13);
      v[3] = s02 + s13;
      v[7] = d02 + d13;
      v[11] = s02 - s13;
      v[15] = d02 - d13;
    }
  }
};

// 32 = 2 x 16: X[k] = E[k] + W32^k O[k], X[k+16] = E[k] - W32^k O[k].
template <>
struct Dft<32> {
  static BBT_HD void run(cf* v) {
    cf ev[16], od[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      ev[i] = v[2 * i];
      od[i] = v[2 * i + 1];
    }
    Dft<16>::run(ev);
    Dft<16>::run(od);
    // cos, sin of k pi/16, k = 0..15.
    const float c[16] = {1.f,
                         0.98078528040323044913f,
                         0.92387953251128675613f,
                         0.83146961230254523708f,
                         0.70710678118654752440f,
                         0.55557023301960222474f,
                         0.38268343236508977173f,
                         0.19509032201612826785f,
                         0.f,
                         -0.19509032201612826785f,
                         -0.38268343236508977173f,
                         -0.55557023301960222474f,
                         -0.70710678118654752440f,
                         -0.83146961230254523708f,
                         -0.92387953251128675613f,
                         -0.98078528040323044913f};
    const float s[16] = {0.f,
                         0.19509032201612826785f,
                         0.38268343236508977173f,
                         0.55557023301960222474f,
                         0.70710678118654752440f,
                         0.83146961230254523708f,
                         0.92387953251128675613f,
                         0.98078528040323044913f,
                         1.f,
                         0.98078528040323044913f,
                         0.92387953251128675613f,
                         0.83146961230254523708f,
                         0.70710678118654752440f,
                         0.55557023301960222474f,
                         0.38268343236508977173f,
                         0.19509032201612826785f};
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      if (k == 0) {
        v[0] = ev[0] + od[0];
        v[16] = ev[0] - od[0];
      } else if (k == 8) {
        const cf o = mul_mi(od[8]);
        v[8] = ev[8] + o;
        v[24] = ev[8] - o;
      } else {
        // X = e + w o and X' = e - w o = 2 e - X, w = (c, -s), as six fused
        // multiply-adds instead of a complex multiply and two additions.
        fma_pm(ev[k], od[k], c[k], -s[k], v[k], v[k + 16]);
      }
    }
  }
};

// ---------------------------------------------------------------------------
// Largest single-CTA transform, and the size of the low-order table of the
// large-N twiddle W_N^m.  Each block FFT of N points takes the table of the
// N-th roots of unity exp(-2 pi i m / N), m < N (compact, so that the few
// entries a stage needs stay in L1).
constexpr int kLog2TwiddleTable = 14;
constexpr int kTwiddleTable = 1 << kLog2TwiddleTable;

// Split of a 2^LOG2N transform into the fewest stages of radix <= 2^LOG2E,
// as equal as possible (larger radices first).
template <int LOG2N, int LOG2E>
struct StagePlan {
  static constexpr int NS = LOG2N == 0 ? 0 : (LOG2N + LOG2E - 1) / LOG2E;
  static BBT_HD constexpr int bits(int i) {
    return NS == 0 ? 0 : LOG2N / NS + (i < LOG2N % NS ? 1 : 0);
  }
  static BBT_HD constexpr int before(int i) {  // log2 of points combined so far
    int b = 0;
    for (int s = 0; s < i; ++s) b += bits(s);
    return b;
  }
};

// Configuration of a block FFT: N points, E elements per thread, THREADS
// threads per CTA, hence G = THREADS*E/N transforms ("lanes") per CTA.
template <int LOG2N_, int LOG2E_, int THREADS_>
struct FftCfg {
  static constexpr int LOG2N = LOG2N_;
  static constexpr int LOG2E = LOG2E_ < LOG2N ? LOG2E_ : LOG2N;
  static constexpr int N = 1 << LOG2N;
  static constexpr int E = 1 << LOG2E;  // elements per thread
  static constexpr int T = N / E;       // threads per FFT
  static constexpr int THREADS = THREADS_ > T ? THREADS_ : T;
  static constexpr int G = THREADS / T;
  using Plan = StagePlan<LOG2N, LOG2E>;
  // Padding of the exchange buffer: one slot per 2^PADSHIFT, matched to the
  // stride of the first stage's stores.
  static constexpr int PADSHIFT = Plan::NS > 0 && Plan::bits(0) > 3
                                      ? Plan::bits(0) : 4;
  static constexpr int NPAD = N + (N >> PADSHIFT);  // exchange slots per FFT
  static constexpr size_t SMEM_BYTES =
      Plan::NS > 1 ? (size_t)G * NPAD * sizeof(float) * 2 : 16;
  // Gathered twiddle tables (apply_twiddles_tab): three runs of Ns entries
  // for every stage after the first; offset of the stage that starts with
  // 2^LOG2NS points combined, and the total.
  template <int LOG2NS>
  static BBT_HD constexpr int twtab_offset() {
    int off = 0;
    for (int s = 1; s < Plan::NS; ++s) {
      if (Plan::before(s) == LOG2NS) return off;
      off += 3 << Plan::before(s);
    }
    return off;
  }
  static BBT_HD constexpr int twtab_size() {
    int off = 0;
    for (int s = 1; s < Plan::NS; ++s) off += 3 << Plan::before(s);
    return off > 0 ? off : 1;
  }
};

// Fill the gathered twiddle tables of a transform from its table of N-th
// roots of unity (all threads of the group take part; synchronise after).
template <class C>
BBT_HD void fill_twtab(cf* tab, const cf* tw, int tid, int nthreads) {
  using P = typename C::Plan;
  int off = 0;
  for (int s = 1; s < P::NS; ++s) {
    const int ns = 1 << P::before(s);
    const int shift = C::LOG2N - P::before(s) - P::bits(s);
    for (int i = tid; i < 3 * ns; i += nthreads) {
      const int which = i / ns, k = i - which * ns;
      const int mult = which == 0 ? 1 : (which == 1 ? 8 : 16);
      // Entries a stage of smaller radix never reads may fall outside the
      // table of roots: leave them alone.
      const long long idx = (long long)(k << shift) * mult;
      if (idx < C::N) tab[off + i] = tw[idx];
    }
    off += 3 * ns;
  }
}

#if defined(__CUDACC__) && defined(__CUDA_ARCH__)
#define BBT_SYNC() __syncthreads()
#define BBT_SYNCWARP() __syncwarp()
#elif defined(BBT_EMULATE)
}  // namespace bbt
void bbt_emu_syncthreads();
void bbt_emu_syncwarp();
namespace bbt {
#define BBT_SYNC() bbt_emu_syncthreads()
#define BBT_SYNCWARP() bbt_emu_syncwarp()
#else
#define BBT_SYNC()
#define BBT_SYNCWARP()
#endif

// Exchange-buffer addressing.  G FFTs ("lanes") share one CTA.
//  LaneFast: consecutive threads work on consecutive lanes (column tiles);
//  slot = padslot(p) * G + g.
//  LaneSlow: consecutive threads work on consecutive elements of one FFT;
//  slot = g * NPAD + padslot(p).
// ``slot(p)`` is the padded position of element p; ``ref(s)`` the storage
// of padded position s.  Offsets between the elements a thread touches in one
// stage are compile-time constants in padded positions (see fft_stage).
// ``sync()`` is the barrier between the writes and the reads of an exchange:
// the whole CTA, or only the warp when all threads of a transform sit in one.
template <int PADSHIFT>
struct SmemLaneFast {
  static constexpr bool kLaneFast = true;
  cf* base;
  int g, G;
  static BBT_HD int slot(int p) { return p + (p >> PADSHIFT); }
  BBT_HD cf& ref(int s) const { return base[s * G + g]; }
  BBT_HD cf& at(int p) const { return ref(slot(p)); }
  static BBT_HD void sync() { BBT_SYNC(); }
  static constexpr bool kTwTab = false;
  const cf* twtab = nullptr;
};
template <int PADSHIFT>
struct SmemLaneSlow {
  static constexpr bool kLaneFast = false;
  cf* base;  // already offset to this lane
  static BBT_HD int slot(int p) { return p + (p >> PADSHIFT); }
  BBT_HD cf& ref(int s) const { return base[s]; }
  BBT_HD cf& at(int p) const { return ref(slot(p)); }
  static BBT_HD void sync() { BBT_SYNC(); }
  static constexpr bool kTwTab = false;
  const cf* twtab = nullptr;
};
// A transform whose threads all belong to one warp (at most 32 x E points):
// exchanges need only a warp-level barrier, so the warps of a CTA run their
// transforms independently of one another.
template <int PADSHIFT>
struct SmemWarp {
  static constexpr bool kLaneFast = false;
  cf* base;  // this transform's private region
  static BBT_HD int slot(int p) { return p + (p >> PADSHIFT); }
  BBT_HD cf& ref(int s) const { return base[s]; }
  BBT_HD cf& at(int p) const { return ref(slot(p)); }
  static BBT_HD void sync() { BBT_SYNCWARP(); }
  // Stage twiddles from gathered tables in shared memory (fill_twtab).
  static constexpr bool kTwTab = true;
  const cf* twtab;
};

// Streaming access to data that is touched once: do not let it displace the
// twiddle tables in L1.
BBT_HD cf ld_stream(const cf* p) {
#if defined(__CUDA_ARCH__)
  float2 w = __ldcs(reinterpret_cast<const float2*>(p));
  return mk(w.x, w.y);
#else
  return *p;
#endif
}
#if defined(__CUDA_ARCH__)
#define BBT_LDGF(p) __ldg(p)
#else
#define BBT_LDGF(p) (*(p))
#endif
// Asynchronous 8-byte copy global -> shared (zeros when !valid), and the
// wait for all copies this thread has issued.
BBT_HD void cp_async8(cf* smem_dst, const cf* gmem_src, bool valid) {
#if defined(__CUDA_ARCH__)
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst),
               "l"(gmem_src), "r"(bytes));
#else
  *smem_dst = valid ? *gmem_src : mk(0.f, 0.f);
#endif
}
BBT_HD void cp_async_wait() {
#if defined(__CUDA_ARCH__)
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#endif
}
BBT_HD void prefetch_l2(const void* p) {
#if defined(__CUDA_ARCH__)
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}

// v[e] <- v[e] * base * step^e (MODE 0), or its conjugate (MODE 1), for
// e < 2^BITS, with pw[b] = step^(2^b): a linear phase ramp at one complex
// multiply per element and per power.  (A real scale goes into ``base``.)
template <int BITS, int MODE>
struct Ramp {
  static BBT_HD void run(cf* v, cf w, const cf* pw) {
    Ramp<BITS - 1, MODE>::run(v, w, pw);
    Ramp<BITS - 1, MODE>::run(v + (1 << (BITS - 1)), cmul(w, pw[BITS - 1]), pw);
  }
};
template <int MODE>
struct Ramp<0, MODE> {
  static BBT_HD void run(cf* v, cf w, const cf*) {
    cf r = cmul(v[0], w);
    v[0] = MODE ? mk(r.x, -r.y) : r;
  }
};

BBT_HD cf ldtw(const cf* tw, int i) {
#if defined(__CUDA_ARCH__)
  float2 w = __ldg(reinterpret_cast<const float2*>(tw) + i);
  return mk(w.x, w.y);
#else
  return tw[i];
#endif
}

// b[r] *= w^r for r < R: w = w1, w^8 = w8 and w^16 = w16 are given (looked up
// by the caller), w^2 and w^4 come from squaring w (BBT_TW_SQUARE >= 1) and
// the other powers are products of two of these.
template <int R>
BBT_HD void apply_twiddle_powers(cf* b, cf w1, cf w8, cf w16) {
  if constexpr (R >= 2) {
    cf w[8];  // w^1 .. w^7
    w[1] = w1;
    if constexpr (R >= 4) {
      w[2] = cmul(w[1], w[1]);
      w[3] = cmul(w[2], w[1]);
    }
    if constexpr (R >= 8) {
      w[4] = cmul(w[2], w[2]);
      w[5] = cmul(w[4], w[1]);
      w[6] = cmul(w[4], w[2]);
      w[7] = cmul(w[4], w[3]);
    }
    constexpr int LOW = R < 8 ? R : 8;
#pragma unroll
    for (int r = 1; r < LOW; ++r) b[r] = cmul(b[r], w[r]);
    if constexpr (R >= 16) {
      const cf hi = w8;
      b[8] = cmul(b[8], hi);
#pragma unroll
      for (int r = 1; r < 8; ++r) b[8 + r] = cmul(b[8 + r], cmul(hi, w[r]));
      if constexpr (R >= 32) {
        const cf hi2 = w16;
        b[16] = cmul(b[16], hi2);
#pragma unroll
        for (int r = 1; r < 8; ++r)
          b[16 + r] = cmul(b[16 + r], cmul(hi2, w[r]));
        cf hi3 = cmul(hi2, hi);
        b[24] = cmul(b[24], hi3);
#pragma unroll
        for (int r = 1; r < 8; ++r)
          b[24 + r] = cmul(b[24 + r], cmul(hi3, w[r]));
      }
    }
  }
}

#ifndef BBT_TW_SQUARE
// 1: powers of two of the twiddle by squaring instead of look-ups; 2: w^8 and
// w^16 looked up (less rounding error, two more loads; the default, see
// BBT_RAMP_SQUARE for the measurements); 0: all looked up.
#define BBT_TW_SQUARE 2
#endif
// b[r] *= w^r for r < R, with w = tw[kk]: powers of two are looked up, the
// others are products of two looked-up or derived values.
template <int R>
BBT_HD void apply_twiddles(cf* b, const cf* __restrict__ tw, int kk) {
#if BBT_TW_SQUARE == 2
  cf w1 = mk(1.f, 0.f), w8 = w1, w16 = w1;
  if constexpr (R >= 2) w1 = ldtw(tw, kk);
  if constexpr (R >= 16) w8 = ldtw(tw, 8 * kk);
  if constexpr (R >= 32) w16 = ldtw(tw, 16 * kk);
  apply_twiddle_powers<R>(b, w1, w8, w16);
#else
  if constexpr (R >= 2) {
    cf w[8];  // w^1 .. w^7
    w[1] = ldtw(tw, kk);
    if constexpr (R >= 4) {
      w[2] = BBT_TW_SQUARE ? cmul(w[1], w[1]) : ldtw(tw, 2 * kk);
      w[3] = cmul(w[2], w[1]);
    }
    if constexpr (R >= 8) {
      w[4] = BBT_TW_SQUARE ? cmul(w[2], w[2]) : ldtw(tw, 4 * kk);
      w[5] = cmul(w[4], w[1]);
      w[6] = cmul(w[4], w[2]);
      w[7] = cmul(w[4], w[3]);
    }
    constexpr int LOW = R < 8 ? R : 8;
#pragma unroll
    for (int r = 1; r < LOW; ++r) b[r] = cmul(b[r], w[r]);
    if constexpr (R >= 16) {
      cf hi = BBT_TW_SQUARE == 1 ? cmul(w[4], w[4]) : ldtw(tw, 8 * kk);
      b[8] = cmul(b[8], hi);
#pragma unroll
      for (int r = 1; r < 8; ++r) b[8 + r] = cmul(b[8 + r], cmul(hi, w[r]));
      if constexpr (R >= 32) {
        cf hi2 = BBT_TW_SQUARE == 1 ? cmul(hi, hi) : ldtw(tw, 16 * kk);
        b[16] = cmul(b[16], hi2);
#pragma unroll
        for (int r = 1; r < 8; ++r)
          b[16 + r] = cmul(b[16 + r], cmul(hi2, w[r]));
        cf hi3 = cmul(hi2, hi);
        b[24] = cmul(b[24], hi3);
#pragma unroll
        for (int r = 1; r < 8; ++r)
          b[24 + r] = cmul(b[24 + r], cmul(hi3, w[r]));
      }
    }
  }
#endif
}

// Twiddles of one stage from a table gathered per stage in shared memory
// (tab[which * Ns + k] = w_k^{1, 8, 16}): consecutive threads read consecutive
// entries, whereas the strided look-ups tw[8 k], tw[16 k] of consecutive k
// touch a different cache line each.
template <int R>
BBT_HD void apply_twiddles_tab(cf* b, const cf* tab, int ns, int k) {
  cf w1 = mk(1.f, 0.f), w8 = w1, w16 = w1;
  if constexpr (R >= 2) w1 = tab[k];
  if constexpr (R >= 16) w8 = tab[ns + k];
  if constexpr (R >= 32) w16 = tab[2 * ns + k];
  apply_twiddle_powers<R>(b, w1, w8, w16);
}

// One Stockham stage: Ns = 2^LOG2NS points already combined, radix 2^LOG2R.
template <class C, int LOG2NS, int LOG2R, class Smem>
BBT_HD void fft_stage(cf* v, int t, const cf* __restrict__ tw, const Smem& sm) {
  constexpr int R = 1 << LOG2R;
  constexpr int Ns = 1 << LOG2NS;
  constexpr int NB = C::E / R;  // butterflies per thread
#pragma unroll
  for (int q = 0; q < NB; ++q) {
    const int j = t + C::T * q;
    const int k = j & (Ns - 1);
    cf b[R];
#pragma unroll
    for (int r = 0; r < R; ++r) b[r] = v[q + r * NB];
    if constexpr (LOG2NS > 0) {
      // tw is the table of N-th roots of unity: exp(-2 pi i m / N), m < N.
      if constexpr (Smem::kTwTab)
        apply_twiddles_tab<R>(b, sm.twtab + C::template twtab_offset<LOG2NS>(),
                              Ns, k);
      else
        apply_twiddles<R>(b, tw, k << (C::LOG2N - LOG2NS - LOG2R));
    }
    Dft<R>::run(b);
    if constexpr ((1 << (LOG2NS + LOG2R)) == C::N) {
#pragma unroll
      for (int r = 0; r < R; ++r) v[q + r * NB] = b[r];
    } else {
      constexpr int PADN = 1 << C::PADSHIFT;
      const int p0 = ((j - k) << LOG2R) + k;
      if constexpr (Ns % PADN == 0) {
        // (p0 + r Ns) >> PADSHIFT = (p0 >> PADSHIFT) + r (Ns >> PADSHIFT).
        const int s0 = Smem::slot(p0);
#pragma unroll
        for (int r = 0; r < R; ++r) sm.ref(s0 + r * (Ns + Ns / PADN)) = b[r];
      } else if constexpr (R * Ns <= PADN) {
        // All R outputs fall in one padding block.
        const int s0 = Smem::slot(p0);
#pragma unroll
        for (int r = 0; r < R; ++r) sm.ref(s0 + r * Ns) = b[r];
      } else {
#pragma unroll
        for (int r = 0; r < R; ++r) sm.at(p0 + r * Ns) = b[r];
      }
    }
  }
  if constexpr ((1 << (LOG2NS + LOG2R)) != C::N) {
    constexpr int PADN = 1 << C::PADSHIFT;
    Smem::sync();
    if constexpr (C::T % PADN == 0) {
      const int s0 = Smem::slot(t);
#pragma unroll
      for (int e = 0; e < C::E; ++e)
        v[e] = sm.ref(s0 + e * (C::T + C::T / PADN));
    } else {
#pragma unroll
      for (int e = 0; e < C::E; ++e) v[e] = sm.at(t + C::T * e);
    }
    Smem::sync();
  }
}

template <class C, int STAGE, class Smem>
struct FftStages {
  static BBT_HD void run(cf* v, int t, const cf* __restrict__ tw,
                         const Smem& sm) {
    using P = typename C::Plan;
    fft_stage<C, P::before(STAGE), P::bits(STAGE), Smem>(v, t, tw, sm);
    if constexpr (STAGE + 1 < P::NS)
      FftStages<C, STAGE + 1, Smem>::run(v, t, tw, sm);
  }
};

// Forward FFT of the N values spread over T threads (v[e] <-> t + T*e).
// All threads of the CTA must call this together (it synchronises).
template <class C, class Smem>
BBT_HD void block_fft(cf* v, int t, const cf* __restrict__ tw, const Smem& sm) {
  if constexpr (C::Plan::NS > 0) FftStages<C, 0, Smem>::run(v, t, tw, sm);
}

// The same transform in two parts: `head` runs every stage that ends with an
// exchange through shared memory, `tail` the last one, which works in
// registers only.  Between the two the exchange buffer is free (every thread
// has passed the barrier after its last read), so a kernel can let an
// asynchronous copy of its next tile land there while the tail runs.
template <class C, int STAGE, int END, class Smem>
struct FftStageRange {
  static BBT_HD void run(cf* v, int t, const cf* __restrict__ tw,
                         const Smem& sm) {
    if constexpr (STAGE < END) {
      using P = typename C::Plan;
      fft_stage<C, P::before(STAGE), P::bits(STAGE), Smem>(v, t, tw, sm);
      FftStageRange<C, STAGE + 1, END, Smem>::run(v, t, tw, sm);
    }
  }
};
template <class C, class Smem>
BBT_HD void block_fft_head(cf* v, int t, const cf* __restrict__ tw,
                           const Smem& sm) {
  if constexpr (C::Plan::NS > 1)
    FftStageRange<C, 0, C::Plan::NS - 1, Smem>::run(v, t, tw, sm);
}
template <class C, class Smem>
BBT_HD void block_fft_tail(cf* v, int t, const cf* __restrict__ tw,
                           const Smem& sm) {
  if constexpr (C::Plan::NS > 0)
    FftStageRange<C, C::Plan::NS - 1, C::Plan::NS, Smem>::run(v, t, tw, sm);
}

}  // namespace bbt

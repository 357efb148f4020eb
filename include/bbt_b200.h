/* bbt_b200.h -- C ABI of the B200 baseband-tasks hot path.
 *
 * The reference (mhvk/baseband-tasks) is pure Python and has no FFI; this ABI
 * is the boundary a drop-in replacement of its hot path exports.  Each entry
 * point names the reference code it replaces (paths relative to
 * baseband_tasks/ in the reference tree).  The Python side binds it with
 * ctypes (baseband_tasks_b200/_cabi.py); INTEGRATION.md shows the stub.
 *
 * Conventions
 *  - every function returns 0 on success, a negative bbt_status on failure;
 *    bbt_last_error() gives a thread-local message; nothing throws or aborts;
 *  - all data pointers are DEVICE pointers owned by the caller (complex64 =
 *    interleaved float re,im; time-major, C-contiguous, as Base.read()
 *    delivers them, base.py:389-438) unless a parameter says "host";
 *  - the library allocates only opaque plan objects (twiddle/chirp tables),
 *    created and destroyed explicitly; plans are immutable after creation and
 *    bound to the device that was current at creation;
 *  - every launch takes a cudaStream_t (as void*) and never synchronises.
 */
#ifndef BBT_B200_H
#define BBT_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum bbt_status {
  BBT_OK = 0,
  BBT_EINVAL = -1,      /* bad argument (ValueError on the Python side) */
  BBT_EUNSUPPORTED = -2,/* size/shape outside what the kernels handle */
  BBT_ECUDA = -3,       /* CUDA runtime error */
  BBT_ENOMEM = -4
};

enum bbt_fft_kind { BBT_C2C = 0, BBT_R2C = 1, BBT_C2R = 2 };
enum bbt_fft_direction { BBT_FORWARD = 0, BBT_BACKWARD = 1 };

typedef struct bbt_fft_plan bbt_fft_plan;
typedef struct bbt_dedisperse_plan bbt_dedisperse_plan;

int bbt_version(void);
const char* bbt_last_error(void);

/* ---- FFT: replaces np.fft.{fft,ifft,rfft,irfft}(a, axis, norm) as called by
 * NumpyFFTBase (fourier/numpy.py:33-49) for the FFT objects FFTMakerBase.__call__
 * creates (fourier/base.py:262-311).  Data are [outer][n][inner]; the
 * transform runs along the middle axis.  Any n >= 2: powers of two up to
 * 16384 run in one kernel, larger ones in four steps, other lengths through
 * Bluestein's algorithm on a power-of-two length; real transforms beyond the
 * single-kernel sizes go through the complex transform.  All but the
 * single-kernel case need a work buffer of bbt_fft_plan_work_bytes().
 * scale multiplies the output (1, 1/n or 1/sqrt(n): fourier/base.py:95-104). */
int bbt_fft_plan_create(bbt_fft_plan** plan, int64_t n, int64_t outer,
                        int64_t inner, int kind, int direction, double scale);
int64_t bbt_fft_plan_work_bytes(const bbt_fft_plan* plan);
int bbt_fft_exec(const bbt_fft_plan* plan, const void* in, void* out,
                 void* work, void* stream);
int bbt_fft_plan_destroy(bbt_fft_plan* plan);

/* ---- Coherent (de)dispersion: replaces Disperse.task, i.e.
 * ifft(fft(x) * phase_factor)[pad_start : pad_start + n_valid]
 * (dispersion.py:115-139) over whole runs of overlap-save frames
 * (base.py:775-795).  Frames are [n][n_series] complex64.
 * The chirp is generated on the device in float64 from
 *   phase = K dm F (1/f_ref - 1/F)^2 1e6 sideband + sample_offset/rate fftfreq,
 *   F = freq + fftfreq sideband      (dm.py:103-105, dispersion.py:117-126)
 * with one (freq, f_ref, sideband) triple per distinct chirp; series_map[s]
 * (host, n_series ints) says which chirp series s uses.  dm is the dispersing
 * DM (Dedisperse passes -dm, dispersion.py:184).  With freq_mhz == NULL no
 * chirp is generated (use bbt_dedisperse_plan_set_response).  hint = 0 lets
 * the library choose the three-pass split and layout; otherwise bits 0-7:
 * log2 of the column-FFT length, bit 8 / 9: force the planar / interleaved
 * work-buffer layout, bits 12 / 13: half-size (two per SM) tiles in the
 * column / row passes, bit 14: full-size row tiles for the interleaved
 * layout (tuning and tests).  Which kernel variants run (persistent kernels
 * fed by bulk / tensor-map copies, row formulations) is the library's choice;
 * bbt_tune_set can force them for A/B measurements. */
int bbt_dedisperse_plan_create(bbt_dedisperse_plan** plan, int64_t n,
                               int64_t n_series, int64_t pad_start,
                               int64_t n_valid, int64_t n_chirp,
                               const int32_t* series_map,
                               const double* freq_mhz, const double* fref_mhz,
                               const int8_t* sideband, double dm,
                               double rate_mhz, double sample_offset,
                               int hint);
/* Replace the chirp by an arbitrary response (host, [n_chirp][n] complex64,
 * natural FFT bin order): Convolve-style reuse (convolution.py:97-119). */
int bbt_dedisperse_plan_set_response(bbt_dedisperse_plan* plan,
                                     const void* host_response);
/* Copy the chirp back to the host in natural bin order ([n_chirp][n]). */
int bbt_dedisperse_plan_get_response(const bbt_dedisperse_plan* plan,
                                     void* host_response);
/* Bytes of scratch bbt_dedisperse_exec needs for a run of n_frames frames
 * (the three-pass work buffer plus the tile counters of the persistent
 * column passes; 0 for single-pass plans).  The buffer is the caller's, 16-byte
 * aligned, and must not be shared by calls that run concurrently. */
int64_t bbt_dedisperse_work_bytes(const bbt_dedisperse_plan* plan,
                                  int64_t n_frames);
/* Frame f reads in + f*in_frame_stride (complex elements) and writes its
 * n_valid - skip samples [pad_start+skip, pad_start+n_valid) to
 * out + f*out_frame_stride.  skip is PaddedTaskBase._frame_offset of a
 * re-anchored last frame (base.py:783-795); 0 otherwise. */
int bbt_dedisperse_exec(const bbt_dedisperse_plan* plan, const void* in,
                        int64_t in_frame_stride, int64_t n_frames,
                        int64_t skip, void* out, int64_t out_frame_stride,
                        void* work, void* stream);
/* The same with Power (functions.py:138-142) fused into the last pass, for
 * streams whose series come in polarization pairs (series 2q, 2q+1 = X, Y of
 * one channel, i.e. the polarization axis is the last one): instead of the
 * voltages X, Y the 16 bytes of a pair hold the four float32 products
 * [|X|^2, |Y|^2, Re(X conj Y), Im(X conj Y)].  `out` is thus the
 * (time, ..., 4) float32 output of Power(Dedisperse(...)) and the voltages
 * never go to memory.  bbt_dedisperse_power_supported tells (1 / 0) whether
 * the plan has this path (three-pass plans with an even number of series);
 * bbt_dedisperse_power_exec returns BBT_EUNSUPPORTED otherwise and the
 * caller runs bbt_dedisperse_exec and bbt_power_exec instead. */
int bbt_dedisperse_power_supported(const bbt_dedisperse_plan* plan);
int bbt_dedisperse_power_exec(const bbt_dedisperse_plan* plan, const void* in,
                              int64_t in_frame_stride, int64_t n_frames,
                              int64_t skip, void* out,
                              int64_t out_frame_stride, void* work,
                              void* stream);
int bbt_dedisperse_plan_destroy(bbt_dedisperse_plan* plan);

/* ---- Detection: replaces Power.task (functions.py:132-143) on (A, 2, B)
 * complex64 -> (A, 4, B) float32, and Square.task (functions.py:15-16,41). */
int bbt_power_exec(const void* in, void* out, int64_t a, int64_t b,
                   void* stream);
int bbt_square_exec(const void* in, void* out, int64_t n, int is_complex,
                    void* stream);
/* out[i] = a[i] * b[i] for n complex64 values: the ``ft *= phase_factor`` of
 * Disperse.task (dispersion.py:137) when the frame length is not one the
 * fused dedispersion plan takes (a power of two) and the three steps run
 * through the FFT plans. */
int bbt_multiply_exec(const void* a, const void* b, void* out, int64_t n,
                      void* stream);

/* ---- Fused Channelize(n) -> Power: replaces channelize.py:73-74 followed by
 * functions.py:132-143 for input [(n_spec*n)][m][2] complex64 -> output
 * [n_spec][n][m][4] float32. */
int bbt_channelize_power_exec(const void* in, void* out, int64_t n, int64_t m,
                              int64_t n_spec, void* stream);
/* ... -> Integrate: additionally integration.py:273-303.  offsets (device,
 * int64) are absolute bin edges in spectra; bins b_first .. b_first+n_bins-1
 * are accumulated (+=) into sum[bin][n][m][4] (float32) and count[bin]
 * (int64) for the part that overlaps spectra [j_first, j_first + n_spec).
 * With average != 0 every contribution is divided by the full width of its
 * bin, offsets[bin+1] - offsets[bin], so that sum ends up holding the averages
 * (the division of integration.py:268-269) without a separate pass. */
int bbt_channelize_power_integrate_exec(const void* in, int64_t n, int64_t m,
                                        int64_t n_spec, int64_t j_first,
                                        const int64_t* offsets,
                                        int64_t b_first, int64_t n_bins,
                                        void* sum, void* count, int average,
                                        void* stream);

/* ---- Integrate: replaces Integrate._integrate (integration.py:273-303) for
 * float32 input [n][inner]; same offsets/accumulate convention as above. */
int bbt_integrate_exec(const void* in, int64_t n, int64_t inner,
                       int64_t i_first, const int64_t* offsets,
                       int64_t b_first, int64_t n_bins, void* sum, void* count,
                       int average, void* stream);

/* ---- Fold: replaces Fold._integrate (integration.py:380-395).  Time bin b
 * covers absolute samples [lo[b], hi[b]) (device int64; the host applies the
 * reference's searchsorted convention).  Phase bins come either from pbin
 * (device int32 per sample of this call) or from the float64 polynomial
 * phase(i) = sum_k coef[k] ((i - i_ref)/rate)^k evaluated in float64 by
 * Horner's rule with individually rounded operations.  Here i counts samples
 * on the grid of the whole observation: the first sample of this call is
 * i_phase (an exact integer, so that the bins do not depend on how a stream
 * was cut into blocks or shared out over GPUs), and i_ref is the (possibly
 * fractional) index on that grid at which the polynomial's time is zero.  With power != 0 the input is
 * [n][inner/4][2] complex64 and the four polarization products are formed
 * on the fly.  sum[bin][n_phase][inner] float32 and count[bin][n_phase]
 * int64 are accumulated (+=). */
int bbt_fold_exec(const void* in, int power, int64_t n, int64_t inner,
                  int64_t i_first, int64_t i_phase, const int64_t* lo,
                  const int64_t* hi,
                  int64_t b_first, int64_t n_bins, const int32_t* pbin,
                  const double* coef, int ncoef, double i_ref, double rate,
                  int n_phase, void* sum, void* count, void* stream);

/* ---- Polyphase filter bank, FIR and FFT fused: replaces
 * PolyphaseFilterBankSamples.ppf (pfb.py:91-100; PolyphaseFilterBank.ppf
 * :145-154 computes the same in the Fourier domain) followed by
 * Channelize.task (channelize.py:73-74).  in: [(n_spec + n_tap - 1) * n][inner]
 * complex64 (is_real = 0), float32 (1) or raw int8 samples (2); response: device float32 [n_tap][n];
 * out: [n_spec][n_chan][inner] complex64 with n_chan = n/2+1 (real) or n. */
int bbt_pfb_exec(const void* in, void* out, const void* response, int64_t n,
                 int64_t n_tap, int64_t inner, int64_t n_spec, int is_real,
                 void* stream);

/* ---- Integer sample shifts: replaces ShiftSamples.task (sampling.py:380-425,
 * used by DisperseSamples/DedisperseSamples, dispersion.py:193-298):
 * out[i][s] = in[i + offset[s]][s] for i < n_out, with offset (device int64,
 * one per series, >= 0) and items of item_bytes = 4 (float32) or 8 (complex64). */
int bbt_shift_exec(const void* in, void* out, const int64_t* offset,
                   int64_t n_out, int64_t n_series, int item_bytes,
                   void* stream);

/* ---- Conversion between real float32 and complex64 samples (n values):
 * to_real == 0 sets the imaginary part to zero, to_real != 0 keeps the real
 * part.  Lets real-valued streams (rfft/irfft in the reference,
 * fourier/numpy.py:41-49) go through the complex dedispersion kernels with a
 * Hermitian response. */
int bbt_convert_exec(const void* in, void* out, int64_t n, int to_real,
                     void* stream);
/* Real-valued streams, two overlap-save frames per complex frame.  The
 * response of a real stream (rfft -> x phase factor -> irfft,
 * dispersion.py:115-139 with fourier/numpy.py:41-49) is a real convolution,
 * which acts on the real and the imaginary part of a complex series
 * separately; so frames 2p and 2p+1 (each n samples, samples_per_frame
 * apart, in a run of n_in real samples x n_series) are packed as
 * z[p] = x[2p] + i x[2p+1] into `out` ([(n_frames+1)/2][n][n_series]
 * complex64, zero beyond the end of the run), go through
 * bbt_dedisperse_exec as (n_frames+1)/2 frames with in_frame_stride =
 * n * n_series, and bbt_unpair_frames_exec writes the real and imaginary
 * parts of the [(n_frames+1)/2][samples_per_frame][n_series] result back as
 * the n_frames * samples_per_frame real output samples: half the transforms
 * and half the traffic of one complex frame per real frame. */
int bbt_pair_frames_exec(const void* in, void* out, int64_t n_in,
                         int64_t samples_per_frame, int64_t n, int64_t n_series,
                         int64_t n_frames, void* stream);
int bbt_unpair_frames_exec(const void* in, void* out,
                           int64_t samples_per_frame, int64_t n_series,
                           int64_t n_frames, void* stream);

/* ---- Packed payload decode (the VDIF-style encodings of baseband, which the
 * reference's coded HDF5 payloads reuse, io/hdf5/payload.py:165-166): value v
 * is the bps-bit code at bits [v*bps, (v+1)*bps) of the byte stream `in`
 * (first value in the least significant bits); out[v] = levels[code], with
 * levels a device table of 1 << bps float32.  bps in {1, 2, 4, 8}; complex
 * samples are (re, im) value pairs, so out can be viewed as complex64.
 * out must be 16-byte aligned. */
int bbt_decode_exec(const void* in, void* out, const float* levels, int64_t n,
                    int bps, void* stream);

/* ---- Averaging: out[b][c] = sum[b][c] / count[b] (NaN for empty bins),
 * the division Integrate._read_frame does (integration.py:268-269). */
int bbt_average_exec(const void* sum, const void* count, void* out,
                     int64_t n_bins, int64_t inner, void* stream);

/* ---- Measurement helpers (not part of the reference surface).
 * bbt_launch_count: kernels launched by this library so far.
 * bbt_profile_enable(1): bracket every launch with CUDA events on its stream;
 * bbt_profile_report: wait for them and write "kernel count total_ms" lines
 * (NUL-terminated) into buf, clearing the records. */
int64_t bbt_launch_count(void);
/* Force a kernel variant for A/B measurements (see DESIGN.md section 4 for
 * the keys); the environment variable BBT_TUNE="key=value,..." does the same
 * at start-up.  Results do not depend on the variant. */
int bbt_tune_set(const char* key, int value);
int bbt_profile_enable(int on);
int bbt_profile_report(char* buf, int64_t size);

/* ---- Measurement helper (not part of the reference surface): copies
 * `rows` chunks of chunk_bytes, row_stride_bytes apart, for each of n_tiles
 * column tiles -- the access pattern of the strided FFT passes. */
int bbt_strided_copy_bench(const void* in, void* out, int64_t rows,
                           int64_t row_stride_bytes, int64_t chunk_bytes,
                           int64_t n_tiles, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BBT_B200_H */

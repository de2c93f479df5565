/* b200dd.h -- C ABI of the B200-native delay-Doppler processor (libb200dd.so).
 *
 * This is the drop-in boundary for blah2's hot path: every entry point replaces the
 * arithmetic behind one public method of the reference's process classes, takes only
 * plain C types (pointers + sizes, no CUDA / torch types in signatures; a stream is an
 * opaque void* that may be NULL) and returns an integer status.  The C++ classes in
 * blah2_b200/dropin/ (same names, constructors and method signatures as the reference)
 * and the Python binding blah2_b200/capi.py are thin callers of this ABI.
 *
 * Reference interface each group replaces (paths relative to the blah2 repository):
 *   b200dd_caf_*     Ambiguity::Ambiguity / Ambiguity::process / getters
 *                    src/process/ambiguity/Ambiguity.h:34-58, Ambiguity.cpp:11-200
 *   b200dd_wh_*      WienerHopf::WienerHopf / WienerHopf::process
 *                    src/process/clutter/WienerHopf.h:68-78, WienerHopf.cpp:7-163
 *   b200dd_det_*     CfarDetector1D::process, Centroid::process, Interpolate::process,
 *                    Map::set_metrics
 *                    src/process/detection/CfarDetector1D.h:46-55, Centroid.h:35-44,
 *                    Interpolate.h:36-45, src/data/Map.cpp:188-206
 *   b200dd_spectrum_*   SpectrumAnalyser::SpectrumAnalyser / SpectrumAnalyser::process
 *                    src/process/spectrum/SpectrumAnalyser.h:48-57, SpectrumAnalyser.cpp:9-74
 *   b200dd_next_hamming   next_hamming, src/process/meta/HammingNumber.h:36
 *
 * Data conventions
 *   host IQ      interleaved complex128 (re, im doubles) -- what IqData holds
 *                (src/data/IqData.h:26).
 *   device IQ    interleaved complex64 (float2).  Every capture format of the reference
 *                (int16 RSPduo, fc32 USRP, int8 HackRF / Kraken) is exactly representable.
 *   map          row-major [nDopplerBins][nDelayBins] complex, row k <-> doppler[k],
 *                column j <-> delay[j]: the layout of Map<complex<double>>::data
 *                (src/data/Map.h:30); complex128 on the host, complex64 on the device.
 *   detections   three parallel double arrays delay (bins) / doppler (Hz) / snr (dB)
 *                (src/data/Detection.h:17-24), in the reference's emission order.
 *
 * Threading: a handle owns one CUDA stream and scratch buffers and is NOT re-entrant
 * (the reference's objects are not either; they are called from one thread,
 * src/blah2.cpp:245).  Different handles may be used from different threads.
 *
 * There is no CPU fallback: every *_process* entry point runs CUDA kernels for sm_100a
 * and fails with B200DD_ERR_CUDA when no usable device is present.
 *
 * Environment (all optional; defaults are the measured best on B200, profiles/r01_summary.md):
 *   B200DD_CAF_LOG2M / B200DD_CAF_PARTS      range-correlation FFT length (2^k) / CTAs per batch
 *   B200DD_CAF_GROUPS                        warp groups per range CTA, each on its own segments (1..4)
 *   B200DD_CAF_TMA=1                         stage IQ segments through shared memory with bulk async copies
 *   B200DD_WH_LOG2M, B200DD_WH_CORR_LOG2M, B200DD_WH_APPLY_LOG2M   WienerHopf FFT plans
 *   B200DD_WH_SOLVE_SPLIT=0                  one-barrier-per-step Toeplitz solve kernel also for <= 448 taps
 *   B200DD_WH_SOLVE_SHORT=0                  generic Toeplitz solve kernel also for <= 992 taps
 *   B200DD_PIPELINE_GRAPH=1 / 0              CUDA-graph replay of the device chain for every buffer triple / never
 *                                            (default: only triples given to b200dd_pipeline_prepare_device)
 * None of them changes results beyond the rounding of a different FFT factorisation.
 */
#ifndef B200DD_H
#define B200DD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200DD_API __attribute__((visibility("default")))
#else
#define B200DD_API
#endif

#define B200DD_OK 0
#define B200DD_ERR_ARG 1         /* null pointer / size mismatch */
#define B200DD_ERR_GEOMETRY 2    /* geometry outside the supported range (see DESIGN.md "limits") */
#define B200DD_ERR_CUDA 3        /* CUDA runtime error; b200dd_last_error() has the text */
#define B200DD_ERR_CAPACITY 4    /* output capacity too small (detections) */
#define B200DD_FILTER_FAILED 10  /* WienerHopf: matrix not positive definite -> y left untouched
                                    (reference returns false, WienerHopf.cpp:111-122) */

/* Text of the last error raised on the calling thread ("" if none). */
B200DD_API const char *b200dd_last_error(void);

/* Library / device probe: returns number of CUDA devices (0 when none), fills optional name. */
B200DD_API int b200dd_device_count(void);
B200DD_API int b200dd_device_name(int device, char *buf, int buflen);

/* next_hamming(value): first 5-smooth number strictly greater than value
 * (src/process/meta/HammingNumber.cpp:38-48). Host arithmetic. */
B200DD_API uint32_t b200dd_next_hamming(uint32_t value);

/* ------------------------------------------------------------------ CAF / Ambiguity */

typedef struct b200dd_caf b200dd_caf;

typedef struct {
  int32_t delay_min;     /* bins  (Ambiguity.h:34 delayMin) */
  int32_t delay_max;     /* bins */
  int32_t doppler_min;   /* Hz */
  int32_t doppler_max;   /* Hz */
  uint32_t fs;           /* Hz */
  uint32_t n_samples;    /* samples per CPI per channel */
  int32_t round_hamming; /* only affects get_nfft(); the GPU FFT length is independent */
  int32_t device;        /* CUDA device ordinal, -1 = current device */
} b200dd_caf_params;

typedef struct {
  uint32_t n_delay_bins;   /* Ambiguity::get_n_delay_bins   */
  uint32_t n_doppler_bins; /* Ambiguity::get_n_doppler_bins */
  uint32_t n_corr;         /* Ambiguity::get_n_corr         */
  uint32_t nfft;           /* Ambiguity::get_nfft           */
  uint32_t n_used;         /* nDopplerBins * nCorr: samples consumed per channel (Ambiguity.cpp:105) */
  double cpi;              /* Ambiguity::get_cpi            */
  double doppler_middle;   /* Ambiguity::get_doppler_middle */
  /* implementation facts (for DESIGN / bench reporting) */
  uint32_t range_fft_len;   /* M of the segmented range FFT */
  uint32_t range_segments;  /* segments per batch */
  uint32_t range_hop;       /* new samples per segment */
  uint32_t doppler_fft_len; /* Bluestein length M2 */
  uint32_t range_parts;     /* CTAs per batch (partial range matrices summed by the Doppler kernel) */
  uint32_t range_groups;    /* warp groups per range CTA, each transforming its own segments of the batch */
} b200dd_caf_geometry;

B200DD_API int b200dd_caf_create(const b200dd_caf_params *params, b200dd_caf **out);
/* The host half of b200dd_caf_create on its own -- the Ambiguity constructor's arithmetic (Ambiguity.cpp:11-66: bin
 * counts, nCorr, cpi, nfft, both axes) and the kernel plan -- without touching a device (params->device is only used
 * to look up the SM count; 148 is assumed when there is none).  delay / doppler are nullable; capacities in elements.
 * This is what lets the CPU test-suite check the PRODUCT's geometry code against the reference's golden values. */
B200DD_API int b200dd_caf_plan(const b200dd_caf_params *params, b200dd_caf_geometry *out, int32_t *delay,
                               uint32_t cap_delay, double *doppler, uint32_t cap_doppler);
B200DD_API void b200dd_caf_destroy(b200dd_caf *h);
B200DD_API int b200dd_caf_get_geometry(const b200dd_caf *h, b200dd_caf_geometry *out);
/* Map axes as the Ambiguity constructor builds them (Ambiguity.cpp:46-59):
 * delay[n_delay_bins] (bins), doppler[n_doppler_bins] (Hz). */
B200DD_API int b200dd_caf_get_axes(const b200dd_caf *h, int32_t *delay, double *doppler);

/* Ambiguity::process on HOST buffers: x, y = n interleaved complex128 samples each
 * (n >= n_used; only the first n_used are consumed, like the FIFO pops at
 * Ambiguity.cpp:108-112).  map_out = [n_doppler_bins][n_delay_bins] complex128.
 * H2D, conversion to float2, kernels and D2H all run on the handle's stream; the call
 * returns after the map is in map_out. */
B200DD_API int b200dd_caf_process_host(b200dd_caf *h, const double *x, const double *y, uint32_t n, double *map_out);

/* Device-resident variant: d_x, d_y = n float2 samples in device memory, d_map =
 * [n_doppler_bins][n_delay_bins] float2 in device memory.  Asynchronous on `stream`
 * (NULL = the handle's stream); no host synchronisation. */
B200DD_API int b200dd_caf_process_device(b200dd_caf *h, const void *d_x, const void *d_y, uint32_t n, void *d_map,
                              void *stream);

/* The two stages separately, for ONE large CPI sharded over several GPUs (BASELINE config 5):
 * range stage on a contiguous block of batches (the rank's 1/world slice of the IQ), then -- after the
 * caller has all-gathered the range matrix -- the Doppler stage on a tile of delay columns.
 *   d_x, d_y : float2 samples of batches [batch0, batch0 + n_batches), i.e. n_batches * n_corr samples
 *   d_R      : out, float2 [n_batches][n_delay_bins]  (rows batch0.. of the range matrix)
 *   d_R (doppler) : in, the COMPLETE range matrix float2 [n_doppler_bins][n_delay_bins]
 *   d_map_tile    : out, float2 [n_doppler_bins][n_cols] = map[:, col0 : col0 + n_cols]
 * Symmetric Doppler windows only (the pre-rotation phase would need the global sample index). */
B200DD_API int b200dd_caf_range_device(b200dd_caf *h, const void *d_x, const void *d_y, uint32_t batch0,
                                       uint32_t n_batches, void *d_R, void *stream);
B200DD_API int b200dd_caf_doppler_device(b200dd_caf *h, const void *d_R, uint32_t col0, uint32_t n_cols,
                                         void *d_map_tile, void *stream);
/* A tile [nDop][n_cols] (as b200dd_caf_doppler_device writes it, possibly gathered from another GPU) into the delay
 * columns [col0, col0 + n_cols) of a row-major [nDop][nDel] map: one strided device copy. */
B200DD_API int b200dd_caf_place_tile_device(b200dd_caf *h, const void *d_tile, uint32_t col0, uint32_t n_cols, void *d_map,
                                            void *stream);
/* All tiles of an EQUAL split of the delay columns over n_tiles ranks (tile t = columns [t b + min(t, r), ...) with
 * b = nDel / n_tiles, r = nDel % n_tiles: the first r tiles one column wider), stored back to back in d_tiles as the
 * gather delivers them, into the row-major map: one kernel. */
B200DD_API int b200dd_caf_place_tiles_device(b200dd_caf *h, const void *d_tiles, uint32_t n_tiles, void *d_map, void *stream);

/* Profiling aid: same as b200dd_caf_process_device but brackets the range-correlation kernel and the
 * Doppler kernel with CUDA events on the launching stream and returns their durations (ms).
 * Synchronises the stream.  Used by bench.py for the roofline figures. */
B200DD_API int b200dd_caf_profile_device(b200dd_caf *h, const void *d_x, const void *d_y, uint32_t n, void *d_map,
                                         void *stream, float *ms_range, float *ms_doppler);

/* Intermediate range matrix of the last process call (row i = batch i, nDelayBins lags;
 * Ambiguity.cpp:106-149) copied to host as complex64 -- parity tests only. */
B200DD_API int b200dd_caf_debug_range_matrix(b200dd_caf *h, float *out);

/* Device pointer to the handle's last map (float2 [nDop][nDel]) for chaining detection
 * without a host round trip; valid until the next process call on this handle. */
B200DD_API void *b200dd_caf_device_map(b200dd_caf *h);
/* Handle's stream (cudaStream_t as void*). */
B200DD_API void *b200dd_caf_stream(b200dd_caf *h);

/* ------------------------------------------------------------------ WienerHopf */

typedef struct b200dd_wh b200dd_wh;

/* WienerHopf(delayMin, delayMax, nSamples): nBins = delayMax - delayMin (no +1,
 * WienerHopf.cpp:12). device = -1 -> current device. */
B200DD_API int b200dd_wh_create(int32_t delay_min, int32_t delay_max, uint32_t n_samples, int32_t device, b200dd_wh **out);
B200DD_API void b200dd_wh_destroy(b200dd_wh *h);

/* WienerHopf::process on HOST buffers.  x: n_samples complex128 (read only);
 * y: n_samples complex128, overwritten with the filtered surveillance channel.
 * Returns B200DD_OK, or B200DD_FILTER_FAILED with y untouched. */
B200DD_API int b200dd_wh_process_host(b200dd_wh *h, const double *x, double *y);

/* Device-resident variant: d_x, d_y float2[n_samples]; d_y_out float2[n_samples] (may
 * alias d_y).  Asynchronous; the success flag is written to the handle's device status
 * word and, when the solve fails, d_y_out receives d_y unchanged.  Use
 * b200dd_wh_last_status() (synchronises the stream) to read it. */
B200DD_API int b200dd_wh_process_device(b200dd_wh *h, const void *d_x, const void *d_y, void *d_y_out, void *stream);
B200DD_API int b200dd_wh_last_status(b200dd_wh *h);
/* Same with complex128 device buffers (double2) in and out -- what the host path and the
 * pipeline's host entry use so that the filter sees the caller's doubles unrounded. */
B200DD_API int b200dd_wh_process_device_f64(b200dd_wh *h, const void *d_x, const void *d_y, void *d_y_out, void *stream);
/* Profiling aid: per-stage CUDA-event durations (ms) of correlation, solve, and weight-spectrum+filter. */
B200DD_API int b200dd_wh_profile_device(b200dd_wh *h, const void *d_x, const void *d_y, void *d_y_out, void *stream,
                                        float *ms_corr, float *ms_solve, float *ms_apply);
/* Device address of the status word (0 ok, 1 failed) for kernels chained after the filter. */
B200DD_API const int *b200dd_wh_device_status(b200dd_wh *h);
/* Filter weights w[nBins] / correlations a, b of the last call as complex128 -- parity tests. */
B200DD_API int b200dd_wh_debug_weights(b200dd_wh *h, double *w, double *a, double *b);
B200DD_API uint32_t b200dd_wh_n_bins(const b200dd_wh *h);

/* ---- one CPI split over several GPUs: this handle filters the samples
 * [chunk_begin, chunk_begin + chunk_len) of an n_samples-sample signal (delayMin <= 0, chunk longer than the
 * filter).  The caller holds LOCAL float2 buffers with halos (b200dd_wh_chunk_halos):
 *   x_loc: x[chunk_begin - x_left .. chunk_begin + chunk_len + x_right)   (indices taken modulo n_samples on the right:
 *          the correlations are circular; the left halo of the first chunk is never read: zero filter history)
 *   y_loc: y[chunk_begin .. chunk_begin + chunk_len + y_right)            (modulo n_samples as well)
 * 1. b200dd_wh_chunk_corr_device -> d_ab = this chunk's share of (a[0..nBins), b[0..nBins)) (2 nBins complex128);
 * 2. the caller sums d_ab over the GPUs (b200dd_comm_allreduce_f64_async, 4 nBins doubles);
 * 3. b200dd_wh_chunk_filter_device solves the SAME system on every GPU and writes the chunk's chunk_len filtered
 *    samples (float2) to d_y_out.  b200dd_wh_last_status reports a failed solve as for a whole signal. */
B200DD_API int b200dd_wh_create_chunk(int32_t delay_min, int32_t delay_max, uint32_t n_samples, uint32_t chunk_begin,
                                      uint32_t chunk_len, int32_t device, b200dd_wh **out);
B200DD_API int b200dd_wh_chunk_halos(const b200dd_wh *h, uint32_t *x_left, uint32_t *x_right, uint32_t *y_right);
B200DD_API int b200dd_wh_chunk_corr_device(b200dd_wh *h, const void *d_x_loc, const void *d_y_loc, void *d_ab, void *stream);
B200DD_API int b200dd_wh_chunk_filter_device(b200dd_wh *h, const void *d_ab, const void *d_x_loc, const void *d_y_loc,
                                             void *d_y_out, void *stream);

/* The FFT plan of the two FFT stages (for measurement: transforms per CPI = 3 corr_segments + 2 corr_ctas in the
 * correlation kernel, 2 filter_blocks + 1 in the filter stage). */
typedef struct {
  uint32_t corr_fft_len, corr_hop, corr_segments, corr_ctas;
  uint32_t filter_fft_len, filter_hop, filter_blocks;
} b200dd_wh_plan;
B200DD_API int b200dd_wh_get_plan(const b200dd_wh *h, b200dd_wh_plan *out);
B200DD_API void *b200dd_wh_stream(b200dd_wh *h);

/* ------------------------------------------------------------------ detection tail */

typedef struct b200dd_det b200dd_det;

typedef struct {
  double pfa;           /* CfarDetector1D(pfa, ...)          CfarDetector1D.h:46 */
  int32_t n_guard;      /* int8_t in the reference */
  int32_t n_train;      /* int8_t */
  int32_t min_delay;    /* int8_t */
  double min_doppler;
  uint32_t n_centroid_delay;   /* Centroid(nDelay, nDoppler, resolutionDoppler)  Centroid.h:35 */
  uint32_t n_centroid_doppler;
  double resolution_doppler;
  int32_t interp_delay;   /* Interpolate(doDelay, doDoppler)  Interpolate.h:36 */
  int32_t interp_doppler;
  int32_t device;
} b200dd_det_params;

B200DD_API int b200dd_det_create(const b200dd_det_params *params, uint32_t max_doppler_bins, uint32_t max_delay_bins,
                      b200dd_det **out);
B200DD_API void b200dd_det_destroy(b200dd_det *h);

/* Map::set_metrics on a device map (float2 [n_dop][n_del]): metrics[0] = noisePower,
 * metrics[1] = maxPower (Map.cpp:188-206).  Synchronises the stream. */
B200DD_API int b200dd_det_set_metrics_device(b200dd_det *h, const void *d_map, uint32_t n_dop, uint32_t n_del, double *metrics,
                                  void *stream);

/* Stage selectors for b200dd_det_process_*: run CFAR only, CFAR+Centroid, or all three. */
#define B200DD_DET_CFAR 1
#define B200DD_DET_CENTROID 2
#define B200DD_DET_INTERPOLATE 3

/* CfarDetector1D -> Centroid -> Interpolate on a device map (blah2.cpp:285-287).
 * delay[n_del] / doppler[n_dop] are the map axes (host), noise_power = Map::noisePower.
 * Outputs (host): up to `cap` detections; *n_out = number found (may exceed cap ->
 * B200DD_ERR_CAPACITY with the first cap filled). */
B200DD_API int b200dd_det_process_device(b200dd_det *h, int last_stage, const void *d_map, uint32_t n_dop, uint32_t n_del,
                              const int32_t *delay, const double *doppler, double noise_power, double *o_delay,
                              double *o_doppler, double *o_snr, uint32_t cap, uint32_t *n_out, void *stream);

/* Fully asynchronous variant for device-resident streams of CPIs: Map::set_metrics is evaluated
 * on the device first (blah2.cpp:279) and its noisePower feeds the detector without a host round
 * trip; nothing is copied back until b200dd_det_chain_fetch (which synchronises the stream).
 * last_stage = 0 runs Map::set_metrics only (detection disabled: the fetch then reports 0 detections).
 * A CFAR stage that finds more detections than the handle's list capacity (min(cells, 2^18)) makes the
 * fetch return B200DD_ERR_CAPACITY: the later stages saw a truncated list. */
B200DD_API int b200dd_det_chain_device_async(b200dd_det *h, int last_stage, const void *d_map, uint32_t n_dop,
                                             uint32_t n_del, const int32_t *delay, const double *doppler,
                                             void *stream);
B200DD_API int b200dd_det_chain_fetch(b200dd_det *h, double *metrics, double *o_delay, double *o_doppler,
                                      double *o_snr, uint32_t cap, uint32_t *n_out, void *stream);

/* Same on a HOST complex128 map (what the drop-in classes hold in Map::data). */
B200DD_API int b200dd_det_process_host(b200dd_det *h, int last_stage, const double *map, uint32_t n_dop, uint32_t n_del,
                            const int32_t *delay, const double *doppler, double noise_power, double *o_delay,
                            double *o_doppler, double *o_snr, uint32_t cap, uint32_t *n_out);

/* Centroid / Interpolate alone on a host detection list (class-level drop-ins). */
B200DD_API int b200dd_det_centroid_host(b200dd_det *h, const double *delay, const double *doppler, const double *snr,
                             uint32_t n, double *o_delay, double *o_doppler, double *o_snr, uint32_t cap,
                             uint32_t *n_out);
B200DD_API int b200dd_det_interpolate_host(b200dd_det *h, const double *delay, const double *doppler, const double *snr,
                                uint32_t n, const double *map, uint32_t n_dop, uint32_t n_del, const int32_t *mdelay,
                                const double *mdoppler, double noise_power, double *o_delay, double *o_doppler,
                                double *o_snr, uint32_t cap, uint32_t *n_out);

/* ------------------------------------------------------------------ SpectrumAnalyser
 *
 * SpectrumAnalyser(n, bandwidth) / SpectrumAnalyser::process(IqData *x)
 * (src/process/spectrum/SpectrumAnalyser.h:48-57, SpectrumAnalyser.cpp:9-74): the stage of the reference's
 * process thread that runs on the reference channel right before the clutter filter (src/blah2.cpp:263-265).
 * decimation = n / bandwidth, nSpectrum = n / decimation, nfft = nSpectrum * decimation (:16-18); the result is
 * every decimation-th bin of the fft-shifted nfft-point spectrum of the first nfft samples (:36-54), which the
 * reference stores with IqData::update_spectrum (:55).  The device kernels never form the nfft-point transform:
 * one HBM-bound folding pass over x and an nSpectrum-point DFT (blah2_b200/csrc/spectrum.cu). */
typedef struct b200dd_spectrum b200dd_spectrum;

typedef struct {
  uint32_t decimation;   /* SpectrumAnalyser.cpp:16 */
  uint32_t n_spectrum;   /* :17, number of complex bins process() produces */
  uint32_t nfft;         /* :18, samples of x consumed (read, not popped) */
  uint32_t n_frequency;  /* entries of the frequency vector the reference's loop (:57-67) produces: its uint32_t
                            counter starts at (2^32 - nSpectrum) / 2, so this is 0 for every realistic size */
  /* implementation facts */
  uint32_t fold_chunks;          /* row chunks of the folding pass (partial sums reduced in a fixed order) */
  uint32_t fold_rows_per_chunk;
} b200dd_spectrum_geometry;

/* Returns B200DD_ERR_GEOMETRY where the reference is undefined (bandwidth <= 0, NaN, or bandwidth > n, which
 * divides by zero at :17) or for more than 65536 spectrum bins. */
B200DD_API int b200dd_spectrum_create(uint32_t n, double bandwidth, int32_t device, b200dd_spectrum **out);
/* The host half of b200dd_spectrum_create on its own (no device needed): geometry and, when `frequency` is not NULL,
 * the n_frequency values of the frequency vector. */
B200DD_API int b200dd_spectrum_plan(uint32_t n, double bandwidth, b200dd_spectrum_geometry *out, double *frequency,
                                    uint32_t cap);
B200DD_API void b200dd_spectrum_destroy(b200dd_spectrum *h);
B200DD_API int b200dd_spectrum_get_geometry(const b200dd_spectrum *h, b200dd_spectrum_geometry *out);
/* The vector SpectrumAnalyser::process hands to IqData::update_frequency (n_frequency doubles, kHz). */
B200DD_API int b200dd_spectrum_get_frequency(const b200dd_spectrum *h, double *frequency, uint32_t cap);

/* SpectrumAnalyser::process on a HOST buffer: x = n >= nfft interleaved complex128 samples (only the first nfft
 * are read); spectrum_out = n_spectrum complex128.  Synchronous. */
B200DD_API int b200dd_spectrum_process_host(b200dd_spectrum *h, const double *x, uint32_t n, double *spectrum_out);
/* Device-resident variants: d_x = float2[n] (or double2[n] for _f64) in device memory; d_spectrum =
 * double2[n_spectrum] in device memory, or NULL to use the handle's own buffer (read with b200dd_spectrum_fetch).
 * Asynchronous on `stream` (NULL = the handle's stream). */
B200DD_API int b200dd_spectrum_process_device(b200dd_spectrum *h, const void *d_x, uint32_t n, void *d_spectrum,
                                              void *stream);
B200DD_API int b200dd_spectrum_process_device_f64(b200dd_spectrum *h, const void *d_x, uint32_t n, void *d_spectrum,
                                                  void *stream);
/* Copies the handle's buffer (result of the last process_device call with d_spectrum == NULL) to the host and
 * synchronises the stream. */
B200DD_API int b200dd_spectrum_fetch(b200dd_spectrum *h, double *spectrum_out, void *stream);
/* Profiling aid: CUDA-event durations (ms) of the folding pass (the HBM-streaming kernel) and of the rest
 * (reduction + nSpectrum-point DFT).  Synchronises the stream. */
B200DD_API int b200dd_spectrum_profile_device(b200dd_spectrum *h, const void *d_x, uint32_t n, void *stream,
                                              float *ms_fold, float *ms_rest);
B200DD_API void *b200dd_spectrum_stream(b200dd_spectrum *h);

/* ------------------------------------------------------------------ whole CPI pipeline
 *
 * One coherent-processing interval through the body of the reference's process thread
 * (src/blah2.cpp:268-287): [WienerHopf::process] -> Ambiguity::process -> Map::set_metrics ->
 * [CfarDetector1D -> Centroid -> Interpolate], with every intermediate (filtered surveillance
 * channel, range matrix, map, detection lists) staying in device memory.  The per-class entry
 * points above remain the drop-in boundary; this is the path a maintainer takes when the caller
 * can hand over a whole CPI at once (INTEGRATION.md "fast path").
 */
typedef struct b200dd_pipeline b200dd_pipeline;

typedef struct {
  b200dd_caf_params caf;
  int32_t clutter_enable;     /* process.clutter.enable   (config/config.yml:28-32) */
  int32_t clutter_delay_min;
  int32_t clutter_delay_max;
  int32_t detection_enable;   /* process.detection.enable (config/config.yml:33-40) */
  b200dd_det_params det;      /* det.resolution_doppler = 1 / tCpi as blah2.cpp:183 passes it */
} b200dd_pipeline_params;

typedef struct {
  int32_t filter_status;   /* B200DD_OK, or B200DD_FILTER_FAILED: the CPI is skipped like blah2.cpp:270-273 --
                              n_detections, noise_power and max_power are then 0 and the map buffer is not a product */
  uint32_t n_detections;
  double noise_power;      /* Map::noisePower */
  double max_power;        /* Map::maxPower   */
} b200dd_cpi_result;

B200DD_API int b200dd_pipeline_create(const b200dd_pipeline_params *params, b200dd_pipeline **out);
B200DD_API void b200dd_pipeline_destroy(b200dd_pipeline *h);
B200DD_API int b200dd_pipeline_get_geometry(const b200dd_pipeline *h, b200dd_caf_geometry *out);
B200DD_API int b200dd_pipeline_get_axes(const b200dd_pipeline *h, int32_t *delay, double *doppler);

/* HOST buffers: x, y = n complex128 samples (n == n_samples when the clutter filter is enabled,
 * >= n_used otherwise).  map_out (nullable) = [nDop][nDel] complex128.  Detections into the three
 * arrays (capacity cap).  Synchronous: returns when the results are in host memory. */
B200DD_API int b200dd_pipeline_process_host(b200dd_pipeline *h, const double *x, const double *y, uint32_t n,
                                            double *map_out, b200dd_cpi_result *result, double *o_delay,
                                            double *o_doppler, double *o_snr, uint32_t cap);

/* Asynchronous form of the host entry: enqueues H2D, all kernels and the D2H of the map on the
 * pipeline's own stream and returns; x, y and map_out must stay valid (pinned memory for true
 * overlap) until b200dd_pipeline_fetch returns.  Two pipelines used alternately overlap the PCIe
 * transfers of one CPI with the kernels of the other. */
B200DD_API int b200dd_pipeline_submit_host(b200dd_pipeline *h, const double *x, const double *y, uint32_t n,
                                           double *map_out);

/* Same, fed with the reference's replay / capture layout: n time instants of little-endian int16
 * I1 Q1 I2 Q2 (src/capture/rspduo/RspDuo.cpp:155-174; channel 1 = reference, channel 2 = surveillance).
 * 8 bytes per instant cross PCIe instead of 32; the de-interleave runs on the device.  int16 is exact
 * in float32, so results equal the complex128 entry point's. */
B200DD_API int b200dd_pipeline_submit_host_rspduo(b200dd_pipeline *h, const int16_t *iq, uint32_t n, double *map_out);

/* DEVICE buffers (float2), asynchronous on `stream` (NULL = the pipeline's stream).  d_map
 * (nullable) receives the float2 map.  Results stay on the device until b200dd_pipeline_fetch. */
B200DD_API int b200dd_pipeline_submit_device(b200dd_pipeline *h, const void *d_x, const void *d_y, uint32_t n,
                                             void *d_map, void *stream);
/* Plan creation for the device path (optional): records the whole chain for this (d_x, d_y, d_map) triple as a CUDA
 * graph -- it runs the chain once on the buffers' current contents, captures and instantiates it and synchronises
 * the stream -- after which b200dd_pipeline_submit_device on the same triple is ONE cudaGraphLaunch.  Results are
 * identical either way.  Worth it for submit -> fetch -> submit loops over a fixed ring of buffers (7-20 % per CPI);
 * with many CPIs enqueued ahead eager launches measured 3 % faster, hence opt-in.  B200DD_PIPELINE_GRAPH=1 records
 * every triple on its second use without this call, B200DD_PIPELINE_GRAPH=0 turns replay off altogether. */
B200DD_API int b200dd_pipeline_prepare_device(b200dd_pipeline *h, const void *d_x, const void *d_y, uint32_t n,
                                              void *d_map, void *stream);
B200DD_API int b200dd_pipeline_fetch(b200dd_pipeline *h, b200dd_cpi_result *result, double *o_delay,
                                     double *o_doppler, double *o_snr, uint32_t cap, void *stream);
B200DD_API void *b200dd_pipeline_stream(b200dd_pipeline *h);

/* Optional first stage of the process thread (src/blah2.cpp:263-265): SpectrumAnalyser(n_samples, bandwidth) on
 * the reference channel of every CPI submitted afterwards, evaluated on the samples already resident on the
 * device (no extra transfer).  *n_spectrum (nullable) receives the number of bins.  The spectrum of the last
 * submitted CPI is read with b200dd_pipeline_fetch_spectrum (call it after b200dd_pipeline_fetch, or it
 * synchronises the pipeline's stream itself). */
B200DD_API int b200dd_pipeline_enable_spectrum(b200dd_pipeline *h, double bandwidth, uint32_t *n_spectrum);
B200DD_API int b200dd_pipeline_fetch_spectrum(b200dd_pipeline *h, double *spectrum_out, uint32_t cap);

/* ------------------------------------------------------------------------------------------------
 * Inter-GPU exchanges (one process per GPU, NCCL over NVLink / NVSwitch).  The reference has no
 * distributed code; these serve the two shardings of SURVEY.md s8(e): independent CPIs round-robin
 * over the GPUs (the only communication: gather of finished maps to one rank) and one large CPI split
 * over the GPUs (all-gather of the range matrix, gather of the delay-column tiles; for the clutter filter
 * an all-reduce of the partial correlations and a halo from the left neighbour).
 * NCCL is resolved with dlopen("libnccl.so.2") at the first call: libb200dd.so has no link-time dependency on it.
 * Every *_async call orders the communicator's own stream after what is enqueued so far on `after` (a CUDA
 * stream, nullable), enqueues the exchange there and returns; b200dd_comm_join makes a compute stream wait for
 * everything enqueued on the communicator so far.  All ranks must issue the same sequence of calls.
 * ------------------------------------------------------------------------------------------------ */
typedef struct b200dd_comm b200dd_comm;
#define B200DD_COMM_ID_BYTES 128

/* rank 0 creates the 128-byte id and hands it to the other ranks by any host-side means */
B200DD_API int b200dd_comm_get_unique_id(uint8_t *id128);
B200DD_API int b200dd_comm_create(int32_t rank, int32_t world, const uint8_t *id128, int32_t device, b200dd_comm **out);
B200DD_API void b200dd_comm_destroy(b200dd_comm *c);
B200DD_API int32_t b200dd_comm_rank(const b200dd_comm *c);
B200DD_API int32_t b200dd_comm_world(const b200dd_comm *c);
B200DD_API void *b200dd_comm_stream(b200dd_comm *c);

/* every rank contributes `bytes` from d_send; rank dst receives rank r's block at d_recv + r * bytes
 * (d_recv may be NULL elsewhere) */
B200DD_API int b200dd_comm_gather_async(b200dd_comm *c, const void *d_send, void *d_recv, size_t bytes, int32_t dst,
                                        void *after);
/* blocks of different sizes: rank r's send_bytes must equal bytes[r]; block r lands at d_recv + offsets[r] */
B200DD_API int b200dd_comm_gatherv_async(b200dd_comm *c, const void *d_send, size_t send_bytes, void *d_recv,
                                         const size_t *bytes, const size_t *offsets, int32_t dst, void *after);
B200DD_API int b200dd_comm_allgatherv_async(b200dd_comm *c, const void *d_send, void *d_recv, const size_t *bytes,
                                            const size_t *offsets, void *after);
/* in-place sum of `count` doubles over the ranks */
B200DD_API int b200dd_comm_allreduce_f64_async(b200dd_comm *c, void *d_buf, size_t count, void *after);
/* one send and / or one receive in one NCCL group (halo exchange): a peer of -1 (or a size of 0) skips that half;
 * the peers' calls must pair up (rank a sends to b <=> rank b receives from a, same size) */
B200DD_API int b200dd_comm_sendrecv_async(b200dd_comm *c, const void *d_send, size_t send_bytes, int32_t send_peer,
                                          void *d_recv, size_t recv_bytes, int32_t recv_peer, void *after);
/* the communicator's stream waits for everything enqueued so far on `stream` (what `after` does inside the *_async
 * calls; call it once per compute stream when an exchange depends on several of them, then pass after = NULL) */
B200DD_API int b200dd_comm_wait_stream(b200dd_comm *c, void *stream);
B200DD_API int b200dd_comm_join(b200dd_comm *c, void *stream);
B200DD_API int b200dd_comm_sync(b200dd_comm *c);

/* Measured FP64 FMA rate of the device (TFLOP/s, 2 flop per FMA): the roofline of the FP64 WienerHopf kernels,
 * which MEASURED_PEAKS.json does not carry.  Takes a few milliseconds. */
B200DD_API int b200dd_ubench_fp64_tflops(int32_t device, double *tflops);

/* Pin the CALLING host thread to the CPUs local to the device's PCIe root (sysfs local_cpulist, intersected with
 * the thread's current affinity mask); call it before allocating pinned staging buffers so that they land on the
 * GPU's NUMA node.  cpulist_out (nullable, `cap` bytes) receives the kernel's list, e.g. "0-31,64-95". */
B200DD_API int b200dd_bind_host_to_device(int32_t device, char *cpulist_out, int32_t cap);

/* Pinned host memory for staging buffers (malloc / free semantics, NULL on failure). */
B200DD_API void *b200dd_host_alloc(size_t bytes);
B200DD_API void b200dd_host_free(void *p);

#ifdef __cplusplus
}
#endif

#endif /* B200DD_H */

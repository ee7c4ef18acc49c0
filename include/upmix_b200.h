/* upmix_b200.h -- C ABI of the B200-native multi-band STFT centre-extraction path.
 *
 * This is the drop-in boundary for the hot path of willleskowitz/upmix.  Plain pointers and sizes,
 * no C++ or torch types, no exceptions: every call returns 0 on success or a negative UPMIX_E_* code,
 * and upmix_last_error() gives the message for the calling thread.  All device memory is owned by the
 * caller (signal, outputs, workspace, streaming state); a plan owns only its read-only tables.  A plan
 * is immutable after creation, belongs to one device, and may be used from one stream at a time.
 *
 * Reference interfaces each entry point replaces (paths under the reference repository):
 *   python-prototype/center_extraction.py   (CE)      bela/upmix.cpp   (BU)
 *
 *   upmix_plan_create        <- the per-band state built by MultiBandExtractorAccu.__init__ (CE:240-271)
 *                               for every band of chain_bands (CE:518-580); C++ twin
 *                               Overlap75UpmixBand::setup / MultiBandUpmix::setup (BU:176-226, 439-471)
 *   upmix_process            <- extract_center_left_right_multi_band_in_memory (CE:477-513), i.e.
 *                               process_all_blocks of every band (CE:426-472) + the band sum (CE:503-511)
 *   upmix_process_segment    <- same, for one time shard of a track (no reference equivalent: the
 *                               reference keeps the whole signal in memory, CE:444-445)
 *   upmix_stream_block       <- process_stereo_chunk / flush_final state carried between calls
 *                               (CE:353-424) and MultiBandUpmix::process (BU:474-493)
 *   upmix_process_host       <- the same call as main.py makes (MP:78-80), with host buffers
 */
#ifndef UPMIX_B200_H
#define UPMIX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct UpmixPlan UpmixPlan;

/* One band, described by host tables (copied at plan creation).
 * n_fft: STFT size, a power of two in [64, 65536].  hop: n_fft/4 (75 % overlap) -- any hop that
 * divides n_fft is accepted for n_fft <= 8192; sizes above 8192 require hop == n_fft/4.
 * ana / syn: analysis and synthesis windows, n_fft floats (CE:257-258).
 * gain: per-bin real band-limit gain, n_fft/2+1 floats -- what _band_limit (CE:334-351) multiplies
 * both spectra with. */
typedef struct UpmixBandDesc {
    int32_t n_fft;       /* power of two in [64, 65536] */
    int32_t hop;         /* even, divides n_fft; above 8192 points: n_fft/4 or n_fft/2 (75 % / 50 % overlap) */
    const float* ana;
    const float* syn;
    const float* gain;
} UpmixBandDesc;

enum {
    UPMIX_OUT_LSCRS = 0, /* three outputs: centre, left-side, right-side (CE:513 order: C, Ls, Rs) */
    UPMIX_OUT_FOLD = 1   /* two outputs: Ls + 0.5 C, Rs + 0.5 C (BU:295-303; main.py "stereo_sum") */
};

enum {
    UPMIX_OK = 0,
    UPMIX_E_INVALID = -1,   /* bad argument */
    UPMIX_E_UNSUPPORTED = -2, /* size / hop outside what the kernels implement */
    UPMIX_E_CUDA = -3,      /* CUDA runtime error (message has the detail) */
    UPMIX_E_WORKSPACE = -4  /* workspace too small */
};

/* Message of the last failing call on this thread (never NULL). */
const char* upmix_last_error(void);

/* ABI version: major*1000 + minor. */
int upmix_version(void);

int upmix_plan_create(int n_bands, const UpmixBandDesc* bands, int out_mode, int device, UpmixPlan** out);
/* Same, with flags.  UPMIX_PLAN_NO_DECIMATE: band-limited bands (every live bin below 512 and below n_fft/16:
 * all but the top band of a crossover set) keep the full-size transform kernels instead of the decimated
 * ones; results agree to float32 rounding.  Block streaming always uses the full-size kernels, so a plan made
 * with this flag gives bit-identical results through upmix_process and upmix_stream_block. */
enum {
    UPMIX_PLAN_NO_DECIMATE = 1,
    /* dense bands of 256 / 512 / 1024 points keep the one-frame-per-CTA kernel instead of the frame-batched one
     * (16 frames per tile); results agree to float32 rounding.  Block streaming always runs the one-frame kernels. */
    UPMIX_PLAN_NO_BATCH = 2
};
int upmix_plan_create_ex(int n_bands, const UpmixBandDesc* bands, int out_mode, int device, int flags, UpmixPlan** out);
int upmix_plan_destroy(UpmixPlan* plan);
int upmix_plan_n_bands(const UpmixPlan* plan);
/* Bands with identical STFTs (size, hop, both windows) are merged into one pipeline: they share the
 * forward transform and, by linearity, the inverse transform and the overlap-add. */
int upmix_plan_n_pipelines(const UpmixPlan* plan);

/* Bytes of device workspace upmix_process / upmix_process_segment need for seg_len output samples
 * per track and n_tracks tracks (negative on error); the figure also covers every shorter segment with
 * the same n_tracks.  Any 256-byte-aligned device buffer will do. */
int64_t upmix_workspace_bytes(const UpmixPlan* plan, int64_t seg_len, int n_tracks);

/* Whole tracks.  L, R: device, planar float32, track t at L + t*in_stride, n_samples each.
 * out_c / out_l / out_r: device float32, track t at out + t*out_stride.  In FOLD mode out_c may be NULL.
 * Asynchronous on `stream` (a cudaStream_t passed as void*; NULL = default stream). */
int upmix_process(const UpmixPlan* plan, const float* L, const float* R, int64_t n_samples, int n_tracks,
                  int64_t in_stride, float* out_c, float* out_l, float* out_r, int64_t out_stride,
                  void* workspace, int64_t workspace_bytes, void* stream);

/* One time shard.  The track has n_total samples; L/R hold its samples [in_begin, in_begin+in_len)
 * (index 0 = sample in_begin); outputs receive samples [seg_begin, seg_end) (index 0 = seg_begin).
 * The input range must cover [seg_begin - halo, seg_end + halo) clipped to the track, halo =
 * upmix_segment_halo(plan); frames keep their global index, so the result is bit-identical to the
 * same samples of an unsharded run. */
int upmix_process_segment(const UpmixPlan* plan, const float* L, const float* R, int64_t in_begin, int64_t in_len,
                          int64_t n_total, int64_t seg_begin, int64_t seg_end, int n_tracks, int64_t in_stride,
                          float* out_c, float* out_l, float* out_r, int64_t out_stride, void* workspace,
                          int64_t workspace_bytes, void* stream);

/* Input margin (samples each side) upmix_process_segment needs: max over bands of n_fft, or of
 * n_fft + hop for bands above 8192 (their frames are transformed in even/odd pairs). */
int64_t upmix_segment_halo(const UpmixPlan* plan);

/* Block streaming (bands with n_fft <= 8192).  `state` is caller-owned device memory of
 * upmix_stream_state_bytes(plan, n_tracks) bytes; upmix_stream_reset clears it.  Each call consumes
 * n_new fresh samples per track (a multiple of every band's hop) and returns n_new output samples
 * delayed by the plan's largest (n_fft - hop): out[i] of call j is sample j*n_new + i - delay of the
 * offline result (zeros before the signal starts).  With hop-aligned blocks of hw samples and
 * n_fft <= 4*hw this is the 3*hw latency of the Bela program (BU:232-237). */
int64_t upmix_stream_state_bytes(const UpmixPlan* plan, int n_tracks);
int64_t upmix_stream_workspace_bytes(const UpmixPlan* plan, int n_new, int n_tracks);
int64_t upmix_stream_delay(const UpmixPlan* plan);
int upmix_stream_reset(const UpmixPlan* plan, void* state, int n_tracks, void* stream);
/* samples_done: samples per track consumed by the previous calls since the last reset (the caller
 * keeps this count; the state itself is plain device memory).  Steady-state blocks (samples_done >= twice the
 * largest n_fft) are replayed as CUDA graphs cached in the plan, one per block position modulo the largest
 * n_fft; in / out pointers may change from block to block at no cost, `state` and `workspace` should stay
 * put (a new pair means a new capture).  UPMIX_GRAPHS=0 in the environment at plan creation disables this. */
int upmix_stream_block(const UpmixPlan* plan, void* state, int64_t samples_done, const float* in_l, const float* in_r,
                       int n_new, int n_tracks, int64_t in_stride, float* out_c, float* out_l, float* out_r,
                       int64_t out_stride, void* workspace, int64_t workspace_bytes, void* stream);

/* One frame of a single-band plan, the stateful API of the prototype
 * (process_stereo_chunk, CE:353-409): blk_l / blk_r hold the n_fft samples of frame `frame_index`
 * (device), `ring` is the caller-owned overlap-add state [track][3][n_fft] floats (zero it to start;
 * its contents are the prototype's accumC/accumL/accumR, which is also what flush_final returns,
 * CE:411-424), out_* receive the hop samples the frame finishes.  frame_index must increase by one
 * per call.  Workspace: upmix_workspace_bytes(plan, hop, n_tracks).  Sizes above 8192 take one track per
 * call and Ls/C/Rs output (their centre is transformed with a zero partner frame, so the result agrees
 * with upmix_process to float32 rounding rather than bit for bit). */
int upmix_frame_step(const UpmixPlan* plan, void* ring, int64_t frame_index, const float* blk_l, const float* blk_r,
                     int n_tracks, int64_t in_stride, float* out_c, float* out_l, float* out_r, int64_t out_stride,
                     void* workspace, int64_t workspace_bytes, void* stream);

/* Host buffers: the call main.py makes (MP:43-50, 78-80).  L / R: host arrays of n elements, float32 (UPMIX_F32) or
 * float64 (UPMIX_F64), element strides stride_l / stride_r (main.py passes the two columns of an interleaved float64
 * [n][2] array: stride 2); pageable or pinned.  out_*: host float32 [n], caller-owned (pageable or pinned; out_c may
 * be NULL in FOLD mode).  The track goes through in time segments -- converted and uploaded chunk by chunk by
 * n_threads workers (0: UPMIX_HOST_THREADS, or the cores divided by LOCAL_WORLD_SIZE, at most 16), processed with upmix_process_segment as soon
 * as a segment's halo'd input has landed, downloaded and copied out by n_threads more workers -- so conversion, both
 * copy directions and the kernels overlap; the result is bit-identical to upmix_process on the same samples.  Pinned
 * buffers (cudaHostAlloc / cudaHostRegister) are used in place.  Device buffers and pinned staging are cached in the
 * plan between calls (one host call at a time per plan); upmix_plan_release_host frees them.  Returns when the outputs
 * are complete. */
enum { UPMIX_F32 = 0, UPMIX_F64 = 1 };
int upmix_process_host_ex(const UpmixPlan* plan, const void* L, const void* R, int dtype, int64_t stride_l, int64_t stride_r,
                          int64_t n_samples, float* out_c, float* out_l, float* out_r, int n_threads);
/* Same with contiguous float32 input and the default thread count. */
int upmix_process_host(const UpmixPlan* plan, const float* L, const float* R, int64_t n_samples, float* out_c,
                       float* out_l, float* out_r);
int upmix_plan_release_host(UpmixPlan* plan);

/* main.py's tail on the device.  upmix_peak3: peaks3[0..2] (device) = max|C|, max|Ls|, max|Rs| over n
 * samples (main.py:85-88); workspace of upmix_peak_workspace_bytes() device bytes.  upmix_export_mix:
 * scales C/Ls/Rs by `scale` (main.py:90-97) and writes interleaved stereo float32 [n][2]:
 * mode 0 "AB" out_a = (Ls+C+Rs, L+R) (main.py:110-119); mode 1 "split" out_a = (Ls,0), out_b = (C,C),
 * out_c = (0,Rs) (main.py:125-141); mode 2 "stereo_sum" out_a = (Ls+C/2, Rs+C/2) (main.py:143-153). */
int64_t upmix_peak_workspace_bytes(void);
int upmix_peak3(const float* c, const float* l, const float* r, int64_t n, float* peaks3, void* workspace,
                int64_t workspace_bytes, void* stream);
int upmix_export_mix(int mode, float scale, const float* c, const float* l, const float* r, const float* in_l,
                     const float* in_r, int64_t n, float* out_a, float* out_b, float* out_c, void* stream);

/* WAV edge (main.py:43-55, 119-153 read/write 16-bit PCM through soundfile): interleaved PCM16 stereo
 * [n][2] -> planar float32 L, R (x/32768) and *peak = max|x| (main.py:53); interleaved float32 stereo ->
 * PCM16 (x 32767, round to nearest even, saturating: libsndfile's
 * float -> PCM_16 scale when sf.write is handed float data).  Workspace: upmix_peak_workspace_bytes(). */
int upmix_pcm16_to_planar(const int16_t* interleaved, int64_t n, float* l, float* r, float* peak, void* workspace,
                          int64_t workspace_bytes, void* stream);
int upmix_stereo_to_pcm16(const float* interleaved, int64_t n, int16_t* out, void* stream);

/* Causal FIR filter, the arithmetic of apply_fir_filter (python-prototype/filter_design.py:54-59:
 * scipy.signal.lfilter(taps, 1.0, wave)): y[i] = sum_k taps[k] x[i-k], x[i<0] = 0, n outputs per track.
 * x, y, taps: device float32; track t at x + t*x_stride / y + t*y_stride; 1 <= n_taps <= 16384; y must
 * not overlap x.  The taps themselves (design_lr4_hp_fir / design_lr4_lp_fir, filter_design.py:25-52)
 * are designed on the host.  Nothing on the centre-extraction path calls this (the prototype does not
 * either: filter_design is imported by no other file). */
int upmix_fir_filter(const float* x, int64_t n, int n_tracks, int64_t x_stride, const float* taps, int n_taps, float* y,
                     int64_t y_stride, void* stream);

/* Measurement helpers (bench.py): number of kernels this library launched since the last reset, and
 * the FP32 FMA throughput of the device (the roofline denominator of this FP32-bound path). */
int64_t upmix_debug_launch_count(int reset);
int upmix_measure_fp32_tflops(int device, double* tflops, int* sm_count);

#ifdef __cplusplus
}
#endif
#endif /* UPMIX_B200_H */

// Kernel-side argument blocks shared by the kernels (upmix_kernels.cu) and the C-ABI host code
// (upmix_capi.cu).  Vocabulary follows the reference: a *band* is one crossover interval with its
// own STFT size (center_extraction.py:518-580), a *frame* is one STFT block starting at f*hop
// (center_extraction.py:448-460), a *hop* is the run of `hop` output samples that frame f finishes
// (center_extraction.py:396-399), a *track* is one stereo signal.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace upmix {

constexpr int COL_R = 16;         // column radix of the large-N (four-step) path
constexpr int FUSED_MAX_N = 8192; // largest STFT size handled by the single-CTA fused kernel
constexpr int LARGE_MAX_N = 65536;

// Decimated ("band-limited") transform of a band whose live bins all lie below P = n_fft / Q (see upmix_dec.cu):
// with n = Q p + q the frame splits into Q sequences of P points, and only bins [0, K] and their mirrors are
// ever needed.  P = 0: the band does not qualify (dense band) or the path is switched off.
struct DecDev {
    int P, Q;                  // points per decimated sequence, number of sequences (n_fft = P * Q)
    int K, KP;                 // highest live bin (max_bin), row length of the spectrum arrays (K + 1 rounded up to 32)
    const float2* tw_last;     // [R0][16]  exp(-2 pi i r k / P): twiddles of the second (radix-16) pass of the P-point transform
    const float2* tw_step;     // [KP]      exp(-2 pi i k / n_fft)
    const float2* tw_base;     // [Q/16][KP] exp(-2 pi i 16 g k / n_fft): twiddle of the first sequence of group g
};

// Frame-batched kernel of a dense band of 256 / 512 / 1024 points (upmix_fb.cuh).  tw_full == nullptr: not used.
struct FbDev {
    const float2* tw_full;     // [RA][RB]  exp(-2 pi i r k / n_fft): second-pass twiddles of the n_fft-point transform
    const float2* tw_half;     // [HA][16]  exp(-2 pi i r k / (n_fft/2)): second pass of the centre's n_fft/2-point transform
};

// Device tables of one band (all in global memory, read-only during processing).
struct BandDev {
    int n_fft;
    int hop;                   // hop the kernels step by.  A band above 8192 points with 50 % overlap (hop = n_fft/2) runs on
                               // the 75 % machinery: hop = n_fft/4 here and frame_step = 2 -- the odd frames do not exist
                               // (their spectra are zero: they add exact zeros to the overlap-add, so the result is the
                               // two-frame sum of center_extraction.py:392-407 bit for bit)
    int frame_step;            // 1, or 2 (see hop)
    const float* ana;          // [n_fft + 1]   analysis window, followed by one zero
    const float* syn;          // [n_fft]       synthesis window / n_fft (the inverse FFT is unnormalised)
    const float* gain;         // [n_gains][gain_stride] band-limit gains of the bands merged into this
                               // pipeline (same n_fft/hop/windows); per bin the non-zero gains come first
    int n_gains;
    int gain_stride;           // >= n_fft/2+1
    int max_bin;               // highest bin with a non-zero gain in any of the merged bands (-1: none)
    const float2* tw_fft;      // per-pass twiddles (fft_device.cuh layout) of the n_fft-point transform
                               // (fused path) or of the n_fft/16-point row transform (large path)
    const float2* tw_inv;      // per-pass twiddles of the inverse n_fft-point transform's plan (fused path; the
                               // fused kernel's inverse starts with the radix its forward transform ends with)
    const float2* tw_half;     // per-pass twiddles of the n_fft/2-point transform (fused path)
    const float2* tw_pack;     // [n_fft/2]     exp(-2*pi*i*k/n_fft) for the real-signal packing (fused path)
    const float2* tw_col;      // [16][n_fft/16] exp(-2*pi*i*k1*n2/n_fft), large path only
    DecDev dec;                // decimated path (dec.P == 0: not used)
    FbDev fb;                  // frame-batched path (fb.tw_full == nullptr: not used)
};

// Where the samples are.  Global sample index s of a track lives at in[s - in_begin] for
// in_begin <= s < in_end; everything else reads as zero (signal start, zero-extended tail, or the
// part of the track another shard owns).  Output sample s goes to out[ch][s - out_begin] when
// seg_begin <= s < seg_end.
struct SegArgs {
    const float* in_l;
    const float* in_r;
    long long in_stride;       // elements between tracks
    long long in_begin, in_end;
    float* out_c;
    float* out_l;
    float* out_r;
    long long out_stride;      // elements between tracks
    long long out_begin;
    long long seg_begin, seg_end;
    long long hop_begin, hop_end;   // global hop indices to produce
    int hops_per_run;               // hops handled by one CTA / thread run (plus warm-up frames)
    int fold;                       // fused kernel: fold the centre in the frequency domain -- emit
                                    // Ls + 0.5 C and Rs + 0.5 C in the Ls / Rs slots, no C transform
    float* state;                   // optional streaming state [track][3][n_fft]: ring carried between calls
    float* state_out;               // where the ring is saved (state == nullptr: unused).  One run per track: the CTA loads
                                    // `state` and saves to `state_out` (may be the same memory).  Several runs per track
                                    // (hops_per_run >= 3): run 0 loads `state`, the others replay their three warm-up frames
                                    // from the input like an offline call, and the run that ends at hop_end saves -- to a
                                    // DIFFERENT buffer, which the host copies over `state` afterwards (run 0 may load late)
    int accum;                      // 0: finished samples are stored;  1: added to what out_* already holds
                                    // (float32, band order = launch order: center_extraction.py:503-511)
    int mix;                        // 0: out_c/out_l/out_r receive C, Ls, Rs;  1: fold-down epilogue --
                                    // out_l receives Ls + 0.5 C, out_r Rs + 0.5 C, out_c is not touched
};

// Scratch of one wave of the large-N path.  Frames [frame0, frame0 + n_frames) of every track,
// frame0 even, n_frames even.
struct WaveArgs {
    float2* a;        // [track][n_frames][16][n2]   column-transformed, twiddled input spectra
    float2* b1;       // [track][n_frames][16][n2]   row-inverse-transformed Ls + i*Rs
    float2* b2;       // [track][n_frames/2][16][n2] row-inverse-transformed C(f even) + i*C(f odd)
    long long frame0;
    int n_frames;
};

// Scratch of one wave of the decimated path: frames [frame0, frame0 + n_frames) of every track.
struct DecWave {
    float2* part;     // [track][n_frames][Q/16][2][KP]  per-group partial sums of Z[k] and Z[n_fft-k] (Q > 16 only)
    float2* spec;     // [track][n_frames][3][KP]        masked spectra: Y[k], Y[n_fft-k] of Ls + i Rs, and C[k]
    long long frame0;
    int n_frames;
};

}  // namespace upmix

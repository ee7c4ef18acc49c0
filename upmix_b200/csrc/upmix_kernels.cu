// Hand-written sm_100a kernels for the multi-band STFT centre-extraction path.
//
//   band_fused_kernel<N>    N <= 8192: whole per-band chain in one CTA per run of hops --
//   (upmix_fused.cuh,       frame load + analysis window -> packed complex FFT (L + iR) in shared
//    upmix_fused_*.cu)      memory -> Hermitian split, band gain, centre mask -> inverse FFT of
//                           Ls + i*Rs (N points) and of C (N/2 points, real-signal packing) ->
//                           synthesis window -> overlap-add ring in shared memory -> finished hops.
//   col_fwd_kernel          N >= 16384 (four-step, N = 16 x N2): radix-16 column DFT of the windowed
//   row_mask_kernel<N2>     frame in registers; N2-point row FFT + mask + inverse row FFT in shared
//   col_inv_ola_kernel      memory; radix-16 inverse column DFT + synthesis window + overlap-add in
//                           registers.
//   band_sum_kernel         sums the per-band outputs in band order (center_extraction.py:503-511)
//                           and applies the output mode (Ls/C/Rs or L+0.5C / R+0.5C fold-down).
//
// Reference behaviour reproduced: center_extraction.py:353-472 (per-band chain), bela/upmix.cpp:238-306.
#include <stdlib.h>

#include "fft_device.cuh"
#include "upmix_kernels.cuh"
#include "upmix_launch.h"
#include "upmix_fused.cuh"

namespace upmix {

#ifndef UPMIX_TW_IN_ROW
#define UPMIX_TW_IN_ROW -1    // four-step twiddles applied by the row kernel (1: one table look-up per point, as
                              // much L2 traffic as the data) or by the column kernels (0: a column's fifteen
                              // twiddles are powers of one number -- four look-ups, eleven multiplies); -1: by
                              // size, as measured (ms per band-hour, row / column): 16384 6.56 / 6.88-6.97,
                              // 32768 6.34 / 6.22, 65536 6.55 / 6.14 -- short rows leave the column kernels,
                              // which sit at the HBM roofline, as the larger share
#endif
static inline bool tw_in_row(int n_fft) { return UPMIX_TW_IN_ROW < 0 ? n_fft / COL_R <= 1024 : UPMIX_TW_IN_ROW != 0; }

// ---------------------------------------------------------------------------------------------
// large-N path, N = 16 * N2
// ---------------------------------------------------------------------------------------------
// K1: one thread per (frame, column n2): windowed samples x[16 rows][n2] -> radix-16 DFT over the
// rows -> A[frame][k1][n2] (the four-step twiddle W_N^{n2*k1} is applied by the row kernel's loads).
template <bool TWROW>
__global__ void __launch_bounds__(128) col_fwd_kernel(const BandDev b, const SegArgs a, const WaveArgs w) {
    const int N2 = b.n_fft / COL_R;
    const int n2 = blockIdx.x * blockDim.x + threadIdx.x;
    const int fl = blockIdx.y;
    const int track = blockIdx.z;
    if (n2 >= N2) return;
    const long long f = w.frame0 + fl;
    const long long s0 = f * b.hop;
    const float* __restrict__ inl = a.in_l + (long long)track * a.in_stride;
    const float* __restrict__ inr = a.in_r + (long long)track * a.in_stride;
    float2 v[COL_R];
    const float* __restrict__ pl = inl + (s0 - a.in_begin);
    const float* __restrict__ pr = inr + (s0 - a.in_begin);
    if (b.frame_step > 1 && f % b.frame_step != 0) {
        // 50 % overlap run on the 75 % machinery: this frame does not exist -- zero spectrum
#pragma unroll
        for (int r = 0; r < COL_R; r++) v[r] = make_float2(0.f, 0.f);
    } else if (s0 >= a.in_begin && s0 + b.n_fft <= a.in_end) {          // whole frame available (block-uniform)
#pragma unroll
        for (int r = 0; r < COL_R; r++) {
            const int n = r * N2 + n2;
            const float wn = __ldg(b.ana + n);
            v[r] = cscale(make_float2(__ldg(pl + n), __ldg(pr + n)), wn);
        }
    } else {
#pragma unroll
        for (int r = 0; r < COL_R; r++) {
            const int n = r * N2 + n2;
            const long long s = s0 + n;
            const bool ok = s >= a.in_begin && s < a.in_end;
            const long long idx = ok ? s - a.in_begin : 0;
            const float wn = __ldg(b.ana + n);
            const float l = __ldg(inl + idx), rr = __ldg(inr + idx);
            v[r] = ok ? make_float2(l * wn, rr * wn) : make_float2(0.f, 0.f);
        }
    }
    Dft<COL_R, -1>::run(v);
    float2* dst = w.a + (((long long)track * w.n_frames + fl) * COL_R) * N2 + n2;
    if constexpr (!TWROW)          // A * W_N^{k1 n2} here; otherwise the row kernel applies it
        apply_powers16<false>(v, __ldg(b.tw_col + N2 + n2), __ldg(b.tw_col + 2 * N2 + n2), __ldg(b.tw_col + 4 * N2 + n2),
                              __ldg(b.tw_col + 8 * N2 + n2));
#pragma unroll
    for (int k1 = 0; k1 < COL_R; k1++) dst[(long long)k1 * N2] = v[k1];
}

template <int N2> struct RowCfg;
#define UPMIX_ROW_CFG(N2_, PLAN_)                                                                \
    template <> struct RowCfg<N2_> {                                                              \
        static constexpr int PLAN = PLAN_;                                                        \
        static_assert(fft_size(PLAN_) == N2_, "plan does not match the size");                   \
        static constexpr int T = 2 * N2_ / fft_radix(PLAN_, 0); /* two rows, one butterfly each */ \
        static constexpr int SMEM = 6 * PADSZ<PLAN_>() * (int)sizeof(float2);                     \
    };
#ifndef UPMIX_ROWPLAN_1024
#define UPMIX_ROWPLAN_1024 mkplan(16, 16, 4)
#endif
#ifndef UPMIX_ROWPLAN_2048
#define UPMIX_ROWPLAN_2048 mkplan(16, 16, 8)
#endif
#ifndef UPMIX_ROWPLAN_4096
#define UPMIX_ROWPLAN_4096 mkplan(16, 16, 16)
#endif
UPMIX_ROW_CFG(1024, UPMIX_ROWPLAN_1024)
UPMIX_ROW_CFG(2048, UPMIX_ROWPLAN_2048)
UPMIX_ROW_CFG(4096, UPMIX_ROWPLAN_4096)

// K2: one CTA per (row pair, frame pair, track).  Rows k1 and 16-k1 of the frame's spectrum are
// mirror images of each other (bin k <-> N-k), so the CTA holds both rows of both frames, finishes
// the forward transform along the rows, applies split/gain/mask, and starts the inverse transform
// (rows) of Ls+iRs for each frame and of C(even frame) + i*C(odd frame).
template <int N2, bool TWROW>
__global__ void __launch_bounds__(RowCfg<N2>::T) row_mask_kernel(const BandDev b, const WaveArgs w) {
    constexpr int T = RowCfg<N2>::T;
    constexpr int PL = RowCfg<N2>::PLAN;
    constexpr int RS = PADSZ<PL>();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* S = reinterpret_cast<float2*>(smem_raw);   // 6 rows: f0a f0b f1a f1b ca cb
    const int tid = threadIdx.x;
    const int pr = blockIdx.x;                          // 0: rows (0, 8) self-mirrored; p: rows (p, 16-p)
    const int fp = blockIdx.y;
    const int track = blockIdx.z;
    const int ka = pr, kb = pr == 0 ? COL_R / 2 : COL_R - pr;
    const long long fbase = ((long long)track * w.n_frames + 2 * fp) * COL_R;
    const float2* __restrict__ A = w.a;
    const float2* __restrict__ tw = b.tw_fft;
    const float2* __restrict__ twc = b.tw_col;
    const float* __restrict__ gain = b.gain;

    // forward row transforms: rows (f0,ka) (f0,kb) then (f1,ka) (f1,kb)
#pragma unroll
    for (int g = 0; g < 2; g++) {
        auto ld = [&](int row, int n, int, int) -> float2 {      // A * W_N^{k1 n}: the four-step twiddle
            const int k1 = row ? kb : ka;
            if constexpr (TWROW) return cmul(A[(fbase + (long long)g * COL_R + k1) * N2 + n], __ldg(twc + k1 * N2 + n));
            else return A[(fbase + (long long)g * COL_R + k1) * N2 + n];
        };
        float2* buf = S + 2 * g * RS;
        auto st = make_store([&](int row, int k, float2 v, NoAux) { buf[row * RS + PAD<PL>(k)] = v; });
        fft_smem<PL, -1, T, 2, false>(buf, tid, tw, ld, st);
    }

    // split / gain / mask over mirror pairs
    const int n_items = pr == 0 ? N2 + 1 : N2;
    {
        constexpr int ITM = (N2 + 1 + T - 1) / T;
        int lo_a[ITM], hi_a[ITM], bin_a[ITM];
        float g_a[ITM];
#pragma unroll
        for (int i = 0; i < ITM; i++) {                 // index arithmetic and gain loads first
            const int it = min(tid + i * T, n_items - 1);
            int lo_row, lo_idx, hi_row, hi_idx, bin;
            if (pr != 0) {
                const int k2 = it;
                if (k2 < N2 / 2) { lo_row = 0; lo_idx = k2; hi_row = 1; hi_idx = N2 - 1 - k2; bin = ka + COL_R * k2; }
                else { lo_row = 1; lo_idx = N2 - 1 - k2; hi_row = 0; hi_idx = k2; bin = kb + COL_R * (N2 - 1 - k2); }
            } else if (it <= N2 / 2) {          // row 0: bin 16*k2 <-> 16*(N2-k2)
                lo_row = 0; lo_idx = it; hi_row = 0; hi_idx = (N2 - it) & (N2 - 1); bin = COL_R * it;
            } else {                            // row 8: bin 8+16*k2 <-> 8+16*(N2-1-k2)
                const int k2 = it - (N2 / 2 + 1);
                lo_row = 1; lo_idx = k2; hi_row = 1; hi_idx = N2 - 1 - k2; bin = COL_R / 2 + COL_R * k2;
            }
            lo_a[i] = lo_row * RS + PAD<PL>(lo_idx);
            hi_a[i] = hi_row * RS + PAD<PL>(hi_idx);
            g_a[i] = __ldg(gain + bin);
            bin_a[i] = bin;
        }
#pragma unroll
        for (int i = 0; i < ITM; i++) {
            if (tid + i * T >= n_items) break;
            const int lo = lo_a[i], hi = hi_a[i];
            const float g = g_a[i];
            float2 c[2];
#pragma unroll
            for (int fr = 0; fr < 2; fr++) {
                float2* buf = S + 2 * fr * RS;
                float2 ylo, yhi;
                mask_bin_merged(buf[lo], buf[hi], g, gain + bin_a[i], b.n_gains, b.gain_stride, ylo, yhi, c[fr]);
                buf[lo] = ylo;
                buf[hi] = yhi;
            }
            float2* cb = S + 4 * RS;
            cb[lo] = cadd(c[0], make_float2(-c[1].y, c[1].x));                          // C_even + i C_odd
            cb[hi] = cadd(make_float2(c[0].x, -c[0].y), make_float2(c[1].y, c[1].x));   // conj C_even + i conj C_odd
        }
    }
    __syncthreads();

    // inverse row transforms, results to the wave scratch
#pragma unroll
    for (int g = 0; g < 3; g++) {
        float2* buf = S + 2 * g * RS;
        auto ld = [&](int row, int n, int, int) -> float2 { return buf[row * RS + PAD<PL>(n)]; };
        float2* dst = g < 2 ? w.b1 + (fbase + (long long)g * COL_R) * N2
                            : w.b2 + (((long long)track * (w.n_frames / 2) + fp) * COL_R) * N2;
        if constexpr (TWROW) {
            auto st = make_store([&](int row, int n) -> float2 { return __ldg(twc + (row ? kb : ka) * N2 + n); },
                                 [&](int row, int n, float2 v, float2 t) {     // * conj W_N^{k1 n}
                                     const int k1 = row ? kb : ka;
                                     dst[(long long)k1 * N2 + n] = cmul(v, make_float2(t.x, -t.y));
                                 });
            fft_smem<PL, +1, T, 2, true>(buf, tid, tw, ld, st);
        } else {
            auto st = make_store([&](int row, int n, float2 v, NoAux) { dst[(long long)(row ? kb : ka) * N2 + n] = v; });
            fft_smem<PL, +1, T, 2, true>(buf, tid, tw, ld, st);
        }
    }
}

// K2, band-limited: every non-zero gain of the pipeline sits below bin 16*ROW_K (BandDev::max_bin; true
// for every band the dynamic-resolution rule sizes above 8192, whose pass band ends near bin 430).
// Then only the first and last ROW_K points of each row carry signal: the forward row transforms
// compute just those (fft_rows_fwd_pruned) into a small spectrum array, the mask runs over 2*ROW_K
// mirror pairs, and the inverse row transforms start from those inputs alone (fft_rows_inv_pruned).
// With the six rows no longer resident the CTA needs one row of shared memory, so it transforms one row
// at a time with half the threads and two (or more) CTAs share an SM: one computes while the other
// waits for its rows to arrive from HBM.
// Register budget per thread: 65536 / (T * REGS) CTAs share an SM.  Without the four-step twiddles the kernel
// fits 80 registers and a third CTA per SM pays (ms per band-hour, 128 / 100 / 85: 65536 6.16 / 6.15 / 5.78,
// 32768 6.24 / 6.02 / 5.93); with them (16384 points) 128 is best (6.58 / 6.54 / 6.80).
#ifndef UPMIX_ROWP_REGS
#define UPMIX_ROWP_REGS(TWROW) ((TWROW) ? 128 : 85)
#endif
template <int N2, bool TWROW>
__global__ void __launch_bounds__(N2 / 16, 65536 / (N2 / 16 * UPMIX_ROWP_REGS(TWROW))) row_mask_pruned_kernel(const BandDev b, const WaveArgs w) {
    constexpr int T = N2 / 16;
    constexpr int PL = RowCfg<N2>::PLAN;
    constexpr int RS = PADSZ<PL>();
    constexpr int K = ROW_K;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* buf = reinterpret_cast<float2*>(smem_raw);             // one row being transformed
    float2* spec = buf + RS;                                       // [3: f0 f1 c][2 rows][2K] band-limited spectra
    const int tid = threadIdx.x;
    const int pr = blockIdx.x;                          // 0: rows (0, 8) self-mirrored; p: rows (p, 16-p)
    const int fp = blockIdx.y;
    const int track = blockIdx.z;
    const int ka = pr, kb = pr == 0 ? COL_R / 2 : COL_R - pr;
    const long long fbase = ((long long)track * w.n_frames + 2 * fp) * COL_R;
    const float2* __restrict__ A = w.a;
    const float2* __restrict__ tw = b.tw_fft;
    const float2* __restrict__ twc = b.tw_col;
    const float* __restrict__ gain = b.gain;

#pragma unroll 1
    for (int s = 0; s < 4; s++) {                       // (frame, row) = (0,a) (0,b) (1,a) (1,b)
        const int g = s >> 1, k1 = (s & 1) ? kb : ka;
        const float2* __restrict__ src = A + (fbase + (long long)g * COL_R + k1) * N2;
        const float2* __restrict__ twr = twc + k1 * N2;
        auto ld = [&](int, int n, int, int) -> float2 {
            if constexpr (TWROW) return cmul(src[n], __ldg(twr + n));       // * W_N^{k1 n}
            else return src[n];                                             // twiddled by col_fwd
        };
        float2* sp = spec + s * 2 * K;
        auto out = [&](int, int idx, float2 v) { sp[idx] = v; };
        fft_rows_fwd_pruned<PL, T, 1>(buf, tid, tw, ld, out);
    }

    // split / gain / mask over the 2K mirror pairs.  Slot idx < K is point idx of the row, idx >= K is
    // point N2 - 2K + idx; items [0, K) start in the first row, [K, 2K) in the second.
    if (tid < 2 * K) {
        const int j = tid & (K - 1);
        const bool second = tid >= K;
        int lo, hi, bin;
        if (pr != 0) {                  // bin k1 + 16 j  <->  (16 - k1) + 16 (N2 - 1 - j)
            lo = (second ? 2 * K : 0) + j;
            hi = (second ? 0 : 2 * K) + 2 * K - 1 - j;
            bin = (second ? kb : ka) + COL_R * j;
        } else if (!second) {           // row 0: bin 16 j <-> 16 (N2 - j)
            lo = j;
            hi = j == 0 ? 0 : 2 * K - j;
            bin = COL_R * j;
        } else {                        // row 8: bin 8 + 16 j <-> 8 + 16 (N2 - 1 - j)
            lo = 2 * K + j;
            hi = 2 * K + 2 * K - 1 - j;
            bin = COL_R / 2 + COL_R * j;
        }
        const float g = __ldg(gain + bin);
        float2 c[2];
#pragma unroll
        for (int fr = 0; fr < 2; fr++) {
            float2* sp = spec + fr * 4 * K;
            float2 ylo, yhi;
            mask_bin_merged(sp[lo], sp[hi], g, gain + bin, b.n_gains, b.gain_stride, ylo, yhi, c[fr]);
            sp[lo] = ylo;
            sp[hi] = yhi;
        }
        float2* cb = spec + 8 * K;
        cb[lo] = cadd(c[0], make_float2(-c[1].y, c[1].x));                          // C_even + i C_odd
        cb[hi] = cadd(make_float2(c[0].x, -c[0].y), make_float2(c[1].y, c[1].x));   // conj C_even + i conj C_odd
        // row 0, point N2-K (slot K): the mirror of bin 16*K, which has no gain -- no item covers it, but
        // the pruned inverse reads it
        if (pr == 0 && tid == 0) spec[K] = spec[4 * K + K] = spec[8 * K + K] = make_float2(0.f, 0.f);
    }
    __syncthreads();

#pragma unroll 1
    for (int s = 0; s < 6; s++) {                       // (f0,a) (f0,b) (f1,a) (f1,b) (c,a) (c,b)
        const int g = s >> 1, k1 = (s & 1) ? kb : ka;
        float2* dst = (g < 2 ? w.b1 + (fbase + (long long)g * COL_R) * N2
                             : w.b2 + (((long long)track * (w.n_frames / 2) + fp) * COL_R) * N2) + (long long)k1 * N2;
        const float2* __restrict__ twr = twc + k1 * N2;
        const float2* sp = spec + s * 2 * K;
        auto in = [&](int, int idx) -> float2 { return sp[idx]; };
        if constexpr (TWROW) {
            auto st = make_store([&](int, int n) -> float2 { return __ldg(twr + n); },
                                 [&](int, int n, float2 v, float2 t) { dst[n] = cmul(v, make_float2(t.x, -t.y)); });   // * conj W_N^{k1 n}
            fft_rows_inv_pruned<PL, T, 1>(buf, tid, tw, in, st);
        } else {
            auto st = make_store([&](int, int n, float2 v, NoAux) { dst[n] = v; });               // twiddled by col_inv_ola
            fft_rows_inv_pruned<PL, T, 1>(buf, tid, tw, in, st);
        }
    }
}

// K3: one thread per (column n2, run of hops, track).  For every frame pair of the run (plus the
// frames before it whose tails reach into the run) the thread finishes the inverse transform down
// its column (radix-16), applies the synthesis window and overlap-adds in registers: a frame shifts
// the 16-row accumulator by 4 rows (hop = N/4 = 4*N2), the 4 rows that fall out are finished samples.
// TWROW = false: the kernel also applies the four-step twiddles, which costs registers: 3 CTAs per SM (168
// registers, no spills) instead of 4 (measured: 65536 6.70 -> 6.14 ms per band-hour)
// ACCUM: the band adds its hops to what the outputs hold (a compile-time variant: the values it prefetches cost
// 12 registers, which the storing variant should not pay for; it keeps the occupancy of its twin -- measured on
// the default six bands, 1-hour track: 4 CTAs per SM with 72 B of spills 29.6 ms, 3 CTAs per SM without 30.8 ms).
template <bool TWROW, bool ACCUM>
__global__ void __launch_bounds__(128, TWROW ? 4 : 3) col_inv_ola_kernel(const BandDev b, const SegArgs a, const WaveArgs w) {
    const int N2 = b.n_fft / COL_R;
    const int n2 = blockIdx.x * blockDim.x + threadIdx.x;
    if (n2 >= N2) return;
    const int track = blockIdx.z;
    const int H = b.hop;
    const long long h0 = a.hop_begin + (long long)blockIdx.y * a.hops_per_run;
    const long long h1 = min(h0 + (long long)a.hops_per_run, a.hop_end);
    if (h0 >= h1) return;
    const long long f_begin = max(0LL, h0 - 3);
    float* outp[3] = {a.out_c + (long long)track * a.out_stride, a.out_l + (long long)track * a.out_stride,
                      a.out_r + (long long)track * a.out_stride};
    const float* __restrict__ syn = b.syn + n2;

    float acc[3][COL_R];
#pragma unroll
    for (int ch = 0; ch < 3; ch++)
#pragma unroll
        for (int i = 0; i < COL_R; i++) acc[ch][i] = 0.f;
    // this column's four-step twiddles W_N^{k1 n2}: four powers kept, the rest rebuilt before every column transform
    float2 tw1, tw2, tw4, tw8;
    if constexpr (!TWROW) {
        tw1 = __ldg(b.tw_col + N2 + n2);
        tw2 = __ldg(b.tw_col + 2 * N2 + n2);
        tw4 = __ldg(b.tw_col + 4 * N2 + n2);
        tw8 = __ldg(b.tw_col + 8 * N2 + n2);
    }

    // finished rows of frame f leave the accumulator; the rest moves up by one hop
    // accum: what the output already holds for the hop frame f finishes (the bands before this one) is requested
    // together with the frame's scratch rows and waits in registers -- loaded at the store, one dependent HBM
    // round trip per sample made this kernel 2.7x slower than its storing twin
    float pv[3][ACCUM ? COL_R / 4 : 1];
    auto request_prev = [&](long long f) {
        if (!ACCUM || f < h0 || f >= h1) return;
        const long long s0 = f * H + n2;
#pragma unroll
        for (int n1 = 0; n1 < COL_R / 4; n1++) {
            const long long s = s0 + n1 * N2;
            const bool in = s >= a.seg_begin && s < a.seg_end;
#pragma unroll
            for (int ch = 0; ch < 3; ch++)
                if (ch > 0 || !a.mix) pv[ch][ACCUM ? n1 : 0] = in ? __ldcs(outp[ch] + (s - a.out_begin)) : 0.f;
        }
    };
    auto emit_shift = [&](long long f) {
        const bool emit = f >= h0 && f < h1;
        const long long s0 = f * H + n2;
#pragma unroll
        for (int n1 = 0; n1 < COL_R / 4; n1++) {
            const long long s = s0 + n1 * N2;
            if (emit && s >= a.seg_begin && s < a.seg_end) {
                const long long o = s - a.out_begin;
                if (a.mix) {                     // fold-down epilogue: Ls + 0.5 C, Rs + 0.5 C
                    const float hc = 0.5f * acc[0][n1];
                    const float vl = acc[1][n1] + hc, vr = acc[2][n1] + hc;
                    __stcs(outp[1] + o, ACCUM ? pv[1][ACCUM ? n1 : 0] + vl : vl);
                    __stcs(outp[2] + o, ACCUM ? pv[2][ACCUM ? n1 : 0] + vr : vr);
                } else {
#pragma unroll
                    for (int ch = 0; ch < 3; ch++) __stcs(outp[ch] + o, ACCUM ? pv[ch][ACCUM ? n1 : 0] + acc[ch][n1] : acc[ch][n1]);
                }
            }
        }
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
#pragma unroll
            for (int i = 0; i < COL_R - COL_R / 4; i++) acc[ch][i] = acc[ch][i + COL_R / 4];
#pragma unroll
            for (int i = COL_R - COL_R / 4; i < COL_R; i++) acc[ch][i] = 0.f;
        }
    };

    for (long long p = f_begin >> 1; 2 * p < h1; ++p) {
        const long long fl = 2 * p - w.frame0;              // local index of the even frame
        const float2* __restrict__ B2 = w.b2 + (((long long)track * (w.n_frames / 2) + (fl >> 1)) * COL_R) * N2 + n2;
        const float2* __restrict__ B1 = w.b1 + (((long long)track * w.n_frames + fl) * COL_R) * N2 + n2;
        float wn[COL_R];
        float codd[COL_R];
#pragma unroll
        for (int n1 = 0; n1 < COL_R; n1++) wn[n1] = __ldg(syn + n1 * N2);
        {
            float2 v[COL_R];
#pragma unroll
            for (int k1 = 0; k1 < COL_R; k1++) v[k1] = B2[(long long)k1 * N2];
            if constexpr (!TWROW) apply_powers16<true>(v, tw1, tw2, tw4, tw8);
            Dft<COL_R, +1>::run(v);                          // v[n1] = (c_even[n], c_odd[n]), n = n1*N2 + n2
#pragma unroll
            for (int n1 = 0; n1 < COL_R; n1++) {
                acc[0][n1] += v[n1].x * wn[n1];
                codd[n1] = v[n1].y;
            }
        }
#pragma unroll
        for (int half = 0; half < 2; half++) {
            float2 v[COL_R];
            request_prev(2 * p + half);
#pragma unroll
            for (int k1 = 0; k1 < COL_R; k1++) v[k1] = B1[((long long)half * COL_R + k1) * N2];
            if constexpr (!TWROW) apply_powers16<true>(v, tw1, tw2, tw4, tw8);
            Dft<COL_R, +1>::run(v);
#pragma unroll
            for (int n1 = 0; n1 < COL_R; n1++) {
                if (half) acc[0][n1] += codd[n1] * wn[n1];
                acc[1][n1] += v[n1].x * wn[n1];
                acc[2][n1] += v[n1].y * wn[n1];
            }
            emit_shift(2 * p + half);
        }
    }
}

// Single-frame variant of K3 for the stateful chunk API (process_stereo_chunk, center_extraction.py:
// 353-409) on the four-step path: the frame sits in slot 0 of a two-frame wave whose partner is zero.
// One thread per column: finish the inverse transform, window, add into the caller's overlap-add ring
// [track][3][N] (frame-local order), emit the first hop and shift the ring by one hop.  A hop is 4 rows
// of the thread's own column, so the in-place shift touches nobody else's samples.
template <bool TWROW>
__global__ void __launch_bounds__(128, TWROW ? 4 : 3) col_inv_frame_kernel(const BandDev b, const WaveArgs w, float* __restrict__ ring,
                                                               float* __restrict__ out_c, float* __restrict__ out_l,
                                                               float* __restrict__ out_r, long long out_stride) {
    const int N = b.n_fft;
    const int N2 = N / COL_R;
    const int n2 = blockIdx.x * blockDim.x + threadIdx.x;
    if (n2 >= N2) return;
    const int track = blockIdx.z;
    const float2* __restrict__ B2 = w.b2 + ((long long)track * (w.n_frames / 2) * COL_R) * N2 + n2;
    const float2* __restrict__ B1 = w.b1 + ((long long)track * w.n_frames * COL_R) * N2 + n2;
    float* rg = ring + (long long)track * 3 * N + n2;
    float* outp[3] = {out_c + (long long)track * out_stride, out_l + (long long)track * out_stride,
                      out_r + (long long)track * out_stride};
    float wn[COL_R];
#pragma unroll
    for (int n1 = 0; n1 < COL_R; n1++) wn[n1] = __ldg(b.syn + n1 * N2 + n2);
    float2 vc[COL_R], v[COL_R];
#pragma unroll
    for (int k1 = 0; k1 < COL_R; k1++) { vc[k1] = B2[(long long)k1 * N2]; v[k1] = B1[(long long)k1 * N2]; }
    if constexpr (!TWROW) {
        const float2 tw1 = __ldg(b.tw_col + N2 + n2), tw2 = __ldg(b.tw_col + 2 * N2 + n2), tw4 = __ldg(b.tw_col + 4 * N2 + n2),
                     tw8 = __ldg(b.tw_col + 8 * N2 + n2);
        apply_powers16<true>(vc, tw1, tw2, tw4, tw8);
        apply_powers16<true>(v, tw1, tw2, tw4, tw8);
    }
    Dft<COL_R, +1>::run(vc);
    Dft<COL_R, +1>::run(v);
#pragma unroll
    for (int n1 = 0; n1 < COL_R; n1++) {
        const float val[3] = {vc[n1].x * wn[n1], v[n1].x * wn[n1], v[n1].y * wn[n1]};
#pragma unroll
        for (int ch = 0; ch < 3; ch++) {
            const float acc = rg[(long long)ch * N + n1 * N2] + val[ch];
            if (n1 < COL_R / 4) outp[ch][n1 * N2 + n2] = acc;
            else rg[(long long)ch * N + (n1 - COL_R / 4) * N2] = acc;
        }
    }
#pragma unroll
    for (int n1 = COL_R - COL_R / 4; n1 < COL_R; n1++)
#pragma unroll
        for (int ch = 0; ch < 3; ch++) rg[(long long)ch * N + n1 * N2] = 0.f;
}

cudaError_t launch_col_inv_frame(const BandDev& b, const WaveArgs& w, float* ring, float* out_c, float* out_l, float* out_r,
                                 long long out_stride, int n_tracks, cudaStream_t st) {
    const int n2 = b.n_fft / COL_R;
    if (tw_in_row(b.n_fft)) col_inv_frame_kernel<true><<<dim3((n2 + 127) / 128, 1, n_tracks), 128, 0, st>>>(b, w, ring, out_c, out_l, out_r, out_stride);
    else col_inv_frame_kernel<false><<<dim3((n2 + 127) / 128, 1, n_tracks), 128, 0, st>>>(b, w, ring, out_c, out_l, out_r, out_stride);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// band summation + output mode
// ---------------------------------------------------------------------------------------------
// ws layout: [band][channel C,Ls,Rs][track][ws_seg] (ws_seg >= seg_len, multiple of 4).  Bands are added in list order, in float32, as
// center_extraction.py:503-511 does.  mode 0: C, Ls, Rs.  mode 1 (fold-down, bela/upmix.cpp:295-303,
// 487-490): out_l = sum_b (Ls_b + 0.5 C_b), out_r = sum_b (Rs_b + 0.5 C_b); out_c untouched.  mode 2: the
// band kernels already folded the centre in (SegArgs::fold): out_l = sum_b slot_l, out_r = sum_b slot_r.
// (upmix_stream_block patches parameters 5, 6, 7 -- out_c, out_l, out_r -- of this kernel's CUDA-graph node, 10 parameters
// in all: keep the order, or change BAND_SUM_* in upmix_capi.cu with it.)
template <int VEC>
__global__ void __launch_bounds__(256) band_sum_kernel(const float* __restrict__ ws, int n_bands, int n_tracks,
                                                       long long seg_len, long long ws_seg,
                                                       float* __restrict__ out_c,
                                                       float* __restrict__ out_l, float* __restrict__ out_r,
                                                       long long out_stride, int mode) {
    const int track = blockIdx.y;
    const long long band_stride = 3LL * n_tracks * ws_seg;
    const long long ch_stride = (long long)n_tracks * ws_seg;
    const long long nvec = seg_len / VEC;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
         i += (long long)gridDim.x * blockDim.x) {
        float c[VEC], l[VEC], r[VEC];
#pragma unroll
        for (int j = 0; j < VEC; j++) c[j] = l[j] = r[j] = 0.f;
        for (int bnd = 0; bnd < n_bands; bnd++) {
            const float* p = ws + bnd * band_stride + track * ws_seg + i * VEC;
            float vc[VEC], vl[VEC], vr[VEC];
            if (VEC == 4) {
                if (mode != 2) *reinterpret_cast<float4*>(vc) = __ldcs(reinterpret_cast<const float4*>(p));
                *reinterpret_cast<float4*>(vl) = __ldcs(reinterpret_cast<const float4*>(p + ch_stride));
                *reinterpret_cast<float4*>(vr) = __ldcs(reinterpret_cast<const float4*>(p + 2 * ch_stride));
            } else {
                vc[0] = mode != 2 ? p[0] : 0.f; vl[0] = p[ch_stride]; vr[0] = p[2 * ch_stride];
            }
#pragma unroll
            for (int j = 0; j < VEC; j++) {
                if (mode == 0) { c[j] += vc[j]; l[j] += vl[j]; r[j] += vr[j]; }
                else if (mode == 1) { l[j] += vl[j] + 0.5f * vc[j]; r[j] += vr[j] + 0.5f * vc[j]; }
                else { l[j] += vl[j]; r[j] += vr[j]; }               // already folded by the band kernels
            }
        }
        const long long o = track * out_stride + i * VEC;
        if (VEC == 4) {
            if (mode == 0) __stcs(reinterpret_cast<float4*>(out_c + o), *reinterpret_cast<float4*>(c));
            __stcs(reinterpret_cast<float4*>(out_l + o), *reinterpret_cast<float4*>(l));
            __stcs(reinterpret_cast<float4*>(out_r + o), *reinterpret_cast<float4*>(r));
        } else {
            if (mode == 0) out_c[o] = c[0];
            out_l[o] = l[0];
            out_r[o] = r[0];
        }
    }
}

// Block streaming: stage = history ++ new block, history <- last D samples of stage, per channel and track (one CTA
// each: the barrier orders the in-place shift of the history).  Replaces six cudaMemcpy2DAsync per block.
// (upmix_stream_block patches parameters 2 and 3 -- in_l, in_r -- of this kernel's CUDA-graph node, 8 parameters in all:
// keep the order, or change STREAM_STAGE_* in upmix_capi.cu with it.)
__global__ void __launch_bounds__(1024) stream_stage_kernel(float* __restrict__ hist, float* __restrict__ stage,
                                                            const float* __restrict__ in_l, const float* __restrict__ in_r,
                                                            long long in_stride, int D, int n_new, int n_tracks) {
    const int ch = blockIdx.x, track = blockIdx.y;
    const int span = D + n_new;
    float* __restrict__ h = hist + ((long long)ch * n_tracks + track) * D;
    float* __restrict__ s = stage + ((long long)ch * n_tracks + track) * span;
    const float* __restrict__ src = (ch ? in_r : in_l) + (long long)track * in_stride;
#pragma unroll 4
    for (int i = threadIdx.x; i < span; i += blockDim.x) s[i] = i < D ? h[i] : __ldg(src + (i - D));
    __syncthreads();
#pragma unroll 4
    for (int j = threadIdx.x; j < D; j += blockDim.x) h[j] = s[n_new + j];
}

// ---------------------------------------------------------------------------------------------
// main.py's tail on the device (main.py:85-97, 110-157): peak of the three outputs, then one scale
// factor and the export mix written as interleaved stereo.
// ---------------------------------------------------------------------------------------------
// partial[3*blockIdx.x + ch] = max |x_ch| over this block's grid-stride share (max is order-free,
// so the two-stage reduction is deterministic)
__global__ void __launch_bounds__(256) peak3_kernel(const float* __restrict__ c, const float* __restrict__ l,
                                                    const float* __restrict__ r, long long n, float* __restrict__ partial) {
    float m[3] = {0.f, 0.f, 0.f};
    const float* src[3] = {c, l, r};
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
#pragma unroll
        for (int ch = 0; ch < 3; ch++) m[ch] = fmaxf(m[ch], fabsf(__ldg(src[ch] + i)));
    }
    __shared__ float sm[3][8];
#pragma unroll
    for (int ch = 0; ch < 3; ch++) {
        float v = m[ch];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if ((threadIdx.x & 31) == 0) sm[ch][threadIdx.x >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3) {
        float v = 0.f;
        for (int w = 0; w < 8; w++) v = fmaxf(v, sm[threadIdx.x][w]);
        partial[3 * blockIdx.x + threadIdx.x] = v;
    }
}

__global__ void peak3_final_kernel(const float* __restrict__ partial, int n_blocks, float* __restrict__ out3) {
    const int ch = threadIdx.x;
    if (ch < 3) {
        float v = 0.f;
        for (int b = 0; b < n_blocks; b++) v = fmaxf(v, partial[3 * b + ch]);
        out3[ch] = v;
    }
}

// mode 0 "AB": (Ls+C+Rs, L+R) ; 1 "split": (Ls,0) (C,C) (0,Rs) ; 2 "stereo_sum": (Ls + C/2, Rs + C/2).
// Ls, C, Rs are scaled first, in float32, like main.py:95-97; out_* are interleaved stereo [n][2].
__global__ void __launch_bounds__(256) export_mix_kernel(const float* __restrict__ c, const float* __restrict__ l,
                                                         const float* __restrict__ r, const float* __restrict__ in_l,
                                                         const float* __restrict__ in_r, long long n, float scale, int mode,
                                                         float2* __restrict__ out_a, float2* __restrict__ out_b,
                                                         float2* __restrict__ out_c) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float vc = __ldg(c + i) * scale, vl = __ldg(l + i) * scale, vr = __ldg(r + i) * scale;
        if (mode == 0) {
            out_a[i] = make_float2((vl + vc) + vr, __ldg(in_l + i) + __ldg(in_r + i));
        } else if (mode == 1) {
            out_a[i] = make_float2(vl, 0.f);
            out_b[i] = make_float2(vc, vc);
            out_c[i] = make_float2(0.f, vr);
        } else {
            out_a[i] = make_float2(vl + 0.5f * vc, vr + 0.5f * vc);
        }
    }
}

// WAV edge: interleaved 16-bit PCM stereo [n][2] -> planar float32 (x / 32768: libsndfile's float view of PCM_16,
// which is what sf.read returns, main.py:43), and interleaved float32 stereo -> 16-bit PCM as sf.write(float data)
// stores it (main.py:119-153): x * 32767 rounded to nearest even -- libsndfile's float -> PCM_16 conversion scales by
// 0x7FFF when it is not asked to clip (pcm.c f2s_array; soundfile is not installed here, so this is pinned from the
// library's documented behaviour, not from a fixture).  Out-of-range values saturate here (libsndfile would wrap);
// main.py scales its outputs to the input's peak (main.py:90-97), so they do not occur on that path.  4 bytes per
// stereo sample cross PCIe in each direction.
__global__ void __launch_bounds__(256) pcm16_to_planar_kernel(const short2* __restrict__ in, long long n, float* __restrict__ l,
                                                              float* __restrict__ r, float* __restrict__ peak_partial) {
    float m = 0.f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const short2 v = in[i];
        const float a = (float)v.x * (1.f / 32768.f), b = (float)v.y * (1.f / 32768.f);
        l[i] = a;
        r[i] = b;
        m = fmaxf(m, fmaxf(fabsf(a), fabsf(b)));
    }
    __shared__ float sm[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        float v = 0.f;
        for (int w = 0; w < 8; w++) v = fmaxf(v, sm[w]);
        peak_partial[blockIdx.x] = v;
    }
}

__global__ void max_final_kernel(const float* __restrict__ partial, int n_blocks, float* __restrict__ out) {
    if (threadIdx.x == 0) {
        float v = 0.f;
        for (int b = 0; b < n_blocks; b++) v = fmaxf(v, partial[b]);
        out[0] = v;
    }
}

__global__ void __launch_bounds__(256) stereo_to_pcm16_kernel(const float2* __restrict__ in, long long n, short2* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float2 v = in[i];
        const float a = fminf(fmaxf(v.x * 32767.f, -32768.f), 32767.f);
        const float b = fminf(fmaxf(v.y * 32767.f, -32768.f), 32767.f);
        out[i] = make_short2((short)__float2int_rn(a), (short)__float2int_rn(b));
    }
}

cudaError_t launch_pcm16_to_planar(const short* in, long long n, float* l, float* r, float* partial, int n_blocks, float* peak,
                                   cudaStream_t st) {
    pcm16_to_planar_kernel<<<n_blocks, 256, 0, st>>>(reinterpret_cast<const short2*>(in), n, l, r, partial);
    max_final_kernel<<<1, 32, 0, st>>>(partial, n_blocks, peak);
    return cudaGetLastError();
}

cudaError_t launch_stereo_to_pcm16(const float* in, long long n, short* out, cudaStream_t st) {
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    stereo_to_pcm16_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float2*>(in), n, reinterpret_cast<short2*>(out));
    return cudaGetLastError();
}

cudaError_t launch_peak3(const float* c, const float* l, const float* r, long long n, float* partial, int n_blocks,
                         float* out3, cudaStream_t st) {
    peak3_kernel<<<n_blocks, 256, 0, st>>>(c, l, r, n, partial);
    peak3_final_kernel<<<1, 32, 0, st>>>(partial, n_blocks, out3);
    return cudaGetLastError();
}

cudaError_t launch_export_mix(const float* c, const float* l, const float* r, const float* in_l, const float* in_r,
                              long long n, float scale, int mode, float* out_a, float* out_b, float* out_c, cudaStream_t st) {
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    if (blocks < 1) blocks = 1;
    export_mix_kernel<<<(unsigned)blocks, 256, 0, st>>>(c, l, r, in_l, in_r, n, scale, mode, reinterpret_cast<float2*>(out_a),
                                                        reinterpret_cast<float2*>(out_b), reinterpret_cast<float2*>(out_c));
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// Causal FIR filter (filter_design.py:54-59: scipy.signal.lfilter(taps, 1.0, wave)):
//   y[i] = sum_k taps[k] * x[i - k],  x[i < 0] = 0,  same length as x.
// A CTA produces FIR_TILE consecutive outputs of one track: the input span and the reversed taps sit in
// shared memory; a thread owns 8 consecutive outputs and slides a 16-sample register window over the
// span, 8 taps per step (two 128-bit loads of samples, two broadcast 128-bit loads of taps, 64 FMAs).
// ---------------------------------------------------------------------------------------------
constexpr int FIR_THREADS = 256, FIR_PER_THREAD = 8, FIR_TILE = FIR_THREADS * FIR_PER_THREAD;
__global__ void __launch_bounds__(FIR_THREADS) fir_kernel(const float* __restrict__ x, long long n, long long x_stride,
                                                          const float* __restrict__ taps, int n_taps, int kpad,
                                                          float* __restrict__ y, long long y_stride) {
    extern __shared__ __align__(16) float fir_smem[];
    float* hr = fir_smem;                 // [kpad]  hr[m] = taps[n_taps-1-m], zero beyond n_taps
    float* xs = fir_smem + kpad;          // [FIR_TILE + kpad + 8]  xs[t] = x[i0 - (n_taps-1) + t]
    const int tid = threadIdx.x;
    const long long i0 = (long long)blockIdx.x * FIR_TILE;
    const float* __restrict__ xt = x + (long long)blockIdx.y * x_stride;
    float* __restrict__ yt = y + (long long)blockIdx.y * y_stride;
    for (int m = tid; m < kpad; m += FIR_THREADS) hr[m] = m < n_taps ? __ldg(taps + (n_taps - 1 - m)) : 0.f;
    const int span = FIR_TILE + kpad + 8;
    for (int t = tid; t < span; t += FIR_THREADS) {
        const long long i = i0 - (n_taps - 1) + t;
        xs[t] = (i >= 0 && i < n) ? __ldg(xt + i) : 0.f;
    }
    __syncthreads();
    const int o = tid * FIR_PER_THREAD;
    float acc[FIR_PER_THREAD];
#pragma unroll
    for (int j = 0; j < FIR_PER_THREAD; j++) acc[j] = 0.f;
    float w[16];
    *reinterpret_cast<float4*>(w) = *reinterpret_cast<const float4*>(xs + o);
    *reinterpret_cast<float4*>(w + 4) = *reinterpret_cast<const float4*>(xs + o + 4);
#pragma unroll 2
    for (int m0 = 0; m0 < kpad; m0 += 8) {
        *reinterpret_cast<float4*>(w + 8) = *reinterpret_cast<const float4*>(xs + o + m0 + 8);
        *reinterpret_cast<float4*>(w + 12) = *reinterpret_cast<const float4*>(xs + o + m0 + 12);
        float h[8];
        *reinterpret_cast<float4*>(h) = *reinterpret_cast<const float4*>(hr + m0);
        *reinterpret_cast<float4*>(h + 4) = *reinterpret_cast<const float4*>(hr + m0 + 4);
#pragma unroll
        for (int mm = 0; mm < 8; mm++)
#pragma unroll
            for (int j = 0; j < FIR_PER_THREAD; j++) acc[j] = fmaf(h[mm], w[j + mm], acc[j]);
#pragma unroll
        for (int j = 0; j < 8; j++) w[j] = w[j + 8];
    }
#pragma unroll
    for (int j = 0; j < FIR_PER_THREAD; j++)
        if (i0 + o + j < n) yt[i0 + o + j] = acc[j];
}

cudaError_t launch_fir(const float* x, long long n, int n_tracks, long long x_stride, const float* taps, int n_taps, float* y,
                       long long y_stride, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const int kpad = (n_taps + 7) / 8 * 8;
    const int smem = (kpad + FIR_TILE + kpad + 8) * (int)sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(fir_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    fir_kernel<<<dim3((unsigned)((n + FIR_TILE - 1) / FIR_TILE), n_tracks), FIR_THREADS, smem, st>>>(x, n, x_stride, taps, n_taps, kpad, y,
                                                                                                      y_stride);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// FP32 peak probe: 8 independent FMA chains per thread, enough warps to fill every SM.  Used by
// bench.py for the roofline denominator (MEASURED_PEAKS.json has no FP32 figure).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fma_peak_kernel(float* out, int iters, float b, float c) {
    float a[8];
#pragma unroll
    for (int j = 0; j < 8; j++) a[j] = (float)(threadIdx.x + j) * 1e-3f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int j = 0; j < 8; j++) a[j] = fmaf(a[j], b, c);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 8; j++) s += a[j];
    if (s == 123.456f) out[blockIdx.x] = s;       // never true in practice; keeps the chains alive
}

cudaError_t launch_fma_peak(float* out, int blocks, int iters, cudaStream_t st) {
    fma_peak_kernel<<<blocks, 256, 0, st>>>(out, iters, 0.999f, 1e-4f);
    return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------
// launchers (called from upmix_capi.cu)
// ---------------------------------------------------------------------------------------------
static unsigned long long g_launches = 0;
unsigned long long& launch_counter() { return g_launches; }
void launch_count_add(unsigned long long n) { g_launches += n; }     // kernels replayed by a CUDA graph
unsigned long long launch_count(bool reset) {
    const unsigned long long v = g_launches;
    if (reset) g_launches = 0;
    return v;
}
// The fused kernel is compiled in four translation units (upmix_fused_*.cu: 24 instantiations each take a while);
// each defines the launcher of its sizes.
cudaError_t launch_band_fused_64_512(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st);
cudaError_t launch_band_fused_1024_2048(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st);
cudaError_t launch_band_fused_4096(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st);
cudaError_t launch_band_fused_8192(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st);
cudaError_t launch_band_fused(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st) {
    g_launches++;
    if (b.n_fft <= 512) return launch_band_fused_64_512(b, a, n_runs, n_tracks, st);
    if (b.n_fft <= 2048) return launch_band_fused_1024_2048(b, a, n_runs, n_tracks, st);
    if (b.n_fft == 4096) return launch_band_fused_4096(b, a, n_runs, n_tracks, st);
    return launch_band_fused_8192(b, a, n_runs, n_tracks, st);
}

// radix plans, for the host-side twiddle tables
void fused_plans(int n_fft, int* fwd, int* inv, int* half) {
    *fwd = *inv = *half = 0;
    switch (n_fft) {
#define UPMIX_CASE(N_) case N_: *fwd = FusedCfg<N_>::FWD; *inv = FusedCfg<N_>::INV; *half = FusedCfg<N_>::HALF; break;
        UPMIX_CASE(64) UPMIX_CASE(128) UPMIX_CASE(256) UPMIX_CASE(512) UPMIX_CASE(1024) UPMIX_CASE(2048)
        UPMIX_CASE(4096) UPMIX_CASE(8192)
#undef UPMIX_CASE
        default: break;
    }
}
int fused_ctas_per_sm(int n_fft) {
    switch (n_fft) {
#define UPMIX_CASE(N_) case N_: return FusedCfg<N_>::MINB;
        UPMIX_CASE(64) UPMIX_CASE(128) UPMIX_CASE(256) UPMIX_CASE(512) UPMIX_CASE(1024) UPMIX_CASE(2048)
        UPMIX_CASE(4096) UPMIX_CASE(8192)
#undef UPMIX_CASE
        default: return 1;
    }
}
int row_plan(int n2) {
    switch (n2) {
        case 1024: return RowCfg<1024>::PLAN;
        case 2048: return RowCfg<2048>::PLAN;
        case 4096: return RowCfg<4096>::PLAN;
        default: return 0;
    }
}

int fused_smem_bytes(int n_fft) {
    switch (n_fft) {
        case 64: return FusedCfg<64>::SMEM;
        case 128: return FusedCfg<128>::SMEM;
        case 256: return FusedCfg<256>::SMEM;
        case 512: return FusedCfg<512>::SMEM;
        case 1024: return FusedCfg<1024>::SMEM;
        case 2048: return FusedCfg<2048>::SMEM;
        case 4096: return FusedCfg<4096>::SMEM;
        case 8192: return FusedCfg<8192>::SMEM;
        default: return -1;
    }
}

cudaError_t launch_col_fwd(const BandDev& b, const SegArgs& a, const WaveArgs& w, int n_tracks, cudaStream_t st) {
    const int n2 = b.n_fft / COL_R;
    if (tw_in_row(b.n_fft)) col_fwd_kernel<true><<<dim3((n2 + 127) / 128, w.n_frames, n_tracks), 128, 0, st>>>(b, a, w);
    else col_fwd_kernel<false><<<dim3((n2 + 127) / 128, w.n_frames, n_tracks), 128, 0, st>>>(b, a, w);
    g_launches++;
    return cudaGetLastError();
}

template <int N2, bool TWROW>
static cudaError_t launch_row_full(const BandDev& b, const WaveArgs& w, int n_tracks, cudaStream_t st) {
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(row_mask_kernel<N2, TWROW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             RowCfg<N2>::SMEM);
        if (e != cudaSuccess) return e;
        attr_done[dev & 63] = true;
    }
    row_mask_kernel<N2, TWROW><<<dim3(COL_R / 2, w.n_frames / 2, n_tracks), RowCfg<N2>::T, RowCfg<N2>::SMEM, st>>>(b, w);
    g_launches++;
    return cudaGetLastError();
}
template <int N2, bool TWROW>
static cudaError_t launch_row_pruned(const BandDev& b, const WaveArgs& w, int n_tracks, cudaStream_t st) {
    constexpr int SMEM = (PADSZ<RowCfg<N2>::PLAN>() + 12 * ROW_K) * (int)sizeof(float2);
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(row_mask_pruned_kernel<N2, TWROW>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) return e;
        attr_done[dev & 63] = true;
    }
    row_mask_pruned_kernel<N2, TWROW><<<dim3(COL_R / 2, w.n_frames / 2, n_tracks), N2 / 16, SMEM, st>>>(b, w);
    g_launches++;
    return cudaGetLastError();
}
// The pruned row kernel needs every non-zero gain below bin 16*ROW_K (UPMIX_ROW_PRUNE=0 forces the full one).
template <int N2>
static cudaError_t launch_row_n(const BandDev& b, const WaveArgs& w, int n_tracks, cudaStream_t st) {
    static const bool allow = [] { const char* e = getenv("UPMIX_ROW_PRUNE"); return !(e && atoi(e) == 0); }();
    const bool twrow = tw_in_row(b.n_fft);
    if (allow && b.max_bin < COL_R * ROW_K)
        return twrow ? launch_row_pruned<N2, true>(b, w, n_tracks, st) : launch_row_pruned<N2, false>(b, w, n_tracks, st);
    return twrow ? launch_row_full<N2, true>(b, w, n_tracks, st) : launch_row_full<N2, false>(b, w, n_tracks, st);
}

cudaError_t launch_row_mask(const BandDev& b, const WaveArgs& w, int n_tracks, cudaStream_t st) {
    switch (b.n_fft / COL_R) {
        case 1024: return launch_row_n<1024>(b, w, n_tracks, st);
        case 2048: return launch_row_n<2048>(b, w, n_tracks, st);
        case 4096: return launch_row_n<4096>(b, w, n_tracks, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_col_inv_ola(const BandDev& b, const SegArgs& a, const WaveArgs& w, int n_runs, int n_tracks,
                               cudaStream_t st) {
    const int n2 = b.n_fft / COL_R;
    const dim3 grid((n2 + 127) / 128, n_runs, n_tracks);
    if (tw_in_row(b.n_fft)) {
        if (a.accum) col_inv_ola_kernel<true, true><<<grid, 128, 0, st>>>(b, a, w);
        else col_inv_ola_kernel<true, false><<<grid, 128, 0, st>>>(b, a, w);
    } else {
        if (a.accum) col_inv_ola_kernel<false, true><<<grid, 128, 0, st>>>(b, a, w);
        else col_inv_ola_kernel<false, false><<<grid, 128, 0, st>>>(b, a, w);
    }
    g_launches++;
    return cudaGetLastError();
}

cudaError_t launch_stream_stage(float* hist, float* stage, const float* in_l, const float* in_r, long long in_stride, int D,
                                int n_new, int n_tracks, cudaStream_t st) {
    g_launches++;
    stream_stage_kernel<<<dim3(2, n_tracks), 1024, 0, st>>>(hist, stage, in_l, in_r, in_stride, D, n_new, n_tracks);
    return cudaGetLastError();
}

cudaError_t launch_band_sum(const float* ws, int n_bands, int n_tracks, long long seg_len, long long ws_seg,
                            float* out_c, float* out_l, float* out_r, long long out_stride, int mode,
                            cudaStream_t st) {
    if (seg_len <= 0) return cudaSuccess;
    const bool vec = (seg_len % 4 == 0) && (ws_seg % 4 == 0) && (out_stride % 4 == 0 || n_tracks == 1) &&
                     ((reinterpret_cast<uintptr_t>(ws) | reinterpret_cast<uintptr_t>(out_l) |
                       reinterpret_cast<uintptr_t>(out_r) | (mode == 0 ? reinterpret_cast<uintptr_t>(out_c) : 0)) % 16 == 0);
    const long long nvec = vec ? seg_len / 4 : seg_len;
    long long blocks = (nvec + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    dim3 grid((unsigned)blocks, n_tracks);
    g_launches++;
    if (vec) band_sum_kernel<4><<<grid, 256, 0, st>>>(ws, n_bands, n_tracks, seg_len, ws_seg, out_c, out_l, out_r, out_stride, mode);
    else band_sum_kernel<1><<<grid, 256, 0, st>>>(ws, n_bands, n_tracks, seg_len, ws_seg, out_c, out_l, out_r, out_stride, mode);
    return cudaGetLastError();
}

}  // namespace upmix

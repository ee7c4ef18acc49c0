// The fused single-CTA band kernel (sizes 64 .. 8192) and its per-size configuration.  Included by
// upmix_kernels.cu (configuration queries only) and by the upmix_fused_*.cu translation units, each of which
// instantiates the kernel for some sizes.
#pragma once
#include <stdlib.h>

#include "fft_device.cuh"
#include "upmix_kernels.cuh"

namespace upmix {

#ifndef UPMIX_REG_CAP
#define UPMIX_REG_CAP 255
#endif
#ifndef UPMIX_MASK_CH
#define UPMIX_MASK_CH 2           // mask rounds per chunk (2 bin quadruples = 4 masks in flight per thread)
#endif
#ifndef UPMIX_TMA_MIN_N
#define UPMIX_TMA_MIN_N 1024      // frames of this size and larger are staged by TMA bulk copies (measured again with the
                                  // final kernels: 1024 4.81 -> 4.70 ms per band-hour; 512 4.81 -> 4.85, 256 5.43 -> 5.61)
#endif

// ---------------------------------------------------------------------------------------------
// fused single-CTA band kernel
// ---------------------------------------------------------------------------------------------
// Per-size configuration of the fused kernel: radix plans of the N-point and N/2-point transforms,
// threads per CTA (= per frame in flight) and the CTAs per SM the register budget is sized for.
template <int N> struct FusedCfg;
#ifndef UPMIX_MEGA
#define UPMIX_MEGA 1              // fuse last forward pass, mask and first inverse passes (mega_phase) where the plans line up
#endif
constexpr int cmax(int a, int b) { return a > b ? a : b; }
#define UPMIX_FUSED_CFG_X(N_, ...) UPMIX_FUSED_CFG(N_, __VA_ARGS__)
#define UPMIX_FUSED_CFG(N_, FWD_, INV_, HALF_, T_, MINB_)                                                \
    template <> struct FusedCfg<N_> {                                                                     \
        static constexpr int FWD = FWD_, INV = INV_, HALF = HALF_, T = T_, MINB = MINB_;                  \
        static constexpr bool TMA = N_ >= UPMIX_TMA_MIN_N;                                               \
        /* the fused middle needs: inverse starts with the forward's last radix RL, the half transform   \
           with RL/2, and one butterfly pair per thread */                                               \
        static constexpr int RL = fft_radix(FWD_, fft_num_passes(FWD_) - 1);                              \
        static constexpr bool MEGA = UPMIX_MEGA && RL >= 4 && fft_radix(INV_, 0) == RL && 2 * T_ * RL == N_; \
        static constexpr bool FUSE_HALF = MEGA && 2 * fft_radix(HALF_, 0) == RL;                          \
        /* registers per thread: the share of the file MINB co-resident CTAs leave, but never the whole \
           file for one CTA -- UPMIX_REG_CAP keeps room for CTAs of other pipelines on the same SM */ \
        static constexpr int MAXREG = (65536 / (T_ * MINB_) / 8 * 8) > UPMIX_REG_CAP ? UPMIX_REG_CAP : (65536 / (T_ * MINB_) / 8 * 8);                            \
        static_assert(fft_size(FWD_) == N_ && fft_size(INV_) == N_ && fft_size(HALF_) == N_ / 2, "plan does not match the size"); \
        static constexpr int ZSZ = cmax(PADSZ<FWD_>(), PADSZ<INV_>());                                    \
        static constexpr int SMEM = (ZSZ + PADSZ<HALF_>()) * (int)sizeof(float2) + 3 * N_ * (int)sizeof(float); \
    };
// Chosen by measurement on B200 (profiles/band_bench.py; see profiles/r01_tuning.md): 16-32 points per
// thread, every lane busy in every pass, and enough registers to avoid spills even if that leaves 8-12
// warps per SM.  Fields: forward plan, inverse plan, half-size plan, threads, CTAs per SM.  From 256 points
// up the forward plan ends with the radix RL = points per thread / 2, the inverse plan starts with it and the
// half plan with RL/2, so the middle of the frame runs fused in registers (mega_phase).  Measured, ms per
// band-hour, separate passes with the previous plans / fused middle / fused middle but the half-size
// transform on its own (its spectrum stored, any plan): 256: 6.81 / 5.91 / 5.91; 512: 5.15 / 5.52 / 5.79;
// 1024: 5.09 / 6.04 / 5.40; 2048: 5.80 / 5.60 / 5.75; 4096: 5.77 / 5.67 / 5.64; 8192: 6.49 / 8.10 (spills) /
// 6.24.  512 and 1024 points (16 points per thread, RL = 8) keep the separate passes: their best plans
// end with a cheap radix-4 pass, which the fusion cannot use.
#ifndef UPMIX_CFG_64
#define UPMIX_CFG_64 mkplan(8, 8), mkplan(8, 8), mkplan(8, 4), 32, 12
#endif
UPMIX_FUSED_CFG_X(64, UPMIX_CFG_64)
#ifndef UPMIX_CFG_128
#define UPMIX_CFG_128 mkplan(8, 4, 4), mkplan(8, 4, 4), mkplan(4, 4, 4), 32, 12
#endif
UPMIX_FUSED_CFG_X(128, UPMIX_CFG_128)
#ifndef UPMIX_CFG_256
#define UPMIX_CFG_256 mkplan(8, 8, 4), mkplan(4, 8, 8), mkplan(8, 4, 4), 32, 12
#endif
UPMIX_FUSED_CFG_X(256, UPMIX_CFG_256)
#ifndef UPMIX_CFG_512
#define UPMIX_CFG_512 mkplan(16, 8, 4), mkplan(16, 8, 4), mkplan(16, 4, 4), 32, 12
#endif
UPMIX_FUSED_CFG_X(512, UPMIX_CFG_512)
#ifndef UPMIX_CFG_1024
#define UPMIX_CFG_1024 mkplan(16, 8, 8), mkplan(16, 16, 4), mkplan(16, 8, 4), 64, 6
#endif
UPMIX_FUSED_CFG_X(1024, UPMIX_CFG_1024)
#ifndef UPMIX_CFG_2048
#define UPMIX_CFG_2048 mkplan(16, 8, 16), mkplan(16, 8, 16), mkplan(8, 8, 16), 64, 4
#endif
UPMIX_FUSED_CFG_X(2048, UPMIX_CFG_2048)
#ifndef UPMIX_CFG_4096
#define UPMIX_CFG_4096 mkplan(16, 16, 16), mkplan(16, 16, 16), mkplan(16, 16, 8), 128, 2
#endif
UPMIX_FUSED_CFG_X(4096, UPMIX_CFG_4096)
#ifndef UPMIX_CFG_8192
#define UPMIX_CFG_8192 mkplan(32, 16, 16), mkplan(16, 16, 32), mkplan(16, 16, 16), 256, 1
#endif
UPMIX_FUSED_CFG_X(8192, UPMIX_CFG_8192)

// MODE selects the mask variant at compile time (the mask is ~45 % of a dense band's instructions):
//   MODE_PLAIN  one band, Ls/C/Rs out;  MODE_FOLD  one band, centre folded per bin (SegArgs::fold);
//   MODE_MERGED several bands share the pipeline (per-bin loop over their gains; fold read at run time).
enum { MODE_PLAIN = 0, MODE_FOLD = 1, MODE_MERGED = 2 };
// FE ("fused emit", hop = N/4, no fold-down epilogue): the last inverse passes treat the four hops of a frame
// differently, at compile time -- the newest hop is stored into the overlap-add ring, the two middle ones are
// accumulated, and the oldest, which this frame completes, goes from registers straight to the output (plus
// what the output already holds when the band accumulates).  No copy-out pass, no clearing, a quarter less
// ring traffic: the ring and the copy-out were 15-18 % of the kernel's shared-memory wavefronts.
#ifndef UPMIX_DIRECT_EMIT
#define UPMIX_DIRECT_EMIT 1
#endif
#ifndef UPMIX_FE_MIN_N
#define UPMIX_FE_MIN_N 256        // smallest size that uses it (measured; see launch_fused_nm)
#endif
template <int N, int MODE, bool FE>
__global__ void __launch_bounds__(FusedCfg<N>::T) __maxnreg__(FusedCfg<N>::MAXREG) band_fused_kernel(const BandDev b, const SegArgs a) {
    constexpr int T = FusedCfg<N>::T;
    constexpr int M = N / 2;
    constexpr int PF = FusedCfg<N>::FWD, PI = FusedCfg<N>::INV, PH = FusedCfg<N>::HALF;
    constexpr bool MEGA = FusedCfg<N>::MEGA;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* Z = reinterpret_cast<float2*>(smem_raw);
    float2* Cz = Z + FusedCfg<N>::ZSZ;
    float* ring = reinterpret_cast<float*>(Cz + PADSZ<PH>());   // [3][N]: C, Ls, Rs

    const int tid = threadIdx.x;
    const int H = b.hop;
    const int K = N / H;
    const bool fold = MODE == MODE_MERGED ? a.fold != 0 : MODE == MODE_FOLD;
    const long long h0 = a.hop_begin + (long long)blockIdx.x * a.hops_per_run;
    const long long h1 = min(h0 + (long long)a.hops_per_run, a.hop_end);
    if (h0 >= h1) return;
    const int track = blockIdx.y;
    const float* __restrict__ inl = a.in_l + (long long)track * a.in_stride;
    const float* __restrict__ inr = a.in_r + (long long)track * a.in_stride;
    float* outp[3] = {a.out_c + (long long)track * a.out_stride, a.out_l + (long long)track * a.out_stride,
                      a.out_r + (long long)track * a.out_stride};
    float* st_ring = a.state ? a.state + (long long)track * 3 * N : nullptr;

    const bool st_load = st_ring && blockIdx.x == 0;
    long long f_begin;
    if (st_load) {
        f_begin = h0;
        const int nb0 = (int)(h0 % K) * H;          // slot 0 of the saved ring = first sample of frame h0
        for (int i = tid; i < 3 * N; i += T) {
            const int ch = i / N, n = i - ch * N;
            ring[ch * N + ((nb0 + n) & (N - 1))] = st_ring[i];
        }
    } else {
        f_begin = max(0LL, h0 - (K - 1));
        for (int i = tid; i < 3 * N; i += T) ring[i] = 0.f;
    }
    __syncthreads();

    const float* __restrict__ ana = b.ana;
    const float* __restrict__ syn = b.syn;
    const float* __restrict__ gain = b.gain;
    const float2* __restrict__ tw = b.tw_fft;
    const float2* __restrict__ twi = b.tw_inv;
    const float2* __restrict__ twh = b.tw_half;
    const float2* __restrict__ twp = b.tw_pack;

    // Frame staging, two variants (FusedCfg<N>::TMA, chosen by measurement):
    //  * TMA (N >= 1024): the raw samples of a frame land in the (then idle) transform buffer Z as two
    //    planar arrays SL[N], SR[N]: one elected thread issues two bulk copies (cp.async.bulk + mbarrier)
    //    for the NEXT frame as soon as the inverse transform of Ls + i Rs has released Z, so the copy
    //    overlaps the centre's inverse transform and the copy-out; pass 0 of the next frame reads SL/SR
    //    from shared memory.  Frames not wholly inside [in_begin, in_end) (track / shard edges) or not
    //    16-byte aligned are filled by all threads instead, zero outside the available samples.
    //  * registers (small N, where a frame turns around too quickly for the extra barrier to pay): the
    //    next frame's samples are requested before the current frame's copy-out and wait in registers;
    //    xin[it][r] is point n = tid + it*T + r*NB0, exactly what pass 0 asks for.
    // Either way the analysis-window values are requested at the same time and wait in registers.
    constexpr bool TMA = FusedCfg<N>::TMA;
    constexpr int R0 = fft_radix(PF, 0);
    constexpr int NB0 = N / R0;
    constexpr int IT0 = (NB0 + T - 1) / T;
    __shared__ __align__(8) uint64_t in_bar;
    float* SL = reinterpret_cast<float*>(Z);
    float* SR = SL + N;
    float xw[IT0][R0];
    float2 xin[TMA ? 1 : IT0][TMA ? 1 : R0];
    const bool tma_ok = TMA && ((reinterpret_cast<uintptr_t>(inl) | reinterpret_cast<uintptr_t>(inr)) & 15) == 0 &&
                        (a.in_begin & 3) == 0 && (H & 3) == 0;
    uint32_t in_phase = 0;
    bool in_by_tma = false;
    if (TMA) {
        if (tid == 0) mbar_init(&in_bar, 1);
        __syncthreads();
    }
    auto stage = [&](long long fr) {          // every thread calls this; TMA: at a point where nobody uses Z
        const long long s0n = fr * H;
        const float* __restrict__ pl = inl + (s0n - a.in_begin);     // pl[n] is sample s0n + n
        const float* __restrict__ pr = inr + (s0n - a.in_begin);
        const bool whole = s0n >= a.in_begin && s0n + N <= a.in_end; // CTA-uniform
        const int lo_n = (int)max(0LL, min((long long)N, a.in_begin - s0n));
        const int hi_n = (int)max(0LL, min((long long)N, a.in_end - s0n));
        if constexpr (TMA) {
            in_by_tma = tma_ok && whole;
            if (in_by_tma) {
                if (tid == 0) {
                    fence_proxy_async_smem();
                    mbar_expect_tx(&in_bar, 2u * N * (uint32_t)sizeof(float));
                    tma_load_1d(SL, pl, N * (uint32_t)sizeof(float), &in_bar);
                    tma_load_1d(SR, pr, N * (uint32_t)sizeof(float), &in_bar);
                }
            } else {
                for (int n = tid; n < N; n += T) {
                    const bool ok = n >= lo_n && n < hi_n;
                    SL[n] = ok ? __ldg(pl + n) : 0.f;
                    SR[n] = ok ? __ldg(pr + n) : 0.f;
                }
            }
#pragma unroll
            for (int it = 0; it < IT0; it++) {
                const int j = tid + it * T;
                if (NB0 % T == 0 || j < NB0) {
#pragma unroll
                    for (int r = 0; r < R0; r++) xw[it][r] = __ldg(ana + j + r * NB0);
                }
            }
        } else if (whole) {
#pragma unroll
            for (int it = 0; it < IT0; it++) {
                const int j = tid + it * T;
                if (NB0 % T == 0 || j < NB0) {
#pragma unroll
                    for (int r = 0; r < R0; r++) {
                        xw[it][r] = __ldg(ana + j + r * NB0);
                        xin[it][r] = make_float2(__ldg(pl + j + r * NB0), __ldg(pr + j + r * NB0));
                    }
                }
            }
        } else {
            // Samples outside [in_begin, in_end) -- before the track, past its end, or another shard's --
            // read as zero: their loads are clamped to a valid sample and their window value is taken
            // from the zero stored after the table (ana[N] == 0), so the loads stay unconditional.
            const bool any = hi_n > lo_n;
            const int lo_c = min(lo_n, N - 1);
#pragma unroll 1
            for (int it = 0; it < IT0; it++) {
                const int j = tid + it * T;
                if (NB0 % T == 0 || j < NB0) {
#pragma unroll
                    for (int r = 0; r < R0; r++) {
                        const int n = j + r * NB0;
                        const bool ok = n >= lo_n && n < hi_n;
                        xw[it][r] = __ldg(ana + (ok ? n : N));
                        xin[it][r] = any ? make_float2(__ldg(pl + (ok ? n : lo_c)), __ldg(pr + (ok ? n : lo_c)))
                                         : make_float2(0.f, 0.f);
                    }
                }
            }
        }
    };
    stage(f_begin);
    if (TMA && !in_by_tma) __syncthreads();

    for (long long f = f_begin; f < h1; ++f) {
        const long long s0 = f * H;
        const int base = (int)(f % K) * H;

        // ---- where this frame's finished hop goes ---------------------------------------------------
        // accum: the samples are added to what the output already holds (the bands before this one, in
        // band order).  Those values are requested here, at the top of the frame, and wait in
        // registers (ITE float4 per channel) through the whole frame, so the copy-out does not stall on HBM
        // (requested after the Ls + iRs inverse they were still in flight at the copy-out: 6 % of the
        // 1024-point kernel's stall samples sat on that one addition).
        const bool emit = f >= h0;
        const int e_lo = (int)max(0LL, min((long long)H, a.seg_begin - s0));
        const int e_hi = emit ? (int)max(0LL, min((long long)H, a.seg_end - s0)) : 0;
        const bool accum = a.accum != 0;
        const bool vec_out = e_lo == 0 && e_hi == H && (H % 4 == 0) && !a.mix &&
                             (((reinterpret_cast<uintptr_t>(outp[1] + (s0 - a.out_begin)) | reinterpret_cast<uintptr_t>(outp[2] + (s0 - a.out_begin)) |
                                (fold ? 0 : reinterpret_cast<uintptr_t>(outp[0] + (s0 - a.out_begin)))) & 15) == 0);   // CTA-uniform
        constexpr int ITE = FE ? 1 : (N / 4 + 4 * T - 1) / (4 * T);   // float4 per thread and channel at hop = N/4
        // (small sizes run 16 CTAs per SM on a 128-register budget: they load at the copy-out instead)
        const bool pre = !FE && N >= 1024 && accum && vec_out && H <= 4 * T * ITE;
        float4 prev[3][ITE];
        auto request_prev = [&]() {
            if (pre) {
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    if (ch == 0 && fold) continue;
                    const float* __restrict__ po = outp[ch] + (s0 - a.out_begin);
#pragma unroll
                    for (int it = 0; it < ITE; it++) {
                        const int i = (tid + it * T) * 4;
                        if (i < H) prev[ch][it] = __ldcs(reinterpret_cast<const float4*>(po + i));
                    }
                }
            }
        };
        // 16 points per thread: 12 registers held through the whole frame; 32 points per thread: 24, requested
        // once the fused middle of the frame (the register peak) is over
        constexpr bool PREV_AT_TOP = N <= 1024;
        if (!FE && PREV_AT_TOP) request_prev();
        if (FE && accum && emit) {
            // the last passes read the output's previous sums themselves; pull their lines towards L2 now
            for (int line = tid; line < (H + 31) / 32; line += T) {
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    if (ch == 0 && fold) continue;
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(outp[ch] + (s0 - a.out_begin) + line * 32));
                }
            }
        }

        // ---- forward: Z = FFT_N( ana * (L + iR) ) -----------------------------------------------
        if (TMA && in_by_tma) {
            mbar_wait(&in_bar, in_phase);
            in_phase ^= 1;
        }
        auto ld_in = [&](int, int n, int it, int r) -> float2 {
            const float wn = xw[it][r];
            if constexpr (TMA) return cscale(make_float2(SL[n], SR[n]), wn);
            else return cscale(xin[it][r], wn);
        };
        auto st_z = make_store([&](int, int k, float2 v, NoAux) { Z[PAD<PF>(k)] = v; });
        if constexpr (MEGA) {
            // all forward passes but the last; the last one runs fused with the mask and the first inverse passes
            stockham_range<PF, 0, fft_num_passes(PF) - 1, -1, T, 1, TMA>(Z, tid, tw, ld_in, st_z);
            mega_phase<N, T, PF, PI, PH, MODE == MODE_MERGED, FusedCfg<N>::FUSE_HALF>(Z, Cz, tid, tw, gain, b.n_gains, b.gain_stride, twp, fold);
        } else {
        fft_smem<PF, -1, T, 1, TMA>(Z, tid, tw, ld_in, st_z);      // TMA: in place, SL/SR live inside Z

        // ---- split / gain / mask; Y1 = Ls + i*Rs in place, C packed for the half-size inverse ---
        // An item is the bin quadruple k, M-k and their mirrors N-k, M+k, for k = 0 .. M/2.  Items
        // 0 .. M/2-1 go in rounds of T threads, CH rounds per chunk; the last one (k = M/2, a single bin
        // pair) is one thread's epilogue, so no round runs for it alone.  A chunk whose gains are all
        // zero only stores zeros (most of a low band's spectrum); otherwise its CH*2 masks are
        // straight-line code, so their MUFU / dependency latencies overlap.
        {
            constexpr int HALF = M / 2;
            constexpr int ROUNDS = (HALF + T - 1) / T;
            // measured (same run, ms per band-hour): 2 rounds per chunk 1024 5.64 -> 5.39, 512 5.49 -> 5.32 (4 spill
            // there); 4 per chunk for the 32-points-per-thread sizes: 8192 7.08 -> 7.03, 4096 6.25 -> 6.20
            constexpr int WANT = N >= 2048 ? 2 * UPMIX_MASK_CH : UPMIX_MASK_CH;
            constexpr int CH = ROUNDS % WANT == 0 ? WANT : (ROUNDS % 2 == 0 ? 2 : 1);
            auto zero_item = [&](int k) {
                const float2 zero = make_float2(0.f, 0.f);
                Z[PAD<PF>(k)] = zero;
                Z[PAD<PF>((N - k) & (N - 1))] = zero;
                Z[PAD<PF>(M - k)] = zero;
                Z[PAD<PF>(M + k)] = zero;
                if (!fold) {
                    Cz[PAD<PH>(k)] = zero;
                    if (k > 0) Cz[PAD<PH>(M - k)] = zero;
                }
            };
            auto mask_item = [&](int k, bool live, float g1, float g2, float2 wp) {
                const int k2 = M - k;
                const int km = (N - k) & (N - 1);
                const float2 a1 = Z[PAD<PF>(k)], b1 = Z[PAD<PF>(km)];
                const float2 a2 = Z[PAD<PF>(k2)], b2 = Z[PAD<PF>(M + k)];
                float2 c1, y1, y1m, c2, y2, y2m;
                if constexpr (MODE == MODE_MERGED) {
                    mask_bin_merged(a1, b1, g1, gain + k, b.n_gains, b.gain_stride, y1, y1m, c1);
                    mask_bin_merged(a2, b2, g2, gain + k2, b.n_gains, b.gain_stride, y2, y2m, c2);
                } else {
                    mask_bin(a1, b1, g1, y1, y1m, c1);
                    mask_bin(a2, b2, g2, y2, y2m, c2);
                }
                if (fold) {
                    // (Ls + C/2) + i (Rs + C/2) = (Ls + i Rs) + (1+i) C/2: the fold-down is linear, so
                    // it is taken here and the centre needs no transform of its own
                    const float2 u1 = cadd(make_float2(c1.x, c1.x), make_float2(-c1.y, c1.y));   // (1+i) C
                    const float2 u2 = cadd(make_float2(c2.x, c2.x), make_float2(-c2.y, c2.y));
                    y1 = caxpy(u1, 0.5f, y1);
                    y1m = caxpy(make_float2(u1.y, u1.x), 0.5f, y1m);
                    y2 = caxpy(u2, 0.5f, y2);
                    y2m = caxpy(make_float2(u2.y, u2.x), 0.5f, y2m);
                }
                if (live) {
                    Z[PAD<PF>(k)] = y1;
                    Z[PAD<PF>(km)] = y1m;
                    Z[PAD<PF>(k2)] = y2;
                    Z[PAD<PF>(M + k)] = y2m;
                }
                if (!fold) {
                    // z[k] = (C[k] + conj C[M-k]) + i e^{+2 pi i k/N} (C[k] - conj C[M-k]);
                    // IFFT_M(z)[m] = c[2m] + i c[2m+1]
                    const float2 A = cadd(c1, make_float2(c2.x, -c2.y));
                    const float2 B = cadd(c1, make_float2(-c2.x, c2.y));
                    const float2 D = cmul(B, make_float2(wp.x, -wp.y));
                    if (live) {
                        Cz[PAD<PH>(k)] = cadd(A, make_float2(-D.y, D.x));
                        if (k > 0) Cz[PAD<PH>(M - k)] = cadd(make_float2(A.x, -A.y), make_float2(D.y, D.x));
                    }
                }
            };
#pragma unroll 1
            for (int it0 = 0; it0 < ROUNDS; it0 += CH) {
                float g1[CH], g2[CH];
                float2 wp[CH];
                bool any = false;
#pragma unroll
                for (int i = 0; i < CH; i++) {
                    const int k = min(tid + (it0 + i) * T, HALF - 1);
                    g1[i] = __ldg(gain + k);
                    g2[i] = __ldg(gain + M - k);
                    wp[i] = __ldg(twp + k);
                    any = any || g1[i] != 0.f || g2[i] != 0.f;   // merged tables: non-zero gains come first
                }
                if (!any) {
#pragma unroll
                    for (int i = 0; i < CH; i++) {
                        const int k = tid + (it0 + i) * T;
                        if (HALF % T == 0 || k < HALF) zero_item(k);
                    }
                    continue;
                }
#pragma unroll
                for (int i = 0; i < CH; i++) {
                    const int kq = tid + (it0 + i) * T;
                    // surplus threads (HALF not a multiple of T) recompute the last item and store nothing
                    mask_item(min(kq, HALF - 1), HALF % T == 0 || kq < HALF, g1[i], g2[i], wp[i]);
                }
            }
            if (tid == T - 1) {                          // k = M/2: bins M/2 and N - M/2, taken twice by the item code
                const float g = __ldg(gain + HALF);
                if (g == 0.f) zero_item(HALF);
                else mask_item(HALF, true, g, g, __ldg(twp + HALF));
            }
        }
        __syncthreads();
        }   // !MEGA

        if (!FE && !PREV_AT_TOP) request_prev();

        // ---- inverse transforms, synthesis window, overlap-add (oldest frame first) --------------
        auto ld_z = [&](int, int n, int, int) -> float2 { return Z[PAD<PF>(n)]; };
        auto st_lr = make_store([&](int, int n) -> float { return __ldg(syn + n); },
                                [&](int, int n, float2 v, float wn) {
                                    const int p = (base + n) & (N - 1);
                                    const float2 acc = caxpy(v, wn, make_float2(ring[N + p], ring[2 * N + p]));
                                    ring[N + p] = acc.x;
                                    ring[2 * N + p] = acc.y;
                                });
        // FE: per-output overlap-add.  Output r of a last-pass butterfly lies in hop (4 r) / R of the frame.
        constexpr int PINV = MEGA ? PI : PF;
        constexpr int RLI = fft_radix(PINV, fft_num_passes(PINV) - 1), RLH = fft_radix(PH, fft_num_passes(PH) - 1);
        static_assert(!FE || N < UPMIX_FE_MIN_N || (RLI % 4 == 0 && RLH % 4 == 0),
                      "fused emit needs last inverse radices that are multiples of 4: an output must lie in one hop of the frame");
        float* __restrict__ oC = outp[0] + (s0 - a.out_begin);
        float* __restrict__ oL = outp[1] + (s0 - a.out_begin);
        float* __restrict__ oR = outp[2] + (s0 - a.out_begin);
        struct AuxLR { float wn, pl, pr; };
        struct AuxC { float2 wn, pc; };
        auto st_lr_fe = make_store_r(
            [&](int, int n, int r) -> AuxLR {
                AuxLR x;
                x.wn = __ldg(syn + n);
                x.pl = x.pr = 0.f;
                if ((4 * r) / RLI == 0 && accum && n >= e_lo && n < e_hi) {
                    x.pl = __ldcs(oL + n);
                    x.pr = __ldcs(oR + n);
                }
                return x;
            },
            [&](int, int n, float2 v, AuxLR x, int r) {
                const int q = (4 * r) / RLI;
                const int p = (base + n) & (N - 1);
                if (q == 3) {                                            // first contribution: store
                    const float2 t = cscale(v, x.wn);
                    ring[N + p] = t.x;
                    ring[2 * N + p] = t.y;
                } else {
                    float2 acc = caxpy(v, x.wn, make_float2(ring[N + p], ring[2 * N + p]));
                    if (q == 0) {                                        // last contribution: the hop is finished
                        if (n >= e_lo && n < e_hi) {
                            if (accum) acc = make_float2(x.pl + acc.x, x.pr + acc.y);
                            __stcs(oL + n, acc.x);
                            __stcs(oR + n, acc.y);
                        }
                    } else {
                        ring[N + p] = acc.x;
                        ring[2 * N + p] = acc.y;
                    }
                }
            });
        const bool c_vec = e_lo == 0 && e_hi == H && (reinterpret_cast<uintptr_t>(oC) & 7) == 0;    // CTA-uniform
        auto st_c_fe = make_store_r(
            [&](int, int m, int r) -> AuxC {
                AuxC x;
                x.wn = __ldg(reinterpret_cast<const float2*>(syn) + m);
                x.pc = make_float2(0.f, 0.f);
                if ((4 * r) / RLH == 0 && accum) {
                    const int n = 2 * m;
                    if (c_vec) x.pc = __ldcs(reinterpret_cast<const float2*>(oC + n));
                    else {
                        if (n >= e_lo && n < e_hi) x.pc.x = __ldcs(oC + n);
                        if (n + 1 >= e_lo && n + 1 < e_hi) x.pc.y = __ldcs(oC + n + 1);
                    }
                }
                return x;
            },
            [&](int, int m, float2 v, AuxC x, int r) {
                const int q = (4 * r) / RLH;
                const int n = 2 * m;
                float2* rq = reinterpret_cast<float2*>(ring + ((base + n) & (N - 1)));
                if (q == 3) {
                    *rq = __fmul2_rn(v, x.wn);
                } else {
                    float2 acc = __ffma2_rn(v, x.wn, *rq);
                    if (q == 0) {
                        if (accum) acc = make_float2(x.pc.x + acc.x, x.pc.y + acc.y);
                        if (c_vec) __stcs(reinterpret_cast<float2*>(oC + n), acc);
                        else {
                            if (n >= e_lo && n < e_hi) oC[n] = acc.x;
                            if (n + 1 >= e_lo && n + 1 < e_hi) oC[n + 1] = acc.y;
                        }
                    } else {
                        *rq = acc;
                    }
                }
            });
        if constexpr (FE) {
            if constexpr (MEGA) stockham_range<PI, 1, fft_num_passes(PI), +1, T, 1, true>(Z, tid, twi, ld_z, st_lr_fe);
            else fft_smem<PF, +1, T, 1, true>(Z, tid, tw, ld_z, st_lr_fe);
        } else {
            if constexpr (MEGA) stockham_range<PI, 1, fft_num_passes(PI), +1, T, 1, true>(Z, tid, twi, ld_z, st_lr);
            else fft_smem<PF, +1, T, 1, true>(Z, tid, tw, ld_z, st_lr);
        }
        if (TMA && f + 1 < h1) stage(f + 1);       // Z is idle from here to the next frame's first pass
        auto ld_c = [&](int, int n, int, int) -> float2 { return Cz[PAD<PH>(n)]; };
        auto st_c = make_store([&](int, int m) -> float2 { return __ldg(reinterpret_cast<const float2*>(syn) + m); },
                               [&](int, int m, float2 v, float2 wn) {
                                   const int p = (base + 2 * m) & (N - 1);
                                   float2* q = reinterpret_cast<float2*>(ring + p);
                                   *q = __ffma2_rn(v, wn, *q);
                               });
        if (!fold) {
            if constexpr (FE) {
                if constexpr (FusedCfg<N>::FUSE_HALF) stockham_range<PH, 1, fft_num_passes(PH), +1, T, 1, true>(Cz, tid, twh, ld_c, st_c_fe);
                else fft_smem<PH, +1, T, 1, true>(Cz, tid, twh, ld_c, st_c_fe);
            } else {
                if constexpr (FusedCfg<N>::FUSE_HALF) stockham_range<PH, 1, fft_num_passes(PH), +1, T, 1, true>(Cz, tid, twh, ld_c, st_c);
                else fft_smem<PH, +1, T, 1, true>(Cz, tid, twh, ld_c, st_c);
            }
        }

        if (!TMA && f + 1 < h1) stage(f + 1);      // register variant: request the next frame before the copy-out

        // ---- emit the hop this frame finished, clear its ring slots --------------------------------
        // mix: the fold-down epilogue Ls + 0.5 C / Rs + 0.5 C (bela/upmix.cpp:295-303).
        if constexpr (FE) continue;                     // already written by the last passes (which end with a barrier)
        if (a.mix) {
            float* __restrict__ pl = outp[1] + (s0 - a.out_begin);
            float* __restrict__ pr = outp[2] + (s0 - a.out_begin);
            float* rc = ring + base;
            float* rl = ring + N + base;
            float* rr = ring + 2 * N + base;
            for (int i = tid; i < H; i += T) {
                const float hc = 0.5f * rc[i];
                float vl = rl[i] + hc, vr = rr[i] + hc;
                rc[i] = rl[i] = rr[i] = 0.f;
                if (i >= e_lo && i < e_hi) {
                    if (accum) { vl = pl[i] + vl; vr = pr[i] + vr; }
                    pl[i] = vl;
                    pr[i] = vr;
                }
            }
        } else {
#pragma unroll
            for (int ch = 0; ch < 3; ch++) {
                if (ch == 0 && fold) continue;                                  // no centre channel when folded
                float* __restrict__ po = outp[ch] + (s0 - a.out_begin);
                float* rg = ring + ch * N + base;
                if (pre) {                                                       // previous sums already in registers
#pragma unroll
                    for (int it = 0; it < ITE; it++) {
                        const int i = (tid + it * T) * 4;
                        if (i < H) {
                            const float4 v = *reinterpret_cast<const float4*>(rg + i);
                            const float4 o = prev[ch][it];
                            *reinterpret_cast<float4*>(rg + i) = make_float4(0.f, 0.f, 0.f, 0.f);
                            __stcs(reinterpret_cast<float4*>(po + i), make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w));
                        }
                    }
                } else if (vec_out) {                                            // CTA-uniform: vector copy-out
                    for (int i = tid * 4; i < H; i += T * 4) {
                        float4 v = *reinterpret_cast<const float4*>(rg + i);
                        *reinterpret_cast<float4*>(rg + i) = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (accum) {
                            const float4 o = __ldcs(reinterpret_cast<const float4*>(po + i));
                            v = make_float4(o.x + v.x, o.y + v.y, o.z + v.z, o.w + v.w);
                        }
                        __stcs(reinterpret_cast<float4*>(po + i), v);
                    }
                } else {
                    for (int i = tid; i < H; i += T) {
                        const float v = rg[i];
                        rg[i] = 0.f;
                        if (i >= e_lo && i < e_hi) po[i] = accum ? po[i] + v : v;
                    }
                }
            }
        }
        __syncthreads();
    }
    if (st_ring && h1 == a.hop_end) {
        // re-base the ring so that slot 0 is the first unfinished sample (frame h1 starts there)
        float* st_save = a.state_out + (long long)track * 3 * N;
        const int nb = (int)(h1 % K) * H;
        for (int i = tid; i < 3 * N; i += T) {
            const int ch = i / N, n = i - ch * N;
            // FE leaves the emitted hop in the ring (the next frame overwrites it): it reads as zero in the state
            st_save[i] = FE && n >= N - H ? 0.f : ring[ch * N + ((nb + n) & (N - 1))];
        }
    }
}

template <int N, int MODE, bool FE>
static cudaError_t launch_fused_nmf(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st) {
    static bool attr_done[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!attr_done[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(band_fused_kernel<N, MODE, FE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             FusedCfg<N>::SMEM);
        if (e != cudaSuccess) return e;
        attr_done[dev & 63] = true;
    }
    band_fused_kernel<N, MODE, FE><<<dim3(n_runs, n_tracks), FusedCfg<N>::T, FusedCfg<N>::SMEM, st>>>(b, a);
    return cudaGetLastError();
}
template <int N, int MODE>
static cudaError_t launch_fused_nm(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st) {
    static const bool allow = [] { const char* e = getenv("UPMIX_DIRECT_EMIT"); return UPMIX_DIRECT_EMIT && !(e && atoi(e) == 0); }();
    // measured, ms per band-hour without / with: 8192 6.22 / 5.75, 4096 5.66 / 5.33, 2048 5.59 / 5.39, 1024 5.09 / 4.87,
    // 512 5.15 / 5.07, 256 5.91 / 6.27 with the half-size first pass fused (spills), 5.79 / 5.49 with the half transform
    // on its own ({8,4,4}): from 256 points up
    if (allow && N >= UPMIX_FE_MIN_N && b.hop * 4 == b.n_fft && !a.mix) return launch_fused_nmf<N, MODE, true>(b, a, n_runs, n_tracks, st);
    return launch_fused_nmf<N, MODE, false>(b, a, n_runs, n_tracks, st);
}
template <int N>
static cudaError_t launch_fused_n(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st) {
    if (b.n_gains > 1) return launch_fused_nm<N, MODE_MERGED>(b, a, n_runs, n_tracks, st);
    return a.fold ? launch_fused_nm<N, MODE_FOLD>(b, a, n_runs, n_tracks, st)
                  : launch_fused_nm<N, MODE_PLAIN>(b, a, n_runs, n_tracks, st);
}


}  // namespace upmix

// Frame-batched band kernel for the dense (top) band of a crossover set, sizes 256 / 512 / 1024, 75 % overlap.
//
// The single-frame fused kernel (upmix_fused.cuh) keeps one frame per CTA: its threads run along the frame, so every
// pass pays for padded, conflict-prone shared-memory strides, per-thread twiddle loads and an overlap-add ring in shared
// memory -- the L1 / shared-memory data pipe is its busiest unit (78 % at 1024 points).  Here a CTA holds SIXTEEN
// consecutive frames and the frame index is the lane: the tile is buf[point][frame] (row stride 17), so
//   * every shared-memory access of every pass is lane-contiguous (conflict-free, no index padding),
//   * pass twiddles, window samples, gains and packing twiddles are the same for the 16 lanes of a half-warp,
//   * an N-point transform is two radix-32/16 passes (1024 = 32 x 32) instead of three,
//   * the overlap-add needs no memory at all: the four frames that finish a hop sit in four neighbouring lanes, so the
//     hop is ((y[f-3] + y[f-2]) + y[f-1]) + y[f] with three warp shuffles, oldest frame first like the reference
//     (center_extraction.py:392-407); what the next tile needs from this one (three partial sums per output) stays in
//     registers, so consecutive tiles of a run need no replayed frames,
//   * the 16 finished hops of a tile are 16*hop CONSECUTIVE output samples: they are transposed through shared memory
//     and leave as fully coalesced 8-byte stores.
// Input: the 19 hops of samples a tile needs are staged with cp.async (8-byte copies, zero registers) into the part of
// the transform buffer that is idle while the centre's inverse transform runs, one tile ahead.
//
// Per tile: window, forward N-point transform of L + iR, split / gain / centre mask on bin quadruples (k, M-k, M+k,
// N-k), inverse N-point transform of Ls + iRs and inverse N/2-point transform of the packed centre, synthesis window,
// overlap-add, emit -- the whole per-band chain of center_extraction.py:353-409 for 16 frames at once.
#pragma once
#include <stdlib.h>

#include "fft_device.cuh"
#include "upmix_kernels.cuh"

namespace upmix {

#ifndef UPMIX_FB1024_F
#define UPMIX_FB1024_F 8          // frames per tile of the 1024-point kernel (16: one CTA of 512 threads per SM; 8: two of 256)
#endif

template <int N> struct FbCfg;
// RA x RB = N (forward and inverse of Ls + iRs), HA x 16 = N/2 (centre); F = frames per tile (= lanes per butterfly row);
// CTAS = CTAs per SM the smem / registers allow
template <> struct FbCfg<1024> { static constexpr int RA = 32, RB = 32, HA = 32, F = UPMIX_FB1024_F, CTAS = F == 16 ? 1 : 2; };
template <> struct FbCfg<512>  { static constexpr int RA = 32, RB = 16, HA = 16, F = 16, CTAS = 2; };
template <> struct FbCfg<256>  { static constexpr int RA = 16, RB = 16, HA = 8,  F = 16, CTAS = 4; };

// Tile layout: buf[point][frame].  Row `p` of the tile starts at fb_ro<F>(p) (in float2):
//   F = 16: p * 17 -- a half-warp is one row, the odd stride keeps every pass conflict-free;
//   F = 8:  (p & 31) * 8 + (p >> 5) * 264 -- a half-warp holds TWO rows (butterfly rows jb, jb + 1), 16 words each, which
//           must fall into different halves of the 32 banks: rows one apart do (8 float2 = 16 words), rows 32 apart --
//           the first-pass scatter j * 32 + r -- do because every block of 32 rows is followed by one row of padding
//           (264 = 33 * 8).
// Both are (p & 31) * S1 + (p >> 5) * S32 with S32 = 32 * S1 for F = 16.
template <int F> struct FbLay { static constexpr int S1 = F == 16 ? 17 : 8, S32 = F == 16 ? 32 * 17 : 33 * 8; };
template <int F> __host__ __device__ __forceinline__ constexpr int fb_ro(int p) { return (p & 31) * FbLay<F>::S1 + (p >> 5) * FbLay<F>::S32; }
// offset of `rows` rows further down, valid when the walk does not leave / cross blocks irregularly: rows a multiple of 32,
// or (F = 16) anything
template <int F> __host__ __device__ constexpr int fb_step(int rows) { return F == 16 ? rows * 17 : (rows / 32) * FbLay<F>::S32 + (rows % 32) * FbLay<F>::S1; }

template <int N> __host__ __device__ constexpr int fb_stage_stride() { return N / 4 + 32 / FbCfg<N>::F; }     // floats per staged hop row: (32 / F) f + m never collides
// the 1024-point kernel (128 registers) keeps the overlap-add carries of lanes 0..2 in
// shared memory: [N/32 butterfly rows][3 lanes][8 + 4 outputs]
template <int N> __host__ __device__ constexpr bool fb_carry_in_smem() { return N >= 1024; }
template <int N> __host__ __device__ constexpr int fb_smem_bytes() {
    constexpr int F = FbCfg<N>::F;
    return (fb_ro<F>(N) + fb_ro<F>(N / 2) + (fb_carry_in_smem<N>() ? (N / 32) * 3 * 12 : 0)) * (int)sizeof(float2);
}

// v[r] *= tw[k * R + r] (CONJ: its conjugate), r = 1..R-1: the twiddles of a second pass, eight at a time (16-byte
// loads) so that they never hold more than 16 registers beside the butterfly's 2R
template <int R, bool CONJ>
__device__ __forceinline__ void fb_apply_tw(const float2* __restrict__ tw, int k, float2 (&v)[R]) {
    const float4* __restrict__ t4 = reinterpret_cast<const float4*>(tw + k * R);
#pragma unroll
    for (int c = 0; c < R / 8; c++) {
        float4 x[4];
#pragma unroll
        for (int m = 0; m < 4; m++) x[m] = __ldg(t4 + 4 * c + m);
#pragma unroll
        for (int m = 0; m < 4; m++) {
            const int r = 8 * c + 2 * m;
            if (r > 0) v[r] = cmul(v[r], make_float2(x[m].x, CONJ ? -x[m].y : x[m].y));
            v[r + 1] = cmul(v[r + 1], make_float2(x[m].z, CONJ ? -x[m].w : x[m].w));
        }
    }
}

enum { FB_PLAIN = 0, FB_FOLD = 1, FB_MERGED = 2 };

template <int N, int MODE, bool ACCUM>
__global__ void __launch_bounds__(FbCfg<N>::F * N / 32, FbCfg<N>::CTAS) band_fb_kernel(const BandDev b, const SegArgs a) {
    constexpr int F = FbCfg<N>::F, S1 = FbLay<F>::S1;
    constexpr int T = F * N / 32, M = N / 2, H = N / 4, JS = N / 32;
    constexpr int RA = FbCfg<N>::RA, RB = FbCfg<N>::RB, HA = FbCfg<N>::HA, HB = 16;
    constexpr int NBA = N / RA, NBB = N / RB, ITA = 32 / RA, ITB = 32 / RB;     // butterflies per sequence / per thread
    constexpr int SY = RB / 4, SC = HB / 4;                                      // last-pass outputs per hop of the frame
    constexpr int HS = fb_stage_stride<N>();
    constexpr int IN_ROWS = F + 3;                                               // hops of input a tile of F frames covers
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* buf = reinterpret_cast<float2*>(smem_raw);                           // [N rows], row p at fb_ro<F>(p)
    float2* cbuf = buf + fb_ro<F>(N);                                            // [M rows] packed centre spectrum
    float* in_l = reinterpret_cast<float*>(buf);                                 // staged input [19][HS] x 2 (inside buf)
    float* in_r = in_l + IN_ROWS * HS;
    float* stage = in_r + IN_ROWS * HS;                                          // finished hops [3][F][HS] (inside buf)
    static_assert((2 * IN_ROWS * HS + 3 * F * HS) * 4 <= fb_ro<F>(N) * 8, "staging does not fit the transform buffer");
    static_assert(F == 16 || (RA == 32 && RB == 32 && HA == 32), "the 8-frame layout is laid out for 32 x 32 points");
    constexpr int RPS = T / (H / 2);                                             // staged / emitted hop rows a pass of the CTA covers

    const int tid = threadIdx.x, q = tid & (F - 1), jb = tid / F;
    const int track = blockIdx.y;
    const bool fold = MODE == FB_MERGED ? a.fold != 0 : MODE == FB_FOLD;
    const long long h0 = a.hop_begin + (long long)blockIdx.x * a.hops_per_run;
    const long long h1 = min(h0 + (long long)a.hops_per_run, a.hop_end);
    if (h0 >= h1) return;
    const float* __restrict__ gl = a.in_l + (long long)track * a.in_stride;
    const float* __restrict__ gr = a.in_r + (long long)track * a.in_stride;
    float* outp[3] = {a.out_c + (long long)track * a.out_stride, a.out_l + (long long)track * a.out_stride,
                      a.out_r + (long long)track * a.out_stride};
    const float* __restrict__ ana = b.ana;
    const float* __restrict__ syn = b.syn;
    const float* __restrict__ gain = b.gain;

    // stage the input of the tile whose first frame is F0: samples [F0*H, F0*H + 19*H) of both channels
    auto stage_input = [&](long long F0) {
        const long long s0 = F0 * H;
        const bool whole = s0 >= a.in_begin && s0 + IN_ROWS * H <= a.in_end &&
                           ((reinterpret_cast<uintptr_t>(gl + (s0 - a.in_begin)) | reinterpret_cast<uintptr_t>(gr + (s0 - a.in_begin))) & 7) == 0;
        if (whole) {
            const float* __restrict__ pl = gl + (s0 - a.in_begin);
            const float* __restrict__ pr = gr + (s0 - a.in_begin);
            // H/2 eight-byte chunks per row: a thread keeps its column and walks down RPS rows at a time
            const int col = 2 * (tid & (H / 2 - 1));
#pragma unroll
            for (int row = tid / (H / 2); row < IN_ROWS; row += RPS) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(in_l + row * HS + col)), "l"(pl + row * H + col) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(in_r + row * HS + col)), "l"(pr + row * H + col) : "memory");
            }
        } else {
            // samples outside [in_begin, in_end) -- before the track, past its end, another shard's -- are zero
            for (int c = tid; c < IN_ROWS * H; c += T) {
                const int row = c / H, col = c - row * H;
                const long long s = s0 + c;
                const bool ok = s >= a.in_begin && s < a.in_end;
                in_l[row * HS + col] = ok ? __ldg(gl + (s - a.in_begin)) : 0.f;
                in_r[row * HS + col] = ok ? __ldg(gr + (s - a.in_begin)) : 0.f;
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    // what the next tile needs from this one, per output of this thread's last-pass butterflies: lane f < 3 keeps the
    // partial sum of the older frames of hop F0 + 16 + f (see the overlap-add below)
    constexpr bool CSM = fb_carry_in_smem<N>();
    float2* carry_sm = cbuf + fb_ro<F>(M) + (jb * 3 + (q < 3 ? q : 0)) * 12;   // (CSM) this thread's 12 carries, lanes 0..2 only
    float2 carry_y[CSM ? 1 : ITB][CSM ? 1 : SY], carry_c[CSM ? 1 : SC];
    if constexpr (CSM) {
        static_assert(ITB * SY + SC == 12, "carry layout");
        if (q < 3) {
#pragma unroll
            for (int r = 0; r < 12; r++) carry_sm[r] = make_float2(0.f, 0.f);
        }
    } else {
#pragma unroll
        for (int it = 0; it < ITB; it++)
#pragma unroll
            for (int r = 0; r < SY; r++) carry_y[it][r] = make_float2(0.f, 0.f);
#pragma unroll
        for (int r = 0; r < SC; r++) carry_c[r] = make_float2(0.f, 0.f);
    }

    // hop = ((y[f-3] + y[f-2]) + y[f-1]) + y[f]: x0..x3 are this lane's contributions to the hops of frames f .. f+3
    auto overlap_add = [&](float2 x0, float2 x1, float2 x2, float2 x3, float2& carry) -> float2 {
        const unsigned FULL = 0xffffffffu;
        float2 t1, t2, t3;
        t1.x = __shfl_up_sync(FULL, x1.x, 1, F); t1.y = __shfl_up_sync(FULL, x1.y, 1, F);
        t2.x = __shfl_up_sync(FULL, x2.x, 2, F); t2.y = __shfl_up_sync(FULL, x2.y, 2, F);
        t3.x = __shfl_up_sync(FULL, x3.x, 3, F); t3.y = __shfl_up_sync(FULL, x3.y, 3, F);
        // lanes 0..2: the older frames are the previous tile's (carry: lane 0 holds (y13 + y14) + y15, lane 1 y14 + y15,
        // lane 2 y15, each the matching hop segment)
        // (selects, no branches: lane 2 adds carry + t2, lane 1 keeps the carry, lane 0 adds nothing but its own frame;
        // adding an exact zero changes nothing)
        const float2 zero = make_float2(0.f, 0.f);
        const float2 A = cadd(q >= 3 ? t3 : carry, q >= 2 ? t2 : zero);
        const float2 B = cadd(A, q >= 1 ? t1 : zero);
        const float2 sum = cadd(B, x0);
        // carry for the next tile (F = 16): lane 13 (x3 + x2@14) + x1@15, lane 14 x3 + x2@15, lane 15 x3 -> lanes 0, 1, 2
        float2 d1, d2;
        d1.x = __shfl_down_sync(FULL, x2.x, 1, F); d1.y = __shfl_down_sync(FULL, x2.y, 1, F);
        d2.x = __shfl_down_sync(FULL, x1.x, 2, F); d2.y = __shfl_down_sync(FULL, x1.y, 2, F);
        const float2 A2 = cadd(x3, q <= F - 2 ? d1 : zero);
        const float2 v = cadd(A2, q == F - 3 ? d2 : zero);
        const int src = (q < 3 ? F - 3 + q : q);
        carry.x = __shfl_sync(FULL, v.x, src, F);
        carry.y = __shfl_sync(FULL, v.y, src, F);
        return sum;
    };

    long long F0 = h0 - 3;                                       // the first tile starts with the three frames before the run
    stage_input(F0);
#pragma unroll 1
    for (; F0 < h1; F0 += F) {
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();

        // ---- forward pass 0 (radix RA): point idx of lane f is sample f*H + idx of the staged span, times ana[idx] ----
        {
            float2 v[ITA][RA];
            const bool live = F0 + q >= 0;
#pragma unroll
            for (int it = 0; it < ITA; it++) {
                const int j = jb + it * JS;
#pragma unroll
                for (int r = 0; r < RA; r++) {
                    const int row = q + (4 * r) / RA, col = j + NBA * (r % (RA / 4));
                    // (frames before the start of the track do not exist, center_extraction.py:448: they add nothing)
                    const float wn = live ? __ldg(ana + j + NBA * r) : 0.f;
                    v[it][r] = cscale(make_float2(in_l[row * HS + col], in_r[row * HS + col]), wn);
                }
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < ITA; it++) {
                const int j = jb + it * JS;
                Dft<RA, -1>::run(v[it]);
                float2* __restrict__ dst = buf + fb_ro<F>(j * RA) + q;
#pragma unroll
                for (int r = 0; r < RA; r++) dst[r * S1] = v[it][r];
            }
        }
        __syncthreads();
        // ---- forward pass 1 (radix RB, NS = RA): butterfly j reads j + r*NBB, twiddle exp(-2 pi i r j / N), writes bin j + r*RA ----
        {
            float2 v[ITB][RB];
#pragma unroll
            for (int it = 0; it < ITB; it++) {
                const float2* __restrict__ src = buf + fb_ro<F>(jb + it * JS) + q;
#pragma unroll
                for (int r = 0; r < RB; r++) v[it][r] = src[r * fb_step<F>(NBB)];
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < ITB; it++) {
                const int j = jb + it * JS;
                fb_apply_tw<RB, false>(b.fb.tw_full, j, v[it]);
                Dft<RB, -1>::run(v[it]);
                float2* __restrict__ dst = buf + fb_ro<F>(j) + q;
#pragma unroll
                for (int r = 0; r < RB; r++) dst[r * fb_step<F>(RA)] = v[it][r];
            }
        }
        __syncthreads();

        // ---- split / gain / centre mask on bin quadruples k, M-k, M+k, N-k (k = 0 .. M/2), in place; packed centre -> cbuf ----
        {
            auto mask_item = [&](int k) {
                const int k2 = M - k, km = (N - k) & (N - 1);
                float2* __restrict__ p1 = buf + fb_ro<F>(k) + q;
                float2* __restrict__ p1m = buf + fb_ro<F>(km) + q;
                float2* __restrict__ p2 = buf + fb_ro<F>(k2) + q;
                float2* __restrict__ p2m = buf + fb_ro<F>(M + k) + q;
                const float g1 = __ldg(gain + k), g2 = __ldg(gain + k2);
                const float2 wp = __ldg(b.tw_pack + k);
                // (no shortcut for zero gains: this kernel serves dense bands; a zero gain gives exact zeros anyway, and
                // straight-line items let the gain / twiddle loads of several items overlap)
                const float2 a1 = *p1, b1 = *p1m, a2 = *p2, b2 = *p2m;
                float2 c1, y1, y1m, c2, y2, y2m;
                if constexpr (MODE == FB_MERGED) {
                    mask_bin_merged(a1, b1, g1, gain + k, b.n_gains, b.gain_stride, y1, y1m, c1);
                    mask_bin_merged(a2, b2, g2, gain + k2, b.n_gains, b.gain_stride, y2, y2m, c2);
                } else {
                    mask_bin(a1, b1, g1, y1, y1m, c1);
                    mask_bin(a2, b2, g2, y2, y2m, c2);
                }
                if (fold) {                                       // (Ls + C/2) + i (Rs + C/2): add (1+i) C / 2
                    const float2 u1 = cadd(make_float2(c1.x, c1.x), make_float2(-c1.y, c1.y));
                    const float2 u2 = cadd(make_float2(c2.x, c2.x), make_float2(-c2.y, c2.y));
                    y1 = caxpy(u1, 0.5f, y1);
                    y1m = caxpy(make_float2(u1.y, u1.x), 0.5f, y1m);
                    y2 = caxpy(u2, 0.5f, y2);
                    y2m = caxpy(make_float2(u2.y, u2.x), 0.5f, y2m);
                }
                *p1 = y1; *p1m = y1m; *p2 = y2; *p2m = y2m;
                if (!fold) {
                    float2 zk, zmk;
                    pack_pair(c1, c2, wp, zk, zmk);
                    cbuf[fb_ro<F>(k) + q] = zk;
                    if (k > 0) cbuf[fb_ro<F>(k2) + q] = zmk;
                }
            };
#pragma unroll 4
            for (int i = 0; i < (M / 2) / JS; i++) mask_item(jb + i * JS);
            if (jb == 0) mask_item(M / 2);                        // bins M/2 and N - M/2: one pair, taken twice by the item code
        }
        __syncthreads();

        // ---- inverse of Ls + i Rs: pass 0 (radix RA, no twiddles), in place ----
        {
            float2 v[ITA][RA];
#pragma unroll
            for (int it = 0; it < ITA; it++) {
                const float2* __restrict__ src = buf + fb_ro<F>(jb + it * JS) + q;
#pragma unroll
                for (int r = 0; r < RA; r++) v[it][r] = src[r * fb_step<F>(NBA)];
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < ITA; it++) {
                Dft<RA, +1>::run(v[it]);
                float2* __restrict__ dst = buf + fb_ro<F>((jb + it * JS) * RA) + q;
#pragma unroll
                for (int r = 0; r < RA; r++) dst[r * S1] = v[it][r];
            }
        }
        __syncthreads();
        // ---- pass 1 (radix RB): output r of butterfly j is sample j + r*RA of the frame, in hop r / SY; synthesis window,
        // overlap-add across lanes, finished hops -> stage ----
        {
            float2 v[ITB][RB];
#pragma unroll
            for (int it = 0; it < ITB; it++) {
                const float2* __restrict__ src = buf + fb_ro<F>(jb + it * JS) + q;
#pragma unroll
                for (int r = 0; r < RB; r++) v[it][r] = src[r * fb_step<F>(NBB)];
            }
            __syncthreads();                                     // buf is free: the stage and the next tile's input live in it
            if (F0 + F < h1) stage_input(F0 + F);
#pragma unroll
            for (int it = 0; it < ITB; it++) {
                const int j = jb + it * JS;
                fb_apply_tw<RB, true>(b.fb.tw_full, j, v[it]);
                Dft<RB, +1>::run(v[it]);
#pragma unroll
                for (int r = 0; r < RB; r++) v[it][r] = cscale(v[it][r], __ldg(syn + j + r * RA));
#pragma unroll
                for (int rr = 0; rr < SY; rr++) {
                    float2 s;
                    if constexpr (CSM) {
                        float2 cy = q < 3 ? carry_sm[it * SY + rr] : make_float2(0.f, 0.f);
                        s = overlap_add(v[it][rr], v[it][SY + rr], v[it][2 * SY + rr], v[it][3 * SY + rr], cy);
                        if (q < 3) carry_sm[it * SY + rr] = cy;
                    } else {
                        s = overlap_add(v[it][rr], v[it][SY + rr], v[it][2 * SY + rr], v[it][3 * SY + rr], carry_y[it][rr]);
                    }
                    const int m = j + rr * RA;
                    stage[(F + q) * HS + m] = s.x;
                    stage[(2 * F + q) * HS + m] = s.y;
                }
            }
        }
        // ---- inverse of the packed centre (N/2 points: HA x 16): c[2m] + i c[2m+1] ----
        if (!fold) {
            constexpr int NBH = M / HA, ITH = (16 / HA) > 0 ? 16 / HA : 1;
            const bool act = HA <= 16 || tid < T / 2;             // radix 32: half the threads have a butterfly
            {
                float2 v[ITH][HA];
                if (act) {
#pragma unroll
                    for (int it = 0; it < ITH; it++) {
                        // (F = 8: NBH = 16 and jb < 16, so row jb + 16 r is row jb + 16 (r & 1) of block r >> 1)
                        const float2* __restrict__ src = cbuf + fb_ro<F>(jb + it * JS) + q;
#pragma unroll
                        for (int r = 0; r < HA; r++) v[it][r] = src[F == 16 ? r * NBH * 17 : fb_ro<F>(r * NBH)];
                    }
                }
                __syncthreads();
                if (act) {
#pragma unroll
                    for (int it = 0; it < ITH; it++) {
                        Dft<HA, +1>::run(v[it]);
                        float2* __restrict__ dst = cbuf + fb_ro<F>((jb + it * JS) * HA) + q;
#pragma unroll
                        for (int r = 0; r < HA; r++) dst[r * S1] = v[it][r];
                    }
                }
            }
            __syncthreads();
            {
                constexpr int NBL = M / HB;                       // = HA: butterflies per sequence of the last pass
                float2 v[HB];
                const int j = jb;                                 // T / 16 = N / 32 = M / 16 butterflies per sequence: one each
                const float2* __restrict__ src = cbuf + fb_ro<F>(j) + q;
#pragma unroll
                for (int r = 0; r < HB; r++) v[r] = src[F == 16 ? r * NBL * 17 : fb_ro<F>(r * NBL)];
                fb_apply_tw<HB, true>(b.fb.tw_half, j, v);
                Dft<HB, +1>::run(v);
#pragma unroll
                for (int r = 0; r < HB; r++) {
                    const float2 wn = __ldg(reinterpret_cast<const float2*>(syn) + j + r * HA);
                    v[r] = __fmul2_rn(v[r], wn);
                }
#pragma unroll
                for (int rr = 0; rr < SC; rr++) {
                    float2 s;
                    if constexpr (CSM) {
                        float2 cy = q < 3 ? carry_sm[ITB * SY + rr] : make_float2(0.f, 0.f);
                        s = overlap_add(v[rr], v[SC + rr], v[2 * SC + rr], v[3 * SC + rr], cy);
                        if (q < 3) carry_sm[ITB * SY + rr] = cy;
                    } else {
                        s = overlap_add(v[rr], v[SC + rr], v[2 * SC + rr], v[3 * SC + rr], carry_c[rr]);
                    }
                    *reinterpret_cast<float2*>(stage + q * HS + 2 * (j + rr * HA)) = s;
                }
            }
        }
        __syncthreads();

        // ---- emit: the tile's 16 hops are 16*H consecutive samples of each output ----
        {
            const long long sb = F0 * H;                          // first sample of the tile's first hop
            // samples [e_lo, e_hi) of the tile's span go out: hops of this run only (the first tile's first three hops
            // are incomplete), inside the segment
            const long long lo = max(max(h0 * H, a.seg_begin), sb), hi = min(min(h1 * H, a.seg_end), sb + (long long)F * H);
            const int e_lo = (int)(lo - sb), e_hi = (int)(hi - sb);
            if (e_hi > e_lo) {
#pragma unroll
                for (int ch = 0; ch < 3; ch++) {
                    if (ch == 0 && fold) continue;
                    float* __restrict__ po = outp[ch] + (sb - a.out_begin);
                    const float* __restrict__ sg = stage + ch * F * HS;
                    const bool vec = (reinterpret_cast<uintptr_t>(po) & 7) == 0;      // CTA-uniform
                    if (vec && e_lo == 0 && e_hi == F * H) {                           // the usual tile: every hop goes out whole
                        float2 pv[(F * H / 2) / T];
                        if (ACCUM) {
#pragma unroll
                            for (int i = 0; i < (F * H / 2) / T; i++) pv[i] = __ldcs(reinterpret_cast<const float2*>(po) + tid + i * T);
                        }
                        // H/2 float2 per hop row: a thread keeps its column, RPS hops further each time
                        const float* __restrict__ sg0 = sg + (tid / (H / 2)) * HS + 2 * (tid & (H / 2 - 1));
#pragma unroll
                        for (int i = 0; i < (F * H / 2) / T; i++) {
                            float2 s = *reinterpret_cast<const float2*>(sg0 + RPS * i * HS);
                            if (ACCUM) s = make_float2(pv[i].x + s.x, pv[i].y + s.y);
                            __stcs(reinterpret_cast<float2*>(po) + tid + i * T, s);
                        }
                    } else if (vec) {
                        float2 pv[(F * H / 2 + T - 1) / T];
                        if (ACCUM) {
#pragma unroll
                            for (int i = 0; i < (F * H / 2) / T; i++) {
                                const int e = 2 * (tid + i * T);
                                if (e >= e_lo && e + 1 < e_hi) pv[i] = __ldcs(reinterpret_cast<const float2*>(po + e));
                            }
                        }
#pragma unroll
                        for (int i = 0; i < (F * H / 2) / T; i++) {
                            const int e = 2 * (tid + i * T);
                            const int f = e / H, m = e - f * H;
                            float2 s = *reinterpret_cast<const float2*>(sg + f * HS + m);
                            if (e >= e_lo && e + 1 < e_hi) {
                                if (ACCUM) s = make_float2(pv[i].x + s.x, pv[i].y + s.y);
                                __stcs(reinterpret_cast<float2*>(po + e), s);
                            } else {
                                if (e >= e_lo && e < e_hi) po[e] = ACCUM ? po[e] + s.x : s.x;
                                if (e + 1 >= e_lo && e + 1 < e_hi) po[e + 1] = ACCUM ? po[e + 1] + s.y : s.y;
                            }
                        }
                    } else {
                        for (int e = tid; e < F * H; e += T) {
                            const int f = e / H, m = e - f * H;
                            const float s = sg[f * HS + m];
                            if (e >= e_lo && e < e_hi) po[e] = ACCUM ? po[e] + s : s;
                        }
                    }
                }
            }
        }
        // (the next iteration starts with a barrier: the stage is read before anything overwrites buf)
    }
}

template <int N, int MODE>
static cudaError_t launch_fb_nm(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st) {
    cudaError_t e;
    if (a.accum) {
        if ((e = cudaFuncSetAttribute(band_fb_kernel<N, MODE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, fb_smem_bytes<N>())) != cudaSuccess) return e;
        band_fb_kernel<N, MODE, true><<<dim3(n_runs, n_tracks), FbCfg<N>::F * N / 32, fb_smem_bytes<N>(), st>>>(b, a);
    } else {
        if ((e = cudaFuncSetAttribute(band_fb_kernel<N, MODE, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, fb_smem_bytes<N>())) != cudaSuccess) return e;
        band_fb_kernel<N, MODE, false><<<dim3(n_runs, n_tracks), FbCfg<N>::F * N / 32, fb_smem_bytes<N>(), st>>>(b, a);
    }
    return cudaGetLastError();
}
template <int N>
static cudaError_t launch_fb_n(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st) {
    if (b.n_gains > 1) return launch_fb_nm<N, FB_MERGED>(b, a, n_runs, n_tracks, st);
    return a.fold ? launch_fb_nm<N, FB_FOLD>(b, a, n_runs, n_tracks, st) : launch_fb_nm<N, FB_PLAIN>(b, a, n_runs, n_tracks, st);
}

}  // namespace upmix

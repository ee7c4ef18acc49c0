// Decimated ("band-limited") band kernels for sm_100a.
//
// Every band but the top one of a crossover set keeps only a sliver of its spectrum: the dynamic-resolution
// rule (center_extraction.py:173-197) puts bin_low near 32..64, so the pass band plus its raised-cosine
// fades (center_extraction.py:282-332) ends below bin ~430 whatever the STFT size -- 0.2 .. 12 % of the
// n_fft/2+1 bins the reference transforms, masks and inverts per frame (center_extraction.py:366-389).
// With K = the highest bin that carries gain, P the power of two above it and n_fft = P*Q, write the
// sample index as n = Q p + q.  Then, for the packed frame z = ana * (L + iR),
//
//     Z[k]       = sum_q  W^{ qk}  F_q[k]          F_q = FFT_P( z[Q p + q] over p ),   W = exp(-2 pi i / n_fft)
//     Z[n_fft-k] = sum_q  W^{-qk}  F_q[P-k]                                            0 <= k <= K
//
// i.e. Q independent P-point transforms (one per decimated sequence) and a twiddled sum over q for the few
// live bins; and the other way round the masked spectrum Y (live bins only) gives every output sample as
//
//     y[Q p + q] = IFFT_P( U_q )[p],   U_q[s] = [s <= K] Y[s] W^{-qs} + [P-s <= K] Y[n_fft-(P-s)] W^{q(P-s)}
//
// -- again Q independent P-point transforms.  Work per frame drops from n_fft log n_fft to n_fft log P plus
// O(K Q), a 65536-point frame never has to exist as a whole (no four-step scratch in HBM: what travels
// between the kernels is the K+1 live bins), and because the sequences are the batch dimension every
// shared-memory access of the transforms is lane-contiguous and every pass twiddle is the same for the 16
// lanes of a half-warp.
//
//   dec_fwd_kernel   CTA = (16 sequences q, one frame): window, P-point transforms, partial twiddled sums
//                    over its 16 q (Horner in W^k); with Q = 16 it also masks (split, gain, centre factor)
//   dec_mask_kernel  Q > 16: adds the Q/16 partial sums in group order and masks
//   dec_inv_kernel   CTA = (16 columns, run of hops): expands the live bins to U_q, P-point inverse, synthesis
//                    window, overlap-add in REGISTERS (the last pass is radix 16 and hop = P/4 points of a
//                    sequence, so output r of a butterfly lies in hop r/4: 12 accumulators per butterfly), emits
//                    the finished hop.  Ls + i Rs and the centre are separate launches of the same kernel: the
//                    centre is real, so two sequences share one complex transform (q and q+16, or -- when there
//                    are only 16 sequences -- q and q+8 of two runs of hops side by side).
//
// Reference behaviour reproduced: center_extraction.py:353-409 per frame, 426-472 over the signal.
#pragma once
#include <stdlib.h>

#include <algorithm>

#include "fft_device.cuh"
#include "upmix_kernels.cuh"
#include "upmix_launch.h"

namespace upmix {

constexpr int DEC_QS = 17;        // row stride (float2) of a 16-column tile: odd, so that column-wise walks
                                  // (lanes = consecutive rows) are as conflict-free as row-wise ones
#ifndef UPMIX_DEC_PREV_AHEAD
#define UPMIX_DEC_PREV_AHEAD 1    // accumulating bands: the previous sums of a hop are requested one frame ahead (two buffers)
#endif
constexpr int DEC_PREV_BUFS = UPMIX_DEC_PREV_AHEAD ? 2 : 1;

// second-pass twiddles of this thread's butterfly: w[r] = exp(-2 pi i r k / P), r = 1..15, as eight 16-byte loads
__device__ __forceinline__ void load_tw16(const float2* __restrict__ tw_last, int k, float2 (&w)[16]) {
    const float4* __restrict__ t4 = reinterpret_cast<const float4*>(tw_last + k * 16);
#pragma unroll
    for (int m = 0; m < 8; m++) {
        const float4 x = __ldg(t4 + m);
        w[2 * m] = make_float2(x.x, x.y);
        w[2 * m + 1] = make_float2(x.z, x.w);
    }
}

// ---------------------------------------------------------------------------------------------
// forward: frame -> live bins
// ---------------------------------------------------------------------------------------------
template <int P, int Q>
__global__ void __launch_bounds__(P / 2, 1024 / P) dec_fwd_kernel(const BandDev b, const SegArgs a, const DecWave w) {
    constexpr int T = P / 2, R0 = P / 16, NB0 = 16, NB1 = R0, IT0 = 32 / R0, JS = P / 32, N = P * Q;
    constexpr bool FUSE_MASK = Q == 16;                          // one group of 16 sequences: the CTA has every live bin
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* buf = reinterpret_cast<float2*>(smem_raw);           // [P][DEC_QS]
    float2* zs = buf + P * DEC_QS;                               // [2][KP] (FUSE_MASK)
    const int tid = threadIdx.x, q = tid & 15, jb = tid >> 4;
    const int fl = blockIdx.x, g = blockIdx.y, track = blockIdx.z;      // frames on x: an hour of 8192-point frames exceeds 65535
    const int K = b.dec.K, KP = b.dec.KP;
    const int qg = g * 16 + q;
    const long long s0 = (w.frame0 + fl) * (long long)b.hop;
    const float* __restrict__ inl = a.in_l + (long long)track * a.in_stride;
    const float* __restrict__ inr = a.in_r + (long long)track * a.in_stride;
    const float* __restrict__ ana = b.ana;

    if (b.frame_step > 1 && (w.frame0 + fl) % b.frame_step != 0) {
        // 50 % overlap run on the 75 % machinery: this frame does not exist -- its live bins are zero
        const float2 zero = make_float2(0.f, 0.f);
        if constexpr (FUSE_MASK) {
            float2* __restrict__ sp = w.spec + (((long long)track * w.n_frames + fl) * 3) * KP;
            for (int k = tid; k <= K; k += T) sp[k] = sp[KP + k] = sp[2 * KP + k] = zero;
        } else {
            float2* __restrict__ pz = w.part + (((long long)track * w.n_frames + fl) * (Q / 16) + g) * 2 * KP;
            for (int k = tid; k <= K; k += T) pz[k] = pz[KP + k] = zero;
        }
        return;
    }

    // pass 0 (radix R0, no twiddles): input point idx of sequence q is sample Q*idx + q of the frame
    {
        float2 v[IT0][R0];
        const bool whole = s0 >= a.in_begin && s0 + N <= a.in_end;                 // CTA-uniform
        if (whole) {
            const float* __restrict__ pl = inl + (s0 - a.in_begin) + qg;
            const float* __restrict__ pr = inr + (s0 - a.in_begin) + qg;
#pragma unroll
            for (int it = 0; it < IT0; it++) {
                const int j = jb + it * JS;
#pragma unroll
                for (int r = 0; r < R0; r++) {
                    const int n = Q * (j + r * NB0);
                    v[it][r] = cscale(make_float2(__ldg(pl + n), __ldg(pr + n)), __ldg(ana + qg + n));
                }
            }
        } else {
            // samples outside [in_begin, in_end) -- before the track, past its end, another shard's -- are zero
#pragma unroll
            for (int it = 0; it < IT0; it++) {
                const int j = jb + it * JS;
#pragma unroll
                for (int r = 0; r < R0; r++) {
                    const int n = Q * (j + r * NB0) + qg;
                    const long long s = s0 + n;
                    const bool ok = s >= a.in_begin && s < a.in_end;
                    const long long idx = ok ? s - a.in_begin : 0;
                    const float wn = __ldg(ana + n);
                    const float l = __ldg(inl + idx), rr = __ldg(inr + idx);
                    v[it][r] = ok ? make_float2(l * wn, rr * wn) : make_float2(0.f, 0.f);
                }
            }
        }
#pragma unroll
        for (int it = 0; it < IT0; it++) {
            const int j = jb + it * JS;
            Dft<R0, -1>::run(v[it]);
            float2* __restrict__ dst = buf + (j * R0) * DEC_QS + q;
#pragma unroll
            for (int r = 0; r < R0; r++) dst[r * DEC_QS] = v[it][r];
        }
    }
    __syncthreads();
    // pass 1 (radix 16, NS = R0): butterfly j reads j + r*R0, twiddle exp(-2 pi i r j / P), writes slot j + r*R0
    {
        float2 u[2][16];
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const int j = jb + it * JS;
            const float2* __restrict__ src = buf + j * DEC_QS + q;
#pragma unroll
            for (int r = 0; r < 16; r++) u[it][r] = src[r * NB1 * DEC_QS];
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const int j = jb + it * JS;
            float2 tw[16];
            load_tw16(b.dec.tw_last, j, tw);
#pragma unroll
            for (int r = 1; r < 16; r++) u[it][r] = cmul(u[it][r], tw[r]);
            Dft<16, -1>::run(u[it]);
            float2* __restrict__ dst = buf + j * DEC_QS + q;
#pragma unroll
            for (int r = 0; r < 16; r++) dst[r * R0 * DEC_QS] = u[it][r];
        }
    }
    __syncthreads();
    // twiddled sum over this group's 16 sequences, one item per live bin and sign:
    //   sum_j F_{16g+j}[slot] W^{+-k(16g+j)} = W^{+-16gk} * Horner_j( F_j, W^{+-k} )
    float2* __restrict__ part = FUSE_MASK ? zs : w.part + (((long long)track * w.n_frames + fl) * (Q / 16) + g) * 2 * KP;
    for (int i = tid; i < 2 * K + 1; i += T) {
        const bool minus = i > K;
        const int k = minus ? i - K : i;
        const int slot = minus ? P - k : k;
        float2 st = __ldg(b.dec.tw_step + k), bs = __ldg(b.dec.tw_base + g * KP + k);
        if (minus) { st.y = -st.y; bs.y = -bs.y; }
        const float2* __restrict__ row = buf + slot * DEC_QS;
        float2 acc = row[15];
#pragma unroll
        for (int j = 14; j >= 0; j--) acc = cfma(acc, st, row[j]);
        part[(minus ? KP : 0) + k] = cmul(acc, bs);
    }
    if constexpr (FUSE_MASK) {
        __syncthreads();
        float2* __restrict__ sp = w.spec + (((long long)track * w.n_frames + fl) * 3) * KP;
        for (int k = tid; k <= K; k += T) {
            const float2 za = zs[k], zb = k ? zs[KP + k] : za;
            float2 ylo, yhi, c;
            mask_bin_merged(za, zb, __ldg(b.gain + k), b.gain + k, b.n_gains, b.gain_stride, ylo, yhi, c);
            if (a.fold) {                                        // (Ls + C/2) + i (Rs + C/2): add (1+i) C / 2
                const float2 uu = cadd(make_float2(c.x, c.x), make_float2(-c.y, c.y));
                ylo = caxpy(uu, 0.5f, ylo);
                yhi = caxpy(make_float2(uu.y, uu.x), 0.5f, yhi);
            }
            sp[k] = ylo;
            sp[KP + k] = yhi;
            sp[2 * KP + k] = c;
        }
    }
}

// Q > 16: partial sums of the Q/16 groups, added in group order, then the mask (upmix_dec.cu)
cudaError_t launch_dec_mask(const BandDev& b, const DecWave& w, int fold, int n_tracks, cudaStream_t st);

// ---------------------------------------------------------------------------------------------
// inverse: live bins -> finished hops
// ---------------------------------------------------------------------------------------------
// CM: what the 16 columns of the tile are
//   DEC_Y    sequences q = 16 g + col of Ls + i Rs            (blockIdx.y = g < Q/16)
//   DEC_C32  centre, sequences q = 32 g + col (real part) and q + 16 (imaginary part)   (blockIdx.y = g < Q/32)
//   DEC_C16  centre when Q = 16: columns 0..7 = sequences col / col + 8 of run 2*blockIdx.x, columns 8..15 the
//            same of run 2*blockIdx.x + 1 (two runs of hops side by side)
enum { DEC_Y = 0, DEC_C32 = 1, DEC_C16 = 2 };

template <int P, int Q, bool CENTRE, bool ACCUM>
__global__ void __launch_bounds__(P / 2, 1024 / P) dec_inv_kernel(const BandDev b, const SegArgs a, const DecWave w) {
    constexpr int R0 = P / 16, NB0 = 16, NB1 = R0, IT0 = 32 / R0, JS = P / 32, H = P * Q / 4;
    constexpr int CM = !CENTRE ? DEC_Y : Q == 16 ? DEC_C16 : DEC_C32;
    constexpr int NST = CM == DEC_C16 ? 2 : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float2* buf = reinterpret_cast<float2*>(smem_raw);           // [P][DEC_QS]
    float* prevbuf = reinterpret_cast<float*>(buf + P * DEC_QS);  // ACCUM: [DEC_PREV_BUFS][2 regions][P/4 rows][16] previous sums of the hop
    constexpr int RSZ = (P / 4) * 16;
    const int tid = threadIdx.x, q = tid & 15, jb = tid >> 4;
    const int g = blockIdx.y, track = blockIdx.z;
    const int K = b.dec.K, KP = b.dec.KP;

    // runs of hops [h0, h1) of the stream(s) of this tile
    long long h0s[NST], h1s[NST];
    bool any_run = false;
#pragma unroll
    for (int s = 0; s < NST; s++) {
        const long long run = (long long)blockIdx.x * NST + s;
        h0s[s] = a.hop_begin + run * a.hops_per_run;
        h1s[s] = min(h0s[s] + (long long)a.hops_per_run, a.hop_end);
        any_run = any_run || h0s[s] < h1s[s];
    }
    if (!any_run) return;

    // emission side: what this lane's column holds
    const int sid = CM == DEC_C16 ? q >> 3 : 0;
    const int qoff1 = CM == DEC_Y ? 16 * g + q : CM == DEC_C32 ? 32 * g + q : (q & 7);
    const int qoff2 = CM == DEC_Y ? qoff1 : CM == DEC_C32 ? qoff1 + 16 : qoff1 + 8;
    float* __restrict__ o1 = (CM == DEC_Y ? a.out_l : a.out_c) + (long long)track * a.out_stride;
    float* __restrict__ o2 = (CM == DEC_Y ? a.out_r : a.out_c) + (long long)track * a.out_stride;
    const float* __restrict__ syn = b.syn;
    const long long my_h0 = NST == 2 ? (sid ? h0s[NST - 1] : h0s[0]) : h0s[0];
    const long long my_h1 = NST == 2 ? (sid ? h1s[NST - 1] : h1s[0]) : h1s[0];

    // expansion side: this thread's slot pair (s, P - s); thread 0 takes P/2 (its own partner) and slot 0
    const int es = tid == 0 ? P / 2 : tid, es2 = P - es;
    const bool has1 = es <= K, has2 = es2 <= K;
    const float2 one = make_float2(1.f, 0.f);
    const float2 st1 = has1 ? __ldg(b.dec.tw_step + es) : one, st2 = has2 ? __ldg(b.dec.tw_step + es2) : one;
    // twiddle of the first sequence of the tile's 16-sequence chunk(s): W^{16 g' s}
    const int gb = CM == DEC_Y ? g : CM == DEC_C32 ? 2 * g : 0;
    const float2 bs1 = has1 ? __ldg(b.dec.tw_base + gb * KP + es) : one, bs2 = has2 ? __ldg(b.dec.tw_base + gb * KP + es2) : one;
    float2 bs1b = one, bs2b = one;                                // second chunk (q + 16) of a DEC_C32 tile
    if (CM == DEC_C32) {
        if (has1) bs1b = __ldg(b.dec.tw_base + (gb + 1) * KP + es);
        if (has2) bs2b = __ldg(b.dec.tw_base + (gb + 1) * KP + es2);
    }

    float2 acc[2][12];                                           // overlap-add: the three unfinished hops of this thread's outputs
#pragma unroll
    for (int it = 0; it < 2; it++)
#pragma unroll
        for (int r = 0; r < 12; r++) acc[it][r] = make_float2(0.f, 0.f);

    const int n_iter = a.hops_per_run + 3;
    // ACCUM (this band adds to what the outputs hold: the bands before it, in band order): the hop a frame finishes is
    // copied asynchronously (cp.async, no registers) from the outputs into shared memory and read there by the last pass
    // -- loaded at the store, the dependent HBM round trips cost 25 % of the kernel.  Two regions of P/4 rows x 16 floats:
    // Ls / Rs rows of the tile (DEC_Y), the two 16-sequence halves (DEC_C32), the two runs (DEC_C16).  Hops cut by a
    // segment edge, or not 16-byte aligned, take the plain loads below.  A frame lasts about as long as a trip to HBM,
    // so the request goes out ONE FRAME AHEAD (two buffers, frame i uses buffer i & 1): issued at the top of frame i - 1,
    // waited for before the last pass of frame i.
    auto request_prev = [&](int i) -> bool {
        const float* gsrc[2];
        bool ok = i < n_iter;
#pragma unroll
        for (int rg = 0; rg < 2; rg++) {
            const int s = NST == 2 ? rg : 0;
            const long long f = h0s[s] - 3 + i;
            const bool valid_s = h0s[s] < h1s[s] && f >= 0 && f < h1s[s];
            const long long sb = f * H;
            ok = ok && valid_s && f >= h0s[s] && sb >= a.seg_begin && sb + H <= a.seg_end;
            const float* base = (CM == DEC_Y ? (rg ? a.out_r : a.out_l) : a.out_c) + (long long)track * a.out_stride + (sb - a.out_begin);
            gsrc[rg] = base + (CM == DEC_Y ? 16 * g : CM == DEC_C32 ? 32 * g + 16 * rg : 0);
            ok = ok && (reinterpret_cast<uintptr_t>(gsrc[rg]) & 15) == 0;
        }
        if (ok) {
            float* pb = prevbuf + (DEC_PREV_BUFS == 2 ? (i & 1) * 2 * RSZ : 0);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int c = tid + k * (P / 2);                 // 2 regions x P chunks of 16 bytes
                const int rg = c / P, row = (c % P) >> 2, c4 = c & 3;
                const float* src = gsrc[rg] + Q * row + 4 * c4;
                const uint32_t dst = smem_u32(pb + rg * RSZ + row * 16 + 4 * c4);
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");     // (an empty group when nothing was requested)
        return ok;
    };
    bool staged_next = false;
    if constexpr (ACCUM && DEC_PREV_BUFS == 2) staged_next = request_prev(0);
#pragma unroll 1
    for (int i = 0; i < n_iter; i++) {
        long long fs[NST];
        bool valid[NST], any = false;
#pragma unroll
        for (int s = 0; s < NST; s++) {
            fs[s] = h0s[s] - 3 + i;
            valid[s] = h0s[s] < h1s[s] && fs[s] >= 0 && fs[s] < h1s[s];
            any = any || valid[s];
        }
        bool staged_prev = false;
        if constexpr (ACCUM) {
            if constexpr (DEC_PREV_BUFS == 2) {
                staged_prev = staged_next;
                staged_next = request_prev(i + 1);
            } else {
                if (any) staged_prev = request_prev(i);
            }
        }
        if (!any) continue;                                      // CTA-uniform: nothing has been accumulated yet
        const float* __restrict__ prevcur = prevbuf + (DEC_PREV_BUFS == 2 ? (i & 1) * 2 * RSZ : 0);

        // ---- expansion: live bins -> U_q[s] for the tile's columns, written column-wise (lanes = slots) ----
        if constexpr (CM == DEC_Y) {
            const float2* __restrict__ sp = w.spec + (((long long)track * w.n_frames + (fs[0] - w.frame0)) * 3) * KP;
            const float2 zero = make_float2(0.f, 0.f);
            const float2 yp1 = has1 ? sp[es] : zero, ym1 = has1 ? sp[KP + es] : zero;
            float2* __restrict__ row1 = buf + es * DEC_QS;
            float2* __restrict__ row2 = buf + es2 * DEC_QS;
            float2 ca = cmul(yp1, cconj(bs1)), cb = cmul(ym1, bs1);
            if (has2) {
                const float2 yp2 = sp[es2], ym2 = sp[KP + es2];
                float2 cc = cmul(ym2, bs2), cd = cmul(yp2, cconj(bs2));
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    row1[c] = cadd(ca, cc);
                    if (es2 != es) row2[c] = cadd(cd, cb);
                    ca = cmul(ca, cconj(st1));
                    cb = cmul(cb, st1);
                    cc = cmul(cc, st2);
                    cd = cmul(cd, cconj(st2));
                }
            } else {
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    row1[c] = ca;
                    row2[c] = cb;
                    ca = cmul(ca, cconj(st1));
                    cb = cmul(cb, st1);
                }
            }
            if (tid == 0) {
                const float2 y0 = sp[0];
#pragma unroll
                for (int c = 0; c < 16; c++) buf[c] = y0;
            }
        } else {
            // centre: Uc_q[s] = C[s] W^{-qs} + conj(C[P-s]) W^{q(P-s)}, Hermitian in s; column = Uc_q + i Uc_q'
            float2* __restrict__ row1 = buf + es * DEC_QS;
            float2* __restrict__ row2 = buf + es2 * DEC_QS;
            const float2 zero = make_float2(0.f, 0.f);
            if constexpr (CM == DEC_C32) {
                const float2* __restrict__ sp = w.spec + (((long long)track * w.n_frames + (fs[0] - w.frame0)) * 3 + 2) * KP;
                const float2 c1 = has1 ? sp[es] : zero, c2 = has2 ? cconj(sp[es2]) : zero;
                float2 e = cmul(c1, cconj(bs1)), f = cmul(c2, bs2), eb = cmul(c1, cconj(bs1b)), fb = cmul(c2, bs2b);
#pragma unroll
                for (int c = 0; c < 16; c++) {
                    const float2 lo = cadd(e, f), hi = cadd(eb, fb);
                    row1[c] = make_float2(lo.x - hi.y, lo.y + hi.x);
                    if (es2 != es) row2[c] = make_float2(lo.x + hi.y, hi.x - lo.y);
                    e = cmul(e, cconj(st1));
                    eb = cmul(eb, cconj(st1));
                    f = cmul(f, st2);
                    fb = cmul(fb, st2);
                }
                if (tid == 0) {
                    const float2 c0 = sp[0];
#pragma unroll
                    for (int c = 0; c < 16; c++) buf[c] = make_float2(c0.x - c0.y, c0.y + c0.x);
                }
            } else {
#pragma unroll
                for (int s = 0; s < NST; s++) {
                    const float2* __restrict__ sp = w.spec + (((long long)track * w.n_frames + (fs[s] - w.frame0)) * 3 + 2) * KP;
                    const float2 c1 = (valid[s] && has1) ? sp[es] : zero, c2 = (valid[s] && has2) ? cconj(sp[es2]) : zero;
                    float2 e = c1, f = c2;                       // Q = 16: one chunk, first twiddle is 1
                    float2 u[16];
#pragma unroll
                    for (int c = 0; c < 16; c++) {
                        u[c] = cadd(e, f);
                        e = cmul(e, cconj(st1));
                        f = cmul(f, st2);
                    }
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const float2 lo = u[c], hi = u[c + 8];
                        row1[8 * s + c] = make_float2(lo.x - hi.y, lo.y + hi.x);
                        if (es2 != es) row2[8 * s + c] = make_float2(lo.x + hi.y, hi.x - lo.y);
                    }
                    if (tid == 0) {
                        const float2 c0 = valid[s] ? sp[0] : zero;
#pragma unroll
                        for (int c = 0; c < 8; c++) buf[8 * s + c] = make_float2(c0.x - c0.y, c0.y + c0.x);
                    }
                }
            }
        }
        __syncthreads();

        // the live bins this thread expands for the NEXT frame: pull them from L2 into L1 while the passes run
        // (the expansion's first multiply waited on these loads for 14 % of the kernel's stall samples)
        if (i + 1 < n_iter) {
#pragma unroll
            for (int s = 0; s < NST; s++) {
                const long long fn = fs[s] + 1;
                if (fn >= 0 && fn < h1s[s]) {
                    const float2* __restrict__ sp = w.spec + (((long long)track * w.n_frames + (fn - w.frame0)) * 3 + (CM == DEC_Y ? 0 : 2)) * KP;
                    if (has1) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + es));
                        if (CM == DEC_Y) asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + KP + es));
                    }
                    if (has2) {
                        asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + es2));
                        if (CM == DEC_Y) asm volatile("prefetch.global.L1 [%0];" ::"l"(sp + KP + es2));
                    }
                }
            }
        }

        // ---- inverse pass 0 (radix R0, no twiddles), in place ----
        {
            float2 v[IT0][R0];
#pragma unroll
            for (int it = 0; it < IT0; it++) {
                const float2* __restrict__ src = buf + (jb + it * JS) * DEC_QS + q;
#pragma unroll
                for (int r = 0; r < R0; r++) v[it][r] = src[r * NB0 * DEC_QS];
            }
            __syncthreads();
#pragma unroll
            for (int it = 0; it < IT0; it++) {
                Dft<R0, +1>::run(v[it]);
                float2* __restrict__ dst = buf + ((jb + it * JS) * R0) * DEC_QS + q;
#pragma unroll
                for (int r = 0; r < R0; r++) dst[r * DEC_QS] = v[it][r];
            }
        }
        if constexpr (ACCUM) {
            // this frame's previous sums have landed (the request for the next frame may still be in flight)
            if constexpr (DEC_PREV_BUFS == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
            else asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncthreads();

        // ---- inverse pass 1 (radix 16): output r of butterfly j is point p = j + r*R0 of the sequence, in hop r/4
        // of the frame; synthesis window, overlap-add in registers, the oldest hop leaves ----
        const long long fme = NST == 2 ? (sid ? fs[NST - 1] : fs[0]) : fs[0];
        const bool emit = fme >= my_h0 && fme < my_h1;
        const long long sbase = fme * H;                          // first sample of the frame
        // samples [e_lo, e_hi) of the frame's first hop go out (all of it, except at segment edges / warm-up frames)
        const int e_lo = (int)max(0LL, min((long long)H, a.seg_begin - sbase));
        const int e_hi = emit ? (int)max(0LL, min((long long)H, a.seg_end - sbase)) : 0;
        float* __restrict__ d1 = o1 + (sbase - a.out_begin);
        float* __restrict__ d2 = o2 + (sbase - a.out_begin);
#pragma unroll
        for (int it = 0; it < 2; it++) {
            const int j = jb + it * JS;
            float2 u[16], tw[16];
            const int m1 = Q * j + qoff1, m2 = Q * j + qoff2;     // + r * R0 * Q: compile-time offsets
            // this band adds to what the outputs hold (the bands before it, in band order): those values are requested
            // first and wait in registers while the butterfly runs
            float2 prev[4];
            if (ACCUM && !staged_prev) {
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const int n1 = m1 + r * R0 * Q, n2 = m2 + r * R0 * Q;
                    prev[r].x = (n1 >= e_lo && n1 < e_hi) ? __ldcs(d1 + n1) : 0.f;
                    prev[r].y = (n2 >= e_lo && n2 < e_hi) ? __ldcs(d2 + n2) : 0.f;
                }
            }
            const float2* __restrict__ src = buf + j * DEC_QS + q;
#pragma unroll
            for (int r = 0; r < 16; r++) u[r] = src[r * NB1 * DEC_QS];
            load_tw16(b.dec.tw_last, j, tw);
#pragma unroll
            for (int r = 1; r < 16; r++) u[r] = cmul(u[r], cconj(tw[r]));
            Dft<16, +1>::run(u);
            const float* __restrict__ syn1 = syn + m1;
            const float* __restrict__ syn2 = syn + m2;
#pragma unroll
            for (int r = 0; r < 16; r++) {
                const float w1 = __ldg(syn1 + r * R0 * Q);
                const float2 ww = CM == DEC_Y ? make_float2(w1, w1) : make_float2(w1, __ldg(syn2 + r * R0 * Q));
                if (r < 4) {
                    const float2 tot = __ffma2_rn(u[r], ww, acc[it][r]);
                    const int n1 = m1 + r * R0 * Q, n2 = m2 + r * R0 * Q;
                    if (ACCUM && staged_prev) {                       // row p = j + r*R0 of the staged hop
                        const float* pb = prevcur + (j + r * R0) * 16;
                        prev[r] = CM == DEC_C16 ? make_float2(pb[sid * RSZ + (q & 7)], pb[sid * RSZ + 8 + (q & 7)])
                                                : make_float2(pb[q], pb[RSZ + q]);
                    }
                    if (n1 >= e_lo && n1 < e_hi) __stcs(d1 + n1, ACCUM ? prev[r].x + tot.x : tot.x);
                    if (n2 >= e_lo && n2 < e_hi) __stcs(d2 + n2, ACCUM ? prev[r].y + tot.y : tot.y);
                }
                if (r >= 4 && r < 12) acc[it][r - 4] = __ffma2_rn(u[r], ww, acc[it][r]);
                if (r >= 12) acc[it][r - 4] = __fmul2_rn(u[r], ww);
            }
        }
        __syncthreads();                                         // buf is free for the next frame's expansion
    }
}


}  // namespace upmix

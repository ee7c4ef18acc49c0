// Shared-memory Stockham FFT building blocks for sm_100a (no cuFFT).
//
// A transform of N = R1*R2*...*Rp complex points lives in one shared-memory buffer of float2 and is
// done in p auto-sorting (Stockham) passes.  In every pass a thread owns whole radix-R butterflies:
// it reads R points at stride N/R (consecutive threads -> consecutive words, conflict-free),
// multiplies by the pass twiddles, runs the radix-R DFT in registers and scatters the R results
// at stride NS (the product of the radices already done).  The buffer is updated in place: all reads
// of a pass finish (barrier) before its writes start.  Several independent rows can share the pass
// (ROWS), which is how the row kernel of the large-N path batches its transforms.
//
// Index padding PAD(i) = i + (i >> 4) keeps the stride-R scatter of the early passes off the same
// 8-byte bank (16 distinct double-banks per half-warp for LDS.64/STS.64).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace upmix {

__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// multiply by DIR*i  (DIR = -1: forward transform, e^{-i...};  DIR = +1: inverse)
template <int DIR>
__device__ __forceinline__ float2 mul_i(float2 a) {
    return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}

// z * exp(DIR * 2*pi*i * m / 16), m in [0, 8), with the trivial cases folded away.
template <int DIR>
__device__ __forceinline__ float2 mul_w16(float2 z, int m) {
    constexpr float C1 = 0.92387953251128674f;   // cos(pi/8)
    constexpr float S1 = 0.38268343236508977f;   // sin(pi/8)
    constexpr float H = 0.70710678118654752f;    // cos(pi/4)
    constexpr float s = DIR < 0 ? -1.f : 1.f;
    switch (m) {
        case 0: return z;
        case 1: return cmul(z, make_float2(C1, s * S1));
        case 2: return make_float2(H * (z.x - s * z.y), H * (z.y + s * z.x));
        case 3: return cmul(z, make_float2(S1, s * C1));
        case 4: return mul_i<DIR>(z);
        case 5: return cmul(z, make_float2(-S1, s * C1));
        case 6: return make_float2(H * (-z.x - s * z.y), H * (s * z.x - z.y));
        default: return cmul(z, make_float2(-C1, s * S1));
    }
}

// In-register radix-R DFT, decimation in time, natural-order output.  R in {2,4,8,16}.
template <int R, int DIR>
struct Dft {
    static __device__ __forceinline__ void run(float2 (&v)[R]) {
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; i++) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        Dft<R / 2, DIR>::run(e);
        Dft<R / 2, DIR>::run(o);
#pragma unroll
        for (int k = 0; k < R / 2; k++) {
            const float2 t = mul_w16<DIR>(o[k], k * (16 / R));
            v[k] = cadd(e[k], t);
            v[k + R / 2] = csub(e[k], t);
        }
    }
};
template <int DIR>
struct Dft<1, DIR> {
    static __device__ __forceinline__ void run(float2 (&)[1]) {}
};

__host__ __device__ constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }

// radix of the next pass when `rem` = N / NS points are still to be combined
__host__ __device__ constexpr int pick_radix(int rem) {
    const int l = ilog2(rem);
    return (l >= 3 && l != 4) ? 8 : (l == 4 || l == 2) ? 4 : 2;
}

__host__ __device__ constexpr int PAD(int i) { return i + (i >> 4); }
__host__ __device__ constexpr int PADSZ(int n) { return n + (n >> 4); }

// One Stockham pass over ROWS rows of N points held in `buf` (row stride PADSZ(N)).
//   FIRST: inputs come from ld(row, idx) instead of buf;  LAST: outputs go to st(row, idx, v).
//   tw[m * tws] = exp(-2*pi*i*m/N)  (forward table; conjugated here for DIR = +1).
template <int N, int R, int NS, int DIR, int T, int ROWS, bool FIRST, bool LAST, bool LD_SMEM, class Ld, class St>
__device__ __forceinline__ void stockham_pass(float2* buf, int tid, const float2* __restrict__ tw, int tws,
                                              Ld& ld, St& st) {
    constexpr int NB = N / R;                      // butterflies per row
    constexpr int TOTAL = NB * ROWS;
    constexpr int IT = (TOTAL + T - 1) / T;
    float2 v[IT][R];
#pragma unroll
    for (int it = 0; it < IT; it++) {
        const int jj = tid + it * T;
        if (TOTAL % T == 0 || jj < TOTAL) {
            const int row = ROWS == 1 ? 0 : jj / NB;
            const int j = ROWS == 1 ? jj : jj - row * NB;
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (FIRST) v[it][r] = ld(row, j + r * NB);
                else v[it][r] = buf[row * PADSZ(N) + PAD(j + r * NB)];
            }
        }
    }
    if (!FIRST || LD_SMEM) __syncthreads();        // in place: every read before any write
#pragma unroll
    for (int it = 0; it < IT; it++) {
        const int jj = tid + it * T;
        if (TOTAL % T == 0 || jj < TOTAL) {
            const int row = ROWS == 1 ? 0 : jj / NB;
            const int j = ROWS == 1 ? jj : jj - row * NB;
            const int k = j & (NS - 1);
            if (NS > 1) {
                constexpr int step = N / (NS * R);
#pragma unroll
                for (int r = 1; r < R; r++) {
                    float2 w = __ldg(&tw[(r * k * step) * tws]);
                    if (DIR > 0) w.y = -w.y;
                    v[it][r] = cmul(v[it][r], w);
                }
            }
            Dft<R, DIR>::run(v[it]);
            const int j0 = (j - k) * R + k;
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (LAST) st(row, j0 + r * NS, v[it][r]);
                else buf[row * PADSZ(N) + PAD(j0 + r * NS)] = v[it][r];
            }
        }
    }
    __syncthreads();
}

template <int N, int NS, int DIR, int T, int ROWS, bool LD_SMEM, class Ld, class St>
__device__ __forceinline__ void stockham_rec(float2* buf, int tid, const float2* __restrict__ tw, int tws,
                                             Ld& ld, St& st) {
    constexpr int R = pick_radix(N / NS);
    constexpr bool LAST = (NS * R == N);
    stockham_pass<N, R, NS, DIR, T, ROWS, NS == 1, LAST, LD_SMEM, Ld, St>(buf, tid, tw, tws, ld, st);
    if constexpr (!LAST) stockham_rec<N, NS * R, DIR, T, ROWS, LD_SMEM, Ld, St>(buf, tid, tw, tws, ld, st);
}

// Full transform of ROWS rows.  ld(row, n) supplies input point n; st(row, k, value) receives output
// point k (natural order).  LD_SMEM says ld reads the same shared buffer (forces the read barrier).
// Ends with a __syncthreads().
template <int N, int DIR, int T, int ROWS, bool LD_SMEM, class Ld, class St>
__device__ __forceinline__ void fft_smem(float2* buf, int tid, const float2* __restrict__ tw, int tws,
                                         Ld ld, St st) {
    static_assert(N >= 8, "transform too small");
    stockham_rec<N, 1, DIR, T, ROWS, LD_SMEM, Ld, St>(buf, tid, tw, tws, ld, st);
}

// ---------------------------------------------------------------------------------------------
// Centre mask (reference: center_extraction.py:373-384; bela/upmix.cpp:363-385), float32.
// coherence = |SL*conj(SR)| / (|SL||SR| + EPS) is evaluated as m / (m + EPS) with m = |SL||SR|
// (identical in exact arithmetic).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void centre_split(float2 sl, float2 sr, float2& c, float2& ls, float2& rs) {
    constexpr float EPS = 1e-12f;
    const float ml = sqrtf(sl.x * sl.x + sl.y * sl.y);
    const float mr = sqrtf(sr.x * sr.x + sr.y * sr.y);
    const float m = ml * mr;
    const float coh = m / (m + EPS);
    const float bal = (ml - mr) / (ml + mr + EPS);
    const float h = 0.5f * (coh * (1.0f - fabsf(bal)));
    c = make_float2(h * (sl.x + sr.x), h * (sl.y + sr.y));
    ls = csub(sl, c);
    rs = csub(sr, c);
}

// Split the spectrum of z = l + i*r at bin k (a = Z[k], b = Z[N-k]) into the two real-signal spectra,
// apply the band gain, run the mask.
__device__ __forceinline__ void split_gain_mask(float2 a, float2 b, float g, float2& c, float2& ls, float2& rs) {
    const float2 sl = make_float2(g * (0.5f * (a.x + b.x)), g * (0.5f * (a.y - b.y)));
    const float2 sr = make_float2(g * (0.5f * (a.y + b.y)), g * (0.5f * (b.x - a.x)));
    centre_split(sl, sr, c, ls, rs);
}

}  // namespace upmix

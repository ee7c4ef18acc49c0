// Shared-memory Stockham FFT building blocks for sm_100a (no cuFFT).
//
// A transform of N = R1*R2*...*Rp complex points lives in one shared-memory buffer of float2 and is
// done in p auto-sorting (Stockham) passes with large in-register radices (up to 32), so that a
// transform of 8192 points needs only three passes (32*16*16) and two trips through shared memory.
// In every pass a thread owns whole radix-R butterflies: it reads R points at stride N/R
// (consecutive threads -> consecutive words, conflict-free), multiplies by the pass twiddles (read
// coalesced from a per-pass table in global memory, L1/L2 resident and frame-invariant), runs the
// radix-R DFT in registers and scatters the R results at stride NS (the product of the radices
// already done).  The buffer is updated in place: all reads of a pass finish (barrier) before its
// writes start.  Several independent rows can share the pass (ROWS), which is how the row kernel of
// the large-N path batches its transforms.
//
// Index padding PAD(i) = i + (i >> log2 R1) keeps the stride-R1 scatter of the first pass off the
// same 8-byte bank (16 distinct double-banks per half-warp for STS.64); later passes have NS >= 16
// and write runs of consecutive words.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#ifndef UPMIX_TW_RECUR
#define UPMIX_TW_RECUR 1          // derive the non-power-of-two pass twiddles instead of loading them
#endif

namespace upmix {

// Packed FP32x2 arithmetic (sm_100: add/mul/fma.f32x2, SASS FADD2 / FMUL2 / FFMA2).  The packed forms take
// operand modifiers -- swapped halves (.LO_HI), one-half negation (.NP), a scalar or an immediate broadcast
// to both halves (.F32) -- so with a complex number in a register pair
//   * complex add / subtract is ONE instruction,
//   * multiplication by +-i is free (it folds into the consuming add as swap + half negation),
//   * a complex multiply is TWO:  t = (a.y, a.x) * (-b.y, b.y);  r = a * (b.x, b.x) + t,
// against 2 / 4 scalar instructions.  The butterflies and the mask are issue-bound (the FMA pipe itself
// is not the limit), so halving their instruction count is what counts.  nvcc maps the intrinsics below
// onto those modifiers (checked in the SASS: FMUL2 R, -R.F32x2.LO_HI.NP, R.F32 ; FFMA2 R, R.F32x2.HI_LO, R.F32, R).
#ifndef UPMIX_SCALAR_CADD
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
#else
__device__ __forceinline__ float2 cadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float2 csub(float2 a, float2 b) { return make_float2(a.x - b.x, a.y - b.y); }
#endif
#ifndef UPMIX_SCALAR_CMUL
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    const float2 t = __fmul2_rn(make_float2(a.y, a.x), make_float2(-b.y, b.y));
    return __ffma2_rn(a, make_float2(b.x, b.x), t);
}
__device__ __forceinline__ float2 cscale(float2 a, float s) { return __fmul2_rn(a, make_float2(s, s)); }
// a * s + b with a real scalar s
__device__ __forceinline__ float2 caxpy(float2 a, float s, float2 b) { return __ffma2_rn(a, make_float2(s, s), b); }
#else
__device__ __forceinline__ float2 cmul(float2 a, float2 b) {
    return make_float2(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
__device__ __forceinline__ float2 cscale(float2 a, float s) { return make_float2(a.x * s, a.y * s); }
__device__ __forceinline__ float2 caxpy(float2 a, float s, float2 b) { return make_float2(fmaf(a.x, s, b.x), fmaf(a.y, s, b.y)); }
#endif

// a * b + c, complex: two packed FMAs
__device__ __forceinline__ float2 cfma(float2 a, float2 b, float2 c) {
    const float2 t = __ffma2_rn(a, make_float2(b.x, b.x), c);
    return __ffma2_rn(make_float2(a.y, a.x), make_float2(-b.y, b.y), t);
}
__device__ __forceinline__ float2 cconj(float2 a) { return make_float2(a.x, -a.y); }

// multiply by DIR*i  (DIR = -1: forward transform, e^{-i...};  DIR = +1: inverse)
template <int DIR>
__device__ __forceinline__ float2 mul_i(float2 a) {
    return DIR < 0 ? make_float2(a.y, -a.x) : make_float2(-a.y, a.x);
}

// z * exp(DIR * 2*pi*i * m / 32), m in [0, 16), with the trivial cases folded away.
template <int DIR>
__device__ __forceinline__ float2 mul_w32(float2 z, int m) {
    constexpr float s = DIR < 0 ? -1.f : 1.f;
    constexpr float H = 0.70710678118654752f;    // cos(pi/4)
    // cos(2 pi m / 32), m = 0..8
    constexpr float C[9] = {1.f, 0.98078528040323045f, 0.92387953251128676f, 0.83146961230254524f, 0.70710678118654752f,
                            0.55557023301960222f, 0.38268343236508977f, 0.19509032201612827f, 0.f};
    if (m == 0) return z;
    if (m == 8) return mul_i<DIR>(z);
    // (1 + s i)/sqrt2 and (-1 + s i)/sqrt2: one packed add (z +- i z) and one packed scale
    if (m == 4) return cscale(cadd(z, make_float2(-s * z.y, s * z.x)), H);
    if (m == 12) return cscale(cadd(z, make_float2(s * z.y, -s * z.x)), -H);
    // general: cos(2 pi m/32) = C[m] (m<8) or -C[16-m]; sin(2 pi m/32) = C[8-m] (m<8) or C[m-8]
    const float c = m < 8 ? C[m] : -C[16 - m];
    const float sn = m < 8 ? C[8 - m] : C[m - 8];
    return cmul(z, make_float2(c, s * sn));
}

// In-register radix-R DFT, decimation in time, natural-order output.  R in {2,4,8,16,32}.
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
            const float2 t = mul_w32<DIR>(o[k], k * (32 / R));
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

// v[k] *= w^k (CONJ: conj(w)^k), k = 1..15, from the four loaded powers w^1, w^2, w^4, w^8: every other power
// is a product of at most four of them, so the rounding stays within a few ulp of a table look-up while
// eleven of fifteen loads go away (the four-step twiddles W_N^{k1 n2} of one column n2).
template <bool CONJ>
__device__ __forceinline__ void apply_powers16(float2 (&v)[16], float2 w1, float2 w2, float2 w4, float2 w8) {
    float2 w[16];
    w[1] = w1; w[2] = w2; w[4] = w4; w[8] = w8;
    w[3] = cmul(w1, w2);
    w[5] = cmul(w1, w4);
    w[6] = cmul(w2, w4);
    w[7] = cmul(w[3], w4);
#pragma unroll
    for (int k = 9; k < 16; k++) w[k] = cmul(w[k - 8], w8);
#pragma unroll
    for (int k = 1; k < 16; k++) v[k] = cmul(v[k], CONJ ? make_float2(w[k].x, -w[k].y) : w[k]);
}

// ---------------------------------------------------------------------------------------------
// TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier helpers: one elected thread moves a frame's
// samples global -> shared asynchronously; consumers wait on the mbarrier's phase parity.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE_%=;\n"
        "bra MBAR_WAIT_%=;\n"
        "MBAR_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// generic-proxy accesses to shared memory made before this fence are ordered before later async-proxy
// (bulk copy) writes of the same locations
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// bytes: multiple of 16; src and dst 16-byte aligned
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// Radix plans.  A plan packs the radices of the passes, 6 bits each, first pass in the low bits.
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int mkplan(int r0, int r1, int r2 = 0, int r3 = 0) {
    return r0 | (r1 << 6) | (r2 << 12) | (r3 << 18);
}
__host__ __device__ constexpr int fft_radix(int plan, int p) { return p > 3 ? 0 : (plan >> (6 * p)) & 63; }
__host__ __device__ constexpr int fft_num_passes(int plan) {
    return fft_radix(plan, 1) == 0 ? 1 : fft_radix(plan, 2) == 0 ? 2 : fft_radix(plan, 3) == 0 ? 3 : 4;
}
// product of the radices of passes 0..p-1
__host__ __device__ constexpr int fft_ns(int plan, int p) {
    return p == 0 ? 1 : fft_ns(plan, p - 1) * fft_radix(plan, p - 1);
}
__host__ __device__ constexpr int fft_size(int plan) { return fft_ns(plan, fft_num_passes(plan)); }
// offset (in float2) of pass p's twiddles inside the transform's table; pass 0 has none
__host__ __device__ constexpr int fft_tw_offset(int plan, int p) {
    return p <= 1 ? 0 : fft_tw_offset(plan, p - 1) + (fft_radix(plan, p - 1) - 1) * fft_ns(plan, p - 1);
}
// total twiddle entries of a transform: table[off_p + (r-1)*NS_p + k] = exp(-2 pi i r k / (NS_p R_p))
__host__ __device__ constexpr int fft_tw_size(int plan) { return fft_tw_offset(plan, fft_num_passes(plan)); }

__host__ __device__ constexpr int pad_shift(int plan) { return ilog2(fft_radix(plan, 0)); }
template <int PLAN>
__host__ __device__ constexpr int PAD(int i) { return i + (i >> pad_shift(PLAN)); }
template <int PLAN>
__host__ __device__ constexpr int PADSZ() { return fft_size(PLAN) + (fft_size(PLAN) >> pad_shift(PLAN)); }

// Output functor of a transform: `pre(row, k)` is called BEFORE the pass barrier and may start global
// loads the store needs (window samples ...); `put(row, k, value, aux)` is called after the butterfly.
// Warps issue in order, so anything loaded in `pre` is in flight across the barrier and the butterfly.
struct NoAux {};
template <class Pre, class Put>
struct StoreFn {
    static constexpr bool WITH_R = false;
    Pre pre;
    Put put;
};
// Variant whose functors also receive r, the output's index within its butterfly: pre(row, k, r), put(row, k,
// value, aux, r).  r is a literal after unrolling, so a store functor can specialise per output at compile time
// (the overlap-add treats the four hops of a frame differently).
template <class Pre, class Put>
struct StoreFnR {
    static constexpr bool WITH_R = true;
    Pre pre;
    Put put;
};
template <class Pre, class Put>
__device__ __forceinline__ StoreFnR<Pre, Put> make_store_r(Pre pre, Put put) { return StoreFnR<Pre, Put>{pre, put}; }
template <class St, bool R = St::WITH_R>
struct AuxOf { using type = decltype(std::declval<St&>().pre(0, 0)); };
template <class St>
struct AuxOf<St, true> { using type = decltype(std::declval<St&>().pre(0, 0, 0)); };
template <class Pre, class Put>
__device__ __forceinline__ StoreFn<Pre, Put> make_store(Pre pre, Put put) { return StoreFn<Pre, Put>{pre, put}; }
template <class Put>
__device__ __forceinline__ auto make_store(Put put) {
    auto pre = [](int, int) { return NoAux{}; };
    return StoreFn<decltype(pre), Put>{pre, put};
}

// One Stockham pass over ROWS rows of N points held in `buf` (row stride PADSZ<PLAN>()).
//   FIRST: inputs come from ld(row, idx, it, r) instead of buf (it, r: which of the thread's
//          butterflies / which input, compile-time after unrolling, so ld may index registers);
//   LAST:  outputs go to st.put(row, idx, v, aux) with aux = st.pre(row, idx) fetched before the barrier.
//   tw: this transform's twiddle table (forward sign; conjugated here for DIR = +1); the pass's
//   twiddles are also fetched before the barrier.
template <int PLAN, int P, int DIR, int T, int ROWS, bool LD_SMEM, class Ld, class St>
__device__ __forceinline__ void stockham_pass(float2* buf, int tid, const float2* __restrict__ tw, Ld& ld, St& st) {
    constexpr int N = fft_size(PLAN);
    constexpr int R = fft_radix(PLAN, P);
    constexpr int NS = fft_ns(PLAN, P);
    constexpr bool FIRST = P == 0;
    constexpr bool LAST = NS * R == N;
    constexpr int NB = N / R;                      // butterflies per row
    constexpr int TOTAL = NB * ROWS;
    constexpr int IT = (TOTAL + T - 1) / T;
    constexpr int RS = PADSZ<PLAN>();
    // padded addresses are affine in r whenever the stride is a multiple of the padding period
    constexpr int PERIOD = 1 << pad_shift(PLAN);
    constexpr bool LD_LIN = NB % PERIOD == 0;
    constexpr int LD_STR = NB + NB / PERIOD;
    constexpr bool ST_LIN = NS % PERIOD == 0 || (NS == 1 && R == PERIOD);
    constexpr int ST_STR = NS == 1 ? 1 : NS + NS / PERIOD;
    float2 v[IT][R];
    float2 w[IT][NS > 1 ? R : 1];
    typename AuxOf<St>::type aux[IT][LAST ? R : 1];
#pragma unroll
    for (int it = 0; it < IT; it++) {
        const int jj = tid + it * T;
        if (TOTAL % T == 0 || jj < TOTAL) {
            const int row = ROWS == 1 ? 0 : jj / NB;
            const int j = ROWS == 1 ? jj : jj - row * NB;
            const int k = j & (NS - 1);
            if (NS > 1) {                          // global loads first: they overlap the barrier
                const float2* __restrict__ twp = tw + fft_tw_offset(PLAN, P) + k;
#pragma unroll
                for (int r = 1; r < R; r++)
                    if (!UPMIX_TW_RECUR || (r & (r - 1)) == 0) w[it][r] = __ldg(twp + (r - 1) * NS);
            }
            if (LAST) {
                const int j0 = (j - k) * R + k;
#pragma unroll
                for (int r = 0; r < R; r++) {
                    if constexpr (St::WITH_R) aux[it][r] = st.pre(row, j0 + r * NS, r);
                    else aux[it][r] = st.pre(row, j0 + r * NS);
                }
            }
            const float2* __restrict__ src = buf + row * RS + PAD<PLAN>(j);
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (FIRST) v[it][r] = ld(row, j + r * NB, it, r);
                else if (LD_LIN) v[it][r] = src[r * LD_STR];
                else v[it][r] = buf[row * RS + PAD<PLAN>(j + r * NB)];
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
#pragma unroll
                for (int r = 1; r < R; r++) {
                    // w_r = w_1^r: only the powers of two are loaded, the rest is one packed complex multiply
                    // away (w_r = w_{r - 2^q} w_{2^q}, at most log2 R - 1 roundings deep) -- the L1 / shared-
                    // memory data pipe is the busiest unit of these kernels, the FMA pipe has room
                    if (UPMIX_TW_RECUR && (r & (r - 1)) != 0) {
                        int hb = 1;
                        while (hb * 2 <= r) hb *= 2;
                        w[it][r] = cmul(w[it][r - hb], w[it][hb]);
                    }
                    float2 ww = w[it][r];
                    if (DIR > 0) ww.y = -ww.y;
                    v[it][r] = cmul(v[it][r], ww);
                }
            }
            Dft<R, DIR>::run(v[it]);
            const int j0 = (j - k) * R + k;
            float2* __restrict__ dst = buf + row * RS + PAD<PLAN>(j0);
#pragma unroll
            for (int r = 0; r < R; r++) {
                if (LAST) {
                    if constexpr (St::WITH_R) st.put(row, j0 + r * NS, v[it][r], aux[it][r], r);
                    else st.put(row, j0 + r * NS, v[it][r], aux[it][r]);
                }
                else if (ST_LIN) dst[r * ST_STR] = v[it][r];
                else buf[row * RS + PAD<PLAN>(j0 + r * NS)] = v[it][r];
            }
        }
    }
    __syncthreads();
}

template <int PLAN, int P, int DIR, int T, int ROWS, bool LD_SMEM, class Ld, class St>
__device__ __forceinline__ void stockham_rec(float2* buf, int tid, const float2* __restrict__ tw, Ld& ld, St& st) {
    stockham_pass<PLAN, P, DIR, T, ROWS, LD_SMEM, Ld, St>(buf, tid, tw, ld, st);
    if constexpr (P + 1 < fft_num_passes(PLAN)) stockham_rec<PLAN, P + 1, DIR, T, ROWS, LD_SMEM, Ld, St>(buf, tid, tw, ld, st);
}

// Passes P .. PEND-1 of a transform (the fused kernel runs the last forward pass and the first inverse pass
// itself, fused with the centre mask).
template <int PLAN, int P, int PEND, int DIR, int T, int ROWS, bool LD_SMEM, class Ld, class St>
__device__ __forceinline__ void stockham_range(float2* buf, int tid, const float2* __restrict__ tw, Ld& ld, St& st) {
    if constexpr (P < PEND) {
        stockham_pass<PLAN, P, DIR, T, ROWS, LD_SMEM, Ld, St>(buf, tid, tw, ld, st);
        stockham_range<PLAN, P + 1, PEND, DIR, T, ROWS, LD_SMEM, Ld, St>(buf, tid, tw, ld, st);
    }
}

// Full transform of ROWS rows of fft_size(PLAN) points.  ld(row, n, it, r) supplies input point n; st
// (a StoreFn) receives output point k (natural order).  LD_SMEM says ld reads the same shared buffer
// (forces the read barrier).  tw = twiddle table of THIS plan (fft_tw_size(PLAN) entries).  Ends with
// a __syncthreads().
template <int PLAN, int DIR, int T, int ROWS, bool LD_SMEM, class Ld, class St>
__device__ __forceinline__ void fft_smem(float2* buf, int tid, const float2* __restrict__ tw, Ld ld, St st) {
    static_assert(fft_num_passes(PLAN) >= 2, "a plan needs at least two passes");
    stockham_rec<PLAN, 0, DIR, T, ROWS, LD_SMEM, Ld, St>(buf, tid, tw, ld, st);
}

// ---------------------------------------------------------------------------------------------
// Band-limited (pruned) row transforms of the four-step path.
//
// The dynamic-resolution rule puts the pass band of every band but the top one below bin ~430
// (bin_low in [32, 64), bin_high = bin_low * f_high/f_low, plus the fade).  For the 16 x N2 four-step
// split that means: of the N2 points of a row only the first K = 32 and -- their mirrors -- the last K
// carry any gain.  The forward row transform then needs only those 2K outputs, and the inverse row
// transform has only those 2K non-zero inputs.  With R0 = R1 = 16 (every row plan) the pruning is
// static per pass:
//   forward  pass 0: full;  pass 1: only outputs r with 16 r in [0,K) u [256-K,256) are consumed;
//            pass 2: only butterflies j in [0,K) (output r = 0) and [NS-K, NS) (output r = R-1) run;
//   inverse  pass 0: only butterflies j in [0,K) (single non-zero input r = 0) and [NB-K, NB) (single
//            input r = 15) run;  pass 1: only inputs r whose block intersects [0,16K) u [N2-16K,N2) are
//            non-zero;  pass 2: full.
// Unneeded outputs are simply not stored (the compiler drops their arithmetic); known-zero inputs are
// skipped by DftNZ, a DFT whose non-zero input set is a template mask.
// ---------------------------------------------------------------------------------------------
constexpr int ROW_K = 32;

__host__ __device__ constexpr unsigned compact_bits(unsigned m, int r, int parity) {
    unsigned o = 0;
    for (int i = 0; i < r / 2; i++)
        if ((m >> (2 * i + parity)) & 1u) o |= 1u << i;
    return o;
}

// In-register radix-R DFT of inputs of which only those with their bit set in NZ are non-zero (the
// others are never read).  Same recursion as Dft<R>, with the all-zero halves skipped.
template <int R, int DIR, unsigned NZ>
struct DftNZ {
    static __device__ __forceinline__ void run(float2 (&v)[R]) {
        constexpr unsigned NE = compact_bits(NZ, R, 0), NO = compact_bits(NZ, R, 1);
        float2 e[R / 2], o[R / 2];
#pragma unroll
        for (int i = 0; i < R / 2; i++) { e[i] = v[2 * i]; o[i] = v[2 * i + 1]; }
        if constexpr (NE != 0) DftNZ<R / 2, DIR, NE>::run(e);
        if constexpr (NO != 0) DftNZ<R / 2, DIR, NO>::run(o);
#pragma unroll
        for (int k = 0; k < R / 2; k++) {
            if constexpr (NO == 0) {
                const float2 x = NE != 0 ? e[k] : make_float2(0.f, 0.f);
                v[k] = x;
                v[k + R / 2] = x;
            } else {
                const float2 t = mul_w32<DIR>(o[k], k * (32 / R));
                if constexpr (NE == 0) {
                    v[k] = t;
                    v[k + R / 2] = make_float2(-t.x, -t.y);
                } else {
                    v[k] = cadd(e[k], t);
                    v[k + R / 2] = csub(e[k], t);
                }
            }
        }
    }
};
template <int DIR, unsigned NZ>
struct DftNZ<1, DIR, NZ> {
    static __device__ __forceinline__ void run(float2 (&)[1]) {}
};

// outputs r of forward pass 1 that some active butterfly of pass 2 reads (NS of pass 2 is 256)
__host__ __device__ constexpr unsigned fwd_p1_out_mask(int K) {
    unsigned m = 0;
    for (int r = 0; r < 16; r++)
        if (16 * r < K || 16 * (r + 1) > 256 - K) m |= 1u << r;
    return m;
}
// inputs r of inverse pass 1 that can be non-zero: pass 0 fills [0, 16K) and [N2 - 16K, N2)
__host__ __device__ constexpr unsigned inv_p1_in_mask(int n2, int K) {
    const int nb1 = n2 / 16;
    unsigned m = 0;
    for (int r = 0; r < 16; r++)
        if (nb1 * r < 16 * K || nb1 * (r + 1) > n2 - 16 * K) m |= 1u << r;
    return m;
}

// Forward transform of ROWS rows when only outputs [0,K) and [N2-K,N2) are needed.  They are handed to
// out(row, idx, value), idx in [0, 2K): idx < K is point idx, idx >= K is point N2 - 2K + idx; buf is
// scratch.  PLAN = {16, 16, R2}.  ld as in fft_smem.  Ends with a __syncthreads().
template <int PLAN, int T, int ROWS, class Ld, class Out>
__device__ __forceinline__ void fft_rows_fwd_pruned(float2* buf, int tid, const float2* __restrict__ tw, Ld ld, Out out) {
    constexpr int N = fft_size(PLAN), RS = PADSZ<PLAN>(), K = ROW_K;
    static_assert(fft_num_passes(PLAN) == 3 && fft_radix(PLAN, 0) == 16 && fft_radix(PLAN, 1) == 16, "row plans are {16,16,R}");
    static_assert(T >= ROWS * 2 * K, "the pruned passes need 2K threads per row");
    auto none = make_store([](int, int, float2, NoAux) {});
    stockham_pass<PLAN, 0, -1, T, ROWS, false, Ld, decltype(none)>(buf, tid, tw, ld, none);
    {   // pass 1: full butterflies, only the consumed outputs are stored
        constexpr int R = 16, NS = 16, NB = N / R, TOTAL = ROWS * NB, IT = (TOTAL + T - 1) / T;
        constexpr unsigned OUT = fwd_p1_out_mask(K);
        constexpr int LD_STR = NB + NB / 16, ST_STR = NS + NS / 16;
        float2 v[IT][R], w[IT][R];
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const int jj = tid + it * T;
            if (TOTAL % T == 0 || jj < TOTAL) {
                const int row = ROWS == 1 ? 0 : jj / NB, j = jj - row * NB, k = j & (NS - 1);
                const float2* __restrict__ twp = tw + fft_tw_offset(PLAN, 1) + k;
#pragma unroll
                for (int r = 1; r < R; r++)
                    if (!UPMIX_TW_RECUR || (r & (r - 1)) == 0) w[it][r] = __ldg(twp + (r - 1) * NS);
                const float2* __restrict__ src = buf + row * RS + PAD<PLAN>(j);
#pragma unroll
                for (int r = 0; r < R; r++) v[it][r] = src[r * LD_STR];
            }
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const int jj = tid + it * T;
            if (TOTAL % T == 0 || jj < TOTAL) {
                const int row = ROWS == 1 ? 0 : jj / NB, j = jj - row * NB, k = j & (NS - 1);
#pragma unroll
                for (int r = 1; r < R; r++) {
                    if (UPMIX_TW_RECUR && (r & (r - 1)) != 0) {      // as in stockham_pass: derived, not loaded
                        int hb = 1;
                        while (hb * 2 <= r) hb *= 2;
                        w[it][r] = cmul(w[it][r - hb], w[it][hb]);
                    }
                    v[it][r] = cmul(v[it][r], w[it][r]);
                }
                Dft<R, -1>::run(v[it]);
                float2* __restrict__ dst = buf + row * RS + PAD<PLAN>((j - k) * R + k);
#pragma unroll
                for (int r = 0; r < R; r++)
                    if ((OUT >> r) & 1u) dst[r * ST_STR] = v[it][r];
            }
        }
        __syncthreads();
    }
    {   // pass 2: 2K butterflies per row, one output each (r = 0 below, r = R-1 above)
        constexpr int R = fft_radix(PLAN, 2), NS = 256, NB = N / R;
        static_assert(NB == 256, "pass 2 of a row plan has 256 butterflies");
        constexpr int LD_STR = NB + NB / 16;
        const bool act = tid < ROWS * 2 * K;
        const int row = tid / (2 * K), a = tid - row * 2 * K;
        const bool low = a < K;                             // warp-uniform: K is a multiple of 32
        const int j = low ? a : NS - 2 * K + a;
        if (act) {
            float2 v[R], w[R];
            const float2* __restrict__ twp = tw + fft_tw_offset(PLAN, 2) + j;
#pragma unroll
            for (int r = 1; r < R; r++) w[r] = __ldg(twp + (r - 1) * NS);
            const float2* __restrict__ src = buf + row * RS + PAD<PLAN>(j);
#pragma unroll
            for (int r = 0; r < R; r++) v[r] = src[r * LD_STR];
#pragma unroll
            for (int r = 1; r < R; r++) v[r] = cmul(v[r], w[r]);
            Dft<R, -1>::run(v);                             // the compiler keeps only the output used
            out(row, a, low ? v[0] : v[R - 1]);
        }
        __syncthreads();
    }
}

// Inverse transform of ROWS rows whose only non-zero inputs are points [0,K) and [N2-K,N2), supplied by
// in(row, idx) with idx as in fft_rows_fwd_pruned.  st as in fft_smem (last pass output functor).
template <int PLAN, int T, int ROWS, class In, class St>
__device__ __forceinline__ void fft_rows_inv_pruned(float2* buf, int tid, const float2* __restrict__ tw, In in, St st) {
    constexpr int N = fft_size(PLAN), RS = PADSZ<PLAN>(), K = ROW_K;
    static_assert(T >= ROWS * 2 * K, "the pruned passes need 2K threads per row");
    {   // pass 0: 2K butterflies per row, one non-zero input each; nothing of buf is read
        constexpr int R = 16, NB = N / R;
        if (tid < ROWS * 2 * K) {
            const int row = tid / (2 * K), a = tid - row * 2 * K;
            const bool low = a < K;
            const int j = low ? a : NB - 2 * K + a;
            const float2 x = in(row, a);
            float2 v[R];
            float2* __restrict__ dst = buf + row * RS + PAD<PLAN>(j * R);      // NS = 1: outputs j*16 + r, contiguous
            if (low) {
                v[0] = x;
                DftNZ<R, +1, 1u>::run(v);
            } else {
                v[R - 1] = x;
                DftNZ<R, +1, 1u << (R - 1)>::run(v);
            }
#pragma unroll
            for (int r = 0; r < R; r++) dst[r] = v[r];
        }
        __syncthreads();
    }
    {   // pass 1: every butterfly, but only the inputs that pass 0 can have filled
        constexpr int R = 16, NS = 16, NB = N / R, TOTAL = ROWS * NB, IT = (TOTAL + T - 1) / T;
        constexpr unsigned NZ = inv_p1_in_mask(N, K);
        constexpr int LD_STR = NB + NB / 16, ST_STR = NS + NS / 16;
        float2 v[IT][R], w[IT][R];
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const int jj = tid + it * T;
            if (TOTAL % T == 0 || jj < TOTAL) {
                const int row = ROWS == 1 ? 0 : jj / NB, j = jj - row * NB, k = j & (NS - 1);
                const float2* __restrict__ twp = tw + fft_tw_offset(PLAN, 1) + k;
                const float2* __restrict__ src = buf + row * RS + PAD<PLAN>(j);
#pragma unroll
                for (int r = 0; r < R; r++)
                    if ((NZ >> r) & 1u) {
                        if (r > 0) w[it][r] = __ldg(twp + (r - 1) * NS);
                        v[it][r] = src[r * LD_STR];
                    }
            }
        }
        __syncthreads();
#pragma unroll
        for (int it = 0; it < IT; it++) {
            const int jj = tid + it * T;
            if (TOTAL % T == 0 || jj < TOTAL) {
                const int row = ROWS == 1 ? 0 : jj / NB, j = jj - row * NB, k = j & (NS - 1);
#pragma unroll
                for (int r = 1; r < R; r++)
                    if ((NZ >> r) & 1u) v[it][r] = cmul(v[it][r], make_float2(w[it][r].x, -w[it][r].y));
                DftNZ<R, +1, NZ>::run(v[it]);
                float2* __restrict__ dst = buf + row * RS + PAD<PLAN>((j - k) * R + k);
#pragma unroll
                for (int r = 0; r < R; r++) dst[r * ST_STR] = v[it][r];
            }
        }
        __syncthreads();
    }
    auto nold = [](int, int, int, int) -> float2 { return make_float2(0.f, 0.f); };
    stockham_pass<PLAN, 2, +1, T, ROWS, true, decltype(nold), St>(buf, tid, tw, nold, st);
}

// ---------------------------------------------------------------------------------------------
// Centre mask (reference: center_extraction.py:373-384; bela/upmix.cpp:363-385), float32.
// coherence = |SL*conj(SR)| / (|SL||SR| + EPS) is evaluated as m / (m + EPS) with m = |SL||SR|
// (identical in exact arithmetic), balance = (|SL|-|SR|) / (|SL|+|SR|+EPS).
// ---------------------------------------------------------------------------------------------
// MUFU-based square root and reciprocal (max rel. error 2^-22), flush-to-zero variants: one SASS
// instruction each (MUFU.SQRT / MUFU.RCP) -- the non-ftz forms carry a denormal rescue (FSETP + two
// predicated FMULs) that the mask does not need: a denormal |P|^2 means a magnitude below 1e-19, where
// the centre factor is ~m/EPS ~ 0 anyway, and the divisor is >= EPS^2 = 1e-24, a normal number.
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// One bin of the frame: a = Z[k], b = Z[N-k] of the packed transform Z = FFT(l + i r).  With
// P = a + conj b and Q = a - conj b the two real-signal spectra are SL = P/2, SR = -i Q/2; after the band
// gain g:  |SL| = g|P|/2, |SR| = g|Q|/2, SL + SR = (g/2)(P.x + Q.y, P.y - Q.x).
//   centerFactor = coherence (1 - |balance|) = m (s - |d|) / ((m + EPS) s),
//   m = |SL||SR|, s = |SL| + |SR| + EPS, d = |SL| - |SR|           (one division instead of two)
//   C = 0.5 centerFactor (SL + SR)
// and, because SL + i SR = g a and conj SL + i conj SR = g b, the packed spectrum of Ls + i Rs is
//   Y[k] = g a - (1+i) C,   Y[N-k] = g b - (1+i) conj C            (Ls = SL - C, Rs = SR - C).
// The magnitudes and the quotient use the approximate (2^-22) square root and division: the factor only
// scales SL + SR, so its relative error (< 1e-6) sits 120 dB below the signal; the denominator is >= 1e-24,
// far from the approximate divider's limits.
__device__ __forceinline__ void mask_bin(float2 a, float2 b, float g, float2& y_lo, float2& y_hi, float2& c) {
    constexpr float EPS = 1e-12f;
    const float2 P = cadd(a, make_float2(b.x, -b.y));          // a + conj b
    const float2 Q = cadd(a, make_float2(-b.x, b.y));          // a - conj b
    const float hg = 0.5f * g;
    const float ml = hg * sqrt_approx(fmaf(P.x, P.x, P.y * P.y));
    const float mr = hg * sqrt_approx(fmaf(Q.x, Q.x, Q.y * Q.y));
    const float m = ml * mr;
    const float sden = ml + mr + EPS;
    const float cf = (m * (sden - fabsf(ml - mr))) * rcp_approx((m + EPS) * sden);
    const float t = (0.5f * cf) * hg;
    c = cscale(cadd(P, make_float2(Q.y, -Q.x)), t);            // t (P - i Q) = t (px + qy, py - qx)
    const float2 u = cadd(make_float2(c.x, c.x), make_float2(-c.y, c.y));   // (1+i) C = (cx - cy, cx + cy)
    y_lo = caxpy(a, g, make_float2(-u.x, -u.y));
    y_hi = caxpy(b, g, make_float2(-u.y, -u.x));               // (1+i) conj C = (cx + cy, cx - cy)
}

// Bands that share an STFT (same size, hop and windows) share the forward transform and -- the inverse
// transform and the overlap-add being linear -- the inverse too: their masked spectra are summed per
// bin.  `gain` points at bin k of the first gain table; tables are `stride` floats apart and sorted per
// bin so that the non-zero gains come first (g0 = the first one, already loaded).
__device__ __forceinline__ void mask_bin_merged(float2 a, float2 b, float g0, const float* __restrict__ gain, int n_gains,
                                                int stride, float2& y_lo, float2& y_hi, float2& c) {
    if (g0 == 0.f) {
        y_lo = y_hi = c = make_float2(0.f, 0.f);
        return;
    }
    mask_bin(a, b, g0, y_lo, y_hi, c);
#pragma unroll 1
    for (int q = 1; q < n_gains; q++) {                  // single-band pipelines never enter
        const float g = __ldg(gain + (long long)q * stride);
        if (g == 0.f) break;
        float2 yl, yh, cc;
        mask_bin(a, b, g, yl, yh, cc);
        y_lo = cadd(y_lo, yl);
        y_hi = cadd(y_hi, yh);
        c = cadd(c, cc);
    }
}

// z[k] = (C[k] + conj C[M-k]) + i e^{+2 pi i k/N} (C[k] - conj C[M-k]) and z[M-k], the packed spectrum whose
// M-point inverse is c[2m] + i c[2m+1] (M = N/2; wk = exp(-2 pi i k / N)).
__device__ __forceinline__ void pack_pair(float2 ck, float2 cmk, float2 wk, float2& zk, float2& zmk) {
    const float2 A = cadd(ck, make_float2(cmk.x, -cmk.y));
    const float2 B = cadd(ck, make_float2(-cmk.x, cmk.y));
    const float2 D = cmul(B, make_float2(wk.x, -wk.y));
    zk = cadd(A, make_float2(-D.y, D.x));
    zmk = cadd(make_float2(A.x, -A.y), make_float2(D.y, D.x));
}

// ---------------------------------------------------------------------------------------------
// Fused middle of a frame: last forward pass -> split / gain / centre mask -> first pass of the inverse
// transform of Ls + i Rs and of the packed centre, all in registers.
//
// The last forward pass (radix RL, NSL = N/RL butterflies) gives butterfly j the bins j + r NSL, and the
// first pass of an inverse transform that starts with the same radix reads exactly those bins.  The mirror
// of bin j + r NSL is (NSL - j) + (RL-1-r) NSL -- a bin of butterfly NSL - j -- so a thread that owns the
// butterfly PAIR (j, NSL - j) holds every mirror pair the mask needs, and both inverse butterflies' inputs.
// Bins k and M - k (M = N/2), which the packed centre spectrum combines, live in the same pair as well.
// Thread 0 owns the two self-mirrored butterflies 0 and NSL/2 (bins 0 and N/2 are their own mirrors).
// Against separate passes this saves, per frame, one store and two loads of the whole spectrum, one store
// and one load of the centre spectrum, three barriers and the mask's index arithmetic: the kernels wait on
// the shared-memory / L1 data pipe (81 % busy at 1024 points), not on arithmetic.
//   Z:  spectrum buffer; read with the forward plan's padding, written with the inverse plan's
//   Cz: packed centre spectrum (half plan's padding), written unless `fold`
// MODE as band_fused_kernel.  Starts after the barrier of the previous forward pass, ends with a barrier.
// ---------------------------------------------------------------------------------------------
template <int N, int T, int PFW, int PIV, int PHV, bool MERGED, bool FUSE_HALF>
__device__ __forceinline__ void mega_phase(float2* Z, float2* Cz, int tid, const float2* __restrict__ twf,
                                           const float* __restrict__ gain, int n_gains, int gain_stride,
                                           const float2* __restrict__ twp, bool fold) {
    constexpr int PL = fft_num_passes(PFW) - 1;
    constexpr int RL = fft_radix(PFW, PL);
    constexpr int RH = RL / 2;
    constexpr int NSL = N / RL;
    constexpr int M = N / 2;
    static_assert(fft_radix(PIV, 0) == RL && (!FUSE_HALF || fft_radix(PHV, 0) == RH) && 2 * T == NSL && RL >= 4, "plans do not line up");
    const bool special = tid == 0;
    const int j1 = tid, j2 = special ? NSL / 2 : NSL - tid;
    float2 v1[RL], v2[RL], w1[RL], w2[RL], wp[RH];
    float g[RL + 1];
    int gk[RL + 1];                                      // bin of g[] (merged pipelines index more gain tables with it)
    {   // everything that comes from global memory first
        const float2* __restrict__ t1 = twf + fft_tw_offset(PFW, PL) + j1;
        const float2* __restrict__ t2 = twf + fft_tw_offset(PFW, PL) + j2;
#pragma unroll
        for (int r = 1; r < RL; r++)
            if (!UPMIX_TW_RECUR || (r & (r - 1)) == 0) {
                w1[r] = __ldg(t1 + (r - 1) * NSL);
                w2[r] = __ldg(t2 + (r - 1) * NSL);
            }
        if (!special) {
#pragma unroll
            for (int r = 0; r < RL; r++) {               // pair r: bins j1 + r NSL  <->  j2 + (RL-1-r) NSL
                const int k = j1 + r * NSL;
                gk[r] = k <= M ? k : N - k;
                g[r] = __ldg(gain + gk[r]);
            }
            gk[RL] = 0;
            g[RL] = 0.f;
#pragma unroll
            for (int r = 0; r < RH; r++) wp[r] = __ldg(twp + j1 + r * NSL);
        } else {                                         // (its pairs' gains are loaded by the lanes that mask them)
#pragma unroll
            for (int r = 0; r < RH; r++)                 // k = (r+1) NSL for r < RL/4 (k = 0 needs none);  then k = NSL/2 + r' NSL
                wp[r] = __ldg(twp + (r < RL / 4 ? (r + 1) * NSL : NSL / 2 + (r - RL / 4) * NSL));
        }
        const float2* __restrict__ s1 = Z + PAD<PFW>(j1);
        const float2* __restrict__ s2 = Z + PAD<PFW>(j2);
        constexpr int LD_STR = NSL + NSL / (1 << pad_shift(PFW));
        static_assert(NSL % (1 << pad_shift(PFW)) == 0, "padded stride must be affine");
#pragma unroll
        for (int r = 0; r < RL; r++) {
            v1[r] = s1[r * LD_STR];
            v2[r] = s2[r * LD_STR];
        }
    }
    __syncthreads();                                     // every read of Z before any write
    // one butterfly after the other: the second one's twiddles need not be live during the first
#pragma unroll
    for (int r = 1; r < RL; r++) {
        if (UPMIX_TW_RECUR && (r & (r - 1)) != 0) {
            int hb = 1;
            while (hb * 2 <= r) hb *= 2;
            w1[r] = cmul(w1[r - hb], w1[hb]);
        }
        v1[r] = cmul(v1[r], w1[r]);
    }
    Dft<RL, -1>::run(v1);
#pragma unroll
    for (int r = 1; r < RL; r++) {
        if (UPMIX_TW_RECUR && (r & (r - 1)) != 0) {
            int hb = 1;
            while (hb * 2 <= r) hb *= 2;
            w2[r] = cmul(w2[r - hb], w2[hb]);
        }
        v2[r] = cmul(v2[r], w2[r]);
    }
    Dft<RL, -1>::run(v2);

    // packed centre bins: zA[r] is bin j1 + r NSL, zB[r] bin j2 + r NSL -- the inputs of the half-size
    // transform's first pass when it starts with radix RH (FUSE_HALF: kept in registers), otherwise
    // stored at their natural place for a transform that runs all its passes
    float2 zA[FUSE_HALF ? RH : 1], zB[FUSE_HALF ? RH : 1];
    auto put_a = [&](int r, float2 z) {
        if constexpr (FUSE_HALF) zA[r] = z;
        else Cz[PAD<PHV>(j1 + r * NSL)] = z;
    };
    auto put_b = [&](int r, float2 z) {
        if constexpr (FUSE_HALF) zB[r] = z;
        else Cz[PAD<PHV>(j2 + r * NSL)] = z;
    };
    auto mask2 = [&](float2& a, float2& bb, int gi, float2& c) {     // mask one mirror pair in place
        float2 ylo, yhi;
        if constexpr (MERGED) mask_bin_merged(a, bb, g[gi], gain + gk[gi], n_gains, gain_stride, ylo, yhi, c);
        else mask_bin(a, bb, g[gi], ylo, yhi, c);
        if (fold) {                                      // (Ls + C/2) + i (Rs + C/2): add (1+i) C / 2
            const float2 u = cadd(make_float2(c.x, c.x), make_float2(-c.y, c.y));
            ylo = caxpy(u, 0.5f, ylo);
            yhi = caxpy(make_float2(u.y, u.x), 0.5f, yhi);
        }
        a = ylo;
        bb = yhi;
    };
    if (!special) {
        // Pairs r and RH + r (bins k and k + M) feed the same packed-centre bins (k and M - k), so they are
        // masked together, two such couples per step: four masks in flight, their centre values consumed
        // at once.  A step whose four gains are all zero only clears its bins (most of a low band's spectrum).
        constexpr int STEP = RH >= 2 ? 2 : 1;
#pragma unroll
        for (int r0 = 0; r0 < RH; r0 += STEP) {
            bool any = false;
#pragma unroll
            for (int r = r0; r < r0 + STEP; r++) any = any || g[r] != 0.f || g[RH + r] != 0.f;
            if (any) {
#pragma unroll
                for (int r = r0; r < r0 + STEP; r++) {
                    float2 ca, cb;
                    mask2(v1[r], v2[RL - 1 - r], r, ca);
                    mask2(v1[RH + r], v2[RH - 1 - r], RH + r, cb);
                    // C[M-k] = conj C[M+k], and bin M+k = j1 + (RH+r) NSL
                    if (!fold) {
                        float2 zk, zmk;
                        pack_pair(ca, make_float2(cb.x, -cb.y), wp[r], zk, zmk);
                        put_a(r, zk);
                        put_b(RH - 1 - r, zmk);
                    }
                }
            } else {
#pragma unroll
                for (int r = r0; r < r0 + STEP; r++) {
                    v1[r] = v2[RL - 1 - r] = v1[RH + r] = v2[RH - 1 - r] = make_float2(0.f, 0.f);
                    if (!fold) {
                        put_a(r, make_float2(0.f, 0.f));
                        put_b(RH - 1 - r, make_float2(0.f, 0.f));
                    }
                }
            }
        }
    }
    // Thread 0's two self-mirrored butterflies hold 2 RL + 1 mirror pairs of their own.  Masked by that one
    // thread they would cost its warp a second mask round of full length, and the CTA a wait at the next barrier
    // (4 % of the 8192-point kernel); instead the first lanes of warp 0 take one pair each: thread 0 parks its 2 RL
    // values in a small shared scratch, lanes 0 .. RL mask one pair apiece, thread 0 collects the results.
    __shared__ float2 sp[2 * RL], sc[RL + 1];
    if (tid < 32) {                                      // warp 0
        if (special) {
#pragma unroll
            for (int r = 0; r < RL; r++) {
                sp[r] = v1[r];
                sp[RL + r] = v2[r];
            }
        }
        __syncwarp();
        if (tid <= RL) {
            // pair p: p <= RH: bins p NSL <-> (RL - p) NSL of butterfly 0 (p = 0 and p = RH are their own mirrors);
            //         p > RH:  bins NSL/2 + r NSL <-> NSL/2 + (RL-1-r) NSL of butterfly NSL/2, r = p - RH - 1
            const int p = tid;
            const int r = p - RH - 1;
            const int ia = p <= RH ? p : RL + r;
            const int ib = p <= RH ? (RL - p) & (RL - 1) : 2 * RL - 1 - r;
            const int bin = p <= RH ? p * NSL : NSL / 2 + r * NSL;
            const float gp = __ldg(gain + bin);
            const float2 a = sp[ia], bb = sp[ib];
            float2 ylo, yhi, c;
            if constexpr (MERGED) mask_bin_merged(a, bb, gp, gain + bin, n_gains, gain_stride, ylo, yhi, c);
            else mask_bin(a, bb, gp, ylo, yhi, c);
            if (fold) {
                const float2 u = cadd(make_float2(c.x, c.x), make_float2(-c.y, c.y));
                ylo = caxpy(u, 0.5f, ylo);
                yhi = caxpy(make_float2(u.y, u.x), 0.5f, yhi);
            }
            __syncwarp((2u << RL) - 1u);                 // every pair read before any is overwritten
            sp[ia] = ylo;
            if (ib != ia) sp[ib] = yhi;
            sc[p] = c;
        }
        __syncwarp();
        if (special) {
#pragma unroll
            for (int r = 0; r < RL; r++) {
                v1[r] = sp[r];
                v2[r] = sp[RL + r];
            }
            if (!fold) {
#pragma unroll
                for (int r = 0; r <= RL / 4; r++) {      // k = r NSL, M - k = (RH - r) NSL
                    float2 zk, zmk;
                    pack_pair(sc[r], sc[RH - r], r == 0 ? make_float2(1.f, 0.f) : wp[r - 1], zk, zmk);
                    put_a(r, zk);
                    if (r > 0) put_a(RH - r, zmk);
                }
#pragma unroll
                for (int r = 0; r < RL / 4; r++) {       // k = NSL/2 + r NSL, M - k = NSL/2 + (RH-1-r) NSL
                    float2 zk, zmk;
                    pack_pair(sc[RH + 1 + r], sc[RH + 1 + RH - 1 - r], wp[RL / 4 + r], zk, zmk);
                    put_b(r, zk);
                    put_b(RH - 1 - r, zmk);
                }
            }
        }
    }

    Dft<RL, +1>::run(v1);                                // first inverse pass: no twiddles, outputs j RL + r
    Dft<RL, +1>::run(v2);
    float2* __restrict__ d1 = Z + PAD<PIV>(j1 * RL);
    float2* __restrict__ d2 = Z + PAD<PIV>(j2 * RL);
#pragma unroll
    for (int r = 0; r < RL; r++) {
        d1[r] = v1[r];
        d2[r] = v2[r];
    }
    if (FUSE_HALF && !fold) {
        Dft<FUSE_HALF ? RH : 1, +1>::run(zA);
        Dft<FUSE_HALF ? RH : 1, +1>::run(zB);
        float2* __restrict__ e1 = Cz + PAD<PHV>(j1 * RH);
        float2* __restrict__ e2 = Cz + PAD<PHV>(j2 * RH);
#pragma unroll
        for (int r = 0; r < (FUSE_HALF ? RH : 1); r++) {
            e1[r] = zA[r];
            e2[r] = zB[r];
        }
    }
    __syncthreads();
}

}  // namespace upmix

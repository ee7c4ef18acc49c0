// Dispatch of the decimated band kernels (upmix_dec.cuh; one translation unit per P: upmix_dec_128/256/512.cu).
#include "fft_device.cuh"
#include "upmix_kernels.cuh"
#include "upmix_launch.h"

namespace upmix {

cudaError_t launch_dec_fwd_128(const BandDev& b, const SegArgs& a, const DecWave& w, int n_tracks, cudaStream_t st);
cudaError_t launch_dec_fwd_256(const BandDev& b, const SegArgs& a, const DecWave& w, int n_tracks, cudaStream_t st);
cudaError_t launch_dec_fwd_512(const BandDev& b, const SegArgs& a, const DecWave& w, int n_tracks, cudaStream_t st);
cudaError_t launch_dec_inv_128(const BandDev& b, const SegArgs& a, const DecWave& w, int n_runs, int n_tracks, bool centre, cudaStream_t st);
cudaError_t launch_dec_inv_256(const BandDev& b, const SegArgs& a, const DecWave& w, int n_runs, int n_tracks, bool centre, cudaStream_t st);
cudaError_t launch_dec_inv_512(const BandDev& b, const SegArgs& a, const DecWave& w, int n_runs, int n_tracks, bool centre, cudaStream_t st);

// Q > 16: partial sums of the Q/16 groups, added in group order, then the mask (center_extraction.py:370-384)
__global__ void __launch_bounds__(128) dec_mask_kernel(const BandDev b, const DecWave w, const int fold) {
    const int k = blockIdx.y * 128 + threadIdx.x;
    if (k > b.dec.K) return;
    const int fl = blockIdx.x, track = blockIdx.z;
    const int KP = b.dec.KP, G = b.dec.Q / 16;
    const float2* __restrict__ p = w.part + (((long long)track * w.n_frames + fl) * G) * 2 * KP + k;
    float2 za = p[0], zb = k ? p[KP] : make_float2(0.f, 0.f);
    for (int g = 1; g < G; g++) {
        za = cadd(za, p[(long long)g * 2 * KP]);
        if (k) zb = cadd(zb, p[(long long)g * 2 * KP + KP]);
    }
    if (k == 0) zb = za;
    float2 ylo, yhi, c;
    mask_bin_merged(za, zb, __ldg(b.gain + k), b.gain + k, b.n_gains, b.gain_stride, ylo, yhi, c);
    if (fold) {
        const float2 uu = cadd(make_float2(c.x, c.x), make_float2(-c.y, c.y));
        ylo = caxpy(uu, 0.5f, ylo);
        yhi = caxpy(make_float2(uu.y, uu.x), 0.5f, yhi);
    }
    float2* __restrict__ sp = w.spec + (((long long)track * w.n_frames + fl) * 3) * KP;
    sp[k] = ylo;
    sp[KP + k] = yhi;
    sp[2 * KP + k] = c;
}

unsigned long long& launch_counter();
cudaError_t launch_dec_mask(const BandDev& b, const DecWave& w, int fold, int n_tracks, cudaStream_t st) {
    dec_mask_kernel<<<dim3(w.n_frames, (b.dec.K + 128) / 128, n_tracks), 128, 0, st>>>(b, w, fold);
    launch_counter()++;
    return cudaGetLastError();
}

// forward transform, twiddled sums and mask of the wave's frames
cudaError_t launch_dec_fwd(const BandDev& b, const SegArgs& a, const DecWave& w, int n_tracks, cudaStream_t st) {
    switch (b.dec.P) {
        case 128: return launch_dec_fwd_128(b, a, w, n_tracks, st);
        case 256: return launch_dec_fwd_256(b, a, w, n_tracks, st);
        case 512: return launch_dec_fwd_512(b, a, w, n_tracks, st);
        default: return cudaErrorInvalidValue;
    }
}

// Ls + i Rs (centre = false: out_l / out_r) or the centre (centre = true: out_c) of hops [a.hop_begin, a.hop_end) in runs
// of a.hops_per_run; the wave must hold frames max(0, hop_begin - 3) .. hop_end - 1
cudaError_t launch_dec_inv(const BandDev& b, const SegArgs& a, const DecWave& w, int n_runs, int n_tracks, bool centre,
                           cudaStream_t st) {
    switch (b.dec.P) {
        case 128: return launch_dec_inv_128(b, a, w, n_runs, n_tracks, centre, st);
        case 256: return launch_dec_inv_256(b, a, w, n_runs, n_tracks, centre, st);
        case 512: return launch_dec_inv_512(b, a, w, n_runs, n_tracks, centre, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace upmix

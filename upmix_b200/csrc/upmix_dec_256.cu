// dec_*_kernel instantiations for P = 256 points per decimated sequence (see upmix_dec.cuh).
#include "upmix_dec.cuh"

namespace upmix {

unsigned long long& launch_counter();

template <int P, int Q>
static cudaError_t dec_fwd_pq(const BandDev& b, const SegArgs& a, const DecWave& w, int n_tracks, cudaStream_t st) {
    constexpr bool FUSE = Q == 16;
    const int smem = (P * DEC_QS + (FUSE ? 2 * b.dec.KP : 0)) * (int)sizeof(float2);
    cudaError_t e = cudaFuncSetAttribute(dec_fwd_kernel<P, Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    dec_fwd_kernel<P, Q><<<dim3(w.n_frames, Q / 16, n_tracks), P / 2, smem, st>>>(b, a, w);
    launch_counter()++;
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    if (!FUSE) return launch_dec_mask(b, w, a.fold, n_tracks, st);
    return cudaSuccess;
}

template <int P, int Q, bool CENTRE, bool ACCUM>
static cudaError_t dec_inv_pqca(const BandDev& b, const SegArgs& a, const DecWave& w, int n_runs, int n_tracks, cudaStream_t st) {
    const int smem = P * DEC_QS * (int)sizeof(float2) + (ACCUM ? DEC_PREV_BUFS * 2 * (P / 4) * 16 * (int)sizeof(float) : 0);
    const int gx = !CENTRE ? Q / 16 : Q == 16 ? 1 : Q / 32;                // tiles per run of hops
    const int gy = CENTRE && Q == 16 ? (n_runs + 1) / 2 : n_runs;          // (two runs per tile)
    cudaError_t e = cudaFuncSetAttribute(dec_inv_kernel<P, Q, CENTRE, ACCUM>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    dec_inv_kernel<P, Q, CENTRE, ACCUM><<<dim3(gy, gx, n_tracks), P / 2, smem, st>>>(b, a, w);
    launch_counter()++;
    return cudaGetLastError();
}

template <int P, int Q>
static cudaError_t dec_inv_pq(const BandDev& b, const SegArgs& a, const DecWave& w, int n_runs, int n_tracks, bool centre, cudaStream_t st) {
    if (centre) return a.accum ? dec_inv_pqca<P, Q, true, true>(b, a, w, n_runs, n_tracks, st) : dec_inv_pqca<P, Q, true, false>(b, a, w, n_runs, n_tracks, st);
    return a.accum ? dec_inv_pqca<P, Q, false, true>(b, a, w, n_runs, n_tracks, st) : dec_inv_pqca<P, Q, false, false>(b, a, w, n_runs, n_tracks, st);
}

cudaError_t launch_dec_fwd_256(const BandDev& b, const SegArgs& a, const DecWave& w, int n_tracks, cudaStream_t st) {
    switch (b.dec.Q) {
        case 16: return dec_fwd_pq<256, 16>(b, a, w, n_tracks, st);
        case 32: return dec_fwd_pq<256, 32>(b, a, w, n_tracks, st);
        case 64: return dec_fwd_pq<256, 64>(b, a, w, n_tracks, st);
        case 128: return dec_fwd_pq<256, 128>(b, a, w, n_tracks, st);
        case 256: return dec_fwd_pq<256, 256>(b, a, w, n_tracks, st);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_dec_inv_256(const BandDev& b, const SegArgs& a, const DecWave& w, int n_runs, int n_tracks, bool centre, cudaStream_t st) {
    switch (b.dec.Q) {
        case 16: return dec_inv_pq<256, 16>(b, a, w, n_runs, n_tracks, centre, st);
        case 32: return dec_inv_pq<256, 32>(b, a, w, n_runs, n_tracks, centre, st);
        case 64: return dec_inv_pq<256, 64>(b, a, w, n_runs, n_tracks, centre, st);
        case 128: return dec_inv_pq<256, 128>(b, a, w, n_runs, n_tracks, centre, st);
        case 256: return dec_inv_pq<256, 256>(b, a, w, n_runs, n_tracks, centre, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace upmix

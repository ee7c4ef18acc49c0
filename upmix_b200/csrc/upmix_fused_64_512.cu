// band_fused_kernel instantiations for sizes 64, 128, 256, 512 (see upmix_fused.cuh).
#include "upmix_fused.cuh"
#include "upmix_launch.h"

namespace upmix {

cudaError_t launch_band_fused_64_512(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st) {
    switch (b.n_fft) {
        case 64: return launch_fused_n<64>(b, a, n_runs, n_tracks, st);
        case 128: return launch_fused_n<128>(b, a, n_runs, n_tracks, st);
        case 256: return launch_fused_n<256>(b, a, n_runs, n_tracks, st);
        case 512: return launch_fused_n<512>(b, a, n_runs, n_tracks, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace upmix

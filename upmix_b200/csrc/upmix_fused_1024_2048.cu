// band_fused_kernel instantiations for sizes 1024, 2048 (see upmix_fused.cuh).
#include "upmix_fused.cuh"
#include "upmix_launch.h"

namespace upmix {

cudaError_t launch_band_fused_1024_2048(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st) {
    switch (b.n_fft) {
        case 1024: return launch_fused_n<1024>(b, a, n_runs, n_tracks, st);
        case 2048: return launch_fused_n<2048>(b, a, n_runs, n_tracks, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace upmix

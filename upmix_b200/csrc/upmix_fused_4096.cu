// band_fused_kernel instantiations for sizes 4096 (see upmix_fused.cuh).
#include "upmix_fused.cuh"
#include "upmix_launch.h"

namespace upmix {

cudaError_t launch_band_fused_4096(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st) {
    switch (b.n_fft) {
        case 4096: return launch_fused_n<4096>(b, a, n_runs, n_tracks, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace upmix

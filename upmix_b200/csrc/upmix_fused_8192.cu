// band_fused_kernel instantiations for sizes 8192 (see upmix_fused.cuh).
#include "upmix_fused.cuh"
#include "upmix_launch.h"

namespace upmix {

cudaError_t launch_band_fused_8192(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st) {
    switch (b.n_fft) {
        case 8192: return launch_fused_n<8192>(b, a, n_runs, n_tracks, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace upmix

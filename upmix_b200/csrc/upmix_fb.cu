// band_fb_kernel instantiations and dispatch (see upmix_fb.cuh).
#include <stdlib.h>

#include "upmix_fb.cuh"
#include "upmix_launch.h"

namespace upmix {

unsigned long long& launch_counter();

// radices of the two passes of the n_fft-point transform and of the first pass of the n_fft/2-point one (the host
// builds the twiddle tables from them); 0: size not served by the frame-batched kernel
void fb_plan(int n_fft, int* ra, int* rb, int* ha) {
    *ra = *rb = *ha = 0;
    // Measured on B200, ms per band-hour of a dense band, frame-batched / one frame per CTA: 256 points 3.71 / 5.65, 512
    // 4.11 / 5.01, 1024 4.71 (8-frame tiles, two CTAs per SM; 5.05 with 16-frame tiles, one CTA of 512 threads) / 4.60 --
    // so 1024 points keep the one-frame kernel unless UPMIX_FB_MAX_N says otherwise.
    static const int max_n = [] { const char* e = getenv("UPMIX_FB_MAX_N"); return e ? atoi(e) : 512; }();
    if (n_fft > max_n) return;
    switch (n_fft) {
        case 256: *ra = FbCfg<256>::RA; *rb = FbCfg<256>::RB; *ha = FbCfg<256>::HA; break;
        case 512: *ra = FbCfg<512>::RA; *rb = FbCfg<512>::RB; *ha = FbCfg<512>::HA; break;
        case 1024: *ra = FbCfg<1024>::RA; *rb = FbCfg<1024>::RB; *ha = FbCfg<1024>::HA; break;
        default: break;
    }
}
int fb_frames_per_tile(int n_fft) { return n_fft == 256 ? FbCfg<256>::F : n_fft == 512 ? FbCfg<512>::F : FbCfg<1024>::F; }
int fb_ctas_per_sm(int n_fft) { return n_fft == 256 ? FbCfg<256>::CTAS : n_fft == 512 ? FbCfg<512>::CTAS : FbCfg<1024>::CTAS; }

cudaError_t launch_band_fb(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st) {
    launch_counter()++;
    switch (b.n_fft) {
        case 256: return launch_fb_n<256>(b, a, n_runs, n_tracks, st);
        case 512: return launch_fb_n<512>(b, a, n_runs, n_tracks, st);
        case 1024: return launch_fb_n<1024>(b, a, n_runs, n_tracks, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace upmix

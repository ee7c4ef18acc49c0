// Launch entry points implemented in upmix_kernels.cu, called by the C-ABI layer (upmix_capi.cu).
#pragma once
#include <cuda_runtime.h>
#include "upmix_kernels.cuh"

namespace upmix {

cudaError_t launch_band_fused(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st);
int fused_smem_bytes(int n_fft);
void fused_plans(int n_fft, int* fwd, int* inv, int* half);
int row_plan(int n2);
int fused_ctas_per_sm(int n_fft);
cudaError_t launch_col_fwd(const BandDev& b, const SegArgs& a, const WaveArgs& w, int n_tracks, cudaStream_t st);
cudaError_t launch_row_mask(const BandDev& b, const WaveArgs& w, int n_tracks, cudaStream_t st);
cudaError_t launch_col_inv_ola(const BandDev& b, const SegArgs& a, const WaveArgs& w, int n_runs, int n_tracks,
                               cudaStream_t st);
cudaError_t launch_col_inv_frame(const BandDev& b, const WaveArgs& w, float* ring, float* out_c, float* out_l, float* out_r,
                                 long long out_stride, int n_tracks, cudaStream_t st);
// frame-batched kernel of dense 256 / 512 / 1024-point bands (upmix_fb.cu)
cudaError_t launch_band_fb(const BandDev& b, const SegArgs& a, int n_runs, int n_tracks, cudaStream_t st);
void fb_plan(int n_fft, int* ra, int* rb, int* ha);
int fb_ctas_per_sm(int n_fft);
int fb_frames_per_tile(int n_fft);
// decimated path (upmix_dec.cu): forward transform + mask of the wave's frames; inverse + overlap-add of a range of hops
cudaError_t launch_dec_fwd(const BandDev& b, const SegArgs& a, const DecWave& w, int n_tracks, cudaStream_t st);
cudaError_t launch_dec_inv(const BandDev& b, const SegArgs& a, const DecWave& w, int n_runs, int n_tracks, bool centre,
                           cudaStream_t st);
cudaError_t launch_band_sum(const float* ws, int n_bands, int n_tracks, long long seg_len, long long ws_seg,
                            float* out_c, float* out_l, float* out_r, long long out_stride, int mode,
                            cudaStream_t st);

// block streaming: stage = history ++ new block; history <- the last D samples of stage ([2][track][D] / [2][track][D + n_new])
cudaError_t launch_stream_stage(float* hist, float* stage, const float* in_l, const float* in_r, long long in_stride, int D,
                                int n_new, int n_tracks, cudaStream_t st);

cudaError_t launch_pcm16_to_planar(const short* in, long long n, float* l, float* r, float* partial, int n_blocks, float* peak,
                                   cudaStream_t st);
cudaError_t launch_stereo_to_pcm16(const float* in, long long n, short* out, cudaStream_t st);
cudaError_t launch_peak3(const float* c, const float* l, const float* r, long long n, float* partial, int n_blocks,
                         float* out3, cudaStream_t st);
cudaError_t launch_export_mix(const float* c, const float* l, const float* r, const float* in_l, const float* in_r,
                              long long n, float scale, int mode, float* out_a, float* out_b, float* out_c, cudaStream_t st);
cudaError_t launch_fir(const float* x, long long n, int n_tracks, long long x_stride, const float* taps, int n_taps, float* y,
                       long long y_stride, cudaStream_t st);
cudaError_t launch_fma_peak(float* out, int blocks, int iters, cudaStream_t st);
unsigned long long launch_count(bool reset);
void launch_count_add(unsigned long long n);

}  // namespace upmix

// A/B of the exchange between the two radix-32 passes of a 1024-point FFT held by ONE WARP (32 points per thread):
// (a) through shared memory (padded 32 x 33 tile: 32 STS.64 + 32 LDS.64 per thread), the way every kernel of the library
//     exchanges data between passes;
// (b) with warp shuffles only (in-register 32 x 32 transpose, five xor stages: 160 SHFL + selects per thread) -- the
//     "warp-shuffle butterflies" BASELINE.json's north_star words the FFT with.
// Both variants run the same Dft<32> butterflies and the same twiddle multiply (fft_device.cuh); the checksum proves
// they compute the same transform.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I upmix_b200/csrc -o /tmp/shuffle_probe profiles/shuffle_exchange_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "fft_device.cuh"

using namespace upmix;

template <bool SHUFFLE>
__global__ void __launch_bounds__(128) fft1024_warp(const float2* __restrict__ tw, float2* out, int iters) {
    __shared__ float2 tile[4][32 * 33];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    float2 v[32], acc = make_float2(0.f, 0.f);
    float2 w[32];
#pragma unroll
    for (int r = 0; r < 32; r++) w[r] = tw[lane * 32 + r];       // exp(-2 pi i r lane / 1024)
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 32; r++) v[r] = make_float2(__sinf(0.37f * (lane + 32 * r) + it), 0.01f * r + acc.y * 1e-9f);
        // pass 0: thread `lane` transforms points lane + 32 r (r = 0..31); output r is sub-bin r of column lane
        Dft<32, -1>::run(v);
#pragma unroll
        for (int r = 1; r < 32; r++) v[r] = cmul(v[r], w[r]);
        // exchange: thread `lane` needs output `lane` of every column c = 0..31
        if constexpr (SHUFFLE) {
#pragma unroll
            for (int s = 16; s >= 1; s >>= 1) {
                const bool up = (lane & s) != 0;
#pragma unroll
                for (int r = 0; r < 32; r++) {
                    if (r & s) continue;
                    const float2 send = up ? v[r] : v[r + s];
                    float2 got;
                    got.x = __shfl_xor_sync(0xffffffffu, send.x, s);
                    got.y = __shfl_xor_sync(0xffffffffu, send.y, s);
                    if (up) v[r] = got; else v[r + s] = got;
                }
            }
        } else {
            float2* t = tile[wid];
#pragma unroll
            for (int r = 0; r < 32; r++) t[r * 33 + lane] = v[r];
            __syncwarp();
#pragma unroll
            for (int c = 0; c < 32; c++) v[c] = t[lane * 33 + c];
            __syncwarp();
        }
        // pass 1: thread `lane` holds sub-bin `lane` of the 32 columns; output r is bin lane + 32 r
        Dft<32, -1>::run(v);
#pragma unroll
        for (int r = 0; r < 32; r++) acc = cadd(acc, cscale(v[r], 1.0f / (1 + r)));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

int main() {
    const int blocks = 148 * 4, threads = 128, iters = 2000;
    float2* tw_h = new float2[1024];
    for (int c = 0; c < 32; c++)
        for (int r = 0; r < 32; r++) {
            const double a = -2.0 * 3.14159265358979323846 * r * c / 1024.0;
            tw_h[c * 32 + r] = make_float2((float)cos(a), (float)sin(a));
        }
    float2 *tw, *out;
    cudaMalloc(&tw, 1024 * sizeof(float2));
    cudaMalloc(&out, blocks * threads * sizeof(float2));
    cudaMemcpy(tw, tw_h, 1024 * sizeof(float2), cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float2* h = new float2[blocks * threads];
    for (int mode = 0; mode < 2; mode++) {
        float best = 1e30f;
        double sum = 0;
        for (int rep = 0; rep < 4; rep++) {
            cudaEventRecord(e0);
            if (mode) fft1024_warp<true><<<blocks, threads>>>(tw, out, iters);
            else fft1024_warp<false><<<blocks, threads>>>(tw, out, iters);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep) best = ms < best ? ms : best;
        }
        cudaMemcpy(h, out, blocks * threads * sizeof(float2), cudaMemcpyDeviceToHost);
        for (int i = 0; i < blocks * threads; i++) sum += (double)h[i].x + (double)h[i].y;
        const double ffts = (double)blocks * (threads / 32) * iters;
        printf("%-28s %8.3f ms  %7.2f G points/s  %6.2f nominal TFLOP/s (5 N log2 N)  checksum %.6e  err %s\n",
               mode ? "warp shuffles (5 xor stages)" : "shared memory (32x33 tile)", best, ffts * 1024 / best / 1e6,
               ffts * 51200 / best / 1e9, sum, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}

// FP32 issue-rate probe: warp-instructions per cycle per SM for FFMA / FADD / FMUL / FADD2 / FFMA2 and a
// butterfly-like FADD+FFMA mix, 8..16 independent chains per thread, enough warps to fill the SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp32_mix_probe profiles/fp32_mix_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float b, float c) {
    float a[16];
#pragma unroll
    for (int j = 0; j < 16; j++) a[j] = threadIdx.x * 1e-3f + j;
    for (int i = 0; i < iters; i++) {
        if (MODE == 0) {
#pragma unroll
            for (int j = 0; j < 16; j++) a[j] = fmaf(a[j], b, c);
        } else if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 16; j++) a[j] = a[j] + c;
        } else if (MODE == 2) {
#pragma unroll
            for (int j = 0; j < 16; j++) a[j] = a[j] * b;
        } else if (MODE == 3) {   // packed add
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float2 r = __fadd2_rn(make_float2(a[j], a[j + 1]), make_float2(c, b));
                a[j] = r.x; a[j + 1] = r.y;
            }
        } else if (MODE == 4) {   // packed fma
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                float2 r = __ffma2_rn(make_float2(a[j], a[j + 1]), make_float2(b, b), make_float2(c, c));
                a[j] = r.x; a[j + 1] = r.y;
            }
        } else if (MODE == 5) {   // register-register adds between chains (butterfly-like): a0 +- a1
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                const float s = a[j] + a[j + 1], d = a[j] - a[j + 1];
                a[j] = s * 0.5f; a[j + 1] = d * 0.5f;
            }
        } else {                  // 3-register FFMA (no constant operand)
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
                a[j] = fmaf(a[j], a[j + 1], a[j]);
                a[j + 1] = fmaf(a[j + 1], b, a[j + 1]);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 16; j++) s += a[j];
    if (s == 123.456f) out[blockIdx.x] = s;
}
template <int MODE>
void run(const char* name, double ops_per_iter, double flops_per_iter) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float* out; cudaMalloc(&out, 4 * sms * 64);
    const int blocks = sms * 8 * 2, iters = 1 << 14;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
        cudaEventRecord(e0); probe<MODE><<<blocks, 256>>>(out, iters, 0.999f, 1e-4f); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) best = ms < best ? ms : best;
    }
    const double thr_inst = ops_per_iter * iters * 256.0 * blocks;       // thread-level instructions
    printf("%-34s %8.3f ms  %7.2f G warp-inst/s/SM  %6.2f TFLOP/s\n", name, best, thr_inst / 32 / (best * 1e-3) / sms / 1e9,
           flops_per_iter * iters * 256.0 * blocks / (best * 1e-3) / 1e12);
}
int main() {
    run<0>("FFMA (reg, const, const)", 16, 32);
    run<1>("FADD (reg + const)", 16, 16);
    run<2>("FMUL (reg * const)", 16, 16);
    run<3>("FADD2 packed", 8, 16);
    run<4>("FFMA2 packed", 8, 32);
    run<5>("butterfly FADD+FADD+2 FMUL", 32, 32);
    run<6>("FFMA 3-register", 16, 32);
    return 0;
}

// Internal: the plan object behind the C ABI (include/upmix_b200.h), shared by upmix_capi.cu (plans, scheduling)
// and upmix_host.cu (host-buffer pipeline).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <vector>

#include "upmix_kernels.cuh"

struct UpmixHostCtx;

struct UpmixPlan {
    int device = 0;
    int out_mode = 0;
    std::vector<upmix::BandDev> bands;  // one entry per pipeline: bands with identical STFTs are merged
    int n_bands_in = 0;          // bands the caller described
    void* tables = nullptr;     // one device allocation holding every table
    int max_large_n = 0;        // largest n_fft handled by the four-step path (0: none)
    int64_t halo = 0;           // input margin a time shard needs on each side
    int64_t delay = 0;          // max over bands of n_fft - hop (block streaming latency)
    int sm_count = 148;
    bool fold_in_freq = false;  // FOLD output and every pipeline fused or decimated: the centre is folded per bin
    bool use_dec = true;        // band-limited pipelines take the decimated path (upmix_dec.cu)
    // Pipelines are independent until the band sum, so they are spread over a few plan-owned streams
    // (forked from / joined to the caller's stream with events): co-resident CTAs of different pipelines
    // fill each other's stalls and launch tails.  Pipelines of the four-step path share one scratch and
    // therefore one stream.
    // aux[0]: four-step pipelines; aux[1 .. N_MAIN]: fused / decimated pipelines, round robin; aux[si + N_MAIN]: the centre's
    // inverse of the decimated pipeline on aux[si] (it only shares the masked spectra with the inverse of Ls + i Rs)
    static constexpr int N_MAIN = 4;
    static constexpr int N_AUX = 1 + 2 * N_MAIN;
    cudaStream_t aux[N_AUX] = {};
    cudaEvent_t ev_fork = nullptr;
    cudaEvent_t ev_join[N_AUX] = {};
    cudaEvent_t ev_side[N_AUX] = {};   // forward + mask of a decimated pipeline done: its centre stream may start
    bool multi_stream = false;
    UpmixHostCtx* host = nullptr;   // buffers of upmix_process_host_ex, created on first use
    // Short calls (the staged band sum: a few seconds of audio, 10-20 launches on four streams) are captured once into
    // a CUDA graph per argument set and replayed: the launch gaps are most of such a call's time.
    struct GraphEntry {
        uint64_t key[16];
        cudaGraphExec_t exec;
        int n_kernels;
        uint64_t last_use;
        // block streaming only: the graph itself (owner of the node handles), the two nodes that see the caller's
        // buffers and the pointers they currently hold (in_l, in_r, out_c, out_l, out_r)
        cudaGraph_t graph = nullptr;
        cudaGraphNode_t stage_node = nullptr, sum_node = nullptr;
        const void* io[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
    };
    std::vector<GraphEntry> graphs;
    std::vector<GraphEntry> stream_graphs;   // steady-state blocks of upmix_stream_block, keyed on the folded position
    cudaStream_t cap_stream = nullptr;
    uint64_t graph_clock = 0;
    bool use_graphs = true;
};


// sets upmix_last_error() for the calling thread and returns `code`
int upmix_fail(int code, const char* fmt, ...);
// frees the host pipeline's cached buffers (upmix_host.cu); called by upmix_plan_destroy
void upmix_host_ctx_destroy(UpmixHostCtx* ctx);

// Host-buffer entry points of the C ABI: upmix_process_host / upmix_process_host_ex.
//
// The call main.py makes (main.py:43-50, 78-80) hands over float64 strided views of an interleaved [n][2] array in
// pageable memory and expects freshly allocated float32 arrays back (center_extraction.py:503-513).  Between such
// buffers and the device this file runs a chunked pipeline:
//
//   input workers   convert / gather a chunk of L and R (float64 or float32, any stride) into a pinned staging slot and
//                   queue its host-to-device copy (one cudaMemcpyAsync per channel and chunk)
//   caller thread   waits (on the device, by events) for the chunks a time segment needs -- its samples plus the
//                   halo -- and queues upmix_process_segment for it: segments start short and double, so the first
//                   results leave early; frames keep their global index, so the result is bit-identical to one
//                   whole-track call
//   output workers  the segment's outputs come down in pieces into pinned slots; workers copy them to the caller's arrays
//
// Buffers that already are pinned (cudaHostAlloc / cudaHostRegister, e.g. torch pin_memory) are used directly, without
// staging or workers.  Device buffers, pinned rings, streams and events are cached in the plan (one host call at a
// time per plan) and freed by upmix_plan_release_host / upmix_plan_destroy.
#include "../../include/upmix_b200.h"

#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#include <vector>

#include "upmix_plan.h"

namespace {

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

constexpr int64_t FIRST_SEG = 384 * 1024;       // first segment (8 s at 48 kHz); sizes double up to ...
constexpr int64_t MAX_SEG = 4320 * 1024;        // ... about 90 s (measured in round 1: 60-120 s segments are best)

}  // namespace

struct UpmixHostCtx {
    std::mutex mu;
    int64_t cap = 0;                 // samples the device buffers hold
    int n_out = 0;
    float* d_in = nullptr;           // [2][cap]
    float* d_out = nullptr;          // [n_out][cap]
    void* ws = nullptr;
    int64_t ws_bytes = 0;
    int64_t chunk = 0;               // samples per staging slot
    int ns_in = 0, ns_out = 0;
    float* pin_in = nullptr;         // [ns_in][2][chunk]
    float* pin_out = nullptr;        // [ns_out][n_out][chunk]
    cudaStream_t s_up = nullptr, s_main = nullptr, s_down = nullptr;
    std::vector<cudaEvent_t> events;
    void free_all() {
        cudaFree(d_in);
        cudaFree(d_out);
        cudaFree(ws);
        cudaFreeHost(pin_in);
        cudaFreeHost(pin_out);
        d_in = d_out = nullptr;
        ws = nullptr;
        pin_in = pin_out = nullptr;
        cap = ws_bytes = chunk = 0;
        ns_in = ns_out = 0;
        for (cudaEvent_t e : events) cudaEventDestroy(e);
        events.clear();
        if (s_up) cudaStreamDestroy(s_up);
        if (s_main) cudaStreamDestroy(s_main);
        if (s_down) cudaStreamDestroy(s_down);
        s_up = s_main = s_down = nullptr;
    }
};

void upmix_host_ctx_destroy(UpmixHostCtx* ctx) {
    if (!ctx) return;
    ctx->free_all();
    delete ctx;
}

namespace {

bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

struct Piece { int64_t a, len; int seg; };

// spin-wait with back-off until `flag` is set or `err` is raised; returns false on error
bool wait_flag(const std::atomic<int>& flag, const std::atomic<int>& err) {
    int spins = 0;
    while (!flag.load(std::memory_order_acquire)) {
        if (err.load(std::memory_order_relaxed)) return false;
        if (++spins < 200) std::this_thread::yield();
        else std::this_thread::sleep_for(std::chrono::microseconds(20));
    }
    return true;
}

}  // namespace

// conversion / copy loops (upmix_simd.cpp: AVX2 with non-temporal stores behind a run-time CPU check)
extern "C" {
void upmix_host_pair_f64(const double* src, int64_t n, float* dl, float* dr);
void upmix_host_pair_f32(const float* src, int64_t n, float* dl, float* dr);
void upmix_host_gather_f64(const double* src, int64_t stride, int64_t n, float* dst);
void upmix_host_gather_f32(const float* src, int64_t stride, int64_t n, float* dst);
void upmix_host_copy(float* dst, const float* src, int64_t n);
}

extern "C" {

int upmix_plan_release_host(UpmixPlan* plan) {
    if (!plan) return upmix_fail(UPMIX_E_INVALID, "plan is NULL");
    if (plan->host) {
        int prev = -1;
        cudaGetDevice(&prev);
        cudaSetDevice(plan->device);
        std::lock_guard<std::mutex> lock(plan->host->mu);
        plan->host->free_all();
        if (prev >= 0) cudaSetDevice(prev);
    }
    return UPMIX_OK;
}

int upmix_process_host_ex(const UpmixPlan* cplan, const void* L, const void* R, int dtype, int64_t stride_l, int64_t stride_r,
                          int64_t n, float* out_c, float* out_l, float* out_r, int n_threads) {
    UpmixPlan* plan = const_cast<UpmixPlan*>(cplan);
    if (!plan) return upmix_fail(UPMIX_E_INVALID, "plan is NULL");
    const int n_out = plan->out_mode == UPMIX_OUT_LSCRS ? 3 : 2;
    if (!L || !R || !out_l || !out_r || (n_out == 3 && !out_c)) return upmix_fail(UPMIX_E_INVALID, "NULL buffer");
    if (dtype != UPMIX_F32 && dtype != UPMIX_F64) return upmix_fail(UPMIX_E_INVALID, "unknown dtype %d", dtype);
    if (stride_l < 1 || stride_r < 1) return upmix_fail(UPMIX_E_INVALID, "strides must be positive (elements)");
    if (n <= 0) return n == 0 ? UPMIX_OK : upmix_fail(UPMIX_E_INVALID, "negative length");
    int prev_dev = -1;
    cudaGetDevice(&prev_dev);
    if (cudaSetDevice(plan->device) != cudaSuccess) return upmix_fail(UPMIX_E_CUDA, "cannot select device %d", plan->device);
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_dev};

    if (!plan->host) {
        static std::mutex create_mu;
        std::lock_guard<std::mutex> g(create_mu);
        if (!plan->host) plan->host = new UpmixHostCtx();
    }
    UpmixHostCtx& c = *plan->host;
    std::lock_guard<std::mutex> lock(c.mu);

#define HOST_CHECK(expr)                                                                              \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) return upmix_fail(UPMIX_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

    if (n_threads <= 0) {
        const char* ev = getenv("UPMIX_HOST_THREADS");
        // default: the cores of the box divided by the ranks sharing it (torchrun's LOCAL_WORLD_SIZE), at most 16
        const char* lw = getenv("LOCAL_WORLD_SIZE");
        const unsigned ranks = lw && atoi(lw) > 0 ? (unsigned)atoi(lw) : 1u;
        n_threads = ev ? atoi(ev) : (int)std::min(16u, std::max(2u, std::thread::hardware_concurrency() / ranks));
        n_threads = std::max(1, std::min(n_threads, 32));
    }
    const bool in_direct = dtype == UPMIX_F32 && stride_l == 1 && stride_r == 1 && is_pinned(L) && is_pinned(R);
    const bool out_direct = is_pinned(out_l) && is_pinned(out_r) && (n_out == 2 || is_pinned(out_c));

    // ---- segments (multiples of the largest hop; first ones short) and the chunks / pieces that feed and drain them ----
    int64_t align = 2;
    for (const upmix::BandDev& b : plan->bands) align = std::max<int64_t>(align, b.hop);
    std::vector<int64_t> bounds(1, 0);
    {
        // UPMIX_HOST_SEG (samples): fixed segment length, for tests and sweeps
        const char* ev = getenv("UPMIX_HOST_SEG");
        const int64_t forced = ev ? std::max<int64_t>(1, atoll(ev)) : 0;
        const char* eg = getenv("UPMIX_HOST_GROWTH");              // segment i + 1 = growth x segment i, for sweeps
        const double growth = eg ? std::max(1.05, atof(eg)) : 2.0;
        const char* em = getenv("UPMIX_HOST_MAXSEG");              // (samples) longest segment, for sweeps
        const int64_t max_seg_len = em ? std::max<int64_t>(FIRST_SEG, atoll(em)) : MAX_SEG;
        const int64_t n_seg = std::max<int64_t>(1, std::min<int64_t>(64, (n + max_seg_len / 2) / max_seg_len));
        const int64_t seg = round_up(forced ? forced : (n + n_seg - 1) / n_seg, align);
        int64_t step = forced ? seg : std::min(seg, round_up(FIRST_SEG, align));
        int64_t pos = 0;
        while (pos < n) {
            pos = std::min(n, pos + step);
            bounds.push_back(pos);
            step = std::min(seg, round_up((int64_t)(growth * (double)step), align));
        }
    }
    const int n_segs = (int)bounds.size() - 1;
    int64_t max_seg = 0;
    for (int i = 0; i < n_segs; i++) max_seg = std::max(max_seg, bounds[i + 1] - bounds[i]);
    const int64_t chunk = std::min<int64_t>(1 << 20, round_up(n, 64));
    const int n_chunks = (int)((n + chunk - 1) / chunk);
    std::vector<Piece> pieces;
    for (int i = 0; i < n_segs; i++)
        for (int64_t a = bounds[i]; a < bounds[i + 1]; a += chunk) pieces.push_back({a, std::min(chunk, bounds[i + 1] - a), i});
    const int n_pieces = (int)pieces.size();

    // ---- cached buffers ----
    if (c.n_out != n_out || c.cap < n) {
        cudaFree(c.d_in);
        cudaFree(c.d_out);
        c.d_in = c.d_out = nullptr;
        c.cap = 0;
        const int64_t cap = round_up(n, 64);
        HOST_CHECK(cudaMalloc(&c.d_in, (size_t)2 * cap * sizeof(float)));
        HOST_CHECK(cudaMalloc(&c.d_out, (size_t)n_out * cap * sizeof(float)));
        c.cap = cap;
        c.n_out = n_out;
    }
    const int64_t wsb = upmix_workspace_bytes(plan, max_seg, 1);
    if (wsb < 0) return (int)wsb;
    if (c.ws_bytes < wsb) {
        cudaFree(c.ws);
        c.ws = nullptr;
        c.ws_bytes = 0;
        HOST_CHECK(cudaMalloc(&c.ws, (size_t)std::max<int64_t>(wsb, 256)));
        c.ws_bytes = wsb;
    }
    const int want_in = in_direct ? 0 : 2 * n_threads, want_out = out_direct ? 0 : 2 * n_threads;
    if (c.chunk < chunk || c.ns_in < want_in) {
        cudaFreeHost(c.pin_in);
        c.pin_in = nullptr;
        c.ns_in = 0;
        if (want_in) HOST_CHECK(cudaHostAlloc(&c.pin_in, (size_t)want_in * 2 * std::max(chunk, c.chunk) * sizeof(float), cudaHostAllocDefault));
        c.ns_in = want_in;
    }
    if (c.chunk < chunk || c.ns_out < want_out) {
        cudaFreeHost(c.pin_out);
        c.pin_out = nullptr;
        c.ns_out = 0;
        if (want_out) HOST_CHECK(cudaHostAlloc(&c.pin_out, (size_t)want_out * 3 * std::max(chunk, c.chunk) * sizeof(float), cudaHostAllocDefault));
        c.ns_out = want_out;
    }
    c.chunk = std::max(chunk, c.chunk);
    const int64_t slot_len = c.chunk;
    if (!c.s_up) {
        HOST_CHECK(cudaStreamCreateWithFlags(&c.s_up, cudaStreamNonBlocking));
        HOST_CHECK(cudaStreamCreateWithFlags(&c.s_main, cudaStreamNonBlocking));
        HOST_CHECK(cudaStreamCreateWithFlags(&c.s_down, cudaStreamNonBlocking));
    }
    // pinned float32 input goes up in place, one piece per segment: what the segment adds to the uploaded range.
    // (UPMIX_HOST_H2D_AHEAD=1 uploads in pieces of its own -- short first, doubling up to 32 MB per channel -- so that the
    // upload runs ahead at full speed whatever the segment lengths are.  Measured, same call, three times each: 43.4-44.1 ms
    // against 41.8-42.3 ms: with both directions busy each moves ~47 GB/s instead of ~54, and the download -- 1.5x the
    // bytes -- is the one that must not be slowed; an upload paced by the segments leaves it more of the link.)
    std::vector<int64_t> h2d_bounds(1, 0);
    static const bool h2d_ahead = [] { const char* e = getenv("UPMIX_HOST_H2D_AHEAD"); return e && atoi(e) != 0; }();
    if (in_direct && h2d_ahead) {
        int64_t step = std::min<int64_t>(n, FIRST_SEG + plan->halo), pos = 0;
        while (pos < n) {
            pos = std::min(n, pos + step);
            h2d_bounds.push_back(pos);
            step = std::min<int64_t>(8 << 20, 2 * step);
        }
    } else if (in_direct) {                                       // one piece per segment: what the segment adds
        for (int i = 0; i < n_segs; i++) {
            const int64_t need = std::min(n, bounds[i + 1] + plan->halo);
            if (need > h2d_bounds.back()) h2d_bounds.push_back(need);
        }
    }
    const int n_h2d = (int)h2d_bounds.size() - 1;
    const int n_up = std::max(std::max(n_chunks, n_segs), n_h2d);   // ev_up: per chunk (staged input) or per piece (pinned input)
    const size_t n_events = (size_t)n_up + n_segs + n_pieces;
    while (c.events.size() < n_events) {
        cudaEvent_t e;
        HOST_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c.events.push_back(e);
    }
    cudaEvent_t* ev_up = c.events.data();
    cudaEvent_t* ev_comp = ev_up + n_up;
    cudaEvent_t* ev_down = ev_comp + n_segs;

    std::vector<std::atomic<int>> up_rec(n_chunks), down_rec(n_pieces), out_done(n_pieces);
    for (auto& f : up_rec) f.store(0);
    for (auto& f : down_rec) f.store(0);
    for (auto& f : out_done) f.store(0);
    std::atomic<int> err(0);
    float* const d_l = c.d_in;
    float* const d_r = c.d_in + c.cap;
    float* const d_o[3] = {n_out == 3 ? c.d_out : nullptr, n_out == 3 ? c.d_out + c.cap : c.d_out,
                           n_out == 3 ? c.d_out + 2 * c.cap : c.d_out + c.cap};
    float* const h_o[3] = {out_c, out_l, out_r};
    const int ch0 = n_out == 3 ? 0 : 1;

    // ---- input side ----
    const bool paired = !in_direct && stride_l == 2 && stride_r == 2 &&
                        (const char*)R == (const char*)L + (dtype == UPMIX_F64 ? 8 : 4);
    static const bool trace = [] { const char* e = getenv("UPMIX_HOST_TRACE"); return e && atoi(e) != 0; }();
    using clk = std::chrono::steady_clock;
    auto secs = [](clk::time_point a, clk::time_point b) { return std::chrono::duration<double>(b - a).count(); };
    const auto t_start = clk::now();
    std::vector<double> t_gather(64, 0.0), t_wait(64, 0.0), t_copyout(64, 0.0);
    auto upload_worker = [&](int w, int n_workers) {
        cudaSetDevice(plan->device);
        for (int ck = w; ck < n_chunks; ck += n_workers) {
            if (err.load()) return;
            const int64_t a = (int64_t)ck * chunk, len = std::min(chunk, n - a);
            const int slot = ck % c.ns_in;
            // the slot's previous chunk (ck - ns_in) was this worker's: its copy must have left the slot
            const auto t0 = clk::now();
            if (ck >= c.ns_in && cudaEventSynchronize(ev_up[ck - c.ns_in]) != cudaSuccess) { err.store(1); return; }
            const auto t1 = clk::now();
            float* sl = c.pin_in + (int64_t)slot * 2 * slot_len;
            float* sr = sl + slot_len;
            if (dtype == UPMIX_F64) {
                if (paired) upmix_host_pair_f64((const double*)L + 2 * a, len, sl, sr);
                else {
                    upmix_host_gather_f64((const double*)L + a * stride_l, stride_l, len, sl);
                    upmix_host_gather_f64((const double*)R + a * stride_r, stride_r, len, sr);
                }
            } else {
                if (paired) upmix_host_pair_f32((const float*)L + 2 * a, len, sl, sr);
                else {
                    upmix_host_gather_f32((const float*)L + a * stride_l, stride_l, len, sl);
                    upmix_host_gather_f32((const float*)R + a * stride_r, stride_r, len, sr);
                }
            }
            const auto t2 = clk::now();
            t_wait[w & 63] += secs(t0, t1);
            t_gather[w & 63] += secs(t1, t2);
            cudaError_t e = cudaMemcpyAsync(d_l + a, sl, (size_t)len * sizeof(float), cudaMemcpyHostToDevice, c.s_up);
            if (e == cudaSuccess) e = cudaMemcpyAsync(d_r + a, sr, (size_t)len * sizeof(float), cudaMemcpyHostToDevice, c.s_up);
            if (e == cudaSuccess) e = cudaEventRecord(ev_up[ck], c.s_up);
            if (e != cudaSuccess) { err.store(1); return; }
            up_rec[ck].store(1, std::memory_order_release);
        }
    };
    // ---- output side ----
    auto download_worker = [&](int w, int n_workers) {
        cudaSetDevice(plan->device);
        for (int p = w; p < n_pieces; p += n_workers) {
            if (!wait_flag(down_rec[p], err)) return;
            if (cudaEventSynchronize(ev_down[p]) != cudaSuccess) { err.store(1); return; }
            const float* slot = c.pin_out + (int64_t)(p % c.ns_out) * 3 * slot_len;
            const auto t0 = clk::now();
            for (int ch = ch0; ch < 3; ch++)
                upmix_host_copy(h_o[ch] + pieces[p].a, slot + (int64_t)ch * slot_len, pieces[p].len);
            t_copyout[w & 63] += secs(t0, clk::now());
            out_done[p].store(1, std::memory_order_release);
        }
    };

    std::vector<std::thread> threads;
    if (in_direct) {
        // pinned float32 input: uploaded in place, one copy per channel for what each segment adds (see the segment loop)
    } else {
        const int nw = std::min(n_threads, n_chunks);
        for (int w = 0; w < nw; w++) threads.emplace_back(upload_worker, w, nw);
    }
    if (!out_direct) {
        const int nw = std::min(n_threads, n_pieces);
        for (int w = 0; w < nw; w++) threads.emplace_back(download_worker, w, nw);
    }

    // ---- caller thread: segments ----
    int rc = UPMIX_OK;
    int next_chunk = 0, next_piece = 0;
    int up_piece = 0, up_waited = 0;
    for (int i = 0; i < n_segs && rc == UPMIX_OK && !err.load(); i++) {
        const int64_t a = bounds[i], b = bounds[i + 1];
        const int64_t need = std::min(n, b + plan->halo);
        if (in_direct) {
            if (i == 0) {                                          // queue the whole upload, piece by piece
                cudaError_t e = cudaSuccess;
                for (int k = 0; k < n_h2d && e == cudaSuccess; k++) {
                    const int64_t p0 = h2d_bounds[k];
                    const size_t bytes = (size_t)(h2d_bounds[k + 1] - p0) * sizeof(float);
                    e = cudaMemcpyAsync(d_l + p0, (const float*)L + p0, bytes, cudaMemcpyHostToDevice, c.s_up);
                    if (e == cudaSuccess) e = cudaMemcpyAsync(d_r + p0, (const float*)R + p0, bytes, cudaMemcpyHostToDevice, c.s_up);
                    if (e == cudaSuccess) e = cudaEventRecord(ev_up[k], c.s_up);
                }
                if (e != cudaSuccess) { err.store(1); break; }
            }
            while (up_piece < n_h2d && h2d_bounds[up_piece] < need) up_piece++;       // pieces 0 .. up_piece - 1 cover [0, need)
            if (up_piece > up_waited) {
                if (cudaStreamWaitEvent(c.s_main, ev_up[up_piece - 1], 0) != cudaSuccess) { err.store(1); break; }
                up_waited = up_piece;
            }
        } else {
            const int last = (int)((need - 1) / chunk);
            for (; next_chunk <= last; next_chunk++) {
                if (!wait_flag(up_rec[next_chunk], err)) break;
                if (cudaStreamWaitEvent(c.s_main, ev_up[next_chunk], 0) != cudaSuccess) err.store(1);
            }
        }
        if (err.load()) break;
        // UPMIX_HOST_NOCOMPUTE=1 (measurement only): the copy pipeline without the kernels -- outputs are whatever the buffer holds
        static const bool no_compute = [] { const char* e = getenv("UPMIX_HOST_NOCOMPUTE"); return e && atoi(e) != 0; }();
        if (!no_compute)
        rc = upmix_process_segment(plan, d_l, d_r, 0, n, n, a, b, 1, c.cap, d_o[0] ? d_o[0] + a : nullptr, d_o[1] + a, d_o[2] + a,
                                   c.cap, c.ws, c.ws_bytes, c.s_main);
        if (rc != UPMIX_OK) break;
        if (cudaEventRecord(ev_comp[i], c.s_main) != cudaSuccess || cudaStreamWaitEvent(c.s_down, ev_comp[i], 0) != cudaSuccess) {
            err.store(1);
            break;
        }
        if (out_direct) {                                     // pinned outputs: one copy per channel and segment, in place
            cudaError_t e = cudaSuccess;
            for (int ch = ch0; ch < 3 && e == cudaSuccess; ch++)
                e = cudaMemcpyAsync(h_o[ch] + a, d_o[ch] + a, (size_t)(b - a) * sizeof(float), cudaMemcpyDeviceToHost, c.s_down);
            if (e != cudaSuccess) { err.store(1); break; }
            continue;
        }
        for (; next_piece < n_pieces && pieces[next_piece].seg == i; next_piece++) {
            const Piece& pc = pieces[next_piece];
            cudaError_t e = cudaSuccess;
            {
                // the slot's previous piece must have been copied out to the caller's arrays
                if (next_piece >= c.ns_out && !wait_flag(out_done[next_piece - c.ns_out], err)) break;
                float* slot = c.pin_out + (int64_t)(next_piece % c.ns_out) * 3 * slot_len;
                for (int ch = ch0; ch < 3 && e == cudaSuccess; ch++)
                    e = cudaMemcpyAsync(slot + (int64_t)ch * slot_len, d_o[ch] + pc.a, (size_t)pc.len * sizeof(float), cudaMemcpyDeviceToHost, c.s_down);
                if (e == cudaSuccess) e = cudaEventRecord(ev_down[next_piece], c.s_down);
                down_rec[next_piece].store(1, std::memory_order_release);
            }
            if (e != cudaSuccess) { err.store(1); break; }
        }
    }
    if (rc != UPMIX_OK || err.load()) err.store(1);          // release every waiter
    const auto t_launched = clk::now();
    for (std::thread& t : threads) t.join();
    const auto t_joined = clk::now();
    const cudaError_t e_up = cudaStreamSynchronize(c.s_up), e_main = cudaStreamSynchronize(c.s_main), e_down = cudaStreamSynchronize(c.s_down);
    if (trace) {
        double g = 0, wt = 0, co = 0;
        for (int i = 0; i < 64; i++) { g += t_gather[i]; wt += t_wait[i]; co += t_copyout[i]; }
        fprintf(stderr, "[upmix host] n=%lld threads=%d in_direct=%d out_direct=%d chunks=%d segs=%d: setup+launch %.1f ms, join %.1f ms, drain %.1f ms; "
                        "workers: gather %.1f ms, slot wait %.1f ms, copy-out %.1f ms (summed over threads)\n",
                (long long)n, n_threads, (int)in_direct, (int)out_direct, n_chunks, n_segs, 1e3 * secs(t_start, t_launched),
                1e3 * secs(t_launched, t_joined), 1e3 * secs(t_joined, clk::now()), 1e3 * g, 1e3 * wt, 1e3 * co);
    }
    if (rc != UPMIX_OK) return rc;
    if (err.load() || e_up != cudaSuccess || e_main != cudaSuccess || e_down != cudaSuccess) {
        const cudaError_t e = e_up != cudaSuccess ? e_up : e_main != cudaSuccess ? e_main : e_down != cudaSuccess ? e_down : cudaGetLastError();
        return upmix_fail(UPMIX_E_CUDA, "host pipeline failed: %s", cudaGetErrorString(e));
    }
    return UPMIX_OK;
#undef HOST_CHECK
}

int upmix_process_host(const UpmixPlan* plan, const float* L, const float* R, int64_t n_samples, float* out_c, float* out_l,
                       float* out_r) {
    return upmix_process_host_ex(plan, L, R, UPMIX_F32, 1, 1, n_samples, out_c, out_l, out_r, 0);
}

}  // extern "C"

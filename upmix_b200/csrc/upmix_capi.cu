// C-ABI layer (include/upmix_b200.h): plan construction, workspace layout, band scheduling.
// No torch types, no exceptions across the boundary.
#include "../../include/upmix_b200.h"

#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include "fft_device.cuh"
#include "upmix_kernels.cuh"
#include "upmix_launch.h"
#include "upmix_plan.h"

using namespace upmix;

thread_local char g_err[512] = "";

int upmix_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

namespace {

#define fail upmix_fail

#define CU_CHECK(expr)                                                                            \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return fail(UPMIX_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline bool is_pow2(int x) { return x > 0 && (x & (x - 1)) == 0; }

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Hops a column thread of col_inv_ola finishes in one run (plus 3 warm-up frames): long runs amortise the
// warm-up, short ones keep the grid full when a wave is short (time shards of the host pipeline, short tracks).
int k3_hops_per_run(int64_t wave_hops, int n_tracks) {
    static const int forced = [] { const char* e = getenv("UPMIX_K3_RUN"); return e ? std::max(4, atoi(e)) : 0; }();
    if (forced) return forced;
    const int64_t parallel = wave_hops * n_tracks;      // hops of a wave over all tracks: what fills the grid
    return parallel >= 4096 && wave_hops >= 128 ? 64 : parallel >= 256 ? 32 : 16;
}

}  // namespace

namespace {

struct Layout {
    int64_t ws_seg = 0;         // per-track stride of a band output in the workspace (floats)
    int64_t band_out_bytes = 0;
    int k3_run = 32;            // large path: hops per column-thread run of col_inv_ola
    int wave_tracks = 1;        // large path: tracks per wave (a big batch goes through in groups of tracks)
    int wave_hops = 0;          // large path: hops finished per wave
    int wave_frames = 0;        // large path: frames resident per wave (even)
    int64_t a_bytes = 0, b1_bytes = 0, b2_bytes = 0;
    int64_t dec_off = 0;        // decimated bands: their scratch regions follow the four-step scratch, one per band
    int64_t total = 0;
};

// Scratch of one decimated band: per frame the masked live bins (3 rows of KP) and, with more than one group of
// 16 sequences, the groups' partial sums (2 rows each).  A long input goes through in waves of hops (and a large
// batch in groups of tracks) so that a band's scratch stays within UPMIX_DEC_WS_MB (default 1024 MiB).
struct DecLayout {
    int wave_tracks = 1;
    int64_t wave_hops = 0;      // hops finished per wave; the wave holds wave_hops + 3 frames per track
    int64_t part_bytes = 0, spec_bytes = 0, total = 0;
};

DecLayout make_dec_layout(const BandDev& b, int64_t seg_len, int n_tracks) {
    const char* ev = getenv("UPMIX_DEC_WS_MB");               // read per call: tests force small waves with it
    const int64_t cap = (int64_t)(ev ? std::max(1, atoi(ev)) : 1024) << 20;
    DecLayout d;
    const int64_t groups = b.dec.Q / 16;
    const int64_t per_frame = ((groups > 1 ? groups * 2 : 0) + 3) * (int64_t)b.dec.KP * (int64_t)sizeof(float2);
    const int64_t seg_hops = (seg_len + b.hop - 1) / b.hop + 1;      // a segment that starts inside a hop touches one more
    const int64_t frames_cap = std::max<int64_t>(64, cap / per_frame);
    if ((seg_hops + 3) * n_tracks <= frames_cap) {
        d.wave_tracks = n_tracks;
        d.wave_hops = seg_hops;
    } else {
        d.wave_tracks = (int)std::max<int64_t>(1, std::min<int64_t>(n_tracks, frames_cap / 1024));
        d.wave_hops = std::min<int64_t>(seg_hops, std::max<int64_t>(61, frames_cap / d.wave_tracks - 3));
    }
    const int64_t frames = (d.wave_hops + 3) * d.wave_tracks;
    d.part_bytes = groups > 1 ? round_up(frames * groups * 2 * b.dec.KP * (int64_t)sizeof(float2), 256) : 0;
    d.spec_bytes = round_up(frames * 3 * b.dec.KP * (int64_t)sizeof(float2), 256);
    d.total = d.part_bytes + d.spec_bytes;
    return d;
}

// Two ways to sum the bands (center_extraction.py:503-511), same float32 additions in the same order:
//  * staged: every pipeline writes its C/Ls/Rs to a workspace slot, band_sum_kernel adds the slots.  The
//    pipelines are independent until then and run on the plan's own streams -- best for short inputs
//    (block streaming, a few seconds of audio), where launch gaps dominate.
//  * direct: the pipelines run one after the other and add their finished samples straight into the
//    caller's outputs (the first one stores).  No per-band slots (12 bytes per sample and band, written
//    and read again), no band-sum pass: 84 -> 60 bytes of HBM traffic per sample for three bands, and
//    the workspace shrinks to the four-step scratch.  Used from UPMIX_DIRECT_MIN samples up (default 3 Mi).
bool direct_sum(int64_t seg_len, int n_tracks) {
    int64_t min_samples = 3LL << 20;   // crossover measured at ~60 s of one 48 kHz track (profiles/direct_min_sweep.py)
    if (const char* ev = getenv("UPMIX_DIRECT_MIN")) min_samples = atoll(ev);
    return seg_len * (int64_t)n_tracks >= min_samples;
}

Layout make_layout(const UpmixPlan* p, int64_t seg_len, int n_tracks, bool staged, bool chunk_api = false) {
    Layout l;
    l.ws_seg = round_up(std::max<int64_t>(seg_len, 1), 64);
    l.band_out_bytes = staged ? round_up((int64_t)p->bands.size() * 3 * n_tracks * l.ws_seg * (int64_t)sizeof(float), 256) : 0;
    l.total = l.band_out_bytes;
    // four-step scratch: bands above 8192 points that do not take the decimated path (the chunk API always takes
    // the four-step path for them)
    int64_t hop_min = INT64_MAX;
    for (const BandDev& b : p->bands)
        if (b.n_fft > FUSED_MAX_N && (!b.dec.P || chunk_api)) hop_min = std::min<int64_t>(hop_min, b.hop);
    if (hop_min != INT64_MAX) {
        const int64_t seg_hops = (seg_len + hop_min - 1) / hop_min + 1;
        // hops per wave over all tracks (tuning knob: UPMIX_WAVE_HOPS).  Measured with the final kernels, ms per
        // band-hour for waves of 1024 / 2048 / 4096 / 8192 hops (runs of 64 hops per column thread): 65536 points
        // 6.65 / 6.02 / 5.65 / 5.61, 16384 points 9.69 / 7.29 / 6.26 / 6.14; the scratch of a 65536-point band is
        // 1.3 MB per hop (10.7 GB for 8192 hops, of 180 GB)
        int64_t wave_total = 8192;
        if (const char* ev = getenv("UPMIX_WAVE_HOPS")) wave_total = std::max(16, atoi(ev));
        // A wave covers `wave_tracks` tracks x `wave_hops` hops.  Waves shorter than 256 hops waste work (six extra
        // frames and three warm-up frames per run), so a large batch goes through in groups of tracks instead.
        const int64_t wt = std::max<int64_t>(1, std::min<int64_t>(n_tracks, wave_total / 256));
        l.wave_tracks = (int)wt;
        int64_t wh = std::max<int64_t>(16, wave_total / wt);
        l.k3_run = k3_hops_per_run(std::min<int64_t>(wh, seg_hops), (int)wt);
        wh = round_up(std::min<int64_t>(wh, round_up(seg_hops, l.k3_run)), l.k3_run);
        l.wave_hops = (int)wh;
        l.wave_frames = (int)wh + 6;
        const int64_t per_frame = (int64_t)p->max_large_n * (int64_t)sizeof(float2);
        l.a_bytes = round_up(wt * l.wave_frames * per_frame, 256);
        l.b1_bytes = l.a_bytes;
        l.b2_bytes = round_up(wt * (l.wave_frames / 2) * per_frame, 256);
        l.total += l.a_bytes + l.b1_bytes + l.b2_bytes;
    }
    l.dec_off = l.total;
    for (const BandDev& b : p->bands)
        if (b.dec.P) l.total += make_dec_layout(b, seg_len, n_tracks).total;
    return l;
}

// Hops per CTA of the fused kernel.  A run replays the (n_fft/hop - 1) frames before it, so long runs are
// cheaper; but the grid should come out as a whole number of waves of co-resident CTAs, or the last wave leaves
// SMs idle (measured, 8192 points, 1-hour track, run = 64 / 96 / 128 / 192 hops: 5.81 / 5.74 / 6.24 / 5.63 ms).
// The smallest number of full waves whose runs stay within `run_max` hops is taken.
int pick_hops_per_run(const UpmixPlan* p, int n_fft, int64_t total_hops, int n_tracks) {
    static const int run_max = [] { const char* e = getenv("UPMIX_RUN_MAX"); return e ? std::max(8, atoi(e)) : 192; }();
    const int64_t slots = (int64_t)p->sm_count * fused_ctas_per_sm(n_fft);
    const int64_t work = total_hops * n_tracks;
    for (int64_t waves = 1;; waves++) {
        // runs per track such that all tracks together fill `waves` waves
        const int64_t runs_per_track = std::max<int64_t>(1, waves * slots / n_tracks);
        const int64_t r = (total_hops + runs_per_track - 1) / runs_per_track;
        if (r <= run_max || waves * slots >= work) return (int)std::max<int64_t>(8, r);
    }
}

// Hops per run of dec_inv_kernel.  A run replays the 3 frames before it, and `ctas_per_run` CTAs share a run's hop range
// (the groups of 16 sequences; 0: two runs share a CTA -- the centre of a 16-sequence band).  The kernel's time is about
// (waves of co-resident CTAs) x (frames per run), so the run length that minimises that product is taken: long runs for
// an hour of audio (3 replayed frames in ~290), one or two hops per run for a few seconds of it (the GPU is not full
// anyway and what counts is the length of the serial chain per CTA).
int pick_dec_hops_per_run(const UpmixPlan* p, const BandDev& b, int64_t total_hops, int n_tracks, int ctas_per_run, int min_run = 1) {
    static const int run_max = [] { const char* e = getenv("UPMIX_DEC_RUN_MAX"); return e ? std::max(1, atoi(e)) : 320; }();
    const int64_t slots = (int64_t)p->sm_count * (1024 / b.dec.P);
    int64_t best_r = 1, best_cost = INT64_MAX;
    auto cost_of = [&](int64_t r) {
        const int64_t runs = (total_hops + r - 1) / r;
        const int64_t ctas = (ctas_per_run ? runs * ctas_per_run : (runs + 1) / 2) * n_tracks;
        return ((ctas + slots - 1) / slots) * (r + 3);
    };
    // min_run: calls whose pipelines run side by side on the plan's streams (short inputs) share the GPU, so the replayed
    // frames of one-hop runs are paid for by the other pipelines: runs of at least three hops there (measured, 10 s of the
    // default six bands: 0.164 ms with one-hop runs, 0.143 / 0.142 / 0.145 / 0.148 with at least 2 / 3 / 4 / 6)
    static const int run_min_env = [] { const char* e = getenv("UPMIX_DEC_RUN_MIN"); return e ? std::max(1, atoi(e)) : 0; }();
    const int64_t run_min = run_min_env ? run_min_env : min_run;
    for (int64_t r = std::min<int64_t>(run_min, total_hops); r <= std::min<int64_t>(run_max, total_hops); r++) {
        const int64_t c = cost_of(r);
        if (c < best_cost || (c == best_cost && r > best_r)) { best_cost = c; best_r = r; }      // ties: fewer, longer runs
    }
    return (int)best_r;
}

// One decimated band over hops [a.hop_begin, a.hop_end) of every track, in waves: forward + mask of the wave's frames,
// then the inverse / overlap-add of Ls + i Rs and of the centre.
// `side` (optional, with its event): a second stream for the centre's inverse, used when the band goes through in ONE wave
// (short inputs, where the length of the serial chain is what counts; the next wave would overwrite the spectra).
int run_dec_band(const UpmixPlan* p, const BandDev& b, const SegArgs& a, int n_tracks, char* scratch, const DecLayout& dl, cudaStream_t st,
                 cudaStream_t side = nullptr, cudaEvent_t ev_side = nullptr, bool* side_used = nullptr, int min_run = 1) {
    const int64_t h_begin = a.hop_begin, h_end = a.hop_end;
    if (dl.wave_tracks < n_tracks || h_end - h_begin > dl.wave_hops || a.fold) side = nullptr;
    if (side_used) *side_used = side != nullptr && h_end > h_begin;
    for (int t0 = 0; t0 < n_tracks; t0 += dl.wave_tracks) {
        const int nt = std::min(dl.wave_tracks, n_tracks - t0);
        SegArgs at = a;
        at.in_l += (int64_t)t0 * a.in_stride;
        at.in_r += (int64_t)t0 * a.in_stride;
        if (at.out_c) at.out_c += (int64_t)t0 * a.out_stride;
        at.out_l += (int64_t)t0 * a.out_stride;
        at.out_r += (int64_t)t0 * a.out_stride;
        for (int64_t w0 = h_begin; w0 < h_end; w0 += dl.wave_hops) {
            const int64_t w1 = std::min<int64_t>(w0 + dl.wave_hops, h_end);
            DecWave w;
            w.part = reinterpret_cast<float2*>(scratch);
            w.spec = reinterpret_cast<float2*>(scratch + dl.part_bytes);
            w.frame0 = std::max<int64_t>(0, w0 - 3);
            w.n_frames = (int)(w1 - w.frame0);
            if (w.n_frames > dl.wave_hops + 3) return fail(UPMIX_E_INVALID, "internal: decimated wave of %d frames exceeds %lld", w.n_frames, (long long)dl.wave_hops + 3);
            SegArgs aw = at;
            aw.hop_begin = w0;
            aw.hop_end = w1;
            CU_CHECK(launch_dec_fwd(b, aw, w, nt, st));
            const int groups = b.dec.Q / 16;
            if (!a.fold && side) {
                CU_CHECK(cudaEventRecord(ev_side, st));
                CU_CHECK(cudaStreamWaitEvent(side, ev_side, 0));
            }
            aw.hops_per_run = pick_dec_hops_per_run(p, b, w1 - w0, nt, groups, min_run);
            CU_CHECK(launch_dec_inv(b, aw, w, (int)((w1 - w0 + aw.hops_per_run - 1) / aw.hops_per_run), nt, false, st));
            if (!a.fold) {
                aw.hops_per_run = pick_dec_hops_per_run(p, b, w1 - w0, nt, groups / 2, min_run);
                CU_CHECK(launch_dec_inv(b, aw, w, (int)((w1 - w0 + aw.hops_per_run - 1) / aw.hops_per_run), nt, true, side ? side : st));
            }
        }
    }
    return UPMIX_OK;
}

// Runs every band of the plan over output samples [seg_begin, seg_end) into the band workspace,
// then sums the bands.  `state` (optional) = per-band overlap-add rings for block streaming.
int run_segment(const UpmixPlan* p, const float* L, const float* R, int64_t in_begin, int64_t in_end,
                int64_t n_total, int64_t seg_begin, int64_t seg_end, int n_tracks, int64_t in_stride, float* out_c,
                float* out_l, float* out_r, int64_t out_stride, void* workspace, int64_t workspace_bytes,
                float* const* band_state, cudaStream_t st, bool force_direct = false, float* const* band_state_tmp = nullptr) {
    const int64_t seg_len = seg_end - seg_begin;
    const int64_t prod_end = std::min(seg_end, n_total);      // nothing is produced past the end of the track
    const bool direct = seg_begin >= 0 && prod_end == seg_end &&
                        (force_direct || (band_state == nullptr && direct_sum(seg_len, n_tracks)));
    const Layout lay = make_layout(p, seg_len, n_tracks, !direct);
    if (workspace_bytes < lay.total)
        return fail(UPMIX_E_WORKSPACE, "workspace too small: %lld bytes given, %lld needed", (long long)workspace_bytes,
                    (long long)lay.total);
    if ((reinterpret_cast<uintptr_t>(workspace) & 255) != 0) return fail(UPMIX_E_INVALID, "workspace must be 256-byte aligned");
    float* ws = reinterpret_cast<float*>(workspace);
    char* scratch = reinterpret_cast<char*>(workspace) + lay.band_out_bytes;
    const int nb = (int)p->bands.size();
    const cudaStream_t caller = st;
    // (block streaming forks too: every band has its own ring and its own output slot)
    const bool fork = p->multi_stream && nb > 1 && !direct;
    bool first = true;
    bool used[UpmixPlan::N_AUX] = {};
    int next_fused = 1;
    int64_t dec_off = lay.dec_off;
    if (fork) CU_CHECK(cudaEventRecord(p->ev_fork, caller));

    for (int bi = 0; bi < nb; bi++) {
        const BandDev& b = p->bands[bi];
        const bool dec = b.dec.P != 0 && band_state == nullptr;
        int side_si = -1;
        const DecLayout dlay = b.dec.P ? make_dec_layout(b, seg_len, n_tracks) : DecLayout();
        char* dec_scratch = reinterpret_cast<char*>(workspace) + dec_off;
        dec_off += dlay.total;
        if (fork) {
            const bool own = b.n_fft <= FUSED_MAX_N || dec;     // four-step bands share one scratch, hence one stream
            const int si = own ? next_fused : 0;
            if (own) next_fused = next_fused == UpmixPlan::N_MAIN ? 1 : next_fused + 1;
            st = p->aux[si];
            side_si = own && dec ? si + UpmixPlan::N_MAIN : -1;
            if (!used[si]) {
                CU_CHECK(cudaStreamWaitEvent(st, p->ev_fork, 0));
                used[si] = true;
            }
        }
        SegArgs a;
        a.in_l = L;
        a.in_r = R;
        a.in_stride = in_stride;
        a.in_begin = in_begin;
        a.in_end = in_end;
        float* base = ws + (int64_t)bi * 3 * n_tracks * lay.ws_seg;
        if (direct) {
            a.out_c = out_c;
            a.out_l = out_l;
            a.out_r = out_r;
            a.out_stride = out_stride;
            // UPMIX_FORCE_ACCUM=1 (measurement only): the first pipeline accumulates too, onto whatever the outputs hold
            static const bool force_accum = [] { const char* e = getenv("UPMIX_FORCE_ACCUM"); return e && atoi(e) != 0; }();
            a.accum = first && !force_accum ? 0 : 1;
            a.mix = p->out_mode == UPMIX_OUT_FOLD && !p->fold_in_freq ? 1 : 0;
            first = false;
        } else {
            a.out_c = base;
            a.out_l = base + (int64_t)n_tracks * lay.ws_seg;
            a.out_r = base + 2 * (int64_t)n_tracks * lay.ws_seg;
            a.out_stride = lay.ws_seg;
            a.accum = 0;
            a.mix = 0;
        }
        a.out_begin = seg_begin;
        a.seg_begin = std::max<int64_t>(seg_begin, 0);
        a.seg_end = prod_end;
        a.state = band_state ? band_state[bi] : nullptr;
        a.state_out = a.state;
        a.fold = p->fold_in_freq ? 1 : 0;
        if (!direct && (seg_begin < 0 || prod_end < seg_end)) {
            // part of the requested range lies outside the track: those samples are zero
            CU_CHECK(cudaMemsetAsync(base, 0, (size_t)3 * n_tracks * lay.ws_seg * sizeof(float), st));
        }
        if (a.seg_end <= a.seg_begin) continue;
        a.hop_begin = a.seg_begin / b.hop;
        a.hop_end = (a.seg_end + b.hop - 1) / b.hop;
        const int64_t total_hops = a.hop_end - a.hop_begin;
        if (dec) {
            const bool two = fork && side_si > 0 && !a.fold;
            bool side_used = false;
            const int rc = run_dec_band(p, b, a, n_tracks, dec_scratch, dlay, st, two ? p->aux[side_si] : nullptr, two ? p->ev_side[side_si] : nullptr,
                                        &side_used, fork ? 3 : 1);
            if (rc) return rc;
            if (side_used) used[side_si] = true;
        } else if (b.fb.tw_full && !a.state && !a.mix) {
            // dense band of 256 / 512 / 1024 points: FT frames per tile; a run of r hops costs ceil((r + 3) / FT) tiles
            const int64_t slots = (int64_t)p->sm_count * fb_ctas_per_sm(b.n_fft);
            const int64_t FT = fb_frames_per_tile(b.n_fft);
            int64_t best_k = 1, best_cost = INT64_MAX;
            for (int64_t k = 1; k <= 1024 / FT; k++) {
                const int64_t r = FT * k - 3, runs = (total_hops + r - 1) / r;
                const int64_t cost = ((runs * n_tracks + slots - 1) / slots) * k;
                if (cost < best_cost || (cost == best_cost && runs * n_tracks > slots / 2)) { best_cost = cost; best_k = k; }
                if (runs == 1) break;
            }
            a.hops_per_run = (int)(FT * best_k - 3);
            const int n_runs = (int)((total_hops + a.hops_per_run - 1) / a.hops_per_run);
            CU_CHECK(launch_band_fb(b, a, n_runs, n_tracks, st));
        } else if (b.n_fft <= FUSED_MAX_N) {
            if (a.state) {
                // the ring is carried between calls.  A block that holds many hops of a small band (32 hops of a 256-point
                // band in a 2048-sample Bela block) would be ONE CTA working through them one after the other -- the longest
                // chain of the block (60 us of 99).  Such a band runs as several CTAs: run 0 loads the ring, the others
                // replay their K-1 warm-up frames from the block itself (they start K-1 hops into it at the earliest), the
                // last one saves the ring to a temporary that is copied over the state afterwards.  Same additions in the
                // same order: bit-identical to the single run.
                const int K = b.n_fft / b.hop;
                int n_runs = 1;
                a.hops_per_run = (int)total_hops;
                if (band_state_tmp && band_state_tmp[bi] && K >= 2 && total_hops >= 2 * K && n_tracks <= 16) {
                    a.hops_per_run = K - 1;
                    n_runs = (int)((total_hops + a.hops_per_run - 1) / a.hops_per_run);
                    a.state_out = band_state_tmp[bi];
                }
                CU_CHECK(launch_band_fused(b, a, n_runs, n_tracks, st));
                if (a.state_out != a.state)
                    CU_CHECK(cudaMemcpyAsync(a.state, a.state_out, (size_t)3 * b.n_fft * n_tracks * sizeof(float), cudaMemcpyDeviceToDevice, st));
            } else {
                a.hops_per_run = pick_hops_per_run(p, b.n_fft, total_hops, n_tracks);
                const int n_runs = (int)((total_hops + a.hops_per_run - 1) / a.hops_per_run);
                CU_CHECK(launch_band_fused(b, a, n_runs, n_tracks, st));
            }
        } else {
            if (a.state) return fail(UPMIX_E_UNSUPPORTED, "block streaming needs n_fft <= %d (band %d has %d)", FUSED_MAX_N, bi, b.n_fft);
            const int64_t per_frame = (int64_t)b.n_fft;
            WaveArgs w;
            w.a = reinterpret_cast<float2*>(scratch);
            w.b1 = reinterpret_cast<float2*>(scratch + lay.a_bytes);
            w.b2 = reinterpret_cast<float2*>(scratch + lay.a_bytes + lay.b1_bytes);
            (void)per_frame;
            a.hops_per_run = lay.k3_run;
            const int64_t h_begin = a.hop_begin, h_end = a.hop_end;
            for (int t0 = 0; t0 < n_tracks; t0 += lay.wave_tracks) {
                const int nt = std::min(lay.wave_tracks, n_tracks - t0);
                SegArgs at = a;                                   // this group of tracks
                at.in_l += (int64_t)t0 * a.in_stride;
                at.in_r += (int64_t)t0 * a.in_stride;
                if (at.out_c) at.out_c += (int64_t)t0 * a.out_stride;
                at.out_l += (int64_t)t0 * a.out_stride;
                at.out_r += (int64_t)t0 * a.out_stride;
                for (int64_t w0 = h_begin; w0 < h_end; w0 += lay.wave_hops) {
                    const int64_t w1 = std::min<int64_t>(w0 + lay.wave_hops, h_end);
                    const int64_t fa = std::max<int64_t>(0, w0 - 3) & ~1LL;
                    const int64_t fb = (w1 + 1) & ~1LL;
                    w.frame0 = fa;
                    w.n_frames = (int)(fb - fa);
                    if (w.n_frames > lay.wave_frames) return fail(UPMIX_E_INVALID, "internal: wave of %d frames exceeds %d", w.n_frames, lay.wave_frames);
                    SegArgs aw = at;
                    aw.hop_begin = w0;
                    aw.hop_end = w1;
                    const int n_runs = (int)((w1 - w0 + lay.k3_run - 1) / lay.k3_run);
                    CU_CHECK(launch_col_fwd(b, aw, w, nt, st));
                    CU_CHECK(launch_row_mask(b, w, nt, st));
                    CU_CHECK(launch_col_inv_ola(b, aw, w, n_runs, nt, st));
                }
            }
        }
    }
    if (fork) {
        for (int si = 0; si < UpmixPlan::N_AUX; si++)
            if (used[si]) {
                CU_CHECK(cudaEventRecord(p->ev_join[si], p->aux[si]));
                CU_CHECK(cudaStreamWaitEvent(caller, p->ev_join[si], 0));
            }
    }
    if (!direct)
        CU_CHECK(launch_band_sum(ws, nb, n_tracks, seg_len, lay.ws_seg, out_c, out_l, out_r, out_stride,
                                 p->fold_in_freq ? 2 : p->out_mode, caller));
    return UPMIX_OK;
}

int check_io(const UpmixPlan* p, const float* L, const float* R, float* out_c, float* out_l, float* out_r, int n_tracks) {
    if (!p) return fail(UPMIX_E_INVALID, "plan is NULL");
    if (!L || !R) return fail(UPMIX_E_INVALID, "input pointer is NULL");
    if (!out_l || !out_r || (p->out_mode == UPMIX_OUT_LSCRS && !out_c)) return fail(UPMIX_E_INVALID, "output pointer is NULL");
    if (n_tracks < 1 || n_tracks > 65535) return fail(UPMIX_E_INVALID, "n_tracks must be in [1, 65535], got %d", n_tracks);
    return UPMIX_OK;
}

}  // namespace

extern "C" {

const char* upmix_last_error(void) { return g_err; }

int upmix_version(void) { return 1000; }

int upmix_plan_create(int n_bands, const UpmixBandDesc* bands, int out_mode, int device, UpmixPlan** out) {
    return upmix_plan_create_ex(n_bands, bands, out_mode, device, 0, out);
}

int upmix_plan_create_ex(int n_bands, const UpmixBandDesc* bands, int out_mode, int device, int flags, UpmixPlan** out) {
    if (!out) return fail(UPMIX_E_INVALID, "out is NULL");
    *out = nullptr;
    if (n_bands < 1 || !bands) return fail(UPMIX_E_INVALID, "need at least one band");
    if (out_mode != UPMIX_OUT_LSCRS && out_mode != UPMIX_OUT_FOLD) return fail(UPMIX_E_INVALID, "unknown out_mode %d", out_mode);
    for (int i = 0; i < n_bands; i++) {
        const UpmixBandDesc& d = bands[i];
        if (!d.ana || !d.syn || !d.gain) return fail(UPMIX_E_INVALID, "band %d: NULL table", i);
        if (!is_pow2(d.n_fft) || d.n_fft < 64 || d.n_fft > LARGE_MAX_N)
            return fail(UPMIX_E_UNSUPPORTED, "band %d: n_fft=%d is not a power of two in [64, %d]", i, d.n_fft, LARGE_MAX_N);
        if (d.hop < 2 || d.n_fft % d.hop != 0 || (d.hop & 1))
            return fail(UPMIX_E_UNSUPPORTED, "band %d: hop=%d must be even and divide n_fft=%d", i, d.hop, d.n_fft);
        if (d.n_fft > FUSED_MAX_N && d.hop * 4 != d.n_fft && d.hop * 2 != d.n_fft)
            return fail(UPMIX_E_UNSUPPORTED, "band %d: n_fft=%d > %d requires hop = n_fft/4 or n_fft/2 (75%% / 50%% overlap), got %d", i, d.n_fft,
                        FUSED_MAX_N, d.hop);
    }
    // hop the kernels step by / frame step (BandDev::hop): 50 % overlap above 8192 points runs as 75 % with every other frame absent
    auto khop = [](const UpmixBandDesc& d) { return d.n_fft > FUSED_MAX_N && d.hop * 2 == d.n_fft ? d.n_fft / 4 : d.hop; };
    auto kstep = [](const UpmixBandDesc& d) { return d.n_fft > FUSED_MAX_N && d.hop * 2 == d.n_fft ? 2 : 1; };
    // Bands whose STFT is identical (size, hop, both windows) run as ONE pipeline: they share the forward
    // transform, get their own mask per bin, and -- inverse transform and overlap-add being linear -- share
    // the inverse as well.  groups[g] = indices of the bands of pipeline g, in list order.
    std::vector<std::vector<int>> groups;
    for (int i = 0; i < n_bands; i++) {
        bool placed = false;
        for (auto& grp : groups) {
            const UpmixBandDesc& h = bands[grp[0]];
            const UpmixBandDesc& d = bands[i];
            if (h.n_fft == d.n_fft && h.hop == d.hop && memcmp(h.ana, d.ana, sizeof(float) * d.n_fft) == 0 &&
                memcmp(h.syn, d.syn, sizeof(float) * d.n_fft) == 0) {
                grp.push_back(i);
                placed = true;
                break;
            }
        }
        if (!placed) groups.push_back(std::vector<int>(1, i));
    }
    // Decimated path (upmix_dec.cu): a pipeline whose live bins all lie below P = 128 / 256 / 512 <= n_fft/16
    // (75 % overlap) splits its frames into n_fft/P sequences of P points.  dec_p[g] = P, or 0.
    bool use_dec = !(flags & UPMIX_PLAN_NO_DECIMATE);
    if (const char* ev = getenv("UPMIX_DEC")) use_dec = use_dec && atoi(ev) != 0;
    bool use_fb = !(flags & UPMIX_PLAN_NO_BATCH);
    if (const char* ev = getenv("UPMIX_FB")) use_fb = use_fb && atoi(ev) != 0;
    std::vector<int> dec_p(groups.size(), 0), dec_k(groups.size(), 0);
    bool any_four_step = false;
    for (size_t gi = 0; gi < groups.size(); gi++) {
        const UpmixBandDesc& d = bands[groups[gi][0]];
        int top = 0;
        for (int m : groups[gi])
            for (int k = d.n_fft / 2; k > top; k--)
                if (bands[m].gain[k] != 0.f) { top = k; break; }
        dec_k[gi] = top;
        const int P = top < 128 ? 128 : top < 256 ? 256 : top < 512 ? 512 : 0;
        if (use_dec && P && khop(d) * 4 == d.n_fft && d.n_fft >= 16 * P) dec_p[gi] = P;
        if (d.n_fft > FUSED_MAX_N && !dec_p[gi]) any_four_step = true;
    }
    // the fold-down epilogue of a plan with four-step bands is applied by the band kernels' copy-out, which the
    // decimated kernels do not have: such plans keep the older paths
    if (out_mode == UPMIX_OUT_FOLD && any_four_step) {
        std::fill(dec_p.begin(), dec_p.end(), 0);
        use_fb = false;      // (the frame-batched kernel has no fold-down epilogue either: staged and direct sums must pick the same kernels)
    }
    int64_t floats = 0;
    for (size_t gi = 0; gi < groups.size(); gi++) {
        const auto& grp = groups[gi];
        const UpmixBandDesc& d = bands[grp[0]];
        floats += round_up(d.n_fft + 1, 64) + round_up(d.n_fft, 64) + (int64_t)grp.size() * round_up(d.n_fft / 2 + 1, 64);
        if (dec_p[gi]) {
            const int64_t kp = round_up(dec_k[gi] + 1, 32);
            floats += 2 * (dec_p[gi] / 16) * 16 + 2 * kp + 2 * (d.n_fft / dec_p[gi] / 16) * kp;
        }
        {
            int ra = 0, rb = 0, ha = 0;
            fb_plan(d.n_fft, &ra, &rb, &ha);
            if (use_fb && ra && !dec_p[gi] && d.hop * 4 == d.n_fft) floats += round_up(2 * ra * rb, 64) + round_up(2 * ha * 16, 64);
        }
        if (d.n_fft > FUSED_MAX_N) {
            floats += 2LL * d.n_fft + round_up(2LL * fft_tw_size(row_plan(d.n_fft / COL_R)), 64);
        } else {
            int pf = 0, pi = 0, ph = 0;
            fused_plans(d.n_fft, &pf, &pi, &ph);
            floats += round_up(2LL * fft_tw_size(pf), 64) + round_up(2LL * fft_tw_size(pi), 64) +
                      round_up(2LL * fft_tw_size(ph), 64) + round_up(2LL * (d.n_fft / 2), 64);
        }
    }
    DeviceGuard guard(device);
    if (!guard.ok) return fail(UPMIX_E_CUDA, "cannot select device %d", device);
    UpmixPlan* p = new (std::nothrow) UpmixPlan();
    if (!p) return fail(UPMIX_E_INVALID, "out of host memory");
    p->device = device;
    p->out_mode = out_mode;
    p->n_bands_in = n_bands;
    int sms = 0;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sms > 0) p->sm_count = sms;
    std::vector<float> host((size_t)floats, 0.f);
    cudaError_t e = cudaMalloc(&p->tables, (size_t)floats * sizeof(float));
    if (e != cudaSuccess) {
        delete p;
        return fail(UPMIX_E_CUDA, "cudaMalloc(%lld) failed: %s", (long long)(floats * 4), cudaGetErrorString(e));
    }
    float* dbase = reinterpret_cast<float*>(p->tables);
    int64_t off = 0;
    // per-pass FFT twiddles of an n-point transform (layout: fft_device.cuh), double -> float32
    auto gen_fft_tw = [&](int plan, float* dst) {
        for (int pass = 1; pass < fft_num_passes(plan); pass++) {
            const int R = fft_radix(plan, pass), NS = fft_ns(plan, pass), o = fft_tw_offset(plan, pass);
            for (int r = 1; r < R; r++)
                for (int k = 0; k < NS; k++) {
                    const double ang = -2.0 * M_PI * (double)(((int64_t)r * k) % ((int64_t)NS * R)) / ((double)NS * R);
                    dst[2 * (o + (r - 1) * NS + k)] = (float)cos(ang);
                    dst[2 * (o + (r - 1) * NS + k) + 1] = (float)sin(ang);
                }
        }
    };
    p->use_dec = use_dec;
    bool four_step = false;
    for (size_t gi = 0; gi < groups.size(); gi++) {
        const auto& grp = groups[gi];
        const UpmixBandDesc& d = bands[grp[0]];
        BandDev b;
        memset(&b.dec, 0, sizeof(b.dec));
        if (dec_p[gi]) {
            // tables of the decimated path, computed in double: second-pass twiddles of the P-point transform
            // [k][r] = exp(-2 pi i r k / P); W^k and W^{16 g k}, W = exp(-2 pi i / n_fft), for the live bins
            const int P = dec_p[gi], Q = d.n_fft / P, R0 = P / 16;
            const int K = dec_k[gi], KP = (int)round_up(K + 1, 32);
            b.dec.P = P;
            b.dec.Q = Q;
            b.dec.K = K;
            b.dec.KP = KP;
            auto put = [&](int64_t at, double turns) {
                host[at] = (float)cos(-2.0 * M_PI * turns);
                host[at + 1] = (float)sin(-2.0 * M_PI * turns);
            };
            for (int k = 0; k < R0; k++)
                for (int r = 0; r < 16; r++) put(off + 2 * (k * 16 + r), (double)((r * k) % P) / P);
            b.dec.tw_last = reinterpret_cast<const float2*>(dbase + off);
            off += 2 * R0 * 16;
            for (int k = 0; k < KP; k++) put(off + 2 * k, (double)k / d.n_fft);
            b.dec.tw_step = reinterpret_cast<const float2*>(dbase + off);
            off += 2 * KP;
            for (int g = 0; g < Q / 16; g++)
                for (int k = 0; k < KP; k++) put(off + 2 * ((int64_t)g * KP + k), (double)(((int64_t)16 * g * k) % d.n_fft) / d.n_fft);
            b.dec.tw_base = reinterpret_cast<const float2*>(dbase + off);
            off += 2 * (int64_t)(Q / 16) * KP;
        }
        if (d.n_fft > FUSED_MAX_N && !dec_p[gi]) four_step = true;
        b.fb.tw_full = b.fb.tw_half = nullptr;
        {
            int ra = 0, rb = 0, ha = 0;
            fb_plan(d.n_fft, &ra, &rb, &ha);
            if (use_fb && ra && !dec_p[gi] && d.hop * 4 == d.n_fft) {
                // frame-batched kernel (upmix_fb.cuh): second-pass twiddles [k][r] of the n_fft- and n_fft/2-point transforms
                for (int k = 0; k < ra; k++)
                    for (int r = 0; r < rb; r++) {
                        const double ang = -2.0 * M_PI * (double)((r * k) % d.n_fft) / d.n_fft;
                        host[off + 2 * (k * rb + r)] = (float)cos(ang);
                        host[off + 2 * (k * rb + r) + 1] = (float)sin(ang);
                    }
                b.fb.tw_full = reinterpret_cast<const float2*>(dbase + off);
                off += round_up(2 * ra * rb, 64);
                const int mh = d.n_fft / 2;
                for (int k = 0; k < ha; k++)
                    for (int r = 0; r < 16; r++) {
                        const double ang = -2.0 * M_PI * (double)((r * k) % mh) / mh;
                        host[off + 2 * (k * 16 + r)] = (float)cos(ang);
                        host[off + 2 * (k * 16 + r) + 1] = (float)sin(ang);
                    }
                b.fb.tw_half = reinterpret_cast<const float2*>(dbase + off);
                off += round_up(2 * ha * 16, 64);
            }
        }
        b.n_fft = d.n_fft;
        b.hop = khop(d);
        b.frame_step = kstep(d);
        b.tw_fft = b.tw_inv = b.tw_half = b.tw_pack = b.tw_col = nullptr;
        memcpy(&host[off], d.ana, sizeof(float) * d.n_fft);      // ana[n_fft] = 0 follows (host is zero-filled)
        b.ana = dbase + off;
        off += round_up(d.n_fft + 1, 64);
        const float inv_n = 1.0f / (float)d.n_fft;   // exact: n_fft is a power of two
        for (int n = 0; n < d.n_fft; n++) host[off + n] = d.syn[n] * inv_n;
        b.syn = dbase + off;
        off += round_up(d.n_fft, 64);
        {   // gain tables of the merged bands; per bin the non-zero gains first (band order kept)
            const int nb = d.n_fft / 2 + 1;
            const int64_t gs = round_up(nb, 64);
            const int ng = (int)grp.size();
            b.max_bin = -1;
            for (int k = 0; k < nb; k++) {
                int q = 0;
                for (int m = 0; m < ng; m++) {
                    const float g = bands[grp[m]].gain[k];
                    if (g != 0.f) host[off + q++ * gs + k] = g;
                }
                if (q) b.max_bin = k;
            }
            b.gain = dbase + off;
            b.n_gains = ng;
            b.gain_stride = (int)gs;
            off += ng * gs;
        }
        if (d.n_fft <= FUSED_MAX_N) {
            int pf = 0, pi = 0, ph = 0;
            fused_plans(d.n_fft, &pf, &pi, &ph);
            gen_fft_tw(pf, &host[off]);
            b.tw_fft = reinterpret_cast<const float2*>(dbase + off);
            off += round_up(2LL * fft_tw_size(pf), 64);
            gen_fft_tw(pi, &host[off]);
            b.tw_inv = reinterpret_cast<const float2*>(dbase + off);
            off += round_up(2LL * fft_tw_size(pi), 64);
            gen_fft_tw(ph, &host[off]);
            b.tw_half = reinterpret_cast<const float2*>(dbase + off);
            off += round_up(2LL * fft_tw_size(ph), 64);
            for (int k = 0; k < d.n_fft / 2; k++) {
                const double ang = -2.0 * M_PI * (double)k / (double)d.n_fft;
                host[off + 2 * k] = (float)cos(ang);
                host[off + 2 * k + 1] = (float)sin(ang);
            }
            b.tw_pack = reinterpret_cast<const float2*>(dbase + off);
            off += round_up(2LL * (d.n_fft / 2), 64);
        } else {
            const int n2 = d.n_fft / COL_R;
            gen_fft_tw(row_plan(n2), &host[off]);
            b.tw_fft = reinterpret_cast<const float2*>(dbase + off);
            off += round_up(2LL * fft_tw_size(row_plan(n2)), 64);
            for (int k1 = 0; k1 < COL_R; k1++)
                for (int c = 0; c < n2; c++) {
                    const int64_t m = ((int64_t)k1 * c) % d.n_fft;
                    const double ang = -2.0 * M_PI * (double)m / (double)d.n_fft;
                    host[off + 2 * ((int64_t)k1 * n2 + c)] = (float)cos(ang);
                    host[off + 2 * ((int64_t)k1 * n2 + c) + 1] = (float)sin(ang);
                }
            b.tw_col = reinterpret_cast<const float2*>(dbase + off);
            off += 2LL * d.n_fft;
            p->max_large_n = std::max(p->max_large_n, d.n_fft);
        }
        // Margin a time shard needs around [seg_begin, seg_end): the first hop may start up to hop-1
        // samples before seg_begin and needs the 3 frames before it -> n_fft; large bands transform the
        // centre of frames 2p and 2p+1 in one complex FFT, so a shard must also see the whole partner
        // frame -> one more hop.
        p->halo = std::max<int64_t>(p->halo, d.n_fft > FUSED_MAX_N ? d.n_fft + d.hop : d.n_fft);
        p->delay = std::max<int64_t>(p->delay, d.n_fft - d.hop);
        p->bands.push_back(b);
    }
    p->fold_in_freq = out_mode == UPMIX_OUT_FOLD && !four_step;
    {
        const char* ev = getenv("UPMIX_STREAMS");
        p->multi_stream = !(ev && atoi(ev) == 0);
        bool ok = cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming) == cudaSuccess;
        for (int si = 0; si < UpmixPlan::N_AUX && ok; si++)
            ok = cudaStreamCreateWithFlags(&p->aux[si], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p->ev_join[si], cudaEventDisableTiming) == cudaSuccess &&
                 cudaEventCreateWithFlags(&p->ev_side[si], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) p->multi_stream = false;
        const char* eg = getenv("UPMIX_GRAPHS");
        p->use_graphs = !(eg && atoi(eg) == 0);
        if (p->use_graphs && cudaStreamCreateWithFlags(&p->cap_stream, cudaStreamNonBlocking) != cudaSuccess) p->use_graphs = false;
    }
    e = cudaMemcpy(p->tables, host.data(), (size_t)floats * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        cudaFree(p->tables);
        delete p;
        return fail(UPMIX_E_CUDA, "table upload failed: %s", cudaGetErrorString(e));
    }
    *out = p;
    return UPMIX_OK;
}

int upmix_plan_destroy(UpmixPlan* plan) {
    if (!plan) return UPMIX_OK;
    DeviceGuard guard(plan->device);
    for (int si = 0; si < UpmixPlan::N_AUX; si++) {
        if (plan->aux[si]) cudaStreamDestroy(plan->aux[si]);
        if (plan->ev_join[si]) cudaEventDestroy(plan->ev_join[si]);
        if (plan->ev_side[si]) cudaEventDestroy(plan->ev_side[si]);
    }
    if (plan->ev_fork) cudaEventDestroy(plan->ev_fork);
    for (UpmixPlan::GraphEntry& g : plan->graphs) cudaGraphExecDestroy(g.exec);
    for (UpmixPlan::GraphEntry& g : plan->stream_graphs) {
        cudaGraphExecDestroy(g.exec);
        cudaGraphDestroy(g.graph);
    }
    if (plan->cap_stream) cudaStreamDestroy(plan->cap_stream);
    upmix_host_ctx_destroy(plan->host);
    cudaFree(plan->tables);
    delete plan;
    return UPMIX_OK;
}

int upmix_plan_n_bands(const UpmixPlan* plan) { return plan ? plan->n_bands_in : fail(UPMIX_E_INVALID, "plan is NULL"); }

int upmix_plan_n_pipelines(const UpmixPlan* plan) { return plan ? (int)plan->bands.size() : fail(UPMIX_E_INVALID, "plan is NULL"); }

int64_t upmix_workspace_bytes(const UpmixPlan* plan, int64_t seg_len, int n_tracks) {
    if (!plan) return fail(UPMIX_E_INVALID, "plan is NULL");
    if (seg_len < 0 || n_tracks < 1) return fail(UPMIX_E_INVALID, "bad seg_len / n_tracks");
    // Enough for this segment length AND any shorter one (callers size one workspace for a run of segments
    // whose last one is shorter): long segments sum directly and need only the four-step scratch, which
    // grows with the length; segments below the direct-sum threshold stage their bands, so the staged
    // layout of the longest such segment is covered as well.
    // upmix_frame_step on a band above 8192 points always takes the four-step path: cover its scratch as well
    int64_t chunk_api = 0;
    if (plan->bands.size() == 1 && plan->bands[0].n_fft > FUSED_MAX_N && plan->bands[0].dec.P && seg_len <= plan->bands[0].hop)
        chunk_api = make_layout(plan, seg_len, n_tracks, false, true).total;
    if (!direct_sum(seg_len, n_tracks)) return std::max(chunk_api, make_layout(plan, seg_len, n_tracks, true).total);
    int64_t lo = 0, hi = seg_len;                      // largest staged length: direct_sum is monotonic in seg_len
    while (lo < hi) {
        const int64_t mid = lo + (hi - lo + 1) / 2;
        if (direct_sum(mid, n_tracks)) hi = mid - 1; else lo = mid;
    }
    return std::max(chunk_api, std::max(make_layout(plan, seg_len, n_tracks, false).total, make_layout(plan, lo, n_tracks, true).total));
}

int64_t upmix_segment_halo(const UpmixPlan* plan) { return plan ? plan->halo : fail(UPMIX_E_INVALID, "plan is NULL"); }

int upmix_process_segment(const UpmixPlan* plan, const float* L, const float* R, int64_t in_begin, int64_t in_len,
                          int64_t n_total, int64_t seg_begin, int64_t seg_end, int n_tracks, int64_t in_stride,
                          float* out_c, float* out_l, float* out_r, int64_t out_stride, void* workspace,
                          int64_t workspace_bytes, void* stream) {
    int rc = check_io(plan, L, R, out_c, out_l, out_r, n_tracks);
    if (rc) return rc;
    if (n_total < 0 || in_begin < 0 || in_len < 0 || in_begin + in_len > n_total)
        return fail(UPMIX_E_INVALID, "input range [%lld, %lld) is not inside the track [0, %lld)", (long long)in_begin,
                    (long long)(in_begin + in_len), (long long)n_total);
    if (seg_begin < 0 || seg_end < seg_begin || seg_end > n_total)
        return fail(UPMIX_E_INVALID, "segment [%lld, %lld) is not inside the track [0, %lld)", (long long)seg_begin,
                    (long long)seg_end, (long long)n_total);
    if (seg_end == seg_begin) return UPMIX_OK;
    const int64_t need_lo = std::max<int64_t>(0, seg_begin - plan->halo), need_hi = std::min(n_total, seg_end + plan->halo);
    if (in_begin > need_lo || in_begin + in_len < need_hi)
        return fail(UPMIX_E_INVALID, "input range [%lld, %lld) does not cover the halo'd segment [%lld, %lld)", (long long)in_begin,
                    (long long)(in_begin + in_len), (long long)need_lo, (long long)need_hi);
    DeviceGuard guard(plan->device);
    if (!guard.ok) return fail(UPMIX_E_CUDA, "cannot select device %d", plan->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    // short calls: replay (or capture) a CUDA graph of exactly this call
    UpmixPlan* mp = const_cast<UpmixPlan*>(plan);
    if (mp->use_graphs && mp->cap_stream && !direct_sum(seg_end - seg_begin, n_tracks)) {
        uint64_t key[16] = {(uint64_t)(uintptr_t)L, (uint64_t)(uintptr_t)R, (uint64_t)in_begin, (uint64_t)in_len, (uint64_t)n_total,
                            (uint64_t)seg_begin, (uint64_t)seg_end, (uint64_t)n_tracks, (uint64_t)in_stride, (uint64_t)(uintptr_t)out_c,
                            (uint64_t)(uintptr_t)out_l, (uint64_t)(uintptr_t)out_r, (uint64_t)out_stride, (uint64_t)(uintptr_t)workspace,
                            (uint64_t)workspace_bytes, 0};
        if (const char* ev = getenv("UPMIX_DIRECT_MIN")) key[15] = (uint64_t)atoll(ev);      // (tests flip the band-sum mode)
        for (UpmixPlan::GraphEntry& g : mp->graphs)
            if (memcmp(g.key, key, sizeof(key)) == 0) {
                g.last_use = ++mp->graph_clock;
                CU_CHECK(cudaGraphLaunch(g.exec, st));
                launch_count_add(g.n_kernels);
                return UPMIX_OK;
            }
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        const bool caller_capturing = st && cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
        if (!caller_capturing && cudaStreamBeginCapture(mp->cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const unsigned long long before = launch_count(false);
            const int rc2 = run_segment(plan, L, R, in_begin, in_begin + in_len, n_total, seg_begin, seg_end, n_tracks, in_stride, out_c,
                                        out_l, out_r, out_stride, workspace, workspace_bytes, nullptr, mp->cap_stream);
            cudaGraph_t graph = nullptr;
            const cudaError_t ee = cudaStreamEndCapture(mp->cap_stream, &graph);
            cudaGraphExec_t exec = nullptr;
            if (rc2 == UPMIX_OK && ee == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                cudaGraphDestroy(graph);
                UpmixPlan::GraphEntry g;
                memcpy(g.key, key, sizeof(key));
                g.exec = exec;
                g.n_kernels = (int)(launch_count(false) - before);
                g.last_use = ++mp->graph_clock;
                if (mp->graphs.size() >= 16) {                  // evict the least recently used
                    size_t lru = 0;
                    for (size_t i = 1; i < mp->graphs.size(); i++)
                        if (mp->graphs[i].last_use < mp->graphs[lru].last_use) lru = i;
                    cudaGraphExecDestroy(mp->graphs[lru].exec);
                    mp->graphs[lru] = g;
                } else {
                    mp->graphs.push_back(g);
                }
                CU_CHECK(cudaGraphLaunch(exec, st));
                return UPMIX_OK;
            }
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();                                 // capture failed: run the call the plain way
            if (rc2 != UPMIX_OK) return rc2;
        }
    }
    return run_segment(plan, L, R, in_begin, in_begin + in_len, n_total, seg_begin, seg_end, n_tracks, in_stride, out_c, out_l,
                       out_r, out_stride, workspace, workspace_bytes, nullptr, st);
}

int upmix_process(const UpmixPlan* plan, const float* L, const float* R, int64_t n_samples, int n_tracks,
                  int64_t in_stride, float* out_c, float* out_l, float* out_r, int64_t out_stride, void* workspace,
                  int64_t workspace_bytes, void* stream) {
    if (n_samples == 0) return check_io(plan, L, R, out_c, out_l, out_r, n_tracks);
    return upmix_process_segment(plan, L, R, 0, n_samples, n_samples, 0, n_samples, n_tracks, in_stride, out_c, out_l, out_r,
                                 out_stride, workspace, workspace_bytes, stream);
}

// ---- block streaming ---------------------------------------------------------------------------
// state layout (floats): [track][2][delay] input history, then per band [track][3][n_fft] rings.
int64_t upmix_stream_delay(const UpmixPlan* plan) { return plan ? plan->delay : fail(UPMIX_E_INVALID, "plan is NULL"); }

int64_t upmix_stream_state_bytes(const UpmixPlan* plan, int n_tracks) {
    if (!plan || n_tracks < 1) return fail(UPMIX_E_INVALID, "bad plan / n_tracks");
    int64_t floats = round_up(2 * plan->delay * n_tracks, 64);
    for (const BandDev& b : plan->bands) floats += round_up(3LL * b.n_fft * n_tracks, 64);
    return floats * (int64_t)sizeof(float);
}

int upmix_stream_reset(const UpmixPlan* plan, void* state, int n_tracks, void* stream) {
    if (!plan || !state) return fail(UPMIX_E_INVALID, "plan / state is NULL");
    DeviceGuard guard(plan->device);
    CU_CHECK(cudaMemsetAsync(state, 0, (size_t)upmix_stream_state_bytes(plan, n_tracks), reinterpret_cast<cudaStream_t>(stream)));
    return UPMIX_OK;
}

}  // extern "C"

namespace {

// bytes of the temporary rings of split bands (see run_segment), one per band
int64_t stream_tmp_ring_bytes(const UpmixPlan* plan, int n_tracks) {
    int64_t floats = 0;
    for (const BandDev& b : plan->bands) floats += round_up(3LL * b.n_fft * n_tracks, 64);
    return round_up(floats * (int64_t)sizeof(float), 256);
}

// parameter lists of stream_stage_kernel / band_sum_kernel (upmix_kernels.cu), whose graph nodes get new I/O pointers per block
constexpr int STREAM_STAGE_NARGS = 8, STREAM_STAGE_IN_L = 2;
constexpr int BAND_SUM_NARGS = 10, BAND_SUM_OUT_C = 5;

struct StreamIo {
    const float* in_l;
    const float* in_r;
    float* out_c;
    float* out_l;
    float* out_r;
};

// One block: stage = history ++ new block, every band over the block's output range (rings carried), band sum.
// Under stream capture `stage_node` / `sum_node` receive the graph nodes of the two kernels that touch the caller's
// input and output buffers (the only per-call pointers).
int stream_block_body(const UpmixPlan* plan, void* state, int64_t pos, const StreamIo& io, int n_new, int n_tracks, int64_t in_stride,
                      int64_t out_stride, void* workspace, int64_t workspace_bytes, cudaStream_t st, cudaGraphNode_t* stage_node,
                      cudaGraphNode_t* sum_node) {
    const int64_t D = plan->delay;
    const int64_t span = D + n_new;
    float* hist = reinterpret_cast<float*>(state);
    float* stage = reinterpret_cast<float*>(workspace);       // [2][track][span]
    const int64_t stage_bytes = round_up(2LL * n_tracks * span * (int64_t)sizeof(float), 256);
    const int64_t tmp_bytes = stream_tmp_ring_bytes(plan, n_tracks);
    auto last_node = [&](cudaGraphNode_t* out) {
        if (!out) return;
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        const cudaGraphNode_t* deps = nullptr;
        size_t nd = 0;
        *out = nullptr;
        if (cudaStreamGetCaptureInfo(st, &cs, nullptr, nullptr, &deps, &nd) == cudaSuccess && cs == cudaStreamCaptureStatusActive && nd == 1)
            *out = deps[0];
    };
    if (n_new <= 65536 && D <= 65536) {
        CU_CHECK(launch_stream_stage(hist, stage, io.in_l, io.in_r, in_stride, (int)D, n_new, n_tracks, st));
        last_node(stage_node);
    } else {
        // stage = history ++ new block (per channel, per track); history <- last D samples of stage
        for (int ch = 0; ch < 2; ch++) {
            const float* src = ch ? io.in_r : io.in_l;
            float* dst = stage + (int64_t)ch * n_tracks * span;
            CU_CHECK(cudaMemcpy2DAsync(dst, span * sizeof(float), hist + (int64_t)ch * n_tracks * D, D * sizeof(float),
                                       D * sizeof(float), n_tracks, cudaMemcpyDeviceToDevice, st));
            CU_CHECK(cudaMemcpy2DAsync(dst + D, span * sizeof(float), src, in_stride * sizeof(float), (size_t)n_new * sizeof(float),
                                       n_tracks, cudaMemcpyDeviceToDevice, st));
            CU_CHECK(cudaMemcpy2DAsync(hist + (int64_t)ch * n_tracks * D, D * sizeof(float), dst + n_new, span * sizeof(float),
                                       D * sizeof(float), n_tracks, cudaMemcpyDeviceToDevice, st));
        }
    }
    std::vector<float*> rings, tmps;
    float* rp = hist + round_up(2 * D * n_tracks, 64);
    float* tp = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + stage_bytes);
    for (const BandDev& b : plan->bands) {
        rings.push_back(rp);
        tmps.push_back(tp);
        rp += round_up(3LL * b.n_fft * n_tracks, 64);
        tp += round_up(3LL * b.n_fft * n_tracks, 64);
    }
    const int64_t seg_begin = pos - D, seg_end = pos + n_new - D;
    const int rc = run_segment(plan, stage, stage + (int64_t)n_tracks * span, seg_begin, pos + n_new, INT64_MAX / 4, seg_begin, seg_end,
                               n_tracks, span, io.out_c, io.out_l, io.out_r, out_stride, reinterpret_cast<char*>(workspace) + stage_bytes + tmp_bytes,
                               workspace_bytes - stage_bytes - tmp_bytes, rings.data(), st, false, tmps.data());
    if (rc == UPMIX_OK) last_node(sum_node);
    return rc;
}

}  // namespace

extern "C" {

int64_t upmix_stream_workspace_bytes(const UpmixPlan* plan, int n_new, int n_tracks) {
    if (!plan || n_new < 1 || n_tracks < 1) return fail(UPMIX_E_INVALID, "bad arguments");
    const int64_t stage = round_up(2LL * n_tracks * (plan->delay + n_new) * (int64_t)sizeof(float), 256);
    return stage + stream_tmp_ring_bytes(plan, n_tracks) + make_layout(plan, n_new, n_tracks, true).total;
}

int upmix_stream_block(const UpmixPlan* plan, void* state, int64_t samples_done, const float* in_l, const float* in_r,
                       int n_new, int n_tracks, int64_t in_stride, float* out_c, float* out_l, float* out_r,
                       int64_t out_stride, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_io(plan, in_l, in_r, out_c, out_l, out_r, n_tracks);
    if (rc) return rc;
    if (!state) return fail(UPMIX_E_INVALID, "state is NULL");
    if (n_new < 1 || samples_done < 0) return fail(UPMIX_E_INVALID, "bad n_new / samples_done");
    int64_t n_max = 0;
    for (const BandDev& b : plan->bands) {
        if (b.n_fft > FUSED_MAX_N) return fail(UPMIX_E_UNSUPPORTED, "block streaming needs n_fft <= %d", FUSED_MAX_N);
        if (n_new % b.hop != 0 || samples_done % b.hop != 0)
            return fail(UPMIX_E_INVALID, "block of %d samples is not a multiple of hop %d", n_new, b.hop);
        n_max = std::max<int64_t>(n_max, b.n_fft);
    }
    const int64_t need = upmix_stream_workspace_bytes(plan, n_new, n_tracks);
    if (workspace_bytes < need) return fail(UPMIX_E_WORKSPACE, "workspace too small: %lld given, %lld needed", (long long)workspace_bytes, (long long)need);
    DeviceGuard guard(plan->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const StreamIo io = {in_l, in_r, out_c, out_l, out_r};

    // Steady state: a block's launches depend on its position only through (position mod n_fft) of each band -- ring
    // slots and frame phases -- once every frame it touches exists (position >= 2 n_max).  So the position is folded into
    // [2 n_max, 3 n_max), which makes the launches of blocks one n_max apart IDENTICAL, and the block is replayed as a
    // CUDA graph: one launch instead of ~20 (the CPU-side launch cost was most of a block's wall time).  The caller's
    // input / output pointers change from block to block; only the staging kernel and the band sum see them, and their
    // graph nodes get the new pointers (cudaGraphExecKernelNodeSetParams) before the replay.
    UpmixPlan* mp = const_cast<UpmixPlan*>(plan);
    const int64_t pos = samples_done >= 2 * n_max ? 2 * n_max + samples_done % n_max : samples_done;
    // (blocks much shorter than the largest STFT would need more graphs than the cache holds: they take the plain path)
    int64_t gcd_a = n_max, gcd_b = n_new;
    while (gcd_b) { const int64_t t = gcd_a % gcd_b; gcd_a = gcd_b; gcd_b = t; }
    const int64_t phases = n_max / gcd_a;
    constexpr size_t STREAM_GRAPHS_MAX = 32;
    if (mp->use_graphs && mp->cap_stream && samples_done >= 2 * n_max && n_new <= 65536 && plan->delay <= 65536 &&
        phases <= (int64_t)STREAM_GRAPHS_MAX) {
        const uint64_t align16 = ((reinterpret_cast<uintptr_t>(out_l) | reinterpret_cast<uintptr_t>(out_r) | reinterpret_cast<uintptr_t>(out_c)) & 15) == 0;
        uint64_t key[16] = {(uint64_t)(uintptr_t)state, (uint64_t)pos, (uint64_t)n_new, (uint64_t)n_tracks, (uint64_t)in_stride, (uint64_t)out_stride,
                            (uint64_t)(uintptr_t)workspace, (uint64_t)workspace_bytes, align16, out_c ? 1u : 0u, 0, 0, 0, 0, 0, 0x5354524dull /* "STRM" */};
        for (UpmixPlan::GraphEntry& g : mp->stream_graphs)
            if (memcmp(g.key, key, sizeof(key)) == 0) {
                g.last_use = ++mp->graph_clock;
                if (g.io[0] != in_l || g.io[1] != in_r) {
                    cudaKernelNodeParams kp;
                    CU_CHECK(cudaGraphKernelNodeGetParams(g.stage_node, &kp));
                    void* args[STREAM_STAGE_NARGS];
                    for (int i = 0; i < STREAM_STAGE_NARGS; i++) args[i] = kp.kernelParams[i];
                    args[STREAM_STAGE_IN_L] = (void*)&in_l;
                    args[STREAM_STAGE_IN_L + 1] = (void*)&in_r;
                    kp.kernelParams = args;
                    CU_CHECK(cudaGraphExecKernelNodeSetParams(g.exec, g.stage_node, &kp));
                    g.io[0] = in_l;
                    g.io[1] = in_r;
                }
                if (g.io[2] != out_c || g.io[3] != out_l || g.io[4] != out_r) {
                    cudaKernelNodeParams kp;
                    CU_CHECK(cudaGraphKernelNodeGetParams(g.sum_node, &kp));
                    void* args[BAND_SUM_NARGS];
                    for (int i = 0; i < BAND_SUM_NARGS; i++) args[i] = kp.kernelParams[i];
                    args[BAND_SUM_OUT_C] = (void*)&out_c;
                    args[BAND_SUM_OUT_C + 1] = (void*)&out_l;
                    args[BAND_SUM_OUT_C + 2] = (void*)&out_r;
                    kp.kernelParams = args;
                    CU_CHECK(cudaGraphExecKernelNodeSetParams(g.exec, g.sum_node, &kp));
                    g.io[2] = out_c;
                    g.io[3] = out_l;
                    g.io[4] = out_r;
                }
                CU_CHECK(cudaGraphLaunch(g.exec, st));
                launch_count_add(g.n_kernels);
                return UPMIX_OK;
            }
        cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
        const bool caller_capturing = st && cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs != cudaStreamCaptureStatusNone;
        if (!caller_capturing && cudaStreamBeginCapture(mp->cap_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const unsigned long long before = launch_count(false);
            cudaGraphNode_t stage_node = nullptr, sum_node = nullptr;
            const int rc2 = stream_block_body(plan, state, pos, io, n_new, n_tracks, in_stride, out_stride, workspace, workspace_bytes,
                                              mp->cap_stream, &stage_node, &sum_node);
            cudaGraph_t graph = nullptr;
            const cudaError_t ee = cudaStreamEndCapture(mp->cap_stream, &graph);
            cudaGraphExec_t exec = nullptr;
            if (rc2 == UPMIX_OK && ee == cudaSuccess && graph && stage_node && sum_node && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
                UpmixPlan::GraphEntry g;
                memcpy(g.key, key, sizeof(key));
                g.exec = exec;
                g.graph = graph;                               // kept: the node handles and their parameter storage live in it
                g.stage_node = stage_node;
                g.sum_node = sum_node;
                g.io[0] = in_l; g.io[1] = in_r; g.io[2] = out_c; g.io[3] = out_l; g.io[4] = out_r;
                g.n_kernels = (int)(launch_count(false) - before);
                g.last_use = ++mp->graph_clock;
                if (mp->stream_graphs.size() >= STREAM_GRAPHS_MAX) {          // evict the least recently used
                    size_t lru = 0;
                    for (size_t i = 1; i < mp->stream_graphs.size(); i++)
                        if (mp->stream_graphs[i].last_use < mp->stream_graphs[lru].last_use) lru = i;
                    cudaGraphExecDestroy(mp->stream_graphs[lru].exec);
                    cudaGraphDestroy(mp->stream_graphs[lru].graph);
                    mp->stream_graphs[lru] = g;
                } else {
                    mp->stream_graphs.push_back(g);
                }
                CU_CHECK(cudaGraphLaunch(exec, st));
                return UPMIX_OK;
            }
            if (exec) cudaGraphExecDestroy(exec);
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();                                 // capture failed: run the block the plain way
            if (rc2 != UPMIX_OK) return rc2;
        }
    }
    return stream_block_body(plan, state, pos, io, n_new, n_tracks, in_stride, out_stride, workspace, workspace_bytes, st, nullptr, nullptr);
}

int upmix_frame_step(const UpmixPlan* plan, void* ring, int64_t frame_index, const float* blk_l, const float* blk_r,
                     int n_tracks, int64_t in_stride, float* out_c, float* out_l, float* out_r, int64_t out_stride,
                     void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = check_io(plan, blk_l, blk_r, out_c, out_l, out_r, n_tracks);
    if (rc) return rc;
    if (!ring) return fail(UPMIX_E_INVALID, "ring is NULL");
    if (plan->bands.size() != 1) return fail(UPMIX_E_INVALID, "upmix_frame_step needs a single-band plan");
    const BandDev& b = plan->bands[0];
    if (frame_index < 0) return fail(UPMIX_E_INVALID, "negative frame index");
    DeviceGuard guard(plan->device);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (b.frame_step > 1) return fail(UPMIX_E_UNSUPPORTED, "frame stepping above %d points needs 75%% overlap", FUSED_MAX_N);
    if (b.n_fft > FUSED_MAX_N) {
        // four-step path: the frame is slot 0 of a two-frame wave, its partner slot is zero
        const Layout lay = make_layout(plan, b.hop, n_tracks, false, true);
        if (workspace_bytes < lay.total) return fail(UPMIX_E_WORKSPACE, "workspace too small: %lld given, %lld needed", (long long)workspace_bytes, (long long)lay.total);
        if (plan->out_mode != UPMIX_OUT_LSCRS) return fail(UPMIX_E_UNSUPPORTED, "frame stepping on the four-step path needs Ls/C/Rs output");
        char* scratch = reinterpret_cast<char*>(workspace) + lay.band_out_bytes;
        WaveArgs w;
        w.a = reinterpret_cast<float2*>(scratch);
        w.b1 = reinterpret_cast<float2*>(scratch + lay.a_bytes);
        w.b2 = reinterpret_cast<float2*>(scratch + lay.a_bytes + lay.b1_bytes);
        w.frame0 = 0;
        w.n_frames = 2;
        SegArgs a;
        memset(&a, 0, sizeof(a));
        a.in_l = blk_l;
        a.in_r = blk_r;
        a.in_stride = in_stride;
        a.in_begin = 0;
        a.in_end = b.n_fft;
        CU_CHECK(cudaMemsetAsync(w.a, 0, (size_t)n_tracks * 2 * b.n_fft * sizeof(float2), st));
        WaveArgs w1 = w;
        w1.n_frames = 1;                       // column transform of the one real frame ...
        if (n_tracks != 1) return fail(UPMIX_E_UNSUPPORTED, "frame stepping on the four-step path handles one track per call");
        CU_CHECK(launch_col_fwd(b, a, w1, n_tracks, st));
        CU_CHECK(launch_row_mask(b, w, n_tracks, st));          // ... paired with the zero frame
        CU_CHECK(launch_col_inv_frame(b, w, reinterpret_cast<float*>(ring), out_c, out_l, out_r, out_stride, n_tracks, st));
        return UPMIX_OK;
    }
    float* rings[1] = {reinterpret_cast<float*>(ring)};
    const int64_t s0 = frame_index * b.hop;
    // one band: its hop goes straight to the caller's outputs (no staging slot, no band sum)
    return run_segment(plan, blk_l, blk_r, s0, s0 + b.n_fft, INT64_MAX / 4, s0, s0 + b.hop, n_tracks, in_stride, out_c, out_l,
                       out_r, out_stride, workspace, workspace_bytes, rings, st, true);
}

// ---- main.py's normalise + export on the device -------------------------------------------------
int64_t upmix_peak_workspace_bytes(void) { return (int64_t)(3 * 1184 + 3) * (int64_t)sizeof(float); }

int upmix_peak3(const float* c, const float* l, const float* r, int64_t n, float* peaks3, void* workspace,
                int64_t workspace_bytes, void* stream) {
    if (!c || !l || !r || !peaks3 || !workspace) return fail(UPMIX_E_INVALID, "NULL pointer");
    if (n < 0) return fail(UPMIX_E_INVALID, "negative length");
    if (workspace_bytes < upmix_peak_workspace_bytes()) return fail(UPMIX_E_WORKSPACE, "peak workspace too small");
    int blocks = (int)std::min<int64_t>(1184, std::max<int64_t>(1, (n + 255) / 256));
    CU_CHECK(launch_peak3(c, l, r, n, reinterpret_cast<float*>(workspace), blocks, peaks3, reinterpret_cast<cudaStream_t>(stream)));
    return UPMIX_OK;
}

int upmix_export_mix(int mode, float scale, const float* c, const float* l, const float* r, const float* in_l,
                     const float* in_r, int64_t n, float* out_a, float* out_b, float* out_c, void* stream) {
    if (mode < 0 || mode > 2) return fail(UPMIX_E_INVALID, "unknown export mode %d", mode);
    if (!c || !l || !r || !out_a) return fail(UPMIX_E_INVALID, "NULL pointer");
    if (mode == 0 && (!in_l || !in_r)) return fail(UPMIX_E_INVALID, "AB export needs the input channels");
    if (mode == 1 && (!out_b || !out_c)) return fail(UPMIX_E_INVALID, "split export needs three outputs");
    if (n <= 0) return n == 0 ? UPMIX_OK : fail(UPMIX_E_INVALID, "negative length");
    CU_CHECK(launch_export_mix(c, l, r, in_l, in_r, n, scale, mode, out_a, out_b, out_c, reinterpret_cast<cudaStream_t>(stream)));
    return UPMIX_OK;
}

int upmix_pcm16_to_planar(const int16_t* interleaved, int64_t n, float* l, float* r, float* peak, void* workspace,
                          int64_t workspace_bytes, void* stream) {
    if (!interleaved || !l || !r || !peak || !workspace) return fail(UPMIX_E_INVALID, "NULL pointer");
    if (n < 0) return fail(UPMIX_E_INVALID, "negative length");
    if (workspace_bytes < upmix_peak_workspace_bytes()) return fail(UPMIX_E_WORKSPACE, "workspace too small");
    const int blocks = (int)std::min<int64_t>(1184, std::max<int64_t>(1, (n + 255) / 256));
    CU_CHECK(launch_pcm16_to_planar(interleaved, n, l, r, reinterpret_cast<float*>(workspace), blocks, peak,
                                    reinterpret_cast<cudaStream_t>(stream)));
    return UPMIX_OK;
}

int upmix_stereo_to_pcm16(const float* interleaved, int64_t n, int16_t* out, void* stream) {
    if (!interleaved || !out) return fail(UPMIX_E_INVALID, "NULL pointer");
    if (n <= 0) return n == 0 ? UPMIX_OK : fail(UPMIX_E_INVALID, "negative length");
    CU_CHECK(launch_stereo_to_pcm16(interleaved, n, out, reinterpret_cast<cudaStream_t>(stream)));
    return UPMIX_OK;
}

int upmix_fir_filter(const float* x, int64_t n, int n_tracks, int64_t x_stride, const float* taps, int n_taps, float* y,
                     int64_t y_stride, void* stream) {
    if (!x || !taps || !y) return fail(UPMIX_E_INVALID, "NULL pointer");
    if (n < 0 || n_tracks < 1 || n_tracks > 65535) return fail(UPMIX_E_INVALID, "bad length / n_tracks");
    if (n_taps < 1 || n_taps > 16384) return fail(UPMIX_E_UNSUPPORTED, "n_taps must be in [1, 16384], got %d", n_taps);
    CU_CHECK(launch_fir(x, n, n_tracks, x_stride, taps, n_taps, y, y_stride, reinterpret_cast<cudaStream_t>(stream)));
    return UPMIX_OK;
}

int64_t upmix_debug_launch_count(int reset) { return (int64_t)launch_count(reset != 0); }

int upmix_measure_fp32_tflops(int device, double* tflops, int* sm_count) {
    if (!tflops) return fail(UPMIX_E_INVALID, "tflops is NULL");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(UPMIX_E_CUDA, "cannot select device %d", device);
    int sms = 0;
    CU_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    if (sm_count) *sm_count = sms;
    float* out = nullptr;
    CU_CHECK(cudaMalloc(&out, sizeof(float) * sms * 64));
    cudaEvent_t e0, e1;
    CU_CHECK(cudaEventCreate(&e0));
    CU_CHECK(cudaEventCreate(&e1));
    const int blocks = sms * 8 * 4, iters = 1 << 15;
    double best = 0.0;
    for (int rep = 0; rep < 6; rep++) {
        CU_CHECK(cudaEventRecord(e0, nullptr));
        CU_CHECK(launch_fma_peak(out, blocks, iters, nullptr));
        CU_CHECK(cudaEventRecord(e1, nullptr));
        CU_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        CU_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * 8.0 * (double)iters * 256.0 * (double)blocks;
        if (rep > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return UPMIX_OK;
}

}  // extern "C"

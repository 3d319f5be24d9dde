// wavefront.cu — wavefront path tracer for sm_100a: generate, extend, shade, connect, accumulate.
//
// Reference stages and what replaces them:
//   ray_gen.comp:39-146        k_wf_generate   thin-lens eye rays (hash RNG), SoA queue write, tile-sharded
//   ray_extend.comp:245-268    k_trace_persistent<ExtendIO>   persistent closest-hit traversal on the queue
//   shade.comp:70-266          k_wf_shade      Disney BSDF, MIS on emissive hits, NEE shadow-ray emission,
//                                              queue compaction with ONE atomic per warp (ballot + popc)
//   ray_shadow.comp:245-269    k_trace_persistent<ConnectIO>  persistent any-hit + atomic accumulate
//   blit.comp:15-22            k_wf_finalize   sqrt(acc / spp)
//   host loop src/lib.rs:1706-1729 -> Wavefront::render: no per-bounce host read-back; counts stay in HBM.
// Path state is 64 B/path (4 x float4 SoA), shadow jobs 48 B — the reference's record sizes
// (structs.glsl:4-9, 172-176) so the algorithmic byte counts of SURVEY §8d apply.
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>

#include "shading.cuh"
#include "shade_path.cuh"
#include "trace_kernel.cuh"
#include "wavefront.h"
#include "wavefront_kernels.cuh"

namespace rfw {

#define WF_CK(x)                          \
    do {                                  \
        cudaError_t e_ = (x);             \
        if (e_ != cudaSuccess) return e_; \
    } while (0)

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static uint32_t morton2(uint32_t x, uint32_t y) {
    auto part = [](uint32_t v) {
        v &= 0xFFFF;
        v = (v | (v << 8)) & 0x00FF00FF;
        v = (v | (v << 4)) & 0x0F0F0F0F;
        v = (v | (v << 2)) & 0x33333333;
        v = (v | (v << 1)) & 0x55555555;
        return v;
    };
    return part(x) | (part(y) << 1);
}

// all tile ids of a tiles_x x tiles_y grid in Morton order (host logic of the multi-GPU sharding: tile k of this
// list belongs to rank k mod world and is that rank's tile number k / world)
std::vector<uint32_t> morton_tile_order(uint32_t tiles_x, uint32_t tiles_y) {
    std::vector<uint32_t> order(tiles_x * tiles_y);
    for (uint32_t i = 0; i < order.size(); i++) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [tiles_x](uint32_t a, uint32_t b) { return morton2(a % tiles_x, a / tiles_x) < morton2(b % tiles_x, b / tiles_x); });
    return order;
}

void Wavefront::release() {
    auto fr = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
    fr(d_owned_tiles); fr(d_morton_tiles);
    for (int i = 0; i < 2; i++) { fr(d_O[i]); fr(d_D[i]); fr(d_T[i]); }
    for (int i = 0; i < 2; i++) { fr(d_shO[i]); fr(d_shD[i]); fr(d_shE[i]); }
    fr(d_S); fr(d_partial); fr(d_term); fr(d_accum); fr(d_output); fr(d_counts); fr(d_stats);
    max_paths = 0;
    wave_capacity = 0;
    for (cudaEvent_t e : stage_events) cudaEventDestroy(e);
    stage_events.clear(); stage_marks.clear(); stage_used = 0;
    for (cudaEvent_t e : sync_events) cudaEventDestroy(e);
    sync_events.clear();
    if (main2) { cudaStreamDestroy(main2); main2 = nullptr; }
    if (side2) { cudaStreamDestroy(side2); side2 = nullptr; }
}

cudaError_t Wavefront::configure(uint32_t w, uint32_t h, uint32_t tile_size, uint32_t rank_, uint32_t world_) {
    release();
    width = w; height = h; tile = tile_size ? tile_size : 64; rank = rank_; world = world_ ? world_ : 1;
    tiles_x = (w + tile - 1) / tile; tiles_y = (h + tile - 1) / tile;
    const uint32_t n_tiles = tiles_x * tiles_y;
    morton_tiles = morton_tile_order(tiles_x, tiles_y);
    std::vector<uint32_t> owned;
    for (uint32_t r = rank; r < n_tiles; r += world) owned.push_back(morton_tiles[r]);  // tile k (Morton order) -> rank k mod n
    n_owned_tiles = (uint32_t)owned.size();
    tiles_per_rank = (n_tiles + world - 1) / world;
    max_paths = n_owned_tiles * tile * tile;
    if (w == 0 || h == 0) return cudaSuccess;
    WF_CK(cudaMalloc(&d_owned_tiles, sizeof(uint32_t) * std::max<size_t>(1, owned.size())));
    WF_CK(cudaMalloc(&d_morton_tiles, sizeof(uint32_t) * n_tiles));
    WF_CK(cudaMemcpy(d_owned_tiles, owned.data(), sizeof(uint32_t) * owned.size(), cudaMemcpyHostToDevice));
    WF_CK(cudaMemcpy(d_morton_tiles, morton_tiles.data(), sizeof(uint32_t) * n_tiles, cudaMemcpyHostToDevice));
    WF_CK(cudaMalloc(&d_accum, (size_t)w * h * sizeof(float4)));
    WF_CK(cudaMalloc(&d_output, (size_t)w * h * sizeof(float4)));
    WF_CK(cudaMalloc(&d_counts, 16 * sizeof(uint32_t)));  // 8 per sub-wave lane
    WF_CK(cudaMalloc(&d_stats, 4 * sizeof(unsigned long long)));
    WF_CK(cudaMemset(d_accum, 0, (size_t)w * h * sizeof(float4)));
    WF_CK(cudaMemset(d_output, 0, (size_t)w * h * sizeof(float4)));
    WF_CK(cudaMemset(d_counts, 0, 16 * sizeof(uint32_t)));
    WF_CK(cudaMemset(d_stats, 0, 4 * sizeof(unsigned long long)));
    return ensure_wave(1);
}

// queues and partial accumulators for waves of `b` samples (grow-only)
cudaError_t Wavefront::ensure_wave(uint32_t b) {
    if (b <= wave_capacity) return cudaSuccess;
    auto fr = [](auto*& p) { if (p) cudaFree(p); p = nullptr; };
    for (int i = 0; i < 2; i++) { fr(d_O[i]); fr(d_D[i]); fr(d_T[i]); fr(d_shO[i]); fr(d_shD[i]); fr(d_shE[i]); }
    fr(d_S); fr(d_partial); fr(d_term);
    wave_capacity = 0;
    const size_t mp = (size_t)std::max<uint32_t>(1, max_paths) * b;
    for (int i = 0; i < 2; i++) {
        WF_CK(cudaMalloc(&d_O[i], mp * sizeof(float4)));
        WF_CK(cudaMalloc(&d_D[i], mp * sizeof(float4)));
        WF_CK(cudaMalloc(&d_T[i], mp * sizeof(float4)));
    }
    WF_CK(cudaMalloc(&d_S, mp * sizeof(float4)));
    for (int i = 0; i < 2; i++) {
        WF_CK(cudaMalloc(&d_shO[i], mp * sizeof(float4)));
        WF_CK(cudaMalloc(&d_shD[i], mp * sizeof(float4)));
        WF_CK(cudaMalloc(&d_shE[i], mp * sizeof(float4)));
    }
    const size_t pb = (size_t)width * height * b * sizeof(float4);
    WF_CK(cudaMalloc(&d_partial, pb));
    WF_CK(cudaMemset(d_partial, 0, pb));
    WF_CK(cudaMalloc(&d_term, pb));
    WF_CK(cudaMemset(d_term, 0, pb));
    wave_capacity = b;
    return cudaSuccess;
}

uint32_t Wavefront::wave_spp_for(uint32_t spp) const {
    // as many samples per wave as fit `wave_paths` path slots: deeper bounces then launch on queues large enough to
    // fill the persistent grids (180 GB of HBM makes the 176 B/path of queue + partial accumulator state cheap)
    const uint64_t mp = std::max<uint32_t>(1, max_paths), npix = std::max<uint64_t>(1, (uint64_t)width * height);
    uint64_t b = std::max<uint64_t>(1, wave_paths / mp);
    b = std::min<uint64_t>(b, 0xFFFFFFFFull / npix);  // partial-accumulator indices are 32-bit
    return (uint32_t)std::min<uint64_t>(b, std::max(1u, spp));
}

static FrameParams make_params(const Wavefront& wf, const RfwCameraView3D& cam, uint32_t sample, uint32_t path_length) {
    FrameParams fp;
    fp.cam = cam;
    fp.width = wf.width; fp.height = wf.height; fp.tile = wf.tile; fp.tiles_x = wf.tiles_x; fp.max_paths = wf.max_paths;
    fp.sample = sample; fp.path_length = path_length;
    fp.wave_spp = 1; fp.npix = wf.width * wf.height;
    fp.clamp_value = wf.clamp_value;
    fp.sky[0] = wf.sky[0]; fp.sky[1] = wf.sky[1]; fp.sky[2] = wf.sky[2];
    fp.blue_noise = wf.d_blue_noise; fp.blue_noise_n = wf.blue_noise_n;
    return fp;
}

cudaError_t Wavefront::clear(cudaStream_t stream) {
    if (!d_accum) return cudaSuccess;
    WF_CK(cudaMemsetAsync(d_accum, 0, (size_t)width * height * sizeof(float4), stream));
    WF_CK(cudaMemsetAsync(d_stats, 0, 4 * sizeof(unsigned long long), stream));
    return cudaSuccess;
}

cudaError_t Wavefront::stage_mark(cudaStream_t stream, int stage) {
    if (!stage_timing) return cudaSuccess;
    if (stage_used == stage_events.size()) {
        cudaEvent_t e;
        WF_CK(cudaEventCreate(&e));
        stage_events.push_back(e);
    }
    WF_CK(cudaEventRecord(stage_events[stage_used++], stream));
    stage_marks.push_back(stage);
    return cudaSuccess;
}

cudaError_t Wavefront::stage_times(float out_ms[5]) {
    for (int k = 0; k < 5; k++) out_ms[k] = 0.0f;
    for (size_t i = 1; i < stage_used; i++) {
        float ms = 0.0f;
        WF_CK(cudaEventElapsedTime(&ms, stage_events[i - 1], stage_events[i]));
        out_ms[stage_marks[i]] += ms;
    }
    return cudaSuccess;
}

// RFWB200_WF_TRACE=1 (diagnostic): a timing event after every kernel of every lane; Wavefront::dump_trace prints when each one
// completed relative to the start of the frame (the streams must have been synchronised)
static const bool g_wf_trace = getenv("RFWB200_WF_TRACE") != nullptr;
cudaError_t Wavefront::trace_mark(cudaStream_t st, const char* what, int lane, int bounce) {
    if (!g_wf_trace) return cudaSuccess;
    cudaEvent_t e;
    WF_CK(cudaEventCreate(&e));
    WF_CK(cudaEventRecord(e, st));
    trace_events.push_back({e, what, lane, bounce});
    return cudaSuccess;
}
void Wavefront::dump_trace() {
    if (trace_events.size() < 2) return;
    fprintf(stderr, "rfwb200 wavefront trace (ms since the first mark; completion times):\n");
    for (size_t i = 1; i < trace_events.size(); i++) {
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, trace_events[0].ev, trace_events[i].ev);
        fprintf(stderr, "  lane %d bounce %d %-10s done at %8.3f\n", trace_events[i].lane, trace_events[i].bounce, trace_events[i].what, ms);
    }
    for (auto& t : trace_events) cudaEventDestroy(t.ev);
    trace_events.clear();
}

// One SUB-WAVE: `n_spp` samples of every owned pixel (sample indices first_sample .. first_sample + n_spp - 1) through generate and
// `depth` x { extend -> shade -> connect } on the lane's own pair of streams, queue halves and counters.  `slot0` = first path slot
// of the lane inside the wave queues, `acc0` = first per-sample accumulator plane of the lane.
cudaError_t Wavefront::enqueue_subwave(int lane, cudaStream_t m, cudaStream_t c, bool two_streams, const SceneView& sv, const ShadeScene& ss, const RfwCameraView3D& cam, uint32_t first_sample,
                                       uint32_t n_spp, uint32_t depth, size_t slot0, size_t acc0) {
    const int shade_blocks = sm_count * 8 * 128 / RFW_SHADE_THREADS;
    const TraceTuning tune{refill_below, sv.two_level ? tri_batch_two_level : tri_batch, tri_blocked, inst_batch, grid_rays_per_thread};
    FrameParams fp = make_params(*this, cam, first_sample, 0);
    fp.wave_spp = n_spp;
    const uint32_t cap = max_paths * n_spp;
    uint32_t* counts = d_counts + 8 * lane;
    float4 *O[2] = {d_O[0] + slot0, d_O[1] + slot0}, *D[2] = {d_D[0] + slot0, d_D[1] + slot0}, *T[2] = {d_T[0] + slot0, d_T[1] + slot0}, *S = d_S + slot0;
    float4 *shO[2] = {d_shO[0] + slot0, d_shO[1] + slot0}, *shD[2] = {d_shD[0] + slot0, d_shD[1] + slot0}, *shE[2] = {d_shE[0] + slot0, d_shE[1] + slot0};
    float4* partial = d_partial + acc0 * fp.npix;
    float4* term = d_term + acc0 * fp.npix;
    cudaEvent_t* ev = sync_events.data() + (size_t)lane * (2 * (size_t)depth + 2);  // [2 b] shade(b) done, [2 b + 1] connect(b) done
    cudaStream_t conn = two_streams ? c : m;
    WF_CK(cudaMemsetAsync(counts, 0, 8 * sizeof(uint32_t), m));
    k_wf_generate<<<(cap + 255) / 256, 256, 0, m>>>(fp, d_owned_tiles, O[0], D[0], counts);
    launches++;
    if (lane == 0) WF_CK(stage_mark(m, 0));
    WF_CK(trace_mark(m, "generate", lane, -1));
    for (uint32_t b = 0; b < depth; b++) {
        const int cur = b & 1, nxt = cur ^ 1, sb = b & 1;
        fp.path_length = b;
        // ---- main stream: extend(b) -> shade(b) ------------------------------------------------------------------
        ExtendIO eio{O[cur], D[cur], counts + cur, S};
        if (sv.two_level) WF_CK((launch_persistent_io<ExtendIO, false, true>(m, sm_count, 0, tune, sv, eio, cap, counts + 4)));
        else WF_CK((launch_persistent_io<ExtendIO, false, false>(m, sm_count, 0, tune, sv, eio, cap, counts + 4)));
        if (lane == 0) WF_CK(stage_mark(m, 1));
        WF_CK(trace_mark(m, "extend", lane, (int)b));
        if (two_streams && b >= 2) WF_CK(cudaStreamWaitEvent(m, ev[2 * (b - 2) + 1], 0));  // shadow queue `sb` is free again: connect(b - 2) is done
        k_wf_shade<<<shade_blocks, RFW_SHADE_THREADS, 0, m>>>(fp, ss, S, O[cur], D[cur], T[cur], O[nxt], D[nxt], T[nxt], shO[sb], shD[sb], shE[sb], term, counts + cur, counts + nxt,
                                                              counts + 2 + sb);
        k_wf_advance_paths<<<1, 1, 0, m>>>(counts, d_stats, cur);
        if (lane == 0) WF_CK(stage_mark(m, 2));
        WF_CK(trace_mark(m, "shade", lane, (int)b));
        if (two_streams) {
            WF_CK(cudaEventRecord(ev[2 * b], m));
            WF_CK(cudaStreamWaitEvent(conn, ev[2 * b], 0));
        }
        // ---- connect(b): beside extend(b + 1) / shade(b + 1) -------------------------------------------------------
        ConnectIO cio{shO[sb], shD[sb], shE[sb], counts + 2 + sb, reinterpret_cast<float*>(partial)};
        if (sv.two_level) WF_CK((launch_persistent_io<ConnectIO, true, true>(conn, sm_count, 0, tune, sv, cio, cap, counts + 5)));
        else WF_CK((launch_persistent_io<ConnectIO, true, false>(conn, sm_count, 0, tune, sv, cio, cap, counts + 5)));
        k_wf_advance_shadow<<<1, 1, 0, conn>>>(counts, d_stats, sb);
        if (two_streams) WF_CK(cudaEventRecord(ev[2 * b + 1], conn));
        if (lane == 0) WF_CK(stage_mark(m, 3));
        WF_CK(trace_mark(conn, "connect", lane, (int)b));
        launches += 5;
    }
    return cudaGetLastError();
}

cudaError_t Wavefront::render(cudaStream_t stream, const SceneView& sv, const ShadeScene& ss, const RfwCameraView3D& cam, uint32_t first_sample, uint32_t spp, uint32_t depth) {
    if (max_paths == 0 || depth == 0) return cudaSuccess;
    const uint32_t wave = wave_spp_for(spp);
    WF_CK(ensure_wave(wave));
    stage_used = 0; stage_marks.clear();
    // per-stage timing needs the stages one after the other: every overlap is switched off for that (diagnostic) run
    const bool two_streams = overlap && side != nullptr && !stage_timing;
    // Two sub-waves in flight: a persistent traversal launch ends on its few longest rays (0.2 - 0.3 ms with a handful of warps
    // busy), and the bounces of ONE wave are a dependency chain — nothing of that wave can fill the hole.  A wave of >= 2
    // samples per pixel is therefore split into two halves (samples [0, h) and [h, n)), each with its own half of the queues,
    // its own counters and its own pair of streams: while one half drains a tail, the other half's kernels fill the SMs.  The
    // per-sample accumulators keep the image independent of the split (k_wf_reduce folds all samples in sample order).
    const bool split = two_streams && split_waves;
    if (split && !main2) {
        int prio_lo = 0, prio_hi = 0;
        WF_CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        WF_CK(cudaStreamCreateWithPriority(&main2, cudaStreamNonBlocking, prio_hi));
        WF_CK(cudaStreamCreateWithPriority(&side2, cudaStreamNonBlocking, prio_lo));
    }
    while (sync_events.size() < 2 * (2 * (size_t)depth + 2) + 2) {
        cudaEvent_t e;
        WF_CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        sync_events.push_back(e);
    }
    cudaEvent_t ev_fork = sync_events[sync_events.size() - 2], ev_join = sync_events[sync_events.size() - 1];
    const size_t lane_events = 2 * (size_t)depth + 2;
    WF_CK(stage_mark(stream, 4));
    WF_CK(trace_mark(stream, "start", 0, -1));
    for (uint32_t s = 0; s < spp; s += wave) {
        const uint32_t wspp = std::min(wave, spp - s);
        const uint32_t half0 = (split && wspp >= 2) ? (wspp + 1) / 2 : wspp;
        // (the previous wave's reduce on `stream` came after all of its connects: queues, counters and accumulators are free)
        if (half0 < wspp) {
            WF_CK(cudaEventRecord(ev_fork, stream));
            WF_CK(cudaStreamWaitEvent(main2, ev_fork, 0));
        }
        WF_CK(enqueue_subwave(0, stream, side, two_streams, sv, ss, cam, first_sample + s, half0, depth, 0, 0));
        if (half0 < wspp) {
            WF_CK(enqueue_subwave(1, main2, side2, true, sv, ss, cam, first_sample + s + half0, wspp - half0, depth, (size_t)max_paths * half0, half0));
            WF_CK(cudaStreamWaitEvent(stream, sync_events[lane_events + 2 * (depth - 1) + 1], 0));  // lane 1's last connect (its connects are ordered among themselves)
            WF_CK(cudaEventRecord(ev_join, main2));                                                // ... and its last shade (terminal accumulators)
            WF_CK(cudaStreamWaitEvent(stream, ev_join, 0));
        }
        if (two_streams) WF_CK(cudaStreamWaitEvent(stream, sync_events[2 * (depth - 1) + 1], 0));  // lane 0's last connect
        FrameParams fp = make_params(*this, cam, first_sample + s, 0);
        fp.wave_spp = wspp;
        k_wf_reduce<<<(max_paths + 255) / 256, 256, 0, stream>>>(fp, d_owned_tiles, d_partial, d_term, d_accum);
        launches++;
        WF_CK(stage_mark(stream, 4));
        WF_CK(trace_mark(stream, "reduce", 0, -1));
    }
    return cudaGetLastError();
}

cudaError_t Wavefront::debug_view(cudaStream_t stream, const SceneView& sv, const ShadeScene& ss, const RfwCameraView3D& cam, uint32_t mode) {
    if (max_paths == 0) return cudaSuccess;
    WF_CK(ensure_wave(1));
    FrameParams fp = make_params(*this, cam, 0, 0);
    const TraceTuning tune{refill_below, sv.two_level ? tri_batch_two_level : tri_batch, tri_blocked, inst_batch, 0};
    WF_CK(cudaMemsetAsync(d_counts, 0, 8 * sizeof(uint32_t), stream));
    k_wf_generate_centre<<<(max_paths + 255) / 256, 256, 0, stream>>>(fp, d_owned_tiles, d_O[0], d_D[0], d_counts);
    ExtendIO eio{d_O[0], d_D[0], d_counts + 0, d_S};
    if (sv.two_level) WF_CK((launch_persistent_io<ExtendIO, false, true>(stream, sm_count, 0, tune, sv, eio, max_paths, d_counts + 4)));
    else WF_CK((launch_persistent_io<ExtendIO, false, false>(stream, sm_count, 0, tune, sv, eio, max_paths, d_counts + 4)));
    k_wf_debug_view<<<sm_count * 8, 128, 0, stream>>>(fp, ss, mode, d_S, d_O[0], d_D[0], d_counts + 0, d_output);
    launches += 3;
    return cudaGetLastError();
}

cudaError_t Wavefront::finalize(cudaStream_t stream, uint32_t sample_count) {
    if (max_paths == 0) return cudaSuccess;
    RfwCameraView3D cam{};
    FrameParams fp = make_params(*this, cam, 0, 0);
    k_wf_finalize<<<(max_paths + 255) / 256, 256, 0, stream>>>(fp, d_owned_tiles, d_accum, d_output, 1.0f / (float)std::max(1u, sample_count));
    launches++;
    return cudaGetLastError();
}

cudaError_t Wavefront::export_tiles(cudaStream_t stream, float* d_out) {
    if (max_paths == 0) return cudaSuccess;
    RfwCameraView3D cam{};
    FrameParams fp = make_params(*this, cam, 0, 0);
    k_wf_export<<<(max_paths + 255) / 256, 256, 0, stream>>>(fp, d_owned_tiles, d_accum, reinterpret_cast<float4*>(d_out));
    launches++;
    return cudaGetLastError();
}

cudaError_t Wavefront::assemble(cudaStream_t stream, const float* d_gathered, uint32_t tiles_per_rank_, uint32_t world_, uint32_t sample_count, float* d_image) {
    RfwCameraView3D cam{};
    FrameParams fp = make_params(*this, cam, 0, 0);
    const size_t total = (size_t)world_ * tiles_per_rank_ * tile * tile;
    if (total == 0) return cudaSuccess;
    k_wf_assemble<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(fp, d_morton_tiles, tiles_x * tiles_y, reinterpret_cast<const float4*>(d_gathered), tiles_per_rank_, world_,
                                                                       1.0f / (float)std::max(1u, sample_count), reinterpret_cast<float4*>(d_image));
    launches++;
    return cudaGetLastError();
}

}  // namespace rfw

// Host SIMT harness, part 2 (test infrastructure, never shipped): the wavefront path tracer's KERNELS themselves
// (rfw_rs_b200/csrc/wavefront_kernels.cuh: k_wf_generate, k_wf_shade with its per-CTA queue-slot reservation, k_wf_advance,
// k_wf_reduce, and the persistent traversal kernel behind the ExtendIO / ConnectIO policies) launched in the order
// Wavefront::render launches them (wavefront.cu), on the lane-thread SIMT machine of simt_machine.h.  The queues, their
// counters, the shadow queue and the per-sample partial accumulators are plain host arrays here.
#include "device_shims.h"
#include "simt_machine.h"

static unsigned char* rfw_host_smem = nullptr;
#define RFW_HOST_SIMT 1
#include "../../rfw_rs_b200/csrc/wavefront_kernels.cuh"

alignas(16) static unsigned char g_smem[65536];
using namespace rfw;

template <class IO, bool ANY>
static bool run_trace(const SceneView& sv, const IO& io, uint32_t* counter, const TraceTuning& tune) {
    *counter = 0;  // launch_persistent_mb clears the work counter before every launch
    rfw_host_smem = g_smem;
    if (sv.two_level) return simt_launch(1, PT_THREADS, [&]() { k_trace_persistent<IO, ANY, true, PT_THREADS, 8, PT_SM_STACK>(sv, io, counter, tune); });
    return simt_launch(1, PT_THREADS, [&]() { k_trace_persistent<IO, ANY, false, PT_THREADS, 8, PT_SM_STACK>(sv, io, counter, tune); });
}

extern "C" {
// One wave of `spp` samples per pixel, `depth` bounces, single rank; acc = h*w*4 floats accumulated into.
// Returns 0, or -1 when a kernel hung (a thread missed a barrier / warp collective).  stats: [0] extension rays, [1] shadow rays.
int wf_render(const void* scene_view, const InstanceShading* inst_table, const RfwDeviceMaterial* mats, uint32_t n_mats, const RfwAreaLight* area, uint32_t na,
              const RfwPointLight* point, uint32_t np, const RfwSpotLight* spot, uint32_t ns, const RfwDirectionalLight* dir, uint32_t nd, const TexDesc* textures,
              uint32_t n_textures, const TexDesc* skybox, const RfwCameraView3D* cam, uint32_t w, uint32_t h, uint32_t tile, const uint32_t* owned_tiles,
              uint32_t n_owned, uint32_t first_sample, uint32_t spp, uint32_t depth, float clamp_value, const float* sky, int shade_ctas, float* acc, uint64_t* stats) {
    const SceneView& sv = *reinterpret_cast<const SceneView*>(scene_view);
    ShadeScene ss;
    memset(&ss, 0, sizeof(ss));
    ss.inst = inst_table; ss.materials = mats; ss.n_materials = n_mats;
    ss.area = area; ss.point = point; ss.spot = spot; ss.dir = dir;
    ss.n_area = (int)na; ss.n_point = (int)np; ss.n_spot = (int)ns; ss.n_dir = (int)nd;
    ss.textures = textures; ss.n_textures = n_textures;
    if (skybox) { ss.has_sky = 1u; ss.sky = *skybox; }
    FrameParams fp;
    memset(&fp, 0, sizeof(fp));
    fp.cam = *cam; fp.width = w; fp.height = h; fp.tile = tile; fp.tiles_x = (w + tile - 1) / tile; fp.max_paths = n_owned * tile * tile;
    fp.sample = first_sample; fp.path_length = 0; fp.wave_spp = spp; fp.npix = w * h;
    fp.clamp_value = clamp_value; fp.sky[0] = sky[0]; fp.sky[1] = sky[1]; fp.sky[2] = sky[2];
    const uint32_t cap = fp.max_paths * spp;
    std::vector<float4> O[2], D[2], T[2], S(cap), shO[2], shD[2], shE[2], partial((size_t)fp.npix * spp, f4(0, 0, 0, 0)), term((size_t)fp.npix * spp, f4(0, 0, 0, 0));
    for (int i = 0; i < 2; i++) { O[i].resize(cap); D[i].resize(cap); T[i].resize(cap); shO[i].resize(cap); shD[i].resize(cap); shE[i].resize(cap); }
    uint32_t counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    unsigned long long st[4] = {0, 0, 0, 0};
    const TraceTuning tune{28, sv.two_level ? 4 : 4, 4, 6, 0};
    g_abort.store(false);
    // generate
    if (!simt_launch((cap + 255) / 256, 256, [&]() { k_wf_generate(fp, owned_tiles, O[0].data(), D[0].data(), counts); })) return -1;
    // Wavefront::render's launch order; connect(b) is issued on a second stream there (beside extend(b + 1) / shade(b + 1)) and
    // runs here at the LATEST point its dependencies allow — after shade(b + 1), right before shade(b + 2) re-uses its shadow
    // queue — so the serial schedule exercises the double buffering the concurrent one relies on
    auto connect = [&](uint32_t b) -> bool {
        const int sb = (int)(b & 1u);
        const ConnectIO cio{shO[sb].data(), shD[sb].data(), shE[sb].data(), counts + 2 + sb, reinterpret_cast<float*>(partial.data())};
        if (!run_trace<ConnectIO, true>(sv, cio, counts + 5, tune)) return false;
        return simt_launch(1, 1, [&]() { k_wf_advance_shadow(counts, st, sb); });
    };
    for (uint32_t b = 0; b < depth; b++) {
        const int cur = (int)(b & 1u), nxt = cur ^ 1, sb = (int)(b & 1u);
        fp.path_length = b;
        const ExtendIO eio{O[cur].data(), D[cur].data(), counts + cur, S.data()};
        if (!run_trace<ExtendIO, false>(sv, eio, counts + 4, tune)) return -1;
        if (b >= 2 && !connect(b - 2)) return -1;
        if (!simt_launch((unsigned)shade_ctas, RFW_SHADE_THREADS, [&]() {
                k_wf_shade(fp, ss, S.data(), O[cur].data(), D[cur].data(), T[cur].data(), O[nxt].data(), D[nxt].data(), T[nxt].data(), shO[sb].data(), shD[sb].data(), shE[sb].data(),
                           term.data(), counts + cur, counts + nxt, counts + 2 + sb);
            })) return -1;
        if (!simt_launch(1, 1, [&]() { k_wf_advance_paths(counts, st, cur); })) return -1;
    }
    for (uint32_t b = depth >= 2 ? depth - 2 : 0; b < depth; b++)
        if (!connect(b)) return -1;
    if (!simt_launch((fp.max_paths + 255) / 256, 256, [&]() { k_wf_reduce(fp, owned_tiles, partial.data(), term.data(), reinterpret_cast<float4*>(acc)); })) return -1;
    if (stats) { stats[0] = st[0]; stats[1] = st[1]; }
    return 0;
}
}

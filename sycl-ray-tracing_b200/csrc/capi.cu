// extern "C" shim of include/b200rt.h: argument checking, device residency, launches. No compute happens on the host.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/b200rt.h"
#include "bvh_build.h"
#include "device_types.h"
#include "kernels.h"
#include "obj_ingest.h"

using namespace b200rt;

struct b200rt_obj { ParsedObjArrays a; };
struct b200rt_hdr { std::vector<float> rgb; int w = 0, h = 0; };

struct b200rt_bvh { FlatBVH flat; };

struct b200rt_scene
{
    int device = 0;
    SceneDev dev{};
    b200rt_bvh_info info{};
    size_t bytes = 0;
    std::vector<void*> allocs;
    MaterialDev* d_mats = nullptr; int mats_cap = 0;
    std::vector<char> mat_used;           // material slots some primitive references (slot 0 of parse_obj is emissive but normally unused)
    int max_mat_index = -1;               // highest material index any primitive references: every later material set must cover it
    // per-scene scratch, grown on demand and kept across calls
    unsigned int* d_work = nullptr;
    unsigned long long* d_rays = nullptr;
    float4* d_tiles = nullptr; size_t tiles_cap = 0;
    float4* d_image = nullptr; size_t image_cap = 0;
    int* d_prim = nullptr; float* d_t = nullptr; size_t prim_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // wavefront integrator state (allocated on first use, grown on demand)
    WfGroup wf[kMaxWfGroups]{}; int wf_groups = 0; std::vector<void*> wf_allocs; int wf_cap[kMaxWfGroups] = {};   // slots allocated per group
    unsigned int* h_active = nullptr;     // pinned, one word per group
    cudaEvent_t fork_event = nullptr;
    float* d_env_lum = nullptr;           // per-texel luminance as the integrator computes it (image.h:80-85), kept for the alias table
    float2* d_alias = nullptr; float alias_total = 0.0f;
    void* bvh_window = nullptr; size_t bvh_window_bytes = 0;   // 8-ary nodes + triangles, kept resident in L2 (see pin_bvh_in_l2)
    bool l2_pinned = false, l2_window_set = false;
    double kernel_times[4] = {};          // B200RT_FLAG_TIME_KERNELS: trace ms, shade ms, trace launches, shade launches of the last render
    WfTimeline timeline;                  // B200RT_FLAG_TIME_INLINE: per-launch events of the last wavefront frame
    bool timeline_valid = false;
    // persistent integrator state (allocated on first use)
    WfBuffers pw{}; int pw_cap = 0, pw_grid = 0; std::vector<void*> pw_allocs;
    // multi-GPU scenes (b200rt_scene_create_multi): this scene is rank 0; replicas[i] is rank i + 1 on another device, owned here
    std::vector<b200rt_scene*> replicas;
    float4* d_gather = nullptr; size_t gather_cap = 0;      // rank 0: world per-rank tile buffers, rank-major (the peer copies' destination)
    cudaStream_t mg_stream = nullptr;                       // per device: the stream a multi-GPU frame runs on
    unsigned char* d_rgba8 = nullptr; size_t rgba8_cap = 0; // output stage scratch (b200rt_render_rgba8)
    // pinned staging for copies from / to pageable caller memory: two chunks, ping-pong
    void* h_stage[2] = { nullptr, nullptr }; cudaEvent_t stage_ev[2] = { nullptr, nullptr };
};

namespace {

thread_local std::string g_error = "";

int fail(int code, const char* fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_error = buf;
    return code;
}

#define CU(expr)                                                                                             \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess) return fail(B200RT_ERR_CUDA, "%s failed: %s", #expr, cudaGetErrorString(e__)); \
    } while (0)

// Makes `device` current for the lifetime of the guard and restores the caller's device afterwards: the library must not
// change the calling thread's current device behind its back (torch and other CUDA users share the thread).
struct DeviceScope
{
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceScope(int device)
    {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != device) err = cudaSetDevice(device);
        else if (err == cudaSuccess) prev = -1;          // nothing to restore
    }
    ~DeviceScope() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceScope(const DeviceScope&) = delete;
    DeviceScope& operator=(const DeviceScope&) = delete;
};
#define ON_DEVICE(dev)                                                                                          \
    DeviceScope device_scope__(dev);                                                                             \
    if (device_scope__.err != cudaSuccess) return fail(B200RT_ERR_CUDA, "cudaSetDevice(%d) failed: %s", (dev), cudaGetErrorString(device_scope__.err))

template <typename T>
int upload(b200rt_scene* s, const T* host, size_t n, const T** dev_out)
{
    T* d = nullptr;
    size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    CU(cudaMalloc(&d, bytes));
    s->allocs.push_back(d);
    s->bytes += bytes;
    if (n) CU(cudaMemcpy(d, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *dev_out = d;
    return B200RT_OK;
}

template <typename T>
int device_alloc(b200rt_scene* s, T** out, size_t n)
{
    T* d = nullptr;
    const size_t bytes = std::max<size_t>(n, 1) * sizeof(T);
    CU(cudaMalloc(&d, bytes));
    s->allocs.push_back(d);
    s->bytes += bytes;
    *out = d;
    return B200RT_OK;
}

int make_params(const b200rt_scene* s, const float* camera17, int w, int h, int spp, int bounces,
                const b200rt_render_options* o, RenderParams& P)
{
    if (!s || !camera17) return fail(B200RT_ERR_ARG, "scene and camera must not be NULL");
    if (w <= 0 || h <= 0) return fail(B200RT_ERR_ARG, "bad frame size %dx%d", w, h);
    if (spp < 0 || bounces < 0) return fail(B200RT_ERR_ARG, "spp and max_bounces must be >= 0");
    b200rt_render_options d;
    b200rt_default_render_options(&d);
    if (o) d = *o;
    if (d.world <= 0) d.world = 1;
    if (d.rank < 0 || d.rank >= d.world) return fail(B200RT_ERR_ARG, "rank %d outside world %d", d.rank, d.world);
    std::memcpy(P.cam.m, camera17, 16 * sizeof(float));
    P.cam.fov_dist = camera17[16];
    P.cam.w = w; P.cam.h = h;
    P.spp = spp; P.max_bounces = bounces;
    P.rank = d.rank; P.world = d.world;
    P.tiles_x = (w + kTileDim - 1) / kTileDim;
    P.tile_skew = tile_skew();
    P.tiles_y = (h + kTileDim - 1) / kTileDim;
    P.n_rank_tiles = b200rt_tiles_for_rank(w, h, d.rank, d.world);
    P.flags = d.flags;
    P.rx0 = P.ry0 = P.rw = P.rh = 0;
    P.sample_begin = 0; P.sample_end = spp; P.stamp0 = 0; P.acc_rng = nullptr; P.acc_sum = nullptr;
    return B200RT_OK;
}

int ensure_scratch(b200rt_scene* s, size_t tile_px, size_t image_px, size_t prim_px)
{
    if (!s->d_work)
    {
        CU(cudaMalloc(&s->d_work, sizeof(unsigned int)));
        CU(cudaMalloc(&s->d_rays, sizeof(unsigned long long)));
        CU(cudaEventCreate(&s->ev0));
        CU(cudaEventCreate(&s->ev1));
    }
    if (tile_px > s->tiles_cap)
    {
        if (s->d_tiles) cudaFree(s->d_tiles);
        s->d_tiles = nullptr; s->tiles_cap = 0;
        CU(cudaMalloc(&s->d_tiles, tile_px * sizeof(float4)));
        s->tiles_cap = tile_px;
    }
    if (image_px > s->image_cap)
    {
        if (s->d_image) cudaFree(s->d_image);
        s->d_image = nullptr; s->image_cap = 0;
        CU(cudaMalloc(&s->d_image, image_px * sizeof(float4)));
        s->image_cap = image_px;
    }
    if (prim_px > s->prim_cap)
    {
        if (s->d_prim) cudaFree(s->d_prim);
        if (s->d_t) cudaFree(s->d_t);
        s->d_prim = nullptr; s->d_t = nullptr; s->prim_cap = 0;
        CU(cudaMalloc(&s->d_prim, prim_px * sizeof(int)));
        CU(cudaMalloc(&s->d_t, prim_px * sizeof(float)));
        s->prim_cap = prim_px;
    }
    return B200RT_OK;
}


// ---- host <-> device copies -----------------------------------------------------------------------------------------------------
// Caller buffers may be pageable (std::vector, numpy) or page-locked (b200rt_host_alloc, cudaHostRegister). Page-locked memory
// is copied directly at the PCIe rate. Pageable memory goes through two pinned chunks owned by the scene: the DMA of one
// chunk overlaps the host memcpy of the other (a plain cudaMemcpy on pageable memory does the same inside the driver, one
// thread and smaller chunks).
constexpr size_t kStageChunk = 8u << 20;

bool host_pointer_is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

void host_memcpy(void* dst, const void* src, size_t bytes)
{
    constexpr size_t kBlock = 1u << 20;
    const long long n_blocks = (long long)((bytes + kBlock - 1) / kBlock);
    if (n_blocks < 4) { std::memcpy(dst, src, bytes); return; }
#pragma omp parallel for schedule(static)
    for (long long b = 0; b < n_blocks; b++)
    {
        const size_t off = (size_t)b * kBlock;
        std::memcpy((char*)dst + off, (const char*)src + off, std::min(kBlock, bytes - off));
    }
}

int ensure_stage(b200rt_scene* s)
{
    if (s->h_stage[0]) return B200RT_OK;
    for (int i = 0; i < 2; i++)
    {
        CU(cudaMallocHost(&s->h_stage[i], kStageChunk));
        CU(cudaEventCreateWithFlags(&s->stage_ev[i], cudaEventDisableTiming));
    }
    return B200RT_OK;
}

// asynchronous with respect to the host only for pinned sources; in both cases `src` may be reused when the call returns
// only after the stream has been synchronised (pinned) or immediately (pageable: it has been copied out)
int copy_to_device(b200rt_scene* s, void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st)
{
    if (!bytes) return B200RT_OK;
    if (host_pointer_is_pinned(src_host)) { CU(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, st)); return B200RT_OK; }
    int rc = ensure_stage(s);
    if (rc) return rc;
    size_t off = 0;
    for (int i = 0; off < bytes; i++, off += kStageChunk)
    {
        const int b = i & 1;
        const size_t n = std::min(kStageChunk, bytes - off);
        if (i >= 2) CU(cudaEventSynchronize(s->stage_ev[b]));            // the DMA that last read this chunk has finished
        host_memcpy(s->h_stage[b], (const char*)src_host + off, n);
        CU(cudaMemcpyAsync((char*)dst_dev + off, s->h_stage[b], n, cudaMemcpyHostToDevice, st));
        CU(cudaEventRecord(s->stage_ev[b], st));
    }
    return B200RT_OK;
}

// blocking: returns when dst_host holds the data
int copy_to_host(b200rt_scene* s, void* dst_host, const void* src_dev, size_t bytes, cudaStream_t st)
{
    if (!bytes) { CU(cudaStreamSynchronize(st)); return B200RT_OK; }
    if (host_pointer_is_pinned(dst_host))
    {
        CU(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, st));
        CU(cudaStreamSynchronize(st));
        return B200RT_OK;
    }
    int rc = ensure_stage(s);
    if (rc) return rc;
    const size_t n_chunks = (bytes + kStageChunk - 1) / kStageChunk;
    auto issue = [&](size_t i) -> cudaError_t {
        const size_t off = i * kStageChunk, n = std::min(kStageChunk, bytes - off);
        cudaError_t e = cudaMemcpyAsync(s->h_stage[i & 1], (const char*)src_dev + off, n, cudaMemcpyDeviceToHost, st);
        return e != cudaSuccess ? e : cudaEventRecord(s->stage_ev[i & 1], st);
    };
    CU(issue(0));
    if (n_chunks > 1) CU(issue(1));
    for (size_t i = 0; i < n_chunks; i++)
    {
        const size_t off = i * kStageChunk, n = std::min(kStageChunk, bytes - off);
        CU(cudaEventSynchronize(s->stage_ev[i & 1]));
        host_memcpy((char*)dst_host + off, s->h_stage[i & 1], n);
        if (i + 2 < n_chunks) CU(issue(i + 2));
    }
    return B200RT_OK;
}

template <typename T>
cudaError_t wf_alloc(b200rt_scene* s, T** p, size_t n)
{
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, n * sizeof(T));
    if (e == cudaSuccess) { s->wf_allocs.push_back(d); *p = (T*)d; }
    return e;
}

// tile groups (independent iteration chains on their own streams). 3 for a full 1080p frame; a rank that holds a fraction of it
// (multi-GPU) is latency-bound per pass and gains a little from a fourth chain (one rank of 8 emulated: 115.9 -> 112.2 ms)
int wavefront_group_count(const RenderParams& P)
{
    const char* e = getenv("B200RT_WF_GROUPS");          // read per frame (tools sweep it inside one process)
    const int env_n = e ? std::min(atoi(e), kMaxWfGroups) : 0;
    if (env_n >= 1) return env_n;
    return (long long)P.n_rank_tiles * kTilePixels < 700000 ? 4 : 3;
}

int ensure_wavefront(b200rt_scene* s, const RenderParams& P)
{
    const int G = wavefront_group_count(P);
    if (!s->h_active)
    {
        CU(cudaMallocHost(&s->h_active, kMaxWfGroups * sizeof(unsigned int)));
        CU(cudaEventCreateWithFlags(&s->fork_event, cudaEventDisableTiming));
        for (int g = 0; g < kMaxWfGroups; g++)
        {
            CU(cudaStreamCreateWithFlags(&s->wf[g].stream, cudaStreamNonBlocking));
            CU(cudaStreamCreateWithFlags(&s->wf[g].detach_stream, cudaStreamNonBlocking));
            CU(cudaEventCreateWithFlags(&s->wf[g].poll_event, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s->wf[g].join_event, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s->wf[g].detach_ready, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&s->wf[g].detach_join, cudaEventDisableTiming));
            s->wf[g].host_active = s->h_active + g;
        }
    }
    // group g of rank r == rank r + g * world of a (world * G) partition
    int need[kMaxWfGroups];
    bool fits = s->wf_groups == G;
    int most = 1;
    for (int g = 0; g < G; g++)
    {
        need[g] = b200rt_tiles_for_rank(P.cam.w, P.cam.h, P.rank + g * P.world, P.world * G) * kTilePixels;
        most = std::max(most, need[g]);
        if (fits && need[g] > s->wf_cap[g]) fits = false;
    }
    if (!fits)
    {
        // (re)allocate every group at the largest group's size: the buffers are then reusable for any rank/world split of
        // this frame size or smaller (the arrays are indexed k * n_slots + slot, so only the capacity matters)
        for (void* p : s->wf_allocs) cudaFree(p);
        s->wf_allocs.clear();
        for (int g = 0; g < G; g++)
        {
            WfBuffers& w = s->wf[g].buf;
            const size_t n = (size_t)most;
            s->wf_cap[g] = most;
            CU(wf_alloc(s, &w.rng, n)); CU(wf_alloc(s, &w.sample, n)); CU(wf_alloc(s, &w.bounce, n)); CU(wf_alloc(s, &w.flags, n));
            CU(wf_alloc(s, &w.final_c, n)); CU(wf_alloc(s, &w.sample_c, n)); CU(wf_alloc(s, &w.thr, n)); CU(wf_alloc(s, &w.thr_next, n));
            CU(wf_alloc(s, &w.ray_o, 5 * n)); CU(wf_alloc(s, &w.ray_d, 5 * n)); CU(wf_alloc(s, &w.side_w, 4 * n));
            CU(wf_alloc(s, &w.res, 5 * n));
            CU(wf_alloc(s, &w.queue, 5 * n)); CU(wf_alloc(s, &w.counters, 8)); CU(wf_alloc(s, &w.rays_total, 1));
            s->wf[g].dmem = WfDetachMem{};          // the study paths' memory is allocated below, when a frame asks for them
            s->wf[g].amem = WfAsyncMem{};
        }
        s->wf_groups = G;
    }
    // study paths, only when a frame asks for them: the early hand-over of a group's lagging pixels (persist.cu: control
    // words, the list, the tail kernel's private queues) and the ray ring + per-chunk counts of the cross-warp continuation (async.cu);
    // sized for the groups' capacity, freed with the rest
    const auto env_on = [](const char* name) { const char* e = getenv(name); return e && atoi(e) != 0; };
    const bool want_detach = s->dev.has_wide && ((P.flags & B200RT_FLAG_WF_DETACH) || env_on("B200RT_WF_DETACH"));
    const bool want_async = s->dev.has_wide && ((P.flags & B200RT_FLAG_WF_ASYNC) || env_on("B200RT_WF_ASYNC"));
    for (int g = 0; g < G; g++)
    {
        WfDetachMem& dm = s->wf[g].dmem;
        if (want_detach && !dm.list)
        {
            dm.list_words = 65536; dm.ctas = wavefront_tail_max_ctas();
            CU(wf_alloc(s, &dm.ctl, (size_t)wavefront_detach_ctl_words())); CU(wf_alloc(s, &dm.list, (size_t)dm.list_words));
            CU(wf_alloc(s, &dm.queue, (size_t)wavefront_detach_queue_words(dm.ctas)));
        }
        WfAsyncMem& a = s->wf[g].amem;
        if (want_async && !a.ray_ring && wavefront_async_fits(s->wf_cap[g]))
        {
            a.ray_log2 = wavefront_async_ray_log2(s->wf_cap[g]); a.chunk_words = wavefront_async_chunk_words(s->wf_cap[g]);
            CU(wf_alloc(s, &a.ray_ring, (size_t)1 << a.ray_log2)); CU(wf_alloc(s, &a.ctrl, 64));
            CU(wf_alloc(s, &a.chunk_cnt, (size_t)a.chunk_words)); CU(wf_alloc(s, &a.chunk_live, (size_t)a.chunk_words));
        }
    }
    for (int g = 0; g < G; g++)
    {
        s->wf[g].buf.n_slots = need[g];
        s->wf[g].buf.tile_stride = G;
        s->wf[g].buf.tile_offset = g;
    }
    return B200RT_OK;
}

template <typename T>
cudaError_t pw_alloc(b200rt_scene* s, T** p, size_t n)
{
    void* d = nullptr;
    cudaError_t e = cudaMalloc(&d, n * sizeof(T));
    if (e == cudaSuccess) { s->pw_allocs.push_back(d); *p = (T*)d; }
    return e;
}

int ensure_persistent(b200rt_scene* s)
{
    const int grid = persistent_grid();
    const int n = persistent_slots(grid);
    s->pw_grid = grid;
    if (n <= s->pw_cap) { s->pw.n_slots = n; return B200RT_OK; }
    for (void* p : s->pw_allocs) cudaFree(p);
    s->pw_allocs.clear(); s->pw_cap = 0;
    WfBuffers& w = s->pw;
    const size_t m = (size_t)n;
    CU(pw_alloc(s, &w.rng, m)); CU(pw_alloc(s, &w.sample, m)); CU(pw_alloc(s, &w.bounce, m)); CU(pw_alloc(s, &w.flags, m));
    CU(pw_alloc(s, &w.final_c, m)); CU(pw_alloc(s, &w.sample_c, m)); CU(pw_alloc(s, &w.thr, m)); CU(pw_alloc(s, &w.thr_next, m));
    CU(pw_alloc(s, &w.ray_o, 5 * m)); CU(pw_alloc(s, &w.ray_d, 5 * m)); CU(pw_alloc(s, &w.side_w, 4 * m));
    CU(pw_alloc(s, &w.res, 5 * m));
    CU(pw_alloc(s, &w.queue, 5 * m)); CU(pw_alloc(s, &w.counters, (size_t)8)); CU(pw_alloc(s, &w.rays_total, (size_t)1));
    w.n_slots = n; w.tile_stride = 1; w.tile_offset = 0;
    s->pw_cap = n;
    return B200RT_OK;
}

// Node caching in L2 (opt-in, B200RT_L2_PERSIST=1): a persisting access-policy window over [8-ary nodes | triangles] on every
// stream that launches traversal kernels keeps what rays fetch resident while the wavefront state streams through L2.
// Measured on C3 (1080p, 64 spp): 1958 Mrays/s with the window vs 2129 without — the set-aside takes L2 away from the
// wavefront state that the shade pass re-reads right after the trace pass wrote it, which costs more than the BVH's
// 78 % -> higher L2 hit rate gains. Off by default for that reason.
int pin_bvh_in_l2(b200rt_scene* s, cudaStream_t extra)
{
    static const bool enabled = []() { const char* e = getenv("B200RT_L2_PERSIST"); return e && atoi(e) != 0; }();
    if (!enabled || !s->bvh_window_bytes) return B200RT_OK;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, s->device);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, s->device);
    if (max_persist <= 0 || max_window <= 0) return B200RT_OK;
    if (!s->l2_pinned)
    {
        CU(cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist));
        s->l2_pinned = true;
    }
    cudaStreamAttrValue attr;
    std::memset(&attr, 0, sizeof(attr));
    attr.accessPolicyWindow.base_ptr = s->bvh_window;
    attr.accessPolicyWindow.num_bytes = std::min(s->bvh_window_bytes, (size_t)max_window);
    // a window larger than the set-aside would thrash itself: pin the fraction that fits
    attr.accessPolicyWindow.hitRatio = (float)std::min(1.0, (double)max_persist / (double)attr.accessPolicyWindow.num_bytes);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    for (int g = 0; g < kMaxWfGroups; g++)
        if (s->wf[g].stream) CU(cudaStreamSetAttribute(s->wf[g].stream, cudaStreamAttributeAccessPolicyWindow, &attr));
    if (extra) CU(cudaStreamSetAttribute(extra, cudaStreamAttributeAccessPolicyWindow, &attr));
    return B200RT_OK;
}

// runs the selected integrator for this rank's tiles on `st`; *launches receives the number of kernels launched
int run_integrator(b200rt_scene* s, const RenderParams& P, int integrator, const float4* fb_in, float4* out_tiles, cudaStream_t st, int* launches)
{
    s->kernel_times[0] = s->kernel_times[1] = s->kernel_times[2] = s->kernel_times[3] = 0.0;
    s->timeline_valid = false;
    SceneDev dev = s->dev;
    if (P.flags & B200RT_FLAG_ENV_ALIAS)
    {
        if (!s->d_alias) return fail(B200RT_ERR_ARG, "B200RT_FLAG_ENV_ALIAS needs b200rt_scene_build_env_alias() first");
        dev.use_alias = 1;
        dev.cdf_total = s->alias_total;      // the normalisation the table was built with (the float running sum can be far off)
    }
    if (integrator == B200RT_INTEGRATOR_WAVEFRONT)
    {
        int rc = ensure_wavefront(s, P);
        if (rc) return rc;
        if (!s->l2_window_set) { if ((rc = pin_bvh_in_l2(s, nullptr))) return rc; s->l2_window_set = true; }
        unsigned int unfinished = 0;
        s->timeline_valid = (P.flags & B200RT_FLAG_TIME_INLINE) && !(P.flags & B200RT_FLAG_TIME_KERNELS);
        CU(run_wavefront(dev, P, s->wf, s->wf_groups, fb_in, out_tiles, st, s->fork_event, launches, s->kernel_times, &unfinished, &s->timeline));
        if (unfinished) return fail(B200RT_ERR_CUDA, "wavefront integrator: %u pixels unfinished after spp * (max_bounces + 1) iterations", unfinished);
        CU(wavefront_sum_rays(s->wf, s->wf_groups, s->d_rays, st));
        *launches += 1;
        return B200RT_OK;
    }
    if (integrator == B200RT_INTEGRATOR_PERSISTENT)
    {
        if (!dev.has_wide) return fail(B200RT_ERR_ARG, "the persistent integrator traverses the 8-ary BVH (leaves of at most 3 triangles)");
        int rc = ensure_persistent(s);
        if (rc) return rc;
        CU(run_persistent(dev, P, s->pw, s->pw_grid, fb_in, out_tiles, st));
        CU(cudaMemcpyAsync(s->d_rays, s->pw.rays_total, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
        *launches = 1;
        return B200RT_OK;
    }
    CU(launch_megakernel(dev, P, fb_in, out_tiles, s->d_work, s->d_rays, st));
    *launches = 1;
    return B200RT_OK;
}

// per-kernel timing fields of the stats struct (the frame must have completed)
int fill_kernel_times(b200rt_scene* s, b200rt_stats* stats)
{
    stats->trace_ms = s->kernel_times[0]; stats->shade_ms = s->kernel_times[1];
    stats->trace_launches = (int)s->kernel_times[2]; stats->shade_launches = (int)s->kernel_times[3];
    stats->trace_union_ms = 0.0; stats->tail_ms = 0.0; stats->tail_launches = 0;
    if (s->timeline_valid)
    {
        WfTimelineSummary t;
        CU(wavefront_timeline_summary(s->timeline, &t));
        stats->trace_ms = t.trace_ms; stats->shade_ms = t.shade_ms;
        stats->trace_launches = t.trace_launches; stats->shade_launches = t.shade_launches;
        stats->trace_union_ms = t.trace_union_ms;
        stats->tail_ms = t.tail_ms; stats->tail_launches = t.tail_launches;
    }
    return B200RT_OK;
}

int convert_materials(const float* materials10, int n, const std::vector<char>& used, std::vector<MaterialDev>& out, int* any_emissive)
{
    out.resize((size_t)std::max(n, 1));
    *any_emissive = 0;
    for (int i = 0; i < n; i++)
    {
        const float* m = materials10 + 10 * (size_t)i;
        MaterialDev& d = out[i];
        d.er = m[0]; d.eg = m[1]; d.eb = m[2];
        d.dr = m[4]; d.dg = m[5]; d.db = m[6];
        d.metalness = m[8]; d.roughness = m[9];
        const bool referenced = i >= (int)used.size() || used[i];
        if (referenced && (m[0] > 0.0f || m[1] > 0.0f || m[2] > 0.0f)) *any_emissive = 1;
    }
    return B200RT_OK;
}

} // namespace

extern "C" {

const char* b200rt_last_error(void) { return g_error.c_str(); }
const char* b200rt_version(void) { return "b200rt 0.1 (sm_100a)"; }

void b200rt_bvh_default_options(b200rt_bvh_options* o)
{
    if (!o) return;
    o->max_leaf_size = 3; o->sah_bins = 16; o->use_diag_slabs = 1; o->num_threads = 0;
}

void b200rt_default_render_options(b200rt_render_options* o)
{
    if (!o) return;
    o->integrator = B200RT_INTEGRATOR_WAVEFRONT; o->flags = 0; o->rank = 0; o->world = 1;
}

int b200rt_bvh_build(const float* tri_xyz9, int n_tri, const b200rt_bvh_options* opts, b200rt_bvh** out)
{
    if (!out) return fail(B200RT_ERR_ARG, "out must not be NULL");
    *out = nullptr;
    if (n_tri < 0 || (n_tri > 0 && !tri_xyz9)) return fail(B200RT_ERR_ARG, "bad triangle buffer");
    if (n_tri >= (1 << 27)) return fail(B200RT_ERR_ARG, "at most 2^27-1 triangles (leaf references pack first<<4|count)");
    b200rt_bvh_options o;
    b200rt_bvh_default_options(&o);
    if (opts) o = *opts;
    b200rt_bvh* b = new (std::nothrow) b200rt_bvh;
    if (!b) return fail(B200RT_ERR_ALLOC, "out of host memory");
    try { build_flat_bvh(tri_xyz9, n_tri, o, b->flat); }
    catch (const std::exception& e) { delete b; return fail(B200RT_ERR_ALLOC, "BVH build failed: %s", e.what()); }
    if (b->flat.info.wide_max_depth > kMaxTraversalDepth || b->flat.info.max_depth > kMaxTraversalDepth) { delete b; return fail(B200RT_ERR_ARG, "BVH depth %d exceeds the traversal stack", b->flat.info.max_depth); }
    *out = b;
    return B200RT_OK;
}

int b200rt_bvh_build_device(const float* tri_xyz9, int n_tri, int device, b200rt_bvh** out)
{
    if (!out) return fail(B200RT_ERR_ARG, "out must not be NULL");
    *out = nullptr;
    if (n_tri < 0 || (n_tri > 0 && !tri_xyz9)) return fail(B200RT_ERR_ARG, "bad triangle buffer");
    if (n_tri >= (1 << 27)) return fail(B200RT_ERR_ARG, "at most 2^27-1 triangles (leaf references pack first<<4|count)");
    if (n_tri <= kWideMaxLeaf) return b200rt_bvh_build(tri_xyz9, n_tri, nullptr, out);      // a single leaf: nothing to build in parallel
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return fail(B200RT_ERR_CUDA, "no CUDA device: b200rt_bvh_build_device has no CPU fallback");
    if (device < 0) CU(cudaGetDevice(&device));
    if (device >= n_dev) return fail(B200RT_ERR_ARG, "device %d of %d", device, n_dev);
    ON_DEVICE(device);
    b200rt_bvh* b = new (std::nothrow) b200rt_bvh;
    if (!b) return fail(B200RT_ERR_ALLOC, "out of host memory");
    std::string err;
    int rc = 1;
    try { rc = build_flat_bvh_device(tri_xyz9, n_tri, device, b->flat, err); }
    catch (const std::exception& e) { err = e.what(); rc = 1; }
    if (rc == 2) { delete b; return b200rt_bvh_build(tri_xyz9, n_tri, nullptr, out); }     // radix tree deeper than the traversal stack: host SAH builder
    if (rc) { delete b; return fail(B200RT_ERR_CUDA, "device BVH build: %s", err.c_str()); }
    *out = b;
    return B200RT_OK;
}

int b200rt_bvh_get_info(const b200rt_bvh* bvh, b200rt_bvh_info* out)
{
    if (!bvh || !out) return fail(B200RT_ERR_ARG, "NULL argument");
    *out = bvh->flat.info;
    if (out->sah_cost == 0.0) out->sah_cost = sah_cost_of(bvh->flat);       // device-built trees: computed on request from the binary records
    return B200RT_OK;
}

int b200rt_bvh_get_arrays(const b200rt_bvh* bvh, const float** axis16, const float** diag16, const float** tris12)
{
    if (!bvh) return fail(B200RT_ERR_ARG, "NULL bvh");
    if (axis16) *axis16 = reinterpret_cast<const float*>(bvh->flat.axis.data());
    if (diag16) *diag16 = reinterpret_cast<const float*>(bvh->flat.diag.data());
    if (tris12) *tris12 = reinterpret_cast<const float*>(bvh->flat.tris.data());
    return B200RT_OK;
}

int b200rt_bvh_get_wide_nodes(const b200rt_bvh* bvh, const void** nodes80, int* n_nodes)
{
    if (!bvh || !nodes80 || !n_nodes) return fail(B200RT_ERR_ARG, "NULL argument");
    *nodes80 = bvh->flat.wide.data();
    *n_nodes = (int)bvh->flat.wide.size();
    return B200RT_OK;
}

int b200rt_bvh_check(const b200rt_bvh* bvh, const float* tri_xyz9, int n_tri)
{
    if (!bvh) return fail(B200RT_ERR_ARG, "NULL bvh");
    int r = check_flat_bvh(bvh->flat, tri_xyz9, n_tri);
    if (r) return fail(100 + r, "BVH invariant %d violated", r);
    return B200RT_OK;
}

void b200rt_bvh_destroy(b200rt_bvh* bvh) { delete bvh; }

} // extern "C"

namespace {

struct SceneArgs
{
    const float* tri_xyz9; int n_tri;
    const int* tri_material; int n_material_indices;
    const float* materials10; int n_materials;
    const int* emissive_tri; int n_emissive;
    const void* spheres20; int n_spheres;
    const float* env_rgba; int env_w, env_h; const float* env_cdf_or_null;
    int env_channels;       // 4 = Image (RGBA); 3 = stbi_loadf's RGB triplets, expanded on the device (utils.cpp:113-121)
};

int validate_scene_args(const SceneArgs& a)
{
    if (a.n_tri < 0 || (a.n_tri > 0 && !a.tri_xyz9)) return fail(B200RT_ERR_ARG, "bad triangle buffer");
    if (a.n_material_indices < a.n_tri || (a.n_material_indices > 0 && !a.tri_material)) return fail(B200RT_ERR_ARG, "need one material index per triangle");
    if (a.n_materials <= 0 || !a.materials10) return fail(B200RT_ERR_ARG, "need at least one material");
    if (a.n_emissive < 0 || (a.n_emissive > 0 && !a.emissive_tri)) return fail(B200RT_ERR_ARG, "bad emissive triangle list");
    if (a.n_spheres < 0 || (a.n_spheres > 0 && !a.spheres20)) return fail(B200RT_ERR_ARG, "bad sphere buffer");
    if (a.env_channels != 3 && a.env_channels != 4) return fail(B200RT_ERR_ARG, "env_channels must be 3 (RGB) or 4 (RGBA)");
    if (!a.env_rgba || a.env_w <= 0 || a.env_h <= 0)
        return fail(B200RT_ERR_ARG, "an environment map is required (the reference samples it unconditionally, render_kernel.cpp:114)");
    for (int i = 0; i < a.n_material_indices; i++)
        if (a.tri_material[i] < 0 || a.tri_material[i] >= a.n_materials) return fail(B200RT_ERR_ARG, "material index %d of primitive %d out of range", a.tri_material[i], i);
    for (int i = 0; i < a.n_emissive; i++)
        if (a.emissive_tri[i] < 0 || a.emissive_tri[i] >= a.n_tri) return fail(B200RT_ERR_ARG, "emissive triangle index out of range");
    return B200RT_OK;
}

// host-side products shared by every device replica of a scene (computed once)
struct SceneHostData
{
    std::vector<float4> bvh_and_tris; size_t n_wide4 = 0;
    std::vector<unsigned char> oct_lut;
    std::vector<int> slot_of_prim;
    std::vector<SphereDev> spheres;
};

void prepare_host_data(const SceneArgs& a, const FlatBVH& f, SceneHostData& h)
{
    // the 8-ary nodes and the triangle stream share one allocation: one L2 access-policy window covers everything a ray fetches
    const size_t nw = f.wide.size() * 5, nt = f.tris.size() * 3;
    h.bvh_and_tris.resize(std::max<size_t>(nw + nt, 1));
    if (nw) std::memcpy(h.bvh_and_tris.data(), f.wide.data(), nw * sizeof(float4));
    if (nt) std::memcpy(h.bvh_and_tris.data() + nw, f.tris.data(), nt * sizeof(float4));
    h.n_wide4 = nw;
    h.oct_lut.resize(8 * 256);
    for (int o = 0; o < 8; o++)
        for (int m = 0; m < 256; m++)
        {
            unsigned r = 0;
            for (int b = 0; b < 8; b++) if (m & (1 << b)) r |= 1u << (b ^ o);
            h.oct_lut[(size_t)o * 256 + m] = (unsigned char)r;
        }
    h.slot_of_prim.assign((size_t)std::max(a.n_tri, 1), 0);
    for (size_t i = 0; i < f.tris.size(); i++) h.slot_of_prim[f.tris[i].prim] = (int)i;
    h.spheres.resize((size_t)std::max(a.n_spheres, 1));
    for (int i = 0; i < a.n_spheres; i++) std::memcpy(&h.spheres[i], (const char*)a.spheres20 + 20 * (size_t)i, 20);
}

// uploads one replica of the scene to `device`
int create_on_device(const SceneArgs& a, const FlatBVH& f, const SceneHostData& h, int device, b200rt_scene** out)
{
    *out = nullptr;
    ON_DEVICE(device);
    b200rt_scene* s = new (std::nothrow) b200rt_scene;
    if (!s) return fail(B200RT_ERR_ALLOC, "out of host memory");
    s->device = device;
    s->info = f.info;
    s->mat_used.assign((size_t)a.n_materials, 0);
    for (int i = 0; i < a.n_material_indices; i++) { s->mat_used[a.tri_material[i]] = 1; s->max_mat_index = std::max(s->max_mat_index, a.tri_material[i]); }
    int rc = B200RT_OK;
    do
    {
        const size_t nt = f.tris.size() * 3;
        if ((rc = upload(s, h.bvh_and_tris.data(), h.n_wide4 + nt, &s->dev.wide))) break;
        s->dev.tris = s->dev.wide + h.n_wide4;
        s->bvh_window = const_cast<float4*>(s->dev.wide);
        s->bvh_window_bytes = (h.n_wide4 + nt) * sizeof(float4);
        s->dev.has_wide = f.wide.empty() ? 0 : 1;
        s->dev.qmagic = 0x43000000u;
        if ((rc = upload(s, h.oct_lut.data(), h.oct_lut.size(), &s->dev.oct_lut))) break;
        if ((rc = upload(s, reinterpret_cast<const float4*>(f.axis.data()), f.axis.size() * 4, &s->dev.axis))) break;
        if ((rc = upload(s, reinterpret_cast<const float4*>(f.diag.data()), f.diag.size() * 4, &s->dev.diag))) break;
        if ((rc = upload(s, h.slot_of_prim.data(), (size_t)a.n_tri, &s->dev.slot_of_prim))) break;
        if ((rc = upload(s, a.tri_material, (size_t)a.n_material_indices, &s->dev.mat_idx))) break;
        if ((rc = upload(s, a.emissive_tri, (size_t)a.n_emissive, &s->dev.emissive))) break;
        if ((rc = upload(s, h.spheres.data(), (size_t)a.n_spheres, &s->dev.spheres))) break;
        const size_t n_texels = (size_t)a.env_w * a.env_h;
        if (a.env_channels == 3)
        {
            float4* d_env = nullptr;
            if ((rc = device_alloc(s, &d_env, n_texels))) break;
            float* tmp = nullptr;
            cudaError_t ce = cudaMalloc(&tmp, n_texels * 3 * sizeof(float));
            if (ce == cudaSuccess) ce = cudaMemcpy(tmp, a.env_rgba, n_texels * 3 * sizeof(float), cudaMemcpyHostToDevice);
            if (ce == cudaSuccess) ce = launch_env_expand_rgb(tmp, n_texels, d_env, 0);
            if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
            cudaFree(tmp);
            if (ce != cudaSuccess) { rc = fail(B200RT_ERR_CUDA, "env RGB -> RGBA: %s", cudaGetErrorString(ce)); break; }
            s->dev.env = d_env;
        }
        else if ((rc = upload(s, reinterpret_cast<const float4*>(a.env_rgba), n_texels, &s->dev.env))) break;
        s->dev.n_tri = a.n_tri; s->dev.n_emissive = a.n_emissive; s->dev.n_spheres = a.n_spheres;
        s->dev.env_w = a.env_w; s->dev.env_h = a.env_h;
        s->dev.has_diag = f.info.has_diag_slabs;
        // env tables (K5, env_tables.cu): per-texel luminance, the running-sum CDF in the reference's serial float order
        // (Utils::compute_env_map_cdf, utils.cpp:126-142) unless the caller brings its own, and the last-column copy the row search probes
        float* d_lum = nullptr; float* d_cdf = nullptr; float* d_row = nullptr;
        if ((rc = device_alloc(s, &d_lum, n_texels)) || (rc = device_alloc(s, &d_cdf, n_texels)) || (rc = device_alloc(s, &d_row, (size_t)a.env_h))) break;
        cudaError_t ce = launch_env_luminance(s->dev.env, n_texels, d_lum, 0);
        if (ce == cudaSuccess)
        {
            if (a.env_cdf_or_null) ce = cudaMemcpy(d_cdf, a.env_cdf_or_null, n_texels * sizeof(float), cudaMemcpyHostToDevice);
            else ce = launch_env_cdf_serial(d_lum, n_texels, d_cdf, 0);
        }
        if (ce == cudaSuccess) ce = launch_env_row_cdf(d_cdf, a.env_w, a.env_h, d_row, 0);
        if (ce == cudaSuccess) ce = cudaMemcpy(&s->dev.cdf_total, d_cdf + n_texels - 1, sizeof(float), cudaMemcpyDeviceToHost);
        if (ce != cudaSuccess) { rc = fail(B200RT_ERR_CUDA, "env tables: %s", cudaGetErrorString(ce)); break; }
        s->dev.cdf = d_cdf; s->dev.row_cdf = d_row; s->d_env_lum = d_lum;
        // guide table of the CDF search (bit-identical texel choice, ~2 probes instead of log2(W) + log2(H)); B200RT_CDF_GUIDE=0 turns it off
        static const int guide_buckets = []() { const char* e = getenv("B200RT_CDF_GUIDE"); int v = e ? atoi(e) : 65536; return v < 0 ? 0 : v; }();
        if (guide_buckets > 0 && n_texels >= 64)
        {
            unsigned int* d_guide = nullptr;
            if ((rc = device_alloc(s, &d_guide, (size_t)guide_buckets + 2))) break;
            float scale = 0.0f; int built = 0;
            ce = build_env_cdf_guide(d_cdf, n_texels, s->dev.cdf_total, guide_buckets, d_guide, &scale, &built, 0);
            if (ce != cudaSuccess) { rc = fail(B200RT_ERR_CUDA, "env CDF guide: %s", cudaGetErrorString(ce)); break; }
            if (built) { s->dev.cdf_guide = d_guide; s->dev.cdf_guide_scale = scale; s->dev.cdf_guide_n = guide_buckets; }
        }
    } while (0);
    if (rc) { b200rt_scene_destroy(s); return rc; }
    rc = b200rt_scene_set_materials(s, a.materials10, a.n_materials);
    if (rc) { b200rt_scene_destroy(s); return rc; }
    *out = s;
    return B200RT_OK;
}

int create_scene(const SceneArgs& a, const b200rt_bvh* bvh_or_null, const int* devices, int n_devices, b200rt_scene** out)
{
    if (!out) return fail(B200RT_ERR_ARG, "out must not be NULL");
    *out = nullptr;
    int rc = validate_scene_args(a);
    if (rc) return rc;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0)
        return fail(B200RT_ERR_CUDA, "no CUDA device: b200rt has no CPU fallback");
    if (n_devices <= 0 || n_devices > 64) return fail(B200RT_ERR_ARG, "bad device count %d", n_devices);
    std::vector<int> devs((size_t)n_devices);
    for (int i = 0; i < n_devices; i++)
    {
        int d = devices ? devices[i] : -1;
        if (d < 0) { if (n_devices > 1) d = i; else CU(cudaGetDevice(&d)); }
        if (d >= n_dev) return fail(B200RT_ERR_ARG, "device %d of %d", d, n_dev);
        devs[i] = d;
    }

    b200rt_bvh* own = nullptr;
    const b200rt_bvh* bvh = bvh_or_null;
    if (!bvh)
    {
        // above ~5 M triangles the host SAH build dominates the whole job (12.5 s at 20 M): build on the GPU instead
        static const int device_build_threshold = []() { const char* e = getenv("B200RT_DEVICE_BUILD_MIN_TRIS"); return e ? atoi(e) : 5000000; }();
        rc = a.n_tri >= device_build_threshold ? b200rt_bvh_build_device(a.tri_xyz9, a.n_tri, devs[0], &own) : b200rt_bvh_build(a.tri_xyz9, a.n_tri, nullptr, &own);
        if (rc) return rc;
        bvh = own;
    }
    if (bvh->flat.info.n_triangles != a.n_tri) { delete own; return fail(B200RT_ERR_ARG, "BVH was built for %d triangles, scene has %d", bvh->flat.info.n_triangles, a.n_tri); }

    SceneHostData h;
    prepare_host_data(a, bvh->flat, h);
    b200rt_scene* first = nullptr;
    rc = create_on_device(a, bvh->flat, h, devs[0], &first);
    for (int i = 1; i < n_devices && !rc; i++)
    {
        b200rt_scene* r = nullptr;
        rc = create_on_device(a, bvh->flat, h, devs[i], &r);
        if (!rc) first->replicas.push_back(r);
    }
    delete own;
    if (rc) { if (first) b200rt_scene_destroy(first); return rc; }
    if (n_devices > 1)
    {
        // peer access towards rank 0's device: the per-rank tile buffers are copied GPU to GPU (NVLink) at the end of a frame
        for (b200rt_scene* r : first->replicas)
        {
            DeviceScope scope(r->device);
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, r->device, first->device) == cudaSuccess && can)
            {
                const cudaError_t pe = cudaDeviceEnablePeerAccess(first->device, 0);
                if (pe != cudaSuccess) cudaGetLastError();      // already enabled (or unsupported: the copy is then staged by the driver)
            }
        }
    }
    *out = first;
    return B200RT_OK;
}

} // namespace

extern "C" {

int b200rt_scene_create(const float* tri_xyz9, int n_tri, const int* tri_material, int n_material_indices,
                        const float* materials10, int n_materials, const int* emissive_tri, int n_emissive,
                        const void* spheres20, int n_spheres,
                        const float* env_rgba, int env_w, int env_h, const float* env_cdf_or_null,
                        const b200rt_bvh* bvh_or_null, int device, b200rt_scene** out)
{
    const SceneArgs a = { tri_xyz9, n_tri, tri_material, n_material_indices, materials10, n_materials, emissive_tri, n_emissive,
                          spheres20, n_spheres, env_rgba, env_w, env_h, env_cdf_or_null, 4 };
    return create_scene(a, bvh_or_null, &device, 1, out);
}

int b200rt_scene_create_multi(const float* tri_xyz9, int n_tri, const int* tri_material, int n_material_indices,
                              const float* materials10, int n_materials, const int* emissive_tri, int n_emissive,
                              const void* spheres20, int n_spheres,
                              const float* env_pixels, int env_channels, int env_w, int env_h, const float* env_cdf_or_null,
                              const b200rt_bvh* bvh_or_null, const int* devices_or_null, int n_devices, b200rt_scene** out)
{
    const SceneArgs a = { tri_xyz9, n_tri, tri_material, n_material_indices, materials10, n_materials, emissive_tri, n_emissive,
                          spheres20, n_spheres, env_pixels, env_w, env_h, env_cdf_or_null, env_channels };
    return create_scene(a, bvh_or_null, devices_or_null, n_devices, out);
}

int b200rt_scene_device_count(const b200rt_scene* s) { return s ? 1 + (int)s->replicas.size() : 0; }

// Vose's alias method in double over per-texel luminances: texel i is accepted with probability prob[i], otherwise alias[i] is taken
static int build_alias_table(const float* lum, size_t n, std::vector<float>& prob, std::vector<int>& alias_out, double* total_out)
{
    double total = 0.0;
    for (size_t i = 0; i < n; i++) total += lum[i] > 0.0f ? (double)lum[i] : 0.0;
    if (!(total > 0.0)) return fail(B200RT_ERR_ARG, "environment map has no luminance to sample");
    std::vector<double> q(n);
    std::vector<unsigned int> alias(n), small, large;
    small.reserve(n); large.reserve(n);
    for (size_t i = 0; i < n; i++)
    {
        q[i] = (lum[i] > 0.0f ? (double)lum[i] : 0.0) / total * (double)n;
        alias[i] = (unsigned int)i;
        (q[i] < 1.0 ? small : large).push_back((unsigned int)i);
    }
    while (!small.empty() && !large.empty())
    {
        const unsigned int lo = small.back(), hi = large.back();
        small.pop_back();
        alias[lo] = hi;
        q[hi] = (q[hi] + q[lo]) - 1.0;
        if (q[hi] < 1.0) { large.pop_back(); small.push_back(hi); }
    }
    for (unsigned int i : large) q[i] = 1.0;
    for (unsigned int i : small) q[i] = 1.0;          // numerical leftovers
    prob.resize(n); alias_out.resize(n);
    for (size_t i = 0; i < n; i++) { prob[i] = (float)q[i]; alias_out[i] = (int)alias[i]; }
    *total_out = total;
    return B200RT_OK;
}

int b200rt_env_alias_table(const float* env_rgba, int env_w, int env_h, float* prob_out, int* alias_out, double* total_out)
{
    if (!env_rgba || env_w <= 0 || env_h <= 0 || !prob_out || !alias_out) return fail(B200RT_ERR_ARG, "bad alias table arguments");
    const size_t n = (size_t)env_w * env_h;
    std::vector<float> lum(n);
    for (size_t i = 0; i < n; i++)
    {
        const float* p = env_rgba + 4 * i;
        lum[i] = (float)(0.3086 * p[0] + 0.6094 * p[1] + 0.0820 * p[2]);      // Image::luminance_of_pixel, image.h:80-85
    }
    std::vector<float> prob; std::vector<int> alias; double total = 0.0;
    int rc = build_alias_table(lum.data(), n, prob, alias, &total);
    if (rc) return rc;
    std::memcpy(prob_out, prob.data(), n * sizeof(float));
    std::memcpy(alias_out, alias.data(), n * sizeof(int));
    if (total_out) *total_out = total;
    return B200RT_OK;
}

int b200rt_scene_build_env_alias(b200rt_scene* s)
{
    if (!s) return fail(B200RT_ERR_ARG, "NULL scene");
    for (b200rt_scene* r : s->replicas) { const int rc = b200rt_scene_build_env_alias(r); if (rc) return rc; }
    if (s->d_alias) return B200RT_OK;
    ON_DEVICE(s->device);
    const size_t n = (size_t)s->dev.env_w * s->dev.env_h;
    if (!n || !s->d_env_lum) return fail(B200RT_ERR_ARG, "scene has no environment map");
    float2* d = nullptr;
    CU(cudaMalloc(&d, n * sizeof(float2)));
    double total = 0.0;
    const cudaError_t ce = build_env_alias_device(s->d_env_lum, n, d, &total, 0);
    if (ce != cudaSuccess)
    {
        cudaFree(d);
        if (ce == cudaErrorInvalidValue) { cudaGetLastError(); return fail(B200RT_ERR_ARG, "environment map has no luminance to sample"); }
        return fail(B200RT_ERR_CUDA, "alias table build failed: %s", cudaGetErrorString(ce));
    }
    // published together, only after the build succeeded
    s->d_alias = d;
    s->bytes += n * sizeof(float2);
    s->alias_total = (float)total;
    s->dev.env_alias = d;
    return B200RT_OK;
}

int b200rt_scene_get_env_alias(b200rt_scene* s, float* prob_out, int* alias_out, double* total_out)
{
    if (!s || !prob_out || !alias_out) return fail(B200RT_ERR_ARG, "NULL argument");
    if (!s->d_alias) return fail(B200RT_ERR_ARG, "no alias table: call b200rt_scene_build_env_alias() first");
    ON_DEVICE(s->device);
    const size_t n = (size_t)s->dev.env_w * s->dev.env_h;
    std::vector<float2> t(n);
    CU(cudaMemcpy(t.data(), s->d_alias, n * sizeof(float2), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < n; i++) { prob_out[i] = t[i].x; std::memcpy(&alias_out[i], &t[i].y, sizeof(int)); }
    if (total_out) *total_out = (double)s->alias_total;
    return B200RT_OK;
}

int b200rt_scene_get_env_cdf(b200rt_scene* s, float* cdf_out)
{
    if (!s || !cdf_out) return fail(B200RT_ERR_ARG, "NULL argument");
    ON_DEVICE(s->device);
    CU(cudaMemcpy(cdf_out, s->dev.cdf, (size_t)s->dev.env_w * s->dev.env_h * sizeof(float), cudaMemcpyDeviceToHost));
    return B200RT_OK;
}

int b200rt_scene_set_materials(b200rt_scene* s, const float* materials10, int n_materials)
{
    if (!s || !materials10 || n_materials <= 0) return fail(B200RT_ERR_ARG, "bad materials");
    if (n_materials <= s->max_mat_index)
        return fail(B200RT_ERR_ARG, "%d materials, but the scene's primitives reference material index %d", n_materials, s->max_mat_index);
    for (b200rt_scene* r : s->replicas) { const int rc = b200rt_scene_set_materials(r, materials10, n_materials); if (rc) return rc; }
    ON_DEVICE(s->device);
    std::vector<MaterialDev> mats;
    int any = 0;
    convert_materials(materials10, n_materials, s->mat_used, mats, &any);
    if (n_materials > s->mats_cap)
    {
        if (s->d_mats) { cudaFree(s->d_mats); s->d_mats = nullptr; }
        CU(cudaMalloc(&s->d_mats, sizeof(MaterialDev) * (size_t)n_materials));
        s->mats_cap = n_materials;
    }
    CU(cudaMemcpy(s->d_mats, mats.data(), sizeof(MaterialDev) * (size_t)n_materials, cudaMemcpyHostToDevice));
    s->dev.mats = s->d_mats;
    s->dev.n_mats = n_materials;
    s->dev.any_emissive_material = any;
    return B200RT_OK;
}

void b200rt_scene_destroy(b200rt_scene* s)
{
    if (!s) return;
    for (b200rt_scene* r : s->replicas) b200rt_scene_destroy(r);
    s->replicas.clear();
    DeviceScope scope(s->device);
    if (s->d_gather) cudaFree(s->d_gather);
    if (s->d_rgba8) cudaFree(s->d_rgba8);
    if (s->mg_stream) cudaStreamDestroy(s->mg_stream);
    s->timeline.destroy();
    for (int i = 0; i < 2; i++) { if (s->h_stage[i]) cudaFreeHost(s->h_stage[i]); if (s->stage_ev[i]) cudaEventDestroy(s->stage_ev[i]); }
    for (void* p : s->allocs) cudaFree(p);
    if (s->d_mats) cudaFree(s->d_mats);
    if (s->d_alias) cudaFree(s->d_alias);
    if (s->d_work) cudaFree(s->d_work);
    if (s->d_rays) cudaFree(s->d_rays);
    if (s->d_tiles) cudaFree(s->d_tiles);
    if (s->d_image) cudaFree(s->d_image);
    if (s->d_prim) cudaFree(s->d_prim);
    if (s->d_t) cudaFree(s->d_t);
    for (void* p : s->wf_allocs) cudaFree(p);
    for (void* p : s->pw_allocs) cudaFree(p);
    if (s->h_active)
    {
        for (int g = 0; g < kMaxWfGroups; g++)
        {
            if (s->wf[g].stream) cudaStreamDestroy(s->wf[g].stream);
            if (s->wf[g].detach_stream) cudaStreamDestroy(s->wf[g].detach_stream);
            if (s->wf[g].poll_event) cudaEventDestroy(s->wf[g].poll_event);
            if (s->wf[g].join_event) cudaEventDestroy(s->wf[g].join_event);
            if (s->wf[g].detach_ready) cudaEventDestroy(s->wf[g].detach_ready);
            if (s->wf[g].detach_join) cudaEventDestroy(s->wf[g].detach_join);
        }
        cudaEventDestroy(s->fork_event);
        cudaFreeHost(s->h_active);
    }
    if (s->ev0) cudaEventDestroy(s->ev0);
    if (s->ev1) cudaEventDestroy(s->ev1);
    delete s;
}

int b200rt_scene_get_bvh_info(const b200rt_scene* s, b200rt_bvh_info* out)
{
    if (!s || !out) return fail(B200RT_ERR_ARG, "NULL argument");
    *out = s->info;
    return B200RT_OK;
}

size_t b200rt_scene_device_bytes(const b200rt_scene* s)
{
    if (!s) return 0;
    size_t b = s->bytes;
    for (const b200rt_scene* r : s->replicas) b += r->bytes;
    return b;
}

int b200rt_tiles_for_rank(int width, int height, int rank, int world)
{
    if (width <= 0 || height <= 0 || world <= 0 || rank < 0 || rank >= world) return 0;
    const int n = ((width + kTileDim - 1) / kTileDim) * ((height + kTileDim - 1) / kTileDim);
    return (n - rank + world - 1) / world;       // tiles rank, rank + world, ... < n
}

int b200rt_render_tiles_device(b200rt_scene* s, const float* camera17, int w, int h, int spp, int bounces,
                               void* dev_tiles, const b200rt_render_options* opts, void* cuda_stream, b200rt_stats* stats)
{
    RenderParams P;
    int rc = make_params(s, camera17, w, h, spp, bounces, opts, P);
    if (rc) return rc;
    if (!dev_tiles) return fail(B200RT_ERR_ARG, "dev_tiles must not be NULL");
    if (opts && (opts->integrator < B200RT_INTEGRATOR_MEGAKERNEL || opts->integrator > B200RT_INTEGRATOR_PERSISTENT))
        return fail(B200RT_ERR_ARG, "unknown integrator %d", opts->integrator);
    ON_DEVICE(s->device);
    if ((rc = ensure_scratch(s, 0, 0, 0))) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int integrator = opts ? opts->integrator : B200RT_INTEGRATOR_WAVEFRONT;
    CU(cudaMemsetAsync(s->d_rays, 0, sizeof(unsigned long long), st));
    if (stats) CU(cudaEventRecord(s->ev0, st));
    int launches = 0;
    if ((rc = run_integrator(s, P, integrator, nullptr, (float4*)dev_tiles, st, &launches))) return rc;
    if (stats)
    {
        CU(cudaEventRecord(s->ev1, st));
        CU(cudaEventSynchronize(s->ev1));
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        std::memset(stats, 0, sizeof(*stats));
        CU(cudaMemcpy(&stats->rays, s->d_rays, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        stats->kernel_ms = ms; stats->total_ms = ms; stats->gpu_launches = launches;
        if ((rc = fill_kernel_times(s, stats))) return rc;
        // pixels of this rank that lie inside the frame
        unsigned long long px = 0;
        for (int k = 0; k < P.n_rank_tiles; k++)
        {
            int tx, ty;
            tile_xy(P.rank + k * P.world, P.tiles_x, P.tile_skew, tx, ty);
            px += (unsigned long long)std::min(kTileDim, w - tx * kTileDim) * std::min(kTileDim, h - ty * kTileDim);
        }
        stats->samples = px * (unsigned long long)spp;
    }
    return B200RT_OK;
}

int b200rt_untile_device(b200rt_scene* s, const void* dev_gathered, int tiles_per_rank_padded, int world, int w, int h,
                         void* dev_image, void* cuda_stream)
{
    if (!s || !dev_gathered || !dev_image || world <= 0 || w <= 0 || h <= 0) return fail(B200RT_ERR_ARG, "bad untile arguments");
    if (tiles_per_rank_padded < b200rt_tiles_for_rank(w, h, 0, world)) return fail(B200RT_ERR_ARG, "tiles_per_rank_padded too small");
    ON_DEVICE(s->device);
    CU(launch_untile((const float4*)dev_gathered, tiles_per_rank_padded, world, -1, w, h, (float4*)dev_image, (cudaStream_t)cuda_stream));
    return B200RT_OK;
}

} // extern "C"

namespace {

unsigned long long pixels_of_rank(const RenderParams& P, int w, int h)
{
    unsigned long long px = 0;
    for (int k = 0; k < P.n_rank_tiles; k++)
    {
        int tx, ty;
        tile_xy(P.rank + k * P.world, P.tiles_x, P.tile_skew, tx, ty);
        px += (unsigned long long)std::min(kTileDim, w - tx * kTileDim) * std::min(kTileDim, h - ty * kTileDim);
    }
    return px;
}

// One rank of a multi-GPU frame, run by its own host thread: render this device's interleaved tiles as mean radiance
// (B200RT_FLAG_LINEAR_TILES) and push them into rank 0's gather buffer with ONE device-to-device copy (NVLink peer copy).
struct RankResult { int rc = B200RT_OK; std::string error; unsigned long long rays = 0; float ms = 0.0f; int launches = 0; };

void render_rank(b200rt_scene* r, b200rt_scene* root, RenderParams P, int integrator, size_t tiles_padded, RankResult* out)
{
    auto body = [&]() -> int {
        ON_DEVICE(r->device);
        int rc = ensure_scratch(r, r == root ? 0 : tiles_padded * kTilePixels, 0, 0);
        if (rc) return rc;
        if (!r->mg_stream) CU(cudaStreamCreateWithFlags(&r->mg_stream, cudaStreamNonBlocking));
        cudaStream_t st = r->mg_stream;
        float4* dst = root->d_gather + (size_t)P.rank * tiles_padded * kTilePixels;
        float4* tiles = r == root ? dst : r->d_tiles;          // rank 0 renders straight into its slice of the gather buffer
        CU(cudaMemsetAsync(r->d_rays, 0, sizeof(unsigned long long), st));
        CU(cudaEventRecord(r->ev0, st));
        if ((rc = run_integrator(r, P, integrator, nullptr, tiles, st, &out->launches))) return rc;
        CU(cudaEventRecord(r->ev1, st));
        if (r != root)
            CU(cudaMemcpyPeerAsync(dst, root->device, tiles, r->device, (size_t)P.n_rank_tiles * kTilePixels * sizeof(float4), st));
        CU(cudaStreamSynchronize(st));
        CU(cudaEventElapsedTime(&out->ms, r->ev0, r->ev1));
        CU(cudaMemcpy(&out->rays, r->d_rays, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        return B200RT_OK;
    };
    out->rc = body();
    if (out->rc) out->error = g_error;
}

// The whole frame: framebuffer in (or Color::Black()), integrator on one or several devices, un-tile, optional RGBA8 output
// stage, result out. fb_out / rgba8_out: exactly one is non-NULL.
int render_frame(b200rt_scene* s, const float* camera17, int w, int h, int spp, int bounces, const float* fb_in, float* fb_out,
                 unsigned char* rgba8_out, int flip_y, const b200rt_render_options* opts, b200rt_stats* stats)
{
    RenderParams P;
    int rc = make_params(s, camera17, w, h, spp, bounces, opts, P);
    if (rc) return rc;
    const int integrator = opts ? opts->integrator : B200RT_INTEGRATOR_WAVEFRONT;
    if ((integrator < B200RT_INTEGRATOR_MEGAKERNEL || integrator > B200RT_INTEGRATOR_PERSISTENT)) return fail(B200RT_ERR_ARG, "unknown integrator %d", integrator);
    if (P.flags & B200RT_FLAG_LINEAR_TILES) return fail(B200RT_ERR_ARG, "B200RT_FLAG_LINEAR_TILES is for b200rt_render_tiles_device");
    const int n_dev = 1 + (int)s->replicas.size();
    if (n_dev > 1 && P.world != 1) return fail(B200RT_ERR_ARG, "a multi-GPU scene partitions the frame itself: rank/world must stay 0/1");
    auto t0 = std::chrono::high_resolution_clock::now();
    ON_DEVICE(s->device);
    const size_t image_px = (size_t)w * h;
    if ((rc = ensure_scratch(s, (size_t)std::max(P.n_rank_tiles, 1) * kTilePixels, image_px, 0))) return rc;
    if (rgba8_out && image_px * 4 > s->rgba8_cap)
    {
        if (s->d_rgba8) cudaFree(s->d_rgba8);
        s->d_rgba8 = nullptr; s->rgba8_cap = 0;
        CU(cudaMalloc(&s->d_rgba8, image_px * 4));
        s->rgba8_cap = image_px * 4;
    }
    cudaStream_t st = 0;
    unsigned long long h2d = 17 * sizeof(float), d2h = 0, rays = 0;
    int launches = 0;
    float kernel_ms = 0.0f;
    const bool fb_zero = !fb_in || (P.flags & B200RT_FLAG_FB_IS_ZERO);

    if (n_dev == 1)
    {
        // single device: the integrator tone-maps each finished pixel against the incoming framebuffer itself
        const bool upload_fb = !fb_zero || (P.world > 1 && fb_in);
        if (upload_fb) { if ((rc = copy_to_device(s, s->d_image, fb_in, image_px * sizeof(float4), st))) return rc; h2d += image_px * sizeof(float4); }
        else if (P.world > 1) CU(launch_fill_f4(s->d_image, image_px, make_float4(0.0f, 0.0f, 0.0f, 1.0f), st));
        CU(cudaMemsetAsync(s->d_rays, 0, sizeof(unsigned long long), st));
        CU(cudaEventRecord(s->ev0, st));
        if ((rc = run_integrator(s, P, integrator, fb_zero ? nullptr : s->d_image, s->d_tiles, st, &launches))) return rc;
        CU(launch_untile(s->d_tiles, P.n_rank_tiles, P.world, P.world > 1 ? P.rank : -1, w, h, s->d_image, st));
        launches += 1;
        CU(cudaEventRecord(s->ev1, st));
    }
    else
    {
        // N devices, one host thread each: interleaved tiles, scene replicated, mean radiance gathered on rank 0's device, where
        // `framebuffer += final; tone map` (render_kernel.cpp:169-180) happens against the incoming framebuffer
        const size_t tiles_padded = (size_t)b200rt_tiles_for_rank(w, h, 0, n_dev);
        const size_t need = tiles_padded * kTilePixels * (size_t)n_dev;
        if (need > s->gather_cap)
        {
            if (s->d_gather) cudaFree(s->d_gather);
            s->d_gather = nullptr; s->gather_cap = 0;
            CU(cudaMalloc(&s->d_gather, need * sizeof(float4)));
            s->gather_cap = need;
        }
        if (!fb_zero) { if ((rc = copy_to_device(s, s->d_image, fb_in, image_px * sizeof(float4), st))) return rc; h2d += image_px * sizeof(float4); }
        else CU(launch_fill_f4(s->d_image, image_px, make_float4(0.0f, 0.0f, 0.0f, 1.0f), st));
        CU(cudaStreamSynchronize(st));
        std::vector<RankResult> res((size_t)n_dev);
        std::vector<std::thread> threads;
        for (int d = 0; d < n_dev; d++)
        {
            RenderParams Pd = P;
            Pd.rank = d; Pd.world = n_dev;
            Pd.n_rank_tiles = b200rt_tiles_for_rank(w, h, d, n_dev);
            Pd.flags |= B200RT_FLAG_LINEAR_TILES;
            b200rt_scene* r = d == 0 ? s : s->replicas[(size_t)d - 1];
            if (d == 0) continue;
            threads.emplace_back(render_rank, r, s, Pd, integrator, tiles_padded, &res[(size_t)d]);
        }
        {
            RenderParams P0 = P;
            P0.rank = 0; P0.world = n_dev;
            P0.n_rank_tiles = b200rt_tiles_for_rank(w, h, 0, n_dev);
            P0.flags |= B200RT_FLAG_LINEAR_TILES;
            render_rank(s, s, P0, integrator, tiles_padded, &res[0]);
        }
        for (std::thread& t : threads) t.join();
        for (int d = 0; d < n_dev; d++)
        {
            if (res[(size_t)d].rc) { g_error = "device rank " + std::to_string(d) + ": " + res[(size_t)d].error; return res[(size_t)d].rc; }
            rays += res[(size_t)d].rays;
            launches += res[(size_t)d].launches;
            kernel_ms = std::max(kernel_ms, res[(size_t)d].ms);
        }
        CU(launch_untile_accumulate(s->d_gather, (int)tiles_padded, n_dev, w, h, s->d_image, st));
        launches += 1;
    }

    // output stage
    if (rgba8_out)
    {
        CU(launch_quantise_rgba8(s->d_image, w, h, flip_y, s->d_rgba8, st));
        launches += 1;
        if ((rc = copy_to_host(s, rgba8_out, s->d_rgba8, image_px * 4, st))) return rc;
        d2h += image_px * 4;
    }
    else
    {
        if ((rc = copy_to_host(s, fb_out, s->d_image, image_px * sizeof(float4), st))) return rc;
        d2h += image_px * sizeof(float4);
    }
    if (stats)
    {
        std::memset(stats, 0, sizeof(*stats));
        if (n_dev == 1)
        {
            CU(cudaEventElapsedTime(&kernel_ms, s->ev0, s->ev1));
            CU(cudaMemcpy(&rays, s->d_rays, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            stats->samples = pixels_of_rank(P, w, h) * (unsigned long long)spp;
        }
        else stats->samples = (unsigned long long)image_px * (unsigned long long)spp;
        stats->rays = rays;
        stats->kernel_ms = kernel_ms;
        if ((rc = fill_kernel_times(s, stats))) return rc;
        stats->gpu_launches = launches;
        stats->h2d_bytes = h2d; stats->d2h_bytes = d2h;
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
    }
    return B200RT_OK;
}

} // namespace

extern "C" {

int b200rt_render(b200rt_scene* s, const float* camera17, int w, int h, int spp, int bounces, float* fb,
                  const b200rt_render_options* opts, b200rt_stats* stats)
{
    if (!fb) return fail(B200RT_ERR_ARG, "framebuffer must not be NULL");
    return render_frame(s, camera17, w, h, spp, bounces, fb, fb, nullptr, 0, opts, stats);
}

int b200rt_render_rgba8(b200rt_scene* s, const float* camera17, int w, int h, int spp, int bounces, const float* fb_in_or_null,
                        int flip_y, unsigned char* out_rgba8, const b200rt_render_options* opts, b200rt_stats* stats)
{
    if (!out_rgba8) return fail(B200RT_ERR_ARG, "out_rgba8 must not be NULL");
    if (opts && opts->world > 1) return fail(B200RT_ERR_ARG, "b200rt_render_rgba8 renders whole frames (rank/world must stay 0/1)");
    return render_frame(s, camera17, w, h, spp, bounces, fb_in_or_null, nullptr, out_rgba8, flip_y, opts, stats);
}

int b200rt_render_region(b200rt_scene* s, const float* camera17, int w, int h, int spp, int bounces, int x0, int y0, int x1, int y1,
                         float* out_rgba, const b200rt_render_options* opts, b200rt_stats* stats)
{
    RenderParams P;
    int rc = make_params(s, camera17, w, h, spp, bounces, opts, P);
    if (rc) return rc;
    if (!out_rgba) return fail(B200RT_ERR_ARG, "out_rgba must not be NULL");
    if (x0 < 0 || y0 < 0 || x1 > w || y1 > h || x0 >= x1 || y0 >= y1) return fail(B200RT_ERR_ARG, "bad region [%d,%d) x [%d,%d) of %dx%d", x0, x1, y0, y1, w, h);
    if ((long long)(x1 - x0) * (y1 - y0) > (1ll << 30)) return fail(B200RT_ERR_ARG, "region too large");
    ON_DEVICE(s->device);
    const size_t px = (size_t)(x1 - x0) * (y1 - y0);
    if ((rc = ensure_scratch(s, px, 0, 0))) return rc;
    P.rank = 0; P.world = 1;
    P.flags &= ~B200RT_FLAG_LINEAR_TILES;
    P.rx0 = x0; P.ry0 = y0; P.rw = x1 - x0; P.rh = y1 - y0;
    cudaStream_t st = 0;
    SceneDev dev = s->dev;
    if (P.flags & B200RT_FLAG_ENV_ALIAS)
    {
        if (!s->d_alias) return fail(B200RT_ERR_ARG, "B200RT_FLAG_ENV_ALIAS needs b200rt_scene_build_env_alias() first");
        dev.use_alias = 1; dev.cdf_total = s->alias_total;
    }
    CU(cudaMemsetAsync(s->d_rays, 0, sizeof(unsigned long long), st));
    CU(cudaEventRecord(s->ev0, st));
    // one lane per pixel through the megakernel's state machine (both integrators produce identical pixels)
    CU(launch_megakernel(dev, P, nullptr, s->d_tiles, s->d_work, s->d_rays, st));
    CU(cudaEventRecord(s->ev1, st));
    if ((rc = copy_to_host(s, out_rgba, s->d_tiles, px * sizeof(float4), st))) return rc;
    if (stats)
    {
        std::memset(stats, 0, sizeof(*stats));
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        CU(cudaMemcpy(&stats->rays, s->d_rays, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        stats->samples = (unsigned long long)px * (unsigned long long)spp;
        stats->kernel_ms = ms; stats->total_ms = ms; stats->gpu_launches = 1;
        stats->h2d_bytes = 17 * sizeof(float); stats->d2h_bytes = px * sizeof(float4);
    }
    return B200RT_OK;
}

struct b200rt_accum
{
    b200rt_scene* scene = nullptr;
    float camera17[17] = {};
    int w = 0, h = 0, spp_total = 0, bounces = 0, done = 0;
    size_t n_slots = 0;                  // tile-major pixel slots of the whole frame
    uint32_t* d_rng = nullptr; float4* d_sum = nullptr; float4* d_tiles = nullptr;
};

int b200rt_accum_create(b200rt_scene* s, const float* camera17, int w, int h, int spp_total, int bounces, b200rt_accum** out)
{
    if (!out) return fail(B200RT_ERR_ARG, "out must not be NULL");
    *out = nullptr;
    if (!s || !camera17) return fail(B200RT_ERR_ARG, "scene and camera must not be NULL");
    if (w <= 0 || h <= 0 || spp_total <= 0 || bounces <= 0) return fail(B200RT_ERR_ARG, "bad accumulator arguments");
    if (!s->replicas.empty()) return fail(B200RT_ERR_ARG, "accumulators run on single-device scenes");
    ON_DEVICE(s->device);
    b200rt_accum* a = new (std::nothrow) b200rt_accum;
    if (!a) return fail(B200RT_ERR_ALLOC, "out of host memory");
    a->scene = s; a->w = w; a->h = h; a->spp_total = spp_total; a->bounces = bounces;
    std::memcpy(a->camera17, camera17, sizeof(a->camera17));
    a->n_slots = (size_t)b200rt_tiles_for_rank(w, h, 0, 1) * kTilePixels;
    cudaError_t e = cudaMalloc(&a->d_rng, a->n_slots * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&a->d_sum, a->n_slots * sizeof(float4));
    if (e == cudaSuccess) e = cudaMalloc(&a->d_tiles, a->n_slots * sizeof(float4));
    if (e == cudaSuccess) e = cudaMemset(a->d_tiles, 0, a->n_slots * sizeof(float4));
    if (e != cudaSuccess) { b200rt_accum_destroy(a); return fail(B200RT_ERR_CUDA, "accumulator allocation: %s", cudaGetErrorString(e)); }
    *out = a;
    return B200RT_OK;
}

int b200rt_accum_samples(const b200rt_accum* a) { return a ? a->done : 0; }

int b200rt_accum_add(b200rt_accum* a, int n_samples, const b200rt_render_options* opts, b200rt_stats* stats)
{
    if (!a) return fail(B200RT_ERR_ARG, "NULL accumulator");
    if (n_samples <= 0 || a->done + n_samples > a->spp_total) return fail(B200RT_ERR_ARG, "%d more samples after %d of %d", n_samples, a->done, a->spp_total);
    b200rt_scene* s = a->scene;
    RenderParams P;
    int rc = make_params(s, a->camera17, a->w, a->h, a->spp_total, a->bounces, opts, P);
    if (rc) return rc;
    if (P.world != 1) return fail(B200RT_ERR_ARG, "accumulators render whole frames (rank/world must stay 0/1)");
    const int integrator = opts ? opts->integrator : B200RT_INTEGRATOR_WAVEFRONT;
    if (integrator != B200RT_INTEGRATOR_WAVEFRONT && integrator != B200RT_INTEGRATOR_PERSISTENT)
        return fail(B200RT_ERR_ARG, "accumulators run on the wavefront or the persistent integrator");
    ON_DEVICE(s->device);
    if ((rc = ensure_scratch(s, 0, 0, 0))) return rc;
    P.flags |= B200RT_FLAG_LINEAR_TILES;
    P.sample_begin = a->done; P.sample_end = a->done + n_samples;
    P.acc_rng = a->d_rng; P.acc_sum = a->d_sum;
    cudaStream_t st = 0;
    CU(cudaMemsetAsync(s->d_rays, 0, sizeof(unsigned long long), st));
    CU(cudaEventRecord(s->ev0, st));
    int launches = 0;
    if ((rc = run_integrator(s, P, integrator, nullptr, a->d_tiles, st, &launches))) return rc;
    CU(cudaEventRecord(s->ev1, st));
    CU(cudaStreamSynchronize(st));
    a->done += n_samples;
    if (stats)
    {
        std::memset(stats, 0, sizeof(*stats));
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        CU(cudaMemcpy(&stats->rays, s->d_rays, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        stats->samples = (unsigned long long)a->w * a->h * (unsigned long long)n_samples;
        stats->kernel_ms = ms; stats->total_ms = ms; stats->gpu_launches = launches;
        if ((rc = fill_kernel_times(s, stats))) return rc;
    }
    return B200RT_OK;
}

int b200rt_accum_resolve(b200rt_accum* a, const float* fb_in, float* fb_out)
{
    if (!a || !fb_out) return fail(B200RT_ERR_ARG, "NULL argument");
    if (a->done <= 0) return fail(B200RT_ERR_ARG, "nothing accumulated yet");
    b200rt_scene* s = a->scene;
    ON_DEVICE(s->device);
    const size_t image_px = (size_t)a->w * a->h;
    int rc = ensure_scratch(s, 0, image_px, 0);
    if (rc) return rc;
    cudaStream_t st = 0;
    if (fb_in) { if ((rc = copy_to_device(s, s->d_image, fb_in, image_px * sizeof(float4), st))) return rc; }
    else CU(launch_fill_f4(s->d_image, image_px, make_float4(0.0f, 0.0f, 0.0f, 1.0f), st));
    CU(launch_untile_accumulate(a->d_tiles, (int)(a->n_slots / kTilePixels), 1, a->w, a->h, s->d_image, st));
    return copy_to_host(s, fb_out, s->d_image, image_px * sizeof(float4), st);
}

void b200rt_accum_destroy(b200rt_accum* a)
{
    if (!a) return;
    if (a->scene)
    {
        DeviceScope scope(a->scene->device);
        cudaFree(a->d_rng); cudaFree(a->d_sum); cudaFree(a->d_tiles);
    }
    delete a;
}

int b200rt_rng_stream(int x, int y, int spp, int n, uint32_t* state_out, float* floats_out)
{
    if (n < 0 || !state_out || (n > 0 && !floats_out)) return fail(B200RT_ERR_ARG, "bad rng_stream arguments");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return fail(B200RT_ERR_CUDA, "no CUDA device: b200rt has no CPU fallback");
    uint32_t* d_state = nullptr; float* d_f = nullptr;
    cudaError_t e = cudaSuccess;
    do
    {
        if ((e = cudaMalloc(&d_state, sizeof(uint32_t))) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_f, sizeof(float) * (size_t)std::max(n, 1))) != cudaSuccess) break;
        if ((e = launch_rng_stream(x, y, spp, n, d_state, d_f, 0)) != cudaSuccess) break;
        if ((e = cudaMemcpy(state_out, d_state, sizeof(uint32_t), cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if (n && (e = cudaMemcpy(floats_out, d_f, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
    } while (0);
    cudaFree(d_state); cudaFree(d_f);
    if (e != cudaSuccess) return fail(B200RT_ERR_CUDA, "rng_stream: %s", cudaGetErrorString(e));
    return B200RT_OK;
}

int b200rt_obj_load(const char* path, b200rt_obj** out)
{
    if (!path || !out) return fail(B200RT_ERR_ARG, "NULL argument");
    *out = nullptr;
    b200rt_obj* o = new (std::nothrow) b200rt_obj;
    if (!o) return fail(B200RT_ERR_ALLOC, "out of host memory");
    std::string err;
    bool ok = false;
    try { ok = load_obj(path, o->a, err); }
    catch (const std::exception& e) { err = e.what(); }
    if (!ok) { delete o; return fail(B200RT_ERR_ARG, "OBJ ingest: %s", err.c_str()); }
    *out = o;
    return B200RT_OK;
}

int b200rt_obj_get(const b200rt_obj* o, const float** tri_xyz9, int* n_tri, const int** tri_material, const float** materials10,
                   int* n_materials, const int** emissive_tri, int* n_emissive)
{
    if (!o) return fail(B200RT_ERR_ARG, "NULL obj");
    if (tri_xyz9) *tri_xyz9 = o->a.tri_xyz9.data();
    if (n_tri) *n_tri = (int)(o->a.tri_xyz9.size() / 9);
    if (tri_material) *tri_material = o->a.tri_material.data();
    if (materials10) *materials10 = o->a.materials10.data();
    if (n_materials) *n_materials = (int)(o->a.materials10.size() / 10);
    if (emissive_tri) *emissive_tri = o->a.emissive_tri.data();
    if (n_emissive) *n_emissive = (int)o->a.emissive_tri.size();
    return B200RT_OK;
}

void b200rt_obj_destroy(b200rt_obj* o) { delete o; }

int b200rt_hdr_load(const char* path, int flip_y, b200rt_hdr** out)
{
    if (!path || !out) return fail(B200RT_ERR_ARG, "NULL argument");
    *out = nullptr;
    b200rt_hdr* h = new (std::nothrow) b200rt_hdr;
    if (!h) return fail(B200RT_ERR_ALLOC, "out of host memory");
    std::string err;
    bool ok = false;
    try { ok = load_hdr(path, flip_y != 0, h->rgb, h->w, h->h, err); }
    catch (const std::exception& e) { err = e.what(); }
    if (!ok) { delete h; return fail(B200RT_ERR_ARG, "HDR ingest: %s", err.c_str()); }
    *out = h;
    return B200RT_OK;
}

int b200rt_hdr_get(const b200rt_hdr* h, const float** rgb, int* width, int* height)
{
    if (!h) return fail(B200RT_ERR_ARG, "NULL hdr");
    if (rgb) *rgb = h->rgb.data();
    if (width) *width = h->w;
    if (height) *height = h->h;
    return B200RT_OK;
}

void b200rt_hdr_destroy(b200rt_hdr* h) { delete h; }

int b200rt_env_cdf_search(b200rt_scene* s, const float* values, int n, int use_guide, int* xy_out)
{
    if (!s || n < 0 || (n > 0 && (!values || !xy_out))) return fail(B200RT_ERR_ARG, "bad env_cdf_search arguments");
    if (n == 0) return B200RT_OK;
    ON_DEVICE(s->device);
    SceneDev dev = s->dev;
    if (!use_guide) dev.cdf_guide = nullptr;
    else if (!dev.cdf_guide) return fail(B200RT_ERR_ARG, "this scene has no CDF guide table (the running sum is not monotone, or B200RT_CDF_GUIDE=0)");
    float* d_v = nullptr; int* d_xy = nullptr;
    cudaError_t e = cudaSuccess;
    do
    {
        if ((e = cudaMalloc(&d_v, sizeof(float) * (size_t)n)) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_xy, sizeof(int) * 2 * (size_t)n)) != cudaSuccess) break;
        if ((e = cudaMemcpy(d_v, values, sizeof(float) * (size_t)n, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = launch_env_cdf_search(dev, d_v, n, d_xy, 0)) != cudaSuccess) break;
        if ((e = cudaMemcpy(xy_out, d_xy, sizeof(int) * 2 * (size_t)n, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
    } while (0);
    cudaFree(d_v); cudaFree(d_xy);
    if (e != cudaSuccess) return fail(B200RT_ERR_CUDA, "env_cdf_search: %s", cudaGetErrorString(e));
    return B200RT_OK;
}

int b200rt_host_alloc(size_t bytes, void** out)
{
    if (!out) return fail(B200RT_ERR_ARG, "out must not be NULL");
    *out = nullptr;
    CU(cudaMallocHost(out, std::max<size_t>(bytes, 1)));
    return B200RT_OK;
}

void b200rt_host_free(void* p) { if (p) cudaFreeHost(p); }

int b200rt_untile_accumulate_device(b200rt_scene* s, const void* dev_gathered, int tiles_per_rank_padded, int world, int w, int h,
                                    void* dev_fb_inout, void* cuda_stream)
{
    if (!s || !dev_gathered || !dev_fb_inout || world <= 0 || w <= 0 || h <= 0) return fail(B200RT_ERR_ARG, "bad untile arguments");
    if (tiles_per_rank_padded < b200rt_tiles_for_rank(w, h, 0, world)) return fail(B200RT_ERR_ARG, "tiles_per_rank_padded too small");
    ON_DEVICE(s->device);
    CU(launch_untile_accumulate((const float4*)dev_gathered, tiles_per_rank_padded, world, w, h, (float4*)dev_fb_inout, (cudaStream_t)cuda_stream));
    return B200RT_OK;
}

int b200rt_quantise_rgba8_device(const void* dev_image_rgba, int w, int h, int flip_y, void* dev_out_rgba8, void* cuda_stream)
{
    if (!dev_image_rgba || !dev_out_rgba8 || w <= 0 || h <= 0) return fail(B200RT_ERR_ARG, "bad quantise arguments");
    CU(launch_quantise_rgba8((const float4*)dev_image_rgba, w, h, flip_y, dev_out_rgba8, (cudaStream_t)cuda_stream));
    return B200RT_OK;
}

int b200rt_quantise_rgba8(const float* image_rgba, int w, int h, int flip_y, unsigned char* out_rgba8)
{
    if (!image_rgba || !out_rgba8 || w <= 0 || h <= 0) return fail(B200RT_ERR_ARG, "bad quantise arguments");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return fail(B200RT_ERR_CUDA, "no CUDA device: b200rt has no CPU fallback");
    const size_t px = (size_t)w * h;
    float4* d_in = nullptr; void* d_out = nullptr;
    cudaError_t e = cudaSuccess;
    do
    {
        if ((e = cudaMalloc(&d_in, px * sizeof(float4))) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_out, px * 4)) != cudaSuccess) break;
        if ((e = cudaMemcpy(d_in, image_rgba, px * sizeof(float4), cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = launch_quantise_rgba8(d_in, w, h, flip_y, d_out, 0)) != cudaSuccess) break;
        if ((e = cudaMemcpy(out_rgba8, d_out, px * 4, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
    } while (0);
    cudaFree(d_in); cudaFree(d_out);
    if (e != cudaSuccess) return fail(B200RT_ERR_CUDA, "quantise_rgba8: %s", cudaGetErrorString(e));
    return B200RT_OK;
}

int b200rt_denoise(const float* image, int channels, int w, int h, float blend, int iterations, float sigma, float* out)
{
    if (!image || !out || w <= 0 || h <= 0 || (channels != 3 && channels != 4)) return fail(B200RT_ERR_ARG, "bad denoise arguments");
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0) return fail(B200RT_ERR_CUDA, "no CUDA device: b200rt has no CPU fallback");
    if (iterations <= 0) iterations = 5;
    if (iterations > 8) iterations = 8;
    if (!(sigma > 0.0f)) sigma = 0.45f;
    const size_t bytes = (size_t)w * h * channels * sizeof(float);
    float *d_in = nullptr, *d_tmp = nullptr, *d_out = nullptr;
    cudaError_t e = cudaSuccess;
    do
    {
        if ((e = cudaMalloc(&d_in, bytes)) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_tmp, bytes)) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_out, bytes)) != cudaSuccess) break;
        if ((e = cudaMemcpy(d_in, image, bytes, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = launch_denoise(d_in, channels, w, h, iterations, sigma, blend, d_tmp, d_out, 0)) != cudaSuccess) break;
        if ((e = cudaMemcpy(out, d_out, bytes, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
    } while (0);
    cudaFree(d_in); cudaFree(d_tmp); cudaFree(d_out);
    if (e != cudaSuccess) return fail(B200RT_ERR_CUDA, "denoise: %s", cudaGetErrorString(e));
    return B200RT_OK;
}

int b200rt_trace_primary_device(b200rt_scene* s, const float* camera17, int w, int h, int sample, int spp_for_seed,
                                void* dev_prim, void* dev_t, const b200rt_render_options* opts, void* cuda_stream, b200rt_stats* stats)
{
    RenderParams P;
    int rc = make_params(s, camera17, w, h, spp_for_seed, 1, opts, P);
    if (rc) return rc;
    if (!dev_prim || !dev_t) return fail(B200RT_ERR_ARG, "output buffers must not be NULL");
    if (sample > 0) return fail(B200RT_ERR_ARG, "only sample 0 (or < 0 = un-jittered) has a path-independent camera ray");
    ON_DEVICE(s->device);
    if ((rc = ensure_scratch(s, 0, 0, 0))) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (stats) CU(cudaEventRecord(s->ev0, st));
    CU(launch_primary(s->dev, P, sample, (int*)dev_prim, (float*)dev_t, st));
    if (stats)
    {
        CU(cudaEventRecord(s->ev1, st));
        CU(cudaEventSynchronize(s->ev1));
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        std::memset(stats, 0, sizeof(*stats));
        stats->kernel_ms = ms; stats->total_ms = ms; stats->gpu_launches = 1;
        unsigned long long px = 0;
        for (int k = 0; k < P.n_rank_tiles; k++)
        {
            int tx, ty;
            tile_xy(P.rank + k * P.world, P.tiles_x, P.tile_skew, tx, ty);
            px += (unsigned long long)std::min(kTileDim, w - tx * kTileDim) * std::min(kTileDim, h - ty * kTileDim);
        }
        stats->rays = px; stats->samples = px;
    }
    return B200RT_OK;
}

int b200rt_trace_primary(b200rt_scene* s, const float* camera17, int w, int h, int sample, int spp_for_seed,
                         int* prim_out, float* t_out, const b200rt_render_options* opts, b200rt_stats* stats)
{
    if (!s) return fail(B200RT_ERR_ARG, "NULL scene");
    if (!prim_out || !t_out) return fail(B200RT_ERR_ARG, "output buffers must not be NULL");
    if (w <= 0 || h <= 0) return fail(B200RT_ERR_ARG, "bad frame size");
    auto t0 = std::chrono::high_resolution_clock::now();
    ON_DEVICE(s->device);
    const size_t px = (size_t)w * h;
    int rc = ensure_scratch(s, 0, 0, px);
    if (rc) return rc;
    if (opts && opts->world > 1)
    {
        // pixels of other ranks read as "miss"; a whole-frame call writes every pixel itself
        CU(cudaMemsetAsync(s->d_prim, 0xff, px * sizeof(int), 0));
        CU(launch_fill_f32(s->d_t, px, -1.0f, 0));
    }
    rc = b200rt_trace_primary_device(s, camera17, w, h, sample, spp_for_seed, s->d_prim, s->d_t, opts, nullptr, stats);
    if (rc) return rc;
    if ((rc = copy_to_host(s, prim_out, s->d_prim, px * sizeof(int), 0))) return rc;
    if ((rc = copy_to_host(s, t_out, s->d_t, px * sizeof(float), 0))) return rc;
    if (stats)
    {
        stats->d2h_bytes = px * 8; stats->h2d_bytes = 17 * sizeof(float);
        stats->total_ms = std::chrono::duration<double, std::milli>(std::chrono::high_resolution_clock::now() - t0).count();
    }
    return B200RT_OK;
}

int b200rt_trace_rays_device(b200rt_scene* s, const void* dev_rays6, int n, int any_hit, void* dev_prim, void* dev_t, void* dev_extra8,
                             const b200rt_render_options* opts, void* cuda_stream, b200rt_stats* stats)
{
    if (!s) return fail(B200RT_ERR_ARG, "NULL scene");
    if (n < 0 || (n > 0 && (!dev_rays6 || !dev_prim || !dev_t))) return fail(B200RT_ERR_ARG, "bad ray buffers");
    ON_DEVICE(s->device);
    int rc = ensure_scratch(s, 0, 0, 0);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (stats) CU(cudaEventRecord(s->ev0, st));
    CU(launch_trace_rays(s->dev, (const float*)dev_rays6, n, any_hit, opts ? opts->flags : 0, (int*)dev_prim, (float*)dev_t, (float*)dev_extra8, st));
    if (stats)
    {
        CU(cudaEventRecord(s->ev1, st));
        CU(cudaEventSynchronize(s->ev1));
        float ms = 0.0f;
        CU(cudaEventElapsedTime(&ms, s->ev0, s->ev1));
        std::memset(stats, 0, sizeof(*stats));
        stats->kernel_ms = ms; stats->total_ms = ms; stats->gpu_launches = 1; stats->rays = (unsigned long long)n;
    }
    return B200RT_OK;
}

int b200rt_trace_rays(b200rt_scene* s, const float* rays6, int n, int any_hit, int* prim_out, float* t_out, float* extra8,
                      const b200rt_render_options* opts)
{
    if (!s) return fail(B200RT_ERR_ARG, "NULL scene");
    if (n < 0 || (n > 0 && (!rays6 || !prim_out || !t_out))) return fail(B200RT_ERR_ARG, "bad ray buffers");
    if (n == 0) return B200RT_OK;
    ON_DEVICE(s->device);
    float* d_rays6 = nullptr; int* d_prim = nullptr; float* d_t = nullptr; float* d_extra = nullptr;
    int rc = B200RT_OK;
    cudaError_t e = cudaSuccess;
    do
    {
        if ((e = cudaMalloc(&d_rays6, sizeof(float) * 6 * (size_t)n)) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_prim, sizeof(int) * (size_t)n)) != cudaSuccess) break;
        if ((e = cudaMalloc(&d_t, sizeof(float) * (size_t)n)) != cudaSuccess) break;
        if (extra8 && (e = cudaMalloc(&d_extra, sizeof(float) * 8 * (size_t)n)) != cudaSuccess) break;
        if ((e = cudaMemcpy(d_rays6, rays6, sizeof(float) * 6 * (size_t)n, cudaMemcpyHostToDevice)) != cudaSuccess) break;
        if ((e = launch_trace_rays(s->dev, d_rays6, n, any_hit, opts ? opts->flags : 0, d_prim, d_t, d_extra, 0)) != cudaSuccess) break;
        if ((e = cudaMemcpy(prim_out, d_prim, sizeof(int) * (size_t)n, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if ((e = cudaMemcpy(t_out, d_t, sizeof(float) * (size_t)n, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
        if (extra8 && (e = cudaMemcpy(extra8, d_extra, sizeof(float) * 8 * (size_t)n, cudaMemcpyDeviceToHost)) != cudaSuccess) break;
    } while (0);
    if (e != cudaSuccess) rc = fail(B200RT_ERR_CUDA, "trace_rays: %s", cudaGetErrorString(e));
    cudaFree(d_rays6); cudaFree(d_prim); cudaFree(d_t); cudaFree(d_extra);
    return rc;
}

} // extern "C"

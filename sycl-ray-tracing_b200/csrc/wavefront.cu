// K3: wavefront integrator with path regeneration (sm_100a). Compiled with -fmad=false (see pt_device.cuh).
//
// One slot per pixel of this rank's tile-major buffer. A slot carries the pixel's xorshift state through all of its
// samples, so the per-pixel RNG stream is consumed in exactly the reference's order (render_kernel.cpp:75-181) and the
// image is bit-identical to the megakernel's. Every iteration is two launches:
//
//   wf_shade : per slot — resolve the side rays of the previous surface interaction (sample_color += (light + env) *
//              throughput, :128), consume the path ray's closest hit (miss -> sky/finish sample, hit -> material fetch,
//              the four side rays of sample_light_sources / sample_environment_map with their MIS weights, the
//              continuation sample, termination tests :130-135), finish samples / pixels, regenerate camera rays, and
//              append every ray it produced to a compacted queue (grouped by ray kind per warp, so 32 consecutive queue
//              entries are the same kind of ray from neighbouring pixels).
//   wf_trace : persistent CTAs walk the queue; one ray per lane; closest-hit, any-hit or shadow traversal by ray kind.
//
// No RNG draw depends on a trace result, which is what allows a whole surface interaction — all four side
// rays and the continuation ray — to be generated in one shade pass and traced in one trace pass.
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "pt_device.cuh"

namespace b200rt {

#ifndef WF_SHADE_MIN_BLOCKS
#define WF_SHADE_MIN_BLOCKS 4
#endif
enum { WF_ALIVE = 1, WF_PENDING = 2, WF_TERMINATED = 4, WF_DONE = 8 };

__device__ __forceinline__ void wf_enqueue(unsigned int* queue, unsigned int* counter, bool pred, unsigned int entry)
{
    const unsigned int mask = __ballot_sync(0xffffffffu, pred);
    if (!mask) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    unsigned int base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned int)__popc(mask));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pred) queue[base + __popc(mask & ((1u << lane) - 1u))] = entry;
}

__device__ __forceinline__ void wf_store_ray(const WfBuffers& B, int k, int slot, v3 o, v3 d, float tmax, int kind)
{
    const size_t i = (size_t)k * B.n_slots + slot;
    B.ray_o[i] = make_float4(o.x, o.y, o.z, tmax);
    B.ray_d[i] = make_float4(d.x, d.y, d.z, __int_as_float(kind));
}

__device__ __forceinline__ void wf_start_sample(const RenderParams& P, const WfBuffers& B, int slot, int x, int y, uint32_t& rng)
{
    const float xj = ((float)x + 0.5f) + xs_float(rng) - 1.0f;      // :88-89
    const float yj = ((float)y + 0.5f) + xs_float(rng) - 1.0f;
    v3 o, d;
    camera_ray(P.cam, xj, yj, o, d);
    wf_store_ray(B, 4, slot, o, d, 0.0f, SIDE_CLOSEST_LIGHT);
}

__device__ __forceinline__ size_t wf_out_index(const WfBuffers& B, int slot)
{
    return ((size_t)(slot >> 8) * B.tile_stride + B.tile_offset) * kTilePixels + (slot & 255);
}

__device__ __forceinline__ bool wf_slot_pixel(const RenderParams& P, int slot, int& x, int& y)
{
    const int unit = slot >> 5, lane = slot & 31;
    const int k = unit >> 3, sub = unit & 7;
    const int tile_id = P.rank + k * P.world;
    const int tx = tile_id % P.tiles_x, ty = tile_id / P.tiles_x;
    x = tx * kTileDim + (sub & 1) * kPatchW + (lane & 7);
    y = ty * kTileDim + (sub >> 1) * kPatchH + (lane >> 3);
    return x < P.cam.w && y < P.cam.h;
}

__global__ void __launch_bounds__(256) wf_init(SceneDev S, RenderParams P, WfBuffers B, const float4* __restrict__ fb_in_rowmajor,
                                               float4* __restrict__ out_tiles)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    bool q_path = false;
    if (slot < B.n_slots)
    {
        int x, y;
        const bool inside = wf_slot_pixel(P, slot, x, y);
        if (!inside) { out_tiles[wf_out_index(B, slot)] = make_float4(0.0f, 0.0f, 0.0f, 0.0f); B.flags[slot] = WF_DONE; }
        else if (P.spp <= 0 || P.max_bounces <= 0)
        {
            const float n = (float)P.spp;
            out_tiles[wf_out_index(B, slot)] = pixel_output(P.flags, fb_in_rowmajor, (size_t)y * P.cam.w + x, CO(0.0f / n, 0.0f / n, 0.0f / n));
            B.flags[slot] = WF_DONE;
        }
        else
        {
            uint32_t rng = pixel_rng(x, y, P.spp);
            wf_start_sample(P, B, slot, x, y, rng);
            B.rng[slot] = rng;
            B.sample[slot] = 0;
            B.bounce[slot] = 0;
            B.final_c[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            B.sample_c[slot] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
            B.thr[slot] = make_float4(1.0f, 1.0f, 1.0f, 0.0f);
            B.flags[slot] = WF_ALIVE;
            q_path = true;
        }
    }
    const unsigned int active = __ballot_sync(0xffffffffu, q_path);
    if ((threadIdx.x & 31) == 0 && active) atomicAdd(&B.counters[2], (unsigned int)__popc(active));
    wf_enqueue(B.queue, &B.counters[3], q_path, ((unsigned int)slot << 3) | 4u);
}

// normal of an analytic sphere hit (Sphere::intersect, sphere.h:47-49), recomputed from the stored hit distance
__device__ __forceinline__ v3 wf_sphere_normal(const SceneDev& S, int prim, v3 p)
{
    for (int i = 0; i < S.n_spheres; i++)
    {
        const SphereDev s = S.spheres[i];
        if (s.prim == prim) return normalize(p - V(s.cx, s.cy, s.cz));
    }
    return V(0.0f, 0.0f, 0.0f);
}

__global__ void __launch_bounds__(256, WF_SHADE_MIN_BLOCKS) wf_shade(SceneDev S, RenderParams P, WfBuffers B, const float4* __restrict__ fb_in_rowmajor,
                                                float4* __restrict__ out_tiles, int parity)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = B.n_slots;
    unsigned int* q_count = &B.counters[3 + parity];
    if (slot == 0)
    {
        B.counters[3 + (parity ^ 1)] = 0;      // the other queue was fully consumed by the previous trace pass
        B.counters[5 + parity] = 0;            // head of the queue this pass fills
    }
    bool q_path = false, q0 = false, q1 = false, q2 = false, q3 = false, pixel_done = false;
    const int flags_in = slot < n ? B.flags[slot] : WF_DONE;
    if (!(flags_in & WF_DONE))
    {
        int x, y;
        wf_slot_pixel(P, slot, x, y);
        uint32_t rng = B.rng[slot];
        int sample = B.sample[slot], bounce = B.bounce[slot];
        float4 t4 = B.thr[slot], s4 = B.sample_c[slot];
        col throughput = CO(t4.x, t4.y, t4.z), sample_color = CO(s4.x, s4.y, s4.z);
        int flags = flags_in;
        bool finish = false;
        const bool trace_light_brdf = S.any_emissive_material || !(P.flags & B200RT_FLAG_SKIP_DEAD_RAYS);

        // A. resolve the side rays of the previous surface interaction
        if (flags & WF_PENDING)
        {
            col c[4];
#pragma unroll
            for (int k = 0; k < 4; k++)
            {
                c[k] = CO(0.0f, 0.0f, 0.0f);
                const size_t i = (size_t)k * n + slot;
                const float4 rd4 = B.ray_d[i];
                const int kind = __float_as_int(rd4.w);
                if (kind == SIDE_NONE) continue;
                const float4 w4 = B.side_w[i];
                if (kind == SIDE_CLOSEST_LIGHT)
                {
                    const float t = B.res_t[i];
                    if (t > 0.0f)
                    {
                        const float4 ro4 = B.ray_o[i];
                        SideRay sr;
                        sr.o = V(ro4.x, ro4.y, ro4.z); sr.d = V(rd4.x, rd4.y, rd4.z);
                        sr.weight = CO(w4.x, w4.y, w4.z); sr.pdf = w4.w; sr.kind = kind; sr.tmax = 0.0f;
                        Hit h;
                        h.t = t; h.prim = B.res_prim[i]; h.slot = B.res_tslot[i]; h.u = h.v = -1.0f;
                        if (h.slot < 0) h.sphere_n = wf_sphere_normal(S, h.prim, sr.o + t * sr.d);
                        c[k] = side_light_hit(S, sr, h);
                    }
                }
                else if (B.res_prim[i] == 0) c[k] = CO(w4.x, w4.y, w4.z);      // unoccluded
            }
            sample_color = sample_color + ((c[0] + c[1]) + (c[3] + c[2])) * throughput;     // light = c0+c1 (:712), env = c3+c2 (:630), :128
            const float4 tn = B.thr_next[slot];
            throughput = CO(tn.x, tn.y, tn.z);
            flags &= ~WF_PENDING;
            if (flags & WF_TERMINATED) finish = true;
        }

        // B. consume the path ray
        if (!finish && (flags & WF_ALIVE))
        {
            const size_t i = (size_t)4 * n + slot;
            const float4 ro4 = B.ray_o[i], rd4 = B.ray_d[i];
            const v3 ro = V(ro4.x, ro4.y, ro4.z), rd = V(rd4.x, rd4.y, rd4.z);
            const float t = B.res_t[i];
            if (!(t > 0.0f))
            {
                if (bounce == 0 && P.max_bounces >= 2)                           // :146-159
                    sample_color = sample_color + env_from_direction(S, rd) * throughput;
                finish = true;
            }
            else
            {
                Hit h;
                h.t = t; h.prim = B.res_prim[i]; h.slot = B.res_tslot[i]; h.u = h.v = -1.0f;
                Surface sf;
                sf.p = ro + t * rd;
                if (h.slot >= 0)
                {
                    const float4 ve1 = __ldg(S.tris + 3 * (size_t)h.slot + 1), ve2 = __ldg(S.tris + 3 * (size_t)h.slot + 2);
                    sf.n = normalize(cross(V(ve1.x, ve1.y, ve1.z), V(ve2.x, ve2.y, ve2.z)));
                }
                else sf.n = wf_sphere_normal(S, h.prim, sf.p);
                sf.view = -rd;
                sf.m = S.mats[__ldg(S.mat_idx + h.prim)];                        // :107-108
                SideRay sr;
                sr.o = sr.d = V(0.0f, 0.0f, 0.0f); sr.tmax = 0.0f; sr.pdf = 0.0f; sr.weight = CO(0.0f, 0.0f, 0.0f);
                side_light_sample(S, sf, rng, sr);
                wf_store_ray(B, 0, slot, sr.o, sr.d, sr.tmax, sr.kind);
                if (sr.kind != SIDE_NONE) { B.side_w[(size_t)0 * n + slot] = make_float4(sr.weight.r, sr.weight.g, sr.weight.b, 0.0f); q0 = true; }
                side_light_brdf(S, sf, rng, sr);
                if (sr.kind != SIDE_NONE && !trace_light_brdf) sr.kind = SIDE_NONE;
                wf_store_ray(B, 1, slot, sr.o, sr.d, 0.0f, sr.kind);
                if (sr.kind != SIDE_NONE) { B.side_w[(size_t)1 * n + slot] = make_float4(sr.weight.r, sr.weight.g, sr.weight.b, sr.pdf); q1 = true; }
                side_env_sample(S, sf, rng, sr);
                wf_store_ray(B, 2, slot, sr.o, sr.d, 0.0f, sr.kind);
                if (sr.kind != SIDE_NONE) { B.side_w[(size_t)2 * n + slot] = make_float4(sr.weight.r, sr.weight.g, sr.weight.b, 0.0f); q2 = true; }
                side_env_brdf(S, sf, rng, sr);
                wf_store_ray(B, 3, slot, sr.o, sr.d, 0.0f, sr.kind);
                if (sr.kind != SIDE_NONE) { B.side_w[(size_t)3 * n + slot] = make_float4(sr.weight.r, sr.weight.g, sr.weight.b, 0.0f); q3 = true; }

                float bpdf;
                v3 ndir = V(0.0f, 0.0f, 0.0f);
                const col brdf = ct_sample(sf.m, sf.view, sf.n, ndir, bpdf, rng);        // :123
                if (bounce == 0) sample_color = sample_color + CO(sf.m.er, sf.m.eg, sf.m.eb);
                flags = WF_PENDING;
                col tnext = throughput;
                if (is_black(brdf) || bpdf < 1.0e-8f || isinf(bpdf)) flags |= WF_TERMINATED;     // :130-135
                else
                {
                    tnext = throughput * ((brdf * smax(0.0f, dot(ndir, sf.n))) / bpdf);       // :137
                    bounce++;
                    if (bounce >= P.max_bounces) flags |= WF_TERMINATED;
                    else
                    {
                        wf_store_ray(B, 4, slot, sf.p + 1.0e-4f * sf.n, ndir, 0.0f, SIDE_CLOSEST_LIGHT);
                        flags |= WF_ALIVE;
                        q_path = true;
                    }
                }
                B.thr_next[slot] = make_float4(tnext.r, tnext.g, tnext.b, 0.0f);
            }
        }

        // C. finish the sample: next sample of this pixel, or the pixel itself
        if (finish)
        {
            float4 f4 = B.final_c[slot];
            col final_color = CO(f4.x, f4.y, f4.z) + sample_color;
            sample++;
            if (sample < P.spp)
            {
                B.final_c[slot] = make_float4(final_color.r, final_color.g, final_color.b, 0.0f);
                wf_start_sample(P, B, slot, x, y, rng);
                throughput = CO(1.0f, 1.0f, 1.0f);
                sample_color = CO(0.0f, 0.0f, 0.0f);
                bounce = 0;
                flags = WF_ALIVE;
                q_path = true;
            }
            else
            {
                const float nspp = (float)P.spp;
                const col mean = CO(final_color.r / nspp, final_color.g / nspp, final_color.b / nspp);
                out_tiles[wf_out_index(B, slot)] = pixel_output(P.flags, fb_in_rowmajor, (size_t)y * P.cam.w + x, mean);
                flags = WF_DONE;
                pixel_done = true;
            }
        }
        B.rng[slot] = rng;
        B.sample[slot] = sample;
        B.bounce[slot] = bounce;
        B.thr[slot] = make_float4(throughput.r, throughput.g, throughput.b, 0.0f);
        B.sample_c[slot] = make_float4(sample_color.r, sample_color.g, sample_color.b, 0.0f);
        B.flags[slot] = flags;
    }
    const unsigned int done_mask = __ballot_sync(0xffffffffu, pixel_done);
    if ((threadIdx.x & 31) == 0 && done_mask) atomicSub(&B.counters[2], (unsigned int)__popc(done_mask));
    const unsigned int s3 = (unsigned int)slot << 3;
    wf_enqueue(B.queue, q_count, q_path, s3 | 4u);
    wf_enqueue(B.queue, q_count, q0, s3 | 0u);
    wf_enqueue(B.queue, q_count, q1, s3 | 1u);
    wf_enqueue(B.queue, q_count, q2, s3 | 2u);
    wf_enqueue(B.queue, q_count, q3, s3 | 3u);
}

// result of one queue entry
__device__ __forceinline__ void wf_store_result(const WfBuffers& B, size_t r, int mode, bool found, const Hit& h)
{
    if (mode == TRACE_CLOSEST)
    {
        B.res_t[r] = found ? h.t : -1.0f;
        B.res_prim[r] = h.prim;
        B.res_tslot[r] = h.slot;
    }
    else B.res_prim[r] = found ? 1 : 0;
}

// ablation variant (B200RT_FLAG_SIMPLE_TRACE): one ray per lane, grid-stride over the queue (the warp waits for its slowest ray)
template <int TL>
__global__ void wf_trace_simple(SceneDev S, WfBuffers B, int parity)
{
    const unsigned int n_rays = B.counters[3 + parity];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(B.rays_total, (unsigned long long)n_rays);
    const int n = B.n_slots;
    for (unsigned int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rays; i += gridDim.x * blockDim.x)
    {
        const unsigned int e = B.queue[i];
        const int slot = (int)(e >> 3), k = (int)(e & 7u);
        const size_t r = (size_t)k * n + slot;
        const float4 ro4 = B.ray_o[r], rd4 = B.ray_d[r];
        const int kind = __float_as_int(rd4.w);
        const int mode = kind == SIDE_SHADOW ? TRACE_SHADOW : (kind == SIDE_CLOSEST_LIGHT ? TRACE_CLOSEST : TRACE_ANY);
        Hit h;
        const bool found = trace_ray<TL>(S, V(ro4.x, ro4.y, ro4.z), V(rd4.x, rd4.y, rd4.z), ro4.w, mode, h);
        wf_store_result(B, r, mode, found, h);
    }
}

// default variant, persistent: every lane owns one traversal state; whenever at least kRefillThreshold lanes of a warp are idle
// they are refilled with new rays from the warp's private block of the queue (blocks of kWarpBlock rays are claimed
// with one atomic), so no lane waits for the slowest ray of its warp (Aila & Laine's persistent threads, per lane).
#ifndef WF_WARP_BLOCK
#define WF_WARP_BLOCK 64
#endif
#ifndef WF_REFILL
#define WF_REFILL 8
#endif
#ifndef WF_LEAF_MIN
#define WF_LEAF_MIN 4
#endif
#ifdef WF_TRACE_MIN_BLOCKS
#define WF_TRACE_BOUNDS __launch_bounds__(128, WF_TRACE_MIN_BLOCKS)
#else
#define WF_TRACE_BOUNDS
#endif
constexpr int kWarpBlock = WF_WARP_BLOCK;
constexpr int kRefillThreshold = WF_REFILL;
constexpr int kLeafThreshold = WF_LEAF_MIN;

// one interface over the binary and the 8-ary traversal state machines, so the persistent kernel below serves every layout
template <int TL>
struct TravOps
{
    typedef Trav State;
    typedef TravStack Stack;
    static __device__ __forceinline__ void init(State& T, v3 o, v3 d, float tmax, int mode) { trav_init<TL == TL_DIAG>(T, o, d, tmax, mode); }
    static __device__ __forceinline__ bool at_node(const State& T) { return T.cur >= 0; }
    static __device__ __forceinline__ void node(const SceneDev& S, State& T, Stack& K) { trav_inner<TL == TL_DIAG>(S, T, K); }
    static __device__ __forceinline__ void leaf(const SceneDev& S, State& T, Stack& K) { trav_leaf(S, T, K); }
};
template <>
struct TravOps<TL_WIDE>
{
    typedef Trav8 State;
    typedef TravStack8 Stack;
    static __device__ __forceinline__ void init(State& T, v3 o, v3 d, float tmax, int mode) { trav8_init(T, o, d, tmax, mode); }
    static __device__ __forceinline__ bool at_node(const State& T) { return trav8_has_node(T); }
    static __device__ __forceinline__ void node(const SceneDev& S, State& T, Stack& K) { trav8_node(S, T, K); }
    static __device__ __forceinline__ void leaf(const SceneDev& S, State& T, Stack& K) { trav8_tris(S, T, K); }
};

template <int TL>
__global__ void WF_TRACE_BOUNDS wf_trace(SceneDev S, WfBuffers B, int parity)
{
    typedef TravOps<TL> Ops;
    const unsigned int FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned int n_rays = B.counters[3 + parity];
    unsigned int* head = &B.counters[5 + parity];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(B.rays_total, (unsigned long long)n_rays);
    const int n = B.n_slots;
    const bool spheres = S.n_spheres != 0;
    // rays claimed per atomic: kWarpBlock when the queue is long, down to one warp's worth when it is short (otherwise a
    // few warps would serialise a short queue while the rest of the machine idles)
    const unsigned int total_warps = gridDim.x * (blockDim.x >> 5);
    const unsigned int warp_block = min((unsigned int)kWarpBlock, max(32u, (n_rays / total_warps) & ~31u));

    typename Ops::State T;
    typename Ops::Stack K;
    T.done = true;
    bool active = false;
    size_t r = 0;
    int mode = TRACE_CLOSEST;
    unsigned int blk_next = 0, blk_end = 0;       // warp-uniform: the warp's private block of queue entries
    bool exhausted = false;                       // warp-uniform: the queue has no more blocks

    for (;;)
    {
        unsigned int idle = __ballot_sync(FULL, !active);
        if (idle)
        {
            if (!exhausted && (idle == FULL || __popc(idle) >= kRefillThreshold))
            {
                if (blk_next >= blk_end)
                {
                    unsigned int b = 0;
                    if (lane == 0) b = atomicAdd(head, warp_block);
                    b = __shfl_sync(FULL, b, 0);
                    if (b >= n_rays) exhausted = true;
                    else { blk_next = b; blk_end = min(b + warp_block, n_rays); }
                }
                if (blk_next < blk_end)
                {
                    const unsigned int idx = blk_next + __popc(idle & ((1u << lane) - 1u));
                    if (!active && idx < blk_end)
                    {
                        const unsigned int e = B.queue[idx];
                        const int slot = (int)(e >> 3), k = (int)(e & 7u);
                        r = (size_t)k * n + slot;
                        const float4 ro4 = B.ray_o[r], rd4 = B.ray_d[r];
                        const int kind = __float_as_int(rd4.w);
                        mode = kind == SIDE_SHADOW ? TRACE_SHADOW : (kind == SIDE_CLOSEST_LIGHT ? TRACE_CLOSEST : TRACE_ANY);
                        Ops::init(T, V(ro4.x, ro4.y, ro4.z), V(rd4.x, rd4.y, rd4.z), ro4.w, spheres ? TRACE_CLOSEST : mode);
                        active = true;
                    }
                    blk_next = min(blk_next + (unsigned int)__popc(idle), blk_end);
                }
                idle = __ballot_sync(FULL, !active);
            }
        }
        // warp-level phase vote: all lanes that hold an inner node step together, or all lanes that hold a leaf run the
        // exact triangle tests together, whichever is the majority — the two kinds of work never interleave inside a warp
        const bool has_inner = active && Ops::at_node(T), has_leaf = active && !Ops::at_node(T);
        const unsigned int m_inner = __ballot_sync(FULL, has_inner), m_leaf = __ballot_sync(FULL, has_leaf);
        if (!(m_inner | m_leaf))
        {
            if (exhausted) break;
            continue;
        }
        if (m_inner && __popc(m_leaf) < max(__popc(m_inner), kLeafThreshold)) { if (has_inner) Ops::node(S, T, K); }
        else if (has_leaf) Ops::leaf(S, T, K);
        if (active && T.done)
        {
            bool found = T.hit.t > 0.0f;
            if (spheres) found = finish_with_spheres(S, T.o, T.d, T.tmax, mode, T.hit);
            wf_store_result(B, r, mode, found, T.hit);
            active = false;
        }
    }
}

// ---- default trace kernel for the 8-ary layout: persistent lanes + warp-cooperative triangle tests --------------------------------
// ncu on the phase-vote kernel above (profiles/r1b_*): the node step ran with 20.7 of 32 lanes, but the exact triangle
// tests with only 7.3 — few lanes of a warp hold a leaf at the same time, and those that do hold different numbers of
// triangles — and that phase took 45 % of the stall samples. Here a lane never tests its own triangles. The node step
// leaves a lane's hit triangles as (owner lane, triangle slot) pairs in a per-warp shared-memory queue and the lane goes
// straight on with its next node; whenever the queue holds 32 pairs, all 32 lanes test one pair each (ray from the owner's
// shared-memory record, Moller-Trumbore exactly as triangle.h:16-60) and fold the result into the owner's 64-bit key
// (t bits << 32 | primitive index) with a shared-memory atomicMin — which is exactly the closest-hit rule of the other
// kernels (smaller t, ties to the lower original index). Owners pick up their new t_best after each drain. A ray is
// finished when it has no node work and no pair left in the queue.
#ifndef WF_COOP_QUEUE
#define WF_COOP_QUEUE 160
#endif
constexpr int kCoopQueue = WF_COOP_QUEUE;   // pairs per warp (> 32); an append that does not fit drains first (a build with 40 runs the GPU
                                            // test suite through that path all the time)
constexpr int kCoopMaxWarps = 4;         // the kernel is launched with 128-thread CTAs
constexpr unsigned long long kNoHit = 0xffffffffffffffffull;

struct CoopWarp
{
    float4 ro[32], rd[32];              // per lane: origin + tmax, direction + trace mode
    unsigned long long best[32];        // per lane: (t bits << 32) | primitive index of the best accepted hit
    unsigned int q[kCoopQueue];         // (owner lane << 27) | triangle slot
    unsigned int pend[32];
};

// all 32 lanes: test whole batches of 32 pairs (and the last partial one if `full`), compact the rest to the queue's front,
// tell every lane whether it still owns a queued pair
__device__ __forceinline__ void coop_drain(const SceneDev& S, CoopWarp& W, const int lane, unsigned int& qcount, const bool full, bool& pending)
{
    unsigned int base = 0;
    while (qcount - base >= 32u || (full && base < qcount))
    {
        const unsigned int nb = min(32u, qcount - base);
        if ((unsigned int)lane < nb)
        {
            const unsigned int e = W.q[base + lane];
            const unsigned int owner = e >> 27, slot = e & 0x07ffffffu;
            const float4 ro4 = W.ro[owner], rd4 = W.rd[owner];
            const float4* tp = S.tris + 3 * (size_t)slot;
            const float4 va = __ldg(tp), ve1 = __ldg(tp + 1), ve2 = __ldg(tp + 2);
            float t, u, v;
            if (tri_test(va, ve1, ve2, V(ro4.x, ro4.y, ro4.z), V(rd4.x, rd4.y, rd4.z), t, u, v))
            {
                const int mode = __float_as_int(rd4.w);
                if (mode != TRACE_SHADOW || t + 1.0e-4f < ro4.w)
                {
                    const unsigned long long key = ((unsigned long long)__float_as_uint(t) << 32) | (unsigned int)__float_as_int(va.w);
                    if (key < W.best[owner]) atomicMin(&W.best[owner], key);
                }
            }
        }
        base += nb;
    }
    __syncwarp();
    const unsigned int rem = qcount - base;
    const unsigned int e = (unsigned int)lane < rem ? W.q[base + lane] : 0u;
    W.pend[lane] = 0u;
    __syncwarp();
    if ((unsigned int)lane < rem) { W.q[lane] = e; W.pend[e >> 27] = 1u; }
    __syncwarp();
    pending = W.pend[lane] != 0u;
    qcount = rem;
}

#ifndef WF_COOP_MIN_BLOCKS
#define WF_COOP_MIN_BLOCKS 7            // <= 72 registers: 28 warps per SM; 32 (64 registers) measured the same, 16 (117 registers) 20 % slower
#endif
__global__ void __launch_bounds__(32 * kCoopMaxWarps, WF_COOP_MIN_BLOCKS) wf_trace_coop(SceneDev S, WfBuffers B, int parity)
{
    __shared__ CoopWarp s_warps[kCoopMaxWarps];
    CoopWarp& W = s_warps[threadIdx.x >> 5];
    const unsigned int FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const unsigned int lanes_below = (1u << lane) - 1u;
    const unsigned int n_rays = B.counters[3 + parity];
    unsigned int* head = &B.counters[5 + parity];
    if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(B.rays_total, (unsigned long long)n_rays);
    const int n = B.n_slots;
    const bool spheres = S.n_spheres != 0;
    const unsigned int total_warps = gridDim.x * (blockDim.x >> 5);
    const unsigned int warp_block = min((unsigned int)kWarpBlock, max(32u, (n_rays / total_warps) & ~31u));

    Trav8 T;
#ifdef WF_COOP_LOCAL_STACK
    TravStack8 K;
#else
    __shared__ uint2 s_stack[kSharedStackDepth * 32 * kCoopMaxWarps];
    TravStack8Shared K;
    K.sh = s_stack + threadIdx.x; K.stride = 32 * kCoopMaxWarps;
#endif
    T.done = true; T.tg = make_uint2(0u, 0u);
    bool active = false, pending = false;
    size_t r = 0;
    int mode = TRACE_CLOSEST;
    unsigned int blk_next = 0, blk_end = 0, qcount = 0;       // warp-uniform
    bool exhausted = false;                                   // warp-uniform

    for (;;)
    {
        // (1) refill idle lanes from the warp's private block of the ray queue
        unsigned int idle = __ballot_sync(FULL, !active);
        if (idle && !exhausted && (idle == FULL || __popc(idle) >= kRefillThreshold))
        {
            if (blk_next >= blk_end)
            {
                unsigned int b = 0;
                if (lane == 0) b = atomicAdd(head, warp_block);
                b = __shfl_sync(FULL, b, 0);
                if (b >= n_rays) exhausted = true;
                else { blk_next = b; blk_end = min(b + warp_block, n_rays); }
            }
            if (blk_next < blk_end)
            {
                const unsigned int idx = blk_next + __popc(idle & lanes_below);
                if (!active && idx < blk_end)
                {
                    const unsigned int e = B.queue[idx];
                    const int slot = (int)(e >> 3), k = (int)(e & 7u);
                    r = (size_t)k * n + slot;
                    const float4 ro4 = B.ray_o[r], rd4 = B.ray_d[r];
                    const int kind = __float_as_int(rd4.w);
                    mode = kind == SIDE_SHADOW ? TRACE_SHADOW : (kind == SIDE_CLOSEST_LIGHT ? TRACE_CLOSEST : TRACE_ANY);
                    const int tmode = spheres ? TRACE_CLOSEST : mode;
                    trav8_init(T, V(ro4.x, ro4.y, ro4.z), V(rd4.x, rd4.y, rd4.z), ro4.w, tmode);
                    W.ro[lane] = ro4;
                    W.rd[lane] = make_float4(rd4.x, rd4.y, rd4.z, __int_as_float(tmode));
                    W.best[lane] = kNoHit;
                    active = true; pending = false;
                }
                blk_next = min(blk_next + (unsigned int)__popc(idle), blk_end);
            }
            __syncwarp();
        }
        const unsigned int m_active = __ballot_sync(FULL, active);
        if (!m_active)
        {
            if (exhausted) break;
            continue;
        }
        // (2) one node step for every lane that has one
        if (active && !T.done)
        {
            trav8_node(S, T, K);
            // the next node group can be fetched from the stack right away (its load overlaps the queue housekeeping below);
            // the pending triangle group lives in T.tg / T.tvalid, which the pop does not touch
            if (!T.done && !(T.ng.y & 0xff000000u)) trav8_pop(T, K);
        }
        // (3) hit triangles -> the warp's pair queue (draining first when it would overflow); the lane moves on
        for (;;)
        {
            const unsigned int cnt = (active ? __popc(T.tg.y) : 0u);
            if (!__any_sync(FULL, cnt != 0u)) break;            // node steps near the root hit no leaf child at all: no scan needed
            unsigned int incl = cnt;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const unsigned int up = __shfl_up_sync(FULL, incl, o);
                if (lane >= o) incl += up;
            }
            const unsigned int total = __shfl_sync(FULL, incl, 31);
            if (!total) break;
            const unsigned int room = kCoopQueue - qcount;
            unsigned int at = incl - cnt;
            if (cnt)
            {
                unsigned int bits = T.tg.y;
                while (bits && at < room)
                {
                    const int b = __ffs((int)bits) - 1;
                    bits &= bits - 1u;
                    W.q[qcount + at] = ((unsigned int)lane << 27) | (unsigned int)trav8_tri_slot(T, b);
                    at++;
                }
                T.tg.y = bits;
                pending = true;
            }
            __syncwarp();
            qcount += min(total, room);
            if (total <= room) break;
            coop_drain(S, W, lane, qcount, false, pending);
        }
        // (4) test queued pairs: whole batches of 32; everything when no lane has node work left or enough lanes wait
        const unsigned int m_node = __ballot_sync(FULL, active && !T.done);
        const unsigned int m_wait = __ballot_sync(FULL, active && T.done && pending);
        const bool full = !m_node || __popc(m_wait) >= kRefillThreshold;
        if (qcount >= 32u || (full && qcount))
        {
            coop_drain(S, W, lane, qcount, full, pending);
            if (active)
            {
                const unsigned long long key = W.best[lane];
                if (key != kNoHit)
                {
                    if (T.mode == TRACE_CLOSEST) T.tbest = __uint_as_float((unsigned int)(key >> 32));
                    else T.done = true;           // occlusion established; the lane only waits for its queued pairs to leave
                }
            }
        }
        // (5) finished rays
        if (active && T.done && !pending)
        {
            const unsigned long long key = W.best[lane];
            Hit h;
            bool found = key != kNoHit;
            h.t = found ? __uint_as_float((unsigned int)(key >> 32)) : -1.0f;
            h.prim = found ? (int)(unsigned int)key : -1;
            h.slot = (found && T.mode == TRACE_CLOSEST) ? __ldg(S.slot_of_prim + h.prim) : -1;
            h.u = h.v = -1.0f;
            if (spheres)
            {
                const float4 ro4 = W.ro[lane], rd4 = W.rd[lane];
                found = finish_with_spheres(S, V(ro4.x, ro4.y, ro4.z), V(rd4.x, rd4.y, rd4.z), ro4.w, mode, h);
            }
            wf_store_result(B, r, mode, found, h);
            active = false;
        }
    }
}

// ---- host driver ---------------------------------------------------------------------------------------------------------------
// The frame's pixels are split into n_groups interleaved tile groups (group g of rank r behaves like rank r + g * world of
// a world * n_groups partition). Each group runs its own trace/shade iteration chain on its own stream, so while one
// group's trace pass drains its longest rays (a single ray is a dependent chain of node fetches; the slowest ray of a
// pass bounds that pass) the other groups' kernels fill the machine. Groups never exchange data.
cudaError_t run_wavefront(const SceneDev& S, const RenderParams& P, const WfGroup* groups, int n_groups, const float4* fb_in_rowmajor,
                          float4* out_tiles, cudaStream_t stream, cudaEvent_t fork_event, int* launches_out, double* kernel_times4,
                          unsigned int* unfinished_out)
{
    const int n_sm = current_sm_count();
    int launches = 0;
    const int tl = traversal_layout(S, P.flags);
    const bool simple = (P.flags & B200RT_FLAG_SIMPLE_TRACE) != 0;
    static const int tb_env = []() { const char* e = getenv("B200RT_TRACE_BLOCK"); int v = e ? atoi(e) : 128; return (v == 64 || v == 128 || v == 256) ? v : 128; }();
    int per_sm = 0;
    typedef void (*TraceKernel)(SceneDev, WfBuffers, int);
    static const TraceKernel kernels[2][3] = { { wf_trace<TL_AXIS>, wf_trace<TL_DIAG>, wf_trace<TL_WIDE> },
                                               { wf_trace_simple<TL_AXIS>, wf_trace_simple<TL_DIAG>, wf_trace_simple<TL_WIDE> } };
    static const bool vote_wide = []() { const char* e = getenv("B200RT_WIDE_TRACE"); return e && e[0] == 'v'; }();   // ablation: phase-vote kernel on the 8-ary layout
    const bool coop = tl == TL_WIDE && !simple && !vote_wide;
    const TraceKernel trace_kernel = coop ? wf_trace_coop : kernels[simple ? 1 : 0][tl];
    const int tb = coop ? 32 * kCoopMaxWarps : tb_env;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, trace_kernel, tb, 0);
    if (per_sm <= 0) per_sm = 1;
    // With several tile groups in flight a persistent trace kernel takes only ~4/7 of the CTAs that fit (B200RT_TRACE_GRID_PCT,
    // default 58): the rest of every SM stays free for the other groups' trace / shade kernels, which then overlap with the
    // whole pass instead of only with its tail (C3, 32 spp: 2 179 vs 2 108 Mrays/s; 72 %: 2 157, 43 %: 2 112).
    static const int grid_pct_env = []() { const char* e = getenv("B200RT_TRACE_GRID_PCT"); int v = e ? atoi(e) : 0; return v > 100 ? 100 : v; }();
    const bool timing_mode = getenv("B200RT_WF_TIMING") != nullptr || (P.flags & B200RT_FLAG_TIME_KERNELS);      // kernels run one at a time: full grid
    const int grid_pct = grid_pct_env >= 10 ? grid_pct_env : ((n_groups > 1 && !timing_mode) ? 58 : 100);
    const int trace_grid = n_sm * std::max(1, (per_sm * grid_pct + 50) / 100);
    static const bool timing_env = getenv("B200RT_WF_TIMING") != nullptr;  // diagnostics: per-kernel times on stderr (serialises the groups)
    const bool timing = timing_env || (P.flags & B200RT_FLAG_TIME_KERNELS);
    int n_trace = 0, n_shade = 0;
    cudaError_t e;

    // fork: every group stream starts after the work already queued on the caller's stream. From here on an error does not
    // return: it ends the launching and falls through to the join, so that no group stream is left forked.
    if ((e = cudaEventRecord(fork_event, stream)) != cudaSuccess) return e;
    cudaError_t err = cudaSuccess;
#define WF_TRY(expr) do { if (err == cudaSuccess) { const cudaError_t e__ = (expr); if (e__ != cudaSuccess) err = e__; } } while (0)
    struct GroupRun { RenderParams P; int parity; bool finished, forked; int slot_grid; long long it, poll_it; bool poll_pending; };
    GroupRun run[kMaxWfGroups];
    for (int g = 0; g < n_groups; g++)
    {
        const WfGroup& G = groups[g];
        GroupRun& R = run[g];
        R.P = P;
        R.P.rank = P.rank + g * P.world;
        R.P.world = P.world * n_groups;
        R.P.n_rank_tiles = G.buf.n_slots / kTilePixels;
        R.parity = 0; R.it = 0; R.poll_it = 0; R.poll_pending = false; R.forked = false;
        R.slot_grid = (G.buf.n_slots + 255) / 256;
        R.finished = R.slot_grid <= 0;
        if (err != cudaSuccess) { R.finished = true; continue; }
        WF_TRY(cudaStreamWaitEvent(G.stream, fork_event, 0));
        if (err != cudaSuccess) { R.finished = true; continue; }
        R.forked = true;
        // an empty group (fewer tiles than groups) still gets its counters cleared: the frame's ray count sums every group
        WF_TRY(cudaMemsetAsync(G.buf.counters, 0, 8 * sizeof(unsigned int), G.stream));
        WF_TRY(cudaMemsetAsync(G.buf.rays_total, 0, sizeof(unsigned long long), G.stream));
        if (R.finished) continue;
        wf_init<<<R.slot_grid, 256, 0, G.stream>>>(S, R.P, G.buf, fb_in_rowmajor, out_tiles);
        launches++;
    }
    const long long max_iters = (long long)P.spp * (P.max_bounces + 1) + 2;
    cudaEvent_t tev[3] = { nullptr, nullptr, nullptr };
    double t_trace = 0.0, t_shade = 0.0;
    if (timing) for (int i = 0; i < 3; i++) cudaEventCreate(&tev[i]);
    int remaining = 0;
    unsigned int unfinished = 0;
    for (int g = 0; g < n_groups; g++) remaining += run[g].finished ? 0 : 1;
    while (remaining > 0 && err == cudaSuccess)
    {
        for (int g = 0; g < n_groups && err == cudaSuccess; g++)
        {
            GroupRun& R = run[g];
            if (R.finished) continue;
            const WfGroup& G = groups[g];
            // has an earlier poll of this group's unfinished-pixel counter landed?
            if (R.poll_pending && cudaEventQuery(G.poll_event) == cudaSuccess)
            {
                R.poll_pending = false;
                if (*G.host_active == 0) { R.finished = true; remaining--; continue; }
            }
            // bounded run-ahead: at most kRunAhead iterations queued behind an unanswered poll (a finished group would
            // otherwise leave a long train of empty launches behind)
            if (R.poll_pending && R.it - R.poll_it >= 8) continue;
            if (R.it >= max_iters)
            {
                // safety net: spp * (max_bounces + 1) iterations finish every pixel. Wait for the last poll; pixels still
                // unfinished now are a bug and are reported to the caller, never dropped silently.
                unsigned int left = 0;
                WF_TRY(cudaStreamSynchronize(G.stream));
                WF_TRY(cudaMemcpy(&left, &G.buf.counters[2], sizeof(left), cudaMemcpyDeviceToHost));
                unfinished += left;
                R.finished = true; remaining--;
                continue;
            }
            if (timing) cudaEventRecord(tev[0], G.stream);
            trace_kernel<<<trace_grid, tb, 0, G.stream>>>(S, G.buf, R.parity);
            if (timing) cudaEventRecord(tev[1], G.stream);
            R.parity ^= 1;
            wf_shade<<<R.slot_grid, 256, 0, G.stream>>>(S, R.P, G.buf, fb_in_rowmajor, out_tiles, R.parity);
            launches += 2;
            WF_TRY(cudaGetLastError());
            if (timing)
            {
                cudaEventRecord(tev[2], G.stream);
                WF_TRY(cudaEventSynchronize(tev[2]));
                float a = 0, b = 0;
                cudaEventElapsedTime(&a, tev[0], tev[1]); cudaEventElapsedTime(&b, tev[1], tev[2]);
                unsigned int nq[8] = {};
                WF_TRY(cudaMemcpy(nq, G.buf.counters, sizeof(nq), cudaMemcpyDeviceToHost));
                t_trace += a; t_shade += b; n_trace++; n_shade++;
                if (timing_env) fprintf(stderr, "wf g%d it %lld: trace %.3f ms  shade %.3f ms  active px %u  next queue %u\n", g, R.it, a, b, nq[2], nq[3 + R.parity]);
            }
            R.it++;
            if (!R.poll_pending && ((R.it & 3) == 0 || R.it >= max_iters))
            {
                WF_TRY(cudaMemcpyAsync(G.host_active, &G.buf.counters[2], sizeof(unsigned int), cudaMemcpyDeviceToHost, G.stream));
                WF_TRY(cudaEventRecord(G.poll_event, G.stream));
                if (err == cudaSuccess) { R.poll_pending = true; R.poll_it = R.it; }
            }
        }
    }
    // join (on every path, error or not): the caller's stream continues after every group that was forked
    for (int g = 0; g < n_groups; g++)
    {
        const WfGroup& G = groups[g];
        if (!run[g].forked) continue;
        const cudaError_t e1 = cudaEventRecord(G.join_event, G.stream);
        const cudaError_t e2 = e1 == cudaSuccess ? cudaStreamWaitEvent(stream, G.join_event, 0) : e1;
        if (err == cudaSuccess && e2 != cudaSuccess) err = e2;
    }
#undef WF_TRY
    if (timing)
    {
        if (timing_env) fprintf(stderr, "wf total: trace %.3f ms  shade %.3f ms  launches %d\n", t_trace, t_shade, launches);
        if (kernel_times4) { kernel_times4[0] = t_trace; kernel_times4[1] = t_shade; kernel_times4[2] = n_trace; kernel_times4[3] = n_shade; }
        for (int i = 0; i < 3; i++) cudaEventDestroy(tev[i]);
    }
    if (launches_out) *launches_out = launches;
    if (unfinished_out) *unfinished_out = unfinished;
    if (err != cudaSuccess) return err;
    return cudaGetLastError();
}

// sums the groups' ray counters into *total (one tiny launch on the caller's stream, after the join)
__global__ void wf_sum_rays(const unsigned long long* a, const unsigned long long* b, const unsigned long long* c, const unsigned long long* d,
                            const unsigned long long* e2, const unsigned long long* f, const unsigned long long* g, const unsigned long long* h,
                            int n, unsigned long long* total)
{
    const unsigned long long* p[8] = { a, b, c, d, e2, f, g, h };
    unsigned long long s = 0;
    for (int i = 0; i < n; i++) s += *p[i];
    *total = s;
}

cudaError_t wavefront_sum_rays(const WfGroup* groups, int n_groups, unsigned long long* total, cudaStream_t stream)
{
    const unsigned long long* p[8];
    for (int i = 0; i < 8; i++) p[i] = groups[i < n_groups ? i : 0].buf.rays_total;
    wf_sum_rays<<<1, 1, 0, stream>>>(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7], n_groups, total);
    return cudaGetLastError();
}

} // namespace b200rt

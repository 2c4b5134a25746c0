// Barrier-free continuation of a pass-synchronous frame that pools the tracing work ACROSS warps (sm_100a). -fmad=false.
//
// wavefront.cu advances every pixel of a tile group by one shading step per trace/shade launch pair; a pass lasts as long as
// its slowest ray (~0.1 ms: a dependent chain of L2 fetches, L1 cold after every launch) plus the shading chain (~0.05 ms),
// however few rays it carries, and a frame is a chain of up to spp * (max_bounces + 1) passes. wf_tail (persist.cu) removes the
// barrier by giving every warp a few slots of its own: the warp then traces its handful of rays itself, one mini-pass at a time.
// Here one persistent launch per tile group has two kinds of warps:
//
//   shader warps : own the group's CHUNKS of 32 consecutive slots (an 8x4 pixel patch; chunk c belongs to shader warp c % W).
//                  A warp walks its chunks; a chunk whose count of outstanding rays is zero is shaded as a whole
//                  (wf_shade_slot: the same device code as every other integrator), its new rays go to the device-wide RAY
//                  ring grouped by kind — 32 neighbouring pixels' rays of one kind side by side, as coherent as a pass's queue.
//   tracer warps : every idle lane holds a ticket of the ray ring (wf_device.cuh); when its ray arrives it traces it with the
//                  same warp-cooperative traversal as wf_trace_coop (coop_trace_queue), writes ONE stamped 16-byte result and
//                  fires ONE decrement of the chunk's count — no fence, no return value, nothing the warp waits for.
//
// A chunk's chain of iterations then costs what ITS rays and ITS shading cost — no pass, no slowest ray of the group, no launch —
// while every ray of the group finds a lane wherever one is free, and L1 stays warm with the top of the BVH for the whole launch.
// Same per-slot state, same device functions, same per-pixel RNG streams: the frame is bit-identical to the other integrators'
// and the ray count is the same.
// Memory ordering: a shader publishes a slot's state and ray records by the release store of its first ring entry; a tracer
// publishes nothing — its result record carries the slot's shading-step count as a stamp (one 16-byte store is single-copy
// atomic), and a shader that finds a chunk's count at zero only shades the slots whose expected results all carry the right
// stamp (a result still in flight just waits for the next look). Per-slot state is read with loads that bypass L1 (wf_ld).
//
// Status: a study path (B200RT_FLAG_WF_ASYNC), bit-identical and tested, NOT the default — a group run this way sustains 1.4 Grays/s
// against the passes' 2.4, because the 6 000-instruction shading step and the traversal loop then share every SM's instruction
// cache (5.3 "no instruction" stalls per issue; DESIGN.md 4.3, profiles/r2d_*).
#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "wf_device.cuh"

namespace b200rt {

#ifndef ASYNC_MIN_BLOCKS
#define ASYNC_MIN_BLOCKS 7
#endif

// the rays a slot's last shading step left behind, from its state (the pass kernels' queue is not carried over):
// PENDING -> the side rays whose kind is set, ALIVE -> the path ray
__device__ __forceinline__ void async_rays_of_state(const WfBuffers& B, int slot, int st, bool q[5])
{
    const int n = B.n_slots;
    q[0] = (st & WF_ALIVE) != 0;
#pragma unroll
    for (int k = 0; k < 4; k++)
        q[k + 1] = (st & WF_PENDING) && __float_as_int(wf_ld(&B.ray_d[(size_t)k * n + slot].w)) != SIDE_NONE;
}

// Have the results of all rays the slot is waiting for arrived? (their stamp is the slot's shading-step count.) The slot's flag
// word, its four side-ray kinds and its five result stamps are loaded together: one L2 round trip.
__device__ __forceinline__ bool async_slot_ready(const WfBuffers& B, int slot, int& st)
{
    const int n = B.n_slots;
    st = wf_ld(B.flags + slot);
    unsigned int stamp_k[5];
    int kind[4];
#pragma unroll
    for (int k = 0; k < 4; k++) kind[k] = __float_as_int(wf_ld(&B.ray_d[(size_t)k * n + slot].w));
#pragma unroll
    for (int k = 0; k < 5; k++) stamp_k[k] = __float_as_uint(wf_ld(&B.res[(size_t)k * n + slot].w));
    if (st & WF_DONE) return false;
    const unsigned int stamp = (unsigned int)st >> kWfSeqShift;
    bool ok = !(st & WF_ALIVE) || stamp_k[4] == stamp;
#pragma unroll
    for (int k = 0; k < 4; k++)
        if ((st & WF_PENDING) && kind[k] != SIDE_NONE) ok = ok && stamp_k[k] == stamp;
    return ok;
}

// pushes the rays of one shading step (q[0] = path ray, q[1..4] = side rays 0..3), grouped by kind like the pass kernels' queue
__device__ __forceinline__ unsigned int async_push_rays(const WfRing& rays, int slot, const bool q[5], bool release)
{
    const unsigned int s3 = (unsigned int)slot << 3;
    const unsigned int payload[5] = { s3 | 4u, s3 | 0u, s3 | 1u, s3 | 2u, s3 | 3u };
    return ring_push5(rays, q, payload, release);
}

// Seeds the ring and the chunk counters from the state the last shade pass (or wf_init) left: every unfinished slot's rays, every
// chunk's count of rays and of unfinished slots. One warp = one chunk.
__global__ void __launch_bounds__(256) wf_async_seed(WfBuffers B, CoopQueueRing A)
{
    const int slot = blockIdx.x * blockDim.x + threadIdx.x;
    bool q[5] = { false, false, false, false, false };
    bool alive = false;
    if (slot < B.n_slots)
    {
        const int st = B.flags[slot];
        alive = !(st & WF_DONE);
        if (alive) async_rays_of_state(B, slot, st, q);
    }
    const unsigned int pushed = async_push_rays(A.rays, slot, q, false);      // the kernel boundary publishes
    const unsigned int n_alive = __popc(__ballot_sync(0xffffffffu, alive));
    if ((threadIdx.x & 31) == 0 && (slot >> 5) < A.n_chunks)
    {
        const unsigned int ci = A.chunk_index((unsigned int)slot >> 5);
        A.cnt[ci] = pushed;
        A.chunk_live[ci] = n_alive;
        if (pushed) atomicAdd(B.rays_total, (unsigned long long)pushed);
    }
}

__global__ void __launch_bounds__(32 * kCoopMaxWarps, ASYNC_MIN_BLOCKS)
wf_async(SceneDev S, RenderParams P, WfBuffers B, CoopQueueRing A, int shader_stride, const float4* __restrict__ fb_in_rowmajor,
         float4* __restrict__ out_tiles)
{
    __shared__ CoopWarp s_warps[kCoopMaxWarps];
    __shared__ uint2 s_stack[kSharedStackDepth * 32 * kCoopMaxWarps];
    const unsigned int FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const int warp_global = blockIdx.x * kCoopMaxWarps + (threadIdx.x >> 5);
    const int sw = warp_global / shader_stride;        // shader warps: every shader_stride-th warp of the grid, the first W of them

    if (warp_global % shader_stride != 0 || sw >= A.W)
    {
        // ---- tracer warp: runs until no pixel of the group is left --------------------------------------------------------------------
        CoopWarp& W = s_warps[threadIdx.x >> 5];
        TravStack8Shared K;
        K.sh = s_stack + threadIdx.x; K.stride = 32 * kCoopMaxWarps;
        coop_trace_queue(S, B, (const unsigned int*)nullptr, A, W, K);
        return;
    }

    // ---- shader warp: walks its chunks until all of them are finished ----------------------------------------------------------------------
    unsigned int* const my_cnt = A.cnt + (size_t)sw * A.K;
    unsigned int* const my_live = A.chunk_live + (size_t)sw * A.K;
    unsigned long long rays = 0;
    for (;;)
    {
        bool any_live = false, shaded = false;
        for (int jb = 0; jb < A.K; jb += 32)
        {
            // lane l looks at chunk jb + l of this warp (always the same lane for a chunk: it also fires the chunk's increments,
            // so its own later loads of the count see them)
            const int j = jb + lane;
            const bool valid = j < A.K && j * A.W + sw < A.n_chunks;
            unsigned int lv = 0u, word = 1u;
            if (valid) { lv = ld_relaxed(my_live + j); word = ld_relaxed(my_cnt + j); }
            any_live = any_live || lv != 0u;
            unsigned int m = __ballot_sync(FULL, lv != 0u && word == 0u);
            while (m)
            {
                const int l = __ffs((int)m) - 1;
                m &= m - 1u;
                const int slot = ((jb + l) * A.W + sw) * 32 + lane;
                bool q[5] = { false, false, false, false, false };
                bool pixel_done = false;
                int st = WF_DONE;
                if (slot < B.n_slots && async_slot_ready(B, slot, st))
                {
                    int x, y;
                    wf_slot_pixel(P, slot, x, y);
                    // one shading step; a step that neither produced a ray nor finished the pixel (a terminated path whose side rays
                    // were all skipped) is followed by the next one right away
                    for (;;)
                    {
                        const ShadeOut R = wf_shade_slot(S, P, B, slot, st, x, y, wf_out_index(B, slot), fb_in_rowmajor, out_tiles);
                        st = R.flags;
                        q[0] = R.q_path; q[1] = R.q0; q[2] = R.q1; q[3] = R.q2; q[4] = R.q3;
                        pixel_done = R.pixel_done;
                        if (pixel_done || q[0] || q[1] || q[2] || q[3] || q[4]) break;
                    }
                }
                // release: the lane's state and ray records become visible before its first ring entry
                const unsigned int pushed = async_push_rays(A.rays, slot, q, true);
                const unsigned int n_done = __popc(__ballot_sync(FULL, pixel_done));
                if (lane == l)
                {
                    if (pushed) red_add_relaxed(my_cnt + j, pushed);
                    if (n_done) { lv -= n_done; st_relaxed(my_live + j, lv); }
                }
                if (lane == 0 && n_done) atomicSub(&B.counters[2], n_done);
                rays += pushed;
                shaded = true;
            }
        }
        if (!__any_sync(FULL, any_live)) break;
        if (!shaded) __nanosleep(200);
    }
    if (lane == 0 && rays) atomicAdd(B.rays_total, rays);
}

static int ceil_log2(size_t v)
{
    int l = 0;
    while (((size_t)1 << l) < v) l++;
    return l;
}

// ring capacity of a tile group of n_slots pixels: every ray that can exist at once (5 per slot) fits with room to spare, and a lap
// is never shorter than 2^20 entries
int wavefront_async_ray_log2(int n_slots) { return std::max(20, ceil_log2((size_t)5 * (size_t)std::max(1, n_slots) + 1)); }
// words of the per-chunk arrays: W * K <= n_chunks + W, W < kAsyncMaxWarps
constexpr int kAsyncMaxWarps = 8192;
int wavefront_async_chunk_words(int n_slots) { return (std::max(1, n_slots) + 31) / 32 + kAsyncMaxWarps; }
bool wavefront_async_fits(int n_slots) { return n_slots > 0 && n_slots < (1 << (kRingPayloadBits - 3)); }

int wavefront_async_max_ctas()
{
    int per_sm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, wf_async, 32 * kCoopMaxWarps, 0);
    if (per_sm <= 0) per_sm = 1;
    return current_sm_count() * per_sm;
}

// M: ring sized by wavefront_async_ray_log2, chunk arrays by wavefront_async_chunk_words (of the capacity the group was allocated for)
cudaError_t launch_wavefront_async(const SceneDev& S, const RenderParams& P, const WfBuffers& B, const WfAsyncMem& M, int ctas,
                                   const float4* fb_in_rowmajor, float4* out_tiles, cudaStream_t stream)
{
    if (B.n_slots <= 0) return cudaSuccess;
    // B200RT_WF_ASYNC_SHADERS = "num/den": that share of the grid's warps shades (spread evenly over the CTAs), the rest traces
    int sh_num = 1, sh_den = 4;
    {
        const char* ev = getenv("B200RT_WF_ASYNC_SHADERS");
        int a = 0, b = 0;
        if (ev && sscanf(ev, "%d/%d", &a, &b) == 2 && a >= 1 && b >= 2 * a) { sh_num = a; sh_den = b; }
    }
    // no more CTAs than could ever have work: a slot is either being shaded or has at most 5 rays in flight
    const int useful = std::max(1, (int)(((size_t)5 * B.n_slots + 32 * kCoopMaxWarps - 1) / (32 * kCoopMaxWarps)));
    const int grid = std::max(1, std::min(std::min(ctas, useful), kAsyncMaxWarps / kCoopMaxWarps - 1));
    const int total_warps = grid * kCoopMaxWarps;
    const int n_chunks = (B.n_slots + 31) / 32;
    const int W = std::max(1, std::min(n_chunks, (int)((long long)total_warps * sh_num / sh_den)));      // <= total_warps / 2: tracers exist
    const int stride = total_warps / W;
    cudaError_t e;
    if ((e = cudaMemsetAsync(M.ray_ring, 0xff, sizeof(unsigned int) << M.ray_log2, stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(M.ctrl, 0, 64 * sizeof(unsigned int), stream)) != cudaSuccess) return e;
    CoopQueueRing A;
    A.rays.buf = M.ray_ring; A.rays.ctrl = M.ctrl; A.rays.log2cap = (unsigned int)M.ray_log2;
    A.cnt = M.chunk_cnt; A.chunk_live = M.chunk_live;
    A.live = B.counters + 2;
    A.n_chunks = n_chunks; A.W = W; A.K = (n_chunks + W - 1) / W;
    if ((size_t)A.W * A.K > (size_t)M.chunk_words) return cudaErrorInvalidValue;
    wf_async_seed<<<(B.n_slots + 255) / 256, 256, 0, stream>>>(B, A);
    wf_async<<<grid, 32 * kCoopMaxWarps, 0, stream>>>(S, P, B, A, stride, fb_in_rowmajor, out_tiles);
    return cudaGetLastError();
}

} // namespace b200rt

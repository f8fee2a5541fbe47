// j40-b200: LF-group kernels (serial modular decoders of the LF image and the HF metadata, parallel post stages)
#include "j40b_kernels.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

namespace j40b {

// the weighted predictor's error row is part of the slice for the classes that use it, and for the kernel that ends stage
// 1 (the varblock placement's occupancy bitmap lies over error, property and sample rows)
__host__ __device__ inline bool lf_chan_needs_wp_row(int K, int stage, int c, bool split) {
    return K == MC_WP || K == MC_REST || (stage == 1 && c == (split ? 2 : 3));
}

// One channel of one serial stage of the LF groups, one decoder class per kernel (lf_chan_body in j40b_exec.h): one warp
// per 32 / G LF groups, G lanes each; `cap`: width of the shared-memory rows. The work list is ordered largest group first.
//
// G < 32: the per-sample instruction stream is the same for every stream of a class, so a warp that carries 2 or 4 streams
// in lock step spends one instruction on 2 or 4 samples (measured over 1024 LF groups: 8.0 G -> 2.15 G warp instructions,
// 29.85 of 32 lanes active, same time per sample). Each group of G lanes keeps the complete decoder state of its stream in
// its own registers and its own slice of shared memory, and talks only to itself (votes and shuffles among whoever is
// converged, barriers under the group's mask), so the groups need not agree on anything: where their control flow differs
// (a refill, a long symbol, an edge sample, an error) the hardware runs them one after the other and joins them again. The
// compiled tree must fit G lanes (host: Batch::lf_tree_lanes). CudaBackend::launch_lf takes G < 32 only on request
// (J40B_LF_LANES): a pipeline of batches turned out not to be bound by issue slots and gains 2 %, a batch alone takes 24 %
// longer ("Lane groups").
//
// `rot`: the grid is at least one block per SM, and block b works as block (b - rot) mod gridDim.x of the work list (those
// beyond the list leave at once), each batch object rotating by a different amount (CudaBackend::turn), so that the
// long-lived warps of the batches in flight -- the work list starts with the big LF groups -- do not depend on the order in
// which the hardware hands blocks to SMs. Measured (J40B_LF_SMHIST): residency is even over the SMs to +- 15 %; the
// rotation itself moved the LF stages of a pipeline from 19.4 to 18.8 ms per step.
// diagnostics (J40B_LF_SMHIST=1, kl_lf_smhist_dump): warp residency of these kernels per SM, in cycles, and the most warps an
// SM held at once
__device__ unsigned long long g_lf_sm_busy[256];
__device__ unsigned int g_lf_sm_now[256], g_lf_sm_max[256];
struct LfSmHist {
    unsigned int smid = 0; long long t0 = 0; bool on;
    __device__ LfSmHist(bool on_) : on(on_) {
        if (!on || (threadIdx.x & 31)) return;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        atomicMax(&g_lf_sm_max[smid & 255], atomicAdd(&g_lf_sm_now[smid & 255], 1u) + 1u);
        t0 = clock64();
    }
    __device__ ~LfSmHist() {
        if (!on || (threadIdx.x & 31)) return;
        atomicAdd(&g_lf_sm_busy[smid & 255], (unsigned long long) (clock64() - t0));
        atomicSub(&g_lf_sm_now[smid & 255], 1u);
    }
};
void kl_lf_smhist_dump() {
    unsigned long long busy[256]; unsigned int mx[256];
    cudaDeviceSynchronize();
    if (cudaMemcpyFromSymbol(busy, g_lf_sm_busy, sizeof(busy)) != cudaSuccess || cudaMemcpyFromSymbol(mx, g_lf_sm_max, sizeof(mx)) != cudaSuccess) return;
    unsigned long long tot = 0;
    for (int i = 0; i < 256; ++i) tot += busy[i];
    fprintf(stderr, "lf smhist (share of warp-cycles in 1/1000, max warps at once):");
    for (int i = 0; i < 160; ++i) fprintf(stderr, " %d:%.1f/%u", i, tot ? 1000.0 * (double) busy[i] / (double) tot : 0.0, mx[i]);
    fprintf(stderr, "\n");
    memset(busy, 0, sizeof(busy)); memset(mx, 0, sizeof(mx));
    cudaMemcpyToSymbol(g_lf_sm_busy, busy, sizeof(busy)); cudaMemcpyToSymbol(g_lf_sm_max, mx, sizeof(mx));
}

template <int K, int G>
__global__ void __launch_bounds__(32) k_lf_chan(const LfWork *items, int n, int stage, int c, int cap, int split, int rot) {
    int blk = (int) blockIdx.x - rot;
    if (blk < 0) blk += (int) gridDim.x;
    if (blk * (32 / G) >= n) return;
    LfSmHist hist((split & 2) != 0);
    __shared__ int32_t div24[64];
    extern __shared__ __align__(16) uint8_t smem[];
    fill_div24(div24, (int) threadIdx.x, 32);
    __syncwarp();
    const bool with_wp = lf_chan_needs_wp_row(K, stage, c, (split & 1) != 0);
    if (G == 32) {
        WarpScratch *ws;
        ModSmem ms = carve_warp_slice(smem, cap, ws, with_wp, LF_TAB_BYTES);
        lf_chan_body<K>(items[blk], stage, c, *ws, ms, div24, (int) threadIdx.x, 32, WarpSync(), (split & 1) != 0);
        return;
    }
    const int grp = (int) threadIdx.x / G, lane = (int) threadIdx.x % G;
    const int i = blk * (32 / G) + grp;
    if (i >= n) return;
    WarpScratch *ws;
    ModSmem ms = carve_warp_slice(smem + (size_t) grp * warp_slice_bytes(cap, with_wp, LF_TAB_BYTES), cap, ws, with_wp, LF_TAB_BYTES);
    ms.lanes = G;
    GroupSync sync;
    sync.s = grp * G;
    sync.m = (G == 32 ? 0xffffffffu : (1u << G) - 1u) << sync.s;
    lf_chan_body<K>(items[i], stage, c, *ws, ms, div24, lane, G, sync, (split & 1) != 0);
}

// Lane-per-stream variant (j40b_modlane.h): every thread owns one LF group; 32 * LANE_WARPS work items per block.
// Shared memory (lane-interleaved): the 64-entry divisor table, 16 property slots and the compiled tree of the current
// channel (LANE_NODE_CAP nodes) per thread.
struct LaneSmem {
    int32_t div24[64];
    int32_t props[16 * 32 * LANE_WARPS];
    int32_t nodes[LANE_NODE_CAP * 4 * 32 * LANE_WARPS];
};
__device__ inline LaneEnv lane_env(LaneSmem &sm) {
    LaneEnv env;
    env.div24 = sm.div24;
    env.props = sm.props + threadIdx.x;
    env.nodes = sm.nodes + threadIdx.x;
    env.ring = nullptr; env.wring = nullptr; env.ring_w = 0;
    env.lstride = 32 * LANE_WARPS;
    return env;
}

template <int STAGE>
__global__ void __launch_bounds__(32 * LANE_WARPS, 1) k_lf_lane(const LfWork *items, int n) {
    __shared__ LaneSmem sm;
    fill_div24(sm.div24, (int) threadIdx.x, (int) blockDim.x);
    __syncthreads();
    const int i = (int) (blockIdx.x * blockDim.x + threadIdx.x);
    const LfWork *w = &items[i < n ? i : n - 1];
    const bool active = i < n && (STAGE == 1 || !*w->err);
    // the fast path (rANS, no LZ77) is taken when every stream of the warp qualifies
    const bool plain = __all_sync(0xffffffffu, !active || spec_is_plain_ans(w->arena, lf_stage_spec_off(*w, STAGE - 1)));
    const LaneEnv env = lane_env(sm);
    if (STAGE == 1) {
        if (plain) lf_decode1_lanes<1>(w, active, env, WarpAny(), WarpSync());
        else lf_decode1_lanes<0>(w, active, env, WarpAny(), WarpSync());
    } else {
        if (plain) lf_decode2_lanes<1>(w, active, env, WarpAny(), WarpSync());
        else lf_decode2_lanes<0>(w, active, env, WarpAny(), WarpSync());
    }
}

// inverse transforms of the HF metadata image + varblock placement behind k_lf_lane<2>: one warp per LF group
__global__ void __launch_bounds__(32) k_lf_place(const LfWork *items) {
    __shared__ uint32_t bitmap[256 * 8];
    lf_place_body(items[blockIdx.x], bitmap, (int) threadIdx.x, 32, WarpSync());
}

__global__ void __launch_bounds__(256) k_lf_post(const LfWork *items) {
    lf_post_body(items[blockIdx.x], (int) threadIdx.x, (int) blockDim.x, BlockSync());
}

__global__ void __launch_bounds__(128) k_lf_llf(const LfWork *items) {
    lf_llf_body(items[blockIdx.x], (int) threadIdx.x, (int) blockDim.x, BlockSync());
}


template <int K, int G>
static bool lf_chan_attr() {
    return kl_carveout(k_lf_chan<K, G>, "J40B_CARVEOUT_LF") &&
           cudaFuncSetAttribute(k_lf_chan<K, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) warp_slice_bytes(LF_ROW_CAP, true, LF_TAB_BYTES) * (32 / G)) == cudaSuccess;
}
template <int G>
static bool lf_chan_attrs() {
    return lf_chan_attr<MC_WP, G>() && lf_chan_attr<MC_GRAD, G>() && lf_chan_attr<MC_WIDE, G>() && lf_chan_attr<MC_GEN, G>() && lf_chan_attr<MC_REST, G>();
}
bool kl_init_lf() {
    return lf_chan_attrs<32>() && lf_chan_attrs<16>() && lf_chan_attrs<8>() && kl_carveout(k_lf_post, "J40B_CARVEOUT_LF") && kl_carveout(k_lf_llf, "J40B_CARVEOUT_LF") &&
           kl_carveout(k_lf_place, "J40B_CARVEOUT_LF") && kl_carveout(k_lf_lane<1>, "J40B_CARVEOUT_LF") && kl_carveout(k_lf_lane<2>, "J40B_CARVEOUT_LF");
}

template <int K, int G>
static void lf_chan_launch(cudaStream_t stream, const LfWork *w, int n, int stage, int c, int cap, bool split, int spread, int turn) {
    const int per = 32 / G, blocks = (n + per - 1) / per;
    static const bool smhist = getenv("J40B_LF_SMHIST") != nullptr;
    // small launches (a single image: latency) are left alone; the others cover every SM and start at a place of their own
    const int grid = blocks >= 16 && blocks < spread ? spread : blocks;
    const int rot = blocks >= 16 ? kl_rotation(turn, grid) : 0;
    k_lf_chan<K, G><<<grid, 32, warp_slice_bytes(cap, lf_chan_needs_wp_row(K, stage, c, split), LF_TAB_BYTES) * per, stream>>>(w, n, stage, c, cap, (split ? 1 : 0) | (smhist ? 2 : 0), rot);
}
template <int G>
static void lf_stage_launch(int stage, cudaStream_t stream, const LfWork *w, int n, int cap, int c0, int c1, bool split, int spread, int turn) {
    for (int c = c0; c < c1; ++c) {
        lf_chan_launch<MC_WP, G>(stream, w, n, stage, c, cap, split, spread, turn);
        lf_chan_launch<MC_GRAD, G>(stream, w, n, stage, c, cap, split, spread, turn);
        lf_chan_launch<MC_WIDE, G>(stream, w, n, stage, c, cap, split, spread, turn);
        lf_chan_launch<MC_GEN, G>(stream, w, n, stage, c, cap, split, spread, turn);
        lf_chan_launch<MC_REST, G>(stream, w, n, stage, c, cap, split, spread, turn);
    }
}
// channels [c0, c1) of a stage (0: LF image, 1: HF metadata + placement): per channel the five class kernels in a row;
// `lanes` (32, 16 or 8) lanes per LF group; `spread`: SMs of the device, `turn`: which batch object this is (see k_lf_chan)
void kl_lf_stage(int stage, cudaStream_t stream, const LfWork *w, int n, int cap, int c0, int c1, bool split, int lanes, int spread, int turn) {
    if (lanes == 8) lf_stage_launch<8>(stage, stream, w, n, cap, c0, c1, split, spread, turn);
    else if (lanes == 16) lf_stage_launch<16>(stage, stream, w, n, cap, c0, c1, split, spread, turn);
    else lf_stage_launch<32>(stage, stream, w, n, cap, c0, c1, split, spread, turn);
}
void kl_lf_lane(int stage, cudaStream_t stream, const LfWork *w, int n) {
    const int per_block = 32 * LANE_WARPS, blocks = (n + per_block - 1) / per_block;
    if (stage == 1) k_lf_lane<1><<<blocks, per_block, 0, stream>>>(w, n);
    else k_lf_lane<2><<<blocks, per_block, 0, stream>>>(w, n);
}
void kl_lf_place(int n, cudaStream_t stream, const LfWork *w) { k_lf_place<<<n, 32, 0, stream>>>(w); }
void kl_lf_post(int n, cudaStream_t stream, const LfWork *w) { k_lf_post<<<n, 256, 0, stream>>>(w); }
void kl_lf_llf(int n, cudaStream_t stream, const LfWork *w) { k_lf_llf<<<n, 128, 0, stream>>>(w); }

} // namespace j40b

// j40-b200: LF-group kernels (serial modular decoders of the LF image and the HF metadata, parallel post stages)
#include "j40b_kernels.h"
#include <stdlib.h>

namespace j40b {

// the weighted predictor's error row is part of the slice for the classes that use it, and for the kernel that ends stage
// 1 (the varblock placement's occupancy bitmap lies over error, property and sample rows)
__host__ __device__ inline bool lf_chan_needs_wp_row(int K, int stage, int c, bool split) {
    return K == MC_WP || K == MC_REST || (stage == 1 && c == (split ? 2 : 3));
}

// One channel of one serial stage of the LF groups, one decoder class per kernel (lf_chan_body in j40b_exec.h): one warp
// per LF group; `cap`: width of the shared-memory rows. The work list is ordered largest group first.
template <int K>
__global__ void __launch_bounds__(32) k_lf_chan(const LfWork *items, int stage, int c, int cap, int split) {
    __shared__ int32_t div24[64];
    extern __shared__ __align__(16) uint8_t smem[];
    fill_div24(div24, (int) threadIdx.x, 32);
    __syncwarp();
    WarpScratch *ws;
    ModSmem ms = carve_warp_slice(smem, cap, ws, lf_chan_needs_wp_row(K, stage, c, split != 0));
    lf_chan_body<K>(items[blockIdx.x], stage, c, *ws, ms, div24, (int) threadIdx.x, 32, WarpSync(), split != 0);
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


bool kl_init_lf() {
    const int lf_smem = (int) warp_slice_bytes(LF_ROW_CAP);
    return cudaFuncSetAttribute(k_lf_chan<MC_WP>, cudaFuncAttributeMaxDynamicSharedMemorySize, lf_smem) == cudaSuccess &&
           cudaFuncSetAttribute(k_lf_chan<MC_GRAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, lf_smem) == cudaSuccess &&
           cudaFuncSetAttribute(k_lf_chan<MC_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, lf_smem) == cudaSuccess &&
           cudaFuncSetAttribute(k_lf_chan<MC_GEN>, cudaFuncAttributeMaxDynamicSharedMemorySize, lf_smem) == cudaSuccess &&
           cudaFuncSetAttribute(k_lf_chan<MC_REST>, cudaFuncAttributeMaxDynamicSharedMemorySize, lf_smem) == cudaSuccess;
}
// channels [c0, c1) of a stage (0: LF image, 1: HF metadata + placement): per channel the five class kernels in a row
void kl_lf_stage(int stage, cudaStream_t stream, const LfWork *w, int n, int cap, int c0, int c1, bool split) {
    const int sp = split ? 1 : 0;
    for (int c = c0; c < c1; ++c) {
        k_lf_chan<MC_WP><<<n, 32, warp_slice_bytes(cap, lf_chan_needs_wp_row(MC_WP, stage, c, split)), stream>>>(w, stage, c, cap, sp);
        k_lf_chan<MC_GRAD><<<n, 32, warp_slice_bytes(cap, lf_chan_needs_wp_row(MC_GRAD, stage, c, split)), stream>>>(w, stage, c, cap, sp);
        k_lf_chan<MC_WIDE><<<n, 32, warp_slice_bytes(cap, lf_chan_needs_wp_row(MC_WIDE, stage, c, split)), stream>>>(w, stage, c, cap, sp);
        k_lf_chan<MC_GEN><<<n, 32, warp_slice_bytes(cap, lf_chan_needs_wp_row(MC_GEN, stage, c, split)), stream>>>(w, stage, c, cap, sp);
        k_lf_chan<MC_REST><<<n, 32, warp_slice_bytes(cap, lf_chan_needs_wp_row(MC_REST, stage, c, split)), stream>>>(w, stage, c, cap, sp);
    }
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

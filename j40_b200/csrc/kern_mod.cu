// j40-b200: modular-frame kernels (group decode, inverse transforms + RGBA8 render)
#include "j40b_kernels.h"

namespace j40b {

// `cap`: width of the shared-memory rows (the widest channel of this launch, rounded up); `spec_cap`: bytes reserved
// for the staged code spec (0 = tables through L1). Both are sized per launch: a block of a 256-pixel-wide group
// then takes ~30 KB instead of the 106 KB of the widest case, and seven of them fit an SM instead of two.
__global__ void __launch_bounds__(32) k_modular(ModWork *items, int cap, int spec_cap) {
    __shared__ int32_t div24[64];
    extern __shared__ __align__(16) uint8_t smem[];
    ModWork &w = items[blockIdx.x];
    const bool staged = spec_cap > 0 && stage_spec_blob(w.arena, w.spec_off, smem, (uint32_t) spec_cap, (int) threadIdx.x, 32);
    fill_div24(div24, (int) threadIdx.x, 32);
    __syncwarp();
    WarpScratch *ws;
    ModSmem ms = carve_warp_slice(smem + spec_cap, cap, ws);
    modular_body(w, *ws, ms, div24, staged ? smem : nullptr, w.arena, (int) threadIdx.x, 32, WarpSync());
}

// Lane-per-stream variant (j40b_modlane.h): every thread owns one sub-bitstream (shared memory as in k_lf_lane)
__global__ void __launch_bounds__(32 * LANE_WARPS, 1) k_mod_lane(ModWork *items, int n) {
    __shared__ int32_t div24[64];
    __shared__ int32_t props[16 * 32 * LANE_WARPS];
    __shared__ int32_t nodes[LANE_NODE_CAP * 4 * 32 * LANE_WARPS];
    fill_div24(div24, (int) threadIdx.x, (int) blockDim.x);
    __syncthreads();
    const int i = (int) (blockIdx.x * blockDim.x + threadIdx.x);
    ModWork *w = &items[i < n ? i : n - 1];
    const bool active = i < n;
    const bool plain = __all_sync(0xffffffffu, !active || spec_is_plain_ans(w->arena, w->spec_off));
    LaneEnv env;
    env.div24 = div24; env.props = props + threadIdx.x; env.nodes = nodes + threadIdx.x;
    env.ring = nullptr; env.wring = nullptr; env.ring_w = 0; env.lstride = 32 * LANE_WARPS;
    if (plain) modular_lanes<1>(w, active, env, WarpAny(), WarpSync());
    else modular_lanes<0>(w, active, env, WarpAny(), WarpSync());
}

// rows go over grid.x together with the column blocks (grid.y is capped at 65535, frames may be 2^18 rows tall)
__global__ void __launch_bounds__(256) k_render(const RenderWork *w, int width, int height, int xblocks) {
    const int y = (int) (blockIdx.x / (unsigned) xblocks), xb = (int) (blockIdx.x % (unsigned) xblocks);
    const int x = xb * (int) blockDim.x + (int) threadIdx.x;
    if (x < width && y < height) render_px(*w, x, y);
}

// delta palettes (palette_delta_body): one thread per restored channel
__global__ void __launch_bounds__(32) k_palette_delta(const RenderWork *w, int num_c) {
    __shared__ int32_t div24[64];
    fill_div24(div24, (int) threadIdx.x, 32);
    __syncwarp();
    if ((int) threadIdx.x < num_c) palette_delta_body(*w, (int) threadIdx.x, div24);
}
void kl_palette_delta(cudaStream_t stream, const RenderWork *w, int num_c) { k_palette_delta<<<1, 32, 0, stream>>>(w, num_c); }

bool kl_init_mod() {
    const int mod_smem = (int) (SPEC_COPY_BYTES + warp_slice_bytes(MOD_ROW_CAP));
    return kl_carveout(k_modular, "J40B_CARVEOUT_MOD") && kl_carveout(k_mod_lane, "J40B_CARVEOUT_MOD") && kl_carveout(k_render, "J40B_CARVEOUT_MOD") && kl_carveout(k_palette_delta, "J40B_CARVEOUT_MOD") &&
           cudaFuncSetAttribute(k_modular, cudaFuncAttributeMaxDynamicSharedMemorySize, mod_smem) == cudaSuccess;
}
void kl_modular(int n, cudaStream_t stream, ModWork *w, int cap, int spec_cap) {
    k_modular<<<n, 32, (size_t) spec_cap + warp_slice_bytes(cap), stream>>>(w, cap, spec_cap);
}
void kl_mod_lane(int n, cudaStream_t stream, ModWork *w) {
    const int per_block = 32 * LANE_WARPS;
    k_mod_lane<<<(n + per_block - 1) / per_block, per_block, 0, stream>>>(w, n);
}
void kl_render(cudaStream_t stream, const RenderWork *w, int width, int height) {
    if (width <= 0 || height <= 0) return;
    const int xblocks = (width + 255) / 256;
    k_render<<<(unsigned) xblocks * (unsigned) height, 256, 0, stream>>>(w, width, height, xblocks);
}

} // namespace j40b

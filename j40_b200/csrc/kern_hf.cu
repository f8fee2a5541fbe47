// j40-b200: pass-group kernels (HF coefficient entropy decode into token lists)
#include "j40b_kernels.h"

namespace j40b {

// varblock lists of the groups (see j40b_hf.h): one warp per group
__global__ void __launch_bounds__(128) k_hf_prep(const HfPrepWork *items, int n) {
    const int i = (int) (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5));
    if (i < n) hf_prep_body(items[i], (int) threadIdx.x & 31, 32, WarpSync());
}

// HF coefficient entropy decode, SIMT: one warp per `lanes` consecutive (pass, group) sections, one section per lane,
// in a warp-uniform two-phase loop (j40b_hf.h). The work list is ordered image by image and pass by pass, so a
// block's lanes nearly always share one code spec (cluster map + alias tables / prefix LUTs), which is staged in
// shared memory by one bulk asynchronous copy (cp.async.bulk, completion on an mbarrier); lanes of another image
// or pass (at the boundaries) and specs that do not fit read their tables from global memory instead.
// `lanes` (<= 32) sections per warp: with small batches fewer lanes per warp give more warps (latency hiding);
// the host picks it from the number of sections (CudaBackend::launch_hf).
// `rot`: as in k_lf_chan (kern_lf.cu) -- block b works as block (b - rot) mod gridDim.x, so that the few dozen blocks of
// the batches in flight do not all start on the same SMs.
__global__ void __launch_bounds__(32 * HF_WARPS) k_hf_group(const HfWork *items, int n, int lanes, int spec_cap, int rot) {
    extern __shared__ __align__(128) uint8_t spec_copy[];
    __shared__ uint16_t ctx_lut[128];
    __shared__ __align__(8) uint64_t bar;
    const int per_block = HF_WARPS * lanes;
    int blk = (int) blockIdx.x - rot;
    if (blk < 0) blk += (int) gridDim.x;
    const int first = blk * per_block;
    if (first >= n) return;
    const HfWork &w0 = items[first];
    const int pass0 = w0.grp->pass;
    const uint32_t spec_off = w0.f->coeff_spec_off[pass0];
    const DCodeSpec *spec = (const DCodeSpec *) (w0.arena + spec_off);
    const uint32_t blob_bytes = (spec->blob_hi - spec->blob_lo + 15u) & ~15u; // lo is 16-byte aligned; the arena is padded
    const bool staged = spec_cap > 0 && blob_bytes <= (uint32_t) spec_cap;
    if (staged && threadIdx.x == 0) {
        const uint32_t bar_s = (uint32_t) __cvta_generic_to_shared(&bar), dst_s = (uint32_t) __cvta_generic_to_shared(spec_copy);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_s), "r"(blob_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(dst_s), "l"(w0.arena + spec->blob_lo), "r"(blob_bytes), "r"(bar_s) : "memory");
    }
    if (threadIdx.x < 64) ctx_lut[threadIdx.x] = (uint16_t) coeff_nnz_ctx2((int) threadIdx.x);
    else if (threadIdx.x < 128) ctx_lut[threadIdx.x] = (uint16_t) (threadIdx.x == 64 ? 0 : coeff_freq_ctx2((int) threadIdx.x - 64));
    __syncthreads(); // the barrier is initialised (and the look-up table written) before anyone waits on it
    if (staged) {
        const uint32_t bar_s = (uint32_t) __cvta_generic_to_shared(&bar);
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar_s) : "memory");
        }
    }
    const int warp = (int) threadIdx.x >> 5, lane = (int) threadIdx.x & 31;
    const int i = first + warp * lanes + lane;
    const bool real = lane < lanes && i < n && items[i < n ? i : first].grp != nullptr; // (padding items have no group)
    const HfWork *w = &items[real ? i : first];
    const bool active = real && !*w->lf_err;
    // the fast path (rANS, no LZ77) is taken when every section of the warp qualifies
    const bool plain = __all_sync(0xffffffffu, !active || hf_is_plain_ans(*w));
    const uint8_t *copy = staged ? spec_copy : nullptr;
    if (plain) hf_lanes_run<1>(w, active, copy, w0.arena, pass0, ctx_lut, WarpAny(), WarpSync());
    else hf_lanes_run<0>(w, active, copy, w0.arena, pass0, ctx_lut, WarpAny(), WarpSync());
}

bool kl_init_hf() { return kl_carveout(k_hf_prep, "J40B_CARVEOUT_HF") && kl_carveout(k_hf_group, "J40B_CARVEOUT_HF"); }
void kl_hf_prep(int n, cudaStream_t stream, const HfPrepWork *w) { k_hf_prep<<<(n + 3) / 4, 128, 0, stream>>>(w, n); }
void kl_hf_group(int blocks, size_t smem, cudaStream_t stream, const HfWork *w, int n, int lanes, int spec_cap, int spread, int turn) {
    const int grid = blocks >= 16 && blocks < spread ? spread : blocks;
    const int rot = blocks >= 16 ? kl_rotation(turn, grid) : 0;
    k_hf_group<<<grid, 32 * HF_WARPS, smem, stream>>>(w, n, lanes, spec_cap, rot);
}

} // namespace j40b

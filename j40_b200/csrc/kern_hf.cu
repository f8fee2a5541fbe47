// j40-b200: pass-group kernel (HF coefficient entropy decode into token lists)
#include "j40b_kernels.h"

namespace j40b {

// HF coefficient entropy decode, SIMT: one warp per 32 consecutive groups (one group per lane). The work
// list is ordered image by image, so a warp's lanes nearly always share one image, whose coefficient code
// spec (cluster map + alias tables / prefix LUTs) is staged in shared memory; lanes of another image (at
// image boundaries) and specs that do not fit read the tables from global memory instead.
// `lanes` (<= 32) groups per warp: with small batches fewer lanes per warp give more warps (latency hiding)
// and less divergence; the host picks it from the number of groups (CudaBackend::launch_hf).
__global__ void __launch_bounds__(32 * HF_WARPS) k_hf_group(const HfWork *items, int n, int lanes, int spec_cap) {
    extern __shared__ __align__(16) uint8_t spec_copy[];
    __shared__ uint16_t ctx_lut[128];
    const int per_block = HF_WARPS * lanes;
    const int first = (int) blockIdx.x * per_block;
    const HfWork &w0 = items[first];
    const bool staged = spec_cap > 0 && stage_spec_blob(w0.arena, w0.f->coeff_spec_off, spec_copy, (uint32_t) spec_cap, (int) threadIdx.x, 32 * HF_WARPS);
    if (threadIdx.x < 64) ctx_lut[threadIdx.x] = (uint16_t) coeff_nnz_ctx2((int) threadIdx.x);
    else if (threadIdx.x < 128) ctx_lut[threadIdx.x] = (uint16_t) (threadIdx.x == 64 ? 0 : coeff_freq_ctx2((int) threadIdx.x - 64));
    __syncthreads();
    const int warp = (int) threadIdx.x >> 5, lane = (int) threadIdx.x & 31;
    const int i = first + warp * lanes + lane;
    if (lane < lanes && i < n) hf_group_body(items[i], staged ? spec_copy : nullptr, w0.arena, ctx_lut);
}


void kl_hf_group(int blocks, size_t smem, cudaStream_t stream, const HfWork *w, int n, int lanes, int spec_cap) {
    k_hf_group<<<blocks, 32 * HF_WARPS, smem, stream>>>(w, n, lanes, spec_cap);
}

} // namespace j40b

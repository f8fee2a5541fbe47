// j40-b200: back-end kernels (tokens -> dequantised coefficients -> inverse transforms -> RGBA8)
#define J40B_KERN_BACK_TU
#include "j40b_kernels.h"
#include <stdio.h>
#include <stdlib.h>

namespace j40b {

static unsigned long long *g_phase = nullptr;
#if defined(J40B_PHASE_CLOCKS)
void kl_back_phase_dump() {
    unsigned long long h[16] = {0};
    cudaError_t e1 = cudaDeviceSynchronize();
    cudaError_t e2 = g_phase ? cudaMemcpy(h, g_phase, sizeof(h), cudaMemcpyDeviceToHost) : cudaSuccess;
    if (e1 != cudaSuccess || e2 != cudaSuccess) fprintf(stderr, "phase dump: %s / %s\n", cudaGetErrorString(e1), cudaGetErrorString(e2));
    unsigned long long tot = 0;
    for (int i = 0; i < 7; ++i) tot += h[i];
    fprintf(stderr, "k_back_tile phase cycles (thread 0 of every block):");
    for (int i = 0; i < 7; ++i) fprintf(stderr, " p%d=%.1f%%", i, 100.0 * (double) h[i] / (double) (tot ? tot : 1));
    fprintf(stderr, " total=%llu\n", tot);
    if (g_phase) cudaMemset(g_phase, 0, sizeof(h));
}
#else
void kl_back_phase_dump() {}
#endif

#ifndef J40B_TILE_PERSIST_DEFAULT
#define J40B_TILE_PERSIST_DEFAULT 3
#endif
#ifndef J40B_TILE_MINB
#define J40B_TILE_MINB 3 // (4: 64 registers per thread; experiment switch, Makefile: TILE_MINB)
#endif
// one block per 64x64-pixel tile of a group (blockIdx.y = tile index inside the 256x256 group)
__global__ void __launch_bounds__(256, J40B_TILE_MINB) k_back_tile(const BackWork *items, unsigned long long *phase) {
    extern __shared__ __align__(16) float tile_coef[];
    __shared__ TileShared ts;
    ts.phase = phase;
    ts.tables_staged = 0;
    back_tile_body(items[blockIdx.x], (int) (blockIdx.y & 3), (int) (blockIdx.y >> 2), tile_coef, ts, (int) threadIdx.x, (int) blockDim.x, BlockSync());
}

// The same as a persistent kernel: `gridDim.x` blocks walk the batch's tiles. A block that has found room on an SM keeps
// it until the batch's tiles are done, instead of every one of the 138 000 tile blocks of a 64-frame batch competing
// anew with the long-lived serial decoders of the other batches in flight (which leaves about one tile block per SM).
__global__ void __launch_bounds__(256, J40B_TILE_MINB) k_back_tile_persistent(const BackWork *items, int ntiles, unsigned long long *phase) {
    extern __shared__ __align__(16) float tile_coef[];
    __shared__ TileShared ts;
    __shared__ __align__(8) uint64_t bar;
    ts.phase = phase;
    // the sRGB threshold table and start LUT (the batch's shared tables, 5 KB) once per block: two bulk asynchronous copies
    // completing on one mbarrier, instead of a cooperative copy loop per tile
    if (threadIdx.x == 0) {
        const DFrame &f0 = *items[0].f;
        const uint32_t bar_s = (uint32_t) __cvta_generic_to_shared(&bar);
        const uint32_t thr_s = (uint32_t) __cvta_generic_to_shared(ts.thr), lut_s = (uint32_t) __cvta_generic_to_shared(ts.lut);
        const uint32_t thr_bytes = sizeof(ts.thr), lut_bytes = sizeof(ts.lut);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_s));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_s), "r"(thr_bytes + lut_bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(thr_s), "l"(f0.srgb_thr), "r"(thr_bytes), "r"(bar_s) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"(lut_s), "l"(f0.srgb_lut), "r"(lut_bytes), "r"(bar_s) : "memory");
        ts.tables_staged = 1;
    }
    __syncthreads();
    {
        const uint32_t bar_s = (uint32_t) __cvta_generic_to_shared(&bar);
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0; selp.u32 %0, 1, 0, p; }"
                         : "=r"(ok) : "r"(bar_s) : "memory");
        }
    }
    for (int t = (int) blockIdx.x; t < ntiles; t += (int) gridDim.x) {
        __syncthreads(); // the previous tile's readers of the shared buffers are done
        const int ti = t & 15;
        back_tile_body(items[t >> 4], ti & 3, ti >> 2, tile_coef, ts, (int) threadIdx.x, (int) blockDim.x, BlockSync());
    }
}

// varblocks the tile kernel leaves out: persistent blocks, each with its own 1 MiB slice of scratch
__global__ void __launch_bounds__(256) k_back_generic(const BackWork *items, int n, float *scratch_pool) {
    float *scratch = scratch_pool + (size_t) blockIdx.x * 4 * 65536;
    for (int i = (int) blockIdx.x; i < n; i += (int) gridDim.x) {
        BackWork w = items[i];
        w.big_scratch = scratch;
        back_generic_body(w, (int) threadIdx.x, (int) blockDim.x, BlockSync());
        __syncthreads();
    }
}


// diagnostics (j40b_batch_debug_dump): one block per varblock
__global__ void __launch_bounds__(128) k_dump_coeffs(DumpWork w) {
    dump_coeffs_body(w, (int) blockIdx.x, (int) threadIdx.x, (int) blockDim.x, BlockSync());
}
void kl_dump_coeffs(int n, cudaStream_t stream, const DumpWork &w) { if (n > 0) k_dump_coeffs<<<n, 128, 0, stream>>>(w); }

bool kl_init_back() {
    return kl_carveout(k_back_tile, "J40B_CARVEOUT_BACK") && kl_carveout(k_back_tile_persistent, "J40B_CARVEOUT_BACK") && kl_carveout(k_back_generic, "J40B_CARVEOUT_BACK") && kl_carveout(k_dump_coeffs, "J40B_CARVEOUT_BACK") &&
           cudaFuncSetAttribute(k_back_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * TILE_CH * 4) == cudaSuccess &&
           cudaFuncSetAttribute(k_back_tile_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * TILE_CH * 4) == cudaSuccess;
}
void kl_back_tile(int n, cudaStream_t stream, const BackWork *w) {
#if defined(J40B_PHASE_CLOCKS)
    if (!g_phase) { cudaMalloc(&g_phase, 16 * sizeof(unsigned long long)); cudaMemset(g_phase, 0, 16 * sizeof(unsigned long long)); }
#endif
    // J40B_TILE_PERSIST=b: persistent variant with b blocks per SM (0 = one block per tile)
    static const int persist = getenv("J40B_TILE_PERSIST") ? atoi(getenv("J40B_TILE_PERSIST")) : J40B_TILE_PERSIST_DEFAULT;
    if (persist > 0) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const long long blocks = (long long) sms * persist < (long long) n * 16 ? (long long) sms * persist : (long long) n * 16;
        k_back_tile_persistent<<<(unsigned) blocks, 256, 3 * TILE_CH * 4, stream>>>(w, n * 16, g_phase);
        return;
    }
    k_back_tile<<<dim3((unsigned) n, 16), 256, 3 * TILE_CH * 4, stream>>>(w, g_phase);
}
void kl_back_generic(int blocks, cudaStream_t stream, const BackWork *w, int n, float *pool) {
    k_back_generic<<<blocks, 256, 0, stream>>>(w, n, pool);
}

} // namespace j40b

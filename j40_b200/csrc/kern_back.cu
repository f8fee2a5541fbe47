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
    back_tile_body(items[blockIdx.x], (int) (blockIdx.y & 3), (int) (blockIdx.y >> 2), tile_coef, ts, (int) threadIdx.x, (int) blockDim.x, BlockSync());
}

// The same as a persistent kernel: `gridDim.x` blocks walk the batch's tiles. A block that has found room on an SM keeps
// it until the batch's tiles are done, instead of every one of the 138 000 tile blocks of a 64-frame batch competing
// anew with the long-lived serial decoders of the other batches in flight (which leaves about one tile block per SM).
__global__ void __launch_bounds__(256, J40B_TILE_MINB) k_back_tile_persistent(const BackWork *items, int ntiles, unsigned long long *phase) {
    extern __shared__ __align__(16) float tile_coef[];
    __shared__ TileShared ts;
    ts.phase = phase;
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
    return cudaFuncSetAttribute(k_back_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * TILE_CH * 4) == cudaSuccess &&
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

// j40-b200: kernel launch interface between the host side of the CUDA backend (j40b_cuda.cu) and the
// translation units that hold the kernels (kern_lf.cu, kern_hf.cu, kern_back.cu, kern_mod.cu; split so that
// they compile in parallel). Each launcher enqueues one kernel on `stream`.
#pragma once
#include "j40b_pipeline.h"
#include <cuda_runtime.h>
#include <stdlib.h>

namespace j40b {

#if defined(__CUDACC__)
struct BlockSync { __device__ void operator()() const { __syncthreads(); } };
#endif

// dynamic shared memory of the serial-decoder kernels: a copy of the code spec's tables
enum { SPEC_COPY_BYTES = 40 * 1024 };
// widest channel the shared-memory row path of the serial decoders takes: LF groups are at most 256 cells
// wide; modular groups at most 1024 pixels
enum { LF_ROW_CAP = 256, MOD_ROW_CAP = 1024 };
// shared memory per LF group for copies of the current channel's alias tables (8 tables of 64 entries, 2 of 256)
enum { LF_TAB_BYTES = 4096 };

// Serial modular decoders (LF image, HF metadata, modular groups): one *warp* per work item, SIMT-uniform: all
// lanes run the same sample; lane j evaluates decision node j and leaf j of the compiled MA tree (two ballots pick
// the leaf), lanes 0..3 the weighted predictor's sub-predictors, every leaf lane pre-fetches its cluster's alias
// entry. The warp's working set (sample rows, weighted-predictor error rows, reference-channel property rows, the
// compiled tree) lives in its slice of shared memory; the CTA's warps can share one staged copy of the code spec.
#if defined(__CUDACC__)
struct WarpSync {
    static constexpr bool kFull = true;
    __device__ void operator()() const { __syncwarp(); }
    __device__ uint32_t mask() const { return 0xffffffffu; }
    __device__ int shift() const { return 0; }
};
// G consecutive lanes of a warp that is shared by 32 / G streams
struct GroupSync {
    static constexpr bool kFull = false;
    uint32_t m; int s;
    __device__ void operator()() const { __syncwarp(m); }
    __device__ uint32_t mask() const { return m; }
    __device__ int shift() const { return s; }
};
#endif

// `with_wp`: the slice has the weighted predictor's error row (decoder classes without it leave it out: 5 of 12.4 KB)
// `tab_bytes`: room for copies of the current channel's alias tables (multiple of 16; 0: none)
__host__ __device__ inline size_t warp_slice_bytes(int cap, bool with_wp = true, int tab_bytes = 0) {
    size_t n = sizeof(WarpScratch) + (sizeof(SimtLane) + sizeof(SimtLeaf)) * SIMT_LANES;
    n += (size_t) cap * (3 * 2 + (with_wp ? 5 * 4 : 0) + SIMT_REF_SLOTS * 4);
    return ((n + 15) & ~(size_t) 15) + (size_t) tab_bytes;
}

#if defined(__CUDACC__)
__device__ inline ModSmem carve_warp_slice(uint8_t *base, int cap, WarpScratch *&ws, bool with_wp = true, int tab_bytes = 0) {
    ws = (WarpScratch *) base;
    ModSmem ms;
    ms.leaves = (SimtLeaf *) (base + sizeof(WarpScratch));
    ms.tab = (SimtLane *) (ms.leaves + SIMT_LANES);
    ms.wp = (int32_t *) (ms.tab + SIMT_LANES);
    ms.refp = ms.wp + (with_wp ? (size_t) cap * 5 : 0);
    ms.rows = cap ? (int16_t *) (ms.refp + (size_t) cap * SIMT_REF_SLOTS) : nullptr;
    ms.info = ws->info;
    ms.cap = cap;
    ms.lanes = SIMT_LANES;
    ms.tabs = tab_bytes ? (uint64_t *) (base + warp_slice_bytes(cap, with_wp)) : nullptr;
    ms.tabs_entries = tab_bytes / 8;
    return ms;
}
#endif

enum { HF_WARPS = 4 };
// warps per block of the lane-per-stream kernels (k_lf_lane, k_mod_lane): one, so that a launch of a few hundred streams
// spreads over as many SMs as it has warps (each warp is latency-bound and wants a scheduler and an L1 of its own)
enum { LANE_WARPS = 1 };
#if defined(__CUDACC__)
struct WarpAny { __device__ bool operator()(bool p) const { return __any_sync(0xffffffffu, p); } };
#endif

// Experiment switch: the preferred split of an SM's 256 KB between shared memory and L1, per kernel family
// (J40B_CARVEOUT_LF / _HF / _BACK / _MOD = percent of shared memory; unset: the driver's choice, which sizes shared memory
// for the most blocks the kernel's registers allow and leaves the rest to L1). Measured: 100 % everywhere changes nothing
// for a pipeline of batches and costs the coefficient kernel a third (its tables beyond the staged spec go through L1).
#if defined(__CUDACC__)
template <class F> inline bool kl_carveout(F *kernel, const char *env) {
    const char *e = getenv(env);
    return !e || atoi(e) < 0 || cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(e)) == cudaSuccess;
}
#endif
// where in a grid of `grid` blocks batch object number `turn` starts its work list: multiples of the golden ratio, which
// spread any number of objects evenly (see k_lf_chan)
inline int kl_rotation(int turn, int grid) {
    const double x = (double) turn * 0.6180339887498949;
    return (int) ((x - (double) (long long) x) * (double) grid) % (grid > 0 ? grid : 1);
}
bool kl_init_lf();   // raises the dynamic shared-memory limits of the kernels in that translation unit
bool kl_init_hf();
bool kl_init_back();
bool kl_init_mod();
void kl_lf_stage(int stage, cudaStream_t stream, const LfWork *w, int n, int cap, int c0, int c1, bool split, int lanes, int spread, int turn); // 5 class kernels per channel
void kl_lf_lane(int stage, cudaStream_t stream, const LfWork *w, int n);
void kl_lf_place(int n, cudaStream_t stream, const LfWork *w);
void kl_lf_post(int n, cudaStream_t stream, const LfWork *w);
void kl_lf_llf(int n, cudaStream_t stream, const LfWork *w);
void kl_hf_prep(int n, cudaStream_t stream, const HfPrepWork *w);
void kl_hf_group(int blocks, size_t smem, cudaStream_t stream, const HfWork *w, int n, int lanes, int spec_cap, int spread, int turn);
void kl_back_tile(int n, cudaStream_t stream, const BackWork *w);
void kl_back_generic(int blocks, cudaStream_t stream, const BackWork *w, int n, float *pool);
void kl_lf_smhist_dump();   // J40B_LF_SMHIST=1: where the serial LF kernels' warps ran (per SM)
void kl_back_phase_dump(); // diagnostic builds (make PHASE_CLOCKS=1): prints and clears the tile kernel's phase counters
void kl_dump_coeffs(int n, cudaStream_t stream, const DumpWork &w); // diagnostics
void kl_modular(int n, cudaStream_t stream, ModWork *w, int cap, int spec_cap);
void kl_mod_lane(int n, cudaStream_t stream, ModWork *w);
void kl_palette_delta(cudaStream_t stream, const RenderWork *w, int num_c);
void kl_render(cudaStream_t stream, const RenderWork *w, int width, int height);

} // namespace j40b

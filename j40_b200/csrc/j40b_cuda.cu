// j40-b200: CUDA kernels (sm_100a), the CUDA backend of the batch pipeline, and the C ABI:
// the ten j40_* entry points of the reference (include/j40.h) plus the j40b_* batch extension
// (include/j40b.h). There is no CPU decoding path in this library: without a usable GPU every decode
// fails loudly with the error code "!gpu".
#include "j40b_kernels.h"
#include "../../include/j40.h"
#include "../../include/j40b.h"
#include <cuda_runtime.h>
#include <atomic>
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include <memory>
#include <vector>

using namespace j40b;

#define EXPORT extern "C" __attribute__((visibility("default")))

static bool cuda_ok(cudaError_t e) { return e == cudaSuccess; }

struct CudaBackend {
    int device = 0;
    cudaStream_t stream = nullptr, side = nullptr; // side: highest priority, for the short generic back-end kernel
    cudaEvent_t ev_side[2] = {nullptr, nullptr};
    cudaStream_t side2 = nullptr;                // normal priority: the sharpness channel of the LF groups (launch_lf)
    cudaEvent_t ev_side2[2] = {nullptr, nullptr};
    bool side2_pending = false;
    bool ok = false;
    int num_sms = 148;
    float *big_pool = nullptr;
    int big_blocks = 0;
    int64_t launches = 0;
    int last_lf_lanes = 0;
    int turn = 0; // which of the process's batch objects this is: rotates where its kernels' blocks start (kern_lf.cu)
    cudaEvent_t ev[10] = {nullptr};
    bool mod_marked = false;
    float kernel_ms[9] = {0};

    bool init(int dev) {
        device = dev;
        int count = 0;
        if (!cuda_ok(cudaGetDeviceCount(&count)) || dev < 0 || dev >= count) return false;
        if (!cuda_ok(cudaSetDevice(dev))) return false;
        cudaDeviceProp prop;
        if (!cuda_ok(cudaGetDeviceProperties(&prop, dev))) return false;
        num_sms = prop.multiProcessorCount;
        { static std::atomic<int> next_turn{0}; turn = next_turn++; }
        if (!cuda_ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking))) return false;
        {
            int least = 0, greatest = 0;
            cudaDeviceGetStreamPriorityRange(&least, &greatest);
            if (!cuda_ok(cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, greatest))) { side = nullptr; cudaGetLastError(); }
            for (auto &e : ev_side) if (!cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming))) return false;
            if (!cuda_ok(cudaStreamCreateWithFlags(&side2, cudaStreamNonBlocking))) { side2 = nullptr; cudaGetLastError(); }
            for (auto &e : ev_side2) if (!cuda_ok(cudaEventCreateWithFlags(&e, cudaEventDisableTiming))) return false;
        }
        for (auto &e : ev) if (!cuda_ok(cudaEventCreate(&e))) return false;
        if (!kl_init_lf() || !kl_init_hf() || !kl_init_back() || !kl_init_mod()) return false;
        ok = true;
        return true;
    }
    void destroy() {
        if (!ok) return;
        cudaSetDevice(device);
        if (big_pool) cudaFree(big_pool);
        for (auto &e : ev) if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
        if (side) cudaStreamDestroy(side);
        if (side2) cudaStreamDestroy(side2);
        for (auto &e : ev_side2) if (e) cudaEventDestroy(e);
        for (auto &e : ev_side) if (e) cudaEventDestroy(e);
        ok = false;
    }
    void *dev_alloc(size_t n) { void *p = nullptr; cudaSetDevice(device); if (!cuda_ok(cudaMalloc(&p, n ? n : 1))) return nullptr; return p; }
    void dev_free(void *p) { cudaSetDevice(device); cudaFree(p); }
    // staging memory: pinned if the driver grants it, plain otherwise (the kind is remembered per pointer)
    void *host_alloc(size_t n) {
        void *p = nullptr;
        if (cuda_ok(cudaHostAlloc(&p, n ? n : 1, cudaHostAllocDefault))) return p;
        cudaGetLastError();
        p = malloc(n ? n : 1);
        if (p) unpinned.push_back(p);
        return p;
    }
    void host_free(void *p) {
        for (size_t i = 0; i < unpinned.size(); ++i) if (unpinned[i] == p) { unpinned.erase(unpinned.begin() + (long) i); free(p); return; }
        cudaFreeHost(p);
    }
    std::vector<void *> unpinned;
    void h2d(void *d, const void *s, size_t n) { cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, stream); }
    void d2h(void *d, const void *s, size_t n) { cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, stream); cudaStreamSynchronize(stream); }
    void d2h_async(void *d, const void *s, size_t n) { cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, stream); }
    void dev_memset(void *d, int v, size_t n) { cudaMemsetAsync(d, v, n, stream); }
    void sync() { cudaStreamSynchronize(stream); }
    int lane_stride() const { return 32 * LANE_WARPS; } // work items per slot of the lane decoders' interleaved buffers

    // J40B_DEBUG_SKIP (measurement aid, bench.py --debug-skip; results are wrong with it): bit 0 leaves out the tile kernel,
    // bit 1 the coefficient kernel, bit 2 the LF-group stage kernels -- what each stage costs the others in a pipeline
    static int debug_skip() { const char *e = getenv("J40B_DEBUG_SKIP"); return e ? atoi(e) : 0; }

    void launch_lf(const LfWork *w, int n, size_t, bool split_ok, int tree_lanes) {
        bool split = split_ok;
        if (debug_skip() & 4) { for (int i : {0, 5, 6, 1}) cudaEventRecord(ev[i], stream); return; }
        // One stream per warp (SIMT-uniform decoder, j40b_modular.h; per channel and decoder class, lf_chan_body) or,
        // with J40B_LF_MODE=lane, one per lane (j40b_modlane.h): 1/20 of the issue slots and no shared memory, but 6-7
        // times the latency per stream (measured: LF image of 64 4K frames 70 ms against 490 ms), which a pipeline can
        // only hide with more batches in flight than fit the device memory at 6 GB per batch of 64 4K frames.
        bool lane_mode = false;
        if (const char *e = getenv("J40B_LF_MODE")) lane_mode = e[0] == 'l';
        cudaEventRecord(ev[0], stream);
        // The sharpness channel off the critical path pays for latency (one 4K image: 210 -> 178 ms), not for throughput: with
        // a dozen batches in flight the extra concurrent kernels cost 8 % (47.7 against 44.0 ms per step). Small launches only.
        split = split && n < 64;
        if (const char *e = getenv("J40B_LF_SPLIT")) split = split_ok && atoi(e) != 0; // (1: also for large launches, 0: never)
        if (lane_mode || !side2) split = false;
        // Lanes per LF group (kern_lf.cu): 32 -- one LF group per warp -- unless J40B_LF_LANES=16|8 asks for 2 or 4 LF groups
        // per warp (the compiled trees must fit that many lanes). Measured with a dozen 64-frame batches in flight: 4 groups
        // per warp need a quarter of the warp instructions and gain 2 % of the pipelined step (43.4 against 44.3 ms), because
        // the step is not bound by issue slots; alone the stages take 24 % longer (a shared warp is as slow as its slowest
        // stream, and carries the union of their branches). Not the default for that.
        int lanes = 32;
        if (const char *e = getenv("J40B_LF_LANES")) { const int v = atoi(e); if ((v == 8 || v == 16 || v == 32) && tree_lanes <= v) lanes = v; }
        last_lf_lanes = lane_mode ? 1 : lanes;
        if (lane_mode) { kl_lf_lane(1, stream, w, n); ++launches; } else { kl_lf_stage(0, stream, w, n, LF_ROW_CAP, 0, 3, false, lanes, num_sms, turn); launches += 15; }
        cudaEventRecord(ev[5], stream);
        kl_lf_post(n, stream, w);
        if (lane_mode) { kl_lf_lane(2, stream, w, n); kl_lf_place(n, stream, w); launches += 2; }
        else if (!split) { kl_lf_stage(1, stream, w, n, LF_ROW_CAP, 0, 4, false, lanes, num_sms, turn); launches += 20; }
        else {
            // the sharpness channel (a third of the stage, never used by j40) on a side stream, next to the LLF, coefficient
            // and tile stages; join_side() makes the end of the decode wait for it
            kl_lf_stage(1, stream, w, n, LF_ROW_CAP, 0, 3, true, lanes, num_sms, turn);
            cudaEventRecord(ev_side2[0], stream);
            cudaStreamWaitEvent(side2, ev_side2[0], 0);
            kl_lf_stage(1, side2, w, n, LF_ROW_CAP, 3, 4, true, lanes, num_sms, turn);
            cudaEventRecord(ev_side2[1], side2);
            side2_pending = true;
            launches += 20;
        }
        cudaEventRecord(ev[6], stream);
        kl_lf_llf(n, stream, w);
        cudaEventRecord(ev[1], stream);
        launches += 2;
    }
    // lanes (sections) per block of the coefficient kernel for a launch of n sections: HF_WARPS warps of `lanes` each --
    // full warps as soon as that still leaves two warps per SM (a warp of 32 lanes costs the same issue slots per iteration
    // as one of 8) and fewer lanes per warp for small launches (a single 4K image has 135 sections), where latency is
    // what counts
    int hf_block_lanes(int n) const {
        int lanes = (n + num_sms * 2 - 1) / (num_sms * 2);
        lanes = lanes < 4 ? 4 : lanes > 32 ? 32 : lanes;
        if (const char *e = getenv("J40B_HF_LANES")) { int v = atoi(e); if (v >= 1 && v <= 32) lanes = v; }
        return HF_WARPS * lanes;
    }
    // `n`: sections including the padding of the work list to whole blocks of `per_block` lanes (Batch::upload)
    void launch_hf(const HfPrepWork *pw, int ngroups, const HfWork *w, int n, size_t spec_bytes, int per_block) {
        kl_hf_prep(ngroups, stream, pw);
        ++launches;
        int spec_cap = spec_bytes <= SPEC_COPY_BYTES ? (int) ((spec_bytes + 15) & ~(size_t) 15) : 0;
        if (const char *e = getenv("J40B_HF_STAGE")) if (!atoi(e)) spec_cap = 0; // experiment: tables through L1, no shared memory
        const int lanes = per_block / HF_WARPS;
        if (!(debug_skip() & 2)) kl_hf_group((n + per_block - 1) / per_block, (size_t) spec_cap, stream, w, n, lanes, spec_cap, num_sms, turn);
        cudaEventRecord(ev[2], stream);
        ++launches;
    }
    void launch_back(const BackWork *w, int n) {
        if (!big_pool) {
            big_blocks = num_sms;
            if (!cuda_ok(cudaMalloc(&big_pool, (size_t) big_blocks * 4 * 65536 * sizeof(float)))) { big_pool = nullptr; big_blocks = 0; }
        }
        // The generic path (varblocks the tile kernel leaves out) next to the tile kernel, on a side stream of the highest
        // priority: it usually has nothing to do and ends in microseconds, but behind the tile kernel on a saturated GPU
        // its 148 blocks waited for slots for as long as a whole tile stage (83 ms measured with 12 batches in flight).
        if (big_pool && side) {
            cudaEventRecord(ev_side[0], stream);
            cudaStreamWaitEvent(side, ev_side[0], 0);
            kl_back_generic(big_blocks < n ? big_blocks : n, side, w, n, big_pool);
            cudaEventRecord(ev_side[1], side);
            ++launches;
        }
        // J40B_TILE_PRIO=1 (experiment): the tile kernel on the high-priority stream as well (short-lived blocks that
        // finish a batch; they compete for slots with the long-lived serial decoders of the other batches in flight)
        static const bool tile_prio = getenv("J40B_TILE_PRIO") && atoi(getenv("J40B_TILE_PRIO")) != 0;
        if (tile_prio && big_pool && side) {
            kl_back_tile(n, side, w);
            cudaEventRecord(ev_side[1], side);
        } else if (!(debug_skip() & 1)) {
            kl_back_tile(n, stream, w);
        }
        cudaEventRecord(ev[3], stream);
        ++launches;
        if (big_pool && side) cudaStreamWaitEvent(stream, ev_side[1], 0);
        else if (big_pool) { kl_back_generic(big_blocks < n ? big_blocks : n, stream, w, n, big_pool); ++launches; }
        cudaEventRecord(ev[4], stream);
    }
    // `spec_bytes`: the image's code-spec blob; `max_w`: its widest channel (sizes the shared-memory rows)
    void launch_mod(ModWork *w, int n, size_t spec_bytes, int max_w) {
        bool lane_mode = false; // (see launch_lf; 8192x8192 lossless: 183 ms per frame against 1019 ms)
        if (const char *e = getenv("J40B_LF_MODE")) lane_mode = e[0] == 'l';
        if (lane_mode) { kl_mod_lane(n, stream, w); ++launches; return; }
        int cap = (max_w + 63) & ~63;
        if (cap > MOD_ROW_CAP || cap <= 0) cap = MOD_ROW_CAP;
        int spec_cap = spec_bytes <= SPEC_COPY_BYTES ? (int) ((spec_bytes + 15) & ~(size_t) 15) : 0;
        // the staged code spec shortens each symbol a little but costs residency: when the groups do not fit the
        // GPU in one wave with it, leave the tables to L1 (8192x8192: 1024 one-warp blocks, 4 vs 10 per SM)
        const size_t slice = warp_slice_bytes(cap), sm_bytes = 227 * 1024;
        const size_t fit_staged = sm_bytes / (slice + (size_t) spec_cap + 1024), fit_plain = sm_bytes / (slice + 1024);
        if (spec_cap && (size_t) n > fit_staged * (size_t) num_sms && fit_plain > fit_staged) spec_cap = 0;
        kl_modular(n, stream, w, cap, spec_cap);
        ++launches;
    }
    void launch_dump(const DumpWork &w, int n) { kl_dump_coeffs(n, stream, w); ++launches; }
    void join_side() { if (side2_pending) { cudaStreamWaitEvent(stream, ev_side2[1], 0); side2_pending = false; } }
    void mark_modular(int which) { cudaEventRecord(ev[8 + which], stream); mod_marked = true; }
    void launch_palette_delta(const RenderWork *w, int num_c) { kl_palette_delta(stream, w, num_c); ++launches; }
    void launch_render(const RenderWork *w, int width, int height) {
        kl_render(stream, w, width, height);
        ++launches;
    }
};

// ---------------------------------------------------------------------------------------------
// batch object

struct j40b_batch {
    CudaBackend be;
    Batch<CudaBackend> *batch = nullptr;
    std::vector<std::pair<const uint8_t *, size_t>> inputs;
    bool uploaded = false, decoded = false;
    cudaEvent_t t0 = nullptr, t1 = nullptr, m[2] = {nullptr, nullptr}, jev = nullptr;
    float last_ms = 0;
    int64_t last_launches = 0;
    uint8_t *read_dst = nullptr; // pending j40b_batch_read_all_async target (re-issued if the decode is retried)
    size_t read_pitch = 0;
};

EXPORT int j40b_gpu_available(void) {
    int count = 0;
    return cuda_ok(cudaGetDeviceCount(&count)) && count > 0;
}

EXPORT j40b_batch *j40b_batch_create(int device) {
    j40b_batch *b = new j40b_batch;
    if (!b->be.init(device)) { delete b; return nullptr; }
    cudaEventCreate(&b->t0);
    cudaEventCreate(&b->t1);
    b->batch = new Batch<CudaBackend>(b->be);
    return b;
}

EXPORT void j40b_batch_destroy(j40b_batch *b) {
    if (!b) return;
    cudaSetDevice(b->be.device);
    delete b->batch;
    if (b->t0) cudaEventDestroy(b->t0);
    if (b->t1) cudaEventDestroy(b->t1);
    if (b->m[0]) cudaEventDestroy(b->m[0]);
    if (b->m[1]) cudaEventDestroy(b->m[1]);
    if (b->jev) cudaEventDestroy(b->jev);
    b->be.destroy();
    delete b;
}

EXPORT int j40b_batch_add(j40b_batch *b, const void *buf, size_t size) {
    if (!b || !buf) return -1;
    b->inputs.push_back({(const uint8_t *) buf, size});
    b->batch->add((const uint8_t *) buf, size);
    b->uploaded = false;
    return (int) b->batch->plans.size() - 1;
}

EXPORT int j40b_batch_add_many(j40b_batch *b, const void *const *bufs, const size_t *sizes, int n, int threads) {
    if (!b || !bufs || !sizes || n < 0) return -1;
    for (int i = 0; i < n; ++i) if (!bufs[i]) return -1;
    const int first = (int) b->batch->plans.size();
    for (int i = 0; i < n; ++i) b->inputs.push_back({(const uint8_t *) bufs[i], sizes[i]});
    b->batch->host_threads = threads;
    b->batch->add_many((const uint8_t *const *) bufs, sizes, (size_t) n, threads);
    b->uploaded = false;
    return first;
}

EXPORT int j40b_batch_reset(j40b_batch *b) {
    if (!b) return -1;
    cudaSetDevice(b->be.device);
    b->be.sync();
    b->batch->reset();
    b->batch->full_token_cap = false;
    b->inputs.clear();
    b->uploaded = b->decoded = false;
    b->read_dst = nullptr;
    return 0;
}

EXPORT int j40b_batch_upload(j40b_batch *b) {
    if (!b) return -1;
    cudaSetDevice(b->be.device);
    bool ok = b->batch->upload();
    b->uploaded = ok;
    b->decoded = false;
    b->read_dst = nullptr;
    return ok && cudaGetLastError() == cudaSuccess ? 0 : -1;
}

EXPORT int j40b_batch_read_all_async(j40b_batch *b, void *dst, size_t pitch) {
    if (!b || !dst || !b->decoded) return -1;
    cudaSetDevice(b->be.device);
    b->read_dst = (uint8_t *) dst;
    b->read_pitch = pitch;
    b->batch->download_all_async(b->read_dst, pitch);
    return 0;
}

EXPORT int j40b_batch_decode(j40b_batch *b) {
    if (!b || !b->uploaded) return -1;
    cudaSetDevice(b->be.device);
    int64_t before = b->be.launches;
    cudaEventRecord(b->t0, b->be.stream);
    b->batch->execute();
    cudaEventRecord(b->t1, b->be.stream);
    b->last_launches = b->be.launches - before;
    b->decoded = true;
    return 0;
}

static void gather_times(j40b_batch *b) {
    cudaEventElapsedTime(&b->last_ms, b->t0, b->t1);
    CudaBackend &be = b->be;
    for (float &k : be.kernel_ms) k = 0;
    // events are only recorded when the corresponding kernels were launched in this decode
    bool vardct = false;
    for (auto &p : b->batch->plans) if (!p->err && !p->df.is_modular && !p->lfg_sec.empty() && !p->pg_sec.empty()) vardct = true;
    if (be.mod_marked) cudaEventElapsedTime(&be.kernel_ms[4], be.ev[8], be.ev[9]); // k_modular + k_render of all modular images
    be.mod_marked = false;
    if (vardct) {
        cudaEventElapsedTime(&be.kernel_ms[0], be.ev[0], be.ev[1]);
        cudaEventElapsedTime(&be.kernel_ms[1], be.ev[1], be.ev[2]);
        cudaEventElapsedTime(&be.kernel_ms[2], be.ev[2], be.ev[3]);
        cudaEventElapsedTime(&be.kernel_ms[3], be.ev[3], be.ev[4]);
        cudaEventElapsedTime(&be.kernel_ms[6], be.ev[0], be.ev[5]); // k_lf_decode<1>
        cudaEventElapsedTime(&be.kernel_ms[7], be.ev[5], be.ev[6]); // k_lf_post + k_lf_decode<2>
        cudaEventElapsedTime(&be.kernel_ms[8], be.ev[6], be.ev[1]); // k_lf_llf
    }
}

EXPORT int j40b_batch_wait(j40b_batch *b) {
    if (!b || !b->decoded) return -1;
    cudaSetDevice(b->be.device);
    b->batch->collect_errors();
    cudaError_t ce = cudaGetLastError();
    gather_times(b);
    if (getenv("J40B_PHASE_DUMP")) kl_back_phase_dump();
    if (getenv("J40B_LF_SMHIST_DUMP")) kl_lf_smhist_dump();
    // internal conditions: a token arena that turned out too small (redo with worst-case capacity), LF-group
    // sub-bitstreams with trees of their own (the host reads them where the device found them; at most one round per
    // stage of an LF group)
    for (int round = 0; round < 6; ++round) {
        bool retry = false;
        for (auto &r : b->batch->results) if (r.err == E_TOKV && !b->batch->full_token_cap) retry = true;
        if (retry) b->batch->full_token_cap = true;
        bool ltre = false;
        for (auto &r : b->batch->results) if (r.err == E_LTRE) ltre = true;
        if (ltre && b->batch->resolve_local_trees()) retry = true;
        if (!retry) break;
        if (b->batch->upload()) {
            cudaEventRecord(b->t0, b->be.stream);
            b->batch->execute();
            cudaEventRecord(b->t1, b->be.stream);
            if (b->read_dst) b->batch->download_all_async(b->read_dst, b->read_pitch);
        }
        b->batch->collect_errors();
        gather_times(b); // describe the decode that produced the result, not the aborted first attempt
    }
    int failed = 0;
    for (auto &r : b->batch->results) failed += r.err != 0;
    if (ce != cudaSuccess) {
        for (auto &r : b->batch->results) if (!r.err) r.err = J40B_4CC('!', 'g', 'p', 'u');
        return (int) b->batch->results.size();
    }
    return failed;
}

EXPORT int j40b_batch_count(const j40b_batch *b) { return b ? (int) b->batch->plans.size() : 0; }
EXPORT uint32_t j40b_batch_error(const j40b_batch *b, int i) {
    if (!b || i < 0 || i >= (int) b->batch->plans.size()) return J40B_4CC('U', 'i', 'd', 'x');
    if (i < (int) b->batch->results.size()) return b->batch->results[(size_t) i].err;
    return b->batch->plans[(size_t) i]->err;
}
EXPORT int j40b_batch_info(const j40b_batch *b, int i, int32_t *w, int32_t *h, int32_t *stride) {
    if (!b || i < 0 || i >= (int) b->batch->results.size()) return -1;
    const ImageResult &r = b->batch->results[(size_t) i];
    if (w) *w = r.width;
    if (h) *h = r.height;
    if (stride) *stride = r.stride;
    return 0;
}
EXPORT const void *j40b_batch_device_pixels(const j40b_batch *b, int i) {
    if (!b || i < 0 || i >= (int) b->batch->results.size() || b->batch->results[(size_t) i].err) return nullptr;
    return b->batch->device_pixels((size_t) i);
}
EXPORT int j40b_batch_read_pixels(j40b_batch *b, int i, void *dst) {
    if (!b || !dst || i < 0 || i >= (int) b->batch->results.size() || b->batch->results[(size_t) i].err) return -1;
    cudaSetDevice(b->be.device);
    b->batch->download_pixels((size_t) i, (uint8_t *) dst);
    return 0;
}
EXPORT size_t j40b_batch_debug_dump(j40b_batch *b, int i, int lf_group, int what, void *dst, size_t cap) {
    if (!b || !dst || !b->decoded || i < 0 || lf_group < 0) return 0;
    cudaSetDevice(b->be.device);
    b->be.sync();
    return b->batch->debug_dump((size_t) i, (size_t) lf_group, what, dst, cap);
}
// A consumer of the device-resident output: image i as a PAM file (P7, RGB_ALPHA), rows copied from
// device memory without the stride padding (cudaMemcpy2D: device pitch -> tight host rows).
EXPORT int j40b_batch_write_pam(j40b_batch *b, int i, const char *path) {
    if (!b || !path || !b->decoded || i < 0 || i >= (int) b->batch->results.size() || b->batch->results[(size_t) i].err) return -1;
    cudaSetDevice(b->be.device);
    const ImageResult &r = b->batch->results[(size_t) i];
    const size_t row = (size_t) r.width * 4;
    std::vector<uint8_t> tight(row * (size_t) r.height);
    b->be.sync();
    if (!cuda_ok(cudaMemcpy2D(tight.data(), row, b->batch->device_pixels((size_t) i), (size_t) r.stride, row, (size_t) r.height, cudaMemcpyDeviceToHost))) return -1;
    FILE *fp = fopen(path, "wb");
    if (!fp) return -1;
    fprintf(fp, "P7\nWIDTH %d\nHEIGHT %d\nDEPTH 4\nMAXVAL 255\nTUPLTYPE RGB_ALPHA\nENDHDR\n", r.width, r.height);
    const bool ok = fwrite(tight.data(), 1, tight.size(), fp) == tight.size();
    return fclose(fp) == 0 && ok ? 0 : -1;
}
EXPORT float j40b_batch_last_decode_ms(const j40b_batch *b) { return b ? b->last_ms : 0.0f; }

// Several batches (each with its own stream) can be in flight at once, e.g. to overlap the latency-bound
// LF-group decode of one batch with the HF / back kernels of another. These three calls time such a
// region on the device: mark(b, 0) ... enqueue on any batches ... join(b, other) for every other batch,
// mark(b, 1); after the batches have been waited for, j40b_batch_mark_ms(b) is the elapsed device time.
EXPORT int j40b_batch_mark(j40b_batch *b, int which) {
    if (!b || which < 0 || which > 1) return -1;
    cudaSetDevice(b->be.device);
    if (!b->m[which]) cudaEventCreate(&b->m[which]);
    return cudaEventRecord(b->m[which], b->be.stream) == cudaSuccess ? 0 : -1;
}
EXPORT int j40b_batch_join(j40b_batch *b, j40b_batch *other) {
    if (!b || !other || b->be.device != other->be.device) return -1;
    cudaSetDevice(b->be.device);
    if (!other->jev) cudaEventCreateWithFlags(&other->jev, cudaEventDisableTiming);
    cudaEventRecord(other->jev, other->be.stream);
    return cudaStreamWaitEvent(b->be.stream, other->jev, 0) == cudaSuccess ? 0 : -1;
}
// Makes everything enqueued on b from now on wait until `other`'s most recently enqueued decode has finished
// stage `stage` (0 = LF stage, 1 = HF stage, 2 = everything). A serving loop uses it to keep batches in
// flight out of phase, so that the latency-bound LF decoders of some overlap the HF / tile kernels of others.
EXPORT int j40b_batch_after(j40b_batch *b, j40b_batch *other, int stage) {
    if (!b || !other || b->be.device != other->be.device || stage < 0 || stage > 2 || !other->decoded) return -1;
    cudaSetDevice(b->be.device);
    const int which = stage == 0 ? 1 : stage == 1 ? 2 : 4;
    return cudaStreamWaitEvent(b->be.stream, other->be.ev[which], 0) == cudaSuccess ? 0 : -1;
}
EXPORT float j40b_batch_mark_ms(j40b_batch *b) {
    float ms = 0;
    if (!b || !b->m[0] || !b->m[1]) return 0;
    cudaSetDevice(b->be.device);
    cudaEventSynchronize(b->m[1]);
    cudaEventElapsedTime(&ms, b->m[0], b->m[1]);
    return ms;
}
// diagnostic: device time from `ref`'s mark 0 to event `which` of b's last decode (0 LF start, 5 LF image decoded,
// 6 HF metadata decoded, 1 LF done, 2 HF done, 3 tiles done, 4 all done); shows how batches in flight interleave
EXPORT float j40b_batch_event_ms(const j40b_batch *b, const j40b_batch *ref, int which) {
    float ms = -1;
    if (!b || !ref || !ref->m[0] || which < 0 || which > 6) return ms;
    cudaSetDevice(b->be.device);
    if (cudaEventElapsedTime(&ms, ref->m[0], b->be.ev[which]) != cudaSuccess) { cudaGetLastError(); return -1; }
    return ms;
}
EXPORT float j40b_batch_kernel_ms(const j40b_batch *b, int which) { return b && which >= 0 && which < 9 ? b->be.kernel_ms[which] : 0.0f; }
EXPORT int64_t j40b_batch_stat(const j40b_batch *b, int what) {
    if (!b) return 0;
    switch (what) {
    case 0: return (int64_t) b->batch->device_bytes();
    case 1: return (int64_t) b->batch->h2d_bytes();
    case 2: return b->last_launches;
    case 3: { int64_t s = 0; for (auto &p : b->batch->plans) s += (int64_t) p->cs_size; return s; }
    case 4: { int64_t s = 0; for (auto &r : b->batch->results) if (!r.err) s += (int64_t) r.width * r.height; return s; }
    case 5: return b->batch->be.last_lf_lanes;
    default: return 0;
    }
}

// ---------------------------------------------------------------------------------------------
// the reference's public API (j40.h:233-272, implementation j40.h:8243-8480)

enum : uint32_t {
    MAGIC_IMAGE = 0x6a3440b2u,       // "valid handle"
    MAGIC_IMAGE_ERR = 0x9d51c3a7u,   // ^ origin: handle carrying only an error code
    MAGIC_IMAGE_OPEN_ERR = 0x42e07f19u, // ^ origin: fopen failed, saved errno in u.saved_errno
    MAGIC_FRAME = 0x1f8b22c4u,
    MAGIC_FRAME_ERR = 0x77a0d9e5u,
    MAGIC_INNER = 0x5cb0a1d3u,
};
enum Origin { O_NONE = 0, O_NEXT, O_FROM_FILE, O_FROM_MEMORY, O_OUTPUT_FORMAT, O_NEXT_FRAME, O_CURRENT_FRAME, O_FRAME_PIXELS, O_ERROR_STRING, O_FREE };
static const char *const ORIGIN_NAMES[] = {"(unknown)", nullptr, "from_file", "from_memory", "output_format", "next_frame", "current_frame", "frame_pixels_*", "error_string", "free"};
enum { O_LAST_ALT_MAGIC = O_FROM_MEMORY };

struct j40__inner {
    uint32_t magic;
    int origin;
    j40_err err;
    int saved_errno;
    char errbuf[256];
    // source
    const uint8_t *data;
    size_t size;
    void *owned;                 // malloc'ed file contents (from_file)
    void *user_buf;
    j40_memory_free_func freefunc;
    // result
    int advanced, rendered;
    int32_t width, height, stride;
    uint8_t *pixels;             // pinned host memory
    bool pixels_pinned;
};

static const struct { char err[5]; const char *msg; } ERROR_STRINGS[] = { // j40.h:8004-8028
    {"Upt0", "`path` parameter is NULL"}, {"Ubf0", "`buf` parameter is NULL"}, {"Uch?", "Bad `channel` parameter"},
    {"Ufm?", "Bad `format` parameter"}, {"Uof?", "Bad `channel` and `format` combination"}, {"Urnd", "Frame is not yet rendered"},
    {"Ufre", "Trying to reuse already freed image"}, {"!mem", "Out of memory"}, {"!jxl", "The JPEG XL signature is not found"},
    {"open", "Failed to open file"}, {"bigg", "Image dimensions are too large to handle"}, {"flen", "File is too lengthy to handle"},
    {"shrt", "Premature end of file"}, {"slim", "Image size limit reached"}, {"elim", "Extra channel number limit reached"},
    {"xlim", "Modular transform limit reached"}, {"tlim", "Meta-adaptive tree size or depth limit reached"},
    {"plim", "ICC profile length limit reached"}, {"fbpp", "Given bits per pixel value is disallowed"},
    {"fblk", "Black extra channel is disallowed"}, {"fm32", "32-bit buffers for modular encoding are disallowed"},
    {"TODO", "Unimplemented feature encountered"}, {"TEST", "Testing-only error occurred"},
    {"!gpu", "No usable CUDA device (j40-b200 has no CPU decoding path)"},
};
#define ERR4(s) J40B_4CC((s)[0], (s)[1], (s)[2], (s)[3])

static j40_err set_alt_magic(j40_err err, int saved_errno, int origin, j40_image *image) {
    if (err == ERR4("open")) {
        image->magic = MAGIC_IMAGE_OPEN_ERR ^ (uint32_t) origin;
        image->u.saved_errno = saved_errno;
        return err;
    }
    image->magic = MAGIC_IMAGE_ERR ^ (uint32_t) origin;
    return image->u.err = err;
}

static j40_err check_image(j40_image *image, int neworigin, j40__inner **out) {
    *out = nullptr;
    if (!image) return ERR4("Uim0");
    if (image->magic != MAGIC_IMAGE) {
        uint32_t origin = image->magic ^ MAGIC_IMAGE_ERR;
        if (0 < origin && origin <= O_LAST_ALT_MAGIC) {
            if (origin == O_NEXT && neworigin) image->magic = MAGIC_IMAGE_ERR ^ (uint32_t) neworigin;
            return image->u.err;
        }
        origin = image->magic ^ MAGIC_IMAGE_OPEN_ERR;
        if (0 < origin && origin <= O_LAST_ALT_MAGIC) return ERR4("open");
        return ERR4("Uim?");
    }
    if (!image->u.inner || image->u.inner->magic != MAGIC_INNER) return ERR4("Uim?");
    *out = image->u.inner;
    return image->u.inner->err;
}

static void free_inner(j40__inner *inner) {
    if (inner->freefunc) inner->freefunc(inner->user_buf);
    free(inner->owned);
    if (inner->pixels) { if (inner->pixels_pinned) cudaFreeHost(inner->pixels); else free(inner->pixels); }
    inner->magic = 0;
    free(inner);
}

EXPORT j40_err j40_error(const j40_image *image) {
    j40__inner *inner;
    return check_image((j40_image *) image, O_NONE, &inner);
}

EXPORT const char *j40_error_string(const j40_image *image) {
    static char static_errbuf[256];
    uint32_t origin = O_NONE;
    j40_err err = 0;
    char *buf = nullptr;
    int saved_errno = 0, corrupted = 0;
    if (!image) {
        snprintf(static_errbuf, sizeof static_errbuf, "`image` parameter is NULL during j40_error_string");
        return static_errbuf;
    }
    if (image->magic == MAGIC_IMAGE) {
        if (image->u.inner && image->u.inner->magic == MAGIC_INNER) {
            origin = (uint32_t) image->u.inner->origin;
            err = image->u.inner->err;
            buf = image->u.inner->errbuf;
            saved_errno = image->u.inner->saved_errno;
        } else corrupted = 1;
    } else {
        origin = image->magic ^ MAGIC_IMAGE_ERR;
        if (0 < origin && origin <= O_LAST_ALT_MAGIC) {
            err = image->u.err;
            buf = static_errbuf;
            if (origin == O_NEXT) origin = O_ERROR_STRING;
        } else {
            origin = image->magic ^ MAGIC_IMAGE_OPEN_ERR;
            if (0 < origin && origin <= O_LAST_ALT_MAGIC) {
                err = ERR4("open");
                buf = static_errbuf;
                saved_errno = image->u.saved_errno;
            } else corrupted = 1;
        }
    }
    if (corrupted) {
        snprintf(static_errbuf, sizeof static_errbuf, "`image` parameter is found corrupted during j40_error_string");
        return static_errbuf;
    }
    const char *msg = nullptr;
    for (const auto &e : ERROR_STRINGS) if (err == ERR4(e.err)) { msg = e.msg; break; }
    if (!msg) {
        snprintf(buf, 256, "Decoding failed (%c%c%c%c) during j40_%s", err >> 24 & 0xff, err >> 16 & 0xff, err >> 8 & 0xff, err & 0xff, ORIGIN_NAMES[origin]);
    } else if (saved_errno) {
        snprintf(buf, 256, "%s during j40_%s: %s", msg, ORIGIN_NAMES[origin], strerror(saved_errno));
    } else {
        snprintf(buf, 256, "%s during j40_%s", msg, ORIGIN_NAMES[origin]);
    }
    return buf;
}

static j40__inner *new_inner() {
    j40__inner *inner = (j40__inner *) calloc(1, sizeof(j40__inner));
    if (inner) inner->magic = MAGIC_INNER;
    return inner;
}

EXPORT j40_err j40_from_memory(j40_image *image, void *buf, size_t size, j40_memory_free_func freefunc) {
    if (!image) return ERR4("Uim0");
    if (!buf) return set_alt_magic(ERR4("Ubf0"), 0, O_FROM_MEMORY, image);
    j40__inner *inner = new_inner();
    if (!inner) return set_alt_magic(ERR4("!mem"), 0, O_FROM_MEMORY, image);
    if (size > (uint64_t) INT64_MAX) { free_inner(inner); return set_alt_magic(ERR4("flen"), 0, O_FROM_MEMORY, image); }
    inner->data = (const uint8_t *) buf;
    inner->size = size;
    inner->user_buf = buf;
    inner->freefunc = freefunc;
    image->magic = MAGIC_IMAGE;
    image->u.inner = inner;
    return 0;
}

EXPORT j40_err j40_from_file(j40_image *image, const char *path) {
    if (!image) return ERR4("Uim0");
    if (!path) return set_alt_magic(ERR4("Upt0"), 0, O_FROM_FILE, image);
    j40__inner *inner = new_inner();
    if (!inner) return set_alt_magic(ERR4("!mem"), 0, O_FROM_FILE, image);
    int saved = errno;
    errno = 0;
    FILE *fp = fopen(path, "rb");
    if (!fp) {
        int e = errno;
        errno = saved;
        free_inner(inner);
        return set_alt_magic(ERR4("open"), e, O_FROM_FILE, image);
    }
    errno = saved;
    // the reference streams the file lazily (j40.h:1238-1262); a read error surfaces in j40_next_frame there,
    // here the file is small enough to be read at once and a failure is reported the same way later
    size_t cap = 1 << 16, len = 0;
    uint8_t *data = (uint8_t *) malloc(cap);
    int read_failed = 0;
    while (data) {
        size_t got = fread(data + len, 1, cap - len, fp);
        len += got;
        if (got == 0) { read_failed = ferror(fp); break; }
        if (len == cap) {
            uint8_t *nd = (uint8_t *) realloc(data, cap * 2);
            if (!nd) { free(data); data = nullptr; break; }
            data = nd;
            cap *= 2;
        }
    }
    fclose(fp);
    if (!data) { free_inner(inner); return set_alt_magic(ERR4("!mem"), 0, O_FROM_FILE, image); }
    inner->owned = data;
    inner->data = data;
    inner->size = len;
    if (read_failed) { inner->err = ERR4("read"); inner->origin = O_NEXT_FRAME; }
    image->magic = MAGIC_IMAGE;
    image->u.inner = inner;
    return 0;
}

EXPORT j40_err j40_output_format(j40_image *image, int32_t channel, int32_t format) {
    j40__inner *inner;
    j40_err err = check_image(image, O_OUTPUT_FORMAT, &inner);
    if (err) return err;
    if (channel != J40_RGBA) { inner->origin = O_OUTPUT_FORMAT; return inner->err = ERR4("Uch?"); }
    if (format != J40_U8X4) { inner->origin = O_OUTPUT_FORMAT; return inner->err = ERR4("Ufm?"); }
    return 0;
}

// The single-image API shares one lazily created batch object (stream, events, device block, pinned staging) per
// process: creating and destroying those per image costs tens of milliseconds. Handles stay independent; decodes
// of different handles are serialised by the mutex (the reference has no threading at all, j40.h:8034).
static std::mutex g_api_mutex;
static j40b_batch *g_api_ctx = nullptr;
static int g_api_dev = -1;

static j40_err advance(j40__inner *inner) {
    if (inner->advanced) return 0;
    // header-level errors do not need a GPU to be diagnosed; the plan is handed to the batch as parsed
    std::unique_ptr<FramePlan> plan(new FramePlan);
    if (uint32_t e = parse_frame(inner->data, inner->size, *plan)) return e;
    std::lock_guard<std::mutex> lock(g_api_mutex);
    int dev = 0;
    if (const char *e = getenv("J40B_DEVICE")) dev = atoi(e);
    if (g_api_ctx && g_api_dev != dev) { j40b_batch_destroy(g_api_ctx); g_api_ctx = nullptr; }
    if (!g_api_ctx) { g_api_ctx = j40b_batch_create(dev); g_api_dev = dev; }
    j40b_batch *b = g_api_ctx;
    if (!b) return ERR4("!gpu");
    j40b_batch_reset(b);
    b->batch->plans.push_back(std::move(plan));
    b->inputs.push_back({inner->data, inner->size});
    const int idx = 0;
    j40_err err = 0;
    if (j40b_batch_upload(b) != 0) err = j40b_batch_error(b, idx) ? j40b_batch_error(b, idx) : ERR4("!gpu");
    if (!err) {
        j40b_batch_decode(b);
        j40b_batch_wait(b);
        err = j40b_batch_error(b, idx);
    }
    if (!err) {
        j40b_batch_info(b, idx, &inner->width, &inner->height, &inner->stride);
        size_t total = (size_t) inner->stride * (size_t) inner->height;
        void *p = nullptr;
        // pinned memory pays for big frames only (cudaHostAlloc itself costs about a millisecond)
        if (total >= (4u << 20) && cuda_ok(cudaHostAlloc(&p, total, cudaHostAllocDefault))) inner->pixels_pinned = true;
        else { cudaGetLastError(); p = malloc(total ? total : 1); }
        if (!p) err = ERR4("!mem");
        else {
            inner->pixels = (uint8_t *) p;
            if (j40b_batch_read_pixels(b, idx, p) != 0) err = ERR4("!gpu");
        }
    }
    j40b_batch_reset(b); // drops the borrowed input pointers; allocations stay for the next image
    if (!err) inner->advanced = 1;
    return err;
}

EXPORT int j40_next_frame(j40_image *image) {
    j40__inner *inner;
    j40_err err = check_image(image, O_NEXT_FRAME, &inner);
    if (err) return 0;
    err = advance(inner);
    if (err) {
        inner->origin = O_NEXT_FRAME;
        inner->err = err;
        return 0;
    }
    if (inner->rendered) return 0;
    inner->rendered = 1;
    return 1;
}

EXPORT j40_frame j40_current_frame(j40_image *image) {
    j40__inner *inner;
    j40_frame frame;
    j40_err err = check_image(image, O_CURRENT_FRAME, &inner);
    frame.magic = MAGIC_FRAME_ERR;
    frame.reserved = 0;
    frame.inner = inner;
    if (err) return frame;
    if (!inner->rendered) {
        if (!j40_next_frame(image)) {
            if (inner->err) return frame;
        }
    }
    frame.magic = MAGIC_FRAME;
    return frame;
}

// the 21x7 placeholder shown on errors: red, alpha spelling "Err" (same picture as j40.h:8429-8446)
static const char *const PLACEHOLDER[7] = {
    "111111111111111111111",
    "100011111111111111111",
    "101111111111111111111",
    "100010001000100010001",
    "101110111011101010111",
    "100010111011100010111",
    "111111111111111111111",
};

EXPORT j40_pixels_u8x4 j40_frame_pixels_u8x4(const j40_frame *frame, int32_t channel) {
    static uint8_t error_data[7 * 21 * 4];
    static bool error_init = false;
    if (!error_init) {
        for (int y = 0; y < 7; ++y) for (int x = 0; x < 21; ++x) {
            uint8_t *p = error_data + (y * 21 + x) * 4;
            p[0] = 255; p[1] = 0; p[2] = 0; p[3] = PLACEHOLDER[y][x] == '1' ? 255 : 0;
        }
        error_init = true;
    }
    j40_pixels_u8x4 error_pixels = {21, 7, 21 * 4, error_data};
    if (!frame || frame->magic != MAGIC_FRAME) return error_pixels;
    j40__inner *inner = frame->inner;
    if (!inner || inner->magic != MAGIC_INNER) return error_pixels;
    if (channel != J40_RGBA) return error_pixels;
    if (!inner->rendered) { inner->origin = O_FRAME_PIXELS; inner->err = ERR4("Urnd"); return error_pixels; }
    j40_pixels_u8x4 px;
    px.width = inner->width;
    px.height = inner->height;
    px.stride_bytes = inner->stride;
    px.data = inner->pixels;
    return px;
}

EXPORT const j40_u8x4 *j40_row_u8x4(j40_pixels_u8x4 pixels, int32_t y) {
    return (const j40_u8x4 *) ((const char *) pixels.data + (size_t) pixels.stride_bytes * (size_t) y);
}

EXPORT void j40_free(j40_image *image) {
    j40__inner *inner;
    if (!image) return;
    check_image(image, O_FREE, &inner);
    if (inner) free_inner(inner);
    image->magic = MAGIC_IMAGE_ERR ^ (uint32_t) O_NEXT;
    image->u.err = ERR4("Ufre");
}

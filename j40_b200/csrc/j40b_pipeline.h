// j40-b200: batch pipeline. Lays a batch of parsed frames out in device memory (one allocation, one
// H2D copy), builds the per-thread-block work items and issues the kernels:
//   VarDCT frame:  lf_group (1 block / LF group) -> hf_group (1 warp / group, serial entropy decode)
//                  -> back (1 block / group: dequant, CfL, IDCT, XYB->sRGB, RGBA8 store)
//   modular frame: modular (1 warp / group) -> render (1 thread / pixel)
// Templated over a backend: CudaBackend (j40b_cuda.cu) is the product; tests/hostemu provides a CPU
// backend that runs the very same kernel bodies single-threaded for the `-m "not gpu"` logic tests.
#pragma once
#include "j40b_exec.h"
#include "j40b_host.h"
#include <memory>
#include <string.h>
#include <stdlib.h>
#include <algorithm>
#include <atomic>
#include <thread>

namespace j40b {

struct ImageResult {
    uint32_t err = 0;
    int32_t width = 0, height = 0, stride = 0;
    size_t rgba_off = 0; // offset of the RGBA8 plane inside the batch's device block
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

template <class BE>
class Batch {
public:
    explicit Batch(BE &be_) : be(be_) {}
    ~Batch() { release(); }

    BE &be;
    std::vector<std::unique_ptr<FramePlan>> plans;
    std::vector<ImageResult> results;
    bool full_token_cap = false;

    // ---- step 1: host parse (no device work)
    void add(const uint8_t *data, size_t size) {
        std::unique_ptr<FramePlan> p(new FramePlan);
        parse_frame(data, size, *p);
        plans.push_back(std::move(p));
    }

    // host parse of many images on `threads` host threads (frames are independent; parse_frame touches only its
    // own FramePlan and the immutable global tables)
    void add_many(const uint8_t *const *bufs, const size_t *sizes, size_t n, int threads) {
        const size_t first = plans.size();
        for (size_t i = 0; i < n; ++i) plans.emplace_back(new FramePlan);
        parallel_for(n, threads, [&](size_t i) { parse_frame(bufs[i], sizes[i], *plans[first + i]); });
    }

    // runs fn(0..n-1) on up to `threads` host threads (<= 0: as many as the machine offers, at most 16)
    template <class F>
    static void parallel_for(size_t n, int threads, F fn) {
        if (threads <= 0) threads = (int) std::min<unsigned>(16u, std::max(1u, std::thread::hardware_concurrency()));
        threads = (int) std::min<size_t>((size_t) threads, n);
        if (threads <= 1) { for (size_t i = 0; i < n; ++i) fn(i); return; }
        std::atomic<size_t> next(0);
        std::vector<std::thread> pool;
        auto work = [&] { for (size_t i; (i = next.fetch_add(1)) < n;) fn(i); };
        for (int t = 1; t < threads; ++t) pool.emplace_back(work);
        work();
        for (auto &t : pool) t.join();
    }
    int host_threads = 0; // for upload()'s copies into the staging buffer

    // forgets the images but keeps the device / pinned allocations for the next upload()
    void reset() { plans.clear(); results.clear(); img.clear(); num_lf = num_hf = num_grp = 0; }

    // ---- step 2: lay out + upload (asynchronous on the backend's stream). Returns false when out of memory.
    bool upload() {
        const GlobalTables &gt = GlobalTables::get();
        size_t n = plans.size();
        results.assign(n, ImageResult());
        img.assign(n, Img());
        // global tables block (shared by all images)
        size_t up = 0;                      // upload blob cursor
        auto up_alloc = [&](size_t bytes) { size_t o = align_up(up, 256); up = o + bytes; return o; };
        size_t gt_dq[17], gt_order[13], gt_thr;
        for (int i = 0; i < 17; ++i) gt_dq[i] = up_alloc(gt.dq[i].size() * 4);
        for (int i = 0; i < 13; ++i) gt_order[i] = up_alloc(gt.order[i].size() * 4);
        gt_thr = up_alloc(256 * 4);
        size_t gt_lut = up_alloc(SRGB_LUT_BYTES);
        size_t work = 0;                    // device-only area cursor (follows the upload blob)
        auto wk_alloc = [&](size_t bytes) { size_t o = align_up(work, 256); work = o + bytes; return o; };
        // everything execute() has to zero before a decode (error words, token ranges) sits in one region, cleared by
        // a single memset per decode instead of one per LF group (hundreds of tiny stream operations per batch)
        size_t zero = 0;
        auto zk_alloc = [&](size_t bytes) { size_t o = align_up(zero, 256); zero = o + bytes; return o; };
        size_t n_lf = 0, n_hf = 0, n_grp = 0, n_mod = 0;
        max_global_blob = max_coeff_blob = max_ec_blob = 0;
        lf_tree_lanes = 1;
        for (size_t k = 0; k < n; ++k) {
            FramePlan &p = *plans[k];
            Img &im = img[k];
            results[k].err = p.err;
            if (p.err) continue;
            const DFrame &d = p.df;
            results[k].width = d.width; results[k].height = d.height;
            results[k].stride = (int32_t) align_up((size_t) d.width * 4 + 1, 32);
            im.frame_off = up_alloc(sizeof(DFrame));
            im.arena_off = up_alloc(p.arena.bytes.size());
            im.cs_off = up_alloc(p.cs_size + 16);
            im.rgba_off = wk_alloc((size_t) results[k].stride * (size_t) d.height);
            results[k].rgba_off = im.rgba_off; // relative to the work area for now
            if (!d.is_modular) {
                if (d.global_spec_off) {
                    const DCodeSpec *gs = (const DCodeSpec *) (p.arena.bytes.data() + d.global_spec_off);
                    max_global_blob = std::max(max_global_blob, (size_t) (gs->blob_hi - gs->blob_lo));
                }
                bool lz_coef = false;
                for (int ps = 0; ps < d.num_passes; ++ps) {
                    const DCodeSpec *cs_ = (const DCodeSpec *) (p.arena.bytes.data() + d.coeff_spec_off[ps]);
                    max_coeff_blob = std::max(max_coeff_blob, (size_t) (cs_->blob_hi - cs_->blob_lo));
                    lz_coef = lz_coef || cs_->lz77_enabled;
                }
                const size_t npass = (size_t) d.num_passes;
                im.nlf = p.lfg_sec.size(); im.npg = p.pg_sec.size(); im.ng = im.npg / npass;
                im.lfg_off = up_alloc(sizeof(DLfGroup) * im.nlf);
                im.grp_off = up_alloc(sizeof(DGroup) * im.npg);
                im.err_off = zk_alloc(4 * (im.nlf + im.npg + 1));
                im.lf.resize(im.nlf);
                bool wp = d.global_tree_uses_wp != 0;
                for (const FramePlan::LocalHeader &lh : p.lfg_local) wp = wp || (lh.present && lh.uses_wp);
                bool lz_mod = d.global_spec_off && ((const DCodeSpec *) (p.arena.bytes.data() + d.global_spec_off))->lz77_enabled;
                for (const FramePlan::LocalHeader &lh : p.lfg_local) // (trees local to LF-group sections bring their own code specs)
                    if (lh.present && !lh.host_err && lh.spec_off && ((const DCodeSpec *) (p.arena.bytes.data() + lh.spec_off))->lz77_enabled) lz_mod = true;
                for (size_t i = 0; i < im.nlf; ++i) {
                    LfBuf &b = im.lf[i];
                    int ggx = (int) (i % (size_t) d.ggcolumns), ggy = (int) (i / (size_t) d.ggcolumns);
                    b.left = ggx * 2048; b.top = ggy * 2048;
                    b.w = std::min(2048, d.width - b.left); b.h = std::min(2048, d.height - b.top);
                    b.w8 = ceil_div(b.w, 8); b.h8 = ceil_div(b.h, 8); b.w64 = ceil_div(b.w, 64); b.h64 = ceil_div(b.h, 64);
                    size_t n8 = (size_t) b.w8 * b.h8, n64 = (size_t) b.w64 * b.h64;
                    size_t cap = (size_t) 1 << ceil_lg32((uint32_t) n8);
                    b.lfq = wk_alloc(2 * 3 * n8);
                    b.lfdeq = wk_alloc(4 * 3 * n8);
                    b.lf = d.skip_adapt_lf_smooth ? b.lfdeq : wk_alloc(4 * 3 * n8);
                    b.lfidx = wk_alloc(n8);
                    b.xfromy = wk_alloc(2 * n64); b.bfromy = wk_alloc(2 * n64);
                    b.blockinfo = wk_alloc(2 * 2 * cap);
                    b.sharp = wk_alloc(2 * n8);
                    b.blocks = wk_alloc(4 * n8);
                    b.varblocks = wk_alloc(sizeof(DVarblock) * n8);
                    b.llf = wk_alloc(4 * 3 * n8);
                    b.wp = wp ? wk_alloc(4 * 2 * 5 * std::max<size_t>(cap, (size_t) b.w8)) : (size_t) -1;
                    b.lz = lz_mod ? wk_alloc(4u << 18) : (size_t) -1;
                    b.vb_tok = zk_alloc(4 * 2 * 3 * n8 * npass);
                    b.llf_scratch = wk_alloc(4 * 2048);
                    b.lane = wk_alloc(sizeof(ModLaneScratch));
                }
                im.grp.resize(im.npg);
                size_t tok_total = 0;
                for (size_t pg = 0; pg < im.npg; ++pg) {
                    GrpBuf &gb = im.grp[pg];
                    const size_t g = pg % im.ng;
                    int grow = (int) (g / (size_t) d.gcolumns), gcol = (int) (g % (size_t) d.gcolumns);
                    gb.gw = std::min(d.width, (gcol + 1) * 256) - gcol * 256;
                    gb.gh = std::min(d.height, (grow + 1) * 256) - grow * 256;
                    // words of the token list: one per non-zero coefficient, three for a value outside 16 bits
                    size_t full = (size_t) 3 * 3 * 64 * ceil_div(gb.gw, 8) * ceil_div(gb.gh, 8);
                    size_t cap = full_token_cap ? full : std::min(full, (size_t) p.pg_sec[pg].size * 3 + 256);
                    if (!full_token_cap && token_squeeze) cap = std::min<size_t>(cap, 64); // tests: force the overflow + retry path
                    gb.tok_first = tok_total; gb.tok_cap = cap;
                    tok_total += cap;
                    gb.lz = lz_coef ? wk_alloc(4u << 18) : (size_t) -1;
                    gb.nonzeros = wk_alloc(3 * 32);
                    gb.vbs = pg < im.ng ? wk_alloc(sizeof(HfVb) * (size_t) ceil_div(gb.gw, 8) * ceil_div(gb.gh, 8)) : im.grp[g].vbs;
                }
                im.tok_off = wk_alloc(sizeof(DToken) * std::max<size_t>(tok_total, 1));
                // extra channels (alpha ...) of a multi-section frame: coded per pass group behind the coefficients; the
                // reference decodes them and then drops the planes (j40.h:7024-7033, 7869-7870). Decoded here into scratch
                // planes for the sake of the errors that decoding can raise.
                im.nec = p.gmod.num_channels > 0 && !p.single_section ? im.npg : 0;
                if (im.nec) {
                    im.ec.resize(im.nec);
                    const int nch = p.gmod.num_channels - p.num_gm_channels;
                    const bool lz_g = d.global_spec_off && ((const DCodeSpec *) (p.arena.bytes.data() + d.global_spec_off))->lz77_enabled;
                    int max_w = 1;
                    for (size_t pg = 0; pg < im.nec; ++pg) {
                        ModBuf &mb = im.ec[pg];
                        mb.gw = im.grp[pg].gw; mb.gh = im.grp[pg].gh;
                        max_w = std::max(max_w, mb.gw);
                        mb.wp = wp ? wk_alloc(4 * 2 * 5 * (size_t) mb.gw) : (size_t) -1;
                        uint32_t capl = 1;
                        while (capl < (size_t) nch * mb.gw * mb.gh && capl < (1u << 20)) capl <<= 1;
                        mb.lz_mask = capl - 1;
                        mb.lz = lz_g ? wk_alloc(4 * (size_t) capl) : (size_t) -1;
                        mb.lane = wk_alloc(sizeof(ModLaneScratch));
                        mb.planes = wk_alloc(2 * (size_t) std::max(1, nch) * mb.gw * mb.gh);
                    }
                    im.ring_w = (size_t) ((max_w + 63) & ~63);
                    const size_t LS = (size_t) be.lane_stride(), slots = (im.nec + LS - 1) / LS;
                    im.ring_off = wk_alloc(2 * slots * 3 * im.ring_w * LS);
                    im.wring_off = wk_alloc(4 * slots * 2 * im.ring_w * 5 * LS);
                    im.mod_off = up_alloc(sizeof(ModWork) * im.nec);
                    if (d.global_spec_off) {
                        const DCodeSpec *gs = (const DCodeSpec *) (p.arena.bytes.data() + d.global_spec_off);
                        max_ec_blob = std::max(max_ec_blob, (size_t) (gs->blob_hi - gs->blob_lo));
                    }
                }
                n_lf += im.nlf; n_hf += im.npg; n_grp += im.ng;
            } else {
                im.err_off = zk_alloc(4 * (p.pg_sec.size() + 2));
                // coded channel list of the global image: palette (meta) channels first, each with its own size
                int nch = d.num_channels;
                for (int c = 0; c < nch; ++c) im.plane[c] = wk_alloc(2 * (size_t) std::max(1, p.gmod.ch[c].w) * std::max(1, p.gmod.ch[c].h));
                const bool has_global = p.num_gm_channels > 0; // work item 0 = the channels coded in LfGlobal
                im.nmod = p.pg_sec.size() + (has_global ? 1 : 0);
                auto spec_lz = [&](uint32_t off) { return off && ((const DCodeSpec *) (p.arena.bytes.data() + off))->lz77_enabled; };
                const bool wp_g = d.global_tree_uses_wp != 0, lz_g = spec_lz(d.global_spec_off);
                int gsize = 1 << d.group_size_shift;
                im.mod.resize(im.nmod);
                for (size_t g = 0; g < im.nmod; ++g) {
                    ModBuf &mb = im.mod[g];
                    const bool global = has_global && g == 0;
                    const size_t pg = g - (has_global ? 1 : 0);
                    int gw, gh;
                    size_t syms = 0;
                    if (global) {
                        gw = gh = 1;
                        for (int c = 0; c < p.num_gm_channels; ++c) {
                            gw = std::max(gw, p.gmod.ch[c].w); gh = std::max(gh, p.gmod.ch[c].h);
                            syms += (size_t) p.gmod.ch[c].w * p.gmod.ch[c].h;
                        }
                    } else {
                        gw = std::min(d.width, ((int) (pg % (size_t) d.gcolumns) + 1) * gsize) - (int) (pg % (size_t) d.gcolumns) * gsize;
                        gh = std::min(d.height, ((int) (pg / (size_t) d.gcolumns) + 1) * gsize) - (int) (pg / (size_t) d.gcolumns) * gsize;
                        syms = (size_t) (nch - p.num_gm_channels) * gw * gh;
                    }
                    mb.gw = gw; mb.gh = gh;
                    const FramePlan::LocalHeader *lh = global ? &p.gmod_local : (pg < p.pg_local.size() ? &p.pg_local[pg] : nullptr);
                    const bool local = lh && lh->present;
                    const bool wp = local ? lh->uses_wp != 0 : wp_g, lz = local ? spec_lz(lh->spec_off) : lz_g;
                    mb.wp = wp ? wk_alloc(4 * 2 * 5 * (size_t) gw) : (size_t) -1;
                    uint32_t capl = 1;
                    while (capl < syms && capl < (1u << 20)) capl <<= 1;
                    mb.lz_mask = capl - 1;
                    mb.lz = lz ? wk_alloc(4 * (size_t) capl) : (size_t) -1;
                    mb.lane = wk_alloc(sizeof(ModLaneScratch));
                }
                {   // row ring of the lane-per-stream decoder: as wide as the widest group, at most 1024
                    int max_w = 1;
                    for (const ModBuf &mb : im.mod) max_w = std::max(max_w, mb.gw);
                    im.ring_w = max_w <= 1024 ? (size_t) ((max_w + 63) & ~63) : 0;
                    const size_t LS = (size_t) be.lane_stride(), slots = (im.nmod + LS - 1) / LS;
                    im.ring_off = im.ring_w ? wk_alloc(2 * slots * 3 * im.ring_w * LS) : 0;
                    im.wring_off = im.ring_w ? wk_alloc(4 * slots * 2 * im.ring_w * 5 * LS) : 0;
                }
                im.mod_off = up_alloc(sizeof(ModWork) * std::max<size_t>(im.nmod, 1));
                im.render_off = up_alloc(sizeof(RenderWork));
                im.ndelta = 0;
                if (p.gmod.nb_transforms > 0 && p.gmod.tr[p.gmod.nb_transforms - 1].kind == 1 && p.gmod.tr[p.gmod.nb_transforms - 1].nb_deltas > 0) {
                    const ModTransform &t = p.gmod.tr[p.gmod.nb_transforms - 1];
                    im.ndelta = t.num_c;
                    for (int c = 0; c < t.num_c && c < MOD_MAX_CH; ++c) {
                        im.dplane[c] = wk_alloc(2 * (size_t) d.width * d.height);
                        im.dwp[c] = t.d_pred == 6 ? wk_alloc(4 * 2 * 5 * (size_t) d.width) : (size_t) -1;
                    }
                }
                n_mod += im.nmod;
            }
        }
        zero_off = wk_alloc(zero);
        zero_bytes = zero;
        for (size_t k = 0; k < n; ++k) {
            if (plans[k]->err) continue;
            img[k].err_off += zero_off;
            for (LfBuf &b : img[k].lf) b.vb_tok += zero_off;
        }
        {
            const size_t LS = (size_t) be.lane_stride(), slots = (n_lf + LS - 1) / LS;
            lf_ring_off = wk_alloc(2 * slots * 3 * LF_RING_W * LS);
            lf_wring_off = wk_alloc(4 * slots * 2 * LF_RING_W * 5 * LS);
        }
        lfw_off = up_alloc(sizeof(LfWork) * std::max<size_t>(n_lf, 1));
        // The coefficient kernel's work list is padded so that no block of `hf_per_block` lanes holds sections of two
        // images or passes: the lanes of a block then all decode with the block's staged copy of the code spec (a lane
        // of another image reads its tables through L1 and holds up its whole warp; with 135 sections per 4K frame and
        // 128 lanes per block nearly every block had such lanes). Padding items have grp == nullptr.
        hf_per_block = n_hf ? std::max(1, be.hf_block_lanes((int) n_hf)) : 1;
        {
            size_t padded = 0;
            for (size_t k = 0; k < n; ++k) {
                const Img &im = img[k];
                if (plans[k]->err || plans[k]->df.is_modular || !im.ng) continue;
                padded += (im.npg / im.ng) * align_up(im.ng, (size_t) hf_per_block);
            }
            n_hf = padded;
        }
        hfw_off = up_alloc(sizeof(HfWork) * std::max<size_t>(n_hf, 1));
        bkw_off = up_alloc(sizeof(BackWork) * std::max<size_t>(n_grp, 1));
        ppw_off = up_alloc(sizeof(HfPrepWork) * std::max<size_t>(n_grp, 1));
        num_lf = n_lf; num_hf = n_hf; num_grp = n_grp;
        upload_bytes = align_up(up, 256);
        work_bytes = align_up(work, 256);
        if (upload_bytes + work_bytes > dev_cap) {
            if (dev) be.dev_free(dev);
            dev_cap = upload_bytes + work_bytes;
            dev = (uint8_t *) be.dev_alloc(dev_cap);
        }
        if (upload_bytes > staging_cap) {
            if (staging) be.host_free(staging);
            staging_cap = upload_bytes + upload_bytes / 8;
            staging = (uint8_t *) be.host_alloc(staging_cap);
        }
        if (!dev || !staging) {
            release();
            for (size_t k = 0; k < n; ++k) if (!results[k].err) results[k].err = E_MEM;
            for (size_t k = 0; k < n; ++k) if (!plans[k]->err) plans[k]->err = E_MEM;
            num_lf = num_hf = num_grp = 0;
            return false;
        }
        // zero everything but the codestream / arena payloads (those are overwritten below, in parallel)
        {
            size_t pos = 0;
            for (size_t k = 0; k < n; ++k) {
                if (plans[k]->err) continue;
                const Img &im = img[k];
                memset(staging + pos, 0, im.arena_off - pos);
                pos = im.cs_off + plans[k]->cs_size; // arena and codestream are adjacent allocations
                memset(staging + im.arena_off + plans[k]->arena.bytes.size(), 0, im.cs_off - (im.arena_off + plans[k]->arena.bytes.size()));
            }
            memset(staging + pos, 0, upload_bytes - pos);
        }
        parallel_for(n, host_threads, [&](size_t k) {
            const FramePlan &p = *plans[k];
            if (p.err) return;
            memcpy(staging + img[k].arena_off, p.arena.bytes.data(), p.arena.bytes.size());
            memcpy(staging + img[k].cs_off, p.cs, p.cs_size);
        });
        uint8_t *dwork = dev + upload_bytes;
        // ---- fill the staging blob with device addresses
        for (int i = 0; i < 17; ++i) memcpy(staging + gt_dq[i], gt.dq[i].data(), gt.dq[i].size() * 4);
        for (int i = 0; i < 13; ++i) memcpy(staging + gt_order[i], gt.order[i].data(), gt.order[i].size() * 4);
        memcpy(staging + gt_thr, gt.srgb_thr, 256 * 4);
        memcpy(staging + gt_lut, gt.srgb_lut, SRGB_LUT_BYTES);
        LfWork *lfw = (LfWork *) (staging + lfw_off);
        HfWork *hfw = (HfWork *) (staging + hfw_off);
        BackWork *bkw = (BackWork *) (staging + bkw_off);
        HfPrepWork *ppw = (HfPrepWork *) (staging + ppw_off);
        size_t ilf = 0, ihf = 0, igrp = 0;
        std::vector<std::pair<int64_t, LfWork>> lf_sorted; // (cells, item)
        for (size_t k = 0; k < n; ++k) {
            FramePlan &p = *plans[k];
            Img &im = img[k];
            if (p.err) continue;
            std::vector<HfWork> hf_img; // this image's sections; sorted and padded into hfw below
            results[k].rgba_off = upload_bytes + im.rgba_off;
            DFrame d = p.df;
            // tables: per-image custom ones live in the arena, library defaults in the shared block
            for (int i = 0; i < 17; ++i) {
                d.dq[i] = (const float *) (p.custom_dq_off[i] ? dev + im.arena_off + p.custom_dq_off[i] : dev + gt_dq[i]);
            }
            for (int ps = 0; ps < MAX_PASSES; ++ps) for (int i = 0; i < 13; ++i) for (int c = 0; c < 3; ++c) {
                d.order[ps][i][c] = (const int32_t *) (p.custom_order_off[ps][i][c] ? dev + im.arena_off + p.custom_order_off[ps][i][c] : dev + gt_order[i]);
            }
            d.srgb_thr = (const float *) (dev + gt_thr);
            d.srgb_lut = dev + gt_lut;
            d.srgb_wrap_hi = gt.srgb_wrap_hi;
            memcpy(staging + im.frame_off, &d, sizeof(d));
            const DFrame *dframe = (const DFrame *) (dev + im.frame_off);
            const uint8_t *darena = dev + im.arena_off, *dcs = dev + im.cs_off;
            uint32_t *derr = (uint32_t *) (dwork + im.err_off);
            if (!d.is_modular) {
                DLfGroup *lg = (DLfGroup *) (staging + im.lfg_off);
                DGroup *gr = (DGroup *) (staging + im.grp_off);
                DToken *dtok = (DToken *) (dwork + im.tok_off);
                for (size_t i = 0; i < im.nlf; ++i) {
                    const LfBuf &b = im.lf[i];
                    DLfGroup &g = lg[i];
                    g.idx = (int32_t) i; g.left = b.left; g.top = b.top; g.width = b.w; g.height = b.h;
                    g.width8 = b.w8; g.height8 = b.h8; g.width64 = b.w64; g.height64 = b.h64;
                    g.sec_off = (uint32_t) p.lfg_sec[i].off; g.sec_size = p.lfg_sec[i].size; g.sec_start_bit = p.lfg_sec[i].start_bit;
                    g.lfq = (int16_t *) (dwork + b.lfq); g.lfdeq = (float *) (dwork + b.lfdeq); g.lf = (float *) (dwork + b.lf);
                    g.lfidx = dwork + b.lfidx; g.xfromy = (int16_t *) (dwork + b.xfromy); g.bfromy = (int16_t *) (dwork + b.bfromy);
                    g.blockinfo = (int16_t *) (dwork + b.blockinfo); g.sharpness = (int16_t *) (dwork + b.sharp);
                    g.blocks = (int32_t *) (dwork + b.blocks); g.varblocks = (DVarblock *) (dwork + b.varblocks);
                    g.llf = (float *) (dwork + b.llf);
                    g.wp_scratch = b.wp == (size_t) -1 ? nullptr : (int32_t *) (dwork + b.wp);
                    g.lz_window = b.lz == (size_t) -1 ? nullptr : (int32_t *) (dwork + b.lz);
                    g.vb_tok = (uint32_t *) (dwork + b.vb_tok);
                    g.nb_varblocks = 0; g.end_bit = 0;
                    LfWork &w = lfw[ilf++];
                    w.f = dframe; w.arena = darena; w.cs = dcs;
                    w.g = (DLfGroup *) (dev + im.lfg_off) + i;
                    // lanes the compiled trees of this LF group's channels need at most (the pruning and the count that
                    // modular_channel_prep / simt_compile_tree do on the device)
                    for (int st = 0; st < 2; ++st) {
                        const bool local = !p.lfg_local.empty() && p.lfg_local[2 * i + (size_t) st].present;
                        const uint32_t toff = local ? p.lfg_local[2 * i + (size_t) st].tree_off : d.global_tree_off;
                        if (!local && !d.have_global_tree) continue;
                        const DTreeNode *tree = (const DTreeNode *) (p.arena.bytes.data() + toff);
                        const int32_t sidx = st == 0 ? 1 + (int32_t) i : 1 + 2 * d.num_lf_groups + (int32_t) i;
                        for (int c = 0; c < (st == 0 ? 3 : 4); ++c) {
                            DTreeNode pruned[PTREE_CAP];
                            bool uses_wp = false;
                            const int nn = prune_tree(tree, c, sidx, pruned, PTREE_CAP, &uses_wp);
                            int inner = 0;
                            for (int k = 0; k < nn; ++k) inner += pruned[k].a < 0;
                            lf_tree_lanes = std::max(lf_tree_lanes, nn == 0 ? 32 : std::max(inner, nn - inner));
                        }
                    }
                    w.err = derr + i;
                    w.llf_scratch = (float *) (dwork + b.llf_scratch);
                    w.lane_scratch = (ModLaneScratch *) (dwork + b.lane);
                    for (int st = 0; st < 2; ++st) {
                        LfLocal &lo = w.local[st];
                        memset(&lo, 0, sizeof(lo));
                        if (p.lfg_local.empty() || !p.lfg_local[2 * i + (size_t) st].present) continue;
                        const FramePlan::LocalHeader &lh = p.lfg_local[2 * i + (size_t) st];
                        lo.present = 1; lo.host_err = lh.host_err; lo.tree_off = lh.tree_off; lo.spec_off = lh.spec_off;
                        lo.uses_wp = lh.uses_wp; lo.start_bit = lh.start_bit; lo.hdr = lh.hdr;
                    }
                    lf_sorted.emplace_back((int64_t) b.w8 * b.h8, w);
                }
                for (size_t pg = 0; pg < im.npg; ++pg) {
                    const GrpBuf &gb = im.grp[pg];
                    DGroup &g = gr[pg];
                    const size_t gi = pg % im.ng;
                    int grow = (int) (gi / (size_t) d.gcolumns), gcol = (int) (gi % (size_t) d.gcolumns);
                    g.idx = (int32_t) gi;
                    g.pass = (int32_t) (pg / im.ng);
                    g.lfg = (grow / 8) * d.ggcolumns + gcol / 8;
                    g.gx8 = (gcol % 8) * 32; g.gy8 = (grow % 8) * 32;
                    g.gw = gb.gw; g.gh = gb.gh;
                    g.sec_off = (uint32_t) p.pg_sec[pg].off; g.sec_size = p.pg_sec[pg].size; g.sec_start_bit = p.pg_sec[pg].start_bit;
                    g.tok_first = (uint32_t) gb.tok_first; g.tok_cap = (uint32_t) gb.tok_cap;
                    g.lz_window = gb.lz == (size_t) -1 ? nullptr : (int32_t *) (dwork + gb.lz);
                    g.nonzeros = dwork + gb.nonzeros;
                    g.tok_used = 0;
                    g.vbs = (HfVb *) (dwork + gb.vbs);
                    g.nvb = 0;
                    hf_img.emplace_back();
                    HfWork &w = hf_img.back();
                    w.f = dframe; w.arena = darena; w.cs = dcs;
                    w.g = (DLfGroup *) (dev + im.lfg_off) + g.lfg;
                    w.grp = (DGroup *) (dev + im.grp_off) + pg;
                    w.geo = (DGroup *) (dev + im.grp_off) + gi;
                    w.tokens = dtok;
                    w.lf_err = derr + g.lfg;
                    w.err = derr + im.nlf + pg;
                    if (pg >= im.ng) continue;
                    BackWork &bw = bkw[igrp];
                    bw.f = dframe; bw.arena = darena; bw.g = w.g; bw.grp = w.grp; bw.tokens = dtok;
                    bw.lf_err = w.lf_err; bw.hf_err = w.err; bw.hf_err_stride = (int32_t) im.ng;
                    bw.rgba = dwork + im.rgba_off; bw.rgba_stride = results[k].stride;
                    bw.big_scratch = nullptr;
                    HfPrepWork &pw = ppw[igrp++];
                    pw.f = dframe; pw.arena = darena; pw.g = w.g; pw.grp = w.grp; pw.lf_err = w.lf_err;
                }
                // pass groups of one image, longest section first: the lanes of a warp (one group each) then
                // carry similar amounts of work, and the long ones start first
                // (pass by pass: the lanes of a block share one staged code spec)
                std::stable_sort(hf_img.begin(), hf_img.end(), [&](const HfWork &a, const HfWork &b2) {
                    const DGroup &ga = gr[a.grp - (DGroup *) (dev + im.grp_off)], &gb2 = gr[b2.grp - (DGroup *) (dev + im.grp_off)];
                    return ga.pass != gb2.pass ? ga.pass < gb2.pass : ga.sec_size > gb2.sec_size;
                });
                for (size_t s0 = 0; s0 < hf_img.size(); s0 += im.ng) { // one pass's sections, then padding up to a whole block
                    for (size_t i = 0; i < im.ng; ++i) hfw[ihf++] = hf_img[s0 + i];
                    for (size_t i = im.ng; i < align_up(im.ng, (size_t) hf_per_block); ++i) memset(&hfw[ihf++], 0, sizeof(HfWork));
                }
                if (im.nec) {
                    ModWork *mw = (ModWork *) (staging + im.mod_off);
                    const int nch = p.gmod.num_channels - p.num_gm_channels;
                    for (size_t pg = 0; pg < im.nec; ++pg) {
                        ModWork &w = mw[pg];
                        const ModBuf &mb = im.ec[pg];
                        memset(&w, 0, sizeof(w));
                        w.f = dframe; w.arena = darena; w.cs = dcs;
                        w.sec_off = (uint32_t) p.pg_sec[pg].off; w.sec_size = p.pg_sec[pg].size;
                        w.start_bit_src = &((DGroup *) (dev + im.grp_off) + pg)->end_bit;
                        w.skip_if = derr + im.nlf + pg;
                        w.err = derr + im.nlf + pg;
                        // stream index of a pass group (j40.h:7012): behind LfGlobal, three per LF group, 17 matrices
                        w.sidx = (int32_t) (1 + 3 * d.num_lf_groups + 17 + (int64_t) pg);
                        w.tree_off = d.global_tree_off; w.spec_off = d.global_spec_off; w.tree_uses_wp = d.global_tree_uses_wp;
                        w.m.num_channels = nch;
                        for (int c = 0; c < nch; ++c) {
                            w.m.ch[c].px = (int16_t *) (dwork + mb.planes) + (size_t) c * mb.gw * mb.gh;
                            w.m.ch[c].stride = mb.gw; w.m.ch[c].w = mb.gw; w.m.ch[c].h = mb.gh;
                        }
                        w.wp_scratch = mb.wp == (size_t) -1 ? nullptr : (int32_t *) (dwork + mb.wp);
                        w.lz_window = mb.lz == (size_t) -1 ? nullptr : (int32_t *) (dwork + mb.lz);
                        w.lz_mask = mb.lz_mask;
                        w.lane_scratch = (ModLaneScratch *) (dwork + mb.lane);
                        const size_t LS = (size_t) be.lane_stride(), slot = pg / LS, lane = pg % LS;
                        w.ring_w = (int32_t) im.ring_w; w.ring_lstride = (int32_t) LS;
                        w.ring = (int16_t *) (dwork + im.ring_off) + slot * 3 * im.ring_w * LS + lane;
                        w.wring = (int32_t *) (dwork + im.wring_off) + slot * 2 * im.ring_w * 5 * LS + lane;
                    }
                }
            } else {
                ModWork *mw = (ModWork *) (staging + im.mod_off);
                RenderWork *rw = (RenderWork *) (staging + im.render_off);
                memset(rw, 0, sizeof(*rw));
                rw->f = dframe;
                for (int c = 0; c < d.num_channels; ++c) rw->plane[c] = (int16_t *) (dwork + im.plane[c]);
                for (int c = 0; c < im.ndelta && c < MOD_MAX_CH; ++c) {
                    rw->dplane[c] = (int16_t *) (dwork + im.dplane[c]);
                    rw->dwp[c] = im.dwp[c] == (size_t) -1 ? nullptr : (int32_t *) (dwork + im.dwp[c]);
                }
                rw->any_err = derr + im.nmod; // summary word written by nobody: per-section words are checked on the host
                rw->rgba = dwork + im.rgba_off; rw->rgba_stride = results[k].stride;
                int gsize = 1 << d.group_size_shift;
                const bool has_global = p.num_gm_channels > 0;
                for (size_t g = 0; g < im.nmod; ++g) {
                    ModWork &w = mw[g];
                    const ModBuf &mb = im.mod[g];
                    memset(&w, 0, sizeof(w));
                    w.f = dframe; w.arena = darena; w.cs = dcs;
                    const bool global = has_global && g == 0;
                    const size_t pg = g - (has_global ? 1 : 0);
                    const SectionRef &s = global ? p.gmod_sec : p.pg_sec[pg];
                    w.sec_off = (uint32_t) s.off; w.sec_size = s.size; w.sec_start_bit = s.start_bit;
                    w.sidx = global ? 0 : (int32_t) (1 + 3 * d.num_lf_groups + 17 + (int64_t) pg);
                    w.header_parsed = global ? 1 : 0;
                    w.is_global = global ? 1 : 0;
                    w.check_end = global && p.single_section ? 1 : 0;
                    w.tree_off = d.global_tree_off; w.spec_off = d.global_spec_off; w.tree_uses_wp = d.global_tree_uses_wp;
                    const FramePlan::LocalHeader *lh = global ? &p.gmod_local : (pg < p.pg_local.size() ? &p.pg_local[pg] : nullptr);
                    if (lh && lh->present) {
                        w.preset_err = lh->host_err;
                        w.tree_off = lh->tree_off; w.spec_off = lh->spec_off; w.tree_uses_wp = lh->uses_wp;
                        if (!global) { w.header_parsed = 1; w.sec_start_bit = lh->start_bit; w.m = lh->hdr; }
                    }
                    if (global) {
                        // the channels coded with the global image: all of them (single group) or the palettes
                        w.m = p.gmod;
                        w.m.num_channels = p.num_gm_channels;
                        for (int c = 0; c < p.num_gm_channels; ++c) {
                            w.m.ch[c].px = (int16_t *) (dwork + im.plane[c]);
                            w.m.ch[c].stride = p.gmod.ch[c].w;
                        }
                    } else {
                        const int gx = (int) (pg % (size_t) d.gcolumns) * gsize, gy = (int) (pg / (size_t) d.gcolumns) * gsize;
                        w.m.num_channels = d.num_channels - p.num_gm_channels;
                        for (int c = 0; c < w.m.num_channels; ++c) {
                            w.m.ch[c].px = (int16_t *) (dwork + im.plane[p.num_gm_channels + c]) + (size_t) gy * d.width + gx;
                            w.m.ch[c].stride = d.width; w.m.ch[c].w = mb.gw; w.m.ch[c].h = mb.gh;
                            w.m.ch[c].hshift = w.m.ch[c].vshift = 0;
                        }
                    }
                    w.wp_scratch = mb.wp == (size_t) -1 ? nullptr : (int32_t *) (dwork + mb.wp);
                    w.lz_window = mb.lz == (size_t) -1 ? nullptr : (int32_t *) (dwork + mb.lz);
                    w.lz_mask = mb.lz_mask;
                    w.err = derr + g;
                    w.lane_scratch = (ModLaneScratch *) (dwork + mb.lane);
                    {
                        const size_t LS = (size_t) be.lane_stride(), slot = g / LS, lane = g % LS;
                        w.ring_w = (int32_t) im.ring_w; w.ring_lstride = (int32_t) LS;
                        w.ring = im.ring_w ? (int16_t *) (dwork + im.ring_off) + slot * 3 * im.ring_w * LS + lane : nullptr;
                        w.wring = im.ring_w ? (int32_t *) (dwork + im.wring_off) + slot * 2 * im.ring_w * 5 * LS + lane : nullptr;
                    }
                }
            }
        }
        // LF groups, largest first: a 4K frame has two big and two tiny LF groups; in image order the
        // one-warp blocks of the serial decoders land two big ones per SM on half of the SMs
        std::stable_sort(lf_sorted.begin(), lf_sorted.end(), [](const std::pair<int64_t, LfWork> &a, const std::pair<int64_t, LfWork> &b2) { return a.first > b2.first; });
        for (size_t i = 0; i < lf_sorted.size(); ++i) {
            lfw[i] = lf_sorted[i].second;
            // the lane-per-stream decoders' row ring: one per warp (LS consecutive items), lane-interleaved
            const size_t LS = (size_t) be.lane_stride(), slot = i / LS, lane = i % LS;
            lfw[i].ring_w = LF_RING_W; lfw[i].ring_lstride = (int32_t) LS;
            lfw[i].ring = (int16_t *) (dwork + lf_ring_off) + slot * 3 * LF_RING_W * LS + lane;
            lfw[i].wring = (int32_t *) (dwork + lf_wring_off) + slot * 2 * LF_RING_W * 5 * LS + lane;
        }
        be.h2d(dev, staging, upload_bytes);
        return true;
    }

    // ---- step 3: kernels (can be repeated: all state they depend on is re-initialised here)
    void execute() {
        if (!dev) return;
        uint8_t *dwork = dev + upload_bytes;
        if (zero_bytes) be.dev_memset(dwork + zero_off, 0, zero_bytes);
        // multi-section frames only: the pass groups of a single-section frame start where its LF group ends
        bool split = true;
        for (size_t k = 0; k < plans.size(); ++k) if (!plans[k]->err && !plans[k]->df.is_modular && plans[k]->single_section) split = false;
        if (num_lf) be.launch_lf((const LfWork *) (dev + lfw_off), (int) num_lf, max_global_blob, split, lf_tree_lanes);
        if (num_hf) be.launch_hf((const HfPrepWork *) (dev + ppw_off), (int) num_grp, (const HfWork *) (dev + hfw_off), (int) num_hf, max_coeff_blob, hf_per_block);
        for (size_t k = 0; k < plans.size(); ++k) { // extra channels behind the coefficients (errors only; planes are scratch)
            const Img &im = img[k];
            if (plans[k]->err || !im.nec) continue;
            int max_w = 0;
            for (const ModBuf &mb : im.ec) max_w = std::max(max_w, mb.gw);
            be.launch_mod((ModWork *) (dev + im.mod_off), (int) im.nec, max_ec_blob ? max_ec_blob : (size_t) -1, max_w);
        }
        if (num_grp) be.launch_back((const BackWork *) (dev + bkw_off), (int) num_grp);
        bool any_mod = false;
        for (size_t k = 0; k < plans.size(); ++k) {
            FramePlan &p = *plans[k];
            Img &im = img[k];
            if (p.err || !p.df.is_modular) continue;
            if (!any_mod) { any_mod = true; be.mark_modular(0); }
            if (im.nmod) {
                int max_w = 0;
                for (const ModBuf &mb : im.mod) max_w = std::max(max_w, (int) mb.gw);
                const DCodeSpec *gs = p.df.global_spec_off ? (const DCodeSpec *) (p.arena.bytes.data() + p.df.global_spec_off) : nullptr;
                if (p.gmod_local.present || !p.pg_local.empty()) gs = nullptr; // local code specs: no common blob to stage
                be.launch_mod((ModWork *) (dev + im.mod_off), (int) im.nmod, gs ? (size_t) (gs->blob_hi - gs->blob_lo) : (size_t) -1, max_w);
            }
            if (im.ndelta) be.launch_palette_delta((const RenderWork *) (dev + im.render_off), im.ndelta);
            be.launch_render((const RenderWork *) (dev + im.render_off), p.df.width, p.df.height);
        }
        if (any_mod) be.mark_modular(1);
        be.join_side(); // (the LF groups' sharpness channels, if they ran on a side stream)
    }

    // ---- step 4: errors (synchronises)
    void collect_errors() {
        be.sync();
        if (!dev) return;
        uint8_t *dwork = dev + upload_bytes;
        for (size_t k = 0; k < plans.size(); ++k) {
            FramePlan &p = *plans[k];
            Img &im = img[k];
            if (p.err) continue;
            size_t nerr = p.df.is_modular ? im.nmod + 2 : im.nlf + im.npg + 1;
            std::vector<uint32_t> e(nerr);
            be.d2h(e.data(), dwork + im.err_off, 4 * nerr);
            uint32_t best = 0;
            int64_t best_rank = INT64_MAX;
            if (!p.df.is_modular) {
                for (size_t i = 0; i < im.nlf; ++i) if (e[i] && p.lfg_sec[i].rank < best_rank) { best = e[i]; best_rank = p.lfg_sec[i].rank; }
                for (size_t g = 0; g < im.npg; ++g) if (e[im.nlf + g] && p.pg_sec[g].rank < best_rank) { best = e[im.nlf + g]; best_rank = p.pg_sec[g].rank; }
            } else {
                const bool has_global = p.num_gm_channels > 0;
                for (size_t g = 0; g < im.nmod; ++g) {
                    int64_t rank = has_global && g == 0 ? -2 : p.pg_sec[g - (has_global ? 1 : 0)].rank;
                    if (e[g] && rank < best_rank) { best = e[g]; best_rank = rank; }
                }
            }
            if (!best) {
                // Nothing may follow the frame (j40.h:8213). The reference only notices trailing bytes that
                // its main read-ahead buffer (64 KiB, j40.h:1676) already holds when it seeks past a
                // multi-section frame (j40.h:1789-1797); single-section frames are checked exactly.
                if (p.single_section) {
                    if (p.cs_size > p.end_codeoff) best = E_EXCS;
                    else if (p.cs_size < p.end_codeoff) best = E_SHRT;
                    else if (p.trailing_box_err) best = p.trailing_box_err;
                    else if (p.trailing_partial_box) best = E_SHRT; // read through to the end of the box, then the next header
                } else if (p.cs_size > p.end_codeoff && p.end_codeoff < 65536) best = E_EXCS;
            }
            results[k].err = best;
        }
    }

    // E_LTRE round: LF groups whose sub-bitstream names a tree of its own reported where its header starts; the host
    // reads header, tree and code spec there. Returns true if anything was read (upload + execute again).
    bool resolve_local_trees() {
        if (!dev) return false;
        bool any = false;
        uint8_t *dwork = dev + upload_bytes;
        for (size_t k = 0; k < plans.size(); ++k) {
            FramePlan &p = *plans[k];
            const Img &im = img[k];
            if (p.err || p.df.is_modular || results[k].err != E_LTRE) continue;
            std::vector<uint32_t> e(im.nlf);
            be.d2h(e.data(), dwork + im.err_off, 4 * im.nlf);
            for (size_t i = 0; i < im.nlf; ++i) {
                if (e[i] != E_LTRE) continue;
                DLfGroup g;
                be.d2h(&g, dev + im.lfg_off + i * sizeof(DLfGroup), sizeof(g));
                if (g.ltree_stage < 0 || g.ltree_stage > 1) continue;
                if (!p.lfg_local.empty() && p.lfg_local[2 * i + (size_t) g.ltree_stage].present) continue; // (cannot happen twice)
                parse_lf_group_local_tree(p, i, g.ltree_stage, g.ltree_bit, g.nb_varblocks);
                any = true;
            }
        }
        return any;
    }

    void download_pixels(size_t k, uint8_t *dst) {
        const ImageResult &r = results[k];
        be.d2h(dst, dev + r.rgba_off, (size_t) r.stride * (size_t) r.height);
    }
    uint8_t *device_pixels(size_t k) { return dev + results[k].rgba_off; }
    // enqueues the D2H copy of every image (pitch bytes apart in dst) behind the kernels; truly asynchronous
    // when dst is pinned host memory
    void download_all_async(uint8_t *dst, size_t pitch) {
        if (!dev) return;
        for (size_t k = 0; k < results.size(); ++k) {
            const ImageResult &r = results[k];
            if (plans[k]->err || !r.height) continue;
            size_t bytes = (size_t) r.stride * (size_t) r.height;
            be.d2h_async(dst + k * pitch, dev + r.rgba_off, bytes < pitch ? bytes : pitch);
        }
    }

    // Diagnostics: an intermediate array of LF group `lfg` of (VarDCT) image k after a decode, in the layout of the
    // reference's j40__lf_group_st members (the numbering follows oracle/ref_harness.c's ref_staged_lf_group_array):
    //   0 blocks i32[h8][w8]; 1 varblocks {coeffoff | qfidx, bits of 1/HfMul} i32[nb][2]; 2 lfindices u8[h8][w8];
    //   3/4/5 llfcoeffs f32[w8*h8] X/Y/B; 6/7/8 coeffs f32[w8*h8*64] as decoded; 9/10 xfromy/bfromy i16[h64][w64];
    //   11 sharpness i16[h8][w8]; 12/13/14 coeffs after dequantisation; 15/16/17 the smoothed LF planes f32[h8][w8].
    // Returns the number of bytes written, 0 if unavailable or `cap` is too small.
    size_t debug_dump(size_t k, size_t lfg, int what, void *dst, size_t cap) {
        if (!dev || k >= plans.size() || plans[k]->err || results[k].err || plans[k]->df.is_modular) return 0;
        const Img &im = img[k];
        if (lfg >= im.nlf) return 0;
        const LfBuf &b = im.lf[lfg];
        uint8_t *dwork = dev + upload_bytes;
        const size_t n8 = (size_t) b.w8 * b.h8, n64 = (size_t) b.w64 * b.h64;
        auto copy = [&](size_t off, size_t bytes) -> size_t { if (bytes > cap) return 0; be.d2h(dst, dwork + off, bytes); return bytes; };
        switch (what) {
        case 0: return copy(b.blocks, n8 * 4);
        case 1: {
            DLfGroup g;
            be.d2h(&g, dev + im.lfg_off + lfg * sizeof(DLfGroup), sizeof(g));
            std::vector<DVarblock> vb((size_t) std::max(0, g.nb_varblocks));
            if (vb.size() * 8 > cap) return 0;
            if (!vb.empty()) be.d2h(vb.data(), dwork + b.varblocks, vb.size() * sizeof(DVarblock));
            int32_t *o = (int32_t *) dst;
            for (size_t i = 0; i < vb.size(); ++i) { o[2 * i] = vb[i].coeffoff | vb[i].qfidx; memcpy(&o[2 * i + 1], &vb[i].hfmul_inv, 4); }
            return vb.size() * 8;
        }
        case 2: return copy(b.lfidx, n8);
        case 3: case 4: case 5: return copy(b.llf + (size_t) (what - 3) * n8 * 4, n8 * 4);
        case 9: return copy(b.xfromy, n64 * 2);
        case 10: return copy(b.bfromy, n64 * 2);
        case 11: return copy(b.sharp, n8 * 2);
        case 15: case 16: case 17: return copy(b.lf + (size_t) (what - 15) * n8 * 4, n8 * 4);
        case 6: case 7: case 8: case 12: case 13: case 14: {
            if (n8 * 64 * 4 > cap) return 0;
            float *planes = (float *) be.dev_alloc(3 * n8 * 64 * 4);
            if (!planes) return 0;
            be.dev_memset(planes, 0, 3 * n8 * 64 * 4);
            DumpWork w;
            w.f = (const DFrame *) (dev + im.frame_off);
            w.g = (const DLfGroup *) (dev + im.lfg_off) + lfg;
            w.tokens = (const DToken *) (dwork + im.tok_off);
            w.out = planes;
            w.dequant = what >= 12 ? 1 : 0;
            be.launch_dump(w, (int) n8);
            const int c = what >= 12 ? what - 12 : what - 6;
            be.d2h(dst, planes + (size_t) c * n8 * 64, n8 * 64 * 4);
            be.dev_free(planes);
            return n8 * 64 * 4;
        }
        default: return 0;
        }
    }

    void release() {
        if (dev) be.dev_free(dev);
        if (staging) be.host_free(staging);
        dev = staging = nullptr;
        dev_cap = staging_cap = 0;
    }

    size_t device_bytes() const { return upload_bytes + work_bytes; }
    size_t h2d_bytes() const { return upload_bytes; }

private:
    struct LfBuf { int left, top, w, h, w8, h8, w64, h64; size_t lfq, lfdeq, lf, lfidx, xfromy, bfromy, blockinfo, sharp, blocks, varblocks, llf, wp, lz, vb_tok, llf_scratch, lane; };
    struct GrpBuf { int gw, gh; size_t tok_first, tok_cap, lz, nonzeros, vbs; };
    struct ModBuf { int gw, gh; size_t wp, lz, lane, planes; uint32_t lz_mask; };
    struct Img {
        size_t frame_off = 0, arena_off = 0, cs_off = 0, lfg_off = 0, grp_off = 0, mod_off = 0, render_off = 0;
        size_t rgba_off = 0, err_off = 0, tok_off = 0, plane[MOD_MAX_CH] = {0};
        size_t dplane[MOD_MAX_CH] = {0}, dwp[MOD_MAX_CH] = {0}; // delta palette: restored channels, WP rows
        int ndelta = 0;
        size_t ring_w = 0, ring_off = 0, wring_off = 0; // modular frames: the lane decoders' row ring
        size_t nlf = 0, ng = 0, npg = 0, nmod = 0, nec = 0; // LF groups, groups, (pass, group) sections, modular sub-bitstreams
        std::vector<LfBuf> lf;
        std::vector<GrpBuf> grp;
        std::vector<ModBuf> mod, ec; // modular frames: sub-bitstreams; VarDCT frames: extra channels per pass group
    };
    std::vector<Img> img;
    uint8_t *dev = nullptr, *staging = nullptr;
    size_t upload_bytes = 0, work_bytes = 0, dev_cap = 0, staging_cap = 0;
    size_t lf_ring_off = 0, lf_wring_off = 0, zero_off = 0, zero_bytes = 0;
    enum { LF_RING_W = 256 }; // LF groups are at most 256 cells wide
    size_t lfw_off = 0, hfw_off = 0, bkw_off = 0, ppw_off = 0, num_lf = 0, num_hf = 0, num_grp = 0;
    bool token_squeeze = getenv("J40B_TEST_TOKEN_SQUEEZE") != nullptr; // see prepare(): exercises the token-arena retry
    size_t max_global_blob = 0, max_coeff_blob = 0, max_ec_blob = 0; // largest code-spec blobs of the batch (shared-memory staging sizes)
    int hf_per_block = 1;  // lanes per block of the coefficient kernel (the work list is padded to it per image and pass)
    int lf_tree_lanes = 1; // most inner nodes / leaves of any LF-group channel's pruned tree in the batch
};

} // namespace j40b

// JPEG XL test-stream writer -- TEST INFRASTRUCTURE (C API for tests/ and bench.py).
#include "vardct.h"
#include "modframe.h"
#include <cstdlib>
#include <sstream>

using namespace jxlgen;

static GenTables g_tables;
static thread_local std::string g_err;

static std::map<std::string, double> parse_kv(const char *s) {
    std::map<std::string, double> m;
    std::stringstream ss(s ? s : "");
    std::string item;
    while (std::getline(ss, item, ',')) {
        size_t eq = item.find('=');
        if (eq == std::string::npos) continue;
        m[item.substr(0, eq)] = std::atof(item.substr(eq + 1).c_str());
    }
    return m;
}
#define GETI(field, key) if (kv.count(key)) p.field = (decltype(p.field)) kv[key]

extern "C" {

__attribute__((visibility("default"))) const char *jxlgen_last_error() { return g_err.c_str(); }

// kind 0: dq weights (idx 0..16, float[n*3]); 1: natural order (idx 0..12, int32[n]);
// 2: forward 64x64 matrix for a special 8x8 transform (idx = dctsel, double[4096])
__attribute__((visibility("default"))) int jxlgen_set_table(int kind, int idx, const void *data, int count) {
    if (kind == 0 && idx >= 0 && idx < 17) g_tables.dq[idx].assign((const float *) data, (const float *) data + count);
    else if (kind == 1 && idx >= 0 && idx < 13) g_tables.order[idx].assign((const int32_t *) data, (const int32_t *) data + count);
    else if (kind == 2 && idx >= 0 && idx < 27) g_tables.fwd[idx].assign((const double *) data, (const double *) data + count);
    else return 0;
    return 1;
}

static void fill_stats(const GenStats &s, int64_t *out) {
    if (!out) return;
    out[0] = s.bytes; out[1] = s.hf_symbols; out[2] = s.lf_symbols; out[3] = s.nonzeros; out[4] = s.num_varblocks;
    out[5] = s.sections; out[6] = s.coef_clusters; out[7] = s.tree_nodes;
    for (int i = 0; i < 27; ++i) out[8 + i] = s.transform_hist[i];
}

// rgb (optional): caller-supplied 8-bit RGB image, else a synthetic photo is generated from the seed
__attribute__((visibility("default"))) int jxlgen_vardct(const char *params, const uint8_t *rgb, uint8_t **out, size_t *out_size, int64_t *stats /*40*/) {
    try {
        auto kv = parse_kv(params);
        VarDCTParams p;
        GETI(width, "width"); GETI(height, "height"); GETI(seed, "seed"); GETI(transform_mix, "mix");
        GETI(global_scale, "global_scale"); GETI(quant_lf, "quant_lf"); GETI(hfmul_base, "hfmul"); GETI(hfmul_var, "hfmul_var");
        GETI(x_qm_scale, "x_qm"); GETI(b_qm_scale, "b_qm"); GETI(use_ans, "ans"); GETI(log_alpha_size, "las");
        GETI(max_clusters, "clusters"); GETI(cfl, "cfl"); GETI(custom_cfl_base, "cfl_base"); GETI(smooth_lf, "smooth");
        GETI(extra_prec, "extra_prec"); GETI(tree_preset, "tree"); GETI(custom_block_ctx, "block_ctx");
        GETI(custom_orders, "orders"); GETI(num_hf_presets, "presets"); GETI(explicit_frame_header, "explicit_fh");
        GETI(container, "container"); GETI(container_jxlp, "jxlp"); GETI(permuted_toc, "permuted"); GETI(lz77_coeffs, "lz77");
        GETI(quant_deadzone, "deadzone"); GETI(force_dctsel, "force"); GETI(alpha, "alpha"); GETI(raw_dq, "raw_dq");
        GETI(passes, "passes"); GETI(coef_spike, "coef_spike"); GETI(raw_dq_lie, "raw_dq_lie"); GETI(lf_local_tree, "lf_local_tree"); GETI(big_take, "big_take"); GETI(big_thr, "big_thr");
        ImageRGB8 im;
        if (rgb) { im.w = p.width; im.h = p.height; im.px.assign(rgb, rgb + (size_t) p.width * (size_t) p.height * 3); }
        else im = synth_photo(p.width, p.height, p.seed);
        VarDCTEncoder enc(p, g_tables);
        std::vector<uint8_t> code = enc.encode(im);
        *out = (uint8_t *) std::malloc(code.size() ? code.size() : 1);
        std::memcpy(*out, code.data(), code.size());
        *out_size = code.size();
        fill_stats(enc.stats, stats);
        return 1;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 0;
    }
}

__attribute__((visibility("default"))) int jxlgen_modular(const char *params, const uint8_t *rgb, uint8_t **out, size_t *out_size, int64_t *stats) {
    try {
        auto kv = parse_kv(params);
        ModularParams p;
        GETI(width, "width"); GETI(height, "height"); GETI(seed, "seed"); GETI(group_shift, "group_shift");
        GETI(rct_type, "rct"); GETI(tree_preset, "tree"); GETI(use_ans, "ans"); GETI(lz77, "lz77"); GETI(alpha, "alpha");
        GETI(container, "container"); GETI(smooth, "smooth"); GETI(max_clusters, "clusters"); GETI(local_tree, "local_tree"); GETI(palette, "palette"); GETI(pal_deltas, "pal_deltas"); GETI(pal_pred, "pal_pred");
        ImageRGB8 im;
        if (rgb) { im.w = p.width; im.h = p.height; im.px.assign(rgb, rgb + (size_t) p.width * (size_t) p.height * 3); }
        else im = synth_photo(p.width, p.height, p.seed);
        ModularFrameEncoder enc(p);
        std::vector<uint8_t> code = enc.encode(im);
        *out = (uint8_t *) std::malloc(code.size() ? code.size() : 1);
        std::memcpy(*out, code.data(), code.size());
        *out_size = code.size();
        fill_stats(enc.stats, stats);
        return 1;
    } catch (const std::exception &e) {
        g_err = e.what();
        return 0;
    }
}

// the synthetic source image itself (for PSNR checks)
__attribute__((visibility("default"))) int jxlgen_synth(int w, int h, uint64_t seed, uint8_t *rgb_out) {
    ImageRGB8 im = synth_photo(w, h, seed);
    std::memcpy(rgb_out, im.px.data(), im.px.size());
    return 1;
}

__attribute__((visibility("default"))) void jxlgen_free(void *p) { std::free(p); }

}

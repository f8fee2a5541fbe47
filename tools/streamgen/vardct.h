// JPEG XL test-stream writer -- TEST INFRASTRUCTURE.
// VarDCT frame writer: forward XYB, varblock layout selection, forward transforms, quantisation,
// chroma-from-luma, LF image + HF metadata as modular sub-images, HF coefficient tokenisation with
// the exact context model the decoder uses (j40.h:6888-7004), section/TOC
// assembly.  Lossy on purpose: only the *syntax and context modelling* have to be exact.
#pragma once
#include "modular.h"
#include "synth.h"
#include <map>

namespace jxlgen {

struct DctSel { int8_t log_rows, log_cols, param_idx, order_idx; };
static const DctSel kDctSel[27] = {
    {3, 3, 0, 0}, {3, 3, 1, 1}, {3, 3, 2, 1}, {3, 3, 3, 1}, {4, 4, 4, 2}, {5, 5, 5, 3}, {4, 3, 6, 4}, {3, 4, 6, 4},
    {5, 3, 7, 5}, {3, 5, 7, 5}, {5, 4, 8, 6}, {4, 5, 8, 6}, {3, 3, 9, 1}, {3, 3, 9, 1}, {3, 3, 10, 1}, {3, 3, 10, 1},
    {3, 3, 10, 1}, {3, 3, 10, 1}, {6, 6, 11, 7}, {6, 5, 12, 8}, {5, 6, 12, 8}, {7, 7, 13, 9}, {7, 6, 14, 10},
    {6, 7, 14, 10}, {8, 8, 15, 11}, {8, 7, 16, 12}, {7, 8, 16, 12},
};
static const int8_t kOrderLog[13][2] = {{3, 3}, {3, 3}, {4, 4}, {5, 5}, {3, 4}, {3, 5}, {4, 5}, {6, 6}, {5, 6}, {7, 7}, {6, 7}, {8, 8}, {7, 8}};
inline bool is_special8(int dctsel) { return dctsel == 1 || dctsel == 2 || dctsel == 3 || (dctsel >= 12 && dctsel <= 17); }

// tables handed in by the harness (derived from the oracle so that the generator needs no decoder code)
struct GenTables {
    std::vector<float> dq[17];      // [n][3] dequantisation weights per parameter set
    std::vector<int32_t> order[13]; // natural coefficient orders
    std::vector<double> fwd[27];    // 64x64 forward matrices for the special 8x8 transforms
};

struct VarDCTParams {
    int width = 256, height = 256;
    uint64_t seed = 0;
    int transform_mix = 1;     // 0 = DCT8x8 only, 1 = "e6-like" mix up to 64x64, 2 = all 27 types
    int global_scale = 4096, quant_lf = 16;
    int hfmul_base = 3, hfmul_var = 2;
    int x_qm_scale = 3, b_qm_scale = 2;
    bool use_ans = true;       // false = prefix codes everywhere
    int log_alpha_size = 6;
    int max_clusters = 24;
    bool cfl = true;           // non-zero per-tile XFromY/BFromY
    bool custom_cfl_base = false;
    bool smooth_lf = true;     // false sets skip_adapt_lf_smooth
    int extra_prec = 0;
    int raw_dq = 0;            // bit i: dequantisation matrix i (only 0 = 8x8 and 4 = 16x16) is sent RAW, as a modular image
    bool alpha = false;        // 8-bit alpha extra channel, coded per pass group with the global tree (multi-group frames only)
    int tree_preset = 1;       // 0 single gradient leaf, 1 WP + property tree, 2 stress tree
    bool custom_block_ctx = false;
    int custom_orders = 0;     // bit mask over the 13 orders
    int num_hf_presets = 1;
    bool explicit_frame_header = true;
    bool container = false;    // wrap into ISO BMFF boxes (jxlc)
    bool container_jxlp = false; // split the codestream over two jxlp boxes
    bool permuted_toc = false;
    bool lz77_coeffs = false;  // enable LZ77 in the coefficient stream (rare in practice)
    float quant_deadzone = 0.55f;
    int force_dctsel = -1;     // >= 0: use this transform wherever it fits (coverage tests)
    int lf_local_tree = 0;     // bit 0: the LF image of every LF group, bit 1: its HF metadata image is coded with a tree of its
                               // own (a copy of the global one, stored with the sub-bitstream, j40.h:3827-3835); bit 2: odd LF groups only
    float raw_dq_lie = 1.0f;   // RAW matrices: the denominator written is this many times the one used for quantising, so that the
                               // decoder's coefficients come out that much larger (out-of-range samples; tests of the int16 wrap, j40.h:7234)
    int coef_spike = 0;        // != 0: quantised values outside 16 bits (the decoder's wide token form). One pass: a few varblocks get
                               // +-spike and the 16-bit boundary values at their last coefficients; several passes: every 7th
                               // position carries +-spike in pass 0 and its opposite in pass 1 (the sum stays small)
    int passes = 1;            // > 1: the quantised coefficients are split over this many passes (the decoder adds them up);
                               // pass p has its own code spec and, with custom_orders, its own coefficient orders
    float big_take = 0.75f;    // transform_mix 1: probability of taking a larger transform where the content allows it
    float big_thr = 0.06f;     // transform_mix 1: activity threshold (scaled by 1/sqrt(cells)) below which it is allowed
};

struct GenStats {
    int64_t bytes = 0, hf_symbols = 0, lf_symbols = 0, nonzeros = 0, num_varblocks = 0;
    int64_t transform_hist[27] = {0};
    int64_t sections = 0;
    int32_t coef_clusters = 0, tree_nodes = 0;
    double psnr_hint = 0;
};

// ------------------------------------------------------------------------------------------
// colour

inline double srgb_to_linear(double v) { return v <= 0.04045 ? v / 12.92 : std::pow((v + 0.055) / 1.055, 2.4); }

struct XYBImage {
    int w = 0, h = 0, pw = 0, ph = 0; // padded to multiples of 8
    std::vector<float> p[3];          // X, Y, B  [ph][pw]
};

inline XYBImage rgb_to_xyb(const ImageRGB8 &im) {
    static const double INV[3][3] = {
        {11.031566901960783, -9.866943921568629, -0.16462299647058826},
        {-3.254147380392157, 4.418770392156863, -0.16462299647058826},
        {-3.6588512862745097, 2.7129230470588235, 1.9459282392156863},
    };
    static const double BIAS = -0.0037930732552754493;
    // forward matrix = inverse of INV
    double a[3][6];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) { a[i][j] = INV[i][j]; a[i][j + 3] = i == j; }
    for (int i = 0; i < 3; ++i) {
        int piv = i;
        for (int r = i + 1; r < 3; ++r) if (std::fabs(a[r][i]) > std::fabs(a[piv][i])) piv = r;
        for (int j = 0; j < 6; ++j) std::swap(a[i][j], a[piv][j]);
        double d = a[i][i];
        for (int j = 0; j < 6; ++j) a[i][j] /= d;
        for (int r = 0; r < 3; ++r) if (r != i) { double f = a[r][i]; for (int j = 0; j < 6; ++j) a[r][j] -= f * a[i][j]; }
    }
    double lut[256];
    for (int i = 0; i < 256; ++i) lut[i] = srgb_to_linear(i / 255.0);
    XYBImage out;
    out.w = im.w; out.h = im.h;
    out.pw = (im.w + 7) / 8 * 8; out.ph = (im.h + 7) / 8 * 8;
    for (int c = 0; c < 3; ++c) out.p[c].assign((size_t) out.pw * (size_t) out.ph, 0.0f);
    double cb = std::cbrt(BIAS);
    for (int y = 0; y < out.ph; ++y) for (int x = 0; x < out.pw; ++x) {
        int sx = std::min(x, im.w - 1), sy = std::min(y, im.h - 1);
        const uint8_t *px = &im.px[((size_t) sy * (size_t) im.w + (size_t) sx) * 3];
        double r = lut[px[0]], g = lut[px[1]], b = lut[px[2]];
        double lms[3];
        for (int i = 0; i < 3; ++i) lms[i] = a[i][3] * r + a[i][4] * g + a[i][5] * b;
        double p[3];
        for (int i = 0; i < 3; ++i) p[i] = std::cbrt(lms[i] - BIAS) + cb;
        size_t o = (size_t) y * (size_t) out.pw + (size_t) x;
        out.p[0][o] = (float) ((p[0] - p[1]) * 0.5);
        out.p[1][o] = (float) ((p[0] + p[1]) * 0.5);
        out.p[2][o] = (float) p[2];
    }
    return out;
}

// ------------------------------------------------------------------------------------------
// forward DCT (orthogonal-cosine basis with DC = mean; the decoder's inverse has unit DC gain)

struct CosTable {
    int n;
    std::vector<double> c; // [k][i] = s_k / n * cos(pi k (2i+1) / 2n)
};
inline const CosTable &cos_table(int n) {
    static std::map<int, CosTable> cache;
    auto it = cache.find(n);
    if (it != cache.end()) return it->second;
    CosTable t;
    t.n = n;
    t.c.resize((size_t) n * (size_t) n);
    for (int k = 0; k < n; ++k) for (int i = 0; i < n; ++i) {
        double s = k ? std::sqrt(2.0) : 1.0;
        t.c[(size_t) k * (size_t) n + (size_t) i] = s / n * std::cos(M_PI * k * (2 * i + 1) / (2.0 * n));
    }
    return cache[n] = t;
}

// in: R x C samples (row-major, stride given); out: coefficients in the decoder's layout
inline void forward_dct2d(const float *in, int stride, int R, int C, std::vector<double> &out) {
    const CosTable &tr = cos_table(R), &tc = cos_table(C);
    std::vector<double> tmp((size_t) R * (size_t) C); // [y][u]
    for (int y = 0; y < R; ++y) for (int u = 0; u < C; ++u) {
        double s = 0;
        const double *cc = &tc.c[(size_t) u * (size_t) C];
        const float *row = in + (size_t) y * (size_t) stride;
        for (int x = 0; x < C; ++x) s += cc[x] * row[x];
        tmp[(size_t) y * (size_t) C + (size_t) u] = s;
    }
    out.assign((size_t) R * (size_t) C, 0.0);
    for (int v = 0; v < R; ++v) for (int u = 0; u < C; ++u) {
        double s = 0;
        const double *cr = &tr.c[(size_t) v * (size_t) R];
        for (int y = 0; y < R; ++y) s += cr[y] * tmp[(size_t) y * (size_t) C + (size_t) u];
        // wide blocks keep [v][u]; square and tall ones are stored transposed ([u][v])
        size_t idx = R < C ? (size_t) v * (size_t) C + (size_t) u : (size_t) u * (size_t) R + (size_t) v;
        out[idx] = s;
    }
}

// ------------------------------------------------------------------------------------------

struct VarBlock {
    int x8, y8;      // position in the LF group's 8x8 grid
    int dctsel;
    int hfmul;       // >= 1
    int qfidx;
    std::vector<int32_t> q[3]; // quantised coefficients, XYB, decoder layout
};

struct LfGroupEnc {
    int left, top, w, h, w8, h8, w64, h64;
    std::vector<VarBlock> vbs;
    std::vector<int32_t> blockmap; // [h8][w8]: index of the varblock whose top-left is here, else -1
    std::vector<int32_t> cover;    // [h8][w8]: index of covering varblock
    std::vector<int16_t> lfq[3];   // quantised LF, XYB order
    std::vector<uint8_t> lfidx;    // [h8][w8]
    std::vector<int16_t> xfromy, bfromy; // [h64][w64]
    std::vector<int16_t> sharp;    // [h8][w8]
};

class VarDCTEncoder {
public:
    VarDCTParams P;
    const GenTables &T;
    GenStats stats;
    Rng rng;

    VarDCTEncoder(const VarDCTParams &p, const GenTables &t) : P(p), T(t), rng(p.seed * 7919 + 17) {}

    // block-context configuration
    int nb_lf_thr[3] = {0, 0, 0}, nb_qf_thr = 0, lf_thr[3][15], qf_thr[15];
    std::vector<uint8_t> block_ctx_map;
    int nb_block_ctx = 15;
    // CfL
    int colour_factor = 84;
    float base_corr_x = 0.0f, base_corr_b = 1.0f;
    int x_factor_lf = 0, b_factor_lf = 0;

    int group_cols = 0, group_rows = 0, lfg_cols = 0, lfg_rows = 0;
    std::vector<LfGroupEnc> lfgs;
    std::vector<int> group_preset;

    std::vector<uint8_t> encode(const ImageRGB8 &im) {
        JG_CHECK(im.w == P.width && im.h == P.height);
        XYBImage xyb = rgb_to_xyb(im);
        setup_contexts();
        for (int i = 0; i < 17; ++i) if ((P.raw_dq >> i) & 1) { // weights as the decoder will see them: integers over 2
            raw_int[i].resize(T.dq[i].size());
            raw_dq[i].resize(T.dq[i].size());
            for (size_t k = 0; k < T.dq[i].size(); ++k) {
                int32_t v = (int32_t) std::lround((double) T.dq[i][k] * 2.0);
                v = std::min(32767, std::max(1, v));
                raw_int[i][k] = v;
                raw_dq[i][k] = (float) v / 2.0f;
            }
        }
        group_cols = (P.width + 255) / 256; group_rows = (P.height + 255) / 256;
        lfg_cols = (P.width + 2047) / 2048; lfg_rows = (P.height + 2047) / 2048;
        int num_groups = group_cols * group_rows, num_lfg = lfg_cols * lfg_rows;
        JG_CHECK(P.num_hf_presets >= 1 && P.num_hf_presets <= num_groups);
        group_preset.resize((size_t) num_groups);
        for (int g = 0; g < num_groups; ++g) group_preset[(size_t) g] = P.num_hf_presets > 1 ? rng.below(P.num_hf_presets) : 0;

        lfgs.resize((size_t) num_lfg);
        for (int gy = 0; gy < lfg_rows; ++gy) for (int gx = 0; gx < lfg_cols; ++gx) {
            LfGroupEnc &g = lfgs[(size_t) gy * (size_t) lfg_cols + (size_t) gx];
            g.left = gx * 2048; g.top = gy * 2048;
            g.w = std::min(2048, P.width - g.left); g.h = std::min(2048, P.height - g.top);
            g.w8 = (g.w + 7) / 8; g.h8 = (g.h + 7) / 8; g.w64 = (g.w + 63) / 64; g.h64 = (g.h + 63) / 64;
            build_lf_group(g, xyb);
        }

        // ---- global MA tree and the modular token streams that use it
        MATree tree = make_tree(num_lfg);
        stats.tree_nodes = (int32_t) tree.nodes.size();
        ModularTokenizer mt(tree);
        std::vector<TokStream> lf_ts((size_t) num_lfg), meta_ts((size_t) num_lfg);
        for (int i = 0; i < num_lfg; ++i) {
            LfGroupEnc &g = lfgs[(size_t) i];
            std::vector<Channel> ch(3);
            static const int YXB[3] = {1, 0, 2};
            for (int c = 0; c < 3; ++c) {
                ch[(size_t) c].w = g.w8; ch[(size_t) c].h = g.h8;
                ch[(size_t) c].px.assign(g.lfq[YXB[c]].begin(), g.lfq[YXB[c]].end());
            }
            mt.run(ch, 1 + i, lf_ts[(size_t) i]);
            std::vector<Channel> mc(4);
            mc[0].w = g.w64; mc[0].h = g.h64; mc[0].px.assign(g.xfromy.begin(), g.xfromy.end());
            mc[1].w = g.w64; mc[1].h = g.h64; mc[1].px.assign(g.bfromy.begin(), g.bfromy.end());
            mc[2].w = (int) g.vbs.size(); mc[2].h = 2;
            for (const VarBlock &vb : g.vbs) mc[2].px.push_back(vb.dctsel);
            for (const VarBlock &vb : g.vbs) mc[2].px.push_back(vb.hfmul - 1);
            mc[3].w = g.w8; mc[3].h = g.h8; mc[3].px.assign(g.sharp.begin(), g.sharp.end());
            mt.run(mc, 1 + 2 * num_lfg + i, meta_ts[(size_t) i]);
            stats.lf_symbols += (int64_t) lf_ts[(size_t) i].size() + (int64_t) meta_ts[(size_t) i].size();
        }
        // extra channel (alpha): one modular sub-stream per pass group, after the HF coefficients (j40.h:7024-7033)
        for (int i = 0; i < 17; ++i) if ((P.raw_dq >> i) & 1) {
            int rows = i == 0 ? 8 : 16, cols = rows;
            JG_CHECK(i == 0 || i == 4);
            std::vector<Channel> ch(3);
            for (int c = 0; c < 3; ++c) {
                ch[(size_t) c].w = cols; ch[(size_t) c].h = rows;
                for (int k = 0; k < rows * cols; ++k) ch[(size_t) c].px.push_back(raw_int[i][(size_t) k * 3 + (size_t) c]);
            }
            mt.run(ch, 1 + 3 * num_lfg + i, raw_ts[i]);
        }
        std::vector<TokStream> ec_ts;
        if (P.alpha && num_groups == 1) {
            // single group: the channel fits a group and is coded with the global image in LfGlobal (j40.h:6327-6337)
            std::vector<Channel> ch(1);
            ch[0].w = P.width; ch[0].h = P.height;
            for (int Y = 0; Y < P.height; ++Y) for (int X = 0; X < P.width; ++X) ch[0].px.push_back(((X / 37 + Y / 53) % 5 == 0) ? 200 + ((X >> 3) & 7) : 255);
            mt.run(ch, 0, ec_global_ts);
            stats.lf_symbols += (int64_t) ec_global_ts.size();
        } else if (P.alpha) {
            ec_ts.resize((size_t) num_groups);
            for (int g = 0; g < num_groups; ++g) {
                int grow = g / group_cols, gcol = g % group_cols;
                int gw = std::min(P.width, (gcol + 1) * 256) - gcol * 256, gh = std::min(P.height, (grow + 1) * 256) - grow * 256;
                std::vector<Channel> ch(1);
                ch[0].w = gw; ch[0].h = gh;
                for (int y = 0; y < gh; ++y) for (int x = 0; x < gw; ++x) {
                    int X = gcol * 256 + x, Y = grow * 256 + y;
                    ch[0].px.push_back(((X / 37 + Y / 53) % 5 == 0) ? 200 + ((X >> 3) & 7) : 255);
                }
                mt.run(ch, 1 + 3 * num_lfg + 17 + g, ec_ts[(size_t) g]);
                stats.lf_symbols += (int64_t) ec_ts[(size_t) g].size();
            }
        }
        EntropyOpts mo;
        mo.use_prefix = !P.use_ans;
        mo.log_alpha_size = 8;
        mo.cfg = {4, 1, 0};
        mo.max_clusters = 16;
        CodeSpec mspec;
        {
            std::vector<const TokStream *> all;
            for (auto &s : lf_ts) all.push_back(&s);
            for (auto &s : meta_ts) all.push_back(&s);
            for (auto &s : ec_ts) all.push_back(&s);
            all.push_back(&ec_global_ts);
            for (int i = 0; i < 17; ++i) all.push_back(&raw_ts[i]);
            mspec.build(tree.num_leaves, mo, all);
        }

        // ---- HF coefficient token streams
        JG_CHECK(P.passes >= 1 && P.passes <= 11 && (P.passes == 1 || num_groups > 1));
        const int npass = P.passes;
        std::vector<TokStream> hf_ts((size_t) num_groups * (size_t) npass); // [pass * num_groups + g]
        std::vector<std::vector<std::vector<int32_t>>> orders_p((size_t) npass);
        for (int ps = 0; ps < npass; ++ps) {
            orders_p[(size_t) ps] = make_orders(ps);
            for (int g = 0; g < num_groups; ++g) tokenize_group(g, orders_p[(size_t) ps], hf_ts[(size_t) ps * (size_t) num_groups + (size_t) g], ps);
        }
        EntropyOpts co;
        co.use_prefix = !P.use_ans;
        co.log_alpha_size = P.lz77_coeffs ? 8 : P.log_alpha_size;
        co.cfg = {4, 1, 0};
        co.max_clusters = P.max_clusters;
        co.lz77 = P.lz77_coeffs;
        if (co.lz77) for (auto &s : hf_ts) lz77_rle(s, co.min_length, 0, 6);
        int num_coef_ctx = 495 * nb_block_ctx * P.num_hf_presets;
        std::vector<CodeSpec> cspecs((size_t) npass);
        for (int ps = 0; ps < npass; ++ps) {
            std::vector<const TokStream *> all;
            for (int g = 0; g < num_groups; ++g) { const TokStream &s = hf_ts[(size_t) ps * (size_t) num_groups + (size_t) g]; all.push_back(&s); stats.hf_symbols += (int64_t) s.size(); }
            cspecs[(size_t) ps].build(num_coef_ctx, co, all);
        }
        stats.coef_clusters = cspecs[0].nclusters;

        // ---- sections
        BitWriter lfglobal;
        write_lf_global(lfglobal, tree, mspec, mo);
        BitWriter hfglobal;
        mspec_ptr = &mspec;
        write_hf_global(hfglobal, num_groups, cspecs, orders_p);
        std::vector<BitWriter> lfsec((size_t) num_lfg), pgsec((size_t) num_groups * (size_t) npass);
        for (int i = 0; i < num_lfg; ++i) {
            BitWriter &bw = lfsec[(size_t) i];
            bw.put((uint64_t) P.extra_prec, 2);
            EntropyOpts lto; // as in write_lf_global
            lto.use_prefix = !P.use_ans;
            lto.log_alpha_size = 8;
            lto.cfg = {4, 1, 0};
            lto.max_clusters = 6;
            const bool here = !(P.lf_local_tree & 4) || (i & 1);
            ModularHeaderOpts mh, mh1, mh2;
            mh1.use_global_tree = !(here && (P.lf_local_tree & 1));
            mh2.use_global_tree = !(here && (P.lf_local_tree & 2));
            write_modular_header_prefix(bw, mh1);
            if (!mh1.use_global_tree) { write_tree(bw, tree, lto); mspec.write(bw); }
            mspec.encode(bw, lf_ts[(size_t) i]);
            int nvb = (int) lfgs[(size_t) i].vbs.size();
            bw.put((uint64_t) (nvb - 1), ceil_lg((uint32_t) (lfgs[(size_t) i].w8 * lfgs[(size_t) i].h8)));
            write_modular_header_prefix(bw, mh2);
            if (!mh2.use_global_tree) { write_tree(bw, tree, lto); mspec.write(bw); }
            mspec.encode(bw, meta_ts[(size_t) i]);
        }
        for (int pg = 0; pg < num_groups * npass; ++pg) {
            const int g = pg % num_groups;
            BitWriter &bw = pgsec[(size_t) pg];
            bw.put((uint64_t) group_preset[(size_t) g], ceil_lg((uint32_t) P.num_hf_presets));
            cspecs[(size_t) (pg / num_groups)].encode(bw, hf_ts[(size_t) pg]);
            if (P.alpha && num_groups > 1) {
                ModularHeaderOpts mh;
                write_modular_header_prefix(bw, mh);
                mspec.encode(bw, ec_ts[(size_t) g]);
            }
        }

        // ---- assemble the codestream
        BitWriter out;
        write_headers(out);
        bool single = num_groups == 1 && npass == 1;
        if (single) {
            // reference order for single-section frames: LfGlobal, HfGlobal, LfGroup, PassGroup (SURVEY B-12)
            BitWriter body;
            body.append_bits(lfglobal);
            body.append_bits(hfglobal);
            body.append_bits(lfsec[0]);
            body.append_bits(pgsec[0]);
            body.pad();
            out.bit(0); // not permuted
            out.pad();
            out.u32((uint32_t) body.bytes.size(), 0, 10, 1024, 14, 17408, 22, 4211712, 30);
            out.pad();
            out.append_bytes(body.bytes);
            stats.sections = 1;
        } else {
            std::vector<BitWriter *> secs;
            secs.push_back(&lfglobal);
            for (auto &s : lfsec) secs.push_back(&s);
            secs.push_back(&hfglobal);
            for (auto &s : pgsec) secs.push_back(&s);
            for (BitWriter *s : secs) s->pad();
            stats.sections = (int64_t) secs.size();
            std::vector<int> fileorder(secs.size()); // fileorder[k] = logical section stored k-th
            for (size_t i = 0; i < secs.size(); ++i) fileorder[i] = (int) i;
            if (P.permuted_toc) {
                // shuffle pass-group sections among themselves (LF groups stay ahead of their groups)
                size_t first = 2 + (size_t) num_lfg;
                for (size_t i = secs.size() - 1; i > first; --i) {
                    size_t j = first + (size_t) rng.below((int) (i - first + 1));
                    std::swap(fileorder[i], fileorder[j]);
                }
                write_toc_permutation(out, fileorder);
            } else {
                out.bit(0);
            }
            out.pad();
            // TOC entries are the sizes in *file* order; the permutation maps them back
            for (size_t k = 0; k < secs.size(); ++k) {
                out.u32((uint32_t) secs[(size_t) fileorder[k]]->bytes.size(), 0, 10, 1024, 14, 17408, 22, 4211712, 30);
            }
            out.pad();
            for (size_t k = 0; k < secs.size(); ++k) out.append_bytes(secs[(size_t) fileorder[k]]->bytes);
        }
        std::vector<uint8_t> code = out.bytes;
        if (P.container) code = wrap_container(code, P.container_jxlp);
        stats.bytes = (int64_t) code.size();
        return code;
    }

private:
    // --------------------------------------------------------------------------------------
    void setup_contexts() {
        static const uint8_t DEF[39] = {
            0, 1, 2, 2, 3, 3, 4, 5, 6, 6, 6, 6, 6,
            7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14,
            7, 8, 9, 9, 10, 11, 12, 13, 14, 14, 14, 14, 14,
        };
        if (!P.custom_block_ctx) {
            block_ctx_map.assign(DEF, DEF + 39);
            nb_block_ctx = 15;
        } else {
            nb_lf_thr[0] = 1; lf_thr[0][0] = 0;          // X > 0
            nb_lf_thr[1] = 2; lf_thr[1][0] = 60; lf_thr[1][1] = 200; // Y
            nb_lf_thr[2] = 0;
            nb_qf_thr = 2; qf_thr[0] = 2; qf_thr[1] = 5; // compared against HfMul-1 (stored +1)
            int lfsize = (nb_lf_thr[0] + 1) * (nb_lf_thr[1] + 1) * (nb_lf_thr[2] + 1);
            int size = 39 * lfsize * (nb_qf_thr + 1);
            block_ctx_map.assign((size_t) size, 0);
            int maxc = 0;
            for (int c = 0; c < 3; ++c) for (int o = 0; o < 13; ++o) for (int q = 0; q <= nb_qf_thr; ++q) for (int l = 0; l < lfsize; ++l) {
                int idx = (o * (nb_qf_thr + 1) + q) * lfsize + l + 13 * (nb_qf_thr + 1) * lfsize * c;
                int v = (DEF[o + 13 * c] + q + (l & 1)) % 16;
                block_ctx_map[(size_t) idx] = (uint8_t) v;
                maxc = std::max(maxc, v);
            }
            // cluster ids must be contiguous from 0: compact them
            std::vector<int> remap(16, -1);
            int next = 0;
            for (uint8_t &v : block_ctx_map) { if (remap[v] < 0) remap[v] = next++; v = (uint8_t) remap[v]; }
            nb_block_ctx = next;
        }
        if (P.custom_cfl_base) {
            colour_factor = 128;
            base_corr_x = 0.0625f;
            base_corr_b = 0.875f;
            x_factor_lf = 3;
            b_factor_lf = -5;
        }
    }

    float kx_lf() const { return base_corr_x + (float) x_factor_lf / (float) colour_factor; }
    float kb_lf() const { return base_corr_b + (float) b_factor_lf / (float) colour_factor; }

    // --------------------------------------------------------------------------------------
    MATree make_tree(int num_lfg) {
        typedef MATree M;
        MATree t;
        if (P.tree_preset == 0) {
            t.flatten(M::Leaf(5));
            return t;
        }
        M::P luma, chroma_x, chroma_b, meta;
        if (P.tree_preset == 1) {
            luma = M::Branch(15, 30, M::Leaf(6),
                     M::Branch(15, -31, M::Branch(15, 5, M::Leaf(6), M::Branch(15, -6, M::Leaf(6), M::Leaf(6))), M::Leaf(6)));
            chroma_x = M::Branch(19, 4, M::Leaf(5), M::Branch(19, 0, M::Leaf(5), M::Leaf(5)));
            chroma_b = M::Branch(19, 6, M::Leaf(5), M::Branch(5, 8, M::Leaf(5), M::Leaf(1)));
            meta = M::Branch(0, 2, M::Branch(7, 3, M::Leaf(1), M::Leaf(2)),          // sharpness
                     M::Branch(0, 1, M::Branch(2, 0, M::Leaf(0), M::Branch(7, 0, M::Leaf(0), M::Leaf(0))), // block info rows
                       M::Leaf(5)));                                                 // x/b from y
        } else {
            // stress tree: many properties and predictors
            luma = M::Branch(8, 2, M::Leaf(13), M::Branch(9, 100, M::Leaf(4), M::Branch(10, 0, M::Leaf(3),
                     M::Branch(11, -2, M::Leaf(10), M::Branch(12, 1, M::Leaf(11), M::Branch(13, 0, M::Leaf(12),
                       M::Branch(14, 0, M::Leaf(7), M::Branch(3, 5, M::Leaf(8, 1), M::Leaf(9, -1)))))))));
            chroma_x = M::Branch(16, 0, M::Leaf(6), M::Branch(17, 3, M::Leaf(5), M::Branch(18, 0, M::Leaf(2), M::Leaf(1))));
            chroma_b = M::Branch(20, 1, M::Leaf(5), M::Branch(23, 2, M::Leaf(6), M::Branch(4, 6, M::Leaf(5), M::Leaf(0))));
            meta = M::Branch(0, 2, M::Branch(6, 3, M::Leaf(2), M::Leaf(1)),
                     M::Branch(0, 1, M::Branch(2, 0, M::Leaf(1), M::Leaf(0)), M::Branch(15, 0, M::Leaf(6), M::Leaf(5))));
        }
        M::P lf = M::Branch(0, 0, M::Branch(0, 1, chroma_b, chroma_x), luma);
        t.flatten(M::Branch(1, num_lfg, meta, lf));
        return t;
    }

    // --------------------------------------------------------------------------------------
    void build_lf_group(LfGroupEnc &g, const XYBImage &xyb) {
        int w8 = g.w8, h8 = g.h8;
        // per-cell activity on Y
        std::vector<float> act((size_t) w8 * (size_t) h8);
        std::vector<float> mean[3];
        for (int c = 0; c < 3; ++c) mean[c].assign((size_t) w8 * (size_t) h8, 0.0f);
        for (int y8 = 0; y8 < h8; ++y8) for (int x8 = 0; x8 < w8; ++x8) {
            for (int c = 0; c < 3; ++c) {
                double s = 0, s2 = 0;
                for (int y = 0; y < 8; ++y) for (int x = 0; x < 8; ++x) {
                    float v = xyb.p[c][(size_t) (g.top + y8 * 8 + y) * (size_t) xyb.pw + (size_t) (g.left + x8 * 8 + x)];
                    s += v; s2 += (double) v * v;
                }
                mean[c][(size_t) y8 * (size_t) w8 + (size_t) x8] = (float) (s / 64);
                if (c == 1) act[(size_t) y8 * (size_t) w8 + (size_t) x8] = (float) std::sqrt(std::max(0.0, s2 / 64 - (s / 64) * (s / 64)));
            }
        }
        // ---- LF quantisation (channels X, Y, B); B and X are coded relative to Y (LF CfL)
        static const float MLF[3] = {1.0f / 4096.0f, 1.0f / 512.0f, 1.0f / 256.0f};
        for (int c = 0; c < 3; ++c) g.lfq[c].assign((size_t) w8 * (size_t) h8, 0);
        for (size_t i = 0; i < (size_t) w8 * (size_t) h8; ++i) {
            float step[3];
            for (int c = 0; c < 3; ++c) step[c] = MLF[c] / (float) (P.global_scale * P.quant_lf) * (float) (65536 >> P.extra_prec);
            int qy = (int) std::lrint(mean[1][i] / step[1]);
            float dy = (float) qy * step[1];
            int qx = (int) std::lrint((mean[0][i] - kx_lf() * dy) / step[0]);
            int qb = (int) std::lrint((mean[2][i] - kb_lf() * dy) / step[2]);
            auto cl = [](int v) { return (int16_t) std::min(32000, std::max(-32000, v)); };
            g.lfq[0][i] = cl(qx); g.lfq[1][i] = cl(qy); g.lfq[2][i] = cl(qb);
        }
        // lf indices (j40.h:6566-6570 nesting: X, then B, then Y)
        g.lfidx.assign((size_t) w8 * (size_t) h8, 0);
        for (size_t i = 0; i < (size_t) w8 * (size_t) h8; ++i) {
            int v = 0;
            for (int k = 0; k < nb_lf_thr[0]; ++k) v += g.lfq[0][i] > lf_thr[0][k];
            v *= nb_lf_thr[0] + 1;
            for (int k = 0; k < nb_lf_thr[2]; ++k) v += g.lfq[2][i] > lf_thr[2][k];
            v *= nb_lf_thr[2] + 1;
            for (int k = 0; k < nb_lf_thr[1]; ++k) v += g.lfq[1][i] > lf_thr[1][k];
            g.lfidx[i] = (uint8_t) v;
        }
        // ---- CfL maps and sharpness
        g.xfromy.assign((size_t) g.w64 * (size_t) g.h64, 0);
        g.bfromy.assign((size_t) g.w64 * (size_t) g.h64, 0);
        if (P.cfl) for (size_t i = 0; i < g.xfromy.size(); ++i) {
            g.xfromy[i] = (int16_t) (rng.below(9) - 4);
            g.bfromy[i] = (int16_t) (rng.below(17) - 8);
        }
        g.sharp.assign((size_t) w8 * (size_t) h8, 0);
        for (size_t i = 0; i < g.sharp.size(); ++i) g.sharp[i] = (int16_t) std::min(7, (int) (act[i] * 200.0f));

        // ---- varblock layout: first free cell in raster order gets the next varblock
        g.blockmap.assign((size_t) w8 * (size_t) h8, -1);
        g.cover.assign((size_t) w8 * (size_t) h8, -1);
        for (int y0 = 0; y0 < h8; ++y0) for (int x0 = 0; x0 < w8; ++x0) {
            if (g.cover[(size_t) y0 * (size_t) w8 + (size_t) x0] >= 0) continue;
            int sel = choose_transform(g, act, x0, y0);
            VarBlock vb;
            vb.x8 = x0; vb.y8 = y0; vb.dctsel = sel;
            int r8 = 1 << (kDctSel[sel].log_rows - 3), c8 = 1 << (kDctSel[sel].log_cols - 3);
            float amax = 0;
            for (int i = 0; i < r8; ++i) for (int j = 0; j < c8; ++j) {
                g.cover[(size_t) (y0 + i) * (size_t) w8 + (size_t) (x0 + j)] = (int32_t) g.vbs.size();
                amax = std::max(amax, act[(size_t) (y0 + i) * (size_t) w8 + (size_t) (x0 + j)]);
            }
            g.blockmap[(size_t) y0 * (size_t) w8 + (size_t) x0] = (int32_t) g.vbs.size();
            // adaptive quantisation: busier blocks are quantised more coarsely (smaller HfMul = coarser)
            int hm = P.hfmul_base + (P.hfmul_var > 0 ? rng.below(P.hfmul_var + 1) : 0) + (amax < 0.01f ? 1 : 0);
            vb.hfmul = std::max(1, hm);
            vb.qfidx = 0;
            for (int k = 0; k < nb_qf_thr; ++k) vb.qfidx += (vb.hfmul - 1) >= qf_thr[k];
            stats.transform_hist[sel]++;
            g.vbs.push_back(vb);
        }
        stats.num_varblocks += (int64_t) g.vbs.size();
        // ---- forward transforms + quantisation
        for (VarBlock &vb : g.vbs) quantize_block(g, vb, xyb);
    }

    int choose_transform(const LfGroupEnc &g, const std::vector<float> &act, int x0, int y0) {
        if (P.force_dctsel >= 0) {
            int sel = P.force_dctsel;
            int r8 = 1 << (kDctSel[sel].log_rows - 3), c8 = 1 << (kDctSel[sel].log_cols - 3);
            bool ok = !(x0 % c8) && !(y0 % r8) && x0 + c8 <= g.w8 && y0 + r8 <= g.h8 &&
                      (x0 >> 5) == ((x0 + c8 - 1) >> 5) && (y0 >> 5) == ((y0 + r8 - 1) >> 5);
            for (int i = 0; ok && i < r8; ++i) for (int j = 0; j < c8; ++j) {
                if (g.cover[(size_t) (y0 + i) * (size_t) g.w8 + (size_t) (x0 + j)] >= 0) { ok = false; break; }
            }
            return ok ? sel : 0;
        }
        if (P.transform_mix == 0) return 0;
        // candidate list, large to small
        static const int E6[] = {18, 19, 20, 5, 10, 11, 4, 8, 9, 6, 7};
        static const int ALL[] = {24, 25, 26, 21, 22, 23, 18, 19, 20, 5, 10, 11, 4, 8, 9, 6, 7};
        const int *cand = P.transform_mix == 2 ? ALL : E6;
        int ncand = P.transform_mix == 2 ? 17 : 11;
        for (int k = 0; k < ncand; ++k) {
            int sel = cand[k];
            int r8 = 1 << (kDctSel[sel].log_rows - 3), c8 = 1 << (kDctSel[sel].log_cols - 3);
            if (x0 % c8 || y0 % r8) continue;
            if (x0 + c8 > g.w8 || y0 + r8 > g.h8) continue;
            // must not cross a 256x256 group boundary: guaranteed by alignment for sizes <= 32 cells
            if ((x0 >> 5) != ((x0 + c8 - 1) >> 5) || (y0 >> 5) != ((y0 + r8 - 1) >> 5)) continue;
            bool free_ = true;
            float amax = 0;
            for (int i = 0; i < r8 && free_; ++i) for (int j = 0; j < c8; ++j) {
                if (g.cover[(size_t) (y0 + i) * (size_t) g.w8 + (size_t) (x0 + j)] >= 0) { free_ = false; break; }
                amax = std::max(amax, act[(size_t) (y0 + i) * (size_t) g.w8 + (size_t) (x0 + j)]);
            }
            if (!free_) continue;
            // smoother content tolerates larger transforms
            float thr = P.big_thr / (float) std::sqrt((double) (r8 * c8));
            double take = P.transform_mix == 2 ? 0.08 : (double) P.big_take;
            if ((P.transform_mix == 2 || amax < thr) && rng.uni() < take) return sel;
        }
        // 8x8 family
        float a = act[(size_t) y0 * (size_t) g.w8 + (size_t) x0];
        double r = rng.uni();
        double special = P.transform_mix == 2 ? 0.6 : (a > 0.02f ? 0.3 : 0.06);
        if (r < special) {
            static const int SP[] = {12, 13, 3, 2, 1, 14, 15, 16, 17};
            return SP[rng.below(9)];
        }
        return 0;
    }

    void quantize_block(const LfGroupEnc &g, VarBlock &vb, const XYBImage &xyb) {
        static const float QM[8] = {1.5625f, 1.25f, 1.0f, 0.8f, 0.64f, 0.512f, 0.4096f, 0.32768f};
        static const float QBIAS[3] = {1.0f - 0.05465007330715401f, 1.0f - 0.07005449891748593f, 1.0f - 0.049935103337343655f};
        const DctSel &d = kDctSel[vb.dctsel];
        int R = 1 << d.log_rows, C = 1 << d.log_cols, size = R * C;
        const std::vector<float> &dq = raw_dq[d.param_idx].empty() ? T.dq[d.param_idx] : raw_dq[d.param_idx];
        JG_CHECK((int) dq.size() == size * 3);
        float mult[3];
        mult[1] = 65536.0f / (float) P.global_scale / (float) vb.hfmul;
        mult[0] = mult[1] * QM[P.x_qm_scale];
        mult[2] = mult[1] * QM[P.b_qm_scale];
        int tx = (vb.x8 / 8), ty = (vb.y8 / 8);
        float kx = base_corr_x + (float) g.xfromy[(size_t) ty * (size_t) g.w64 + (size_t) tx] / (float) colour_factor;
        float kb = base_corr_b + (float) g.bfromy[(size_t) ty * (size_t) g.w64 + (size_t) tx] / (float) colour_factor;
        std::vector<double> coef[3];
        for (int c = 0; c < 3; ++c) {
            const float *src = &xyb.p[c][(size_t) (g.top + vb.y8 * 8) * (size_t) xyb.pw + (size_t) (g.left + vb.x8 * 8)];
            if (is_special8(vb.dctsel)) {
                const std::vector<double> &F = T.fwd[vb.dctsel];
                JG_CHECK(F.size() == 64 * 64);
                coef[c].assign(64, 0.0);
                for (int k = 0; k < 64; ++k) {
                    double s = 0;
                    for (int p = 0; p < 64; ++p) s += F[(size_t) k * 64 + (size_t) p] * src[(size_t) (p >> 3) * (size_t) xyb.pw + (size_t) (p & 7)];
                    coef[c][(size_t) k] = s;
                }
            } else {
                // blocks hanging over the padded image edge are clamped to the padded area by construction
                forward_dct2d(src, xyb.pw, R, C, coef[c]);
            }
        }
        // LLF region in the decoder layout: min/8 rows x max/8 columns, row length max(R,C)
        int lmin = std::min(R, C) / 8, lmax = std::max(R, C) / 8, rowlen = std::max(R, C);
        if (is_special8(vb.dctsel)) { lmin = lmax = 1; rowlen = 8; }
        auto is_llf = [&](int i) { return (i / rowlen) < lmin && (i % rowlen) < lmax; };
        std::vector<float> deqy((size_t) size, 0.0f);
        for (int pass = 0; pass < 3; ++pass) {
            int c = pass == 0 ? 1 : pass == 1 ? 0 : 2; // Y first (needed for CfL)
            vb.q[c].assign((size_t) size, 0);
            for (int i = 0; i < size; ++i) {
                if (is_llf(i)) continue;
                double v = coef[c][(size_t) i];
                if (c == 0) v -= (double) kx * deqy[(size_t) i];
                if (c == 2) v -= (double) kb * deqy[(size_t) i];
                double qf = v * dq[(size_t) i * 3 + (size_t) c] / mult[c];
                int q = (int) (qf < 0 ? -std::floor(-qf + (1.0 - P.quant_deadzone)) : std::floor(qf + (1.0 - P.quant_deadzone)));
                q = std::min(30000, std::max(-30000, q));
                vb.q[c][(size_t) i] = q;
                if (q) stats.nonzeros++;
                if (c == 1) {
                    float a = (float) q;
                    a = std::abs(q) <= 1 ? a * QBIAS[1] : a - 0.145f / a;
                    deqy[(size_t) i] = a * mult[1] / dq[(size_t) i * 3 + 1];
                }
            }
        }
        if (P.coef_spike && P.passes == 1 && (vb.x8 * 5 + vb.y8 * 3) % 23 == 0) {
            const int32_t sp[6] = {P.coef_spike, -P.coef_spike, 32767, -32767, 32768, -32768};
            for (int c = 0; c < 3; ++c) for (int k = 0; k < 2; ++k)
                vb.q[c][(size_t) (size - 1 - k * 3 - c)] = sp[(vb.x8 + vb.y8 + c * 2 + k) % 6];
        }
    }

    // --------------------------------------------------------------------------------------
    std::vector<std::vector<int32_t>> make_orders(int pass = 0) {
        // orders[idx*3 + c]; custom ones are small perturbations of the natural order
        std::vector<std::vector<int32_t>> o(13 * 3);
        for (int i = 0; i < 13; ++i) for (int c = 0; c < 3; ++c) {
            o[(size_t) i * 3 + (size_t) c] = T.order[i];
            if ((P.custom_orders >> i) & 1) {
                std::vector<int32_t> &v = o[(size_t) i * 3 + (size_t) c];
                int size = (int) v.size(), skip = size / 64;
                int lim = std::min(size, skip + 200); // keep the Lehmer code short
                Rng r2(P.seed * 131 + (uint64_t) (i * 3 + c) + (uint64_t) pass * 977);
                for (int k = skip; k + 1 < lim; ++k) if (r2.uni() < 0.3) std::swap(v[(size_t) k], v[(size_t) k + 1 + (size_t) r2.below(std::min(3, lim - k - 1))]);
            }
        }
        return o;
    }

    // the share of quantised coefficient v at scan index i that pass `pass` codes; the shares add up to v.
    // Low frequencies go to the early passes, the rest is split by value so that positions meet in several passes.
    int32_t pass_share(int32_t v, int i, int size, int pass) const {
        const int n = P.passes;
        if (n == 1) return v;
        if (P.coef_spike && i % 7 == 3) {
            const int32_t sp = i % 14 == 3 ? P.coef_spike : -P.coef_spike;
            return pass == 0 ? v + sp : pass == 1 ? -sp : 0;
        }
        if (i < size / 8) { // low frequencies: passes 0 and 1 share the value
            if (pass == 0) return v / 2;
            if (pass == 1) return v - v / 2;
            return 0;
        }
        // the rest: round-robin over the passes by position, except that every 5th goes to the last pass with an
        // opposite-sign part in pass 0 (sums that cancel)
        if (i % 5 == 0 && v != 0 && n > 1) return pass == n - 1 ? v + 1 : pass == 0 ? -1 : 0;
        return pass == (i % n) ? v : 0;
    }

    void tokenize_group(int gidx, const std::vector<std::vector<int32_t>> &orders, TokStream &ts, int pass = 0) {
        static const int8_t FREQ_CTX[64] = {
            -1, 0, 2, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26, 28,
            30, 30, 32, 32, 34, 34, 36, 36, 38, 38, 40, 40, 42, 42, 44, 44,
            46, 46, 46, 46, 48, 48, 48, 48, 50, 50, 50, 50, 52, 52, 52, 52,
            54, 54, 54, 54, 56, 56, 56, 56, 58, 58, 58, 58, 60, 60, 60, 60,
        };
        static const int16_t NNZ_CTX[64] = {
            0, 0, 62, 124, 124, 186, 186, 186, 186, 246, 246, 246, 246, 304, 304, 304,
            304, 304, 304, 304, 304, 360, 360, 360, 360, 360, 360, 360, 360, 360, 360, 360,
            360, 412, 412, 412, 412, 412, 412, 412, 412, 412, 412, 412, 412, 412, 412, 412,
            412, 412, 412, 412, 412, 412, 412, 412, 412, 412, 412, 412, 412, 412, 412, 412,
        };
        int grow = gidx / group_cols, gcol = gidx % group_cols;
        const LfGroupEnc &g = lfgs[(size_t) (grow / 8) * (size_t) lfg_cols + (size_t) (gcol / 8)];
        int gx8 = (gcol % 8) * 32, gy8 = (grow % 8) * 32;
        int gw = std::min(P.width, (gcol + 1) * 256) - gcol * 256, gh = std::min(P.height, (grow + 1) * 256) - grow * 256;
        int gw8 = (gw + 7) / 8, gh8 = (gh + 7) / 8;
        int lfsize = (nb_lf_thr[0] + 1) * (nb_lf_thr[1] + 1) * (nb_lf_thr[2] + 1);
        int ctxoff = 495 * nb_block_ctx * group_preset[(size_t) gidx];
        std::vector<int8_t> nonzeros((size_t) gw8 * (size_t) gh8 * 3, 0);
        for (int y8 = 0; y8 < gh8; ++y8) for (int x8 = 0; x8 < gw8; ++x8) {
            int ggx8 = x8 + gx8, ggy8 = y8 + gy8, nzpos = y8 * gw8 + x8;
            int vi = g.blockmap[(size_t) ggy8 * (size_t) g.w8 + (size_t) ggx8];
            if (vi < 0) continue;
            const VarBlock &vb = g.vbs[(size_t) vi];
            const DctSel &d = kDctSel[vb.dctsel];
            int log_size = d.log_rows + d.log_cols;
            int lfidx = g.lfidx[(size_t) ggy8 * (size_t) g.w8 + (size_t) ggx8];
            int bctx0 = (d.order_idx * (nb_qf_thr + 1) + vb.qfidx) * lfsize + lfidx;
            int bctxc = 13 * (nb_qf_thr + 1) * lfsize;
            for (int cyxb = 0; cyxb < 3; ++cyxb) {
                static const int YXB[3] = {1, 0, 2};
                int c = YXB[cyxb];
                const std::vector<int32_t> &order = orders[(size_t) d.order_idx * 3 + (size_t) c];
                std::vector<int32_t> q = vb.q[c];
                if (P.passes > 1) {
                    // shares are assigned per *position* (the order tables differ between passes)
                    const std::vector<int32_t> &nat = T.order[d.order_idx];
                    std::vector<int32_t> share(q.size(), 0);
                    for (int i = 1 << (log_size - 6); i < (1 << log_size); ++i) share[(size_t) nat[(size_t) i]] = pass_share(q[(size_t) nat[(size_t) i]], i, 1 << log_size, pass);
                    q.swap(share);
                }
                int bctx = block_ctx_map[(size_t) (bctx0 + bctxc * cyxb)];
                int pred = x8 > 0 ? (y8 > 0 ? (nonzeros[(size_t) (nzpos - 1) * 3 + (size_t) c] + nonzeros[(size_t) (nzpos - gw8) * 3 + (size_t) c] + 1) >> 1
                                            : nonzeros[(size_t) (nzpos - 1) * 3 + (size_t) c])
                                  : (y8 > 0 ? nonzeros[(size_t) (nzpos - gw8) * 3 + (size_t) c] : 32);
                int nzctx = ctxoff + bctx + (pred < 8 ? pred : 4 + pred / 2) * nb_block_ctx;
                int size = 1 << log_size, first = 1 << (log_size - 6);
                int nz = 0;
                for (int i = first; i < size; ++i) nz += q[(size_t) order[(size_t) i]] != 0;
                // the decoder bounds nnz by 63 * size/64; coefficient positions below `first` are LLF
                JG_CHECK(nz <= (63 << (log_size - 6)));
                ts.push_back({(uint32_t) nzctx, (uint32_t) nz, 0});
                int qnz = (nz + first - 1) / first;
                for (int i = 0; i < (1 << (d.log_rows - 3)); ++i) for (int j = 0; j < (1 << (d.log_cols - 3)); ++j) {
                    nonzeros[(size_t) (nzpos + i * gw8 + j) * 3 + (size_t) c] = (int8_t) qnz;
                }
                int cctx = ctxoff + 458 * bctx + 37 * nb_block_ctx;
                int prev = nz <= (1 << (log_size - 4));
                for (int i = first; nz > 0 && i < size; ++i) {
                    int ctx = cctx + NNZ_CTX[(nz + first - 1) / first] + FREQ_CTX[i >> (log_size - 6)] + prev;
                    int32_t v = q[(size_t) order[(size_t) i]];
                    ts.push_back({(uint32_t) ctx, pack_signed(v), 0});
                    prev = v != 0;
                    nz -= prev;
                }
            }
        }
    }

    // --------------------------------------------------------------------------------------
    void write_headers(BitWriter &bw) {
        bw.put(0xff, 8); bw.put(0x0a, 8);
        write_size_header(bw, P.width, P.height);
        if (!P.alpha) {
            bw.bit(1); // ImageMetadata all_default: 8-bit, XYB, sRGB
        } else {
            bw.bit(0);        // !all_default
            bw.bit(0);        // extra_fields
            bw.bit(0);        // integer samples
            bw.u32(8, 8, 0, 10, 0, 12, 0, 1, 6);
            bw.bit(1);        // modular_16bit_buffers
            bw.u32(1, 0, 0, 1, 0, 2, 4, 1, 12); // one extra channel
            bw.bit(1);        // d_alpha
            bw.bit(1);        // xyb_encoded
            bw.bit(1);        // ColourEncoding all_default
            bw.u64(0);        // extensions
        }
        bw.bit(1); // default_m
        bw.pad();  // frame header starts byte-aligned
        if (!P.explicit_frame_header && !P.alpha) {
            JG_CHECK(P.smooth_lf && P.x_qm_scale == 3 && P.b_qm_scale == 2);
            bw.bit(1);
            return;
        }
        bw.bit(0);
        bw.put(0, 2);       // regular frame
        bw.bit(0);          // VarDCT
        bw.u64(P.smooth_lf ? 0 : 128); // flags
        bw.put(0, 2);       // log_upsampling
        if (P.alpha) bw.put(0, 2); // ec upsampling
        bw.put((uint64_t) P.x_qm_scale, 3);
        bw.put((uint64_t) P.b_qm_scale, 3);
        if (P.passes <= 3) bw.put((uint64_t) (P.passes - 1), 2); // num_passes: selectors 1, 2, 3
        else { bw.put(3, 2); bw.put((uint64_t) (P.passes - 4), 3); }
        if (P.passes > 1) {
            bw.put(0, 2);   // num_ds = 0
            for (int i = 0; i + 1 < P.passes; ++i) bw.put((uint64_t) (i & 3), 2); // shift[i]: parsed and ignored by j40
        }
        bw.bit(0);          // have_crop
        bw.u32(0, 0, 0, 1, 0, 2, 0, 3, 2); // blend mode: replace
        if (P.alpha) bw.u32(0, 0, 0, 1, 0, 2, 0, 3, 2); // blend mode (alpha)
        bw.bit(1);          // is_last
        bw.u32(0, 0, 0, 0, 4, 16, 5, 48, 10); // name length 0
        // restoration filter: all_default must be 0 (SURVEY B-1); cjxl-like gaborish + 2 EPF iterations
        bw.bit(0);
        bw.bit(1);          // gab
        bw.bit(0);          // gab_custom
        bw.put(2, 2);       // epf_iters
        bw.bit(0);          // sharp_custom
        bw.bit(0);          // weight_custom
        bw.bit(0);          // sigma_custom
        bw.u64(0);          // restoration extensions
        bw.u64(0);          // frame extensions
    }

public:
    static void write_size_header(BitWriter &bw, int w, int h) {
        auto ratio_of = [&](int hh) -> int {
            if (w == hh) return 1;
            if ((uint64_t) w == (uint64_t) hh * 6 / 5) return 2;
            if ((uint64_t) w == (uint64_t) hh * 4 / 3) return 3;
            if ((uint64_t) w == (uint64_t) hh * 3 / 2) return 4;
            if ((uint64_t) w == (uint64_t) hh * 16 / 9) return 5;
            if ((uint64_t) w == (uint64_t) hh * 5 / 4) return 6;
            if (w == hh * 2) return 7;
            return 0;
        };
        int ratio = ratio_of(h);
        bool div8 = h % 8 == 0 && h <= 256 && (ratio != 0 || (w % 8 == 0 && w <= 256));
        bw.bit(div8);
        if (div8) bw.put((uint64_t) (h / 8 - 1), 5); else bw.u32((uint32_t) h, 1, 9, 1, 13, 1, 18, 1, 30);
        bw.put((uint64_t) ratio, 3);
        if (ratio == 0) {
            if (div8) bw.put((uint64_t) (w / 8 - 1), 5); else bw.u32((uint32_t) w, 1, 9, 1, 13, 1, 18, 1, 30);
        }
    }

    static std::vector<uint8_t> wrap_container(const std::vector<uint8_t> &code, bool jxlp) {
        std::vector<uint8_t> out = {0, 0, 0, 0x0c, 'J', 'X', 'L', ' ', 0x0d, 0x0a, 0x87, 0x0a,
                                    0, 0, 0, 0x14, 'f', 't', 'y', 'p', 'j', 'x', 'l', ' ', 0, 0, 0, 0, 'j', 'x', 'l', ' '};
        auto box = [&](const char *type, const uint8_t *data, size_t n, const uint8_t *prefix, size_t np) {
            uint32_t size = (uint32_t) (8 + np + n);
            out.push_back((uint8_t) (size >> 24)); out.push_back((uint8_t) (size >> 16));
            out.push_back((uint8_t) (size >> 8)); out.push_back((uint8_t) size);
            out.insert(out.end(), type, type + 4);
            out.insert(out.end(), prefix, prefix + np);
            out.insert(out.end(), data, data + n);
        };
        if (!jxlp) {
            box("jxlc", code.data(), code.size(), nullptr, 0);
        } else {
            size_t cut = std::min(code.size(), std::max<size_t>(1, code.size() / 3));
            // NOTE: index flags as the *reference* reads them (top bit clear = last box), see j40.h:1550
            uint8_t i0[4] = {0x80, 0, 0, 0}, i1[4] = {0, 0, 0, 1};
            box("jxlp", code.data(), cut, i0, 4);
            static const uint8_t junk[5] = {1, 2, 3, 4, 5};
            box("xtra", junk, 5, nullptr, 0);
            box("jxlp", code.data() + cut, code.size() - cut, i1, 4);
        }
        return out;
    }

    TokStream ec_global_ts; // alpha of a single-group frame
    std::vector<float> raw_dq[17];   // weights actually in force where a matrix is sent RAW (multiples of 1/2)
    std::vector<int32_t> raw_int[17];
    TokStream raw_ts[17];
    const CodeSpec *mspec_ptr = nullptr;

private:
    void write_lf_global(BitWriter &bw, const MATree &tree, const CodeSpec &mspec, const EntropyOpts &mo) {
        bw.bit(1); // LF dequant defaults
        bw.u32((uint32_t) P.global_scale, 1, 11, 2049, 11, 4097, 12, 8193, 16);
        bw.u32((uint32_t) P.quant_lf, 16, 0, 1, 5, 1, 8, 1, 16);
        if (!P.custom_block_ctx) {
            bw.bit(1);
        } else {
            bw.bit(0);
            for (int i = 0; i < 3; ++i) {
                bw.put((uint64_t) nb_lf_thr[i], 4);
                for (int j = 0; j < nb_lf_thr[i]; ++j) {
                    uint32_t u = pack_signed(lf_thr[i][j]);
                    bw.u32(u, 0, 4, 16, 8, 272, 16, 65808, 32 > 30 ? 30 : 32);
                }
            }
            bw.put((uint64_t) nb_qf_thr, 4);
            for (int i = 0; i < nb_qf_thr; ++i) bw.u32((uint32_t) (qf_thr[i] - 1), 0, 2, 4, 3, 12, 5, 44, 8);
            // block context cluster map (max 16 clusters): written through the generic cluster-map coder
            CodeSpec tmp;
            tmp.num_ctx = (int) block_ctx_map.size();
            BitWriter cm;
            write_block_ctx_map(bw);
        }
        if (!P.custom_cfl_base) {
            bw.bit(1);
        } else {
            bw.bit(0);
            bw.u32((uint32_t) colour_factor, 84, 0, 256, 0, 2, 8, 258, 16);
            bw.f16(base_corr_x);
            bw.f16(base_corr_b);
            bw.put((uint64_t) (x_factor_lf + 127), 8);
            bw.put((uint64_t) (b_factor_lf + 127), 8);
        }
        bw.bit(1); // global tree present
        EntropyOpts to = mo;
        to.log_alpha_size = 8;
        to.cfg = {4, 1, 0};
        to.max_clusters = 6;
        write_tree(bw, tree, to);
        mspec.write(bw);
        // no global modular channels for a VarDCT frame without extra channels; with one, the global modular
        // header follows (the channel itself is larger than a group and therefore coded per pass group)
        if (P.alpha) {
            ModularHeaderOpts gh;
            write_modular_header_prefix(bw, gh);
            mspec.encode(bw, ec_global_ts); // (an empty ANS stream still carries its final state)
        }
    }

    void write_block_ctx_map(BitWriter &bw) {
        // reuse CodeSpec's private writer through a tiny friend-free trick: build a spec whose
        // cluster map *is* the block context map and emit only that part
        struct Pub : CodeSpec { using CodeSpec::CodeSpec; };
        int n = (int) block_ctx_map.size();
        // simple mode when possible (nbits <= 3), else complex with MTF
        int nbits = ceil_lg((uint32_t) nb_block_ctx);
        if (nbits <= 3) {
            bw.bit(1);
            bw.put((uint64_t) nbits, 2);
            for (int i = 0; i < n; ++i) bw.put(block_ctx_map[(size_t) i], nbits);
        } else {
            bw.bit(0);
            bw.bit(0); // no MTF
            TokStream ts;
            for (int i = 0; i < n; ++i) ts.push_back({0, block_ctx_map[(size_t) i], 0});
            EntropyOpts no;
            no.log_alpha_size = 5;
            no.cfg = {5, 0, 0};
            no.max_clusters = 1;
            CodeSpec nested;
            std::vector<const TokStream *> v{&ts};
            nested.build(1, no, v);
            nested.write(bw);
            nested.encode(bw, ts);
        }
    }

    void write_hf_global(BitWriter &bw, int num_groups, const std::vector<CodeSpec> &cspecs, const std::vector<std::vector<std::vector<int32_t>>> &orders_p) {
        if (!P.raw_dq) {
            bw.bit(1); // default dequantisation matrices
        } else {
            bw.bit(0);
            for (int i = 0; i < 17; ++i) {
                if (!((P.raw_dq >> i) & 1)) { bw.put(0, 3); continue; } // library default
                bw.put(7, 3); // RAW (j40.h:4705-4743)
                bw.f16(2.0f * P.raw_dq_lie); // denominator
                ModularHeaderOpts mh;
                write_modular_header_prefix(bw, mh);
                mspec_ptr->encode(bw, raw_ts[i]);
            }
        }
        bw.put((uint64_t) (P.num_hf_presets - 1), ceil_lg((uint32_t) num_groups));
        // HfPass, once per pass
        for (size_t ps = 0; ps < cspecs.size(); ++ps) {
            int used = P.custom_orders & 0x1fff;
            if (!used) bw.put(2, 2); // selector 2: used_orders = 0
            else { bw.put(3, 2); bw.put((uint64_t) used, 13); }
            if (used) {
                TokStream ts;
                for (int i = 0; i < 13; ++i) if ((used >> i) & 1) {
                    for (int c = 0; c < 3; ++c) lehmer_tokens(T.order[i], orders_p[ps][(size_t) i * 3 + (size_t) c], ts);
                }
                EntropyOpts po;
                po.use_prefix = !P.use_ans;
                po.log_alpha_size = 8;
                po.cfg = {4, 1, 0};
                po.max_clusters = 4;
                CodeSpec pspec;
                std::vector<const TokStream *> v{&ts};
                pspec.build(8, po, v);
                pspec.write(bw);
                pspec.encode(bw, ts);
            }
            cspecs[ps].write(bw);
        }
    }

    // Lehmer code of `target` relative to `natural` for positions >= size/64 (j40.h:5428-5475)
    static void lehmer_tokens(const std::vector<int32_t> &natural, const std::vector<int32_t> &target, TokStream &ts) {
        int size = (int) natural.size(), skip = size / 64;
        std::vector<int32_t> cur(natural.begin() + skip, natural.end());
        std::vector<int32_t> lehmer;
        int n = size - skip;
        int last_nonzero = -1;
        for (int p = 0; p < n; ++p) {
            int32_t want = target[(size_t) (skip + p)];
            int x = 0;
            while (cur[(size_t) (p + x)] != want) ++x;
            lehmer.push_back(x);
            if (x) {
                last_nonzero = p;
                int32_t t = cur[(size_t) (p + x)];
                for (int k = p + x; k > p; --k) cur[(size_t) k] = cur[(size_t) k - 1];
                cur[(size_t) p] = t;
            }
        }
        int end = last_nonzero + 1;
        auto ctx_of = [](uint32_t v) { return (uint32_t) std::min(7, ceil_lg(v + 1)); };
        ts.push_back({ctx_of((uint32_t) size), (uint32_t) end, 0});
        uint32_t prev = 0;
        for (int i = 0; i < end; ++i) {
            ts.push_back({ctx_of(prev), (uint32_t) lehmer[(size_t) i], 0});
            prev = (uint32_t) lehmer[(size_t) i];
        }
    }

    void write_toc_permutation(BitWriter &bw, const std::vector<int> &fileorder) {
        // The decoder reads sizes in file order into sections[], then applies the Lehmer-coded
        // permutation so that sections[] ends up in logical order: after the permutation position i
        // must hold the entry that was read at file position k where fileorder[k] == i.
        int n = (int) fileorder.size();
        std::vector<int32_t> natural((size_t) n), target((size_t) n);
        for (int i = 0; i < n; ++i) natural[(size_t) i] = i;
        for (int k = 0; k < n; ++k) target[(size_t) fileorder[(size_t) k]] = k;
        bw.bit(1);
        TokStream ts;
        // skip = 0 here: emulate lehmer_tokens with skip 0
        {
            std::vector<int32_t> cur = natural;
            std::vector<int32_t> lehmer;
            int last_nonzero = -1;
            for (int p = 0; p < n; ++p) {
                int x = 0;
                while (cur[(size_t) (p + x)] != target[(size_t) p]) ++x;
                lehmer.push_back(x);
                if (x) {
                    last_nonzero = p;
                    int32_t t = cur[(size_t) (p + x)];
                    for (int k = p + x; k > p; --k) cur[(size_t) k] = cur[(size_t) k - 1];
                    cur[(size_t) p] = t;
                }
            }
            int end = last_nonzero + 1;
            auto ctx_of = [](uint32_t v) { return (uint32_t) std::min(7, ceil_lg(v + 1)); };
            ts.push_back({ctx_of((uint32_t) n), (uint32_t) end, 0});
            uint32_t prev = 0;
            for (int i = 0; i < end; ++i) { ts.push_back({ctx_of(prev), (uint32_t) lehmer[(size_t) i], 0}); prev = (uint32_t) lehmer[(size_t) i]; }
        }
        EntropyOpts po;
        po.use_prefix = !P.use_ans;
        po.log_alpha_size = 8;
        po.cfg = {4, 1, 0};
        po.max_clusters = 4;
        CodeSpec ps;
        std::vector<const TokStream *> v{&ts};
        ps.build(8, po, v);
        ps.write(bw);
        ps.encode(bw, ts);
    }
};

} // namespace jxlgen

// j40-b200: VarDCT group decoding -- device functions (also compiled for the CPU kernel-logic tests).
//
// Replaces, per SURVEY.md §8(a):
//   a6  j40__hf_coeffs                       j40.h:6888-7004  -> j40b_hf.h
//   a10 j40__lf_quant / j40__smooth_lf       j40.h:6492-6583  -> lf_dequant(), lf_smooth()
//   a11 j40__hf_metadata (+LLF forward DCT)  j40.h:6585-6720, 5944 -> place_varblocks(), llf_from_lf()
//   a12 j40__dequant_hf                      j40.h:7053-7097  -> inside varblock_to_pixels()
//   a14 inverse DCT family                   j40.h:5802-5990  -> idct_cols(), inverse_dct2d()
//   a15 special 8x8 transforms               j40.h:6002-6246  -> inverse_special()
//   a16 CfL + IDCT dispatch + crop           j40.h:7099-7204
//   a17 XYB -> linear -> sRGB -> quantise    j40.h:7208-7237  } fused into the RGBA8 store
//   a18 j40__render_to_u8x4_rgba             j40.h:7910-7957  }
// Float arithmetic follows the reference operation by operation (one rounding per operation, same
// association); only data movement and the order of independent operations differ.
#pragma once
#include "j40b_modular.h"
#include <math.h>
#include "j40b_tables.inc"

namespace j40b {

static const float h_half_secants[256] = J40B_HALF_SECANTS_INIT;
static const float h_lf2llf[64] = J40B_LF2LLF_INIT;
static const float h_afv_basis[256] = J40B_AFV_BASIS_INIT;
#if defined(__CUDACC__)
static __device__ const float d_half_secants[256] = J40B_HALF_SECANTS_INIT;
static __device__ const float d_lf2llf[64] = J40B_LF2LLF_INIT;
static __device__ const float d_afv_basis[256] = J40B_AFV_BASIS_INIT;
#endif
#if defined(__CUDA_ARCH__)
#define J40B_HALF_SECANT(i) d_half_secants[i]
#define J40B_LF2LLF(i) d_lf2llf[i]
#define J40B_AFV(i) d_afv_basis[i]
#else
#define J40B_HALF_SECANT(i) h_half_secants[i]
#define J40B_LF2LLF(i) h_lf2llf[i]
#define J40B_AFV(i) h_afv_basis[i]
#endif
#define J40B_SQRT2 1.4142135623730951f

// DctSelect -> (log rows, log columns, dequant parameter set, coefficient order), ISO/IEC 18181-1 table
struct DctSelectInfo { int8_t log_rows, log_columns, param_idx, order_idx; };
J40B_HD J40B_INLINE DctSelectInfo dct_select_info(int dctsel) {
    const int8_t T[27][4] = {
        {3, 3, 0, 0}, {3, 3, 1, 1}, {3, 3, 2, 1}, {3, 3, 3, 1}, {4, 4, 4, 2}, {5, 5, 5, 3}, {4, 3, 6, 4}, {3, 4, 6, 4},
        {5, 3, 7, 5}, {3, 5, 7, 5}, {5, 4, 8, 6}, {4, 5, 8, 6}, {3, 3, 9, 1}, {3, 3, 9, 1}, {3, 3, 10, 1}, {3, 3, 10, 1},
        {3, 3, 10, 1}, {3, 3, 10, 1}, {6, 6, 11, 7}, {6, 5, 12, 8}, {5, 6, 12, 8}, {7, 7, 13, 9}, {7, 6, 14, 10},
        {6, 7, 14, 10}, {8, 8, 15, 11}, {8, 7, 16, 12}, {7, 8, 16, 12},
    };
    DctSelectInfo r = {T[dctsel][0], T[dctsel][1], T[dctsel][2], T[dctsel][3]};
    return r;
}

struct NoSync {
    static constexpr bool kFull = true;
    J40B_HD void operator()() const {}
    J40B_HD uint32_t mask() const { return 0xffffffffu; }
    J40B_HD int shift() const { return 0; }
};

// ---------------------------------------------------------------------------------------------
// frame-level constants the kernels need (one per image, in device memory)

enum { MAX_PASSES = 11 }; // j40.h:5259: 1, 2, 3 or 4 + u(3)

struct DFrame {
    int32_t width, height;
    int32_t is_modular, xyb_encoded, bpp;
    int32_t group_size_shift, num_groups, num_lf_groups, gcolumns, grows, ggcolumns, ggrows;
    // VarDCT
    int32_t global_scale, quant_lf;
    float m_lf_scaled[3];
    int32_t skip_adapt_lf_smooth;
    int32_t nb_lf_thr[3], lf_thr[3][15], nb_qf_thr, qf_thr[15];
    int32_t nb_block_ctx, block_ctx_size, num_hf_presets;
    float inv_colour_factor, base_corr_x, base_corr_b, kx_lf, kb_lf;
    float x_qm_mult, b_qm_mult;      // 0.8^(qm_scale - 2)
    float quant_bias[3], quant_bias_num;
    float opsin_inv_mat[9], opsin_bias[3], cbrt_opsin_bias[3], itscale;
    // table arena offsets (bytes)
    uint32_t block_ctx_map_off;      // uint8_t[block_ctx_size]
    uint32_t global_tree_off;        // DTreeNode[]
    uint32_t global_spec_off;        // DCodeSpec (0 = none)
    int32_t num_passes;
    uint32_t coeff_spec_off[MAX_PASSES]; // DCodeSpec per pass
    // device pointers, filled in by the executor (shared library tables or per-image custom ones)
    const float *dq[17];             // float[n][3] per parameter set
    const int32_t *order[MAX_PASSES][13][3]; // int32_t[size] per pass, order and channel
    const float *srgb_thr;           // float[256]: smallest v whose 8-bit output is >= k+1; [255] = NaN
    const uint8_t *srgb_lut;         // uint8[SRGB_LUT_N + 1]: number of thresholds <= b / SRGB_LUT_N
    float srgb_wrap_hi;              // samples >= this wrap around in the reference's int16 cast (srgb_u8_wrapped)
    int32_t global_tree_uses_wp, have_global_tree;
    // modular frames
    int32_t num_channels, num_gm_channels, alpha_channel; // alpha_channel < 0: opaque
    int32_t nb_global_transforms;
    ModTransform global_tr[MOD_MAX_TRANSFORMS];
    int32_t nb_meta_channels;  // palette channels in front of the coded channel list (modular frames)
    int32_t num_out_channels;  // channels after the inverse global transforms (3 colours + extra channels)
    WPParams global_wp;
};

struct DVarblock {
    int32_t coeffoff;   // multiple of 64, offset into the LF group's coefficient space (>> 6 = LLF offset)
    float hfmul_inv;    // 1 / HfMul
    uint16_t x8, y8;    // cell position inside the LF group
    uint8_t dctsel, qfidx;
    uint16_t pad;
};

// per LF group working set (device pointers)
struct DLfGroup {
    int32_t idx, left, top, width, height, width8, height8, width64, height64;
    uint32_t sec_off, sec_size;  // byte span of the LfGroup section in the codestream buffer
    uint64_t sec_start_bit;      // non-zero only for single-section frames
    int16_t *lfq;        // [3][h8*w8] as decoded (Y, X, B)
    float *lfdeq;        // [3][h8*w8] XYB, dequantised
    float *lf;           // [3][h8*w8] XYB, after adaptive smoothing (== lfdeq when smoothing is skipped)
    uint8_t *lfidx;      // [h8*w8]
    int16_t *xfromy, *bfromy; // [h64*w64]
    int16_t *blockinfo;  // [2][h8*w8] (row length nb_varblocks)
    int16_t *sharpness;  // [h8*w8]
    int32_t *blocks;     // [h8*w8]
    DVarblock *varblocks; // [h8*w8]
    float *llf;          // [3][h8*w8]
    int32_t *wp_scratch; // [2*max(w8, nb_varblocks)*5] or null
    int32_t *lz_window;  // [1 << 18] or null
    int32_t nb_varblocks; // written by the kernel
    int32_t has_big;      // written by the kernel: some varblock is larger than 64x64
    uint64_t end_bit;     // written by the kernel (single-section frames continue from here)
    uint32_t *vb_tok;     // [num_passes][3][h8*w8][2] {first token, count}, written by the pass-group kernel (zeroed before)
    // hand-over between the LF kernels (decode LF image -> post-process -> decode HF metadata -> LLF)
    uint64_t mid_bit;     // bit position after the LF image
    int32_t placed;       // stage 1: channels 0-2 decoded and the varblocks placed (by the kernel that ends channel 2, `split` mode)
    // hand-over between the per-channel kernels of a stage (lf_chan_body): where the next channel starts
    uint64_t chan_bit;
    uint32_t chan_ans;
    int32_t chan_copy[3]; // LZ77: num_to_copy, copy_pos, num_decoded
    uint64_t ltree_bit;   // E_LTRE: where the modular header that names a local tree starts (bits from the section start)
    int32_t ltree_stage;  // E_LTRE: 0 = LF image, 1 = HF metadata
    int32_t meta_overrun; // the HF metadata decode ran past the end of the section (lane-per-stream path: read by lf_place_body)
    int32_t extra_prec;
    int32_t nb_tr1, nb_tr2;               // transforms of the LF image / of the HF metadata image
    ModTransform tr1[MOD_MAX_TRANSFORMS], tr2[MOD_MAX_TRANSFORMS];
};

// One non-zero HF coefficient as the entropy kernel leaves it for the back end: 4 bytes, the index in the scan order
// (< 65536 for every transform; the back end applies the order table) and the value. A value outside -32767..32767
// (corrupt or crafted streams only) takes three words: {pos, -32768}, then {0, low half}, {0, high half}; position 0 is
// a varblock's first LLF coefficient and never coded in an HF section, so the two extension words identify themselves
// to a reader that visits the list one word per thread.
struct DToken {
    uint32_t w;
    J40B_HD J40B_INLINE uint32_t pos() const { return w & 0xffffu; }
    J40B_HD J40B_INLINE int32_t val() const { return (int32_t) w >> 16; }
    J40B_HD J40B_INLINE bool is_ext() const { return (w & 0xffffu) == 0; }
    J40B_HD static J40B_INLINE DToken make(uint32_t pos, int32_t val16) { DToken t; t.w = pos | (uint32_t) val16 << 16; return t; }
};
constexpr int32_t TOKEN_WIDE = -32768;
J40B_HD J40B_INLINE bool token_needs_wide(int32_t v) { return (uint32_t) (v + 32767) > 65534u; }
// value of the token at t[0]; t[1], t[2] are read only behind the wide marker
J40B_HD J40B_INLINE int32_t token_value(const DToken *t) {
    const int32_t v = t[0].val();
    if (v != TOKEN_WIDE) return v;
    return (int32_t) ((t[1].w >> 16) | (t[2].w & 0xffff0000u));
}

struct HfVb;
// one record per (pass, group); the records of pass p follow those of pass p - 1 (num_groups apart)
struct DGroup {
    int32_t idx, lfg;            // group index, LF group index
    int32_t gx8, gy8;            // cell offset inside the LF group
    int32_t gw, gh;              // pixel size
    int32_t pass;
    uint32_t sec_off, sec_size;
    uint64_t sec_start_bit;
    uint32_t tok_first, tok_cap; // slice of the image's token array
    int32_t *lz_window;          // or null
    uint8_t *nonzeros;           // [3 * 32] scratch of the coefficient decoder (see HfLane::colbuf)
    uint32_t tok_used;           // written by the kernel
    HfVb *vbs;                   // the group's varblocks in decoding order (shared by the passes; written by hf_prep)
    int32_t nvb;                 // written by hf_prep (pass 0's record only)
    uint64_t end_bit;            // written by the kernel: bit position behind the coefficients
};

// =============================================================================================
// LF group: dequantisation, LF indices, adaptive smoothing (parallel over samples)

// j40.h:6544-6571
J40B_HD inline void lf_dequant(const DFrame &f, const DLfGroup &g, int extra_prec, int tid, int nth) {
    const int n = g.width8 * g.height8;
    const int YXB2XYB[3] = {1, 0, 2};
    for (int c = 0; c < 3; ++c) {
        float mult_lf = J40B_FMUL(J40B_FDIV(f.m_lf_scaled[c], (float) (f.global_scale * f.quant_lf)), (float) (65536 >> extra_prec));
        const int16_t *src = g.lfq + (size_t) YXB2XYB[c] * n;
        float *dst = g.lfdeq + (size_t) c * n;
        for (int i = tid; i < n; i += nth) dst[i] = J40B_FMUL((float) src[i], mult_lf);
    }
    const int16_t *cx = g.lfq + (size_t) 1 * n, *cy = g.lfq, *cb = g.lfq + (size_t) 2 * n;
    for (int i = tid; i < n; i += nth) {
        uint8_t v = 0;
        for (int k = 0; k < f.nb_lf_thr[0]; ++k) v = (uint8_t) (v + (cx[i] > f.lf_thr[0][k]));
        v = (uint8_t) (v * (f.nb_lf_thr[0] + 1));
        for (int k = 0; k < f.nb_lf_thr[2]; ++k) v = (uint8_t) (v + (cb[i] > f.lf_thr[2][k]));
        v = (uint8_t) (v * (f.nb_lf_thr[2] + 1));
        for (int k = 0; k < f.nb_lf_thr[1]; ++k) v = (uint8_t) (v + (cy[i] > f.lf_thr[1][k]));
        g.lfidx[i] = v;
    }
}

// j40.h:6492-6542; reads lfdeq, writes lf (the reference works in place through line buffers, which
// amounts to reading only unsmoothed neighbours)
J40B_HD inline void lf_smooth(const DFrame &f, const DLfGroup &g, int tid, int nth) {
    const float W0 = 0.05226273532324128f, W1 = 0.20345139757231578f, W2 = 0.0334829185968739f;
    const int w = g.width8, h = g.height8, n = w * h;
    float inv_m_lf[3];
    for (int c = 0; c < 3; ++c) {
        inv_m_lf[c] = J40B_FDIV(J40B_FDIV((float) (f.global_scale * f.quant_lf), f.m_lf_scaled[c]), 65536.0f);
    }
    for (int i = tid; i < n; i += nth) {
        int y = i / w, x = i - y * w;
        if (y < 1 || y >= h - 1 || x < 1 || x >= w - 1) {
            for (int c = 0; c < 3; ++c) g.lf[(size_t) c * n + i] = g.lfdeq[(size_t) c * n + i];
            continue;
        }
        float wa[3], gap = 0.5f;
        for (int c = 0; c < 3; ++c) {
            const float *p = g.lfdeq + (size_t) c * n + i;
            float r0 = J40B_FADD(J40B_FADD(J40B_FMUL(p[-w - 1], W2), J40B_FMUL(p[-w], W1)), J40B_FMUL(p[-w + 1], W2));
            float r1 = J40B_FADD(J40B_FADD(J40B_FMUL(p[-1], W1), J40B_FMUL(p[0], W0)), J40B_FMUL(p[1], W1));
            float r2 = J40B_FADD(J40B_FADD(J40B_FMUL(p[w - 1], W2), J40B_FMUL(p[w], W1)), J40B_FMUL(p[w + 1], W2));
            wa[c] = J40B_FADD(J40B_FADD(r0, r1), r2);
            float d = J40B_FSUB(wa[c], p[0]);
            d = J40B_FMUL(d < 0 ? -d : d, inv_m_lf[c]);
            if (gap < d) gap = d;
        }
        gap = J40B_FSUB(3.0f, J40B_FMUL(4.0f, gap));
        gap = gap > 0.0f ? gap : 0.0f; // j40__maxf(0.0f, x)
        for (int c = 0; c < 3; ++c) {
            float s = g.lfdeq[(size_t) c * n + i];
            g.lf[(size_t) c * n + i] = J40B_FADD(J40B_FMUL(J40B_FSUB(wa[c], s), gap), s);
        }
    }
}

// =============================================================================================
// varblock placement (serial; j40.h:6636-6704)

J40B_HD inline void place_varblocks(const DFrame &f, DLfGroup &g, ErrSlot &es, const BitReader &br) {
    const int w8 = g.width8, h8 = g.height8, nvb = g.nb_varblocks;
    const int log_gsize8 = f.group_size_shift - 3;
    const int16_t *info0 = g.blockinfo, *info1 = g.blockinfo + nvb;
    int voff = 0, coeffoff = 0;
    g.has_big = 0;
    for (int y0 = 0; y0 < h8; ++y0) for (int x0 = 0; x0 < w8; ++x0) {
        if (g.blocks[y0 * w8 + x0]) continue;
        if (voff >= nvb) { es.set(br, E_VBLK); return; }
        int dctsel = info0[voff];
        if (dctsel < 0 || dctsel >= 27) { es.set(br, E_DCT); return; }
        DctSelectInfo d = dct_select_info(dctsel);
        int vw8 = 1 << (d.log_columns - 3), vh8 = 1 << (d.log_rows - 3);
        int x1 = x0 + vw8 - 1, y1 = y0 + vh8 - 1;
        if (!(x1 < w8 && (x0 >> log_gsize8) == (x1 >> log_gsize8))) { es.set(br, E_VBLK); return; }
        if (!(y1 < h8 && (y0 >> log_gsize8) == (y1 >> log_gsize8))) { es.set(br, E_VBLK); return; }
        for (int i = 0; i < vh8; ++i) for (int j = 0; j < vw8; ++j) g.blocks[(y0 + i) * w8 + x0 + j] = 1 << 20 | voff;
        g.blocks[y0 * w8 + x0] = (dctsel + 2) << 20 | voff;
        DVarblock vb;
        vb.coeffoff = coeffoff;
        int m1 = info1[voff];
        int qf = 0;
        for (int j = 0; j < f.nb_qf_thr; ++j) qf += m1 >= f.qf_thr[j];
        vb.qfidx = (uint8_t) qf;
        vb.hfmul_inv = J40B_FDIV(1.0f, J40B_FADD((float) m1, 1.0f));
        vb.x8 = (uint16_t) x0; vb.y8 = (uint16_t) y0;
        vb.dctsel = (uint8_t) dctsel;
        // bit 0: not handled by the tile kernel (larger than 64x64, or straddling a 64x64-pixel tile)
        vb.pad = (uint16_t) ((d.log_columns + d.log_rows > 12 || (x0 & 7) + vw8 > 8 || (y0 & 7) + vh8 > 8) ? 1 : 0);
        g.varblocks[voff] = vb;
        if (vb.pad & 1) g.has_big = 1;
        coeffoff += 1 << (d.log_columns + d.log_rows);
        ++voff;
    }
    if (voff != nvb) es.set(br, E_VBLK);
}

// The same with an occupancy bitmap in shared memory (`bitmap`: height8 rows of 8 words, zeroed here) instead of
// reading the `blocks` map back from global memory: a group is 32 cells wide, i.e. one bitmap word, and a
// varblock may not cross a group boundary, so a varblock's cells of one row always lie in one word. Lane 0
// places the varblocks (position, coefficient offset, `blocks` cells; the order of the stores is the
// reference's, so overlapping varblocks of malformed streams end up the same); then all lanes fill in the
// per-varblock attributes.
template <class Sync>
J40B_HD inline void place_varblocks_warp(const DFrame &f, DLfGroup &g, ErrSlot &es, const BitReader &br, uint32_t *bitmap,
                                         int lane, int nlanes, Sync sync) {
    const int w8 = g.width8, h8 = g.height8, nvb = g.nb_varblocks;
    const int words = (w8 + 31) >> 5;
    for (int i = lane; i < h8 * 8; i += nlanes) bitmap[i] = 0;
    sync();
    if (lane == 0) {
        const int log_gsize8 = f.group_size_shift - 3;
        const int16_t *info0 = g.blockinfo;
        int voff = 0, coeffoff = 0, big = 0;
        for (int y0 = 0; y0 < h8 && !es.err; ++y0) for (int wi = 0; wi < words && !es.err; ++wi) {
            const int nbits = w8 - wi * 32 < 32 ? w8 - wi * 32 : 32;
            const uint32_t valid = nbits == 32 ? 0xffffffffu : ((1u << nbits) - 1);
            for (;;) {
                const uint32_t freeb = ~bitmap[y0 * 8 + wi] & valid;
                if (!freeb) break;
                const int x0 = wi * 32 + floor_lg32(freeb & (0u - freeb));
                if (voff >= nvb) { es.set(br, E_VBLK); break; }
                const int dctsel = info0[voff];
                if (dctsel < 0 || dctsel >= 27) { es.set(br, E_DCT); break; }
                const DctSelectInfo d = dct_select_info(dctsel);
                const int vw8 = 1 << (d.log_columns - 3), vh8 = 1 << (d.log_rows - 3);
                const int x1 = x0 + vw8 - 1, y1 = y0 + vh8 - 1;
                if (!(x1 < w8 && (x0 >> log_gsize8) == (x1 >> log_gsize8))) { es.set(br, E_VBLK); break; }
                if (!(y1 < h8 && (y0 >> log_gsize8) == (y1 >> log_gsize8))) { es.set(br, E_VBLK); break; }
                const uint32_t mask = (vw8 >= 32 ? 0xffffffffu : ((1u << vw8) - 1)) << (x0 & 31);
                for (int i = 0; i < vh8; ++i) {
                    bitmap[(y0 + i) * 8 + wi] |= mask;
                    int32_t *row = g.blocks + (size_t) (y0 + i) * w8 + x0;
                    for (int j = 0; j < vw8; ++j) row[j] = 1 << 20 | voff;
                }
                g.blocks[y0 * w8 + x0] = (dctsel + 2) << 20 | voff;
                DVarblock vb;
                vb.coeffoff = coeffoff;
                vb.qfidx = 0;
                vb.hfmul_inv = 0.0f;
                vb.x8 = (uint16_t) x0; vb.y8 = (uint16_t) y0;
                vb.dctsel = (uint8_t) dctsel;
                // bit 0: not handled by the tile kernel (larger than 64x64, or straddling a 64x64-pixel tile)
                vb.pad = (uint16_t) ((d.log_columns + d.log_rows > 12 || (x0 & 7) + vw8 > 8 || (y0 & 7) + vh8 > 8) ? 1 : 0);
                g.varblocks[voff] = vb;
                big |= vb.pad & 1;
                coeffoff += 1 << (d.log_columns + d.log_rows);
                ++voff;
            }
        }
        if (!es.err && voff != nvb) es.set(br, E_VBLK);
        g.has_big = big;
        bitmap[0] = es.err; // hand the outcome to the other lanes
    }
    sync();
    if (bitmap[0]) { if (lane != 0) es.set_raw(bitmap[0]); return; }
    const int16_t *info1 = g.blockinfo + nvb;
    for (int v = lane; v < nvb; v += nlanes) {
        const int m1 = info1[v];
        int qf = 0;
        for (int j = 0; j < f.nb_qf_thr; ++j) qf += m1 >= f.qf_thr[j];
        g.varblocks[v].qfidx = (uint8_t) qf;
        g.varblocks[v].hfmul_inv = J40B_FDIV(1.0f, J40B_FADD((float) m1, 1.0f));
    }
}

// ---------------------------------------------------------------------------------------------
// 1-D transforms over `rep` interleaved columns (element i of column r at [i * rep + r])

// j40__inverse_dct (j40.h:5802-5933) restated iteratively: the recursion's buffers alternate between
// `in` and `out` by depth, so all sub-transforms of one depth can run side by side.
template <class Sync>
J40B_HD inline void idct_cols(float *out, float *in, int t, int rep, int tid, int nth, Sync sync) {
    const int N = 1 << t;
    if (t <= 0) {
        for (int i = tid; i < rep; i += nth) out[i] = in[i];
        sync();
        return;
    }
    const int total = N * rep;
    // going down: de-interleave even/odd, B^T on the odd half
    for (int d = 0; d + 1 < t; ++d) {
        float *src = (d & 1) ? out : in, *dst = (d & 1) ? in : out;
        const int n = N >> d, half = n >> 1;
        for (int e = tid; e < total; e += nth) {
            int r = e % rep, pos = e / rep;
            int off = pos & ~(n - 1), i = pos & (n - 1);
            float v;
            if (i < half) v = src[(off + 2 * i) * rep + r];
            else if (i == half) v = J40B_FMUL(J40B_SQRT2, src[(off + 1) * rep + r]);
            else { int k = i - half; v = J40B_FADD(src[(off + 2 * k - 1) * rep + r], src[(off + 2 * k + 1) * rep + r]); }
            dst[pos * rep + r] = v;
        }
        sync();
    }
    { // size-2 butterflies
        int d = t - 1;
        float *src = (d & 1) ? out : in, *dst = (d & 1) ? in : out;
        for (int e = tid; e < (total >> 1); e += nth) {
            int r = e % rep, pair = e / rep;
            float x = src[(2 * pair) * rep + r], y = src[(2 * pair + 1) * rep + r];
            dst[(2 * pair) * rep + r] = J40B_FADD(x, y);
            dst[(2 * pair + 1) * rep + r] = J40B_FSUB(x, y);
        }
        sync();
    }
    // coming back up: (H_n)^T W^c_n
    for (int d = t - 2; d >= 0; --d) {
        float *src = (d & 1) ? out : in, *dst = (d & 1) ? in : out;
        const int n = N >> d, half = n >> 1;
        for (int e = tid; e < (total >> 1); e += nth) {
            int r = e % rep, q = e / rep;
            int blk = q / half, i = q - blk * half, off = blk * n;
            float mult = J40B_HALF_SECANT(half + i);
            float x = src[(off + i) * rep + r], y = src[(off + half + i) * rep + r];
            float ym = J40B_FMUL(y, mult);
            dst[(off + i) * rep + r] = J40B_FADD(x, ym);
            dst[(off + n - 1 - i) * rep + r] = J40B_FSUB(x, ym);
        }
        sync();
    }
}

// j40__forward_dct_unscaled (j40.h:5768-5883), same iterative restatement
template <class Sync>
J40B_HD inline void fdct_cols(float *out, float *in, int t, int rep, int tid, int nth, Sync sync) {
    const int N = 1 << t;
    if (t <= 0) {
        for (int i = tid; i < rep; i += nth) out[i] = in[i];
        sync();
        return;
    }
    const int total = N * rep;
    for (int d = 0; d + 1 < t; ++d) {
        float *src = (d & 1) ? out : in, *dst = (d & 1) ? in : out;
        const int n = N >> d, half = n >> 1;
        for (int e = tid; e < (total >> 1); e += nth) {
            int r = e % rep, q = e / rep;
            int blk = q / half, i = q - blk * half, off = blk * n;
            float mult = J40B_HALF_SECANT(half + i);
            float x = src[(off + i) * rep + r], y = src[(off + n - 1 - i) * rep + r];
            dst[(off + i) * rep + r] = J40B_FADD(x, y);
            dst[(off + half + i) * rep + r] = J40B_FMUL(J40B_FSUB(x, y), mult);
        }
        sync();
    }
    {
        int d = t - 1;
        float *src = (d & 1) ? out : in, *dst = (d & 1) ? in : out;
        for (int e = tid; e < (total >> 1); e += nth) {
            int r = e % rep, pair = e / rep;
            float x = src[(2 * pair) * rep + r], y = src[(2 * pair + 1) * rep + r];
            dst[(2 * pair) * rep + r] = J40B_FADD(x, y);
            dst[(2 * pair + 1) * rep + r] = J40B_FSUB(x, y);
        }
        sync();
    }
    for (int d = t - 2; d >= 0; --d) {
        float *src = (d & 1) ? out : in, *dst = (d & 1) ? in : out;
        const int n = N >> d, half = n >> 1;
        for (int e = tid; e < total; e += nth) {
            int r = e % rep, pos = e / rep;
            int off = pos & ~(n - 1), i = pos & (n - 1);
            float v;
            if (!(i & 1)) v = src[(off + (i >> 1)) * rep + r];
            else if (i == n - 1) v = src[(off + n - 1) * rep + r];
            else if (i == 1) v = J40B_FADD(J40B_FMUL(J40B_SQRT2, src[(off + half) * rep + r]), src[(off + half + 1) * rep + r]);
            else { int k = i >> 1; v = J40B_FADD(src[(off + half + k) * rep + r], src[(off + half + k + 1) * rep + r]); }
            dst[pos * rep + r] = v;
        }
        sync();
    }
}

// j40__inverse_dct2d (j40.h:5972-5990): buf holds coefficients (transposed layout, row length
// max(R, C)) and receives R x C samples; scratch is as large as buf.
template <class Sync>
J40B_HD inline void inverse_dct2d(float *buf, float *scratch, int log_rows, int log_columns, int tid, int nth, Sync sync) {
    const int R = 1 << log_rows, C = 1 << log_columns, n = R * C;
    if (log_columns > log_rows) {
        for (int e = tid; e < n; e += nth) { int y = e / C, x = e - y * C; scratch[x * R + y] = buf[e]; }
    } else {
        for (int e = tid; e < n; e += nth) scratch[e] = buf[e];
    }
    sync();
    idct_cols(buf, scratch, log_columns, R, tid, nth, sync); // buf: [C][R]
    for (int e = tid; e < n; e += nth) { int y = e / R, x = e - y * R; scratch[x * C + y] = buf[e]; }
    sync();
    idct_cols(buf, scratch, log_rows, C, tid, nth, sync);    // buf: [R][C]
}

// j40__forward_dct2d_scaled_for_llf (j40.h:5944-5970): buf = rows x columns LF samples, result in the
// coefficient layout; scratch as large as buf
template <class Sync>
J40B_HD inline void forward_dct2d_llf(float *buf, float *scratch, int log_rows, int log_columns, int tid, int nth, Sync sync) {
    const int R = 1 << log_rows, C = 1 << log_columns, n = R * C;
    fdct_cols(scratch, buf, log_rows, C, tid, nth, sync);                        // scratch: [R][C]
    for (int e = tid; e < n; e += nth) { int y = e / C, x = e - y * C; buf[x * R + y] = scratch[e]; }
    sync();
    fdct_cols(scratch, buf, log_columns, R, tid, nth, sync);                     // scratch: [C][R]
    for (int e = tid; e < n; e += nth) {
        int y = e / R, x = e - y * R;
        scratch[e] = J40B_FMUL(scratch[e], J40B_FMUL(J40B_LF2LLF(R + x), J40B_LF2LLF(C + y)));
    }
    sync();
    if (log_columns > log_rows) {
        for (int e = tid; e < n; e += nth) { int y = e / R, x = e - y * R; buf[x * C + y] = scratch[e]; }
    } else {
        for (int e = tid; e < n; e += nth) buf[e] = scratch[e];
    }
    sync();
}

// 1-D inverse DCT of N points in place, the recursion of j40__inverse_dct_core (j40.h:5802-5841) unrolled
template <int N> struct Idct1D {
    J40B_HD static J40B_INLINE void run(float *v) {
        float a[N / 2], b[N / 2];
#pragma unroll
        for (int i = 0; i < N / 2; ++i) a[i] = v[2 * i];
        b[0] = J40B_FMUL(J40B_SQRT2, v[1]);
#pragma unroll
        for (int i = 1; i < N / 2; ++i) b[i] = J40B_FADD(v[2 * i - 1], v[2 * i + 1]);
        Idct1D<N / 2>::run(a);
        Idct1D<N / 2>::run(b);
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            float t = J40B_FMUL(b[i], J40B_HALF_SECANT(N / 2 + i));
            v[i] = J40B_FADD(a[i], t);
            v[N - 1 - i] = J40B_FSUB(a[i], t);
        }
    }
};
template <> struct Idct1D<2> {
    J40B_HD static J40B_INLINE void run(float *v) {
        float x = v[0], y = v[1];
        v[0] = J40B_FADD(x, y);
        v[1] = J40B_FSUB(x, y);
    }
};
template <> struct Idct1D<1> { J40B_HD static J40B_INLINE void run(float *) {} };

// `rep` interleaved columns (element i of column r at [i * REP + r]), each through Idct1D: what
// idct_cols computes, as straight-line code for one thread with compile-time sizes
template <int T, int REP>
J40B_HD J40B_INLINE void idct_cols_fixed(float *out, const float *in) {
#pragma unroll
    for (int r = 0; r < REP; ++r) {
        float v[1 << T];
#pragma unroll
        for (int i = 0; i < (1 << T); ++i) v[i] = in[i * REP + r];
        Idct1D<(1 << T)>::run(v);
#pragma unroll
        for (int i = 0; i < (1 << T); ++i) out[i * REP + r] = v[i];
    }
}

// ---------------------------------------------------------------------------------------------
// special 8x8 transforms (one thread each; j40.h:5992-6246)

J40B_HD J40B_INLINE void aux_idct2x2(float *out, const float *in, int x, int y, int S2) {
    int p = y * 8 + x, q = (y * 2) * 8 + (x * 2);
    float c00 = in[p], c01 = in[p + S2], c10 = in[p + S2 * 8], c11 = in[p + S2 * 9];
    out[q + 0] = J40B_FADD(J40B_FADD(J40B_FADD(c00, c01), c10), c11);
    out[q + 1] = J40B_FSUB(J40B_FSUB(J40B_FADD(c00, c01), c10), c11);
    out[q + 8] = J40B_FSUB(J40B_FADD(J40B_FSUB(c00, c01), c10), c11);
    out[q + 9] = J40B_FADD(J40B_FSUB(J40B_FSUB(c00, c01), c10), c11);
}

J40B_HD inline void inverse_dct2x2_pyramid(float *buf) { // DctSelect 2
    float scratch[64];
    aux_idct2x2(buf, buf, 0, 0, 1);
    #pragma unroll
    for (int i = 0; i < 64; ++i) scratch[i] = buf[i];
    #pragma unroll
    for (int y = 0; y < 2; ++y)
        #pragma unroll
        for (int x = 0; x < 2; ++x) aux_idct2x2(scratch, buf, x, y, 2);
    #pragma unroll
    for (int y = 0; y < 4; ++y)
        #pragma unroll
        for (int x = 0; x < 4; ++x) aux_idct2x2(buf, scratch, x, y, 4);
}

J40B_HD inline void inverse_dct4x4_quad(float *buf) { // DctSelect 3
    float scratch[64];
    aux_idct2x2(buf, buf, 0, 0, 1);
    idct_cols_fixed<2, 16>(scratch, buf);
    #pragma unroll
    for (int y = 0; y < 8; ++y)
        #pragma unroll
        for (int x = 0; x < 8; ++x) buf[x * 8 + y] = scratch[y * 8 + x];
    idct_cols_fixed<2, 16>(scratch, buf);
    #pragma unroll
    for (int y = 0; y < 4; ++y)
        #pragma unroll
        for (int x = 0; x < 4; ++x) {
        buf[y * 8 + x] = scratch[(y * 2) * 8 + (x * 2)];
        buf[y * 8 + (x + 4)] = scratch[(y * 2 + 1) * 8 + (x * 2)];
        buf[(y + 4) * 8 + x] = scratch[(y * 2) * 8 + (x * 2 + 1)];
        buf[(y + 4) * 8 + (x + 4)] = scratch[(y * 2 + 1) * 8 + (x * 2 + 1)];
    }
}

J40B_HD inline void inverse_hornuss(float *buf) { // DctSelect 1
    float scratch[64];
    #pragma unroll
    for (int i = 0; i < 64; ++i) scratch[i] = buf[i];
    aux_idct2x2(scratch, buf, 0, 0, 1);
    #pragma unroll
    for (int y = 0; y < 2; ++y)
        #pragma unroll
        for (int x = 0; x < 2; ++x) {
        int pos00 = y * 8 + x, pos11 = (y + 2) * 8 + (x + 2);
        float rsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        #pragma unroll
        for (int iy = 0; iy < 4; ++iy)
            #pragma unroll
            for (int ix = 0; ix < 4; ++ix) {
            rsum[ix] = J40B_FADD(rsum[ix], scratch[(y + iy * 2) * 8 + (x + ix * 2)]);
        }
        float s = J40B_FSUB(J40B_FADD(J40B_FADD(J40B_FADD(rsum[0], rsum[1]), rsum[2]), rsum[3]), scratch[pos00]);
        float sample11 = J40B_FSUB(scratch[pos00], J40B_FMUL(s, 0.0625f));
        scratch[pos00] = scratch[pos11];
        scratch[pos11] = 0.0f;
        #pragma unroll
        for (int iy = 0; iy < 4; ++iy)
            #pragma unroll
            for (int ix = 0; ix < 4; ++ix) {
            buf[(4 * y + iy) * 8 + (4 * x + ix)] = J40B_FADD(scratch[(y + iy * 2) * 8 + (x + ix * 2)], sample11);
        }
    }
}

J40B_HD inline void inverse_dct8x4(float *buf) { // DctSelect 13 ("DCT32" in the reference: 8 columns x 4 rows halves)
    float scratch[64];
    float tmp = J40B_FADD(buf[0], buf[8]);
    buf[8] = J40B_FSUB(buf[0], buf[8]);
    buf[0] = tmp;
    // buf viewed as 4 rows x 16 columns; IDCT-4 down the 16 columns
    idct_cols_fixed<2, 16>(scratch, buf);
    // scratch viewed 8x8 again, transpose into buf
    #pragma unroll
    for (int y = 0; y < 8; ++y)
        #pragma unroll
        for (int x = 0; x < 8; ++x) buf[x * 8 + y] = scratch[y * 8 + x];
    idct_cols_fixed<3, 8>(scratch, buf);
    // columns 01234567 -> 02461357
    #pragma unroll
    for (int y = 0; y < 8; ++y)
        #pragma unroll
        for (int x = 0; x < 8; ++x) buf[y * 8 + (((x & 1) << 2) | (x >> 1))] = scratch[y * 8 + x];
}

J40B_HD inline void inverse_dct4x8(float *buf) { // DctSelect 12 ("DCT23")
    float scratch[64];
    #pragma unroll
    for (int i = 0; i < 64; ++i) scratch[i] = buf[i];
    scratch[0] = J40B_FADD(buf[0], buf[8]);
    scratch[8] = J40B_FSUB(buf[0], buf[8]);
    #pragma unroll
    for (int y = 0; y < 8; ++y)
        #pragma unroll
        for (int x = 0; x < 8; ++x) buf[x * 8 + y] = scratch[y * 8 + x];
    idct_cols_fixed<3, 8>(scratch, buf);
    #pragma unroll
    for (int y = 0; y < 8; ++y)
        #pragma unroll
        for (int x = 0; x < 8; ++x) buf[x * 8 + y] = scratch[y * 8 + x];
    // buf viewed as 4 rows x 16 columns
    idct_cols_fixed<2, 16>(scratch, buf);
    // rows 01234567 -> 02461357
    #pragma unroll
    for (int y = 0; y < 8; ++y) {
        int oy = ((y & 1) << 2) | (y >> 1);
        #pragma unroll
        for (int x = 0; x < 8; ++x) buf[oy * 8 + x] = scratch[y * 8 + x];
    }
}

J40B_HD inline void inverse_afv(float *buf, int flipx, int flipy) { // DctSelect 14..17
    float scratch[64];
    float *bufafv = buf, *buf22 = buf + 16, *buf23 = buf + 32, *buf32 = buf23;
    float *scratchafv = scratch, *scratch22 = scratch + 16, *scratch23 = scratch + 32, *scratch32 = scratch23;
    #pragma unroll
    for (int y = 0; y < 8; y += 2)
        #pragma unroll
        for (int x = 0; x < 8; ++x) {
        scratch[(x % 2) * 16 + (y / 2) * 4 + (x / 2)] = buf[y * 8 + x];
    }
    #pragma unroll
    for (int y = 1; y < 8; y += 2)
        #pragma unroll
        for (int x = 0; x < 8; ++x) {
        scratch32[x * 4 + (y / 2)] = buf[y * 8 + x];
    }
    scratchafv[0] = J40B_FMUL(J40B_FADD(J40B_FADD(buf[0], buf[1]), buf[8]), 4.0f);
    scratch22[0] = J40B_FADD(J40B_FSUB(buf[0], buf[1]), buf[8]);
    scratch32[0] = J40B_FSUB(buf[0], buf[8]);
    // 16x16 basis times the 16 AFV coefficients, accumulated in index order from 0
    #pragma unroll
    for (int i = 0; i < 16; ++i) {
        float sum = 0.0f;
        #pragma unroll
        for (int j = 0; j < 16; ++j) sum = J40B_FADD(sum, J40B_FMUL(scratchafv[j], J40B_AFV(i * 16 + j)));
        bufafv[i] = sum;
    }
    idct_cols_fixed<2, 4>(buf22, scratch22);
    idct_cols_fixed<3, 4>(buf32, scratch32);
    #pragma unroll
    for (int y = 0; y < 4; ++y) {
        #pragma unroll
        for (int x = 0; x < 4; ++x) scratchafv[y * 4 + x] = bufafv[y * 4 + x];
        #pragma unroll
        for (int x = 0; x < 4; ++x) scratch22[x * 4 + y] = buf22[y * 4 + x];
    }
    #pragma unroll
    for (int y = 0; y < 8; ++y) {
        #pragma unroll
        for (int x = 0; x < 4; ++x) scratch23[x * 8 + y] = buf32[y * 4 + x];
    }
    idct_cols_fixed<2, 4>(buf22, scratch22);
    idct_cols_fixed<2, 8>(buf23, scratch23);
    #pragma unroll
    for (int i = 16; i < 64; ++i) scratch[i] = buf[i];
    #pragma unroll
    for (int y = 0; y < 4; ++y) {
        int fy = flipy ? 7 - y : y;
        int afv22pos = fy * 8;
        int dct22pos = (flipy * 4 + y) * 8 + (!flipx * 4);
        int dct23pos = (!flipy * 4 + y) * 8;
        #pragma unroll
        for (int x = 0; x < 4; ++x) buf[afv22pos + (flipx ? 7 - x : x)] = scratchafv[y * 4 + x];
        #pragma unroll
        for (int x = 0; x < 4; ++x) buf[dct22pos + x] = scratch22[y * 4 + x];
        #pragma unroll
        for (int x = 0; x < 8; ++x) buf[dct23pos + x] = scratch23[y * 8 + x];
    }
}

J40B_HD J40B_INLINE bool is_special_8x8(int dctsel) { return dctsel == 1 || dctsel == 2 || dctsel == 3 || (dctsel >= 12 && dctsel <= 17); }

J40B_HD inline void inverse_special(int dctsel, float *buf) {
    switch (dctsel) {
    case 1: inverse_hornuss(buf); break;
    case 2: inverse_dct2x2_pyramid(buf); break;
    case 3: inverse_dct4x4_quad(buf); break;
    case 12: inverse_dct4x8(buf); break;
    case 13: inverse_dct8x4(buf); break;
    case 14: inverse_afv(buf, 0, 0); break;
    case 15: inverse_afv(buf, 1, 0); break;
    case 16: inverse_afv(buf, 0, 1); break;
    case 17: inverse_afv(buf, 1, 1); break;
    }
}

// =============================================================================================
// LLF coefficients of one varblock from the (smoothed) LF planes (j40.h:6669-6683).
// scratch: 2 * (vh8*vw8) floats when the varblock is larger than 8x8.
template <class Sync>
J40B_HD inline void llf_from_lf(const DLfGroup &g, const DVarblock &vb, float *scratch, int tid, int nth, Sync sync) {
    DctSelectInfo d = dct_select_info(vb.dctsel);
    const int n8 = g.width8 * g.height8;
    if (d.log_rows <= 3 && d.log_columns <= 3) {
        for (int c = tid; c < 3; c += nth) g.llf[(size_t) c * n8 + (vb.coeffoff >> 6)] = g.lf[(size_t) c * n8 + vb.y8 * g.width8 + vb.x8];
        sync();
        return;
    }
    const int vw8 = 1 << (d.log_columns - 3), vh8 = 1 << (d.log_rows - 3), n = vw8 * vh8;
    for (int c = 0; c < 3; ++c) {
        float *dst = g.llf + (size_t) c * n8 + (vb.coeffoff >> 6);
        for (int e = tid; e < n; e += nth) {
            int i = e / vw8, j = e - i * vw8;
            dst[e] = g.lf[(size_t) c * n8 + (vb.y8 + i) * g.width8 + vb.x8 + j];
        }
        sync();
        forward_dct2d_llf(dst, scratch, d.log_rows - 3, d.log_columns - 3, tid, nth, sync);
    }
}

// =============================================================================================
// back half: tokens -> coefficients -> samples -> RGBA8 for one varblock

// smallest index k in [0, 255] such that thr[k] > v, i.e. the number of thresholds <= v
J40B_HD J40B_INLINE int srgb_u8_from_linear(const float *thr, float v) {
    int lo = 0, hi = 255; // answer in [0, 255]
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (thr[mid] <= v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Same result through a start table of SRGB_LUT_N + 1 entries: lut[b] = number of thresholds <= b / SRGB_LUT_N
// (b = floor(SRGB_LUT_N v) with v clamped to [0, 1]; the product is exact, SRGB_LUT_N being a power of two). The
// thresholds are at least 1 / (255 * 12.92) apart, wider than a bucket, so at most one more lies inside bucket b (the
// host checks that when it builds the table) and one branch-free step finishes the search. thr[255] is NaN, which ends
// the search at 255; v <= 0 and NaN (clamped to 0; the reference's cast gives 0 after clamping) fall into bucket 0
// below thr[0] > 0. Used by the tile kernel.
enum { SRGB_LUT_N = 4096, SRGB_LUT_BYTES = 4100 };
J40B_HD J40B_INLINE int srgb_u8_lut(const float *thr, const uint8_t *lut, float v) {
    const float vc = fminf(fmaxf(v, 0.0f), 1.0f);
    const int code = lut[(int) J40B_FMUL(vc, (float) SRGB_LUT_N)];
    return code + (thr[code] <= v ? 1 : 0);
}

// Samples outside the range the threshold table covers. The reference converts the sRGB-encoded value with an unchecked
// `(int16_t)` cast (j40.h:7234): on x86-64 that is a truncating float -> int32 conversion (0x80000000 for NaN and for
// values outside the int32 range) of which the low 16 bits are kept, so that huge samples wrap around before the render
// step clamps them (j40.h:7950). `srgb_wrap_hi` (host, same libm as the table) is the smallest v whose encoded value
// reaches 32768; on the negative side (the linear branch, exact arithmetic) the first wrap is near v = -9.947. powf of
// the host's libm is replaced by a double-precision pow rounded to float on the device; an integer step of the
// encoded value takes >= 480 float steps of v here, so a last-bit difference in powf almost never reaches the result.
J40B_HD inline int srgb_u8_wrapped(float v, int bpp) {
    float s;
    if (v <= 0.0031308f) s = J40B_FMUL(12.92f, v);
    else {
#if defined(__CUDA_ARCH__)
        const float p = (float) pow((double) v, (double) (1.0f / 2.4f));
#else
        const float p = powf(v, 1.0f / 2.4f);
#endif
        s = J40B_FSUB(J40B_FMUL(1.055f, p), 0.055f);
    }
    const float q = J40B_FADD(J40B_FMUL((float) ((1 << bpp) - 1), s), 0.5f);
    int32_t t;
    if (!(q >= -2147483648.0f && q < 2147483648.0f)) t = (int32_t) 0x80000000u; // cvttss2si's "integer indefinite"
    else t = (int32_t) q;                                                        // truncation towards zero
    const int32_t i16 = (int32_t) (int16_t) (uint16_t) ((uint32_t) t & 0xffffu);
    const int32_t maxpixel = (1 << bpp) - 1, half = 1 << (bpp - 1);
    const int32_t p8 = imin(imax(0, i16), maxpixel);
    return (p8 * 255 + half) / maxpixel;
}
J40B_HD J40B_INLINE bool srgb_needs_wrap(float v, float wrap_hi) { return v >= wrap_hi || v < -9.9f; }

// Working buffers: coef[3] (X, Y, B) each `size` floats, scratch `size` floats; `size` = R*C.
// Executed by `nth` cooperating threads (a warp or a block) separated by `sync`.
template <class Sync>
J40B_HD inline void varblock_to_pixels(const DFrame &f, const uint8_t *arena, const DLfGroup &g, const DVarblock &vb, int voff,
                                       const DToken *tokens, float *coefx, float *coefy, float *coefb, float *scratch,
                                       uint8_t *rgba, int32_t rgba_stride, int tid, int nth, Sync sync) {
    const DctSelectInfo d = dct_select_info(vb.dctsel);
    const int R = 1 << d.log_rows, C = 1 << d.log_columns, size = R * C;
    const int n8 = g.width8 * g.height8;
    float *coef[3] = {coefx, coefy, coefb};
    // 1. zero + scatter the decoded (quantised) coefficients
    for (int c = 0; c < 3; ++c) for (int i = tid; i < size; i += nth) coef[c][i] = 0.0f;
    sync();
    for (int pass = 0; pass < f.num_passes; ++pass) {
        for (int c = 0; c < 3; ++c) {
            const uint32_t *slot = g.vb_tok + (((size_t) pass * 3 + c) * n8 + voff) * 2;
            const uint32_t first = slot[0], cnt = slot[1];
            const int32_t *order = f.order[pass][d.order_idx][c];
            // positions within one varblock-channel and pass are distinct (a scan order is a permutation); the
            // passes add up (j40.h:6989): integer-valued floats, exact in any order
            for (uint32_t k = tid; k < cnt; k += nth) {
                const DToken *t = tokens + first + k;
                if (t->is_ext()) continue;
                const int32_t pos = order[t->pos()];
                coef[c][pos] = J40B_FADD(coef[c][pos], (float) token_value(t));
            }
        }
        sync();
    }
    // 2. dequantise (j40.h:7078-7094)
    const float *dq = f.dq[d.param_idx];
    float mult[3];
    mult[1] = J40B_FMUL(J40B_FDIV(65536.0f, (float) f.global_scale), vb.hfmul_inv);
    mult[0] = J40B_FMUL(mult[1], f.x_qm_mult);
    mult[2] = J40B_FMUL(mult[1], f.b_qm_mult);
    for (int c = 0; c < 3; ++c) for (int i = tid; i < size; i += nth) {
        float v = coef[c][i];
        if (-1.0f <= v && v <= 1.0f) v = J40B_FMUL(v, f.quant_bias[c]);
        else v = J40B_FSUB(v, J40B_FDIV(f.quant_bias_num, v));
        coef[c][i] = J40B_FMUL(v, J40B_FDIV(mult[c], dq[i * 3 + c]));
    }
    sync();
    // 3. chroma from luma + LLF insertion (j40.h:7138-7175)
    const float kx_hf = J40B_FADD(f.base_corr_x, J40B_FMUL(f.inv_colour_factor, (float) g.xfromy[(vb.y8 / 8) * g.width64 + (vb.x8 / 8)]));
    const float kb_hf = J40B_FADD(f.base_corr_b, J40B_FMUL(f.inv_colour_factor, (float) g.bfromy[(vb.y8 / 8) * g.width64 + (vb.x8 / 8)]));
    for (int i = tid; i < size; i += nth) {
        float y = coefy[i];
        coefx[i] = J40B_FADD(coefx[i], J40B_FMUL(y, kx_hf));
        coefb[i] = J40B_FADD(coefb[i], J40B_FMUL(y, kb_hf));
    }
    sync();
    {
        const int lmin = d.log_rows < d.log_columns ? d.log_rows : d.log_columns;
        const int lmax = d.log_rows < d.log_columns ? d.log_columns : d.log_rows;
        const int vh8 = 1 << (lmin - 3), vw8 = 1 << (lmax - 3);
        const float *l0 = g.llf + (size_t) 0 * n8 + (vb.coeffoff >> 6);
        const float *l1 = g.llf + (size_t) 1 * n8 + (vb.coeffoff >> 6);
        const float *l2 = g.llf + (size_t) 2 * n8 + (vb.coeffoff >> 6);
        for (int e = tid; e < vh8 * vw8; e += nth) {
            int y = e / vw8, x = e - y * vw8;
            int p = y * vw8 * 8 + x;
            coefx[p] = J40B_FADD(l0[e], J40B_FMUL(l1[e], f.kx_lf));
            coefy[p] = l1[e];
            coefb[p] = J40B_FADD(l2[e], J40B_FMUL(l1[e], f.kb_lf));
        }
    }
    sync();
    // 4. inverse transforms
    if (is_special_8x8(vb.dctsel)) {
        for (int c = tid; c < 3; c += nth) inverse_special(vb.dctsel, coef[c]);
        sync();
    } else {
        for (int c = 0; c < 3; ++c) inverse_dct2d(coef[c], scratch, d.log_rows, d.log_columns, tid, nth, sync);
    }
    // 5. XYB -> sRGB -> RGBA8, cropped to the image (j40.h:7208-7237, 7941-7952)
    const float *thr = f.srgb_thr;
    const int px0 = g.left + vb.x8 * 8, py0 = g.top + vb.y8 * 8;
    const int effw = imin(f.width - px0, C), effh = imin(f.height - py0, R);
    for (int e = tid; e < effw * effh; e += nth) {
        int y = e / effw, x = e - y * effw;
        int i = y * C + x;
        float sx = coefx[i], sy = coefy[i], sb = coefb[i];
        float p[3] = {J40B_FADD(sy, sx), J40B_FSUB(sy, sx), sb};
        float lin[3];
        for (int c = 0; c < 3; ++c) {
            float pp = J40B_FSUB(p[c], f.cbrt_opsin_bias[c]);
            lin[c] = J40B_FMUL(J40B_FADD(J40B_FMUL(J40B_FMUL(pp, pp), pp), f.opsin_bias[c]), f.itscale);
        }
        uint8_t *o = rgba + (size_t) (py0 + y) * (size_t) rgba_stride + (size_t) (px0 + x) * 4;
        uint8_t out[4];
        for (int c = 0; c < 3; ++c) {
            float v = J40B_FADD(J40B_FADD(J40B_FMUL(lin[0], f.opsin_inv_mat[c * 3 + 0]), J40B_FMUL(lin[1], f.opsin_inv_mat[c * 3 + 1])),
                                J40B_FMUL(lin[2], f.opsin_inv_mat[c * 3 + 2]));
            out[c] = (uint8_t) (srgb_needs_wrap(v, f.srgb_wrap_hi) ? srgb_u8_wrapped(v, 8) : srgb_u8_from_linear(thr, v));
        }
        out[3] = 255;
        o[0] = out[0]; o[1] = out[1]; o[2] = out[2]; o[3] = out[3];
    }
    sync();
}

} // namespace j40b

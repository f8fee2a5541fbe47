// j40-b200: host-side front end. Parses everything that is serial and tiny (container, image and frame
// headers, TOC, LfGlobal, HfGlobal: SURVEY.md §2 "OUT OF SCOPE (host parse)" rows) and turns it into a
// FramePlan: device-ready tables plus the byte spans of the sections the kernels decode.
#pragma once
#include "j40b_vardct.h"
#include <string>
#include <vector>

namespace j40b {

struct Arena {
    std::vector<uint8_t> bytes;
    Arena() { bytes.resize(16, 0); } // offset 0 is reserved as "none"
    uint32_t alloc(size_t n, size_t align = 8) {
        size_t off = (bytes.size() + align - 1) / align * align;
        bytes.resize(off + n, 0);
        return (uint32_t) off;
    }
    template <class T> T *at(uint32_t off) { return (T *) (bytes.data() + off); }
};

struct SectionRef {
    uint64_t off = 0;     // byte offset in the linearised codestream
    uint32_t size = 0;    // bytes available to this section
    uint64_t start_bit = 0;
    int32_t rank = 0;     // position in the reference's decoding order (for first-error selection)
};

struct ImageInfo {
    int32_t width = 0, height = 0, bpp = 8, exp_bits = 0;
    int modular_16bit_buffers = 1, num_extra_channels = 0, xyb_encoded = 1, want_icc = 0;
    int cspace_grey = 0;
    struct EC { int type, bpp, exp_bits, dim_shift, alpha_associated; } ec[4];
    float intensity_target = 255.0f;
    float opsin_inv_mat[3][3], opsin_bias[3], quant_bias[3], quant_bias_num;
    int anim = 0, anim_have_timecodes = 0;
};

struct FrameInfo {
    int is_last = 1, type = 0, is_modular = 0;
    int has_noise = 0, has_patches = 0, has_splines = 0, use_lf_frame = 0, skip_adapt_lf_smooth = 0;
    int do_ycbcr = 0, jpeg_upsampling = 0, group_size_shift = 8, x_qm_scale = 3, b_qm_scale = 2, num_passes = 1;
    int32_t width = 0, height = 0;
    int32_t grows = 0, gcolumns = 0, ggrows = 0, ggcolumns = 0;
    int64_t num_groups = 0, num_lf_groups = 0;
};

struct FramePlan {
    uint32_t err = 0;
    uint32_t deferred_err = 0; // a host-side error the reference would only raise after everything else in the host's part
    // linearised codestream (points into the caller's buffer for bare codestreams)
    const uint8_t *cs = nullptr;
    size_t cs_size = 0;
    std::vector<uint8_t> cs_owned;

    ImageInfo im;
    FrameInfo fh;
    DFrame df;          // table offsets refer to `arena`; pointer-typed members are filled by the executor
    Arena arena;        // per-image tables (code specs, tree, block context map, custom dq / orders)
    bool single_section = false;
    uint32_t trailing_box_err = 0;     // container: a misplaced box follows the last codestream box beyond 64 KiB (`box?`); single-section frames run into it
    bool trailing_partial_box = false; // container: a truncated box header follows the last codestream box (beyond 64 KiB)
    std::vector<SectionRef> lfg_sec, pg_sec; // per LF group / per (pass, group): index pass * num_groups + group
    uint64_t end_codeoff = 0;
    // custom (non-library) tables: 0 = use the process-wide default tables
    uint32_t custom_dq_off[17] = {0};
    uint32_t custom_order_off[MAX_PASSES][13][3] = {{{0}}};
    // modular frames: global image
    ModImage gmod;               // channel geometry (px pointers are assigned by the executor)
    SectionRef gmod_sec;         // where the globally-coded channels start (inside LfGlobal)
    int num_gm_channels = 0;
    bool gmod_has_stream = false; // an entropy-coded stream (possibly empty) follows the global header
    int rank_lf_global = 0;
    // modular sub-bitstreams whose header names a tree of their own (j40.h:3827-3835): header, tree and code
    // spec are read on the host; the device starts at `start_bit` with these tables
    struct LocalHeader {
        bool present = false;
        ModImage hdr;            // wp parameters and transforms as parsed (channel geometry of the sub-image)
        uint32_t tree_off = 0, spec_off = 0;
        int32_t uses_wp = 0;
        uint64_t start_bit = 0;  // first bit of the channel data, relative to the section start
        uint32_t host_err = 0;   // the header failed to parse: reported through the section's error slot
    };
    std::vector<LocalHeader> lfg_local;  // VarDCT frames: [2 * LF group + stage] once Batch::resolve_local_trees has read them (else empty)
    LocalHeader gmod_local;              // the global image's
    std::vector<LocalHeader> pg_local;   // per pass group of a modular frame (empty: none is local)
};

// process-wide immutable tables (library dequantisation matrices, natural orders, sRGB thresholds)
struct GlobalTables {
    std::vector<float> dq[17];        // [n][3]
    std::vector<int32_t> order[13];
    float srgb_thr[256]; // [255] = NaN: ends the search of srgb_u8_lut()
    float srgb_wrap_hi;     // smallest sample whose encoded value reaches 32768 (the reference's int16 cast wraps from there)
    uint8_t srgb_lut[4100]; // SRGB_LUT_BYTES
    static const GlobalTables &get();
};

// Parses `data` up to and including HfGlobal. Returns 0 or a four-character error code.
uint32_t parse_frame(const uint8_t *data, size_t size, FramePlan &plan);

// Reads the modular header, MA tree and code spec that start `start_bit` bits into LF-group section `lfg` of a VarDCT
// frame (stage 0: the LF image, 1: the HF metadata image with `nb_varblocks` varblocks) into plan.lfg_local.
void parse_lf_group_local_tree(FramePlan &plan, size_t lfg, int stage, uint64_t start_bit, int32_t nb_varblocks);

// helpers shared with tests
std::vector<float> compute_dq_matrix_default(int idx);
std::vector<int32_t> compute_natural_order(int log_rows, int log_columns);
void compute_srgb_thresholds(int bpp, float *thr255);

} // namespace j40b

"""Seeded test-stream catalogue shared by the CPU (emulator) and GPU parity tests."""

# known-answer vectors hand-assembled in SURVEY.md Appendix C (hex of the whole file, expected outcome)
KNOWN_ANSWER = {
    "V1_modular_8x8_local_tree": ("ff0a4106088100000c00897c3e", (0, 0, 0, 255)),
    "V2_vardct_8x8": ("ff0a4106012800790097cf1f0a2b1f9301", (144, 144, 144, 255)),
    "V3_vardct_264x8_5sections": (
        "ff0a38000e0e011c00070800000000790097cf1f0a014c555555550000000000000000c00100000000000000000000000000291f", None),
    "V4_vardct_8x8_ans": ("ff0a41060148007900970fba01c0958f01e838041840032800", (144, 144, 144, 255)),
}

VARDCT_CASES = [
    # (name, width, height, seed, options)
    ("c1_256_dct8_single_section", 256, 256, 1, dict(mix=0, tree=0, cfl=0)),
    ("e6like_ans", 520, 392, 10, dict(mix=1, tree=1)),
    ("e6like_prefix", 520, 392, 11, dict(mix=1, tree=1, ans=0)),
    ("all_transforms_stress_tree", 520, 392, 12, dict(mix=2, tree=2)),
    ("block_ctx_orders_presets", 520, 392, 13, dict(mix=1, tree=1, block_ctx=1, orders=0x1f, presets=2)),
    ("no_smooth_extra_prec", 520, 392, 14, dict(mix=1, tree=2, smooth=0, extra_prec=1)),
    ("cfl_base_qm", 520, 392, 15, dict(mix=1, tree=1, cfl_base=1, x_qm=2, b_qm=4)),
    ("default_frame_header", 520, 392, 16, dict(mix=1, explicit_fh=0)),
    ("container_jxlc", 520, 392, 17, dict(mix=1, tree=1, container=1)),
    ("container_jxlp", 520, 392, 18, dict(mix=1, tree=1, container=1, jxlp=1)),
    ("permuted_toc", 520, 392, 19, dict(mix=1, tree=1, permuted=1)),
    ("lz77_coeffs_ans", 520, 392, 20, dict(mix=1, lz77=1)),
    ("lz77_coeffs_prefix", 520, 392, 21, dict(mix=1, ans=0, lz77=1)),
    ("tiny_8x8", 8, 8, 5, dict(mix=1, tree=1)),
    ("ragged_264x8", 264, 8, 5, dict(mix=1, tree=1)),
    ("ragged_257x255", 257, 255, 5, dict(mix=1, tree=1)),
    ("wide_1000x300", 1000, 300, 5, dict(mix=1, tree=1)),
    ("two_lf_groups_2100x300", 2100, 300, 6, dict(mix=1, tree=1, hfmul=6)),
    # VarDCT frames with an alpha extra channel (coded per pass group behind the coefficients; the reference
    # decodes and then discards it: A = 255)
    ("alpha_extra_channel_ans", 520, 392, 22, dict(mix=1, tree=1, alpha=1)),
    ("alpha_extra_channel_prefix", 300, 520, 23, dict(mix=1, tree=2, alpha=1, ans=0)),
    # ... in a single-group frame: coded between LfGlobal and HfGlobal, decoded (and dropped) by the host
    ("alpha_single_group", 200, 100, 24, dict(mix=1, tree=1, alpha=1)),
    ("alpha_single_group_256_prefix", 256, 256, 25, dict(mix=1, tree=2, alpha=1, ans=0)),
    # RAW dequantisation matrices (modular images inside HfGlobal, j40.h:4705-4743) for 8x8 and 16x16
    ("raw_dq_8x8_single_group", 200, 136, 26, dict(mix=1, tree=1, raw_dq=1)),
    ("raw_dq_8x8_16x16", 520, 392, 27, dict(mix=1, tree=1, raw_dq=0x11)),
    ("raw_dq_prefix_alpha", 300, 200, 28, dict(mix=1, tree=2, raw_dq=0x11, ans=0, alpha=1)),
    # more than one pass (j40.h:6844-6866, 7847-7854): per-pass orders and code specs, coefficients added up
    ("passes2", 520, 392, 33, dict(mix=1, tree=1, passes=2)),
    ("passes3_custom_orders", 520, 392, 34, dict(mix=1, tree=1, passes=3, orders=0x1f)),
    ("passes4_prefix", 520, 392, 35, dict(mix=1, tree=1, passes=4, ans=0)),
    ("passes2_presets_block_ctx", 520, 392, 36, dict(mix=1, tree=1, passes=2, presets=2, block_ctx=1)),
    ("passes3_permuted_toc", 520, 392, 37, dict(mix=1, tree=1, passes=3, permuted=1)),
    ("passes5_lz77", 300, 264, 38, dict(mix=1, tree=1, passes=5, lz77=1)),
    ("passes2_dct128", 512, 512, 39, dict(force=21, hfmul=12, tree=1, passes=2)),
    ("passes2_all_transforms", 520, 392, 40, dict(mix=2, tree=2, passes=2)),
    # LF-group sub-bitstreams with MA trees of their own (j40.h:3827-3835 reached from 6722-6783): the device reports where
    # the header starts, the host reads header + tree + code spec there, the batch is decoded again (E_LTRE round)
    ("lf_local_tree_lf_image", 520, 392, 41, dict(mix=1, tree=1, lf_local_tree=1)),
    ("lf_local_tree_hf_meta_prefix", 520, 392, 42, dict(mix=1, tree=1, lf_local_tree=2, ans=0)),
    ("lf_local_tree_both_two_lf_groups", 2100, 300, 43, dict(mix=1, tree=2, lf_local_tree=7, hfmul=6)),
    ("lf_local_tree_single_group", 200, 100, 44, dict(mix=1, tree=1, lf_local_tree=3, lz77=1)),
    # coefficient values outside 16 bits (three-word form of the token list), cancelling over the passes
    ("wide_tokens_two_passes", 520, 392, 45, dict(mix=1, tree=1, passes=2, coef_spike=40000)),
    ("wide_tokens_three_passes_all_transforms", 300, 264, 46, dict(mix=2, tree=1, passes=3, coef_spike=5000000)),
]

# samples far outside [0, 1]: a RAW dequantisation matrix whose written denominator is `raw_dq_lie` times the one the
# writer quantised with. The reference's unchecked (int16_t) cast wraps such samples around (j40.h:7234).
WRAP_CASES = [
    ("int16_wrap_x8", 136, 72, 5, dict(mix=0, tree=1, raw_dq=1, raw_dq_lie=8, hfmul=4)),
    ("int16_wrap_x512", 136, 72, 5, dict(mix=0, tree=1, raw_dq=1, raw_dq_lie=512, hfmul=4)),
    ("int16_wrap_x16384_mixed", 264, 136, 6, dict(mix=1, tree=1, raw_dq=0x11, raw_dq_lie=16384, hfmul=4)),
    # ... and single-pass coefficient values outside 16 bits (+-70000, +-32767, +-32768) in a few varblocks
    ("wide_tokens_single_pass", 264, 136, 7, dict(mix=1, tree=1, coef_spike=70000)),
    ("wide_tokens_single_pass_dct64", 256, 256, 8, dict(force=18, tree=1, hfmul=12, coef_spike=1 << 21)),
]

MODULAR_CASES = [
    ("fjxl_like_prefix_lz77", 600, 400, 2, dict()),
    ("ans_no_lz77", 600, 400, 2, dict(ans=1, lz77=0)),
    ("wp_tree", 600, 400, 2, dict(tree=2)),
    ("alpha", 600, 400, 2, dict(alpha=1)),
    ("group_shift7", 600, 400, 2, dict(group_shift=7)),
    ("group_shift9", 600, 400, 2, dict(group_shift=9)),
    ("no_rct", 600, 400, 2, dict(rct=-1)),
    ("single_leaf", 600, 400, 2, dict(tree=0, lz77=0)),
    ("ans_lz77", 600, 400, 2, dict(ans=1, lz77=1)),
    ("container", 600, 400, 2, dict(container=1)),
    ("rct13", 600, 400, 2, dict(rct=13)),
    ("rct27", 600, 400, 2, dict(rct=27)),
    ("rct41", 600, 400, 2, dict(rct=41)),
    ("single_group_8x8", 8, 8, 4, dict()),
    ("single_group_256", 256, 256, 4, dict(tree=2, ans=1, alpha=1)),
    ("ragged_257x129", 257, 129, 4, dict(tree=2, ans=1, alpha=1)),
    # trees and code specs local to a sub-bitstream (j40.h:3827-3835), mixed with the global one / without any
    ("local_tree_odd_groups", 600, 400, 6, dict(local_tree=1)),
    ("local_tree_all_groups_wp_ans", 600, 400, 6, dict(local_tree=2, ans=1, lz77=0, tree=2, alpha=1)),
    ("local_tree_single_group", 200, 100, 6, dict(local_tree=1, ans=1)),
    # palette transform (j40.h:3762-3792, 4402-4490): table coded with the global image, index channel per group,
    # implicit entries (index < 0, index >= nb_colours) in a few rows
    ("palette_single_group", 200, 100, 7, dict(palette=1)),
    ("palette_groups_prefix_lz77", 600, 400, 7, dict(palette=1)),
    ("palette_alpha_wp_ans", 600, 400, 7, dict(palette=1, alpha=1, ans=1, lz77=0, tree=2)),
    ("palette_local_trees_shift7", 300, 300, 7, dict(palette=1, group_shift=7, local_tree=1)),
    # delta palettes (nb_deltas > 0, j40.h:4416-4480): entries below nb_deltas are added to a prediction from the restored
    # neighbours -- gradient, weighted predictor, the 7-tap predictor 13, select
    ("palette_delta_gradient", 600, 400, 7, dict(palette=1, pal_deltas=3, pal_pred=5)),
    ("palette_delta_wp_single_group", 200, 100, 7, dict(palette=1, pal_deltas=40, pal_pred=6)),
    ("palette_delta_wp_groups", 600, 400, 7, dict(palette=1, pal_deltas=40, pal_pred=6, ans=1, lz77=0)),
    ("palette_delta_pred13_alpha", 600, 400, 7, dict(palette=1, pal_deltas=200, pal_pred=13, alpha=1)),
    ("palette_delta_all_select_shift7", 300, 300, 7, dict(palette=1, pal_deltas=1000, pal_pred=4, group_shift=7, local_tree=1)),
]


# Error-code differences on corrupt input that are known and documented ("not built"): (reference, ours).
# Everything else must agree exactly. (Round 1 tolerated 10 % mismatches; the last known class -- a bit flip that turns
# an LF group's modular header into one with a tree of its own -- went away with the E_LTRE round.)
ALLOWED_CODE_MISMATCHES = set()


def force_cases():
    return [(f"force_dctsel_{sel}", 512, 512, 3, dict(force=sel, hfmul=12, tree=1)) for sel in range(27)]


def make(gen, kind, w, h, seed, opts):
    fn = gen.vardct if kind == "vardct" else gen.modular
    data, _ = fn(w, h, seed=seed, **opts)
    return data


def corruptions(data: bytes, seed: int, n: int):
    """Deterministic corrupt variants: truncations, bit flips, appended bytes."""
    import random
    rnd = random.Random(seed)
    out = []
    for i in range(n):
        kind = i % 3
        if kind == 0:
            cut = rnd.randrange(2, max(3, len(data)))
            out.append((f"truncate@{cut}", data[:cut]))
        elif kind == 1:
            b = bytearray(data)
            pos = rnd.randrange(0, len(b))
            b[pos] ^= 1 << rnd.randrange(8)
            out.append((f"flip@{pos}", bytes(b)))
        else:
            out.append((f"append{i}", data + bytes([rnd.randrange(256)])))
    return out

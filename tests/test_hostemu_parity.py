"""CPU checks of the kernel logic: the kernel bodies of j40_b200/csrc/j40b_exec.h, executed single-threaded by
tests/hostemu, must reproduce the oracle's RGBA bytes and error codes exactly. (The same comparisons against
the real CUDA library are in test_gpu_parity.py, marked `gpu`.)"""
import numpy as np
import pytest

from tests import streams


def _cmp(oracle, emu, data):
    a, ea, _, sa = oracle.decode(data)
    b, eb, sb = emu.decode(data)
    assert ea == eb
    if a is None:
        assert b is None
    else:
        assert sa == sb
        assert np.array_equal(a, b), int((a != b).sum())


def test_known_answer(oracle, emu):
    for name, (hx, _) in streams.KNOWN_ANSWER.items():
        _cmp(oracle, emu, bytes.fromhex(hx))


@pytest.mark.parametrize("case", streams.VARDCT_CASES, ids=[c[0] for c in streams.VARDCT_CASES])
def test_vardct(oracle, emu, gen, case):
    _, w, h, seed, opts = case
    _cmp(oracle, emu, streams.make(gen, "vardct", w, h, seed, opts))


@pytest.mark.parametrize("case", streams.force_cases()[::3], ids=[c[0] for c in streams.force_cases()[::3]])
def test_each_transform(oracle, emu, gen, case):
    _, w, h, seed, opts = case
    _cmp(oracle, emu, streams.make(gen, "vardct", 256, 256, seed, opts))


@pytest.mark.parametrize("case", streams.MODULAR_CASES, ids=[c[0] for c in streams.MODULAR_CASES])
def test_modular(oracle, emu, gen, case):
    _, w, h, seed, opts = case
    _cmp(oracle, emu, streams.make(gen, "modular", w, h, seed, opts))


def test_error_codes_on_corrupt_streams(oracle, emu, gen):
    base = [streams.make(gen, "vardct", 264, 136, 3, dict(mix=1, tree=1)),
            streams.make(gen, "vardct", 64, 64, 4, dict(mix=1, tree=1, ans=0)),
            streams.make(gen, "modular", 300, 200, 5, dict())]
    total = mismatches = 0
    for bi, data in enumerate(base):
        for name, bad in streams.corruptions(data, bi, 60):
            a, ea, _, _ = oracle.decode(bad)
            b, eb, _ = emu.decode(bad)
            total += 1
            # success/failure must agree; when both succeed the pixels must agree
            assert (ea == "") == (eb == ""), (bi, name, ea, eb)
            if ea == "":
                assert np.array_equal(a, b), (bi, name)
            elif ea != eb and (ea, eb) not in streams.ALLOWED_CODE_MISMATCHES:
                mismatches += 1
    assert mismatches == 0, (mismatches, total)


def test_error_codes_on_corrupt_streams_newer_features(oracle, emu, gen):
    """palette, local MA trees, RAW dequantisation matrices, extra channels: corrupt variants must fail or decode
    exactly like the reference. (Round 1 skipped the extra-channel data of multi-group VarDCT frames, which the
    reference decodes and then drops, and so missed damage confined to it; it is decoded on the device now.)"""
    base = {
        "palette": streams.make(gen, "modular", 300, 280, 3, dict(palette=1)),
        "palette_single_alpha": streams.make(gen, "modular", 120, 90, 3, dict(palette=1, alpha=1)),
        "palette_delta_wp": streams.make(gen, "modular", 300, 280, 3, dict(palette=1, pal_deltas=20, pal_pred=6)),
        "vardct_lf_local_trees": streams.make(gen, "vardct", 264, 136, 8, dict(mix=1, tree=1, lf_local_tree=3)),
        "vardct_passes3": streams.make(gen, "vardct", 300, 264, 9, dict(mix=1, tree=1, passes=3)),
        "local_tree": streams.make(gen, "modular", 300, 280, 4, dict(local_tree=1)),
        "local_tree_all_ans": streams.make(gen, "modular", 300, 280, 4, dict(local_tree=2, ans=1, lz77=0, tree=2)),
        "raw_dq": streams.make(gen, "vardct", 264, 136, 5, dict(mix=1, tree=1, raw_dq=0x11)),
        "alpha_single_group": streams.make(gen, "vardct", 200, 100, 6, dict(mix=1, tree=1, alpha=1)),
        "alpha_multi_group": streams.make(gen, "vardct", 300, 264, 7, dict(mix=1, tree=1, alpha=1)),
    }
    total = mismatches = 0
    for bi, (name, data) in enumerate(sorted(base.items())):
        for cname, bad in streams.corruptions(data, 100 + bi, 30):
            a, ea, _, _ = oracle.decode(bad)
            b, eb, _ = emu.decode(bad)
            total += 1
            assert (ea == "") == (eb == ""), (name, cname, ea, eb)
            if ea == "":
                assert np.array_equal(a, b), (name, cname)
            elif ea != eb and (ea, eb) not in streams.ALLOWED_CODE_MISMATCHES:
                mismatches += 1
    assert mismatches == 0, (mismatches, total)


def test_token_arena_overflow_is_retried(oracle, emu, gen, monkeypatch):
    """J40B_TEST_TOKEN_SQUEEZE shrinks the first token-arena estimate to 64 tokens per group: every group reports the
    internal code `tokv` and the decode is repeated with worst-case capacity (same logic as j40b_batch_wait)"""
    monkeypatch.setenv("J40B_TEST_TOKEN_SQUEEZE", "1")
    _cmp(oracle, emu, streams.make(gen, "vardct", 520, 392, 81, dict(mix=1, tree=1)))
    _cmp(oracle, emu, streams.make(gen, "vardct", 64, 64, 82, dict(mix=1, tree=1, ans=0)))


def test_trailing_bytes_behind_the_container(oracle, emu, gen):
    """a truncated box header behind the codestream box: the reference reports `shrt` only if it lies within the
    first 64 KiB of the file (its first buffer fill); complete trailing boxes are ignored either way"""
    small = streams.make(gen, "modular", 300, 100, 305, dict(tree=0, lz77=0, alpha=1, container=1))
    large = streams.make(gen, "modular", 300, 136, 305, dict(tree=0, lz77=0, alpha=1, container=1))
    assert len(small) < 65536 < len(large)
    for data in (small, large):
        for tail in (b"\x07", b"\x00\x00\x00", b"\x00\x00\x00\x10abcd1234", b"\x00\x00\x00\x10abcd12345678"):
            _cmp(oracle, emu, data + tail)
    assert oracle.decode(small + b"\x07")[1] == "shrt" and oracle.decode(large + b"\x07")[1] == ""
    # ... but a single-section frame is read through to the end of its box and then runs into the broken header
    single = streams.make(gen, "modular", 200, 300, 746, dict(tree=1, lz77=1, alpha=1, group_shift=9, container=1))
    assert len(single) > 65536
    for tail in (b"\x07", b"\x00" * 5, b"\x00\x00\x00\x10abcd1234"):
        _cmp(oracle, emu, single + tail)
    assert oracle.decode(single + b"\x07")[1] == "shrt"


def test_too_many_transforms_is_an_error_not_an_overflow(oracle, emu):
    """regression (found by the ASan/UBSan corruption sweep): an LF-group modular header announcing more than eight
    transforms used to be copied into the eight-entry list of the LF group before the error was looked at"""
    import os
    data = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bad_too_many_transforms.jxl"), "rb").read()
    a, ea, _, _ = oracle.decode(data)
    b, eb, _ = emu.decode(data)
    assert ea == eb == "xlim" and a is None and b is None


def _boxes(data):
    """splits a container file into (type, payload) after the 32 bytes of signature + ftyp"""
    out, pos = [], 32
    while pos < len(data):
        size = int.from_bytes(data[pos:pos + 4], "big")
        typ = data[pos + 4:pos + 8]
        end = len(data) if size == 0 else pos + size
        out.append((typ, data[pos + 8:end]))
        pos = end
    return out


def _box(typ, payload):
    return (8 + len(payload)).to_bytes(4, "big") + typ + payload


def test_duplicate_boxes_behind_the_codestream(oracle, emu, gen):
    """a second jxll / jxlc / jxlp box behind the last codestream box is `box?` in the reference (advisor finding,
    round 1) when the reference gets to see its header: within the first 64 KiB of the file, or behind a single-section
    frame of any size (read through to its end); a large multi-section frame is left by a seek and decodes"""
    for opts, w, h, want in ((dict(tree=0, lz77=0, container=1), 120, 90, "box?"), (dict(container=1), 600, 400, ""),
                             (dict(container=1, group_shift=9), 200, 300, "box?")):
        data = streams.make(gen, "modular", w, h, 11, opts)
        head, boxes = data[:32], _boxes(data)
        cs = [b for b in boxes if b[0] == b"jxlc"]
        assert len(cs) == 1
        jxlc = _box(b"jxlc", cs[0][1])
        jxll = _box(b"jxll", b"\x05")
        free = _box(b"free", b"\0" * 70000)
        half = len(cs[0][1]) // 2
        jxlp0 = _box(b"jxlp", b"\x80\0\0\0" + cs[0][1][:half])
        jxlp1 = _box(b"jxlp", b"\0\0\0\x01" + cs[0][1][half:])  # clear top bit: the reference's "last" (App. B quirk)
        variants = [head + jxlc + jxll + jxll, head + jxll + jxlc + jxll, head + jxlc + jxlc,
                    head + jxlp0 + jxlp1 + jxlp1, head + free + jxlc + jxll + jxll, head + jxlc + free + jxlc,
                    head + jxll + jxlc + free + jxll, head + jxlp0 + jxlp1 + jxlc, head + jxlc + jxlp1]
        for v in variants:
            assert oracle.decode(v)[1] == want
            _cmp(oracle, emu, v)
        _cmp(oracle, emu, head + jxll + jxlc)
        _cmp(oracle, emu, head + jxlp0 + jxlp1)


@pytest.mark.parametrize("opts", [dict(mix=1, tree=1), dict(mix=2, tree=2, block_ctx=1, orders=0x1f), dict(mix=1, tree=1, passes=3, smooth=0, extra_prec=1)],
                         ids=["e6like", "all_transforms_custom", "passes3_nosmooth"])
def test_intermediate_arrays_match_the_reference(oracle, emu, gen, opts):
    """block map, varblocks, LF indices, CfL maps, LLF and HF coefficients (before and after dequantisation) of the
    kernel bodies against the reference's own arrays (two LF groups: 2100 pixels wide)"""
    from tests import intermediates
    data, _ = gen.vardct(2100, 200, seed=50, hfmul=8, **opts)
    n = intermediates.check(oracle, data, lambda gg, what, out: emu.dump(data, gg, what, out))
    assert n > 100000


_LANE_CASES = [("vardct",) + c for c in streams.VARDCT_CASES] + [("modular",) + c for c in streams.MODULAR_CASES]


@pytest.mark.parametrize("case", _LANE_CASES, ids=[c[1] for c in _LANE_CASES])
def test_lane_per_stream_decoders(oracle, emu, gen, case, monkeypatch):
    """the lane-per-stream serial decoders (j40b_modlane.h: lf_decode1_lanes, lf_decode2_lanes + lf_place_body,
    modular_lanes) run by the emulator instead of the warp-per-stream ones"""
    monkeypatch.setenv("HOSTEMU_LANE", "1")
    kind, _, w, h, seed, opts = case
    _cmp(oracle, emu, streams.make(gen, kind, w, h, seed, opts))


def test_error_codes_on_corrupt_streams_lane_mode(oracle, emu, gen, monkeypatch):
    monkeypatch.setenv("HOSTEMU_LANE", "1")
    base = [streams.make(gen, "vardct", 264, 136, 3, dict(mix=1, tree=1)),
            streams.make(gen, "vardct", 64, 64, 4, dict(mix=1, tree=1, ans=0)),
            streams.make(gen, "modular", 300, 200, 5, dict())]
    for bi, data in enumerate(base):
        # (the same corruptions as test_error_codes_on_corrupt_streams: the reference itself crashes on some others)
        for name, bad in streams.corruptions(data, bi, 60):
            a, ea, _, _ = oracle.decode(bad)
            b, eb, _ = emu.decode(bad)
            assert (ea == "") == (eb == ""), (bi, name, ea, eb)
            if ea == "":
                assert np.array_equal(a, b), (bi, name)


@pytest.mark.parametrize("case", streams.WRAP_CASES, ids=[c[0] for c in streams.WRAP_CASES])
def test_out_of_range_samples_wrap_like_the_reference(oracle, emu, gen, case):
    """the reference's (int16_t) cast of the sRGB-encoded sample (j40.h:7234) wraps around for huge samples; round 1
    saturated them (the one known pixel divergence on streams both decoders accept)"""
    _, w, h, seed, opts = case
    data = streams.make(gen, "vardct", w, h, seed, opts)
    a, ea, _, _ = oracle.decode(data)
    assert ea == "" and len(np.unique(a[..., :3])) > 2
    _cmp(oracle, emu, data)


def test_local_tree_with_lz77_needs_a_window(oracle, emu):
    """regression (found by the round-2 ASan/UBSan sweep): a corrupt LF-group header naming a tree of its own whose code
    spec enables LZ77 while the global one does not -- the LZ77 window is sized from the local specs as well"""
    import os
    data = open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bad_local_tree_lz77.jxl"), "rb").read()
    a, ea, _, _ = oracle.decode(data)
    b, eb, _ = emu.decode(data)
    assert ea == eb != "" and a is None and b is None


@pytest.mark.parametrize("case", streams.VARDCT_CASES[1::4], ids=[c[0] for c in streams.VARDCT_CASES[1::4]])
def test_without_the_sharpness_split(oracle, emu, gen, case, monkeypatch):
    """multi-section frames decode the sharpness channel of the LF groups behind everything else (the emulator runs it
    last, like the side stream of the device); HOSTEMU_NO_SPLIT keeps it on the critical path (what single-section
    frames always do). Both orders must give the reference's result -- and the reference's error on corrupt input."""
    monkeypatch.setenv("HOSTEMU_NO_SPLIT", "1")
    _, w, h, seed, opts = case
    data = streams.make(gen, "vardct", w, h, seed, opts)
    _cmp(oracle, emu, data)
    for name, bad in streams.corruptions(data, 1, 12):
        a, ea, _, _ = oracle.decode(bad)
        b, eb, _ = emu.decode(bad)
        assert ea == eb, (name, ea, eb)
        if ea == "":
            assert np.array_equal(a, b), name


@pytest.mark.parametrize("case", streams.VARDCT_CASES[2::5], ids=[c[0] for c in streams.VARDCT_CASES[2::5]])
def test_trees_that_do_not_fit_the_lane_group(oracle, emu, gen, case, monkeypatch):
    """The device gives an LF group 32, 16 or 8 lanes (kern_lf.cu) and the compiled MA tree may have as many inner nodes and
    leaves; the host sizes the groups from the pruned trees (Batch::lf_tree_lanes, reported here by the emulator). A tree
    that does not fit is walked node by node instead (class MC_REST): HOSTEMU_SIMT_LANES=4 sends every tree of these streams
    that way."""
    _, w, h, seed, opts = case
    data = streams.make(gen, "vardct", w, h, seed, opts)
    _cmp(oracle, emu, data)
    need = emu.last_tree_lanes()
    assert 1 <= need <= 32
    if opts.get("tree") == 1 and not opts.get("lf_local_tree"):
        assert need == 5
    monkeypatch.setenv("HOSTEMU_SIMT_LANES", "4")
    _cmp(oracle, emu, data)


@pytest.mark.parametrize("block", [1, 32, 128])
def test_coefficient_work_list_padding(oracle, emu, gen, block, monkeypatch):
    """the coefficient kernel's work list is padded to whole blocks per image and pass (Batch::hf_per_block; padding items
    have no group): 6 groups x 3 passes and a single group, with blocks of 1 (no padding), 32 and 128 lanes"""
    monkeypatch.setenv("HOSTEMU_HF_BLOCK", str(block))
    _cmp(oracle, emu, streams.make(gen, "vardct", 520, 392, 91, dict(mix=1, tree=1, passes=3, orders=0x1f)))
    _cmp(oracle, emu, streams.make(gen, "vardct", 200, 136, 92, dict(mix=1, tree=1)))

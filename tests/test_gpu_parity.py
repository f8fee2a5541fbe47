"""GPU parity tests (run with `-m gpu` on a B200): libj40b200.so, through its C ABI, must return exactly the
oracle's RGBA bytes, strides and error codes. The oracle is the unmodified reference compiled into
oracle/_ref/libj40ref.so (built where /root/reference exists; it travels to the GPU box with the repository)."""
import numpy as np
import pytest

import j40_b200 as J
from tests import streams

pytestmark = pytest.mark.gpu


def _cmp(oracle, data):
    a, ea, ma, sa = oracle.decode(data)
    b, eb, mb, sb = J.decode(data)
    assert ea == eb, (ea, eb, ma, mb)
    if a is None:
        assert b is None
    else:
        assert sa == sb and a.shape == b.shape
        assert np.array_equal(a, b), int((a != b).sum())
    return ea


def test_gpu_present():
    assert J.gpu_available()


def test_known_answer(oracle):
    for name, (hx, _) in streams.KNOWN_ANSWER.items():
        _cmp(oracle, bytes.fromhex(hx))


@pytest.mark.parametrize("case", streams.VARDCT_CASES, ids=[c[0] for c in streams.VARDCT_CASES])
def test_vardct(oracle, gen, case):
    _, w, h, seed, opts = case
    assert _cmp(oracle, streams.make(gen, "vardct", w, h, seed, opts)) == ""


@pytest.mark.parametrize("case", streams.force_cases(), ids=[c[0] for c in streams.force_cases()])
def test_each_transform(oracle, gen, case):
    _, w, h, seed, opts = case
    assert _cmp(oracle, streams.make(gen, "vardct", w, h, seed, opts)) == ""


@pytest.mark.parametrize("case", streams.MODULAR_CASES, ids=[c[0] for c in streams.MODULAR_CASES])
def test_modular(oracle, gen, case):
    _, w, h, seed, opts = case
    assert _cmp(oracle, streams.make(gen, "modular", w, h, seed, opts)) == ""


def test_1080p_and_4k_bit_exact(oracle, gen):
    # BASELINE.json configs[1] (1920x1080, 40 groups) and configs[2] (3840x2160, 135 groups + 4 LF groups)
    for (w, h, seed) in [(1920, 1080, 7), (3840, 2160, 8)]:
        data, st = gen.vardct(w, h, seed=seed, mix=1, tree=1, hfmul=10)
        assert _cmp(oracle, data) == ""


def test_modular_2048_lossless_roundtrip(oracle, gen):
    # fjxl-shaped: prefix codes + LZ77, YCoCg; exact reconstruction of the (posterised) source is a
    # size-independent property on top of the oracle comparison
    w = h = 2048
    data, _ = gen.modular(w, h, seed=9)
    px, err, _, _ = J.decode(data)
    assert err == ""
    src = (gen.synth(w, h, 9) >> 2) << 2
    assert np.array_equal(px[..., :3], src) and (px[..., 3] == 255).all()
    _cmp(oracle, data)


def test_batch_api_matches_single_decodes(oracle, gen):
    datas = [streams.make(gen, "vardct", 520, 392, 30 + i, dict(mix=1, tree=1)) for i in range(5)]
    datas.append(streams.make(gen, "modular", 300, 200, 5, dict()))
    datas.append(b"not a jxl file")
    datas.append(datas[0][: len(datas[0]) // 2])
    b = J.Batch(0)
    for d in datas:
        b.add(d)
    b.upload()
    b.decode()
    b.wait()
    for i, d in enumerate(datas):
        a, ea, _, sa = oracle.decode(d)
        assert b.error(i) == ea, i
        if a is not None:
            assert np.array_equal(b.read_pixels(i), a), i
    # decoding the same uploaded batch again gives the same answer (state is re-initialised)
    b.decode()
    b.wait()
    a, _, _, _ = oracle.decode(datas[2])
    assert np.array_equal(b.read_pixels(2), a)
    assert b.last_decode_ms() > 0 and b.stat(2) >= 3
    b.close()


def test_error_codes_on_corrupt_streams(oracle, gen):
    base = [streams.make(gen, "vardct", 264, 136, 3, dict(mix=1, tree=1)),
            streams.make(gen, "vardct", 64, 64, 4, dict(mix=1, tree=1, ans=0)),
            streams.make(gen, "modular", 300, 200, 5, dict())]
    total = mismatches = 0
    for bi, data in enumerate(base):
        for name, bad in streams.corruptions(data, bi, 45):
            a, ea, _, _ = oracle.decode(bad)
            b, eb, _, _ = J.decode(bad)
            total += 1
            assert (ea == "") == (eb == ""), (bi, name, ea, eb)
            if ea == "":
                assert np.array_equal(a, b), (bi, name)
            elif ea != eb and (ea, eb) not in streams.ALLOWED_CODE_MISMATCHES:
                mismatches += 1
    assert mismatches == 0, (mismatches, total)


def test_batch_reuse_and_async_readback(oracle, gen):
    """j40b_batch_reset keeps the allocations; j40b_batch_read_all_async copies every image behind the kernels.
    Two batch objects in flight at once (the pipelined serving pattern bench.py's e2e figure uses)."""
    import torch
    sets = [[streams.make(gen, "vardct", 264 + 8 * k, 200, 40 + 3 * k + i, dict(mix=1, tree=1)) for i in range(3)] for k in range(4)]
    objs = [J.Batch(0), J.Batch(0)]
    outs = [torch.empty((3, 300 * 4 * 300), dtype=torch.uint8, pin_memory=True).numpy() for _ in objs]
    for rnd in range(2):
        for k, b in enumerate(objs):
            b.reset()
            for d in sets[2 * rnd + k]:
                b.add(d)
            b.upload()
            b.decode()
            b.read_all_async(outs[k])
        for k, b in enumerate(objs):
            assert b.wait() == 0
            for i, d in enumerate(sets[2 * rnd + k]):
                a, ea, _, sa = oracle.decode(d)
                w, h, s = b.info(i)
                assert (w, h, s) == (a.shape[1], a.shape[0], sa)
                got = outs[k][i][: h * s].reshape(h, s)[:, : w * 4].reshape(h, w, 4)
                assert np.array_equal(got, a), (rnd, k, i)
    for b in objs:
        b.close()


def test_token_arena_overflow_is_retried(oracle, gen, monkeypatch):
    """the per-group token arena is sized from the section's byte count; a group with more non-zero coefficients
    than that estimate reports the internal code `tokv` and the batch is decoded again with worst-case capacity.
    J40B_TEST_TOKEN_SQUEEZE shrinks the first estimate to 64 tokens so that every group takes that path."""
    monkeypatch.setenv("J40B_TEST_TOKEN_SQUEEZE", "1")
    datas = [streams.make(gen, "vardct", 520, 392, 80 + i, dict(mix=1, tree=1)) for i in range(2)]
    b = J.Batch(0)
    for d in datas:
        b.add(d)
    b.upload()
    b.decode()
    assert b.wait() == 0
    for i, d in enumerate(datas):
        want, _, _, _ = oracle.decode(d)
        assert b.error(i) == "" and np.array_equal(b.read_pixels(i), want)
    b.close()


def test_phase_control_and_stage_timeline(oracle, gen):
    """j40b_batch_after orders one batch behind a stage of another; j40b_batch_event_ms reports when the stages of
    a decode ended relative to a mark (the diagnostics DESIGN.md's pipelining analysis is based on)"""
    datas = [streams.make(gen, "vardct", 520, 392, 70 + i, dict(mix=1, tree=1)) for i in range(3)]
    a, b = J.Batch(0), J.Batch(0)
    for d in datas:
        a.add(d)
        b.add(d)
    a.upload(); b.upload()
    a.mark(0)
    a.decode()
    assert b.after(a, 2) == 0          # b starts only when a's decode has finished completely
    b.decode()
    a.join(b)
    a.mark(1)
    assert a.wait() == 0 and b.wait() == 0
    total = a.mark_ms()
    ev_a = [a.event_ms(a, i) for i in (0, 5, 6, 1, 2, 3, 4)]
    ev_b = [b.event_ms(a, i) for i in (0, 5, 6, 1, 2, 3, 4)]
    assert all(x >= 0 for x in ev_a + ev_b) and ev_a == sorted(ev_a) and ev_b == sorted(ev_b)
    assert ev_b[1] >= ev_a[-1] - 1e-3 and ev_b[-1] <= total + 1e-3   # b's first kernel ended after a's last
    km = a.kernel_ms()
    assert abs(km["lf_group"] - (km["lf_image"] + km["lf_hfmeta"] + km["lf_llf"])) < 0.05
    for i, d in enumerate(datas):
        want, _, _, _ = oracle.decode(d)
        assert np.array_equal(a.read_pixels(i), want) and np.array_equal(b.read_pixels(i), want)
    a.close(); b.close()


def test_sharded_decode_two_ranks_one_gpu(oracle, gen):
    """the sharding helper with the real GPU decoder: two 'ranks' (sequential here) cover the list once"""
    import hashlib
    from j40_b200.sharding import decode_shard, gpu_decode_batch
    datas = [streams.make(gen, "vardct", 136, 72 + 8 * i, 60 + i, dict(mix=1)) for i in range(5)]
    st = sorted(decode_shard(datas, 0, 2, gpu_decode_batch(0)) + decode_shard(datas, 1, 2, gpu_decode_batch(0)))
    assert [i for i, _, _ in st] == list(range(5))
    for i, err, digest in st:
        a, ea, _, _ = oracle.decode(datas[i])
        assert err == ea == "" and digest == hashlib.sha256(a.tobytes()).hexdigest()


def test_reference_dj40_unchanged_writes_the_oracle_pixels(oracle, gen, tmp_path):
    """the reference's own CLI, built unchanged against include/j40.h (oracle/Makefile), on the GPU"""
    import os, subprocess
    from PIL import Image
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "dj40_b200")
    if not os.path.exists(exe):
        pytest.skip("dj40_b200 not built")
    data = streams.make(gen, "vardct", 520, 392, 77, dict(mix=1, tree=1))
    (tmp_path / "a.jxl").write_bytes(data)
    r = subprocess.run([exe, str(tmp_path / "a.jxl"), str(tmp_path / "a.png")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "520x392 frame read." in r.stderr
    a, ea, _, _ = oracle.decode(data)
    got = np.array(Image.open(tmp_path / "a.png").convert("RGBA"))
    assert np.array_equal(got, a)


@pytest.mark.parametrize("opts", [dict(mix=1, tree=1), dict(mix=2, tree=2, block_ctx=1, orders=0x1f), dict(mix=1, tree=1, passes=3, smooth=0, extra_prec=1)],
                         ids=["e6like", "all_transforms_custom", "passes3_nosmooth"])
def test_intermediate_arrays_match_the_reference(oracle, gen, opts):
    """north_star: float intermediates within 1e-5 of the reference. j40b_batch_debug_dump returns the device's block
    map, varblock records, LF indices, CfL maps, LLF coefficients and HF coefficients (as decoded and dequantised);
    oracle.ref.Staged replays the reference up to the same points. Expected and asserted: bit-equal."""
    from tests import intermediates
    data, _ = gen.vardct(2100, 200, seed=50, hfmul=8, **opts)
    b = J.Batch(0)
    b.add(data)
    b.upload()
    b.decode()
    assert b.wait() == 0
    n = intermediates.check(oracle, data, lambda gg, what, out: b.debug_dump(0, gg, what, out))
    assert n > 100000
    b.close()


def test_c4_8192_modular_bit_exact(oracle, gen):
    """BASELINE.json configs[3]: 8192x8192 lossless modular (fjxl-shaped: prefix codes + LZ77, YCoCg, 1024 groups):
    bit-exact against the reference and exact reconstruction of the source"""
    import numpy as np
    w = h = 8192
    src = np.tile((gen.synth(2048, 2048, 9) >> 2) << 2, (4, 4, 1))
    data, _ = gen.modular(w, h, seed=9, rgb=src)
    px, err, _, _ = J.decode(data)
    assert err == ""
    assert np.array_equal(px[..., :3], src) and (px[..., 3] == 255).all()
    a, ea, _, sa = oracle.decode(data)
    assert ea == "" and np.array_equal(a, px)


def test_tall_narrow_modular_frame(oracle, gen):
    """more than 65535 rows (the render kernel's rows used to sit on grid.y; advisor finding, round 1)"""
    data, _ = gen.modular(24, 70000, seed=3, tree=0, lz77=0)
    _cmp(oracle, data)


@pytest.mark.parametrize("case", [c for c in streams.VARDCT_CASES if c[0].startswith("passes")], ids=lambda c: c[0])
def test_multi_pass_through_the_batch_api(oracle, gen, case):
    """several images with different pass counts in one batch"""
    _, w, h, seed, opts = case
    datas = [streams.make(gen, "vardct", w, h, seed, opts), streams.make(gen, "vardct", 264, 200, seed, dict(mix=1, tree=1))]
    b = J.Batch(0)
    for d in datas:
        b.add(d)
    b.upload()
    b.decode()
    assert b.wait() == 0
    for i, d in enumerate(datas):
        a, ea, _, _ = oracle.decode(d)
        assert np.array_equal(b.read_pixels(i), a)
    b.close()


_LANE_CASES = [("vardct",) + c for c in streams.VARDCT_CASES[::2]] + [("modular",) + c for c in streams.MODULAR_CASES[::2]]


@pytest.mark.parametrize("case", _LANE_CASES, ids=[c[1] for c in _LANE_CASES])
def test_lane_per_stream_decoders(oracle, gen, case, monkeypatch):
    """the lane-per-stream serial decoders (k_lf_lane, k_lf_place, k_mod_lane; j40b_modlane.h; opt-in: J40B_LF_MODE=lane)
    on small images"""
    monkeypatch.setenv("J40B_LF_MODE", "lane")
    kind, _, w, h, seed, opts = case
    _cmp(oracle, streams.make(gen, kind, w, h, seed, opts))


@pytest.mark.parametrize("lanes", [8, 16])
def test_lf_groups_sharing_a_warp(oracle, gen, lanes, monkeypatch):
    """2 or 4 LF groups per warp in the serial LF kernels (kern_lf.cu, G = 16 / 8 lanes per group; J40B_LF_LANES), on a small
    batch: frames of different sizes, coding tools and
    decoder classes side by side, so that the groups of a warp disagree on geometry and control flow; a corrupt stream
    among them (its group leaves early)"""
    monkeypatch.setenv("J40B_LF_LANES", str(lanes))
    opts = [dict(mix=1, tree=1), dict(mix=1, tree=1, ans=0), dict(mix=2, tree=1), dict(mix=1, tree=1, lz77=1),
            dict(mix=1, tree=1, alpha=1), dict(mix=0, tree=1, hfmul=4), dict(mix=1, tree=1, block_ctx=1, orders=0x1f)]
    datas = [streams.make(gen, "vardct", 136 + 72 * (i % 5), 72 + 56 * (i % 4), 300 + i, opts[i % len(opts)]) for i in range(21)]
    datas.append(streams.make(gen, "vardct", 2100, 300, 6, dict(mix=1, tree=1, hfmul=6)))  # two LF groups
    datas.insert(2, datas[3][: len(datas[3]) * 2 // 3])  # truncated inside the LF group's section or a pass group's
    b = J.Batch(0)
    b.add_many(datas)
    b.upload()
    b.decode()
    b.wait()
    assert b.stat(5) == lanes, "the trees of these streams were expected to fit the lane groups"
    for i, d in enumerate(datas):
        a, ea, _, _ = oracle.decode(d)
        assert b.error(i) == ea, (i, b.error(i), ea)
        if ea == "":
            assert np.array_equal(b.read_pixels(i), a), i
    b.close()


def test_lane_mode_large_batch_matches(oracle, gen, monkeypatch):
    """80 small VarDCT frames + a modular frame of 72 groups in one batch through the lane kernels (full warps, lanes of
    different geometry side by side)"""
    monkeypatch.setenv("J40B_LF_MODE", "lane")
    datas = [streams.make(gen, "vardct", 136 + 8 * (i % 5), 72 + 8 * (i % 3), 200 + i, dict(mix=1, tree=1)) for i in range(80)]
    datas.append(streams.make(gen, "modular", 2300, 2000, 3, dict(tree=2, ans=1, lz77=0)))  # 72 groups
    b = J.Batch(0)
    b.add_many(datas)
    b.upload()
    b.decode()
    assert b.wait() == 0
    for i in list(range(0, 80, 7)) + [80]:
        a, ea, _, _ = oracle.decode(datas[i])
        assert ea == "" and np.array_equal(b.read_pixels(i), a), i
    b.close()


def test_error_codes_on_corrupt_streams_lane_mode(oracle, gen, monkeypatch):
    monkeypatch.setenv("J40B_LF_MODE", "lane")
    base = [streams.make(gen, "vardct", 264, 136, 3, dict(mix=1, tree=1)),
            streams.make(gen, "vardct", 64, 64, 4, dict(mix=1, tree=1, ans=0)),
            streams.make(gen, "modular", 300, 200, 5, dict())]
    for bi, data in enumerate(base):
        # (the same corruptions as test_error_codes_on_corrupt_streams: the reference itself crashes on some others)
        for name, bad in streams.corruptions(data, bi, 45):
            a, ea, _, _ = oracle.decode(bad)
            b, eb, _, _ = J.decode(bad)
            assert (ea == "") == (eb == ""), (bi, name, ea, eb)
            if ea == "":
                assert np.array_equal(a, b), (bi, name)


@pytest.mark.parametrize("case", streams.WRAP_CASES, ids=[c[0] for c in streams.WRAP_CASES])
def test_out_of_range_samples_wrap_like_the_reference(oracle, gen, case):
    """the reference's (int16_t) cast (j40.h:7234) wraps huge samples around. On the device the host libm's powf is a
    double-precision pow rounded to float (srgb_u8_wrapped): a last-bit difference can move the encoded value across
    an integer only once in several hundred float steps, so at most a couple of pixels of these ~10^4 may differ."""
    _, w, h, seed, opts = case
    data = streams.make(gen, "vardct", w, h, seed, opts)
    a, ea, _, sa = oracle.decode(data)
    b, eb, _, sb = J.decode(data)
    assert ea == eb == "" and sa == sb
    assert int((a != b).any(axis=-1).sum()) <= 2


def test_pam_writer_consumes_device_output(oracle, gen, tmp_path):
    """j40b_batch_write_pam: device-resident pixels to a PAM file without the stride padding"""
    data = streams.make(gen, "vardct", 264, 136, 90, dict(mix=1, tree=1))
    b = J.Batch(0)
    b.add(data)
    b.upload()
    b.decode()
    assert b.wait() == 0
    path = tmp_path / "a.pam"
    assert b.write_pam(0, str(path)) == 0
    raw = path.read_bytes()
    head, body = raw.split(b"ENDHDR\n", 1)
    assert head == b"P7\nWIDTH 264\nHEIGHT 136\nDEPTH 4\nMAXVAL 255\nTUPLTYPE RGB_ALPHA\n"
    a, _, _, _ = oracle.decode(data)
    assert body == a.tobytes()
    b.close()

"""Pins the oracle (the unmodified reference, compiled by oracle/Makefile) with the known-answer vectors of
SURVEY.md App. C, and pins the product's host-side tables and float kernels against it function by function."""
import numpy as np
import pytest

from tests.streams import KNOWN_ANSWER


def test_known_answer_vectors(oracle):
    for name, (hx, colour) in KNOWN_ANSWER.items():
        px, err, msg, stride = oracle.decode(bytes.fromhex(hx))
        assert err == "", (name, msg)
        if colour is not None:
            assert (px.reshape(-1, 4) == np.array(colour, np.uint8)).all(), name
    px, _, _, stride = oracle.decode(bytes.fromhex(KNOWN_ANSWER["V3_vardct_264x8_5sections"][0]))
    assert px.shape == (8, 264, 4) and stride == 1088
    # 8-pixel vertical stripes 144 / 0
    assert (px[:, 0:8, :3] == 144).all() and (px[:, 8:16, :3] == 0).all() and (px[..., 3] == 255).all()


def test_spec_ordered_single_section_is_rejected(oracle):
    # the reference reads LfGlobal, HfGlobal, LfGroup, PassGroup in that order
    v2 = bytes.fromhex(KNOWN_ANSWER["V2_vardct_8x8"][0])
    assert oracle.decode(v2)[1] == ""


def test_dq_matrices_match_reference(oracle, emu):
    for idx in range(17):
        want = oracle.default_dq_matrix(idx)
        buf = np.zeros(want.size, np.float32)
        n = emu.lib().hostemu_dq_matrix(idx, buf.ctypes.data, buf.size)
        assert n == want.shape[0]
        assert np.array_equal(buf.view(np.uint32), want.reshape(-1).view(np.uint32)), idx


def test_natural_orders_match_reference(oracle, emu):
    for lr, lc in [(3, 3), (4, 4), (5, 5), (3, 4), (3, 5), (4, 5), (6, 6), (5, 6), (7, 7), (6, 7), (8, 8), (7, 8)]:
        want = oracle.natural_order(lr, lc)
        got = np.zeros(want.size, np.int32)
        assert emu.lib().hostemu_natural_order(lr, lc, got.ctypes.data) == want.size
        assert np.array_equal(got, want), (lr, lc)


def test_srgb_threshold_table_matches_reference_quantiser(oracle, emu):
    thr = np.zeros(255, np.float32)
    emu.lib().hostemu_srgb_thresholds(8, thr.ctypes.data)
    assert (np.diff(thr) > 0).all()
    rng = np.random.default_rng(0)
    # dense random floats over the reachable range plus values straddling every threshold
    vals = np.concatenate([
        rng.uniform(-0.5, 1.5, 20000).astype(np.float32),
        np.exp(rng.uniform(-20, 2, 20000)).astype(np.float32),
        np.nextafter(thr, np.float32(-np.inf)), thr, np.nextafter(thr, np.float32(np.inf)),
    ])
    L = emu.lib()
    # bucket edges of the tile kernel's start table (j40b_vardct.h srgb_u8_lut) and their neighbours
    edges = (np.arange(0, 4097, dtype=np.float32) / np.float32(4096.0)).astype(np.float32)
    vals = np.concatenate([vals, edges, np.nextafter(edges, np.float32(-np.inf)), np.nextafter(edges, np.float32(np.inf)),
                           np.array([np.nan, np.inf, -np.inf, 0.0, -0.0, 1.0, 2.0, 1e30], np.float32)]).astype(np.float32)
    lut_out = np.zeros(vals.size, np.uint8)
    L.hostemu_srgb_lut_lookup(vals.ctypes.data, int(vals.size), lut_out.ctypes.data)
    for k, v in enumerate(vals):
        if np.isfinite(v) and v < 100.0:
            want = min(max(int(oracle.lib().ref_srgb_quant(float(v), 8)), 0), 255)
        else:
            want = 255 if v > 0 else 0   # saturating device behaviour; the reference's int16 cast is undefined here
        got = L.hostemu_srgb_lookup(thr.ctypes.data, float(v))
        assert got == want, (float(v), got, want)
        assert int(lut_out[k]) == want, (float(v), int(lut_out[k]), want)


@pytest.mark.parametrize("dctsel", list(range(27)))
def test_inverse_transforms_bit_exact(oracle, emu, dctsel):
    dims = [(3, 3)] * 4 + [(4, 4), (5, 5), (4, 3), (3, 4), (5, 3), (3, 5), (5, 4), (4, 5)] + [(3, 3)] * 6 + \
           [(6, 6), (6, 5), (5, 6), (7, 7), (7, 6), (6, 7), (8, 8), (8, 7), (7, 8)]
    lr, lc = dims[dctsel]
    rng = np.random.default_rng(dctsel)
    for trial in range(3):
        coef = (rng.standard_normal(1 << (lr + lc)) * (10.0 if trial else 0.01)).astype(np.float32)
        if trial == 2:
            coef[rng.random(coef.size) < 0.8] = 0
        want = oracle.inverse_transform(dctsel, coef)
        got = coef.copy()
        emu.lib().hostemu_inverse_transform(dctsel, got.ctypes.data)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (dctsel, trial)


def test_forward_llf_bit_exact(oracle, emu):
    rng = np.random.default_rng(7)
    for lr in range(0, 6):
        for lc in range(0, 6):
            if lr == 0 and lc == 0:
                continue
            blk = rng.standard_normal(1 << (lr + lc)).astype(np.float32)
            want = oracle.forward_llf(blk, lr, lc)
            got = blk.copy()
            emu.lib().hostemu_forward_llf(got.ctypes.data, lr, lc)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (lr, lc)

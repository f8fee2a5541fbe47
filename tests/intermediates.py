"""Shared by the CPU (emulator) and GPU tests: the decoder's intermediate arrays against the reference's own
(oracle.ref.Staged = a step-by-step replay of j40__advance): block map, varblock records, LF indices, chroma-from-luma
and sharpness maps bit for bit; LLF coefficients, coefficients as decoded and after j40__dequant_hf within 1e-5
relative (north_star: "float intermediates match within 1e-5"; they are expected -- and checked -- to be bit-equal)."""
import numpy as np

TOL = 1e-5
# name in oracle.ref.Staged.WHAT -> (`what` of j40b_batch_debug_dump, stage of the reference replay)
ARRAYS = [("blocks", 0, 0), ("varblocks", 1, 0), ("lfindices", 2, 0), ("llf_x", 3, 0), ("llf_y", 4, 0), ("llf_b", 5, 0),
          ("coef_x", 6, 0), ("coef_y", 7, 0), ("coef_b", 8, 0), ("xfromy", 9, 0), ("bfromy", 10, 0), ("sharpness", 11, 0),
          ("coef_x", 12, 1), ("coef_y", 13, 1), ("coef_b", 14, 1)]


def check(oracle, data, dump):
    """dump(lf_group, what, out_array) -> bytes written. Returns the number of float values compared."""
    st = oracle.Staged(data)
    assert st.err == "", st.err
    compared = 0
    try:
        for stage in (0, 1):
            assert st.advance(stage) == ""
            for gg in range(st.info["num_lf_groups"]):
                for name, what, at in ARRAYS:
                    if at != stage:
                        continue
                    want = st.lf_group_array(gg, name)
                    got = np.zeros_like(want)
                    n = dump(gg, what, got)
                    assert n == want.nbytes, (name, what, n, want.nbytes)
                    if want.dtype == np.float32:
                        scale = np.maximum(np.abs(want), 1e-30)
                        bad = np.abs(got - want) > TOL * scale
                        assert not bad.any(), (name, what, gg, int(bad.sum()))
                        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (name, what, gg, "within 1e-5 but not bit-equal")
                        compared += want.size
                    else:
                        assert np.array_equal(got, want), (name, what, gg)
    finally:
        st.close()
    return compared

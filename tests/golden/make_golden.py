#!/usr/bin/env python
"""Generates the committed golden fixtures: small JPEG XL codestreams from tools/streamgen together with the
RGBA8 output of the UNMODIFIED reference (oracle/_ref, built from /root/reference/j40.h in this container).

    python tests/golden/make_golden.py        # rewrites tests/golden/*.jxl and golden.json

The reference ships no test vectors of its own; these pin the oracle's output across rebuilds
and travel to the GPU box, where /root/reference does not exist. Pixels are stored as a SHA-256 of the tight
h*w*4 RGBA buffer plus the first and last pixel rows (enough to localise a mismatch)."""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref          # noqa: E402
from tools import streamgen     # noqa: E402
from tests.streams import KNOWN_ANSWER  # noqa: E402

CASES = {
    "vardct_dct8_64x64": lambda: streamgen.vardct(64, 64, seed=1, mix=0, tree=0),
    "vardct_mix_200x136_wp": lambda: streamgen.vardct(200, 136, seed=2, mix=1, tree=1),
    "vardct_mix_stress_tree_prefix": lambda: streamgen.vardct(136, 200, seed=3, mix=1, tree=2, ans=0),
    "vardct_multigroup_520x392": lambda: streamgen.vardct(520, 392, seed=4, mix=1, tree=1, hfmul=10, hfmul_var=4),
    "vardct_container_jxlp_lz77": lambda: streamgen.vardct(96, 80, seed=5, mix=1, container=1, jxlp=1, lz77=1),
    "vardct_permuted_orders_presets": lambda: streamgen.vardct(300, 280, seed=6, mix=1, permuted=1, orders=0x1f, presets=2, block_ctx=1),
    "vardct_alpha_extra_channel_264x264": lambda: streamgen.vardct(264, 264, seed=9, mix=1, tree=1, alpha=1),
    "vardct_raw_dq_alpha_136x120": lambda: streamgen.vardct(136, 120, seed=12, mix=1, tree=1, raw_dq=0x11, alpha=1),
    "modular_rgb_rct_300x200": lambda: streamgen.modular(300, 200, seed=7),
    "modular_local_trees_300x280": lambda: streamgen.modular(300, 280, seed=10, local_tree=1),
    "modular_palette_300x280": lambda: streamgen.modular(300, 280, seed=11, palette=1),
    "modular_alpha_wp_ans": lambda: streamgen.modular(130, 70, seed=8, alpha=1, tree=2, ans=1, lz77=0),
}


def main():
    index = {}
    for name, make in CASES.items():
        data = make()[0]
        px, err, msg, stride = ref.decode(data)
        assert err == "", (name, err, msg)
        open(os.path.join(HERE, name + ".jxl"), "wb").write(data)
        index[name] = {"bytes": len(data), "width": int(px.shape[1]), "height": int(px.shape[0]), "stride": int(stride),
                       "sha256": hashlib.sha256(px.tobytes()).hexdigest(),
                       "first_row": px[0].tobytes().hex() if px.shape[1] <= 64 else px[0, :64].tobytes().hex(),
                       "last_row": px[-1].tobytes().hex() if px.shape[1] <= 64 else px[-1, :64].tobytes().hex()}
    for name, (hexdata, _) in KNOWN_ANSWER.items():
        data = bytes.fromhex(hexdata)
        px, err, msg, stride = ref.decode(data)
        index["ka_" + name] = {"hex": hexdata, "error": err,
                               "sha256": hashlib.sha256(px.tobytes()).hexdigest() if not err else "",
                               "width": int(px.shape[1]) if not err else 0, "height": int(px.shape[0]) if not err else 0}
    # corrupted variants of one stream: the reference's error code is part of the contract
    base = CASES["vardct_mix_200x136_wp"]()[0]
    for tag, data in (("truncated_half", base[: len(base) // 2]), ("truncated_tail", base[:-3]),
                      ("bitflip_mid", base[: len(base) // 2] + bytes([base[len(base) // 2] ^ 0x10]) + base[len(base) // 2 + 1:])):
        px, err, msg, stride = ref.decode(data)
        open(os.path.join(HERE, "bad_" + tag + ".jxl"), "wb").write(data)
        index["bad_" + tag] = {"bytes": len(data), "error": err, "sha256": hashlib.sha256(px.tobytes()).hexdigest() if not err else ""}
    json.dump(index, open(os.path.join(HERE, "golden.json"), "w"), indent=1, sort_keys=True)
    print("wrote", len(index), "golden entries")


if __name__ == "__main__":
    main()

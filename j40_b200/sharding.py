"""Multi-GPU use of the decoder: images are independent (one j40__inner per handle in the reference,
j40.h:8329), so a list of images shards over ranks at image granularity with **no collective on the data
path** (SURVEY.md §8e). One process per GPU; each rank decodes its shard with its own `Batch`; results stay on
that rank's device. The only cross-rank step is an optional host-side gather of per-image status (a few bytes
per image) so that rank 0 can report which images failed.

`shard_indices` is deterministic and identical on every rank, so no rank needs to be told its share.
"""
from __future__ import annotations

import hashlib
from typing import Callable, List, Optional, Sequence, Tuple


def shard_indices(n_items: int, rank: int, world: int, sizes: Optional[Sequence[int]] = None) -> List[int]:
    """Indices of the images rank `rank` of `world` decodes.

    Without `sizes`: round-robin (image i -> rank i mod world), which keeps neighbouring frames of a sequence
    on different GPUs. With `sizes` (compressed bytes, a proxy for entropy-decode time): longest-processing-time
    greedy assignment, ties broken by index, which balances ragged batches.
    """
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    if sizes is None:
        return list(range(rank, n_items, world))
    if len(sizes) != n_items:
        raise ValueError("sizes must have one entry per item")
    load = [0] * world
    mine: List[int] = []
    for i in sorted(range(n_items), key=lambda k: (-int(sizes[k]), k)):
        r = min(range(world), key=lambda k: (load[k], k))
        load[r] += int(sizes[i])
        if r == rank:
            mine.append(i)
    return sorted(mine)


def decode_shard(datas: Sequence[bytes], rank: int, world: int, decode_batch: Callable[[List[bytes]], List[Tuple[str, bytes]]],
                 balance: bool = False) -> List[Tuple[int, str, str]]:
    """Decodes this rank's share with `decode_batch` (list of codestreams -> list of (error code, RGBA bytes))
    and returns [(global index, error code, sha256 of the pixels)] for the local images."""
    idx = shard_indices(len(datas), rank, world, [len(d) for d in datas] if balance else None)
    out = decode_batch([datas[i] for i in idx]) if idx else []
    assert len(out) == len(idx)
    return [(i, err, hashlib.sha256(px).hexdigest() if not err else "") for i, (err, px) in zip(idx, out)]


def gather_status(local: List[Tuple[int, str, str]], dist=None) -> List[Tuple[int, str, str]]:
    """Host-side gather of the per-image status on every rank (torch.distributed all_gather_object; works with
    the gloo and nccl backends). Without an initialised process group returns the local list."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return sorted(local)
    parts: List[Optional[list]] = [None] * dist.get_world_size()
    dist.all_gather_object(parts, local)
    return sorted(x for p in parts for x in p)


def gpu_decode_batch(device: int) -> Callable[[List[bytes]], List[Tuple[str, bytes]]]:
    """The product decoder as a `decode_batch` callable: one `Batch` on CUDA device `device`."""
    from . import Batch

    def run(datas: List[bytes]) -> List[Tuple[str, bytes]]:
        b = Batch(device)
        try:
            for d in datas:
                b.add(d)
            b.upload()
            b.decode()
            b.wait()
            res = []
            for i in range(len(datas)):
                e = b.error(i)
                res.append((e, b"" if e else b.read_pixels(i).tobytes()))
            return res
        finally:
            b.close()
    return run

"""Multi-GPU path: images shard over ranks with no data-path collective. The host-side logic
(j40_b200/sharding.py) is exercised here with two gloo ranks on the CPU; each rank decodes its share with the
CPU kernel-logic emulator standing in for the GPU batch decoder (the GPU variant is in test_gpu_parity.py)."""
import hashlib
import os
import socket

import pytest

from j40_b200.sharding import shard_indices, decode_shard, gather_status


def test_shard_indices_partition_round_robin_and_balanced():
    for n in (0, 1, 5, 64, 257):
        for world in (1, 2, 3, 8):
            parts = [shard_indices(n, r, world) for r in range(world)]
            assert sorted(i for p in parts for i in p) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
    sizes = [100, 1, 1, 1, 50, 50, 98, 3]
    parts = [shard_indices(len(sizes), r, 2, sizes) for r in range(2)]
    assert sorted(parts[0] + parts[1]) == list(range(len(sizes)))
    loads = [sum(sizes[i] for i in p) for p in parts]
    assert abs(loads[0] - loads[1]) <= 4
    with pytest.raises(ValueError):
        shard_indices(4, 2, 2)
    with pytest.raises(ValueError):
        shard_indices(4, 0, 2, [1, 2])


def _streams():
    from tools import streamgen
    datas = [streamgen.vardct(72 + 8 * i, 40 + 16 * (i % 3), seed=10 + i, mix=i % 2, tree=i % 3)[0] for i in range(5)]
    datas.append(streamgen.modular(50, 30, seed=3)[0])
    datas.append(datas[0][: len(datas[0]) // 2])  # a truncated stream: error status must travel too
    return datas


def _emu_batch(datas):
    from tests import hostemu
    out = []
    for d in datas:
        px, err, _ = hostemu.decode(d)
        out.append((err, b"" if err else px.tobytes()))
    return out


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        datas = _streams()
        local = decode_shard(datas, rank, world, _emu_batch, balance=(rank >= 0))
        assert [i for i, _, _ in local] == shard_indices(len(datas), rank, world, [len(d) for d in datas])
        allst = gather_status(local, dist)
        q.put((rank, [i for i, _, _ in local], allst))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharded_decode_matches_oracle(oracle):
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    datas = _streams()
    want = []
    for i, d in enumerate(datas):
        px, err, _, _ = oracle.decode(d)
        want.append((i, err, "" if err else hashlib.sha256(px.tobytes()).hexdigest()))
    assert any(e for _, e, _ in want) and any(not e for _, e, _ in want)
    locals_ = sorted(i for _, idx, _ in res for i in idx)
    assert locals_ == list(range(len(datas)))           # every image decoded exactly once
    for _, _, allst in res:
        assert allst == want                             # every rank sees the full, oracle-identical status

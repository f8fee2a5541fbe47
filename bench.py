#!/usr/bin/env python
"""bench.py -- throughput of the JPEG XL group-decode hot path on B200 (BASELINE.json metric: Mpixels/s decoded,
4K VarDCT batch), with roofline, CPU baseline and end-to-end numbers.

  python bench.py --gpus 1 --steps 5 --warmup 3                    # our CUDA path
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU, batch-sharded
  python bench.py --impl reference --gpus 1 --steps 2 --warmup 1   # the reference j40.h on the host CPU cores

A "step" is one decode of this rank's batch of synthetic 4K VarDCT frames ("d1/e6-like" streams from
tools/streamgen; no JPEG XL encoder exists offline). Frames are independent, so ranks share nothing: no
collective on the data path (weak scaling: the per-GPU batch is fixed).
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# a dozen batch objects (CUDA streams) are in flight at once: with the default of 8 hardware work queues, streams
# that share a queue serialise behind each other (measured: +4 % device-resident, +15 % end to end with 32)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

STREAM_OPTS = dict(mix=1, tree=1, hfmul=10, hfmul_var=4)  # the "d1/e6-like" preset


def workload_name(w, h, frames, distinct):
    return (f"batch of {frames} independent {w}x{h} VarDCT frames per GPU ({distinct} distinct, cycled), "
            f"tools/streamgen d1/e6-like preset {STREAM_OPTS}")


def host_memory_budget():
    """bytes of host memory available to this process tree: MemAvailable, capped by the cgroup limit if there is one"""
    avail = 1 << 62
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                avail = int(line.split()[1]) * 1024
    except Exception:
        pass
    for path in ("/sys/fs/cgroup/memory.max", "/sys/fs/cgroup/memory/memory.limit_in_bytes"):
        try:
            v = open(path).read().strip()
            if v.isdigit():
                avail = min(avail, int(v))
        except Exception:
            pass
    return avail


def _gen_one(args):
    w, h, seed = args
    from tools import streamgen
    data, st = streamgen.vardct(w, h, seed=seed, **STREAM_OPTS)
    return data, st


def make_streams(w, h, seeds):
    from tools import streamgen
    streamgen._ensure_tables()  # oracle-derived tables, inherited by forked workers
    procs = max(1, min(len(seeds), (os.cpu_count() or 2) - 1, 32))
    if procs == 1:
        return [_gen_one((w, h, s)) for s in seeds]
    with mp.get_context("fork").Pool(procs) as pool:
        return pool.map(_gen_one, [(w, h, s) for s in seeds])


class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.sm_max = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.sm_max = float(out[1])
                for n, v in zip(names, out[2:]):
                    if "Active" in v and "Not" not in v:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.1)

    def result(self):
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def _ref_worker(args):
    data, reps = args
    from oracle import ref
    good, secs = ref.time_decode(data, reps)
    return good, secs


def run_reference(args, rank, world):
    """The reference's own CPU implementation (oracle/_ref = unmodified j40.h) on all host cores."""
    if rank != 0:
        return
    w, h = args.width, args.height
    cores = os.cpu_count() or 1
    nproc = max(1, min(cores, 64))
    datas = [d for d, _ in make_streams(w, h, list(range(min(args.distinct, nproc))))]
    work = [(datas[i % len(datas)], 1) for i in range(nproc)]
    ctx = mp.get_context("fork")
    with ctx.Pool(nproc) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_worker, work[: max(1, nproc // 4)])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            res = pool.map(_ref_worker, work)
            assert all(g == 1 for g, _ in res), "reference failed to decode a bench stream"
        dt = time.perf_counter() - t0
    mpix = nproc * args.steps * w * h / dt / 1e6
    line = {
        "impl": "reference", "metric": "Mpixels/s decoded (4K VarDCT batch)", "value": mpix, "unit": "Mpix/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(w, h, args.frames_per_gpu, args.distinct),
                   "implementation": "reference j40.h (unmodified, -O3 -ffp-contract=off) on the host CPU, one process per frame",
                   "frames_per_step": nproc, "processes": nproc,
                   "sample": f"each step decodes {nproc} frames of the workload (one per process) instead of all {args.frames_per_gpu}"},
        "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": nproc, "kind": "reference",
                         "sample": f"{nproc} frames per step, one process per frame, {args.steps} steps"},
        "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--frames-per-gpu", type=int, default=64)
    ap.add_argument("--distinct", type=int, default=8, help="distinct streams per rank (cycled to fill the batch)")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--streams", type=int, default=12, help="batch objects (CUDA streams) the timed steps are pipelined over")
    ap.add_argument("--e2e-sets", type=int, default=2, help="groups of `streams` batch objects the end-to-end loop alternates between")
    ap.add_argument("--lag", type=int, default=0, help="step s starts its LF stage when step s-lag has finished its own "
                    "(keeps the batches in flight out of phase); 0 = no phase control (default: measured best), -1 = streams/2")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import j40_b200 as J
    from oracle import ref

    if not J.gpu_available():
        print(json.dumps({"error": "no CUDA device; j40_b200 has no CPU decoding path"}))
        sys.exit(1)
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # keep NCCL's version banner off stdout: rank 0 prints one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    w, h = args.width, args.height
    F = args.frames_per_gpu
    gen = make_streams(w, h, [rank * args.distinct + i for i in range(args.distinct)])
    datas = [g[0] for g in gen]
    stats = [g[1] for g in gen]
    frames = [datas[i % len(datas)] for i in range(F)]
    comp_bytes = sum(len(d) for d in frames)
    pixels = F * w * h

    # ---- parity gate: no number counts unless the GPU output equals the reference's, byte for byte
    want, e0, _, _ = ref.decode(datas[0])
    assert e0 == "", "oracle rejected a bench stream"

    # ---- device-resident throughput: inputs + tables already in HBM, kernels only
    b = J.Batch(local_rank)
    for d in frames:
        b.add(d)
    b.upload()
    b.decode()
    failed = b.wait()
    assert failed == 0, [b.error(i) for i in range(F) if b.error(i)]
    assert np.array_equal(b.read_pixels(0), want), "GPU output differs from the reference"
    # serial pass (one batch, steps back to back): per-kernel device times for the roofline section
    step_ms, kernel_ms = [], []
    for it in range(args.warmup + 2):
        b.decode()
        b.wait()
        if it >= args.warmup:
            step_ms.append(b.last_decode_ms())
            kernel_ms.append(b.kernel_ms())
    serial_ms = sum(step_ms) / len(step_ms)
    dev_bytes = b.stat(0)
    launches_per_step = b.stat(2)
    # timed region: K steps pipelined over M batch objects (one CUDA stream each), so that the latency-bound
    # LF-group kernel of one step overlaps the HF / back kernels of its neighbours. All K steps run inside the
    # region; device time is taken with CUDA events on stream 0 after joining every stream.
    M = max(1, min(args.streams, args.steps))
    lag = M // 2 if args.lag < 0 else min(args.lag, M - 1)

    def submit_decode(s_):
        # phase control: without it all batches in flight run their LF stages together, then their HF stages ...
        if lag and s_ >= lag:
            batches[s_ % M].after(batches[(s_ - lag) % M], 0)
        batches[s_ % M].decode()

    batches = [b]
    for m in range(1, M):
        bm = J.Batch(local_rank)
        for i in range(F):
            bm.add(frames[(i + m) % F])
        bm.upload()
        batches.append(bm)
    for _ in range(args.warmup):
        for k in range(M):
            submit_decode(k)
        for bm in batches:
            assert bm.wait() == 0
    sampler = ClockSampler(local_rank)
    sampler.start()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    b.mark(0)
    for s_ in range(args.steps):
        submit_decode(s_)
    for bm in batches[1:]:
        b.join(bm)
    b.mark(1)
    for bm in batches:
        assert bm.wait() == 0
    torch.cuda.synchronize()
    total_ms = b.mark_ms()
    if os.environ.get("J40B_TIMELINE"):
        # when each stage of the last decode on every batch object ended, relative to the start of the region
        for k, bm in enumerate(batches):
            print("timeline batch %d: " % k + " ".join("%s=%.1f" % (n, bm.event_ms(b, i)) for n, i in
                  [("lf0", 0), ("lfimg", 5), ("hfmeta", 6), ("lf", 1), ("hf", 2), ("tiles", 3), ("end", 4)]), file=sys.stderr)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches = launches_per_step * args.steps
    for bm in batches[1:]:
        dev_bytes += bm.stat(0)
    t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
    if dist:
        dist.barrier()
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * pixels * args.steps / (total_ms_max / 1e3) / 1e6

    # ---- end to end through the C ABI with HOST buffers, every step: host parse of all frames + one H2D of
    # codestreams and tables + kernels + D2H of every decoded frame into pinned host memory. The batch objects
    # (device + pinned staging allocations) are reused across steps, as a serving loop would, and steps are
    # pipelined over them so that the parse / H2D / D2H of one step overlap the kernels of its neighbours.
    e2e = None
    if not args.skip_e2e:
        # Serving loop: `sets` groups of W batch objects. A group is submitted as one wave (all W decodes back to
        # back) and the next wave of the same group only after all of its D2H copies have landed; while one group
        # copies out and is re-parsed / re-uploaded, the other one computes. Waves matter: resubmitting each
        # object as soon as its own copy is done (the obvious rolling scheme) spreads the batches evenly over all
        # phases, and a phase-staggered mix of LF / HF / tile kernels runs ~40 % slower than waves.
        W = len(batches)
        sets = max(1, args.e2e_sets)
        # every object in flight owns a pinned destination for its frames (2.1 GB at 64 x 4K): stay well inside the
        # host memory this rank can count on (all ranks of the node allocate the same)
        per_obj = F * h * b.info(0)[2]
        budget = host_memory_budget() // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
        if sets * W * per_obj > 0.5 * budget:
            sets = 1
        if W * per_obj > 0.5 * budget:
            W = max(2, int(0.5 * budget / per_obj))
        if sets > 1 and (sets - 1) * W * b.stat(0) > 0.8 * torch.cuda.mem_get_info()[0]:
            sets = 1  # not enough free HBM for a second group of batch objects
        objs = list(batches[:W])
        for m in range(W, sets * W):
            bm = J.Batch(local_rank)
            objs.append(bm)
        E = len(objs)
        pitch = h * b.info(0)[2]
        host_out = [torch.empty((F, pitch), dtype=torch.uint8, pin_memory=True) for _ in range(E)]
        out_np = [t_.numpy() for t_ in host_out]

        host_t = [0.0] * 6
        skip = os.environ.get("J40B_E2E_SKIP", "")  # diagnostics only ("d2h"): such a run is not an e2e number

        def submit(k):
            bm = objs[k]
            t = [time.perf_counter()]
            bm.reset(); t.append(time.perf_counter())
            bm.add_many(frames); t.append(time.perf_counter())
            bm.upload(); t.append(time.perf_counter())
            bm.decode(); t.append(time.perf_counter())
            if skip != "d2h":
                bm.read_all_async(out_np[k])
            t.append(time.perf_counter())
            for i in range(5):
                host_t[i] += t[i + 1] - t[i]

        for k in range(E):          # warm-up (also pages the pinned buffers in)
            submit(k)
        for k in range(E):
            assert objs[k].wait() == 0
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        n_waves = max(3 * sets, -(-args.steps // W))
        n_e2e = n_waves * W
        host_t[:] = [0.0] * 6
        t0 = time.perf_counter()
        for wv in range(n_waves):
            ids = range((wv % sets) * W, (wv % sets) * W + W)
            if wv >= sets:
                tw = time.perf_counter()
                for k in ids:
                    assert objs[k].wait() == 0      # the group's previous wave, including its D2H
                host_t[5] += time.perf_counter() - tw
            for k in ids:
                submit(k)
        for k in range(E):
            assert objs[k].wait() == 0
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if os.environ.get("J40B_TIMELINE"):
            print("e2e host seconds (whole run): reset %.3f add_many %.3f upload %.3f decode %.3f read_async %.3f wait %.3f of %.3f"
                  % (*host_t, dt), file=sys.stderr)
        h2d = batches[0].stat(1)
        te = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if dist:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        stride = b.info(0)[2]
        got = out_np[0][0].reshape(h, stride)[:, : w * 4].reshape(h, w, 4)
        assert np.array_equal(got, want), "end-to-end output differs from the reference"
        e2e = {"value": world * pixels * n_e2e / float(te.item()) / 1e6, "unit": "Mpix/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(F * pitch), "steps": n_e2e,
               "includes": "host parse + H2D + kernels + D2H of all frames into pinned host memory, "
                           f"{n_waves} waves of {W} steps over {sets} groups of {W} reused batch objects"}
        for bm in objs[W:]:
            bm.close()
    for bm in batches[1:]:
        bm.close()
    b.close()

    if rank != 0:
        if dist:
            dist.destroy_process_group()
        return

    # ---- roofline (HBM): algorithmic bytes = compressed bytes read once + RGBA8 written once. Every
    # kernel of a step works on the whole batch, so the bytes "one launch processes" are the step's bytes; the
    # dominant kernel is the longest single kernel of the serial pass (CUDA events on the batch's stream, each
    # kernel alone on the GPU). `step_*` is the same figure for the whole pipelined step of the timed region.
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg_bytes = comp_bytes + 4 * pixels
    ms_step = total_ms / args.steps
    kavg = {k: sum(x[k] for x in kernel_ms) / len(kernel_ms) for k in kernel_ms[0]}
    names = {"lf_image": "k_lf_decode<1>", "lf_hfmeta": "k_lf_post+k_lf_decode<2>", "lf_llf": "k_lf_llf", "hf_group": "k_hf_group",
             "back": "k_back_tile", "back_big": "k_back_generic", "modular": "k_modular", "render": "k_render"}
    dominant = max(names, key=lambda k: kavg.get(k, 0.0))
    traffic = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        ent = tr.get(names[dominant])
        if ent:
            traffic = ent["dram_bytes_per_frame"] * F
    except Exception:
        pass
    achieved = alg_bytes / (kavg[dominant] / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": names[dominant], "kernel_ms": kavg[dominant],
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "the path is bound by the serial entropy decoders' dependent-issue latency, not by HBM (SURVEY 8d)",
                "all_kernel_ms": kavg, "serial_ms_per_step": serial_ms,
                "step_achieved": alg_bytes / (ms_step / 1e3) / 1e9, "step_frac": alg_bytes / (ms_step / 1e3) / 1e9 / peak,
                "back_tile_achieved": alg_bytes / (kavg["back"] / 1e3) / 1e9 if kavg.get("back") else None}

    # ---- CPU baseline: the reference itself, one thread, bounded sample of the same workload
    reps = 6
    good, secs = ref.time_decode(datas[0], reps)
    cpu = {"value": w * h / statistics.median(secs) / 1e6, "unit": "Mpix/s", "cores": 1, "kind": "reference",
           "sample": f"{reps} decodes of one {w}x{h} bench frame through j40_from_memory..j40_frame_pixels_u8x4, median",
           "host_cores_available": os.cpu_count()}

    line = {
        "metric": "Mpixels/s decoded (4K VarDCT batch)", "value": value, "unit": "Mpix/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(w, h, F, args.distinct),
                   "frames_per_gpu": F, "groups_per_gpu": F * ((w + 255) // 256) * ((h + 255) // 256),
                   "compressed_bytes_per_gpu": comp_bytes, "bits_per_pixel": 8.0 * comp_bytes / pixels,
                   "hf_symbols_per_pixel": sum(s["hf_symbols"] for s in stats) / (len(stats) * w * h),
                   "l2": "inputs+outputs per step (%.1f GB) exceed L2" % ((comp_bytes + 4 * pixels) / 1e9),
                   "pipelining": f"{args.steps} steps over {M} batch objects / CUDA streams, LF stage of step s gated on step s-{lag}",
                   "parallelism": f"batch-sharded x{world}, no collective"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
        "clocks": sampler.result(), "device_bytes": int(dev_bytes),
    }
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- throughput of the JPEG XL group-decode hot path on B200 (BASELINE.json metric: Mpixels/s decoded,
4K VarDCT batch), with roofline, CPU baseline, end-to-end and single-image latency numbers.

  python bench.py --gpus 1 --steps 5 --warmup 3                    # our CUDA path, BASELINE config 3/5 (4K VarDCT)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU, batch-sharded
  python bench.py --impl reference --gpus 1 --steps 2 --warmup 1   # the reference j40.h on the host CPU cores
  python bench.py --workload c2 | c4                               # 1920x1080 VarDCT / 8192x8192 lossless modular
  python bench.py --scaling strong --total-frames 256              # C5 as written: one fixed batch of 256 distinct frames

A "step" is one decode of this rank's batch of synthetic frames (streams from tools/streamgen; no JPEG XL encoder
exists offline). Frames are independent, so ranks share nothing: no collective on the data path. Weak scaling keeps
the per-GPU batch fixed; strong scaling shards one fixed list of frames with j40_b200.sharding.shard_indices.
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
# many batch objects (CUDA streams) are in flight at once: with the default of 8 hardware work queues, streams
# that share a queue serialise behind each other (measured: +4 % device-resident, +15 % end to end with 32)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

# "d1": >= 1.5 bpp with an 8x8-majority transform histogram (measured 1.85 bpp, 65 % of the varblocks
# in the 8x8 family, 0.36 non-zeros per pixel). "light": round 1's preset (0.9 bpp, large transforms), for history.
PRESETS = {
    "d1": dict(mix=1, tree=1, hfmul=18, hfmul_var=6, big_take=0.35, big_thr=0.03),
    "light": dict(mix=1, tree=1, hfmul=10, hfmul_var=4),
}
WORKLOADS = {
    # name: (kind, width, height, frames per GPU, distinct streams per rank, frames per batch object)
    "c3": ("vardct", 3840, 2160, 64, 8, 64),
    "c2": ("vardct", 1920, 1080, 256, 16, 256),
    "c4": ("modular", 8192, 8192, 2, 1, 1),
}


def stream_opts(args):
    return dict(PRESETS[args.preset]) if WORKLOADS[args.workload][0] == "vardct" else {}


def workload_name(args):
    kind, w, h = WORKLOADS[args.workload][:3]
    if kind == "modular":
        what = "lossless modular frames (fjxl-shaped: prefix codes + LZ77, YCoCg, gradient-predictor MA tree, group shift 8)"
    else:
        what = f"VarDCT frames, tools/streamgen preset '{args.preset}' {PRESETS[args.preset]}"
    if args.scaling == "strong":
        return f"one batch of {args.total_frames} distinct {w}x{h} {what}, sharded over the GPUs"
    return f"batch of {args.frames_per_gpu} independent {w}x{h} {what} per GPU ({args.distinct} distinct, cycled)"


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


def _gen_one(a):
    kind, w, h, seed, opts = a
    from tools import streamgen
    if kind == "modular":
        import numpy as np
        # an 8192x8192 source tiled from a 2048x2048 synthetic photo (the generator's noise synthesis is the slow part)
        t = min(2048, w, h)
        rgb = None
        if w % t == 0 and h % t == 0 and w > t:
            rgb = np.tile(streamgen.synth(t, t, seed), (h // t, w // t, 1))
        return streamgen.modular(w, h, seed=seed, rgb=rgb, **opts)
    return streamgen.vardct(w, h, seed=seed, **opts)


def make_streams(kind, w, h, seeds, opts, share=1):
    from tools import streamgen
    if kind == "vardct":
        streamgen._ensure_tables()  # oracle-derived tables, inherited by forked workers
    procs = max(1, min(len(seeds), ((os.cpu_count() or 2) - 1) // max(1, share), 32))
    if procs == 1 or kind == "modular":  # (the modular writer is multi-threaded itself)
        return [_gen_one((kind, w, h, s, opts)) for s in seeds]
    with mp.get_context("fork").Pool(procs) as pool:
        return pool.map(_gen_one, [(kind, w, h, s, opts) for s in seeds])


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


def _ref_worker(a):
    data, reps = a
    from oracle import ref
    good, secs = ref.time_decode(data, reps)
    return good, secs


def run_reference(args, rank, world):
    """The reference's own CPU implementation (oracle/_ref = unmodified j40.h) on all host cores."""
    if rank != 0:
        return
    kind, w, h = WORKLOADS[args.workload][:3]
    cores = os.cpu_count() or 1
    nproc = max(1, min(cores, 64))
    if kind == "modular":
        nproc = min(nproc, 4)  # ~3 GB of planes per 8192x8192 decode
    datas = [d for d, _ in make_streams(kind, w, h, list(range(min(args.distinct, nproc))), stream_opts(args))]
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
        "impl": "reference", "metric": metric_name(args), "value": mpix, "unit": "Mpix/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32" if kind == "vardct" else "i16",
        "data": "synthetic",
        "config": {"workload": workload_name(args),
                   "implementation": "reference j40.h (unmodified, -O3 -ffp-contract=off) on the host CPU, one process per frame",
                   "frames_per_step": nproc, "processes": nproc,
                   "sample": f"each step decodes {nproc} frames of the workload (one per process)"},
        "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": nproc, "kind": "reference",
                         "sample": f"{nproc} frames per step, one process per frame, {args.steps} steps"},
        "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def metric_name(args):
    return {"c3": "Mpixels/s decoded (4K VarDCT batch)", "c2": "Mpixels/s decoded (1080p VarDCT batch)",
            "c4": "Mpixels/s decoded (8192x8192 lossless modular)"}[args.workload]


def d2h_ceiling(torch, nbytes, reps, dist):
    """what a bare device-to-pinned-host copy of one step's output achieves on this box, all ranks at once: the
    ceiling of any end-to-end number that returns every decoded frame to the host"""
    n = min(nbytes, 1 << 30)
    src = torch.empty(n, dtype=torch.uint8, device="cuda")
    dst = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    del src, dst
    return n * reps / float(t.item()) / 1e9  # GB/s per rank with all ranks copying


def latency_section(J, ref, np, args):
    """single-image latency through j40_from_memory ... j40_frame_pixels_u8x4 (what a dj40.c user sees), next to the
    reference on one host core, for BASELINE configs C1-C3 (and C4 when it is the workload)"""
    from tools import streamgen
    cases = [("c1_256x256_vardct_dct8", "vardct", 256, 256, dict(mix=0, tree=0, cfl=0)),
             ("c2_1920x1080_vardct", "vardct", 1920, 1080, PRESETS[args.preset]),
             ("c3_3840x2160_vardct", "vardct", 3840, 2160, PRESETS[args.preset])]
    if args.workload == "c4":
        cases.append(("c4_8192x8192_modular", "modular", 8192, 8192, {}))
    out = {}
    for name, kind, w, h, opts in cases:
        data, _ = _gen_one((kind, w, h, 1, dict(opts)))
        want, e0, _, _ = ref.decode(data)
        ms = []
        for _ in range(4):
            t0 = time.perf_counter()
            got, err, _, _ = J.decode(data)
            ms.append((time.perf_counter() - t0) * 1e3)
            assert err == "" and np.array_equal(got, want), "single-image decode differs from the reference"
        _, secs = ref.time_decode(data, 1 if kind == "modular" else 2)
        out[name] = {"ours_ms": min(ms[1:]), "reference_ms": min(secs) * 1e3}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=24)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--preset", default="d1", choices=sorted(PRESETS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--total-frames", type=int, default=256, help="strong scaling: size of the fixed batch (distinct frames)")
    ap.add_argument("--frames-per-gpu", type=int, default=0, help="weak scaling: frames per GPU and step (0 = the workload's default)")
    ap.add_argument("--distinct", type=int, default=0, help="weak scaling: distinct streams per rank, cycled to fill the batch")
    ap.add_argument("--obj-frames", type=int, default=0, help="frames per batch object (0 = the workload's default); experiments")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-latency", action="store_true")
    ap.add_argument("--debug-skip", type=int, default=0, help="measurement aid, NOT a bench value: leave stages out of the pipelined "
                    "region (bit 0 tile kernel, bit 1 coefficient kernel, bit 2 LF-group kernels) to see what each costs the others")
    ap.add_argument("--streams", type=int, default=12, help="batch objects (CUDA streams) the timed steps are pipelined over")
    ap.add_argument("--e2e-sets", type=int, default=2, help="groups of `streams` batch objects the end-to-end loop alternates between")
    ap.add_argument("--e2e-mode", default="rolling", choices=["waves", "rolling"], help="waves: a group of objects is resubmitted when all of it has "
                    "landed in host memory; rolling: every object is resubmitted as soon as its own frames have landed")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_world = max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    kind, w, h, F_default, distinct_default, obj_frames = WORKLOADS[args.workload]
    if args.obj_frames > 0:
        obj_frames = args.obj_frames
    args.frames_per_gpu = args.frames_per_gpu or F_default
    args.distinct = args.distinct or distinct_default
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import j40_b200 as J
    from j40_b200.sharding import shard_indices
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
    host_threads = max(1, min(16, (os.cpu_count() or 1) // local_world))  # parse threads of this rank

    # ---- this rank's frames
    opts = stream_opts(args)
    if args.scaling == "strong":
        mine = shard_indices(args.total_frames, rank, world)   # frame i of the fixed batch = seed i
        gen = make_streams(kind, w, h, list(mine), opts, share=local_world)
        datas = [g[0] for g in gen]
        frames = datas
    else:
        gen = make_streams(kind, w, h, [rank * args.distinct + i for i in range(args.distinct)], opts, share=local_world)
        datas = [g[0] for g in gen]
        frames = [datas[i % len(datas)] for i in range(args.frames_per_gpu)]
    stats = [g[1] for g in gen]
    F = len(frames)
    comp_bytes = sum(len(d) for d in frames)
    pixels = F * w * h
    # a step = this rank's frames, in chunks of at most `obj_frames` frames per batch object
    chunks = [frames[i:i + obj_frames] for i in range(0, F, obj_frames)]
    C_ = len(chunks)

    # ---- parity gate: no number counts unless the GPU output of EVERY distinct frame equals the reference's
    want = {}
    for d in datas:
        if d not in want:
            px, e0, _, _ = ref.decode(d)
            assert e0 == "", "oracle rejected a bench stream"
            want[d] = px

    def check_batch(bm, chunk):
        seen = set()
        for i, d in enumerate(chunk):
            if d in seen:
                continue
            seen.add(d)
            assert np.array_equal(bm.read_pixels(i), want[d]), "GPU output differs from the reference"

    # ---- device-resident throughput: inputs + tables already in HBM, kernels only
    M = max(C_, (max(1, args.streams) // C_) * C_)   # batch objects in flight: whole steps
    batches = []
    for m in range(M):
        bm = J.Batch(local_rank)
        chunk = chunks[m % C_]
        rot = (m // C_) % max(1, len(chunk))         # (another rotation of the same frames per object in flight)
        chunk = chunk[rot:] + chunk[:rot]
        bm.add_many(chunk, host_threads)
        bm.upload()
        bm.decode()
        assert bm.wait() == 0, [bm.error(i) for i in range(len(chunk)) if bm.error(i)]
        check_batch(bm, chunk)
        batches.append(bm)
    b = batches[0]
    # serial pass (the first step's objects, one after the other): per-kernel device times for the roofline section
    step_ms, kernel_ms = [], []
    for it in range(args.warmup + 2):
        tot, ks = 0.0, None
        for bm in batches[:C_]:
            bm.decode()
            bm.wait()
            tot += bm.last_decode_ms()
            k = bm.kernel_ms()
            ks = k if ks is None else {n: ks[n] + k[n] for n in k}
        if it >= args.warmup:
            step_ms.append(tot)
            kernel_ms.append(ks)
    serial_ms = sum(step_ms) / len(step_ms)
    dev_bytes = sum(bm.stat(0) for bm in batches)
    launches_per_step = sum(bm.stat(2) for bm in batches[:C_])
    # timed region: K steps pipelined over the M batch objects (one CUDA stream each), so that the latency-bound
    # LF-group kernels of one step overlap the HF / back kernels of its neighbours. All K steps run inside the
    # region; device time is taken with CUDA events on stream 0 after joining every stream.
    if args.debug_skip:
        os.environ["J40B_DEBUG_SKIP"] = str(args.debug_skip)
        args.skip_e2e = args.skip_latency = True
    for _ in range(args.warmup):
        for bm in batches:
            bm.decode()
        for bm in batches:
            assert bm.wait() == 0
    if os.environ.get("J40B_LF_SMHIST"):   # diagnostics: clear the per-SM residency counters of the serial LF kernels
        os.environ["J40B_LF_SMHIST_DUMP"] = "1"
        batches[0].wait()
        os.environ.pop("J40B_LF_SMHIST_DUMP")
    sampler = ClockSampler(local_rank)
    sampler.start()
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    b.mark(0)
    for s_ in range(args.steps):
        for c in range(C_):
            batches[(s_ * C_ + c) % M].decode()
    for bm in batches[1:]:
        b.join(bm)
    b.mark(1)
    stage_sum = {}
    for bm in batches:
        assert bm.wait() == 0
    torch.cuda.synchronize()
    total_ms = b.mark_ms()
    for bm in batches[:min(M, args.steps * C_)]:      # stage times of the last decode of each object, as stretched by co-running
        for n, v in bm.kernel_ms().items():
            stage_sum[n] = stage_sum.get(n, 0.0) + v
    if os.environ.get("J40B_LF_SMHIST"):   # ... and print them for the timed region
        os.environ["J40B_LF_SMHIST_DUMP"] = "1"
        batches[0].wait()
        os.environ.pop("J40B_LF_SMHIST_DUMP")
    os.environ.pop("J40B_DEBUG_SKIP", None)
    if os.environ.get("J40B_TIMELINE"):
        for k, bm in enumerate(batches):
            print("timeline batch %d: " % k + " ".join("%s=%.1f" % (n, bm.event_ms(b, i)) for n, i in
                  [("lf0", 0), ("lfimg", 5), ("hfmeta", 6), ("lf", 1), ("hf", 2), ("tiles", 3), ("end", 4)]), file=sys.stderr)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    launches = launches_per_step * args.steps
    t = torch.tensor([total_ms, float(pixels)], dtype=torch.float64, device="cuda")
    tmax = t.clone()
    if dist:
        dist.barrier()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    total_ms_max = float(tmax[0].item())
    job_pixels = float(t[1].item()) if dist else float(pixels)   # all ranks' pixels per step
    value = job_pixels * args.steps / (total_ms_max / 1e3) / 1e6

    # ---- end to end through the C ABI with HOST buffers, every step: host parse of all frames + one H2D of
    # codestreams and tables + kernels + D2H of every decoded frame into pinned host memory. The batch objects
    # (device + pinned staging allocations) are reused across steps, as a serving loop would, and steps are
    # pipelined over them so that the parse / H2D / D2H of one step overlap the kernels of its neighbours.
    e2e = None
    if not args.skip_e2e:
        # Serving loop: `sets` groups of W batch objects. A group is submitted as one wave (all W decodes back to
        # back) and the next wave of the same group only after all of its D2H copies have landed; while one group
        # copies out and is re-parsed / re-uploaded, the other one computes.
        stride = b.info(0)[2]
        pitch = h * stride
        W = M
        sets = max(1, args.e2e_sets)
        per_obj = max(len(c) for c in chunks) * pitch
        budget = host_memory_budget() // local_world
        if sets * W * per_obj > 0.5 * budget:
            sets = 1
        if W * per_obj > 0.5 * budget:
            W = max(C_, int(0.5 * budget / per_obj) // C_ * C_)
        if sets > 1 and (sets - 1) * W * max(bm.stat(0) for bm in batches) > 0.8 * torch.cuda.mem_get_info()[0]:
            sets = 1  # not enough free HBM for a second group of batch objects
        ceiling = d2h_ceiling(torch, per_obj, 4, dist)
        objs = list(batches[:W])
        for m in range(W, sets * W):
            objs.append(J.Batch(local_rank))
        E = len(objs)
        obj_chunk = [chunks[k % C_] for k in range(E)]
        host_out = [torch.empty((len(obj_chunk[k]), pitch), dtype=torch.uint8, pin_memory=True) for k in range(E)]
        out_np = [t_.numpy() for t_ in host_out]
        host_t = [0.0] * 6
        skip = os.environ.get("J40B_E2E_SKIP", "")  # diagnostics only ("d2h"): such a run is not an end-to-end number

        def submit(k):
            bm = objs[k]
            tt = [time.perf_counter()]
            bm.reset(); tt.append(time.perf_counter())
            bm.add_many(obj_chunk[k], host_threads); tt.append(time.perf_counter())
            bm.upload(); tt.append(time.perf_counter())
            bm.decode(); tt.append(time.perf_counter())
            if skip != "d2h":
                bm.read_all_async(out_np[k])
            tt.append(time.perf_counter())
            for i in range(5):
                host_t[i] += tt[i + 1] - tt[i]

        for k in range(E):          # warm-up (also pages the pinned buffers in)
            submit(k)
        for k in range(E):
            assert objs[k].wait() == 0
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
        n_waves = max(3 * sets, -(-args.steps * C_ // W))
        n_obj = n_waves * W
        host_t[:] = [0.0] * 6
        t0 = time.perf_counter()
        for wv in range(n_waves):
            ids = range((wv % sets) * W, (wv % sets) * W + W)
            if args.e2e_mode == "rolling":
                for k in ids:
                    if wv >= sets:
                        tw = time.perf_counter()
                        assert objs[k].wait() == 0  # this object's previous decode, including its D2H
                        host_t[5] += time.perf_counter() - tw
                    submit(k)
                continue
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
        e2e_pixels = sum(len(obj_chunk[(wv % sets) * W + j]) for wv in range(n_waves) for j in range(W)) * w * h
        h2d = sum(bm.stat(1) for bm in objs[:C_])
        te = torch.tensor([dt, float(e2e_pixels)], dtype=torch.float64, device="cuda")
        tem = te.clone()
        if dist:
            dist.all_reduce(tem, op=dist.ReduceOp.MAX)
            dist.all_reduce(te, op=dist.ReduceOp.SUM)
        # every distinct frame of every object, as it arrived in host memory
        for k in range(E if skip != "d2h" else 0):
            seen = set()
            for i, d in enumerate(obj_chunk[k]):
                if d in seen:
                    continue
                seen.add(d)
                got = out_np[k][i].reshape(h, stride)[:, : w * 4].reshape(h, w, 4)
                assert np.array_equal(got, want[d]), "end-to-end output differs from the reference"
        e2e_value = (float(te[1].item()) if dist else float(e2e_pixels)) / float(tem[0].item()) / 1e6
        d2h_step = F * pitch
        e2e = {"value": e2e_value, "unit": "Mpix/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h_step), "steps": n_obj // C_,
               "d2h_gbs_per_gpu": e2e_value / world * 1e6 * (pitch / (w * h)) / 1e9,
               "ceiling_gbs_per_gpu": ceiling,
               "frac_of_ceiling": e2e_value / world * 1e6 * (pitch / (w * h)) / 1e9 / ceiling,
               "ceiling": "bare cudaMemcpyAsync of one batch object's output into pinned host memory, all ranks at once",
               "host_parse_threads": host_threads,
               "includes": "host parse + H2D + kernels + D2H of all frames into pinned host memory, "
                           f"{n_waves} waves of {W} batch objects over {sets} groups of {W} reused objects, {args.e2e_mode}"}
        for bm in objs[W:]:
            bm.close()
        del host_out, out_np

    latency = None
    if rank == 0 and not args.skip_latency:
        latency = latency_section(J, ref, np, args)

    lf_lanes = int(b.stat(5))   # lanes per LF group the serial LF kernels of the last decode used (32, 16 or 8)
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
    names = {"lf_image": "k_lf_chan stage 0 (LF image: 3 channels x 5 class kernels)",
             "lf_hfmeta": "k_lf_post + k_lf_chan stage 1 (HF metadata: 4 channels x 5 class kernels, varblock placement)",
             "lf_llf": "k_lf_llf", "hf_group": "k_hf_prep+k_hf_group", "back": "k_back_tile", "back_big": "k_back_generic",
             "modular": "k_modular+k_render"}
    traffic_key = {"lf_image": "k_lf_chan", "lf_hfmeta": "k_lf_chan", "hf_group": "k_hf_prep+k_hf_group", "back": "k_back_tile"}
    dominant = max(names, key=lambda k: kavg.get(k, 0.0))
    traffic = traffic_total = None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum per launch and frame from the committed ncu launch list (profiles/)
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        ent = tr.get(traffic_key.get(dominant, ""))
        if ent and ent.get("preset", args.preset) == args.preset:
            # (k_lf_chan: both stages' launches together; the dominant stage is about half of them)
            traffic = ent["dram_bytes_per_frame"] * F
        tot = [v["dram_bytes_per_frame"] for v in tr.values() if isinstance(v, dict) and "dram_bytes_per_frame" in v]
        if tot:
            traffic_total = sum(tot) * F
    except Exception:
        pass
    achieved = alg_bytes / (kavg[dominant] / 1e3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_total": traffic_total,
                "kernel": names[dominant], "kernel_ms": kavg[dominant],
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if peaks else "fallback 6.65 TB/s (B200_PROFILING.md)",
                "algorithmic_bytes_per_launch": alg_bytes,
                "note": "the path is bound by the serial entropy decoders' dependent-issue latency, not by HBM (SURVEY 8d)",
                "all_kernel_ms": kavg, "serial_ms_per_step": serial_ms,
                "stage_ms_in_region": {k: v / max(1, min(M, args.steps * C_)) * C_ for k, v in stage_sum.items()},
                "step_achieved": alg_bytes / (ms_step / 1e3) / 1e9, "step_frac": alg_bytes / (ms_step / 1e3) / 1e9 / peak,
                "back_tile_achieved": alg_bytes / (kavg["back"] / 1e3) / 1e9 if kavg.get("back") else None}

    # ---- CPU baseline: the reference itself, one thread, bounded sample of the same workload
    reps = 6 if kind == "vardct" else 1
    good, secs = ref.time_decode(datas[0], reps)
    cpu = {"value": w * h / statistics.median(secs) / 1e6, "unit": "Mpix/s", "cores": 1, "kind": "reference",
           "sample": f"{reps} decodes of one {w}x{h} bench frame through j40_from_memory..j40_frame_pixels_u8x4, median",
           "host_cores_available": os.cpu_count()}

    npx = len(stats) * w * h
    hist = [sum(s["transform_hist"][i] for s in stats) for i in range(27)]
    line = {
        "metric": metric_name(args), "value": value, "unit": "Mpix/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32" if kind == "vardct" else "i16", "data": "synthetic",
        "config": {"workload": workload_name(args), "preset": args.preset if kind == "vardct" else None,
                   "frames_per_gpu": F, "batch_objects_per_step": C_, "lf_lanes_per_lf_group": lf_lanes if kind == "vardct" else None,
                   "groups_per_gpu": F * ((w + 255) // 256) * ((h + 255) // 256),
                   "compressed_bytes_per_gpu": comp_bytes, "bits_per_pixel": 8.0 * comp_bytes / pixels,
                   "hf_symbols_per_pixel": sum(s["hf_symbols"] for s in stats) / npx,
                   "lf_symbols_per_pixel": sum(s["lf_symbols"] for s in stats) / npx,
                   "nonzeros_per_pixel": sum(s["nonzeros"] for s in stats) / npx,
                   "varblocks_per_frame": sum(s["num_varblocks"] for s in stats) / len(stats),
                   "transform_hist": hist,
                   "share_8x8_family": (sum(hist[i] for i in (0, 1, 2, 3, 12, 13, 14, 15, 16, 17)) / max(1, sum(hist))) if kind == "vardct" else None,
                   "l2": "inputs+outputs per step (%.1f GB) exceed L2" % ((comp_bytes + 4 * pixels) / 1e9),
                   "pipelining": f"{args.steps} steps over {M} batch objects / CUDA streams, free-running",
                   "parity_gate": "every distinct frame of every batch object compared with the reference before timing, and again as it arrives in host memory in the end-to-end loop",
                   "parallelism": f"batch-sharded x{world}, no collective"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "latency": latency, "gpu_launches": int(launches),
        **({"INVALID_debug_skip": args.debug_skip} if args.debug_skip else {}),
        "clocks": sampler.result(), "device_bytes": int(dev_bytes),
    }
    print(json.dumps(line))
    if dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- throughput of the Global Patch Collider inference path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config 0..4]
                  [--forest tau|zero|deep] [--shape 1024x436] [--batch 256]

--config picks a BASELINE.json configuration (default 1, the one the headline metric is quoted on):
  0  sparsematch, defaultTauForest, 1024x436           1  sparsematch, defaultZeroForest, 1024x436
  2  batch of 1920x1080 pairs (32 per GPU), tau forest  3  3840x2160 pair through a 4-level pyramid
  4  deep random forest (16 x 12) on 1080p frames (32 per GPU)
--forest / --shape / --batch override the configuration's values.

A "step" is one pass of the whole hot path (kernel A1: box + Sobel, kernel A2: fern hashing on
TMA-staged tiles, kernel B: per-row matching, scans, kernel C: ordered support emission) over one
batch of synthetic stereo pairs.

  value      stereo pairs/s, inputs resident in HBM, timed with CUDA events on the launching
             stream (max over ranks); `mpix_per_s` = value * 2*W*H / 1e6 (BASELINE.json's second unit)
  e2e        same metric through the C ABI with HOST (pinned) buffers: H2D of the images and D2H
             of the support lists inside the timed region (gpc_match_batch)
  roofline   dominant kernel's algorithmic bytes / its CUDA-event duration vs the measured HBM peak
  cpu_baseline  the unmodified reference (oracle/_ref) on the box's host cores, bounded sample

Multi-GPU (torchrun, one rank per GPU): pairs shard across ranks, no data-path collective;
weak scaling (every rank processes its own batch).  `--impl reference` times the reference's CPU
path only (rank 0), none of this repo's kernels.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FORESTS = {"tau": os.path.join(ROOT, "forests", "defaultTauForest.txt"),
           "zero": os.path.join(ROOT, "forests", "defaultZeroForest.txt"),
           "deep": os.path.join(ROOT, "forests", "deepRandomForest16x12.txt")}
METRIC = "stereo pairs/s (sparsematch: preprocessImage x2 + rectifiedMatch)"
# BASELINE.json configs[i] -> (forest, shape, pairs per step per GPU)
CONFIGS = {0: ("tau", "1024x436", 256), 1: ("zero", "1024x436", 256), 2: ("tau", "1920x1080", 32),
           3: ("tau", "3840x2160", 1), 4: ("deep", "1920x1080", 32)}
CONFIG_NAMES = {0: "sparsematch", 1: "sparsematch", 2: "batch of 1080p pairs, one pair stream per GPU",
                3: "4-level pyramid (3840x2160, 1920x1080, 960x540, 480x270), per-level hashing and matching",
                4: "deep random forest (16 trees x 12 tests, first 32 tests as in the reference) on a 1080p sequence"}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=1, choices=[0, 1, 2, 3, 4])
    ap.add_argument("--forest", default=None, choices=list(FORESTS))
    ap.add_argument("--shape", default=None)
    ap.add_argument("--batch", type=int, default=None, help="pairs per step per GPU")
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic pairs (seeds 1234+i), tiled to the batch")
    ap.add_argument("--sparse", action="store_true", help="low-texture variant of the synthetic pairs (SURVEY.md 8d): ~25 %% of the candidates")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--extended", action="store_true",
                    help="configs[4] with ALL tests of the deep forest (multi-word states, gpc_match_pair_wide) instead of the "
                         "reference's first 32; parity unpinned (no reference result exists), end-to-end line only")
    ap.add_argument("--pool", action="store_true",
                    help="ONE process drives --gpus N devices through gpc_pool_match_batch (library-level multi-GPU driver); "
                         "prints an end-to-end line only")
    args = ap.parse_args()
    forest, shape, batch = CONFIGS[args.config]
    args.forest = args.forest or forest
    args.shape = args.shape or shape
    args.batch = args.batch or batch
    return args


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def workload_name(args, w, h):
    return (f"configs[{args.config}]: {CONFIG_NAMES[args.config]}, forests/{os.path.basename(FORESTS[args.forest])}, "
            f"synthetic {w}x{h} stereo pairs" + (" (low-texture variant)" if args.sparse else ""))


def config_dict(args, w, h):
    """The SAME dict in both arms (the driver compares them): what is computed, not how much of it an arm samples."""
    P = w * h
    return {"workload": workload_name(args, w, h), "shape": f"{w}x{h}", "forest": os.path.basename(FORESTS[args.forest]),
            "settings": "gradientThreshold 5, verticalTolerance 0, dispHigh 128, epipolarMode, sort matcher (sparsematch.cpp:29-34)",
            "pairs_per_step_per_gpu": args.batch, "distinct_pairs": min(args.distinct, args.batch),
            "sharding": "pairs round-robin over the GPUs, no collective",
            "l2": f"inputs larger than L2: {2 * args.batch * P / 1e6:.0f} MB raw + {8 * args.batch * P / 1e6:.0f} MB hash per step vs 126 MB L2"}


def support_digest(supp):
    """FNV-1a-64 variant over an ordered support list [n, 3] int32 (x, y, d as float bits) -- SURVEY.md appendix C.
    Plain Python on purpose: the product arm of bench.py never touches oracle/."""
    supp = np.ascontiguousarray(supp).reshape(-1, 3)
    d = supp[:, 2].copy().view(np.float32).astype(np.int64)
    vals = np.stack([supp[:, 0].astype(np.int64), supp[:, 1].astype(np.int64), d], axis=1).reshape(-1) & 0xFFFFFFFF
    hsh = 1469598103934665603
    for v in vals.tolist():
        hsh = ((hsh ^ v) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % hsh


def golden_records(args, w, h):
    """Reference-generated count + digest per seed for this workload (tests/golden/bench_pairs.json), or None."""
    p = os.path.join(ROOT, "tests", "golden", "bench_pairs.json")
    if not os.path.exists(p):
        return None
    with open(p) as f:
        g = json.load(f)
    key = f"{w}x{h}/{args.forest}" + ("/sparse" if args.sparse else "")
    recs = g["cases"].get(key)
    return {r["seed"]: r for r in recs} if recs else None


def verify_batch(args, w, h, B, pair_ids, d_out, d_n, d_nc, torch):
    """Checks the batch the timed region just produced: every pair's counts against its seed's golden record, the
    ordered support list of the first distinct pairs by digest, and every tiled copy bit-identical to its original."""
    gold = golden_records(args, w, h)
    distinct = min(args.distinct, B)
    n = d_n.cpu().numpy().astype(np.int64)
    nc = d_nc.cpu().numpy().astype(np.int64)
    seed_of = lambda j: 1234 + int(pair_ids[j]) % distinct        # pair g of the job is the synthetic pair of seed 1234 + g mod distinct
    checked = 0
    first = {}
    for j in range(B):
        sd = seed_of(j)
        if sd not in first:
            first[sd] = j
            if gold and sd in gold:
                rec = gold[sd]
                got = d_out[j, :int(n[j])].cpu().numpy()
                assert (int(nc[j, 0]), int(nc[j, 1]), int(n[j])) == (rec["n_cand_l"], rec["n_cand_r"], rec["n_supports"]), \
                    f"pair {j} (seed {sd}): counts {nc[j].tolist()} {int(n[j])} != golden {rec}"
                dg = support_digest(got)
                assert dg == rec["digest"], f"pair {j} (seed {sd}): digest {dg} != golden {rec['digest']}"
                checked += 1
        else:
            k = first[sd]
            assert int(n[j]) == int(n[k]) and (nc[j] == nc[k]).all(), f"pair {j} differs from pair {k} (same seed)"
            assert bool(torch.equal(d_out[j, :int(n[j])], d_out[k, :int(n[k])])), f"pair {j} differs from pair {k} (same seed)"
    return {"pairs_in_batch": B, "digest_checked_pairs": checked, "copies_identical": B - len(first),
            "against": "tests/golden/bench_pairs.json (count + FNV digest of the unmodified reference, scripts/make_golden_bench.py)"
                       if checked else "no golden record for this workload: copies compared with their originals only"}


def copy_ceiling(torch, h_img, d_img, h_out, d_flat, n_out_bytes, dist, reps=3):
    """The box's bare copy ceiling for this step's traffic: pinned H2D of the images and D2H of the supports on two
    streams at once (all ranks at the same time), no kernels.  Returns seconds per step (max over ranks)."""
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    n_el = n_out_bytes // 4
    def once():
        with torch.cuda.stream(s_up):
            d_img.copy_(h_img, non_blocking=True)
        with torch.cuda.stream(s_dn):
            h_out.view(-1)[:n_el].copy_(d_flat[:n_el], non_blocking=True)
    once()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / reps
    if dist is not None:
        t = torch.tensor([sec], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    return sec


def ncu_traffic(kernel, n_pixels):
    """DRAM bytes per launch from the committed ncu capture (profiles/traffic.json: bytes per pixel)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None, None
    with open(p) as f:
        t = json.load(f)
    if kernel not in t.get("dram_bytes_per_pixel", {}):
        return None, None
    return t["dram_bytes_per_pixel"][kernel] * n_pixels, t.get("source")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_near_gpu(index):
    """Pin this rank's host thread (and so, by first touch, its pinned staging buffers) to the CPU cores of
    the GPU's NUMA node: the end-to-end number is bound by host-memory / PCIe traffic when several ranks
    stream at once.  Best effort; returns a note for the JSON line."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, b, rest = bus.split(":")
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{b.lower()}:{rest.lower()}/local_cpulist"
        cpus = set()
        for part in open(path).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"host thread bound to {len(cpus)} cores local to GPU {index}"
    except Exception as e:                                        # noqa: BLE001 - measurement nicety only
        return f"no NUMA binding ({type(e).__name__})"
    return "no NUMA binding"


def make_images(w, h, batch, distinct, sparse=False):
    from opengpc_b200.synth import synth_batch
    base = synth_batch(w, h, min(distinct, batch), seed0=1234, sparse=sparse)
    reps = (batch + len(base) - 1) // len(base)
    return np.ascontiguousarray(np.tile(base, (reps, 1, 1, 1))[:batch])


def cpu_reference_run(images, forest_path, threads, iters, disp_high=128):
    """Times the compiled, unmodified reference (oracle/_ref) or, if absent, the C port."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oraclelib import Oracle, Reference, settings
    n_pairs, _, h, w = images.shape
    if Reference.available():
        ref = Reference()
        sec, tot = ref.time_pairs(images, forest_path, threads=threads, iters=iters, disp_high=disp_high)
        return threads * iters / sec, "reference", tot
    o = Oracle()   # scalar port, single thread
    f = o.read_forest(forest_path)
    t0 = time.perf_counter()
    tot = 0
    n = max(1, iters)
    for i in range(n):
        s, _, _ = o.pair(images[i % n_pairs, 0], images[i % n_pairs, 1], f, settings(disp_high=disp_high))
        tot += len(s)
    return n / (time.perf_counter() - t0), "port", tot


def run_reference_arm(args, w, h):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    images = make_images(w, h, min(args.distinct, 8), args.distinct, args.sparse)
    # one step = every host thread runs the sparsematch window on one pair (bounded sample)
    rates = []
    for _ in range(args.warmup):
        cpu_reference_run(images, FORESTS[args.forest], cores, 1)
    t0 = time.perf_counter()
    kind = "reference"
    for _ in range(args.steps):
        r, kind, _ = cpu_reference_run(images, FORESTS[args.forest], cores, 1)
        rates.append(r)
    wall = time.perf_counter() - t0
    threads = cores if kind == "reference" else 1
    value = threads * args.steps / wall
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * wall / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "mpix_per_s": value * 2 * w * h / 1e6,
            "config": config_dict(args, w, h),
            "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": threads, "kind": kind,
                             "sample": f"{threads} pairs per step ({threads} host threads x 1 pair), "
                                       f"t0..t2 window of sparsematch.cpp:45-52"},
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_pyramid(args, g, ctx, images, w, h, world, rank, dist, torch):
    """configs[3]: one 4K pair through 4 pyramid levels (gpc_match_pyramid, host buffers: the upload
    of the pair and the download of every level's supports are inside the timed region)."""
    settings = g.sparsematch_settings()
    levels = 4
    h_img = torch.from_numpy(images[:1]).pin_memory()              # [1, 2, h, w]
    P = w * h
    cap = (w - 26) * (h - 26)
    h_out = torch.empty((cap, 3), dtype=torch.int32).pin_memory()
    h_off = torch.zeros(levels + 1, dtype=torch.int64).pin_memory()

    def step():
        ctx.match_pyramid_raw(h_img.data_ptr(), h_img.data_ptr() + P, w, h, levels, settings, h_out.data_ptr(), cap, h_off.data_ptr())

    sampler = ClockSampler(int(os.environ.get("LOCAL_RANK", "0")))
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        step()
    l0 = ctx.launches
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    offs = h_off.numpy().copy()
    supp = h_out[:int(offs[-1])]
    # check the levels the timed steps produced against the reference-generated fixture (tests/golden/pyramid.json)
    verified = {"levels_checked": 0}
    gp = os.path.join(ROOT, "tests", "golden", "pyramid.json")
    if os.path.exists(gp) and args.forest == "tau" and not args.sparse and (world == 1 or rank == 0):
        with open(gp) as f:
            for case in json.load(f)["cases"]:
                if (case["w"], case["h"], case["seed"], case["forest"]) == (w, h, 1234, "tau"):
                    for lv in case["levels"][:levels]:
                        got = supp[int(offs[lv["level"]]):int(offs[lv["level"] + 1])].numpy()
                        assert len(got) == lv["n_supports"] and support_digest(got) == lv["digest"], f"pyramid level {lv['level']} differs from the golden record"
                        verified["levels_checked"] += 1
                    verified["against"] = "tests/golden/pyramid.json (the unmodified reference run per level)"
    from opengpc_b200.shard import reduce_timing
    ms, launches = reduce_timing(sec * 1e3, ctx.launches - l0, dist, "cuda")
    clocks = sampler.stop()
    if rank != 0:
        return
    value = world * args.steps / (ms / 1e3)
    pix = sum(2 * (w >> l) * (h >> l) for l in range(levels))
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "mpix_per_s": value * pix / 1e6,
            "config": config_dict(args, w, h), "levels": levels, "supports_per_level": np.diff(offs).tolist(),
            "note": "host-buffer API only: value == e2e (upload + per-level download inside the timed region)",
            "verified": verified, "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 2 * w * h, "d2h_bytes_per_step": int(len(supp)) * 12,
                    "api": "gpc_match_pyramid"}}
    if not args.no_cpu_baseline and world == 1:
        # the reference has no pyramid: its sparsematch window per level on the identically down-sampled images,
        # every host core running one pair of that level
        from opengpc_b200.synth import downsample2x
        cores = os.cpu_count() or 1
        lv_imgs, sec_per_pair, kind = images[:1], 0.0, "reference"
        for l in range(levels):
            rate, kind, _ = cpu_reference_run(lv_imgs, FORESTS[args.forest], cores, 1, disp_high=128 >> l)
            sec_per_pair += 1.0 / rate
            lv_imgs = np.ascontiguousarray(np.stack([np.stack([downsample2x(lv_imgs[0, 0]), downsample2x(lv_imgs[0, 1])])]))
        threads = cores if kind == "reference" else 1
        line["cpu_baseline"] = {"value": 1.0 / sec_per_pair, "unit": "pairs/s", "cores": threads, "kind": kind,
                                "sample": f"{levels} levels x {threads} pairs ({threads} host threads x 1 pair per level), "
                                          f"sparsematch.cpp:45-52 window per level"}
    print(json.dumps(line), flush=True)


def run_pool(args, w, h):
    """--pool: the whole job in one process.  gpc_pool_match_batch over --gpus devices, host (pinned, portable) buffers,
    N x batch pairs per step; uploads, kernels and downloads of all devices inside the timed region."""
    import ctypes as C
    import opengpc_b200 as g
    N, B, P = args.gpus, args.batch * args.gpus, w * h
    lib = g.load_library()
    base = make_images(w, h, min(args.distinct, B), args.distinct, args.sparse)
    img_ptr = lib.gpc_host_alloc(2 * B * P)
    images = np.ctypeslib.as_array(C.cast(img_ptr, C.POINTER(C.c_uint8)), shape=(B, 2, h, w))
    for j in range(B):
        images[j] = base[j % len(base)]
    gold = golden_records(args, w, h)
    est = int(sum(gold[1234 + (j % len(base))]["n_supports"] for j in range(B))) if gold else B * (w - 26) * (h - 26) // 2
    cap = est + 1024
    out_ptr = lib.gpc_host_alloc(cap * 12)
    out = np.ctypeslib.as_array(C.cast(out_ptr, C.POINTER(C.c_int32)), shape=(cap, 3))
    offs = np.zeros(B + 1, np.int64)
    settings = g.sparsematch_settings()
    sampler = ClockSampler(0)
    sampler.start()
    with g.Pool(list(range(N)), max_w=w, max_h=h, max_batch_per_device=(B + N - 1) // N + 16) as pool:
        pool.set_forest(FORESTS[args.forest])
        step = lambda: pool.match_batch_raw(img_ptr, B, w, h, settings, out_ptr, cap, offs.ctypes.data)
        for _ in range(max(args.warmup, 3)):
            step()
        l0 = pool.launches
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step()                                                  # synchronous: returns with the results on the host
        sec = time.perf_counter() - t0
        launches = pool.launches - l0
    clocks = sampler.stop()
    checked = 0
    for j in range(B):
        rec = gold.get(1234 + (j % len(base))) if gold else None
        if rec:
            assert int(offs[j + 1] - offs[j]) == rec["n_supports"], (j, int(offs[j + 1] - offs[j]), rec)
            if j < len(base) or j >= B - 2:
                assert support_digest(out[int(offs[j]):int(offs[j + 1])]) == rec["digest"], j
                checked += 1
    value = B * args.steps / sec
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": N, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "mpix_per_s": value * 2 * P / 1e6, "config": config_dict(args, w, h),
            "note": "one process, gpc_pool_match_batch: value IS the end-to-end figure (host buffers, copies inside the timed region)",
            "verified": {"pairs_in_batch": B, "counts_checked": B if gold else 0, "digest_checked_pairs": checked},
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 2 * B * P, "d2h_bytes_per_step": int(offs[B]) * 12 + (B + 1) * 8,
                    "api": "gpc_pool_match_batch (one context + host thread per GPU, chunks round-robin, supports gathered in pair order)"}}
    print(json.dumps(line), flush=True)
    lib.gpc_host_free(img_ptr)
    lib.gpc_host_free(out_ptr)


def run_extended(args, w, h):
    """--extended: every test of the forest file (192 for the deep forest) through gpc_match_pair_wide, pair by pair,
    host buffers (upload and download inside the timed region)."""
    import opengpc_b200 as g
    B = args.batch
    images = make_images(w, h, B, args.distinct, args.sparse)
    settings = g.sparsematch_settings()
    tests = g.read_forest_tests(FORESTS[args.forest])
    sampler = ClockSampler(0)
    sampler.start()
    with g.Context(device=0, max_w=w, max_h=h, max_batch=1) as ctx:
        ctx.set_wide_forest(tests)
        counts = []
        for j in range(min(B, 3)):                                  # warm-up (also sizes the workspaces)
            counts.append(len(ctx.match_pair_wide(images[j, 0], images[j, 1], settings)[0]))
        l0 = ctx.launches
        t0 = time.perf_counter()
        n = 0
        for _ in range(args.steps):
            for j in range(B):
                supp, _, _ = ctx.match_pair_wide(images[j, 0], images[j, 1], settings)
                if j < len(counts):
                    assert len(supp) == counts[j], "extended mode is not repeatable"
                n += 1
        sec = time.perf_counter() - t0
        launches = ctx.launches - l0
    clocks = sampler.stop()
    value = n / sec
    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": 1, "steps": args.steps, "warmup": 3,
            "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic", "mpix_per_s": value * 2 * w * h / 1e6, "config": config_dict(args, w, h),
            "extended": {"tests": int(len(tests)), "state_words": (len(tests) + 31) // 32, "supports_first_pairs": counts,
                         "parity": "unpinned: the reference keeps the first 32 tests (inference.hpp:426); checked against the "
                                   "scalar restatement in tests/test_wide.py"},
            "note": "pair by pair through gpc_match_pair_wide: value IS the end-to-end figure",
            "clocks": clocks, "gpu_launches": launches,
            "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 2 * B * w * h, "d2h_bytes_per_step": int(sum(counts)) * 12 * B // max(len(counts), 1),
                    "api": "gpc_match_pair_wide"}}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    w, h = (int(v) for v in args.shape.lower().split("x"))
    if args.impl == "reference":
        run_reference_arm(args, w, h)
        return
    if args.extended:
        run_extended(args, w, h)
        return
    if args.pool:
        run_pool(args, w, h)
        return

    import torch
    import opengpc_b200 as g

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa_note = bind_near_gpu(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    B = args.batch
    P = w * h
    # the job is world x B pairs; rank r owns pairs r, r + world, ... (opengpc_b200/shard.py) -- independent units, no collective
    from opengpc_b200.shard import gather_support_counts, shard_pairs
    mine = shard_pairs(world * B, rank, world)
    base = make_images(w, h, min(args.distinct, B), args.distinct, args.sparse)
    images = np.ascontiguousarray(base[mine % len(base)])          # [B, 2, h, w] uint8
    settings = g.sparsematch_settings()
    ctx = g.Context(device=local_rank, max_w=w, max_h=h, max_batch=B)
    ctx.set_forest(FORESTS[args.forest])
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)

    if args.config == 3:
        run_pyramid(args, g, ctx, images, w, h, world, rank, dist, torch)
        return
    cap = (w - 26) * (h - 26) * 6 // 10                           # device-resident capacity per pair
    with torch.cuda.stream(stream):
        d_img = torch.from_numpy(images).cuda(non_blocking=False)
        d_out = torch.empty((B, cap, 3), dtype=torch.int32, device="cuda")
        d_n = torch.zeros(B, dtype=torch.int32, device="cuda")
        d_nc = torch.zeros((B, 2), dtype=torch.int32, device="cuda")

    def step():
        ctx.match_batch_device(d_img.data_ptr(), B, w, h, settings, d_out.data_ptr(), cap, d_n.data_ptr(), d_nc.data_ptr())

    def barrier():
        stream.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident throughput ------------------------------------------------------------
    sampler = ClockSampler(local_rank)                            # samples clocks from warm-up to the end of the timed work
    sampler.start()
    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step()
    barrier()
    n_sup = d_n.cpu().numpy().astype(np.int64)
    assert (n_sup <= cap).all(), "device capacity per pair too small for this workload"
    launches0 = ctx.launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = ctx.launches - launches0
    # per-kernel times: a second, separate pass with CUDA events between the kernels (with events in place the library
    # runs the batch as one serial launch sequence, so these add up to slightly more than ms_per_step)
    ctx.enable_kernel_timing(True)
    with torch.cuda.stream(stream):
        for _ in range(max(3, args.steps // 2)):
            step()
    barrier()
    kms, kruns = ctx.kernel_times()
    ctx.enable_kernel_timing(False)
    from opengpc_b200.shard import reduce_timing
    ms_total, launches = reduce_timing(ms_total, launches, dist, "cuda")     # max over ranks, sum over ranks
    ms_per_step = ms_total / args.steps
    value = world * B / (ms_per_step / 1e3)
    if os.environ.get("GPC_BENCH_NO_VERIFY"):                      # kernel experiments that deliberately break the result
        verified = {"skipped": "GPC_BENCH_NO_VERIFY set: this line is NOT a valid measurement"}
    else:
        verified = verify_batch(args, w, h, B, mine, d_out, d_n, d_nc, torch)   # the batch the timed steps produced
        job_counts = gather_support_counts(n_sup, mine, world * B, dist, "cuda")          # results gathered per pair of the job
        verified["job_pairs"] = int(world * B)
        verified["job_supports"] = int(job_counts.sum())

    # ---- end to end through the C ABI with host buffers -----------------------------------------
    e2e = None
    if not args.no_e2e:
        h_img = torch.from_numpy(images).pin_memory()
        tot_sup = int(n_sup.sum())
        h_out = torch.empty((max(tot_sup, 1) + 1024, 3), dtype=torch.int32).pin_memory()
        h_off = torch.zeros(B + 1, dtype=torch.int64).pin_memory()

        def e2e_step():
            ctx.match_batch_raw(h_img.data_ptr(), B, w, h, settings, h_out.data_ptr(), h_out.shape[0], h_off.data_ptr())

        for _ in range(max(args.warmup, 3)):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()                                            # synchronous: returns with results on the host
        barrier()
        sec = time.perf_counter() - t0
        if dist is not None:
            t = torch.tensor([sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        assert int(h_off[B]) == tot_sup
        offs = h_off.numpy()
        assert (np.diff(offs) == n_sup).all(), "end-to-end support counts differ from the device-resident run"
        j0 = 0
        assert support_digest(h_out[int(offs[j0]):int(offs[j0 + 1])].numpy()) == support_digest(d_out[j0, :int(n_sup[j0])].cpu().numpy())
        clocks = sampler.stop()
        ceil_sec = copy_ceiling(torch, h_img, d_img, h_out, d_out.view(-1), tot_sup * 12, dist)
        # single-pair latency through the same API (SURVEY.md 8d asks for it beside the batched throughput)
        lat = []
        for _ in range(12):
            t1 = time.perf_counter()
            ctx.match_batch_raw(h_img.data_ptr(), 1, w, h, settings, h_out.data_ptr(), h_out.shape[0], h_off.data_ptr())
            lat.append(time.perf_counter() - t1)
        single_pair_ms = 1e3 * float(np.median(lat[2:]))
        e2e = {"value": world * B * args.steps / sec, "unit": "pairs/s",
               "h2d_bytes_per_step": int(2 * B * P), "d2h_bytes_per_step": int(tot_sup * 12 + (B + 1) * 8),
               "single_pair_latency_ms": single_pair_ms,
               "copy_ceiling_pairs_per_s": world * B / ceil_sec,
               "copy_ceiling_gbs": world * (2 * B * P + tot_sup * 12) / ceil_sec / 1e9,
               "copy_ceiling_note": "bare pinned H2D of the images + D2H of the supports on two streams, all ranks at once, no kernels",
               "api": "gpc_match_batch (pinned host buffers; upload, kernels and download pipelined over 3 streams)"}
        e2e["frac_of_ceiling"] = e2e["value"] / e2e["copy_ceiling_pairs_per_s"]

    if args.no_e2e:
        clocks = sampler.stop()

    # ---- the other Sintel forest, device-resident only (same run, fewer steps) ---------------------
    other = None
    if args.config in (0, 1) and world == 1 and not args.no_e2e:
        of = "tau" if args.forest == "zero" else "zero"
        ctx.set_forest(FORESTS[of])
        with torch.cuda.stream(stream):
            for _ in range(3):
                step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(max(3, args.steps // 2)):
                step()
            e1.record(stream)
        barrier()
        other = {f"configs[{1 - args.config}] ({os.path.basename(FORESTS[of])}) pairs/s, device-resident":
                 B * max(3, args.steps // 2) / (e0.elapsed_time(e1) / 1e3)}
        ctx.set_forest(FORESTS[args.forest])
        # the two other matching modes of the reference API, end to end through gpc_match_batch (pinned host buffers)
        nb = min(B, 64)
        for label, st in (("global mode (library defaults: epipolarMode(false), verticalTolerance 1, threshold 10)",
                           g.make_settings(thr=10, disp_high=128, vt=1, epipolar=False)),
                          ("useHashtable(true), sparsematch settings",
                           g.make_settings(thr=5, disp_high=128, vt=0, epipolar=True, use_hashtable=True))):
            for _ in range(2):
                ctx.match_batch_raw(h_img.data_ptr(), nb, w, h, st, h_out.data_ptr(), h_out.shape[0], h_off.data_ptr())
            barrier()
            t0 = time.perf_counter()
            for _ in range(3):
                ctx.match_batch_raw(h_img.data_ptr(), nb, w, h, st, h_out.data_ptr(), h_out.shape[0], h_off.data_ptr())
            barrier()
            other[f"{label}: pairs/s end to end, batch {nb}"] = 3 * nb / (time.perf_counter() - t0)
        # the reference's SSE=OFF result mode (GPC_RESULTS_NAIVE), device-resident like `value`
        ctx.set_result_mode(True)
        with torch.cuda.stream(stream):
            for _ in range(3):
                step()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(max(3, args.steps // 2)):
                step()
            e1.record(stream)
        barrier()
        other["GPC_RESULTS_NAIVE (the reference's SSE=OFF build), same forest: pairs/s, device-resident"] = \
            B * max(3, args.steps // 2) / (e0.elapsed_time(e1) / 1e3)
        ctx.set_result_mode(False)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    peak, peak_src = peaks()
    per_kernel_ms = {k: v / max(kruns, 1) for k, v in kms.items()}
    dominant = max(per_kernel_ms, key=per_kernel_ms.get)
    mean_sup = float(n_sup.mean())
    alg_bytes = {"smooth_sobel": 4.0 * P * B,                     # per pair: read 2 u8 images, write 2 u8 smoothed images
                 "hash_tiles": 10.0 * P * B,                      # read 2 u8 smoothed images, write 2 u32 hash images
                 "match_rows": (8.0 * P + 4.0 * mean_sup) * B,    # read 2 hash images, write staged matches
                 "scans": 0.0, "emit_supports": 16.0 * mean_sup * B}
    dom_ms = per_kernel_ms[dominant]
    achieved = alg_bytes[dominant] / (dom_ms / 1e3) / 1e9 if dom_ms > 0 else 0.0
    path_bytes = (10.0 * P + 12.0 * mean_sup)                     # B_alg per pair
    traffic, traffic_src = ncu_traffic(dominant, (2 * B if dominant in ("smooth_sobel", "hash_tiles") else B) * P)
    path_gbs = path_bytes * value / world / 1e9                   # B_alg x pairs/s of one GPU
    kernel_frac = {k: (alg_bytes[k] / (v / 1e3) / 1e9 / peak if v > 0 else None) for k, v in per_kernel_ms.items()}
    all_traffic = {k: ncu_traffic(k, (2 * B if k in ("smooth_sobel", "hash_tiles") else B) * P)[0] for k in per_kernel_ms}
    roofline = {"bound": "hbm", "achieved": path_gbs, "peak": peak, "unit": "GB/s", "frac": path_gbs / peak,
                "definition": "whole path: (10*W*H + 12*Ns) bytes per pair x pairs/s per GPU / measured HBM peak (SURVEY.md 8d)",
                "path_bytes_per_pair": path_bytes, "peak_source": peak_src,
                "kernel": dominant, "kernel_achieved": achieved, "kernel_frac": achieved / peak,
                "algorithmic_bytes_per_launch": alg_bytes[dominant],
                "traffic": traffic, "traffic_source": traffic_src,
                "kernel_ms_per_step": per_kernel_ms, "kernel_fracs": kernel_frac, "kernel_traffic": all_traffic}

    line = {"metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "mpix_per_s": value * 2 * P / 1e6,
            "config": config_dict(args, w, h), "supports_per_pair": mean_sup, "host": numa_note, "verified": verified,
            "clocks": clocks, "gpu_launches": launches, "roofline": roofline}
    if e2e:
        line["e2e"] = e2e
    if other:
        line["other_configs"] = other

    # ---- CPU baseline: the unmodified reference on this box's host cores -------------------------
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        iters = max(1, min(8, int(round(20.0 / (0.09 * (P / 446464.0) * cores))) or 1))
        rate, kind, _ = cpu_reference_run(images[:min(8, B)], FORESTS[args.forest], cores if True else 1, iters)
        threads = cores if kind == "reference" else 1
        line["cpu_baseline"] = {"value": rate, "unit": "pairs/s", "cores": threads, "kind": kind,
                                "sample": f"{threads * iters} pairs ({threads} host threads x {iters}), same synthetic "
                                          f"workload, t0..t2 window of sparsematch.cpp:45-52"}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

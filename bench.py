#!/usr/bin/env python
"""bench.py -- one JSON line per run (see the measurement contract in DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|small]
  python bench.py --impl reference ...      # the reference's CPU path on the host cores

A "step" = one full CCD pass (AABB build -> sort -> sweep -> Tight-Inclusion narrow phase
-> earliest TOI, vertex-face then edge-edge) over one synthetic two-frame scene.
  value : ms/step with the mesh already resident in HBM (sccd_ccd / sccd_ccd_sharded)
  e2e   : ms/step through the host-pointer entry point (sccd_ccd_host / sccd_ccd_sharded_host):
          pinned host buffers -> H2D -> pipeline -> TOI back on the host, all inside the timed
          region
Workload: config 2 (the 1M-primitive scene the metric is quoted on) for a plain `python
bench.py`; config 4 (~50M AABBs) for EVERY launch under torch.distributed.run, world size 1
included, so that one scaling series is one scene ("strong" scaling).
N > 1 : one process per GPU; torch.distributed (gloo) is the rendezvous only.  The ranks split
        the step inside the library (csrc/shard.cu): element slices -> 8-byte records exchanged by
        owning cell range (NCCL send/recv) -> local sort / sweep / narrow phase -> NCCL
        all-reduce(min) of the TOI.
"""
import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ccd_step_time"
UNIT = "ms/step"
PARAMS = dict(ms=0.0, max_iter=-1, tol=1e-6, allow_zero_toi=True)  # tests/test_narrow_phase.cu:41-45


DESCS = {
    "small": "cloth 31x31 over UV sphere (~6K boxes)",
    "c1": "config 1: cloth 101x101 over UV sphere (~62K boxes)",
    "c2": "config 2: cloth-ball, cloth 409x409 + icosphere L5 (~1.06M primitives)",
    "c3": "config 3: 10,000 x 602-primitive blobs, heavy-tailed pile (~6.0M boxes)",
    "c4": "config 4: 83,000 blobs in an x-slab (~50M boxes)",
}


def make_desc(name):
    return DESCS[name]


def make_scene(scenes, name):
    if name not in DESCS:
        raise SystemExit(f"unknown workload {name}")
    gen = {"small": lambda: scenes.cloth_on_sphere(31, seed=7, sphere="uv"), "c1": scenes.scene_c1,
           "c2": scenes.scene_c2, "c3": scenes.scene_c3, "c4": scenes.scene_c4}[name]
    return gen(), DESCS[name]


class ClockSampler:
    """SM clock / throttle reasons DURING the timed region: NVML from a thread of this process
    (a second `nvidia-smi -lms` process per rank holds driver locks long enough to slow the
    step it is supposed to watch), nvidia-smi only if NVML cannot be loaded."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index, period_s=0.02):
        self.rows, self.proc, self.h, self.stop_flag = [], None, None, False
        self.sm, self.reasons, self.max_mhz = [], set(), None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.period = period_s
            self.t = threading.Thread(target=self._poll, daemon=True)
            self.t.start()
            return
        except Exception:
            self.h = None
        exe = shutil.which("nvidia-smi")
        if exe:
            try:
                self.proc = subprocess.Popen(
                    [exe, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                     "-i", str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.t = threading.Thread(target=self._read, daemon=True)
                self.t.start()
            except Exception:
                self.proc = None

    def _poll(self):
        nv = self.nv
        bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                mhz = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.time(), mhz, r))
                for name, bit in bits.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.h is not None:
            self.stop_flag = True
            self.t.join(timeout=1.0)
            sm = [m for (t, m, _) in self.rows if t0 <= t <= t1] or [m for (_, m, _) in self.rows]
            if not sm:
                return None
            return {"sm_mhz": statistics.median(sm), "sm_max_mhz": self.max_mhz,
                    "reasons": sorted(self.reasons), "samples": len(sm), "source": "nvml"}
        if not self.proc:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except Exception:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return None
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


def cpu_step(orc, scene):
    """One CCD step on the host cores: the UNMODIFIED reference CPU broad phase
    (oracle/_ref, TBB replaced by an OpenMP stub) when it is built, else the oracle port;
    the narrow phase is the oracle port (the reference has no CPU narrow phase)."""
    import numpy as np
    t0 = time.perf_counter()
    if orc.ref_cpu() is not None:
        r = orc.ref_cpu_broad_phase(scene, r=PARAMS["ms"])
        vf, ee, kind = r["vf"], r["ee"], "reference"
    else:
        vb, eb, fb = orc.build_boxes(scene, PARAMS["ms"])
        vf = orc.sort_and_sweep_two_lists(vb, fb, 0)[0]
        ee = orc.sort_and_sweep(eb, 0)[0]
        kind = "port"
    t1 = time.perf_counter()
    toi = 1.0
    for pairs, is_vf in ((vf, True), (ee, False)):
        q = orc.gather_queries(scene, np.ascontiguousarray(pairs), is_vf)
        toi, _, _ = orc.narrow_phase(q, is_vf, PARAMS["ms"], PARAMS["max_iter"], PARAMS["tol"],
                                     PARAMS["allow_zero_toi"], toi, per_query=True)
    t2 = time.perf_counter()
    return {"ms": (t2 - t0) * 1e3, "broad_ms": (t1 - t0) * 1e3, "narrow_ms": (t2 - t1) * 1e3,
            "kind": kind, "toi": toi, "n_pairs": [len(vf), len(ee)]}


def host_threads():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU reference arm is entitled to
    all host cores, so the count is set explicitly before any OpenMP runtime starts."""
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        pass
    os.environ["OMP_NUM_THREADS"] = str(n)
    return n


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path on the host cores
    (oracle/_ref when it was built, else the oracle port), all host threads.  Config 4 is too
    big for a CPU step of a few minutes (the reference's one-axis CPU sweep is O(N^5/3) on a
    dense pile), so there the arm times a NAMED SAMPLE SCENE -- the same generator with 1/16 of
    the instances -- and reports that scene's own time: nothing is extrapolated, and
    `config.workload` says what was run."""
    if rank != 0:
        return
    threads = host_threads()
    from _pkg import load_package
    from oracle import orc
    sccd = load_package()
    if args.workload == "c4":
        n_inst = 83_000 // 16
        scene = sccd.scenes.scene_c4(n_inst=n_inst)
        nb = scene['V0'].shape[0] + scene['E'].shape[0] + scene['F'].shape[0]
        desc = (f"config 4 SAMPLE SCENE: {n_inst} blobs in an x-slab ({nb} boxes; 1/16 of the "
                "instances of config 4, same generator and density) -- its own time, not scaled")
    else:
        scene, desc = make_scene(sccd.scenes, args.workload)
    budget_s = 150.0
    t_start = time.perf_counter()
    steps = []
    warm = min(args.warmup, 1)
    for i in range(warm + args.steps):
        r = cpu_step(orc, scene)
        if i >= warm:
            steps.append(r)
        if time.perf_counter() - t_start > budget_s and steps:
            break
    ms = statistics.mean(s["ms"] for s in steps)
    cores = orc.lib().orc_num_threads()
    sample = (f"{len(steps)} full step(s) of `config.workload` (time-bounded to ~{int(budget_s)} s); "
              f"broad phase = {steps[0]['kind']} CPU sort_and_sweep "
              f"({statistics.mean(s['broad_ms'] for s in steps):.0f} ms, oneTBB replaced by an OpenMP stub), "
              f"narrow phase = oracle port with OpenMP ({statistics.mean(s['narrow_ms'] for s in steps):.0f} ms; "
              f"the reference has no CPU narrow phase); OMP_NUM_THREADS={threads}")
    line = {
        "impl": "reference", "metric": METRIC, "value": ms, "unit": UNIT, "n_gpus": args.gpus,
        "steps": len(steps), "warmup": warm, "ms_per_step": ms, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, **PARAMS},
        "cpu_baseline": {"value": ms, "unit": UNIT, "cores": cores, "kind": steps[0]["kind"],
                         "sample": sample},
        "e2e": {"value": ms, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "toi": steps[0]["toi"], "n_pairs": steps[0]["n_pairs"],
    }
    print(json.dumps(line), flush=True)


def ref_cuda_subprocess(workload, calls=10, timeout_s=300):
    """The UNMODIFIED reference CUDA ccd() on the same box (oracle/_ref/libref_sccd_cuda.so), in
    a subprocess so a crash / hang cannot take the bench down: 3 untimed calls, then `calls`
    timed ones; the median is the number to compare with.  {'unavailable': why} otherwise."""
    code = f"""
import sys, json, statistics
sys.path.insert(0, {ROOT!r})
from _pkg import load_package
from oracle import orc
import bench
sccd = load_package()
if orc.ref_cuda(False) is None:
    print(json.dumps({{"unavailable": "oracle/_ref/libref_sccd_cuda.so not built"}})); sys.exit(0)
scene, _ = bench.make_scene(sccd.scenes, {workload!r})
for i in range(3):
    r = orc.ref_cuda_ccd(scene, **bench.PARAMS)
out = []
for i in range({calls}):
    r = orc.ref_cuda_ccd(scene, **bench.PARAMS)
    out.append(r["ms"])
b = orc.ref_cuda_broad_phase(scene, want_pairs=False)
print(json.dumps({{"median_ms": statistics.median(out), "min_ms": min(out), "max_ms": max(out),
                  "calls": len(out), "warmup_calls": 3, "all_ms": out, "toi": r["toi"],
                  "broad_phase_ms": b["ms"], "n_vf": b["n_vf"], "n_ee": b["n_ee"],
                  "what": "unmodified reference CUDA ccd() (host mesh in, TOI out), wall clock "
                          "inside the shim around the call"}}))
"""
    try:
        p = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                           timeout=timeout_s)
        for line in reversed(p.stdout.strip().splitlines()):
            if line.startswith("{"):
                return json.loads(line)
        return {"unavailable": f"rc={p.returncode}: {(p.stderr or p.stdout)[-300:]}"}
    except subprocess.TimeoutExpired:
        return {"unavailable": f"timed out after {timeout_s} s"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None,
                    help="c1|c2|c3|c4|small.  Default: config 2 (the 1M-primitive scene the metric "
                         "is quoted on) for a plain `python bench.py`; config 4 (~50M AABBs) for "
                         "EVERY launch under torch.distributed.run, world size 1 included, so "
                         "that one scaling series is one scene")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    args = ap.parse_args()
    under_launcher = "RANK" in os.environ
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload is None:
        args.workload = "c4" if (under_launcher or max(world, args.gpus) > 1) else "c2"
    if args.impl == "reference":
        return run_reference(args, rank, world)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from _pkg import load_package
    sccd = load_package()
    K = sccd.capi
    torch.cuda.set_device(local)
    if world > 1:
        # torch.distributed is the launcher's rendezvous only (gloo): it carries rank 0's NCCL id
        # and the max-over-ranks of the timings.  The data path -- record exchange, all-gather of
        # the host mesh, min-TOI all-reduce -- is NCCL inside the library (sccd_comm_create).
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("gloo")
    scene, desc = make_scene(sccd.scenes, args.workload)
    nV, nE, nF = scene["V0"].shape[0], scene["E"].shape[0], scene["F"].shape[0]

    stream = torch.cuda.current_stream().cuda_stream
    ctx = sccd.Context(local, stream)
    ctx.upload_mesh(scene["V0"], scene["V1"], scene["E"], scene["F"])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # pinned host copies for the end-to-end arm (column-major, as the C ABI takes them)
    pinned = {k: torch.from_numpy(np.ascontiguousarray(v.T)).pin_memory() for k, v in scene.items()}
    host_args = (pinned["V0"].data_ptr(), pinned["V1"].data_ptr(), pinned["E"].data_ptr(),
                 pinned["F"].data_ptr())

    def timed(fn, steps, warmup, sampler=False):
        # the sampler starts before the warm-up (NVML initialisation stalls the first CUDA calls
        # that follow it); only samples taken inside [t0, t1] are reported
        smp = ClockSampler(local) if sampler else None
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.time()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
              for _ in range(steps)]
        stats, toi = [], None
        for a, b in ev:
            flush.zero_()                 # evict L2 between timed iterations (untimed)
            a.record()
            toi = fn()
            b.record()
            stats.append(ctx.stats())
        barrier()
        t1 = time.time()
        clocks = smp.stop(t0, t1) if smp else None
        per_step = [a.elapsed_time(b) for a, b in ev]
        timed.last_steps = per_step
        return max_over_ranks(sum(per_step) / steps), toi, stats, clocks

    single_ms = single_toi = single_pairs = None
    if world > 1:
        # the same scene on ONE GPU (every rank, plain pipeline) so that the line carries its own
        # strong-scaling reference, timed like the steps below
        ms1, single_toi, st1, _ = timed(lambda: ctx.ccd(**PARAMS), 3, 2)
        single_ms = ms1
        single_pairs = st1[-1]["n_pairs"]
        uid = [sccd.Context.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_create(uid[0], rank, world)
        step_resident = lambda: ctx.ccd_sharded(**PARAMS)
        step_e2e = lambda: ctx.ccd_sharded_host(*host_args, sizes=(nV, nE, nF), **PARAMS)
    else:
        step_resident = lambda: ctx.ccd(**PARAMS)
        step_e2e = lambda: ctx.ccd_host(*host_args, sizes=(nV, nE, nF), **PARAMS)

    ms, toi, stats, clocks = timed(step_resident, args.steps, args.warmup, sampler=True)
    ms_steps = list(timed.last_steps)

    def avg(key, *idx, src=None):
        vals = []
        for s in (src or stats):
            v = s[key]
            for i in idx:
                v = v[i]
            vals.append(v)
        return float(sum(vals)) / len(vals)

    def allsum(x):   # whole-job counts (sum over ranks)
        if world == 1:
            return x
        t = torch.tensor(x, dtype=torch.float64)
        dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    n_pairs = allsum([avg("n_pairs", 0), avg("n_pairs", 1)])
    if world > 1:
        assert toi == single_toi, ("sharded TOI differs from the single-GPU TOI", toi, single_toi)
        assert [int(v) for v in n_pairs] == list(single_pairs), \
            ("the ranks' pair lists do not add up to the single-GPU lists", n_pairs, single_pairs)
    e2e_ms, toi2, _, _ = timed(step_e2e, args.steps, 1)
    assert toi == toi2, (toi, toi2)
    # what a simulator with a fixed topology pays per step (sccd_update_vertices: only the two
    # vertex frames cross PCIe).  Extra information; the contract's e2e is the line above.
    e2e_v = None
    if world == 1:
        try:
            def step_e2e_vertices():
                ctx.update_vertices(pinned["V0"].data_ptr(), pinned["V1"].data_ptr(), nV=nV,
                                    host=True)
                return ctx.ccd(**PARAMS)
            v_ms, toi3, _, _ = timed(step_e2e_vertices, args.steps, 1)
            if toi3 == toi:
                e2e_v = {"value": v_ms, "unit": UNIT, "h2d_bytes_per_step": 2 * 24 * nV,
                         "d2h_bytes_per_step": 8,
                         "note": "pinned vertex frames -> sccd_update_vertices + sccd_ccd"}
        except Exception as exc:  # never let the extra arm cost the bench line
            e2e_v = {"error": repr(exc)[:200]}

    # ---- kernel-level pass (untimed for `value`): every solver round gets its own event pair
    # (profile mode 2: both lists on one stream, so that a kernel's event pair brackets that
    # kernel alone -- with two streams the pairs include waiting behind the other list's kernels)
    ctx.set_option(K.OPT_PROFILE, 2)
    prof = []
    for _ in range(3):
        flush.zero_()
        step_resident()
        prof.append(ctx.stats())
    ctx.set_option(K.OPT_PROFILE, 0)
    pavg = lambda key, *idx: avg(key, *idx, src=prof)

    n_checks = allsum([avg("n_box_checks", 0), avg("n_box_checks", 1)])
    n_cand = allsum([avg("n_candidates", 0), avg("n_candidates", 1)])
    loc_pairs = [avg("n_pairs", 0), avg("n_pairs", 1)]
    loc_recs = [avg("n_records", 0), avg("n_records", 1)]
    loc_culled = [avg("n_culled", 0), avg("n_culled", 1)]
    n_boxes = [nV + nF, nE]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_source = "MEASURED_PEAKS.json (measured copy bandwidth)" if peaks else \
        "fallback 6650 GB/s (B200_PROFILING.md)"
    dfma_peak = ctx.measure_fp64_peak()        # thread-level DFMA/s of this device, measured now
    fp64_peak_tflops = 2.0 * dfma_peak / 1e12
    traffic_db = {}
    try:
        traffic_db = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json")))
    except Exception:
        pass

    # One entry per KERNEL: measured device time per launch of this rank, the
    # roofline that bounds it, and its algorithmic work per launch.
    #   HBM kernels : bytes every box / record / query has to cross the memory system once
    #   solver      : FP64 flops = box checks x (VF 60 DFMA + 36 DADD = 156, EE 48 + 36 = 132)
    kernels = {}

    def hbm(name, ms_, nbytes, what):
        if ms_ > 0:
            gbs = nbytes / (ms_ * 1e-3) / 1e9
            kernels[name] = {"key": name.split(" ")[0],
                             "ms": ms_, "bound": "hbm", "achieved": gbs, "peak": hbm_peak,
                             "unit": "GB/s", "frac": gbs / hbm_peak,
                             "algorithmic_bytes_per_launch": nbytes, "work": what}

    def fp64(name, ms_, checks, flop_per_check):
        if ms_ > 0:
            tf = checks * flop_per_check / (ms_ * 1e-3) / 1e12
            kernels[name] = {"key": name, "ms": ms_, "bound": "fp64", "achieved": tf,
                             "peak": fp64_peak_tflops,
                             "unit": "TFLOP/s", "frac": tf / fp64_peak_tflops,
                             "algorithmic_flop_per_launch": checks * flop_per_check,
                             "work": f"{checks:.0f} box checks x {flop_per_check} FP64 flop"}

    share = 1.0 / world
    if world == 1:
        hbm("boxes (vertex_boxes + element_boxes, 2 launches)", pavg("ms_k_boxes"),
            48 * nV + 8 * nE + 12 * nF + 64 * (nV + nE + nF), "48 B/vertex + indices in, 64 B/box out")
    else:
        hbm("boxes (vertex_boxes + 2 x list_boxes of the sample, 3 launches)", pavg("ms_k_boxes"),
            (48 + 96) * nV + (64 + 60) * (nV + nE + nF) / 16,
            "replicated: vertex table + vertex boxes of all vertices, 1-in-16 box sample")
    hbm("gather (2 launches)", pavg("ms_k_gather"), 2 * 64 * (loc_recs[0] + loc_recs[1]),
        "64 B exact record in + out per sweep record")
    for k, nm in ((0, "vf"), (1, "ee")):
        passes = (int(pavg("key_bits", k)) + 7) // 8
        hbm(f"radix_sort_{nm} ({passes} digit passes + histogram)", pavg("ms_k_sort", k),
            16 * loc_recs[k] * passes + 8 * loc_recs[k],
            "8 B (key, index) record read + written per digit pass, read once for the histogram")
        hbm(f"sweep_count_{nm}", pavg("ms_k_sweep_count", k), 64 * loc_recs[k],
            "64 B per sweep record read once")
        hbm(f"sweep_place_{nm}", pavg("ms_k_sweep_fill", k), 8 * loc_pairs[k], "8 B per pair written")
        hbm(f"narrow_cull_{nm}", pavg("ms_k_cull", k), (8 + 192 + 4) * loc_pairs[k],
            "pair + 8 vertices (192 B) in, survivor index out, per query")
        for r in range(5):
            fp64(f"narrow_round{r}_{nm}", pavg("ms_k_round", k, r), pavg("n_round_checks", k, r),
                 156 if k == 0 else 132)
    for v in kernels.values():      # DRAM bytes per step of the same kernels from the ncu capture
        if v["key"].startswith("narrow_round"):   # (the solver launches of a pass add up under one key)
            v["key"] = "narrow_solve_" + v["key"][-2:]
        v["traffic"] = traffic_db.get(f"{args.workload}:{v['key']}")
    dom = max(kernels, key=lambda k_: kernels[k_]["ms"]) if kernels else None
    roofline = None
    if dom:
        d = kernels[dom]
        roofline = {"kernel": dom, "bound": d["bound"], "achieved": d["achieved"], "peak": d["peak"],
                    "unit": d["unit"], "frac": d["frac"],
                    "traffic": traffic_db.get(f"{args.workload}:{d['key']}"),
                    "traffic_source": "profiles/dram_traffic.json (ncu --set full capture of one "
                                      "whole step of this workload, summed per kernel family; all "
                                      "solver launches of a pass under one figure; not measured in "
                                      "this run)",
                    "kernel_ms": d["ms"], "work_per_launch": d["work"],
                    "peak_source": (peak_source if d["bound"] == "hbm" else
                                    "FP64 peak = 2 x DFMA/s measured on this device in this run by "
                                    "a register-only micro-benchmark (sccd_measure_fp64_peak)"),
                    "note": ("dominant kernel by measured device time among ALL kernels of the "
                             "step (per-kernel event pairs, SCCD_OPT_PROFILE pass); the solver "
                             "rounds are bound by the FP64 pipe / dependent-chain latency, "
                             "everything else by HBM"),
                    "all_kernels": kernels}
    # the HBM-bound kernel with the most time, for the memory-system view of the same step
    hbm_only = {k_: v for k_, v in kernels.items() if v["bound"] == "hbm"}
    dom_h = max(hbm_only, key=lambda k_: hbm_only[k_]["ms"]) if hbm_only else None
    # (stage and kernel times come from the SCCD_OPT_PROFILE pass: the timed steps above only
    # carry the total, their event records were a sixth of the host's work per step)
    narrow_ms = pavg("ms_k_narrow", 0) + pavg("ms_k_narrow", 1)
    stage_ms = {"build": pavg("ms_build"), "sort": pavg("ms_sort"), "sweep_vf": pavg("ms_sweep", 0),
                "sweep_ee": pavg("ms_sweep", 1), "narrow_vf": pavg("ms_narrow", 0),
                "narrow_ee": pavg("ms_narrow", 1), "exchange": pavg("ms_exchange"),
                "total_device_profiled": pavg("ms_total"), "total_device": avg("ms_total")}
    rank_stage_ms = None
    if world > 1:
        rank_stage_ms = [None] * world
        mine = dict(stage_ms)
        mine.update(records=loc_recs, pairs=loc_pairs, records_sent=[avg("n_records_sent", 0),
                                                                    avg("n_records_sent", 1)],
                    k_boxes=pavg("ms_k_boxes"), k_expand=[pavg("ms_k_expand", 0), pavg("ms_k_expand", 1)],
                    k_sort=[pavg("ms_k_sort", 0), pavg("ms_k_sort", 1)], k_gather=pavg("ms_k_gather"),
                    host_syncs=avg("n_host_syncs"))
        dist.all_gather_object(rank_stage_ms, mine)
    fp64_instr = 96 * n_checks[0] + 84 * n_checks[1]
    line = {
        "metric": METRIC, "value": ms, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "n_vertices": nV, "n_edges": nE, "n_faces": nF,
                   "workload_rule": ("config 2 for a plain `python bench.py`; config 4 for every "
                                     "launch under torch.distributed.run (world size 1 included)"),
                   "l2": "256 MiB device write between timed steps (flush)", **PARAMS,
                   "parallelism": (f"{world} ranks: element slices -> 8 B records exchanged by cell "
                                   "range (NCCL send/recv) -> local sort/sweep/narrow -> NCCL min"
                                   if world > 1 else "single GPU")},
        "clocks": clocks,
        "e2e": {"value": e2e_ms, "unit": UNIT,
                "h2d_bytes_per_step": 2 * 24 * nV + 8 * nE + 12 * nF, "d2h_bytes_per_step": 8,
                "note": ("sccd_ccd_sharded_host: each rank copies 1/N of the pinned host mesh "
                         "H2D, NCCL all-gather of the slices (whole job bytes)" if world > 1
                         else "pinned host mesh -> sccd_ccd_host")},
        "e2e_vertices_only": e2e_v,
        "gpu_launches": int(avg("n_launches")) * args.steps,
        "host_syncs_per_step": avg("n_host_syncs"),
        "roofline": roofline,
        "roofline_hbm": ({"kernel": dom_h, **{k_: hbm_only[dom_h][k_] for k_ in
                                              ("achieved", "peak", "unit", "frac", "ms")},
                          "traffic": traffic_db.get(f"{args.workload}:{hbm_only[dom_h]['key']}")}
                         if dom_h else None),
        "toi": toi, "n_pairs": n_pairs, "ms_steps_rank0": ms_steps,
        "stage_ms_per_rank": rank_stage_ms,
        "single_gpu_ms_same_workload": single_ms,
        "speedup_vs_single_gpu": (single_ms / ms if single_ms else None),
        "n_prefilter_survivors": n_cand, "n_box_checks": n_checks,
        "narrow_queries_per_s": (sum(n_pairs) / (narrow_ms * 1e-3)) if narrow_ms > 0 else None,
        "narrow_box_checks_per_s": (sum(n_checks) / (narrow_ms * 1e-3)) if narrow_ms > 0 else None,
        "narrow_fp64_instr_per_s": (fp64_instr / (narrow_ms * 1e-3)) if narrow_ms > 0 else None,
        "sweep_candidate_tests_per_s": ((n_cand[0] + n_cand[1]) / ((pavg("ms_k_sweep_count", 0)
                                        + pavg("ms_k_sweep_count", 1)) * 1e-3)
                                        if pavg("ms_k_sweep_count", 0) + pavg("ms_k_sweep_count", 1) > 0 else None),
        # queue load balance: items every solver round read, box checks per round,
        # sub-boxes handed on, and whether a bounded item list was ever full -- this rank
        "narrow_load_balance": {"round_items": [[avg("n_round_items", k, r) for r in range(6)] for k in (0, 1)],
                                "round_checks": [[avg("n_round_checks", k, r) for r in range(5)] for k in (0, 1)],
                                "culled": loc_culled, "skipped": [avg("n_skipped", 0), avg("n_skipped", 1)],
                                "donated": [avg("n_donated", 0), avg("n_donated", 1)],
                                "queue_overflow": avg("queue_overflow")},
        "stage_ms": stage_ms,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload in ("small", "c1", "c2"):
        from oracle import orc
        host_threads()
        r = cpu_step(orc, scene)
        assert r["toi"] == toi and r["n_pairs"] == [int(n_pairs[0]), int(n_pairs[1])], \
            ("GPU result differs from the CPU baseline", r["toi"], toi, r["n_pairs"], n_pairs)
        line["cpu_baseline"] = {
            "value": r["ms"], "unit": UNIT, "cores": orc.lib().orc_num_threads(), "kind": r["kind"],
            "sample": (f"1 full step of the same workload: broad phase = {r['kind']} CPU "
                       f"sort_and_sweep ({r['broad_ms']:.0f} ms; oneTBB replaced by an OpenMP stub), "
                       f"narrow phase = oracle port with OpenMP ({r['narrow_ms']:.0f} ms; the "
                       "reference has no CPU narrow phase); result checked equal to the GPU's")}
    elif rank == 0 and not args.no_cpu_baseline:
        line["cpu_baseline"] = {
            "value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "reference",
            "sample": ("not run inside this line: one CPU step of this workload takes minutes "
                       "(O(N^5/3) one-axis sweep on a dense pile); see `bench.py --impl reference`, "
                       "which times a named 1/16 sample scene")}
    if rank == 0 and world == 1 and not args.no_ref_cuda and args.workload in ("small", "c1", "c2"):
        line["reference_cuda"] = ref_cuda_subprocess(args.workload)
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- one JSON line per run (see the measurement contract in DESIGN.md).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c1|c3|c4|small]
  python bench.py --impl reference ...      # the reference's CPU path on the host cores

A "step" = one full CCD pass (AABB build -> sort -> sweep -> Tight-Inclusion narrow phase
-> earliest TOI, vertex-face then edge-edge) over one synthetic two-frame scene.
  value : ms/step with the mesh already resident in HBM (sccd_ccd)
  e2e   : ms/step through the host-pointer entry point (sccd_ccd_host): pinned host
          buffers -> H2D -> pipeline -> TOI back on the host, all inside the timed region
N > 1 : one process per GPU (torchrun); every rank sweeps its owner slice of the sorted
        lists and solves the pairs it found; NCCL all-reduce(min) of the TOI ("strong").
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


def run_reference(args, rank, world):
    """Reference arm: the reference's own CPU implementation of the path on the host cores
    (oracle/_ref when it was built, else the oracle port), all host threads.  Config 4 is too
    big for a CPU step of a few minutes, so its step is a BOUNDED SAMPLE: the same pile
    generator with 1/16 of the instances (same density), time scaled by 16 -- stated in
    `sample`."""
    if rank != 0:
        return
    from _pkg import load_package
    from oracle import orc
    sccd = load_package()
    scale = 1
    if args.workload == "c4":
        scale = 16
        scene = sccd.scenes.scene_c4(n_inst=83_000 // scale)
        desc = make_desc("c4")
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
    ms = statistics.mean(s["ms"] for s in steps) * scale
    cores = orc.lib().orc_num_threads()
    what = (f"{len(steps)} full step(s) of the workload" if scale == 1 else
            f"{len(steps)} step(s) of a 1/{scale} sample of the workload (same generator and "
            f"density, {scene['V0'].shape[0] + scene['E'].shape[0] + scene['F'].shape[0]} boxes), "
            f"time x{scale}")
    sample = (f"{what} (time-bounded to ~{int(budget_s)} s); "
              f"broad phase = {steps[0]['kind']} CPU sort_and_sweep "
              f"({statistics.mean(s['broad_ms'] for s in steps):.0f} ms, oneTBB replaced by an OpenMP stub), "
              f"narrow phase = oracle port with OpenMP ({statistics.mean(s['narrow_ms'] for s in steps):.0f} ms; "
              "the reference has no CPU narrow phase)")
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


def ref_cuda_subprocess(workload, timeout_s=240):
    """Reference CUDA ccd() on the same box, in a subprocess so a crash/hang cannot take the
    bench down.  Returns dict or {'unavailable': why}."""
    code = f"""
import sys, json
sys.path.insert(0, {ROOT!r})
from _pkg import load_package
from oracle import orc
import bench
sccd = load_package()
if orc.ref_cuda(False) is None:
    print(json.dumps({{"unavailable": "oracle/_ref/libref_sccd_cuda.so not built"}})); sys.exit(0)
scene, _ = bench.make_scene(sccd.scenes, {workload!r})
out = []
for i in range(3):
    r = orc.ref_cuda_ccd(scene, **bench.PARAMS)
    out.append(r["ms"])
b = orc.ref_cuda_broad_phase(scene, want_pairs=False)
print(json.dumps({{"ccd_ms": out, "toi": r["toi"], "broad_ms": b["ms"], "n_vf": b["n_vf"], "n_ee": b["n_ee"]}}))
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
                    help="c1|c2|c3|c4|small; default: c2 on one GPU (config 2, the scene the "
                         "metric is quoted on), c4 (config 4, ~50M AABBs) on several")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload is None:
        args.workload = "c2" if max(world, args.gpus) == 1 else "c4"
    if args.impl == "reference":
        return run_reference(args, rank, world)
    args.warmup = max(args.warmup, 3)

    import numpy as np
    import torch
    import torch.distributed as dist
    from _pkg import load_package
    sccd = load_package()
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    scene, desc = make_scene(sccd.scenes, args.workload)
    nV, nE, nF = scene["V0"].shape[0], scene["E"].shape[0], scene["F"].shape[0]

    stream = torch.cuda.current_stream().cuda_stream
    ctx = sccd.Context(local, stream)
    ctx.upload_mesh(scene["V0"], scene["V1"], scene["E"], scene["F"])
    sharded = sccd.multigpu.ShardedCCD(ctx) if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    toi_t = torch.zeros(1, dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        if world > 1:   # sharded sweep, pair rebalancing, NCCL min-TOI (multigpu.py)
            ctx.reset_stats()
            return sharded.ccd(**PARAMS)
        return ctx.ccd(**PARAMS)

    # pinned host copies for the end-to-end arm
    pinned = {k: torch.from_numpy(np.ascontiguousarray(v.T)).pin_memory() for k, v in scene.items()}

    packed = sccd.multigpu.pack_mesh(scene["V0"], scene["V1"], scene["E"], scene["F"], world) \
        if world > 1 else None

    def step_e2e():
        if world > 1:
            # every rank copies 1/world of the host mesh over its own PCIe link; the slices
            # are all-gathered over NVLink (multigpu.gather_mesh)
            ctx.reset_stats()
            sharded.upload_mesh_host(packed[0], packed[1], (nV, nE, nF))
            return sharded.ccd(**PARAMS)
        return ctx.ccd_host(pinned["V0"].data_ptr(), pinned["V1"].data_ptr(), pinned["E"].data_ptr(),
                            pinned["F"].data_ptr(), sizes=(nV, nE, nF), **PARAMS)

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
        ms = sum(per_step) / steps
        timed.last_steps = per_step
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, toi, stats, clocks

    single_ms = None
    if world > 1:
        # the same scene on ONE GPU (every rank, unsharded, untimed except on rank 0) so the
        # line carries its own strong-scaling reference
        single = ctx.ccd(**PARAMS)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(2):
            ctx.ccd(**PARAMS)
        torch.cuda.synchronize()
        single_ms = (time.perf_counter() - t0) / 2 * 1e3
        sharded = sccd.multigpu.ShardedCCD(ctx)   # re-arm the shard after the unsharded runs
        barrier()
    ms, toi, stats, clocks = timed(step_resident, args.steps, args.warmup, sampler=True)
    ms_steps = list(timed.last_steps)
    rank_stage_ms = None
    if world > 1:
        assert toi == single, ("sharded TOI differs from the single-GPU TOI", toi, single)
        sharded.profile = True            # one extra untimed step with per-stage events
        step_resident()
        sharded.profile = False
        rank_stage_ms = sharded.last.pop("ms", None)
        gathered = [None] * world
        dist.all_gather_object(gathered, rank_stage_ms)
        rank_stage_ms = gathered
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

    def avg(key, idx=None):
        vals = [s[key] if idx is None else s[key][idx] for s in stats]
        return float(sum(vals)) / len(vals)

    # whole-job counts (sum over ranks)
    def allsum(x):
        if world == 1:
            return x
        t = torch.tensor(x, dtype=torch.float64, device="cuda")
        dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    n_pairs = allsum([avg("n_pairs", 0), avg("n_pairs", 1)])
    n_checks = allsum([avg("n_box_checks", 0), avg("n_box_checks", 1)])
    n_cand = allsum([avg("n_candidates", 0), avg("n_candidates", 1)])
    n_boxes = [nV + nF, nE]
    k_ms = {
        "boxes": avg("ms_k_boxes"), "gather": avg("ms_k_gather"),
        "sweep_count_vf": avg("ms_k_sweep_count", 0), "sweep_count_ee": avg("ms_k_sweep_count", 1),
        "sweep_fill_vf": avg("ms_k_sweep_fill", 0), "sweep_fill_ee": avg("ms_k_sweep_fill", 1),
        "narrow_vf": avg("ms_k_narrow", 0), "narrow_ee": avg("ms_k_narrow", 1),
    }
    stage_ms = {"build": avg("ms_build"), "sort": avg("ms_sort"), "sweep_vf": avg("ms_sweep", 0),
                "sweep_ee": avg("ms_sweep", 1), "narrow_vf": avg("ms_narrow", 0),
                "narrow_ee": avg("ms_narrow", 1), "total_device": avg("ms_total")}
    # algorithmic bytes per launch.  The sweep reads each box once
    # (64 B) in the pass that finds the pairs and writes each pair once (8 B) in the pass that
    # places them; on N > 1 GPUs a rank's sweep / gather only covers its own share of the boxes.
    loc_pairs = [avg("n_pairs", 0), avg("n_pairs", 1)]
    share = 1.0 / world
    alg_bytes = {
        "boxes": 48 * nV + 8 * nE + 12 * nF + 64 * (nV + nE + nF),
        "gather": 2 * 64 * (nV + nE + nF) * share,
        "sweep_count_vf": 64 * n_boxes[0] * share, "sweep_count_ee": 64 * n_boxes[1] * share,
        "sweep_fill_vf": 8 * loc_pairs[0], "sweep_fill_ee": 8 * loc_pairs[1],
        "narrow_vf": (8 + 192 + 8) * loc_pairs[0], "narrow_ee": (8 + 192 + 8) * loc_pairs[1],
    }
    dom = max(k_ms, key=lambda k: k_ms[k])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = alg_bytes[dom] / (k_ms[dom] * 1e-3) / 1e9 if k_ms[dom] > 0 else 0.0
    traffic = None
    try:  # per-launch DRAM bytes of the same kernel from the committed ncu capture, if any
        traffic = json.load(open(os.path.join(ROOT, "profiles", "dram_traffic.json"))).get(
            f"{args.workload}:{dom}")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6650 GB/s",
                "algorithmic_bytes_per_launch": alg_bytes[dom], "kernel_ms": k_ms[dom],
                "all_kernels": {k: {"ms": k_ms[k], "GBps": (alg_bytes[k] / (k_ms[k] * 1e-3) / 1e9
                                                             if k_ms[k] > 0 else 0.0)} for k in k_ms}}
    narrow_ms = k_ms["narrow_vf"] + k_ms["narrow_ee"]
    # FP64 work of the narrow phase (96 / 84 arithmetic instr per box check) against
    # the FP64 pipe rate measured on this device by a register-only DFMA micro-benchmark
    fp64_instr = 96 * n_checks[0] + 84 * n_checks[1]
    dfma_peak = ctx.measure_fp64_peak() * world
    fp64 = {"bound": "fp64 pipe", "kernel": "narrow_vf + narrow_ee",
            "achieved": (fp64_instr / (narrow_ms * 1e-3)) if narrow_ms > 0 else None,
            "peak": dfma_peak, "unit": "thread-level FP64 instr/s (peak: measured DFMA/s)",
            "frac": (fp64_instr / (narrow_ms * 1e-3) / dfma_peak) if narrow_ms > 0 else None,
            "note": "96 (VF) / 84 (EE) FP64 arithmetic instructions per box check, min/max and "
                    "compares not counted"}
    line = {
        "metric": METRIC, "value": ms, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": False, "scaling": "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc, "n_vertices": nV, "n_edges": nE, "n_faces": nF,
                   "l2": "256 MiB device write between timed steps (flush)", **PARAMS,
                   "parallelism": f"cell-range shards x{world}" if world > 1 else "single GPU"},
        "clocks": clocks,
        "e2e": {"value": e2e_ms, "unit": UNIT,
                "h2d_bytes_per_step": 2 * 24 * nV + 8 * nE + 12 * nF, "d2h_bytes_per_step": 8,
                "note": ("whole job: each rank copies 1/N of the mesh H2D, NCCL all-gather of the "
                         "slices" if world > 1 else "pinned host mesh -> sccd_ccd_host")},
        "e2e_vertices_only": e2e_v,
        "gpu_launches": int(avg("n_launches")) * args.steps,
        "roofline": roofline, "roofline_fp64": fp64,
        "toi": toi, "n_pairs": n_pairs,
        "pairs_per_rank": (sharded.last if sharded else None),
        "stage_ms_per_rank": rank_stage_ms, "ms_steps_rank0": ms_steps,
        "single_gpu_ms_same_workload": single_ms,
        "speedup_vs_single_gpu": (single_ms / ms if single_ms else None), "n_prefilter_survivors": n_cand, "n_box_checks": n_checks,
        "narrow_queries_per_s": (sum(n_pairs) / (narrow_ms * 1e-3)) if narrow_ms > 0 else None,
        "narrow_box_checks_per_s": (sum(n_checks) / (narrow_ms * 1e-3)) if narrow_ms > 0 else None,
        "narrow_fp64_instr_per_s": (fp64_instr / (narrow_ms * 1e-3)) if narrow_ms > 0 else None,
        "stage_ms": stage_ms, "kernel_ms": k_ms,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import orc
        r = cpu_step(orc, scene)
        assert r["toi"] == toi and r["n_pairs"] == [int(n_pairs[0]), int(n_pairs[1])], \
            ("GPU result differs from the CPU baseline", r["toi"], toi, r["n_pairs"], n_pairs)
        line["cpu_baseline"] = {
            "value": r["ms"], "unit": UNIT, "cores": orc.lib().orc_num_threads(), "kind": r["kind"],
            "sample": (f"1 full step of the same workload: broad phase = {r['kind']} CPU "
                       f"sort_and_sweep ({r['broad_ms']:.0f} ms; oneTBB replaced by an OpenMP stub), "
                       f"narrow phase = oracle port with OpenMP ({r['narrow_ms']:.0f} ms; the "
                       "reference has no CPU narrow phase); result checked equal to the GPU's")}
    if rank == 0 and world == 1 and not args.no_ref_cuda:
        line["reference_cuda"] = ref_cuda_subprocess(args.workload)
    if rank == 0:
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Headline benchmark: Monte-Carlo shots/s on the [[144,12,12]] BB code (10 rounds, sliding window W=5 F=3, p=1e-3).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--shots S] [--precision f64|f32]

One process per GPU (torchrun sets RANK/LOCAL_RANK/WORLD_SIZE); a step is one pass of the hot path -- sample S shots,
decode them through the sliding window, count logical errors -- with everything resident on the device.  Shots are
partitioned by global shot index (weak scaling: S per GPU per step); there is no data-path collective, only a final
all-reduce of the error counters and a MAX of the device times.  Rank 0 prints ONE JSON line.

  value     shots/s, whole job, CUDA-event time on the launching stream (max over ranks)
  e2e       same metric through the drop-in calls get_stim_mem_result -> sliding_window_bposd_circuit_mem with host
            numpy buffers (D2H of the detection events, H2D again for decoding, D2H of the predictions) inside the timed region
  roofline  dominant kernel (BP).  Its messages never leave shared memory, so the binding roofline is the SM front end:
            bound "issue" = warp-instructions issued per second (instructions per edge-iteration from the committed ncu
            capture profiles/bp_inst_model.json x the edge-iterations this run executed / CUDA-event time of the BP launches)
            against 148 SMs x 4 schedulers x the SM clock sampled during the run.  The HBM message-streaming model of
            SURVEY 8(d) is kept beside it as hbm_model (frac > 1 there only says the messages are on chip), with the real
            DRAM traffic per launch (ncu) and the compulsory I/O bytes.
  cpu_baseline  the oracle's C restatement of the same path (OpenMP, all host cores) on a bounded sample, rank 0, N=1

--impl reference times the reference's CPU path alone: the real stim + ldpc wheels under the unmodified reference package
(baseline/run_reference.py, kind "reference") when they import; on this image they do not exist (not in the wheelhouse, no
network), so the arm falls back to the oracle's C port of the same algorithm (kind "port") and says so.
"""
from __future__ import annotations

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

WORKLOAD = "bb144_r10_p1e-3"
HEADLINE_WORKLOAD = WORKLOAD
W, F = 5, 3
BP_KW = dict(max_iter=10, osd_order=0, bp_method="minimum_sum", schedule="parallel", osd_method="osd_0")
SEED = 20260101
METRIC = "Monte-Carlo shots/sec, [[144,12,12]] BB 10-round window p=1e-3"


def load_workload():
    g = os.path.join(ROOT, "tests", "golden", "circuits")
    if os.path.exists(os.path.join(g, WORKLOAD + ".stim")):
        with open(os.path.join(g, WORKLOAD + ".stim")) as f:
            text = f.read()
    else:                                                  # the large circuits are committed compressed
        import gzip
        with gzip.open(os.path.join(g, WORKLOAD + ".stim.gz"), "rb") as f:
            text = f.read().decode()
    with open(os.path.join(g, WORKLOAD + ".json")) as f:
        meta = json.load(f)
    if RATE[0] is not None:
        # threshold sweeps: the builders write every noise rate as "%.10f" of the one p of ErrorModel(p, p, p, p), so another rate
        # of the same circuit is a substitution of that literal
        old, new = "(%.10f)" % float(meta["p"]), "(%.10f)" % RATE[0]
        assert old in text, "fixture rate literal not found"
        text = text.replace(old, new)
    hz = np.zeros(meta["hz_shape"], dtype=np.uint8)
    for i, r in enumerate(meta["hz_rows"]):
        hz[i, r] = 1
    lz = np.zeros(meta["lz_shape"], dtype=np.uint8)
    for i, r in enumerate(meta["lz_rows"]):
        lz[i, r] = 1
    return text, hz, lz


def config(shots, precision):
    return {"workload": "%s %s, W=%d F=%d, %s %s BP max_iter=%d + %s order %d" % (
                WORKLOAD, "custom circuit" if WORKLOAD.startswith("bb") else "circuit", W, F, BP_KW["bp_method"], "flooding" if BP_KW["schedule"] == "parallel" else "serial", BP_KW["max_iter"],
                BP_KW["osd_method"], BP_KW["osd_order"]),
            "shots_per_step_per_gpu": int(shots), "precision": precision, "seed": SEED, "rate_override": RATE[0],
            "l2": "flushed between timed steps (256 MiB write); per-step message/LLR working set also exceeds L2"}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_arm(shots, nthreads=0):
    """The oracle's C restatement of the path (frame sampler + window loop with BP/OSD-0 in fp64), OpenMP over shots."""
    from oracle import cref, dem as odem, stimtext, windows as owin
    text, hz, lz = load_workload()
    fc = stimtext.parse_flat(text)
    wins = owin.plan(odem.analyze(fc), hz.shape[0], W, F)
    # all host cores, whatever OMP_NUM_THREADS says (torchrun sets it to 1 for its workers)
    threads = nthreads or max(cref.num_threads(), os.cpu_count() or 1)
    t0 = time.perf_counter()
    det, obs = cref.sample(fc, SEED, 0, shots, nthreads=threads)
    pred, stats = cref.sw_decode(wins, hz.shape[0], lz.shape[0], det, nthreads=threads, max_iter=BP_KW["max_iter"],
                                 bp_method=BP_KW["bp_method"], schedule=BP_KW["schedule"], precision="f64",
                                 osd_method=BP_KW["osd_method"], osd_order=BP_KW["osd_order"])
    fails = int(np.any(pred != obs, axis=1).sum())
    dt = time.perf_counter() - t0
    return shots / dt, dt, threads, fails


REAL_REF_WHY = ["not probed"]
RATE = [None]


def real_reference_arm(args):
    """The unmodified reference on real stim + ldpc (baseline/run_reference.py), one process per host core; None when the
    wheels are not importable (then REAL_REF_WHY says why)."""
    script = os.path.join(ROOT, "baseline", "run_reference.py")
    try:
        pr = json.loads(subprocess.run([sys.executable, script, "--probe"], capture_output=True, text=True, timeout=300).stdout.strip().split("\n")[-1])
    except Exception as e:
        REAL_REF_WHY[0] = "probe failed: %r" % (e,)
        return None
    if not pr.get("available"):
        REAL_REF_WHY[0] = pr.get("why", "unknown")
        return None
    procs = os.cpu_count() or 1
    rates, times = [], []
    shots = 16 * procs
    for i in range(args.warmup + args.steps):
        out = subprocess.run([sys.executable, script, "--workload", WORKLOAD, "--shots", str(shots), "--procs", str(procs)],
                             capture_output=True, text=True, timeout=3600).stdout.strip().split("\n")[-1]
        r = json.loads(out)
        if i == 0:                               # size the sample for ~10 s per step
            shots = int(max(procs, min(200000, r["shots_per_s"] * 10.0)) // procs * procs)
        if i >= args.warmup:
            rates.append(r["shots_per_s"])
            times.append(r["wall_s"])
    rate, ms = float(np.mean(rates)), float(np.mean(times) * 1e3)
    sample = "%d shots per step, unmodified reference on stim+ldpc, one process per host core" % shots
    return {"impl": "reference", "metric": METRIC, "value": rate, "unit": "shots/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config(shots, "f64"),
            "cpu_baseline": {"value": rate, "unit": "shots/s", "cores": procs, "kind": "reference", "sample": sample},
            "e2e": {"value": rate, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}


def calibrated_cpu_sample(budget_s=12.0):
    rate, dt, threads, _ = cpu_arm(256)
    shots = int(max(256, min(200000, rate * budget_s)) // 64 * 64)
    return shots


def inst_model(precision):
    """profiles/bp_inst_model.json: warp-instructions and shared-memory wavefronts per edge-iteration of the BP kernel, taken
    from the committed ncu --set full capture (smsp__inst_executed.sum, l1tex__data_pipe_lsu_wavefronts_mem_shared.sum of one
    launch / the edge-iterations that launch ran); None when no model for this precision / schedule is committed."""
    try:
        with open(os.path.join(ROOT, "profiles", "bp_inst_model.json")) as f:
            m = json.load(f)
        key = "%s_%s_%s" % (precision, BP_KW["bp_method"], BP_KW["schedule"])
        e = m.get(key)
        # the counts belong to the kernel variant the capture ran: another workload picks another variant / layout
        return e if e and (WORKLOAD + " ") in e.get("source", "") else None
    except Exception:
        return None


def ncu_traffic(precision):
    """DRAM bytes of one BP launch (read + write) from the committed ncu --set full capture of the same launch geometry
    (65 536 shots per launch); None when no capture for this precision is committed."""
    path = os.path.join(ROOT, "profiles", "r02_bp_kernel_%s_ncu_full.txt" % precision)
    if not os.path.exists(path):
        path = os.path.join(ROOT, "profiles", "r01_bp_kernel_%s_ncu_full.txt" % precision)
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    total, seen = 0.0, 0
    try:
        with open(path) as f:
            for line in f:
                if line.startswith("dram__bytes_read.sum ") or line.startswith("dram__bytes_write.sum "):
                    u = line[line.index("[") + 1:line.index("]")]
                    total += float(line[line.index("]") + 1:].split()[0]) * unit[u]
                    seen += 1
    except Exception:
        return None
    return total if seen == 2 else None


def roofline(precision, agg, clocks, peaks):
    """The BP kernel against the limit that binds it (issue slots), with the shared-memory pipe and the SURVEY 8(d) HBM model."""
    bp_s = agg["bp_ms"] / 1e3
    bp_launches = max(1, agg["bp_launches"])
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    hbm_ach = agg["bp_alg_bytes"] / bp_s / 1e9 if bp_s > 0 else 0.0
    sm_mhz = clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0
    issue_peak = 148 * 4 * sm_mhz * 1e6 / 1e9              # G warp-instructions / s: one per scheduler and cycle
    smem_peak = 148 * sm_mhz * 1e6 / 1e9                   # G shared-memory wavefronts / s: one per SM and cycle
    m = inst_model(precision)                              # None unless a capture of this workload and setting is committed
    ei = agg.get("bp_edge_iters", 0.0)
    esz = 4 if precision == "f32" else 8
    io_bytes = max(0.0, agg["bp_alg_bytes"] - 4.0 * esz * ei)           # syndrome in + commit / carry out: the compulsory HBM bytes
    kernel = {("minimum_sum", "parallel"): "bp_kernel_ms2", ("product_sum", "parallel"): "bp_kernel_compact<PS>"}.get(
        (BP_KW["bp_method"], BP_KW["schedule"]), "bp_kernel_serial_slab")
    if WORKLOAD != HEADLINE_WORKLOAD and BP_KW["schedule"] == "parallel":
        kernel = "flooding BP kernel of this workload's windows (bp_kernel_ms2 / bp_kernel_compact / bp_kernel<global slab> by size)"
    out = {"kernel": kernel, "bound": "issue", "unit": "Gwarp-inst/s", "peak": issue_peak,
           "peak_source": "148 SMs x 4 schedulers x %.0f MHz (SM clock sampled during the timed region)" % sm_mhz,
           "achieved": None, "frac": None, "edge_iters_per_s": ei / bp_s if bp_s > 0 else 0.0,
           "ms_per_launch": agg["bp_ms"] / bp_launches,
           "traffic": None,
           "alg_io_bytes_per_launch": io_bytes / bp_launches,
           "hbm_model": {"bound": "hbm", "achieved": hbm_ach, "peak": hbm_peak, "unit": "GB/s", "model_frac": hbm_ach / hbm_peak,
                         "peak_source": "MEASURED_PEAKS.json (measured)" if peaks else "fallback 6650",
                         "alg_bytes_per_launch": agg["bp_alg_bytes"] / bp_launches,
                         "note": "SURVEY 8(d) message-streaming model (iterations x 4 x nnz x sizeof(msg) + syndrome in + commit/carry out). "
                                 "The messages are shared-memory resident: model_frac > 1 is not an HBM figure, `traffic` is the DRAM bytes "
                                 "one launch really moves (ncu)"}}
    if m and BP_KW["schedule"] == "parallel":
        # the committed capture ran 65 536 shots per BP launch; the posterior rows and syndromes a launch moves scale with its shots
        t, spl = ncu_traffic(precision), agg.get("windows", 0) / bp_launches
        out["traffic"] = t * spl / 65536.0 if (t and spl) else t
        out["traffic_note"] = "dram__bytes_read + write of a 65 536-shot BP launch (ncu --set full, profiles/) x shots per launch of this run / 65 536"
    if m and bp_s > 0:
        inst = m["warp_inst_per_edge_iter"] * ei
        out["achieved"] = inst / bp_s / 1e9
        out["frac"] = out["achieved"] / issue_peak
        out["warp_inst_per_edge_iter"] = m["warp_inst_per_edge_iter"]
        out["thread_inst_per_edge_iter"] = 32 * m["warp_inst_per_edge_iter"]
        out["inst_source"] = m.get("source")
        if m.get("smem_wavefronts_per_edge_iter"):
            wf = m["smem_wavefronts_per_edge_iter"] * ei / bp_s / 1e9
            out["smem"] = {"achieved": wf, "peak": smem_peak, "unit": "Gwavefront/s", "frac": wf / smem_peak}
    else:
        out["note"] = "no ncu instruction count committed for this workload / decoder setting (profiles/bp_inst_model.json): " \
                      "the issue fraction is not stated; edge_iters_per_s and hbm_model are live"
    return out


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of one GPU, sampled while the timed region runs: through NVML every 10 ms when pynvml is
    importable (a 0.6 s region gives ~50 samples), else through nvidia-smi (the profiling recipe's clocks line; ~5 samples / s)."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [x.strip() for x in vis.split(",") if x.strip()]
        if ids and all(x.isdigit() for x in ids) and index < len(ids):
            index = int(ids[index])                   # NVML / nvidia-smi count physical devices
        self.index, self.rows, self._halt, self.source = index, [], threading.Event(), "nvidia-smi"

    def _run_nvml(self):
        import pynvml as nv
        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = [nv.nvmlClocksThrottleReasonHwSlowdown, nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                nv.nvmlClocksThrottleReasonSwThermalSlowdown, nv.nvmlClocksThrottleReasonSwPowerCap]
        nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)          # fails here, before the first sample, if NVML cannot serve this GPU
        self.source = "nvml"
        while not self._halt.is_set():
            r = int(get_reasons(h))
            self.rows.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx)] + ["Active" if r & b else "Not Active" for b in bits])
            self._halt.wait(0.01)

    def run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            self.source = "nvidia-smi"
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(timeout=5)
        sm = [int(r[0]) for r in self.rows if r[0].isdigit()]
        mx = [int(r[1]) for r in self.rows if r[1].isdigit()]
        reasons = sorted({n for r in self.rows for n, v in zip(self.NAMES, r[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows), "source": self.source}


# ------------------------------------------------------------------------------------------------ main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--shots", type=int, default=262144, help="shots per step per GPU")
    ap.add_argument("--e2e-shots", type=int, default=262144)
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--workload", default=WORKLOAD, help="circuit fixture under tests/golden/circuits (default: the headline workload)")
    ap.add_argument("--bp-method", default=None, help="secondary points: minimum_sum (headline) | product_sum")
    ap.add_argument("--schedule", default=None, help="secondary points: parallel (headline) | serial")
    ap.add_argument("--osd-method", default=None, help="secondary points: osd_0 (headline) | osd_cs | osd_e")
    ap.add_argument("--osd-order", type=int, default=None)
    ap.add_argument("--lanes", type=int, default=0, help="concurrent sub-batches per device batch (0 = engine default)")
    ap.add_argument("--capacity", type=int, default=0, help="shots per device batch of the decoder (0 = engine default, 262144)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--rate", type=float, default=None, help="noise rate substituted for the fixture's (threshold sweeps)")
    ap.add_argument("--fanout", action="store_true", help="single process, e2e through the drop-in calls spread over every visible GPU")
    args = ap.parse_args()
    globals()["WORKLOAD"] = args.workload
    RATE[0] = args.rate
    if args.bp_method:
        BP_KW["bp_method"] = args.bp_method
    if args.schedule:
        BP_KW["schedule"] = args.schedule
    if args.osd_method:
        BP_KW["osd_method"] = args.osd_method
    if args.osd_order is not None:
        BP_KW["osd_order"] = args.osd_order
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank != 0:
            return
        real = real_reference_arm(args)
        if real is not None:
            print(json.dumps(real))
            return
        shots = calibrated_cpu_sample(8.0)
        vals = []
        for i in range(args.warmup + args.steps):
            rate, dt, threads, fails = cpu_arm(shots)
            if i >= args.warmup:
                vals.append((rate, dt))
        rate = float(np.mean([v[0] for v in vals])) if vals else 0.0
        ms = float(np.mean([v[1] for v in vals]) * 1e3) if vals else 0.0
        sample = "%d shots per step of the same workload (frame sampling + sliding-window BP/OSD-0 in fp64), OpenMP over shots" % shots
        print(json.dumps({"impl": "reference", "metric": METRIC, "value": rate, "unit": "shots/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config(shots, "f64"),
                          "cpu_baseline": {"value": rate, "unit": "shots/s", "cores": threads, "kind": "port", "sample": sample},
                          "e2e": {"value": rate, "unit": "shots/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "note": "real stim/ldpc unavailable (%s); this arm is the oracle's C port of the reference's CPU path "
                                  "(-O3 -march=x86-64-v3, OpenMP over shots, all host cores)" % REAL_REF_WHY[0]}))
        return

    import torch
    import torch.distributed as dist
    import quits_b200 as qb

    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    text, hz, lz = load_workload()
    # --gpus N means N devices, one per rank: the drop-in calls of this process stay on this rank's device (left alone they would
    # spread a call over every visible GPU; --fanout measures exactly that, from one process)
    qb.set_devices(None if args.fanout else [local])
    ctx = qb.Context.default(local)
    circuit = qb.Circuit(text)
    mc = qb.MonteCarlo(circuit, hz.shape[0], W, F, ctx=ctx, precision=args.precision, profile=True, lanes=args.lanes, capacity=args.capacity, **BP_KW)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    S = args.shots

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(i):
        # global shot range of this rank for step i: disjoint across ranks and steps
        shot0 = (i * world + rank) * S
        return mc.run(S, SEED, shot0)

    for i in range(args.warmup):
        step(i)
        flush.zero_()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    counts = np.zeros(1 + mc.K, dtype=np.uint64)
    agg = {}
    wall0 = time.perf_counter()
    for i in range(args.steps):
        c, st = step(args.warmup + i)
        counts += c
        for k, v in st.items():
            agg[k] = max(agg.get(k, 0), v) if k == "osd_max_columns" else agg.get(k, 0) + v
        flush.zero_()
        torch.cuda.synchronize()
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    dev_ms = agg["total_ms"]                      # CUDA events on the launching stream, summed over the K steps
    t = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    cnt = torch.tensor(counts.astype(np.int64), device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    dev_ms_max, wall_ms_max = (float(x) for x in t.tolist())
    total_shots = S * args.steps * world
    value = total_shots / (dev_ms_max / 1e3)

    line = None
    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        line = {"metric": METRIC, "value": value, "unit": "shots/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": args.precision, "data": "synthetic", "config": config(S, args.precision), "clocks": clocks,
                "gpu_launches": int(agg["bp_launches"] + agg["osd_launches"] + agg["frame_launches"] + agg["other_launches"]),
                "wall_ms_per_step": wall_ms_max / args.steps,
                "logical_errors": int(cnt[0].item()), "shots_total": int(total_shots),
                "decoder_stats": {"bp_edge_iters": agg.get("bp_edge_iters", 0.0), "bp_converged_frac": agg["bp_converged"] / max(1, agg["windows"]),
                                  "bp_iters_per_window": agg["bp_iterations"] / max(1, agg["windows"]),
                                  "osd_calls_per_shot": agg["osd_calls"] / max(1, agg["shots"]),
                                  "osd_columns_per_call": agg["osd_columns"] / max(1, agg["osd_calls"]),
                                  "osd_pivots_per_call": agg["osd_pivots"] / max(1, agg["osd_calls"]),
                                  "osd_max_columns": agg["osd_max_columns"], "osd_fast_path_overflows": agg["osd_overflows"]},
                "kernel_ms_per_step": {"frame": agg["frame_ms"] / args.steps, "bp": agg["bp_ms"] / args.steps,
                                       "osd": agg["osd_ms"] / args.steps},
                "roofline": roofline(args.precision, agg, clocks, peaks)}

    # ---- e2e through the drop-in API with host buffers (same stream of work, smaller batch)
    if not args.no_e2e:
        Se = args.e2e_shots
        dec32 = qb.SlidingWindowDecoder(circuit, hz.shape[0], W, F, ctx=ctx, precision="f32", **BP_KW) if args.precision == "f32" else None

        phase = [0.0, 0.0, 0.0]                      # host wall clock of the three calls over the timed e2e steps (this rank)

        def e2e_step(i):
            shot_seed = SEED + 1000 + i * world + rank
            ta = time.perf_counter()
            det, obs = qb.get_stim_mem_result(circuit, Se, seed=shot_seed)
            tb = time.perf_counter()
            pred = qb.sliding_window_bposd_circuit_mem(det, circuit, hz, lz, W, F, **BP_KW) if dec32 is None else dec32.decode(det)
            tc = time.perf_counter()
            n_bad = qb.count_logical_errors(obs, pred)
            td = time.perf_counter()
            phase[0] += tb - ta; phase[1] += tc - tb; phase[2] += td - tc
            return n_bad
        # the caller's pL reduction: the reference's idiom np.any((obs - pred) % 2, axis=1) (tests/test_sliding_window.py:83) costs
        # 32 ms on 262144 x 12 int64 -- a fifth of the step -- for an integer modulo of 0/1 values; count_logical_errors is the
        # same predicate over row blocks on the host's cores
        e2e_step(0)
        barrier()
        phase[:] = [0.0, 0.0, 0.0]
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for i in range(n_e2e):
            e2e_step(1 + i)
        barrier()
        dt = time.perf_counter() - t0
        print("e2e rank %d host ms per step: sample %.1f decode %.1f reduce %.1f" % (rank, 1e3 * phase[0] / n_e2e, 1e3 * phase[1] / n_e2e,
                                                                                     1e3 * phase[2] / n_e2e), file=sys.stderr)
        te = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        # same loop through the packed entry points of the C ABI (qb_sample_packed -> qb_sw_decode_packed): host buffers again, but
        # one bit per detector instead of one byte, and the comparison on packed words
        dec_p = dec32 if dec32 is not None else qb.SlidingWindowDecoder(circuit, hz.shape[0], W, F, ctx=ctx, **BP_KW)

        def e2e_packed_step(i):
            detp, obsp = circuit.sample(Se, SEED + 2000 + i * world + rank, packed=True)
            predp = dec_p.decode_packed(detp)
            return int(np.count_nonzero((predp ^ obsp).any(axis=1)))
        e2e_packed_step(0)
        barrier()
        t0 = time.perf_counter()
        for i in range(n_e2e):
            e2e_packed_step(1 + i)
        barrier()
        tp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        if rank == 0:
            D, K = circuit.num_detectors, circuit.num_observables
            DW, KW = (D + 63) // 64, (K + 63) // 64
            line["e2e_packed"] = {"value": Se * n_e2e * world / float(tp.item()), "unit": "shots/s", "h2d_bytes_per_step": Se * DW * 8,
                                  "d2h_bytes_per_step": Se * (DW + KW) * 8 + Se * KW * 8,
                                  "path": "Circuit.sample(packed=True) -> SlidingWindowDecoder.decode_packed (qb_sample_packed / qb_sw_decode_packed, "
                                          "u64 bit rows in host memory)"}
            if args.fanout:
                line["e2e_devices"] = qb.active_devices()
            line["e2e"] = {"value": Se * n_e2e * world / float(te.item()), "unit": "shots/s", "h2d_bytes_per_step": Se * D,
                           "d2h_bytes_per_step": Se * (D + K) + Se * K * 8, "shots_per_step_per_gpu": Se, "steps": n_e2e,
                           "host_ms_per_step": {"get_stim_mem_result": 1e3 * phase[0] / n_e2e, "sliding_window_bposd_circuit_mem": 1e3 * phase[1] / n_e2e,
                                                "count_logical_errors": 1e3 * phase[2] / n_e2e, "rank": 0},
                           "path": "get_stim_mem_result -> sliding_window_bposd_circuit_mem (numpy host buffers, host wall clock)"}
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            shots = calibrated_cpu_sample(12.0)
            rate, dt, threads, fails = cpu_arm(shots)
            line["cpu_baseline"] = {"value": rate, "unit": "shots/s", "cores": threads, "kind": "port",
                                    "sample": "%d shots of the same workload in %.1f s (oracle C port: frame sampler + window loop, fp64, OpenMP)" % (shots, dt)}
        print(json.dumps(line))
    # orderly teardown: drop the engine objects while the CUDA context is alive, leave NCCL, then return normally so that the
    # interpreter's exit hooks (the driver records which shared libraries the process loaded) run
    sys.stdout.flush()
    del mc
    qb.clear_decoder_cache()
    ctx.synchronize()
    del flush
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    sys.stdout.flush()
    sys.stderr.flush()


if __name__ == "__main__":
    main()

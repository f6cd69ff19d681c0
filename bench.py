#!/usr/bin/env python
"""bench.py — Rips H0+H1 diagrams/sec on the full-dataset-shaped EEG batch (BASELINE.json
configs[1]: 1,416 recordings x 5 bands x 60 windows of 47x47 float32 distance matrices ->
H0/H1 diagrams + persistence features + the 1416x220 feature table).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

INPUT = BASELINE.md §5(b)'s generator (tools/synth.py): recording `rec` is
default_rng(20261018 + rec) -> x = A(47x8)/sqrt(8) @ S(8x15000) + 0.5 E, ONE mixing matrix per
recording, band-passed into the five EEG bands (zero-phase Butterworth), cut into 60 one-second
windows, Pearson correlation, d = sqrt(2(1-r)), float32 -- built (untimed setup) by the repo's own
signal chain on the device for the repo arm and by the scipy/numpy restatement of the reference's
notebooks (oracle/signal_ref.py) for the reference arm.

One process per GPU (torchrun for N>1).  A step = one pass of the hot path over one batch.
`value`   : whole-job diagrams/s with the distance matrices resident in HBM (CUDA events on the
            launch stream, barrier + synchronize on both sides, max over ranks).  Weak scaling: every
            rank has its own 1,416 recordings; the strong-scaling partition of ONE 1,416-recording
            batch over the ranks is reported under `strong_scaling` when N > 1.
`e2e`     : the same metric through the C-ABI host entry tda_eeg_features_host: pinned host DENSE
            float32 47x47 matrices in (the configuration's stated input), host diagrams / features /
            table out, copies inside the timed region.  Next to it: the condensed-triangle entry
            (ripser's internal C++ format), the dense float64 entry (what compute_eeg_persistence
            receives in the reference) and a features-only read-back.
`roofline`: dominant kernel (rips_small tier 1) timed with CUDA events around its launches.
`cpu_baseline` / --impl reference : the CPU oracle (Ripser-style C++, OpenMP over diagrams) on the
            box's host cores, on a bounded sample of the same workload (the sample is named in
            `config` / `cpu_baseline.sample`).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

R_REC, N_BANDS, N_WIN, N_CH = 1416, 5, 60, 47
THRESH = 2.0
CAP1 = 128
METRIC = "rips_h0h1_diagrams_per_sec"
UNIT = "diagrams/s"
CPU_SAMPLE_RECORDINGS = 141  # ~10% of the workload: 42,300 diagrams (~15-30 core-seconds)
GENERATOR = ("BASELINE.md 5(b): default_rng(20261018+rec), x = A(47x8)/sqrt8 @ S(8x15000) + 0.5 E per recording, "
             "5-band zero-phase Butterworth, 60 windows of 250 samples, corrcoef, sqrt(2(1-r)), float32")
NCU_CAPTURE = os.path.join("profiles", "r02_rips_small_bench_ncu.json")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--recordings", type=int, default=R_REC, help="debug: smaller workload")
    ap.add_argument("--no-secondary", action="store_true", help="skip the audio / stress-cloud / raw-EEG side measurements")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.first = 0

    def mark(self):
        """samples taken before this call (spin-up under load) are not part of the timed region"""
        self.first = len(self.lines)

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines[self.first:]:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_reference_rate(D_host, threads, repeats=1):
    """diagrams/s of the CPU oracle on D_host (numpy (B,47,47) f32) with `threads` host threads."""
    from oracle import rips as orips
    orips.lib()
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        orips.rips_h01_batched(D_host, THRESH, cap1=CAP1, nthreads=threads)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return len(D_host) / best, best


def _cpu_distances_one(rec):
    """reference arm setup: one recording of the stated generator through the reference's own signal
    path on the CPU (oracle/signal_ref.py: scipy sosfiltfilt per channel, numpy corrcoef per window)"""
    import numpy as np
    from oracle import signal_ref
    from tools.synth import raw_eeg
    return signal_ref.eeg_distances(raw_eeg(rec), overlap=0.0).astype(np.float32)      # (5, 60, 47, 47)


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path for this metric.  ripser itself is not
    installable here (no wheel, no network), so this is the oracle port of its algorithm
    (oracle/rips_cpu.cpp) with all host threads, on a bounded SAMPLE of the workload: the first
    CPU_SAMPLE_RECORDINGS recordings of the stated generator.  Rank 0 only; no GPU is touched."""
    if rank != 0:
        return
    import multiprocessing as mp
    import numpy as np
    nrec = min(CPU_SAMPLE_RECORDINGS, args.recordings)
    B = nrec * N_BANDS * N_WIN
    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(min(cores, nrec)) as pool:      # (before any OpenMP runtime exists)
        D = np.concatenate(pool.map(_cpu_distances_one, range(nrec))).reshape(B, N_CH, N_CH)
    from oracle import rips as orips
    orips.lib()
    for _ in range(args.warmup):
        orips.rips_h01_batched(D[: max(B // 10, 1)], THRESH, cap1=CAP1, nthreads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        orips.rips_h01_batched(D, THRESH, cap1=CAP1, nthreads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = B / dt
    sample = f"{nrec} of {R_REC} recordings x {N_BANDS} bands x {N_WIN} windows = {B} diagrams per step"
    cfg = workload_config(args, world)
    cfg["sample"] = sample
    cfg["diagrams_per_step"] = B
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "per_core_value": val / cores},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }), flush=True)


def timed_ms(fn, n=3, warm=1):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def secondary_workloads(dev, x_raw):
    """Other BASELINE.json configurations, reported next to the headline (never part of `value`):
    audio Takens clouds (config c) and the 1,000 / 2,000-point scaling-stress clouds (config e)
    through the grid-cooperative engine, device-resident distance matrices, CUDA events; and the
    whole chain from raw EEG for all the recordings of the batch."""
    import torch
    from tda_eeg_audio_b200 import rips_h01_batched
    from tools.large_bench import takens_clouds
    out = {}
    for name, B, n in (("audio_takens_124pt", 8192, 124), ("audio_takens_248pt", 4096, 248),
                       ("stress_1000pt", 296, 1000), ("stress_2000pt", 148, 2000),
                       # bigger batches: the sweep of a cloud is serial and a few clouds of a batch take several
                       # times the mean, so the device-side queue has something to balance
                       ("stress_1000pt_1184clouds", 1184, 1000), ("stress_2000pt_592clouds", 592, 2000),
                       # the audio path at the batch size of a real run (1,416 recordings x 300 windows arrive as
                       # chunks of 32,768 clouds): a batch of 8,192 lasts as long as its heaviest cloud
                       ("audio_takens_124pt_32768clouds", 32768, 124), ("audio_takens_248pt_32768clouds", 32768, 248)):
        D = takens_clouds(B, n, dev=dev)
        buf = {}
        ms = timed_ms(lambda: rips_h01_batched(D, THRESH, cap1=4 * n, want_pairs=False, out=buf, engine="large"))
        out[name] = {"clouds": B, "points": n, "ms": round(ms, 3), "diagrams_per_s": round(B / ms * 1e3, 1),
                     "mean_h1_bars": round(float(buf["counts"][:, 1].float().mean()), 2),
                     "status_nonzero": int((buf["status"] != 0).sum())}
        del D, buf
        torch.cuda.empty_cache()
    try:
        out["raw_eeg_to_features"] = raw_eeg_workload(dev, x_raw)
    except Exception as exc:  # a side measurement must not cost the others
        out["raw_eeg_to_features"] = {"error": repr(exc)}
    return out


def raw_eeg_workload(dev, x, rec_chunk=None):
    """SURVEY.md §8(d) config (b) "end-to-end from raw EEG" for ALL the recordings of the batch: raw
    47-channel EEG (R, 47, 15000) float64, device-resident -> zero-phase band-pass into 5 bands ->
    60 windows -> correlation distances -> Rips H0+H1 -> features -> (R, 220) table.  CUDA events,
    per-stage times from the library's own event timers."""
    import torch
    from tda_eeg_audio_b200 import _lib, dsp, pipeline
    R = x.shape[0]
    D = torch.empty((R, N_BANDS, N_WIN, N_CH, N_CH), dtype=torch.float32, device=dev)
    st = {}

    def to_D():
        dsp.eeg_distances_from_raw(x, overlap=0.0, rec_chunk=rec_chunk, out=D)

    def run():
        to_D()
        return pipeline.eeg_features_from_distances(D, thresh=THRESH, cap1=CAP1, state=st)

    ms = timed_ms(run, n=3, warm=1)
    ms_D = timed_ms(to_D, n=3, warm=0)
    _lib.profile_enable(True)
    run()
    torch.cuda.synchronize()
    stages = {}
    for name in ("iir_sos_forward", "iir_sos_backward", "corrdist_mma", "rips_small_w1", "pers_features"):
        tms, _ = _lib.profile_query(name)
        stages[name] = round(tms, 3)
    _lib.profile_enable(False)
    samples = R * N_CH * 15000 * N_BANDS
    nwin = R * N_BANDS * N_WIN
    iir_ms = stages["iir_sos_forward"] + stages["iir_sos_backward"]
    res = {"recordings": R, "ms": round(ms, 3), "raw_eeg_to_distance_matrices_ms": round(ms_D, 3),
           "recordings_per_s": round(R / ms * 1e3, 1),
           "diagrams_per_s": round(nwin / ms * 1e3, 1), "stage_ms": stages, "input_bytes": x.numel() * 8,
           "iir": {"algorithmic_GBps": round(samples * 16 / iir_ms / 1e6, 1) if iir_ms else None,
                   "fp64_ops_per_s_no_fma_T": round(samples * 72 / iir_ms / 1e9, 2) if iir_ms else None},
           "gram": {"fp64_TFLOPs_full_square": round(nwin * 2 * N_CH * N_CH * 250 / stages["corrdist_mma"] / 1e9, 2)
                    if stages["corrdist_mma"] else None}}
    # the reference's own code path for the signal stages (scipy sosfiltfilt per channel, numpy corrcoef
    # per window: notebooks 1-2, restated in oracle/signal_ref.py and pinned by tests/golden/notebooks.npz)
    # on ONE recording, one process, as the reported CPU baseline of these stages
    try:
        import numpy as np
        from oracle import signal_ref
        x0 = x[0].cpu().numpy()
        t0 = time.perf_counter()
        ref = signal_ref.eeg_distances(x0, overlap=0.0)            # (5, 60, 47, 47) float64
        cpu_ms = (time.perf_counter() - t0) * 1e3
        got = D[0].double().cpu().numpy()
        off = ~np.eye(N_CH, dtype=bool)
        err = (np.abs(got - ref) / np.maximum(ref, 1e-30))[..., off].max()
        res["cpu_signal_stages"] = {"ms_per_recording": round(cpu_ms, 1), "cores": 1, "kind": "port",
                                    "max_rel_err_gpu_vs_cpu_distances": float(err)}
    except Exception as exc:
        res["cpu_signal_stages"] = {"error": repr(exc)}
    del D, st
    torch.cuda.empty_cache()
    return res


def workload_config(args, world):
    return {"workload": f"EEG batch: {args.recordings} recordings x {N_BANDS} bands x {N_WIN} windows of "
                        f"{N_CH}x{N_CH} f32 correlation-distance matrices -> Rips H0+H1 (thresh 2.0) + 11x2 "
                        f"features/window + {args.recordings}x220 table, per GPU",
            "generator": GENERATOR,
            "diagrams_per_gpu": args.recordings * N_BANDS * N_WIN, "parallelism": f"recordings sharded x{world}, "
            "NCCL allgather of the feature table", "l2": "inputs (3.75 GB) >> L2 (126 MB), no flush needed",
            "cap1": CAP1}


def host_topology():
    nodes = []
    try:
        nodes = sorted(d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit())
    except OSError:
        pass
    return {"cpu_count": os.cpu_count(), "numa_nodes": len(nodes) or None,
            "affinity_cpus": len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else None}


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    from tda_eeg_audio_b200 import _lib, pipeline
    from tda_eeg_audio_b200.dist import shard_range
    from tda_eeg_audio_b200.rips import tier_counts
    from tools.synth import eeg_distance_matrices

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    # stdout carries exactly one JSON line: whatever libraries print meanwhile (NCCL announces its
    # version on stdout when NCCL_DEBUG is set) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    # every rank keeps to its own slice of the host cores (generator threads, pinned staging buffers
    # are first-touched from there)
    topo = host_topology()
    if world > 1 and hasattr(os, "sched_setaffinity"):
        cpus = sorted(os.sched_getaffinity(0))
        per = max(len(cpus) // world, 1)
        mine = cpus[local_rank * per:(local_rank + 1) * per] or cpus
        try:
            os.sched_setaffinity(0, mine)
        except OSError:
            pass
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    R, Bd, Wn, N = args.recordings, N_BANDS, N_WIN, N_CH
    B = R * Bd * Wn
    # ---- inputs of the stated generator, resident in HBM (setup, untimed); weak scaling: rank r
    #      holds recordings [r R, (r+1) R)
    t_setup = time.perf_counter()
    want_raw = rank == 0 and world == 1 and not args.no_secondary
    D, x_raw = eeg_distance_matrices(rank * R, R, dev, step=250, keep_raw=want_raw)
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t_setup
    state = {}
    gathered = torch.empty((world * R, Bd * 44), dtype=torch.float64, device=dev) if world > 1 else None

    def step():
        res = pipeline.eeg_features_from_distances(D, thresh=THRESH, cap1=CAP1, state=state)
        if world > 1:
            dist.all_gather_into_tensor(gathered, res["table"])
        return res

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        res = step()
    barrier()
    st_all = res["rips"]["status"]
    trunc = int((st_all & 1).sum().item())
    bad = int((st_all & 4).sum().item())
    mean_h1 = float(res["rips"]["counts"][:, 1].float().mean().item())
    max_h1 = int(res["rips"]["counts"][:, 1].max().item())
    n_bars = float(res["rips"]["counts"].sum().item())
    tiers = tier_counts(res["rips"], N)

    # ---- timed region (device-resident inputs)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # nvidia-smi needs a moment before its first sample: keep the GPU under the same load meanwhile.
    # A FIXED number of untimed steps on EVERY rank (step() contains a collective when world > 1).
    for _ in range(12):
        step()
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    launches = _lib.launch_count() - l0
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * B / (ms_per_step * 1e-3)

    # ---- strong scaling (N > 1): ONE batch of R recordings (the stated generator's recordings 0..R-1)
    #      partitioned over the ranks in contiguous ranges, all-gather of the (R, 220) table
    strong = None
    if world > 1:
        lo, hi = shard_range(R, rank, world)
        Ds, _ = eeg_distance_matrices(lo, hi - lo, dev, step=250)
        st_s = {}
        from tda_eeg_audio_b200.dist import allgather_rows

        def strong_step():
            r_ = pipeline.eeg_features_from_distances(Ds, thresh=THRESH, cap1=CAP1, state=st_s)
            return allgather_rows(r_["table"], R)

        for _ in range(3):
            tab = strong_step()
        barrier()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(args.steps):
            tab = strong_step()
        s1.record()
        barrier()
        ts = torch.tensor([s0.elapsed_time(s1)], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        s_ms = float(ts.item()) / args.steps
        strong = {"scaling": "strong", "recordings_total": R, "recordings_per_rank": hi - lo,
                  "diagrams_total": B, "ms_per_step": s_ms, "value": B / (s_ms * 1e-3), "unit": UNIT,
                  "table_rows_gathered": int(tab.shape[0]),
                  "note": "same 1,416-recording batch as N=1 (recordings 0..1415 of the generator), contiguous "
                          "recording ranges per rank, NCCL all-gather of the feature table inside the timed region"}
        del Ds, st_s

    # ---- roofline of the dominant kernel: CUDA events around its launches (separate pass)
    _lib.profile_enable(True)
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    k_ms, k_n = _lib.profile_query("rips_small_w1")     # tier 1 of 47-point windows (one-word masks)
    parts = {}
    for name in ("rips_small_w1", "rips_small_w2", "rips_small_w4", "rips_small_w64", "pers_features", "aggregate_windows"):
        tms, tn = _lib.profile_query(name)
        parts[name] = round(tms / max(args.steps, 1), 4)
    _lib.profile_enable(False)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    # 4 N^2 per matrix read + per emitted bar 2 x f32 (birth, death) written; the 8 B of
    # simplex indices per bar count only when the step writes them (it does not: want_pairs=False)
    bytes_per_bar = 8.0
    alg_bytes = 4.0 * N * N * B + bytes_per_bar * n_bars
    k_avg_ms = k_ms / max(k_n, 1)
    achieved = alg_bytes / (k_avg_ms * 1e-3) / 1e9 if k_avg_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "rips_small_kernel<1,false,47,12>", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "peak_source": "MEASURED_PEAKS.json (measured, burst copy bandwidth)" if peaks else "fallback 6650 GB/s",
                "kernel_ms": k_avg_ms, "algorithmic_bytes_per_launch": alg_bytes,
                "algorithmic_bytes_per_diagram": alg_bytes / B, "bytes_per_bar": bytes_per_bar,
                "kernel_ms_per_step": parts,
                "note": "formally HBM-scored; the kernel is shared-memory/issue bound (see DESIGN.md); `traffic` is "
                        "null because DRAM bytes are not measurable inside this run -- see `ncu_capture`"}
    # numbers of a COMMITTED ncu capture of the same kernel on the same workload, scaled to this batch:
    # constants read from a file, not measurements of this run
    ncu_capture = None
    try:
        ncu = json.load(open(os.path.join(ROOT, NCU_CAPTURE)))
        sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
        slots = 148 * 4 * sm_mhz * 1e6                      # warp-instruction issue slots per second
        ips = ncu["warp_instructions_per_window"] * B / (k_avg_ms * 1e-3)
        ncu_capture = {"source": NCU_CAPTURE + " (ncu --set full of this kernel on this workload, committed; "
                                               "NOT measured in this run)",
                       "dram_traffic_bytes_scaled_to_this_batch": ncu["traffic_bytes"] * B / ncu["windows"],
                       "warp_instructions_per_window": ncu["warp_instructions_per_window"],
                       "issue_slot_utilisation_with_this_runs_kernel_time": ips / slots}
    except Exception:
        pass

    # ---- e2e: pinned host buffers through the C-ABI host entries (H2D + D2H inside the timed region)
    from tda_eeg_audio_b200.rips import condense
    h_D = torch.empty((B, N, N), dtype=torch.float32, pin_memory=True)
    h_D.copy_(D.view(B, N, N))
    h_bd0 = torch.empty((B, N, 2), dtype=torch.float32, pin_memory=True)
    h_bd1 = torch.empty((B, CAP1, 2), dtype=torch.float32, pin_memory=True)
    h_cnt = torch.empty((B, 2), dtype=torch.int32, pin_memory=True)
    h_st = torch.empty((B,), dtype=torch.int32, pin_memory=True)
    h_feats = torch.empty((B, 2, 11), dtype=torch.float64, pin_memory=True)
    h_table = torch.empty((R, Bd * 44), dtype=torch.float64, pin_memory=True)
    g_host = torch.empty((world * R, Bd * 44), dtype=torch.float64, pin_memory=True) if world > 1 else None
    d2h_full = (h_bd0.numel() + h_bd1.numel()) * 4 + h_cnt.numel() * 4 + h_st.numel() * 4 + \
        (h_feats.numel() + h_table.numel()) * 8
    d2h_feats = h_cnt.numel() * 4 + h_st.numel() * 4 + (h_feats.numel() + h_table.numel()) * 8

    def measure_e2e(fn, name, h_in, layout, diagrams=True):
        def e2e_step():
            rc = fn(h_in.data_ptr(), R, Bd, Wn, N, THRESH, CAP1, h_bd0.data_ptr() if diagrams else None,
                    h_bd1.data_ptr() if diagrams else None, h_cnt.data_ptr(), h_st.data_ptr(), h_feats.data_ptr(),
                    h_table.data_ptr(), local_rank)
            if rc != 0:
                raise RuntimeError(f"{name} rc={rc}")
            if world > 1:
                tb = h_table.to(dev, non_blocking=True)
                dist.all_gather_into_tensor(gathered, tb)
                g_host.copy_(gathered, non_blocking=True)
                torch.cuda.synchronize()

        h_table.zero_()
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            e2e_step()
        barrier()
        sec = (time.perf_counter() - t0) / args.steps
        tt = torch.tensor([sec], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        sec = float(tt.item())
        h2d = h_in.numel() * h_in.element_size()
        ok = bool(torch.equal(h_table.to(dev), res["table"]))
        return {"value": world * B / sec, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h_full if diagrams else d2h_feats,
                "ms_per_step": sec * 1e3, "matches_device_path": ok, "h2d_gbs_per_rank": h2d / sec / 1e9,
                "input": layout, "returns": "diagrams + features + table" if diagrams else "features + table",
                "api": f"{name} (C-ABI, pinned host buffers, 3-stream chunk pipeline)"}

    e2e = measure_e2e(lib.tda_eeg_features_host, "tda_eeg_features_host", h_D,
                      "dense 47x47 float32 matrices (the configuration's stated input; HEADLINE entry)")
    e2e["features_only"] = measure_e2e(lib.tda_eeg_features_host, "tda_eeg_features_host", h_D,
                                       "dense 47x47 float32 matrices", diagrams=False)
    h_Dc = torch.empty((B, N * (N - 1) // 2), dtype=torch.float32, pin_memory=True)
    for b0 in range(0, B, 65536):
        h_Dc[b0:b0 + 65536].copy_(condense(D.view(B, N, N)[b0:b0 + 65536]))
    e2e["condensed"] = measure_e2e(
        lib.tda_eeg_features_condensed_host, "tda_eeg_features_condensed_host", h_Dc,
        "condensed upper triangle, 1081 float32 per window (the vector ripser.py builds internally for its C++ core)")
    del h_Dc
    if world == 1:
        h_D64 = torch.empty((B, N, N), dtype=torch.float64, pin_memory=True)
        h_D64.copy_(h_D)
        e2e["dense_float64"] = measure_e2e(
            lib.tda_eeg_features_f64_host, "tda_eeg_features_f64_host", h_D64,
            "dense 47x47 float64 matrices (the argument compute_eeg_persistence receives in the reference); "
            "symmetrise / clamp / float32 cast on the device")
        del h_D64
    # the host's own ceiling: a plain pinned host->device copy loop, all ranks at once
    try:
        nbytes = min(h_D.numel() * 4, 1 << 30)
        src = h_D.view(-1)[: nbytes // 4]
        dst = torch.empty_like(src, device=dev)
        dst.copy_(src, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            dst.copy_(src, non_blocking=True)
        barrier()
        sec = (time.perf_counter() - t0) / 4
        tt = torch.tensor([sec], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e["host_ceiling"] = {"plain_pinned_h2d_gbs_per_rank": nbytes / float(tt.item()) / 1e9,
                               "aggregate_gbs": world * nbytes / float(tt.item()) / 1e9, "ranks_copying_at_once": world,
                               "topology": topo}
        del dst
    except Exception as exc:
        e2e["host_ceiling"] = {"error": repr(exc)}

    # ---- CPU baseline on a bounded sample of the same matrices (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1:
        nrec = min(CPU_SAMPLE_RECORDINGS, R)
        nb = nrec * Bd * Wn
        Dh = h_D[:nb].numpy()
        cores = os.cpu_count() or 1
        rate, dt = cpu_reference_rate(Dh, cores)
        rate1, _ = cpu_reference_rate(Dh[: max(nb // 16, 1)], 1)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{nrec} of {R} recordings x {Bd} bands x {Wn} windows = {nb} diagrams "
                         f"({dt:.2f} s wall, oracle/rips_cpu.cpp, OpenMP dynamic)",
               "single_thread_value": rate1, "per_core_value": rate / cores}

    secondary = None
    if rank == 0 and world == 1 and not args.no_secondary:
        del h_D, h_bd0, h_bd1, h_feats
        try:
            secondary = secondary_workloads(dev, x_raw)
        except Exception as exc:  # never lose the headline line to a secondary measurement
            secondary = {"error": repr(exc)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world), "roofline": roofline, "ncu_capture": ncu_capture,
            "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks,
            "quality": {"mean_h1_bars": mean_h1, "max_h1_bars": max_h1, "h1_truncated_windows": trunc,
                        "internal_overflow": bad,
                        "windows_finished_per_tier": {f"W{k}": v for k, v in tiers.items()}},
            "strong_scaling": strong, "setup_s": round(t_setup, 1),
            "secondary": secondary,
        }
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

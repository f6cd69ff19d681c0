"""BASELINE.json configs[3] end to end on synthetic recordings: raw 47-channel EEG + raw 44.1 kHz audio
-> per-band EEG diagrams / 220-feature table, audio envelope -> Takens diagrams, matched EEG-audio
W_H0 / W_H1 per (recording, band) and the mismatched-audio control (W_H1 against the audio of the
subject's first recording of the other condition), recordings sharded over the ranks, results
all-gathered (NCCL).

    python tools/run_config_d.py [R_total]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/run_config_d.py [R_total]

Prints one JSON line with per-stage milliseconds (CUDA events, max over ranks).  The reference does
this with tda_eeg_classification_v2.py + tda_eeg_audio_comparison.py + matched_vs_mismatched.py."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from tda_eeg_audio_b200 import audio, dsp, pipeline, rips_h01_batched
from tda_eeg_audio_b200.dist import allgather_rows, gather_reference_diagrams, shard_range
from tda_eeg_audio_b200.wasserstein import wasserstein_batched


def main():
    R_total = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = shard_range(R_total, rank, world)
    R = hi - lo
    T_eeg, T_aud, n_sub = 15000, 2646000, 45
    # ---- synthetic raw inputs of this rank's recordings (SURVEY.md §8d generators, seed = recording id)
    x = torch.empty((R, 47, T_eeg), dtype=torch.float64, device=dev)
    a = torch.empty((R, T_aud), dtype=torch.float64, device=dev)
    t = torch.arange(T_aud, device=dev, dtype=torch.float64) / 44100.0
    for k, rec in enumerate(range(lo, hi)):
        g = torch.Generator(device=dev); g.manual_seed(20261018 + rec)
        A = torch.randn((47, 8), generator=g, device=dev, dtype=torch.float64) / 8 ** 0.5
        x[k] = A @ torch.randn((8, T_eeg), generator=g, device=dev, dtype=torch.float64) + \
            0.5 * torch.randn((47, T_eeg), generator=g, device=dev, dtype=torch.float64)
        ph = torch.rand(2, generator=g, device=dev, dtype=torch.float64) * 6.283185307179586
        a[k] = (1 + 0.6 * torch.sin(6.283185307179586 * 3.1 * t + ph[0]) + 0.3 * torch.sin(6.283185307179586 * 6.7 * t + ph[1])) * \
            torch.randn(T_aud, generator=g, device=dev, dtype=torch.float64)
    torch.cuda.synchronize()
    ev = {}

    def stage(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        ev[name] = (e0, e1)
        return out

    def run_once():
        if world > 1:
            dist.barrier()
        D = stage("eeg_raw_to_distances", lambda: dsp.eeg_distances_from_raw(x, overlap=0.75))     # (R, 5, 238, 47, 47)
        feat = stage("eeg_rips_features_table", lambda: pipeline.eeg_features_from_distances(D, cap1=128))
        env = stage("audio_resample_envelope", lambda: audio.audio_envelope_from_raw(a))
        aud = stage("audio_bands_takens_rips", lambda: pipeline.audio_diagrams_from_envelope(env, max_windows=15))
        Rr, nb, ns = aud["shape"]
        idx = torch.from_numpy(aud["idx"]).to(dev)
        er = stage("eeg_rips_selected_windows",
                   lambda: rips_h01_batched(D[:, :, idx].contiguous().view(-1, 47, 47), thresh=2.0, cap1=128, want_pairs=False))
        w0, w1 = stage("wasserstein_matched", lambda: pipeline.cross_wasserstein(er, aud["rips"]))
        matched0 = torch.nanmean(w0.view(R, nb, ns), dim=2)
        matched1 = torch.nanmean(w1.view(R, nb, ns), dim=2)

        def mismatched():
            ref = pipeline.mismatch_reference_recording(R_total, n_subjects=n_sub)
            wanted = sorted(set(int(v) for v in ref if v >= 0))
            items = nb * ns
            cap = aud["rips"]["bd1"].shape[1]
            bd, cnt = gather_reference_diagrams(aud["rips"]["bd1"].view(R, items, cap, 2), aud["rips"]["counts"][:, 1].contiguous().view(R, items),
                                                lo, hi, wanted)
            out = torch.full((R, nb), float("nan"), dtype=torch.float64, device=dev)
            if not wanted:
                return out
            pos = {w: i for i, w in enumerate(wanted)}
            rows = [k for k, rec in enumerate(range(lo, hi)) if ref[rec] >= 0]
            if not rows:
                return out
            ia = torch.cat([torch.arange(items, device=dev) + k * items for k in rows]).to(torch.int32)
            ib = torch.cat([torch.arange(items, device=dev) + pos[int(ref[lo + k])] * items for k in rows]).to(torch.int32)
            wm = wasserstein_batched(er["bd1"], er["counts"][:, 1], bd.view(-1, cap, 2), cnt.view(-1), ia, ib)
            out[torch.as_tensor(rows, device=dev)] = torch.nanmean(wm.view(len(rows), nb, ns), dim=2)
            return out
        mism1 = stage("wasserstein_mismatched(+exchange)", mismatched)
        res_local = torch.stack([matched0, matched1, mism1], dim=2)                                # (R, 5, 3)
        table, res = stage("allgather", lambda: (allgather_rows(feat["table"], R_total), allgather_rows(res_local, R_total)))
        return feat, table, res

    run_once()                      # warm-up pass (cuFFT plan, allocator, first-touch)
    torch.cuda.synchronize()
    ev.clear()
    feat, table, res = run_once()   # timed pass
    torch.cuda.synchronize()
    ms = {k: e0.elapsed_time(e1) for k, (e0, e1) in ev.items()}
    tms = torch.tensor(list(ms.values()), dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    if rank == 0:
        tot = float(tms.sum())
        r = res.cpu().numpy()
        print(json.dumps({"config": "d: raw EEG + raw audio -> features, diagrams, matched / mismatched Wasserstein",
                          "recordings": R_total, "n_gpus": world, "stage_ms": {k: round(float(v), 2) for k, v in zip(ms, tms)},
                          "total_ms": round(tot, 2), "recordings_per_s": round(R_total / tot * 1e3, 2),
                          "table_shape": list(table.shape), "table_finite": bool(torch.isfinite(table).all()),
                          "mean_W_H0_matched": float(np.nanmean(r[:, :, 0])), "mean_W_H1_matched": float(np.nanmean(r[:, :, 1])),
                          "mean_W_H1_mismatched": float(np.nanmean(r[:, :, 2])) if np.isfinite(r[:, :, 2]).any() else None,
                          "h1_truncated_windows": int((feat["rips"]["status"] & 1).sum())}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

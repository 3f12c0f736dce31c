#!/usr/bin/env python
"""Observation-encoder throughput and step+encode throughput (report step-only,
step+gray and step+default-symbol separately, each against the HBM roofline).

  python tools/bench_encode.py [--envs 65536] [--steps 200]

One JSON line per observation setting. `encode` = k_encode alone (HBM-write bound, algorithmic bytes
per env from SURVEY.md §8d); `step+encode` = rg_step followed by rg_encode every step, as
DeviceRogueEnv does for a trainer. Experiment script: not the driver's bench contract."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rogue-gym_b200", "python"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=200)
    args = ap.parse_args()
    import numpy as np
    import torch
    from rogue_gym_python import _cabi
    from rogue_gym_python.rollout import Shard, synthetic_actions

    peak = 6550.1
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    n, K = args.envs, args.steps
    sh = Shard("{}", 0, n)
    stream = torch.cuda.ExternalStream(sh.stream())
    acts = torch.from_numpy(np.stack([synthetic_actions(t, sh.env_ids) for t in range(K + 50)])).cuda()
    for t in range(50):
        sh.step_device(acts.data_ptr() + t * n)
    sh.quiesce()
    sh.sync()
    C_ = sh.W * sh.H
    settings = [("gray, no status", 0, 0, 0), ("gray + 9 status + hist", 0, 0x1FF, 1),
                ("symbol + 9 status (reference default ImageSetting())", 1, 0x1FF, 0)]
    # the compact observation (rg_encode_compact): symbol ids + status vector (+ visited map)
    for name, with_hist in (("compact: symbol ids u8 + status i32[9]", False), ("compact + visited map", True)):
        sym = torch.empty((n, sh.H, sh.W), dtype=torch.uint8, device="cuda")
        stat = torch.empty((n, 9), dtype=torch.int32, device="cuda")
        hist_t = torch.empty((n, sh.H, sh.W), dtype=torch.uint8, device="cuda") if with_hist else None

        def encode_c():
            _cabi.check(sh.L.rg_encode_compact(sh.h, sym.data_ptr(), stat.data_ptr(), hist_t.data_ptr() if with_hist else None), sh.h)

        for _ in range(3):
            encode_c()
        sh.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record(stream)
        for _ in range(reps):
            encode_c()
        e1.record(stream)
        sh.sync()
        enc_ms = e0.elapsed_time(e1) / reps
        bytes_env = 2 * C_ + 40 + 36 + ((C_ // 8 + C_) if with_hist else 0)
        enc_gbs = bytes_env * n / (enc_ms * 1e-3) / 1e9
        e0.record(stream)
        for t in range(50, 50 + K):
            sh.step_device(acts.data_ptr() + t * n)
            encode_c()
        sh.quiesce()
        e1.record(stream)
        sh.sync()
        se_ms = e0.elapsed_time(e1) / K
        print(json.dumps({
            "observation": name, "envs": n,
            "encode": {"ms": enc_ms, "envs_per_sec": n / (enc_ms * 1e-3), "bytes_per_env": bytes_env,
                       "roofline": {"bound": "hbm", "achieved": enc_gbs, "peak": peak, "unit": "GB/s", "frac": enc_gbs / peak}},
            "step+encode": {"ms_per_step": se_ms, "env_steps_per_sec": n / (se_ms * 1e-3),
                            "roofline_frac": (bytes_env + 7858) * n / (se_ms * 1e-3) / 1e9 / peak},
        }), flush=True)
    for name, mode, flag, hist in settings:
        ch = sh.L.rg_encode_channels(sh.h, mode, flag, hist)
        out = torch.empty((n, ch, sh.H, sh.W), dtype=torch.float32, device="cuda")
        cc = C.c_int()

        def encode():
            _cabi.check(sh.L.rg_encode(sh.h, mode, flag, hist, out.data_ptr(), C.byref(cc)), sh.h)

        for _ in range(3):
            encode()
        sh.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record(stream)
        for _ in range(reps):
            encode()
        e1.record(stream)
        sh.sync()
        enc_ms = e0.elapsed_time(e1) / reps
        bytes_env = C_ + 40 + (C_ // 8 if hist else 0) + 4 * C_ * ch
        enc_gbs = bytes_env * n / (enc_ms * 1e-3) / 1e9
        e0.record(stream)
        for t in range(50, 50 + K):
            sh.step_device(acts.data_ptr() + t * n)
            encode()
        sh.quiesce()
        e1.record(stream)
        sh.sync()
        se_ms = e0.elapsed_time(e1) / K
        print(json.dumps({
            "observation": name, "channels": ch, "envs": n,
            "encode": {"ms": enc_ms, "envs_per_sec": n / (enc_ms * 1e-3), "bytes_per_env": bytes_env,
                       "roofline": {"bound": "hbm", "achieved": enc_gbs, "peak": peak, "unit": "GB/s", "frac": enc_gbs / peak}},
            "step+encode": {"ms_per_step": se_ms, "env_steps_per_sec": n / (se_ms * 1e-3),
                            "roofline_frac": (bytes_env + 7858) * n / (se_ms * 1e-3) / 1e9 / peak},
        }), flush=True)
        del out
        torch.cuda.empty_cache()
    sh.close()


if __name__ == "__main__":
    main()

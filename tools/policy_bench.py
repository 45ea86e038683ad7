#!/usr/bin/env python
"""Development aid: policy-forward time at rollout batch sizes (CUDA events on the launching stream, L2 flushed between
iterations), as TFLOP/s against the measured bf16 peak.  8.68 MFLOP per observation."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mujoco_rl_manipulate_unknown_objects_b200.policy import GripperPolicy

FLOP_PER_OBS = 2 * (225 * 256 * 32 + 36 * 512 * 64 + 16 * 576 * 64 + 1024 * 512 + 514 * 256 + 256 * 256 + 256 * 12)
peak = 1408.4
try:
    peak = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
except Exception:
    pass
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for n in [int(x) for x in (sys.argv[1:] or ["4096", "8192", "16384"])]:
    pol = GripperPolicy(max_envs=n)
    obs = torch.randint(0, 256, (n, 5, 64, 64), dtype=torch.uint8, device="cuda")
    out = torch.empty((n, 6), device="cuda")
    for _ in range(5):
        pol(obs, out=out)
    ts = []
    for _ in range(20):
        if not os.environ.get("NOFLUSH"):
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); pol(obs, out=out); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = sorted(ts)[len(ts) // 2]
    tf = n * FLOP_PER_OBS / (ms * 1e-3) / 1e12
    print("N=%6d  policy forward %.3f ms  %.1f TFLOP/s  (%.1f%% of %.0f TFLOP/s bf16)  obs read %.0f GB/s" % (n, ms, tf, 100 * tf / peak, peak, n * 20480 / (ms * 1e-3) / 1e9))
    pol.close()

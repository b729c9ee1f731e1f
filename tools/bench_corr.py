#!/usr/bin/env python
"""Micro-benchmark of the correspondence streaming kernel (k_corr) alone: CUDA-event time per launch with a 256 MB
L2 flush between launches.  For kernel-only durations run it under
    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:k_corr ...
    python tools/bench_corr.py [B C]..."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynhor_b200.corr import _CorrSums, plan  # noqa: E402

cases = [(300, 10000), (512, 50000)]
if len(sys.argv) > 2:
    a = [int(x) for x in sys.argv[1:]]
    cases = list(zip(a[0::2], a[1::2]))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for B, C in cases:
    g = torch.Generator(device="cuda").manual_seed(0)
    rec = torch.rand(B, C, 6, device="cuda", generator=g)
    rec[..., :3] -= 0.5
    R = torch.eye(3, device="cuda").repeat(B, 1, 1)
    T = torch.tensor([0.0, 0.0, 2.0], device="cuda").repeat(B, 1, 1)
    s, K = torch.ones(1, device="cuda"), torch.tensor([[1.2, 0, 0.5], [0, 1.2, 0.5], [0, 0, 1]], device="cuda").repeat(B, 1, 1)
    ts = []
    for it in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _CorrSums.apply(R, T, s, rec, K, 256, 1.0)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = float(np.median(ts[2:]))
    print(json.dumps({"B": B, "C": C, "bytes": B * C * 24, "ms_event_incl_launch": ms,
                      "GBs": B * C * 24 / ms / 1e6, "plan": plan(B, C)}))

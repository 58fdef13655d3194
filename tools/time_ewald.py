#!/usr/bin/env python
"""Ewald kernel timing at the bench shape: (a) one call per event pair (includes the host launch latency when the GPU
is idle), (b) R back-to-back calls per event pair (steady-state API throughput), (c) the same launches replayed from a
CUDA graph (device time of the kernel alone).  usage: python tools/time_ewald.py [case] [B ...]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_case

tag = sys.argv[1] if len(sys.argv) > 1 else "n29_rs1"
Bs = [int(a) for a in sys.argv[2:]] or [2048, 8192]
c = load_case(tag)
G = cg.kpoints(2, 15)
for B in Bs:
    reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
    x = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B], device="cuda")
    f = lambda: cg.potential_energy(x, c["kappa"], G, c["L"], c["rs"])
    for _ in range(5): f()
    torch.cuda.synchronize()
    def ev(fn, R):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(R): fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / R * 1e3
    single = min(ev(f, 1) for _ in range(10))
    loop = min(ev(f, 200) for _ in range(3))
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): f()
        torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(20): f()
    torch.cuda.synchronize()
    graph = min(ev(g.replay, 5) for _ in range(3)) / 20
    fl = cg.flops.flop_ewald(c["n"]) * B
    print("%s B=%d: single call %.1f us | 200 back-to-back %.1f us/call | CUDA-graph replay %.2f us/launch = %.3e evals/s, %.2f TFLOP/s"
          % (tag, B, single, loop, graph, B / graph * 1e6, fl / graph * 1e6 / 1e12), flush=True)

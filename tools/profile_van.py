#!/usr/bin/env python
"""Tiny driver for ncu: VAN sampler at a shipped shape.  usage: python tools/profile_van.py [tag] [B]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import coulombgas_b200 as cg
from coulombgas_b200.fixtures import load_van
tag = sys.argv[1] if len(sys.argv) > 1 else "n29_rs1"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2048
c = load_van(tag)
v = cg.VanSampler(c["ktable"], c["n"], c["M"])
for it in range(2):
    s, lp = v.sampler(c["params"], 5 + it, B, return_log_prob=True)
torch.cuda.synchronize()
print("ok", float(lp.mean()))

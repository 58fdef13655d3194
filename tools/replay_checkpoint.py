#!/usr/bin/env python
"""Replay a run written by the reference's main.py through the B200 kernels (SURVEY §8 f/N1).

Reads `<run_dir>/epoch_<N>.pkl` (walkers, trained flow parameters; no jax needed), re-shards the walkers over the
ranks of this job, runs `acc_steps` sampling steps (Metropolis sweep + wrap + Ewald energy, main.py:334-346 minus the
parts that stay in JAX) and prints one data.txt-format line (main.py:369-374) with the V, S and accept-rate columns
filled; F, E, K are `nan` (kinetic energy is not on this path).  The occupied momenta are drawn every step
from the checkpoint's own VAN parameters (src/VMC.py:22, `cgb_van_sample_f64io`; S = -<log p>); --ground-state uses the
ground state for every walker instead, --state-idx file.npy fixed (batch, n) int32 indices.

  python tools/replay_checkpoint.py RUN_DIR [--epoch N] [--acc-steps K] [--seed S]
  torchrun --nproc-per-node 8 tools/replay_checkpoint.py RUN_DIR ...
"""
import argparse
import glob
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("run_dir")
    ap.add_argument("--epoch", type=int, default=None)
    ap.add_argument("--acc-steps", type=int, default=None)
    ap.add_argument("--seed", type=int, default=42)
    ap.add_argument("--state-idx", default=None)
    ap.add_argument("--ground-state", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import coulombgas_b200 as cg
    from coulombgas_b200 import interop

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
    if world > 1:
        dist.init_process_group("nccl")
    f = interop.parse_run_dirname(args.run_dir)
    if f["dim"] != 2 or (f["depth"], f["spsize"], f["tpsize"]) != (2, 16, 16):
        raise SystemExit("unsupported run: the kernels cover dim=2, depth=2, spsize=tpsize=16")
    epoch = args.epoch
    if epoch is None:
        epoch = max(int(os.path.basename(p)[6:12]) for p in glob.glob(os.path.join(args.run_dir, "epoch_*.pkl")))
    ck = interop.load_data(interop.ckpt_filename(epoch, args.run_dir))
    n, rs, L = f["n"], f["rs"], interop.box_length(f["n"], 2)
    try:           # the order of degenerate orbitals the shipped runs were trained with, when it is known
        ktable, _ = interop.sp_indices_twist(2, f["Emax"], f["twist"], order="shipped")
    except ValueError:
        ktable, _ = interop.sp_indices_twist(2, f["Emax"], f["twist"])
    x = torch.as_tensor(interop.shard_walkers(ck["x"], rank, world)).cuda()
    B = x.shape[0]
    van = None
    if args.state_idx:
        si = interop.shard_walkers(np.load(args.state_idx)[:, :, None], rank, world)[:, :, 0].astype(np.int32)
    elif args.ground_state or not ck.get("params_van"):
        si = np.tile(interop.ground_state_indices(ktable.shape[0], n), (B, 1))
    else:
        van = cg.VanSampler(ktable, n, ktable.shape[0])
        params_van = torch.as_tensor(cg.pack_van_params(ck["params_van"])).cuda()
        si = np.zeros((B, n), np.int32)
    si = torch.as_tensor(si).cuda()
    G = cg.kpoints(2, f["Gmax"])
    Vconst = n * rs / L * cg.madelung(2, f["kappa"], G)                    # main.py:171
    logp = cg.make_logp(ktable, L)
    params = torch.as_tensor(cg.pack_params(ck["params_flow"])).cuda()
    acc_steps = args.acc_steps or f["acc_steps"]
    key = args.seed + rank
    V_mean = V2_mean = acc_sum = 0.0
    S_mean = S2_mean = float("nan") if van is None else 0.0
    for _ in range(acc_steps):
        if van is not None:
            si, lp = van.sampler(params_van, key ^ 0x5DEECE66D, B, return_log_prob=True)          # src/VMC.py:22
            sm, s2 = cg.potential_stats(-lp)                                                      # same all-reduce helper
            S_mean += float(sm) / acc_steps; S2_mean += float(s2) / acc_steps
        key, x, acc = cg.sample_x(key, si, logp, x, params, f["mc_steps"], f["mc_stddev"], L)
        V = cg.potential_energy(x, f["kappa"], G, L, rs) + Vconst               # src/VMC.py:40
        vm, v2 = cg.potential_stats(V)
        V_mean += float(vm) / acc_steps; V2_mean += float(v2) / acc_steps; acc_sum += float(acc)     # main.py:283-291
    nan = float("nan")
    V_std = interop.observable_std(V_mean, V2_mean, B * world, acc_steps)
    S_std = nan if van is None else interop.observable_std(S_mean, S2_mean, B * world, acc_steps)
    if rank == 0:
        sys.stdout.write(interop.format_data_line(epoch, F=nan, F_std=nan, E=nan, E_std=nan, K=nan, K_std=nan, V=V_mean,
                                                  V_std=V_std, S=S_mean, S_std=S_std, accept_rate=acc_sum / acc_steps, rs=rs))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

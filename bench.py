#!/usr/bin/env python
"""bench.py — headline benchmark of the forward-sampling hot path (BASELINE.json metric).

One "step" = one sampling step of the reference's training loop for this rank's shard of walkers
(main.py:334-338 -> src/VMC.py:8-25 -> src/MCMC.py): a Metropolis sweep of mc_steps=50 all-electron proposals
(51 evaluations of log p per walker), the wrap into [0,L), one Ewald energy evaluation of the final walkers
(src/VMC.py:40) and ONE all-reduce of the 5-entry statistics vector (accept rate, V mean, V^2 mean: the pmean sites
of src/MCMC.py:39 and src/VMC.py:46-53).  Per step: 3 kernel launches of libcgb200.so (prologue, sweep, Ewald) + one
constant-bank memcpy, captured in a CUDA graph; the per-device PRNG key lives on the device.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config c1..c5] [--scaling weak|strong] [--batch B]
                  [--impl ours|reference]
  torchrun ... bench.py --gpus N ...      (one rank per GPU, NCCL)

Workloads (BASELINE.json configs[0..4] = c1..c5); default c3 = the configuration the metric is quoted on:
  c1 n=13 rs=10 Emax=25 batch 256   c2 n=13 rs=10 Emax=25 batch 4096   c3 n=29 rs=1 Emax=25 batch 2048
  c4 n=49 rs=0.5 Emax=36 batch 512  c5 n=57 rs=0.25 Emax=49 batch 512      (dim 2, Gmax=15, kappa=10, mc_steps 50)
--scaling weak (default): every rank holds the config's batch (value = N x batch x 50 / time);
--scaling strong: the config's batch is the GLOBAL batch, sharded over the ranks exactly as main.py:231-238 does.
The line always carries both: `value` for the chosen mode and a "strong" / "weak" companion object for the other.

Prints ONE JSON line (rank 0).  `value` = walker-steps/s with inputs resident in HBM (device timed, CUDA events on the
launching stream, max over ranks); `e2e` = the same through the public host-buffer API with H2D/D2H copies inside the
timed region; `roofline` = FP32 work rate of the sweep kernel vs the FFMA peak measured by a saturating micro-kernel in
this same process; `parity` = the benchmarked kernel at the benchmarked shape against the float64 oracle on identical
walkers, noise and uniforms (mismatch fraction of the accept masks, max relative error of log p); `cpu_baseline` = the
oracle's timing on this box's host cores on that same bounded sample.  --impl reference times that CPU path alone.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONFIGS = {
    "c1": dict(tag="n13_rs10", n=13, rs=10.0, Emax=25, batch=256, devices=1, name="configs[0]"),
    "c2": dict(tag="n13_rs10", n=13, rs=10.0, Emax=25, batch=4096, devices=8, name="configs[1]"),
    "c3": dict(tag="n29_rs1", n=29, rs=1.0, Emax=25, batch=2048, devices=8, name="configs[2]"),
    "c4": dict(tag="n49_rs0.5", n=49, rs=0.5, Emax=36, batch=512, devices=8, name="configs[3]"),
    "c5": dict(tag="n57_rs0.25", n=57, rs=0.25, Emax=49, batch=512, devices=8, name="configs[4]"),
}
COMMON = dict(dim=2, Gmax=15, kappa=10, mc_steps=50, mc_stddev=0.1)
METRIC = "metropolis_walker_steps_per_sec"
UNIT = "walker-steps/s"
CPU_SAMPLE_WALKERS = {13: 128, 29: 64, 49: 24, 57: 16}      # oracle sample: these walkers x the full 50-step chain (~10-30 s)


def load_inputs(cfg, B, rank=0):
    """Synthetic inputs of the named shape: the shipped checkpoint walkers tiled to B with seeded
    N(0,0.1^2) jitter; trained flow parameters (init-scale for n=13: no shipped flow checkpoint); state_idx even=ground
    state, odd=sorted random subset."""
    from coulombgas_b200.fixtures import load_case
    c = load_case(cfg["tag"])
    rng = np.random.default_rng(42 + rank)
    reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
    x = np.tile(c["x"], (reps, 1, 1))[:B] + 0.1 * rng.standard_normal((B, c["n"], 2))
    x = x - c["L"] * np.floor(x / c["L"])
    M = c["ktable"].shape[0]
    si = np.empty((B, c["n"]), dtype=np.int32)
    rs = np.random.default_rng(43 + rank)
    for b in range(B):
        si[b] = np.arange(M - c["n"], M) if b % 2 == 0 else np.sort(rs.choice(M, size=c["n"], replace=False))
    return c, x, si


def make_config(ckey, cfg, B, world, scaling):
    return {"workload": "%s (%s): n=%d dim=2 rs=%g Theta=0.15 Emax=%d Gmax=15 (|G|=708) kappa=10, batch %d walkers/GPU "
                        "(%s scaling: global batch %d), mc_steps=50 (51 log p evals/walker) + 1 Ewald eval/walker + 1 all-reduce "
                        "per step" % (cfg["name"], ckey, cfg["n"], cfg["rs"], cfg["Emax"], B, scaling, B * world),
            "inputs": "shipped checkpoint walkers tiled + N(0,0.1) jitter, trained flow params, in-kernel Philox (device key)",
            "l2": "256 MiB write between timed iterations (L2 flush); working set itself is < L2",
            "parallelism": "dp%d (walkers sharded, 1 all-reduce of 5 doubles per step)" % world}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag, self.proc = index, [], False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([t.strip() for t in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_sample(c, x, si, noise, unif, walkers, threads):
    """The restated reference algorithm (oracle/: vmap'd flow + jacfwd + LAPACK slogdet, Metropolis chain of the FULL
    mc_steps, pairwise-cos Ewald) on `threads` host cores for the first `walkers` walkers of the batch.
    Returns (trace dict, x_final, V, seconds for the sweep, seconds for the Ewald)."""
    import torch
    from oracle import logpsi as ol, mcmc as om, potential as op, tables
    torch.set_num_threads(threads)
    S = noise.shape[0]
    xs, sis = x[:walkers], si[:walkers]
    lp = lambda y: ol.logp_batch(y, c["params"], sis, c["ktable"], c["L"], chunk=32)
    ol.logp_batch(xs[:2], c["params"], sis[:2], c["ktable"], c["L"])     # warm-up (torch.func first-call overhead)
    t0 = time.perf_counter()
    xo, acc, tr = om.mcmc(lp, xs, noise[:, :walkers], unif[:, :walkers], S, COMMON["mc_stddev"], trace=True)
    t1 = time.perf_counter()
    G = tables.kpoints(2, COMMON["Gmax"])
    V = op.potential_energy(om.wrap(xo, c["L"]), c["kappa"], G, c["L"], c["rs"])
    t2 = time.perf_counter()
    return tr, xo, V, t1 - t0, t2 - t1


def parity_report(trg, tro, unif, walkers, rtol=1e-5):
    """Free-running comparison of the GPU chain with the oracle chain up to the first divergence of each walker
    (tests/test_gpu_parity.py::test_mcmc_vs_oracle): mismatching decisions are legitimate only at near-ties."""
    S = tro["accept"].shape[0]
    acc_g, acc_r = trg["accept"][:, :walkers].astype(bool), tro["accept"]
    diverged = np.zeros(walkers, bool)
    mism = near = compared = 0
    max_rel = 0.0
    for s in range(S):
        live = ~diverged
        ref = tro["logp_prop"][s][live]
        if live.any():
            max_rel = max(max_rel, float(np.max(np.abs(trg["logp_prop"][s + 1][:walkers][live] - ref) / np.abs(ref))))
        bad = live & (acc_g[s] != acc_r[s])
        compared += int(live.sum())
        if bad.any():
            prev = tro["logp_init"] if s == 0 else tro["logp"][s - 1]
            margin = np.abs((tro["logp_prop"][s] - prev) - np.log(unif[s][:walkers]))[bad]
            near += int((margin < 2 * rtol * np.abs(tro["logp_prop"][s][bad])).sum())
        mism += int(bad.sum())
        diverged |= bad
    return {"walkers": walkers, "mc_steps": S, "decisions_compared": compared, "accept_mismatch_frac": mism / max(compared, 1),
            "mismatches": mism, "mismatches_at_near_ties": near, "max_rel_logp_err": max_rel, "tolerance": rtol,
            "walkers_identical_to_the_end": int((~diverged).sum())}


def run_reference(args, ckey, cfg):
    """--impl reference: the reference's CPU implementation of the path.  JAX/jaxlib/haiku are not installable in this
    image (no wheels, no network), so this is the float64 restatement of the reference algorithm as
    written (oracle/), all host threads, on a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    threads = os.cpu_count() or 1
    B = args.batch or (cfg["batch"] if args.scaling == "weak" else cfg["batch"] // world)
    c, x, si = load_inputs(cfg, 256)
    walkers, sub_steps = CPU_SAMPLE_WALKERS[cfg["n"]], 3
    S = COMMON["mc_steps"]
    vals, secs = [], []
    rng = np.random.default_rng(7)
    for it in range(args.warmup + args.steps):
        noise, unif = rng.standard_normal((sub_steps, walkers, c["n"], 2)), rng.random((sub_steps, walkers))
        _, _, _, t_sweep, t_ew = oracle_sample(c, x, si, noise, unif, walkers, threads)
        per_walker = t_sweep / walkers * (S + 1) / (sub_steps + 1) + t_ew / walkers     # a full step: 51 evals + Ewald
        if it >= args.warmup:
            vals.append(S / per_walker); secs.append(t_sweep + t_ew)
    value = float(np.mean(vals))
    sample = ("%d walkers x (%d proposals + 1 initial log p) + Ewald per step, scaled to the 50-proposal sweep; "
              "torch-f64 vmap+jacfwd+LAPACK" % (walkers, sub_steps))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": world * B * S / value * 1e3,
            "ms_per_step_note": "extrapolated time of one full step of this arm's batch at the measured rate; the timed "
                                "sample itself took %.1f ms per repetition" % (float(np.mean(secs)) * 1e3),
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": make_config(ckey, cfg, B, world, args.scaling),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--batch", type=int, default=0, help="walkers per GPU (overrides the config / scaling rule)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the captured CUDA graph")
    args = ap.parse_args()
    ckey, cfg = args.config, CONFIGS[args.config]
    if args.impl == "reference":
        return run_reference(args, ckey, cfg)

    import torch
    import torch.distributed as dist
    import coulombgas_b200 as cg

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner to STDOUT when the communicator is created; keep stdout for the one JSON line:
        # file descriptor 1 points at stderr while the group is set up and the first collective runs
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    S, n = COMMON["mc_steps"], cfg["n"]
    if cfg["batch"] % world != 0:
        raise SystemExit("Batch size must be divisible by the number of GPU devices (main.py:231-233)")
    B_weak, B_strong = cfg["batch"], cfg["batch"] // world
    B = args.batch or (B_weak if args.scaling == "weak" else B_strong)
    Bmax = max(B, B_weak)
    c, x_host_all, si_host_all = load_inputs(cfg, Bmax, rank)
    G = cg.kpoints(2, COMMON["Gmax"])
    Vconst = n * c["rs"] / c["L"] * cg.madelung(2, COMMON["kappa"], G)
    logp = cg.make_logp(c["ktable"], c["L"])
    params = torch.as_tensor(cg.pack_params(c["params"]), device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)       # > 126 MB L2

    def make_step(Bx, seed):
        x = torch.as_tensor(x_host_all[:Bx], device=dev).clone()
        si = torch.as_tensor(si_host_all[:Bx], device=dev)
        st = cg.SamplingStep(logp, si, params, x, cg.device_key(seed + rank), S, COMMON["mc_stddev"], COMMON["kappa"], G,
                             c["rs"], Vconst)
        if not args.no_graph:
            st.capture(warmup=2)
        return st

    def timed(st, steps, warmup):
        """mean ms per step over `steps` steps (CUDA events around each step, L2 flushed in between), max over ranks"""
        for _ in range(warmup):
            st.step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(steps):
            flush.fill_(k & 0xFF)                      # L2 flush between timed iterations (outside the event pair)
            ev[k][0].record()
            st.step()
            ev[k][1].record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        ms = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms), wall

    step = make_step(B, 1000)
    for _ in range(args.warmup):
        step.step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local)
    sampler.start()

    # nvidia-smi needs a moment to start (longer than the whole timed region on a cold box): keep the GPU under the same
    # load with untimed steps until its first sample has arrived, so that the samples that follow fall into the timed region
    def agree(flag):                       # the untimed extra steps contain collectives: every rank must take the same number
        if world == 1:
            return flag
        t = torch.tensor([1.0 if flag else 0.0], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return bool(t.item() > 0)

    t_wait = time.perf_counter()
    while agree(not sampler.rows and time.perf_counter() - t_wait < 5.0):
        step.step()
        torch.cuda.synchronize()
    n_before = len(sampler.rows)
    ms, t_wall = timed(step, args.steps, 0)
    acc_rate, v_mean, v2_mean = step.summary()
    t_wait = time.perf_counter()
    while agree(len(sampler.rows) < n_before + 3 and time.perf_counter() - t_wait < 2.0):
        step.step()
        torch.cuda.synchronize()
    clocks = sampler.finish()
    clocks["window"] = "timed region, extended by untimed steps of the same workload until nvidia-smi (100 ms period) delivered 3 samples"
    value = world * B * S / (ms * 1e-3)

    # ---- the other scaling mode beside the headline (same code path, same step)
    other_mode = "strong" if args.scaling == "weak" else "weak"
    B_other = B_strong if other_mode == "strong" else B_weak
    other = None
    if not args.batch:
        if B_other == B:
            other = {"scaling": other_mode, "global_batch": world * B_other, "walkers_per_gpu": B_other, "value": value,
                     "ms_per_step": ms, "note": "identical to the headline at this N"}
        else:
            st_o = make_step(B_other, 3000)
            ms_o, _ = timed(st_o, args.steps, args.warmup)
            other = {"scaling": other_mode, "global_batch": world * B_other, "walkers_per_gpu": B_other,
                     "value": world * B_other * S / (ms_o * 1e-3), "unit": UNIT, "ms_per_step": ms_o}

    # ---- kernel-only timings for the roofline (sweep kernel alone, Ewald kernel alone), CUDA events
    x = step.x
    si = torch.as_tensor(si_host_all[:B], device=dev)
    bound = logp.bind(params, si)

    def time_fn(fn, reps):
        fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return float(np.mean(ts))
    xk = x.clone()
    kkey = cg.device_key(77 + rank)
    ms_sweep = time_fn(lambda: cg.mcmc(bound, xk, kkey, S, COMMON["mc_stddev"], wrap_L=c["L"], advance_key=True, _want_rate=False), 5)
    # the Ewald kernel runs ~10 us: a single launch per event pair would time the host launch latency, so 20 launches
    # are captured in a CUDA graph and the replay is timed (device time per launch); fallback: 200 back-to-back calls
    few = lambda: cg.potential_energy(xk, COMMON["kappa"], G, c["L"], c["rs"])
    ew_reps, ew_mode = 20, "CUDA-graph replay of 20 launches"
    try:
        gs = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(gs):
            for _ in range(3):
                few()
            torch.cuda.synchronize()
            with torch.cuda.graph(graph, stream=gs):
                for _ in range(ew_reps):
                    few()
        torch.cuda.synchronize()
        ms_ewald = time_fn(graph.replay, 5) / ew_reps
    except Exception:                                  # pragma: no cover
        ew_reps, ew_mode = 200, "200 back-to-back calls"
        ms_ewald = time_fn(lambda: [few() for _ in range(ew_reps)], 3) / ew_reps
    # ---- the caller right before the sweep (SURVEY §8 f/N2): VAN draw of the occupied orbitals, reported beside the headline
    van_info = None
    try:
        from coulombgas_b200.fixtures import load_van
        vtag = {"n13_rs10": "n13_pretrained"}.get(cfg["tag"], cfg["tag"])
        vc = load_van(vtag)
        van = cg.VanSampler(vc["ktable"], vc["n"], vc["M"])
        pv = torch.as_tensor(cg.pack_van_params(vc["params"]), device=dev)
        ms_van = time_fn(lambda: van.sampler(pv, kkey, B, return_log_prob=True), 5)
        van_info = {"kernel": "van_kernel", "ms": ms_van, "walkers_per_sec": world * B / (ms_van * 1e-3),
                    "note": "draws + log_prob for the same batch; not part of the timed step (north_star keeps the VAN outside the path)"}
    except Exception as e:                                   # pragma: no cover
        van_info = {"error": repr(e)}
    blocks = 148 * 8
    ffma = max(2 * 8 * 20000 * blocks * 256 / cg.microbench(0, blocks, 20000) for _ in range(3)) / 1e12
    mufu = max(8 * 20000 * blocks * 256 / cg.microbench(1, blocks, 20000) for _ in range(3)) / 1e12
    flop = cg.flops.flop_sweep(n, S) * B
    achieved = flop / (ms_sweep * 1e-3) / 1e12
    # DRAM traffic of one sweep launch at the headline shape, from the committed ncu --set full capture (profiles/)
    traffic = None
    for tname in ("r02_sweep_traffic.json", "r01_sweep_traffic.json"):
        tfile = os.path.join(ROOT, "profiles", tname)
        if os.path.exists(tfile):
            tj = json.load(open(tfile))
            if tj.get("batch") == B and tj.get("mc_steps") == S and tj.get("n", 29) == n:
                traffic = tj.get("dram_bytes_per_launch")
                break
    peaks_file = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_file)) if os.path.exists(peaks_file) else {}
    wsup = n in (9, 13, 21, 25, 29)
    kname = ("wsweep_kernel<%d, %d>" % (n, 0 if B > 444 else (4 if B > 148 else 8))) if (wsup and B > 64 and not (n <= 16 and 444 < B <= 768)) \
        else "sweep_kernel<%d>" % max(2, (n + 7) // 8)

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H inside the timed region
    xh = torch.as_tensor(x_host_all[:B]).pin_memory()
    sih = torch.as_tensor(si_host_all[:B]).pin_memory()
    xo = torch.empty_like(xh).pin_memory()
    Vo = torch.empty(B, dtype=torch.float64).pin_memory()
    stats_h = torch.empty(8, dtype=torch.float64).pin_memory()
    stats_d = torch.zeros(8, dtype=torch.float64, device=dev)
    ekey = cg.device_key(5000 + rank)

    def e2e_step():
        xd = xh.to(dev, non_blocking=True)
        sd = sih.to(dev, non_blocking=True)
        cg.mcmc(logp.bind(params, sd), xd, ekey, S, COMMON["mc_stddev"], wrap_L=c["L"], stats=stats_d, advance_key=True,
                _want_rate=False)
        V = cg.potential_energy(xd, COMMON["kappa"], G, c["L"], c["rs"], stats=stats_d)
        if world > 1:
            dist.all_reduce(stats_d[:5])
        xo.copy_(xd, non_blocking=True)
        Vo.copy_(V, non_blocking=True)
        stats_h.copy_(stats_d, non_blocking=True)
    for _ in range(2):
        e2e_step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ke = max(3, args.steps // 2)
    a.record()
    for _ in range(ke):
        e2e_step()
    b.record()
    torch.cuda.synchronize()
    e2e_ms = torch.tensor([a.elapsed_time(b) / ke], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * B * S / (float(e2e_ms) * 1e-3)
    h2d = xh.numel() * 8 + sih.numel() * 4
    d2h = xo.numel() * 8 + Vo.numel() * 8 + 8 * 8

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32 (f64 I/O, range reduction, single-particle stream, log-det sums and Metropolis test)", "data": "synthetic",
        "config": make_config(ckey, cfg, B, world, args.scaling),
        other_mode: other,
        "ewald_evals_per_sec": world * B / (ms_ewald * 1e-3),
        "sweep_kernel_ms": ms_sweep, "ewald_kernel_ms": ms_ewald, "ewald_timing": ew_mode,
        "ewald_roofline": {"bound": "fp32", "achieved": cg.flops.flop_ewald(n) * B / (ms_ewald * 1e-3) / 1e12, "peak": ffma,
                           "unit": "TFLOP/s", "frac": cg.flops.flop_ewald(n) * B / (ms_ewald * 1e-3) / 1e12 / ffma,
                           "kernel": "ewald_kernel"},
        "accept_rate": acc_rate, "V_mean_over_rs2": v_mean / c["rs"] ** 2,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": 3 * args.steps,      # per step: wsweep_prologue (or prep_params) + sweep kernel + ewald_kernel (+ 1 memcpy node, + NCCL)
        "step_mode": "eager launches" if args.no_graph else "CUDA graph replay (3 kernels + constant-bank memcpy%s)" % (" + NCCL all-reduce" if world > 1 else ""),
        "clocks": clocks,
        "roofline": {"bound": "fp32", "achieved": achieved, "peak": ffma, "unit": "TFLOP/s", "frac": achieved / ffma,
                     "traffic": traffic, "kernel": kname,
                     "note": "compute-bound path (>2000 FLOP/B): achieved = (8n^3+895n^2+1760n)*(mc_steps+1)*B algorithmic "
                             "FLOP / CUDA-event kernel time; peak = FFMA-saturating micro-kernel measured in this run "
                             "(%s); MUFU peak %.2f Top/s"
                             % ("MEASURED_PEAKS.json holds only HBM and bf16 peaks: %s" % json.dumps(peaks) if peaks else
                                "MEASURED_PEAKS.json absent on this box; it holds only HBM and bf16 peaks anyway", mufu)},
        "wall_s_timed_region": t_wall,
        "van_sampler": van_info,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # ---- parity at the benchmarked shape + CPU baseline, one oracle run serves both: the benchmarked kernel (library's
        # own choice at this batch) with explicit noise on all B walkers vs the float64 oracle on the first `walkers` of them
        threads = os.cpu_count() or 1
        walkers = min(CPU_SAMPLE_WALKERS[n], B)
        rng = np.random.default_rng(44)
        noise, unif = rng.standard_normal((S, B, n, 2)), rng.random((S, B))
        x0 = x_host_all[:B].copy()
        xg, accg, trg = cg.mcmc(bound, x0.copy(), cg.ExplicitNoise(noise, unif), S, COMMON["mc_stddev"], return_trace=True)
        tro, xor_, Vor, t_sweep, t_ew = oracle_sample(c, x0, si_host_all[:B], noise, unif, walkers, threads)
        line["parity"] = parity_report(trg, tro, unif, walkers)
        line["parity"]["kernel"] = kname
        Vg = cg.potential_energy(xor_ - c["L"] * np.floor(xor_ / c["L"]), COMMON["kappa"], G, c["L"], c["rs"])
        line["parity"]["max_rel_ewald_err"] = float(np.max(np.abs(Vg - Vor)) / np.max(np.abs(Vor)))
        line["accept_mismatch_frac"] = line["parity"]["accept_mismatch_frac"]
        line["max_rel_logp_err"] = line["parity"]["max_rel_logp_err"]
        v = walkers * S / (t_sweep + t_ew)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": "%d walkers x full step (50 proposals + initial log p + Ewald): restated reference "
                                          "algorithm (torch-f64 vmap+jacfwd+LAPACK), %.1f s; the same run is the parity check"
                                          % (walkers, t_sweep + t_ew),
                                "ewald_evals_per_sec": walkers / t_ew}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        # Orderly exit: CUDA graphs that hold NCCL kernels must be gone before the communicator is torn down (a
        # destroy_process_group() with live captured collectives was observed to hang); after the last barrier every rank has
        # finished its work, so the processes leave without the teardown handshake.
        step.graph = None
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    main()

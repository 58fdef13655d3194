"""GPU tests of the throughput mode's random streams and of the boundary features around them (run on the B200 box).

The headline benchmark runs with the in-kernel generator, so it gets its own evidence: Philox4x32-10 against the Random123
known-answer vectors and the NumPy restatement (oracle/philox.py), the mapping of its words to the Metropolis uniforms
(bit-exact) and to the Box-Muller proposal noise (1e-6), moments and Kolmogorov-Smirnov tests on >= 2e6 draws, the
device-resident key, two streams with two parameter sets, and the fused sampling step (statistics vector, CUDA graph)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import coulombgas_b200 as cg
    return cg


def test_philox_known_answer_vectors(cg):
    from oracle import philox
    rng = np.random.default_rng(1)
    kat = [list(c) + list(k) for c, k, _ in philox.KAT]
    rnd = rng.integers(0, 2 ** 32, size=(1000, 6), dtype=np.uint64).astype(np.uint32)
    allk = np.concatenate([np.array(kat, dtype=np.uint32), rnd])
    _, _, out = cg.debug_rng(7, 1, 1, 1, kat=allk)
    for t, (_, _, want) in enumerate(philox.KAT):
        assert tuple(int(v) for v in out[t]) == want
    ref = philox.philox4x32_10(allk[:, :4], allk[:, 4:])
    assert np.array_equal(out, ref)


def test_stream_mapping_matches_restatement(cg):
    from oracle import philox
    seed, B, S, n = 0x1234567890ABCDEF, 37, 9, 29
    nrm, uni = cg.debug_rng(seed, B, S, n)
    assert np.array_equal(uni.cpu().numpy(), philox.uniforms(seed, B, S))          # integer -> float64: bit-exact
    ref = philox.normals(seed, B, S, n)
    assert np.abs(nrm.cpu().numpy() - ref).max() < 2e-6                            # FP32 logf / sincospif / sqrtf
    # the device-resident key gives the same streams as the host value
    nrm2, uni2 = cg.debug_rng(cg.device_key(seed), B, S, n)
    assert torch.equal(nrm, nrm2) and torch.equal(uni, uni2)


def test_normal_and_uniform_statistics(cg):
    """>= 2e6 draws each: first four moments of the proposal noise within 5 standard errors, KS distance against N(0,1) /
    U[0,1) at the 1e-3 level, no correlation between the two Box-Muller outputs or between neighbouring particles."""
    from scipy import stats
    B, S, n = 2048, 20, 29
    nrm, uni = cg.debug_rng(20260117, B, S, n)
    z = nrm.cpu().numpy().astype(np.float64).reshape(-1)
    N = z.size
    assert N >= 2_000_000
    assert abs(z.mean()) < 5 / np.sqrt(N)
    assert abs(z.var() - 1.0) < 5 * np.sqrt(2.0 / N)
    assert abs((z ** 3).mean()) < 5 * np.sqrt(15.0 / N)
    assert abs((z ** 4).mean() - 3.0) < 5 * np.sqrt(96.0 / N)
    assert abs(np.abs(z).max()) > 4.5                                 # the tails are populated
    sub = z[::7][:400_000]
    assert stats.kstest(sub, "norm").pvalue > 1e-3
    zz = nrm.cpu().numpy().astype(np.float64)
    assert abs((zz[..., 0] * zz[..., 1]).mean()) < 5 / np.sqrt(N / 2)                  # cos / sin branch
    assert abs((zz[:, :, :-1, 0] * zz[:, :, 1:, 0]).mean()) < 5 / np.sqrt(N / 2)       # neighbouring particles
    assert abs((zz[:-1] * zz[1:]).mean()) < 5 / np.sqrt(N)                             # consecutive steps
    B2, S2 = 8192, 256
    _, u = cg.debug_rng(99, B2, S2, 1)
    u = u.cpu().numpy().reshape(-1)
    assert u.min() >= 0.0 and u.max() < 1.0
    assert abs(u.mean() - 0.5) < 5 / np.sqrt(12.0 * u.size)
    assert abs(u.var() - 1.0 / 12.0) < 5 * np.sqrt(1.0 / 180.0 / u.size)
    assert stats.kstest(u[::5][:400_000], "uniform").pvalue > 1e-3


@pytest.mark.parametrize("kernel", ["warp", "cta", "spec4"])
def test_device_key_equals_host_key(cg, cases, kernel):
    """A sweep keyed by a device-resident (2,) uint32 key (src/VMC.py:20: the traced per-device key) is bit-identical to the
    same sweep keyed by the host integer; CGB_FLAG_ADVANCE_KEY replaces the key by key + golden ratio before use."""
    c = cases("n29_rs1")
    cg.set_sweep_kernel(kernel)
    try:
        B, S = 96, 7
        logp = cg.make_logp(c["ktable"], c["L"])
        si = torch.as_tensor(c["state_idx"][:B], device="cuda")
        bound = logp.bind(c["params"], si)
        x0 = torch.as_tensor(c["x"][:B], device="cuda")
        seed = 0xDEADBEEF12345678
        xa, ra = cg.mcmc(bound, x0.clone(), seed, S, 0.1)
        key = cg.device_key(seed)
        xb, rb = cg.mcmc(bound, x0.clone(), key, S, 0.1)
        assert torch.equal(xa, xb) and float(ra) == float(rb)
        assert key.cpu().numpy().view(np.uint32).tolist() == [seed >> 32, seed & 0xFFFFFFFF]     # untouched
        xc, rc = cg.mcmc(bound, x0.clone(), key, S, 0.1, advance_key=True)
        nxt = (seed + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        assert key.cpu().numpy().view(np.uint32).tolist() == [nxt >> 32, nxt & 0xFFFFFFFF]
        xd, rd = cg.mcmc(bound, x0.clone(), nxt, S, 0.1)
        assert torch.equal(xc, xd) and not torch.equal(xc, xa)
    finally:
        cg.set_sweep_kernel(None)


def test_two_streams_two_parameter_sets(cg, cases):
    """The warp kernel's weights live in a per-device constant bank: sweeps with DIFFERENT parameters issued on two streams
    of one device must each see their own parameters (the library serialises them through its per-device event)."""
    c = cases("n29_rs1")
    cg.set_sweep_kernel("warp")
    try:
        B, S = 600, 12
        logp = cg.make_logp(c["ktable"], c["L"])
        reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
        x0 = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B], device="cuda")
        si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
        pa = cg.pack_params(c["params"])
        pb = pa.copy()
        pb[:96] *= 1.7                                            # a different two-particle layer
        ba, bb = logp.bind(pa, si), logp.bind(pb, si)
        ref_a, _ = cg.mcmc(ba, x0.clone(), 5, S, 0.1)
        ref_b, _ = cg.mcmc(bb, x0.clone(), 5, S, 0.1)
        assert not torch.equal(ref_a, ref_b)
        torch.cuda.synchronize()
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        for _ in range(4):
            xa, xb = x0.clone(), x0.clone()
            torch.cuda.synchronize()
            with torch.cuda.stream(s1):
                cg.mcmc(ba, xa, 5, S, 0.1)
            with torch.cuda.stream(s2):
                cg.mcmc(bb, xb, 5, S, 0.1)
            with torch.cuda.stream(s1):
                xa2 = x0.clone()
                cg.mcmc(ba, xa2, 5, S, 0.1)
            torch.cuda.synchronize()
            assert torch.equal(xa, ref_a) and torch.equal(xb, ref_b) and torch.equal(xa2, ref_a)
    finally:
        cg.set_sweep_kernel(None)


@pytest.mark.timeout(240)
def test_two_streams_segmented_chains(cg, cases):
    """Segmented chains hand a walker from CTA to CTA through flags in global memory; a segment only ever waits for one that was
    ticketed earlier, i.e. for a RUNNING CTA, so two such sweeps sharing the device (each wants every CTA slot) cannot deadlock
    and must each reproduce their own result."""
    c = cases("n57_rs0.25")
    prev = cg.set_sweep_kernel("cta")
    cg.set_debug(segments=True)
    try:
        B, S = 350, 6
        logp = cg.make_logp(c["ktable"], c["L"])
        reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
        x0 = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B], device="cuda")
        si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
        pa = cg.pack_params(c["params"])
        pb = pa.copy()
        pb[:96] *= 1.3
        ba, bb = logp.bind(pa, si), logp.bind(pb, si)
        ref_a, _ = cg.mcmc(ba, x0.clone(), 5, S, 0.1)
        ref_b, _ = cg.mcmc(bb, x0.clone(), 6, S, 0.1)
        assert not torch.equal(ref_a, ref_b)
        torch.cuda.synchronize()
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
        for _ in range(3):
            xa, xb = x0.clone(), x0.clone()
            torch.cuda.synchronize()
            with torch.cuda.stream(s1):
                cg.mcmc(ba, xa, 5, S, 0.1)
            with torch.cuda.stream(s2):
                cg.mcmc(bb, xb, 6, S, 0.1)
            torch.cuda.synchronize()
            assert torch.equal(xa, ref_a) and torch.equal(xb, ref_b)
    finally:
        cg.set_debug(segments=None)
        cg.set_sweep_kernel(prev)


def test_alternating_batch_sizes_one_process(cg, cases):
    """Launch configurations are cached per kernel instantiation; the dynamic-shared-memory attribute belongs to the function,
    not to the launch shape: a small batch (2 warps per CTA) between two large ones (4 warps per CTA) must not shrink it."""
    c = cases("n29_rs1")
    cg.set_sweep_kernel("warp")
    try:
        logp = cg.make_logp(c["ktable"], c["L"])
        ref = {}
        for B in (2048, 1000, 2048, 300, 2048, 1000):
            reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
            x = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B], device="cuda")
            si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
            xo, acc = cg.mcmc(logp.bind(c["params"], si), x.clone(), 11, 4, 0.1)
            torch.cuda.synchronize()
            if B in ref:
                assert torch.equal(xo, ref[B])
            ref[B] = xo
    finally:
        cg.set_sweep_kernel(None)


def test_argument_validation_precedes_the_launch(cg, cases):
    """mcmc() must reject a wrong wrap_L / shape BEFORE the kernel advances the donated walker buffer."""
    c = cases("n13_rs10")
    logp = cg.make_logp(c["ktable"], c["L"])
    B = 8
    si = torch.as_tensor(c["state_idx"][:B], device="cuda")
    x = torch.as_tensor(c["x"][:B], device="cuda")
    keep = x.clone()
    with pytest.raises(ValueError):
        cg.mcmc(logp.bind(c["params"], si), x, 1, 5, 0.1, wrap_L=c["L"] * 1.01)
    with pytest.raises(ValueError):
        cg.mcmc(logp.bind(c["params"], si), x.reshape(B, -1), 1, 5, 0.1)
    with pytest.raises(ValueError):
        cg.mcmc(logp.bind(c["params"], si), x, 1, 0, 0.1)
    torch.cuda.synchronize()
    assert torch.equal(x, keep)


@pytest.mark.parametrize("tag,B", [("n29_rs1", 2048), ("n29_rs1", 256), ("n57_rs0.25", 64)])
def test_sampling_step_statistics_and_graph(cg, cases, tag, B):
    """The fused sampling step (prologue + sweep + Ewald, statistics from the kernels' epilogues): its statistics vector
    equals what the separate public calls give on the same key sequence, a captured CUDA graph replays it with a fresh key
    each time (device key advanced by the prologue), and graph replays equal eager steps."""
    c = cases(tag)
    n, S = c["n"], 6
    logp = cg.make_logp(c["ktable"], c["L"])
    reps = (B + c["x"].shape[0] - 1) // c["x"].shape[0]
    x0 = torch.as_tensor(np.tile(c["x"], (reps, 1, 1))[:B], device="cuda")
    si = torch.as_tensor(np.tile(c["state_idx"], (reps, 1))[:B], device="cuda")
    params = torch.as_tensor(cg.pack_params(c["params"]), device="cuda")
    G = cg.kpoints(2, 15)
    Vconst = n * c["rs"] / c["L"] * cg.madelung(2, 10, G)
    seed = 424242
    # reference: separate calls, host keys following the documented advance rule
    xr = x0.clone()
    want = []
    k = seed
    for _ in range(3):
        k = (k + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
        xr, acc = cg.mcmc(logp.bind(params, si), xr, k, S, 0.1, wrap_L=c["L"])
        V = cg.potential_energy(xr, 10, G, c["L"], c["rs"]) + Vconst
        want.append((float(acc), float(V.mean()), float((V * V).mean()), xr.clone()))
    # eager fused steps
    xe = x0.clone()
    st = cg.SamplingStep(logp, si, params, xe, cg.device_key(seed), S, 0.1, 10, G, c["rs"], Vconst)
    for i in range(3):
        st.step()
        a, vm, v2 = st.summary()
        assert abs(a - want[i][0]) < 1e-12 and abs(vm - want[i][1]) < 1e-9 * abs(want[i][1])
        assert abs(v2 - want[i][2]) < 1e-9 * want[i][2]
        assert torch.equal(xe, want[i][3])
    # graph replays (capture runs 2 warm-up steps + the capture pass does not execute): keys continue the same sequence
    xg = x0.clone()
    sg = cg.SamplingStep(logp, si, params, xg, cg.device_key(seed), S, 0.1, 10, G, c["rs"], Vconst)
    sg.capture(warmup=0)
    for i in range(3):
        sg.step()
        a, vm, v2 = sg.summary()
        assert abs(a - want[i][0]) < 1e-12 and abs(vm - want[i][1]) < 1e-9 * abs(want[i][1])
        assert torch.equal(xg, want[i][3])

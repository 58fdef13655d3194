"""CPU model of the speculative ("prefetching") Metropolis scheduler of the small-batch kernels (DESIGN.md section 3.5;
csrc/cgb_wsweep.cu SPECW, csrc/cgb_sweep.cu SPECC): same node numbering, same choice of the extra node, same replay of the
decisions.  Property: for ANY log p function and any noise / uniforms keyed by step, the speculative schedule produces
exactly the chain of src/MCMC.py:22-34 run step by step -- the GPU tests check the kernels against the sequential kernels,
this test pins the scheduling logic itself (runs without a GPU)."""
import numpy as np
import pytest


def sequential(logp, x0, noise, unif, sigma):
    x, lp, acc = x0.copy(), logp(x0), []
    for s in range(len(unif)):
        xp = x + sigma * noise[s]                               # MCMC.py:26
        lpp = logp(xp)
        a = bool(unif[s] < np.exp(lpp - lp))                    # MCMC.py:28-29 (NaN -> reject)
        if a:
            x, lp = xp, lpp                                     # MCMC.py:31
        acc.append(a)
    return x, lp, acc


def speculative(logp, x0, noise, unif, sigma, W):
    """W evaluators per walker.  Evaluator w < W-1 takes node w+1 of the binary decision tree (heap numbering: the child
    of node m is 2m + accepted); the last one takes log p of the initial walker in the first round and afterwards node W
    (all rejected) or 2W-1 (all accepted), whichever the running acceptance favours."""
    S = len(unif)
    x, lp, t, nacc, acc, rounds, evals = x0.copy(), None, 0, 0, [], 0, 0
    while t < S or rounds == 0:
        first = t == 0
        extra = 2 * W - 1 if 2 * nacc > t else W
        nodes = list(range(1, W)) + [0 if first else extra]
        val = {}
        for node in nodes:                                      # "concurrently"
            xv = x.copy()
            if node > 0:
                depth = node.bit_length() - 1
                step = t + 1 + depth
                if step > S:
                    continue                                    # inactive evaluator
                for j in range(depth):
                    if (node >> (depth - 1 - j)) & 1:
                        xv = xv + sigma * noise[t + j]          # accepted moves above this node, in step order
                xv = xv + sigma * noise[step - 1]
            val[node] = logp(xv)
            evals += 1
        if first:
            lp = val[0]
        node, adv = 1, 0
        while t + adv < S:
            if not (node < W or (not first and node == extra)):
                break                                           # not evaluated in this round
            s = t + adv + 1
            a = bool(unif[s - 1] < np.exp(val[node] - lp))
            if a:
                lp = val[node]
                nacc += 1
                x = x + sigma * noise[s - 1]
            acc.append(a)
            node = 2 * node + int(a)
            adv += 1
        assert adv >= min(int(np.log2(W)), S - t) or S == 0
        t += adv
        rounds += 1
    return x, lp, acc, rounds, evals


@pytest.mark.parametrize("W", [2, 4, 8])
@pytest.mark.parametrize("S", [0, 1, 2, 3, 7, 50])
def test_speculative_schedule_reproduces_the_sequential_chain(W, S):
    for seed in range(20):
        rng = np.random.default_rng(1000 * W + 10 * S + seed)
        d = 6
        A = rng.normal(size=(d, d))
        scale = rng.choice([0.3, 1.0, 3.0])                     # acceptance from ~0.9 down to ~0.1

        def logp(x):
            return float(-scale * np.sum((A @ x) ** 2) + np.sin(3.0 * x[0]))
        x0 = rng.normal(size=d)
        noise, unif = rng.normal(size=(S, d)), rng.uniform(size=S)
        if S > 2 and seed == 0:
            unif[1] = 0.0                                       # u = 0 accepts anything finite (MCMC.py:29: u in [0,1))
        xs, lps, accs = sequential(logp, x0, noise, unif, 0.1)
        xq, lpq, accq, rounds, evals = speculative(logp, x0, noise, unif, 0.1, W)
        assert accq == accs
        assert np.array_equal(xq, xs)                           # same float64 additions in the same order
        assert lpq == lps
        D = int(np.log2(W))
        assert rounds <= max(1, -(-S // D))                     # at least D steps per round
        assert evals <= W * rounds


def test_nan_logp_rejects_in_both_schedules():
    rng = np.random.default_rng(3)
    noise, unif = rng.normal(size=(9, 2)), rng.uniform(size=9)

    def logp(x):
        return float("nan") if x[0] > 0.05 else float(-np.sum(x * x))
    x0 = np.zeros(2)
    xs, lps, accs = sequential(logp, x0, noise, unif, 0.1)
    for W in (4, 8):
        xq, lpq, accq, _, _ = speculative(logp, x0, noise, unif, 0.1, W)
        assert accq == accs and np.array_equal(xq, xs)

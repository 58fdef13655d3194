"""NumPy model of the *kernel's* algorithm (analytic Jacobian, twist-free Slater matrix, common-mode
removal) — a test helper that documents, in executable form, the arithmetic the CUDA kernels do
 and lets the CPU suite check that algorithm against the as-written oracle in float64
and estimate its float32 error budget.  Not part of the product."""
import numpy as np


def pack_params(params):
    """haiku dict -> the flat 1074-float layout [Wt,bt,W0,b0,W1,b1,Wf,bf]."""
    g = lambda k: params[k]
    return np.concatenate([
        g("fermi_net/~/linear_2")["w"].ravel(), g("fermi_net/~/linear_2")["b"].ravel(),
        g("fermi_net/~/linear")["w"].ravel(), g("fermi_net/~/linear")["b"].ravel(),
        g("fermi_net/~/linear_1")["w"].ravel(), g("fermi_net/~/linear_1")["b"].ravel(),
        g("fermi_net/linear")["w"].ravel(), g("fermi_net/linear")["b"].ravel()]).astype(np.float64)


def unpack(flat, dt):
    o = 0
    def take(*shape):
        nonlocal o
        sz = int(np.prod(shape)); a = flat[o:o + sz].reshape(shape).astype(dt); o += sz
        return a
    Wt, bt, W0, b0, W1, b1, Wf, bf = take(5, 16), take(16), take(9, 16), take(16), take(48, 16), take(16), \
        take(16, 2), take(2)
    return Wt, bt, W0[4:9], b0, W1[0:16], W1[16:32], W1[32:48], b1, Wf, bf


def _softplus(a):
    return np.maximum(a, 0) + np.log1p(np.exp(-np.abs(a)))


def _sigmoid(a):
    return 1.0 / (1.0 + np.exp(-a))


def logp_model(x, flat_params, kint, L, dt=np.float64, return_parts=False):
    """x (n,2) float64; kint (n,2) integer part m_j of the occupied k-rows (twist dropped: a pure
    row phase of D).  dt = arithmetic type of everything that is FP32 on the GPU."""
    n = x.shape[0]
    Wt, bt, W0c, b0, W1a, W1b, W1c, b1, Wf, bf = unpack(flat_params, dt)
    u = x / L
    u = u - np.floor(u)                                   # f64 range reduction
    ch, sh = np.cos(np.pi * u).astype(dt), np.sin(np.pi * u).astype(dt)     # half angles
    # ---- phase 1
    shij = sh[:, None, :] * ch[None, :, :] - ch[:, None, :] * sh[None, :, :]
    chij = ch[:, None, :] * ch[None, :, :] + sh[:, None, :] * sh[None, :, :]
    c = 1 - 2 * shij * shij
    s = 2 * shij * chij
    d = np.sqrt((shij * shij).sum(-1))
    t0 = np.concatenate([c, s, d[..., None]], -1)         # (n,n,5)
    a = t0 @ Wt + bt
    sig_t = _sigmoid(a)
    m = _softplus(a).mean(1)                              # (n,16)
    m0 = t0.mean(1)                                       # (n,5)
    # ---- single-particle stream
    pre0 = m0 @ W0c + b0
    h, sig0 = _softplus(pre0), _sigmoid(pre0)
    hbar = h.mean(0)
    g = h @ W1a + hbar @ W1b + m @ W1c + b1
    sig1 = _sigmoid(g)
    sp = h + _softplus(g)
    delta = (sp - sp.mean(0)) @ Wf                        # common mode removed (global phase only)
    w = u + (delta / dt(L)).astype(np.float64)            # box units
    # ---- per-particle contractions
    WfS = Wf.T[None, :, :] * sig1[:, None, :]             # (n,2,16): Wf[e,a]*sig1[i,e]
    M = Wf.T[None] + WfS @ W1a.T                          # (n,2,16)
    Q = (M * sig0[:, None, :]) @ W0c.T                    # (n,2,5)
    R = WfS @ W1c.T                                       # (n,2,16)
    T = WfS @ W1b.T                                       # (n,2,16)
    # ---- phase 2
    alpha = dt(2 * np.pi / (L * n))
    offd = (1 - np.eye(n)).astype(dt)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv4d = np.where(np.eye(n, dtype=bool), 0, 1 / (4 * d)).astype(dt)
    E = np.zeros((n, n, 5, 2), dt)                        # E[i,k] = (1/n) d t0_ik / d x_k
    for comp in range(2):
        E[:, :, 0 + comp, comp] = alpha * s[..., comp] * offd
        E[:, :, 2 + comp, comp] = -alpha * c[..., comp] * offd
        E[:, :, 4, comp] = -alpha * s[..., comp] * inv4d
    F = sig_t[..., None] * np.einsum("fc,ikfd->ikcd", Wt, E)            # (n,n,16,2)
    Jloc = np.einsum("iaf,ikfd->ikad", Q, E) + np.einsum("iac,ikcd->ikad", R, F)
    PE = sig0[:, None, :, None] * np.einsum("fc,ikfd->ikcd", W0c, E)    # P_i E_ik
    H = (PE.sum(0) - sig0[:, :, None] * np.einsum("fc,kfd->kcd", W0c, E.sum(1))) / dt(n)   # (n,16,2)
    Jb = Jloc + np.einsum("iac,kcd->ikad", T, H)
    Jb = Jb * offd[:, :, None, None]
    diag = np.eye(2, dtype=dt)[None] - Jb.sum(1)
    Jb[np.arange(n), np.arange(n)] = diag
    J = Jb.transpose(0, 2, 1, 3).reshape(2 * n, 2 * n)
    # ---- Slater matrix without twist / common phase
    ph = (w[:, None, :] * kint[None, :, :]).sum(-1)
    ph = ph - np.rint(ph)
    D = (np.cos(2 * np.pi * ph).astype(dt) + 1j * np.sin(2 * np.pi * ph).astype(dt))
    ld_D = np.linalg.slogdet(D.astype(np.complex64 if dt == np.float32 else np.complex128))[1] - n * np.log(L)
    ld_J = np.linalg.slogdet(J)[1]
    logp = 2 * float(ld_D) + float(ld_J)
    if return_parts:
        return logp, dict(J=J, w=w, logabsdet_slater=float(ld_D), logjacdet=float(ld_J))
    return logp


def split_ktable(ktable):
    """k = m + twist with m integer and one common fractional twist (src/orbitals.py:51-55)."""
    tw = ktable[0] - np.rint(ktable[0])
    m = np.rint(ktable - tw)
    assert np.abs(ktable - tw - m).max() < 1e-12
    return m.astype(np.int32), tw

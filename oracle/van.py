"""CPU oracle (test infrastructure only) for the autoregressive momentum sampler that feeds the sweep — SURVEY §8 f/N2.

Restates, in NumPy float64:
  Transformer.__call__        src/autoregressive.py:69-96   (embedding -> tanh -> [causal MHA + residual, 2-layer tanh MLP
                                                             + residual] x num_layers -> tanh -> output linear -> shift by x1hat)
  CausalSelfAttention         src/autoregressive.py:15-30   = haiku.MultiHeadAttention with a lower-triangular mask
  _mask / _logits             src/sampler.py:6-29
  sampler                     src/sampler.py:30-38           (n sequential categorical draws)
  log_prob                    src/sampler.py:40-44

Third-party arithmetic that is NOT under /root/reference (dm-haiku, unpinned: CI installs git main):
hk.MultiHeadAttention — restated from its published algorithm: per-head projections `query`/`key`/`value`
(Linear(num_heads*key_size)), logits q.k / sqrt(key_size), masked entries set to -1e30, softmax over keys, heads
concatenated, final `linear` projection to model_size.  The parameter names and shapes agree with the shipped
checkpoints (`transformer/layer{i}_attn/{query,key,value,linear}`); numerically this boundary is pinned only through the
reference's own property test (exact normalisation over all C(10,4) states, tests/test_sampler.py:45-69) and the
entropy column S of the shipped data.txt logs (tests).  jax.random.categorical(key, logits) = argmax(logits + Gumbel):
the Gumbel noise is an INPUT here (JAX's threefry stream is parity-unpinned, as for the Metropolis noise).
"""
import numpy as np


def _lin(p, name, x):
    return x @ np.asarray(p[name]["w"], np.float64) + np.asarray(p[name]["b"], np.float64)


def _softmax(a):
    a = a - a.max(axis=-1, keepdims=True)
    e = np.exp(a)
    return e / e.sum(axis=-1, keepdims=True)


def transformer(params, x, num_layers=2, num_heads=4, prefix="transformer"):
    """src/autoregressive.py:69-96.  x (n, dim) float64 -> (n, output_size)."""
    x = np.asarray(x, np.float64)
    T = x.shape[0]
    h = np.tanh(_lin(params, prefix + "/embedding_mlp", x))                      # :72-75
    model = h.shape[-1]
    ks = model // num_heads
    causal = np.tril(np.ones((T, T), bool))                                       # :27
    for i in range(num_layers):
        a = "%s/layer%d_attn" % (prefix, i)
        q = _lin(params, a + "/query", h).reshape(T, num_heads, ks)
        k = _lin(params, a + "/key", h).reshape(T, num_heads, ks)
        v = _lin(params, a + "/value", h).reshape(T, num_heads, ks)
        logits = np.einsum("thd,Thd->htT", q, k) / np.sqrt(ks)
        logits = np.where(causal[None], logits, -1e30)
        w = _softmax(logits)
        att = np.einsum("htT,Thd->thd", w, v).reshape(T, model)
        h = h + _lin(params, a + "/linear", att)                                  # :82-83
        m = "%s/layer%d_mlp" % (prefix, i)
        h = h + _lin(params, m + "/linear_1", np.tanh(_lin(params, m + "/linear", h)))   # :84-85, DenseBlock :43-48
    h = np.tanh(h)                                                                # :87
    out = _lin(params, prefix + "/output_mlp", h)                                 # :88-90
    x1hat = np.asarray(params[prefix]["x1hat"], np.float64)
    return np.vstack((x1hat, out[:-1]))                                           # :92-94


def mask(state_idx, n, num_states):
    """src/sampler.py:6-10."""
    m = np.tril(np.ones((n, num_states)), k=num_states - n)
    lb = np.concatenate(([-1], np.asarray(state_idx)[:-1]))
    return np.where(np.arange(num_states) > lb[:, None], m, 0.0)


def masked_logits(params, sp_indices, state_idx, num_layers=2, num_heads=4):
    """src/sampler.py:12-29 -> (n, num_states)."""
    state_idx = np.asarray(state_idx)
    n, M = state_idx.shape[0], sp_indices.shape[0]
    logits = transformer(params, np.asarray(sp_indices, np.float64)[state_idx], num_layers, num_heads)
    return np.where(mask(state_idx, n, M) != 0, logits, -1e50)


def log_prob(params, sp_indices, state_idx, num_layers=2, num_heads=4):
    """src/sampler.py:40-44."""
    lg = masked_logits(params, sp_indices, state_idx, num_layers, num_heads)
    mx = lg.max(axis=-1, keepdims=True)
    lsm = lg - (mx + np.log(np.exp(lg - mx).sum(axis=-1, keepdims=True)))
    return float(lsm[np.arange(lg.shape[0]), np.asarray(state_idx)].sum())


def log_prob_batch(params, sp_indices, state_indices, **kw):
    return np.array([log_prob(params, sp_indices, s, **kw) for s in np.asarray(state_indices)])


def sample(params, sp_indices, n, gumbel, num_layers=2, num_heads=4):
    """src/sampler.py:30-38 with categorical(key, logits) = argmax(logits + gumbel).  gumbel (batch, n, num_states)."""
    gumbel = np.asarray(gumbel, np.float64)
    B = gumbel.shape[0]
    out = np.zeros((B, n), dtype=np.int32)
    for b in range(B):
        for i in range(n):
            lg = masked_logits(params, sp_indices, out[b], num_layers, num_heads)
            out[b, i] = int(np.argmax(lg[i] + gumbel[b, i]))
    return out


def pack_params(params, num_layers=2, prefix="transformer"):
    """Flat float64 vector in the order of include/cgb200.h (cgb_van_*): x1hat, embedding w b, per layer
    [query w b, key w b, value w b, linear w b, mlp linear w b, mlp linear_1 w b], output w b."""
    parts = [np.asarray(params[prefix]["x1hat"], np.float64).ravel()]
    def wb(name):
        parts.append(np.asarray(params[name]["w"], np.float64).ravel())
        parts.append(np.asarray(params[name]["b"], np.float64).ravel())
    wb(prefix + "/embedding_mlp")
    for i in range(num_layers):
        for leaf in ("query", "key", "value", "linear"):
            wb("%s/layer%d_attn/%s" % (prefix, i, leaf))
        wb("%s/layer%d_mlp/linear" % (prefix, i))
        wb("%s/layer%d_mlp/linear_1" % (prefix, i))
    wb(prefix + "/output_mlp")
    return np.concatenate(parts)

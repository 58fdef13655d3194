"""TEST INFRASTRUCTURE (oracle): Philox4x32-10 and the mapping of its words to the throughput mode's proposal noise and
Metropolis uniforms, restated in NumPy.  Only tests/ may import this.

The reference draws its noise from jax's threefry (src/MCMC.py:24-29) -- third-party arithmetic that is absent here and
"parity unpinned" (DESIGN.md §5); parity mode therefore takes the noise as an input.  The kernels' own generator is
Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), pinned by the
known-answer vectors of the Random123 distribution (kat_vectors: philox4x32 10) below."""
import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85

# (counter[4], key[2]) -> output[4]   — Random123 kat_vectors, "philox4x32 10"
KAT = [((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
        (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff, 0xffffffff, 0xffffffff, 0xffffffff), (0xffffffff, 0xffffffff),
        (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


def philox4x32_10(ctr, key):
    """ctr (..., 4), key (..., 2) uint32 arrays -> (..., 4) uint32."""
    c = [np.asarray(ctr)[..., i].astype(np.uint64) for i in range(4)]
    k = [np.asarray(key)[..., i].astype(np.uint64) for i in range(2)]
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = np.uint64(M0) * c[0], np.uint64(M1) * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & mask, p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ k[0], lo1, hi0 ^ c[3] ^ k[1], lo0]
        k = [(k[0] + np.uint64(W0)) & mask, (k[1] + np.uint64(W1)) & mask]
    return np.stack(c, axis=-1).astype(np.uint32)


def _key(seed):
    return np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)


def uniforms(seed, B, S):
    """Metropolis uniforms (S, B) of the sweep kernels: counter (b, s+1, 0xffffffff, 1), top 53 bits of words 0|1."""
    b, s = np.meshgrid(np.arange(B, dtype=np.uint32), np.arange(1, S + 1, dtype=np.uint32))
    ctr = np.stack([b, s, np.full_like(b, 0xFFFFFFFF), np.ones_like(b)], axis=-1)
    r = philox4x32_10(ctr, np.broadcast_to(_key(seed), ctr.shape[:-1] + (2,))).astype(np.uint64)
    return ((r[..., 0] << np.uint64(32) | r[..., 1]) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53


def normals(seed, B, S, n):
    """Proposal noise (S, B, n, 2): counter (b, s+1, i, 0); Box-Muller on u1 = (top 23 bits of word 0 + 1/2) 2^-23 (refined by
    word 2 when those bits are all zero), u2 = (top 23 bits of word 1 + 1/2) 2^-23: (r cos 2 pi u2, r sin 2 pi u2)."""
    s, b, i = np.meshgrid(np.arange(1, S + 1, dtype=np.uint32), np.arange(B, dtype=np.uint32), np.arange(n, dtype=np.uint32),
                          indexing="ij")
    ctr = np.stack([b, s, i, np.zeros_like(b)], axis=-1)
    r = philox4x32_10(ctr, np.broadcast_to(_key(seed), ctr.shape[:-1] + (2,)))
    k = (r[..., 0] >> 9).astype(np.float64)
    u1 = np.where(k > 0, (k + 0.5) * 2.0 ** -23, ((r[..., 2] >> 9).astype(np.float64) + 0.5) * 2.0 ** -46)
    u2 = ((r[..., 1] >> 9).astype(np.float64) + 0.5) * 2.0 ** -23
    rad = np.sqrt(-2.0 * np.log(u1))
    return np.stack([rad * np.cos(2 * np.pi * u2), rad * np.sin(2 * np.pi * u2)], axis=-1)

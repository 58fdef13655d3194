"""Data-parallel plumbing of the hot path: one process per GPU replaces the reference's single-process
jax.pmap (src/VMC.py:8-11).  Walkers shard by contiguous blocks exactly as main.py:231-238 does
(batch // num_devices per device, batch % num_devices == 0 enforced); parameters, k-table and G are replicated.
The only cross-device traffic on the path is the mean of per-device scalars (jax.lax.pmean at src/MCMC.py:39
and the V entries of src/VMC.py:46-53): one all-reduce of a tiny float64 vector over the default
torch.distributed group (NCCL over NVLink on GPUs, gloo in the CPU tests).  Pure torch — no CUDA needed."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(batch, num_devices=None, rank=None):
    """[lo, hi) of this rank's walkers.  Mirrors main.py:231-238: the batch must divide evenly."""
    r, w = world()
    num_devices = w if num_devices is None else num_devices
    rank = r if rank is None else rank
    if batch % num_devices != 0:
        raise ValueError("Batch size must be divisible by the number of GPU devices. "
                         "Got batch = %d for %d devices now." % (batch, num_devices))
    per = batch // num_devices
    return rank * per, (rank + 1) * per


def psum(vec):
    """In-place sum over ranks of a small tensor (no-op for a single process)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM)
    return vec


def pmean_of_means(local_means):
    """jax.lax.pmean of per-device means (src/MCMC.py:39, src/VMC.py:46): equal shards, so the mean of the
    per-device means is the global mean.  local_means: 1-D float64 tensor."""
    _, w = world()
    return psum(local_means.clone()) / w


def accept_and_energy_stats(n_accept, n_proposals, V):
    """One all-reduce for the hot path's statistics: [accepted, proposed, sum V, sum V^2, walkers] ->
    (accept_rate, V_mean, V2_mean) over ALL ranks."""
    V = V.to(torch.float64)
    vec = torch.stack([torch.as_tensor(n_accept, dtype=torch.float64, device=V.device).reshape(()),
                       torch.as_tensor(float(n_proposals), dtype=torch.float64, device=V.device),
                       V.sum(), (V * V).sum(),
                       torch.as_tensor(float(V.numel()), dtype=torch.float64, device=V.device)])
    psum(vec)
    return vec[0] / vec[1], vec[2] / vec[4], vec[3] / vec[4]

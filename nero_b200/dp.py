"""Ray-sharded data parallelism for the stage-I hot path: one process per GPU, weights replicated, every rank
renders rows [rank*R/W, (rank+1)*R/W) of the global ray batch, ONE all-reduce of the flat fp32 gradient buffer per
step (NCCL over NVLink; gloo in the CPU tests).  The reference has no multi-GPU path at all
(train/trainer.py:68-69 raises NotImplementedError) -- SURVEY.md section 8e.

Equivalence with a single-GPU step on the whole batch: `loss_rgb` is a mean over rays (equal shards -> mean of
means is exact); `gradient_error` is a mean over the data-dependent number of inner samples N_in, so each rank
rescales its eikonal term by W*N_in_local/sum(N_in) (one extra scalar all-reduce) before backward; `loss_occ`
is a mean over <= 2048 points selected per call and stays a per-rank mean (documented deviation).
"""
import torch
import torch.distributed as dist


def shard_slice(n_rays, rank, world):
    per = n_rays // world
    return slice(rank * per, (rank + 1) * per)


def global_mean_weight(n_local, world, group=None):
    """Factor that turns `mean over my N_in samples`, after averaging over ranks, into the mean over ALL samples."""
    if world == 1:
        return 1.0
    t = torch.tensor([float(n_local)], device='cuda' if dist.get_backend(group) == 'nccl' else 'cpu')
    dist.all_reduce(t, group=group)
    return world * float(n_local) / max(float(t.item()), 1.0)


def sync_gradients(flat, world, group=None):
    """In-place average of the flat gradient buffer across ranks (a single collective)."""
    if world > 1:
        dist.all_reduce(flat, group=group)
        flat.div_(world)
    return flat


def flatten_grads(params):
    """For modules whose grads are not already views of one flat buffer (CPU tests with the oracle)."""
    return torch.cat([p.grad.reshape(-1) for p in params])

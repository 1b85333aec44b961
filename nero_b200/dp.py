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
    """Factor that turns `mean over my N_in samples`, after averaging over ranks, into the mean over ALL samples:
    W * N_in_local / sum_r N_in_r.  Returned as a 0-dim tensor on the collective's device (no host sync); 1.0 for W = 1."""
    if world == 1:
        return 1.0
    dev = 'cuda' if dist.get_backend(group) == 'nccl' else 'cpu'
    # n_local may be the engine's device-side count (graph mode: it is never read back to the host)
    t = n_local.detach().reshape(1).to(dev, torch.float32) if torch.is_tensor(n_local) else torch.tensor([float(n_local)], device=dev)
    tot = t.clone()
    dist.all_reduce(tot, group=group)
    return (world * t / torch.clamp(tot, min=1.0))[0]


def sync_gradients(flat, world, group=None, params=None):
    """In-place average of the flat gradient buffer across ranks (a single collective).  With `params` given, every
    parameter's .grad must be a view into `flat` in parameter order (engine.Grads hands them out that way): a gradient
    that lives elsewhere (e.g. created by an external backward before Grads.ensure ran) would silently be skipped by
    the all-reduce and let the ranks diverge, so it is packed into the buffer first."""
    if params is not None:
        pack_stray_grads(flat, params)
    if world > 1:
        dist.all_reduce(flat, group=group)
        flat.div_(world)
    return flat


def pack_stray_grads(flat, params):
    """Make every p.grad a view of `flat` (parameter order), copying gradients that were allocated elsewhere."""
    off = 0
    for p in params:
        n = p.numel()
        view = flat[off:off + n].view_as(p)
        if p.grad is None:
            view.zero_()
            p.grad = view
        elif p.grad.data_ptr() != view.data_ptr():
            view.copy_(p.grad)
            p.grad = view
        off += n
    assert off == flat.numel(), 'flat gradient buffer does not match the parameter list'
    return flat


def param_checksum_spread(params, world, group=None):
    """max - min over ranks of sum(|p|) over all parameters (0 when the replicas are identical)."""
    s = torch.stack([p.detach().double().abs().sum() for p in params]).sum().reshape(1)
    if world == 1:
        return 0.0
    lo, hi = s.clone(), s.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
    return float((hi - lo).item())


def flatten_grads(params):
    """For modules whose grads are not already views of one flat buffer (CPU tests with the oracle)."""
    return torch.cat([p.grad.reshape(-1) for p in params])

"""Multi-GPU partitioning of the ray-march path: one process per GPU, a grid replica per GPU.

Render: rays are independent given a read-only grid (the reference exploits the same fact with `omp parallel for`
over rows, example_tracer.cpp:630), so VIEWS are dealt round-robin to ranks and no collective sits on the data path.
Training: data-parallel over the ray batch; each rank accumulates its shard's gradient into its own 13*N-float
gradient planes and ONE exchange step — a sum all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests) —
makes every replica's gradient the full-batch gradient.  The reference's loss is an un-normalised sum over rays
(example_tracer.cpp:258-268), so data parallelism is mathematically a larger single-GPU batch; TV + Adam then run
replicated and keep the replicas bit-identical.
"""


def shard_views(n_views, rank, world):
    """global view indices rendered by `rank`: round-robin, every view exactly once over the ranks"""
    return list(range(rank, n_views, world))


def shard_rays(n_rays, rank, world):
    """[begin, end) of the contiguous slice of a packed ray batch that `rank` back-propagates"""
    base, extra = divmod(n_rays, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_gradient(grad, dist):
    """sum the gradient planes over the ranks in place (no-op on one rank); grad = torch tensor viewing the planes"""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(grad, op=dist.ReduceOp.SUM)
    return grad

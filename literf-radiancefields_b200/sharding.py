"""Multi-GPU partitioning of the ray-march path: one process per GPU, a grid replica per GPU.

Render: rays are independent given a read-only grid (the reference exploits the same fact with `omp parallel for`
over rows, example_tracer.cpp:630), so VIEWS are dealt round-robin to ranks and no collective sits on the data path.
Training: data-parallel over the ray batch; each rank accumulates its shard's gradient into its own 13*N-float
gradient planes and ONE exchange step — a sum all-reduce (NCCL over NVLink on the GPU box, gloo in the CPU tests) —
makes every replica's gradient the full-batch gradient.  The reference's loss is an un-normalised sum over rays
(example_tracer.cpp:258-268), so data parallelism is mathematically a larger single-GPU batch; TV + Adam then run
replicated and keep the replicas bit-identical.

PeerExchange is the faster exchange on an NVLink box: the planes of every rank are mapped into every process (CUDA IPC),
and rank r does TV + gradient sum + Adam for z-slab r in one kernel that reads the other ranks' gradients and writes
every replica's parameters through peer memory (lrf_train_percell_peer_async): reduce-scatter + Adam + all-gather fused,
with the Adam work divided by the number of GPUs.  The replicas stay bit-identical because each parameter is computed
once, by its slab's owner.
"""


def shard_views(n_views, rank, world):
    """global view indices rendered by `rank`: round-robin, every view exactly once over the ranks"""
    return list(range(rank, n_views, world))


def shard_rows(height, rank, world, align=4):
    """[y0, y1) of the row band of ONE frame that `rank` renders (latency mode).  Bands are multiples of `align` rows (a warp of the
    render kernel covers an 8 x 4 pixel tile) except the last; every row belongs to exactly one rank; ranks may get an empty band."""
    units = (height + align - 1) // align
    base, extra = divmod(units, world)
    u0 = rank * base + min(rank, extra)
    u1 = u0 + base + (1 if rank < extra else 0)
    return min(u0 * align, height), min(u1 * align, height)


def gather_frame(frame, height, width, dist):
    """latency mode: every rank has rendered its row band into its own full-frame tensor `frame` (uint32/int32 [height * width]);
    rank 0 ends up with the complete frame.  One collective on the data path: a gather of the bands (all-gather into rank-ordered
    slices when the bands are equal, point-to-point band copies otherwise)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return frame
    world, rank = dist.get_world_size(), dist.get_rank()
    bands = [shard_rows(height, r, world) for r in range(world)]
    if rank == 0:
        reqs = [dist.irecv(frame[y0 * width:y1 * width], src=r) for r, (y0, y1) in enumerate(bands) if r != 0 and y1 > y0]
        for q in reqs:
            q.wait()
    else:
        y0, y1 = bands[rank]
        if y1 > y0:
            dist.send(frame[y0 * width:y1 * width], dst=0)
    return frame


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


class PeerExchange:
    """per-iteration exchange over NVLink peer memory.  `tracer` = this rank's RayMarcherExample, `dist` = an initialised
    torch.distributed (NCCL) group used for the handle exchange and as the inter-GPU barrier on the tracer's stream."""

    def __init__(self, tracer, dist, torch):
        self.t, self.dist, self.torch = tracer, dist, torch
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        mine = (tracer.ipc_export(0), tracer.ipc_export(1), tracer.ipc_export(2))
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine)
        tracer.peer_attach_ipc(self.world, self.rank, [e[0] for e in everyone], [e[1] for e in everyone])
        tracer.peer_attach_flags_ipc([e[2] for e in everyone])
        self.stream = torch.cuda.ExternalStream(tracer.lib.lrf_stream(tracer.h))
        self.flag = torch.zeros(1, device="cuda")
        self.use_nccl_barrier = False
        dist.barrier()                       # every rank has mapped every flag array before the first epoch is stored

    def barrier(self):
        """all ranks' work queued so far on their tracer streams is complete before anything queued after it starts.
        Peer-memory flag barrier (lrf_peer_barrier_async, ~10 us); the 1-element NCCL all-reduce it replaces (~170 us at
        8 GPUs) stays available as use_nccl_barrier."""
        if self.use_nccl_barrier:
            with self.torch.cuda.stream(self.stream):
                self.dist.all_reduce(self.flag)
        else:
            self.t.peer_barrier_async()

    def percell(self, it, tv_done=False):
        """between the backward of iteration `it` and the next one.  tv_done: the TV gradient of the own slab is already in
        the local gradient (lrf_rays_grad_tv_dev_async ran it beside the backward)"""
        self.barrier()                       # every rank's gradient is complete
        if tv_done:
            self.t.adam_peer_async(it)       # own slab: sum over ranks + Adam, parameters written to every replica
        else:
            self.t.train_percell_peer_async(it)  # the same preceded by the slab's TV gradient
        self.barrier()                       # every replica is fully written, every gradient fully read
        self.t.zero_grad_async()

    def close(self):
        self.t.sync()
        self.dist.barrier()                  # nobody unmaps while a peer's kernel may still store into the arrays
        timed_out = self.t.peer_barrier_timed_out()
        self.t.peer_detach()
        self.torch.cuda.synchronize()
        self.dist.barrier()                  # every rank has unmapped its peers before anyone frees or re-allocates its planes
        if timed_out:
            raise RuntimeError("a peer-memory barrier timed out: a rank did not arrive within 2 s")


class McExchange(PeerExchange):
    """the per-iteration exchange through NVLink multicast objects (include/lrf.h lrf_mc_*; adam_mc_kernel): multimem.ld_reduce sums
    the gradient replicas in the switch, multimem.st writes the new parameters into every replica.  Same call sequence as
    PeerExchange (barrier, own-slab pass, barrier, zero); the barrier is the same peer-memory flag barrier.  The multicast object is
    created by rank 0 and reaches the other processes as a POSIX file descriptor over a unix socket (SCM_RIGHTS)."""

    def __init__(self, tracer, dist, torch):
        import os
        import socket
        self.t, self.dist, self.torch = tracer, dist, torch
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        if not tracer.mc_supported():
            raise RuntimeError("NVLink multicast objects are not supported on this device / driver")
        tracer.sync()
        torch.cuda.synchronize()
        dist.barrier()                       # nothing of an earlier group is in flight anywhere: the planes are about to move
        tracer.mc_prepare(self.world, self.rank)
        path = [None]
        srv = None
        if self.rank == 0:
            path[0] = "/tmp/lrf_mc_%d_%s.sock" % (os.getpid(), os.environ.get("MASTER_PORT", "0"))
            if os.path.exists(path[0]):
                os.unlink(path[0])
            srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            srv.bind(path[0])
            srv.listen(self.world)
        dist.broadcast_object_list(path, src=0)
        if self.rank == 0:
            fd = tracer.mc_create()
            for _ in range(self.world - 1):
                conn, _ = srv.accept()
                socket.send_fds(conn, [b"mc"], [fd])
                conn.recv(1)                  # the peer has imported the object
                conn.close()
            os.close(fd)
            srv.close()
            os.unlink(path[0])
        else:
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(path[0])
            _, fds, _, _ = socket.recv_fds(c, 16, 1)
            tracer.mc_import(fds[0])
            os.close(fds[0])
            c.send(b"k")
            c.close()
        tracer.mc_add_device()
        dist.barrier()                       # every device is part of the object before anyone binds memory to it
        tracer.mc_bind()
        flags = [None] * self.world
        dist.all_gather_object(flags, tracer.ipc_export(2))
        tracer.peer_attach_flags_ipc(flags)
        self.stream = torch.cuda.ExternalStream(tracer.lib.lrf_stream(tracer.h))
        self.flag = torch.zeros(1, device="cuda")
        self.use_nccl_barrier = False
        dist.barrier()                       # every rank has bound its planes and mapped every flag array

"""Batch data-parallel sharding of the scattering transform over the GPUs of one box.

The path shards naturally over the batch: every signal is independent (the reference core never
mixes batch entries - all primitives act on trailing axes, kymatio/scattering2d/backend/
torch_backend.py:118-127) and the only shared state is the read-only filter bank, replicated per
rank.  One process per GPU (torchrun / torch.distributed); no collective on the compute path.  The
only exchange is an optional all-gather of the coefficient blocks when the caller wants one tensor
on every rank; its backward is the matching reduce-scatter.
"""
import torch
import torch.distributed as dist

__all__ = ["shard_bounds", "shard_batch", "gather_batch", "ShardedScattering", "PeerGatherScattering"]


def shard_bounds(n, rank, world):
    """Contiguous, balanced slice [lo, hi) of n items for `rank` (first n % world ranks get one more)."""
    base, extra = divmod(int(n), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(x, rank=None, world=None, group=None):
    """This rank's slice of a batch that every rank holds in full (dim 0)."""
    if world is None:
        world = dist.get_world_size(group) if dist.is_initialized() else 1
    if rank is None:
        rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


class _AllGatherBatch(torch.autograd.Function):
    """all_gather along dim 0 with per-rank sizes from shard_bounds; backward = reduce-scatter (sum)."""

    @staticmethod
    def forward(ctx, x, total, group):
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        ctx.group, ctx.total = group, total
        bounds = [shard_bounds(total, r, world) for r in range(world)]
        assert x.shape[0] == bounds[rank][1] - bounds[rank][0], "local batch does not match shard_bounds"
        out = x.new_empty((total,) + tuple(x.shape[1:]))
        if total % world == 0:
            dist.all_gather_into_tensor(out, x.contiguous(), group=group)
        else:
            # uneven shards: pad every block to the largest one, gather, then compact
            maxn = max(hi - lo for lo, hi in bounds)
            pad = x.new_zeros((maxn,) + tuple(x.shape[1:]))
            pad[: x.shape[0]] = x
            buf = x.new_empty((world * maxn,) + tuple(x.shape[1:]))
            dist.all_gather_into_tensor(buf, pad, group=group)
            for r, (lo, hi) in enumerate(bounds):
                out[lo:hi] = buf[r * maxn: r * maxn + (hi - lo)]
        return out

    @staticmethod
    def backward(ctx, grad):
        group, total = ctx.group, ctx.total
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        grad = grad.contiguous()
        lo, hi = shard_bounds(total, rank, world)
        backend = dist.get_backend(group)
        if total % world == 0 and backend == "nccl":
            out = grad.new_empty((hi - lo,) + tuple(grad.shape[1:]))
            dist.reduce_scatter_tensor(out, grad, op=dist.ReduceOp.SUM, group=group)
            return out, None, None
        # gloo (CPU tests) / uneven shards: all-reduce then slice (on a copy: `grad` may be the caller's own tensor)
        grad = grad.clone()
        dist.all_reduce(grad, op=dist.ReduceOp.SUM, group=group)
        return grad[lo:hi].clone(), None, None


def gather_batch(x_local, total=None, group=None):
    """Differentiable all-gather of per-rank batch slices into the full (total, ...) tensor."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return x_local
    if total is None:
        t = torch.tensor([x_local.shape[0]], device=x_local.device, dtype=torch.int64)
        dist.all_reduce(t, group=group)
        total = int(t.item())
    return _AllGatherBatch.apply(x_local, total, group)


class ShardedScattering(torch.nn.Module):
    """Wrap a scattering module: each rank transforms its slice of the batch; `gather=True` returns the
    full coefficient tensor on every rank (one all-gather), `gather=False` returns the local block."""

    def __init__(self, scattering, gather=True, group=None, input_is_sharded=False):
        super().__init__()
        self.scattering, self.gather, self.group = scattering, gather, group
        self.input_is_sharded = input_is_sharded

    def forward(self, x):
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return self.scattering(x)
        if self.input_is_sharded:
            local, total = x, None
        else:
            local, total = shard_batch(x, group=self.group).contiguous(), x.shape[0]
        y = self.scattering(local)
        return gather_batch(y, total, self.group) if self.gather else y


class PeerGatherScattering(torch.nn.Module):
    """Batch-sharded 2-D scattering whose all-gather is FUSED into the kernels that produce the coefficients.

    Every rank owns a (total_batch, K, oh, ow) result buffer in symmetric memory (torch.distributed._symmetric_memory:
    each rank's buffer is mapped into every other rank's address space over NVLink / NVSwitch).  A rank transforms its
    slice of the batch with ``scat_plan2d_forward_peers``: the low-pass stages of the tile kernels store every coefficient
    plane into the rank's own buffer AND into the same block of every peer's buffer (remote stores over NVLink, a few KB
    per path, issued while the SMs are busy with the next paths).  One cross-rank barrier later every rank holds the full
    tensor - there is no separate gather pass over the 227 MB per rank that NCCL's all_gather would read again.
    No-grad / inference path (``ShardedScattering`` keeps the differentiable NCCL gather); float32; ``scattering`` is a
    ``kymatio_b200.Scattering2D``.  Two buffers alternate so that a rank may still read step i while step i+1 is written.
    """

    def __init__(self, scattering, group=None, multicast=None):
        """multicast: True / False / None (= use NVLS multimem stores when the symmetric-memory handle offers a multicast
        address and there are more than two ranks: one store per value instead of one per peer)."""
        super().__init__()
        self.scattering, self.group, self.multicast = scattering, group, multicast
        self._bufs, self._hdls, self._step = {}, {}, 0
        self.last_mode = None

    def _symm_buffers(self, total, eng, device):
        import torch.distributed._symmetric_memory as symm_mem
        key = (total, eng.K, eng.out_h, eng.out_w, device.index)
        if key not in self._bufs:
            group = self.group if self.group is not None else dist.group.WORLD
            bufs, hdls = [], []
            for _ in range(2):
                t = symm_mem.empty((total, eng.K, eng.out_h, eng.out_w), dtype=torch.float32, device=device)
                hdls.append(symm_mem.rendezvous(t, group))
                bufs.append(t)
            self._bufs[key], self._hdls[key] = bufs, hdls
        return self._bufs[key], self._hdls[key]

    @torch.no_grad()
    def forward(self, x_local, total=None):
        """x_local: this rank's contiguous slice (shard_bounds) of the batch -> the full (total, K, oh, ow) tensor."""
        S = self.scattering
        if not dist.is_initialized() or dist.get_world_size(self.group) == 1:
            return S(x_local)
        world, rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        if total is None:
            t = torch.tensor([x_local.shape[0]], device=x_local.device, dtype=torch.int64)
            dist.all_reduce(t, group=self.group)
            total = int(t.item())
        lo, hi = shard_bounds(total, rank, world)
        assert x_local.shape[0] == hi - lo, "local batch does not match shard_bounds"
        x = x_local.reshape((-1,) + tuple(x_local.shape[-2:])).contiguous()
        eng = S._engine(x.dtype, x.device)
        phi, psi = S.load_filters()
        eng.bind(phi, psi)
        bufs, hdls = self._symm_buffers(total, eng, x.device)
        i = self._step & 1
        self._step += 1
        buf, hdl = bufs[i], hdls[i]
        block = eng.K * eng.out_h * eng.out_w * 4
        mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
        use_mc = (self.multicast if self.multicast is not None else world > 2) and mc != 0
        if use_mc:
            self.last_mode = "multicast"
            eng.forward(x, out=buf[lo:hi], multicast_ptr=mc + lo * block)
        else:
            self.last_mode = "unicast"
            peers = [int(hdl.buffer_ptrs[r]) + lo * block for r in range(world) if r != rank]
            eng.forward(x, out=buf[lo:hi], peer_ptrs=peers)
        hdl.barrier(channel=0)          # every rank's kernels (and their remote stores) have completed
        return buf

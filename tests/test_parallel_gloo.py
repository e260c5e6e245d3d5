"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the batch sharding and the
differentiable gather (the CUDA scattering itself is replaced by a stand-in per-sample op)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from kymatio_b200.parallel import ShardedScattering, gather_batch, shard_batch, shard_bounds


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 5, 8, 256, 4096, 4099):
        for world in (1, 2, 3, 4, 8):
            bounds = [shard_bounds(n, r, world) for r in range(world)]
            assert bounds[0][0] == 0 and bounds[-1][1] == n
            assert all(bounds[i][1] == bounds[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in bounds]
            assert max(sizes) - min(sizes) <= 1


class _PerSample(torch.nn.Module):
    """Stand-in for the scattering: independent per batch entry, like the real transform."""

    def forward(self, x):
        return torch.stack([x.abs().sum(dim=(-1, -2)), (x ** 2).mean(dim=(-1, -2))], dim=1)


def _worker(rank, world, port, total):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        x = torch.randn(total, 6, 5, requires_grad=True)
        ref = _PerSample()(x)
        wrapped = ShardedScattering(_PerSample(), gather=True)
        y = wrapped(x)
        assert y.shape == ref.shape and torch.allclose(y, ref)
        # every rank back-propagates the same full-tensor loss: gradients sum over ranks
        w = torch.arange(ref.numel(), dtype=torch.float32).reshape(ref.shape)
        (y * w).sum().backward()
        gref, = torch.autograd.grad((ref * w).sum(), x)
        lo, hi = shard_bounds(total, rank, world)
        assert torch.allclose(x.grad[lo:hi], world * gref[lo:hi], atol=1e-5)
        # local (no gather) mode and explicit pieces
        yl = ShardedScattering(_PerSample(), gather=False)(x.detach())
        assert torch.allclose(yl, ref[lo:hi])
        assert torch.equal(shard_batch(x.detach()), x.detach()[lo:hi])
        assert torch.allclose(gather_batch(yl), ref.detach())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [8, 7])
def test_sharded_gather_world2(total):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, total), nprocs=2, join=True)

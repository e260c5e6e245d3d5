"""torchrun check: batch-sharded scattering + NCCL gather equals the single-GPU result; gradients too."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kymatio_b200 import Scattering2D  # noqa: E402
from kymatio_b200.parallel import ShardedScattering, shard_bounds  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(1)
S = Scattering2D(2, (32, 32)).cuda()
for total in (8, 7):
    x = torch.randn(total, 32, 32, device="cuda")
    dist.broadcast(x, 0)
    ref = S(x)
    y = ShardedScattering(S, gather=True)(x)
    assert y.shape == ref.shape and torch.allclose(y, ref, atol=1e-6), (total, float((y - ref).abs().max()))
    xg = x.clone().requires_grad_(True)
    (ShardedScattering(S, gather=True)(xg) ** 2).sum().backward()
    xr = x.clone().requires_grad_(True)
    (S(xr) ** 2).sum().backward()
    lo, hi = shard_bounds(total, rank, world)
    assert torch.allclose(xg.grad[lo:hi], world * xr.grad[lo:hi], rtol=1e-3, atol=1e-5)
    assert float(xg.grad[:lo].abs().sum() + xg.grad[hi:].abs().sum()) == 0.0
# the all-gather fused into the producing kernels (peer stores over NVLink into symmetric memory)
from kymatio_b200.parallel import PeerGatherScattering  # noqa: E402
try:
    P = PeerGatherScattering(S)
    for total in (8, 7, 8):
        x = torch.randn(total, 32, 32, device="cuda")
        dist.broadcast(x, 0)
        ref = S(x)
        lo, hi = shard_bounds(total, rank, world)
        y = P(x[lo:hi].contiguous(), total)
        torch.cuda.synchronize()
        assert y.shape == ref.shape and torch.allclose(y, ref, atol=1e-6), ("peer gather", total, float((y - ref).abs().max()))
    S3 = Scattering2D(3, (64, 64)).cuda()
    for mcast in (True, False):
        Pm = PeerGatherScattering(S3, multicast=mcast)
        x = torch.randn(5 * world, 64, 64, device="cuda")
        dist.broadcast(x, 0)
        lo, hi = shard_bounds(5 * world, rank, world)
        y = Pm(x[lo:hi].contiguous(), 5 * world)
        torch.cuda.synchronize()
        assert torch.allclose(y, S3(x), atol=1e-6), ("peer gather", mcast, Pm.last_mode)
        if rank == 0:
            print("peer_gather mode", Pm.last_mode, "ok")
    P3 = PeerGatherScattering(S3)
    x = torch.randn(6 * world, 64, 64, device="cuda")
    dist.broadcast(x, 0)
    lo, hi = shard_bounds(6 * world, rank, world)
    y = P3(x[lo:hi].contiguous(), 6 * world)
    torch.cuda.synchronize()
    assert torch.allclose(y, S3(x), atol=1e-6)
    if rank == 0:
        print("peer_gather ok: world", world)
except ImportError as e:                      # symmetric memory not in this torch build
    if rank == 0:
        print("peer_gather skipped:", e)
if rank == 0:
    print("dist_check ok: world", world)
dist.destroy_process_group()

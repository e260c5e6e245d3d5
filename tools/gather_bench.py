"""torchrun: cost of handing every rank the full coefficient tensor of the batch-sharded C2 forward (256 images/GPU):
local only vs NCCL all_gather vs peer stores (unicast) vs NVLS multicast stores fused into the producing kernels."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from kymatio_b200 import Scattering2D  # noqa: E402
from kymatio_b200.parallel import PeerGatherScattering, gather_batch  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
S = Scattering2D(3, (256, 256)).to(dev)
x = torch.randn(B, 256, 256, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, steps=8, warm=4):
    with torch.no_grad():
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for e0, e1 in ev:
            flush.zero_()
            e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
    t = torch.tensor([sum(a.elapsed_time(b) for a, b in ev) / steps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


out = {"world": world, "batch_per_gpu": B, "local_only_ms": timed(lambda: S(x)),
       "nccl_all_gather_ms": timed(lambda: gather_batch(S(x), world * B))}
for name, mc in (("peer_unicast_ms", False), ("peer_multicast_ms", True)):
    P = PeerGatherScattering(S, multicast=mc)
    out[name] = timed(lambda: P(x, world * B))
    out[name.replace("_ms", "_mode")] = P.last_mode
    y = P(x, world * B)
    torch.cuda.synchronize()
    out[name.replace("_ms", "_own_block_exact")] = bool(torch.equal(y[rank * B:(rank + 1) * B], S(x)))
    del P, y
if rank == 0:
    print(json.dumps(out))
dist.destroy_process_group()

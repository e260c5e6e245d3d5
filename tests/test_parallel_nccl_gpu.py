"""The one collective north_star names - the coefficient all-gather of the batch-sharded scattering - on real GPUs
over NCCL: tools/dist_check.py under torchrun with 2 ranks (values and gradients against the single-GPU result).
Self-skips on a box with fewer than 2 GPUs."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2])
def test_sharded_gather_nccl(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "dist_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "dist_check ok: world %d" % world in r.stdout


def test_second_device_in_one_process():
    """One process driving a GPU other than cuda:0 while cuda:0 is current: the large-shared-memory opt-in is per device
    (csrc/common.cuh once_per_device) and the 2-D plan must be created under the target device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import numpy as np
    from kymatio_b200 import Scattering2D
    x = torch.randn(3, 64, 64)
    S0 = Scattering2D(3, (64, 64)).to("cuda:0")
    y0 = S0(x.to("cuda:0"))
    torch.cuda.set_device(0)
    S1 = Scattering2D(3, (64, 64)).to("cuda:1")
    y1 = S1(x.to("cuda:1"))
    assert y1.device.index == 1
    assert np.allclose(y0.cpu().numpy(), y1.cpu().numpy(), atol=1e-6)
    from conftest import import_reference
    if import_reference():
        import kymatio_b200.kymatio_plugin as plugin
        plugin.install()
        from kymatio.torch import HarmonicScattering3D, Scattering1D
        for dev in ("cuda:0", "cuda:1"):
            s1 = Scattering1D(J=5, shape=2048, Q=(4, 1), backend="torch_b200").to(dev)
            s3 = HarmonicScattering3D(J=1, shape=(16, 16, 16), L=1, backend="torch_b200").to(dev)
            a = s1(torch.ones(2, 2048, device=dev))
            b = s3(torch.ones(2, 16, 16, 16, device=dev))
            assert a.device == torch.device(dev) and b.device == torch.device(dev)
            assert torch.isfinite(a).all() and torch.isfinite(b).all()


def test_data_parallel_replicas_forward_and_backward():
    """nn.DataParallel over two GPUs (the reference's multi-GPU recipe: class-level backends shared by the replicas,
    SURVEY 8(b) threading): replicas run on their own threads, in the forward and - through autograd's per-device
    threads - in the backward, where each rebuilds its graph from its own kept spectra."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from kymatio_b200 import Scattering2D
    torch.manual_seed(5)
    S = Scattering2D(4, (224, 224)).cuda(0)
    x = torch.randn(6, 224, 224, device="cuda:0")
    w = torch.randn(6, S.forward(x[:1]).shape[1], 14, 14, device="cuda:0")

    def run(module):
        xi = x.clone().requires_grad_(True)
        y = module(xi)
        (y * w).sum().backward()
        return y.detach(), xi.grad

    y1, g1 = run(S)
    y2, g2 = run(torch.nn.DataParallel(S, device_ids=[0, 1]))
    assert torch.equal(y1, y2.to(y1.device))
    assert (g1 - g2).abs().max() <= 1e-5 * g1.abs().max()

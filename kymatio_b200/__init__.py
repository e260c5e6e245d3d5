"""kymatio_b200 - the ``torch_b200`` wavelet-scattering backend: hand-written sm_100a CUDA
(libscat_b200.so, C ABI in include/scat_b200.h) behind kymatio's torch frontends.

    from kymatio_b200 import Scattering2D
    S = Scattering2D(J=3, shape=(256, 256)).cuda()
    Sx = S(x)                       # x: (..., 256, 256) CUDA float32

To use it from an installed, unmodified kymatio instead::

    import kymatio_b200.kymatio_plugin as plugin; plugin.install()
    from kymatio.torch import Scattering2D
    S = Scattering2D(J=3, shape=(256, 256), backend='torch_b200').cuda()
"""
from .scattering2d import Scattering2D  # noqa: F401
from .graph import GraphedScattering  # noqa: F401

__version__ = "0.1.0"
__all__ = ["Scattering2D", "GraphedScattering"]

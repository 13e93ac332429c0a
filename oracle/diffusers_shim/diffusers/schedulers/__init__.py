import os
import sys

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "..", "..", "..")))
from oracle.vx_oracle import DDIM as _DDIM  # noqa: E402


class DDIMScheduler(_DDIM):
    """The restated DDIM of oracle/vx_oracle.py exposed under the diffusers name."""


class _Stub:
    def __init__(self, *a, **k):
        raise NotImplementedError


DPMSolverMultistepScheduler = EulerAncestralDiscreteScheduler = EulerDiscreteScheduler = _Stub
LMSDiscreteScheduler = PNDMScheduler = _Stub

"""B200-native drop-in for the hot path of PassiveTracerFlows.jl's TracerAdvectionDiffusion module.

``from ptf_b200 import TracerAdvectionDiffusion as TAD`` mirrors
``import PassiveTracerFlows.TracerAdvectionDiffusion`` with the new device type ``B200``.
All numerics run in libptf_b200.so (hand-written sm_100a CUDA + cuFFT); importing this package without the
built library raises — there is no CPU or PyTorch fallback.
"""
from . import _capi
from . import parallel
from . import tracer_advection_diffusion as TracerAdvectionDiffusion
from . import multilayerqg as MultiLayerQG
from .tracer_advection_diffusion import (B200, Device, ExpressionFlow, OneDAdvectingFlow, Problem, SeparableFlow,
                                         ThreeDAdvectingFlow, TracerProblem, TwoDAdvectingFlow, gridpoints, noflow,
                                         set_c, step_until, stepforward, updatevars)

__all__ = ["B200", "Device", "Problem", "set_c", "updatevars", "stepforward", "step_until", "OneDAdvectingFlow",
           "TwoDAdvectingFlow", "ThreeDAdvectingFlow", "SeparableFlow", "ExpressionFlow", "TracerProblem", "gridpoints", "noflow",
           "TracerAdvectionDiffusion", "MultiLayerQG"]

_capi.load()   # fail loudly at import time when the CUDA library is missing

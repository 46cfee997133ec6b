"""deepphysinet_b200 - the decoder-query + PDE-residual hot path of flyakon/DeepPhysiNet as
hand-written sm_100a CUDA behind the reference's own Python call surface.

    from deepphysinet_b200 import InterfacePhysics, PhysicsNet     # drop-in for the reference classes
    from deepphysinet_b200.functional import pde_residual           # the fused autograd.Function

See DESIGN.md for the path and its boundary, INTEGRATION.md for how the reference binds to it.
"""
from .config import PhysicsConsts  # noqa: F401
from .functional import DecoderWeights, decoder_values, pde_residual, set_default_mode  # noqa: F401
from .interface import InterfacePhysics  # noqa: F401
from .physics_net import PhysicsNet, VariableNet  # noqa: F401
from .trainer import TrainStep  # noqa: F401

__version__ = "0.1.0"

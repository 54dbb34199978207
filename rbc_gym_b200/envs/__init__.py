from .rbc2d import RBCField, RayleighBenardConvection2DEnv, noise_initial_fields
from .vector import RBCVectorEnv2D

__all__ = ["RBCField", "RayleighBenardConvection2DEnv", "RBCVectorEnv2D", "noise_initial_fields"]

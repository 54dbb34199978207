from .rbc2d import RBCField, RayleighBenardConvection2DEnv, noise_initial_fields
from .rbc3d import RayleighBenardConvection3DEnv, noise_initial_fields_3d
from .sb3 import RBCSB3VecEnv
from .vector import RBCVectorEnv2D, RBCVectorEnv3D

__all__ = ["RBCField", "RayleighBenardConvection2DEnv", "RayleighBenardConvection3DEnv", "RBCVectorEnv2D", "RBCVectorEnv3D", "RBCSB3VecEnv",
           "noise_initial_fields", "noise_initial_fields_3d"]

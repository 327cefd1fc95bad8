"""finite_element_method_b200 — B200-native (sm_100a) drop-in for the stiffness-assembly hot path of
RomanShushakov/finite_element_method: truss / beam / plate local stiffness + deterministic
assembly into the global FP64 CSR matrix. The product is libfemgpu.so (C ABI in include/femgpu.h);
this package is its Python host mirror (`FEM`) plus synthetic mesh generators for the benchmarks.
"""
from .fem import (BEAM, ELEMENT_RESULT_COMPONENTS, PLATE, TRUSS, FEM, DOFParameter, FemError,  # noqa: F401
                  SeparatedStiffnessMatrix, SeparatedStiffnessMatrixSparse)
from . import meshes  # noqa: F401

__all__ = ["FEM", "FemError", "DOFParameter", "SeparatedStiffnessMatrix", "SeparatedStiffnessMatrixSparse", "TRUSS", "BEAM", "PLATE",
           "ELEMENT_RESULT_COMPONENTS", "meshes"]

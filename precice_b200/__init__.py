"""precice_b200 — B200-native (sm_100a) backend for preCICE's radial-basis-function data mapping.

The product is ``librbfmap.so`` (hand-written CUDA kernels behind the C ABI of ``include/rbfmap.h``);
this package is the thin host-side mirror of the reference's mapping classes used by tests and bench.
"""
from .mapping import (MappingError, PartitionOfUnityMapping, RadialBasisFctMapping, as_xyz, device_count,  # noqa: F401
                      eval_basis, nccl_unique_id, partition_rows)

__all__ = ["MappingError", "PartitionOfUnityMapping", "RadialBasisFctMapping", "as_xyz", "device_count", "eval_basis", "nccl_unique_id",
           "partition_rows"]

"""blaze_b200 -- B200-native drop-in for ingonyama-zk/blaze's primitive clients.

Python host-side mirror of the reference's Rust surface (`driver_client`, `ingo_msm`,
`ingo_ntt`, `ingo_hash`) over the C ABI of libblaze_b200.so (include/blaze_b200.h).
Importing the package does not load the CUDA library; the first client constructor does, and it
raises if the library or a CUDA device is missing (no CPU fallback).
"""
from . import error                                           # noqa: F401
from .driver_client import CardType, DriverClient, DriverConfig, DriverPrimitive, DMA_RW   # noqa: F401
from .ingo_msm import (Curve, MSMClient, MSMImageParametrs, MSMInit, MSMInput, MSMParams, MSMResult,   # noqa: F401
                       PointMemoryType, PRECOMPUTE_FACTOR, PRECOMPUTE_FACTOR_BASE)

from .ingo_ntt import NTT, DistributedNTT, NTTClient, NTTInput, NttInit          # noqa: F401

from .ingo_hash import (Hash, PoseidonClient, PoseidonInitializeParameters, PoseidonResult, TreeMode,   # noqa: F401
                        num_of_elements_in_base_layer, num_of_elements_oct_tree)

__version__ = "0.1.0"

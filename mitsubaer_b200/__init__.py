"""mitsubaer_b200 — B200-native (sm_100a CUDA) implementation of MitsubaER's
refractive-radiative-transfer hot path behind a C ABI (include/mitsubaer_b200.h).

Importing this package loads mitsubaer_b200/libmitsubaer_b200.so; there is no CPU path.
"""
import os as _os

from . import _abi
from ._abi import MerError, lib

if _os.environ.get("MER_B200_DEFER_LOAD") != "1":
    lib.mer_abi_version()  # maps libmitsubaer_b200.so now: ImportError / AttributeError if it is missing or stale
from .plugins import (EikonalVolPathIntegrator, GridDataSource, HGPhaseFunction, HeterogeneousRefractiveMedium,
                      SplineDataSource, develop, make_volume_desc)
from . import fields

__all__ = ["MerError", "lib", "SplineDataSource", "GridDataSource", "HGPhaseFunction",
           "HeterogeneousRefractiveMedium", "EikonalVolPathIntegrator", "develop", "make_volume_desc", "fields"]


def device_count():
    return int(lib.mer_device_count())


def kernel_launch_count():
    return int(lib.mer_kernel_launch_count())

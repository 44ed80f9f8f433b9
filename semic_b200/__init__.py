"""semic_b200 -- B200-native SEMIC time step (surface_energy_and_mass_balance) behind the reference's API.

    semic_b200.surface_physics   the f90wrap-style module (also importable as `SurfacePhysics`)
    semic_b200.engine            device-resident series, multi-day runs, host-streamed runs, ensemble misfit
    semic_b200.driver            example.f90 / run_particles.f90 day loops on top of the engine

All numerics run in libsemic_b200.so (hand-written CUDA, sm_100a).  No CPU fallback exists.
"""
from . import _capi, surface_physics          # noqa: F401
from ._capi import SemicError                 # noqa: F401

__all__ = ["surface_physics", "engine", "driver", "SemicError"]

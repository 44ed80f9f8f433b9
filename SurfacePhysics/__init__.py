"""Drop-in name of the reference's f90wrap package (f2py/Makefile:49, f2py/test.py:3):

    import SurfacePhysics as sp
    sp.surface_physics.surface_energy_and_mass_balance(state, par, bnd, day, year)
"""
from semic_b200 import surface_physics          # noqa: F401

"""pychem_b200 -- B200-native (sm_100a, FP64) replacement for pychem's two-electron hot path:
ERI generation (Methods/_c_ints.c + Methods/c_ints/two_electron_*.c driven by
Methods/integrals.py) and J/K digestion (Methods/hartree_fock.py), behind the reference's own
call surface.  See DESIGN.md / INTEGRATION.md.
"""
__all__ = ["structures", "basis_table", "engine", "integrals", "hartree_fock", "dist"]

"""Host-side compatibility modules for the reference's import surface.

`_c_ints` (built by pychem_b200/setup_c_ints.py from csrc/c_ints_shim.c): pychem's 11 legacy
entry points.  `path()` is the directory to put on sys.path so that the reference's
`import _c_ints` (Methods/integrals.py:8) resolves here."""
import os


def path():
    return os.path.dirname(os.path.abspath(__file__))

"""Helpers to drive the REAL reference (oracle/_ref/pychem_py3, built by build_ref.py).

TEST INFRASTRUCTURE ONLY.  Used to validate oracle/eri_oracle.c, to mint tests/golden/ and as
the `reference` arm / cpu_baseline of bench.py.  Raises ImportError when oracle/_ref is absent.
"""
import configparser
import contextlib
import io
import os
import sys
import tempfile
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.path.join(HERE, "_ref", "pychem_py3")

INPUT_TEMPLATE = """[{name}]
Method = "{method}"
Job_Type = "{job_type}"
Basis_Sets = ["{basis}"]
Multiplicity = {mult}
Charge = {charge}
Coords_Units = "ANGSTROM"
Coords = {coords}
Reference = "{reference}"
Max_SCF_Iterations = {maxiter}
MP2_type = "AFTER"
{extra}
"""


def available():
    return os.path.isfile(os.path.join(REF_ROOT, "pychem.py"))


class _NS:
    pass


_MODS = None


def modules():
    """Import the reference's modules (Py3 copy).  Returns a namespace object."""
    global _MODS
    if _MODS is not None:
        return _MODS
    if not available():
        raise ImportError("oracle/_ref not built (run python oracle/build_ref.py where "
                          "/root/reference exists)")
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import pychem
        from Methods import hartree_fock, integrals, mp2, noci, properties
        from Util import structures
        from Data import constants
    ns = _NS()
    ns.pychem, ns.hartree_fock, ns.integrals = pychem, hartree_fock, integrals
    ns.mp2, ns.noci, ns.structures, ns.constants = mp2, noci, structures, constants
    ns.properties = properties
    _MODS = ns
    return ns


def write_input(path, name, coords, basis, method="HF", reference="RHF", mult=1, charge=0,
                maxiter=50, extra="", job_type="Energy"):
    with open(path, "w") as fh:
        fh.write(INPUT_TEMPLATE.format(name=name, method=method, basis=basis, mult=mult,
                                       charge=charge, coords=repr(coords), reference=reference,
                                       maxiter=maxiter, extra=extra, job_type=job_type))


def molecule_from_input(input_file):
    """(molecule, settings) for the first section of an input file, without running anything."""
    ns = modules()
    parser = configparser.ConfigParser()
    parser.read(input_file)
    section = parser.sections()[0]
    with contextlib.redirect_stdout(io.StringIO()):
        molecule, settings = ns.structures.process_input(section, parser)
    return molecule, settings


def build_molecule(coords, basis, **kw):
    modules()
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "x.inp")
        write_input(path, "x", coords, basis, **kw)
        return molecule_from_input(path)


def run(input_file, quiet=True, files=None):
    """pychem.main(input_file) in a scratch directory (it writes <section>.out into cwd).
    `files`: {name: 2-D array} written there first with numpy.savetxt -- the MO files of an
    SCF_Guess = "READ" job (<MO_Read_Basis>_<state>.alpha_MOs / .beta_MOs, hartree_fock.py:35-41)."""
    ns = modules()
    input_file = os.path.abspath(input_file)
    cwd = os.getcwd()
    td = tempfile.mkdtemp()
    os.chdir(td)
    try:
        if files:
            import numpy
            for name, arr in files.items():
                numpy.savetxt(os.path.join(td, name), arr, fmt="%.17e")
        if quiet:
            with contextlib.redirect_stdout(io.StringIO()):
                molecule = ns.pychem.main(input_file)
        else:
            molecule = ns.pychem.main(input_file)
        text = ""
        for name in sorted(os.listdir(td)):
            if name.endswith(".out"):
                with open(os.path.join(td, name)) as fh:
                    text += fh.read()
        molecule.OutText = text          # what the reference wrote to <section>.out
        return molecule
    finally:
        os.chdir(cwd)

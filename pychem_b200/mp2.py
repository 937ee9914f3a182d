"""GPU-backed mirror of the reference's Methods/mp2.py.

    do(settings, molecule, states=None) -> MP2 correlation energy of the last state

Same name, arguments, printed lines and return value as Methods/mp2.py:16-119.  The AO->MO
four-index transform (mp2.py:43-69, O(N^6) Python loops in the reference) runs as four quarter
transforms on the FP64 tensor cores (csrc/pc_mp2.cu, DMMA) and the energy sums (mp2.py:77-94) as
device reductions; the spin-component scaling and the output lines stay here, on the host.

Needs the dense tensor on the device, i.e. `stored` mode of pychem_b200.hartree_fock
(molecule.CoulombIntegrals is what the reference reads at mp2.py:46).

Note on MP2-SCS: the reference is Python-2 code in which `1/3` is 0 and `6/5` is 1
(mp2.py:99-101); run under Python 3 these are true divisions, which is what this mirror does.
"""
import ctypes

import numpy as np

from . import _lib, hartree_fock as _hf


def _printf():
    """The reference's Util.printf when it is importable (the drop-in case), else a stub."""
    try:
        from Util import printf
        return printf
    except Exception:
        class _P:
            @staticmethod
            def delimited_text(outfile, text):
                if outfile is not None:
                    outfile.write(text + "\n")

            @staticmethod
            def text_value(outfile, *pairs):
                if outfile is not None:
                    outfile.write(" ".join(str(x) for x in pairs) + "\n")
        return _P


def mp2_sums(molecule, state, same_spin=True):
    """(Eaa, Eab, Ebb) of one electronic state, unscaled -- mp2.py:43-94."""
    st = _hf._STATE.get(id(molecule))
    if st is None or st.get("G_dev") is None:
        raise _lib.PychemB200Error("MP2 needs the dense tensor on the device: run evaluate_2e_ints in "
                                   "stored mode first (PYCHEM_B200_MODE=stored)")
    db = st["db"]
    N = db.nbf
    Ca = np.ascontiguousarray(state.Alpha.MOs, dtype=np.float64)
    Cb = np.ascontiguousarray(state.Beta.MOs, dtype=np.float64)
    Ea = np.ascontiguousarray(state.Alpha.Energies, dtype=np.float64)
    Eb = np.ascontiguousarray(state.Beta.Energies, dtype=np.float64)
    out = [ctypes.c_double() for _ in range(3)]
    P = lambda a: ctypes.c_void_p(a.ctypes.data)      # noqa: E731
    _lib.check(db.lib.pc_mp2_energy(db.device, N, ctypes.c_void_p(st["G_dev"].data_ptr()), P(Ca), P(Cb),
                                    P(Ea), P(Eb), int(molecule.NAlphaElectrons), int(molecule.NBetaElectrons),
                                    int(bool(same_spin)), *[ctypes.byref(x) for x in out]), mp2=True)
    return out[0].value, out[1].value, out[2].value


def do(settings, molecule, states=None):
    printf = _printf()
    if states is None:
        printf.delimited_text(settings.OutFile, " MP2 calculations for all electronic states ")
        states = molecule.States
    total = 0.0
    for state_index, state in enumerate(states):
        Eaa, Eab, Ebb = mp2_sums(molecule, state, same_spin="P2-SOS" not in settings.Method)
        if "P2-SCS" in settings.Method:
            print("Doing SCS ")
            Eaa *= 1 / 3
            Ebb *= 1 / 3
            Eab *= 6 / 5
        elif "P2-SOS" in settings.Method:
            print("Doing SOS")
            Eaa *= 0
            Ebb *= 0
            Eab *= 1.3
        total = Eaa + Eab + Ebb
        printf.text_value(settings.OutFile, " State: ", state_index, " Total MP2 energy: ",
                          state.TotalEnergy + total)
    printf.delimited_text(settings.OutFile, " End of MP2 calculations ")
    return total


def install(reference_mp2):
    """Rebind `do` inside the reference's Methods.mp2 module (pychem.py:123 calls mp2.do)."""
    saved = reference_mp2.do
    reference_mp2.do = do

    def uninstall():
        reference_mp2.do = saved
    return uninstall

"""ctypes binding of the C ABI in include/pychem_b200.h.

There is NO CPU fallback: if the shared library is missing or a call fails, this raises.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PYCHEM_B200_LIB", os.path.join(HERE, "libpychem_b200.so"))

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int)
c_llp = ctypes.POINTER(ctypes.c_longlong)
c_vp = ctypes.c_void_p

# every symbol include/pychem_b200.h declares: name -> argtypes (restype is int unless noted)
SIGNATURES = {
    "pc_last_error": [],
    "pc_device_count": [c_ip],
    "pc_basis_create": [ctypes.c_int, ctypes.c_int, c_ip, c_ip, c_ip, c_ip, c_dp, c_dp, c_dp,
                        ctypes.POINTER(c_vp)],
    "pc_basis_destroy": [c_vp],
    "pc_basis_set_ints_type": [c_vp, ctypes.c_int, ctypes.c_double],
    "pc_basis_nbf": [c_vp, c_ip],
    "pc_basis_stream": [c_vp, ctypes.POINTER(c_vp)],
    "pc_schwarz": [c_vp, c_dp, c_dp],
    "pc_plan": [c_vp, ctypes.c_double, ctypes.c_int, ctypes.c_int, c_llp, c_llp, c_llp, c_llp],
    "pc_eri_quartets": [c_vp, ctypes.c_int, c_ip, c_llp, c_dp],
    "pc_eri_tensor": [c_vp, c_vp, c_vp],
    "pc_jk_stored": [c_vp] + [c_vp] * 7,
    "pc_jk_direct_accumulate": [c_vp, ctypes.c_int] + [c_vp] * 4,
    "pc_jk_finalize": [c_vp, ctypes.c_int] + [c_vp] * 4,
    "pc_jk_direct_accumulate_auto": [c_vp, c_vp, c_vp, c_vp, c_vp, c_ip],
    "pc_jk_direct": [c_vp, ctypes.c_int] + [c_vp] * 6,
    "pc_jk_direct_auto": [c_vp] + [c_vp] * 6 + [c_ip],
    "pc_jk_classify": [c_vp, c_vp, c_vp, c_vp, c_ip],
    "pc_jk_stored_batch": [c_vp, c_vp, ctypes.c_int, c_vp, c_vp],
    "pc_jk_direct_batch": [c_vp, ctypes.c_int, c_vp, c_vp],
    "pc_jk_direct_batch_accumulate": [c_vp, ctypes.c_int, c_vp, c_vp],
    "pc_jk_finalize_batch": [c_vp, ctypes.c_int, c_vp, c_vp],
    "pc_launch_count": [c_vp, c_llp],
    "pc_set_profiling": [c_vp, ctypes.c_int],
    "pc_plan_items": [c_vp, ctypes.c_int, c_ip, c_ip, c_ip, c_llp, ctypes.POINTER(ctypes.c_float), c_dp],
    "pc_one_electron": [c_vp, ctypes.c_int, c_dp, c_dp, c_vp, c_vp],
    "pc_plan_segments_host": [ctypes.c_int, c_dp, ctypes.c_int, c_ip, ctypes.c_int, c_dp, ctypes.c_int, c_ip,
                              ctypes.c_int, ctypes.c_int, ctypes.c_double, ctypes.c_int, c_ip, c_llp, c_ip, c_llp],
    "pc_fp64_peak": [ctypes.c_int, c_dp],
    "pc_mp2_energy": [ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, ctypes.c_int, ctypes.c_int,
                      ctypes.c_int, c_dp, c_dp, c_dp],
    "pc_mp2_last_error": [],
    "pc_mp2_release": [],
    "pc_dgemm_dmma": [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, c_vp, c_vp, c_vp],
}

_LIB = None


class PychemB200Error(RuntimeError):
    pass


def load():
    """Load libpychem_b200.so (built by pychem_b200/build.py).  Raises if it is not there."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise PychemB200Error(
                "%s is missing: build it with `python -m pychem_b200.build` "
                "(there is no CPU fallback for the GPU path)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the ABI lost a symbol
            fn.argtypes = argtypes
            fn.restype = ctypes.c_char_p if name.endswith("last_error") else ctypes.c_int
        _LIB = lib
    return _LIB


def check(status, mp2=False):
    if status != 0:
        lib = load()
        msg = (lib.pc_mp2_last_error() if mp2 else lib.pc_last_error()).decode()
        raise PychemB200Error(msg or "pychem_b200: unknown error")

#!/usr/bin/env python
"""Build the host-side `_c_ints` module (pychem's 11 legacy entry points, csrc/c_ints_shim.c).

The counterpart of the reference's Methods/setup.py:4-15 -- a setuptools Extension with numpy's
include directory -- for the interpreter at hand (distutils / numpy.distutils, which upstream's
script imports, no longer exist).  Usage, mirroring upstream's install.sh:

    python pychem_b200/setup_c_ints.py build_ext --inplace      # -> pychem_b200/compat/_c_ints*.so

Put pychem_b200/compat on sys.path (or copy the module next to Methods/) and the reference's
`import _c_ints` / `from Methods import _c_ints` resolve to it.
"""
import os
import sys

import numpy
from setuptools import Extension, setup

HERE = os.path.dirname(os.path.abspath(__file__))


def main(argv=None):
    os.chdir(HERE)
    setup(name="pychem_b200_c_ints",
          version="0.2",
          script_args=argv if argv is not None else sys.argv[1:],
          ext_modules=[Extension("compat._c_ints",
                                 sources=[os.path.join("csrc", "c_ints_shim.c")],
                                 include_dirs=[numpy.get_include(), os.path.join(HERE, "csrc")],
                                 extra_compile_args=["-O2", "-std=gnu99"],
                                 libraries=["m"])])


def build_inplace(quiet=True):
    """Called by pychem_b200.build: compile only when the sources are newer than the module."""
    import glob
    import subprocess
    built = glob.glob(os.path.join(HERE, "compat", "_c_ints*.so"))
    srcs = [os.path.join(HERE, "csrc", "c_ints_shim.c"), os.path.join(HERE, "csrc", "pc_boys_table.h"), __file__]
    if built and all(os.path.getmtime(built[0]) >= os.path.getmtime(s) for s in srcs):
        return built[0]
    cmd = [sys.executable, os.path.abspath(__file__), "build_ext", "--inplace", "--build-temp", os.path.join(HERE, "build", "c_ints")]
    subprocess.check_call(cmd, stdout=subprocess.DEVNULL if quiet else None, cwd=HERE)
    return glob.glob(os.path.join(HERE, "compat", "_c_ints*.so"))[0]


if __name__ == "__main__":
    main()

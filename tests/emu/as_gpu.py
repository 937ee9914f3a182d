"""TEST INFRASTRUCTURE ONLY: pytest plugin that points the GPU test files at the host emulation,
to check THEIR logic (shapes, variants, tolerances, fixtures) in a container without a GPU:

    python -m pytest -p tests.emu.as_gpu tests/test_gpu_zzz_f_shells.py -m gpu -q

`pychem_b200.engine.DeviceBasis` becomes tests.emu.emu_engine.EmuDeviceBasis and
`torch.cuda.is_available()` answers True.  Tests that touch CUDA tensors directly still fail (at
that line), and full-size cases are far too slow for the emulation -- select with -k.  A pass here
says nothing about the GPU; it only keeps a mistake in a test from surfacing on the GPU box first.
"""
import torch

from pychem_b200 import engine, integrals
from tests.emu import emu_engine

torch.cuda.is_available = lambda: True
emu_engine.load()
engine.DeviceBasis = emu_engine.EmuDeviceBasis
integrals.DeviceBasis = emu_engine.EmuDeviceBasis

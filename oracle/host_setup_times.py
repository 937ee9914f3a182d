#!/usr/bin/env python
"""Host-side setup cost of the library at (H2O)32 6-31G**, timed WITHOUT a GPU: the same host code
(pc_basis_create, the host part of pc_schwarz, pc_plan) runs in the host-emulation build of the
library sources (tests/emu), where device copies are memmoves and the Schwarz kernels are emulated.
Test/measurement infrastructure (lives under oracle/ with the other developer tools that are not
product code).  Usage: python oracle/host_setup_times.py > profiles/r1e_host_setup_times.json"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pychem_b200 import structures as S  # noqa: E402
from tests.emu import emu_engine as emu  # noqa: E402

emu.load()
out = {"workload": "(H2O)32 6-31G**, N = 768, 73 920 shell pairs", "host_threads": os.cpu_count(),
       "note": "host emulation build (g++ -O1) of pychem_b200/csrc; pc_schwarz includes the EMULATED diagonal-quartet kernels"}
mol = S.Molecule(S.water_cluster(32), "6-31G**")
t = time.time(); db = emu.EmuBasis(mol); out["pc_basis_create_s"] = round(time.time() - t, 4)
t = time.time(); db.schwarz(); out["pc_schwarz_s"] = round(time.time() - t, 4)
t = time.time(); c = db.plan(1e-8, 0, 1); out["pc_plan_s"] = round(time.time() - t, 4)
t = time.time(); db.plan(1e-8, 0, 8); out["pc_plan_rank0_of_8_s"] = round(time.time() - t, 4)
out["surviving_quartets"] = c["all_quartets"]
print(json.dumps(out))

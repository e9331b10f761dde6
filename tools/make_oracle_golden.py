"""Regression fixture for the CPU oracle: a short fp64 trajectory of rodent_0 (5 substeps, CG 4/4).
The reference itself (mujoco-mjx) cannot be run in this image, so this pins the ORACLE against silent edits; the
reference-derived fixtures are tests/golden/notebook_kat.json (tools/extract_notebook_kat.py).
    python tools/make_oracle_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from brax_rodent_run_b200 import mjcf, model_blob  # noqa: E402
from oracle import oracle  # noqa: E402

oracle.build()
m = mjcf.FlatModel.load(os.path.join(ROOT, "brax_rodent_run_b200", "assets", "rodent_0.npz"))
rng = np.random.default_rng(7)
qpos0 = m.qpos0 + rng.uniform(-.01, .01, m.nq)
qpos0[2] = 0.05
qvel0 = rng.uniform(-.01, .01, m.nv)
ctrl = rng.uniform(-1, 1, (5, m.nu))
o = oracle.Oracle(model_blob.pack(m), "f64")
o.set_options(0, 4, 4)
o.init(qpos0, qvel0)
qpos, qvel = [], []
for t in range(5):
    o.set("ctrl", ctrl[t])
    o.step(1)
    qpos.append(o.get("qpos"))
    qvel.append(o.get("qvel"))
np.savez(os.path.join(ROOT, "tests", "golden", "oracle_traj.npz"), qpos0=qpos0, qvel0=qvel0, ctrl=ctrl, qpos=np.array(qpos),
         qvel=np.array(qvel), iterations=4, ls_iterations=4)
print("wrote tests/golden/oracle_traj.npz")

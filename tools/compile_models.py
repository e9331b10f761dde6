"""Compile the reference's MJCF rodent models into FlatModel .npz assets.

/root/reference does not exist on the GPU box, so the compiled tables (derived data, not sources)
are committed under brax_rodent_run_b200/assets/.  Re-run whenever mjcf.py changes:
    python tools/compile_models.py [/root/reference/models]
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from brax_rodent_run_b200 import mjcf  # noqa: E402

src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/models"
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "brax_rodent_run_b200", "assets")
os.makedirs(dst, exist_ok=True)
for name in ("rodent_0", "rodent_new", "rodent_optimized", "rodent_pair"):
    m = mjcf.load_xml(os.path.join(src, name + ".xml"))
    m.save(os.path.join(dst, name + ".npz"))
    print(f"{name}: nq={m.nq} nv={m.nv} nu={m.nu} nbody={m.nbody} ngeom={m.ngeom} ncon={m.ncon} nefc={m.nefc}")

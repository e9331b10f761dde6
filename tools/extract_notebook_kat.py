"""Extract the notebook-derived known answers into tests/golden/notebook_kat.json.

Source: /root/reference/Env_step.ipynb -- the executed cell that prints `state.pipeline_state.qfrc_actuator`
(73 values) and the one that prints `data.q` (74 values) for `reset(PRNGKey(0))`.  These are OUTPUTS of the
reference (MJX) run by its authors; they are the only numeric fixtures the reference holds for this path.
"""
import json
import os
import re

nb = json.load(open("/root/reference/Env_step.ipynb"))


def arrays():
    for cell in nb["cells"]:
        for out in cell.get("outputs", []):
            txt = "".join(out.get("data", {}).get("text/plain", []))
            if txt.startswith("Array(["):
                body = txt[txt.index("[") + 1: txt.rindex("]")]
                vals = [float(x) for x in re.findall(r"[-+]?\d\.\d+e[-+]\d+|[-+]?\d+\.\d*", body)]
                yield vals


found = {len(v): v for v in arrays() if len(v) in (73, 74)}
out = {"source": "Env_step.ipynb (reference notebook outputs)", "qfrc_actuator": found[73], "qpos": found[74]}
dst = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "notebook_kat.json")
json.dump(out, open(dst, "w"), indent=0)
print("wrote", dst, len(found[73]), len(found[74]))

"""FlatModel -> (dir, idata, fdata) blob shared by the CPU oracle and the CUDA library.

The field list is parsed from include/rr_model_fields.h so that C and Python cannot drift."""
from __future__ import annotations

import os
import re
from typing import List, Tuple

import numpy as np

from .mjcf import FlatModel

_HDR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "rr_model_fields.h")


def _parse_header() -> Tuple[List[Tuple[str, str]], List[str], List[str]]:
    src = open(_HDR).read()
    body = re.search(r"#define RR_MODEL_FIELDS\(I, F\)(.*?)\n\n", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = [(k, n) for k, n in re.findall(r"\b([IF])\((\w+)\)", body)]

    def enum_names(prefix):
        blk = re.search(r"enum \{\s*(" + prefix + r".*?)\};", src, re.S).group(1)
        blk = re.sub(r"/\*.*?\*/", "", blk, flags=re.S)
        return [n for n in re.findall(r"\b(" + prefix + r"\w+)", blk) if not n.endswith("_COUNT")]

    return fields, enum_names("RR_OI_"), enum_names("RR_OF_")


FIELDS, OPT_I, OPT_F = _parse_header()
FIELD_INDEX = {n: i for i, (_, n) in enumerate(FIELDS)}


def pack(m: FlatModel):
    """Return (dir int32[2*NF], idata int32[], fdata float64[])."""
    opt_i = np.zeros(len(OPT_I), dtype=np.int32)
    for k, name in enumerate(OPT_I):
        key = name[len("RR_OI_"):].lower()
        opt_i[k] = int(getattr(m, "nM" if key == "nm" else key))
    g = np.asarray(m.gravity, dtype=np.float64)
    opt_f_vals = dict(timestep=m.timestep, gravity_x=g[0], gravity_y=g[1], gravity_z=g[2], tolerance=m.tolerance,
                      ls_tolerance=m.ls_tolerance, impratio=m.impratio, meaninertia=m.meaninertia)
    opt_f = np.array([opt_f_vals[name[len("RR_OF_"):].lower()] for name in OPT_F], dtype=np.float64)
    dir_ = np.zeros(2 * len(FIELDS), dtype=np.int32)
    ichunks, fchunks = [], []
    ioff = foff = 0
    for k, (kind, name) in enumerate(FIELDS):
        if name == "opt_i":
            a = opt_i
        elif name == "opt_f":
            a = opt_f
        else:
            a = m.arrays[name]
        if kind == "I":
            a = np.ascontiguousarray(a, dtype=np.int32).ravel()
            dir_[2 * k], dir_[2 * k + 1] = ioff, a.size
            ichunks.append(a)
            ioff += a.size
        else:
            a = np.ascontiguousarray(a, dtype=np.float64).ravel()
            dir_[2 * k], dir_[2 * k + 1] = foff, a.size
            fchunks.append(a)
            foff += a.size
    idata = np.concatenate(ichunks) if ichunks else np.zeros(0, np.int32)
    fdata = np.concatenate(fchunks) if fchunks else np.zeros(0, np.float64)
    return dir_, idata, fdata

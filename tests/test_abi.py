"""The C-ABI libraries load and export every symbol include/rr_b200.h declares; error codes map to the exceptions the
reference raises (no compute calls on the CUDA library here: this file runs without a GPU)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_asset, synthetic_track


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "rr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rr_[a-z_0-9]+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ("rr_model_create", "rr_env_create", "rr_env_init", "rr_env_step", "rr_env_step_host", "rr_gae", "rr_last_error"):
        assert s in syms


def test_cuda_library_exports_all_symbols():
    from brax_rodent_run_b200 import build
    path = build.build_cuda()  # cross-compiles for sm_100a; no GPU needed
    L = ctypes.CDLL(path)
    for s in declared_symbols():
        assert hasattr(L, s), s
    assert L.rr_prof_count() > 0


def test_emulator_exports_all_symbols(emu_lib):
    L = ctypes.CDLL(emu_lib)
    for s in declared_symbols():
        assert hasattr(L, s), s


def test_product_path_refuses_cpu():
    """No CPU fallback: the product constructor only accepts CUDA devices."""
    from brax_rodent_run_b200.env import Rodent
    with pytest.raises(RuntimeError):
        Rodent(synthetic_track(), device="cpu", model="rodent_0")


def test_missing_library_fails_loudly(tmp_path):
    from brax_rodent_run_b200 import _lib
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.load(str(tmp_path / "nope.so"))


def test_error_mapping(emu_lib):
    """Constructor errors mirror the reference: KeyError on a bad solver name (Rodent_Env_Brax.py:42-45), ValueError on a
    bad model, NotImplementedError on unsupported features."""
    from brax_rodent_run_b200 import _lib
    from brax_rodent_run_b200.env import Rodent
    with pytest.raises(KeyError):
        Rodent(synthetic_track(), device="cpu", model="rodent_0", solver="sor", _lib_path=emu_lib)
    with pytest.raises(NotImplementedError):
        Rodent(synthetic_track(), device="cpu", model="rodent_0", solver="newton", _lib_path=emu_lib)
    with pytest.raises(ValueError):
        Rodent(synthetic_track(), device="cpu", model="no_such_model", _lib_path=emu_lib)
    with pytest.raises(ValueError):
        Rodent(synthetic_track(), device="cpu", xml_path="/nonexistent.xml", _lib_path=emu_lib)
    L = _lib.load(emu_lib)
    h = ctypes.c_void_p()
    bad = np.zeros(4, np.int32)
    rc = L.rr_model_create(bad.ctypes.data_as(_lib.c_i), 4, bad.ctypes.data_as(_lib.c_i), 4,
                           np.zeros(4).ctypes.data_as(_lib.c_d), 4, ctypes.byref(h))
    assert rc == _lib.RR_EINVAL and b"directory" in L.rr_last_error()
    with pytest.raises(ValueError):
        _lib.check(L, rc)


def test_step_requires_task_and_buffers(emu_lib):
    from brax_rodent_run_b200 import _lib
    from brax_rodent_run_b200.env import Rodent
    env = Rodent(synthetic_track(), device="cpu", model="rodent_0", _lib_path=emu_lib)
    buf = _lib.RRBuffers()
    rc = env._L.rr_env_step(env._env, ctypes.byref(buf), None, 10, None)
    assert rc == _lib.RR_EINVAL
    rc = env._L.rr_env_init(env._env, ctypes.byref(buf), None)
    assert rc == _lib.RR_EINVAL

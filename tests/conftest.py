"""Shared fixtures.

Markers:  gpu  -- needs a CUDA device (parity tests proper; call the product library through the C ABI).
Everything else runs on CPU: oracle vs golden vectors, loader / host logic, ABI symbol checks, and the same
parity cases against tests/emu (the step-kernel text compiled for the host with lanes as fibers; test-only).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

EMU_DIR = os.path.join(ROOT, "tests", "emu")
EMU_LIB = os.path.join(EMU_DIR, "librr_emu.so")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device")


def build_emu(force=False):
    from brax_rodent_run_b200 import build as rb
    deps = rb.sources() + [os.path.join(EMU_DIR, "rr_emu.cpp")]
    if not force and rb.up_to_date(EMU_LIB, deps):
        return EMU_LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", EMU_LIB,
                           os.path.join(EMU_DIR, "rr_emu.cpp")])
    return EMU_LIB


@pytest.fixture(scope="session")
def emu_lib():
    return build_emu()


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import oracle
    oracle.build()
    return oracle


def load_asset(name):
    from brax_rodent_run_b200 import mjcf
    return mjcf.FlatModel.load(os.path.join(ROOT, "brax_rodent_run_b200", "assets", name + ".npz"))


@pytest.fixture(scope="session")
def rodent0():
    return load_asset("rodent_0")


def synthetic_track(n=250):
    """straight line x = 0.002 t, y = 0, z = 0.055."""
    return np.stack([0.002 * np.arange(n), np.zeros(n), np.full(n, 0.055)], 1).astype(np.float32)


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def backend_params():
    """(device, lib_path) pairs: the emulator on CPU always, the product library on a GPU when marked."""
    return [pytest.param("emu", id="emu"), pytest.param("cuda", id="cuda", marks=pytest.mark.gpu)]


@pytest.fixture
def make_env(request, emu_lib):
    """Factory: make_env(backend, track, **kw) -> Rodent."""
    from brax_rodent_run_b200.env import Rodent

    def factory(backend, track, **kw):
        if backend == "emu":
            return Rodent(track, device="cpu", _lib_path=emu_lib, **kw)
        if not has_cuda():
            pytest.skip("no CUDA device")
        return Rodent(track, device="cuda:0", **kw)

    return factory

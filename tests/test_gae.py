"""ppo.losses.compute_gae (brax) restated in numpy vs the rr_gae kernel through the C ABI."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import backend_params, has_cuda


def gae_numpy(truncation, termination, rewards, values, bootstrap, lambda_, discount):
    mask = 1 - truncation
    v_tp1 = np.concatenate([values[1:], bootstrap[None]], 0)
    deltas = (rewards + discount * (1 - termination) * v_tp1 - values) * mask
    acc = np.zeros_like(bootstrap)
    out = np.zeros_like(values)
    for t in range(values.shape[0] - 1, -1, -1):
        acc = deltas[t] + discount * (1 - termination[t]) * mask[t] * lambda_ * acc
        out[t] = acc
    vs = out + values
    vs_tp1 = np.concatenate([vs[1:], bootstrap[None]], 0)
    adv = (rewards + discount * (1 - termination) * vs_tp1 - values) * mask
    return vs, adv


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("T,B", [(10, 257), (1, 5), (32, 64)])
def test_gae(backend, T, B, emu_lib):
    from brax_rodent_run_b200 import _lib
    if backend == "cuda" and not has_cuda():
        pytest.skip("no CUDA device")
    L = _lib.load(emu_lib if backend == "emu" else None)
    dev = "cpu" if backend == "emu" else "cuda:0"
    rng = np.random.default_rng(T * 1000 + B)
    rewards = rng.normal(size=(T, B)).astype(np.float32)
    values = rng.normal(size=(T, B)).astype(np.float32)
    boot = rng.normal(size=(B,)).astype(np.float32)
    done = (rng.uniform(size=(T, B)) < 0.1).astype(np.float32)
    trunc = ((rng.uniform(size=(T, B)) < 0.05) * done).astype(np.float32)
    term = done * (1 - trunc)
    t = {k: torch.tensor(v, device=dev) for k, v in dict(r=rewards, v=values, b=boot, te=term, tr=trunc).items()}
    vs, adv = torch.empty((T, B), device=dev), torch.empty((T, B), device=dev)
    p = lambda x: ctypes.c_void_p(x.data_ptr())
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream) if backend == "cuda" else None
    _lib.check(L, L.rr_gae(p(t["r"]), p(t["v"]), p(t["b"]), p(t["te"]), p(t["tr"]), T, B, 0.97, 0.95, p(vs), p(adv), stream))
    want_vs, want_adv = gae_numpy(trunc.astype(np.float64), term.astype(np.float64), rewards.astype(np.float64),
                                  values.astype(np.float64), boot.astype(np.float64), 0.95, 0.97)
    assert np.abs(vs.cpu().numpy() - want_vs).max() < 1e-5
    assert np.abs(adv.cpu().numpy() - want_adv).max() < 1e-5

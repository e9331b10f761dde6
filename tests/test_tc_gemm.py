"""rr_tc_plan / rr_tc_launch (include/rr_b200.h): the learner's grouped GEMM with fused epilogues against plain torch.

On the CPU the emulator backend runs the contract's plain loops (the problem-list logic, majors, epilogues, the bias-gradient
column); on a GPU the same cases go through the tcgen05 kernel (TF32 products, fp32 accumulation: tolerance 2e-3 relative to
the magnitude of the result, as cuBLAS's TF32 path)."""
import pytest
import torch

from brax_rodent_run_b200 import _lib, tc_gemm
from conftest import backend_params, has_cuda


def _setup(backend, emu_lib):
    if backend == "emu":
        return _lib.load(emu_lib), torch.device("cpu"), 1e-5
    if not has_cuda():
        pytest.skip("no CUDA device")
    return _lib.load(), torch.device("cuda:0"), 2e-3


def _silu_grad(z):
    s = torch.sigmoid(z)
    return s * (1 + z * (1 - s))


def _close(got, want, tol, scale=None):
    """`scale`: magnitude of the summed products (TF32 rounds every product; a sum that cancels keeps their absolute error)."""
    scale = max(float(want.abs().max()), 1e-6, scale or 0.0)
    err = float((got - want).abs().max()) / scale
    assert err <= tol, err


# (m, n, k): tile-aligned, ragged in every direction, the value head (n = 1), the policy head (n = 60), k not a multiple of 4
SHAPES = [(128, 64, 32), (256, 256, 1264), (200, 60, 32), (130, 1, 256), (77, 33, 50), (512, 256, 256), (5, 7, 3)]


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("m,n,k", SHAPES)
def test_forward_linear_and_silu(backend, emu_lib, m, n, k):
    L, dev, tol = _setup(backend, emu_lib)
    g = torch.Generator().manual_seed(m * 1000 + n * 10 + k)
    x = torch.randn(m, k, generator=g).to(dev)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(dev)
    b = torch.randn(n, generator=g).to(dev)
    y0, y1, z1 = (torch.full((m, n), float("nan"), device=dev) for _ in range(3))
    grp = tc_gemm.TcGroup(L, [tc_gemm.problem(x, w, y0, bias=b),
                              tc_gemm.problem(x, w, y1, bias=b, epi=tc_gemm.EPI_SILU, aux_out=z1)], dev)
    grp.launch()
    want = x.double() @ w.double().T + b.double()
    _close(y0.double(), want, tol)
    _close(z1.double(), want, tol)
    _close(y1.double(), torch.nn.functional.silu(want), tol)


@pytest.mark.parametrize("backend", backend_params())
@pytest.mark.parametrize("m,n,k", SHAPES)
def test_dgrad_and_wgrad(backend, emu_lib, m, n, k):
    """dX = (dY W) * silu'(Z) with W read MN-major; dW = dY' X and db = dY' 1 with both operands MN-major."""
    L, dev, tol = _setup(backend, emu_lib)
    g = torch.Generator().manual_seed(m * 1000 + n * 10 + k + 1)
    x = torch.randn(m, k, generator=g).to(dev)
    z = torch.randn(m, k, generator=g).to(dev)
    w = (torch.randn(n, k, generator=g) / n ** 0.5).to(dev)
    dy = torch.randn(m, n, generator=g).to(dev)
    dx = torch.full((m, k), float("nan"), device=dev)
    dw = torch.full((n, k), float("nan"), device=dev)
    db = torch.full((n,), float("nan"), device=dev)
    # the same weight gradient with the ones stored as a column of the activation buffer (pitch k + 4)
    xbuf = torch.zeros(m, k + 4, device=dev)
    xbuf[:, :k], xbuf[:, k] = x, 1.0
    dw2 = torch.full((n, k), float("nan"), device=dev)
    db2 = torch.full((n,), float("nan"), device=dev)
    grp = tc_gemm.TcGroup(L, [tc_gemm.problem(dy, w, dx, b_t=True, epi=tc_gemm.EPI_DSILU, aux_in=z),
                              tc_gemm.problem(dy, x, dw, a_t=True, b_t=True, ones_out=db),
                              tc_gemm.problem(dy, xbuf[:, :k + 1], dw2, a_t=True, b_t=True, ones_out=db2, ones_stored=True)], dev)
    grp.launch()
    _close(dw2.double(), dy.double().T @ x.double(), tol)
    _close(db2.double(), dy.double().sum(0), tol, scale=float(dy.abs().sum(0).max()) / 10)
    _close(dx.double(), (dy.double() @ w.double()) * _silu_grad(z.double()), tol)
    _close(dw.double(), dy.double().T @ x.double(), tol)
    _close(db.double(), dy.double().sum(0), tol, scale=float(dy.abs().sum(0).max()) / 10)


@pytest.mark.parametrize("backend", backend_params())
def test_strided_views_and_errors(backend, emu_lib):
    """Operands may be column slices of wider arrays (leading dimension > row length); bad problems raise ValueError."""
    L, dev, tol = _setup(backend, emu_lib)
    g = torch.Generator().manual_seed(5)
    big = torch.randn(96, 300, generator=g).to(dev)
    x, w = big[:, 4:132], torch.randn(48, 128, generator=g).to(dev)
    out = torch.zeros(96, 100, device=dev)
    grp = tc_gemm.TcGroup(L, [tc_gemm.problem(x, w, out[:, 52:100])], dev)
    grp.launch()
    _close(out[:, 52:100].double(), x.double() @ w.double().T, tol)
    assert float(out[:, :52].abs().max()) == 0.0
    with pytest.raises(ValueError):
        p = tc_gemm.problem(x, w, out[:, 52:100])
        p["epi"] = 2  # no aux_in
        tc_gemm.TcGroup(L, [p], dev)


@pytest.mark.parametrize("backend", backend_params())
def test_flat_adam_matches_torch_adam(backend, emu_lib):
    """rr_adam_step: torch.optim.Adam's arithmetic (the reference's optax.adam) on a flat buffer, step counter on the device."""
    from brax_rodent_run_b200.tc_learner import FlatAdam
    L, dev, _ = _setup(backend, emu_lib)
    g0 = torch.Generator().manual_seed(0)
    p, g = torch.randn(70001, generator=g0).to(dev), torch.randn(70001, generator=g0).to(dev)
    ref = torch.nn.Parameter(p.clone())
    opt = torch.optim.Adam([ref], lr=3e-4, eps=1e-8)
    fa = FlatAdam(L, p, g, 3e-4)
    for _ in range(6):
        ref.grad = g.clone()
        opt.step()
        fa.step()
        g.mul_(0.7).add_(0.1)
    assert float(fa.state[0]["step"]) == 6.0
    assert float((p - ref.data).abs().max()) < 2e-6
    # the gradient given as partial sums (the split weight-gradient launch's workspace): summed, stored and stepped in one pass
    parts = torch.randn(4, 70001, generator=g0).to(dev)
    pc, gc = p.clone(), torch.zeros_like(g)
    fc = FlatAdam(L, pc, gc, 3e-4)
    fc.load_state_dict(fa.state_dict())
    fc.partials = parts
    fc.step()
    want_g = ((parts[0] + parts[1]) + parts[2]) + parts[3]
    assert torch.equal(gc, want_g) and fc.partials is None
    g.copy_(want_g)
    fa.step()
    assert torch.equal(pc, p)
    sd = fa.state_dict()
    fb = FlatAdam(L, p.clone(), g, 3e-4)
    fb.load_state_dict(sd)
    assert torch.equal(fb.state[0]["exp_avg"], fa.state[0]["exp_avg"]) and float(fb.state[0]["step"]) == 7.0


@pytest.mark.parametrize("backend", backend_params())
def test_gather_rows_matches_index_select(backend, emu_lib):
    """rr_gather_rows: the learner's minibatch gather (seven tensors, one launch) against torch indexing."""
    import ctypes
    L, dev, _ = _setup(backend, emu_lib)
    g0 = torch.Generator().manual_seed(1)
    T, N, rows = 3, 50, 16
    srcs = [torch.randn(T, N, 1264, generator=g0), torch.randn(T, N, 30, generator=g0), torch.randn(T, N, generator=g0),
            torch.randn(N, 1264, generator=g0), torch.randn(T, N, 7, generator=g0)]
    srcs = [s.to(dev) for s in srcs]
    idx = torch.randperm(N, generator=g0)[:rows].to(dev)
    dsts = [torch.zeros((s.shape[0], rows) + tuple(s.shape[2:]), device=dev) if i != 3 else torch.zeros(rows, 1264, device=dev)
            for i, s in enumerate(srcs)]
    items = (_lib.RRGatherItem * len(srcs))()
    for i, (it, s, d) in enumerate(zip(items, srcs, dsts)):
        lead = 1 if i == 3 else 2
        it.src, it.dst = s.data_ptr(), d.data_ptr()
        it.outer, it.src_rows = (1 if lead == 1 else s.shape[0]), s.shape[lead - 1]
        it.inner = int(torch.tensor(s.shape[lead:]).prod()) if s.dim() > lead else 1
    stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream) if dev.type == "cuda" else None
    _lib.check(L, L.rr_gather_rows(items, len(srcs), ctypes.c_void_p(idx.data_ptr()), rows, stream))
    for i, (s, d) in enumerate(zip(srcs, dsts)):
        assert torch.equal(d, s[idx] if i == 3 else s[:, idx]), i
    with pytest.raises(ValueError):
        _lib.check(L, L.rr_gather_rows(items, 9, ctypes.c_void_p(idx.data_ptr()), rows, stream))


@pytest.mark.gpu
@pytest.mark.parametrize("m,n,k", [(256, 256, 1264), (200, 60, 96), (130, 1, 256)])
def test_cp_async_fallback_path(emu_lib, monkeypatch, m, n, k):
    """Without tensor maps (RR_TC_NO_TMA: what happens when the driver entry point is unavailable, and what unaligned operands
    always take) every operand is staged by cp.async into the same swizzled layouts: interior, ragged-row and generic loaders."""
    L, dev, tol = _setup("cuda", emu_lib)
    g = torch.Generator().manual_seed(m + n + k)
    x = torch.randn(m, k, generator=g).to(dev)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).to(dev)
    dy = torch.randn(m, n, generator=g).to(dev)
    b = torch.randn(n, generator=g).to(dev)
    y, dx = torch.empty(m, n, device=dev), torch.empty(m, k, device=dev)
    dw, db = torch.empty(n, k, device=dev), torch.empty(n, device=dev)
    probs = lambda: [tc_gemm.problem(x, w, y, bias=b), tc_gemm.problem(dy, w, dx, b_t=True),
                     tc_gemm.problem(dy, x, dw, a_t=True, b_t=True, ones_out=db)]
    with_tma = tc_gemm.TcGroup(L, probs(), dev)
    monkeypatch.setenv("RR_TC_NO_TMA", "1")
    grp = tc_gemm.TcGroup(L, probs(), dev)
    assert all(a == 0 and bb == 0 for a, bb in grp.tma)
    assert any(a or bb for a, bb in with_tma.tma) or k % 4 != 0
    grp.launch()
    _close(y.double(), x.double() @ w.double().T + b.double(), tol)
    _close(dx.double(), dy.double() @ w.double(), tol)
    _close(dw.double(), dy.double().T @ x.double(), tol)
    _close(db.double(), dy.double().sum(0), tol, scale=float(dy.abs().sum(0).max()) / 10)

"""Grouped TF32 tensor-core GEMMs with fused epilogues (rr_tc_plan / rr_tc_launch in include/rr_b200.h): the contractions of
the PPO learner's two MLPs (brax ppo.train under jax.grad, brax_rodent_run_ppo.py:97-114, 200) as hand-written tcgen05
kernels.  A `TcGroup` is a planned list of problems with a device-resident copy of the descriptors, so that `launch()` is
one kernel launch and capturable in a CUDA graph.
"""
from __future__ import annotations

import ctypes
from typing import Dict, List, Optional

import torch

from . import _lib

EPI_LINEAR, EPI_SILU, EPI_DSILU = 0, 1, 2


def problem(a: torch.Tensor, b: torch.Tensor, d: torch.Tensor, *, a_t: bool = False, b_t: bool = False,
            bias: Optional[torch.Tensor] = None, epi: int = EPI_LINEAR, aux_in: Optional[torch.Tensor] = None,
            aux_out: Optional[torch.Tensor] = None, ones_out: Optional[torch.Tensor] = None, ones_stored: bool = False) -> Dict:
    """D = epi(A B' + bias) with logical A [m, k], B [n, k].  `a` is A stored [m, k] (or A' stored [k, m] with a_t), `b` is B
    stored [n, k] (or B' stored [k, n] with b_t); all 2-D fp32 with unit stride along the last axis.  `ones_out` [m]: also
    return the sum of A over k (needs b_t) -- or, with `ones_stored`, B's last row is a stored row of ones: D has n - 1 columns and
    the product's last column goes to ones_out."""
    for t in (a, b, d, bias, aux_in, aux_out, ones_out):
        if t is not None:
            assert t.dtype == torch.float32 and (t.dim() == 1 or t.stride(-1) == 1), "fp32, unit stride along the last axis"
    m, k = (a.shape[1], a.shape[0]) if a_t else (a.shape[0], a.shape[1])
    n, kb = (b.shape[1], b.shape[0]) if b_t else (b.shape[0], b.shape[1])
    n_d = n - 1 if ones_stored else n
    assert k == kb and tuple(d.shape) == (m, n_d), (a.shape, b.shape, d.shape, a_t, b_t)
    aux = aux_in if aux_in is not None else aux_out
    if aux is not None:
        assert tuple(aux.shape) == (m, n_d)
    if bias is not None:
        assert bias.numel() == n_d
    if ones_out is not None:
        assert (b_t or ones_stored) and ones_out.numel() == m
    return dict(a=a, b=b, d=d, bias=bias, aux_in=aux_in, aux_out=aux_out, ones_out=ones_out, m=m, n=n, k=k, lda=a.stride(0),
                ldb=b.stride(0), ldd=d.stride(0), ldaux=aux.stride(0) if aux is not None else 0, a_mn=int(a_t), b_mn=int(b_t),
                epi=epi, b_ones=(2 if ones_stored else 1) if ones_out is not None else 0)


class TcGroup:
    def __init__(self, L, problems: List[Dict], device, prof: Optional[torch.Tensor] = None):
        """`prof`: int64 [>= tiles, 16] device tensor for the kernel's per-CTA cycle counters (tools/tc_learner_timing.py)."""
        self.L, self.n, self.device = L, len(problems), torch.device(device)
        self._keep = problems  # the tensors whose addresses are baked into the descriptors
        arr = (_lib.RRTcProblem * self.n)()
        for q, p in zip(arr, problems):
            for name in ("a", "b", "d", "bias", "aux_in", "aux_out", "ones_out"):
                t = p[name]
                if t is not None:
                    assert t.device.type == self.device.type
                setattr(q, name, t.data_ptr() if t is not None else None)
            for name in ("m", "n", "k", "lda", "ldb", "ldd", "ldaux", "a_mn", "b_mn", "epi", "b_ones"):
                setattr(q, name, int(p[name]))
            if prof is not None:
                addr = prof.data_ptr()
                q.reserved[0], q.reserved[1] = ctypes.c_int32(addr & 0xFFFFFFFF).value, ctypes.c_int32(addr >> 32).value
        tiles, smem = ctypes.c_int32(), ctypes.c_int32()
        rec_bytes = L.rr_tc_record_bytes()
        self._records = (ctypes.c_uint8 * (rec_bytes * self.n + 128))()
        base = (ctypes.addressof(self._records) + 127) // 128 * 128   # records hold 128-byte aligned tensor maps
        _lib.check(L, L.rr_tc_plan(arr, self.n, ctypes.byref(tiles), ctypes.byref(smem), ctypes.c_void_p(base)))
        self.tiles, self.smem = tiles.value, smem.value
        self.tma = [(q.reserved[2] & 1, (q.reserved[2] >> 1) & 1) for q in arr]
        self._host = arr
        if self.device.type == "cuda":
            raw = torch.frombuffer(bytearray(ctypes.string_at(base, rec_bytes * self.n)), dtype=torch.uint8)
            self._dev = raw.to(self.device)
            assert self._dev.data_ptr() % 128 == 0
            self._ptr = self._dev.data_ptr()
        else:  # emulator backend: "device" pointers are host pointers
            self._ptr = base

    def launch(self) -> None:
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream) if self.device.type == "cuda" else None
        _lib.check(self.L, self.L.rr_tc_launch(ctypes.c_void_p(self._ptr), self.n, self.tiles, self.smem, stream))

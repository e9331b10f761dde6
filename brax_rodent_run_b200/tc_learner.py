"""Forward and hand-written backward of the PPO learner's policy / value MLPs on the tcgen05 GEMM kernel (tc_gemm.py).

Replaces, for one minibatch, what autograd + cuBLAS did in ~95 launches (brax: `jax.grad(compute_ppo_loss)` inside ppo.train,
brax_rodent_run_ppo.py:97-114, 200) by 12 grouped launches:
  forward   6: layer l of the policy net, the value net and the bootstrap value net (next observation) in ONE launch each;
               bias + SiLU fused, the pre-activations kept for the backward pass
  loss      2: rr_ppo_loss (d loss / d logits, d loss / d baseline)
  dgrad     5: dZ_{l-1} = (dZ_l W_l) * silu'(Z_{l-1}) for both nets per launch (W read MN-major: no transposed copy)
  wgrad     1: all eleven dW_l = dZ_l' X_l in one launch (each split four ways over the rows; the partial sums are added in a fixed
               order by one torch.sum); the activation buffers carry a column of ones, so [dW_l | db_l] = dZ_l' [X_l | 1] is one
               product; the parameters' .grad tensors are views of one flat buffer, which is also the NCCL all-reduce bucket
The parameters stay ordinary nn.Linear weights (acting, export and the optimizer are unchanged).
"""
from __future__ import annotations

from typing import List

import torch
from torch import nn

from . import tc_gemm
from .tc_gemm import EPI_DSILU, EPI_LINEAR, EPI_SILU, TcGroup, problem


def flatten_parameters(params, device):
    """Re-home the parameters in ONE flat fp32 buffer (16-byte aligned pieces; .data and .grad become views): the gradient
    buffer is the weight-gradient launch's output and the NCCL bucket, the parameter buffer is what Adam steps over."""
    sizes = [((p.numel() + 3) // 4) * 4 for p in params]
    flat_p, flat_g = torch.zeros(sum(sizes), device=device), torch.zeros(sum(sizes), device=device)
    offsets, off = {}, 0
    for p, sz in zip(params, sizes):
        view = flat_p[off:off + p.numel()].view_as(p)
        view.copy_(p.data)
        p.data = view
        p.grad = flat_g[off:off + p.numel()].view_as(p)
        offsets[id(p)] = off
        off += sz
    return flat_p, flat_g, offsets


class FlatAdam:
    """torch.optim.Adam's arithmetic (optax.adam in brax's ppo.train) over the flat parameter buffer: one elementwise kernel
    (rr_adam_step) instead of the multi-tensor kernel over 22 small tensors (48 us per step on the B200), capturable."""

    def __init__(self, L, flat_param, flat_grad, lr, betas=(0.9, 0.999), eps=1e-8):
        self.L, self.p, self.g, self.lr, self.betas, self.eps = L, flat_param, flat_grad, lr, betas, eps
        self.partials = None   # [nsplit, n]: when set, the next step() first forms grad = sum of the partials (TcLearner.backward)
        self.state = {0: dict(step=torch.zeros(1, device=flat_param.device), exp_avg=torch.zeros_like(flat_param),
                              exp_avg_sq=torch.zeros_like(flat_param))}

    def zero_grad(self, set_to_none: bool = False):
        pass  # the weight-gradient launch overwrites the whole buffer

    def step(self):
        import ctypes
        from . import _lib
        st, c = self.state[0], lambda t: ctypes.c_void_p(t.data_ptr())
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.p.device).cuda_stream) if self.p.is_cuda else None
        if self.partials is not None:
            _lib.check(self.L, self.L.rr_adam_step_sum(c(self.p), c(self.g), c(self.partials), self.partials.shape[0], c(st["exp_avg"]),
                                                       c(st["exp_avg_sq"]), c(st["step"]), self.p.numel(), self.lr, self.betas[0],
                                                       self.betas[1], self.eps, stream))
            self.partials = None
            return
        _lib.check(self.L, self.L.rr_adam_step(c(self.p), c(self.g), c(st["exp_avg"]), c(st["exp_avg_sq"]), c(st["step"]),
                                               self.p.numel(), self.lr, self.betas[0], self.betas[1], self.eps, stream))

    def state_dict(self):
        return {k: v.clone() for k, v in self.state[0].items()}

    def load_state_dict(self, d):
        for k, v in d.items():
            self.state[0][k].copy_(v)


def _linears(net: nn.Sequential) -> List[nn.Linear]:
    return [m for m in net if isinstance(m, nn.Linear)]


class TcLearner:
    def __init__(self, L, policy: nn.Sequential, value: nn.Sequential, rows: int, boot_rows: int, device, flat_grad=None,
                 grad_offsets=None):
        self.L, self.device = L, torch.device(device)
        self.lp, self.lv = _linears(policy), _linears(value)
        M, Mb, dev = rows, boot_rows, self.device
        k0 = self.lp[0].in_features
        assert self.lv[0].in_features == k0
        new = lambda r, c: torch.zeros(r, c, device=dev)

        def with_ones(r, c):
            """[r, c] activations stored with pitch c + 4 and a column of ones at index c: the weight-gradient launch reads
            [activations | 1] as ONE operand (by TMA), which makes the bias gradient the product's last column."""
            buf = new(r, c + 4)
            buf[:, c] = 1.0
            return buf[:, :c], buf[:, :c + 1]

        self.x, self.x1 = with_ones(M, k0)      # x: [M, k0] view (what the forward reads / the gather fills), x1: [x | 1]
        self.xb = new(Mb, k0)
        # activations (h), pre-activations (z) and their gradients (dz) of the hidden layers; the heads' outputs
        hp = [with_ones(M, l.out_features) for l in self.lp[:-1]]
        hv = [with_ones(M, l.out_features) for l in self.lv[:-1]]
        self.hp, self.hp1 = [a for a, _ in hp], [b for _, b in hp]
        self.hv, self.hv1 = [a for a, _ in hv], [b for _, b in hv]
        self.zp = [new(M, l.out_features) for l in self.lp[:-1]]
        self.dzp = [torch.zeros_like(z) for z in self.zp]
        self.zv = [new(M, l.out_features) for l in self.lv[:-1]]
        self.dzv = [torch.zeros_like(z) for z in self.zv]
        self.hb = [new(Mb, l.out_features) for l in self.lv[:-1]]
        self.logits, self.grad_logits = new(M, self.lp[-1].out_features), new(M, self.lp[-1].out_features)
        self.baseline, self.grad_baseline = new(M, 1), new(M, 1)
        self.bootstrap = new(Mb, 1)
        # static gradient tensors: views of ONE flat buffer (16-byte aligned pieces), which is also the all-reduce bucket
        if flat_grad is None:
            _, flat_grad, grad_offsets = flatten_parameters([p for lin in self.lp + self.lv for p in (lin.weight, lin.bias)], dev)
        self.flat_grad, self._grad_off = flat_grad, grad_offsets
        # the weight gradients reduce over all M rows: split them over `splits` CTAs per output tile (partial sums in a
        # workspace with the flat buffer's layout, summed in a fixed order afterwards -> deterministic)
        self.splits = 4 if (M % 4 == 0 and M // 4 >= 256) else 1
        self.ws = torch.zeros(self.splits, self.flat_grad.numel(), device=dev) if self.splits > 1 else None

        def fwd(lins, l, inp, h, z, out):
            lin, last = lins[l], l == len(lins) - 1
            if last:
                return problem(inp, lin.weight.data, out, bias=lin.bias.data, epi=EPI_LINEAR)
            return problem(inp, lin.weight.data, h[l], bias=lin.bias.data, epi=EPI_SILU, aux_out=z[l] if z is not None else None)

        self.fwd_groups = []
        for l in range(max(len(self.lp), len(self.lv))):
            probs = []
            if l < len(self.lp):
                probs.append(fwd(self.lp, l, self.x if l == 0 else self.hp[l - 1], self.hp, self.zp, self.logits))
            if l < len(self.lv):
                probs.append(fwd(self.lv, l, self.x if l == 0 else self.hv[l - 1], self.hv, self.zv, self.baseline))
                probs.append(fwd(self.lv, l, self.xb if l == 0 else self.hb[l - 1], self.hb, None, self.bootstrap))
            self.fwd_groups.append(TcGroup(L, probs, dev))

        # dgrad: walk both nets from the heads down; step s handles layer (n - 1 - s) of each net that still has a hidden input
        def dgrad_chain(lins, dz, z, g_head):
            chain = []
            for l in range(len(lins) - 1, 0, -1):  # dZ_{l-1} from dZ_l
                dz_l = g_head if l == len(lins) - 1 else dz[l]
                chain.append(problem(dz_l, lins[l].weight.data, dz[l - 1], b_t=True, epi=EPI_DSILU, aux_in=z[l - 1]))
            return chain

        cp, cv = dgrad_chain(self.lp, self.dzp, self.zp, self.grad_logits), dgrad_chain(self.lv, self.dzv, self.zv, self.grad_baseline)
        self.dgrad_groups = []
        # align the chains at their ends so that both finish in the last launch (the value net is one layer deeper)
        steps = max(len(cp), len(cv))
        for s in range(steps):
            probs = []
            ip, iv = s - (steps - len(cp)), s - (steps - len(cv))
            if ip >= 0:
                probs.append(cp[ip])
            if iv >= 0:
                probs.append(cv[iv])
            self.dgrad_groups.append(TcGroup(L, probs, dev))

        self.ones_row = torch.ones(1, M, device=dev)

        def wgrad(lins, dz, h, h1, g_head, l, sp):
            dz_l = g_head if l == len(lins) - 1 else dz[l]
            inp, inp1 = (self.x, self.x1) if l == 0 else (h[l - 1], h1[l - 1])
            w, b = lins[l].weight, lins[l].bias
            rows = slice(sp * (M // self.splits), (sp + 1) * (M // self.splits))
            ow, ob = self._grad_off[id(w)], self._grad_off[id(b)]
            target = self.flat_grad if self.splits == 1 else self.ws[sp]
            dw, db = target[ow:ow + w.numel()].view_as(w), target[ob:ob + b.numel()]
            if w.shape[0] == 1:
                # the value head: dZ is a single column (pitch 4 bytes: no tensor map, slow cp.async path as the 128-row operand).
                # Transposed instead: dW' [K_in, 1] = X' g with g read as the row vector [1, M], db = g . 1
                g_row = dz_l.view(1, M)
                return [problem(inp[rows], g_row[:, rows], dw.view(-1, 1), a_t=True),
                        problem(g_row[:, rows], self.ones_row[:, rows], db.view(1, 1))]
            return [problem(dz_l[rows], inp1[rows], dw, a_t=True, b_t=True, ones_out=db, ones_stored=True)]

        probs = []
        for sp in range(self.splits):
            for l in range(len(self.lv)):
                probs += wgrad(self.lv, self.dzv, self.hv, self.hv1, self.grad_baseline, l, sp)
            for l in range(len(self.lp)):
                probs += wgrad(self.lp, self.dzp, self.hp, self.hp1, self.grad_logits, l, sp)
        self.wgrad_group = TcGroup(L, probs, dev)
        self.launches_per_update = len(self.fwd_groups) + len(self.dgrad_groups) + 1

    def forward(self) -> None:
        """x, xb -> logits [M, 2A], baseline [M, 1], bootstrap [Mb, 1] (+ the activations the backward pass needs)."""
        for g in self.fwd_groups:
            g.launch()

    def backward(self, defer_sum_to=None) -> None:
        """grad_logits, grad_baseline -> .grad of every weight and bias.  `defer_sum_to` (a FlatAdam): leave the sum of the split
        weight gradients' partials to the optimizer's kernel (one launch less; the .grad tensors are complete after its step())."""
        for g in self.dgrad_groups:
            g.launch()
        self.wgrad_group.launch()
        if self.splits > 1:
            if defer_sum_to is not None:
                defer_sum_to.partials = self.ws
            else:
                torch.sum(self.ws, dim=0, out=self.flat_grad)

"""Fused AdamW over the engine's flat buffers -- the B200-side of ``optimizer.step()`` (train_caption.py:111-112,133).

Same update rule as ``torch.optim.AdamW(lr, weight_decay)`` with torch defaults (betas (0.9, 0.999), eps 1e-8, decoupled
weight decay applied to every trainable parameter, no parameter groups -- exactly what the reference loop builds); one
kernel updates the fp32 masters, both moments and the bf16 compute copies, and folds in the 1/world gradient average."""
from __future__ import annotations

import torch

from . import engine, ops


class FusedAdamW:
    def __init__(self, model, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, grad_scale=1.0):
        self.model = model
        self.store = engine.prepare(model)
        self.defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        self.param_groups = [dict(self.defaults, params=self.store.train_params)]   # LR schedules poke param_groups[i]['lr']
        self.m = torch.zeros_like(self.store.master_t)
        self.v = torch.zeros_like(self.store.master_t)
        self.t = 0
        self.grad_scale = grad_scale

    def zero_grad(self, set_to_none: bool = True):
        """Gradients live in one flat buffer that the engine's backward overwrites; nothing to do per parameter."""
        return None

    @torch.no_grad()
    def step(self):
        st = self.store
        g = self.param_groups[0]
        self.t += 1
        ops.adamw_step(st.master_t, st.grad_t, self.m, self.v, st.c16_t, g["lr"], g["betas"][0], g["betas"][1], g["eps"],
                       g["weight_decay"], self.t, self.grad_scale)
        st.mark_fresh()

    @torch.no_grad()
    def step_range(self, lo: int, hi: int, first: bool, last: bool):
        """The update restricted to elements [lo, hi) of the flat buffers (multiples of 8).  A data-parallel / graphed caller can
        update the decoder slice [0, store.n_train_dec) as soon as its gradients are final -- on another stream, while the encoder
        backward is still running -- and the rest afterwards: ``first`` advances the step count, ``last`` republishes derived
        weights (conv packs).  step() == step_range(0, n, True, True)."""
        st = self.store
        g = self.param_groups[0]
        if first:
            self.t += 1
        if hi > lo:
            ops.adamw_step(st.master_t[lo:hi], st.grad_t[lo:hi], self.m[lo:hi], self.v[lo:hi], st.c16_t[lo:hi], g["lr"], g["betas"][0],
                           g["betas"][1], g["eps"], g["weight_decay"], self.t, self.grad_scale)
        if last:
            st.mark_fresh()

    def _layout(self):
        """(name, offset, numel) of every trainable parameter in the flat buffers -- stored with the moments so that a checkpoint
        does not depend on the buffer layout (which follows the backward's completion order and may change between versions)."""
        st = self.store
        names = {id(p): n for n, p in self.model.named_parameters()}
        return [(names[id(p)], st._offset[id(p)][1], p.numel()) for p in st.train_params]

    def state_dict(self):
        return {"m": self.m, "v": self.v, "t": self.t, "layout": self._layout(),
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        mine, theirs = self._layout(), sd.get("layout")
        if theirs is None or [tuple(x) for x in theirs] == mine:
            if sd["m"].numel() != self.m.numel():
                raise ValueError(f"optimizer state holds {sd['m'].numel()} elements, this model's trainable buffer {self.m.numel()}")
            self.m.copy_(sd["m"]); self.v.copy_(sd["v"])
        else:                                                   # saved under another layout: move every parameter's slice by name
            src = {n: (o, k) for n, o, k in theirs}
            missing = [n for n, _, _ in mine if n not in src]
            if missing:
                raise KeyError(f"optimizer state has no moments for {len(missing)} trainable parameters, e.g. {missing[:3]}")
            for n, o, k in mine:
                so, sk = src[n]
                if sk != k:
                    raise ValueError(f"optimizer state of {n}: {sk} elements, expected {k}")
                self.m[o:o + k].copy_(sd["m"][so:so + k]); self.v[o:o + k].copy_(sd["v"][so:so + k])
        self.t = int(sd["t"])
        for g, saved in zip(self.param_groups, sd.get("param_groups", [])):     # lr (schedules), betas, eps, weight_decay
            g.update({k: v for k, v in saved.items() if k != "params"})

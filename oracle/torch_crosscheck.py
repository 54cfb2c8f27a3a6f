"""Independent torch-CPU statement of the same graph  --  TEST INFRASTRUCTURE ONLY.

Purpose: the numpy oracle (``path_attention_oracle.py``) hand-derives the backward pass; this file
states only the *forward* graph of the reference (tensorflow_model.py:222-230, 236-265) with torch
ops and lets torch autograd differentiate it, the way TF autodiff differentiates the reference.
Agreement between the two is one of the three pins listed in the oracle's header.  It is also the
multi-threaded CPU implementation ``bench.py`` times as the reference arm (oneDNN/MKL SGEMM is the
same class of kernel TF-CPU/Eigen would run).  Never imported by the product.
"""
from __future__ import annotations

import numpy as np
import torch


def to_torch(params, dtype=torch.float32, requires_grad=True):
    return {k: torch.tensor(np.asarray(v), dtype=dtype, requires_grad=requires_grad)
            for k, v in params.items()}


def forward(tp, src, path, tgt, mask, keep=1.0, dropout_mask=None):
    """tensorflow_model.py:236-265 with torch ops.  Index/mask arguments are numpy or torch."""
    src = torch.as_tensor(src, dtype=torch.long)
    path = torch.as_tensor(path, dtype=torch.long)
    tgt = torch.as_tensor(tgt, dtype=torch.long)
    dt = tp["W"].dtype
    mask = torch.as_tensor(mask, dtype=dt)
    B, C = src.shape
    x = torch.cat([tp["tok"][src], tp["path"][path], tp["tok"][tgt]], dim=-1)      # :238-243
    if dropout_mask is not None and keep < 1.0:
        dm = torch.as_tensor(dropout_mask, dtype=dt).reshape(B, C, -1)
        x = (x * (1.0 / keep)) * dm                                                  # :245-246
    h = torch.tanh(x.reshape(B * C, -1) @ tp["W"])                                  # :248-252
    z = (h @ tp["a"].reshape(-1, 1)).reshape(B, C, 1)                               # :254-256
    z = z + torch.log(mask).unsqueeze(2)                                            # :257-259
    alpha = torch.softmax(z, dim=1)                                                 # :260
    v = (h.reshape(B, C, -1) * alpha).sum(dim=1)                                    # :262-263
    return v, alpha.squeeze(2)


def loss(tp, v, target):
    """tensorflow_model.py:226-230."""
    logits = v @ tp["tgt"].t()
    target = torch.as_tensor(target, dtype=torch.long)
    per = torch.nn.functional.cross_entropy(logits, target, reduction="none")
    return per.sum() / v.shape[0], logits


def loss_and_grads(params, src, path, tgt, mask, target, keep=1.0, dropout_mask=None,
                   dtype=torch.float32):
    tp = to_torch(params, dtype)
    v, alpha = forward(tp, src, path, tgt, mask, keep, dropout_mask)
    L, logits = loss(tp, v, target)
    L.backward()
    grads = {k: t.grad.detach().numpy() for k, t in tp.items()}
    return float(L.detach()), grads, dict(v=v.detach().numpy(), alpha=alpha.detach().numpy(),
                                          logits=logits.detach().numpy())


class TorchCpuTrainer:
    """Multi-threaded CPU train step (forward + backward + TF1-style dense Adam) used as the
    timed CPU baseline.  Parameters live as torch tensors; Adam is restated by hand because
    torch.optim.Adam's epsilon placement differs from TF1's (tensorflow_model.py:232 [TF-lib])."""

    def __init__(self, params, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8, threads=None):
        if threads:
            torch.set_num_threads(int(threads))
        self.tp = to_torch(params, torch.float32)
        self.m = {k: torch.zeros_like(t) for k, t in self.tp.items()}
        self.v = {k: torch.zeros_like(t) for k, t in self.tp.items()}
        self.t = 0
        self.hp = (lr, beta1, beta2, eps)

    def train_step(self, src, path, tgt, mask, target, keep=1.0, dropout_mask=None) -> float:
        for t in self.tp.values():
            t.grad = None
        v, _ = forward(self.tp, src, path, tgt, mask, keep, dropout_mask)
        L, _ = loss(self.tp, v, target)
        L.backward()
        lr, b1, b2, eps = self.hp
        self.t += 1
        lr_t = lr * (1.0 - b2 ** self.t) ** 0.5 / (1.0 - b1 ** self.t)
        with torch.no_grad():
            for k, p in self.tp.items():
                g = p.grad
                self.m[k].mul_(b1).add_(g, alpha=1.0 - b1)
                self.v[k].mul_(b2).addcmul_(g, g, value=1.0 - b2)
                p.addcdiv_(self.m[k], self.v[k].sqrt().add_(eps), value=-lr_t)
        return float(L.detach())

    def forward_loss(self, src, path, tgt, mask, target) -> float:
        """Forward + full-softmax loss only (BASELINE configs[2]); no dropout, no gradient."""
        with torch.no_grad():
            v, _ = forward(self.tp, src, path, tgt, mask)
            L, _ = loss(self.tp, v, target)
        return float(L)

    def sampled_train_step(self, src, path, tgt, mask, target, sampled, logq_true, logq_sampled, keep=1.0,
                           dropout_mask=None) -> float:
        """The sampled-softmax step of path_attention_oracle.sampled_softmax_loss_and_grads with torch autograd, then
        the same dense Adam (TF1 applies IndexedSlices gradients to every row)."""
        for t in self.tp.values():
            t.grad = None
        v, _ = forward(self.tp, src, path, tgt, mask, keep, dropout_mask)
        target_t = torch.as_tensor(target, dtype=torch.long)
        sampled_t = torch.as_tensor(sampled, dtype=torch.long)
        Y = self.tp["tgt"]
        l_true = (v * Y[target_t]).sum(dim=1) - torch.as_tensor(logq_true)
        l_samp = v @ Y[sampled_t].t() - torch.as_tensor(logq_sampled)[None, :]
        hit = sampled_t[None, :] == target_t[:, None]
        l_samp = torch.where(hit, torch.full_like(l_samp, -1e9), l_samp)
        logits = torch.cat([l_true[:, None], l_samp], dim=1)
        L = torch.nn.functional.cross_entropy(logits, torch.zeros(v.shape[0], dtype=torch.long), reduction="sum") / v.shape[0]
        L.backward()
        lr, b1, b2, eps = self.hp
        self.t += 1
        lr_t = lr * (1.0 - b2 ** self.t) ** 0.5 / (1.0 - b1 ** self.t)
        with torch.no_grad():
            for k, p in self.tp.items():
                g = p.grad
                self.m[k].mul_(b1).add_(g, alpha=1.0 - b1)
                self.v[k].mul_(b2).addcmul_(g, g, value=1.0 - b2)
                p.addcdiv_(self.m[k], self.v[k].sqrt().add_(eps), value=-lr_t)
        return float(L.detach())

    def numpy_params(self):
        return {k: t.detach().numpy() for k, t in self.tp.items()}

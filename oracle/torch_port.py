"""CPU timing baseline ("port") -- TEST / BENCH INFRASTRUCTURE ONLY.

The reference's hot loop (models/fatchord_version.py:194-241 with
utils/distribution.py:87-123) restated with the same PyTorch CPU operators the reference
executes per step (cat, addmm via F.linear, the fused gru_cell, relu, uniform_, log, max,
one-hot select, exp, clamp), so that timing it on the GPU box's host cores measures what
the reference's own CPU path costs there.  /root/reference does not exist on the GPU box,
so the reference itself cannot be timed there; `tests/test_oracle_vs_reference.py` checks
(in the build container) that this port returns the reference's samples under the same
seed, and `bench.py` reports which one was timed (`cpu_baseline.kind = "port"`).
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.nn.functional as F

LOG_SCALE_MIN = float(np.log(1e-14))


def _mol_sample(y: torch.Tensor) -> torch.Tensor:
    """utils/distribution.py:87-123 on y = (1, 30, B); draws from torch's default generator."""
    nr_mix = y.size(1) // 3
    y = y.transpose(1, 2)
    logit_probs = y[:, :, :nr_mix]
    temp = logit_probs.new_empty(logit_probs.size()).uniform_(1e-5, 1.0 - 1e-5)
    temp = logit_probs - torch.log(-torch.log(temp))
    _, argmax = temp.max(dim=-1)
    one_hot = F.one_hot(argmax, nr_mix).float()
    means = torch.sum(y[:, :, nr_mix:2 * nr_mix] * one_hot, dim=-1)
    log_scales = torch.clamp(torch.sum(y[:, :, 2 * nr_mix:3 * nr_mix] * one_hot, dim=-1), min=LOG_SCALE_MIN)
    u = means.new_empty(means.size()).uniform_(1e-5, 1.0 - 1e-5)
    x = means + torch.exp(log_scales) * (torch.log(u) - torch.log(1. - u))
    return torch.clamp(torch.clamp(x, min=-1.), max=1.)


@torch.no_grad()
def generate_segments_torch(sd: dict, mels: torch.Tensor, aux: torch.Tensor, *, mode="MOL", steps=None,
                            n_classes=30, consume_gru_init=True):
    """mels (B, S, 80), aux (B, S, 128) already folded (the reference materialises them,
    fatchord_version.py:188-190).  sd: state_dict-like of torch CPU tensors.
    Returns ((B, steps) samples, seconds spent in the loop)."""
    B, S, _ = mels.shape
    S = steps or S
    H = sd["rnn1.weight_hh_l0"].shape[1]
    d = aux.shape[2] // 4
    if consume_gru_init:                               # fatchord_version.py:178-179
        torch.nn.GRUCell(H, H)
        torch.nn.GRUCell(H + d, H)
    g = lambda k: sd[k]
    h1 = torch.zeros(B, H)
    h2 = torch.zeros(B, H)
    x = torch.zeros(B, 1)
    aux_split = [aux[:, :, d * i:d * (i + 1)] for i in range(4)]
    output = []
    t0 = time.perf_counter()
    for i in range(S):
        m_t = mels[:, i, :]
        a1_t, a2_t, a3_t, a4_t = (a[:, i, :] for a in aux_split)
        x = torch.cat([x, m_t, a1_t], dim=1)
        x = F.linear(x, g("I.weight"), g("I.bias"))
        h1 = torch._VF.gru_cell(x, h1, g("rnn1.weight_ih_l0"), g("rnn1.weight_hh_l0"),
                                g("rnn1.bias_ih_l0"), g("rnn1.bias_hh_l0"))
        x = x + h1
        h2 = torch._VF.gru_cell(torch.cat([x, a2_t], dim=1), h2, g("rnn2.weight_ih_l0"), g("rnn2.weight_hh_l0"),
                                g("rnn2.bias_ih_l0"), g("rnn2.bias_hh_l0"))
        x = x + h2
        x = F.relu(F.linear(torch.cat([x, a3_t], dim=1), g("fc1.weight"), g("fc1.bias")))
        x = F.relu(F.linear(torch.cat([x, a4_t], dim=1), g("fc2.weight"), g("fc2.bias")))
        logits = F.linear(x, g("fc3.weight"), g("fc3.bias"))
        if mode == "MOL":
            sample = _mol_sample(logits.unsqueeze(0).transpose(1, 2))
            output.append(sample.view(-1))
            x = sample.transpose(0, 1)
        else:
            posterior = F.softmax(logits, dim=1)
            distrib = torch.distributions.Categorical(posterior)
            sample = 2 * distrib.sample().float() / (n_classes - 1.) - 1.
            output.append(sample)
            x = sample.unsqueeze(-1)
    elapsed = time.perf_counter() - t0
    return torch.stack(output).transpose(0, 1).numpy(), elapsed

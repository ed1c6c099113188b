"""CPU oracle for the WaveRNN.generate() hot path -- TEST INFRASTRUCTURE ONLY.

A plain numpy (float32) restatement of the reference algorithm, used solely as
the checker for the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this package;
the product path (`wavernn_b200/`) never does and fails loudly without its CUDA
library.

Pinning: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4, 8c), so this oracle is pinned against outputs of the
reference itself: `tests/golden/make_golden.py` runs the UNMODIFIED reference
(imported from /root/reference through `oracle/ref_shim.py`) and commits its
outputs as fixtures; `tests/test_oracle_golden.py` checks every function below
against those fixtures, and `tests/test_oracle_vs_reference.py` re-checks live
whenever /root/reference is present.

Every function cites the reference lines it follows (paths relative to the
reference checkout).  RNG is made explicit: the reference draws, per step,
uniform_(1e-5, 1-1e-5) of shape (1,B,10) then (1,B) from torch's default CPU
generator (utils/distribution.py:106,118); here the caller passes those draws
in as `uniforms[S, 11*B]` with row t = [B*10 mixture draws, fold-major | B
logistic draws] (the layout `torch.empty(S, 11*B).uniform_()` reproduces, see
`replay_uniforms` in tests/helpers.py).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
LOG_SCALE_MIN = float(np.log(1e-14))  # utils/distribution.py:96-97


# ----------------------------------------------------------------------------
# weights
# ----------------------------------------------------------------------------
HOT_KEYS = (
    "I.weight", "I.bias",
    "rnn1.weight_ih_l0", "rnn1.weight_hh_l0", "rnn1.bias_ih_l0", "rnn1.bias_hh_l0",
    "rnn2.weight_ih_l0", "rnn2.weight_hh_l0", "rnn2.bias_ih_l0", "rnn2.bias_hh_l0",
    "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias",
)


def hot_weights(state_dict) -> dict:
    """Pulls the 16 hot-path tensors (models/fatchord_version.py:115-123,273-279)
    out of a state_dict-like mapping as float32 numpy arrays, [out, in] row-major."""
    out = {}
    for k in HOT_KEYS:
        v = state_dict[k]
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        out[k] = np.ascontiguousarray(v, dtype=F32)
    return out


# ----------------------------------------------------------------------------
# fold / unfold  (models/fatchord_version.py:281-405)
# ----------------------------------------------------------------------------
def fold_geometry(total_len: int, target: int, overlap: int):
    """fatchord_version.py:319-330.  Returns (num_folds, padded_len)."""
    num_folds = (total_len - overlap) // (target + overlap)
    extended_len = num_folds * (overlap + target) + overlap
    remaining = total_len - extended_len
    padded = total_len
    if remaining != 0:
        num_folds += 1
        padded = total_len + (target + 2 * overlap - remaining)
    return num_folds, padded


def pad_time(x: np.ndarray, pad: int, side: str = "both") -> np.ndarray:
    """fatchord_version.py:281-291 for a (T, C) array (batch of one dropped)."""
    t, c = x.shape
    total = t + 2 * pad if side == "both" else t + pad
    out = np.zeros((total, c), dtype=x.dtype)
    if side in ("before", "both"):
        out[pad:pad + t] = x
    elif side == "after":
        out[:t] = x
    return out


def fold_with_overlap(x: np.ndarray, target: int, overlap: int) -> np.ndarray:
    """fatchord_version.py:293-340.  x: (L, F) -> (num_folds, target+2*overlap, F)."""
    total_len, feats = x.shape
    num_folds, padded = fold_geometry(total_len, target, overlap)
    if padded != total_len:
        x = pad_time(x, padded - total_len, side="after")
    seg = target + 2 * overlap
    folded = np.zeros((num_folds, seg, feats), dtype=x.dtype)
    for i in range(num_folds):
        s = i * (target + overlap)
        folded[i] = x[s:s + seg]
    return folded


def xfade_and_unfold(y: np.ndarray, target: int, overlap: int) -> np.ndarray:
    """fatchord_version.py:342-405.  y: (num_folds, target+2*overlap) float64.
    Does NOT mutate its argument (the reference does, in place)."""
    y = np.array(y, dtype=np.float64, copy=True)
    num_folds, length = y.shape
    target = length - 2 * overlap
    total_len = num_folds * (target + overlap) + overlap
    silence_len = overlap // 2
    fade_len = overlap - silence_len
    t = np.linspace(-1, 1, fade_len, dtype=np.float64)
    fade_in = np.concatenate([np.zeros(silence_len), np.sqrt(0.5 * (1 + t))])
    fade_out = np.concatenate([np.ones(silence_len), np.sqrt(0.5 * (1 - t))])
    y[:, :overlap] *= fade_in
    y[:, -overlap:] *= fade_out
    out = np.zeros(total_len, dtype=np.float64)
    for i in range(num_folds):
        s = i * (target + overlap)
        out[s:s + length] += y[i]
    return out


def decode_mu_law(y: np.ndarray, mu: int) -> np.ndarray:
    """utils/dsp.py:98-103 with from_labels=False."""
    mu = mu - 1
    return np.sign(y) / mu * ((1 + mu) ** np.abs(y) - 1)


def epilogue(samples: np.ndarray, *, batched: bool, target: int, overlap: int,
             wave_len: int, hop_length: int, mode: str, mu_law: bool, n_classes: int) -> np.ndarray:
    """fatchord_version.py:243-258: float64, mu-law decode, xfade, crop, fade-out."""
    out = np.asarray(samples).astype(np.float64)
    if mu_law and mode == "RAW":
        out = decode_mu_law(out, n_classes)
    out = xfade_and_unfold(out, target, overlap) if batched else out[0]
    fade = np.linspace(1, 0, 20 * hop_length)
    out = out[:wave_len].copy()
    out[-20 * hop_length:] *= fade      # raises for wave_len < 20*hop, as the reference does
    return out


# ----------------------------------------------------------------------------
# UpsampleNetwork in numpy (fatchord_version.py:13-89), eval-mode batch norm
# ----------------------------------------------------------------------------
def _bn_eval(x, sd, prefix, eps=1e-5):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
    scale = (w / np.sqrt(rv + F32(eps))).astype(F32)
    return (x - rm[:, None]) * scale[:, None] + b[:, None]


def upsample_network(sd: dict, mel_padded: np.ndarray, *, pad: int, scales=(5, 5, 11),
                     res_blocks: int = 10):
    """mel_padded: (80, T+2*pad) float32 (already zero-padded, :185).
    Returns (mels_up (T*hop, 80), aux (T*hop, res_out)) as fatchord_version.py:82-89."""
    g = lambda k: np.asarray(sd["upsample." + k], dtype=F32)
    sdn = {k[len("upsample."):]: np.asarray(v, dtype=F32) for k, v in sd.items()
           if k.startswith("upsample.") and not k.endswith("num_batches_tracked")}
    x = mel_padded.astype(F32)
    # MelResNet (:42-48): conv_in k=2*pad+1 valid, no bias
    w = g("resnet.conv_in.weight")                      # (C, 80, k)
    k = w.shape[2]
    t_out = x.shape[1] - k + 1
    y = np.zeros((w.shape[0], t_out), dtype=F32)
    for j in range(k):
        y += w[:, :, j] @ x[:, j:j + t_out]
    y = np.maximum(_bn_eval(y, sdn, "resnet.batch_norm"), 0)
    for i in range(res_blocks):                         # ResBlock (:21-28)
        p = f"resnet.layers.{i}"
        r = y
        z = sdn[p + ".conv1.weight"][:, :, 0] @ y
        z = np.maximum(_bn_eval(z, sdn, p + ".batch_norm1"), 0)
        z = sdn[p + ".conv2.weight"][:, :, 0] @ z
        z = _bn_eval(z, sdn, p + ".batch_norm2")
        y = z + r
    aux = sdn["resnet.conv_out.weight"][:, :, 0] @ y + sdn["resnet.conv_out.bias"][:, None]
    total = int(np.prod(scales))
    aux = np.repeat(aux, total, axis=1)                 # Stretch2d (:57-61): nearest repeat
    m = x
    for li, s in enumerate(scales):                     # up_layers (:73-80)
        m = np.repeat(m, s, axis=1)
        kern = sdn[f"up_layers.{2 * li + 1}.weight"].reshape(-1)   # (2s+1,)
        mp = np.pad(m, ((0, 0), (s, s)))
        acc = np.zeros_like(m)
        for j in range(2 * s + 1):                      # cross-correlation, zero padding s
            acc += kern[j] * mp[:, j:j + m.shape[1]]
        m = acc
    indent = pad * total
    m = m[:, indent:-indent]
    return np.ascontiguousarray(m.T), np.ascontiguousarray(aux.T.astype(F32))


# ----------------------------------------------------------------------------
# the per-sample step (fatchord_version.py:201-229) and samplers
# ----------------------------------------------------------------------------
def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x, dtype=F32))).astype(F32)


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """torch.nn.GRUCell semantics, gate rows ordered [r, z, n]
    (fatchord_version.py:273-279 re-wraps nn.GRU weights into nn.GRUCell)."""
    gi = x @ w_ih.T + b_ih
    gh = h @ w_hh.T + b_hh
    H = h.shape[1]
    r = _sigmoid(gi[:, :H] + gh[:, :H])
    z = _sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = np.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:]).astype(F32)
    return ((F32(1) - z) * n + z * h).astype(F32)


def mol_sample(logits: np.ndarray, u_mix: np.ndarray, u_log: np.ndarray) -> np.ndarray:
    """utils/distribution.py:87-123 with explicit uniforms.
    logits (B, 30) = [10 mixture logits | 10 means | 10 log-scales] (:99-115);
    u_mix (B, 10), u_log (B,) both already in [1e-5, 1-1e-5]."""
    nr = logits.shape[1] // 3
    g = logits[:, :nr] - np.log(-np.log(u_mix, dtype=F32), dtype=F32)       # :107
    k = np.argmax(g, axis=1)                                                # :108 (first max)
    rows = np.arange(logits.shape[0])
    mean = logits[rows, nr + k]                                             # :113
    log_scale = np.maximum(logits[rows, 2 * nr + k], F32(LOG_SCALE_MIN))    # :114-115
    x = mean + np.exp(log_scale, dtype=F32) * (np.log(u_log, dtype=F32)
                                               - np.log(F32(1) - u_log, dtype=F32))  # :119
    return np.clip(x, F32(-1), F32(1)).astype(F32)                          # :121


def raw_sample(logits: np.ndarray, expo: np.ndarray, n_classes: int):
    """fatchord_version.py:231-237 with explicit Exp(1) draws: softmax, then
    Categorical.sample() == torch.multinomial(probs,1,True) == argmax(probs / e)
    with e ~ Exp(1) drawn per class.  Returns (sample, class idx)."""
    z = logits - logits.max(axis=1, keepdims=True)
    p = np.exp(z, dtype=F32)
    p = (p / p.sum(axis=1, keepdims=True, dtype=F32)).astype(F32)
    k = np.argmax(p / expo, axis=1)
    return (F32(2) * k.astype(F32) / F32(n_classes - 1.0) - F32(1)).astype(F32), k


def generate_segments(w: dict, mels_up: np.ndarray, aux: np.ndarray, *, n_seg: int, seg_len: int,
                      seg_stride: int, uniforms: np.ndarray | None = None, mode: str = "MOL",
                      expo: np.ndarray | None = None, x_force: np.ndarray | None = None,
                      want_logits: bool = False, steps: int | None = None):
    """The hot loop fatchord_version.py:194-241 over folds addressed as strided
    windows of the un-folded conditioning stream: fold b, step t reads row
    b*seg_stride + t of mels_up (L,80) / aux (L,128); rows >= L read as zero
    (== pad_tensor(side='after') at :327-330).  Unbatched: n_seg=1, seg_len=L.

    uniforms: (S, 11*n_seg) MOL draws (module docstring); expo: (S, n_seg, n_classes)
    Exp(1) draws for RAW.  x_force (S, n_seg): teacher forcing -- step t's input
    sample is x_force[t-1] instead of the generated one (x_force[-1] unused).
    Returns samples (n_seg, S) float32 [, logits (S, n_seg, n_classes)].
    """
    B, S = n_seg, (steps if steps is not None else seg_len)
    L = mels_up.shape[0]
    H = w["rnn1.weight_hh_l0"].shape[1]
    d = aux.shape[1] // 4
    n_classes = w["fc3.weight"].shape[0]
    h1 = np.zeros((B, H), F32)
    h2 = np.zeros((B, H), F32)
    x = np.zeros((B, 1), F32)
    out = np.zeros((B, S), F32)
    logits_all = np.zeros((S, B, n_classes), F32) if want_logits else None
    base = np.arange(B) * seg_stride
    zero_m = np.zeros((1, mels_up.shape[1]), F32)
    zero_a = np.zeros((1, aux.shape[1]), F32)
    mels_z = np.concatenate([mels_up.astype(F32), zero_m])
    aux_z = np.concatenate([aux.astype(F32), zero_a])
    for t in range(S):
        idx = np.minimum(base + t, L)                      # row L is the zero row
        m_t = mels_z[idx]
        a_t = aux_z[idx]
        a1, a2, a3, a4 = (a_t[:, d * i:d * (i + 1)] for i in range(4))
        if x_force is not None and t > 0:
            x = x_force[t - 1].reshape(B, 1).astype(F32)
        xi = np.concatenate([x, m_t, a1], axis=1)                              # :208
        xx = xi @ w["I.weight"].T + w["I.bias"]                                # :209
        h1 = gru_cell(xx, h1, w["rnn1.weight_ih_l0"], w["rnn1.weight_hh_l0"],
                      w["rnn1.bias_ih_l0"], w["rnn1.bias_hh_l0"])              # :210
        xx = xx + h1                                                           # :212
        h2 = gru_cell(np.concatenate([xx, a2], axis=1), h2, w["rnn2.weight_ih_l0"],
                      w["rnn2.weight_hh_l0"], w["rnn2.bias_ih_l0"], w["rnn2.bias_hh_l0"])  # :213-214
        xx = xx + h2                                                           # :216
        xx = np.maximum(np.concatenate([xx, a3], axis=1) @ w["fc1.weight"].T + w["fc1.bias"], 0)  # :217-218
        xx = np.maximum(np.concatenate([xx, a4], axis=1) @ w["fc2.weight"].T + w["fc2.bias"], 0)  # :220-221
        logits = (xx @ w["fc3.weight"].T + w["fc3.bias"]).astype(F32)          # :223
        if want_logits:
            logits_all[t] = logits
        if mode == "MOL":
            u = uniforms[t]
            s = mol_sample(logits, u[:10 * B].reshape(B, 10), u[10 * B:11 * B])  # :226
        elif mode == "RAW":
            s, _ = raw_sample(logits, expo[t], n_classes)                      # :231-237
        else:
            raise RuntimeError("Unknown model mode value - ", mode)           # :239
        out[:, t] = s
        x = s.reshape(B, 1)
    return (out, logits_all) if want_logits else out


def generate(w: dict, sd_upsample: dict, mel: np.ndarray, *, batched: bool, target: int, overlap: int,
             uniforms: np.ndarray | None, mode: str = "MOL", mu_law: bool = False, pad: int = 2,
             hop_length: int = 275, scales=(5, 5, 11), res_blocks: int = 10, expo=None,
             return_pre: bool = False):
    """Whole generate() (fatchord_version.py:169-264) minus the wav write.
    mel: (80, T) float32 in [0,1]."""
    T = mel.shape[1]
    wave_len = (T - 1) * hop_length                                            # :184
    mel_p = pad_time(mel.T.astype(F32), pad, "both").T                         # :185
    mels_up, aux = upsample_network(sd_upsample, mel_p, pad=pad, scales=scales, res_blocks=res_blocks)
    L = mels_up.shape[0]
    if batched:
        n_seg, _ = fold_geometry(L, target, overlap)
        seg_len, stride = target + 2 * overlap, target + overlap
    else:
        n_seg, seg_len, stride = 1, L, L
    samples = generate_segments(w, mels_up, aux, n_seg=n_seg, seg_len=seg_len, seg_stride=stride,
                                uniforms=uniforms, mode=mode, expo=expo)
    n_classes = w["fc3.weight"].shape[0]
    wav = epilogue(samples, batched=batched, target=target, overlap=overlap, wave_len=wave_len,
                   hop_length=hop_length, mode=mode, mu_law=(mu_law and mode == "RAW"), n_classes=n_classes)
    return (wav, samples) if return_pre else wav

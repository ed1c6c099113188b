"""GPU-side test plumbing: drive the C-ABI engine directly with torch device buffers."""
from __future__ import annotations

import numpy as np
import torch

from wavernn_b200 import cabi


def run_engine(model, m_up, aux, *, n_seg, seg_len, seg_stride, uniforms=None, expo=None, x_force=None,
               want_logits=False, steps=0, precision="fp16", engine="auto", seg_first=0, philox_seed=0):
    """model: our WaveRNN on cuda; m_up/aux: numpy (L,80)/(L,128).  Returns numpy samples
    (n_seg, S) [, logits (S, n_seg, n_classes)] and the engine name."""
    dev = next(model.parameters()).device
    eng = cabi.Engine(model.hot_state(), n_classes=model.n_classes, mode=model.mode, precision=precision,
                      engine=engine, device=dev.index or 0)
    try:
        S = steps or seg_len
        t = lambda a: None if a is None else torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float32, device=dev)
        d_m, d_a, d_u, d_e, d_x = t(m_up), t(aux), t(uniforms), t(expo), t(x_force)
        out = torch.full((n_seg, S), float("nan"), dtype=torch.float32, device=dev)
        lg = torch.zeros((S, n_seg, model.n_classes), dtype=torch.float32, device=dev) if want_logits else None
        p = lambda a: 0 if a is None else a.data_ptr()
        eng.generate(mels_up=p(d_m), aux=p(d_a), L=m_up.shape[0], n_seg=n_seg, seg_len=seg_len,
                     seg_stride=seg_stride, out=p(out), seg_first=seg_first, steps=steps, uniforms=p(d_u),
                     expo=p(d_e), philox_seed=philox_seed, x_force=p(d_x), logits_out=p(lg),
                     stream=torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize(dev)
        eng.check()
        name = eng.name
        res = out.cpu().numpy()
        return (res, lg.cpu().numpy(), name) if want_logits else (res, name)
    finally:
        eng.close()

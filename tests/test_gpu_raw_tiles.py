"""RAW 9-bit head on tiles of more than 24 folds (the `BIG` kernel instantiation: batched conditioning loads, 64 KB
candidate buffer).  Written after the round-1 GPU budget was spent, so this file is the first time that combination
runs on hardware -- it is deliberately the last GPU test file in collection order."""
import numpy as np
import pytest

import helpers
from gpu_helpers import run_engine

pytestmark = pytest.mark.gpu


def test_raw_head_large_tiles_agree_with_small_tiles_and_simt_engine():
    model = helpers.make_model(0, "RAW", "cuda")
    rs = np.random.RandomState(1)
    n_seg, seg_len, stride = 70, 40, 25                       # 2 tiles: 64 folds (big kernel) + 6 folds (small kernel)
    L = (n_seg - 1) * stride + 30                             # the last folds run past the end of the stream
    m_up, aux = rs.rand(L, 80).astype(np.float32), rs.randn(L, 128).astype(np.float32)
    kw = dict(seg_len=seg_len, seg_stride=stride, philox_seed=9)
    big, name = run_engine(model, m_up, aux, n_seg=n_seg, engine="tcgen05", **kw)
    assert name.startswith("tcgen05") and big.shape == (n_seg, seg_len)
    labels = (big + 1.0) * 255.5
    assert np.isfinite(big).all() and np.abs(labels - np.round(labels)).max() <= 1e-3 and labels.min() >= -1e-3 and labels.max() <= 511 + 1e-3
    # the same folds in tiles of <= 24 (small kernel): fold-keyed Philox + identical arithmetic -> identical picks
    for f0, n in ((0, 20), (20, 24), (44, 20), (64, 6)):
        off = f0 * stride
        part, _ = run_engine(model, m_up[off:], aux[off:], n_seg=n, seg_first=f0, engine="tcgen05", **kw)
        assert np.array_equal(part, big[f0:f0 + n]), (f0, n)
    # and the independent SIMT engine on the same Philox stream
    simt, sname = run_engine(model, m_up, aux, n_seg=n_seg, engine="simt", **kw)
    agree = (simt == big).mean()
    print(f"RAW 70 folds, tcgen05 (64 + 6 tiles) vs {sname}: identical picks {agree:.3f}")
    assert agree >= 0.9


def test_generate_many_raw_head_equals_sequential_calls_and_the_oracle(tmp_path):
    """SURVEY 8f-1 for the RAW head (reference :231-237): several utterances in one job -- per-utterance Exp(1)
    streams drawn in the order of sequential generate() calls, mu-law expansion in the epilogue.  Class picks are
    discrete: the batched job must reproduce the per-utterance waveforms exactly, and the oracle's to the fp16
    near-tie rate."""
    import torch
    from oracle import wavernn_oracle as O
    model = helpers.make_model(0, "RAW", "cuda")
    mels = [helpers.make_mel(T, seed) for T, seed in ((30, 0), (24, 3), (35, 5))]
    torch.manual_seed(99)
    seq = [model.generate(m, None, True, 2750, 275, True) for m in mels]
    torch.manual_seed(99)
    many = model.generate_many(mels, [tmp_path / f"{i}.wav" for i in range(3)], 2750, 275, True)
    assert model.gen_stats["utterances"] == 3 and model.gen_stats["engine"].startswith("tcgen05")
    for a, b in zip(seq, many):
        assert np.array_equal(a, b)
    sd = helpers.state_numpy(model)
    w = O.hot_weights(sd)
    torch.manual_seed(99)
    for m, got in zip(mels, many):
        B, _ = O.fold_geometry(m.shape[-1] * 275, 2750, 275)
        torch.nn.GRUCell(512, 512); torch.nn.GRUCell(544, 512)
        e = torch.empty(3300, B, 512)
        for t in range(3300):
            e[t].exponential_()
        ref = O.generate(w, sd, m[0].numpy(), batched=True, target=2750, overlap=275, uniforms=None, mode="RAW", expo=e.numpy(), mu_law=True)
        same = (np.abs(got - ref) <= 1e-9).mean()
        print(f"generate_many RAW vs oracle, T={m.shape[-1]} ({B} folds): identical samples {same:.4f}")
        assert got.shape == ref.shape and same >= 0.97
    # philox mode: no host draws at all, fold-keyed streams -> one job == sequential calls
    model.gen_rng = "philox"
    seq = [model.generate(m, None, True, 2750, 275, True) for m in mels]
    many = model.generate_many(mels, [None] * 3, 2750, 275, True)
    model.gen_rng = "torch"
    assert np.array_equal(seq[0], many[0])     # utterance 0 starts at global fold 0 in both

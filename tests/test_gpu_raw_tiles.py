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

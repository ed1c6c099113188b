"""The engine-arithmetic emulation (folded order, fp16/bf16 operand rounding) against the
reference-order oracle.  States the FP tolerance of the product path on the CPU."""
import numpy as np
import pytest

import helpers
from oracle import contract as C
from oracle import wavernn_oracle as O


@pytest.fixture(scope="module")
def setup():
    model = helpers.make_model(0, "MOL")
    sd = helpers.state_numpy(model)
    w = O.hot_weights(sd)
    mel_p = O.pad_time(helpers.make_mel(30, 0)[0].numpy().T, 2).T
    m_up, aux = O.upsample_network(sd, mel_p, pad=2)
    U = helpers.replay_uniforms(1234, 3300, 3)
    kw = dict(n_seg=3, seg_len=3300, seg_stride=3025, uniforms=U, steps=400)
    ref, lref = O.generate_segments(w, m_up, aux, want_logits=True, **kw)
    return w, m_up, aux, kw, ref, lref


def test_rounders():
    x = np.array([1.0, 1.00390625, 1.001953125, 65504.0, 1e6, -1e6, 3e-8, 0.1], np.float32)
    assert C.bf16_round(x)[1] == np.float32(1.0)                    # tie to even
    assert C.fp16_round(x)[4] == np.float32(65504.0) and C.fp16_round(x)[5] == np.float32(-65504.0)
    assert abs(C.fp16_round(x)[7] - 0.1) < 1e-4 and abs(C.bf16_round(x)[7] - 0.1) < 1e-3


def test_folded_fp32_equals_reference_order(setup):
    w, m_up, aux, kw, ref, lref = setup
    out, lg = C.generate_segments(w, m_up, aux, precision="fp32", want_logits=True, **kw)
    np.testing.assert_allclose(lg, lref, rtol=0, atol=2e-5)
    np.testing.assert_allclose(out, ref, rtol=0, atol=2e-5)


@pytest.mark.parametrize("precision,tol", [("fp16", 2e-3), ("bf16", 1.5e-2)])
def test_operand_rounding_teacher_forced_tolerance(setup, precision, tol):
    """Teacher-forced (reference samples fed back): per-step logits of the rounded-operand
    arithmetic stay within `tol` of the fp32 reference for this model."""
    w, m_up, aux, kw, ref, lref = setup
    _, lg = C.generate_segments(w, m_up, aux, precision=precision, want_logits=True, x_force=ref.T.copy(), **kw)
    assert np.abs(lg - lref).max() < tol

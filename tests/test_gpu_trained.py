"""GPU parity where it is hard: the reference's shipped LJSpeech checkpoint (trained weights make the sampler
chaotic) and BASELINE configs[0] at its stated size.  Fixtures: tests/golden/make_golden_trained.py (unmodified
reference, CPU, build container).  Run on the GPU box: pytest -m gpu.

Stated tolerances for the fp16 tensor-core engine on TRAINED weights (calibrated with the CPU
emulation oracle/contract.py, which predicts max 4.4e-2 / median 1.3e-4 / p99.9 1.6e-2 for the teacher-forced logits
and a first 1e-3 departure at step 180 of the worst fold):
  * teacher-forced logits vs reference logits, 600 steps x 4 folds (|logit| up to 11.6):
        max <= 1e-1, 99.9th percentile <= 3e-2, median <= 1e-3
  * free-running samples vs the reference's: |diff| <= 1e-3 for the first 25 steps of every fold; the step at which a
    fold first departs by 1e-3 / 1e-2 has a median over the folds >= 300 / >= 600 (a flipped Gumbel-argmax legitimately
    snowballs -- the fp32 restatement does it too -- and where it happens depends on the accumulation order)
  * fp32 strict engine, free-running: <= 1e-4 for the first 1000 steps
  * full free run (4 x 12,100 Tacotron mel; 19 x 12,100 cfg2): per-fold sample std within 25 % / 15 % of the
    reference's, mixture-component frequencies within 0.03 absolute
"""
import numpy as np
import pytest
import torch

import helpers
from gpu_helpers import run_engine
from oracle import wavernn_oracle as O
from wavernn_b200 import cabi

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def trained():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    model = helpers.make_model(0, "MOL", "cpu")
    missing = model.load_state_dict(helpers.pretrained_state_dict(), strict=False)
    assert not missing.missing_keys and not missing.unexpected_keys
    model = model.to("cuda")
    sd = helpers.state_numpy(model)
    return dict(model=model, sd=sd)


@pytest.fixture(scope="module")
def taco(trained):
    g = helpers.load_golden("trained_tacotron.npz")
    mel = helpers.tacotron_mels()[int(g["sentence"])]
    mel_p = O.pad_time(mel.T, 2).T
    m_up, aux = O.upsample_network(trained["sd"], mel_p, pad=2)
    U = helpers.replay_uniforms(int(g["seed"]), 12100, 4)
    return dict(g=g, mel=mel, m_up=m_up, aux=aux, U=U, kw=dict(n_seg=4, seg_len=12100, seg_stride=11550))


def test_trained_teacher_forced_logits(trained, taco):
    g = taco["g"]
    for prec, (tmax, tp999, tmed) in (("fp16", (1e-1, 3e-2, 1e-3)), ("fp32", (1e-3, 5e-4, 1e-5))):
        out, lg, name = run_engine(trained["model"], taco["m_up"], taco["aux"], uniforms=taco["U"],
                                   x_force=g["raw"].T.copy(), want_logits=True, steps=600, precision=prec, **taco["kw"])
        e = np.abs(lg - g["logits"])
        print(f"{name} trained teacher-forced logits: max {e.max():.3e} p99.9 {np.quantile(e, 0.999):.3e} median {np.median(e):.3e}")
        assert name.startswith("tcgen05" if prec == "fp16" else "simt")
        assert e.max() <= tmax and np.quantile(e, 0.999) <= tp999 and np.median(e) <= tmed


def test_trained_free_running_prefix_and_statistics(trained, taco):
    g = taco["g"]
    out, lg, name = run_engine(trained["model"], taco["m_up"], taco["aux"], uniforms=taco["U"], want_logits=True, **taco["kw"])
    assert name.startswith("tcgen05")
    d = np.abs(out - g["raw"])
    first3 = [helpers.first_exceed(d[i], 1e-3) for i in range(4)]
    first2 = [helpers.first_exceed(d[i], 1e-2) for i in range(4)]
    print(f"{name} trained free run: first step off by 1e-3 per fold {first3}, by 1e-2 {first2}")
    # chaotic system: WHERE a fold leaves the reference depends on the accumulation order (CPU emulation of the same
    # contract: [1618, 511, 180, never]; B200 round 2: [3541, 965, 42, 3138]), so the bound is on the ensemble
    big = 12100
    assert d[:, :25].max() <= 1e-3
    assert np.median([big if f is None else f for f in first3]) >= 300
    assert np.median([big if f is None else f for f in first2]) >= 600
    assert np.isfinite(out).all() and np.abs(out).max() <= 1.0
    # distribution statistics of the whole run
    std_ref, std = g["std"], out.std(axis=1)
    print("per-fold std ours", std, "reference", std_ref)
    assert np.all(np.abs(std - std_ref) <= 0.25 * std_ref)
    picks = helpers.mol_component_picks(lg, taco["U"])
    freq = np.bincount(picks.ravel(), minlength=10) / picks.size
    freq_ref = g["hist"] / g["hist"].sum()
    print("component frequencies ours", np.round(freq, 4), "reference", np.round(freq_ref, 4))
    assert np.abs(freq - freq_ref).max() <= 0.03
    # the strict engine tracks the reference for a long prefix
    out32, _ = run_engine(trained["model"], taco["m_up"], taco["aux"], uniforms=taco["U"], steps=1000, precision="fp32", **taco["kw"])
    assert np.abs(out32 - g["raw"][:, :1000]).max() <= 1e-4


def test_trained_public_generate_on_tacotron_mel(trained, taco, tmp_path):
    """The drop-in call with the shipped checkpoint and a real Tacotron mel: quiet leading frames keep the two runs
    together for the first ~1,000 samples of the waveform; the whole waveform must have the reference's loudness."""
    model, g = trained["model"], taco["g"]
    torch.manual_seed(int(g["seed"]))
    wav = model.generate(torch.from_numpy(taco["mel"]).unsqueeze(0), tmp_path / "t.wav", True, 11000, 550, True)
    assert wav.shape == g["wav"].shape and model.gen_stats["engine"].startswith("tcgen05")
    d = np.abs(wav - g["wav"])
    print("generate() trained vs reference wav: first 1e-2 departure at", helpers.first_exceed(d, 1e-2), "rms", wav.std(), g["wav"].std())
    assert d[:400].max() <= 1e-2
    assert abs(wav.std() - g["wav"].std()) <= 0.25 * g["wav"].std()


def test_trained_cfg2_full_size_statistics(trained):
    """BASELINE configs[1] with the checkpoint: 19 folds x 12,100 steps, free running."""
    g = helpers.load_golden("trained_cfg2.npz")
    mel_p = O.pad_time(helpers.make_mel(800, 0)[0].numpy().T, 2).T
    m_up, aux = O.upsample_network(trained["sd"], mel_p, pad=2)
    U = helpers.replay_uniforms(int(g["seed"]), 12100, 19)
    out, lg, name = run_engine(trained["model"], m_up, aux, uniforms=U, want_logits=True, n_seg=19, seg_len=12100, seg_stride=11550)
    assert name.startswith("tcgen05") and np.isfinite(out).all()
    d = np.abs(out - g["raw"])
    first = [helpers.first_exceed(d[i], 1e-3) for i in range(19)]
    print("cfg2 trained: first 1e-3 departure per fold", first)
    # (the CPU emulation of the fp16 contract departs at steps 20 .. 1240, median 562, on this input)
    assert d[:, :15].max() <= 1e-3 and np.median([12100 if f is None else f for f in first]) >= 100
    std_ref, std = g["std"], out.std(axis=1)
    print("std ours/ref", np.round(std / std_ref, 3))
    assert abs(out.std() - g["raw"].std()) <= 0.15 * g["raw"].std()
    picks = helpers.mol_component_picks(lg, U)
    freq, freq_ref = np.bincount(picks.ravel(), minlength=10) / picks.size, g["hist"] / g["hist"].sum()
    print("component frequencies ours", np.round(freq, 4), "reference", np.round(freq_ref, 4))
    assert np.abs(freq - freq_ref).max() <= 0.03


def test_cfg1_at_stated_size_matches_reference_fixture():
    """BASELINE configs[0]: one 80 x 100 random mel, MoL, unbatched, random-init weights -> 27,500 sequential steps."""
    g = helpers.load_golden("trained_cfg1.npz")
    model = helpers.make_model(0, "MOL", "cuda")
    sd = helpers.state_numpy(model)
    mel_p = O.pad_time(helpers.make_mel(100, 0)[0].numpy().T, 2).T
    m_up, aux = O.upsample_network(sd, mel_p, pad=2)
    L = 100 * 275
    U = helpers.replay_uniforms(int(g["seed"]), L, 1)
    for prec, tol in (("fp32", 1e-4), ("fp16", 2e-2)):
        out, name = run_engine(model, m_up, aux, n_seg=1, seg_len=L, seg_stride=L, uniforms=U, precision=prec)
        d = np.abs(out - g["raw"]).max()
        print(f"{name} cfg1 27,500 steps vs reference: max {d:.3e}")
        assert out.shape == (1, L) and d <= tol
    mel = helpers.make_mel(100, 0)
    torch.manual_seed(int(g["seed"]))
    wav = model.generate(mel, None, False, 11000, 550, False)
    assert wav.shape == g["wav"].shape and np.abs(wav - g["wav"]).max() <= 2e-2


def test_generate_host_entry_point_matches_device_entry_point():
    """wrnn_generate_host (host pointers in, host pointers out) == wrnn_generate on the same job, and == the reference
    fixture; both engines, teacher forcing and logits included."""
    g = helpers.load_golden("mol_batched.npz")
    model = helpers.make_model(0, "MOL", "cuda")
    sd = helpers.state_numpy(model)
    mel_p = O.pad_time(helpers.make_mel(30, 0)[0].numpy().T, 2).T
    m_up, aux = O.upsample_network(sd, mel_p, pad=2)
    U = helpers.replay_uniforms(1234, 3300, 3)
    kw = dict(n_seg=3, seg_len=3300, seg_stride=3025)
    for prec, tol in (("fp16", 2e-2), ("fp32", 1e-4)):
        eng = cabi.Engine(model.hot_state(), n_classes=30, mode="MOL", precision=prec)
        try:
            out_h, lg_h = eng.generate_host(mels_up=m_up, aux=aux, uniforms=U, want_logits=True, **kw)
            name = eng.name
            forced, lg_f = eng.generate_host(mels_up=m_up, aux=aux, uniforms=U, x_force=g["raw"].T.copy(), steps=600,
                                             want_logits=True, **kw)
        finally:
            eng.close()
        out_d, lg_d, _ = run_engine(model, m_up, aux, uniforms=U, want_logits=True, precision=prec, **kw)
        assert np.array_equal(out_h, out_d) and np.array_equal(lg_h, lg_d)
        print(f"{name} generate_host vs reference raw: {np.abs(out_h - g['raw']).max():.3e}; teacher-forced logits {np.abs(lg_f - g['logits']).max():.3e}")
        assert np.abs(out_h - g["raw"]).max() <= tol
        assert np.abs(lg_f - g["logits"]).max() <= (5e-3 if prec == "fp16" else 1e-4)

"""The reference's own CLIs, UNMODIFIED, on top of wavernn_b200 (north_star: "gen_wavernn.py and gen_tacotron.py drop
in unchanged"; reference gen_wavernn.py:38-65,112-137, gen_tacotron.py:78-92,139-163).

Container-only (needs /root/reference; no GPU): a copy of the reference checkout is made under tmp (the original is
read-only and the scripts create their output directories inside the checkout, utils/paths.py:9-46), the scripts are
byte-compared with the originals, and run in a subprocess through tests/dropin_harness.py, which replaces nothing but
the CUDA engine behind the C ABI (CPU stand-in = the numpy oracle).  Both integration routes of INTEGRATION.md are
covered: the import-time launcher `python -m wavernn_b200.dropin <script>` on a pristine checkout, and the one-file swap
(dropin/models/fatchord_version.py copied over models/fatchord_version.py).
"""
import filecmp
import json
import os
import shutil
import subprocess
import sys
import zipfile
from pathlib import Path

import numpy as np
import pytest
import torch
from scipy.io import wavfile

import helpers
from oracle import ref_shim
from oracle import wavernn_oracle as O

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not mounted")
ROOT = Path(__file__).resolve().parent.parent
REF = Path(ref_shim.REF_ROOT)


@pytest.fixture(scope="module")
def checkout(tmp_path_factory):
    """Pristine copy of the reference (code only) + the shipped checkpoints unpacked where the scripts look for them."""
    dst = tmp_path_factory.mktemp("ref") / "WaveRNN"
    shutil.copytree(REF, dst, ignore=shutil.ignore_patterns("pretrained", "notebooks", "assets", ".git*"))
    for p in dst.rglob("*"):
        os.chmod(p, 0o755 if p.is_dir() else 0o644)
    os.chmod(dst, 0o755)
    for zname, sub in (("ljspeech.wavernn.mol.800k.zip", "ljspeech_mol.wavernn"),
                       ("ljspeech.tacotron.r2.180k.zip", "ljspeech_lsa_smooth_attention.tacotron")):
        d = dst / "checkpoints" / sub
        d.mkdir(parents=True)
        with zipfile.ZipFile(REF / "pretrained" / zname) as z:
            z.extract("latest_weights.pyt", d)
    return dst


def _run(checkout, mode, script, args, seed, tmp_path):
    for name in ("gen_wavernn.py", "gen_tacotron.py", "hparams.py", "utils/paths.py", "utils/dsp.py", "models/tacotron.py"):
        assert filecmp.cmp(checkout / name, REF / name, shallow=False), f"{name} differs from the reference's"
    rec = tmp_path / f"record_{mode}.json"
    env = dict(os.environ, PYTHONPATH=str(ROOT), PYTHONWARNINGS="ignore")
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "dropin_harness.py"), mode, str(rec), str(seed), script] + args,
                       cwd=checkout, env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return json.loads(rec.read_text()), r.stdout


def _oracle_wav(mel, seed, batched, target, overlap):
    sd = {k: v.numpy() for k, v in helpers.pretrained_state_dict().items()}
    T = mel.shape[-1]
    if batched:
        B, _ = O.fold_geometry(T * 275, target, overlap)
        S = target + 2 * overlap
    else:
        B, S = 1, T * 275
    U = helpers.replay_uniforms(seed, S, B)
    return O.generate(O.hot_weights(sd), sd, mel, batched=batched, target=target, overlap=overlap, uniforms=U)


@pytest.mark.parametrize("mode", ["launcher", "swap"])
def test_gen_wavernn_gen_from_file_runs_unchanged(checkout, mode, tmp_path):
    """gen_wavernn.py --file mel.npy: WaveRNN(**hp...) (:112-123), .load(latest_weights.pyt) (:129), gen_from_file
    (:38-65) -> model.generate(mel, save_str, batched, target, overlap, hp.mu_law)."""
    fv = checkout / "models" / "fatchord_version.py"
    if mode == "swap":
        shutil.copyfile(ROOT / "dropin" / "models" / "fatchord_version.py", fv)
    else:
        shutil.copyfile(REF / "models" / "fatchord_version.py", fv)
        assert filecmp.cmp(fv, REF / "models" / "fatchord_version.py", shallow=False)
    mel = helpers.tacotron_mels()[15][:, :40]                      # 40 frames of a real mel, (80, T) in [0, 1]
    np.save(tmp_path / "utt.npy", mel)
    rec, out = _run(checkout, mode, "gen_wavernn.py", ["--file", str(tmp_path / "utt.npy"), "--batched", "-t", "1100", "-o", "110",
                                                       "--force_cpu"], 4321, tmp_path)
    assert rec["waveRNN_is_ours"]
    assert rec["module_file"] == (str(fv) if mode == "swap" else str(ROOT / "wavernn_b200" / "vocoder.py"))
    # ctor: exactly the keyword call of gen_wavernn.py:112-123 with hparams.py values
    assert rec["ctor"] == [dict(rnn_dims=512, fc_dims=512, bits=9, pad=2, upsample_factors=[5, 5, 11], feat_dims=80,
                                compute_dims=128, res_out_dims=128, res_blocks=10, hop_length=275, sample_rate=22050,
                                mode="MOL", n_positional=0)]
    assert rec["load"] == [str(checkout / "checkpoints" / "ljspeech_mol.wavernn" / "latest_weights.pyt")]
    (g,) = rec["generate"]
    save = checkout / "model_outputs" / "ljspeech_mol.wavernn" / "__utt__797k_steps_gen_batched_target1100_overlap110.wav"
    assert g == dict(mel_shape=[1, 80, 40], mel_type="Tensor", save_path=str(save), batched=True, target=1100, overlap=110,
                     mu_law=True, step=797232, training_before=True, wav_len=39 * 275, wav_dtype="float64", training_after=True)
    assert rec["engine_ctor"]["n_classes"] == 30 and rec["engine_ctor"]["mode"] == "MOL"
    # side effect: float32 wav at hp.sample_rate, equal to the oracle's generate() on the same mel / seed
    sr, data = wavfile.read(save)
    want = _oracle_wav(mel, 4321, True, 1100, 110)
    assert sr == 22050 and data.dtype == np.float32 and data.shape == want.shape
    np.testing.assert_allclose(data, want.astype(np.float32), rtol=0, atol=1e-4)
    assert "Trainable Parameters" in out


def test_gen_tacotron_wavernn_loop_runs_unchanged(checkout, tmp_path):
    """gen_tacotron.py -i <text> wavernn: Tacotron (reference, CPU) -> (m + 4) / 8 clipped -> voc_model.generate(m,
    save_path, batched, hp.voc_target, hp.voc_overlap, hp.mu_law) (:139-163) -- on a pristine checkout via the launcher."""
    shutil.copyfile(REF / "models" / "fatchord_version.py", checkout / "models" / "fatchord_version.py")
    rec, out = _run(checkout, "launcher", "gen_tacotron.py", ["--input_text", "Thank you.", "--force_cpu", "wavernn", "--batched"],
                    99, tmp_path)
    assert rec["waveRNN_is_ours"]
    assert rec["ctor"][0]["mode"] == "MOL" and rec["ctor"][0]["n_positional"] == 0
    assert rec["load"] == [str(checkout / "checkpoints" / "ljspeech_mol.wavernn" / "latest_weights.pyt")]
    (g,) = rec["generate"]
    assert g["batched"] is True and g["target"] == 11000 and g["overlap"] == 550 and g["mu_law"] is True
    assert g["mel_type"] == "Tensor" and g["mel_shape"][:2] == [1, 80] and g["step"] == 797232
    save = Path(g["save_path"])
    assert save.parent == checkout / "model_outputs" / "ljspeech_lsa_smooth_attention.tacotron" and save.name.startswith("__input_Thank you._wavernn_batched_")
    mel = np.load(str(save) + ".mel.npy")[0]
    assert mel.min() >= 0.0 and mel.max() <= 1.0
    sr, data = wavfile.read(save)
    want = _oracle_wav(mel, 99, True, 11000, 550)
    assert sr == 22050 and data.shape == want.shape == ((mel.shape[1] - 1) * 275,)
    # trained weights amplify the 1e-7 difference between torch's and the oracle's upsampling (12,100 chaotic steps)
    np.testing.assert_allclose(data, want.astype(np.float32), rtol=0, atol=5e-3)

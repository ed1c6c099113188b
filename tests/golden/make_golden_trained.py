"""Trained-checkpoint fixtures from the UNMODIFIED reference (container-only; needs /root/reference):

    python tests/golden/make_golden_trained.py

Round-1 parity ran on random-init weights only -- a contracting system where fp16 operand rounding lands at 7e-5.
With the shipped LJSpeech checkpoint the sampler is chaotic (a Gumbel-argmax flip snowballs; even an fp32
re-association leaves 1e-4 within ~2,000 steps), so parity with trained weights is stated as
  (a) teacher-forced logits (inputs forced to the reference's own samples): a tolerance per step,
  (b) free-running prefix agreement: |ours - reference| small for a stated number of steps,
  (c) distribution statistics of a full free run: sample std, mixture-component histogram.
This script commits what those checks need:
  pretrained/ljspeech.wavernn.mol.800k.zip  -- the reference's shipped checkpoint, byte for byte (the GPU box has no
                                               /root/reference; `WaveRNN.load` reads its `latest_weights.pyt`)
  trained_tacotron.npz  -- sentence 9 of tacotron_mels.npz (164 frames -> 4 folds x 12,100 steps, hparams fold
                           geometry) through reference generate(): raw samples, wav, logits of the first 600 steps,
                           per-fold std, mixture-component histogram
  trained_cfg2.npz      -- BASELINE configs[1] (torch.rand mel, T=800 -> 19 folds x 12,100) with the checkpoint:
                           raw samples (fp16-packed is NOT used: fp32), std, component histogram
  trained_cfg1.npz      -- BASELINE configs[0] at its stated size (T=100, unbatched, 27,500 steps), random-init
                           weights seed 0 (as the config says), raw samples + wav
"""
import shutil
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent))

from oracle import ref_shim  # noqa: E402
import helpers  # noqa: E402


def run(model, mel, batched, target, overlap, seed, n_logits):
    rec = []
    h = model.fc3.register_forward_hook(lambda m, i, o: rec.append(o.detach().numpy().copy()))
    try:
        r = ref_shim.ref_generate(model, mel, batched, target, overlap, seed=seed)
    finally:
        h.remove()
    logits = np.stack(rec)                                     # (S, B, 30)
    S, B = logits.shape[:2]
    U = helpers.replay_uniforms(seed, S, B)
    picks = helpers.mol_component_picks(logits, U)
    hist = np.bincount(picks.ravel(), minlength=10).astype(np.int64)
    return r, logits[:n_logits].astype(np.float32), hist


def main():
    (HERE / "pretrained").mkdir(exist_ok=True)
    shutil.copyfile(Path(ref_shim.REF_ROOT) / "pretrained" / "ljspeech.wavernn.mol.800k.zip",
                    HERE / "pretrained" / "ljspeech.wavernn.mol.800k.zip")
    model = ref_shim.build_reference_model(seed=0, mode="MOL", pretrained=True)
    model.eval()

    mels = helpers.tacotron_mels()
    mel = torch.from_numpy(mels[9]).unsqueeze(0)
    r, logits, hist = run(model, mel, True, 11000, 550, 1234, 600)
    np.savez_compressed(HERE / "trained_tacotron.npz", sentence=9, T=mel.shape[-1], target=11000, overlap=550, seed=1234,
                        raw=r["raw"], wav=r["wav"], logits=logits, hist=hist, std=r["raw"].std(axis=1))
    print("tacotron mel 9:", r["raw"].shape, "std", r["raw"].std(), "hist", hist)

    mel = helpers.make_mel(800, 0)
    r, logits, hist = run(model, mel, True, 11000, 550, 1234, 0)
    np.savez_compressed(HERE / "trained_cfg2.npz", T=800, target=11000, overlap=550, seed=1234,
                        raw=r["raw"], hist=hist, std=r["raw"].std(axis=1), wav_std=np.float64(r["wav"].std()))
    print("cfg2 trained:", r["raw"].shape, "std", r["raw"].std(), "hist", hist)

    model0 = ref_shim.build_reference_model(seed=0, mode="MOL")
    model0.eval()
    mel = helpers.make_mel(100, 0)
    r = ref_shim.ref_generate(model0, mel, False, 11000, 550, seed=1234)
    np.savez_compressed(HERE / "trained_cfg1.npz", T=100, seed=1234, raw=r["raw"], wav=r["wav"])
    print("cfg1:", r["raw"].shape)


if __name__ == "__main__":
    main()

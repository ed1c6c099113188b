"""Generates the committed fixtures in this directory from the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shim.py).  Run in the build container:

    python tests/golden/make_golden.py

The reference has no tests or golden vectors of its own, so these
files ARE the pin: seeded inputs -> the reference's own outputs.  Weights are not stored
(17 MB); they are re-created from the seed, and a fingerprint of every tensor is stored
so a drift in torch's initialisers would be detected rather than silently compared.
"""
import sys
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent))

from oracle import ref_shim  # noqa: E402
import helpers  # noqa: E402


def capture_logits(model, fn):
    """Runs fn() while recording fc3 outputs (per step logits)."""
    rec = []
    h = model.fc3.register_forward_hook(lambda m, i, o: rec.append(o.detach().numpy().copy()))
    try:
        out = fn()
    finally:
        h.remove()
    return out, np.stack(rec)


def main():
    ref = ref_shim.load_reference()
    from utils.distribution import sample_from_discretized_mix_logistic
    from utils.dsp import decode_mu_law

    # ---- A/B: MoL, random init seed 0 --------------------------------------------------
    model = ref_shim.build_reference_model(seed=0, mode="MOL")
    model.eval()
    np.savez_compressed(HERE / "weights_mol_seed0.npz", **helpers.weight_fingerprint(helpers.state_numpy(model)))
    mel = helpers.make_mel(30, seed=0)
    with torch.no_grad():
        mp = model.pad_tensor(mel.transpose(1, 2), pad=2, side="both").transpose(1, 2)
        m_up, aux = model.upsample(mp)
    r, logits = capture_logits(model, lambda: ref_shim.ref_generate(model, mel, True, 2750, 275, seed=1234))
    np.savez_compressed(HERE / "mol_batched.npz", T=30, target=2750, overlap=275, seed=1234,
                        raw=r["raw"], wav=r["wav"], logits=logits[:600],
                        mels_up_rows=m_up[0, ::97].numpy(), aux_rows=aux[0, ::97].numpy(),
                        mels_up_sum=np.float64(m_up.double().sum()), aux_sum=np.float64(aux.double().sum()))
    mel_u = helpers.make_mel(22, seed=1)
    r = ref_shim.ref_generate(model, mel_u, False, 11000, 550, seed=77)
    np.savez_compressed(HERE / "mol_unbatched.npz", T=22, seed=77, mel_seed=1, raw=r["raw"], wav=r["wav"])
    # exact fold (no padding): L = 5*(550+55)+55 -> T cannot be chosen freely (L = T*275); use target/overlap
    # such that (L - overlap) % (target+overlap) == 0:  T=33 -> L=9075 = 3*(2750+275)... 9075-275 = 8800 no.
    # ---- C: RAW 9-bit ------------------------------------------------------------------
    model_r = ref_shim.build_reference_model(seed=0, mode="RAW", bits=9)
    model_r.eval()
    np.savez_compressed(HERE / "weights_raw_seed0.npz", **helpers.weight_fingerprint(helpers.state_numpy(model_r)))
    r, logits = capture_logits(model_r, lambda: ref_shim.ref_generate(model_r, mel, True, 2750, 275, mu_law=True, seed=1234))
    expo = helpers.replay_expo(1234, 3300, 3, 512)
    np.savez_compressed(HERE / "raw_batched.npz", T=30, target=2750, overlap=275, seed=1234, raw=r["raw"],
                        wav=r["wav"], logits=logits[:64], expo_head=expo[:8],
                        expo_sum=np.float64(expo.astype(np.float64).sum()))
    # ---- E: small functions -----------------------------------------------------------
    x = torch.arange(1 * 1000 * 3, dtype=torch.float32).reshape(1, 1000, 3)
    f_pad = model.fold_with_overlap(x, 300, 30).numpy()          # remaining != 0 -> padded last fold
    x2 = x[:, :3 * 330 + 30]
    f_exact = model.fold_with_overlap(x2, 300, 30).numpy()       # exact
    rs = np.random.RandomState(5)
    y = rs.randn(4, 360)
    xf = model.xfade_and_unfold(y.copy(), 300, 30)
    torch.manual_seed(99)
    lg = torch.randn(1, 30, 64) * 2
    torch.manual_seed(100)
    smp = sample_from_discretized_mix_logistic(lg).numpy()
    torch.manual_seed(100)
    u_mix = torch.empty(1, 64, 10).uniform_(1e-5, 1 - 1e-5).numpy()
    u_log = torch.empty(1, 64).uniform_(1e-5, 1 - 1e-5).numpy()
    grid = np.linspace(-1, 1, 101)
    np.savez_compressed(HERE / "functions.npz", fold_pad=f_pad, fold_exact=f_exact, xfade_in=y, xfade_out=xf,
                        mol_logits=lg.numpy(), mol_sample=smp, mol_u_mix=u_mix, mol_u_log=u_log,
                        mulaw_in=grid, mulaw_out=decode_mu_law(grid, 512, False))
    print("wrote", sorted(p.name for p in HERE.glob("*.npz")))


if __name__ == "__main__":
    main()

"""Live pin of the oracle and the CPU timing port against the UNMODIFIED reference.
Only runs where /root/reference is mounted (the build container); the committed fixtures
in tests/golden carry the same comparison to the GPU box."""
import numpy as np
import pytest
import torch

import helpers
from oracle import ref_shim, torch_port
from oracle import wavernn_oracle as O

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference checkout not mounted")


def test_oracle_tracks_reference_with_pretrained_weights():
    """Trained weights make the sampler chaotic: agreement is a PREFIX property."""
    model = ref_shim.build_reference_model(seed=0, mode="MOL", pretrained=True)
    model.eval()
    sd = helpers.state_numpy(model)
    mel = helpers.make_mel(30, 0)
    r = ref_shim.ref_generate(model, mel, True, 2750, 275, seed=1234)
    U = helpers.replay_uniforms(1234, 3300, 3)
    wav, pre = O.generate(O.hot_weights(sd), sd, mel[0].numpy(), batched=True, target=2750, overlap=275,
                          uniforms=U, return_pre=True)
    d = np.abs(pre - r["raw"])
    assert d[:, :1000].max() < 1e-4
    assert abs(pre.std() - r["raw"].std()) < 0.2 * r["raw"].std()


def test_torch_port_reproduces_reference_samples():
    model = ref_shim.build_reference_model(seed=0, mode="MOL")
    model.eval()
    mel = helpers.make_mel(30, 0)
    r = ref_shim.ref_generate(model, mel, True, 2750, 275, seed=1234)
    with torch.no_grad():
        mp = model.pad_tensor(mel.transpose(1, 2), pad=2, side="both").transpose(1, 2)
        m_up, aux = model.upsample(mp)
        mf, af = model.fold_with_overlap(m_up, 2750, 275), model.fold_with_overlap(aux, 2750, 275)
    torch.manual_seed(1234)
    out, _ = torch_port.generate_segments_torch({k: v.detach() for k, v in model.state_dict().items()}, mf, af)
    np.testing.assert_array_equal(out, r["raw"])


def test_dropin_class_loads_pretrained_checkpoint():
    ours = helpers.make_model(0, "MOL")
    missing = ours.load_state_dict(ref_shim.load_pretrained_state_dict(), strict=False)
    assert not missing.missing_keys and not missing.unexpected_keys
    assert ours.get_step() == 797232

"""BASELINE.json's full sizes on the GPU.

cfg2' (T=802 frames -> 20 folds x 12,100 steps, zero-padded last fold) is checked sample by sample against the CPU
emulation of the engine arithmetic (the random-init model is not chaotic, so all 242,000 samples can be compared) and
against the independent SIMT engine.  cfg5 (4096 folds) is checked through size-independent properties: every 64-fold
tile of the big job equals the same folds run on their own (tiling / sharding invariance under the fold-keyed
in-kernel RNG), everything finite and inside [-1, 1]."""
import numpy as np
import pytest
import torch

import helpers
from gpu_helpers import run_engine
from oracle import contract as C
from oracle import wavernn_oracle as O
from wavernn_b200 import cabi
from wavernn_b200.sharding import fold_geometry

pytestmark = pytest.mark.gpu


def test_cfg2_full_size_matches_emulation_and_simt_engine():
    model = helpers.make_model(0, "MOL", "cuda")
    sd = helpers.state_numpy(model)
    w = O.hot_weights(sd)
    T = 802
    mel_p = O.pad_time(helpers.make_mel(T, 0)[0].numpy().T, 2).T
    m_up, aux = O.upsample_network(sd, mel_p, pad=2)
    geo = fold_geometry(T * 275, 11000, 550)
    assert geo.n_seg == 20 and geo.padded_len > geo.total_len          # zero-padded last fold
    U = helpers.replay_uniforms(1234, geo.seg_len, geo.n_seg)
    kw = dict(n_seg=geo.n_seg, seg_len=geo.seg_len, seg_stride=geo.seg_stride, uniforms=U)
    out_tc, name = run_engine(model, m_up, aux, engine="tcgen05", **kw)
    out_simt, _ = run_engine(model, m_up, aux, engine="simt", **kw)
    emu = C.generate_segments(w, m_up, aux, precision="fp16", **kw)
    d_emu, d_simt = np.abs(out_tc - emu).max(), np.abs(out_tc - out_simt).max()
    print(f"{name} cfg2' 20 x 12100: vs emulation {d_emu:.3e}, vs simt-fp16 {d_simt:.3e}")
    assert out_tc.shape == (20, 12100) and np.isfinite(out_tc).all()
    assert d_emu <= 4e-3 and d_simt <= 4e-3          # 12,100 recurrent steps of fp16-operand rounding differences
    # the last fold runs past the end of the stream: its tail is generated from zero conditioning, not garbage
    assert np.abs(out_tc[-1, -2000:] - emu[-1, -2000:]).max() <= 4e-3


def test_cfg5_4096_folds_tiling_and_shard_invariance():
    model = helpers.make_model(0, "MOL", "cuda")
    dev = torch.device("cuda")
    T, hop, target, overlap, steps = 172_034, 275, 11_000, 550, 48
    geo = fold_geometry(T * hop, target, overlap)
    assert geo.n_seg == 4096 and geo.padded_len == geo.total_len
    torch.manual_seed(0)
    mel = torch.rand(1, 80, T, device=dev)
    with torch.no_grad():
        mp = torch.nn.functional.pad(mel, (2, 2))
        mel_fr = mp[0].transpose(0, 1).contiguous()
        aux_fr = model.eval().upsample.resnet(mp)[0].transpose(0, 1).contiguous()
        taps = model.upsample_taps(dev)
    eng = cabi.Engine(model.hot_state(), n_classes=30, mode="MOL", precision="fp16", engine="tcgen05", device=0)

    def run(f0, n, cond_mode=cabi.COND_AUTO, tables=False):
        row0 = (torch.arange(f0, f0 + n, device=dev, dtype=torch.int64) * geo.seg_stride).contiguous()
        end = torch.full_like(row0, T * hop)
        out = torch.full((n, steps), float("nan"), device=dev)
        eng.generate(mels_up=0, aux=0, L=T * hop, n_seg=n, seg_len=geo.seg_len, seg_stride=geo.seg_stride,
                     out=out.data_ptr(), seg_first=f0, steps=steps, philox_seed=11,
                     fold_row0=row0.data_ptr() if tables else 0, fold_row_end=end.data_ptr() if tables else 0,
                     mel_frames=mel_fr.data_ptr(), aux_frames=aux_fr.data_ptr(), up_taps=taps.data_ptr(), hop=hop,
                     cond_mode=cond_mode, stream=torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        eng.check()
        return out.cpu().numpy()

    big = run(0, 4096, cabi.COND_IN_KERNEL)
    assert eng.launch_count == 64 and np.isfinite(big).all() and np.abs(big).max() <= 1.0
    assert np.array_equal(run(0, 4096), big)                       # default: per-tile conditioning pre-pass, same samples
    assert eng.launch_count == 64 + 128
    assert np.array_equal(run(64 * 9, 64, tables=True), big[64 * 9:64 * 10])      # explicit fold tables (generate_many's path)
    for f0, n in ((0, 64), (64 * 37, 64), (4096 - 64, 64), (512 * 5, 512)):       # a tile, a middle tile, the last, one rank's shard of 8
        assert np.array_equal(run(f0, n), big[f0:f0 + n]), (f0, n)
    eng.close()

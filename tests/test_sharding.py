"""Fold->rank partition and the N>1 generate() path on CPU (gloo, world_size 2).  The CUDA
kernel is replaced by the oracle here (tests may do that; the product never does)."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers
from oracle import wavernn_oracle as O
from wavernn_b200.sharding import fold_geometry, shard_folds, unbatched_geometry


def test_fold_geometry_matches_oracle_for_many_lengths():
    for L in list(range(600, 5000, 37)) + [220000, 220550, 47309350]:
        for target, overlap in ((300, 30), (2750, 275), (11000, 550)):
            if L <= overlap + 1:
                continue
            g = fold_geometry(L, target, overlap)
            assert (g.n_seg, g.padded_len) == O.fold_geometry(L, target, overlap)
            assert g.seg_len == target + 2 * overlap and g.seg_stride == target + overlap
            assert (g.n_seg - 1) * g.seg_stride + g.seg_len <= g.padded_len


def test_cfg_sizes_from_survey():
    assert fold_geometry(800 * 275, 11000, 550).n_seg == 19          # cfg2, no padding
    assert fold_geometry(802 * 275, 11000, 550).n_seg == 20          # cfg2', zero-padded last fold
    assert fold_geometry(172034 * 275, 11000, 550).n_seg == 4096     # cfg5 (A)
    assert fold_geometry(286722 * 275, 18700, 550).n_seg == 4096     # cfg5 (B)
    assert unbatched_geometry(27500).n_seg == 1


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_shards_partition_folds_and_cover_their_rows(world):
    geo = fold_geometry(61 * 275, 2750, 275)
    seen = []
    for r in range(world):
        s = shard_folds(geo, r, world, 275)
        seen += list(range(s.seg_first, s.seg_first + s.n_seg))
        if s.n_seg:
            assert s.row_lo == s.seg_first * geo.seg_stride
            assert s.frame_lo * 275 <= s.row_lo and s.frame_hi * 275 >= s.row_hi
            assert s.row_hi <= geo.total_len
    assert seen == list(range(geo.n_seg))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = helpers.make_model(0, "MOL")
        sd = helpers.state_numpy(model)
        w = O.hot_weights(sd)
        calls = []

        def fake_kernel(self, mels_padded, geo, shard, device, **kw):
            """Stands in for the CUDA launch: same inputs (this rank's conditioning slice and
            RNG slice), oracle arithmetic."""
            u_all, _ = self._reference_draws(geo, geo.seg_len)
            f0, n, B = shard.seg_first, shard.n_seg, geo.n_seg
            u = torch.cat([u_all[:, 10 * f0:10 * (f0 + n)], u_all[:, 10 * B + f0:10 * B + f0 + n]], 1).numpy()
            m_up, aux = self.conditioning(mels_padded, shard.frame_lo, shard.frame_hi)
            off = shard.row_lo - shard.frame_lo * self.hop_length
            rows = shard.row_hi - shard.row_lo
            calls.append((f0, n))
            out = O.generate_segments(w, m_up[off:off + rows].numpy(), aux[off:off + rows].numpy(), n_seg=n,
                                      seg_len=geo.seg_len, seg_stride=geo.seg_stride, uniforms=u)
            return torch.from_numpy(out)

        type(model)._run_segments = fake_kernel
        type(model)._require_cuda = lambda self: torch.device("cpu")
        mel = helpers.make_mel(22, 0)
        torch.manual_seed(1234)
        wav = model.generate(mel, None, True, 550, 55, False)
        q.put((rank, wav, calls))
    finally:
        dist.destroy_process_group()


def test_two_rank_generate_equals_single_rank_oracle():
    """world_size=2 over gloo: each rank generates its folds from its own conditioning slice,
    all-gather, overlap-add == the single-process oracle result."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    model = helpers.make_model(0, "MOL")
    sd = helpers.state_numpy(model)
    geo = fold_geometry(22 * 275, 550, 55)
    U = helpers.replay_uniforms(1234, geo.seg_len, geo.n_seg)
    ref = O.generate(O.hot_weights(sd), sd, helpers.make_mel(22, 0)[0].numpy(), batched=True, target=550,
                     overlap=55, uniforms=U)
    assert res[0][2] != res[1][2] and res[0][2][0][1] + res[1][2][0][1] == geo.n_seg
    for _, wav, _ in res:
        np.testing.assert_allclose(wav, ref, rtol=0, atol=2e-5)


# ---------------------------------------------------------------------------------------------
# generate_many: host plumbing (utterances laid end to end, fold tables, RNG slices, rank sharding) with the CUDA
# pieces replaced by CPU stand-ins that read the same raw pointers the C ABI would receive
# ---------------------------------------------------------------------------------------------
def _view(ptr, shape, dtype=np.float32):
    n = int(np.prod(shape))
    ct = {np.float32: ctypes.c_float, np.int64: ctypes.c_int64}[dtype]
    return np.ctypeslib.as_array((ct * n).from_address(ptr)).reshape(shape)


def _install_generate_many_fakes(model, w, calls):
    import types
    from wavernn_b200 import vocoder

    def fake_expand(*, mel_frames, aux_frames, up_taps, hop, row_lo, n_rows, mels_up, aux, stream=0):
        n_fr = n_rows // hop + 4
        mf, af, tp = _view(mel_frames, (n_fr, 80)), _view(aux_frames, (n_fr, 128)), _view(up_taps, (hop, 5))
        r = np.arange(row_lo, row_lo + n_rows)
        fr, ph = r // hop, r % hop
        m = np.zeros((n_rows, 80), np.float32)
        for d in range(5):                                   # same tap order as the kernel
            m = m + tp[ph, d][:, None] * mf[fr + d]
        _view(mels_up, (n_rows, 80))[:] = m
        _view(aux, (n_rows, 128))[:] = af[fr]

    class FakeEngine:
        name, grid_ctas, launch_count = "fake-oracle", 0, 0

        def generate(self, *, mels_up, aux, L, n_seg, seg_len, seg_stride, out, seg_first=0, uniforms=0, fold_row0=0,
                     fold_row_end=0, **kw):
            m, a = _view(mels_up, (L, 80)), _view(aux, (L, 128))
            r0, r1 = _view(fold_row0, (n_seg,), np.int64), _view(fold_row_end, (n_seg,), np.int64)
            u = _view(uniforms, (seg_len, 11 * n_seg)).copy()
            # one private window per fold (zero beyond its end), laid out with stride == seg_len for the oracle
            mm, aa = np.zeros((n_seg * seg_len, 80), np.float32), np.zeros((n_seg * seg_len, 128), np.float32)
            for f in range(n_seg):
                k = int(min(seg_len, r1[f] - r0[f]))
                mm[f * seg_len:f * seg_len + k] = m[r0[f]:r0[f] + k]
                aa[f * seg_len:f * seg_len + k] = a[r0[f]:r0[f] + k]
            calls.append((seg_first, n_seg))
            _view(out, (n_seg, seg_len))[:] = O.generate_segments(w, mm, aa, n_seg=n_seg, seg_len=seg_len, seg_stride=seg_len, uniforms=u)

        def check(self):
            pass

    type(model)._require_cuda = lambda self: torch.device("cpu")
    type(model)._kernel_conditioning_ok = lambda self: True
    type(model)._get_engine = lambda self, device: FakeEngine()
    vocoder.cabi.expand_conditioning = fake_expand
    torch.cuda.current_stream = lambda device=None: types.SimpleNamespace(cuda_stream=0, synchronize=lambda: None)


_MANY = ((22, 0), (26, 3), (31, 5))          # (frames, mel seed) of the utterances


def _many_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        model = helpers.make_model(0, "MOL")
        model.gen_native_rng = False                      # plain torch operators: the reference's own sequence
        calls = []
        _install_generate_many_fakes(model, O.hot_weights(helpers.state_numpy(model)), calls)
        mels = [helpers.make_mel(T, s) for T, s in _MANY]
        torch.manual_seed(99)
        wavs = model.generate_many(mels, [None] * len(mels), 550, 55, False)
        q.put((rank, wavs, calls, model.gen_stats.get("world")))
    finally:
        dist.destroy_process_group()


def test_generate_many_two_ranks_equals_per_utterance_oracle():
    """generate_many over gloo (2 ranks): the job's folds are split across the ranks, all-gathered, and each
    utterance's waveform equals the oracle's generate() for that utterance under sequential-call RNG order."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_many_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=400) for _ in procs], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    model = helpers.make_model(0, "MOL")
    sd = helpers.state_numpy(model)
    w = O.hot_weights(sd)
    torch.manual_seed(99)                                  # the draws of three sequential reference generate() calls
    refs = []
    for T, s in _MANY:
        geo = fold_geometry(T * 275, 550, 55)
        mel = helpers.make_mel(T, s)                       # (re-seeds torch: restore the stream position afterwards)
        refs.append((geo, mel))
    torch.manual_seed(99)
    want = []
    for geo, mel in refs:
        torch.nn.GRUCell(512, 512); torch.nn.GRUCell(544, 512)
        U = torch.empty(geo.seg_len, 11 * geo.n_seg).uniform_(1e-5, 1 - 1e-5).numpy()
        want.append(O.generate(w, sd, mel[0].numpy(), batched=True, target=550, overlap=55, uniforms=U))
    n_total = sum(g.n_seg for g, _ in refs)
    (f0a, na), (f0b, nb) = res[0][2][0], res[1][2][0]
    assert res[0][3] == 2 and f0a == 0 and f0b == na and na + nb == n_total          # contiguous split of the whole job
    for _, wavs, _, _ in res:
        assert len(wavs) == len(want)
        for got, ref in zip(wavs, want):
            assert got.shape == ref.shape
            np.testing.assert_allclose(got, ref, rtol=0, atol=1e-4)

"""Fold->rank partition and the N>1 generate() path on CPU (gloo, world_size 2).  The CUDA
kernel is replaced by the oracle here (tests may do that; the product never does)."""
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

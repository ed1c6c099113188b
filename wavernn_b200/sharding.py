"""Fold geometry and the fold->rank partition of the generate() path.

Folds (reference `fold_with_overlap`, models/fatchord_version.py:293-340) are
independent for the whole sample loop -- each starts from zero state (:194-196) -- and
only meet again in `xfade_and_unfold` (:397-403).  So they shard across ranks with
no data-path collective; one all-gather of the (n_seg, seg_len) sample blocks feeds the
overlap-add.  Everything here is pure host logic (ints / numpy / CPU-or-GPU torch
tensors) so it is exercised by world_size-2 gloo tests without a GPU.
"""
from __future__ import annotations

from dataclasses import dataclass


@dataclass(frozen=True)
class FoldGeometry:
    total_len: int      # L: un-folded conditioning length (samples)
    target: int
    overlap: int
    n_seg: int          # number of folds B
    seg_len: int        # target + 2*overlap (steps per fold)
    seg_stride: int     # target + overlap
    padded_len: int     # L plus the zero padding fold_with_overlap appends

    @property
    def unfolded_len(self) -> int:
        return self.n_seg * self.seg_stride + self.overlap


def fold_geometry(total_len: int, target: int, overlap: int) -> FoldGeometry:
    """Index math of fatchord_version.py:319-330."""
    if target <= 0 or overlap < 0:
        raise ValueError("target must be > 0 and overlap >= 0")
    n = (total_len - overlap) // (target + overlap)
    remaining = total_len - (n * (overlap + target) + overlap)
    padded = total_len
    if remaining != 0:
        n += 1
        padded = total_len + (target + 2 * overlap - remaining)
    return FoldGeometry(total_len, target, overlap, n, target + 2 * overlap, target + overlap, padded)


def unbatched_geometry(total_len: int) -> FoldGeometry:
    """batched=False: one segment covering the whole stream (fatchord_version.py:192)."""
    return FoldGeometry(total_len, total_len, 0, 1, total_len, total_len, total_len)


@dataclass(frozen=True)
class RankShard:
    rank: int
    world: int
    seg_first: int      # global index of this rank's first fold
    n_seg: int          # folds on this rank (may be 0)
    row_lo: int         # first conditioning row (sample) this rank reads
    row_hi: int         # one past the last EXISTING row it reads (<= L)
    frame_lo: int       # mel frames [frame_lo, frame_hi) cover rows [row_lo, row_hi)
    frame_hi: int
    n_seg_max: int      # max folds on any rank (all-gather block size)


def shard_folds(geo: FoldGeometry, rank: int, world: int, hop_length: int) -> RankShard:
    """Contiguous balanced partition: rank r owns folds [B*r//R, B*(r+1)//R)."""
    if not (0 <= rank < world):
        raise ValueError("bad rank/world")
    f0 = geo.n_seg * rank // world
    f1 = geo.n_seg * (rank + 1) // world
    n_max = max(geo.n_seg * (r + 1) // world - geo.n_seg * r // world for r in range(world))
    if f1 == f0:
        return RankShard(rank, world, f0, 0, 0, 0, 0, 0, n_max)
    row_lo = f0 * geo.seg_stride
    row_hi = min((f1 - 1) * geo.seg_stride + geo.seg_len, geo.total_len)
    frame_lo = row_lo // hop_length
    frame_hi = -(-row_hi // hop_length)
    return RankShard(rank, world, f0, f1 - f0, row_lo, row_hi, frame_lo, frame_hi, n_max)


def gather_segments(local, shard: RankShard, geo: FoldGeometry, group=None):
    """All-gather of the per-rank (n_seg, S) float32 sample blocks into (B, S) on every
    rank (NCCL over NVLink on GPUs, gloo in CPU tests).  `local` is a torch tensor."""
    import torch
    import torch.distributed as dist
    if shard.world == 1:
        return local
    S = local.shape[1] if local.dim() == 2 else geo.seg_len
    block = torch.zeros((shard.n_seg_max, S), dtype=torch.float32, device=local.device)
    if shard.n_seg:
        block[:shard.n_seg] = local
    full = torch.empty((shard.world * shard.n_seg_max, S), dtype=torch.float32, device=local.device)
    dist.all_gather_into_tensor(full, block, group=group)
    pieces = []
    for r in range(shard.world):
        n_r = geo.n_seg * (r + 1) // shard.world - geo.n_seg * r // shard.world
        pieces.append(full[r * shard.n_seg_max: r * shard.n_seg_max + n_r])
    return torch.cat(pieces, dim=0)

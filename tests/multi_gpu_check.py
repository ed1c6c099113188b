"""Run under torchrun with N >= 2 ranks (one per GPU): every rank calls the public
WaveRNN.generate() with the same mel and the same torch seed; folds are sharded across ranks,
all-gathered over NCCL and overlap-added.  The result must equal the committed reference
fixture (and therefore the single-GPU result) within the fp16 tolerance."""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import helpers  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    g = helpers.load_golden("mol_batched.npz")
    model = helpers.make_model(0, "MOL", f"cuda:{local}")
    mel = helpers.make_mel(30, 0)
    torch.manual_seed(1234)
    wav = model.generate(mel, None, True, 2750, 275, False)
    err = float(np.abs(wav - g["wav"]).max())
    # philox mode: result must not depend on the number of ranks -> compare with a 1-rank run on rank 0
    model.gen_rng = "philox"
    model.gen_philox_seed = 7
    wav_p = model.generate(mel, None, True, 2750, 275, False)
    # generate_many: the folds of four utterances sharded as one job == one sharded generate() per utterance
    model.gen_rng = "torch"
    mels = [helpers.make_mel(T, seed) for T, seed in ((30, 0), (26, 3), (41, 5), (22, 7))]
    torch.manual_seed(99)
    seq = [model.generate(m, None, True, 2750, 275, False) for m in mels]
    torch.manual_seed(99)
    many = model.generate_many(mels, [None] * 4, 2750, 275, False)
    many_ok = model.gen_stats["world"] == world and all(a.shape == b.shape and np.abs(a - b).max() <= 1e-6 for a, b in zip(seq, many))
    ok = err <= 2e-2 and np.isfinite(wav_p).all() and many_ok
    flag = torch.tensor([1.0 if ok else 0.0, err], device=f"cuda:{local}")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    gathered = [torch.zeros(wav_p.shape[0], dtype=torch.float64, device=f"cuda:{local}") for _ in range(world)]
    dist.all_gather(gathered, torch.from_numpy(wav_p).to(f"cuda:{local}"))
    same = all(torch.equal(gathered[0], x) for x in gathered)
    if rank == 0:
        print(f"multi-gpu check: world={world} engine={model.gen_stats.get('engine')} max|wav-ref|={err:.3e} "
               f"generate_many==generate: {many_ok} philox identical on all ranks={same} -> {'OK' if flag[0].item() == 1.0 and same else 'FAIL'}")
    dist.destroy_process_group()
    sys.exit(0 if (flag[0].item() == 1.0 and same) else 1)


if __name__ == "__main__":
    main()

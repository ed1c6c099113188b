"""Subprocess harness of tests/test_dropin_cli.py -- TEST INFRASTRUCTURE ONLY (container, CPU).

Executes an UNMODIFIED reference CLI (gen_wavernn.py / gen_tacotron.py) with wavernn_b200's WaveRNN underneath, on a
box without a GPU: the only thing replaced is the CUDA engine behind the C ABI (`cabi.Engine` -> a CPU stand-in that
reads the same raw pointers the library would receive and runs the numpy oracle), exactly as tests/test_sharding.py
does for the multi-rank host logic.  Everything else -- the reference script, its hparams / paths / text front end /
Tacotron, our WaveRNN ctor, load(), generate() prologue, conditioning, RNG replay, epilogue, wav write -- is the real
code.  Records the drop-in surface the script exercised (ctor kwargs, load path, generate arguments) as JSON.

    python tests/dropin_harness.py <launcher|swap> <record.json> <seed> <script.py> [script args...]
"""
import ctypes
import json
import runpy
import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from oracle import ref_shim  # noqa: E402
from oracle import wavernn_oracle as O  # noqa: E402


def _view(ptr, shape, dtype=np.float32):
    n = int(np.prod(shape))
    ct = {np.float32: ctypes.c_float, np.int64: ctypes.c_int64}[dtype]
    return np.ctypeslib.as_array((ct * n).from_address(ptr)).reshape(shape)


class StandInEngine:
    """Same constructor and methods as wavernn_b200.cabi.Engine; the loop runs in the numpy oracle."""
    name, grid_ctas, launch_count = "cpu-stand-in(oracle)", 0, 0

    def __init__(self, weights, *, rnn_dims=512, fc_dims=512, feat_dims=80, aux_dims=32, n_classes=30, mode="MOL",
                 precision="fp16", engine="auto", device=0):
        self.w = O.hot_weights({k: v for k, v in weights.items()})
        self.mode, self.n_classes = mode, n_classes
        RECORD["engine_ctor"] = dict(rnn_dims=rnn_dims, fc_dims=fc_dims, feat_dims=feat_dims, aux_dims=aux_dims,
                                     n_classes=n_classes, mode=mode, precision=precision, engine=engine)

    def generate(self, *, mels_up, aux, L, n_seg, seg_len, seg_stride, out, seg_first=0, steps=0, uniforms=0, **kw):
        m, a = _view(mels_up, (L, 80)), _view(aux, (L, 128))
        u = _view(uniforms, (seg_len, 11 * n_seg)).copy()
        _view(out, (n_seg, seg_len))[:] = O.generate_segments(self.w, m, a, n_seg=n_seg, seg_len=seg_len,
                                                              seg_stride=seg_stride, uniforms=u)
        RECORD.setdefault("engine_jobs", []).append(dict(L=int(L), n_seg=n_seg, seg_len=seg_len, seg_stride=int(seg_stride)))

    def check(self):
        pass

    def close(self):
        pass


RECORD = {}


def main():
    mode, record_path, seed, script = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4]
    args = sys.argv[5:]
    ref_shim.install_import_stubs()
    ref_shim.install_text_stubs()
    from wavernn_b200 import cabi, vocoder
    cabi.Engine = StandInEngine
    W = vocoder.WaveRNN
    W._require_cuda = lambda self: torch.device("cpu")
    torch.cuda.current_stream = lambda device=None: types.SimpleNamespace(cuda_stream=0, synchronize=lambda: None)
    init, load, generate = W.__init__, W.load, W.generate

    def rec_init(self, *a, **kw):
        RECORD.setdefault("ctor", []).append({k: (list(v) if isinstance(v, tuple) else v) for k, v in kw.items()} | {"n_positional": len(a)})
        init(self, *a, **kw)
        self.gen_conditioning, self.gen_epilogue, self.gen_native_rng = "torch", "host", False   # CPU tensors only

    def rec_load(self, path):
        RECORD.setdefault("load", []).append(str(path))
        return load(self, path)

    def rec_generate(self, mels, save_path, batched, target, overlap, mu_law):
        RECORD.setdefault("generate", []).append(dict(mel_shape=list(np.asarray(mels).shape), mel_type=type(mels).__name__,
                                                      save_path=str(save_path), batched=bool(batched), target=int(target),
                                                      overlap=int(overlap), mu_law=bool(mu_law), step=self.get_step(),
                                                      training_before=bool(self.training)))
        np.save(str(save_path) + ".mel.npy", np.asarray(mels))
        torch.manual_seed(seed + len(RECORD["generate"]) - 1)       # known generator state at the call
        wav = generate(self, mels, save_path, batched, target, overlap, mu_law)
        RECORD["generate"][-1].update(wav_len=int(len(wav)), wav_dtype=str(wav.dtype), training_after=bool(self.training))
        return wav

    W.__init__, W.load, W.generate = rec_init, rec_load, rec_generate
    try:
        if mode == "launcher":
            from wavernn_b200 import dropin
            dropin.main([script] + args)
        else:                                                        # the checkout already carries the one-file swap
            sys.path.insert(0, str(Path(script).resolve().parent))
            sys.argv = [script] + args
            runpy.run_path(script, run_name="__main__")
        import models.fatchord_version as mfv
        RECORD["module_file"] = str(getattr(mfv, "__file__", ""))
        RECORD["waveRNN_is_ours"] = mfv.WaveRNN is W
    finally:
        Path(record_path).write_text(json.dumps(RECORD, indent=1))


if __name__ == "__main__":
    main()

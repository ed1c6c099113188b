"""ctypes binding of include/wavernn_b200.h -- the only way the Python host reaches the
CUDA engine.  No torch types cross this boundary: raw pointers (tensor.data_ptr())
and a raw cudaStream_t.  Import never fails (so CPU-only tools can load the package)
but `load()` raises if the shared library has not been built: there is no fallback.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

LIB_NAME = "libwavernn_b200.so"
LIB_PATH = Path(__file__).with_name(LIB_NAME)

WRNN_OK, WRNN_E_INVALID, WRNN_E_CUDA, WRNN_E_NO_DEVICE, WRNN_E_WATCHDOG, WRNN_E_BUSY = 0, -1, -2, -3, -4, -5
MODE_MOL, MODE_RAW = 0, 1
PREC_F16, PREC_FP32, PREC_BF16 = 0, 1, 2
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TCGEN05, ENGINE_STREAM = 0, 1, 2, 3
COND_AUTO, COND_EXPAND, COND_IN_KERNEL = 0, 1, 2
ABI_VERSION = 5

EXPORTS = ("wrnn_abi_version", "wrnn_last_error", "wrnn_create", "wrnn_destroy", "wrnn_generate",
           "wrnn_check", "wrnn_generate_host", "wrnn_engine_name", "wrnn_grid_ctas", "wrnn_launch_count",
           "wrnn_mt19937_uniform", "wrnn_mt19937_uniform_cols", "wrnn_epilogue", "wrnn_expand_conditioning")

_fp = C.POINTER(C.c_float)


class WrnnCfg(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("rnn_dims", "fc_dims", "feat_dims", "aux_dims", "n_classes",
                                         "mode", "precision", "engine")]


WEIGHT_FIELDS = ("I_weight", "I_bias",
                 "rnn1_weight_ih", "rnn1_weight_hh", "rnn1_bias_ih", "rnn1_bias_hh",
                 "rnn2_weight_ih", "rnn2_weight_hh", "rnn2_bias_ih", "rnn2_bias_hh",
                 "fc1_weight", "fc1_bias", "fc2_weight", "fc2_bias", "fc3_weight", "fc3_bias")
# state_dict key for each field (reference models/fatchord_version.py:115-123)
WEIGHT_KEYS = ("I.weight", "I.bias",
               "rnn1.weight_ih_l0", "rnn1.weight_hh_l0", "rnn1.bias_ih_l0", "rnn1.bias_hh_l0",
               "rnn2.weight_ih_l0", "rnn2.weight_hh_l0", "rnn2.bias_ih_l0", "rnn2.bias_hh_l0",
               "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias", "fc3.weight", "fc3.bias")


class WrnnWeights(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in WEIGHT_FIELDS]


class WrnnJob(C.Structure):
    _fields_ = [("mels_up", C.c_void_p), ("aux", C.c_void_p), ("L", C.c_int64), ("seg_stride", C.c_int64),
                ("n_seg", C.c_int32), ("seg_len", C.c_int32), ("seg_first", C.c_int32), ("steps", C.c_int32),
                ("uniforms", C.c_void_p), ("expo", C.c_void_p),
                ("philox_seed", C.c_uint64), ("philox_offset", C.c_uint64),
                ("out", C.c_void_p), ("x_force", C.c_void_p), ("logits_out", C.c_void_p),
                ("fold_row0", C.c_void_p), ("fold_row_end", C.c_void_p),
                ("mel_frames", C.c_void_p), ("aux_frames", C.c_void_p), ("up_taps", C.c_void_p),
                ("hop", C.c_int32), ("cond_mode", C.c_int32), ("uniforms_ready", C.c_void_p)]


_lib = None


def is_built() -> bool:
    return LIB_PATH.is_file()


def load() -> C.CDLL:
    """dlopen the engine.  Raises RuntimeError if it is missing (no CPU fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("WAVERNN_B200_LIB", str(LIB_PATH))
    if not os.path.isfile(path):
        raise RuntimeError(
            f"wavernn_b200: CUDA engine {path} is not built. Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` (or `make -C wavernn_b200/csrc`). There is no CPU fallback for generate().")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    lib.wrnn_abi_version.restype = C.c_int
    lib.wrnn_last_error.restype = C.c_char_p
    lib.wrnn_create.restype = C.c_int
    lib.wrnn_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(WrnnCfg), C.POINTER(WrnnWeights), C.c_int]
    lib.wrnn_destroy.restype = None
    lib.wrnn_destroy.argtypes = [C.c_void_p]
    lib.wrnn_generate.restype = C.c_int
    lib.wrnn_generate.argtypes = [C.c_void_p, C.POINTER(WrnnJob), C.c_void_p]
    lib.wrnn_check.restype = C.c_int
    lib.wrnn_check.argtypes = [C.c_void_p]
    lib.wrnn_generate_host.restype = C.c_int
    lib.wrnn_generate_host.argtypes = [C.c_void_p, C.POINTER(WrnnJob)]
    lib.wrnn_engine_name.restype = C.c_char_p
    lib.wrnn_engine_name.argtypes = [C.c_void_p]
    lib.wrnn_grid_ctas.restype = C.c_int
    lib.wrnn_grid_ctas.argtypes = [C.c_void_p]
    lib.wrnn_launch_count.restype = C.c_int64
    lib.wrnn_launch_count.argtypes = [C.c_void_p]
    lib.wrnn_expand_conditioning.restype = C.c_int
    lib.wrnn_expand_conditioning.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int64, C.c_int64,
                                             C.c_void_p, C.c_void_p, C.c_void_p]
    lib.wrnn_epilogue.restype = C.c_int
    lib.wrnn_epilogue.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_int32, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p]
    lib.wrnn_mt19937_uniform.restype = C.c_int32
    lib.wrnn_mt19937_uniform.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_void_p, C.c_uint64, C.c_float, C.c_float]
    lib.wrnn_mt19937_uniform_cols.restype = C.c_int32
    lib.wrnn_mt19937_uniform_cols.argtypes = [C.c_void_p, C.c_int32, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint64,
                                              C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_float, C.c_float]
    if lib.wrnn_abi_version() != ABI_VERSION:
        raise RuntimeError(f"wavernn_b200: ABI mismatch (lib {lib.wrnn_abi_version()} != binding {ABI_VERSION})")
    _lib = lib
    return lib


class EngineError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"wavernn_b200 engine error {code}: {msg}")
        self.code = code


def _check(lib, rc: int):
    if rc != 0:
        raise EngineError(rc, (lib.wrnn_last_error() or b"").decode(errors="replace"))


class Engine:
    """Owns one `wrnn_t*`.  `weights` maps WEIGHT_KEYS -> object with .data_ptr()
    (torch tensors, fp32, contiguous, host or device) or raw ints."""

    def __init__(self, weights: dict, *, rnn_dims=512, fc_dims=512, feat_dims=80, aux_dims=32, n_classes=30,
                 mode="MOL", precision="fp16", engine="auto", device: int = 0):
        self._h = C.c_void_p()
        self.lib = load()
        cfg = WrnnCfg(rnn_dims, fc_dims, feat_dims, aux_dims, n_classes,
                      {"MOL": MODE_MOL, "RAW": MODE_RAW}[mode],
                      {"fp16": PREC_F16, "bf16": PREC_BF16, "fp32": PREC_FP32}[precision],
                      {"auto": ENGINE_AUTO, "simt": ENGINE_SIMT, "tcgen05": ENGINE_TCGEN05, "stream": ENGINE_STREAM}[engine])
        w = WrnnWeights()
        keep = []
        for field, key in zip(WEIGHT_FIELDS, WEIGHT_KEYS):
            t = weights[key]
            if hasattr(t, "data_ptr"):
                t = t.detach()
                if t.dtype is not _torch().float32 or not t.is_contiguous():
                    t = t.float().contiguous()
                keep.append(t)
                setattr(w, field, t.data_ptr())
            else:
                setattr(w, field, int(t))
        _check(self.lib, self.lib.wrnn_create(C.byref(self._h), C.byref(cfg), C.byref(w), int(device)))
        del keep
        self.n_classes, self.mode, self.precision, self.device = n_classes, mode, precision, device

    @property
    def name(self) -> str:
        return self.lib.wrnn_engine_name(self._h).decode()

    @property
    def grid_ctas(self) -> int:
        return self.lib.wrnn_grid_ctas(self._h)

    @property
    def launch_count(self) -> int:
        return self.lib.wrnn_launch_count(self._h)

    def generate(self, *, mels_up: int, aux: int, L: int, n_seg: int, seg_len: int, seg_stride: int, out: int,
                 seg_first: int = 0, steps: int = 0, uniforms: int = 0, expo: int = 0, philox_seed: int = 0,
                 philox_offset: int = 0, x_force: int = 0, logits_out: int = 0, fold_row0: int = 0, fold_row_end: int = 0,
                 mel_frames: int = 0, aux_frames: int = 0, up_taps: int = 0, hop: int = 0, cond_mode: int = 0,
                 uniforms_ready: int = 0, stream: int = 0):
        """All buffer arguments are raw device addresses (ints).  Asynchronous."""
        job = WrnnJob(mels_up, aux, L, seg_stride, n_seg, seg_len, seg_first, steps, uniforms or None,
                      expo or None, philox_seed, philox_offset, out, x_force or None, logits_out or None,
                      fold_row0 or None, fold_row_end or None, mel_frames or None, aux_frames or None, up_taps or None,
                      hop, cond_mode, uniforms_ready or None)
        _check(self.lib, self.lib.wrnn_generate(self._h, C.byref(job), C.c_void_p(stream or None)))

    def generate_host(self, *, mels_up, aux, n_seg: int, seg_len: int, seg_stride: int, seg_first: int = 0, steps: int = 0,
                      uniforms=None, expo=None, philox_seed: int = 0, philox_offset: int = 0, x_force=None,
                      want_logits: bool = False):
        """wrnn_generate_host: the same job with HOST buffers (numpy float32 arrays) -- the entry a non-torch caller
        binds.  Copies in, runs, copies out, synchronises and checks.  Returns samples (n_seg, S)[, logits (S, n_seg, C)]."""
        import numpy as np
        f32 = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float32)
        mels_up, aux, uniforms, expo, x_force = f32(mels_up), f32(aux), f32(uniforms), f32(expo), f32(x_force)
        S = steps or seg_len
        out = np.empty((n_seg, S), dtype=np.float32)
        logits = np.empty((S, n_seg, self.n_classes), dtype=np.float32) if want_logits else None
        ptr = lambda a: None if a is None else a.ctypes.data
        job = WrnnJob(ptr(mels_up), ptr(aux), mels_up.shape[0], seg_stride, n_seg, seg_len, seg_first, steps, ptr(uniforms),
                      ptr(expo), philox_seed, philox_offset, ptr(out), ptr(x_force), ptr(logits), None, None, None, None,
                      None, 0, 0, None)
        _check(self.lib, self.lib.wrnn_generate_host(self._h, C.byref(job)))
        return (out, logits) if want_logits else out

    def check(self):
        _check(self.lib, self.lib.wrnn_check(self._h))

    def close(self):
        if self._h:
            self.lib.wrnn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _torch():
    import torch
    return torch


def expand_conditioning(*, mel_frames: int, aux_frames: int, up_taps: int, hop: int, row_lo: int, n_rows: int,
                        mels_up: int, aux: int, stream: int = 0):
    """wrnn_expand_conditioning with raw device addresses.  Asynchronous on `stream`."""
    lib = load()
    _check(lib, lib.wrnn_expand_conditioning(mel_frames, aux_frames, up_taps, hop, row_lo, n_rows, mels_up, aux, stream or None))


def epilogue(*, samples: int, n_seg: int, seg_len: int, seg_stride: int, overlap: int, fade_in: int, fade_out: int,
             mu_table: int, n_classes: int, tail: int, tail_len: int, wave_len: int, wav: int, stream: int = 0):
    """wrnn_epilogue with raw device addresses (ints; 0 = NULL).  Asynchronous on `stream`."""
    lib = load()
    _check(lib, lib.wrnn_epilogue(samples, n_seg, seg_len, seg_stride, overlap, fade_in or None, fade_out or None,
                                  mu_table or None, n_classes, tail or None, tail_len, wave_len, wav, stream or None))


# ---------------------------------------------------------------------------------------------
# torch CPU generator replay (parity mode's host-side cost; see csrc/wrnn_hostrng.cu)
# ---------------------------------------------------------------------------------------------
# Layout of torch.get_rng_state() for the CPU generator (CPUGeneratorImplStateLegacy): seed u64 @0, left i32 @8,
# seeded i32 @12, next u64 @16, 624 state words as u64 @24, then the cached-normal fields (untouched here).
_MT_N, _MT_OFF, _STATE_BYTES = 624, 24, 5056


def torch_rng_uniform(skip: int, n: int, lo: float, hi: float, out=None, generator=None, row_len: int = 0, cols=None):
    """Advance torch's CPU generator by `skip` discarded 32-bit outputs and then fill `out` (a contiguous CPU float32
    tensor with n elements, e.g. pinned) with what `torch.empty(n).uniform_(lo, hi)` would have produced -- natively.
    The generator (default: the global one) is left in exactly the state torch would have left it in.
    With `row_len` and `cols=((a_lo, a_hi), (b_lo, b_hi))` the n draws are read as an (n // row_len, row_len) matrix of
    which only those two column ranges are kept (`out`: rows x (a_hi-a_lo + b_hi-b_lo)); the rest is skipped unconverted."""
    import numpy as np
    import torch
    lib = load()
    st = (generator.get_state() if generator is not None else torch.get_rng_state()).clone()
    a = st.numpy()
    if a.size != _STATE_BYTES:
        raise RuntimeError("unexpected torch CPU generator state size")
    left = int(a[8:12].view(np.int32)[0])
    if not (1 <= left <= _MT_N):
        raise RuntimeError("unexpected torch CPU generator state (left)")
    words = a[_MT_OFF:_MT_OFF + 8 * _MT_N].view(np.uint64)
    mt = np.ascontiguousarray(words.astype(np.uint32))
    if cols is not None:
        (a_lo, a_hi), (b_lo, b_hi) = cols
        rows = n // row_len
        assert row_len > 0 and rows * row_len == n
        n_out = rows * ((a_hi - a_lo) + (b_hi - b_lo))
    else:
        n_out = n
    if out is None:
        out = torch.empty(n_out, dtype=torch.float32)
    assert out.dtype == torch.float32 and out.is_contiguous() and out.numel() >= n_out and out.device.type == 'cpu'
    if cols is not None:
        pos = lib.wrnn_mt19937_uniform_cols(mt.ctypes.data, _MT_N + 1 - left, int(skip), out.data_ptr(), rows, int(row_len),
                                            int(a_lo), int(a_hi), int(b_lo), int(b_hi), float(lo), float(hi))
    else:
        pos = lib.wrnn_mt19937_uniform(mt.ctypes.data, _MT_N + 1 - left, int(skip), out.data_ptr() if n else None, int(n),
                                       float(lo), float(hi))
    if pos < 0:
        raise RuntimeError("wrnn_mt19937_uniform rejected its arguments")
    words[:] = mt
    a[8:12].view(np.int32)[0] = _MT_N + 1 - pos
    a[16:24].view(np.uint64)[0] = pos
    if generator is not None:
        generator.set_state(st)
    else:
        torch.set_rng_state(st)
    return out


_rng_replay_ok = None


def torch_rng_replay_ok() -> bool:
    """One-time self-check of the native replay against torch itself (private generators, a skip that crosses
    blocks, the reference's uniform range).  False -> callers use torch's own operators (slower, same numbers)."""
    global _rng_replay_ok
    if _rng_replay_ok is None:
        import torch
        try:
            g1, g2 = torch.Generator(), torch.Generator()
            g1.manual_seed(20240607); g2.manual_seed(20240607)
            torch.empty(1531).uniform_(-0.04, 0.04, generator=g1)
            want = torch.empty(2000).uniform_(1e-5, 1.0 - 1e-5, generator=g1)
            got = torch_rng_uniform(1531, 2000, 1e-5, 1.0 - 1e-5, generator=g2)
            tail_w = torch.empty(700).uniform_(generator=g1)
            tail_g = torch.empty(700).uniform_(generator=g2)
            ok = torch.equal(want, got) and torch.equal(tail_w, tail_g)
            want2 = torch.empty(40, 55).uniform_(1e-5, 1.0 - 1e-5, generator=g1)
            got2 = torch_rng_uniform(0, 40 * 55, 1e-5, 1.0 - 1e-5, generator=g2, row_len=55, cols=((10, 30), (52, 54)))
            ok = ok and torch.equal(torch.cat([want2[:, 10:30], want2[:, 52:54]], 1).reshape(-1), got2)
            ok = ok and torch.equal(torch.empty(5).uniform_(generator=g1), torch.empty(5).uniform_(generator=g2))
            _rng_replay_ok = bool(ok)
        except Exception:
            _rng_replay_ok = False
    return _rng_replay_ok

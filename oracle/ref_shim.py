"""Import the UNMODIFIED reference (fatchord/WaveRNN) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (wavernn_b200/) may
import this file.  It only works in the build container, where the reference
checkout is mounted read-only at /root/reference; the GPU box has no such
directory, so everything here is used solely to (a) pin `oracle/wavernn_oracle.py`
against the real reference and (b) generate the committed fixtures under
`tests/golden/` (see `tests/golden/make_golden.py`).

The reference star-imports matplotlib and librosa (models/fatchord_version.py:5-6
-> utils/display.py:1-3, utils/dsp.py:3) which are absent here, and uses
`np.cumproduct` (fatchord_version.py:68) which NumPy 2 removed.  Three in-process
shims make it importable without touching its files.

Because the reference's top-level packages are called `models` and `utils`,
always run this in a subprocess / fresh interpreter that does not have a
same-named package of ours on sys.path.
"""
import io
import os
import sys
import types
import zipfile

import numpy as np

REF_ROOT = os.environ.get("WAVERNN_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "models", "fatchord_version.py"))


_ref_mod = None
_captured_wavs = []


def install_import_stubs():
    """The three in-process shims that make the reference importable here: fake `librosa`
    (only `librosa.output.write_wav` is reached on this path, utils/dsp.py:22-23) and `matplotlib`, and the
    `np.cumproduct` alias NumPy 2 dropped (fatchord_version.py:68)."""
    if "librosa" not in sys.modules:
        lib = types.ModuleType("librosa")
        lib.output = types.SimpleNamespace(
            write_wav=lambda path, x, sr: _captured_wavs.append((str(path), np.array(x), sr)))
        sys.modules["librosa"] = lib
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        mpl.interactive = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules.update({"matplotlib": mpl, "matplotlib.pyplot": plt})
    if not hasattr(np, "cumproduct"):
        np.cumproduct = np.cumprod


def load_reference():
    """Returns the reference module `models.fatchord_version` (cached)."""
    global _ref_mod
    if _ref_mod is not None:
        return _ref_mod
    if not available():
        raise RuntimeError(f"reference checkout not found at {REF_ROOT}")
    sys.path.insert(0, REF_ROOT)
    install_import_stubs()
    import models.fatchord_version as ref  # noqa: E402  (unmodified reference)
    from utils import hparams as hp  # noqa: E402
    if not hp.is_configured():
        hp.configure(os.path.join(REF_ROOT, "hparams.py"))
    ref.stream = lambda msg: None  # silence gen_display
    _ref_mod = ref
    return ref


def default_kwargs(mode="MOL", bits=9):
    """ctor kwargs as gen_wavernn.py:112-123 builds them from hparams.py:20-60."""
    return dict(rnn_dims=512, fc_dims=512, bits=bits, pad=2, upsample_factors=(5, 5, 11),
                feat_dims=80, compute_dims=128, res_out_dims=128, res_blocks=10,
                hop_length=275, sample_rate=22050, mode=mode)


def build_reference_model(seed=0, mode="MOL", bits=9, pretrained=False):
    import contextlib
    import torch
    ref = load_reference()
    torch.manual_seed(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        model = ref.WaveRNN(**default_kwargs(mode, bits))
    if pretrained:
        sd = load_pretrained_state_dict()
        model.load_state_dict(sd, strict=False)
    return model


def load_pretrained_state_dict():
    import torch
    zpath = os.path.join(REF_ROOT, "pretrained", "ljspeech.wavernn.mol.800k.zip")
    with zipfile.ZipFile(zpath) as z:
        blob = z.read("latest_weights.pyt")
    return torch.load(io.BytesIO(blob), map_location="cpu")


def ref_generate(model, mels, batched, target, overlap, mu_law=False, seed=1234,
                 capture=True):
    """Runs the reference's own generate() on CPU under torch.manual_seed(seed).

    Returns dict(wav=final float64 waveform, pre=(B,S) float64 pre-xfade samples
    (post mu-law decode, as handed to xfade_and_unfold, fatchord_version.py:250-251),
    raw=(B,S) float32 samples exactly as stacked at :243).
    """
    import torch
    ref = load_reference()
    grabbed = {}
    if capture:
        orig_xfade = model.xfade_and_unfold

        def spy(y, target, overlap):
            grabbed["pre"] = y.copy()
            return orig_xfade(y, target, overlap)
        model.xfade_and_unfold = spy
        orig_stack = torch.stack

        def stack_spy(tensors, *a, **k):
            out = orig_stack(tensors, *a, **k)
            grabbed["raw"] = out.transpose(0, 1).detach().cpu().numpy().copy()
            return out
        ref.torch.stack = stack_spy
    try:
        torch.manual_seed(seed)
        wav = model.generate(mels, "/dev/null.wav", batched, target, overlap, mu_law)
    finally:
        if capture:
            del model.xfade_and_unfold
            ref.torch.stack = orig_stack
    model.eval()
    out = dict(wav=np.asarray(wav), raw=grabbed.get("raw"))
    out["pre"] = grabbed.get("pre")
    return out


# --------------------------------------------------------------------------------------
# Tacotron front half of gen_tacotron.py (BASELINE configs[3]); container-only, fixtures only
# --------------------------------------------------------------------------------------
_ONES = ("zero one two three four five six seven eight nine ten eleven twelve thirteen fourteen fifteen sixteen "
         "seventeen eighteen nineteen").split()
_TENS = "zero ten twenty thirty forty fifty sixty seventy eighty ninety".split()


def _number_to_words(n, andword="", zero="zero", group=0):
    """Minimal stand-in for inflect.engine().number_to_words (utils/text/numbers.py:3,7): cardinals below one
    million, which is all the fixture sentences contain."""
    n = int(str(n).replace(",", "").split(".")[0] or 0)
    if n < 20:
        return _ONES[n]
    if n < 100:
        return _TENS[n // 10] + ("-" + _ONES[n % 10] if n % 10 else "")
    if n < 1000:
        return _ONES[n // 100] + " hundred" + (" " + _number_to_words(n % 100) if n % 100 else "")
    return _number_to_words(n // 1000) + " thousand" + (" " + _number_to_words(n % 1000) if n % 1000 else "")


def install_text_stubs():
    """`unidecode` and `inflect` are absent here: ASCII pass-through and a small cardinal speller."""
    if "unidecode" not in sys.modules:
        uni = types.ModuleType("unidecode")
        uni.unidecode = lambda s: s.encode("ascii", "ignore").decode("ascii")
        sys.modules["unidecode"] = uni
    if "inflect" not in sys.modules:
        inf = types.ModuleType("inflect")
        inf.engine = lambda: types.SimpleNamespace(number_to_words=_number_to_words)
        sys.modules["inflect"] = inf


def build_reference_tacotron():
    """The reference's Tacotron with the shipped checkpoint, exactly as gen_tacotron.py:94-111 builds and loads it
    (`.load()`, so the legacy `r` key is honoured, tacotron.py:452-454)."""
    import contextlib
    import tempfile
    import torch
    load_reference()
    install_text_stubs()
    from models.tacotron import Tacotron
    from utils import hparams as hp
    from utils.text.symbols import symbols
    with contextlib.redirect_stdout(io.StringIO()):
        tts = Tacotron(embed_dims=hp.tts_embed_dims, num_chars=len(symbols), encoder_dims=hp.tts_encoder_dims,
                       decoder_dims=hp.tts_decoder_dims, n_mels=hp.num_mels, fft_bins=hp.num_mels,
                       postnet_dims=hp.tts_postnet_dims, encoder_K=hp.tts_encoder_K, lstm_dims=hp.tts_lstm_dims,
                       postnet_K=hp.tts_postnet_K, num_highways=hp.tts_num_highways, dropout=hp.tts_dropout,
                       stop_threshold=hp.tts_stop_threshold)
    zpath = os.path.join(REF_ROOT, "pretrained", "ljspeech.tacotron.r2.180k.zip")
    with zipfile.ZipFile(zpath) as z, tempfile.TemporaryDirectory() as d:
        z.extract("latest_weights.pyt", d)
        tts.load(os.path.join(d, "latest_weights.pyt"))
    return tts


def tacotron_mels(sentences, seed=0):
    """gen_tacotron.py:113-143: text -> ids -> Tacotron.generate -> (m + 4) / 8 clipped to [0, 1]."""
    import torch
    tts = build_reference_tacotron()
    from utils import hparams as hp
    from utils.text import text_to_sequence
    out = []
    for i, s in enumerate(sentences):
        torch.manual_seed(seed + i)
        _, m, _ = tts.generate(text_to_sequence(s.strip(), hp.tts_cleaner_names))
        m = (m + 4) / 8
        np.clip(m, 0, 1, out=m)
        out.append(m.astype(np.float32))
    return out

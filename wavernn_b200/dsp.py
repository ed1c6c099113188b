"""Host-side DSP helpers on the generate() epilogue (reference utils/dsp.py:8-23,
92-103).  librosa is not a dependency: wavs are written with scipy."""
import math

import numpy as np

from .hp import hparams as hp


def label_2_float(x, bits):
    return 2 * x / (2 ** bits - 1.) - 1.


def float_2_label(x, bits):
    assert abs(x).max() <= 1.0
    x = (x + 1.) * (2 ** bits - 1) / 2
    return x.clip(0, 2 ** bits - 1)


def encode_mu_law(x, mu):
    mu = mu - 1
    fx = np.sign(x) * np.log(1 + mu * np.abs(x)) / np.log(1 + mu)
    return np.floor((fx + 1) / 2 * mu + 0.5)


def decode_mu_law(y, mu, from_labels=True):
    """utils/dsp.py:98-103."""
    if from_labels:
        y = label_2_float(y, math.log2(mu))
    mu = mu - 1
    return np.sign(y) / mu * ((1 + mu) ** np.abs(y) - 1)


def save_wav(x, path, sample_rate=None):
    """float32 wav at hp.sample_rate (utils/dsp.py:22-23 wrote through
    librosa.output.write_wav, removed in librosa>=0.8)."""
    from scipy.io import wavfile
    sr = int(sample_rate if sample_rate is not None else hp.sample_rate)
    wavfile.write(str(path), sr, np.asarray(x).astype(np.float32))

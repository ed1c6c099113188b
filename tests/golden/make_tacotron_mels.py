"""BASELINE configs[3] fixture: the 16 mel spectrograms the reference's Tacotron (shipped checkpoint, CPU) produces
for 16 sentences -- the 6 of the reference's sentences.txt plus 10 more -- exactly as gen_tacotron.py:113-143 prepares
them for `voc_model.generate`.  Run in the build container (needs /root/reference):

    python tests/golden/make_tacotron_mels.py

The text front-end's two missing third-party modules are stubbed (oracle/ref_shim.py::install_text_stubs).  Stored as
uint16-quantised [0, 1] values (the mels are clipped to that range; 1.5e-5 resolution) to keep the fixture small; the
quantised values ARE the fixture: every consumer (reference run, oracle, CUDA path) reads the same de-quantised mels.
"""
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))

from oracle import ref_shim  # noqa: E402

EXTRA = [
    "The quick brown fox jumps over the lazy dog.",
    "A persistent kernel keeps the hidden state on the chip.",
    "She sells sea shells by the sea shore.",
    "How much wood would a woodchuck chuck?",
    "Speech synthesis turns text into audio one sample at a time.",
    "The weather tomorrow will be sunny with a light breeze from the west.",
    "Please remember to call your mother on Sunday.",
    "It was the best of times, it was the worst of times.",
    "To be, or not to be, that is the question.",
    "Thank you for listening.",
]


def main():
    with open(Path(ref_shim.REF_ROOT) / "sentences.txt") as f:
        sentences = [l.strip() for l in f if l.strip()] + EXTRA
    assert len(sentences) == 16
    mels = ref_shim.tacotron_mels(sentences, seed=0)
    arrays = {f"mel_{i:02d}": np.round(m * 65535.0).astype(np.uint16) for i, m in enumerate(mels)}
    np.savez_compressed(HERE / "tacotron_mels.npz", sentences=np.array(sentences), **arrays)
    print("frames:", [m.shape[1] for m in mels], "total", sum(m.shape[1] for m in mels))


if __name__ == "__main__":
    main()

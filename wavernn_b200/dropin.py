"""Run an UNMODIFIED reference CLI (gen_wavernn.py, gen_tacotron.py, quick_start.py, train_wavernn.py) with the B200
vocoder underneath:

    cd /path/to/WaveRNN && python -m wavernn_b200.dropin gen_wavernn.py --file mel.npy --batched

The scripts reach the vocoder only through `from models.fatchord_version import WaveRNN` (gen_wavernn.py:3,
gen_tacotron.py:2, quick_start.py:2, train_wavernn.py:8).  `install()` registers a module of that name -- the same
re-export as dropin/models/fatchord_version.py -- in sys.modules before the script is executed, so the import resolves
to wavernn_b200.vocoder while `models.tacotron`, `utils.*`, `hparams.py` and the checkpoints stay the reference's own.
Nothing in the reference checkout is modified.
"""
from __future__ import annotations

import os
import runpy
import sys
import types

SHIM_NAME = "models.fatchord_version"
EXPORTS = ("WaveRNN", "UpsampleNetwork", "MelResNet", "ResBlock", "Stretch2d")


def install() -> types.ModuleType:
    """Registers the replacement `models.fatchord_version` (idempotent).  Returns the module."""
    from . import vocoder
    mod = sys.modules.get(SHIM_NAME)
    if mod is not None and getattr(mod, "__wavernn_b200__", False):
        return mod
    mod = types.ModuleType(SHIM_NAME, "wavernn_b200 drop-in for the reference's models/fatchord_version.py")
    for name in EXPORTS:
        setattr(mod, name, getattr(vocoder, name))
    mod.__wavernn_b200__ = True
    mod.__file__ = vocoder.__file__
    sys.modules[SHIM_NAME] = mod
    return mod


def main(argv=None) -> int:
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv or argv[0] in ("-h", "--help"):
        print(__doc__)
        return 0 if argv else 2
    script = os.path.abspath(argv[0])
    if not os.path.isfile(script):
        print(f"wavernn_b200.dropin: no such script: {argv[0]}", file=sys.stderr)
        return 2
    script_dir = os.path.dirname(script)
    if script_dir in sys.path:
        sys.path.remove(script_dir)
    sys.path.insert(0, script_dir)          # what `python script.py` does: the script's packages (models/, utils/) win
    install()
    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())

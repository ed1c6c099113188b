"""Process-global hyper-parameter singleton, mirroring `utils/__init__.py:40-104`
of the reference: `hparams.configure(path)` copies the module-level names of a
python file onto the singleton exactly once; touching it before that raises
AttributeError, configuring twice raises RuntimeError (reference :51-61)."""
import re
from importlib.util import module_from_spec, spec_from_file_location
from pathlib import Path
from typing import Union


def _load_module(name: str, path: Path):
    if not Path(path).exists():
        raise FileNotFoundError('"%s" doesn\'t exist!' % path)
    spec = spec_from_file_location(name, path)
    if spec is None:
        raise ValueError('could not load module from "%s"' % path)
    mod = module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _HParams:
    def __init__(self, path: Union[str, Path, None] = None):
        self._configured = False
        if path is not None:
            self.configure(path)

    def __getattr__(self, item):
        # only reached when normal lookup fails
        if item.startswith("_"):
            raise AttributeError(item)
        if not self.__dict__.get("_configured", False):
            raise AttributeError("HParams not configured yet. Call self.configure()")
        raise AttributeError(f"hparams has no attribute {item!r}")

    def is_configured(self) -> bool:
        return self._configured

    def configure(self, path: Union[str, Path]):
        if self.is_configured():
            raise RuntimeError("Cannot reconfigure hparams!")
        path = Path(path).expanduser()
        if not path.exists():
            raise FileNotFoundError(f"Could not find hparams file {path}")
        if path.suffix != ".py":
            raise ValueError("`path` must be a python file")
        mod = _load_module("hparams", path)
        dunder = re.compile(r"^__.+__$")
        for name, value in vars(mod).items():
            if dunder.match(name):
                continue
            if name in self.__dict__:
                raise AttributeError(
                    f"module at `path` cannot contain attribute {name} as it "
                    "overwrites an attribute of the same name in utils.hparams")
            setattr(self, name, value)
        self._configured = True


hparams = _HParams()

DEFAULT_HPARAMS_FILE = str(Path(__file__).with_name("hparams_default.py"))

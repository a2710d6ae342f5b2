"""Import shim: the package directory is named after the reference repo (`qwen-image-edit-streamdiffusion_b200/`),
which is not a valid Python identifier, so it is loaded here under the importable name `qie_b200`."""
import importlib.util
import sys
from pathlib import Path

_dir = Path(__file__).resolve().parent / "qwen-image-edit-streamdiffusion_b200"
_spec = importlib.util.spec_from_file_location("qie_b200", _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["qie_b200"] = _mod
_spec.loader.exec_module(_mod)

"""Import shim: the package directory is `mlx-audio-swift_b200/` (not a valid Python identifier), so
`import mlx_audio_swift_b200` loads it from there and registers it under this name."""
import importlib.util
import sys
from pathlib import Path

_dir = Path(__file__).resolve().parent / "mlx-audio-swift_b200"
_spec = importlib.util.spec_from_file_location(
    "mlx_audio_swift_b200", _dir / "__init__.py", submodule_search_locations=[str(_dir)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mlx_audio_swift_b200"] = _mod
_spec.loader.exec_module(_mod)

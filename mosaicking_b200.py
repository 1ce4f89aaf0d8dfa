"""Import alias: `import mosaicking_b200` loads the package that lives in ./mosaic-library_b200/
(the directory name carries a hyphen, so it cannot be imported by a plain import statement)."""
import importlib.util
import sys
from pathlib import Path

_root = Path(__file__).resolve().parent / "mosaic-library_b200"
_spec = importlib.util.spec_from_file_location("mosaicking_b200", _root / "__init__.py",
                                               submodule_search_locations=[str(_root)])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["mosaicking_b200"] = _mod
_spec.loader.exec_module(_mod)

"""Import shim: the package directory is `qwen3-tts_b200/` (hyphen, as the project layout names it), which
Python cannot import by name; this module loads it under the importable name `qwen3_tts_b200`."""
import importlib.util
import os
import sys

_d = os.path.join(os.path.dirname(os.path.abspath(__file__)), "qwen3-tts_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_d, "__init__.py"),
                                               submodule_search_locations=[_d])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)

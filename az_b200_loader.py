"""Import helper: the package directory is named `alphazero-multi-game_b200` (hyphen), which Python's import
statement cannot spell.  `load()` registers it under the module name `alphazero_multi_game_b200`."""
import importlib.util
import os
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG_DIR = os.path.join(ROOT, "alphazero-multi-game_b200")
NAME = "alphazero_multi_game_b200"


def load():
    if NAME in sys.modules:
        return sys.modules[NAME]
    spec = importlib.util.spec_from_file_location(NAME, os.path.join(PKG_DIR, "__init__.py"),
                                                  submodule_search_locations=[PKG_DIR])
    mod = importlib.util.module_from_spec(spec)
    sys.modules[NAME] = mod
    spec.loader.exec_module(mod)
    return mod

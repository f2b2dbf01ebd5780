"""Importable alias of the `codim-ipc_b200/` package directory (a hyphen is not a valid Python
identifier, so the real sources live in `codim-ipc_b200/` and this module points there)."""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "codim-ipc_b200"))
from ._api import *  # noqa: F401,F403

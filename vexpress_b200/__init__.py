"""Importable alias of the package directory ``v-express_b200/`` (a hyphen cannot be imported).

``import vexpress_b200`` resolves sub-modules from ``v-express_b200/``; nothing lives here.
"""
import os as _os

__path__.insert(0, _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "v-express_b200"))
from ._version import __version__  # noqa: E402,F401

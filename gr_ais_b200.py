"""Import shim: the package directory is `gr-ais_b200/` (a name Python cannot
import directly), so this module lends it the importable name `gr_ais_b200`."""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "gr-ais_b200")]

with open(_os.path.join(__path__[0], "__init__.py")) as _fh:
    exec(compile(_fh.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _fh

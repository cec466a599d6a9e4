"""Importable alias of the package directory ``normalizing-flows-pytorch_b200/`` (hyphens are not valid in a
Python identifier).  ``import nfb200`` executes that directory's ``__init__.py`` with this module as the package,
so all submodules live exactly once, as ``nfb200.*``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), 'normalizing-flows-pytorch_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _f

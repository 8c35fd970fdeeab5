"""Importable alias of the package directory ``rvdd-release_b200/`` (a hyphen cannot be imported directly).

``import rvdd_release_b200`` executes ``rvdd-release_b200/__init__.py`` with this module's ``__path__`` pointing at
that directory, so ``rvdd_release_b200.flow_utils`` etc. resolve to the files there.
"""
import os as _os

__path__ = [_os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "rvdd-release_b200")]
with open(_os.path.join(__path__[0], "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(__path__[0], "__init__.py"), "exec"))
del _f

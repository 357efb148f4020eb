"""Import shim: the package sources live in ``baseband-tasks_b200/`` (a name
Python cannot import directly); this makes them importable as
``baseband_tasks_b200``."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(
    _os.path.abspath(__file__))), 'baseband-tasks_b200')
__path__.insert(0, _real)
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _f

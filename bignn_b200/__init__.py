"""Import alias: the package directory `bilevel-graph-neural-network_b200/` is named after
the reference repository and is not a valid Python identifier, so `import bignn_b200`
resolves its submodules there."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      'bilevel-graph-neural-network_b200')
__path__ = [_real]
with open(_os.path.join(_real, '__init__.py')) as _f:
    exec(compile(_f.read(), _os.path.join(_real, '__init__.py'), 'exec'))
del _f

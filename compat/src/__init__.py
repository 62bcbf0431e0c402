"""Shim package: the reference's `src` namespace.

`compat/` goes first on sys.path, so this file is the `src` package the reference's scripts import.
It must not hide the reference's own sub-packages (`src.models`, `src.training`, `src.utils`): the
search path is extended with every other `src/` directory on sys.path, compat first, so
`src.data.*` resolves to the B200 shims here and everything else to the reference's files
(R/train_segmented.py:8-13, R/realtime_analyzer_parallel.py:18-20).
"""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)

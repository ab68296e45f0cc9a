"""B200-native implementation of Chafa's canvas rendering hot path.

The product is the C-ABI shared library ``libchafa_b200.so`` (sources in ``csrc/``,
headers in ``/include``).  ``capi`` is a thin ctypes binding used by tests and bench.py.
"""

from . import capi  # noqa: F401

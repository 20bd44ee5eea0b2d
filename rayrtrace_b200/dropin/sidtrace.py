"""Drop-in module named ``sidtrace``.

Put this directory first on ``sys.path`` (``rayrtrace_b200.dropin_path()``) and
the reference's unmodified ``ray.py`` (``from sidtrace import trace``, ray.py:4)
runs on the B200 library instead of the C extension built by setup.py:5-12.
"""
from rayrtrace_b200.sidtrace import check, trace, trace_batch, Field, invalidate  # noqa: F401

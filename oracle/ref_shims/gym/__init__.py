"""Minimal stand-in for `gym` (absent here) so the read-only Python reference can be imported
in the build container to generate golden vectors. Test infrastructure only."""
from . import spaces  # noqa: F401


class Env:
    pass

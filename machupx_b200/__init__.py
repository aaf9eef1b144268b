"""machupx_b200 — B200-native lifting-line solve path with the MachUpX Scene API.

Same module-level exports as the reference (machupX/__init__.py:9-11)."""
from .airfoil import DatabaseBoundsError  # noqa: F401
from .exceptions import MaxIterationError, SolverNotConvergedError  # noqa: F401
from .scene import Scene  # noqa: F401
from . import helpers  # noqa: F401  (machupX.helpers: parse_input, unit and quaternion helpers)

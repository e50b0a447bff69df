"""coupe_b200 — B200-native recursive coordinate / inertial bisection behind
coupe's own interfaces (the coupe-ffi C ABI and a mirror of the `Partition`
trait).  See DESIGN.md."""
from .api import (BackendError, Context, Error, Group, InputLenMismatch, Rcb, Rib,  # noqa: F401
                  default_context)
from . import _lib  # noqa: F401
from . import tools  # noqa: F401
from .multi_jagged import Grid, MultiJagged, axis_sort  # noqa: F401

__all__ = ["Rcb", "Rib", "Context", "Group", "Error", "InputLenMismatch", "BackendError", "default_context", "tools",
           "MultiJagged", "axis_sort", "Grid"]

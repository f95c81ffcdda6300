"""`th.eb` namespace of the reference (theseus/embodied/__init__.py): the cost functions with a CUDA schema."""
from .core import Between, Difference, Local, Reprojection  # noqa: F401

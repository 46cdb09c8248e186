"""Mirror of the reference's src/criterions/__init__.py:4-8."""
from creamfl_b200.criterions import MCSoftContrastiveLoss, get_criterion  # noqa: F401

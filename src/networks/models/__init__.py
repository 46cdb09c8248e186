"""Mirror of the reference's src/networks/models/__init__.py:6-7."""
from creamfl_b200.towers import PCME, get_model  # noqa: F401

"""Mirror of the reference's src/networks/language_model.py:28-130."""
from creamfl_b200.clients import TextClient as EncoderText  # noqa: F401

"""Mirror of the reference's src/networks/resnet_client.py:220-232."""
from creamfl_b200.clients import ImageClient as ResNet, resnet18_client  # noqa: F401

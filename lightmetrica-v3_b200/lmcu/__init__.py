"""lmcu — B200-native ray-tracing hot path for Lightmetrica v3 (host side of liblmcu.so)."""
from . import abi, scenes  # noqa: F401
from .api import LmcuError, Scene, device_count  # noqa: F401

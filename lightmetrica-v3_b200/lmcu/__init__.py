"""lmcu — B200-native ray-tracing hot path for Lightmetrica v3 (host side of liblmcu.so)."""
from . import abi, scenes  # noqa: F401
from .api import (LmcuError, Scene, cache_stats, cache_trim, device_count, enable_peer_access, ipc_close, ipc_open,  # noqa: F401
                  probe_l2_bandwidth)

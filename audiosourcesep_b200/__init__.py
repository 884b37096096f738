"""B200-native Glow / NCSN / BASIS separation hot path."""
from .config import GlowConfig, NCSNConfig, get_config, dict2namespace  # noqa: F401

__version__ = "0.1.0"

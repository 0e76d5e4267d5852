"""B200-native transition-matrix engine for audio/video textures (hot path only)."""
__version__ = "0.1.0"

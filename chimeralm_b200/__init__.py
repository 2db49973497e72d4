"""chimeralm_b200: B200-native implementation of ChimeraLM's `predict` hot path."""

__version__ = "0.1.0"

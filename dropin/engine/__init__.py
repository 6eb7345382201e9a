"""`from engine import TorchEngine` (reference train.py:12) -> the B200 implementation."""
from plainlm_b200.engine import TorchEngine  # noqa: F401

__all__ = ['TorchEngine']

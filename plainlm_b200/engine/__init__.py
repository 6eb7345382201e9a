from .engine import TorchEngine

__all__ = ['TorchEngine']

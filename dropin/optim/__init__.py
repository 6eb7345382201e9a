"""`from optim import intialize_optimizer, initialize_scheduler` (reference engine.py:9) -> the B200 implementation."""
from plainlm_b200.optim import intialize_optimizer, initialize_scheduler  # noqa: F401

__all__ = ['intialize_optimizer', 'initialize_scheduler']

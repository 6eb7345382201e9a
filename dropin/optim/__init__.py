"""Shim: the reference's `engine/engine.py:9` does `from optim import intialize_optimizer, initialize_scheduler`; with
`dropin/` ahead of the reference on PYTHONPATH those two names resolve to the B200 implementation."""
import plainlm_b200.optim as _impl

intialize_optimizer = _impl.intialize_optimizer  # (sic) the reference's spelling
initialize_scheduler = _impl.initialize_scheduler
__all__ = ('initialize_scheduler', 'intialize_optimizer')

"""plainlm_b200 — B200-native (sm_100a) implementation of plainLM's data-parallel transformer training step.

Sub-packages mirror the reference's own layout for this path (`models`, `engine`, `optim`); `dropin/` at the repo
root re-exports them under the reference's top-level names so the reference's train.py imports them unchanged.
"""

__version__ = '0.1.0'

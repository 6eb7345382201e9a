"""`from torch_utils import pytorch_setup, destroy_ddp` (reference train.py:8) -> the B200 implementation."""
from plainlm_b200.torch_utils import pytorch_setup, destroy_ddp  # noqa: F401

"""`from models import construct_model` (reference train.py:11) -> the B200 implementation."""
from plainlm_b200.models import construct_model, get_param_groups  # noqa: F401
from plainlm_b200.models import transformer, components, embeddings, construct  # noqa: F401

__all__ = ['construct_model', 'get_param_groups']

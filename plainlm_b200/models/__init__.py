from .construct import construct_model, get_param_groups

__all__ = ['construct_model', 'get_param_groups']

from .init_optim import intialize_optimizer, initialize_scheduler

__all__ = ['intialize_optimizer', 'initialize_scheduler']

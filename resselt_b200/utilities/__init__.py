from . import state_dict

__all__ = ['state_dict']

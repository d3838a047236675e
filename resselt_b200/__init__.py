"""resselt_b200 — B200-native forward-inference engine with resselt's loading API.

Public surface identical to /root/reference/resselt/__init__.py:6-26.
"""
from typing import Mapping

from .archs import internal_registry

__version__ = '0.1.0'


def add(arch):
    """Register a new architecture."""
    return internal_registry.add(arch)


def get(id: str):
    """Get architecture by ID."""
    return internal_registry.get(id)


def load_from_file(path: str):
    """Detect the architecture of a checkpoint file and load it."""
    return internal_registry.load_from_file(path)


def load_from_state_dict(state_dict: Mapping[str, object]):
    """Detect the architecture of a state dict and load it."""
    return internal_registry.load_from_state_dict(state_dict)


__all__ = ['add', 'get', 'load_from_file', 'load_from_state_dict']

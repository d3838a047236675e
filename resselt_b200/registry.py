"""Architecture registry and safe checkpoint loading.

Contract: /root/reference/resselt/registry.py:14-116 — ``Registry.{add,get,load_from_file,
load_from_state_dict}``, ``ArchitectureNotFound`` and the restricted unpickler used for
``.pth``/``.ckpt`` files.  ``load_from_state_dict`` canonicalises the dict, asks every registered
architecture (insertion order) to ``detect`` it, builds the module with ``load`` and finishes with a
*strict* ``load_state_dict``.
"""
from __future__ import annotations

import os
import pickle
from types import SimpleNamespace
from typing import Dict, Iterator, Mapping

import torch

from .factory import Architecture
from .utilities.state_dict import canonicalize_state_dict


class ArchitectureNotFound(Exception):
    pass


# globals a plain tensor state dict needs; anything else is refused (reference :20-39)
_ALLOWED_GLOBALS = frozenset(
    {
        ('collections', 'OrderedDict'),
        ('typing', 'OrderedDict'),
        ('torch._utils', '_rebuild_tensor_v2'),
        ('torch', 'BFloat16Storage'),
        ('torch', 'FloatStorage'),
        ('torch', 'HalfStorage'),
        ('torch', 'IntStorage'),
        ('torch', 'LongStorage'),
        ('torch', 'DoubleStorage'),
    }
)


class RestrictedUnpickler(pickle.Unpickler):
    def find_class(self, module: str, name: str):
        if (module, name) not in _ALLOWED_GLOBALS:
            raise pickle.UnpicklingError(f"Global '{module}.{name}' is forbidden")
        return super().find_class(module, name)


RestrictedUnpickle = SimpleNamespace(
    Unpickler=RestrictedUnpickler,
    __name__='pickle',
    load=lambda *args, **kwargs: RestrictedUnpickler(*args, **kwargs).load(),
)

_PICKLE_EXTENSIONS = ('.pth', '.ckpt')


def _read_pickled(path: str):
    # weights_only=False: torch hands the stream to our allow-listing unpickler instead of its own
    return torch.load(path, pickle_module=RestrictedUnpickle, weights_only=False)


class Registry:
    def __init__(self):
        self.store: Dict[str, Architecture] = {}

    def __contains__(self, uid: str) -> bool:
        return uid in self.store

    def __iter__(self) -> Iterator[Architecture]:
        return iter(list(self.store.values()))

    def __len__(self) -> int:
        return len(self.store)

    def add(self, arch: Architecture):
        self.store[arch.id] = arch

    def get(self, uid: str) -> Architecture:
        # unknown ids surface as KeyError, exactly like the reference's dict lookup (registry.py:74)
        arch = self.store[uid]
        if not arch:
            raise ArchitectureNotFound
        return arch

    def load_from_file(self, path: str):
        ext = os.path.splitext(path)[1].lower()
        if ext == '.pt':
            # TorchScript archive first, plain pickle as the fallback (reference :81-93)
            try:
                state_dict = torch.jit.load(path).state_dict()
            except RuntimeError:
                try:
                    state_dict = _read_pickled(path)
                except Exception:
                    state_dict = None
                if state_dict is None:
                    raise
        elif ext in _PICKLE_EXTENSIONS:
            state_dict = _read_pickled(path)
        elif ext == '.safetensors':
            import safetensors.torch

            state_dict = safetensors.torch.load_file(path)
        else:
            raise ValueError(f'Unsupported model file extension {ext}. Please try a supported model type.')
        return self.load_from_state_dict(state_dict)

    def load_from_state_dict(self, state_dict: Mapping[str, object]):
        state_dict = canonicalize_state_dict(state_dict)
        for arch in self.store.values():
            if arch.detect(state_dict):
                model = arch.load(state_dict)
                model.load_state_dict(state_dict)
                return model
        raise ArchitectureNotFound

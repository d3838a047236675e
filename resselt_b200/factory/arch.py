"""Plugin contract: ``Architecture`` + ``ModelMetadata``.

Same surface as /root/reference/resselt/factory/arch.py:12-36: an architecture has an ``id``,
a ``detect(state_dict)`` predicate and an abstract ``load(state_dict) -> nn.Module`` whose result
carries ``parameters_info`` (a ``ModelMetadata``).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from dataclasses import dataclass
from typing import Generic, Mapping, Sequence, TypeVar

import torch

from .key_condition import KeyCondition

T = TypeVar('T', bound=torch.nn.Module, covariant=True)


@dataclass
class ModelMetadata:
    # field order (in, out, upscale, name) is part of the reference contract (factory/arch.py:12-19)
    in_channels: int
    out_channels: int
    upscale: int | Sequence[int]
    name: str


class Architecture(ABC, Generic[T]):
    def __init__(self, uid: str, detect: KeyCondition):
        self.id = uid
        self._detect = detect

    def detect(self, state_dict: Mapping[str, object]) -> bool:
        return self._detect(state_dict)

    @abstractmethod
    def load(self, state_dict: Mapping[str, object]) -> T:
        raise NotImplementedError

    def _enhance_model(self, model: T, in_channels: int, out_channels: int, upscale, name) -> T:
        model.parameters_info = ModelMetadata(
            in_channels=in_channels, out_channels=out_channels, upscale=upscale, name=name
        )
        return model

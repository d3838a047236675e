"""State-dict key predicates used by architecture plugins to recognise a checkpoint.

Mirrors the behaviour of the reference's ``KeyCondition``
(/root/reference/resselt/factory/key_condition.py:6-32): a condition is a tree of
``all``/``any`` nodes whose leaves are key names; a leaf holds when the key is present.
"""
from __future__ import annotations

from typing import Literal, Mapping, Union

Clause = Union[str, 'KeyCondition']


class KeyCondition:
    __slots__ = ('_kind', '_keys')

    def __init__(self, kind: Literal['all', 'any'], keys: tuple[Clause, ...]):
        if kind not in ('all', 'any'):
            raise ValueError(f'unknown KeyCondition kind {kind!r}')
        self._kind = kind
        self._keys = tuple(keys)

    @staticmethod
    def has_all(*keys: Clause) -> 'KeyCondition':
        return KeyCondition('all', keys)

    @staticmethod
    def has_any(*keys: Clause) -> 'KeyCondition':
        return KeyCondition('any', keys)

    def __call__(self, state_dict: Mapping[str, object]) -> bool:
        fold = all if self._kind == 'all' else any
        return fold((k(state_dict) if isinstance(k, KeyCondition) else k in state_dict) for k in self._keys)

    def __repr__(self) -> str:
        return f'KeyCondition.has_{self._kind}{self._keys!r}'

"""Base class of the engine-backed ``nn.Module``s the architecture plugins return.

The module owns parameters/buffers under exactly the names of the reference checkpoint (so the
registry's strict ``load_state_dict`` passes, /root/reference/resselt/registry.py:112-113) and
lazily compiles them into a native plan per (device, compute dtype).  ``forward`` is a C-ABI call;
there is no PyTorch-op fallback.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, Optional, Tuple

import numpy as np
import torch
from torch import nn

from .plan import Plan, PlanBuilder


class ParamGroup(nn.Module):
    """Name-space node of the parameter tree (mirrors the reference's sub-module names)."""


ParamSpec = Tuple[str, Tuple[int, ...], str]  # (dotted name, shape, kind)
# kinds: 'conv_w[*gain]', 'bias:<fan_in>', 'prelu', 'ones', 'zeros', 'buffer_zeros', 'normal:<std>'


def _init_tensor(shape, kind: str, rng: np.random.RandomState) -> torch.Tensor:
    if kind == 'buffer_tensor':  # the "shape" slot carries the (deterministic) value itself
        return shape.clone()
    if kind == 'buffer_long':
        return torch.zeros(shape, dtype=torch.int64)
    if kind == 'buffer_var':
        return torch.from_numpy(rng.uniform(0.5, 1.5, size=shape).astype(np.float32))
    if kind.startswith('buffer_normal:'):
        return torch.from_numpy(np.asarray(rng.normal(0.0, float(kind.split(':')[1]), size=shape), dtype=np.float32))
    if kind.startswith('conv_w'):  # 'conv_w' or 'conv_w*<gain>'
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else shape[0]
        bound = (float(kind.split('*')[1]) if '*' in kind else 1.0) / math.sqrt(max(fan_in, 1))
        a = rng.uniform(-bound, bound, size=shape)
    elif kind.startswith('bias:'):
        bound = 1.0 / math.sqrt(max(int(kind.split(':')[1]), 1))
        a = rng.uniform(-bound, bound, size=shape)
    elif kind == 'prelu':
        a = np.full(shape, 0.25) + rng.uniform(-0.1, 0.1, size=shape)
    elif kind == 'ones':
        a = np.ones(shape)
    elif kind in ('zeros', 'buffer_zeros'):
        a = np.zeros(shape)
    elif kind.startswith('normal:'):
        a = rng.normal(0.0, float(kind.split(':')[1]), size=shape)
    elif kind.startswith('affine_w'):
        a = 1.0 + rng.uniform(-0.2, 0.2, size=shape)
    else:
        raise ValueError(f'unknown init kind {kind}')
    return torch.from_numpy(np.asarray(a, dtype=np.float32).reshape(shape))


def random_state_dict(specs: Iterable[ParamSpec], seed: int = 0) -> Dict[str, torch.Tensor]:
    """Deterministic (numpy RandomState) weights for a parameter spec — used by tests, bench and fixtures."""
    rng = np.random.RandomState(seed)
    return {name: _init_tensor(shape, kind, rng) for name, shape, kind in specs}


class EngineModule(nn.Module):
    def __init__(self, specs: Iterable[ParamSpec], in_channels: int, out_channels: int, upscale: int, seed: int = 0,
                 plan_io: Optional[Tuple[int, int, int]] = None):
        super().__init__()
        # what the caller sees ...
        self.in_channels, self.out_channels, self.upscale = in_channels, out_channels, upscale
        # ... and what the native plan is built for (differs only when host-side glue reshapes the input first)
        self._plan_io = plan_io if plan_io is not None else (in_channels, out_channels, upscale)
        self._plans: Dict[tuple, Plan] = {}
        self._plan_stamp: Dict[tuple, tuple] = {}
        rng = np.random.RandomState(seed)
        for name, shape, kind in specs:
            self._register(name, _init_tensor(shape, kind, rng), is_buffer=kind.startswith('buffer'))

    # ---------------------------------------------------------------- parameter tree
    def _register(self, dotted: str, value: torch.Tensor, is_buffer: bool) -> None:
        node: nn.Module = self
        *path, leaf = dotted.split('.')
        for part in path:
            child = node._modules.get(part)
            if child is None:
                child = ParamGroup()
                node.add_module(part, child)
            node = child
        if is_buffer:
            node.register_buffer(leaf, value)
        else:
            node.register_parameter(leaf, nn.Parameter(value, requires_grad=False))

    def _weights(self) -> Dict[str, torch.Tensor]:
        """fp64 host copies of every parameter/buffer, keyed like the checkpoint."""
        return {k: v.detach().to('cpu', torch.float64) for k, v in self.state_dict().items()}

    def _stamp(self) -> tuple:
        """Fingerprint of the weights the cached native plan was packed from: one (storage address, version counter) pair per
        parameter / buffer.  It changes when a tensor is replaced (``load_state_dict``, ``.to()``) or written in place through
        autograd-visible ops (``p.mul_()``, ``p.copy_()``).  It does NOT see writes that bypass the version counter —
        ``p.data.mul_(2)``, ``p.data.copy_(...)`` (the idiom of EMA / model-interpolation tools) keep both the address and the
        version: call ``invalidate()`` (or ``refresh()``) after such edits, otherwise the plan keeps the old packed weights."""
        return tuple((t.data_ptr(), 0 if t.is_inference() else t._version)
                     for group in (self.parameters(), self.buffers()) for t in group)

    # ---------------------------------------------------------------- plan management
    def build_plan(self, pb: PlanBuilder, w: Dict[str, torch.Tensor]) -> None:  # pragma: no cover - abstract
        raise NotImplementedError

    def _plan_variant(self):
        """Hashable tag of the plan flavour the next forward needs (None: the module has one plan per device and dtype).
        Modules whose output geometry depends on a call argument (SpanPP's ``scale``) override this and ``_plan_io_for_variant``."""
        return None

    def _plan_io_for_variant(self) -> Tuple[int, int, int]:
        return self._plan_io

    def invalidate(self) -> None:
        """Drop every cached native plan; the next forward re-merges, re-packs and re-uploads the weights.  Needed only after
        weight edits the fingerprint cannot see (writes through ``.data``; see ``_stamp``)."""
        self._plans.clear()
        self._plan_stamp.clear()

    refresh = invalidate

    def _apply(self, fn, *args, **kwargs):
        self.invalidate()
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self.invalidate()
        return super().load_state_dict(*args, **kwargs)

    @staticmethod
    def compute_dtype_for(x_dtype: torch.dtype) -> torch.dtype:
        # bf16 tensors take the tcgen05 path; fp32 and fp16 tensors the fp32 CUDA-core path
        return torch.bfloat16 if x_dtype == torch.bfloat16 else torch.float32

    def plan_for(self, device: torch.device, x_dtype: torch.dtype) -> Plan:
        if device.type != 'cuda':
            raise RuntimeError(
                'resselt_b200 modules run on CUDA sm_100 (B200) devices only; there is no CPU path. '
                f'Got an input on {device}.'
            )
        index = device.index if device.index is not None else torch.cuda.current_device()
        cdt = self.compute_dtype_for(x_dtype)
        variant = self._plan_variant()
        key = (index, cdt) if variant is None else (index, cdt, variant)
        stamp = self._stamp()
        plan = self._plans.get(key)
        if plan is None or self._plan_stamp.get(key) != stamp:
            pb = PlanBuilder(cdt, *(self._plan_io if variant is None else self._plan_io_for_variant()),
                             base_divisor=getattr(self, '_plan_base_divisor', 1))
            self.build_plan(pb, self._weights())
            plan = pb.finalize(torch.device('cuda', index))
            self._plans[key] = plan
            self._plan_stamp[key] = stamp
        return plan

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.plan_for(x.device, x.dtype).forward(x)

    def forward_into(self, x: torch.Tensor, out: torch.Tensor) -> torch.Tensor:
        """Forward writing into a caller-owned NCHW tensor (streaming runners reuse their output slots); repeated calls
        with the same tensors replay a captured CUDA graph (Plan.forward, ``graph=True``)."""
        return self.plan_for(x.device, x.dtype).forward(x, out=out, graph=True)

    @property
    def receptive_radius(self) -> int:
        """Input pixels beyond a tile edge that influence the tile's output (exact halo for tiled_forward)."""
        raise NotImplementedError

    @property
    def tile_multiple(self) -> int:
        """Tile origins / sizes of an exact-halo tiling must be multiples of this (1 unless the model re-grids its input)."""
        return 1

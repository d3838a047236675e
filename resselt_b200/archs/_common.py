"""Helpers shared by the architecture plugins: parameter-spec builders and weight algebra."""
from __future__ import annotations

from typing import Dict, List

import torch

from ..engine import ParamSpec


def conv_specs(prefix: str, cin: int, cout: int, k: int, gain: float = 1.0) -> List[ParamSpec]:
    """Names/shapes of one ``nn.Conv2d(cin, cout, k)`` with bias (``gain`` scales the random-init weight range)."""
    fan_in = cin * k * k
    kind = 'conv_w' if gain == 1.0 else f'conv_w*{gain}'
    return [(f'{prefix}.weight', (cout, cin, k, k), kind), (f'{prefix}.bias', (cout,), f'bias:{fan_in}')]


def conv3xc_specs(prefix: str, cin: int, cout: int, gain: int = 2) -> List[ParamSpec]:
    """Parameter names of a re-parameterisable Conv3XC (reference span/arch.py:59-122):
    ``sk`` 1x1, ``conv.0`` 1x1 expand, ``conv.1`` 3x3, ``conv.2`` 1x1 reduce, and the (dead) ``eval_conv``."""
    return (
        conv_specs(f'{prefix}.sk', cin, cout, 1)
        + conv_specs(f'{prefix}.conv.0', cin, cin * gain, 1)
        + conv_specs(f'{prefix}.conv.1', cin * gain, cout * gain, 3)
        + conv_specs(f'{prefix}.conv.2', cout * gain, cout, 1)
        + conv_specs(f'{prefix}.eval_conv', cin, cout, 3)
    )


def merge_conv3xc(w: Dict[str, torch.Tensor], prefix: str):
    """Collapse 1x1 -> 3x3 -> 1x1 plus the parallel 1x1 skip into a single 3x3 (weight, bias), in fp64.

    Closed form of what the reference recomputes on every forward (span/arch.py:124-150):
        W[o,i,:,:] = sum_{n,m} w3[o,n] * w2[n,m,:,:] * w1[m,i]   (+ sk_w[o,i] at the centre tap)
        b[o]       = sum_n w3[o,n] * (sum_{m,kh,kw} w2[n,m,kh,kw] * b1[m] + b2[n]) + b3[o] + sk_b[o]
    The checkpoint's ``eval_conv.*`` tensors are never read: the reference overwrites them before use.
    """
    f64 = torch.float64
    w1, b1 = w[f'{prefix}.conv.0.weight'].to(f64)[:, :, 0, 0], w[f'{prefix}.conv.0.bias'].to(f64)
    w2, b2 = w[f'{prefix}.conv.1.weight'].to(f64), w[f'{prefix}.conv.1.bias'].to(f64)
    w3, b3 = w[f'{prefix}.conv.2.weight'].to(f64)[:, :, 0, 0], w[f'{prefix}.conv.2.bias'].to(f64)
    sk_w, sk_b = w[f'{prefix}.sk.weight'].to(f64)[:, :, 0, 0], w[f'{prefix}.sk.bias'].to(f64)
    merged = torch.einsum('on,nmhw,mi->oihw', w3, w2, w1)
    merged[:, :, 1, 1] += sk_w
    inner_bias = torch.einsum('nmhw,m->n', w2, b1) + b2
    bias = w3 @ inner_bias + b3 + sk_b
    return merged, bias

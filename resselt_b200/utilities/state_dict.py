"""Checkpoint helpers shared by the architecture plugins.

Behavioural contract: /root/reference/resselt/utilities/state_dict.py:5-96.
"""
from __future__ import annotations

import math
from typing import Mapping

_WRAPPER_KEYS = ('state_dict', 'params_ema', 'params-ema', 'params', 'model', 'net')
_COMMON_PREFIXES = ('module.', 'netG.')


def remove_common_prefix(state_dict: Mapping[str, object], prefixes) -> Mapping[str, object]:
    """Strip a prefix only when *every* key carries it (reference :5-13)."""
    if len(state_dict) == 0:
        return state_dict
    for prefix in prefixes:
        if all(key.startswith(prefix) for key in state_dict):
            cut = len(prefix)
            state_dict = {key[cut:]: value for key, value in state_dict.items()}
    return state_dict


def canonicalize_state_dict(state_dict: Mapping[str, object]) -> Mapping[str, object]:
    """Unwrap one level of trainer wrapping and drop DataParallel/GAN prefixes (reference :16-36)."""
    for wrapper in _WRAPPER_KEYS:
        inner = state_dict.get(wrapper) if hasattr(state_dict, 'get') else None
        if isinstance(inner, dict):
            state_dict = inner
            break
    return remove_common_prefix(state_dict, _COMMON_PREFIXES)


def pixelshuffle_scale(ps_size: int, channels: int) -> int:
    """Upscale factor of a conv feeding PixelShuffle: ps_size = channels * r^2 (reference :39-40)."""
    return math.isqrt(ps_size // channels)


def dysample_scale(ds_size: int) -> int:
    """Upscale factor from a DySample offset conv (reference :43-44)."""
    return math.isqrt(ds_size // 8)


def get_seq_len(state_dict: Mapping[str, object], seq_key: str) -> int:
    """Length of the ``nn.Sequential``/``ModuleList`` stored under ``seq_key`` (reference :76-96)."""
    prefix = seq_key + '.'
    top = -1
    for key in state_dict:
        if key.startswith(prefix):
            top = max(top, int(key[len(prefix):].split('.', 1)[0]))
    return top + 1


def get_pixelshuffle_params(state_dict: Mapping[str, object], upsample_key: str = 'upsample', default_nf: int = 64):
    """(upscale, num_feat) of an alternating conv/PixelShuffle ``Upsample`` stack (reference :47-73)."""
    upscale, num_feat = 1, default_nf
    for i in range(0, 10, 2):
        weight = state_dict.get(f'{upsample_key}.{i}.weight')
        if weight is None:
            break
        num_feat = weight.shape[1]
        upscale *= math.isqrt(weight.shape[0] // num_feat)
    return upscale, num_feat

"""Architecture plugins and the module-level registry.

The reference registers every ``Architecture`` subclass found by an ``os.walk`` of its ``archs``
package (/root/reference/resselt/archs/__init__.py:7-28), i.e. in filesystem order.  Here the
in-scope plugins are registered explicitly, in a fixed order.
"""
from ..registry import Registry
from .compact import CompactArch, SRVGGNetCompact
from .dat import DAT, DatArch
from .esrgan import ESRGANArch, RRDBNet
from .plksr import PLKSR, PLKSRArch, RealPLKSR
from .span import SPAN, SPANArch
from .spanplus import SpanPlus, SpanPlusArch
from .gaterv3 import GateRV3, GateRV3Arch
from .rtmosr import RTMoSR, RTMoSRArch
from .spanpp import SpanPP, SpanPPArch
from .swinir import SwinIR, SwinIRArch

internal_registry = Registry()
for _arch in (SPANArch, SpanPlusArch, SpanPPArch, RTMoSRArch, GateRV3Arch, CompactArch, ESRGANArch, PLKSRArch, DatArch, SwinIRArch):
    internal_registry.add(_arch())

__all__ = ['internal_registry', 'SPAN', 'SPANArch', 'SpanPlus', 'SpanPlusArch', 'SRVGGNetCompact', 'CompactArch', 'RRDBNet', 'ESRGANArch', 'RealPLKSR', 'PLKSR', 'PLKSRArch', 'DAT', 'DatArch', 'SwinIR', 'SwinIRArch', 'SpanPP', 'SpanPPArch', 'RTMoSR', 'RTMoSRArch', 'GateRV3', 'GateRV3Arch']

"""The C-ABI library: loads without a GPU, exports every symbol include/*.h declares, validates arguments, and the
product path refuses to run without a device (no CPU fallback).  No compute calls here."""
import ctypes as C
import glob
import os
import re

import pytest
import torch

from resselt_b200.engine import INPUT, OUTPUT, PlanBuilder
from resselt_b200.engine import native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    N.build_library()
    return N.lib()


def test_header_symbols_are_exported(lib):
    declared = set()
    for header in glob.glob(os.path.join(ROOT, 'include', '*.h')):
        text = re.sub(r'/\*.*?\*/', '', open(header).read(), flags=re.S)
        declared |= set(re.findall(r'\b(rsb_[a-z_]+)\s*\(', text))
    assert declared, 'no declarations found'
    assert declared == set(N.EXPORTED_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/ but not exported'
    assert lib.rsb_version() == 204


def test_desc_struct_sizes_match_header_layout():
    # 4-byte fields + 8-byte pointers, natural alignment (what a C compiler produces for the header's structs)
    assert C.sizeof(N.ConvDesc) == 208
    assert C.sizeof(N.GroupNormDesc) == 56


def test_plan_building_and_argument_validation(lib):
    pb = PlanBuilder(torch.bfloat16, 3, 3, 2)
    a = pb.buffer(48)
    b = pb.buffer(12)
    w = torch.zeros(48, 3, 3, 3)
    pb.conv(INPUT, a, w, torch.zeros(48), act=N.ACT_SILU)
    with pytest.raises(N.NativeError) as e:  # spatial conv in place
        pb.conv(a, a, torch.zeros(48, 48, 3, 3))
    assert e.value.code == -1 and 'in place' in str(e.value)
    with pytest.raises(N.NativeError):  # even kernel
        pb.conv(a, b, torch.zeros(12, 48, 2, 2))
    with pytest.raises(N.NativeError):  # output channel mismatch for PixelShuffle(2): 3*4 = 12 expected
        pb.conv(a, OUTPUT, torch.zeros(16, 48, 3, 3), ps=2)
    with pytest.raises(N.NativeError):  # PReLU without slopes
        pb.conv(a, b, torch.zeros(12, 48, 3, 3), act=N.ACT_PRELU)
    pb.conv(a, OUTPUT, torch.zeros(12, 48, 3, 3), ps=2)
    assert lib.rsb_plan_num_ops(pb._h) == 2
    flops = C.c_double()
    assert lib.rsb_plan_flops(pb._h, 1, 10, 10, C.byref(flops)) == 0
    assert flops.value == 2 * 100 * (48 * 27 + 12 * 48 * 9)
    nbytes = C.c_size_t()
    assert lib.rsb_plan_workspace_bytes(pb._h, 1, 16, 16, C.byref(nbytes)) == 0
    # planar-8 bf16 planes of 16x16 pixels x 16 B: 48 ch -> 6, 12 ch -> 2 (whole 16-channel K steps),
    # + the hidden planar copy of the normalised input feeding the 3x3 stem (3 -> 16 channels -> 2 planes)
    assert nbytes.value == (6 + 2 + 2) * 16 * 16 * 16


@pytest.mark.skipif(torch.cuda.is_available(), reason='checks the no-device behaviour')
def test_no_device_means_loud_failure(lib):
    from resselt_b200.archs import SRVGGNetCompact

    assert lib.rsb_device_count() == 0
    pb = PlanBuilder(torch.float32, 3, 3, 1)
    pb.conv(INPUT, OUTPUT, torch.zeros(3, 3, 3, 3))
    with pytest.raises(N.NativeError) as e:
        pb.finalize(torch.device('cuda', 0))
    assert e.value.code == -5
    with pytest.raises(RuntimeError, match='no CPU path'):
        SRVGGNetCompact(num_feat=16, num_conv=1)(torch.rand(1, 3, 8, 8))


def test_product_does_not_import_oracle():
    for path in glob.glob(os.path.join(ROOT, 'resselt_b200', '**', '*.py'), recursive=True):
        src = open(path).read()
        assert not re.search(r'^\s*(from|import)\s+oracle\b', src, flags=re.M), f'{path} imports the test oracle'


def test_descriptor_struct_sizes_match_the_library():
    import ctypes as C

    from resselt_b200.engine import native as N

    lib = N.lib()
    for which, struct in enumerate((N.ConvDesc, N.GroupNormDesc, N.OpDesc, N.OpInfo)):
        assert lib.rsb_abi_struct_size(which) == C.sizeof(struct), struct.__name__
    assert lib.rsb_abi_struct_size(99) == -1

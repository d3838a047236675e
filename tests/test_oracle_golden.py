"""The CPU oracle against outputs of the reference itself (tests/golden, made by oracle/make_golden.py),
plus the semantic known-answer checks SURVEY.md §8c lists."""
import pytest
import torch
import torch.nn.functional as F

import oracle
from oracle.sr_forward import conv3xc_merged
from conftest import golden_case, golden_index, norm_err


@pytest.mark.parametrize('name', sorted(golden_index()))
def test_oracle_matches_reference_fixture(name):
    kind, sd, x, y_ref, meta = golden_case(name)
    y32 = oracle.forward_by_name(kind, sd, x, torch.float32)
    y64 = oracle.forward_by_name(kind, sd, x, torch.float64)
    assert y32.shape == y_ref.shape
    # fp32 re-association noise only (the reference forward itself is fp32)
    assert norm_err(y32, y_ref) <= 2e-6
    assert norm_err(y64, y_ref) <= 2e-6
    # SpanPP's loader hands its scale LIST to the metadata (spanpp/__init__.py:132); its forward defaults to eval_base_scale = 2
    upscale = 2 if isinstance(meta['upscale'], list) else meta['upscale']
    if kind == 'RTMoSR':  # the reference's loader reports upscale = 2 whatever the model's scale is (rtmosr/__init__.py:104)
        upscale = y_ref.shape[2] // x.shape[2]
    assert y_ref.shape[1] == meta['out_channels'] and y_ref.shape[2] == x.shape[2] * upscale


def test_dead_eval_conv_tensors_do_not_matter():
    # the reference overwrites eval_conv.* on every forward (span/arch.py:152-154): randomising them changes nothing
    kind, sd, x, y_ref, _ = golden_case('span_x2_f48')
    g = torch.Generator().manual_seed(0)
    sd2 = {k: (torch.randn(v.shape, generator=g) if '.eval_conv.' in k else v) for k, v in sd.items()}
    assert torch.equal(oracle.forward_by_name(kind, sd, x), oracle.forward_by_name(kind, sd2, x))


def test_conv3xc_closed_form_equals_sequential_composition():
    # interior pixels of conv1x1 -> conv3x3 -> conv1x1 (+ 1x1 skip) equal one merged 3x3 conv
    _, sd, _, _, _ = golden_case('span_x2_f48')
    x = torch.randn(1, 48, 12, 12, dtype=torch.float64, generator=torch.Generator().manual_seed(3))
    p = 'block_2.c2_r'
    g = lambda k: sd[f'{p}.{k}'].double()
    seq = F.conv2d(F.conv2d(F.conv2d(x, g('conv.0.weight'), g('conv.0.bias')), g('conv.1.weight'), g('conv.1.bias')),
                   g('conv.2.weight'), g('conv.2.bias')) + F.conv2d(x, g('sk.weight'), g('sk.bias'))[:, :, 1:-1, 1:-1]
    k, b = conv3xc_merged(sd, p, torch.float64)
    merged = F.conv2d(x, k, b)
    assert (seq - merged).abs().max() < 1e-12


def test_engine_weight_merge_matches_oracle_merge():
    from resselt_b200.archs._common import merge_conv3xc

    _, sd, _, _, _ = golden_case('spanplus_x2_b4')
    w = {k: v.double() for k, v in sd.items()}
    for p in ('feats.0', 'feats.1.block_1.c1_r', 'feats.1.conv_2'):
        k1, b1 = merge_conv3xc(w, p)
        k2, b2 = conv3xc_merged(sd, p, torch.float64)
        assert (k1 - k2).abs().max() < 1e-13 and (b1 - b2).abs().max() < 1e-13


def test_spab_second_output_is_activated():
    from oracle.sr_forward import _conv3xc, _spab

    _, sd, _, _, _ = golden_case('span_x2_f48')
    x = torch.randn(1, 48, 8, 8, generator=torch.Generator().manual_seed(5))
    _, o1 = _spab(sd, 'block_6', x, F.silu)
    assert torch.allclose(o1, F.silu(_conv3xc(sd, 'block_6.c1_r', x)))


@pytest.mark.parametrize('lk_type,kw', [('PLK', dict(kernel_size=13)), ('SparsePLK', dict()), ('RectSparsePLK', dict(kernel_size=15))])
def test_plksr_large_kernel_layers_merge_into_one_dense_kernel(lk_type, kw):
    # the engine runs every partial large-kernel layer of the original PLKSR as ONE dense conv (archs/plksr.py::_dense_lk_kernel):
    # the merged kernel must equal the sum of the (dilated / rectangular) branch convs the reference evaluates (plksr.py:153-164, 116-117)
    from resselt_b200.archs import PLKSR
    from resselt_b200.archs.plksr import _dense_lk_kernel

    m = PLKSR(n_blocks=1, upscaling_factor=2, lk_type=lk_type, seed=9, **kw)
    w = {k: v.double() for k, v in m.state_dict().items()}
    k, b = _dense_lk_kernel(w, 'feats.1.lk', lk_type, m.pdim, m.kmax, m.sparse_dilations, with_idt=True)
    x = torch.randn(2, m.pdim, 23, 19, dtype=torch.float64, generator=torch.Generator().manual_seed(2))
    cw = lambda name, **kk: F.conv2d(x, w[f'feats.1.lk.{name}.weight'], w[f'feats.1.lk.{name}.bias'], **kk)
    if lk_type == 'PLK':
        ref = cw('conv', padding=m.kernel_size // 2)
    elif lk_type == 'SparsePLK':
        ref = sum(cw(f'convs.{j}', padding=(5 // 2) * d, dilation=d) for j, d in enumerate(m.sparse_dilations))
    else:
        mm, nn_ = m.kernel_size, m.kernel_size // 3
        ref = cw('mn_conv', padding=(mm // 2, nn_ // 2)) + cw('nm_conv', padding=(nn_ // 2, mm // 2)) + cw('nn_conv', padding=nn_ // 2)
    ref = ref + x  # with_idt
    assert (F.conv2d(x, k, b, padding=m.kmax // 2) - ref).abs().max() < 1e-12


def test_pointwise_conv_merged_into_3x3_is_exact_including_borders():
    """merge_pointwise_into_conv (SPAN's upsampler(conv_cat(cat)), span/arch.py:247-248): one 3x3 conv + border bias table equals
    conv3x3(zero-padded conv1x1 output) at every pixel — corners, edges, 1-pixel-high and 1-pixel-wide images included."""
    from resselt_b200.archs._common import merge_pointwise_into_conv

    g = torch.Generator().manual_seed(4)
    w1, b1 = torch.randn(6, 10, 1, 1, generator=g).double(), torch.randn(6, generator=g).double()
    wk, bk = torch.randn(5, 6, 3, 3, generator=g).double(), torch.randn(5, generator=g).double()
    wm, bm, border = merge_pointwise_into_conv(w1, b1, wk, bk)
    for H, W in ((7, 9), (1, 6), (5, 1), (1, 1), (2, 2)):
        x = torch.randn(2, 10, H, W, generator=g).double()
        ref = F.conv2d(F.conv2d(x, w1, b1), wk, bk, padding=1)
        got = F.conv2d(x, wm, bm, padding=1)
        ys, xs = torch.arange(H).view(-1, 1), torch.arange(W).view(1, -1)
        mask = (ys == 0).long() + 2 * (ys == H - 1).long() + 4 * (xs == 0).long() + 8 * (xs == W - 1).long()
        got = got + border[mask].permute(2, 0, 1).unsqueeze(0)
        assert (got - ref).abs().max() < 1e-12, (H, W)

"""GPU parity tests (run with -m gpu on a B200): every call goes through the C ABI (ctypes -> libresselt_b200.so).

Bars (BASELINE.json north_star): fp32 path max-abs <= 1e-4 (range-normalised), bf16 path PSNR >= 50 dB
against the fp32 reference, tile seams bit-identical to the untiled device output."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
import resselt_b200
from conftest import golden_case, golden_index, norm_err, psnr
from resselt_b200.archs import DAT, PLKSR, SPAN, GateRV3, RealPLKSR, RRDBNet, RTMoSR, SpanPlus, SpanPP, SRVGGNetCompact, SwinIR
from resselt_b200.engine import INPUT, OUTPUT, PlanBuilder
from resselt_b200.engine import native as N
from resselt_b200.runner import FramePipeline, tiled_forward

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'
FP32_TOL = 1e-4
BF16_PSNR_DB = 50.0


def _load(sd, dtype=torch.float32):
    m = resselt_b200.load_from_state_dict(dict(sd)).eval().to(DEV)
    return m.bfloat16() if dtype == torch.bfloat16 else m


def test_native_library_is_the_path_that_runs():
    assert N.lib().rsb_device_count() >= 1
    m = _load(SRVGGNetCompact(num_feat=16, num_conv=1, upscale=2).state_dict())
    y = m(torch.rand(1, 3, 8, 8, device=DEV))
    assert y.shape == (1, 3, 16, 16) and y.is_cuda
    assert m.plan_for(torch.device(DEV), torch.float32).launches_per_forward == 3


@pytest.mark.parametrize('name', sorted(golden_index()))
def test_golden_fixtures_fp32_and_bf16(name):
    kind, sd, x, y_ref, meta = golden_case(name)
    with torch.inference_mode():
        y32 = _load(sd)(x.to(DEV)).cpu()
        y16 = _load(sd, torch.bfloat16)(x.to(DEV, torch.bfloat16)).float().cpu()
    assert y32.shape == y_ref.shape and y32.dtype == torch.float32
    assert norm_err(y32, y_ref) <= FP32_TOL, f'fp32 path vs reference output: {norm_err(y32, y_ref):.3e}'
    assert psnr(y16, y_ref) >= BF16_PSNR_DB, f'bf16 path PSNR {psnr(y16, y_ref):.2f} dB'


@pytest.mark.parametrize(
    'kind,model,shape',
    [
        ('SPAN', SPAN(feature_channels=48, upscale=2, seed=21), (1, 3, 67, 93)),
        ('SPAN', SPAN(feature_channels=48, upscale=4, seed=22), (2, 3, 33, 40)),
        ('SPANPlus', SpanPlus(blocks=[4], feature_channels=48, upscale=2, seed=23), (1, 3, 70, 50)),
        ('SPANPlus', SpanPlus(blocks=[4], feature_channels=48, upscale=2, upsampler='dys', seed=43), (1, 3, 70, 50)),   # DySample head (default upsampler)
        ('SPANPlus', SpanPlus(blocks=[2], feature_channels=48, upscale=4, upsampler='dys', seed=44), (2, 3, 33, 47)),
        ('RealPLKSR', RealPLKSR(n_blocks=4, upscaling_factor=4, dysample=True, seed=45), (1, 3, 48, 56)),
        ('RealPLKSR', RealPLKSR(dim=32, n_blocks=2, upscaling_factor=3, kernel_size=13, dysample=True, seed=46), (2, 3, 21, 30)),  # odd factor: groups = 3
        ('Compact', SRVGGNetCompact(num_feat=64, num_conv=16, upscale=4, seed=24), (3, 3, 45, 61)),
        ('Compact', SRVGGNetCompact(num_feat=64, num_conv=16, upscale=1, seed=25), (1, 3, 40, 40)),
        ('SpanPP', SpanPP(feature_channels=48, implicit_dim=64, latent_layers=2, seed=52), (2, 3, 37, 50)),
        ('GateRV3', GateRV3(scale=2, span_blocks=2, num_latent=3, seed=56), (1, 3, 70, 90)),                              # full depth (2, 2, 4, 8) U-Net, dim 32, reflect pad to 16
        ('GateRV3', GateRV3(dim=48, enc_blocks=(1, 1, 1), dec_blocks=(1, 1, 1), num_latent=2, scale=4, upsample='pixelshuffle', upsample_mid_dim=32, seed=57), (2, 3, 32, 40)),
        ('RTMoSR', RTMoSR(scale=2, dim=32, n_blocks=2, seed=53), (1, 3, 45, 61)),                                  # odd size: reflect pad + crop
        ('RTMoSR', RTMoSR(scale=4, dim=64, ffn_expansion=2, n_blocks=3, seed=54), (2, 3, 40, 36)),
        ('RTMoSR', RTMoSR(scale=2, dim=32, n_blocks=1, unshuffle_mod=True, seed=55), (1, 3, 41, 54)),              # pixel-unshuffle front end
        ('ESRGAN', RRDBNet(num_blocks=23, scale=4, seed=26), (1, 3, 48, 40)),           # full depth: 69 dense blocks
        ('ESRGAN', RRDBNet(num_blocks=2, scale=8, seed=27), (1, 3, 19, 23)),
        ('ESRGAN', RRDBNet(num_blocks=2, scale=4, key_style='new', seed=28), (2, 3, 24, 17)),  # Real-ESRGAN key names
        ('ESRGAN', RRDBNet(in_nc=12, out_nc=3, num_blocks=2, scale=4, shuffle_factor=2, seed=29), (1, 3, 33, 27)),
        ('RealPLKSR', RealPLKSR(n_blocks=28, upscaling_factor=4, seed=30), (1, 3, 48, 56)),  # full depth
        ('RealPLKSR', RealPLKSR(dim=32, n_blocks=3, upscaling_factor=2, kernel_size=13, use_ea=False, seed=31), (2, 3, 21, 30)),
        ('PLKSR', PLKSR(n_blocks=28, upscaling_factor=4, seed=40), (1, 3, 48, 56)),                                   # full depth, DCCM + 17x17 PLK + EA
        ('PLKSR', PLKSR(n_blocks=3, upscaling_factor=2, ccm_type='ICCM', lk_type='SparsePLK', seed=41), (2, 3, 33, 30)),  # dilated branches -> one 17x17
        ('PLKSR', PLKSR(dim=32, n_blocks=3, upscaling_factor=3, ccm_type='CCM', lk_type='RectSparsePLK', kernel_size=15, use_ea=False, seed=42),
         (1, 3, 40, 41)),
        ('DAT', DAT(upscale=4, seed=32), (1, 3, 64, 64)),                                    # default 6x6 blocks, 180 ch, 8x32 windows
        ('DAT', DAT(depth=[3, 3], num_heads=[6, 6], upscale=2, seed=33), (2, 3, 37, 45)),      # padding + shifted-window masks
        ('DAT', DAT(embed_dim=60, split_size=[4, 8], depth=[3, 2], num_heads=[2, 2], upscale=2, img_size=32, seed=34), (1, 3, 50, 30)),
        ('DAT', DAT(embed_dim=60, split_size=[8, 32], depth=[3, 2], num_heads=[6, 6], upscale=3, upsampler='pixelshuffledirect', resi_connection='3conv', seed=47),
         (1, 3, 40, 70)),                                                                      # DAT-light layout: one-step head, 3conv residual connections
        ('SwinIR', SwinIR(upscale=4, seed=35), (1, 3, 64, 64)),                                # classical SR: 6x6 blocks, 180 ch, window 8
        ('SwinIR', SwinIR(depths=[2], num_heads=[6], upscale=2, seed=39), (1, 3, 136, 200)),       # > 2 tiles per CTA in every layer
        ('SwinIR', SwinIR(embed_dim=60, depths=[6, 6, 6, 6], num_heads=[6, 6, 6, 6], upscale=2, upsampler='pixelshuffledirect', seed=36), (2, 3, 37, 45)),
        ('SwinIR', SwinIR(embed_dim=240, depths=[2, 2], num_heads=[8, 8], upscale=4, upsampler='nearest+conv', resi_connection='3conv', seed=37),
         (1, 3, 40, 52)),                                                                      # real-world SR large model shape
        ('SwinIR', SwinIR(in_chans=1, embed_dim=48, depths=[2, 2], num_heads=[6, 6], window_size=7, img_size=126, img_range=255.0, upsampler='', seed=38),
         (1, 1, 33, 47)),                                                                      # JPEG-artifact model: window 7, residual head
    ],
)
def test_against_oracle_on_seeded_inputs(kind, model, shape):
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(shape[2] * 1000 + shape[3]))
    ref = oracle.forward_by_name(kind, sd, x, torch.float32)
    with torch.inference_mode():
        y32 = _load(sd)(x.to(DEV)).cpu()
        m16 = _load(sd, torch.bfloat16)
        xb = x.to(DEV, torch.bfloat16)
        y16 = m16(xb).float().cpu()
        plan = m16.plan_for(torch.device(DEV), torch.bfloat16)
        plan.force_direct = True  # same plan on the CUDA-core kernels: cross-checks the tcgen05 kernel on device
        y16_direct = m16(xb).float().cpu()
        plan.force_direct = False
    assert norm_err(y32, ref) <= FP32_TOL
    assert psnr(y16, ref) >= BF16_PSNR_DB
    assert psnr(y16, y16_direct) >= BF16_PSNR_DB  # both are bf16-rounded layer by layer; they agree to rounding noise


LAYER_CASES = [
    # cin, cout, k, n, H, W, act
    (48, 48, 3, 1, 32, 40, N.ACT_NONE),
    (48, 48, 3, 1, 37, 45, N.ACT_SILU),
    (64, 64, 3, 2, 33, 17, N.ACT_MISH),
    (48, 12, 3, 1, 32, 24, N.ACT_NONE),
    (192, 48, 1, 1, 40, 40, N.ACT_NONE),
    (16, 16, 17, 1, 48, 40, N.ACT_NONE),
    (64, 128, 3, 1, 32, 32, N.ACT_LRELU),
    (32, 256, 3, 1, 16, 24, N.ACT_NONE),
    (48, 48, 3, 1, 5, 7, N.ACT_GELU),      # image smaller than one 16x8 tile
    (80, 40, 3, 1, 19, 23, N.ACT_SIGMOID),  # channel counts that are not multiples of 16
    (384, 192, 1, 1, 200, 176, N.ACT_NONE),  # 6 K chunks over a 4-stage ring, several tiles per CTA
    (192, 32, 3, 1, 136, 200, N.ACT_NONE),   # 3 K chunks per tile, two issuing warps alternating tiles
]


@pytest.mark.parametrize('cin,cout,k,n,H,W,act', LAYER_CASES)
def test_single_conv_layer_tensor_core_vs_fp64(cin, cout, k, n, H, W, act):
    g = torch.Generator().manual_seed(cin * 1000 + cout + k)
    x = torch.randn(n, cin, H, W, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * k * k) ** 0.5
    bias = torch.randn(cout, generator=g)
    pb = PlanBuilder(torch.bfloat16, cin, cout, 1)
    a, b = pb.buffer(cin), pb.buffer(cout)
    pb.conv(INPUT, a, torch.eye(cin).view(cin, cin, 1, 1))
    pb.conv(a, b, wt, bias, act=act, act_param=0.2)
    pb.conv(b, OUTPUT, torch.eye(cout).view(cout, cout, 1, 1))
    plan = pb.finalize(torch.device(DEV))
    xd = x.to(DEV, torch.bfloat16)
    got = {}
    for mode in ('tc', 'direct'):
        plan.force_direct = mode == 'direct'
        y = plan.forward(xd)
        got[mode] = (plan.read_buffer(b).double().cpu(), y.double().cpu())
    q = lambda t: t.to(torch.bfloat16).double()
    ref = F.conv2d(q(x), q(wt), bias.double(), padding=k // 2)
    ref = {N.ACT_NONE: lambda t: t, N.ACT_SILU: F.silu, N.ACT_MISH: F.mish, N.ACT_LRELU: lambda t: F.leaky_relu(t, 0.2),
           N.ACT_GELU: F.gelu, N.ACT_SIGMOID: torch.sigmoid}[act](ref)
    scale = float(ref.abs().max())
    for mode in ('tc', 'direct'):
        buf, out = got[mode]
        assert float((buf - ref).abs().max()) / scale < 8e-3, mode  # bf16 output rounding (2^-9 relative) + fp32 accumulation
        assert torch.equal(out, buf), 'identity 1x1 read-out must reproduce the buffer exactly'


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize(
    'model,hw,tile',
    [
        (SPAN(feature_channels=48, upscale=2, seed=31), (90, 120), (40, 56)),
        (SRVGGNetCompact(num_feat=64, num_conv=16, upscale=4, seed=32), (70, 64), (32, 32)),
        (SpanPlus(blocks=[4], feature_channels=48, upscale=2, seed=33), (64, 100), (30, 64)),
        (RRDBNet(num_blocks=1, scale=4, seed=34), (60, 72), (24, 40)),
        # Real-ESRGAN x2: pixel-unshuffle front end -> halo and tile origins in multiples of shuffle_factor, odd image size
        (RRDBNet(in_nc=12, out_nc=3, num_blocks=1, scale=4, shuffle_factor=2, seed=35), (91, 118), (44, 60)),
    ],
)
def test_tile_seams_bit_identical(model, hw, tile, dtype):
    m = _load(model.state_dict(), dtype)
    x = torch.rand(1, 3, *hw, generator=torch.Generator().manual_seed(7)).to(DEV, dtype)
    with torch.inference_mode():
        full = m(x)
        tiled = tiled_forward(m, x, m.upscale, tile, halo=m.receptive_radius)
        short = tiled_forward(m, x, m.upscale, tile, halo=2)
    assert torch.equal(full, tiled), 'exact-halo tiling must reproduce the untiled output bit for bit'
    assert not torch.equal(full, short), 'a too-small halo must be visible (guards against a vacuous test)'


def test_subpixel_and_dual_destination_epilogues():
    """The two buffer-destination modes ESRGAN / RealPLKSR rely on, in isolation, tensor-core vs CUDA-core vs fp64."""
    from resselt_b200.archs.esrgan import upconv_phase_kernels

    g = torch.Generator().manual_seed(77)
    x = torch.randn(1, 64, 21, 19, generator=g)
    wt = torch.randn(64, 64, 3, 3, generator=g) / 24
    bias = torch.randn(64, generator=g)
    w2 = torch.randn(64, 64, 3, 3, generator=g) / 24
    pb = PlanBuilder(torch.bfloat16, 64, 64, 2)
    a, up, side, rest = pb.buffer(64), pb.buffer(64, scale=2), pb.buffer(16), pb.buffer(64)
    pb.conv(INPUT, a, torch.eye(64).view(64, 64, 1, 1))
    for phase, wk, pad in upconv_phase_kernels(wt.double()):
        pb.conv(a, up, wk, bias, dst_ps=2, dst_phase=phase, pad=pad, act=N.ACT_LRELU, act_param=0.2)
    pb.conv(a, side, w2, None, dst2=rest.slice(16, 48))          # channels 0..15 -> side, 16..63 -> rest[16:]
    pb.conv(up, OUTPUT, torch.eye(64).view(64, 64, 1, 1), ps=1)
    plan = pb.finalize(torch.device(DEV))
    q = lambda t: t.to(torch.bfloat16).double()
    ref_up = F.leaky_relu(F.conv2d(F.interpolate(q(x), scale_factor=2, mode='nearest'), wt.double(), bias.double(), padding=1), 0.2)
    ref_split = F.conv2d(q(x), q(w2), None, padding=1)
    for direct in (False, True):
        plan.force_direct = direct
        y = plan.forward(x.to(DEV, torch.bfloat16)).double().cpu()
        assert float((y - ref_up).abs().max()) / float(ref_up.abs().max()) < 1.2e-2   # pre-summed taps are rounded to bf16 once
        got_side, got_rest = plan.read_buffer(side).double().cpu(), plan.read_buffer(rest.slice(16, 48)).double().cpu()
        scale = float(ref_split.abs().max())
        assert float((got_side - ref_split[:, :16]).abs().max()) / scale < 8e-3
        assert float((got_rest - ref_split[:, 16:]).abs().max()) / scale < 8e-3


def test_full_size_1080p_properties():
    """At the benchmark size the oracle is too slow to be the checker; use size-independent properties instead:
    run-to-run determinism, exact-halo tiling == untiled (bit for bit), and a crop computed on its own with the
    exact halo equals the same crop of the full frame."""
    m = _load(SPAN(feature_channels=48, upscale=2, seed=3).state_dict(), torch.bfloat16)
    x = torch.rand(1, 3, 1080, 1920, generator=torch.Generator().manual_seed(11)).to(DEV, torch.bfloat16)
    with torch.inference_mode():
        y1 = m(x).clone()
        y2 = m(x)
        assert torch.equal(y1, y2)
        assert torch.isfinite(y1.float()).all()
        tiled = tiled_forward(m, x, 2, (544, 960), halo=m.receptive_radius)
        assert torch.equal(y1, tiled)
        r = m.receptive_radius
        y0, x0, hh, ww = 400, 700, 128, 160
        crop = m(x[:, :, y0 - r:y0 + hh + r, x0 - r:x0 + ww + r].contiguous())
        assert torch.equal(crop[:, :, 2 * r:2 * (r + hh), 2 * r:2 * (r + ww)], y1[:, :, 2 * y0:2 * (y0 + hh), 2 * x0:2 * (x0 + ww)])
        # and a 256x256 corner of the full frame against the CPU oracle computed on the padded neighbourhood
        sd = {k: v.float().cpu() for k, v in m.state_dict().items()}
        ref = oracle.forward_by_name('SPAN', sd, x[:, :, :256 + r, :256 + r].float().cpu(), torch.float32)[:, :, :512, :512]
        assert psnr(y1[:, :, :512, :512].float().cpu(), ref) >= BF16_PSNR_DB


def test_module_behaviour_dtype_moves_and_reload():
    proto = SRVGGNetCompact(num_feat=32, num_conv=3, upscale=2, seed=41)
    m = _load(proto.state_dict())
    x = torch.rand(2, 3, 20, 28, generator=torch.Generator().manual_seed(1)).to(DEV)
    with torch.inference_mode():
        y = m(x)
        assert torch.equal(y, m(x)), 'forward must be stateless'
        # non-contiguous input view
        xt = x.permute(0, 1, 3, 2).contiguous().permute(0, 1, 3, 2)
        assert not xt.is_contiguous() and torch.equal(m(xt), y)
        # half input/outputs ride the fp32 path
        yh = m.half()(x.half())
        assert yh.dtype == torch.float16 and psnr(yh.float().cpu(), y.cpu()) > 60
        m = m.float()
        # new weights through load_state_dict -> plan is rebuilt
        other = SRVGGNetCompact(num_feat=32, num_conv=3, upscale=2, seed=42).state_dict()
        m.load_state_dict(other)
        y_new = m(x)
        ref = oracle.forward_by_name('Compact', {k: v.clone() for k, v in other.items()}, x.cpu(), torch.float32)
        assert norm_err(y_new.cpu(), ref) <= FP32_TOL and not torch.equal(y_new, y)
    with pytest.raises(RuntimeError, match='channels'):
        m(torch.rand(1, 4, 8, 8, device=DEV))
    with pytest.raises(RuntimeError, match='no CPU path'):
        m(torch.rand(1, 3, 8, 8))


def test_frame_pipeline_matches_direct_forward():
    m = _load(SPAN(feature_channels=48, upscale=2, seed=51).state_dict(), torch.bfloat16)
    frames = [torch.rand(1, 3, 64, 80, generator=torch.Generator().manual_seed(i)).to(torch.bfloat16).pin_memory() for i in range(7)]
    outs = FramePipeline(m, 2, torch.device(DEV), depth=3).run(frames)
    with torch.inference_mode():
        for f, o in zip(frames, outs):
            assert not o.is_cuda and torch.equal(o, m(f.to(DEV)).cpu())


def test_workspace_argument_checks():
    import ctypes as C

    m = _load(SRVGGNetCompact(num_feat=16, num_conv=1, upscale=2).state_dict())
    plan = m.plan_for(torch.device(DEV), torch.float32)
    x = torch.rand(1, 3, 8, 8, device=DEV)
    y = torch.empty(1, 3, 16, 16, device=DEV)
    ws = torch.empty(2048, dtype=torch.uint8, device=DEV)
    rc = plan._lib.rsb_plan_forward(plan._h, x.data_ptr(), N.F32, 1, 8, 8, y.data_ptr(), N.F32, ws.data_ptr(), 16, None, 0)
    assert rc == -4 and b'workspace too small' in plan._lib.rsb_last_error()


@pytest.mark.parametrize('C,n,H,W', [(180, 1, 37, 45), (60, 2, 16, 24), (64, 1, 9, 130), (240, 1, 20, 33), (96, 1, 64, 64), (360, 1, 12, 12)])
def test_layernorm_op_bf16_against_fp64(C, n, H, W):
    # register-resident bf16 LayerNorm (4 threads per pixel) for C <= 256, generic three-pass kernel above that
    g = torch.Generator().manual_seed(C * 7 + H)
    x = torch.randn(n, C, H, W, generator=g) * 2.0 + 0.5
    gamma, beta = 1.0 + 0.2 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    pb = PlanBuilder(torch.bfloat16, C, C, 1)
    a, b = pb.buffer(C), pb.buffer(C)
    pb.conv(INPUT, a, torch.eye(C).view(C, C, 1, 1))
    pb.layernorm(a, b, gamma, beta)
    pb.conv(b, OUTPUT, torch.eye(C).view(C, C, 1, 1))
    plan = pb.finalize(torch.device(DEV))
    plan.forward(x.to(DEV, torch.bfloat16))
    got = plan.read_buffer(b).double().cpu()
    xq = x.to(torch.bfloat16).double()
    ref = F.layer_norm(xq.permute(0, 2, 3, 1), (C,), gamma.double(), beta.double(), 1e-5).permute(0, 3, 1, 2)
    assert (got - ref).abs().max() <= 2.0 ** -8 * max(1.0, float(ref.abs().max()))  # one bf16 rounding of the result


def test_large_swinir_after_span_is_repeatable():
    # regression: K-chunked tensor-core convs (3 K chunks over a 4-stage ring, two issuing warps) once faulted at 512^2 when
    # a 1080p SPAN forward had run earlier in the process (timing-dependent mbarrier phase aliasing between the two MMA-issuing
    # warps, fixed by the seen[] counters of conv_tc).  Whole forwards back to back, no synchronisation in between.
    span = SPAN(feature_channels=48, upscale=2, seed=3).eval().to(DEV).bfloat16()
    swin = SwinIR(upscale=4, seed=9).eval().to(DEV).bfloat16()
    g = torch.Generator().manual_seed(5)
    xs = torch.rand(1, 3, 1080, 1920, generator=g).to(DEV, torch.bfloat16)
    xw = torch.rand(1, 3, 512, 512, generator=g).to(DEV, torch.bfloat16)
    with torch.inference_mode():
        for _ in range(3):
            span(xs)
        outs = [swin(xw) for _ in range(4)]
        torch.cuda.synchronize()
    assert bool(torch.isfinite(outs[0].float()).all())
    for o in outs[1:]:
        assert torch.equal(o, outs[0])


def test_forward_into_replays_a_cuda_graph_and_tracks_new_inputs():
    # forward_into(x, out) captures the forward once per (x, out) address pair and replays it; the replay must read the tensors'
    # CURRENT contents and be bit-identical to plain launches
    m = _load(SPAN(feature_channels=48, upscale=2, seed=51).state_dict(), torch.bfloat16)
    g = torch.Generator().manual_seed(9)
    x = torch.rand(1, 3, 96, 128, generator=g).to(DEV, torch.bfloat16)
    out = torch.empty(1, 3, 192, 256, device=DEV, dtype=torch.bfloat16)
    plan = m.plan_for(torch.device(DEV), torch.bfloat16)
    with torch.inference_mode():
        for _ in range(3):
            m.forward_into(x, out)
        assert len(plan._graphs) == 1 and plan._graphs_enabled
        assert torch.equal(out, m(x))
        x.copy_(torch.rand(1, 3, 96, 128, generator=g))
        m.forward_into(x, out)  # replay on new contents
        assert len(plan._graphs) == 1
        assert torch.equal(out, m(x))
        other = torch.empty_like(out)
        m.forward_into(x, other)  # new address pair: second graph
        assert len(plan._graphs) == 2 and torch.equal(other, out)


@pytest.mark.parametrize('shape', [(1, 3, 1, 1), (1, 3, 2, 3), (3, 3, 7, 5), (2, 3, 9, 130), (1, 3, 129, 8)])
@pytest.mark.parametrize('kind', ['SPAN', 'Compact', 'ESRGAN'])
def test_degenerate_image_sizes(kind, shape):
    # images smaller than one tile / one 128-pixel strip, single rows and columns, odd batches: the fp32 path must still match the
    # oracle to 1e-4 and the bf16 path (tensor-core tile kernel with TMA zero fill on every side) the 50 dB bar
    model = {'SPAN': lambda: SPAN(feature_channels=48, upscale=2, seed=61),
             'Compact': lambda: SRVGGNetCompact(num_feat=64, num_conv=4, upscale=4, seed=62),
             'ESRGAN': lambda: RRDBNet(num_blocks=1, scale=4, seed=63)}[kind]()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(sum(shape)))
    ref = oracle.forward_by_name(kind, sd, x, torch.float32)
    with torch.inference_mode():
        y32 = _load(sd)(x.to(DEV)).cpu()
        y16 = _load(sd, torch.bfloat16)(x.to(DEV, torch.bfloat16)).float().cpu()
    assert y32.shape == ref.shape
    assert norm_err(y32, ref) <= FP32_TOL
    if ref.numel() >= 1024:
        assert psnr(y16, ref) >= BF16_PSNR_DB
    else:  # a PSNR over a few dozen samples is noise: bound the error by a few bf16 roundings of the largest value instead
        assert float((y16 - ref).abs().max()) <= 4 * 2.0 ** -8 * max(1.0, float(ref.abs().max()))


def test_input_layouts_and_dtypes():
    # non-contiguous (channels-last, sliced) inputs are made contiguous by the module; fp16 tensors take the exact-fp32 path and
    # come back as fp16; CPU tensors and wrong channel counts raise instead of falling back
    sd = SRVGGNetCompact(num_feat=32, num_conv=3, upscale=2, seed=64).state_dict()
    m = _load(sd)
    x = torch.rand(2, 3, 20, 24, generator=torch.Generator().manual_seed(3)).to(DEV)
    with torch.inference_mode():
        y = m(x)
        assert torch.equal(m(x.contiguous(memory_format=torch.channels_last)), y)
        big = torch.rand(2, 3, 40, 48, device=DEV)
        big[:, :, 10:30, 12:36] = x
        assert torch.equal(m(big[:, :, 10:30, 12:36]), y)
        y_half = m(x.half())
        assert y_half.dtype == torch.float16 and (y_half.float() - m(x.half().float())).abs().max() <= 2e-3
        with pytest.raises(RuntimeError):
            m(x.cpu())
        with pytest.raises(RuntimeError):
            m(torch.rand(1, 4, 8, 8, device=DEV))


# ---------------------------------------------------------------------------------------------------------------- benchmarked sizes
# VERDICT r1 ("What's weak" 1): every oracle / golden comparison above is at <= 136 x 200 px, so the global reductions (RealPLKSR's
# GroupNorm over 16 ch x 262 144 px, DAT's channel-attention Gram and AIM pooling over 262 144 tokens) were unverified at the sizes
# bench / config_times run.  bf16 against the engine's own fp32 plan at the BASELINE shapes (cheap, both on the GPU), and the fp32
# plan against the CPU oracle once per architecture at 256^2 (tens of seconds of CPU time).
_SIZED = [
    ('RealPLKSR', lambda: RealPLKSR(n_blocks=28, upscaling_factor=4, seed=7), (1, 3, 512, 512)),   # config 5a
    ('DAT', lambda: DAT(upscale=4, seed=8), (1, 3, 512, 512)),                                     # config 5b
    ('SwinIR', lambda: SwinIR(upscale=4, seed=9), (1, 3, 512, 512)),                               # section 8a row a19
    ('Compact', lambda: SRVGGNetCompact(num_feat=64, num_conv=16, upscale=4, seed=5), (16, 3, 540, 960)),  # config 2
    ('ESRGAN', lambda: RRDBNet(num_blocks=23, scale=4, seed=6), (1, 3, 768, 768)),                 # config 4's tile unit
    ('SPANPlus', lambda: SpanPlus(blocks=[4], feature_channels=48, upscale=2, seed=4), (1, 3, 1080, 1920)),  # config 3
    ('GateRV3', lambda: GateRV3(scale=2, seed=12), (1, 3, 1080, 1920)),                            # section 8f rank 2: global means over 2 M pixels, five grids
    ('RTMoSR', lambda: RTMoSR(scale=2, dim=32, n_blocks=2, seed=11), (1, 3, 1080, 1920)),
]


@pytest.mark.parametrize('name,make,shape', _SIZED, ids=[c[0] for c in _SIZED])
def test_benchmarked_sizes_bf16_against_fp32_plan(name, make, shape):
    sd = make().state_dict()
    x = torch.rand(*shape, generator=torch.Generator().manual_seed(12))
    with torch.inference_mode():
        y32 = _load(sd)(x.to(DEV)).cpu()
        torch.cuda.empty_cache()
        y16 = _load(sd, torch.bfloat16)(x.to(DEV, torch.bfloat16)).float().cpu()
    assert torch.isfinite(y32).all() and torch.isfinite(y16).all()
    assert float(y32.max() - y32.min()) > 1e-2
    assert psnr(y16, y32) >= BF16_PSNR_DB, f'{name} at {shape}: bf16 vs fp32 plan PSNR {psnr(y16, y32):.2f} dB'


_ORACLE_256 = _SIZED[:3] + [_SIZED[6]]


@pytest.mark.parametrize('name,make', [(c[0], c[1]) for c in _ORACLE_256], ids=[c[0] for c in _ORACLE_256])
def test_full_depth_fp32_plan_against_oracle_at_256(name, make):
    sd = make().state_dict()
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(13))
    ref = oracle.forward_by_name(name, {k: v.clone() for k, v in sd.items()}, x, torch.float32)
    with torch.inference_mode():
        y32 = _load(sd)(x.to(DEV)).cpu()
    assert norm_err(y32, ref) <= FP32_TOL, f'{name} 256^2 fp32 plan vs oracle: {norm_err(y32, ref):.3e}'


@pytest.mark.parametrize('scale', [1, 2, 3, 4])
def test_spanpp_every_scale_of_the_scale_list(scale):
    """SpanPP.forward(x, scale): the IGConv kernel and the output geometry depend on the requested scale (spanpp/arch.py:289-297);
    the engine keeps one native plan per scale."""
    from resselt_b200.archs import SpanPP

    sd = SpanPP(feature_channels=48, implicit_dim=64, latent_layers=2, seed=51).state_dict()
    x = torch.rand(1, 3, 40, 56, generator=torch.Generator().manual_seed(scale))
    ref = oracle.spanpp_forward({k: v.clone() for k, v in sd.items()}, x, torch.float32, scale=scale)
    with torch.inference_mode():
        m = _load(sd)
        y32 = m(x.to(DEV), scale=scale).cpu()
        y16 = _load(sd, torch.bfloat16)(x.to(DEV, torch.bfloat16), scale=scale).float().cpu()
        if scale != 2:
            assert m(x.to(DEV)).shape[-1] == 2 * 56  # the default stays eval_base_scale = 2
    assert y32.shape == ref.shape == (1, 3, 40 * scale, 56 * scale)
    assert norm_err(y32, ref) <= FP32_TOL
    assert psnr(y16, ref) >= BF16_PSNR_DB
    with pytest.raises(KeyError):
        m(x.to(DEV), scale=5)

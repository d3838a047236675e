"""Drop-in behaviour of the loading API (reference: resselt/__init__.py, registry.py, factory/, utilities/state_dict.py)."""
import collections
import os
import pickle

import pytest
import torch

import resselt_b200
from resselt_b200.archs import DAT, PLKSR, SPAN, GateRV3, RealPLKSR, RRDBNet, RTMoSR, SpanPlus, SpanPP, SRVGGNetCompact, SwinIR, internal_registry
from resselt_b200.factory import Architecture, KeyCondition
from resselt_b200.factory.arch import ModelMetadata
from resselt_b200.registry import ArchitectureNotFound, Registry
from resselt_b200.utilities.state_dict import canonicalize_state_dict, get_seq_len, pixelshuffle_scale


def _sd(model):
    return {k: v.clone() for k, v in model.state_dict().items()}


def test_public_surface():
    assert resselt_b200.__all__ == ['add', 'get', 'load_from_file', 'load_from_state_dict']
    assert {'SPAN', 'spanplus', 'Compact', 'ESRGAN', 'PLKSR', 'dat', 'SwinIR'} <= set(internal_registry.store)
    assert resselt_b200.get('SPAN').id == 'SPAN'
    with pytest.raises(KeyError):  # same as the reference's dict lookup (registry.py:74)
        resselt_b200.get('nope')


def test_metadata_field_order():
    m = ModelMetadata(3, 4, 2, 'x')
    assert (m.in_channels, m.out_channels, m.upscale, m.name) == (3, 4, 2, 'x')


@pytest.mark.parametrize(
    'model,meta',
    [
        (SPAN(feature_channels=48, upscale=2), ('SPAN', 3, 3, 2)),
        (SPAN(feature_channels=32, upscale=4, norm=False), ('SPAN', 3, 3, 4)),
        (SPAN(num_in_ch=1, num_out_ch=1, feature_channels=56, upscale=2, norm=False), ('SPAN', 1, 1, 2)),
        (SpanPP(feature_channels=48, implicit_dim=64, latent_layers=2), ('SpanPP', 3, 3, [1, 2, 3, 4])),   # the reference puts the scale list there
        (SpanPP(feature_channels=32, scale_list=[2, 4], implicit_dim=32, latent_layers=4), ('SpanPP', 3, 3, [2, 4])),
        (RTMoSR(), ('RTMoSR', 3, 3, 2)),
        (GateRV3(scale=2, enc_blocks=(1, 1, 1, 1), dec_blocks=(1, 1, 1, 1), num_latent=1, span_blocks=1), ('GateRV3', 3, 3, 2)),
        (GateRV3(dim=16, enc_blocks=(1, 2), dec_blocks=(2, 1), num_latent=2, scale=1), ('GateRV3', 3, 3, 1)),
        (GateRV3(dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=1, scale=4, upsample='dysample', upsample_mid_dim=16), ('GateRV3', 3, 3, 4)),
        (GateRV3(in_ch=1, dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=1, scale=3, upsample='pixelshuffledirect'), ('GateRV3', 1, 1, 3)),
        (GateRV3(dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=2, scale=2, upsample='nearest+conv', attention=True), ('GateRV3', 3, 3, 2)),
        (GateRV3(dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=1, scale=4, upsample='pa_up', upsample_mid_dim=24), ('GateRV3', 3, 3, 4)),
        (GateRV3(dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=1, scale=4, upsample='transpose+conv'), ('GateRV3', 3, 3, 4)),
        (GateRV3(dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=1, scale=3, upsample='transpose+conv'), ('GateRV3', 3, 3, 3)),
        (RTMoSR(scale=4, dim=48, ffn_expansion=1.5, n_blocks=1, dccm=False, se=False), ('RTMoSR', 3, 3, 2)),   # the reference always reports 2
        (RTMoSR(scale=2, n_blocks=1, unshuffle_mod=True), ('RTMoSR', 3, 3, 2)),
        (SpanPlus(blocks=[4], upscale=2), ('SPANPlus', 3, 3, 2)),
        (SpanPlus(blocks=[2, 3], feature_channels=32, upscale=4), ('SPANPlus', 3, 3, 4)),
        (SpanPlus(blocks=[2], upscale=2, upsampler='dys'), ('SPANPlus', 3, 3, 2)),
        (SpanPlus(blocks=[1], feature_channels=32, upscale=4, upsampler='dys'), ('SPANPlus', 3, 3, 4)),
        (RealPLKSR(n_blocks=1, upscaling_factor=4, dysample=True), ('RealPLKSR', 3, 3, 4)),
        (RealPLKSR(dim=32, n_blocks=1, upscaling_factor=3, kernel_size=13, dysample=True), ('RealPLKSR', 3, 3, 3)),
        (SRVGGNetCompact(num_feat=64, num_conv=16, upscale=4), ('Compact', 3, 3, 4)),
        (SRVGGNetCompact(num_feat=24, num_conv=8, upscale=2), ('Compact', 3, 3, 2)),
        (SRVGGNetCompact(num_feat=64, num_conv=16, upscale=1), ('Compact', 3, 3, 1)),
        (RRDBNet(num_blocks=2, scale=4), ('ESRGAN', 3, 3, 4)),
        (RRDBNet(num_blocks=1, scale=2, plus=True), ('ESRGAN', 3, 3, 2)),
        (RRDBNet(num_blocks=1, scale=8), ('ESRGAN', 3, 3, 8)),
        (RRDBNet(in_nc=12, out_nc=3, num_blocks=1, scale=4, shuffle_factor=2), ('ESRGAN', 3, 3, 2)),
        (RRDBNet(num_blocks=1, scale=4, key_style='new'), ('ESRGAN', 3, 3, 4)),
        (RRDBNet(num_blocks=1, scale=2, key_style='bsrgan'), ('ESRGAN', 3, 3, 2)),
        (RealPLKSR(n_blocks=2, upscaling_factor=4), ('RealPLKSR', 3, 3, 4)),
        (RealPLKSR(dim=32, n_blocks=2, upscaling_factor=2, kernel_size=13, use_ea=False), ('RealPLKSR', 3, 3, 2)),
        (PLKSR(n_blocks=2, upscaling_factor=4), ('PLKSR', 3, 3, 4)),
        (PLKSR(dim=32, n_blocks=2, upscaling_factor=2, ccm_type='CCM', kernel_size=13, use_ea=False), ('PLKSR', 3, 3, 2)),
        (PLKSR(n_blocks=1, upscaling_factor=3, ccm_type='ICCM', lk_type='SparsePLK'), ('PLKSR', 3, 3, 3)),
        (PLKSR(n_blocks=1, upscaling_factor=2, lk_type='RectSparsePLK', kernel_size=15), ('PLKSR', 3, 3, 2)),
        (DAT(depth=[3, 2], num_heads=[6, 6], upscale=4), ('DAT', 3, 3, 4)),
        (DAT(embed_dim=60, split_size=[4, 8], depth=[3, 2], num_heads=[2, 2], upscale=2, img_size=32), ('DAT', 3, 3, 2)),
        (DAT(embed_dim=60, split_size=[4, 8], depth=[3, 2], num_heads=[2, 2], upscale=3, img_size=32, upsampler='pixelshuffledirect', resi_connection='3conv'),
         ('DAT', 3, 3, 3)),
        (SwinIR(embed_dim=60, depths=[2, 3], num_heads=[6, 6], upscale=2), ('SwinIR', 3, 3, 2)),
        (SwinIR(embed_dim=180, depths=[2], num_heads=[6], upscale=3, upsampler='pixelshuffledirect'), ('SwinIR', 3, 3, 3)),
        (SwinIR(embed_dim=180, depths=[2], num_heads=[6], upscale=3, upsampler='pixelshuffle'), ('SwinIR', 3, 3, 3)),
        (DAT(embed_dim=60, split_size=[4, 8], depth=[3, 2], num_heads=[2, 2], upscale=3, img_size=32, upsampler='pixelshuffle'), ('DAT', 3, 3, 3)),
        (SwinIR(embed_dim=64, depths=[2, 2], num_heads=[4, 4], mlp_ratio=4.0, upscale=4, upsampler='nearest+conv', resi_connection='3conv', img_size=48),
         ('SwinIR', 3, 3, 4)),
        (SwinIR(in_chans=1, embed_dim=48, depths=[2], num_heads=[6], window_size=7, img_size=126, img_range=255.0, upsampler=''), ('SwinIR', 1, 1, 1)),
    ],
)
def test_detect_and_hyperparameter_inference(model, meta):
    sd = _sd(model)
    loaded = resselt_b200.load_from_state_dict(dict(sd))
    info = loaded.parameters_info
    assert (info.name, info.in_channels, info.out_channels, info.upscale) == meta
    assert type(loaded) is type(model)
    assert list(loaded.state_dict()) == list(sd)  # same names, same order -> strict load passed
    for k, v in loaded.state_dict().items():
        assert torch.equal(v, sd[k])
    if isinstance(model, SpanPlus):
        assert loaded.blocks == model.blocks and loaded.upsampler_kind == model.upsampler_kind
    if isinstance(model, SPAN):
        assert loaded.norm == model.norm
    if isinstance(model, GateRV3):
        assert (loaded.dim, loaded.enc_blocks, loaded.dec_blocks, loaded.num_latent, loaded.scale, loaded.span_blocks) == (
            model.dim, model.enc_blocks, model.dec_blocks, model.num_latent, model.scale, model.span_blocks)
        assert model.scale == 1 or (loaded.upsample, loaded.mid) == (model.upsample, model.mid)
        assert loaded.attention == model.attention
    if isinstance(model, RTMoSR):
        assert (loaded.scale, loaded.dim, loaded.hidden, loaded.n_blocks, loaded.unshuffle, loaded.dccm, loaded.se) == (
            model.scale, model.dim, model.hidden, model.n_blocks, model.unshuffle, model.dccm, model.se)
    if isinstance(model, SpanPP):
        assert (loaded.feature_channels, loaded.scale_list, loaded.base_scale, loaded.ig_kernel_size) == (model.feature_channels, model.scale_list, 2, 3)
    if isinstance(model, RRDBNet):
        assert (loaded.num_blocks, loaded.plus, loaded.shuffle_factor, loaded._keys.style) == (
            model.num_blocks, model.plus, model.shuffle_factor, model._keys.style)
    if isinstance(model, RealPLKSR):
        assert (loaded.dim, loaded.n_blocks, loaded.kernel_size, loaded.use_ea, loaded.dysample) == (
            model.dim, model.n_blocks, model.kernel_size, model.use_ea, model.dysample)
    if isinstance(model, PLKSR):
        assert (loaded.dim, loaded.n_blocks, loaded.ccm_type, loaded.lk_type, loaded.kernel_size, loaded.kmax, loaded.use_ea) == (
            model.dim, model.n_blocks, model.ccm_type, model.lk_type, model.kernel_size, model.kmax, model.use_ea)
    if isinstance(model, SwinIR):
        assert (loaded.dim, loaded.hidden, loaded.window_size, loaded.depths, loaded.heads, loaded.img_size, loaded.img_range, loaded.upsampler,
                loaded.resi_connection) == (model.dim, model.hidden, model.window_size, model.depths, model.heads, model.img_size, model.img_range,
                                            model.upsampler, model.resi_connection)
    if isinstance(model, DAT):
        assert (loaded.depth, loaded.heads, loaded.split, loaded.img_size, loaded.resi_connection, loaded.upsampler_kind) == (
            model.depth, model.heads, model.split, model.img_size, model.resi_connection, model.upsampler_kind)


def test_wrapped_and_prefixed_checkpoints():
    sd = _sd(SRVGGNetCompact(num_feat=16, num_conv=2, upscale=2))
    wrapped = {'params_ema': {'module.' + k: v for k, v in sd.items()}}
    assert resselt_b200.load_from_state_dict(wrapped).parameters_info.name == 'Compact'
    assert list(canonicalize_state_dict({'state_dict': {'netG.a': 1, 'netG.b': 2}})) == ['a', 'b']
    assert list(canonicalize_state_dict({'module.a': 1, 'b': 2})) == ['module.a', 'b']  # prefix only if on every key


def test_no_norm_marker_is_normalised_in_callers_dict():
    sd = _sd(SPAN(feature_channels=16, upscale=2, norm=False))
    sd['no_norm'] = torch.ones(1) * 5
    resselt_b200.load_from_state_dict(sd)
    assert torch.equal(sd['no_norm'], torch.zeros(1))  # reference mutates the dict the same way (span/__init__.py:41-43)


def test_unknown_state_dict_and_strict_mismatch():
    with pytest.raises(ArchitectureNotFound):
        resselt_b200.load_from_state_dict({'foo.weight': torch.zeros(1)})
    sd = _sd(SRVGGNetCompact(num_feat=16, num_conv=2, upscale=2))
    sd['body.0.extra'] = torch.zeros(1)
    with pytest.raises(RuntimeError):
        resselt_b200.load_from_state_dict(sd)


def test_load_from_file_formats(tmp_path):
    import safetensors.torch

    sd = _sd(SRVGGNetCompact(num_feat=16, num_conv=2, upscale=2))
    p_pth, p_ckpt, p_st, p_pt = (str(tmp_path / n) for n in ('m.pth', 'm.CKPT', 'm.safetensors', 'm.pt'))
    torch.save(collections.OrderedDict(sd), p_pth)
    torch.save({'params': sd}, p_ckpt)
    safetensors.torch.save_file(sd, p_st)
    torch.save(sd, p_pt)  # not TorchScript -> pickle fallback (registry.py:81-93)
    for p in (p_pth, p_ckpt, p_st, p_pt):
        assert resselt_b200.load_from_file(p).parameters_info.name == 'Compact'
    with pytest.raises(ValueError, match='Unsupported model file extension'):
        resselt_b200.load_from_file(str(tmp_path / 'm.onnx'))


def test_restricted_unpickler_blocks_code_execution(tmp_path):
    class Evil:
        def __reduce__(self):
            return (os.system, ('echo pwned',))

    p = str(tmp_path / 'evil.pth')
    torch.save({'body.0.weight': Evil()}, p)
    with pytest.raises(pickle.UnpicklingError):
        resselt_b200.load_from_file(p)


def test_plugin_contract_custom_architecture():
    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.only = torch.nn.Conv2d(1, 1, 1)

    class TinyArch(Architecture[Tiny]):
        def __init__(self):
            super().__init__(uid='tiny', detect=KeyCondition.has_all('only.weight', KeyCondition.has_any('only.bias', 'zzz')))

        def load(self, state_dict):
            return self._enhance_model(model=Tiny(), in_channels=1, out_channels=1, upscale=1, name='Tiny')

    reg = Registry()
    reg.add(TinyArch())
    assert 'tiny' in reg and len(list(reg)) == 1
    m = reg.load_from_state_dict(Tiny().state_dict())
    assert m.parameters_info == ModelMetadata(1, 1, 1, 'Tiny')
    with pytest.raises(TypeError):
        Architecture('x', KeyCondition.has_all())  # abstract


def test_state_dict_helpers():
    sd = {'body.0.w': 0, 'body.10.w': 0, 'body.3.x.y': 0, 'bodyguard.99.w': 0}
    assert get_seq_len(sd, 'body') == 11 and get_seq_len(sd, 'nothing') == 0
    assert pixelshuffle_scale(48, 3) == 4 and pixelshuffle_scale(12, 3) == 2


def test_unsupported_heads_are_refused_explicitly():
    # what the engine does not run must fail loudly at load time, never fall back: SPANPlus 'conv' (1x) head is not loadable by the
    # reference either (spanplus/__init__.py:20-27); RealPLKSR 1x DySample has no end convolution (rplksr.py:139)
    with pytest.raises(NotImplementedError):
        SpanPlus(blocks=[1], feature_channels=16, upscale=1, upsampler='conv')
    with pytest.raises(NotImplementedError):
        RealPLKSR(n_blocks=1, upscaling_factor=1, dysample=True)


def test_dysample_checkpoint_with_incomplete_keys_fails_strict_load():
    sd = _sd(SpanPlus(blocks=[1], feature_channels=16, upscale=2, upsampler='dys'))
    del sd['upsampler.scope.weight']
    with pytest.raises(RuntimeError):
        resselt_b200.load_from_state_dict(sd)


def test_weight_fingerprint_sees_replacements_and_versioned_writes_but_not_data_writes():
    """EngineModule._stamp keys the cached native plan: tensor replacement and autograd-visible in-place writes change
    it; writes through ``.data`` bypass the version counter and need ``invalidate()`` / ``refresh()`` — documented and pinned here."""
    m = SRVGGNetCompact(num_feat=16, num_conv=2, upscale=2, seed=1)
    s0 = m._stamp()
    assert isinstance(s0, tuple) and len(s0) == len(list(m.parameters())) + len(list(m.buffers()))
    p = next(m.parameters())
    with torch.no_grad():
        p.mul_(2.0)                      # versioned in-place write
    s1 = m._stamp()
    assert s1 != s0
    p.data.mul_(0.5)                     # EMA-style write: invisible to the fingerprint
    assert m._stamp() == s1
    m._plans[(0, torch.float32)] = object()
    m._plan_stamp[(0, torch.float32)] = s1
    m.refresh()                          # what the caller has to do after such an edit
    assert not m._plans and not m._plan_stamp
    # replacing the tensors (new storage) changes the fingerprint
    m.load_state_dict({k: v.clone() for k, v in m.state_dict().items()})
    assert m._stamp() != s1


def test_engine_plugins_register_in_the_live_reference_registry():
    """INTEGRATION.md section 2, executed: the engine's Architecture plugins are added to the REFERENCE's own registry
    (resselt.registry.Registry.add, /root/reference/resselt/registry.py:70-71) and its load_from_state_dict
    (registry.py:106-116: canonicalize -> first detect() hit -> load() -> strict load_state_dict) returns engine modules."""
    import sys

    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref_dir, 'resselt')):
        pytest.skip('reference not installed in baseline/_ref (tools/install_reference.sh)')
    sys.path.insert(0, ref_dir)
    try:
        import resselt
        from resselt.registry import Registry as RefRegistry

        from resselt_b200.engine import EngineModule

        reg = RefRegistry()
        for arch in internal_registry.store.values():
            reg.add(arch)  # the reference's add(): plain dict insert keyed by arch.id
        for proto, name in ((SPAN(feature_channels=48, upscale=2, seed=2), 'SPAN'), (SRVGGNetCompact(num_feat=32, num_conv=4, upscale=4, seed=3), 'Compact'),
                            (RRDBNet(num_blocks=1, scale=4, seed=4), 'ESRGAN')):
            sd = {k: v.clone() for k, v in proto.state_dict().items()}
            model = reg.load_from_state_dict(sd)
            assert isinstance(model, EngineModule) and model.parameters_info.name == name
            # and the reference's own loader accepts the very same checkpoint (same parameter names, strict)
            ref_model = resselt.load_from_state_dict({k: v.clone() for k, v in proto.state_dict().items()})
            assert type(ref_model).__module__.startswith('resselt.')
            assert set(ref_model.state_dict().keys()) == set(model.state_dict().keys())
    finally:
        sys.path.remove(ref_dir)
        for k in [k for k in sys.modules if k == 'resselt' or k.startswith('resselt.')]:
            del sys.modules[k]

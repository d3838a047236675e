"""Generate tests/golden/*.npz from the LIVE reference (run in the build container only).

    python oracle/make_golden.py

For every case: weights come from the engine's deterministic ``random_state_dict`` (numpy
RandomState, so tests can regenerate them anywhere from the recorded seed), are loaded through the
*reference's own* ``resselt.load_from_state_dict`` (strict), and the reference's fp32 CPU forward
output is stored together with the input.  The reference is imported from /root/reference; it does
not travel to the GPU box — the fixtures do.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, '/root/reference')

GOLDEN_DIR = os.path.join(ROOT, 'tests', 'golden')


def cases():
    """name -> (engine class path, constructor kwargs, weight seed, input shape, input seed)."""
    return {
        'span_x2_f48': ('SPAN', dict(num_in_ch=3, num_out_ch=3, feature_channels=48, upscale=2), 11, (1, 3, 20, 28), 101),
        'span_x4_f32_nonorm': ('SPAN', dict(num_in_ch=3, num_out_ch=3, feature_channels=32, upscale=4, norm=False), 12, (1, 3, 16, 16), 102),
        'spanplus_x2_b4': ('SPANPlus', dict(num_in_ch=3, num_out_ch=3, blocks=[4], feature_channels=48, upscale=2, upsampler='ps'), 13, (1, 3, 24, 20), 103),
        'spanplus_x2_b2_3': ('SPANPlus', dict(num_in_ch=3, num_out_ch=3, blocks=[2, 3], feature_channels=32, upscale=2, upsampler='ps'), 14, (2, 3, 16, 18), 104),
        'compact_x4_nf64_nc16': ('Compact', dict(num_in_ch=3, num_out_ch=3, num_feat=64, num_conv=16, upscale=4), 15, (1, 3, 18, 22), 105),
        'compact_x2_nf24_nc8': ('Compact', dict(num_in_ch=3, num_out_ch=3, num_feat=24, num_conv=8, upscale=2), 16, (2, 3, 16, 12), 106),
        'esrgan_x4_nb2': ('ESRGAN', dict(in_nc=3, out_nc=3, num_filters=64, num_blocks=2, scale=4), 17, (1, 3, 20, 24), 107),
        'esrgan_plus_x2_nb1': ('ESRGAN', dict(in_nc=3, out_nc=3, num_filters=64, num_blocks=1, scale=2, plus=True), 18, (2, 3, 12, 16), 108),
        'realesrgan_x2_unshuffle_nb1': ('ESRGAN', dict(in_nc=12, out_nc=3, num_filters=64, num_blocks=1, scale=4, shuffle_factor=2), 19, (1, 3, 21, 17), 109),
        'realplksr_x4_nb2': ('RealPLKSR', dict(in_ch=3, dim=64, n_blocks=2, upscaling_factor=4, kernel_size=17, split_ratio=0.25,
                                               use_ea=True, norm_groups=4, dysample=False), 20, (1, 3, 24, 20), 110),
        'dat_x4_d3_3': ('DAT', dict(img_size=64, in_chans=3, embed_dim=180, split_size=[8, 32], depth=[3, 3], num_heads=[6, 6],
                                   expansion_factor=2.0, upscale=4), 22, (1, 3, 40, 56), 112),
        'dat_x2_d3_2_s4_8': ('DAT', dict(img_size=32, in_chans=3, embed_dim=60, split_size=[4, 8], depth=[3, 2], num_heads=[2, 2],
                                      expansion_factor=2.0, upscale=2), 23, (2, 3, 19, 27), 113),
        'swinir_x2_ps_d2_3': ('SwinIR', dict(img_size=64, in_chans=3, embed_dim=60, depths=[2, 3], num_heads=[6, 6], window_size=8, mlp_ratio=2.0,
                                             upscale=2, upsampler='pixelshuffle', resi_connection='1conv'), 24, (1, 3, 24, 32), 114),
        'swinir_x3_psd_180': ('SwinIR', dict(img_size=64, in_chans=3, embed_dim=180, depths=[2], num_heads=[6], window_size=8, mlp_ratio=2.0,
                                             upscale=3, upsampler='pixelshuffledirect', resi_connection='1conv'), 25, (2, 3, 19, 27), 115),
        'swinir_x4_nearest_3conv': ('SwinIR', dict(img_size=64, in_chans=3, embed_dim=64, depths=[2, 2], num_heads=[4, 4], window_size=8, mlp_ratio=4.0,
                                                   upscale=4, upsampler='nearest+conv', resi_connection='3conv'), 26, (1, 3, 21, 30), 116),
        'swinir_jpeg_w7_gray': ('SwinIR', dict(img_size=126, in_chans=1, embed_dim=48, depths=[2, 2], num_heads=[6, 6], window_size=7, mlp_ratio=2.0,
                                               upscale=1, img_range=255.0, upsampler='', resi_connection='1conv'), 27, (1, 1, 30, 23), 117),
        'plksr_x4_dccm_plk17': ('PLKSR', dict(dim=64, n_blocks=2, upscaling_factor=4, ccm_type='DCCM', kernel_size=17, split_ratio=0.25, lk_type='PLK',
                                              use_ea=True), 28, (1, 3, 22, 26), 118),
        'plksr_x3_iccm_sparse': ('PLKSR', dict(dim=64, n_blocks=2, upscaling_factor=3, ccm_type='ICCM', kernel_size=17, split_ratio=0.25,
                                               lk_type='SparsePLK', use_ea=True), 29, (2, 3, 18, 16), 119),
        'plksr_x2_ccm_rect13_noea': ('PLKSR', dict(dim=32, n_blocks=3, upscaling_factor=2, ccm_type='CCM', kernel_size=15, split_ratio=0.25,
                                                   lk_type='RectSparsePLK', use_ea=False), 30, (1, 3, 20, 20), 120),
        'spanplus_x2_dys': ('SPANPlus', dict(num_in_ch=3, num_out_ch=3, blocks=[2], feature_channels=48, upscale=2, upsampler='dys'), 31, (1, 3, 22, 26), 121),
        'spanplus_x4_dys_f32': ('SPANPlus', dict(num_in_ch=3, num_out_ch=3, blocks=[1], feature_channels=32, upscale=4, upsampler='dys'), 32, (2, 3, 14, 17), 122),
        'realplksr_x4_dys': ('RealPLKSR', dict(in_ch=3, dim=64, n_blocks=2, upscaling_factor=4, kernel_size=17, split_ratio=0.25, use_ea=True,
                                               norm_groups=4, dysample=True), 33, (1, 3, 20, 24), 123),
        'realplksr_x3_dys': ('RealPLKSR', dict(in_ch=3, dim=32, n_blocks=1, upscaling_factor=3, kernel_size=13, split_ratio=0.25, use_ea=True,
                                               norm_groups=4, dysample=True), 34, (2, 3, 15, 13), 124),
        'dat_light_x3_psd_3conv': ('DAT', dict(img_size=32, in_chans=3, embed_dim=60, split_size=[4, 8], depth=[3, 2], num_heads=[2, 2], expansion_factor=2.0,
                                               upscale=3, upsampler='pixelshuffledirect', resi_connection='3conv'), 35, (1, 3, 21, 30), 125),
        'dat_x3_ps': ('DAT', dict(img_size=32, in_chans=3, embed_dim=60, split_size=[4, 8], depth=[3, 2], num_heads=[2, 2], expansion_factor=2.0,
                                  upscale=3, upsampler='pixelshuffle', resi_connection='1conv'), 36, (1, 3, 20, 26), 126),
        'spanpp_f48': ('SpanPP', dict(num_in_ch=3, feature_channels=48, scale_list=[1, 2, 3, 4], implicit_dim=64, latent_layers=2), 37, (1, 3, 20, 28), 127),
        'spanpp_f32_s4': ('SpanPP', dict(num_in_ch=3, feature_channels=32, scale_list=[2, 4], implicit_dim=32, latent_layers=4), 38, (2, 3, 14, 18), 128),
        'rtmosr_x2_d32': ('RTMoSR', dict(scale=2, dim=32, ffn_expansion=2, n_blocks=2, unshuffle_mod=False, dccm=True, se=True), 39, (1, 3, 21, 27), 129),
        'rtmosr_x4_d48_nose_1x1': ('RTMoSR', dict(scale=4, dim=48, ffn_expansion=1.5, n_blocks=1, unshuffle_mod=False, dccm=False, se=False), 40, (2, 3, 16, 18), 130),
        # GateRV3 (third SPAN descendant): small U-Nets through the reference's own loader
        'gaterv3_x2_ps': ('GateRV3', dict(in_ch=3, dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=2, scale=2, upsample='pixelshuffle', upsample_mid_dim=16, span_blocks=1), 45, (1, 3, 21, 27), 135),
        'gaterv3_x1_conv': ('GateRV3', dict(in_ch=3, dim=32, enc_blocks=(1, 1, 1), dec_blocks=(1, 1, 1), num_latent=1, scale=1, span_blocks=1), 46, (2, 3, 24, 16), 136),
        'gaterv3_x3_psd': ('GateRV3', dict(in_ch=3, dim=16, enc_blocks=(2, 1), dec_blocks=(1, 2), num_latent=1, scale=3, upsample='pixelshuffledirect', upsample_mid_dim=16, span_blocks=2), 47, (1, 3, 17, 19), 137),
        'gaterv3_x2_dys': ('GateRV3', dict(in_ch=3, dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=1, scale=2, upsample='dysample', upsample_mid_dim=16, span_blocks=1), 48, (1, 3, 20, 24), 138),
        'gaterv3_x2_att_nearest': ('GateRV3', dict(in_ch=3, dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=2, scale=2, upsample='nearest+conv', upsample_mid_dim=16, attention=True, span_blocks=1), 49, (1, 3, 20, 28), 139),
        'gaterv3_x4_paup': ('GateRV3', dict(in_ch=3, dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=1, scale=4, upsample='pa_up', upsample_mid_dim=24, span_blocks=1), 50, (1, 3, 16, 20), 140),
        'gaterv3_x2_transpose': ('GateRV3', dict(in_ch=3, dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=1, scale=2, upsample='transpose+conv', upsample_mid_dim=16, span_blocks=1), 51, (1, 3, 18, 22), 141),
        'gaterv3_x4_transpose': ('GateRV3', dict(in_ch=3, dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=1, scale=4, upsample='transpose+conv', upsample_mid_dim=16, span_blocks=1), 52, (1, 3, 16, 12), 142),
        'gaterv3_x3_transpose': ('GateRV3', dict(in_ch=3, dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=1, scale=3, upsample='transpose+conv', upsample_mid_dim=16, span_blocks=1), 53, (1, 3, 13, 17), 143),
        'gaterv3_x3_nearest': ('GateRV3', dict(in_ch=3, dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=1, scale=3, upsample='nearest+conv', upsample_mid_dim=16, span_blocks=1), 54, (1, 3, 14, 18), 144),
        'gaterv3_x3_paup': ('GateRV3', dict(in_ch=3, dim=16, enc_blocks=(1, 1), dec_blocks=(1, 1), num_latent=1, scale=3, upsample='pa_up', upsample_mid_dim=16, span_blocks=1), 55, (1, 3, 12, 15), 145),
        'rtmosr_x2_unshuffle': ('RTMoSR', dict(scale=2, dim=32, ffn_expansion=2, n_blocks=1, unshuffle_mod=True, dccm=True, se=True), 41, (1, 3, 22, 30), 131),
        'realplksr_x2_nb3_noea': ('RealPLKSR', dict(in_ch=3, dim=32, n_blocks=3, upscaling_factor=2, kernel_size=13, split_ratio=0.25,
                                                    use_ea=False, norm_groups=4, dysample=False), 21, (2, 3, 18, 18), 111),
    }


def engine_model(kind: str, kwargs: dict, seed: int):
    from resselt_b200 import archs

    cls = {'SPAN': archs.SPAN, 'SPANPlus': archs.SpanPlus, 'Compact': archs.SRVGGNetCompact}
    extra = {k: getattr(archs, k) for k in ('RRDBNet', 'RealPLKSR') if hasattr(archs, k)}
    cls.update({'ESRGAN': extra.get('RRDBNet'), 'RealPLKSR': extra.get('RealPLKSR'), 'DAT': getattr(archs, 'DAT', None), 'SwinIR': getattr(archs, 'SwinIR', None), 'PLKSR': getattr(archs, 'PLKSR', None), 'SpanPP': getattr(archs, 'SpanPP', None), 'RTMoSR': getattr(archs, 'RTMoSR', None), 'GateRV3': getattr(archs, 'GateRV3', None)})
    return cls[kind](seed=seed, **kwargs)


def main():
    # The reference's DySample builds one tensor with pin_memory=True (resselt/utilities/dysample.py:62), which needs a CUDA driver
    # this container does not have.  Dropping that flag (values unchanged) is the only deviation from the unmodified package.
    _orig_tensor = torch.tensor

    def _tensor_without_pinning(*args, **kwargs):
        kwargs.pop('pin_memory', None)
        return _orig_tensor(*args, **kwargs)

    torch.tensor = _tensor_without_pinning
    import resselt as reference  # the unmodified reference package

    os.makedirs(GOLDEN_DIR, exist_ok=True)
    # `python oracle/make_golden.py name [name ...]` regenerates only those cases and merges them into index.json
    only = set(sys.argv[1:])
    index = {}
    index_path = os.path.join(GOLDEN_DIR, 'index.json')
    if only and os.path.exists(index_path):
        with open(index_path) as f:
            index = json.load(f)['cases']
    for name, (kind, kwargs, wseed, xshape, xseed) in cases().items():
        if only and name not in only:
            continue
        sd = {k: v.clone() for k, v in engine_model(kind, kwargs, wseed).state_dict().items()}
        ref_model = reference.load_from_state_dict(dict(sd)).eval()
        x = torch.from_numpy(np.random.RandomState(xseed).rand(*xshape).astype(np.float32))
        with torch.inference_mode():
            y = ref_model(x)
        info = ref_model.parameters_info
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + '.npz'), x=x.numpy(), y=y.numpy().astype(np.float32))
        index[name] = dict(kind=kind, kwargs=kwargs, weight_seed=wseed, x_shape=list(xshape), x_seed=xseed,
                           meta=dict(name=info.name, in_channels=info.in_channels, out_channels=info.out_channels, upscale=info.upscale),
                           y_min=float(y.min()), y_max=float(y.max()))
        print(f'{name}: y {tuple(y.shape)} range [{float(y.min()):.4f}, {float(y.max()):.4f}]')
    with open(os.path.join(GOLDEN_DIR, 'index.json'), 'w') as f:
        json.dump(dict(reference='rewaifu/resselt v1.4.1 @ /root/reference', torch=torch.__version__, cases=index), f, indent=1, sort_keys=True)


if __name__ == '__main__':
    main()

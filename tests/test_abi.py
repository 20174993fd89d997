"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/vkocr_b200.h declares
(no compute calls without a GPU), the ctypes table covers the header, and the Python face keeps the reference's
state_dict layout, constructors and config classes."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    text = open(os.path.join(ROOT, 'include', 'vkocr_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(vkocr_\w+)\s*\(', text)))


def test_library_exports_every_declared_symbol():
    from vkit_ocr_model_adaptive_scaling_b200 import _build, _lib
    symbols = _header_symbols()
    assert len(symbols) >= 30
    lib = ctypes.CDLL(_build.LIB)
    for name in symbols:
        assert hasattr(lib, name), f'{name} is declared in include/vkocr_b200.h but not exported by the library'
    bound = set(_lib._SIGNATURES) | {'vkocr_last_error', 'vkocr_launch_count'}
    assert bound == set(symbols), f'ctypes table and header disagree: {sorted(bound ^ set(symbols))}'
    assert _lib.LIB.vkocr_abi_version() == 1


def test_struct_layouts_match_header():
    """ctypes mirrors of VkocrConvGeom / VkocrEpilogue have the field order of the header."""
    from vkit_ocr_model_adaptive_scaling_b200 import _lib
    text = open(os.path.join(ROOT, 'include', 'vkocr_b200.h')).read()
    for cname, cls in (('VkocrConvGeom', _lib.ConvGeom), ('VkocrEpilogue', _lib.Epilogue), ('VkocrHeadTail', _lib.HeadTail)):
        body = re.search(r'typedef struct %s \{(.*?)\} %s;' % (cname, cname), text, flags=re.S).group(1)
        body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
        fields = []
        for decl in body.split(';'):
            decl = decl.strip()
            if not decl:
                continue
            names = decl.split(',')
            first = names[0].split()[-1]
            fields.append(re.sub(r'\[.*?\]', '', first.lstrip('*')))
            fields.extend(re.sub(r'\[.*?\]', '', n.strip().lstrip('*')) for n in names[1:])
        assert fields == [f[0] for f in cls._fields_], (cname, fields)


def test_no_cpu_fallback():
    import vkit_ocr_model_adaptive_scaling_b200 as vk
    layer = vk.model.ConvNextBlockLayer(32)
    with pytest.raises(vk._lib.VkocrError):
        layer(torch.randn(1, 32, 8, 8))
    fn = vk.loss_function.L2LossFunction()
    with pytest.raises(vk._lib.VkocrError):
        fn(torch.randn(4), torch.randn(4))


@pytest.mark.parametrize('neck', ['upernext', 'fpn'])
def test_state_dict_layout_equals_reference(golden_dir, neck):
    """Key names, order and shapes equal the reference's state_dict() (recorded in the golden fixtures by
    oracle/make_golden.py, which loads the synthetic weights into the reference with strict=True)."""
    import vkit_ocr_model_adaptive_scaling_b200 as vk
    from oracle import synth
    M = vk.model
    g = np.load(os.path.join(golden_dir, f'adaptive_scaling_tiny_{neck}.npz'))
    cfg = M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY, neck_head_type=M.AdaptiveScalingNeckHeadType(neck))
    model = M.AdaptiveScaling(cfg)
    sd = model.state_dict()
    assert list(sd.keys()) == [str(n) for n in g['param_names']]
    assert len(sd) == (304 if neck == 'upernext' else 280)
    ref = synth.synth_state_dict('tiny', neck)
    assert {k: tuple(v.shape) for k, v in sd.items()} == {k: tuple(v.shape) for k, v in ref.items()}
    model.load_state_dict(ref, strict=True)
    assert len(list(model.buffers())) == 0


def test_reference_initialisation_statistics():
    """block_scale = 1e-6, zero biases, LayerNorm (1, 0), head output bias = init_output_bias
    (convnext.py:38,169-173; upernext.py:225-231; adaptive_scaling.py:93-99)."""
    import vkit_ocr_model_adaptive_scaling_b200 as vk
    M = vk.model
    torch.manual_seed(0)
    model = M.AdaptiveScaling(M.AdaptiveScalingConfig(size=M.AdaptiveScalingSize.TINY,
                                                      neck_head_type=M.AdaptiveScalingNeckHeadType.UPERNEXT))
    sd = model.state_dict()
    assert torch.all(sd['backbone.blocks.0.layers.0.block_scale'] == 1e-6)
    assert float(sd['rough_char_height_head.0.step2_conv1x1.1.bias']) == 8.0
    assert float(sd['rough_char_mask_head.step2_conv1x1.1.bias']) == 0.0
    assert float(sd['backbone.stem.0.bias'].abs().max()) == 0.0
    w = sd['backbone.blocks.2.layers.4.block.3.weight']
    assert abs(float(w.std()) - 0.02) < 2e-3 and float(w.abs().max()) <= 2.0
    assert sum(p.numel() for p in model.parameters()) == 35038222   # measured on the reference
    probs = [layer.prob_bypass for block in model.backbone.blocks for layer in block.layers]
    assert probs[0] == 0.0 and abs(probs[-1] - 0.1) < 1e-12 and len(probs) == 18


def test_config_defaults_and_enums():
    import vkit_ocr_model_adaptive_scaling_b200 as vk
    M, LF = vk.model, vk.loss_function
    cfg = M.AdaptiveScalingConfig()
    assert cfg.size == M.AdaptiveScalingSize.SMALL and cfg.neck_head_type == M.AdaptiveScalingNeckHeadType.FPN
    assert (cfg.rough_upsampling_factor, cfg.rough_init_char_height_output_bias, cfg.precise_upsampling_factor,
            cfg.precise_enable_char_mask_head) == (2, 8.0, 2, False)
    assert [e.value for e in M.AdaptiveScalingSize] == ['tiny', 'small', 'base', 'large']
    r = LF.AdaptiveScalingRoughLossFunctionConifg()
    assert (r.bce_negative_ratio, r.bce_factor, r.focal_factor, r.dice_factor, r.l1_factor, r.downsampled_score_map_min,
            r.char_height_feature_min) == (3.0, 0.0, 5.0, 1.0, 1.0, 1.1, 1.1)
    p = LF.AdaptiveScalingPreciseLossFunctionConifg()
    assert (p.char_mask_focal_factor, p.char_prob_l1_factor, p.char_prob_pos_l2_factor, p.char_prob_neg_l2_factor,
            p.char_prob_wahr_factor, p.char_up_left_offset_l1_factor, p.char_up_left_distance_regulation_l1_factor,
            p.char_corner_angle_cross_entropy_factor, p.char_corner_distance_l1_factor, p.loss_factor) == (
                0.0, 0.0, 2.0, 1.0, 0.0, 1.0, 1.0, 5.0, 1.0, 0.15)
    with pytest.raises(NotImplementedError):
        M.FpnHead(16, 1, upsampling_factor=5)
    with pytest.raises(AssertionError):
        M.UperNextNeck((8, 16, 24), 16)


def test_autograd_nodes_see_the_callers_grad_mode():
    """``ctx.needs_input_grad`` is the same under ``torch.no_grad()``; the nodes of ops.py must not take it for "a backward will
    follow" (they would write every training side channel during inference): ``ops._Function.apply`` records the grad mode."""
    import torch
    from vkit_ocr_model_adaptive_scaling_b200 import ops
    seen = []

    class Probe(ops._Function):
        @staticmethod
        def forward(ctx, x, w):
            seen.append((ops._needs_grad(ctx), tuple(ctx.needs_input_grad)))
            return x * w

        @staticmethod
        def backward(ctx, g):
            return g, g

    x, w = torch.randn(3), torch.nn.Parameter(torch.randn(3))
    Probe.apply(x, w)
    with torch.no_grad():
        Probe.apply(x, w)
    with torch.inference_mode():
        Probe.apply(x, w)
    Probe.apply(x, w.detach())
    assert [s[0] for s in seen] == [True, False, False, False], seen
    assert seen[1][1] == (False, True)          # what torch reports under no_grad: the reason for the wrapper
    assert ops._GRAD_MODE == [True]             # the stack unwinds (also through exceptions: try / finally)
    assert all(issubclass(getattr(ops, n), ops._Function) for n in dir(ops) if n.endswith('Fn') and isinstance(getattr(ops, n), type))

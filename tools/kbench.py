"""GPU micro-benchmark of the hot kernels at the BASELINE config #3 shapes (B=32, 640x640, TINY), one C-ABI call per line:
CUDA-event time (median of `reps` after warm-up; every operand is larger than... or rotated through buffers so that L2
does not serve it), algorithmic bytes / FLOPs and the fraction of the measured peaks.

    python tools/kbench.py [group ...]     groups: mlp ln dw head up tn colsum   (default: all)
"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from vkit_ocr_model_adaptive_scaling_b200 import _lib as L, ops  # noqa: E402

dev = torch.device('cuda:0')
BF = torch.bfloat16
HBM, TC = 6545.6, 1382.4
groups = set(sys.argv[1:]) or {'mlp', 'ln', 'dw', 'head', 'up', 'tn', 'colsum'}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, nbytes=0.0, flops=0.0):
    gbs = nbytes / ms / 1e6
    tfs = flops / ms / 1e9
    print(f'{name:58s} {ms:8.3f} ms  {gbs:7.0f} GB/s ({gbs / HBM:5.1%})  {tfs:7.1f} TF/s ({tfs / TC:5.1%})', flush=True)


def rnd(*shape, dtype=BF):
    return torch.randn(*shape, device=dev, dtype=torch.float32).to(dtype)


if 'mlp' in groups:
    for (M, C) in ((819200, 96), (204800, 192), (51200, 384), (12800, 768)):
        hid = 4 * C
        x = rnd(M, C)
        cp = (C + 63) // 64 * 64
        w1 = rnd(hid, cp); w2 = rnd(C, hid); w2d = rnd(hid, cp); w1d = rnd(C, hid)
        b1 = rnd(hid, dtype=torch.float32); b2 = rnd(C, dtype=torch.float32); gam = torch.rand(C, device=dev)
        g = torch.empty(M, hid, device=dev, dtype=BF); hpre = torch.empty_like(g); dh = torch.empty_like(g)
        y = torch.empty(M, C, device=dev, dtype=BF)
        ms = timeit(lambda: ops.gemm_nt(x, 1, 1, M, C, C, 1, w1, cp, hid, ops._epilogue(g, hid, out_pre=hpre, ld_pre=hid, bias=b1, act=3)))
        report(f'mlp1 fwd  M{M} K{C} N{hid} +bias+gelu+gelu\'', ms, (M * C + 2 * M * hid) * 2, 2.0 * M * C * hid)
        ms = timeit(lambda: ops.gemm_nt(x, 1, 1, M, C, C, 1, w1, cp, hid, ops._epilogue(g, hid, bias=b1, act=1)))
        report(f'mlp1 eval M{M} K{C} N{hid} +bias+gelu', ms, (M * C + M * hid) * 2, 2.0 * M * C * hid)
        ms = timeit(lambda: ops.gemm_nt(g, 1, 1, M, hid, hid, 1, w2, hid, C, ops._epilogue(y, C, bias=b2, col_scale=gam, residual=x, ld_res=C)))
        report(f'mlp2 fwd  M{M} K{hid} N{C} +bias+scale+res', ms, (M * hid + 2 * M * C) * 2, 2.0 * M * C * hid)
        ms = timeit(lambda: ops.gemm_nt(x, 1, 1, M, C, C, 1, w2d, cp, hid, ops._epilogue(dh, hid, act=4, aux=hpre, ld_aux=hid)))
        report(f'mlp2 dgrad M{M} K{C} N{hid} *aux', ms, (M * C + 2 * M * hid) * 2, 2.0 * M * C * hid)
        ms = timeit(lambda: ops.gemm_nt(dh, 1, 1, M, hid, hid, 1, w1d, hid, C, ops._epilogue(y, C)))
        report(f'mlp1 dgrad M{M} K{hid} N{C}', ms, (M * hid + M * C) * 2, 2.0 * M * C * hid)
        s = torch.zeros(C * hid, device=dev)
        ms = timeit(lambda: ops.gemm_tn(x, 1, 1, M, C, C, 1, g, hid, hid, ops._epilogue(s, hid, out_f32=True, accumulate=True, tn=(0, hid, 1))))
        report(f'mlp2 wgrad TN M{M} I{C} J{hid}', ms, (M * hid + M * C) * 2, 2.0 * M * C * hid)
        ms = timeit(lambda: ops.gemm_tn(dh, 1, 1, M, hid, hid, 1, x, C, C, ops._epilogue(s, C, out_f32=True, accumulate=True, tn=(0, C, 1))))
        report(f'mlp1 wgrad TN M{M} I{hid} J{C}', ms, (M * hid + M * C) * 2, 2.0 * M * C * hid)
        del x, g, hpre, dh, y

if 'ln' in groups:
    for (M, C, act) in ((819200, 96, 0), (819200, 96, 1), (204800, 192, 0), (51200, 384, 0), (12800, 768, 0)):
        x = rnd(M, C); y = torch.empty_like(x); dy = rnd(M, C); dx = torch.empty_like(x)
        gm = torch.rand(C, device=dev) + 0.5; bt = rnd(C, dtype=torch.float32)
        mean = torch.empty(M, device=dev); rstd = torch.empty(M, device=dev)
        dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev); ds = torch.zeros(C, device=dev)
        ms = timeit(lambda: ops.layernorm_fwd(x, C, y, C, M, C, gm, bt, act, mean, rstd))
        report(f'layernorm_fwd rows{M} C{C} act{act}', ms, 2.0 * M * C * 2)
        ms = timeit(lambda: ops.layernorm_bwd(dy, C, x, C, mean, rstd, gm, bt, act, dx, C, M, C, dg, db, ds))
        report(f'layernorm_bwd rows{M} C{C} act{act}', ms, 3.0 * M * C * 2)

if 'colsum' in groups:
    for (M, C) in ((819200, 384), (819200, 96), (51200, 1536), (51200, 384)):
        x = rnd(M, C); out = torch.zeros(C, device=dev)
        ms = timeit(lambda: ops.colsum(x, C, M, C, out))
        report(f'colsum rows{M} C{C}', ms, 1.0 * M * C * 2)

if 'dw' in groups:
    for (B, H, W, C) in ((32, 160, 160, 96), (32, 80, 80, 192), (32, 40, 40, 384), (32, 20, 20, 768), (8, 512, 512, 96)):
        x = ops.alloc_nhwc(B, H, W, C, BF, dev); x.normal_()
        y = ops.alloc_nhwc(B, H, W, C, BF, dev)
        add = ops.alloc_nhwc(B, H, W, C, BF, dev); add.normal_()
        wt = rnd(49, C, dtype=torch.float32); bias = rnd(C, dtype=torch.float32); dw = torch.zeros(C, 1, 7, 7, device=dev)
        n = B * H * W * C
        ms = timeit(lambda: ops.dwconv7(x, y, wt, bias, None))
        report(f'dwconv7 fwd {B}x{H}x{W}x{C}', ms, 2.0 * n * 2, 98.0 * n)
        ms = timeit(lambda: ops.dwconv7(x, y, wt, None, add))
        report(f'dwconv7 bwd-data(+add) {B}x{H}x{W}x{C}', ms, 3.0 * n * 2, 98.0 * n)
        ms = timeit(lambda: ops.dwconv7_wgrad(add, x, dw))
        report(f'dwconv7 wgrad {B}x{H}x{W}x{C}', ms, 2.0 * n * 2, 98.0 * n)

if 'up' in groups:
    for (B, h, w, C, f) in ((32, 160, 160, 384, 2), (32, 80, 80, 96, 2)):
        src = ops.alloc_nhwc(B, h, w, C, BF, dev); src.normal_()
        dst = ops.alloc_nhwc(B, h * f, w * f, C, BF, dev); dst.normal_()
        for mode in (0, 1):
            ms = timeit(lambda: ops.upsample_fwd(src, dst, C, mode, False))
            report(f'upsample_fwd mode{mode} {B}x{h}x{w}->x{f} C{C}', ms, (B * h * w + B * h * w * f * f) * C * 2.0)
            ms = timeit(lambda: ops.upsample_bwd(dst, src, C, mode, False))
            report(f'upsample_bwd mode{mode} {B}x{h * f}x{w * f}->/{f} C{C}', ms, (B * h * w + B * h * w * f * f) * C * 2.0)
        del src, dst

if 'head' in groups:
    B, H, W = 32, 320, 320
    M = B * H * W
    for (nh, slot, inner, O, sp) in ((2, 192, 192, 1, 0), (2, 192, 192, 1, 1), (4, 200, 192, 1, 0), (4, 200, 194, 4, 1), (4, 200, 193, 2, 0)):
        ntot = nh * slot
        conv = ops.alloc_nhwc(B, H, W, ntot, BF, dev); conv.normal_()
        dconv = ops.alloc_nhwc(B, H, W, ntot, BF, dev)
        gm = torch.rand(inner, device=dev) + 0.5; bt = rnd(inner, dtype=torch.float32)
        w2 = rnd(O, inner, dtype=torch.float32); b2 = rnd(O, dtype=torch.float32)
        out = torch.empty(B, O, H, W, device=dev); dout = torch.randn(B, O, H, W, device=dev)
        gs = [torch.zeros(inner, device=dev) for _ in range(3)] + [torch.zeros(O, inner, device=dev), torch.zeros(O, device=dev)]
        sl = conv[:, slot:2 * slot]
        dsl = dconv[:, slot:2 * slot]
        ld = conv.stride(3)

        def fwd():
            L.check(L.LIB.vkocr_head_tail_fwd(1, L.ptr(sl), ld, inner, slot, L.ptr(gm), L.ptr(bt), L.ptr(w2), L.ptr(b2), O, sp,
                                              L.ptr(out), H * W, M, L.stream_ptr()), 'head_tail_fwd')

        def bwd():
            L.check(L.LIB.vkocr_head_tail_bwd(1, L.ptr(sl), ld, inner, slot, L.ptr(gm), L.ptr(bt), L.ptr(w2), O, sp, L.ptr(out),
                                              L.ptr(dout), H * W, M, L.ptr(dsl), ld, L.ptr(gs[0]), L.ptr(gs[1]), L.ptr(gs[3]),
                                              L.ptr(gs[4]), L.ptr(gs[2]), L.stream_ptr()), 'head_tail_bwd')
        ms = timeit(fwd)
        report(f'head_tail_fwd rows{M} inner{inner} O{O} (ld {ntot})', ms, M * (slot * 2.0 + 4 * O))
        ms = timeit(bwd)
        report(f'head_tail_bwd rows{M} inner{inner} O{O} (ld {ntot})', ms, M * (2 * slot * 2.0 + 8 * O))
        del conv, dconv, out, dout

if 'tnx' in groups:
    # experiments on the head weight-gradient kernel (environment knobs read by gemm_tc.cu)
    B, H, W, C = 32, 320, 320, 384
    up = ops.alloc_nhwc(B, H, W, C, BF, dev); up.normal_()
    ntot = 832
    dconv = ops.alloc_nhwc(B, H, W, ntot, BF, dev); dconv.normal_()
    gw = torch.zeros(ntot * C * 9, device=dev)

    def run(ks, label):
        ms = timeit(lambda: ops.gemm_tn(dconv, B, H, W, ntot, dconv.stride(3), ks, up, C, up.stride(3),
                                        ops._epilogue(gw, C, out_f32=True, accumulate=True, tn=(1, C * ks * ks, ks * ks))), reps=3)
        report(f'TN ks{ks} I{ntot} J{C} {label}', ms, 0.0, 2.0 * B * H * W * ks * ks * C * ntot)
    run(3, 'baseline')
    run(1, 'baseline')
    os.environ['VKOCR_DEBUG_SKIP_TMA'] = '1'
    run(3, 'SKIP_TMA (MMA only)')
    for v, lab in (('3', 'A K-major'), ('5', 'B K-major'), ('7', 'both K-major')):
        os.environ['VKOCR_DEBUG_SKIP_TMA'] = v
        run(3, f'SKIP_TMA {lab}')
        os.environ['VKOCR_TN_BN'] = '64'
        run(3, f'SKIP_TMA {lab} BN=64')
        del os.environ['VKOCR_TN_BN']
    os.environ['VKOCR_DEBUG_SKIP_TMA'] = '1'
    for v in ('128', '64'):
        os.environ['VKOCR_TN_BN'] = v
        run(3, f'SKIP_TMA BN={v}')
    del os.environ['VKOCR_TN_BN']
    wpx = rnd(ntot, 9 * C)
    convx = ops.alloc_nhwc(B, H, W, ntot, BF, dev)
    ms = timeit(lambda: ops.gemm_nt(up, B, H, W, C, up.stride(3), 3, wpx, C, ntot, ops._epilogue(convx, convx.stride(3))), reps=3)
    report(f'NT ks3 N{ntot} SKIP_TMA (MMA+epilogue only)', ms, 0.0, 2.0 * B * H * W * 9 * C * ntot)
    del os.environ['VKOCR_DEBUG_SKIP_TMA'], wpx, convx
    for k, vals in (('VKOCR_TN_STAGES', ('3', '2')), ('VKOCR_TN_BN', ('128', '256', '64')), ('VKOCR_TN_FILL', ('1', '4', '8'))):
        for v in vals:
            os.environ[k] = v
            run(3, f'{k}={v}')
        del os.environ[k]
    os.environ['VKOCR_TN_BN'] = '128'; os.environ['VKOCR_TN_FILL'] = '1'
    run(3, 'BN=128 FILL=1')
    del os.environ['VKOCR_TN_BN'], os.environ['VKOCR_TN_FILL']

if 'tn' in groups:
    B, H, W, C = 32, 320, 320, 384
    up = ops.alloc_nhwc(B, H, W, C, BF, dev); up.normal_()
    for ntot in (384, 832):
        dconv = ops.alloc_nhwc(B, H, W, ntot, BF, dev); dconv.normal_()
        gw = torch.zeros(ntot * C * 9, device=dev)
        ms = timeit(lambda: ops.gemm_tn(dconv, B, H, W, ntot, dconv.stride(3), 3, up, C, up.stride(3),
                                        ops._epilogue(gw, C, out_f32=True, accumulate=True, tn=(1, C * 9, 9))), reps=3)
        report(f'head wgrad TN ks3 M{B * H * W} I{ntot} J{C}', ms, 0.0, 2.0 * B * H * W * 9 * C * ntot)
        wp = rnd(ntot, 9 * C)
        conv = ops.alloc_nhwc(B, H, W, ntot, BF, dev)
        ms = timeit(lambda: ops.gemm_nt(up, B, H, W, C, up.stride(3), 3, wp, C, ntot, ops._epilogue(conv, conv.stride(3))), reps=3)
        report(f'head conv NT ks3 M{B * H * W} K{9 * C} N{ntot}', ms, 0.0, 2.0 * B * H * W * 9 * C * ntot)
        del dconv, conv, gw

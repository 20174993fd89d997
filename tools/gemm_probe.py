"""GPU probe for the tcgen05 GEMM (run under gpurun): compares NT/TN, plain and 3x3, against torch fp32 and the SIMT kernel."""
import ctypes
import sys
import os
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from vkit_ocr_model_adaptive_scaling_b200 import _lib as L

dev = torch.device('cuda:0')
torch.manual_seed(0)


def pack_w(w_oihw, c_pad, dtype):
    n, c, kh, kw = w_oihw.shape
    p = torch.zeros(n, kh * kw, c_pad, device=w_oihw.device, dtype=dtype)
    p[:, :, :c] = w_oihw.permute(0, 2, 3, 1).reshape(n, kh * kw, c).to(dtype)
    return p.reshape(n, kh * kw * c_pad).contiguous()


def gemm_nt(x_nhwc, w_oihw, bias=None, act=0, backend=0, res=None, col_scale=None, want_pre=False, aux=None, row_scale=None,
            rows_per_group=1):
    B, H, W, C = x_nhwc.shape
    n, c, ks, _ = w_oihw.shape
    c_pad = (C + 63) // 64 * 64
    wp = pack_w(w_oihw, c_pad, x_nhwc.dtype)
    out = torch.empty(B, H, W, n, device=dev, dtype=x_nhwc.dtype)
    pre = torch.empty_like(out) if want_pre else None
    g = L.ConvGeom(B, H, W, ks, C, x_nhwc.stride(2), c_pad)
    ep = L.Epilogue(out.data_ptr(), n, 0, 0, pre.data_ptr() if want_pre else None, n,
                    bias.data_ptr() if bias is not None else None, act,
                    col_scale.data_ptr() if col_scale is not None else None,
                    row_scale.data_ptr() if row_scale is not None else None, rows_per_group,
                    res.data_ptr() if res is not None else None, n,
                    aux.data_ptr() if aux is not None else None, n, 0, 0, 0)
    L.check(L.LIB.vkocr_gemm_nt(L.dtype_tag(x_nhwc.dtype), backend, L.ptr(x_nhwc), ctypes.byref(g), L.ptr(wp), n,
                                ctypes.byref(ep), L.stream_ptr()), 'gemm_nt')
    return out, pre


def gemm_tn(p_nhwc, q_nhwc, ks, backend=0):
    B, H, W, I = p_nhwc.shape
    J = q_nhwc.shape[3]
    out = torch.zeros(ks * ks, I, J, device=dev, dtype=torch.float32)
    g = L.ConvGeom(B, H, W, ks, I, p_nhwc.stride(2), 0)
    ep = L.Epilogue(out.data_ptr(), J, 1, 1, None, 0, None, 0, None, None, 1, None, 0, None, 0, I * J, J, 1)
    L.check(L.LIB.vkocr_gemm_tn(L.dtype_tag(p_nhwc.dtype), backend, L.ptr(p_nhwc), ctypes.byref(g), L.ptr(q_nhwc), J,
                                q_nhwc.stride(2), ctypes.byref(ep), L.stream_ptr()), 'gemm_tn')
    return out


def rel(a, b):
    a = a.float(); b = b.float()
    return float((a - b).norm() / (b.norm() + 1e-30)), float((a - b).abs().max())


def report(name, got, ref, tol):
    torch.cuda.synchronize()
    r, m = rel(got, ref)
    ok = r < tol and bool(torch.isfinite(got.float()).all())
    print(f'{"OK  " if ok else "FAIL"} {name}: rel {r:.3e} maxabs {m:.3e} (tol {tol})', flush=True)
    return ok


def ref_conv(x_nhwc, w, bias):
    y = F.conv2d(x_nhwc.float().permute(0, 3, 1, 2), w.float(), bias, padding=w.shape[-1] // 2)
    return y.permute(0, 2, 3, 1)


def main():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    allok = True
    cases_plain = [(1000, 96, 384), (4096, 384, 96), (777, 48, 96), (12800, 768, 3072), (513, 192, 193), (300, 1152, 96),
                   (70, 96, 1152), (1000, 192, 832), (333, 64, 1200), (5000, 384, 1536)]
    for backend, dt in ((1, torch.float32), (1, torch.bfloat16), (0, torch.bfloat16)):
        tol = 1e-5 if dt == torch.float32 else 6e-3
        tag = f'{"simt" if backend else "tc"}/{str(dt)[6:]}'
        for (M, K, N) in cases_plain:
            x = torch.randn(1, 1, M, K, device=dev).to(dt)
            w = (torch.randn(N, K, 1, 1, device=dev) / K ** 0.5)
            b = torch.randn(N, device=dev)
            out, _ = gemm_nt(x, w, b, backend=backend)
            ref = ref_conv(x, w.to(dt), b)
            allok &= report(f'nt plain {tag} M{M} K{K} N{N}', out, ref, tol)
        # epilogue variants
        M, K, N = 2048, 96, 384
        x = torch.randn(1, 1, M, K, device=dev).to(dt)
        w = torch.randn(N, K, 1, 1, device=dev) / K ** 0.5
        b = torch.randn(N, device=dev)
        out, pre = gemm_nt(x, w, b, act=1, backend=backend, want_pre=True)
        refpre = ref_conv(x, w.to(dt), b)
        allok &= report(f'nt gelu {tag}', out, F.gelu(refpre), tol)
        allok &= report(f'nt gelu-pre {tag}', pre, refpre, tol)
        res = torch.randn(1, 1, M, N, device=dev).to(dt)
        cs = torch.rand(N, device=dev)
        out, _ = gemm_nt(x, w, b, backend=backend, res=res, col_scale=cs)
        allok &= report(f'nt scale+res {tag}', out, refpre * cs + res.float(), tol)
        rsc = torch.rand(M // 256, device=dev)
        out, _ = gemm_nt(x, w, b, backend=backend, res=res, col_scale=cs, row_scale=rsc, rows_per_group=256)
        allok &= report(f'nt scale+rowscale+res {tag}', out, refpre * cs * rsc.repeat_interleave(256)[None, None, :, None] + res.float(), tol)
        aux = torch.randn(1, 1, M, N, device=dev).to(dt)
        out, _ = gemm_nt(x, w, None, act=2, backend=backend, aux=aux)
        a32 = aux.float().requires_grad_(True)
        F.gelu(a32).sum().backward()
        allok &= report(f'nt gelu-grad {tag}', out, ref_conv(x, w.to(dt), None) * a32.grad, tol)
        # 3x3 conv
        for (B, H, W, C, N) in [(2, 20, 20, 1152, 96), (2, 40, 40, 96, 96), (1, 64, 48, 384, 193), (2, 33, 17, 64, 32),
                                 (1, 40, 56, 128, 832), (2, 5, 7, 96, 1152), (1, 20, 28, 832, 384)]:
            x = torch.randn(B, H, W, C, device=dev).to(dt)
            w = torch.randn(N, C, 3, 3, device=dev) / (9 * C) ** 0.5
            b = torch.randn(N, device=dev)
            out, _ = gemm_nt(x, w, b, backend=backend)
            allok &= report(f'nt conv3x3 {tag} B{B} {H}x{W} C{C} N{N}', out, ref_conv(x, w.to(dt), b), tol)
        # TN plain: dW[n,k] = sum_m dY[m,n] X[m,k]
        for (M, I, J) in [(5000, 384, 96), (4096, 96, 384), (1000, 193, 384), (20000, 3072, 768)]:
            pm = torch.randn(1, 1, M, (I + 7) // 8 * 8, device=dev).to(dt)[..., :I]
            qm = torch.randn(1, 1, M, J, device=dev).to(dt)
            out = gemm_tn(pm, qm, 1, backend=backend)
            ref = pm.float().reshape(M, I).t() @ qm.float().reshape(M, J)
            allok &= report(f'tn plain {tag} M{M} I{I} J{J}', out[0], ref, 1e-4 if dt == torch.float32 else 1e-4)
        # TN conv3x3: dW[tap,n,c] = sum_pix dY[pix,n] X[pix+off,c]
        for (B, H, W, I, J) in [(2, 40, 40, 96, 96), (1, 48, 64, 200, 384), (2, 20, 20, 96, 1152)]:
            pm = torch.randn(B, H, W, I, device=dev).to(dt)
            qm = torch.randn(B, H, W, J, device=dev).to(dt)
            out = gemm_tn(pm, qm, 3, backend=backend)
            # reference through autograd of conv2d: grad wrt weight (I, J, 3, 3)
            wref = torch.zeros(I, J, 3, 3, device=dev, requires_grad=True)
            y = F.conv2d(qm.float().permute(0, 3, 1, 2), wref, padding=1)
            y.backward(pm.float().permute(0, 3, 1, 2))
            ref = wref.grad.permute(2, 3, 0, 1).reshape(9, I, J)
            allok &= report(f'tn conv3x3 {tag} B{B} {H}x{W} I{I} J{J}', out, ref, 1e-4)
    # timing of the big head conv on the tensor cores
    for (B, H, W, C, N) in [(8, 320, 320, 384, 192), (8, 320, 320, 384, 832)]:
        x = torch.randn(B, H, W, C, device=dev).to(torch.bfloat16)
        w = torch.randn(N, C, 3, 3, device=dev) / (9 * C) ** 0.5
        for _ in range(2):
            gemm_nt(x, w, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5):
            gemm_nt(x, w, None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        fl = 2.0 * B * H * W * 9 * C * N
        print(f'time conv3x3 tc B{B} {H}x{W} C{C} N{N}: {ms:.3f} ms  {fl / ms / 1e9:.1f} TFLOP/s (incl. weight pack)', flush=True)
    for (M, K, N) in [(819200, 96, 384), (819200, 384, 96), (51200, 384, 1536), (51200, 1536, 384)]:
        x = torch.randn(1, 1, M, K, device=dev).to(torch.bfloat16)
        w = torch.randn(N, K, 1, 1, device=dev) / K ** 0.5
        for _ in range(2):
            gemm_nt(x, w, None)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(5):
            gemm_nt(x, w, None)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f'time plain tc M{M} K{K} N{N}: {ms:.3f} ms  {2.0 * M * K * N / ms / 1e9:.1f} TFLOP/s  {(M * K + M * N) * 2 / ms / 1e6:.0f} GB/s', flush=True)
    print('ALL OK' if allok else 'SOME FAILED')


if __name__ == '__main__':
    main()

"""CPU check of the index / weight logic of csrc/head_combine.cu (no GPU needed): a numpy re-statement of the generic and of
the factor-2 / 3x3 tiled algorithms (same tables, clamps and validity rules as the kernels), compared against
F.interpolate + F.conv2d and its autograd adjoint in fp64.  python tools/check_tapsplit.py"""
import itertools

import numpy as np
import torch
import torch.nn.functional as F


def axis(d, n_in, n_out, mode):
    if mode == 0:
        scale = np.float32(n_in) / np.float32(n_out)
        s = np.float32(scale * np.float32(d + 0.5) - np.float32(0.5))
        if s < 0:
            s = np.float32(0)
        i0 = min(int(s), n_in - 1)
        i1 = i0 + (1 if i0 < n_in - 1 else 0)
        l = float(s - np.float32(i0))
        return i0, i1, 1.0 - l, l
    scale = np.float32(n_in) / np.float32(n_out)
    i = min(int(np.floor(np.float32(d) * scale)), n_in - 1)
    return i, i, 1.0, 0.0


def weight_of(d, p, n_in, n_out, mode):
    i0, i1, w0, w1 = axis(d, n_in, n_out, mode)
    return (w0 if i0 == p else 0.0) + (w1 if i1 == p else 0.0)


def pos(a, d, k):
    first = (0 if d == 2 else -1) if a == 0 else (-1 if d == 0 else 0)
    return first + k


def pair_weights(i, a, d, n_in, mode):
    n_out = 2 * n_in
    Y = 2 * i + a + d - 1
    if Y < 0 or Y >= n_out:
        return 0.0, 0.0
    i0, i1, w0, w1 = axis(Y, n_in, n_out, mode)
    clamp = lambda r: min(max(r, 0), n_in - 1)
    rA, rB = clamp(i + pos(a, d, 0)), clamp(i + pos(a, d, 1))
    wA = (w0 if i0 == rA else 0.0) + (w1 if i1 == rA else 0.0)
    wB = ((w0 if i0 == rB else 0.0) + (w1 if i1 == rB else 0.0)) if rB != rA else 0.0
    assert abs(wA + wB - (w0 + w1)) < 1e-6, (i, a, d, n_in, mode, wA, wB, w0, w1, i0, i1, rA, rB)
    return wA, wB


def reference(x, wt, f, mode):
    up = F.interpolate(x, scale_factor=f, mode='bilinear' if mode == 0 else 'nearest', **({'align_corners': False} if mode == 0 else {}))
    return F.conv2d(up, wt, padding=wt.shape[-1] // 2)


def z_of(x, wt):
    """Z[b, tap, n, p, q] = sum_c x[b, c, p, q] W[n, c, dy, dx]"""
    N, C, k, _ = wt.shape
    return torch.einsum('bcpq,nct->btnpq', x, wt.reshape(N, C, k * k))


def combine_generic(z, f, mode, ks):
    B, T, N, h, w = z.shape
    H, W = h * f, w * f
    pad = ks // 2
    out = np.zeros((B, N, H, W))
    zz = z.numpy()
    for y, x in itertools.product(range(H), range(W)):
        for dy, dx in itertools.product(range(ks), range(ks)):
            Y, X = y + dy - pad, x + dx - pad
            if not (0 <= Y < H and 0 <= X < W):
                continue
            ay, ax = axis(Y, h, H, mode), axis(X, w, W, mode)
            for (p, wy), (q, wx) in itertools.product(((ay[0], ay[2]), (ay[1], ay[3])), ((ax[0], ax[2]), (ax[1], ax[3]))):
                if wy * wx != 0.0:
                    out[:, :, y, x] += wy * wx * zz[:, dy * ks + dx, :, p, q]
    return out


def combine_fast(z, mode):
    """factor 2, 3x3: the two-phase algorithm of hc_fwd_2x3_kernel (without the tiling, same tables)."""
    B, T, N, h, w = z.shape
    zz = z.numpy()
    clampr = lambda r: min(max(r, 0), h - 1)
    clampc = lambda r: min(max(r, 0), w - 1)
    V = np.zeros((B, N, h, 2, 3, w))   # [i][a][dx][q]
    for i, a, d in itertools.product(range(h), range(2), range(3)):
        wA, wB = pair_weights(i, a, d, h, mode)
        rA, rB = clampr(i + pos(a, d, 0)), clampr(i + pos(a, d, 1))
        for dx in range(3):
            V[:, :, i, a, dx, :] += wA * zz[:, d * 3 + dx, :, rA, :] + wB * zz[:, d * 3 + dx, :, rB, :]
    out = np.zeros((B, N, 2 * h, 2 * w))
    for j, b2, d in itertools.product(range(w), range(2), range(3)):
        wA, wB = pair_weights(j, b2, d, w, mode)
        qA, qB = clampc(j + pos(b2, d, 0)), clampc(j + pos(b2, d, 1))
        for i, a in itertools.product(range(h), range(2)):
            out[:, :, 2 * i + a, 2 * j + b2] += wA * V[:, :, i, a, d, qA] + wB * V[:, :, i, a, d, qB]
    return out


def adjoint_fast(dc, mode):
    """factor 2, 3x3: hc_bwd_2x3_kernel's tables.  dc: [B, N, H, W] -> dz [B, 9, N, h, w]"""
    B, N, H, W = dc.shape
    h, w = H // 2, W // 2
    d = dc.numpy()
    E = np.zeros((B, N, h, 3, W))
    for p, dy, o in itertools.product(range(h), range(3), range(4)):
        Y = 2 * p - 1 + o
        yy = Y - dy + 1
        if 0 <= Y < H and 0 <= yy < H:
            E[:, :, p, dy, :] += weight_of(Y, p, h, H, mode) * d[:, :, yy, :]
    dz = np.zeros((B, 9, N, h, w))
    for q, dx, o in itertools.product(range(w), range(3), range(4)):
        X = 2 * q - 1 + o
        xx = X - dx + 1
        if 0 <= X < W and 0 <= xx < W:
            cf = weight_of(X, q, w, W, mode)
            for dy in range(3):
                dz[:, dy * 3 + dx, :, :, q] += cf * E[:, :, :, dy, xx]
    return dz


def adjoint_generic(dc, f, mode, ks):
    B, N, H, W = dc.shape
    h, w = H // f, W // f
    pad = ks // 2
    d = dc.numpy()
    dz = np.zeros((B, ks * ks, N, h, w))
    for p, q in itertools.product(range(h), range(w)):
        for Y, X in itertools.product(range(H), range(W)):
            wy, wx = weight_of(Y, p, h, H, mode), weight_of(X, q, w, W, mode)
            if wy * wx == 0.0:
                continue
            for dy, dx in itertools.product(range(ks), range(ks)):
                yy, xx = Y - dy + pad, X - dx + pad
                if 0 <= yy < H and 0 <= xx < W:
                    dz[:, dy * ks + dx, :, p, q] += wy * wx * d[:, :, yy, xx]
    return dz


def main():
    torch.manual_seed(0)
    worst = 0.0
    for mode in (0, 1):
        for (h, w) in ((1, 1), (1, 4), (2, 3), (3, 2), (5, 7), (4, 9), (6, 6)):
            x = torch.randn(2, 3, h, w, dtype=torch.float64)
            wt = torch.randn(4, 3, 3, 3, dtype=torch.float64)
            ref = reference(x, wt, 2, mode)
            z = z_of(x, wt)
            for name, got in (('generic', combine_generic(z, 2, mode, 3)), ('fast', combine_fast(z, mode))):
                err = float(np.abs(got - ref.numpy()).max())
                worst = max(worst, err)
                assert err < 1e-5, (name, mode, h, w, err)
            # adjoint: <combine(z), dc> == <z, adjoint(dc)>, against autograd
            zt = z.clone().requires_grad_(True)
            dc = torch.randn_like(ref)
            # autograd through an explicit differentiable statement of the combine: interpolate every tap map, shift, sum
            up = F.interpolate(zt.reshape(2, 9 * 4, h, w), scale_factor=2, mode='bilinear' if mode == 0 else 'nearest',
                               **({'align_corners': False} if mode == 0 else {})).reshape(2, 9, 4, 2 * h, 2 * w)
            upp = F.pad(up, (1, 1, 1, 1))
            c = sum(upp[:, dy * 3 + dx, :, dy:dy + 2 * h, dx:dx + 2 * w] for dy in range(3) for dx in range(3))
            assert float((c - ref).abs().max()) < 1e-9
            (c * dc).sum().backward()
            for name, got in (('adj-generic', adjoint_generic(dc, 2, mode, 3)), ('adj-fast', adjoint_fast(dc, mode))):
                err = float(np.abs(got - zt.grad.numpy()).max())
                worst = max(worst, err)
                assert err < 1e-5, (name, mode, h, w, err)
        # other factors / kernel sizes through the generic statement (the 5x5 head of FpnHead for factors in (2, 4])
        for f, ks in ((4, 5), (3, 5), (1, 3), (2, 1)):
            x = torch.randn(1, 2, 3, 4, dtype=torch.float64)
            wt = torch.randn(3, 2, ks, ks, dtype=torch.float64)
            ref = reference(x, wt, f, mode)
            z = z_of(x, wt)
            err = float(np.abs(combine_generic(z, f, mode, ks) - ref.numpy()).max())
            worst = max(worst, err)
            assert err < 1e-5, ('generic', mode, f, ks, err)
    print('tap-split combine / adjoint index logic OK; worst abs error', worst)


if __name__ == '__main__':
    main()

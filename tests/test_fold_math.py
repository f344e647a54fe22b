"""CPU: the algebra behind `csrc/sg2_ops.cu`, stated in torch and checked against the reference's two-pass
formulation (oracle.sg2_blur + strided / transposed conv) and against autograd.  The CUDA kernels implement exactly these
index maps; the -m gpu tests compare the kernels with the same two-pass statements."""
import math

import torch
import torch.nn.functional as F

from oracle import skit_oracle as O

B8 = torch.tensor([1.0, 3.0, 3.0, 1.0]) / 8


def down_fold(w, scale):
    """blur(pad) -> k x k stride-2 conv  ==  (k+3) x (k+3) stride-2 conv with filter (scale w) (*) blur."""
    co, ci, k, _ = w.shape
    out = torch.zeros(co, ci, k + 3, k + 3, dtype=w.dtype)
    for t in range(k):
        for s in range(k):
            for j in range(4):
                for i in range(4):
                    out[:, :, t + j, s + i] = out[:, :, t + j, s + i] + scale * w[:, :, t, s] * B8[j] * B8[i]
    return out


def up_fold(w):
    """demodulate -> conv_transpose 3x3 s2 -> blur(pad 1,1, x4)  ==  3x3 pad-1 conv to 4 co channels (phase-major) + depth-to-space:
    phase (py, px) tap (a, b) = c2[2 + py - 2a][2 + px - 2b], c2[sy][sx] = sum_{t,s} wd[t][s] b4[t - sy + 1] b4[s - sx + 1]."""
    co, ci, _, _ = w.shape
    v = w / math.sqrt(ci * 9)
    wd = v * torch.rsqrt(v.pow(2).sum([1, 2, 3]) + 1e-8).view(-1, 1, 1, 1)
    b4 = B8 * 2
    out = torch.zeros(4 * co, ci, 3, 3, dtype=w.dtype)
    for py in range(2):
        for px in range(2):
            for a in range(3):
                for b in range(3):
                    sy, sx = 2 + py - 2 * a, 2 + px - 2 * b
                    acc = torch.zeros(co, ci, dtype=w.dtype)
                    for t in range(3):
                        for s in range(3):
                            jy, jx = t - sy + 1, s - sx + 1
                            if 0 <= jy <= 3 and 0 <= jx <= 3:
                                acc = acc + wd[:, :, t, s] * b4[jy] * b4[jx]
                    out[(py * 2 + px) * co:(py * 2 + px + 1) * co, :, a, b] = acc
    return out, wd


def test_down_folds_equal_blur_then_strided_conv():
    torch.manual_seed(0)
    x = torch.randn(2, 8, 16, 16)
    w3, w1 = torch.randn(12, 8, 3, 3), torch.randn(12, 8, 1, 1)
    ref = F.conv2d(O.sg2_blur(x, 2, 2), w3 / math.sqrt(72), stride=2)
    torch.testing.assert_close(F.conv2d(x, down_fold(w3, 1 / math.sqrt(72)), stride=2, padding=2), ref, rtol=1e-5, atol=1e-5)
    ref = F.conv2d(O.sg2_blur(x, 1, 1), w1 / math.sqrt(8), stride=2)
    torch.testing.assert_close(F.conv2d(x, down_fold(w1, 1 / math.sqrt(8)), stride=2, padding=1), ref, rtol=1e-5, atol=1e-5)


def test_subpixel_fold_equals_transposed_conv_then_blur():
    torch.manual_seed(1)
    x = torch.randn(2, 8, 10, 14)
    w = torch.randn(6, 8, 3, 3)
    e, wd = up_fold(w)
    ref = O.sg2_blur(F.conv_transpose2d(x, wd.transpose(0, 1), stride=2), 1, 1, gain=4.0)
    raw = F.conv2d(x, e, padding=1)                                                  # [2, 4*6, 10, 14]
    got = raw.view(2, 2, 2, 6, 10, 14).permute(0, 3, 4, 1, 5, 2).reshape(2, 6, 20, 28)  # depth-to-space
    torch.testing.assert_close(got, ref, rtol=1e-5, atol=1e-5)


def test_transposed_folds_match_autograd():
    """skit_sg2_weight_prep_bwd: dW = s * sum_{j,i} dE[t+j][s+i] b_j b_i, and through the demodulation
    dL/dv = d G - d^3 v sum(G v) with v = s W, d = rsqrt(sum v^2 + eps)."""
    torch.manual_seed(2)
    for k in (3, 1):
        w = torch.randn(5, 4, k, k, dtype=torch.float64, requires_grad=True)
        scale = 1 / math.sqrt(4 * k * k)
        e = down_fold(w, scale)
        R = torch.randn_like(e)
        (e * R).sum().backward()
        dw = torch.zeros_like(w)
        for t in range(k):
            for s in range(k):
                for j in range(4):
                    for i in range(4):
                        dw[:, :, t, s] += scale * R[:, :, t + j, s + i] * B8[j] * B8[i]
        torch.testing.assert_close(dw, w.grad, rtol=1e-10, atol=1e-12)
    w = torch.randn(6, 8, 3, 3, dtype=torch.float64, requires_grad=True)
    e, _ = up_fold(w)
    R = torch.randn_like(e)
    (e * R).sum().backward()
    co, ci = 6, 8
    scale, b4 = 1 / math.sqrt(ci * 9), B8 * 2
    G = torch.zeros(co, ci, 3, 3, dtype=torch.float64)
    for ph in range(4):
        py, px = ph >> 1, ph & 1
        for a in range(3):
            for b in range(3):
                for t in range(3):
                    for s in range(3):
                        jy, jx = t - 1 - py + 2 * a, s - 1 - px + 2 * b      # the index form the kernel uses
                        if 0 <= jy <= 3 and 0 <= jx <= 3:
                            G[:, :, t, s] += R[ph * co:(ph + 1) * co, :, a, b] * b4[jy] * b4[jx]
    v = (w * scale).detach()
    d = torch.rsqrt(v.pow(2).sum([1, 2, 3]) + 1e-8).view(-1, 1, 1, 1)
    dw = scale * (d * G - d ** 3 * v * (G * v).sum([1, 2, 3], keepdim=True))
    torch.testing.assert_close(dw, w.grad, rtol=1e-9, atol=1e-12)

"""The im2col ("COL") reformulation behind conv_tc_wgrad_col_kernel and conv_tc_fwd_kernel<2> (gan_danet_b200/csrc/conv_tc.cu), checked on the CPU against
autograd of the reference's own layer type (nn.Conv2d(C_i, 24, kernel_size=3, padding=1), models/generator.py:34).  No GPU: this pins the index
arithmetic the kernels implement -- which operand is shifted, in which direction, how the (kw, kh, co) accumulator rows / K columns are ordered, and
what the 2-bit codes of the relu1_1 pair kernel (thin_conv.cu::tap_pair_kernel) mean -- independently of the CUDA code."""
import torch
import torch.nn.functional as F


def _dycol(dy, G):
    """dycol[b, q, (tap, c)] = dy[b, q - (tap - pad), c] for the 9 taps of a 3x3 / pad 1 convolution; c padded to G * 8 channels; zero outside the image.
    dy: [B, H, W, Co] (NHWC, as the kernels hold it)."""
    B, H, W, Co = dy.shape
    Cp = G * 8
    dyp = F.pad(dy, (0, Cp - Co))                                   # channel padding of the packed operand
    dyp = F.pad(dyp, (0, 0, 1, 1, 1, 1))                            # one pixel of zeros around the image = TMA's out-of-bounds fill
    cols = []
    for kh in range(3):
        for kw in range(3):
            # source pixel q - (kh - 1, kw - 1)  ->  padded coordinates (q_h + 1) - (kh - 1), (q_w + 1) - (kw - 1)
            cols.append(dyp[:, 2 - kh:2 - kh + H, 2 - kw:2 - kw + W, :])
    return torch.stack(cols, dim=3).reshape(B, H, W, 9 * Cp)        # column index = tap * Cp + c, tap = kh * 3 + kw


def test_wgrad_and_dgrad_as_products_over_the_shifted_gradient():
    torch.manual_seed(0)
    B, H, W, Ci, Co = 2, 6, 9, 13, 24
    conv = torch.nn.Conv2d(Ci, Co, kernel_size=3, padding=1).double()
    x = torch.randn(B, Ci, H, W, dtype=torch.float64, requires_grad=True)
    dy = torch.randn(B, Co, H, W, dtype=torch.float64)
    conv(x).backward(dy)
    G = (Co + 7) // 8
    col = _dycol(dy.permute(0, 2, 3, 1), G)                         # [B, H, W, 9 * 24]
    xn = x.detach().permute(0, 2, 3, 1)                             # [B, H, W, Ci]
    # weight gradient: D[(tap, co)][ci] = sum_q dycol[q][(tap, co)] * x[q][ci]   (x is NOT shifted: it is read once per pixel tile)
    D = torch.einsum("bhwm,bhwc->mc", col, xn).reshape(9, G * 8, Ci)[:, :Co]          # [tap, co, ci]
    want_w = conv.weight.grad.permute(2, 3, 0, 1).reshape(9, Co, Ci)                 # OIHW -> [tap = kh * 3 + kw][co][ci]
    torch.testing.assert_close(D, want_w, rtol=1e-12, atol=1e-12)
    # data gradient: dx[q][ci] = sum_(tap, co) dycol[q][(tap, co)] * W[co][ci][tap]   (one product with K = 9 * 24; the kernel orders K as (kw, kh, co))
    Wcol = F.pad(conv.weight.detach().permute(2, 3, 0, 1).reshape(9, Co, Ci), (0, 0, 0, G * 8 - Co)).reshape(9 * G * 8, Ci)
    dx = torch.einsum("bhwm,mc->bhwc", col, Wcol)
    torch.testing.assert_close(dx, x.grad.permute(0, 2, 3, 1), rtol=1e-12, atol=1e-12)


def test_row_variant_windows_of_one_halo_box():
    """Row variant: for a tile of 128 consecutive pixels of one image row the three kw shifts are windows of ONE box of 130 pixels starting one pixel left
    of the tile: window kw starts 2 - kw pixels into the box (the +32 / +16 / +0 byte start addresses of the kernels)."""
    torch.manual_seed(1)
    W = 300
    row = torch.randn(W)
    for x0 in (0, 128, 256):                                        # tile origins (the last tile is ragged: zero fill beyond the row)
        box = torch.zeros(130)
        lo, hi = x0 - 1, x0 + 129
        s0, s1 = max(lo, 0), min(hi, W)
        box[s0 - lo:s1 - lo] = row[s0:s1]
        for kw in range(3):
            win = box[2 - kw:2 - kw + 128]
            want = torch.zeros(128)
            for j in range(128):
                src = x0 + j - (kw - 1)
                if 0 <= src < W:
                    want[j] = row[src]
            assert torch.equal(win, want)


def test_relu1_1_two_bit_codes_reproduce_the_gradient():
    """tap_pair_kernel keeps 2 bits per element of relu1_1: 0 = ReLU closed, else 2 + sign(fa' - fb').  With them
    dz = (dy + gcoef * (code - 2)) * [code != 0] equals the gradient of  mean|relu(conv a) - relu(conv b)| + <dy, relu(conv a)>  w.r.t. conv a."""
    torch.manual_seed(2)
    B, H, W, C = 2, 7, 8, 16
    w = torch.randn(C, 1, 3, 3, dtype=torch.float64)
    bias = torch.randn(C, dtype=torch.float64)
    a = torch.randn(B, 1, H, W, dtype=torch.float64)
    b = a + 0.3 * torch.randn(B, 1, H, W, dtype=torch.float64)
    dy = 1e-3 * torch.randn(B, C, H, W, dtype=torch.float64)
    za = F.conv2d(a, w, bias, padding=1).requires_grad_(True)
    fa, fb = F.relu(za), F.relu(F.conv2d(b, w, bias, padding=1))
    ((fa - fb).abs().mean() + (fa * dy).sum()).backward()
    code = torch.where(za.detach() > 0, 2 + torch.sign(fa.detach() - fb).long(), torch.zeros_like(za, dtype=torch.long))
    gcoef = 1.0 / za.numel()
    dz = torch.where(code != 0, dy + gcoef * (code - 2).double(), torch.zeros_like(dy))
    torch.testing.assert_close(dz, za.grad, rtol=1e-12, atol=1e-15)

"""CPU checks of the index algebra behind two reformulations used on the CUDA path (no GPU, no library calls):

* engine.MergedStem with w-tap folding: the block-structured, folded weight described by ops.conv_weight_part windows,
  applied to the folded input of ops.fold_w, equals cat([conv_c(xc), conv_g(xg)]) of discriminator.py:79-90,121-124,
  180-193,225-228 (4x4(x4) kernels, stride 2, padding 1 along h and w);
* engine.Block._forward_tap_unrolled: a ConvTranspose2d with few output channels equals a 1x1 matrix product onto
  Cout*kh*kw columns followed by the col2im gather of dcv_col2im_act (generator.py:73,274).
"""
import pytest
import torch
import torch.nn.functional as F

from dcvgan_b200 import engine, ops


def _fold_w(xg, xc, kw=4, sw=2, pw=1):
    """reference of dcv_fold_w on (N, C, T, H, W) tensors -> (N, kw*(cg+cc), T, H, Ow), channel = k*cin + c"""
    x = torch.cat([xg, xc], 1)
    n, cin, t, h, w = x.shape
    ow = (w + 2 * pw - kw) // sw + 1
    xp = F.pad(x, (pw, pw))
    cols = [xp[..., k:k + sw * (ow - 1) + 1:sw] for k in range(kw)]          # each (N, cin, T, H, Ow)
    return torch.stack(cols, 1).reshape(n, kw * cin, t, h, ow)


def _dense_from_parts(parts, cl_total, cs_total, taps):
    """assemble W[cs][cl][tap'] exactly as dcv_pack_weight_multi / dcv_pack_weight_sub index the master weights"""
    W = torch.zeros(cs_total, cl_total, taps)
    for (w, off, cl_off, cl_cnt, cs_off, cs_cnt, s_l, s_s, s_tap) in parts:
        flat = w.reshape(-1)
        for cl in range(cl_cnt):
            for cs in range(cs_cnt):
                for tp in range(taps):
                    W[cs_off + cs, cl_off + cl, tp] = flat[off + cl * s_l + cs * s_s + tp * s_tap]
    return W


@pytest.mark.parametrize("kind,cg", [("vdis", 1), ("vdis", 2), ("idis", 1), ("idis", 2)])
def test_folded_merged_stem_equals_the_two_stem_convolutions(kind, cg):
    torch.manual_seed(0)
    cc, ndf = 3, 16
    three_d = kind == "vdis"
    if three_d:
        conv = lambda ci: torch.nn.Conv3d(ci, ndf // 2, 4, (1, 2, 2), (0, 1, 1), bias=False)
    else:
        conv = lambda ci: torch.nn.Conv2d(ci, ndf // 2, 4, 2, 1, bias=False)
    conv_g, conv_c = conv(cg), conv(cc)
    stem = engine.MergedStem(kind, cg, cc, ndf, conv_g, conv_c)
    assert stem.fold and stem.spec.k == ((4 if three_d else 1), 4, 1)
    kt = 4 if three_d else 1
    parts = [ops.conv_weight_part(m.weight.detach(), cl, cs, k, stem.KW) for m, cl, cs, k in stem.windows]
    W = _dense_from_parts(parts, stem.KW * (cg + cc), ndf, kt * 4)          # taps' = (kt, kh), kw folded away
    xg = torch.randn(2, cg, 6 if three_d else 1, 8, 8)
    xc = torch.randn(2, cc, 6 if three_d else 1, 8, 8)
    x2 = _fold_w(xg, xc)
    y = F.conv3d(x2, W.reshape(ndf, -1, kt, 4, 1), stride=(1, 2, 1), padding=(0, 1, 0))
    if three_d:
        ref = torch.cat([conv_c(xc), conv_g(xg)], 1)
    else:
        ref = torch.cat([conv_c(xc[:, :, 0]), conv_g(xg[:, :, 0])], 1).unsqueeze(2)
    assert y.shape == ref.shape
    assert torch.allclose(y, ref, atol=1e-5), float((y - ref).abs().max())


@pytest.mark.parametrize("cin,cout,k,s,p,hw", [(8, 3, 3, 1, 1, 6), (8, 1, 4, 2, 1, 5), (4, 2, 4, 2, 1, 4)])
def test_tap_unrolled_transposed_convolution(cin, cout, k, s, p, hw):
    torch.manual_seed(1)
    w = torch.randn(cin, cout, k, k)
    x = torch.randn(2, cin, hw, hw)
    ref = F.conv_transpose2d(x, w, stride=s, padding=p)
    taps = k * k
    P = torch.einsum("nchw,cj->nhwj", x, w.reshape(cin, cout * taps))       # the 1x1 product, columns j = co*taps + tap
    oh = (hw - 1) * s - 2 * p + k
    y = torch.zeros(2, cout, oh, oh)
    for o_h in range(oh):                                                   # the gather of col2im_act_kernel
        for o_w in range(oh):
            for kh in range(k):
                th = o_h + p - kh
                if th < 0 or th % s or th // s >= hw:
                    continue
                for kw in range(k):
                    tw = o_w + p - kw
                    if tw < 0 or tw % s or tw // s >= hw:
                        continue
                    for co in range(cout):
                        y[:, co, o_h, o_w] += P[:, th // s, tw // s, co * taps + kh * k + kw]
    assert torch.allclose(y, ref, atol=1e-4), float((y - ref).abs().max())

"""Per-op parity of the CUDA kernels (called through the C ABI) against plain PyTorch fp32 on CPU.

fp32 kernels: 1e-4 norm-relative.  bf16 / tcgen05 kernels: inputs are rounded to bf16 on both sides,
the reference is then computed in fp32, tolerance 1e-2 norm-relative (bf16 output rounding ~4e-3).
"""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from helpers import bf16_round, cos_sim, from_act, rel_err, to_act  # noqa: E402


def _ops():
    import dcvgan_b200
    dcvgan_b200.require_device()
    from dcvgan_b200 import ops
    return ops


# (name, kind, cin, cout, k, s, p, N, in_spatial)
CONV_CASES = [
    ("conv2d_k4s2", "conv", 8, 24, (1, 4, 4), (1, 2, 2), (0, 1, 1), 3, (1, 16, 16)),
    ("conv2d_k3s1", "conv", 2, 16, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, (1, 8, 8)),
    ("conv2d_head", "conv", 32, 1, (1, 4, 4), (1, 2, 2), (0, 1, 1), 4, (1, 8, 8)),
    ("conv2d_2to1", "conv", 16, 16, (1, 4, 4), (1, 2, 2), (0, 1, 1), 5, (1, 2, 2)),
    ("conv3d_k4", "conv", 3, 8, (4, 4, 4), (1, 2, 2), (0, 1, 1), 2, (7, 8, 8)),
    ("conv3d_k4_wide", "conv", 16, 40, (4, 4, 4), (1, 2, 2), (0, 1, 1), 2, (5, 8, 8)),
    ("convT_k4s2", "convT", 24, 8, (1, 4, 4), (1, 2, 2), (0, 1, 1), 3, (1, 8, 8)),
    ("convT_k4s1p0", "convT", 10, 16, (1, 4, 4), (1, 1, 1), (0, 0, 0), 6, (1, 1, 1)),
    ("convT_k3s1", "convT", 16, 3, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, (1, 8, 8)),
    ("convT_1to2", "convT", 12, 16, (1, 4, 4), (1, 2, 2), (0, 1, 1), 4, (1, 1, 1)),
]

TC_CASES = [
    ("tc_conv2d_64", "conv", 64, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), 4, (1, 32, 32)),
    ("tc_conv2d_128_256", "conv", 128, 256, (1, 4, 4), (1, 2, 2), (0, 1, 1), 8, (1, 8, 8)),
    ("tc_conv2d_small", "conv", 64, 128, (1, 4, 4), (1, 2, 2), (0, 1, 1), 16, (1, 2, 2)),
    ("tc_conv2d_c32", "conv", 32, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), 4, (1, 16, 16)),
    ("tc_conv2d_c16_n24", "conv", 16, 24, (1, 4, 4), (1, 2, 2), (0, 1, 1), 4, (1, 16, 16)),
    ("tc_conv3d_64_128", "conv", 64, 128, (4, 4, 4), (1, 2, 2), (0, 1, 1), 2, (13, 32, 32)),
    ("tc_conv3d_128_256", "conv", 128, 256, (4, 4, 4), (1, 2, 2), (0, 1, 1), 3, (10, 16, 16)),
    ("tc_convT_128_64", "convT", 128, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), 4, (1, 16, 16)),
    ("tc_convT_512_256", "convT", 512, 256, (1, 4, 4), (1, 2, 2), (0, 1, 1), 8, (1, 4, 4)),
    ("tc_convT_k3_n3", "convT", 128, 3, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, (1, 32, 32)),
    ("tc_convT_64_1", "convT", 64, 1, (1, 4, 4), (1, 2, 2), (0, 1, 1), 2, (1, 32, 32)),
    # halo-tile kernel: odd batch (masked second tile), 128-wide N tile with 4 fused phases, 3x3 stride-1 both ways
    ("halo_convT_128_128_odd", "convT", 128, 128, (1, 4, 4), (1, 2, 2), (0, 1, 1), 3, (1, 16, 16)),
    ("halo_convT_256_64", "convT", 256, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), 5, (1, 32, 32)),
    ("halo_conv_k3_64_64", "conv", 64, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), 3, (1, 32, 32)),
    ("halo_convT_k3_64_32", "convT", 64, 32, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, (1, 16, 24)),
    # weight gradient with row-halo sharing: ragged tiles (S 12 x 20, odd batch), two channel chunks, 3-D with odd T
    ("wgrad_halo_ragged", "conv", 64, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), 3, (1, 24, 40)),
    ("wgrad_halo_128_64", "conv", 128, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), 5, (1, 8, 16)),
    ("wgrad_halo_3d", "conv", 64, 32, (4, 4, 4), (1, 2, 2), (0, 1, 1), 3, (6, 8, 16)),
]


def _torch_fwd(kind, x, w, s, p, three_d):
    if kind == "conv":
        return F.conv3d(x, w, stride=s, padding=p) if three_d else F.conv2d(x[:, :, 0], w[:, :, 0], stride=s[1:], padding=p[1:]).unsqueeze(2)
    return F.conv_transpose2d(x[:, :, 0], w[:, :, 0], stride=s[1:], padding=p[1:]).unsqueeze(2)


def _run_case(ops, case, dtype, impl_tc, tol):
    name, kind, cin, cout, k, s, p, n, sp = case
    torch.manual_seed(sum(map(ord, name)) % 1000)
    three_d = k[0] > 1
    spec = ops.ConvSpec(kind, cin, cout, k, s, p)
    x = torch.randn(n, cin, *sp)
    w = torch.randn((cout, cin, *k) if kind == "conv" else (cin, cout, *k)) * 0.1
    if dtype == torch.bfloat16:
        x, w = bf16_round(x), bf16_round(w)
    x.requires_grad_(True)
    w.requires_grad_(True)
    y_ref = _torch_fwd(kind, x, w, s, p, three_d)
    dy = torch.randn_like(y_ref)
    if dtype == torch.bfloat16:
        dy = bf16_round(dy)
    y_ref.backward(dy)

    from dcvgan_b200._lib import IMPL_SIMT, IMPL_TC
    g = spec.geom(n, sp)
    xa = to_act(x.detach(), dtype)
    out_sp = spec.out_spatial(sp)
    ya = ops.Act.empty(n, *out_sp, cout, dtype)
    wdev = (w.detach().squeeze(2) if (kind == "convT" or not three_d) else w.detach()).contiguous().cuda()
    impl = IMPL_TC if impl_tc else IMPL_SIMT
    if impl_tc:
        assert ops.lib().dcv_conv_tc_supported(C.byref(g), spec.fwd_dir), "case should be tcgen05-eligible"
    # forward
    wp = ops.pack_weight(spec, g, spec.fwd_dir, impl, wdev)
    ops.conv(g, spec.fwd_dir, impl, xa, wp, ya)
    e_fwd = rel_err(from_act(ya), y_ref.detach())
    # data gradient
    dya = to_act(dy, dtype)
    dxa = ops.Act.empty(n, *sp, cin, dtype)
    impl_b = impl
    if impl_tc and not ops.lib().dcv_conv_tc_supported(C.byref(g), spec.bwd_dir):
        impl_b = IMPL_SIMT
    wpb = ops.pack_weight(spec, g, spec.bwd_dir, impl_b, wdev)
    ops.conv(g, spec.bwd_dir, impl_b, dya, wpb, dxa)
    e_dx = rel_err(from_act(dxa), x.grad)
    # weight gradient
    dw = torch.full_like(wdev, 7.0)
    xl, xs = (xa, dya) if kind == "conv" else (dya, xa)
    wimpl = IMPL_TC if (impl_tc and ops.lib().dcv_wgrad_tc_supported(C.byref(g))) else IMPL_SIMT
    ops.wgrad(spec, g, xl, xs, dw, accumulate=False, impl=wimpl)
    torch.cuda.synchronize()
    wg = w.grad.squeeze(2) if (kind == "convT" or not three_d) else w.grad
    e_dw = rel_err(dw.cpu(), wg)
    # accumulate=True adds on top
    ops.wgrad(spec, g, xl, xs, dw, accumulate=True, impl=wimpl)
    torch.cuda.synchronize()
    e_dw2 = rel_err(dw.cpu(), 2 * wg)
    print(f"{name}: fwd {e_fwd:.2e} dx {e_dx:.2e} dw {e_dw:.2e} dw2 {e_dw2:.2e} impl_b={impl_b} wimpl={wimpl}")
    assert e_fwd < tol and e_dx < tol and e_dw < tol and e_dw2 < tol, (name, e_fwd, e_dx, e_dw, e_dw2)


@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_simt_fp32(case):
    _run_case(_ops(), case, torch.float32, False, 1e-4)


@pytest.mark.parametrize("case", CONV_CASES + TC_CASES[:3], ids=[c[0] for c in CONV_CASES + TC_CASES[:3]])
def test_conv_simt_bf16(case):
    _run_case(_ops(), case, torch.bfloat16, False, 1e-2)


@pytest.mark.parametrize("case", TC_CASES, ids=[c[0] for c in TC_CASES])
def test_conv_tcgen05(case):
    _run_case(_ops(), case, torch.bfloat16, True, 1e-2)


TF32_CASES = [c for c in TC_CASES if c[2] % 32 == 0 and c[3] % 32 == 0]


@pytest.mark.parametrize("case", TF32_CASES, ids=[c[0] for c in TF32_CASES])
def test_conv_tcgen05_tf32(case):
    """tcgen05 kind::tf32 variant of the persistent convolution kernel (fp32 tensors, fp32 packed weights, fp32 output
    through the TMA-store epilogue): forward and data gradient against PyTorch fp32.  TF32 keeps 10 mantissa bits of each
    operand (products exact, fp32 accumulation): norm-relative error ~3e-4, gate 1e-3 (the north star's fp32/TF32 bar)."""
    ops = _ops()
    from dcvgan_b200._lib import IMPL_TC_TF32
    name, kind, cin, cout, k, s, p, n, sp = case
    torch.manual_seed(sum(map(ord, name)) % 1000)
    three_d = k[0] > 1
    spec = ops.ConvSpec(kind, cin, cout, k, s, p)
    x = torch.randn(n, cin, *sp, requires_grad=True)
    w = (torch.randn((cout, cin, *k) if kind == "conv" else (cin, cout, *k)) * 0.1).requires_grad_(True)
    y_ref = _torch_fwd(kind, x, w, s, p, three_d)
    dy = torch.randn_like(y_ref)
    y_ref.backward(dy)
    g = spec.geom(n, sp)
    assert ops.lib().dcv_conv_tf32_supported(C.byref(g), spec.fwd_dir) and ops.lib().dcv_conv_tf32_supported(C.byref(g), spec.bwd_dir)
    xa = to_act(x.detach(), torch.float32)
    ya = ops.Act.empty(n, *spec.out_spatial(sp), cout, torch.float32)
    wdev = (w.detach().squeeze(2) if (kind == "convT" or not three_d) else w.detach()).contiguous().cuda()
    ops.conv(g, spec.fwd_dir, IMPL_TC_TF32, xa, ops.pack_weight(spec, g, spec.fwd_dir, IMPL_TC_TF32, wdev), ya)
    e_fwd = rel_err(from_act(ya), y_ref.detach())
    dya = to_act(dy, torch.float32)
    dxa = ops.Act.empty(n, *sp, cin, torch.float32)
    ops.conv(g, spec.bwd_dir, IMPL_TC_TF32, dya, ops.pack_weight(spec, g, spec.bwd_dir, IMPL_TC_TF32, wdev), dxa)
    e_dx = rel_err(from_act(dxa), x.grad)
    print(f"{name} [tf32]: fwd {e_fwd:.2e} dx {e_dx:.2e}")
    assert e_fwd < 1e-3 and e_dx < 1e-3, (name, e_fwd, e_dx)


VARIANTS = [{"nohalo": 1}, {"mt": 1}, {"mt": 4}, {"no_tma_store": 1}, {"wgrad_waves": 2}, {"no_tapgroup": 1}, {"sm_reserve": 16},
            {"pdl": 1}, {"no_nsplit": 1}, {"no_wgrad_halo": 1}, {"wgrad_halo4": 1}, {"no_wgrad_whalo": 1}, {"wgrad_g": 4}, {"wgrad_g": 2}, {"wgrad_g": 2, "no_wgrad_whalo": 1}, {"wgrad_ns": 64}]


@pytest.mark.parametrize("tune", VARIANTS, ids=["+".join(f"{k}={v}" for k, v in e.items()) for e in VARIANTS])
@pytest.mark.parametrize("case", [TC_CASES[0], TC_CASES[1], TC_CASES[5], TC_CASES[7], TC_CASES[8], TC_CASES[9], TC_CASES[13], TC_CASES[14],
                                  TC_CASES[15], TC_CASES[17]],
                         ids=lambda c: c[0])
def test_conv_tcgen05_kernel_variants(case, tune):
    """Every tuning switch of the tcgen05 path (row-halo sharing off, forced M-tile counts, direct-store epilogue, two split
    waves in wgrad, no grouped taps, a reduced persistent grid) must give the same numbers as the default configuration:
    each variant is checked against PyTorch like the default path.  The switches are explicit library state
    (dcv_set_tuning), not environment variables."""
    ops = _ops()
    for k, v in tune.items():
        ops.check(ops.lib().dcv_set_tuning(k.encode(), v))
    try:
        _run_case(ops, case, torch.bfloat16, True, 1e-2)
    finally:
        for k in tune:
            ops.check(ops.lib().dcv_set_tuning(k.encode(), 0))


PADDED_CASES = [
    ("pad_conv3d_3_32", "conv", 3, 32, (4, 4, 4), (1, 2, 2), (0, 1, 1), 2, (7, 16, 16)),
    ("pad_conv3d_1_32", "conv", 1, 32, (4, 4, 4), (1, 2, 2), (0, 1, 1), 3, (6, 16, 16)),
    ("pad_conv2d_k3_25_64", "conv", 25, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, (1, 16, 16)),
    ("pad_convT_64_2", "convT", 64, 2, (1, 4, 4), (1, 2, 2), (0, 1, 1), 3, (1, 16, 16)),
    ("pad_convT_k3_128_3", "convT", 128, 3, (1, 3, 3), (1, 1, 1), (0, 1, 1), 2, (1, 16, 16)),
    ("pad_convT_50_512_1x1", "convT", 50, 512, (1, 4, 4), (1, 1, 1), (0, 0, 0), 32, (1, 1, 1)),
    ("pad_convT_266_256_1x1", "convT", 266, 256, (1, 4, 4), (1, 2, 2), (0, 1, 1), 32, (1, 1, 1)),
    ("pad_conv3d_head_256_1", "conv", 256, 1, (4, 4, 4), (1, 2, 2), (0, 1, 1), 3, (7, 8, 8)),
    ("pad_conv3d_32_64", "conv", 32, 64, (4, 4, 4), (1, 2, 2), (0, 1, 1), 2, (12, 32, 32)),
    ("pad_conv2d_96_192", "conv", 96, 192, (1, 4, 4), (1, 2, 2), (0, 1, 1), 4, (1, 16, 16)),
]


@pytest.mark.parametrize("case", PADDED_CASES, ids=[c[0] for c in PADDED_CASES])
def test_conv_tcgen05_padded_channels(case):
    """Tiny / odd channel counts run on the tensor cores through zero-padded activation channels
    (dcv_geom.wCl / wCs): forward, data gradient and weight gradient, all three via tcgen05."""
    ops = _ops()
    from dcvgan_b200._lib import IMPL_TC
    name, kind, cin, cout, k, s, p, n, sp = case
    torch.manual_seed(sum(map(ord, name)) % 1000)
    three_d = k[0] > 1
    spec = ops.ConvSpec(kind, cin, cout, k, s, p)
    x = bf16_round(torch.randn(n, cin, *sp)).requires_grad_(True)
    w = bf16_round(torch.randn((cout, cin, *k) if kind == "conv" else (cin, cout, *k)) * 0.1).requires_grad_(True)
    y_ref = _torch_fwd(kind, x, w, s, p, three_d)
    dy = bf16_round(torch.randn_like(y_ref))
    y_ref.backward(dy)
    xa = to_act(x.detach(), torch.bfloat16)              # Act.empty pads the channel count to a multiple of 16
    out_sp = spec.out_spatial(sp)
    ya = ops.Act.empty(n, *out_sp, cout, torch.bfloat16)
    g = spec.geom(n, sp, xa.cp, ya.cp)
    assert g.Cl % 16 == 0 and g.Cs % 16 == 0
    assert ops.lib().dcv_conv_tc_supported(C.byref(g), spec.fwd_dir) and ops.lib().dcv_conv_tc_supported(C.byref(g), spec.bwd_dir)
    assert ops.lib().dcv_wgrad_tc_supported(C.byref(g))
    wdev = (w.detach().squeeze(2) if (kind == "convT" or not three_d) else w.detach()).contiguous().cuda()
    wp = ops.pack_weight(spec, g, spec.fwd_dir, IMPL_TC, wdev)
    ops.conv(g, spec.fwd_dir, IMPL_TC, xa.padded_to(xa.cp), wp, ya.padded_to(ya.cp))
    e_fwd = rel_err(from_act(ya), y_ref.detach())
    pad = ya.padded_to(ya.cp).torch()[..., ya.c:]
    assert pad.numel() == 0 or float(pad.float().abs().max()) == 0.0      # padding channels stay zero
    dya = to_act(dy, torch.bfloat16)
    dxa = ops.Act.empty(n, *sp, cin, torch.bfloat16)
    wpb = ops.pack_weight(spec, g, spec.bwd_dir, IMPL_TC, wdev)
    ops.conv(g, spec.bwd_dir, IMPL_TC, dya.padded_to(dya.cp), wpb, dxa.padded_to(dxa.cp))
    e_dx = rel_err(from_act(dxa), x.grad)
    dw = torch.full_like(wdev, 7.0)
    xl, xs = (xa, dya) if kind == "conv" else (dya, xa)
    ops.wgrad(spec, g, xl.padded_to(xl.cp), xs.padded_to(xs.cp), dw, accumulate=False, impl=IMPL_TC)
    torch.cuda.synchronize()
    wg = w.grad.squeeze(2) if (kind == "convT" or not three_d) else w.grad
    e_dw = rel_err(dw.cpu(), wg)
    print(f"{name}: fwd {e_fwd:.2e} dx {e_dx:.2e} dw {e_dw:.2e}  geom Cl={g.Cl} Cs={g.Cs} wCl={g.wCl} wCs={g.wCs}")
    assert e_fwd < 1e-2 and e_dx < 1e-2 and e_dw < 1e-2


@pytest.mark.parametrize("cin,n,hw,slice_out", [(1, 3, (64, 64), True), (2, 2, (64, 64), False), (1, 2, (16, 32), False), (2, 5, (8, 64), True)])
def test_img_conv_direct_kernels(cin, n, hw, slice_out):
    """Inconv (generator.py:171-176: Conv2d(C, 64, 3, 1, 1) + LeakyReLU(0.01)) through the direct HBM-bound kernels:
    forward, and the one-pass backward (activation derivative + weight gradient + data gradient) against PyTorch fp32 on
    bf16-rounded operands.  slice_out: the output / its gradient live in the upper channel half of a 128-channel buffer,
    as in the U-Net concat buffer."""
    ops = _ops()
    torch.manual_seed(40 + cin + n)
    H, W = hw
    spec = ops.ConvSpec("conv", cin, 64, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    x = bf16_round(torch.randn(n, cin, H, W)).requires_grad_(True)
    w = bf16_round(torch.randn(64, cin, 3, 3) * 0.2).requires_grad_(True)  # the kernels round the fp32 master weight to bf16
    y_ref = F.leaky_relu(F.conv2d(x, w, padding=1), 0.01)
    dy = bf16_round(torch.randn_like(y_ref))
    y_ref.backward(dy)
    xa = to_act(x.detach(), torch.bfloat16)
    if slice_out:
        buf = ops.Act.empty(n, 1, H, W, 128, torch.bfloat16)
        ya = buf.ch(64, 128)
    else:
        ya = ops.Act.empty(n, 1, H, W, 64, torch.bfloat16)
    g = spec.geom(n, (1, H, W), xa.cp, ya.cp if not slice_out else 64)
    assert ops.img_conv_ok(spec, g, ops.IMG_FWD, xa, ya) and ops.img_conv_ok(spec, g, ops.IMG_BWD, xa, ya)
    from dcvgan_b200._lib import ACT_LEAKY
    wdev = w.detach().cuda().contiguous()
    ops.img_conv_fwd(spec, g, xa, wdev, ya, ACT_LEAKY, 0.01)
    e_fwd = rel_err(from_act(ya), y_ref.detach().unsqueeze(2))
    # The backward takes the LeakyReLU sign from the stored activation.  Fed with the kernel's own bf16 output the ~0.2 % of
    # elements whose pre-activation lies within rounding distance of zero change sign against the fp32 reference and their
    # dz changes by a factor 100 (measured 3e-2 in dx / dw); the arithmetic is therefore checked on the reference's output
    # rounded to bf16 (same signs), the self-consistent case only loosely.
    dya = to_act(dy, torch.bfloat16)
    ya_own = ya
    ya = to_act(y_ref.detach().unsqueeze(2), torch.bfloat16)
    dxa = ops.Act.empty(n, 1, H, W, cin, torch.bfloat16)
    dw = torch.full_like(wdev, 3.0)
    ops.img_conv_bwd(spec, g, dya, ya, xa, wdev, ACT_LEAKY, 0.01, dw, False, dxa)
    torch.cuda.synchronize()
    e_dx = rel_err(from_act(dxa), x.grad.unsqueeze(2))
    e_dw = rel_err(dw.cpu(), w.grad)
    ops.img_conv_bwd(spec, g, dya, ya, xa, wdev, ACT_LEAKY, 0.01, dw, True, None)
    torch.cuda.synchronize()
    e_dw2 = rel_err(dw.cpu(), 2 * w.grad)
    pad = dxa.padded_to(dxa.cp).torch()[..., dxa.c:]
    assert float(pad.float().abs().max()) == 0.0                           # padding channels of dx stay zero
    dxo = ops.Act.empty(n, 1, H, W, cin, torch.bfloat16)
    dwo = torch.zeros_like(wdev)
    ops.img_conv_bwd(spec, g, dya, ya_own, xa, wdev, ACT_LEAKY, 0.01, dwo, False, dxo)
    torch.cuda.synchronize()
    e_own = max(rel_err(from_act(dxo), x.grad.unsqueeze(2)), rel_err(dwo.cpu(), w.grad))
    print(f"img_conv C={cin} n={n} {H}x{W}: fwd {e_fwd:.2e} dx {e_dx:.2e} dw {e_dw:.2e} dw2 {e_dw2:.2e}; with its own stored output {e_own:.2e}")
    assert e_fwd < 5e-3 and e_dx < 5e-3 and e_dw < 5e-3 and e_dw2 < 5e-3 and e_own < 6e-2


@pytest.mark.parametrize("n,hw", [(2, (64, 64)), (3, (8, 32))])
def test_img_conv_outconv_backward(n, hw):
    """Outconv = ConvTranspose2d(128, 3, 3, 1, 1) (generator.py:272-277): its data gradient (3 -> 128 channels) and weight
    gradient through the image-side mma.sync kernels against PyTorch fp32 on bf16-rounded operands; the data gradient is
    written into a 128-channel buffer, the weight gradient also in accumulate mode."""
    ops = _ops()
    from dcvgan_b200._lib import ACT_NONE
    torch.manual_seed(60 + n)
    H, W = hw
    spec = ops.ConvSpec("convT", 128, 3, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    x = bf16_round(torch.randn(n, 128, H, W)).requires_grad_(True)
    w = bf16_round(torch.randn(128, 3, 3, 3) * 0.1).requires_grad_(True)
    y = F.conv_transpose2d(x, w, stride=1, padding=1)
    dy = bf16_round(torch.randn_like(y))
    y.backward(dy)
    xa = to_act(x.detach(), torch.bfloat16)
    dza = to_act(dy, torch.bfloat16)                     # 3 real channels in a 16-channel pitch
    g = spec.geom(n, (1, H, W), xa.cp, dza.cp)
    assert ops.img_conv_ok(spec, g, ops.IMG_FWD, dza, xa) and ops.img_conv_ok(spec, g, ops.IMG_WGRAD, dza, xa)
    wdev = w.detach().cuda().contiguous()
    dxa = ops.Act.empty(n, 1, H, W, 128, torch.bfloat16)
    ops.img_conv_fwd(spec, g, dza, wdev, dxa, ACT_NONE, 0.0)
    dw = torch.full_like(wdev, 2.0)
    ops.img_conv_bwd(spec, g, xa, None, dza, wdev, ACT_NONE, 0.0, dw, False, None)
    torch.cuda.synchronize()
    e_dx = rel_err(from_act(dxa), x.grad.unsqueeze(2))
    e_dw = rel_err(dw.cpu(), w.grad)
    ops.img_conv_bwd(spec, g, xa, None, dza, wdev, ACT_NONE, 0.0, dw, True, None)
    torch.cuda.synchronize()
    e_dw2 = rel_err(dw.cpu(), 2 * w.grad)
    # forward: ConvTranspose2d + Tanh in one pass over the 128-channel input
    from dcvgan_b200._lib import ACT_TANH
    assert ops.img_conv_ok(spec, g, ops.IMG_SCATTER, dza, xa)
    ya = ops.Act.empty(n, 1, H, W, 3, torch.bfloat16)
    ops.img_conv_scatter(spec, g, xa, wdev, ya, ACT_TANH, 0.0)
    e_fwd = rel_err(from_act(ya), torch.tanh(y.detach()).unsqueeze(2))
    assert float(ya.padded_to(ya.cp).torch()[..., 3:].float().abs().max()) == 0.0
    print(f"outconv n={n} {H}x{W}: fwd {e_fwd:.2e} dx {e_dx:.2e} dw {e_dw:.2e} dw2 {e_dw2:.2e}")
    assert e_fwd < 5e-3 and e_dx < 5e-3 and e_dw < 5e-3 and e_dw2 < 5e-3


@pytest.mark.parametrize("n,hw,gamma", [(2, (64, 64), True), (3, (8, 32), True), (2, (16, 16), False)])
def test_outconv_batchnorm_on_load(n, hw, gamma):
    """dcv_prebn: Outconv's forward kernel normalises + ReLUs the first 64 channels of its 128-channel input while loading it.
    Against the materialised form (dcv_bn_act into a copy of the buffer, then the same kernel): bit-identical output; the
    second half (the Inconv skip, negative values included) is untouched."""
    ops = _ops()
    from dcvgan_b200._lib import ACT_LEAKY, ACT_NONE, ACT_TANH
    torch.manual_seed(90 + n)
    H, W = hw
    spec = ops.ConvSpec("convT", 128, 3, (1, 3, 3), (1, 1, 1), (0, 1, 1))
    cat = to_act(bf16_round(torch.randn(n, 128, 1, H, W) * 1.5 + 0.2), torch.bfloat16)      # [z of up_blocks.5 | skip]
    z = cat.ch(0, 64)
    mean, invstd = (torch.randn(64) * 0.3).cuda(), (torch.rand(64) + 0.5).cuda()
    gd, bd = ((torch.randn(64) * 0.2 + 1).cuda(), (torch.randn(64) * 0.2).cuda()) if gamma else (None, None)
    mat = cat.like()
    mat.base.copy_(cat.base)
    ops.bn_act(z, mean, invstd, gd, bd, None, ACT_LEAKY, 0.0, mat.ch(0, 64))
    assert float(from_act(mat)[:, 64:].min()) < 0 and float(from_act(mat)[:, :64].min()) == 0.0
    pre = {"mean": mean, "invstd": invstd, "gamma": gd, "beta": bd, "c0": 0, "slope": 0.0}
    wdev = (torch.randn(128, 3, 3, 3) * 0.05).cuda()
    y0, y1 = ops.Act.empty(n, 1, H, W, 3, torch.bfloat16), ops.Act.empty(n, 1, H, W, 3, torch.bfloat16)
    g = spec.geom(n, (1, H, W), cat.cp, y0.cp)
    assert ops.img_conv_ok(spec, g, ops.IMG_SCATTER, y0, cat)
    ops.img_conv_scatter(spec, g, mat, wdev, y0, ACT_TANH, 0.0)
    ops.img_conv_scatter(spec, g, cat, wdev, y1, ACT_TANH, 0.0, pre)
    torch.cuda.synchronize()
    assert torch.equal(y0.base, y1.base)
    assert float(y0.base.float().abs().max()) > 0


@pytest.mark.parametrize("kind", [0, 1, 2, 3, 4])
def test_head_with_fused_loss_term(kind):
    """dcv_head_loss (discriminator head + its adversarial-loss term + dL/dlogits in one launch) against the two-launch form
    dcv_conv + dcv_loss_fwd_bwd: identical logits, loss and gradient (same bf16 logits, same summation order); first call
    writes the loss, second accumulates."""
    ops = _ops()
    from dcvgan_b200._lib import IMPL_TC
    torch.manual_seed(70 + kind)
    spec = ops.ConvSpec("conv", 256, 1, (4, 4, 4), (1, 2, 2), (0, 1, 1))
    n, sp = 6, (7, 8, 8)
    x = to_act(bf16_round(torch.randn(n, 256, *sp)), torch.bfloat16)
    w = (torch.randn(1, 256, 4, 4, 4) * 0.02).cuda()
    out_sp = spec.out_spatial(sp)
    y0 = ops.Act.empty(n, *out_sp, 1, torch.bfloat16)
    y1 = ops.Act.empty(n, *out_sp, 1, torch.bfloat16)
    g = spec.geom(n, sp, x.cp, y0.cp)
    wp = ops.pack_weight(spec, g, spec.fwd_dir, IMPL_TC, w)
    assert ops.head_loss_ok(g, spec.fwd_dir, IMPL_TC, x)
    # reference: two launches
    ops.conv(g, spec.fwd_dir, IMPL_TC, x, wp, y0.padded_to(y0.cp))
    loss0 = torch.zeros(1, device="cuda")
    dy0 = ops.Act.empty(n, *out_sp, 1, torch.bfloat16)
    ops.loss_fwd_bwd_act(y0, kind, loss0, False, dy0, 1.0)
    ops.loss_fwd_bwd_act(y0, kind, loss0, True, None, 1.0)
    # fused
    loss1 = torch.zeros(1, device="cuda")
    dy1 = ops.head_loss(g, spec.fwd_dir, x, wp, y1.padded_to(y1.cp), kind, loss1, False, True)
    assert ops.head_loss(g, spec.fwd_dir, x, wp, y1.padded_to(y1.cp), kind, loss1, True, False) is None
    torch.cuda.synchronize()
    assert torch.equal(from_act(y1), from_act(y0))
    assert torch.equal(from_act(dy1), from_act(dy0))
    assert float(dy1.padded_to(dy1.cp).torch()[..., 1:].float().abs().max()) == 0.0      # padding channels of dL/dlogits are zero
    assert abs(float(loss1) - float(loss0)) <= 1e-6 * max(1.0, abs(float(loss0))), (float(loss1), float(loss0))


def test_conv_accumulate_into_output():
    """dcv_conv_accumulate (TMA reduce-add epilogue): y += dgrad(x) equals the separate convolution followed by an add, for a
    strided data gradient (sub-pixel scatter) written into a channel slice of a wider buffer (the U-Net skip-gradient case)."""
    ops = _ops()
    from dcvgan_b200._lib import IMPL_TC
    torch.manual_seed(81)
    spec = ops.ConvSpec("conv", 64, 128, (1, 4, 4), (1, 2, 2), (0, 1, 1))
    n, sp = 5, (1, 32, 32)
    w = (torch.randn(128, 64, 4, 4) * 0.05).cuda()
    dz = to_act(bf16_round(torch.randn(n, 128, 16, 16)), torch.bfloat16)
    g = spec.geom(n, sp)
    wpb = ops.pack_weight(spec, g, spec.bwd_dir, IMPL_TC, w)
    base = bf16_round(torch.randn(n, 128, 32, 32))
    buf0, buf1 = to_act(base, torch.bfloat16), to_act(base, torch.bfloat16)
    skip0, skip1 = buf0.ch(64, 128), buf1.ch(64, 128)
    tmp = ops.Act.empty(n, 1, 32, 32, 64, torch.bfloat16)
    ops.conv(g, spec.bwd_dir, IMPL_TC, dz, wpb, tmp)
    ops.axpy(tmp, skip0, True)
    assert ops.conv_accumulate_ok(g, spec.bwd_dir, IMPL_TC, dz, skip1)
    ops.conv_accumulate(g, spec.bwd_dir, dz, wpb, skip1)
    torch.cuda.synchronize()
    a, b = from_act(buf1), from_act(buf0)
    assert torch.equal(a[:, :64], b[:, :64])                                        # the other half of the buffer is untouched
    e = rel_err(a[:, 64:], b[:, 64:])
    print(f"conv_accumulate vs conv + add: {e:.2e}")
    assert e < 2e-3          # one bf16 rounding of the partial result before the add instead of fp32 -> bf16 after it


def test_conv_channel_slices():
    """input read from / output written into channel slices of wider buffers (concat elimination)"""
    ops = _ops()
    from dcvgan_b200._lib import IMPL_SIMT, IMPL_TC
    for dtype, impl in ((torch.float32, IMPL_SIMT), (torch.bfloat16, IMPL_TC)):
        torch.manual_seed(3)
        spec = ops.ConvSpec("conv", 64, 32, (1, 4, 4), (1, 2, 2), (0, 1, 1))
        x = bf16_round(torch.randn(2, 64, 16, 16))
        w = bf16_round(torch.randn(32, 64, 4, 4) * 0.1)
        y_ref = F.conv2d(x, w, stride=2, padding=1)
        wide_in = ops.Act.empty(2, 1, 16, 16, 96, dtype, zero=True)
        ops.to_channels_last(x.cuda().unsqueeze(2), wide_in.ch(32, 96))
        wide_out = ops.Act.empty(2, 1, 8, 8, 72, dtype, zero=True)
        g = spec.geom(2, (1, 16, 16))
        wp = ops.pack_weight(spec, g, spec.fwd_dir, impl, w.cuda())
        ops.conv(g, spec.fwd_dir, impl, wide_in.ch(32, 96), wp, wide_out.ch(40, 72))
        full = from_act(wide_out, 4)
        assert rel_err(full[:, 40:72], y_ref) < 1e-2
        assert float(full[:, :40].abs().max()) == 0.0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("act", ["relu", "leaky"])
def test_batchnorm_act_dropout(dtype, act):
    ops = _ops()
    from dcvgan_b200._lib import ACT_LEAKY
    torch.manual_seed(1)
    n, c, h, w = 6, 48, 8, 8
    slope = 0.0 if act == "relu" else 0.2
    tol = 1e-4 if dtype == torch.float32 else 2e-2
    z = torch.randn(n, c, h, w) * 2 + 0.5
    if dtype == torch.bfloat16:
        z = bf16_round(z)
    z.requires_grad_(True)
    gamma = (torch.randn(c) * 0.1 + 1).requires_grad_(True)
    beta = (torch.randn(c) * 0.1).requires_grad_(True)
    rm, rv = torch.randn(c) * 0.1, torch.rand(c) + 0.5
    rm_ref, rv_ref = rm.clone(), rv.clone()
    drop = (torch.rand(n, c) < 0.5).float() * 2.0
    u = F.batch_norm(z, rm_ref, rv_ref, gamma, beta, training=True, momentum=0.1, eps=1e-5)
    a_ref = F.leaky_relu(u * drop[:, :, None, None], slope)
    da = torch.randn_like(a_ref)
    a_ref.backward(da)

    za = to_act(z.detach(), dtype)
    rmd, rvd = rm.cuda(), rv.cuda()
    nbt = torch.tensor(7, dtype=torch.int64, device="cuda")
    mean, invstd = ops.bn_batch_stats(za, 1e-5, 0.1, rmd, rvd, nbt)
    assert int(nbt) == 8          # nn.BatchNorm's num_batches_tracked, incremented on the device
    aa = za.like()
    gd, bd, dd = gamma.detach().cuda(), beta.detach().cuda(), drop.cuda().contiguous()
    ops.bn_act(za, mean, invstd, gd, bd, dd, ACT_LEAKY, slope, aa)
    assert rel_err(from_act(aa, 4), a_ref.detach()) < tol
    torch.cuda.synchronize()
    assert rel_err(rmd.cpu(), rm_ref) < 1e-5 and rel_err(rvd.cpu(), rv_ref) < 1e-5
    daa = to_act(da, dtype)
    dza = za.like()
    dg, db = torch.empty(c, device="cuda"), torch.empty(c, device="cuda")
    ops.bn_act_bwd(daa, aa, za, mean, invstd, gd, bd, dd, ACT_LEAKY, slope, dza, dg, db)
    assert rel_err(from_act(dza, 4), z.grad) < tol
    # the vectorised (Leaky)ReLU kernels take the activation sign from z, never from `a`: poisoning `a` changes nothing
    if c % 8 == 0:
        dz2 = za.like()
        junk = aa.like()
        junk.base.fill_(float("nan"))
        ops.bn_act_bwd(daa, junk, za, mean, invstd, gd, bd, dd, ACT_LEAKY, slope, dz2, dg, db)
        assert torch.equal(dz2.base, dza.base)
    torch.cuda.synchronize()
    assert rel_err(dg.cpu(), gamma.grad) < tol and rel_err(db.cpu(), beta.grad) < tol
    # eval-mode statistics
    m2, i2 = ops.bn_eval_stats(rmd, rvd, 1e-5)
    ops.bn_act(za, m2, i2, gd, bd, None, ACT_LEAKY, slope, aa)
    a_eval = F.leaky_relu(F.batch_norm(z.detach(), rm_ref, rv_ref, gamma.detach(), beta.detach(), training=False, eps=1e-5), slope)
    assert rel_err(from_act(aa, 4), a_eval) < tol


def test_batchnorm_wide_channels():
    ops = _ops()
    from dcvgan_b200._lib import ACT_NONE
    torch.manual_seed(2)
    z = torch.randn(4, 300, 2, 2)
    za = to_act(z, torch.float32)
    mean, invstd = ops.bn_batch_stats(za, 1e-5, 0.1, None, None)
    torch.cuda.synchronize()
    assert rel_err(mean.cpu(), z.mean((0, 2, 3))) < 1e-5
    assert rel_err(invstd.cpu(), 1 / torch.sqrt(z.var((0, 2, 3), unbiased=False) + 1e-5)) < 1e-5


def test_elementwise_misc():
    ops = _ops()
    from dcvgan_b200._lib import ACT_LEAKY, ACT_TANH
    torch.manual_seed(4)
    for dtype, tol in ((torch.float32, 1e-6), (torch.bfloat16, 1e-2)):
        x = torch.randn(2, 5, 6, 4, 4)
        xa = to_act(x, dtype)
        xr = from_act(xa)
        assert rel_err(xr, x) < tol
        # temporal difference and its adjoint
        d = ops.Act.empty(2, 5, 4, 4, 5, dtype)
        ops.tdiff(xa, d)
        assert rel_err(from_act(d), xr[:, :, 1:] - xr[:, :, :-1]) < tol
        dy = torch.randn(2, 5, 5, 4, 4)
        dya = to_act(dy, dtype)
        dxa = xa.like()
        ops.tdiff_bwd(dya, dxa, False)
        xg = xr.clone().requires_grad_(True)
        ((xg[:, :, 1:] - xg[:, :, :-1]) * from_act(dya)).sum().backward()
        assert rel_err(from_act(dxa), xg.grad) < tol
        # softmax fwd/bwd over channels
        sa = xa.like()
        ops.softmax(xa, sa)
        xg = xr.clone().requires_grad_(True)
        sm = torch.softmax(xg, 1)
        assert rel_err(from_act(sa), sm.detach()) < tol
        gy = torch.randn_like(sm)
        sm.backward(gy)
        ga = to_act(gy, dtype)
        dz = xa.like()
        ops.softmax_bwd(ga, sa, dz)
        assert rel_err(from_act(dz), xg.grad) < max(tol, 2e-2 if dtype == torch.bfloat16 else 1e-5)
        # segmentation remap
        ra = xa.like()
        ops.segm_remap(xa, ra)
        ref = torch.full_like(xr, -1.0).scatter_(1, xr.argmax(1, keepdim=True), 1.0)
        assert torch.equal(from_act(ra), ref)
        # noise, axpy, frame extract / scatter, activation backward
        nz = torch.randn(xa.rows * xa.c, device="cuda")
        na = xa.like()
        ops.add_noise(xa, nz, 0.2, na)
        ref = xr + 0.2 * nz.cpu().view(2, 6, 4, 4, 5).permute(0, 4, 1, 2, 3)
        assert rel_err(from_act(na), ref) < tol
        ops.axpy(xa, na, True)
        assert rel_err(from_act(na), ref + xr) < tol
        fr = ops.Act.empty(2, 1, 4, 4, 5, dtype)
        ops.frame_extract(xa, 3, fr)
        assert torch.equal(from_act(fr)[:, :, 0], xr[:, :, 3])
        acc = xa.like()
        ops.copy_cl(xa, acc)
        ops.frame_scatter(acc, 3, fr, True)
        ref = xr.clone()
        ref[:, :, 3] *= 2
        assert rel_err(from_act(acc), ref) < tol
        for act, slope, fn in ((ACT_LEAKY, 0.2, lambda v: F.leaky_relu(v, 0.2)), (ACT_TANH, 0.0, torch.tanh)):
            xg = xr.clone().requires_grad_(True)
            a = fn(xg)
            a.backward(torch.ones_like(xr))
            aa = to_act(a.detach(), dtype)
            ones = to_act(torch.ones_like(xr), dtype)
            dz = xa.like()
            ops.act_bwd(ones, aa, act, slope, dz)
            assert rel_err(from_act(dz), xg.grad) < max(tol, 2e-2 if dtype == torch.bfloat16 else 1e-5)


def test_layout_strided_frames():
    """to/from channels-last with the strided views the reference produces (x[:, :, t], permuted videos)"""
    ops = _ops()
    torch.manual_seed(5)
    v = torch.randn(3, 2, 7, 8, 8, device="cuda")
    frame = v[:, :, 4]
    fa = ops.Act.empty(3, 1, 8, 8, 2, torch.float32)
    ops.to_channels_last(frame, fa)
    assert torch.equal(from_act(fa, 4), frame.cpu())
    mem = torch.zeros(3, 7, 2, 8, 8, device="cuda")
    out = mem.permute(0, 2, 1, 3, 4)  # (B,C,T,H,W) view of (B,T,C,H,W) memory, like generator.py:136-139
    va = to_act(v, torch.float32)
    ops.from_channels_last(va, out)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), v.cpu())
    ops.from_channels_last(va, out, accumulate=True)
    torch.cuda.synchronize()
    assert torch.equal(out.cpu(), 2 * v.cpu())


@pytest.mark.parametrize("b,t,d", [(5, 16, 10), (32, 16, 10), (37, 7, 10), (6, 9, 12)])
def test_gru_trajectory(b, t, d):
    """d = 10: register-resident kernels (one gate row per lane; 32 rows fill the 32-warp BPTT block, 37 wrap around);
    d = 12: the generic kernels"""
    ops = _ops()
    torch.manual_seed(6)
    cell = torch.nn.GRUCell(d, d)
    h0 = torch.randn(b, d)
    eps = torch.randn(t, b, d)
    h, hs = h0, []
    for i in range(t):
        h = cell(eps[i], h)
        hs.append(h)
    hs_ref = torch.stack(hs, 1)
    dhs = torch.randn_like(hs_ref)
    hs_ref.backward(dhs)
    P = [p.detach().cuda().contiguous() for p in (cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)]
    hs_dev = ops.gru_traj_fwd(h0.cuda(), eps.cuda(), *P)
    torch.cuda.synchronize()
    assert rel_err(hs_dev.cpu(), hs_ref.detach()) < 1e-5
    grads = [torch.empty_like(p) for p in P]
    ops.gru_traj_bwd(h0.cuda(), eps.cuda(), hs_dev, dhs.cuda().contiguous(), *P, *grads)
    torch.cuda.synchronize()
    for gdev, pref in zip(grads, (cell.weight_ih, cell.weight_hh, cell.bias_ih, cell.bias_hh)):
        assert rel_err(gdev.cpu(), pref.grad) < 1e-4


def test_losses():
    ops = _ops()
    from dcvgan_b200 import _lib
    torch.manual_seed(7)
    y = (torch.randn(4, 4, 4, 4) * 3).requires_grad_(True)
    bce = torch.nn.BCEWithLogitsLoss(reduction="sum")
    refs = {
        _lib.LOSS_BCE_ONES: lambda v: bce(v, torch.ones_like(v)) / v.numel(),
        _lib.LOSS_BCE_ZEROS: lambda v: bce(v, torch.zeros_like(v)) / v.numel(),
        _lib.LOSS_HINGE_REAL: lambda v: torch.mean(F.relu(1.0 - v)),
        _lib.LOSS_HINGE_FAKE: lambda v: torch.mean(F.relu(1.0 + v)),
        _lib.LOSS_SOFTPLUS_NEG: lambda v: torch.mean(F.softplus(-v)),
    }
    for kind, fn in refs.items():
        y.grad = None
        l_ref = fn(y)
        l_ref.backward()
        yd = y.detach().cuda().contiguous()
        out = torch.zeros(1, device="cuda")
        dy = torch.empty_like(yd)
        ops.loss_fwd_bwd(yd, kind, out, False, dy, 1.0)
        ops.loss_fwd_bwd(yd, kind, out, True, None, 1.0)
        torch.cuda.synchronize()
        assert abs(float(out) - 2 * float(l_ref)) < 1e-5 * max(1.0, abs(float(l_ref)))
        assert rel_err(dy.cpu(), y.grad) < 1e-5


def test_adam_matches_torch():
    ops = _ops()
    torch.manual_seed(8)
    shapes = [(3, 5, 4, 4), (4097,), (10,), (64, 64, 4, 4), (1,)]
    params = [torch.randn(s) for s in shapes]
    ref = [p.clone().requires_grad_(True) for p in params]
    opt = torch.optim.Adam(ref, lr=2e-4, betas=(0.5, 0.999), weight_decay=1e-5)
    dev = [p.cuda() for p in params]
    m = [torch.zeros_like(p) for p in dev]
    v = [torch.zeros_like(p) for p in dev]
    for step in range(1, 4):
        grads = [torch.randn(s) for s in shapes]
        for r, g in zip(ref, grads):
            r.grad = g.clone()
        opt.step()
        ops.adam_multi(dev, [g.cuda() for g in grads], m, v, 2e-4, 0.5, 0.999, 1e-8, 1e-5, step)
    torch.cuda.synchronize()
    for d, r in zip(dev, ref):
        assert rel_err(d.cpu(), r.detach()) < 1e-6
    for mm, r in zip(m, ref):
        assert rel_err(mm.cpu(), opt.state[r]["exp_avg"]) < 1e-6


@pytest.mark.parametrize("case", [TC_CASES[0], TC_CASES[5], TC_CASES[7], TC_CASES[13], ("tc_convT_phase_edge", "convT", 64, 64, (1, 4, 4), (1, 2, 2), (0, 1, 1), 3, (1, 12, 20))],
                         ids=lambda c: c[0])
def test_conv_fused_batchnorm_statistics(case):
    """dcv_conv_stats: per-channel sum / sum of squares accumulated in the convolution's TMA-store epilogue equal the
    statistics of the stored bf16 output (what dcv_bn_stats reads back), including tiles that overhang the tensor."""
    ops = _ops()
    from dcvgan_b200._lib import IMPL_TC
    name, kind, cin, cout, k, s, p, n, sp = case
    torch.manual_seed(5)
    spec = ops.ConvSpec(kind, cin, cout, k, s, p)
    x = bf16_round(torch.randn(n, cin, *sp))
    w = bf16_round(torch.randn((cout, cin, *k) if kind == "conv" else (cin, cout, *k)) * 0.1)
    g = spec.geom(n, sp)
    xa = to_act(x, torch.bfloat16)
    ya = ops.Act.empty(n, *spec.out_spatial(sp), cout, torch.bfloat16)
    three_d = k[0] > 1
    wdev = (w.squeeze(2) if (kind == "convT" or not three_d) else w).contiguous().cuda()
    wp = ops.pack_weight(spec, g, spec.fwd_dir, IMPL_TC, wdev)
    slots = ops.conv_stats_slots(g, spec.fwd_dir, xa, ya)
    assert slots > 0, "64-multiple output channels must take the fused path"
    partials = ops.conv_stats(g, spec.fwd_dir, xa, wp, ya, slots)
    mean, invstd = ops.bn_finalize(partials, ya.rows, 1e-5, 0.1, None, None)
    mean2, invstd2 = ops.bn_batch_stats(ya, 1e-5, 0.1, None, None)
    y = from_act(ya)
    y_ref = _torch_fwd(kind, x, w, s, p, three_d)
    assert rel_err(y, y_ref) < 1e-2
    assert rel_err(mean.cpu(), mean2.cpu()) < 1e-4 and rel_err(invstd.cpu(), invstd2.cpu()) < 1e-4, (
        rel_err(mean.cpu(), mean2.cpu()), rel_err(invstd.cpu(), invstd2.cpu()))

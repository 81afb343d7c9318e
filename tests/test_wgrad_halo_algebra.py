"""CPU check of the index algebra behind the halo-sharing plans of the tcgen05 weight gradient (csrc/tc_conv.cu,
wgrad_tc_kernel / wgrad_halo_tile; no GPU, no library calls).

The weight gradient of a 4x4(x kt) stride-2 convolution (discriminator.py:94-99,197-204; generator.py:205-211,239-248) is
dW[tap][cl][cs] = sum over output pixels (i, j) of L[2 i + kh - ph][2 j + kw - pw][cl] * S[i][j][cs].  The kernel does not
load one L box per tap: the taps kh and kh + 2 (and, in the column-halo plans, kw and kw + 2) read the same stride-2
lattice of L one tile row (one pixel) apart, so they are served by ONE box with a halo row (and column).  This file restates
that decomposition with numpy-style indexing and checks it against the direct formula, and checks that the tile
enumeration covers every (tap, channel chunk) exactly once.
"""
import itertools

import pytest
import torch


def halo_tile(mode, kt, kh, kw, clchunks, tile):
    """restatement of wgrad_halo_tile: tile index -> (kt index, kh class, kw of row block 0, channel chunk)"""
    if mode == 2:
        dw = tile % 2
        tile //= 2
        clc = tile % clchunks
        tile //= clchunks
        tc = tile % 2 + 2 * dw
        tile //= 2
    else:
        clc = tile % clchunks
        tile //= clchunks
        tc = tile % kw
        tile //= kw
    return tile // 2, tile % 2, tc, clc


@pytest.mark.parametrize("mode,kt,clchunks", [(1, 1, 1), (1, 4, 2), (2, 1, 1), (2, 4, 2), (2, 1, 4)])
def test_tiles_cover_every_tap_and_chunk_once(mode, kt, clchunks):
    kh = kw = 4
    ntiles = kt * kh * kw * clchunks // 2                       # a 128-row tile = the pair (kh, kh + 2) of one 64-channel chunk
    seen = set()
    for tile in range(ntiles):
        ta, cls, tc, clc = halo_tile(mode, kt, kh, kw, clchunks, tile)
        assert 0 <= ta < kt and cls in (0, 1) and 0 <= tc < kw and 0 <= clc < clchunks
        for d in (0, 1):                                        # row block d of the tile is the tap kh = class + 2 d
            key = (ta, cls + 2 * d, tc, clc)
            assert key not in seen
            seen.add(key)
    assert seen == set(itertools.product(range(kt), range(kh), range(kw), range(clchunks)))
    if mode == 2:                                               # tiles 2b and 2b + 1 share a box: same (kt, kh class, kw class, chunk)
        for b in range(ntiles // 2):
            t0, t1 = halo_tile(2, kt, kh, kw, clchunks, 2 * b), halo_tile(2, kt, kh, kw, clchunks, 2 * b + 1)
            assert t0[0] == t1[0] and t0[1] == t1[1] and t0[3] == t1[3] and t0[2] % 2 == t1[2] % 2 and t1[2] == t0[2] + 2


@pytest.mark.parametrize("h0,w0", [(0, 0), (4, 8), (12, 8)])       # tiles of the 16 x 16 output map, incl. both image borders
def test_halo_box_views_equal_the_per_tap_gather(h0, w0):
    """One box of (bh + 1) x (bw + 1) lattice pixels per parity class, read at row offset kh // 2 and pixel offset kw // 2,
    holds exactly the pixels the tap (kh, kw) gathers for a bh x bw tile of output pixels (zero outside the image)."""
    torch.manual_seed(0)
    H = W = 32
    bh, bw, ph, pw = 4, 8, 1, 1
    L = torch.randn(H, W)
    Lp = torch.zeros(H + 8, W + 8)
    Lp[4:4 + H, 4:4 + W] = L                                    # zero padding, as TMA's out-of-bounds fill

    def at(h, w):
        return Lp[h + 4, w + 4]
    for ch, cw in itertools.product((0, 1), (0, 1)):
        # the box TMA delivers: lattice rows 2 (h0 + r) + ch - ph, r = 0 .. bh; columns 2 (w0 + c) + cw - pw, c = 0 .. bw
        box = torch.stack([torch.stack([at(2 * (h0 + r) + ch - ph, 2 * (w0 + c) + cw - pw) for c in range(bw + 1)])
                           for r in range(bh + 1)])
        for dh, dw in itertools.product((0, 1), (0, 1)):
            kh, kw = ch + 2 * dh, cw + 2 * dw
            view = box[dh:dh + bh, dw:dw + bw]                  # descriptor start: dh box rows + dw pixels further
            direct = torch.stack([torch.stack([at(2 * (h0 + i) + kh - ph, 2 * (w0 + j) + kw - pw) for j in range(bw)])
                                  for i in range(bh)])
            assert torch.equal(view, direct), (kh, kw)


def test_k_step_offsets_inside_the_box():
    """The MMA walks a tile in K steps of 16 pixels (w fastest).  Inside the halo box of the column-halo plan a tile row is 8 of
    the 9 pixels of a box row, so consecutive 8-pixel groups are one box-row pitch apart (SBO = 9 pixels), a K step is two box
    rows, and every (t, n) slab of the tile adds one more box row (its halo row): offset(k) = k * 2 * rp + slab(k) * rp."""
    bw, bh, rp, slabs = 8, 4, 9, 2
    slab_shift = (bw * bh).bit_length() - 1
    box_px = lambda slab, r, c: (slab * (bh + 1) + r) * rp + c            # pixel index inside the box
    for k in range(slabs * bw * bh // 16):
        slab = (k * 16) >> slab_shift
        off = k * 2 * rp + slab * rp
        for g in range(2):                                               # the two 8-pixel groups of the K step
            q = k * 16 + g * 8                                           # first tile pixel of the group
            r, c = (q % (bw * bh)) // bw, q % bw
            assert off + g * rp == box_px(q // (bw * bh), r, c)

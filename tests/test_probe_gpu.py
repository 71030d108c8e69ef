"""Hardware property the halo-tile kernel (csrc/conv_halo.cu) stands on: a tcgen05 A operand may be a ROW-SHIFTED VIEW of a
128-byte-swizzled shared-memory tile - the descriptor's start address moved by an arbitrary number of 128-byte rows, base-offset
field left 0 - and still read rows [shift, shift + 128) of what the TMA unit wrote (the swizzle is a function of the address).
Checked with the probe of csrc/probe.cu (libdeepsent_dev.so) for every shift a 3x3 tap produces on the padded grids in use
(r * Wp + s for Wp = 16, 30, 58) and a few others."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_row_shifted_view_of_a_swizzled_tile_reads_the_shifted_rows():
    from tumblr_emotions_b200 import ops
    from tumblr_emotions_b200._lib import use_dev
    dev = use_dev(True)
    ops.init(0)
    try:
        a = torch.randn(256, 64, generator=torch.Generator().manual_seed(0)).bfloat16().cuda()
        eye = torch.eye(64).bfloat16().cuda()
        shifts = sorted({r * wp + s for wp in (16, 30, 58) for r in range(3) for s in range(3)} | {3, 7, 8, 9, 15, 127, 128})
        for shift in shifts:
            d = torch.full((128, 64), -1.0, device="cuda")
            dev.probe_umma_row_shift(a.data_ptr(), eye.data_ptr(), shift, 0, d.data_ptr(), torch.cuda.current_stream().cuda_stream)
            torch.cuda.synchronize()
            assert torch.equal(d, a[shift:shift + 128].float()), shift
    finally:
        use_dev(False)
        ops.init(0)

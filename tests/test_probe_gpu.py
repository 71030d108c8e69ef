"""Opt-in development probe (DS_RUN_PROBES=1): shifted-view UMMA operands (csrc/probe.cu).  Not a parity test - it answers a
hardware question for the next kernel and is skipped in the normal GPU suite."""
import os

import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("DS_RUN_PROBES") != "1", reason="development probe: set DS_RUN_PROBES=1")]


def test_unshifted_view_reproduces_the_tile():
    from tumblr_emotions_b200 import ops
    from tumblr_emotions_b200._lib import use_dev
    lib = lambda: use_dev(True)
    ops.init(0)
    a = torch.randn(256, 64, generator=torch.Generator().manual_seed(0)).bfloat16().cuda()
    eye = torch.eye(64).bfloat16().cuda()
    for shift in (0, 8, 16, 128):          # whole swizzle atoms: must work with either encoding
        d = torch.zeros(128, 64, device="cuda")
        lib().probe_umma_row_shift(a.data_ptr(), eye.data_ptr(), shift, 0, d.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert torch.equal(d, a[shift:shift + 128].float()), shift

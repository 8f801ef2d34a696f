"""Probe for csrc/wgrad.cu: structured inputs that expose which (row, column) pairs the tensor-core product really combines
(ones: the reduction length; row / column index ramps: operand orientation; identity rows: the pairing along the reduction).
This is how the shared-memory layout of the MN-major tf32 operands was found: with the common 16-byte-atom 128B swizzle the
MMA returns zeros, with SWIZZLE_128B_BASE32B (TMA: SWIZZLE_128B_ATOM_32B), LBO = 8 KB, SBO = 512 B every pattern is exact."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clvln_b200  # noqa: E402,F401
from clvln_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
for R, M, N in [(64, 128, 128), (256, 256, 384), (77, 132, 260)]:
    for name, dy, x in [("ones", torch.ones(R, M), torch.ones(R, N)),
                        ("rand", torch.randn(R, M), torch.randn(R, N)),
                        ("m-index x ones", torch.arange(M).float().repeat(R, 1), torch.ones(R, N)),
                        ("ones x n-index", torch.ones(R, M), torch.arange(N).float().repeat(R, 1)),
                        ("r-delta", torch.eye(R, M), torch.eye(R, N))]:
        dy, x = dy.to(dev), x.to(dev)
        out = ops.wgrad_tc(dy, x)
        ref = dy.double().t() @ x.double()
        err = float((out.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-9))
        print(f"R{R} M{M} N{N} {name}: max-rel {err:.2e}", flush=True)

"""Debug probe for csrc/wgrad.cu: structured inputs that expose which (row, column) pairs the MMA really combines, over
a few shared-memory descriptor settings (VLN_WGRAD_DBG = lbo,sbo,kstep,idesc(hex),layout_type)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clvln_b200
from clvln_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
for dbg in ["8192,512,1024,0,1", "8192,1024,1024,0,1", "8192,512,512,0,1", "512,8192,1024,0,1", "8192,1024,1024,0,2"]:
    os.environ["VLN_WGRAD_DBG"] = dbg
    print("== DBG", dbg)
    for R, M, N in [(64, 128, 128), (256, 256, 384)]:
        for name, dy, x in [("ones", torch.ones(R, M), torch.ones(R, N)),
                            ("rand", torch.randn(R, M), torch.randn(R, N)),
                            ("m-index x ones", torch.arange(M).float().repeat(R, 1), torch.ones(R, N)),
                            ("ones x n-index", torch.ones(R, M), torch.arange(N).float().repeat(R, 1)),
                            ("r-delta", torch.eye(R, M), torch.eye(R, N))]:
            dy, x = dy.to(dev), x.to(dev)
            out = ops.wgrad_tc(dy, x)
            ref = dy.double().t() @ x.double()
            torch.cuda.synchronize()
            err = float((out.double() - ref).abs().max() / ref.abs().max().clamp_min(1e-9))
            print(f"  R{R} M{M} N{N} {name}: err {err:.4f} out[0,:4] {out[0, :4].tolist()} out[:4,0] {out[:4, 0].tolist()} nonzero {int((out != 0).sum())}",
                  flush=True)

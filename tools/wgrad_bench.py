"""Times the weight-gradient / input-gradient kernels of csrc/wgrad.cu at the shapes of an EnvDrop iteration (B = 64, paired
128-row steps x 35), alone, with CUDA events; beside them the library TF32 GEMM of the same product.  One JSON line each."""
import json, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clvln_b200
from clvln_b200 import ops
dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = True


def timeit(fn, n=20):
    """n back-to-back launches replayed as one CUDA graph (no host time between them), CUDA events around the replay."""
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(n):
                fn()
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


shapes = [("gates W_ih|W_hh", 4480, 2048, 2752), ("attn out", 4480, 512, 1024), ("visual in", 4480, 2176, 512),
          ("cand", 4480, 2176, 512), ("text in", 5120, 512, 512), ("encoder W_ih", 5120, 1024, 256), ("action emb", 4480, 64, 128)]
for name, R, M, N in shapes:
    dy, x = torch.randn(R, M, device=dev), torch.randn(R, N, device=dev)
    us = timeit(lambda: ops.wgrad_tc(dy, x))
    lib = timeit(lambda: dy.t() @ x)
    print(json.dumps({"kernel": "wgrad_tf32", "name": name, "R": R, "M": M, "N": N, "us": round(us, 2), "tflops": round(2 * R * M * N / us / 1e6, 1),
                      "operand_GBs": round((R * M + R * N + M * N) * 4 / us / 1e3, 1), "library_tf32_us": round(lib, 2)}), flush=True)
dy, w = torch.randn(5120, 1024, device=dev), torch.randn(1024, 256, device=dev)
us = timeit(lambda: ops.dgrad_tc(dy, w))
lib = timeit(lambda: dy @ w)
print(json.dumps({"kernel": "dgrad_tf32", "name": "encoder dx", "M": 5120, "R": 1024, "N": 256, "us": round(us, 2),
                  "tflops": round(2 * 5120 * 1024 * 256 / us / 1e6, 1), "library_tf32_us": round(lib, 2)}), flush=True)
a, v = torch.randn(35, 128, 80, device=dev), torch.randn(35, 128, 512, device=dev)
us = timeit(lambda: ops.seq_outer_sum(a, v))
lib = timeit(lambda: torch.bmm(a.permute(1, 2, 0), v.transpose(0, 1)))
print(json.dumps({"kernel": "seq_outer_sum", "n": 35, "B": 128, "L": 80, "H": 512, "us": round(us, 2), "library_us": round(lib, 2)}), flush=True)

"""cuBLAS (torch.matmul) on the prefill / ViT / mapper GEMM shapes: the library yardstick next to tools/sweep_gemm_bn.py."""
import sys, torch
cases = [(2560, 4800, 1600), (2560, 1600, 1600), (2560, 6400, 1600), (2560, 1600, 6400),
         (3200, 2304, 768), (3200, 768, 768), (3200, 3072, 768), (3200, 768, 3072),
         (5120, 4800, 1600), (5120, 1600, 1600), (5120, 6400, 1600), (5120, 1600, 6400), (8192, 8192, 8192)]
for tokens, features, K in cases:
    n = max(2, int(400e6 // (features * K * 2)))
    W = (torch.randn(n, features, K, device="cuda") * 0.02).bfloat16()
    x = torch.randn(tokens, K, device="cuda").bfloat16()
    out = torch.empty(tokens, features, device="cuda", dtype=torch.bfloat16)
    for _ in range(2):
        for i in range(n):
            torch.matmul(x, W[i].t(), out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        for i in range(n):
            torch.matmul(x, W[i].t(), out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * n)
    print("%5d x %5d x %5d cuBLAS %.1fus/%.0fTF" % (tokens, features, K, us, 2.0 * tokens * features * K / us / 1e6))

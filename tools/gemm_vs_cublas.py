"""The seven prefill / encoder GEMM shapes of DESIGN.md 3.2: this library (auto tile selection, PDL-chained as in the engine)
against cuBLAS (torch.matmul, bf16), both graph-replayed back to back over rotating weights so that neither side is limited by
its host-side launch cost.  python tools/gemm_vs_cublas.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc

cfg = cc.EngineConfig(lm_d=128, lm_layers=1, lm_heads=2, lm_vocab=503, lm_n_pos=64, map_kind="none", vit=False, max_images=8, max_ctx=32)
eng = cc.Engine(cfg)
cases = [(2560, 4800, 1600), (2560, 1600, 1600), (2560, 6400, 1600), (2560, 1600, 6400), (5120, 4800, 1600), (5120, 6400, 1600),
         (3200, 3072, 768), (3200, 2304, 768), (3200, 768, 3072)]


def replay_time(fn, n):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(n):
            fn(i)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n):
                fn(i)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (5 * n)


for tokens, features, K in cases:
    n = max(8, int(300e6 // (features * K * 2)))
    W = (torch.randn(n, features, K, device="cuda") * 0.02).bfloat16()
    x = torch.randn(tokens, K, device="cuda").bfloat16()
    bias = torch.zeros(features, device="cuda")
    out = torch.empty(tokens, features, device="cuda", dtype=torch.bfloat16)
    fl = 2.0 * tokens * features * K
    ours = replay_time(lambda i: eng.op_linear(x, W[i], bias, "none", None, torch.bfloat16), n)
    cub = replay_time(lambda i: torch.matmul(x, W[i].t(), out=out), n)
    print("%5d x %5d x %5d  ours %6.1f us %5.0f TF/s   cuBLAS %6.1f us %5.0f TF/s   ratio %.2f" % (
        tokens, features, K, ours, fl / ours / 1e6, cub, fl / cub / 1e6, cub / ours))
    del W

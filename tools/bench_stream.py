"""Decode-shaped weight-streaming GEMMs of a GPT-J step (16 rows) and the lm_heads, HBM-cold (rotating weights), captured in
a CUDA graph so that launches are back to back: python tools/bench_stream.py [tokens]   (CCB_GEMM_STREAM=0: the
one-tile-per-CTA kernel)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc

tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 16
cases = [(4096, 4096), (12288, 4096), (16384, 4096), (4096, 16384), (50400, 4096), (50257, 1600)]
cfg = cc.EngineConfig(lm_d=128, lm_layers=1, lm_heads=2, lm_vocab=503, lm_n_pos=64, map_kind="none", vit=False, max_images=8, max_ctx=32)
eng = cc.Engine(cfg)
for features, K in cases:
    n = max(2, int(1200e6 // (features * K * 2)))
    W = (torch.randn(n, features, K, device="cuda") * 0.02).bfloat16()
    x = torch.randn(tokens, K, device="cuda").bfloat16()
    bias = torch.zeros(features, device="cuda")
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(n):
            eng.op_linear(x, W[i], bias, "none", None, torch.bfloat16)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n):
                eng.op_linear(x, W[i], bias, "none", None, torch.bfloat16)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * n)
    print("tokens %d features %5d K %5d: %7.2f us  %5.0f GB/s" % (tokens, features, K, us, features * K * 2 / us / 1e3))
    del W, g

"""Tile-width sweep of the persistent GEMMs on the short problems (ViT, attention c_proj), graph-replayed so that the host's
per-call cost (~15 us through ctypes, as long as these kernels) does not hide the kernel time:
python tools/sweep_gemm_graph.py [tokens features K]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc

cfg = cc.EngineConfig(lm_d=128, lm_layers=1, lm_heads=2, lm_vocab=503, lm_n_pos=64, map_kind="none", vit=False, max_images=8, max_ctx=32)
eng = cc.Engine(cfg)
cases = [(3200, 2304, 768), (3200, 768, 768), (3200, 3072, 768), (3200, 768, 3072), (2560, 1600, 1600), (5120, 1600, 1600)]
if len(sys.argv) > 3:
    cases = [(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]))]


def time_one(x, W, bias, orient, bn):
    n = W.shape[0]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for i in range(n):
            eng.op_linear(x, W[i], bias, "none", None, torch.bfloat16, orient, bn, 0)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for i in range(n):
                eng.op_linear(x, W[i], bias, "none", None, torch.bfloat16, orient, bn, 0)
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
    fl = 2.0 * tokens * features * K
    us = time_one(x, W, bias, 1, 0)
    print("%5d x %5d x %5d auto %.1fus/%.0fTF" % (tokens, features, K, us, fl / us / 1e6))
    for orient in (4, 3):
        line = []
        for bn in (256, 224, 192, 160, 128, 96, 64):
            try:
                us = time_one(x, W, bias, orient, bn)
                line.append("%d:%.1f/%.0f" % (bn, us, fl / us / 1e6))
            except Exception as e:
                line.append("%d:err" % bn)
        print("      %s " % ("pair" if orient == 3 else "one ") + "  ".join(line))
    del W

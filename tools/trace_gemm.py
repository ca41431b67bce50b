"""(Needs a library built with `python tools/build.py --tuning`: the timeline stamps are compiled out otherwise.)
Per-CTA timeline of the decode-shaped GEMM (ccb_debug_gemm_trace)."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc
tokens = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cases = [(4800, 1600, 0), (1600, 6400, 0), (50257, 1600, 0), (4800, 1600, 1)]
cfg = cc.EngineConfig(lm_d=128, lm_layers=1, lm_heads=2, lm_vocab=503, lm_n_pos=64, map_kind="none", vit=False, max_images=8, max_ctx=32)
eng = cc.Engine(cfg)
trace = torch.zeros(4096 * 8, dtype=torch.int64, device="cuda")
names = ["entry", "setup", "first_tile", "mma_issued", "acc_ready", "partial", "epi_done", "exit"]
for features, K, split in cases:
    n = 8
    W = (torch.randn(n, features, K, device="cuda") * 0.02).bfloat16()
    x = torch.randn(tokens, K, device="cuda").bfloat16()
    bias = torch.zeros(features, device="cuda")
    for i in range(n):
        eng.op_linear(x, W[i], bias, "none", None, torch.bfloat16, 0, 0, split)
    torch.cuda.synchronize()
    eng.lib.ccb_debug_gemm_trace(eng._h, C.c_void_p(trace.data_ptr()), 0, 1)
    for i in range(3):
        trace.zero_()
        torch.cuda.synchronize()
        eng.op_linear(x, W[i], bias, "none", None, torch.bfloat16, 0, 0, split)
        torch.cuda.synchronize()
    eng.lib.ccb_debug_gemm_trace(eng._h, None, 0, 1)
    t = trace.cpu().view(-1, 8)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = (t - t0).float() / 1e3
    rel[t == 0] = float("nan")
    print("features %d K %d split %s: %d CTAs, kernel span %.2f us" % (features, K, split or "auto", t.shape[0], float((t[:, 7].max() - t0)) / 1e3))
    for k, nm in enumerate(names):
        col = rel[:, k]
        col = col[~torch.isnan(col)]
        if col.numel():
            print("   %-11s min %7.2f  median %7.2f  max %7.2f us (n=%d)" % (nm, col.min(), col.median(), col.max(), col.numel()))
    del W

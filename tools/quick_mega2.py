"""Decode step on the second-generation persistent kernel vs the first and the operator chain (GPT2-XL, synthetic weights).
   python tools/quick_mega2.py [B] [T] [mode]"""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc
from clipcap_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
mode = sys.argv[3] if len(sys.argv) > 3 else "greedy"
beam = 5 if mode == "beam" else 1
cfg = cc.EngineConfig(max_images=B, max_beam=beam, max_ctx=80)
eng = cc.Engine(cfg)
synthetic.load_synthetic(eng)
info = (C.c_int * 4)()
eng.lib.ccb_debug_mega_info(eng._h, info)
print("mega info: v1 ctas %d, v2 ctas %d, enabled %d, v2 enabled %d" % tuple(info))
images = synthetic.synthetic_images(B, cfg, device="cuda")
p = eng.gen_params(mode, T, stop_token=-1, max_stops=0, top_p=0.9 if mode == "sample" else 0.0, beam_size=beam, seed=1)
res = {}
for flag in (0, 2, 3):
    eng.lib.ccb_debug_set_mega(eng._h, flag)
    for it in range(3):
        tokens, lengths, scores = eng.caption_images(images, p)
        torch.cuda.synchronize()
        pre, dec, steps = eng.last_timing()
    print("mega=%d: prefill %.2f ms, decode %.3f ms/step over %d steps" % (flag, pre, dec / max(steps, 1), steps), flush=True)
    res[flag] = tokens.cpu()
for other in (2, 3):
    a, b = res[0], res[other]
    a2, b2 = a.reshape(-1, a.shape[-1]), b.reshape(-1, b.shape[-1])
    print("mega=%d vs chain: rows identical %d / %d; tokens identical %.4f" % (other, int((a2 == b2).all(dim=-1).sum()), a2.shape[0], float((a2 == b2).float().mean())))
print("chain:", res[0].reshape(-1, res[0].shape[-1])[0].tolist())
print("v2   :", res[3].reshape(-1, res[3].shape[-1])[0].tolist())

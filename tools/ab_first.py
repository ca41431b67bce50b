"""Is the FIRST persistent-kernel generate call of a process right?  (GPT2-XL synthetic, greedy)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc
from clipcap_b200 import synthetic
B, T = 64, 6
cfg = cc.EngineConfig(max_images=B, max_beam=1, max_ctx=80)
eng = cc.Engine(cfg)
synthetic.load_synthetic(eng)
images = synthetic.synthetic_images(B, cfg, device="cuda")
p = eng.gen_params("greedy", T, stop_token=-1, max_stops=0)
def run(flag):
    eng.lib.ccb_debug_set_mega(eng._h, flag)
    t, l, s = eng.caption_images(images, p)
    torch.cuda.synchronize()
    return t.cpu()
m1 = run(1); m2 = run(1); r = run(0)
print("first mega == ref:", bool((m1 == r).all()), " second mega == ref:", bool((m2 == r).all()))
print(m1[0].tolist(), m2[0].tolist(), r[0].tolist())

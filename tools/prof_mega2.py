"""A few decode steps on the five-phase kernel for ncu (python tools/prof_mega2.py [B] [T])."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc
from clipcap_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 6
cfg = cc.EngineConfig(max_images=B, max_beam=1, max_ctx=80)
eng = cc.Engine(cfg)
synthetic.load_synthetic(eng)
images = synthetic.synthetic_images(B, cfg, device="cuda")
eng.lib.ccb_debug_set_mega(eng._h, 3)
p = eng.gen_params("greedy", T, stop_token=-1, max_stops=0)
eng.caption_images(images, p)
torch.cuda.synchronize()
print("done", eng.last_timing())

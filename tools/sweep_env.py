"""A/B of one of the persistent decode kernel's environment switches (CCB_MEGA_GRP, ...) on the config-2 path: ms per decode step per value.
   python tools/sweep_env.py [B] [greedy] [v,v,...] [VAR]  (greedy only: the seed is used to force re-capture)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc
from clipcap_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mode = sys.argv[2] if len(sys.argv) > 2 else "greedy"
pfs = [int(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 1]
var = sys.argv[4] if len(sys.argv) > 4 else "CCB_MEGA_GRP"
cfg = cc.EngineConfig(max_images=B, max_beam=1, max_ctx=80)
eng = cc.Engine(cfg)
synthetic.load_synthetic(eng)
torch.cuda.empty_cache()
images = synthetic.synthetic_images(B, cfg, device="cuda")
ref = None
for i, pf in enumerate(pfs):
    os.environ[var] = str(pf)
    # (the seed is part of the graph key: a new one forces a fresh capture, which reads the environment)
    p = eng.gen_params(mode, 32, stop_token=-1, max_stops=0, top_p=0.9 if mode == "sample" else 0.0, seed=100 + i)
    best = 1e9
    for it in range(4):
        tokens, _, _ = eng.caption_images(images, p)
        torch.cuda.synchronize()
        pre, dec, steps = eng.last_timing()
        best = min(best, dec / steps)
    tok = tokens.cpu()
    if ref is None:
        ref = tok
    print(var + " %3d: decode %.3f ms/step (best of 4), tokens equal to pf[0]: %s" % (pf, best, bool((tok == ref).all())), flush=True)

"""Minimal driver for profiling the persistent decode kernel: GPT2-XL synthetic, B rows, T new tokens, `iters` calls."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc
from clipcap_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 3
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 2
cfg = cc.EngineConfig(max_images=B, max_beam=1, max_ctx=80)
eng = cc.Engine(cfg)
synthetic.load_synthetic(eng)
torch.cuda.empty_cache()
images = synthetic.synthetic_images(B, cfg, device="cuda")
p = eng.gen_params("greedy", T, stop_token=-1, max_stops=0)
profile_last = os.environ.get("CCB_PROFILE_LAST") == "1"   # with: ncu --profile-from-start off
for it in range(iters):
    if profile_last and it == iters - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    tokens, lengths, scores = eng.caption_images(images, p)
    torch.cuda.synchronize()
    pre, dec, steps = eng.last_timing()
    print("iter %d: prefill %.2f ms decode %.3f ms/step" % (it, pre, dec / max(steps, 1)))
if profile_last:
    torch.cuda.profiler.stop()
print(tokens[0].tolist())

"""Soak test of the generation paths (GPT2-XL synthetic): many caption calls in alternating modes / batch sizes on one
engine; checks determinism of greedy / beam across repeats and that no call traps (bounded spins turn a protocol bug into a
launch error)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc
from clipcap_b200 import synthetic

rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 20
cfg = cc.EngineConfig(max_images=64, max_beam=5, max_ctx=80)
eng = cc.Engine(cfg)
synthetic.load_synthetic(eng)
images = synthetic.synthetic_images(64, cfg, device="cuda")
ref = {}
t0 = time.time()
calls = 0
for r in range(rounds):
    for mode, n, kw in (("greedy", 64, {}), ("beam", 51, {"beam_size": 5}), ("sample", 37, {"top_p": 0.9, "seed": 3}), ("greedy", 1, {}),
                        ("beam", 64, {"beam_size": 5}), ("sample", 64, {"top_p": 0.8, "typ_p": 0.5, "seed": 4}), ("greedy", 17, {})):
        p = eng.gen_params(mode, 16 + (r % 3) * 8, stop_token=-1, max_stops=0, **kw)
        tok, ln, sc = eng.caption_images(images[:n], p)
        torch.cuda.synchronize()
        calls += 1
        key = (mode, n, p.max_new_tokens)
        t = tok.cpu()
        if key in ref:
            assert torch.equal(ref[key], t), ("non-deterministic result", key, r)
        else:
            ref[key] = t
print("soak ok: %d calls in %.1f s, %d distinct configurations, all repeats bit-identical" % (calls, time.time() - t0, len(ref)))

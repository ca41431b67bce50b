"""Config 5 of BASELINE.json on one GPU's share: ViT-B/32 + 8-layer Transformer mapper (d = 4096) + GPT-J-6B (bf16,
rotary 64, parallel block, untied biased head), top-p 0.9 sampling, 16 images per GPU (128 over 8 GPUs), 32 new tokens."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc
from clipcap_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 32
cfg = cc.EngineConfig(lm_arch="gptj", lm_d=4096, lm_layers=28, lm_heads=16, lm_vocab=50400, lm_n_pos=2048, lm_rotary_dim=64,
                      map_heads=8, max_images=B, max_beam=1, max_ctx=80)
t0 = time.time()
eng = cc.Engine(cfg)
sds = synthetic.load_synthetic(eng)
del sds
torch.cuda.empty_cache()
print("engine ready in %.1fs, device bytes %.2f GB" % (time.time() - t0, eng.device_bytes / 1e9))
images = synthetic.synthetic_images(B, cfg, device="cuda")
p = eng.gen_params("sample", T, stop_token=-1, max_stops=0, top_p=0.9, seed=1)
profile_last = os.environ.get("CCB_PROFILE_LAST") == "1"
for it in range(3):
    if profile_last and it == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tokens, lengths, scores = eng.caption_images(images, p)
    e1.record()
    torch.cuda.synchronize()
    pre, dec, steps = eng.last_timing()
    ms = e0.elapsed_time(e1)
    print("iter %d: total %.2f ms (%.0f captions/s) prefill+first %.2f ms decode %.2f ms (%d steps, %.3f ms/step)" % (
        it, ms, B / ms * 1e3, pre, dec, steps, dec / max(steps, 1)))
if profile_last:
    torch.cuda.profiler.stop()
print(tokens[0].tolist()[:16])

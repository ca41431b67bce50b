"""Quick timing of the config-2 path (ViT-B/32 + mapper + GPT2-XL greedy, B=64, 32 new tokens) on one GPU."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc
from clipcap_b200 import synthetic

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
mode = sys.argv[2] if len(sys.argv) > 2 else "greedy"
T = int(sys.argv[3]) if len(sys.argv) > 3 else 32
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4
beam = 5 if mode == "beam" else 1
all_feats = len(sys.argv) > 5 and sys.argv[5] == "all"   # use_all_vit_features: 50 ViT tokens -> TransformerMapperAllFeatures
cfg = cc.EngineConfig(max_images=B, max_beam=beam, max_ctx=80, **(dict(map_kind="transformer_all", map_clip_len=50) if all_feats else {}))
t0 = time.time()
eng = cc.Engine(cfg)
sds = synthetic.load_synthetic(eng)
del sds
torch.cuda.empty_cache()
print("engine ready in %.1fs, device bytes %.2f GB" % (time.time() - t0, eng.device_bytes / 1e9))
images = synthetic.synthetic_images(B, cfg, device="cuda")
p = eng.gen_params(mode, T, stop_token=-1, max_stops=0, top_p=0.9 if mode == "sample" else 0.0, beam_size=beam, seed=1)
profile_last = os.environ.get("CCB_PROFILE_LAST") == "1"
for it in range(iters):
    if profile_last and it == iters - 1:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    torch.cuda.synchronize()
    l0 = eng.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    tokens, lengths, scores = eng.caption_images(images, p)
    e1.record()
    torch.cuda.synchronize()
    pre, dec, steps = eng.last_timing()
    ms = e0.elapsed_time(e1)
    print("iter %d: total %.2f ms (%.0f captions/s) prefill+first %.2f ms decode %.2f ms (%d steps, %.3f ms/step) launches %d" % (
        it, ms, B / ms * 1e3, pre, dec, steps, dec / max(steps, 1), eng.launch_count - l0))
if profile_last:
    torch.cuda.profiler.stop()
print(tokens[0].tolist()[:16], "checksum", int((tokens.to(torch.int64) * torch.arange(1, tokens.numel() + 1, device=tokens.device).view_as(tokens)).sum().item()))

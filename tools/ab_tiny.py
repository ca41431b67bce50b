"""A/B of the persistent decode kernel against the operator-per-kernel chain on the tiny golden fixture."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import clipcap_b200 as cc

fx = torch.load(os.path.join(ROOT, "tests", "golden", "tiny_gpt2.pt"), weights_only=False)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 3
PT = int(sys.argv[2]) if len(sys.argv) > 2 else 4
cfg = cc.EngineConfig(
    lm_arch="gpt2", lm_d=fx["d"], lm_layers=2, lm_heads=fx["heads"], lm_vocab=fx["V"], lm_n_pos=64,
    map_dim_clip=fx["dim_clip"], map_clip_len=fx["CL"], map_prefix_len=fx["P"], map_heads=fx["map_heads"],
    map_layers=2, vit_image=fx["vit_image"], vit_patch=fx["vit_patch"], vit_width=fx["vit_width"],
    vit_layers=fx["vit_layers"], vit_heads=fx["vit_heads"], vit_out=fx["dim_clip"], max_images=32, max_beam=5,
    max_ctx=32, page_tokens=PT)
eng = cc.Engine(cfg, 0)
eng.load_state_dict(fx["sd_lm"], prefix="language_model.")
eng.load_state_dict(fx["sd_mapper"], prefix="clip_project.")
eng.load_state_dict(fx["sd_vit"], prefix="visual.")
prefix = fx["prefix"].cuda()
if N > prefix.shape[0]:
    prefix = prefix.repeat((N + 2) // 3, 1, 1)[:N].contiguous()
import ctypes as C
d = fx["d"]
T = int(sys.argv[3]) if len(sys.argv) > 3 else 10
def grab(which, numel, dtype):
    out = torch.empty(numel, dtype=dtype, device="cuda")
    eng.lib.ccb_debug_copy_buffer(eng._h, which, C.c_void_p(out.data_ptr()), out.numel() * out.element_size(), None)
    torch.cuda.synchronize()
    return out.float().cpu()
for mode, kw in (("greedy", {}), ("beam", {"beam_size": 5})):
    res = {}
    bufs = {}
    for flag in (0, 1):
        print("covered:", eng.lib.ccb_debug_set_mega(eng._h, flag))
        p = eng.gen_params(mode, T, stop_token=-1, max_stops=0, **kw)
        for rep in range(int(os.environ.get("REPS", "1"))):
            t, l, s = eng.generate(prefix, p)
            torch.cuda.synchronize()
        res[flag] = t.cpu()
        R = t.numel() // T
        bufs[flag] = {n: grab(w, R * k, dt) for n, w, k, dt in (("h", 0, d, torch.float32), ("x", 1, d, torch.bfloat16), ("att", 2, d, torch.bfloat16), ("mlp", 3, 4 * d, torch.bfloat16))}
    for n in bufs[0]:
        a, b = bufs[0][n], bufs[1][n]
        print("  %-4s max|off| %.4f  max|diff| %.5f" % (n, float(a.abs().max()), float((a - b).abs().max())))
        if n == "att":
            dd = (a - b).abs().view(-1, fx["heads"], 64)
            print("     per (row, head) max diff:", [[round(float(v), 3) for v in r] for r in dd.amax(dim=-1)][:6])
            print("     off row0 head0[:8]", [round(float(v), 3) for v in a.view(-1, 64)[0, :8]])
            print("     on  row0 head0[:8]", [round(float(v), 3) for v in b.view(-1, 64)[0, :8]])
    print(mode, "identical:", bool((res[0] == res[1]).all()))
    print(" off", res[0].reshape(-1, T)[:4].tolist())
    print(" on ", res[1].reshape(-1, T)[:4].tolist())

import sys, os, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import clipcap_b200 as cc
fx = torch.load(os.path.join(ROOT, "tests", "golden", "tiny_gpt2.pt"), weights_only=False)
cfg = cc.EngineConfig(
    lm_arch="gpt2", lm_d=fx["d"], lm_layers=2, lm_heads=fx["heads"], lm_vocab=fx["V"], lm_n_pos=64,
    map_dim_clip=fx["dim_clip"], map_clip_len=fx["CL"], map_prefix_len=fx["P"], map_heads=fx["map_heads"],
    map_layers=2, vit_image=fx["vit_image"], vit_patch=fx["vit_patch"], vit_width=fx["vit_width"],
    vit_layers=fx["vit_layers"], vit_heads=fx["vit_heads"], vit_out=fx["dim_clip"], max_images=32, max_beam=5,
    max_ctx=32, page_tokens=4)
eng = cc.Engine(cfg, 0)
eng.load_state_dict(fx["sd_lm"], prefix="language_model.")
eng.load_state_dict(fx["sd_mapper"], prefix="clip_project.")
eng.load_state_dict(fx["sd_vit"], prefix="visual.")
A = fx["prefix"].cuda()
d = fx["d"]; R = 3
p = eng.gen_params("greedy", 2, stop_token=-1, max_stops=0)
def grab(which, numel, dtype):
    out = torch.empty(numel, dtype=dtype, device="cuda")
    eng.lib.ccb_debug_copy_buffer(eng._h, which, C.c_void_p(out.data_ptr()), out.numel() * out.element_size(), None)
    torch.cuda.synchronize()
    return out.float().cpu()
def run():
    eng.lib.ccb_debug_set_mega(eng._h, 1)
    t, l, s = eng.generate(A, p)
    torch.cuda.synchronize()
    return {n: grab(w, R * k, dt) for n, w, k, dt in (("h", 0, d, torch.float32), ("x", 1, d, torch.bfloat16), ("att", 2, d, torch.bfloat16), ("mlp", 3, 4 * d, torch.bfloat16))}
a = run(); b = run(); c = run()
for n in a:
    print("%-4s first vs second: %.5f   second vs third: %.5f   max|second| %.3f" % (n, float((a[n] - b[n]).abs().max()), float((b[n] - c[n]).abs().max()), float(b[n].abs().max())))
dd = (a["att"] - b["att"]).abs().view(R, -1, 64).amax(-1)
print("att diff per (row, head):", dd.tolist())
dm = (a["mlp"] - b["mlp"]).abs().view(R, -1, 128).amax(-1)
print("mlp diff per (row, 128-col tile):", dm.tolist())
dh = (a["h"] - b["h"]).abs().view(R, -1, 64).amax(-1)
print("h diff per (row, 64-col block):", dh.tolist())
# second engine in the same process: is ITS first call right?
eng2 = cc.Engine(cfg, 0)
eng2.load_state_dict(fx["sd_lm"], prefix="language_model.")
eng2.load_state_dict(fx["sd_mapper"], prefix="clip_project.")
eng2.load_state_dict(fx["sd_vit"], prefix="visual.")
eng_old = eng
eng = eng2
a2 = run(); b2 = run()
print("second engine: first vs second call att diff %.5f; vs first engine's steady att %.5f" % (float((a2["att"] - b2["att"]).abs().max()), float((b2["att"] - b["att"]).abs().max())))

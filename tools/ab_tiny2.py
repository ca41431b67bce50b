import sys, os
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
torch.manual_seed(0)
B = (A + 0.5 * torch.randn_like(A)).contiguous()
T = 6
p = eng.gen_params("greedy", T, stop_token=-1, max_stops=0)
def run(flag, x):
    eng.lib.ccb_debug_set_mega(eng._h, flag)
    t, l, s = eng.generate(x, p)
    torch.cuda.synchronize()
    return t.cpu().tolist()
order = sys.argv[1] if len(sys.argv) > 1 else "mega_first"
if order == "mega_first":
    print("mega A #1", run(1, A)); print("mega B #2", run(1, B)); print("mega A #3", run(1, A)); print("mega B #4", run(1, B))
    print("ref  A   ", run(0, A)); print("ref  B   ", run(0, B))
else:
    print("ref  A   ", run(0, A)); print("ref  B   ", run(0, B))
    print("mega A #1", run(1, A)); print("mega B #2", run(1, B)); print("mega A #3", run(1, A))

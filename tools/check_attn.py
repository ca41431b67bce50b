import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import clipcap_b200 as cc
eng = cc.Engine(cc.EngineConfig(lm_layers=1, map_kind="none", vit=False, max_images=4, max_ctx=32, lm_vocab=512))
torch.manual_seed(0)
for (B, S, H, hd, causal) in [(3, 5, 2, 32, False), (3, 8, 8, 16, False), (3, 8, 8, 24, False), (3, 4, 2, 64, True), (2, 9, 25, 64, True), (2, 17, 3, 64, True), (64, 40, 25, 64, True), (2, 50, 12, 64, False), (2, 80, 8, 200, False), (1, 128, 2, 128, True)]:
    d = H * hd
    qkv = torch.randn(B * S, 3 * d, device="cuda").bfloat16()
    out = eng.op_attention(qkv, B, S, H, hd, causal=causal).float()
    q, k, v = [t.view(B, S, H, hd).transpose(1, 2) for t in qkv.float().split(d, dim=-1)]
    sc = (q @ k.transpose(-1, -2)) / hd ** 0.5
    if causal:
        sc = sc.masked_fill(torch.triu(torch.ones(S, S, dtype=torch.bool, device="cuda"), 1), float("-inf"))
    ref = (sc.softmax(-1) @ v).transpose(1, 2).reshape(B * S, d)
    err = (out - ref).abs().max().item() / ref.abs().max().item()
    print((B, S, H, hd, causal), "rel err %.4f" % err, "nan" if torch.isnan(out).any() else "")

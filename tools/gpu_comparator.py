"""The bar on the same box: the reference's generation loop in PyTorch eager
bf16 on the B200, against this library, on config 2's language-model part (GPT2-XL, prefix 40, 32 new tokens, batch 64).

The reference's language model IS `transformers.GPT2LMHeadModel` (lms/GPT2.py:6), called with `inputs_embeds` and no cache
(lms/GPT2.py:17-19); its loop (inference.py:70-148 with beam_size 1, inference.py:219-292) re-runs the whole sequence for
every token, one image at a time.  Three eager variants on the same weights and the same prefix embeddings:

  ref_loop      the reference as written: batch 1, no KV cache, full re-forward per token (timed on a few images, scaled)
  batched       the same re-forward, all 64 rows at once
  hf_cache      all 64 rows, HF `use_cache=True` (past_key_values) -- the best stock eager path

and this library's prefill + decode for the same 64 prefixes (Engine.generate).  CUDA events, after warm-up.
    python tools/gpu_comparator.py [out.json]
"""
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import transformers
import clipcap_b200 as cc
from clipcap_b200 import synthetic

B, T, P = 64, 32, 40
cfg = cc.EngineConfig(max_images=B, max_beam=1, max_ctx=80)
eng = cc.Engine(cfg)
sds = synthetic.load_synthetic(eng)
images = synthetic.synthetic_images(B, cfg, device="cuda")
prefix = eng.map_prefix(eng.vit_encode(images)).clone()          # [64, 40, 1600] f32

hf_cfg = transformers.GPT2Config(vocab_size=cfg.lm_vocab, n_positions=cfg.lm_n_pos, n_embd=cfg.lm_d, n_layer=cfg.lm_layers, n_head=cfg.lm_heads)
with torch.device("cuda"):
    hf = transformers.GPT2LMHeadModel(hf_cfg)
hf.load_state_dict(sds["lm"], strict=False)
hf.tie_weights()
hf = hf.to(torch.bfloat16).eval()
del sds
torch.cuda.empty_cache()
wte = hf.transformer.wte.weight
pb = prefix.to(torch.bfloat16)


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        out = fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out


@torch.no_grad()
def ref_loop(n_images):
    toks = []
    for i in range(n_images):                      # inference.py: one image at a time
        emb, row = pb[i:i + 1], []
        for _ in range(T):
            logits = hf(inputs_embeds=emb).logits[:, -1, :]          # full re-forward (no cache)
            nxt = logits.float().argmax(-1)
            row.append(nxt)
            emb = torch.cat((emb, wte[nxt][:, None, :]), dim=1)
        toks.append(torch.stack(row, 1))
    return torch.cat(toks)


@torch.no_grad()
def batched():
    emb, rows = pb, []
    for _ in range(T):
        nxt = hf(inputs_embeds=emb).logits[:, -1, :].float().argmax(-1)
        rows.append(nxt)
        emb = torch.cat((emb, wte[nxt][:, None, :]), dim=1)
    return torch.stack(rows, 1)


@torch.no_grad()
def hf_cache():
    out = hf(inputs_embeds=pb, use_cache=True)
    past, rows = out.past_key_values, []
    nxt = out.logits[:, -1, :].float().argmax(-1)
    rows.append(nxt)
    for _ in range(T - 1):
        out = hf(inputs_embeds=wte[nxt][:, None, :], past_key_values=past, use_cache=True)
        past = out.past_key_values
        nxt = out.logits[:, -1, :].float().argmax(-1)
        rows.append(nxt)
    return torch.stack(rows, 1)


def ours():
    p = eng.gen_params("greedy", T, stop_token=-1, max_stops=0)
    return eng.generate(prefix, p)[0]


res = {"config": "GPT2-XL language-model part of config 2: 64 prefixes x 40, 32 new tokens, greedy, bf16", "gpu": torch.cuda.get_device_name(0),
       "torch": torch.__version__, "transformers": transformers.__version__}
n_ref = 2
ms, _ = timed(lambda: ref_loop(n_ref), 1)
res["ref_loop"] = {"ms_per_64": ms * B / n_ref, "captions_per_s": 1e3 * n_ref / ms, "note": "timed on %d images, scaled to 64 (cost is linear in images)" % n_ref}
ms, tb = timed(batched, 2)
res["batched_no_cache"] = {"ms_per_64": ms, "captions_per_s": 1e3 * B / ms}
ms, tc = timed(hf_cache, 3)
res["hf_use_cache"] = {"ms_per_64": ms, "captions_per_s": 1e3 * B / ms}
ms, to = timed(ours, 5)
res["clipcap_b200"] = {"ms_per_64": ms, "captions_per_s": 1e3 * B / ms}
res["speedup_vs_ref_loop"] = res["ref_loop"]["ms_per_64"] / ms
res["speedup_vs_hf_use_cache"] = res["hf_use_cache"]["ms_per_64"] / ms
res["rows_identical_to_hf_use_cache"] = int((to.long() == tc).all(-1).sum())
res["first_token_identical_to_hf_use_cache"] = int((to.long()[:, 0] == tc[:, 0]).sum())
print(json.dumps(res, indent=1))
if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        json.dump(res, f, indent=1)

"""Non-trivial LayerNorm gains / offsets and linear biases.

The reference-generated fixtures keep HF's default initialisation (LayerNorm weight 1 / bias 0, zero linear biases), so a
wrong bias vector or a swapped gamma / beta would not show there.  Here every 1-D parameter of the fixture state dicts is
re-drawn (LayerNorm gains 1 + 0.3 N(0,1), offsets and biases 0.3 N(0,1), rounded to bf16):
  * CPU: the oracle against the classes the reference instantiates (lms/GPT2.py:6 GPT2LMHeadModel, lms/GPTJ.py:5
    GPTJForCausalLM from the installed transformers) on those weights -- pins the oracle's use of every such vector;
  * GPU: the CUDA path (persistent decode kernel and operator chain) against the oracle on the same weights.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import clipcap_oracle as orc  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
TOL = 2e-2


def perturbed(sd, seed):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in sd.items():
        if torch.is_tensor(v) and v.is_floating_point() and v.dim() == 1 and ("ln" in k or "norm" in k or k.endswith("bias")):
            gain = k.endswith("weight")
            v = ((1.0 if gain else 0.0) + 0.3 * torch.randn(v.shape, generator=g)).to(torch.bfloat16)
        out[k] = v
    return out


def f32(sd):
    return {k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v) for k, v in sd.items()}


def load_fixture(arch):
    fx = torch.load(os.path.join(GOLDEN, "tiny_%s.pt" % arch), weights_only=False)
    fx["sd_lm"] = perturbed(fx["sd_lm"], 101)
    fx["sd_mapper"] = perturbed(fx["sd_mapper"], 102)
    fx["sd_vit"] = perturbed(fx["sd_vit"], 103)
    n = sum(1 for k, v in fx["sd_lm"].items() if v.dim() == 1)
    assert n >= 10
    return fx


@pytest.mark.parametrize("arch", ["gpt2", "gptj"])
def test_oracle_lm_against_transformers_with_random_affine_params(arch):
    import transformers
    fx = load_fixture(arch)
    if arch == "gpt2":
        hf = transformers.GPT2LMHeadModel(transformers.GPT2Config(vocab_size=fx["V"], n_positions=64, n_embd=fx["d"], n_layer=2,
                                                                  n_head=fx["heads"]))
    else:
        hf = transformers.GPTJForCausalLM(transformers.GPTJConfig(vocab_size=fx["V"], n_positions=64, n_embd=fx["d"], n_layer=2,
                                                                  n_head=fx["heads"], rotary_dim=fx["rotary_dim"]))
    res = hf.load_state_dict(f32(fx["sd_lm"]), strict=False)
    assert not res.unexpected_keys
    assert all(("attn.bias" in k) or ("masked_bias" in k) or ("embed_positions" in k) or k == "lm_head.weight" for k in res.missing_keys), res.missing_keys
    hf = hf.float().eval()
    lm = orc.OracleLM(f32(fx["sd_lm"]), arch, fx["heads"], fx["rotary_dim"])
    emb = torch.cat((fx["prefix"], lm.get_embedding_text(fx["tokens"])), dim=1)
    mask = torch.cat((torch.ones(3, fx["P"], dtype=torch.bool), fx["mask"]), dim=1)
    with torch.no_grad():
        want = hf(inputs_embeds=emb, attention_mask=mask.long()).logits
    got = lm.logits(emb, mask)
    assert (got - want).abs().max().item() <= 5e-5 * want.abs().max().item()


@pytest.mark.gpu
@pytest.mark.parametrize("arch", ["gpt2", "gptj"])
def test_cuda_path_with_random_affine_params(arch):
    import clipcap_b200 as cc
    fx = load_fixture(arch)
    cfg = cc.EngineConfig(
        lm_arch=fx["arch"], lm_d=fx["d"], lm_layers=2, lm_heads=fx["heads"], lm_vocab=fx["V"], lm_n_pos=64,
        lm_rotary_dim=fx["rotary_dim"], map_dim_clip=fx["dim_clip"], map_clip_len=fx["CL"], map_prefix_len=fx["P"],
        map_heads=fx["map_heads"], map_layers=2, vit_image=fx["vit_image"], vit_patch=fx["vit_patch"],
        vit_width=fx["vit_width"], vit_layers=fx["vit_layers"], vit_heads=fx["vit_heads"], vit_out=fx["dim_clip"],
        max_images=32, max_beam=5, max_ctx=32, max_lm_tokens=32 * 16, page_tokens=4)
    eng = cc.Engine(cfg)
    eng.load_state_dict(fx["sd_lm"], prefix="language_model.")
    eng.load_state_dict(fx["sd_mapper"], prefix="clip_project.")
    eng.load_state_dict(fx["sd_vit"], prefix="visual.")
    eng.check_weights()
    lm = orc.OracleLM(f32(fx["sd_lm"]), arch, fx["heads"], fx["rotary_dim"])

    def rel(a, b):
        return (a.float().cpu() - b).abs().max().item() / max(b.abs().max().item(), 1e-9)

    feat_ref = orc.vit_forward(f32(fx["sd_vit"]), fx["images"], fx["vit_heads"], fx["vit_patch"])
    assert rel(eng.vit_encode(fx["images"]), feat_ref) <= TOL
    prefix_ref = orc.mapper_forward(f32(fx["sd_mapper"]), feat_ref, fx["CL"], fx["map_heads"])
    assert rel(eng.map_prefix(feat_ref), prefix_ref) <= TOL
    emb = torch.cat((prefix_ref, lm.get_embedding_text(fx["tokens"])), dim=1)
    mask = torch.cat((torch.ones(3, fx["P"], dtype=torch.bool), fx["mask"]), dim=1)
    assert rel(eng.lm_forward(emb, attention_mask=mask), lm.logits(emb, mask)) <= TOL
    # KV-cached greedy decode, persistent kernel (GPT-2) and operator chain: every step's choice must be the oracle's arg-max
    # of the teacher-forced logits up to near-ties below the tolerance
    T = 10
    for mega in (1, 0):
        covered = eng.lib.ccb_debug_set_mega(eng._h, mega)
        p = eng.gen_params("greedy", T, stop_token=-1, max_stops=0)
        tokens, _, _ = eng.generate(prefix_ref, p)
        tokens = tokens.cpu().long()
        full = torch.cat((prefix_ref, lm.get_embedding_text(tokens[:, :T - 1])), dim=1)
        pred = lm.logits(full)[:, fx["P"] - 1:]
        agree = pred.argmax(-1) == tokens
        scale = (pred.max() - pred.min()).item()
        for r, t in (~agree).nonzero().tolist():
            assert (pred[r, t].max() - pred[r, t, tokens[r, t]]).item() <= TOL * scale, (arch, mega, covered, r, t)
        assert agree.float().mean().item() >= 0.9
    eng.close()

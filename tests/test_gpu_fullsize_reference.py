"""Parity at BASELINE.json's FULL sizes (config 2: ViT-B/32 + 8-layer mapper (d = 1600, S = 80) + GPT2-XL) against fp32
references evaluated on the GPU box itself:

  * the language model against `transformers.GPT2LMHeadModel` -- the very class lms/GPT2.py:6 instantiates and
    lms/GPT2.py:17-19 calls -- in fp32 on the same bf16-rounded weights: logits within the north-star tolerance
    (max |delta| <= 2e-2 max |ref|) and greedy tokens compared position by position on the reference's own histories;
  * ViT features and prefix embeddings against the fp32 oracle restatement (oracle/clipcap_oracle.py, pinned to the
    reference modules by tests/test_oracle_golden.py) run on CUDA tensors.

Random-init weights give top-1 margins that can be below the bf16 noise of the activations, so token
agreement is counted per position with teacher forcing (a flip cannot cascade) and every disagreement must be a near-tie
in the reference logits.
"""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import clipcap_oracle as orc  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 2e-2


@pytest.fixture(scope="module")
def full():
    import clipcap_b200 as cc
    from clipcap_b200 import synthetic
    cfg = cc.EngineConfig(max_images=32, max_beam=1, max_ctx=80)
    eng = cc.Engine(cfg)
    sds = synthetic.load_synthetic(eng)          # fp32 tensors holding bf16-rounded values, on the device
    images = synthetic.synthetic_images(32, cfg, device="cuda")
    yield eng, cfg, sds, images
    eng.close()


def rel(a, b):
    return (a.float() - b.float()).abs().max().item() / max(b.float().abs().max().item(), 1e-9)


def test_vit_and_mapper_match_the_fp32_oracle_at_full_size(full):
    eng, cfg, sds, images = full
    feat = eng.vit_encode(images[:8])
    ref_feat = orc.vit_forward(sds["vit"], images[:8], cfg.vit_heads, cfg.vit_patch)
    assert rel(feat, ref_feat) <= TOL
    prefix = eng.map_prefix(ref_feat)            # same input for both sides: isolates the mapper
    ref_prefix = orc.mapper_forward(sds["mapper"], ref_feat, cfg.map_clip_len, cfg.map_heads, "relu")
    assert prefix.shape == ref_prefix.shape == (8, cfg.map_prefix_len, cfg.lm_d)
    assert rel(prefix, ref_prefix) <= TOL


def test_gpt2_xl_against_transformers_fp32(full):
    transformers = pytest.importorskip("transformers")
    eng, cfg, sds, images = full
    hf_cfg = transformers.GPT2Config(vocab_size=cfg.lm_vocab, n_positions=cfg.lm_n_pos, n_embd=cfg.lm_d, n_layer=cfg.lm_layers,
                                     n_head=cfg.lm_heads)
    with torch.device("cuda"):
        hf = transformers.GPT2LMHeadModel(hf_cfg)            # what lms/GPT2.py:6 builds
    missing = hf.load_state_dict(sds["lm"], strict=False)
    assert all(k.endswith(("attn.bias", "attn.masked_bias", "lm_head.weight")) for k in missing.missing_keys), missing
    assert not missing.unexpected_keys
    hf.tie_weights()
    hf = hf.float().eval()

    N, T = 32, 16
    prefix = eng.map_prefix(eng.vit_encode(images[:N]))       # [32, 40, 1600]
    p = eng.gen_params("greedy", T, stop_token=-1, max_stops=0)
    tokens, _, _ = eng.generate(prefix, p)
    torch.cuda.synchronize()
    tokens = tokens.long()
    emb = torch.cat([prefix, eng.embed_tokens(tokens[:, :T - 1])], dim=1)       # [32, 55, 1600]
    with torch.no_grad():
        ref_logits = hf(inputs_embeds=emb).logits.float()                       # lms/GPT2.py:17-19
    # (a6) teacher-forced logits of the CUDA path over the same embeddings
    ours = eng.lm_forward(emb)
    assert rel(ours, ref_logits) <= TOL
    # greedy tokens of the KV-cached decode against the reference's argmax on the same history
    P = prefix.shape[1]
    pred = ref_logits[:, P - 1:P - 1 + T]
    agree = pred.argmax(-1) == tokens
    frac = agree.float().mean().item()
    scale = (pred.max() - pred.min()).item()
    bad = (~agree).nonzero()
    for r, t in bad.tolist():
        margin = (pred[r, t].max() - pred[r, t, tokens[r, t]]).item()
        assert margin <= TOL * scale, (r, t, margin, scale)   # only near-ties may differ
    assert frac >= 0.99, frac        # (per-image numbers at 256 images x 32 tokens: tests/test_gpu_config_parity.py)
    print("greedy tokens identical to transformers fp32 at %.2f %% of %d positions" % (100 * frac, agree.numel()))


def test_beam_search_against_the_reference_loop_at_full_size(full):
    """generate_beam (inference.py:70-148, restated in the oracle: batch-1 loop in fp32 on the GPU, with the oracle's own
    KV cache -- numerically the same forward, pinned in tests/test_oracle_golden.py) for GPT2-XL, beam 5, one image at a
    time, against the on-device beam search over a batch of images."""
    eng_small, cfg, sds, images = full
    import clipcap_b200 as cc
    cfg_b = cc.EngineConfig(max_images=8, max_beam=5, max_ctx=80)
    eng = cc.Engine(cfg_b)
    for k, pre in (("lm", "language_model."), ("mapper", "clip_project."), ("vit", "visual.")):
        eng.load_state_dict(sds[k], prefix=pre)
    eng.check_weights()
    N, T = 6, 8
    prefix = eng.map_prefix(eng.vit_encode(images[:N]))
    p = eng.gen_params("beam", T, stop_token=-1, max_stops=0, beam_size=5)
    tok, ln, sc = eng.generate(prefix, p)
    torch.cuda.synchronize()
    tok, sc = tok.cpu(), sc.cpu()
    lm = orc.OracleLM(sds["lm"], "gpt2", cfg.lm_heads)
    same_best, same_set = 0, 0
    for i in range(N):
        with torch.no_grad():
            rt, rl, rs, order = orc.generate_beam(lm, prefix[i:i + 1].float(), beam_size=5, entry_length=T, stop_token=-1,
                                                  use_cache=True)
        rt, rs, order = rt.cpu(), rs.cpu(), order.cpu()
        best_ref = rt[order[0]].tolist()
        best = tok[i, int(sc[i].argmax())].tolist()
        same_best += best == best_ref
        same_set += sorted(map(tuple, tok[i].tolist())) == sorted(map(tuple, rt.tolist()))
        if best == best_ref:
            assert abs(float(sc[i].max()) - float(rs[order[0]])) <= 2e-2 * max(1.0, abs(float(rs[order[0]])))
    eng.close()
    print("beam 5, GPT2-XL: best caption identical for %d / %d images, all five beams for %d / %d" % (same_best, N, same_set, N))
    # random-init log-probabilities are nearly flat, so losing beams swap on near-ties; the winner must hold
    assert same_best >= N - 1


def test_config5_mapper_4096_wide_at_full_size():
    """The mapper of BASELINE config 5 (GPT-J-6B: embedding width 4096, 8 heads -> head_dim 512, S = 40 + 40) against the fp32
    oracle: prefix embeddings within 2e-2 (the head_dim-512 attention takes the scalar prefill kernel)."""
    import clipcap_b200 as cc
    from clipcap_b200 import synthetic
    cfg = cc.EngineConfig(lm_arch="gptj", lm_d=4096, lm_layers=1, lm_heads=16, lm_vocab=50400, lm_n_pos=2048, lm_rotary_dim=64,
                          map_heads=8, max_images=8, max_beam=1, max_ctx=80)
    eng = cc.Engine(cfg)
    sds = synthetic.load_synthetic(eng)
    images = synthetic.synthetic_images(8, cfg, device="cuda")
    feat = eng.vit_encode(images)
    prefix = eng.map_prefix(feat)
    ref = orc.mapper_forward(sds["mapper"], feat.float(), cfg.map_clip_len, cfg.map_heads, "relu")
    assert prefix.shape == ref.shape == (8, cfg.map_prefix_len, 4096)
    assert rel(prefix, ref) <= TOL
    # ... and the prefix drives the language model: first-token logits of the one-layer GPT-J against the oracle
    lm = orc.OracleLM(sds["lm"], "gptj", cfg.lm_heads, cfg.lm_rotary_dim)
    ours = eng.lm_forward(prefix, last_only=True)
    with torch.no_grad():
        want = lm.logits(ref)[:, -1]
    assert rel(ours, want) <= TOL
    eng.close()


def test_clip_text_tower_at_full_size():
    """CLIP ViT-B/32 text tower (49408 x 512, 77 tokens, 12 layers, 8 heads) against the fp32 oracle on the GPU."""
    import clipcap_b200 as cc
    from clipcap_b200 import synthetic
    cfg = cc.EngineConfig(lm_d=128, lm_layers=1, lm_heads=2, lm_vocab=503, lm_n_pos=64, map_kind="none", vit=False, max_images=4,
                          max_ctx=32, text=True, max_texts=40)
    eng = cc.Engine(cfg)
    sds = synthetic.load_synthetic(eng)
    g = torch.Generator().manual_seed(3)
    B = 40
    tokens = torch.zeros(B, 77, dtype=torch.int64)
    for b in range(B):
        n = 2 + (b * 7) % 74
        tokens[b, 0] = 49406
        tokens[b, 1:n] = torch.randint(0, 49406, (n - 1,), generator=g)
        tokens[b, n] = 49407
    feats = eng.clip_encode_text(tokens)
    ref = orc.clip_text_forward(sds["text"], tokens.cuda(), cfg.text_heads)
    assert feats.shape == ref.shape == (B, 512)
    assert rel(feats, ref) <= TOL
    eng.close()


def test_all_vit_features_path_at_full_size():
    """use_all_vit_features at ViT-B/32 + d = 1600 sizes: 50 projected ViT tokens per image and the
    TransformerMapperAllFeatures mapper (S = 90) against the fp32 oracle on the GPU."""
    import clipcap_b200 as cc
    from clipcap_b200 import synthetic
    cfg = cc.EngineConfig(lm_layers=1, lm_vocab=2048, map_kind="transformer_all", map_clip_len=50, max_images=8, max_ctx=64)
    eng = cc.Engine(cfg)
    sds = synthetic.load_synthetic(eng)
    images = synthetic.synthetic_images(8, cfg, device="cuda")
    toks = eng.vit_encode(images)
    ref_toks = orc.vit_forward(sds["vit"], images, cfg.vit_heads, cfg.vit_patch, all_tokens=True)
    assert toks.shape == ref_toks.shape == (8, 50, 512)
    assert rel(toks, ref_toks) <= TOL
    prefix = eng.map_prefix(ref_toks)
    ref_prefix = orc.mapper_all_forward(sds["mapper"], ref_toks, cfg.map_heads, "relu")
    assert prefix.shape == ref_prefix.shape == (8, 40, 1600)
    assert rel(prefix, ref_prefix) <= TOL
    eng.close()


def test_gptj_6b_against_transformers_fp32():
    """Config 5's language model (GPT-J-6B shapes: d = 4096, 28 layers, 16 heads of 256, rotary 64, untied biased head)
    against `transformers.GPTJForCausalLM` in fp32 -- the class lms/GPTJ.py:5 instantiates -- on the same bf16-rounded
    weights: teacher-forced logits within 2e-2, greedy tokens of the KV-cached decode equal to the reference's argmax."""
    transformers = pytest.importorskip("transformers")
    import clipcap_b200 as cc
    from clipcap_b200 import synthetic
    cfg = cc.EngineConfig(lm_arch="gptj", lm_d=4096, lm_layers=28, lm_heads=16, lm_vocab=50400, lm_n_pos=2048, lm_rotary_dim=64,
                          map_kind="none", vit=False, max_images=8, max_ctx=48)
    eng = cc.Engine(cfg)
    sd = synthetic.lm_state_dict(cfg, 1234, "cuda")
    eng.load_state_dict(sd, prefix="language_model.")
    eng.check_weights()
    hf_cfg = transformers.GPTJConfig(vocab_size=cfg.lm_vocab, n_positions=cfg.lm_n_pos, n_embd=cfg.lm_d, n_layer=cfg.lm_layers,
                                     n_head=cfg.lm_heads, rotary_dim=cfg.lm_rotary_dim)
    with torch.device("cuda"):
        hf = transformers.GPTJForCausalLM(hf_cfg)
    res = hf.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys
    assert all(("attn.bias" in k) or ("masked_bias" in k) or ("embed_positions" in k) for k in res.missing_keys), res.missing_keys
    hf = hf.float().eval()
    N, P, T = 8, 12, 8
    torch.manual_seed(5)
    prefix = (0.05 * torch.randn(N, P, cfg.lm_d, device="cuda")).contiguous()
    p = eng.gen_params("greedy", T, stop_token=-1, max_stops=0)
    tokens, _, _ = eng.generate(prefix, p)
    torch.cuda.synchronize()
    tokens = tokens.long()
    emb = torch.cat([prefix, eng.embed_tokens(tokens[:, :T - 1])], dim=1)
    with torch.no_grad():
        ref_logits = hf(inputs_embeds=emb).logits.float()
    ours = eng.lm_forward(emb)
    assert rel(ours, ref_logits) <= TOL
    pred = ref_logits[:, P - 1:P - 1 + T]
    agree = pred.argmax(-1) == tokens
    scale = (pred.max() - pred.min()).item()
    for r, t in (~agree).nonzero().tolist():
        assert (pred[r, t].max() - pred[r, t, tokens[r, t]]).item() <= TOL * scale, (r, t)
    assert agree.float().mean().item() >= 0.95
    print("GPT-J-6B: greedy tokens identical to transformers fp32 at %.1f %% of %d positions" % (100 * agree.float().mean().item(), agree.numel()))
    eng.close()

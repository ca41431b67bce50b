"""Generates tests/golden/*.pt by RUNNING THE REFERENCE ITSELF (build container only).

    PYTHONDONTWRITEBYTECODE=1 python tools/make_golden.py

The reference (/root/reference) has no tests or golden vectors, so the fixtures are the outputs of its own,
unmodified modules -- layers.TransformerMapper, lms.GPT2 / lms.GPTJ (HF transformers), model.CLIPCaptionModel,
inference.generate_beam / generate_no_beam / top_k_top_p_filtering / repetition_penalty_apply,
evaluate_model.generate_no_beam, sampling.top_k_top_p_filtering_batch / repetition_penalty_apply -- on seeded
tiny models whose weights are rounded to bf16 once (the comparand both sides share).  The OpenAI `clip` package
is not installed; HF CLIPVisionModelWithProjection (same arithmetic) stands in and its
weights are exported under the OpenAI names.  The fixtures hold weights + inputs + outputs, so the tests that
consume them (tests/test_oracle_golden.py on CPU, tests/test_gpu_parity.py on the B200) never need the reference.
"""
import os
import types
import math
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_harness  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def bf16_round_(module):
    with torch.no_grad():
        for p in module.parameters():
            p.copy_(p.bfloat16().float())
    return module


def export_clip_vision(hf):
    """HF CLIPVisionModelWithProjection -> OpenAI `visual.*` names."""
    s = hf.state_dict()
    o = {}
    o["conv1.weight"] = s["vision_model.embeddings.patch_embedding.weight"]
    o["class_embedding"] = s["vision_model.embeddings.class_embedding"]
    o["positional_embedding"] = s["vision_model.embeddings.position_embedding.weight"]
    o["ln_pre.weight"] = s["vision_model.pre_layrnorm.weight"]
    o["ln_pre.bias"] = s["vision_model.pre_layrnorm.bias"]
    o["ln_post.weight"] = s["vision_model.post_layernorm.weight"]
    o["ln_post.bias"] = s["vision_model.post_layernorm.bias"]
    o["proj"] = s["visual_projection.weight"].t().contiguous()
    n = hf.config.num_hidden_layers
    for l in range(n):
        a = "vision_model.encoder.layers.%d." % l
        b = "transformer.resblocks.%d." % l
        o[b + "ln_1.weight"], o[b + "ln_1.bias"] = s[a + "layer_norm1.weight"], s[a + "layer_norm1.bias"]
        o[b + "ln_2.weight"], o[b + "ln_2.bias"] = s[a + "layer_norm2.weight"], s[a + "layer_norm2.bias"]
        o[b + "attn.in_proj_weight"] = torch.cat([s[a + "self_attn.%s_proj.weight" % n_] for n_ in "qkv"], 0)
        o[b + "attn.in_proj_bias"] = torch.cat([s[a + "self_attn.%s_proj.bias" % n_] for n_ in "qkv"], 0)
        o[b + "attn.out_proj.weight"], o[b + "attn.out_proj.bias"] = s[a + "self_attn.out_proj.weight"], s[a + "self_attn.out_proj.bias"]
        o[b + "mlp.c_fc.weight"], o[b + "mlp.c_fc.bias"] = s[a + "mlp.fc1.weight"], s[a + "mlp.fc1.bias"]
        o[b + "mlp.c_proj.weight"], o[b + "mlp.c_proj.bias"] = s[a + "mlp.fc2.weight"], s[a + "mlp.fc2.bias"]
    return {k: v.detach().clone() for k, v in o.items()}


class TokenizerStub:
    """The tokenizer surface the generate loops touch (lms/GPT2.py:22-48); ids in, ids out."""

    def __init__(self, stop_id, bos=None, special=()):
        self.stop_id = stop_id
        self.bos_token_id = bos
        self.all_special_ids = list(special)

    def encode_text(self, text, *a, **k):
        return [self.stop_id]

    def decode_tokens(self, tokens):
        return [int(t) for t in tokens]


def pack_sd(sd):
    return {k: v.detach().bfloat16() if v.is_floating_point() else v.detach().clone() for k, v in sd.items()}


def make_model_fixture(ref, arch, seed):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection, GPT2Config, GPTJConfig
    torch.manual_seed(seed)
    V, d, heads, P, CL, dim_clip = 503, 128, 2, 4, 4, 64
    vit_cfg = CLIPVisionConfig(hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                               image_size=64, patch_size=32, projection_dim=dim_clip, hidden_act="quick_gelu")
    vit = bf16_round_(CLIPVisionModelWithProjection(vit_cfg).eval())
    if arch == "gpt2":
        lm = ref.lms.GPT2(GPT2Config(vocab_size=V, n_positions=64, n_embd=d, n_layer=2, n_head=heads))
        # HF's 0.02 init makes a 2-layer tied-embedding model echo its input token; larger block weights make
        # the layers matter and a peakier embedding keeps top-1 margins above bf16 noise
        with torch.no_grad():
            lm.transformer.wte.weight.mul_(4.0)
            for n_, p_ in lm.transformer.h.named_parameters():
                if p_.dim() == 2:
                    p_.mul_(6.0)
    else:
        V = 520
        lm = ref.lms.GPTJ(GPTJConfig(vocab_size=V, n_positions=64, n_embd=d, n_layer=2, n_head=heads, rotary_dim=16))
        with torch.no_grad():
            lm.transformer.wte.weight.mul_(4.0)
            lm.lm_head.weight.mul_(8.0)
            lm.lm_head.bias.normal_(0, 0.1)
            for n_, p_ in lm.transformer.h.named_parameters():
                if p_.dim() == 2:
                    p_.mul_(6.0)
    lm = bf16_round_(lm.eval())
    stop_id = 13
    tok = TokenizerStub(stop_id, bos=V - 1, special=[V - 1])

    class VisualEncoder(torch.nn.Module):  # stands in for clip_model.visual / encode_image
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x):
            return self.m(pixel_values=x).image_embeds

    model = ref.model.CLIPCaptionModel(
        language_model=lm, tokenizer=tok, visual_encoder=VisualEncoder(vit), validator=None,
        train_visual_encoder=False, use_all_vit_features=False, prefix_size=dim_clip, prefix_length=P,
        clip_prefix_length=CL, num_attention_heads=8, num_layers=2, mlp_ratio=4.0, prefix_init_std=1.0,
        act_fn_name="relu", pos_embeddings=False)
    bf16_round_(model.clip_project)
    model.eval()

    N = 3
    images = torch.randn(N, 3, 64, 64)
    fx = {"arch": arch, "V": V, "d": d, "heads": heads, "P": P, "CL": CL, "dim_clip": dim_clip, "map_heads": 8,
          "rotary_dim": 16 if arch == "gptj" else 0, "vit_heads": 2, "vit_patch": 32, "vit_image": 64, "vit_width": 64,
          "vit_layers": 2, "images": images}
    with torch.no_grad():
        feat = model.visual_encoder(images).float()                       # encode_image (inference.py:311)
        prefix = model.clip_project(feat)                                 # inference.py:312
        tokens = torch.randint(0, V - 1, (N, 6))
        mask = torch.ones(N, 6, dtype=torch.bool)
        mask[1, 4:] = False
        mask[2, 2:] = False
        logits_tf = model(tokens, feat, mask).logits                      # model.py:132-149
        logits_prefix = lm.call(inputs_embeds=prefix).logits              # lms/GPT2.py:17-19
        fx.update(feat=feat, prefix=prefix, tokens=tokens, mask=mask, logits_tf=logits_tf, logits_prefix=logits_prefix)

        # choose the stop id so that stopping actually happens: the most frequent token of a free-running greedy decode
        tok.stop_id = -1
        free = [ref.inference.generate_beam(model, tok, prefix[i:i + 1], beam_size=1, entry_length=10)[0] for i in range(N)]
        flat = [t for seq in free for t in seq[3:]]
        tok.stop_id = max(set(flat), key=flat.count)
        fx["stop_id"] = tok.stop_id
        fx["greedy"] = [ref.inference.generate_beam(model, tok, prefix[i:i + 1], beam_size=1, entry_length=10)[0] for i in range(N)]
        fx["beam5"] = [ref.inference.generate_beam(model, tok, prefix[i:i + 1], beam_size=5, entry_length=10)[0] for i in range(N)]
        fx["beam3_T2"] = [ref.inference.generate_beam(model, tok, prefix[i:i + 1], beam_size=3, entry_length=8, temperature=2.0)[0] for i in range(N)]

        # nucleus sampling, both variants; the global RNG is seeded so the oracle can replay the Exp(1) draws
        fx["nobeam_seed"] = 1234 + seed
        import contextlib
        import io
        outs, seeds = [], []
        for i in range(N):
            # inference.py:287 crashes on a one-token caption (`tokens.squeeze()` is 0-d): such seeds are skipped
            s_i = fx["nobeam_seed"] + i
            while True:
                torch.manual_seed(s_i)
                try:
                    with contextlib.redirect_stdout(io.StringIO()):
                        outs.append(ref.inference.generate_no_beam(model, tok, prefix[i:i + 1], entry_length=8, repetition_penalty=1.2))
                    break
                except TypeError:
                    s_i += 1000
            seeds.append(s_i)
        fx["nobeam_inference"] = outs
        fx["nobeam_inference_seeds"] = seeds
        outs = []
        for i in range(N):
            torch.manual_seed(fx["nobeam_seed"] + 100 + i)
            outs.append(ref.evaluate_model.generate_no_beam(model, prefix[i:i + 1], top_p_values=[0.3, 0.9],
                                                            max_decode_length=8, repetition_penalty=1.2, max_stops=2))
        fx["nobeam_evaluate"] = outs
    fx["sd_lm"] = pack_sd(lm.state_dict())
    fx["sd_mapper"] = pack_sd(model.clip_project.state_dict())
    fx["sd_vit"] = pack_sd(export_clip_vision(vit))
    return fx


def make_sampler_fixture(ref, seed):
    torch.manual_seed(seed)
    B, V = 6, 1031
    logits = torch.randn(B, V) * 3.0
    logits[2, 5] = logits[2, 77] = logits[2].max() + 1.0            # exact ties at the top
    logits[4] = torch.round(logits[4])                               # many ties everywhere
    fx = {"logits": logits}
    f = ref.sampling.top_k_top_p_filtering_batch
    fx["topp_0.9"] = f(logits.clone(), top_k=0, top_p=0.9)
    fx["topp_0.1"] = f(logits.clone(), top_k=0, top_p=0.1)
    fx["topk_40"] = f(logits.clone(), top_k=40, top_p=0.0)
    fx["topk_0.05"] = f(logits.clone(), top_k=0.05, top_p=0.0)
    fx["topk_40_topp_0.5"] = f(logits.clone(), top_k=40, top_p=0.5)
    tp = torch.tensor([0.1, 0.3, 0.5, 0.7, 0.9, 0.95])
    tk = torch.tensor([0, 1, 5, 50, 0, 2000])
    fx["top_p_rows"], fx["top_k_rows"] = tp, tk
    fx["rows"] = f(logits.clone(), top_k=tk.clone(), top_p=tp.clone())
    hist = torch.randint(0, V, (B, 7))
    hist[0, 3] = hist[0, 1]                                           # duplicate ids in the history
    fx["history"] = hist
    fx["rep_1.2"] = ref.sampling.repetition_penalty_apply(logits.clone(), hist, 1.2)
    fx["rep_1.2_inference_row0"] = ref.inference.repetition_penalty_apply(logits[0].clone(), hist[0], 1.2)
    fx["topp1d_0.8"] = torch.stack([ref.inference.top_k_top_p_filtering(logits[i].clone(), top_p=0.8, top_k=0) for i in range(B)])
    fx["topk1d_7"] = torch.stack([ref.evaluate_model.top_k_top_p_filtering(logits[i].clone(), top_p=0.0, top_k=7) for i in range(B)])
    # torch.multinomial draws (n = 1 and n = 2 without replacement) with a seeded generator
    probs = torch.softmax(fx["topp_0.9"], -1)
    g = torch.Generator().manual_seed(99)
    fx["multinomial_seed"] = 99
    fx["multinomial_1"] = torch.multinomial(probs, 1, generator=g)
    fx["multinomial_2"] = torch.multinomial(probs, 2, replacement=False, generator=g)
    return fx


def make_typical_fixture(ref, seed):
    """sampling.typical_filtering (sampling.py:72-102) on raw, top-p filtered and per-row budgets."""
    torch.manual_seed(seed)
    B, V = 6, 1031
    logits = torch.randn(B, V) * 3.0
    logits[3] = logits[3] * 0.2                                       # a flat row (high entropy)
    logits[4] = torch.round(logits[4])                                # many ties
    fx = {"logits": logits}
    f = ref.sampling.typical_filtering
    for tp in (0.2, 0.5, 0.9):
        fx["typ_%s" % tp] = f(logits.clone(), typ_p=tp)
    rows = torch.tensor([0.1, 0.25, 0.5, 0.75, 0.9, 0.95])
    fx["typ_rows_p"] = rows
    fx["typ_rows"] = f(logits.clone(), typ_p=rows.clone())
    nucleus = ref.sampling.top_k_top_p_filtering_batch(logits.clone(), top_k=0, top_p=0.9)
    fx["topp_0.9_typ_0.5"] = f(nucleus.clone(), typ_p=0.5)            # the order of sampling.generate (:205-206)
    return fx


def make_allfeatures_fixture(ref, seed):
    """The fork's default configuration (train.py:73 use_all_vit_features=True): every projected ViT token
    (the patched VisionTransformer.forward of inference.py:421-444; HF CLIPVisionModelWithProjection stands in for the
    un-installed OpenAI clip: tokens = last_hidden_state @ visual_projection^T, no post_layernorm) ->
    layers.TransformerMapperAllFeatures (layers/Transformer.py:164-203) with and without pos_embeddings ->
    CLIPCaptionModel.forward / generate_beam of the UNMODIFIED reference."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection, GPT2Config
    torch.manual_seed(seed)
    V, d, heads, P, dim_clip = 503, 128, 2, 4, 64
    vit_cfg = CLIPVisionConfig(hidden_size=64, intermediate_size=256, num_hidden_layers=2, num_attention_heads=2,
                               image_size=64, patch_size=32, projection_dim=dim_clip, hidden_act="quick_gelu")
    vit = bf16_round_(CLIPVisionModelWithProjection(vit_cfg).eval())
    T = (64 // 32) ** 2 + 1
    lm = ref.lms.GPT2(GPT2Config(vocab_size=V, n_positions=64, n_embd=d, n_layer=2, n_head=heads))
    with torch.no_grad():
        lm.transformer.wte.weight.mul_(4.0)
        for n_, p_ in lm.transformer.h.named_parameters():
            if p_.dim() == 2:
                p_.mul_(6.0)
    lm = bf16_round_(lm.eval())
    tok = TokenizerStub(13, bos=V - 1, special=[V - 1])

    class VisualTokens(torch.nn.Module):   # what clip_model.visual.forward returns after the fork's patch
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x):
            h = self.m.vision_model(pixel_values=x).last_hidden_state
            return self.m.visual_projection(h)

    fx = {"V": V, "d": d, "heads": heads, "P": P, "T": T, "dim_clip": dim_clip, "map_heads": 8, "vit_heads": 2, "vit_patch": 32,
          "vit_image": 64, "vit_width": 64, "vit_layers": 2, "stop_id": 13}
    images = torch.randn(3, 3, 64, 64)
    fx["images"] = images
    fx["sd_lm"] = pack_sd(lm.state_dict())
    fx["sd_vit"] = pack_sd(export_clip_vision(vit))
    for pos in (True, False):
        model = ref.model.CLIPCaptionModel(
            language_model=lm, tokenizer=tok, visual_encoder=VisualTokens(vit), validator=None, train_visual_encoder=False,
            use_all_vit_features=True, prefix_size=dim_clip, prefix_length=P, clip_prefix_length=T, num_attention_heads=8,
            num_layers=2, mlp_ratio=4.0, prefix_init_std=1.0, act_fn_name="relu", pos_embeddings=pos)
        bf16_round_(model.clip_project)
        model.eval()
        key = "pos" if pos else "nopos"
        with torch.no_grad():
            tokens_feat = model.visual_encoder(images).float()
            prefix = model.clip_project(tokens_feat)
            fx["vit_tokens"] = tokens_feat
            fx["prefix_" + key] = prefix
            cap = torch.randint(0, V, (3, 6))
            fx["cap_tokens"] = cap if "cap_tokens" not in fx else fx["cap_tokens"]
            fx["forward_logits_" + key] = model(fx["cap_tokens"], tokens_feat, torch.ones(3, 6, dtype=torch.bool)).logits
            fx["beam5_" + key] = [ref.inference.generate_beam(model, tok, prefix[i:i + 1], beam_size=5, entry_length=10)[0]
                                  for i in range(3)]
        fx["sd_mapper_" + key] = pack_sd(model.clip_project.state_dict())
    return fx


def make_mapper_acts_fixture(ref, seed):
    """layers.TransformerMapper (layers/Transformer.py:133-161) of the UNMODIFIED reference with every activation
    parse_act_fn accepts besides relu (:117-130): elu, gelu, selu and geglu (fc1 twice as wide, :74, :112-114)."""
    fx = {"dim_clip": 64, "d": 128, "P": 4, "CL": 4, "map_heads": 8}
    torch.manual_seed(seed)
    fx["feat"] = torch.randn(3, 64)
    for act in ("elu", "gelu", "selu", "geglu"):
        m = ref.layers.TransformerMapper(dim_clip=64, dim_embedding=128, prefix_length=4, clip_length=4, num_heads=8, num_layers=2,
                                         mlp_ratio=4.0, prefix_init_std=1.0, act_fn_name=act)
        bf16_round_(m).eval()
        with torch.no_grad():
            fx["prefix_" + act] = m(fx["feat"])
        fx["sd_" + act] = pack_sd(m.state_dict())
    return fx


def clip_tokenize_ids(texts, ctx, sot, eot):
    """Stand-in for clip.tokenize(texts, truncate=True) on `texts` that are lists of token ids (TokenizerStub.decode_tokens):
    [sot] + ids + [eot], truncated to `ctx` with the end-of-text id kept last, zero-padded -- clip.tokenize's own layout."""
    out = torch.zeros(len(texts), ctx, dtype=torch.int64)
    for i, t in enumerate(texts):
        ids = [sot] + [int(x) for x in (t if isinstance(t, (list, tuple)) else [t])] + [eot]
        if len(ids) > ctx:
            ids = ids[:ctx]
            ids[-1] = eot
        out[i, :len(ids)] = torch.tensor(ids)
    return out


def make_clip_guided_fixture(ref, seed):
    """evaluate_model.generate_clip_guided (evaluate_model.py:182-312) of the UNMODIFIED reference on the tiny GPT-2 fixture's
    language model, prefixes and image embeddings, with a tiny CLIP text tower (HF CLIPTextModelWithProjection, arg-max
    pooling = OpenAI's rule) as `clip_model.encode_text` and `clip.tokenize` replaced by `clip_tokenize_ids`."""
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection, GPT2Config
    base = torch.load(os.path.join(OUT, "tiny_gpt2.pt"), weights_only=False)
    torch.manual_seed(seed)
    V, d = base["V"], base["d"]
    lm = ref.lms.GPT2(GPT2Config(vocab_size=V, n_positions=64, n_embd=d, n_layer=2, n_head=base["heads"]))
    lm.load_state_dict({k: v.float() for k, v in base["sd_lm"].items()}, strict=False)
    lm.tie_weights()
    lm.eval()
    Vc, ctx, w, heads, out = V + 2, 24, 64, 2, base["dim_clip"]
    cfg = CLIPTextConfig(vocab_size=Vc, hidden_size=w, intermediate_size=4 * w, num_hidden_layers=2, num_attention_heads=heads,
                         max_position_embeddings=ctx, projection_dim=out, hidden_act="quick_gelu", eos_token_id=2,
                         bos_token_id=0, pad_token_id=1)
    m = bf16_round_(CLIPTextModelWithProjection(cfg).eval())
    with torch.no_grad():
        for n_, p_ in m.named_parameters():
            if p_.dim() == 2 and "embedding" not in n_:
                p_.mul_(4.0)
    m = bf16_round_(m)
    sot, eot = Vc - 2, Vc - 1

    class ClipModel:
        def encode_text(self, tokens):
            with torch.no_grad():
                return m(input_ids=tokens).text_embeds

    tok = TokenizerStub(base["stop_id"], bos=V - 1, special=[V - 1])
    model = types.SimpleNamespace(tokenizer=tok, language_model=lm)
    ref.evaluate_model.clip.tokenize = lambda texts, truncate=True: clip_tokenize_ids(texts, ctx, sot, eot)
    fx = {"Vc": Vc, "ctx": ctx, "w": w, "heads": heads, "out": out, "sot": sot, "eot": eot, "sd_text": pack_sd(export_clip_text(m)),
          "runs": []}
    for kw in (dict(max_decode_length=8, look_ahead=2, branching_factor=3, repetition_penalty=1.2),
               dict(max_decode_length=6, look_ahead=1, branching_factor=2, repetition_penalty=1.0)):
        outs = [ref.evaluate_model.generate_clip_guided("cpu", base["feat"][i:i + 1], model, ClipModel(), base["prefix"][i:i + 1], **kw)
                for i in range(base["prefix"].shape[0])]
        fx["runs"].append({"kw": kw, "captions": outs})
    return fx


def export_clip_text(hf):
    """HF CLIPTextModelWithProjection -> OpenAI CLIP text-tower names (token_embedding, positional_embedding,
    transformer.resblocks.N.*, ln_final, text_projection)."""
    s = hf.state_dict()
    o = {"token_embedding.weight": s["text_model.embeddings.token_embedding.weight"],
         "positional_embedding": s["text_model.embeddings.position_embedding.weight"],
         "ln_final.weight": s["text_model.final_layer_norm.weight"], "ln_final.bias": s["text_model.final_layer_norm.bias"],
         "text_projection": s["text_projection.weight"].t().contiguous()}
    for l in range(hf.config.num_hidden_layers):
        a = "text_model.encoder.layers.%d." % l
        b = "transformer.resblocks.%d." % l
        o[b + "ln_1.weight"], o[b + "ln_1.bias"] = s[a + "layer_norm1.weight"], s[a + "layer_norm1.bias"]
        o[b + "ln_2.weight"], o[b + "ln_2.bias"] = s[a + "layer_norm2.weight"], s[a + "layer_norm2.bias"]
        o[b + "attn.in_proj_weight"] = torch.cat([s[a + "self_attn.%s_proj.weight" % n_] for n_ in "qkv"], 0)
        o[b + "attn.in_proj_bias"] = torch.cat([s[a + "self_attn.%s_proj.bias" % n_] for n_ in "qkv"], 0)
        o[b + "attn.out_proj.weight"], o[b + "attn.out_proj.bias"] = s[a + "self_attn.out_proj.weight"], s[a + "self_attn.out_proj.bias"]
        o[b + "mlp.c_fc.weight"], o[b + "mlp.c_fc.bias"] = s[a + "mlp.fc1.weight"], s[a + "mlp.fc1.bias"]
        o[b + "mlp.c_proj.weight"], o[b + "mlp.c_proj.bias"] = s[a + "mlp.fc2.weight"], s[a + "mlp.fc2.bias"]
    return {k: v.detach().clone() for k, v in o.items()}


def make_clip_text_fixture(ref, seed):
    """encode_text + cos_sim (sampling.py:14-37).  OpenAI `clip` is not installed: HF CLIPTextModelWithProjection with
    eos_token_id = 2 (features read at the arg-max token id, exactly OpenAI's rule) stands in for its text tower, the
    similarity is the reference's own sampling.cos_sim."""
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection
    torch.manual_seed(seed)
    V, ctx, w, heads, out = 131, 16, 64, 2, 32
    cfg = CLIPTextConfig(vocab_size=V, hidden_size=w, intermediate_size=4 * w, num_hidden_layers=2, num_attention_heads=heads,
                         max_position_embeddings=ctx, projection_dim=out, hidden_act="quick_gelu", eos_token_id=2,
                         bos_token_id=0, pad_token_id=1)
    m = bf16_round_(CLIPTextModelWithProjection(cfg).eval())
    with torch.no_grad():
        for n_, p_ in m.named_parameters():
            if p_.dim() == 2 and "embedding" not in n_:
                p_.mul_(4.0)
    m = bf16_round_(m)
    B = 5
    tokens = torch.zeros(B, ctx, dtype=torch.int64)
    for b in range(B):
        n = 3 + 2 * b
        tokens[b, 0] = V - 2                                  # start-of-text
        tokens[b, 1:n] = torch.randint(0, V - 2, (n - 1,))
        tokens[b, n] = V - 1                                  # end-of-text: the largest id
    with torch.no_grad():
        feats = m(input_ids=tokens).text_embeds.float()
    image_features = torch.randn(1, out)
    fx = {"V": V, "ctx": ctx, "w": w, "heads": heads, "out": out, "tokens": tokens, "text_features": feats,
          "image_features": image_features, "sims": ref.sampling.cos_sim(feats, image_features),
          "sd_text": pack_sd(export_clip_text(m))}
    return fx


class FakeDecoder(torch.nn.Module):
    """A deterministic stand-in for the BLIP text decoder `sampling.generate` drives (no BLIP checkout offline): logits of
    the last position depend on the last two tokens and on the row's encoder state; no randomness."""

    def __init__(self, E, W):
        super().__init__()
        self.E, self.W = E, W
        self.config = types.SimpleNamespace(output_attentions=False, output_hidden_states=False)

    def forward(self, input_ids=None, encoder_hidden_states=None, encoder_attention_mask=None, return_dict=True, **kw):
        e = self.E[input_ids.cpu()]
        prev = torch.cat((e[:, :1] * 0, e[:, :-1]), dim=1)
        h = e + 0.5 * prev + encoder_hidden_states.cpu().mean(dim=1, keepdim=True)
        logits = torch.tanh(h) @ self.W
        logits[:, :, 3] += 1.2 * torch.arange(logits.shape[1], dtype=logits.dtype).view(1, -1)   # EOS (id 3) grows likelier
        return {"logits": logits}


def make_generate_fixture(ref, seed):
    """sampling.generate (sampling.py:165-268) of the unmodified reference with the fake decoder: per-row budgets, EOS
    masking before min_length, forced completion on a high EOS probability, alternate-sample fallback."""
    torch.manual_seed(seed)
    V, hdim, B, eos = 211, 24, 12, 3
    E, W = torch.randn(V, hdim), torch.randn(hdim, V) * 1.5
    dec = FakeDecoder(E, W)
    enc = torch.randn(B, 5, hdim) * 0.3
    enc_mask = torch.ones(B, 5, dtype=torch.long)
    prompt = torch.randint(4, V, (B, 3))
    fx = {"V": V, "hdim": hdim, "eos": eos, "E": E, "W": W, "enc": enc, "enc_mask": enc_mask, "prompt": prompt, "runs": []}
    cases = [
        dict(top_p=torch.linspace(0.3, 0.95, B), top_k=0, typ_p=0.0, min_length=torch.full((B,), 2), max_length=torch.arange(4, 4 + B),
             repetition_penalty=1.3, min_alternate_prob=0, force_eos_log_prob=math.log(0.9)),
        dict(top_p=torch.full((B,), 0.9), top_k=torch.tensor([0, 5, 50, 2, 0, 3, 0, 0, 20, 0, 7, 0]), typ_p=torch.linspace(0.2, 0.9, B),
             min_length=torch.arange(B) % 4, max_length=torch.full((B,), 9), repetition_penalty=None, min_alternate_prob=0.05,
             force_eos_log_prob=math.log(0.5)),
        dict(top_p=torch.full((B,), 0.0), top_k=20, typ_p=0.6, min_length=torch.zeros(B, dtype=torch.long),
             max_length=torch.full((B,), 6), repetition_penalty=1.1, min_alternate_prob=0.2, force_eos_log_prob=1.0),
    ]
    for ci, kw in enumerate(cases):
        torch.manual_seed(1000 + ci)
        out = ref.sampling.generate(dec, prompt.clone(), enc.clone(), enc_mask.clone(), eos_token_id=eos, **{k: (v.clone() if torch.is_tensor(v) else v) for k, v in kw.items()})
        fx["runs"].append({"seed": 1000 + ci, "kwargs": kw, "results": [[t.clone() if torch.is_tensor(t) else t for t in grp] for grp in out]})
    return fx


def make_loss_fixture():
    """The teacher-forced loss of model.py:204-211 (training_step) == evaluate_model.py:505-514 (validation), evaluated
    with the reference's own expression on the logits the unmodified reference model produced for the committed model
    fixtures (tiny_gpt2.pt / tiny_gptj.pt: `logits_tf = model(tokens, feat, mask).logits`)."""
    import torch.nn.functional as nnf
    out = {}
    for arch in ("gpt2", "gptj"):
        fx = torch.load(os.path.join(OUT, "tiny_%s.pt" % arch), weights_only=False)
        tokens, mask, P = fx["tokens"].clone(), fx["mask"], fx["P"]
        tokens[~mask] = 0                                                   # model.py:205
        logits = fx["logits_tf"][:, P - 1: -1]                              # model.py:209
        loss = nnf.cross_entropy(logits.reshape(-1, logits.shape[-1]), tokens.flatten(), ignore_index=0)  # model.py:210
        rows = nnf.cross_entropy(logits.reshape(-1, logits.shape[-1]), tokens.flatten(), ignore_index=0, reduction="none")
        out[arch] = {"loss": loss, "row_loss": rows, "tokens": tokens}
    return out


def main():
    ref = ref_harness.load_reference()
    os.makedirs(OUT, exist_ok=True)
    if "--clip-guided" in sys.argv:
        fx = make_clip_guided_fixture(ref, 27)
        path = os.path.join(OUT, "tiny_clip_guided.pt")
        torch.save(fx, path)
        print(path, os.path.getsize(path) // 1024, "KiB", [r["captions"] for r in fx["runs"]])
        return
    if "--acts" in sys.argv:
        fx = make_mapper_acts_fixture(ref, 26)
        path = os.path.join(OUT, "tiny_mapper_acts.pt")
        torch.save(fx, path)
        print(path, os.path.getsize(path) // 1024, "KiB", {k: tuple(v.shape) for k, v in fx.items() if k.startswith("prefix_")})
        return
    if "--loss" in sys.argv:
        fx = make_loss_fixture()
        path = os.path.join(OUT, "tiny_loss.pt")
        torch.save(fx, path)
        print(path, os.path.getsize(path), "B", {k: float(v["loss"]) for k, v in fx.items()})
        return
    for arch, seed in (() if "--only-typical" in sys.argv else (("gpt2", 11), ("gptj", 12))):
        fx = make_model_fixture(ref, arch, seed)
        path = os.path.join(OUT, "tiny_%s.pt" % arch)
        torch.save(fx, path)
        print(path, os.path.getsize(path) // 1024, "KiB", "stop_id", fx["stop_id"])
        for k in ("greedy", "beam5", "beam3_T2", "nobeam_inference", "nobeam_evaluate"):
            print("  ", k, fx[k])
    if "--only-typical" not in sys.argv:
        fx = make_sampler_fixture(ref, 21)
        path = os.path.join(OUT, "sampler.pt")
        torch.save(fx, path)
        print(path, os.path.getsize(path) // 1024, "KiB")
    if ("--only-typical" not in sys.argv) or "--allfeatures" in sys.argv:
        fx = make_allfeatures_fixture(ref, 23)
        path = os.path.join(OUT, "tiny_allfeatures.pt")
        torch.save(fx, path)
        print(path, os.path.getsize(path) // 1024, "KiB", fx["beam5_pos"], fx["beam5_nopos"])
    if ("--only-typical" not in sys.argv) or "--cliptext" in sys.argv:
        fx = make_clip_text_fixture(ref, 24)
        path = os.path.join(OUT, "tiny_clip_text.pt")
        torch.save(fx, path)
        print(path, os.path.getsize(path) // 1024, "KiB", fx["sims"].reshape(-1).tolist())
    if "--only-typical" not in sys.argv or "--generate" in sys.argv:
        fx = make_generate_fixture(ref, 25)
        path = os.path.join(OUT, "sampling_generate.pt")
        torch.save(fx, path)
        print(path, os.path.getsize(path) // 1024, "KiB", [[tuple(g[0].shape) for g in r["results"]] for r in fx["runs"]])
    fx = make_typical_fixture(ref, 22)
    path = os.path.join(OUT, "typical.pt")
    torch.save(fx, path)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()

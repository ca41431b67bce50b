"""Seeded random-init weights and synthetic images of the shapes BASELINE.json names (there are no checkpoints
or datasets offline).  Weights follow the initialisers of the modules the reference instantiates: HF GPT-2 /
GPT-J (`initializer_range` 0.02, LayerNorm 1/0), torch `nn.Linear` default init for the mapper with
`prefix_const ~ N(0, prefix_init_std)` (layers/Transformer.py:151), HF CLIP vision init for the ViT stand-in.
All tensors are rounded to bf16 once so that the CUDA path and the fp32 oracle share identical weights.
"""
import math
from typing import Dict

import torch

from .engine import EngineConfig


def _rn(g, shape, std, device):
    return (torch.randn(shape, generator=g, device=device) * std).bfloat16().float()


def _uniform(g, shape, bound, device):
    return ((torch.rand(shape, generator=g, device=device) * 2 - 1) * bound).bfloat16().float()


def _linear(sd, name, out_f, in_f, g, device, bias=True):
    b = 1.0 / math.sqrt(in_f)
    sd[name + ".weight"] = _uniform(g, (out_f, in_f), b, device)
    if bias:
        sd[name + ".bias"] = _uniform(g, (out_f,), b, device)


def _ln(sd, name, d, device):
    sd[name + ".weight"] = torch.ones(d, device=device)
    sd[name + ".bias"] = torch.zeros(d, device=device)


def lm_state_dict(cfg: EngineConfig, seed: int = 1234, device="cpu", std: float = 0.02, wte_std: float = 0.02) -> Dict[str, torch.Tensor]:
    g = torch.Generator(device=device).manual_seed(seed)
    d, V, L = cfg.lm_d, cfg.lm_vocab, cfg.lm_layers
    sd = {"transformer.wte.weight": _rn(g, (V, d), wte_std, device)}
    if cfg.lm_arch == "gpt2":
        sd["transformer.wpe.weight"] = _rn(g, (cfg.lm_n_pos, d), std, device)
    for l in range(L):
        p = "transformer.h.%d." % l
        _ln(sd, p + "ln_1", d, device)
        if cfg.lm_arch == "gpt2":
            _ln(sd, p + "ln_2", d, device)
            sd[p + "attn.c_attn.weight"] = _rn(g, (d, 3 * d), std, device)
            sd[p + "attn.c_attn.bias"] = torch.zeros(3 * d, device=device)
            sd[p + "attn.c_proj.weight"] = _rn(g, (d, d), std / math.sqrt(2 * L), device)
            sd[p + "attn.c_proj.bias"] = torch.zeros(d, device=device)
            sd[p + "mlp.c_fc.weight"] = _rn(g, (d, 4 * d), std, device)
            sd[p + "mlp.c_fc.bias"] = torch.zeros(4 * d, device=device)
            sd[p + "mlp.c_proj.weight"] = _rn(g, (4 * d, d), std / math.sqrt(2 * L), device)
            sd[p + "mlp.c_proj.bias"] = torch.zeros(d, device=device)
        else:
            for n in ("q_proj", "k_proj", "v_proj", "out_proj"):
                sd[p + "attn.%s.weight" % n] = _rn(g, (d, d), std, device)
            sd[p + "mlp.fc_in.weight"] = _rn(g, (4 * d, d), std, device)
            sd[p + "mlp.fc_in.bias"] = torch.zeros(4 * d, device=device)
            sd[p + "mlp.fc_out.weight"] = _rn(g, (d, 4 * d), std, device)
            sd[p + "mlp.fc_out.bias"] = torch.zeros(d, device=device)
    _ln(sd, "transformer.ln_f", d, device)
    if cfg.lm_arch == "gptj":
        sd["lm_head.weight"] = _rn(g, (V, d), wte_std, device)
        sd["lm_head.bias"] = torch.zeros(V, device=device)
    return sd


def mapper_state_dict(cfg: EngineConfig, seed: int = 1235, device="cpu", prefix_init_std: float = 1.0):
    g = torch.Generator(device=device).manual_seed(seed)
    d = cfg.lm_d
    sd = {}
    if cfg.map_kind == "mlp":
        hidden = cfg.map_hidden or d * cfg.map_prefix_len // 2
        _linear(sd, "model.0", hidden, cfg.map_dim_clip, g, device)
        _linear(sd, "model.2", d * cfg.map_prefix_len, hidden, g, device)
        return sd
    hidden = cfg.map_hidden or int(d * cfg.map_mlp_ratio)
    if cfg.map_kind == "transformer_all":     # layers/Transformer.py:178-185
        _linear(sd, "linear", d, cfg.map_dim_clip, g, device)
        sd["pos_embeddings"] = _rn(g, (cfg.map_clip_len, d), 1.0, device)
    else:
        _linear(sd, "linear", cfg.map_clip_len * d, cfg.map_dim_clip, g, device)
    sd["prefix_const"] = _rn(g, (cfg.map_prefix_len, d), prefix_init_std, device)
    for l in range(cfg.map_layers):
        p = "transformer.layers.%d." % l
        _ln(sd, p + "norm1", d, device)
        _ln(sd, p + "norm2", d, device)
        _linear(sd, p + "attn.to_queries", d, d, g, device, bias=False)
        _linear(sd, p + "attn.to_keys_values", 2 * d, d, g, device, bias=False)
        _linear(sd, p + "attn.project", d, d, g, device)
        _linear(sd, p + "mlp.fc1", hidden, d, g, device)
        _linear(sd, p + "mlp.fc2", d, hidden, g, device)
    return sd


def vit_state_dict(cfg: EngineConfig, seed: int = 1236, device="cpu"):
    """OpenAI CLIP `visual.*` names; scales follow CLIP.initialize_parameters / HF CLIPVisionModel init."""
    g = torch.Generator(device=device).manual_seed(seed)
    w, L = cfg.vit_width, cfg.vit_layers
    S = (cfg.vit_image // cfg.vit_patch) ** 2 + 1
    sd = {
        "conv1.weight": _rn(g, (w, 3, cfg.vit_patch, cfg.vit_patch), 0.02, device),
        "class_embedding": _rn(g, (w,), w ** -0.5, device),
        "positional_embedding": _rn(g, (S, w), w ** -0.5, device),
        "proj": _rn(g, (w, cfg.vit_out), w ** -0.5, device),
    }
    _ln(sd, "ln_pre", w, device)
    _ln(sd, "ln_post", w, device)
    attn_std, proj_std, fc_std = w ** -0.5, (w ** -0.5) * ((2 * L) ** -0.5), (2 * w) ** -0.5
    for l in range(L):
        p = "transformer.resblocks.%d." % l
        _ln(sd, p + "ln_1", w, device)
        _ln(sd, p + "ln_2", w, device)
        sd[p + "attn.in_proj_weight"] = _rn(g, (3 * w, w), attn_std, device)
        sd[p + "attn.in_proj_bias"] = torch.zeros(3 * w, device=device)
        sd[p + "attn.out_proj.weight"] = _rn(g, (w, w), proj_std, device)
        sd[p + "attn.out_proj.bias"] = torch.zeros(w, device=device)
        sd[p + "mlp.c_fc.weight"] = _rn(g, (4 * w, w), fc_std, device)
        sd[p + "mlp.c_fc.bias"] = torch.zeros(4 * w, device=device)
        sd[p + "mlp.c_proj.weight"] = _rn(g, (w, 4 * w), proj_std, device)
        sd[p + "mlp.c_proj.bias"] = torch.zeros(w, device=device)
    return sd


def text_state_dict(cfg: EngineConfig, seed: int = 1237, device="cpu"):
    """OpenAI CLIP text-tower names; scales follow CLIP.initialize_parameters."""
    g = torch.Generator(device=device).manual_seed(seed)
    w, L = cfg.text_width, cfg.text_layers
    sd = {"token_embedding.weight": _rn(g, (cfg.text_vocab, w), 0.02, device),
          "positional_embedding": _rn(g, (cfg.text_ctx, w), 0.01, device),
          "text_projection": _rn(g, (w, cfg.text_out), w ** -0.5, device)}
    _ln(sd, "ln_final", w, device)
    attn_std, proj_std, fc_std = w ** -0.5, (w ** -0.5) * ((2 * L) ** -0.5), (2 * w) ** -0.5
    for l in range(L):
        p = "transformer.resblocks.%d." % l
        _ln(sd, p + "ln_1", w, device)
        _ln(sd, p + "ln_2", w, device)
        sd[p + "attn.in_proj_weight"] = _rn(g, (3 * w, w), attn_std, device)
        sd[p + "attn.in_proj_bias"] = torch.zeros(3 * w, device=device)
        sd[p + "attn.out_proj.weight"] = _rn(g, (w, w), proj_std, device)
        sd[p + "attn.out_proj.bias"] = torch.zeros(w, device=device)
        sd[p + "mlp.c_fc.weight"] = _rn(g, (4 * w, w), fc_std, device)
        sd[p + "mlp.c_fc.bias"] = torch.zeros(4 * w, device=device)
        sd[p + "mlp.c_proj.weight"] = _rn(g, (w, 4 * w), proj_std, device)
        sd[p + "mlp.c_proj.bias"] = torch.zeros(w, device=device)
    return sd


def synthetic_images(n: int, cfg: EngineConfig, seed: int = 0, device="cpu", dtype=torch.float32):
    """randn images in the already-normalised space of CLIP's preprocessing."""
    g = torch.Generator(device=device).manual_seed(seed)
    return torch.randn(n, 3, cfg.vit_image, cfg.vit_image, generator=g, device=device).to(dtype)


def load_synthetic(engine, seed: int = 1234, device=None, **lm_kwargs):
    """Fill an Engine with seeded random-init weights; returns the three state dicts (on `device`)."""
    device = device or engine.device
    cfg = engine.cfg
    lm = lm_state_dict(cfg, seed, device, **lm_kwargs)
    engine.load_state_dict(lm, prefix="language_model.")
    sds = {"lm": lm}
    if cfg.map_kind != "none":
        sds["mapper"] = mapper_state_dict(cfg, seed + 1, device)
        engine.load_state_dict(sds["mapper"], prefix="clip_project.")
    if cfg.vit:
        sds["vit"] = vit_state_dict(cfg, seed + 2, device)
        engine.load_state_dict(sds["vit"], prefix="visual.")
    if getattr(cfg, "text", False):
        sds["text"] = text_state_dict(cfg, seed + 3, device)
        engine.load_state_dict(sds["text"], prefix="clip_text.")
    engine.check_weights()
    return sds

"""CPU oracle for the caption-generation hot path -- TEST INFRASTRUCTURE ONLY.

A plain PyTorch fp32 restatement of the reference algorithm (andreaskoepf/CLIP-Image-Captioning, paths relative
to the reference tree) and of the third-party arithmetic it calls (HF transformers GPT-2 / GPT-J, OpenAI CLIP
ViT), operating on plain `state_dict`s.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product path (clip-image-captioning_b200/) never does.

Pinning: the reference ships no tests or golden vectors.  The oracle is pinned against
the reference code itself: tools/make_golden.py imports the unmodified reference modules (layers/, lms/,
model.py, inference.py, evaluate_model.py, sampling.py) plus HF transformers 5.5.0 (GPT2LMHeadModel,
GPTJForCausalLM, CLIPVisionModelWithProjection as the OpenAI-ViT stand-in) in the build container, runs them on
seeded tiny models and stores inputs / weights / outputs under tests/golden/; tests/test_oracle_golden.py checks
every function below against those fixtures.  The upstream-ClipCap MLP mapper (absent from this fork) has no
reference code to run: `mlp_mapper_forward` is "parity unpinned".
"""
import math
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


# ------------------------------------------------------------------------------------------------ activations
def quick_gelu(x):  # OpenAI clip/model.py QuickGELU
    return x * torch.sigmoid(1.702 * x)


def gelu_new(x):  # HF activations.NewGELUActivation (GPT-2 / GPT-J "gelu_new")
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * torch.pow(x, 3.0))))


def geglu(x):
    """layers/Transformer.py:112-114 (fc1 is then 2 x hidden wide, :74)."""
    x, gate = x.chunk(2, dim=-1)
    return x * F.gelu(gate)


ACTS = {"relu": F.relu, "elu": F.elu, "gelu": F.gelu, "selu": F.selu, "geglu": geglu}  # layers/Transformer.py:117-130


# ------------------------------------------------------------------------------------------------ ViT
def vit_forward(sd: SD, images: torch.Tensor, heads: int, patch: int, all_tokens: bool = False) -> torch.Tensor:
    """OpenAI CLIP VisionTransformer.forward (mirrored at inference.py:422-442).
    `sd` uses the OpenAI names (conv1.weight, class_embedding, positional_embedding, ln_pre.*,
    transformer.resblocks.N.{ln_1,attn.in_proj_*,attn.out_proj,ln_2,mlp.c_fc,mlp.c_proj}, ln_post.*, proj)."""
    x = F.conv2d(images.float(), sd["conv1.weight"], stride=patch)            # [B, w, g, g]
    B, w = x.shape[0], x.shape[1]
    x = x.reshape(B, w, -1).permute(0, 2, 1)                                   # [B, g*g, w]
    cls = sd["class_embedding"].to(x.dtype) + torch.zeros(B, 1, w, dtype=x.dtype, device=x.device)
    x = torch.cat([cls, x], dim=1) + sd["positional_embedding"]
    x = F.layer_norm(x, (w,), sd["ln_pre.weight"], sd["ln_pre.bias"], 1e-5)
    hd = w // heads
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.resblocks."))
    for l in range(n_layers):
        p = "transformer.resblocks.%d." % l
        y = F.layer_norm(x, (w,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        qkv = F.linear(y, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
        q, k, v = qkv.view(B, -1, 3, heads, hd).unbind(2)
        att = torch.einsum("bnhd,bmhd->bhnm", q, k) * hd ** -0.5
        att = att.softmax(-1)
        y = torch.einsum("bhnm,bmhd->bnhd", att, v).reshape(B, -1, w)
        x = x + F.linear(y, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        y = F.layer_norm(x, (w,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        y = quick_gelu(F.linear(y, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]))
        x = x + F.linear(y, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
    if all_tokens:  # vit_forward_patch, inference.py:421-444: no ln_post, project every token
        return x @ sd["proj"]
    x = F.layer_norm(x[:, 0, :], (w,), sd["ln_post.weight"], sd["ln_post.bias"], 1e-5)
    return x @ sd["proj"]


# ------------------------------------------------------------------------------------------------ mapper
def mapper_attention(sd: SD, p: str, x: torch.Tensor, heads: int) -> torch.Tensor:
    """MultiHeadAttention.forward with y=None, mask=None (layers/MultiHeadAttention.py:17-43)."""
    b, n, c = x.shape
    q = F.linear(x, sd[p + "to_queries.weight"], sd.get(p + "to_queries.bias")).reshape(b, n, heads, c // heads)
    kv = F.linear(x, sd[p + "to_keys_values.weight"], sd.get(p + "to_keys_values.bias")).reshape(b, n, 2, heads, c // heads)
    k, v = kv[:, :, 0], kv[:, :, 1]
    att = torch.einsum("bnhd,bmhd->bnmh", q, k) * (c // heads) ** -0.5
    att = att.softmax(dim=2)
    out = torch.einsum("bnmh,bmhd->bnhd", att, v).reshape(b, n, c)
    return F.linear(out, sd[p + "project.weight"], sd[p + "project.bias"])


def mapper_forward(sd: SD, feat: torch.Tensor, clip_length: int, heads: int, act: str = "relu") -> torch.Tensor:
    """TransformerMapper.forward (layers/Transformer.py:153-161) over Transformer / TransformerLayer /
    MLPTransformer (:52-64, :106-109, :81-87).  `sd` = TransformerMapper.state_dict()."""
    B = feat.shape[0]
    d = sd["prefix_const"].shape[1]
    x = F.linear(feat.float(), sd["linear.weight"], sd["linear.bias"]).view(B, clip_length, -1)
    x = torch.cat((x, sd["prefix_const"].unsqueeze(0).expand(B, -1, -1)), dim=1)
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    fn = ACTS[act]
    for l in range(n_layers):
        p = "transformer.layers.%d." % l
        y = F.layer_norm(x, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        x = x + mapper_attention(sd, p + "attn.", y, heads)
        y = F.layer_norm(x, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
        y = fn(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        x = x + F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x[:, clip_length:]


def mapper_all_forward(sd: SD, feats: torch.Tensor, heads: int, act: str = "relu") -> torch.Tensor:
    """TransformerMapperAllFeatures.forward (layers/Transformer.py:186-203): feats [B, T, dim_clip] (every projected ViT
    token, inference.py:421-444) -> linear per token (+ pos_embeddings when present) || prefix_const -> transformer ->
    the prefix rows.  `sd` = TransformerMapperAllFeatures.state_dict()."""
    B, T = feats.shape[0], feats.shape[1]
    d = sd["prefix_const"].shape[1]
    x = F.linear(feats.float(), sd["linear.weight"], sd["linear.bias"])
    if "pos_embeddings" in sd:
        x = x + sd["pos_embeddings"].unsqueeze(0)
    x = torch.cat((x, sd["prefix_const"].unsqueeze(0).expand(B, -1, -1)), dim=1)
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.layers."))
    fn = ACTS[act]
    for l in range(n_layers):
        p = "transformer.layers.%d." % l
        y = F.layer_norm(x, (d,), sd[p + "norm1.weight"], sd[p + "norm1.bias"], 1e-5)
        x = x + mapper_attention(sd, p + "attn.", y, heads)
        y = F.layer_norm(x, (d,), sd[p + "norm2.weight"], sd[p + "norm2.bias"], 1e-5)
        y = fn(F.linear(y, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
        x = x + F.linear(y, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x[:, T:]


def mlp_mapper_forward(sd: SD, feat: torch.Tensor, prefix_length: int) -> torch.Tensor:
    """Upstream ClipCap MLP mapper (not in this fork; README.md:36 only): Linear -> Tanh -> Linear, viewed
    [B, P, d].  PARITY UNPINNED: no reference code or golden vector exists for it."""
    h = torch.tanh(F.linear(feat.float(), sd["model.0.weight"], sd["model.0.bias"]))
    return F.linear(h, sd["model.2.weight"], sd["model.2.bias"]).view(feat.shape[0], prefix_length, -1)


def clip_text_forward(sd: SD, tokens: torch.Tensor, heads: int) -> torch.Tensor:
    """CLIP.encode_text of OpenAI clip (clip/model.py, pinned nowhere in the reference tree: `clip` is an unpinned git
    dependency, requirements.txt; call sites sampling.py:31, evaluate_model.py ClipScoring): token_embedding +
    positional_embedding -> causal pre-LN transformer with QuickGELU -> ln_final -> features at the end-of-text token
    (the arg-max token id) @ text_projection.  `sd` uses the OpenAI names."""
    x = sd["token_embedding.weight"][tokens] + sd["positional_embedding"]
    B, S, w = x.shape
    hd = w // heads
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.resblocks."))
    causal = torch.ones(S, S, dtype=torch.bool, device=x.device).tril()
    for l in range(n_layers):
        p = "transformer.resblocks.%d." % l
        y = F.layer_norm(x, (w,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        qkv = F.linear(y, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
        q, k, v = (t.view(B, S, heads, hd).transpose(1, 2) for t in qkv.split(w, dim=2))
        att = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
        att = att.masked_fill(~causal, float("-inf")).softmax(-1)
        a = (att @ v).transpose(1, 2).reshape(B, S, w)
        x = x + F.linear(a, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        y = F.layer_norm(x, (w,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        y = quick_gelu(F.linear(y, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]))
        x = x + F.linear(y, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
    x = F.layer_norm(x, (w,), sd["ln_final.weight"], sd["ln_final.bias"], 1e-5)
    x = x[torch.arange(B, device=x.device), tokens.argmax(dim=-1)]
    return x @ sd["text_projection"]


def cos_sim(a, b, normalize=True):
    """sampling.py:14-18."""
    if normalize:
        a = a / torch.norm(a, dim=-1, keepdim=True)
        b = b / torch.norm(b, dim=-1, keepdim=True)
    return a @ b.T


# ------------------------------------------------------------------------------------------------ GPT-2
def _causal_attention(q, k, v, key_mask=None, q_offset=0):
    """q [B,H,Sq,hd], k/v [B,H,Sk,hd]; query i sits at absolute position q_offset + i."""
    hd = q.shape[-1]
    att = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    Sq, Sk = q.shape[2], k.shape[2]
    pos_q = torch.arange(Sq, device=q.device).unsqueeze(1) + q_offset
    allowed = torch.arange(Sk, device=q.device).unsqueeze(0) <= pos_q
    att = att.masked_fill(~allowed, float("-inf"))
    if key_mask is not None:
        att = att.masked_fill(~key_mask.bool()[:, None, None, :], float("-inf"))
    return att.softmax(-1) @ v


def gpt2_forward(sd: SD, embeds: torch.Tensor, heads: int, attention_mask: Optional[torch.Tensor] = None,
                 past: Optional[list] = None, eps: float = 1e-5, last_only: bool = False):
    """GPT2.call(inputs_embeds=E, attention_mask=M) (lms/GPT2.py:17-19 -> HF GPT2LMHeadModel.forward,
    modeling_gpt2.py GPT2Model.forward / GPT2Block / GPT2Attention / GPT2MLP): positions 0..S-1 over
    prefix+text, Conv1D weights [in, out], gelu_new, tied lm_head.  With `past` (list of (k, v) per layer) only
    the new positions are computed -- numerically the KV-cached formulation of the same forward.
    Returns (logits [B,S,V] or [B,V], present)."""
    B, S, d = embeds.shape
    hd = d // heads
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.h."))
    off = past[0][0].shape[2] if past else 0
    h = embeds.float() + sd["transformer.wpe.weight"][off:off + S]
    present = []
    for l in range(n_layers):
        p = "transformer.h.%d." % l
        y = F.layer_norm(h, (d,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], eps)
        qkv = y @ sd[p + "attn.c_attn.weight"] + sd[p + "attn.c_attn.bias"]
        q, k, v = (t.view(B, S, heads, hd).transpose(1, 2) for t in qkv.split(d, dim=2))
        if past:
            k = torch.cat((past[l][0], k), dim=2)
            v = torch.cat((past[l][1], v), dim=2)
        present.append((k, v))
        a = _causal_attention(q, k, v, attention_mask, off).transpose(1, 2).reshape(B, S, d)
        h = h + (a @ sd[p + "attn.c_proj.weight"] + sd[p + "attn.c_proj.bias"])
        y = F.layer_norm(h, (d,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], eps)
        y = gelu_new(y @ sd[p + "mlp.c_fc.weight"] + sd[p + "mlp.c_fc.bias"])
        h = h + (y @ sd[p + "mlp.c_proj.weight"] + sd[p + "mlp.c_proj.bias"])
    if last_only:
        h = h[:, -1:, :]
    h = F.layer_norm(h, (d,), sd["transformer.ln_f.weight"], sd["transformer.ln_f.bias"], eps)
    logits = h @ sd["transformer.wte.weight"].t()
    return (logits[:, 0] if last_only else logits), present


# ------------------------------------------------------------------------------------------------ GPT-J
def _rotate_every_two(x):  # HF modeling_gptj.py rotate_every_two
    x1, x2 = x[..., ::2], x[..., 1::2]
    return torch.stack((-x2, x1), dim=-1).flatten(-2)


def _gptj_rotary(x, positions, rotary_dim):
    """x [B,S,H,hd]; HF create_sinusoidal_positions + apply_rotary_pos_emb on the first rotary_dim dims."""
    inv_freq = 1.0 / (10000 ** (torch.arange(0, rotary_dim, 2, dtype=torch.int64, device=x.device).float() / rotary_dim))
    ang = positions.float()[:, None] * inv_freq[None, :]
    sin = torch.repeat_interleave(torch.sin(ang), 2, dim=-1)[None, :, None, :]
    cos = torch.repeat_interleave(torch.cos(ang), 2, dim=-1)[None, :, None, :]
    rot, rest = x[..., :rotary_dim], x[..., rotary_dim:]
    rot = rot * cos + _rotate_every_two(rot) * sin
    return torch.cat((rot, rest), dim=-1)


def gptj_forward(sd: SD, embeds: torch.Tensor, heads: int, rotary_dim: int,
                 attention_mask: Optional[torch.Tensor] = None, past: Optional[list] = None, eps: float = 1e-5,
                 last_only: bool = False):
    """GPTJ.call (lms/GPTJ.py:16-18 -> HF GPTJForCausalLM: GPTJBlock parallel attention + MLP on ln_1(h),
    bias-free q/k/v/out, interleaved rotary on the first rotary_dim dims, untied biased lm_head)."""
    B, S, d = embeds.shape
    hd = d // heads
    n_layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("transformer.h."))
    off = past[0][0].shape[2] if past else 0
    pos = torch.arange(off, off + S, device=embeds.device)
    h = embeds.float()
    present = []
    for l in range(n_layers):
        p = "transformer.h.%d." % l
        y = F.layer_norm(h, (d,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], eps)
        q = F.linear(y, sd[p + "attn.q_proj.weight"]).view(B, S, heads, hd)
        k = F.linear(y, sd[p + "attn.k_proj.weight"]).view(B, S, heads, hd)
        v = F.linear(y, sd[p + "attn.v_proj.weight"]).view(B, S, heads, hd)
        q = _gptj_rotary(q, pos, rotary_dim).transpose(1, 2)
        k = _gptj_rotary(k, pos, rotary_dim).transpose(1, 2)
        v = v.transpose(1, 2)
        if past:
            k = torch.cat((past[l][0], k), dim=2)
            v = torch.cat((past[l][1], v), dim=2)
        present.append((k, v))
        a = _causal_attention(q, k, v, attention_mask, off).transpose(1, 2).reshape(B, S, d)
        a = F.linear(a, sd[p + "attn.out_proj.weight"])
        m = gelu_new(F.linear(y, sd[p + "mlp.fc_in.weight"], sd[p + "mlp.fc_in.bias"]))
        m = F.linear(m, sd[p + "mlp.fc_out.weight"], sd[p + "mlp.fc_out.bias"])
        h = a + m + h
    if last_only:
        h = h[:, -1:, :]
    h = F.layer_norm(h, (d,), sd["transformer.ln_f.weight"], sd["transformer.ln_f.bias"], eps)
    logits = F.linear(h, sd["lm_head.weight"], sd["lm_head.bias"])
    return (logits[:, 0] if last_only else logits), present


class OracleLM:
    """The slice of the reference LM wrappers the generation loops use (lms/GPT2.py:11-19, lms/GPTJ.py:8-18)."""

    def __init__(self, sd: SD, arch: str, heads: int, rotary_dim: int = 0, eps: float = 1e-5):
        self.sd, self.arch, self.heads, self.rotary_dim, self.eps = sd, arch, heads, rotary_dim, eps

    def get_embedding_size(self) -> int:
        return self.sd["transformer.wte.weight"].shape[1]

    def get_embedding_text(self, tokens: torch.Tensor) -> torch.Tensor:
        return F.embedding(tokens.long(), self.sd["transformer.wte.weight"])

    def forward(self, embeds, attention_mask=None, past=None, last_only=False):
        if self.arch == "gpt2":
            return gpt2_forward(self.sd, embeds, self.heads, attention_mask, past, self.eps, last_only)
        return gptj_forward(self.sd, embeds, self.heads, self.rotary_dim, attention_mask, past, self.eps, last_only)

    def logits(self, embeds, attention_mask=None):  # == language_model.call(...).logits
        return self.forward(embeds, attention_mask)[0]


def caption_model_forward(lm: OracleLM, mapper_fn, tokens, prefix, mask):
    """CLIPCaptionModel.forward (model.py:132-149): logits over [clip_project(prefix) ; wte(tokens)] with an
    all-ones prefix mask prepended to `mask`."""
    emb_text = lm.get_embedding_text(tokens)
    proj = mapper_fn(prefix)
    emb = torch.cat((proj, emb_text), dim=1)
    full_mask = torch.cat((torch.ones(proj.shape[:-1], dtype=torch.bool), mask.bool()), dim=1)
    return lm.logits(emb, full_mask)


def caption_loss(lm: OracleLM, mapper_fn, tokens, prefix, mask, prefix_length: int, ignore_index: int = 0):
    """The teacher-forced loss of model.py:204-211 (training_step) and evaluate_model.py:505-514 (validation):
    positions outside `mask` become token 0, the logits of positions P-1 .. P+L-2 predict tokens 0 .. L-1, mean negative
    log-likelihood over the targets that are not `ignore_index` (0: padding AND any genuine token 0, as in the reference).
    Written out (log-softmax by hand) rather than through F.cross_entropy, which is what the reference calls."""
    tokens = tokens.clone()
    tokens[~mask.bool()] = 0
    logits = caption_model_forward(lm, mapper_fn, tokens, prefix, mask)[:, prefix_length - 1: -1].double()
    lse = torch.logsumexp(logits, dim=-1)
    nll = lse - logits.gather(-1, tokens[..., None]).squeeze(-1)
    keep = tokens != ignore_index
    return (nll * keep).sum() / keep.sum(), (nll * keep).float()


# ------------------------------------------------------------------------------------------------ logit processors
# Tie rule.  The reference sorts with torch.sort(descending=True), whose order among EQUAL logits is unspecified
# (it differs between the CPU and CUDA back ends).  Which of several logits tied exactly at the nucleus boundary
# survive is therefore not defined by the reference; the oracle pins it with a stable sort (lowest index first),
# and so does the CUDA kernel.  Everything else (count kept, every non-tied element) is identical to the reference.
def top_k_top_p_filtering(logits, top_k=0, top_p=0.0, filter_value=-float("inf")):
    """inference.py:24-51 / evaluate_model.py:67-94 (1-D).  Works on a copy."""
    logits = logits.clone()
    assert logits.dim() == 1
    top_k = min(int(top_k), logits.size(-1))
    if top_k > 0:
        logits[logits < torch.topk(logits, top_k)[0][..., -1, None]] = filter_value
    if top_p > 0.0:
        sorted_logits, sorted_indices = torch.sort(logits, descending=True, stable=True)
        cumulative_probs = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
        remove = cumulative_probs > top_p
        remove[..., 1:] = remove[..., :-1].clone()
        remove[..., 0] = 0
        logits[sorted_indices[remove]] = filter_value
    return logits


def repetition_penalty_apply(logits, tokens, penalty):
    """inference.py:53-57 / sampling.py:65-69 (gather -> where -> scatter).  Works on a copy."""
    logits = logits.clone()
    tok = torch.gather(logits, -1, tokens)
    tok = torch.where(tok < 0, tok * penalty, tok / penalty)
    logits.scatter_(-1, tokens, tok)
    return logits


def top_k_top_p_filtering_batch(logits, top_k=0, top_p=0.0, filter_value=float("-inf")):
    """sampling.py:114-162: scalar or per-row top_k (int, fraction of V, or tensor) and top_p (float or [B])."""
    logits = logits.clone()
    batch_size, num_logits = logits.size(0), logits.size(-1)
    if type(top_k) == float:
        top_k = max(1, int(top_k * num_logits)) if 0 < top_k < 1 else int(top_k)
    if type(top_k) == int:
        if top_k > 0:
            cutoff = torch.topk(logits, k=top_k, largest=True).values[:, -1:]
            logits[logits < cutoff] = filter_value
    elif torch.any(top_k > 0):
        top_k = top_k.clamp_max(num_logits)
        for i in range(batch_size):
            k = top_k[i]
            if k <= 0:
                continue
            if k < 1:
                k = max(1, int(k * num_logits))
            cutoff = torch.topk(logits[i], k=int(k), largest=True).values[-1]
            logits[i][logits[i] < cutoff] = filter_value
    if (type(top_p) == float and top_p > 0.0) or (torch.is_tensor(top_p) and torch.any(top_p > 0)):
        if torch.is_tensor(top_p) and top_p.size(-1) != 1:
            top_p = top_p.unsqueeze(-1)
        sorted_logits, sorted_indices = torch.sort(logits, descending=True, dim=-1, stable=True)
        cumulative_probs = torch.cumsum(F.softmax(sorted_logits, dim=-1), dim=-1)
        remove = cumulative_probs > top_p
        remove[:, 1:] = remove[:, :-1].clone()
        remove[:, 0] = False
        remove = remove.scatter(dim=-1, index=sorted_indices, src=remove)
        logits = logits.masked_fill(remove, filter_value)
    return logits


def typical_filtering(logits, typ_p=0.25, min_tokens_to_keep=1, filter_value=float("-inf")):
    """sampling.py:72-102 (typical decoding, Meister et al.): keep the tokens whose information content is closest to
    the entropy of the (already filtered) distribution until their mass reaches typ_p; ties at the cutoff survive.
    typ_p float or per-row tensor [B] / [B, 1]; rows with typ_p <= 0 are only touched when some row is > 0, exactly as
    the reference (a zero budget keeps the single most typical value)."""
    if (type(typ_p) == float and typ_p > 0.0) or (torch.is_tensor(typ_p) and torch.any(typ_p > 0)):
        if torch.is_tensor(typ_p) and typ_p.size(-1) != 1:
            typ_p = typ_p.unsqueeze(-1)
        normalized = F.log_softmax(logits, dim=-1)
        p = normalized.exp()
        entropy = -torch.nansum(normalized * p, dim=-1, keepdim=True)
        shifted_scores = torch.abs(normalized + entropy)
        sorted_scores, sorted_indices = torch.sort(shifted_scores, descending=False, dim=-1, stable=True)
        sorted_p = p.gather(dim=-1, index=sorted_indices)
        cumulative_probs = torch.cumsum(sorted_p, dim=-1)
        last_ind = torch.sum(cumulative_probs < typ_p, dim=-1, keepdim=True)
        last_ind = last_ind.clamp_max(logits.size(-1) - 1)   # (the reference indexes out of range when the mass is never reached)
        remove = sorted_scores > sorted_scores.gather(dim=-1, index=last_ind)
        if min_tokens_to_keep > 1:
            remove[:, :min_tokens_to_keep] = False
        remove = remove.scatter(dim=-1, index=sorted_indices, src=remove)
        logits = logits.masked_fill(remove, filter_value)
    return logits


def multinomial_from_noise(probs: torch.Tensor, q: torch.Tensor, n: int = 1) -> torch.Tensor:
    """torch.multinomial(p, n, replacement=False) == topk(p / q, n) with q ~ Exp(1) drawn by
    `empty_like(p).exponential_(1, generator)` (ATen multinomial kernel; verified against torch.multinomial in
    tests/test_oracle_golden.py).  probs/q [..., V] -> indices [..., n]."""
    return torch.topk(probs / q, n, dim=-1).indices


# ------------------------------------------------------------------------------------------------ generation loops
def generate_beam(lm: OracleLM, embeds: torch.Tensor, beam_size: int = 5, entry_length: int = 67,
                  temperature: float = 1.0, stop_token: int = 13, use_cache: bool = False):
    """inference.py:70-148 for ONE image (embeds [1, P, d]); token ids instead of decoded text.
    Returns (tokens [beam, t] int64, seq_lengths [beam] f32, scores [beam] = scores / seq_lengths, order).
    use_cache=False re-runs the full forward every step exactly like the reference."""
    tokens = None
    scores = None
    seq_lengths = torch.ones(beam_size, device=embeds.device)
    has_stopped = torch.zeros(beam_size, dtype=torch.bool, device=embeds.device)
    past = None
    step_in = embeds
    for _ in range(entry_length):
        if use_cache:
            logits, past = lm.forward(step_in, past=past, last_only=True)
        else:
            logits = lm.logits(embeds)[:, -1, :]
        logits = logits / (temperature if temperature > 0 else 1.0)
        logits = logits.softmax(-1).log()
        if scores is None:
            scores, next_tokens = logits.topk(beam_size, -1)
            embeds = embeds.expand(beam_size, *embeds.shape[1:])
            if past is not None:
                past = [(k.expand(beam_size, *k.shape[1:]), v.expand(beam_size, *v.shape[1:])) for k, v in past]
            next_tokens, scores = next_tokens.permute(1, 0), scores.squeeze(0)
            tokens = next_tokens
        else:
            logits[has_stopped] = -float("inf")
            logits[has_stopped, 0] = 0
            scores_sum = scores[:, None] + logits
            seq_lengths[~has_stopped] += 1
            scores_sum_average = scores_sum / seq_lengths[:, None]
            scores_sum_average, next_tokens = scores_sum_average.view(-1).topk(beam_size, -1)
            src = torch.div(next_tokens, scores_sum.shape[1], rounding_mode="trunc")
            seq_lengths = seq_lengths[src]
            next_tokens = (next_tokens % scores_sum.shape[1]).unsqueeze(1)
            tokens = torch.cat((tokens[src], next_tokens), dim=1)
            embeds = embeds[src]
            if past is not None:
                past = [(k[src], v[src]) for k, v in past]
            scores = scores_sum_average * seq_lengths
            has_stopped = has_stopped[src]
        nxt = lm.get_embedding_text(next_tokens.squeeze(-1)).view(embeds.shape[0], 1, -1)
        embeds = torch.cat((embeds, nxt), dim=1)
        step_in = nxt
        has_stopped = has_stopped + next_tokens.eq(stop_token).squeeze(-1)
        if has_stopped.all():
            break
    scores = scores / seq_lengths
    order = scores.argsort(descending=True)
    return tokens, seq_lengths, scores, order


def generate_no_beam(lm: OracleLM, embeds: torch.Tensor, top_p_values: Sequence[float], noise, entry_length: int = 67,
                     temperature: float = 1.0, stop_token: int = 13, repetition_penalty: float = 1.2,
                     max_stops: int = 1, special_ids: Sequence[int] = (), bos_token: Optional[int] = None,
                     use_cache: bool = False) -> List[List[int]]:
    """inference.py:219-292 (max_stops=1, special_ids=(), bos_token=None) and evaluate_model.py:104-179
    (max_stops=3, special_ids=[50256], bos_token=50256, output stripped of special ids) for ONE image.
    `noise(i, step)` returns the Exp(1) tensor [V] that torch.multinomial would draw for caption i at `step`
    (the RNG contract).  sentence_length_penalty_apply (inference.py:59-68) compares logit
    VALUES with the stop token ID and is a no-op unless a logit equals the id exactly; it is restated as such."""
    if bos_token is not None:
        bos = lm.get_embedding_text(torch.full((embeds.shape[0], 1), bos_token, dtype=torch.int64))
        embeds = torch.cat((embeds, bos), dim=1)
    embeds_init = embeds
    out = []
    for ci, top_p in enumerate(top_p_values):
        tokens: List[int] = []
        embeds = embeds_init
        past, step_in = None, embeds_init
        stops = 0
        for step in range(entry_length):
            if use_cache:
                logits, past = lm.forward(step_in, past=past, last_only=True)
                logits = logits[0]
            else:
                logits = lm.logits(embeds)[0, -1, :]
            if repetition_penalty != 1.0 and tokens:
                logits = repetition_penalty_apply(logits, torch.tensor(tokens, dtype=torch.int64), repetition_penalty)
            logits = logits / (temperature if temperature > 0 else 1.0)
            logits = top_k_top_p_filtering(logits, top_p=top_p, top_k=0)
            probs = F.softmax(logits, dim=-1)
            nxt = int(multinomial_from_noise(probs, noise(ci, step), 1)[0])
            tokens.append(nxt)
            step_in = lm.get_embedding_text(torch.tensor([[nxt]], dtype=torch.int64))
            embeds = torch.cat((embeds, step_in), dim=1)
            if nxt == stop_token:
                stops += 1
            if stops >= max_stops or nxt in special_ids:
                break
        out.append([t for t in tokens if t not in special_ids])
    return out


def generate_clip_guided(lm: OracleLM, embeds: torch.Tensor, image_embedding: torch.Tensor, tokenize_fn, encode_text_fn,
                         bos_token: int, special_ids: Sequence[int], max_decode_length: int = 75,
                         repetition_penalty: float = 1.2, look_ahead: int = 5, branching_factor: int = 3) -> List[int]:
    """evaluate_model.py:182-312 for ONE image (`greedy = True`, step_by_step=False: the other branch reads an undefined
    name): BOS after the prefix (:249-258), then rounds of a depth-first tree search -- at every node the `branching_factor`
    largest (repetition-penalised, :213-216) logits, depth min(look_ahead, room left) (:268), a special id ends a branch
    (:243-245) -- whose leaves are scored by CLIP: decode -> clip.tokenize -> encode_text -> cosine similarity with the
    image embedding (:276-285); the best leaf's tokens are accepted whole (:303).  `tokenize_fn(list of token lists)` and
    `encode_text_fn(clip tokens)` stand for `clip.tokenize(tokenizer.decode_tokens(..))` / `clip_model.encode_text`
    (neither BPE vocabulary exists offline).  Returns the token ids without the special ones (:310)."""
    if image_embedding.dim() == 3 and image_embedding.shape[-2] > 1:
        image_embedding = image_embedding[:, 0, :]
    embeds = torch.cat((embeds, lm.get_embedding_text(torch.full((1, 1), bos_token, dtype=torch.int64))), dim=1)
    tokens: List[int] = []

    def branch(cands, emb, toks, depth):
        logits = lm.logits(emb)[0, -1, :]
        if repetition_penalty != 1.0 and toks:
            logits = repetition_penalty_apply(logits, torch.tensor(toks, dtype=torch.int64), repetition_penalty)
        for t in logits.topk(branching_factor).indices.tolist():
            nt = toks + [t]
            ne = torch.cat((emb, lm.get_embedding_text(torch.tensor([[t]], dtype=torch.int64))), dim=1)
            stop = t in special_ids
            if depth == 0 or stop:
                cands.append((nt, ne, stop))
            else:
                branch(cands, ne, nt, depth - 1)

    while True:
        cands = []
        branch(cands, embeds, tokens, min(look_ahead, max_decode_length - len(tokens)))
        feats = encode_text_fn(tokenize_fn([c[0] for c in cands])).float()
        img = image_embedding / torch.norm(image_embedding)
        feats = feats / torch.norm(feats, dim=-1, keepdim=True)
        best = int((img @ feats.T).argmax())
        tokens, embeds, stop = cands[best]
        if stop or len(tokens) >= max_decode_length:
            break
    return [t for t in tokens if t not in special_ids]


def generate_greedy(lm: OracleLM, embeds: torch.Tensor, entry_length: int, stop_token: int = 13,
                    use_cache: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Batched greedy = generate_beam(beam_size=1) applied per row (batched generation is the
    per-row application of the batch-1 rules).  Returns tokens [B, entry_length] (rows keep decoding after their
    stop token; only the first `lengths[b]` ids are the caption) and lengths [B]."""
    B = embeds.shape[0]
    tokens = torch.zeros(B, entry_length, dtype=torch.int64, device=embeds.device)
    lengths = torch.zeros(B, dtype=torch.int64, device=embeds.device)
    done = torch.zeros(B, dtype=torch.bool, device=embeds.device)
    past, step_in = None, embeds
    for t in range(entry_length):
        if use_cache:
            logits, past = lm.forward(step_in, past=past, last_only=True)
        else:
            logits = lm.logits(embeds)[:, -1, :]
        nxt = logits.argmax(-1)
        tokens[:, t] = nxt
        lengths[~done] = t + 1
        done = done | nxt.eq(stop_token)
        step_in = lm.get_embedding_text(nxt).unsqueeze(1)
        embeds = torch.cat((embeds, step_in), dim=1)
    return tokens, lengths

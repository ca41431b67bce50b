"""Batched logit processors with the reference's signatures (sampling.py:65-69, 72-102, 114-162), executed by the fused
sampler kernel on the device.  They return NEW tensors (the reference mutates / masked_fills)."""
from typing import Optional, Union

import torch

from .engine import Engine


def _engine_of(logits, engine):
    if engine is None:
        raise ValueError("pass engine=<clipcap_b200.Engine> (the processors run on its device)")
    return engine


def repetition_penalty_apply(logits: torch.Tensor, tokens: torch.Tensor, penalty: float, engine: Engine = None):
    """sampling.py:65-69 / inference.py:53-57.  logits [B, V] (or [V]), tokens [B, t] (or [t])."""
    eng = _engine_of(logits, engine)
    squeeze = logits.dim() == 1
    L = logits.unsqueeze(0) if squeeze else logits
    tk = tokens.unsqueeze(0) if tokens.dim() == 1 else tokens
    p = eng.gen_params("sample", 1, repetition_penalty=penalty, q_noise=torch.ones(1, L.shape[1]))
    p.q_ld = 0  # every row reads the same all-ones noise row
    out = eng.sample(L, p, history=tk, return_filtered=True)[1]
    return out[0] if squeeze else out


def top_k_top_p_filtering_batch(logits: torch.Tensor, top_k: Union[int, float, torch.Tensor] = 0,
                                top_p: Union[float, torch.Tensor] = 0.0, filter_value=float("-inf"),
                                engine: Engine = None):
    """sampling.py:114-162: top_k int / fraction of V / per-row tensor; top_p float / per-row tensor."""
    if filter_value != float("-inf"):
        raise ValueError("only filter_value=-inf is supported")
    eng = _engine_of(logits, engine)
    squeeze = logits.dim() == 1
    L = logits.unsqueeze(0) if squeeze else logits
    B, V = L.shape
    kw = {}
    if torch.is_tensor(top_k):
        k = top_k.clone().float()
        frac = (k > 0) & (k < 1)
        k[frac] = torch.clamp((k[frac] * V).floor(), min=1)
        kw["top_k_rows"] = k.clamp(min=0, max=V).to(torch.int32)
        kw["top_k"] = 1
    else:
        if isinstance(top_k, float):
            top_k = max(1, int(top_k * V)) if 0 < top_k < 1 else int(top_k)
        kw["top_k"] = min(int(top_k), V)
    if torch.is_tensor(top_p):
        kw["top_p_rows"] = top_p.reshape(-1).float()
        kw["top_p"] = 1.0
    else:
        kw["top_p"] = float(top_p)
    p = eng.gen_params("sample", 1, q_noise=torch.ones(1, V), **kw)
    p.q_ld = 0
    out = eng.sample(L, p, return_filtered=True)[1]
    return out[0] if squeeze else out


def top_k_top_p_filtering(logits, top_k=0, top_p=0.0, filter_value=-float("inf"), engine: Engine = None):
    """inference.py:24-51 / evaluate_model.py:67-94 (1-D)."""
    return top_k_top_p_filtering_batch(logits, top_k=int(top_k), top_p=float(top_p), filter_value=filter_value, engine=engine)


def typical_filtering(logits: torch.Tensor, typ_p: Union[float, torch.Tensor] = 0.25, min_tokens_to_keep: int = 1,
                      filter_value=float("-inf"), engine: Engine = None):
    """sampling.py:72-102 (typical decoding): keep the tokens whose information content is closest to the entropy until
    their mass reaches typ_p (float or per-row tensor).  A tensor with no positive entry leaves the logits untouched, a
    tensor with some positive entry filters every row, as in the reference."""
    if filter_value != float("-inf"):
        raise ValueError("only filter_value=-inf is supported")
    if min_tokens_to_keep > 1:
        raise ValueError("min_tokens_to_keep > 1 is not supported (the reference never passes it)")
    eng = _engine_of(logits, engine)
    kw = {}
    if torch.is_tensor(typ_p):
        if not bool(torch.any(typ_p > 0)):
            return logits.clone()
        kw["typ_p_rows"] = typ_p.reshape(-1).float()
    elif typ_p > 0.0:
        kw["typ_p"] = float(typ_p)
    else:
        return logits.clone()
    squeeze = logits.dim() == 1
    L = logits.unsqueeze(0) if squeeze else logits
    p = eng.gen_params("sample", 1, q_noise=torch.ones(1, L.shape[1]), **kw)
    p.q_ld = 0
    out = eng.sample(L, p, return_filtered=True)[1]
    return out[0] if squeeze else out


def cos_sim(a: torch.Tensor, b: torch.Tensor, normalize: bool = True) -> torch.Tensor:
    """sampling.py:14-18."""
    if normalize:
        a = a / torch.norm(a, dim=-1, keepdim=True)
        b = b / torch.norm(b, dim=-1, keepdim=True)
    return a @ b.T


def clip_rank(engine: Engine, image: torch.Tensor, text_tokens: torch.Tensor):
    """sampling.py:21-37 on the device: cosine similarity of every candidate caption with the image, both towers on the
    engine.  `image`: one preprocessed image [3, H, W] (Engine.preprocess_images) ; `text_tokens`: [n, 77] ids from
    clip.tokenize (the tokenizer itself is host code and stays the caller's).  Returns a list of n floats."""
    image_features = engine.vit_encode(image.unsqueeze(0) if image.dim() == 3 else image, all_tokens=False)
    text_features = engine.clip_encode_text(text_tokens)
    return cos_sim(text_features, image_features).reshape(-1).tolist()

"""Batched logit processors with the reference's signatures (sampling.py:65-69, 72-102, 114-162), executed by the fused
sampler kernel on the device.  They return NEW tensors (the reference mutates / masked_fills)."""
import math
from typing import Callable, Optional, Union

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


def _filter_kwargs(V: int, top_k, top_p, typ_p):
    """Per-step sampler parameters from the reference's mixed scalar / tensor arguments (sampling.py:114-162, 72-102)."""
    kw = {}
    if torch.is_tensor(top_k):
        if bool(torch.any(top_k > 0)):
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
        if bool(torch.any(top_p > 0)):
            kw["top_p_rows"] = top_p.reshape(-1).float()
            kw["top_p"] = 1.0
    else:
        kw["top_p"] = float(top_p)
    if torch.is_tensor(typ_p):
        if bool(torch.any(typ_p > 0)):
            kw["typ_p_rows"] = typ_p.reshape(-1).float()
    elif typ_p > 0.0:
        kw["typ_p"] = float(typ_p)
    return kw


@torch.no_grad()
def generate(model, inputs: Optional[torch.Tensor], encoder_hidden_states, encoder_attention_mask, eos_token_id, top_p, top_k,
             typ_p, min_length, max_length, repetition_penalty: Optional[float] = None, min_alternate_prob=0,
             force_eos_log_prob=math.log(0.9), engine: Engine = None,
             noise_fn: Optional[Callable[[int, int], torch.Tensor]] = None):
    """sampling.py:165-268: the batched sampling loop with per-row min / max length, top-p / top-k / typical budgets, EOS
    masking before the minimum length, forced completion on a high EOS probability, and the alternate-sample fallback
    (`torch.multinomial(p, 2)`).  `model.forward(input_ids=..., encoder_hidden_states=..., ...)` is called exactly as in the
    reference (any decoder with that surface); everything per step after the logits -- EOS mask, repetition penalty,
    filters, the two draws without replacement -- is ONE launch of the fused sampler kernel.  `noise_fn(rows, V)` supplies
    the Exp(1) noise of the multinomial contract (default: the device RNG); returns the reference's list of
    [tokens, min_length, max_length, top_p, eos_log_probs] groups in completion order."""
    eng = _engine_of(inputs, engine)
    dev = eng.device
    total_max_length = int(max_length.max())
    results = []
    inputs = inputs.to(dev)
    min_length, max_length = min_length.to(dev), max_length.to(dev)
    top_p = top_p.to(dev) if torch.is_tensor(top_p) else top_p
    top_k = top_k.to(dev) if torch.is_tensor(top_k) else top_k
    typ_p = typ_p.to(dev) if torch.is_tensor(typ_p) else typ_p
    eos_probs = torch.empty(inputs.size(0), 0, device=dev)
    cfg = getattr(model, "config", None)
    for i in range(total_max_length):
        if inputs.size(0) == 0:
            break
        outputs = model.forward(input_ids=inputs, encoder_hidden_states=encoder_hidden_states,
                                encoder_attention_mask=encoder_attention_mask, return_dict=True,
                                output_attentions=getattr(cfg, "output_attentions", False),
                                output_hidden_states=getattr(cfg, "output_hidden_states", False))
        last = outputs["logits"][:, -1, :].to(dev, torch.float32).clone()
        B, V = last.shape
        eos_prob = torch.log_softmax(last, dim=-1)[:, eos_token_id]          # raw_p[:, eos].log()
        last[i < min_length, eos_token_id] = float("-inf")
        q = noise_fn(B, V).to(dev, torch.float32) if noise_fn is not None else torch.empty(B, V, device=dev).exponential_(1)
        kw = _filter_kwargs(V, top_k, top_p, typ_p)
        use_pen = repetition_penalty is not None and repetition_penalty > 0
        p = eng.gen_params("sample", 1, q_noise=q, repetition_penalty=float(repetition_penalty) if use_pen else 1.0, **kw)
        nxt, filtered, alt = eng.sample(last, p, history=inputs if use_pen else None, return_filtered=True, return_alt=True)
        next_token = nxt.long().unsqueeze(-1)
        completed = torch.logical_or(next_token.squeeze(-1) == eos_token_id, max_length <= i)
        if force_eos_log_prob < 0:
            completed = torch.logical_or(completed, eos_prob > force_eos_log_prob)
        if bool(torch.any(completed)):
            results.append([inputs[completed], min_length[completed], max_length[completed],
                            top_p[completed] if torch.is_tensor(top_p) else top_p, eos_probs[completed]])
            if min_alternate_prob > 0:
                potential_continue = torch.logical_and(completed, max_length > i)
                if bool(torch.any(potential_continue)):
                    alternate_sample = alt.long().unsqueeze(-1)
                    probs = torch.softmax(filtered, dim=-1)
                    alternate_probs = torch.gather(probs, -1, alternate_sample)
                    potential_continue = torch.logical_and(potential_continue, alternate_sample.squeeze(-1) != eos_token_id)
                    potential_continue = torch.logical_and(potential_continue, alternate_probs.squeeze(-1) > min_alternate_prob)
                    if bool(torch.any(potential_continue)):
                        next_token[potential_continue] = alternate_sample[potential_continue]
                        completed = torch.logical_and(completed, torch.logical_not(potential_continue))
            keep = torch.logical_not(completed)
            inputs, eos_prob, eos_probs, next_token = inputs[keep], eos_prob[keep], eos_probs[keep], next_token[keep]
            if torch.is_tensor(top_p):
                top_p = top_p[keep]
            if torch.is_tensor(top_k):
                top_k = top_k[keep]
            if torch.is_tensor(typ_p):
                typ_p = typ_p[keep]
            min_length, max_length = min_length[keep], max_length[keep]
            keep_h = keep.to(encoder_hidden_states.device)
            encoder_hidden_states = encoder_hidden_states[keep_h]
            encoder_attention_mask = encoder_attention_mask[keep_h]
        inputs = torch.cat([inputs, next_token], dim=-1)
        eos_probs = torch.cat([eos_probs, eos_prob.unsqueeze(-1)], dim=-1)
    if inputs.size(0) > 0:
        results.append([inputs, min_length, max_length, top_p, eos_probs])
    return results


def unique_captions(outputs, num_prompt_tokens: int = 0, special_ids=(), unique: bool = True):
    """The collection step of `sample` (sampling.py:311-323) on token ids: walks the result groups of `generate` in order,
    strips the prompt and the special ids (what `tokenizer.decode(o, skip_special_tokens=True)[len(prompt):]` removes before
    the comparison) and keeps the first occurrence of every caption with its [min_length, max_length, top_p] and EOS
    log-probabilities.  As in the reference, nothing is collected when `unique` is false (`if unique and ... not in
    captions`).  Returns (captions as id tuples, parameters, stats)."""
    special = set(int(t) for t in special_ids)
    captions, parameters, stats = [], [], []
    for group in outputs:
        tokens, min_len, max_len, top_p, eos_probs = group[0], group[1], group[2], group[3], group[4]
        for i, row in enumerate(tokens.tolist()):
            cap = tuple(t for t in row[num_prompt_tokens:] if t not in special)
            if unique and cap not in captions:
                captions.append(cap)
                tp = top_p[i].item() if torch.is_tensor(top_p) else top_p
                parameters.append([int(min_len[i]), int(max_len[i]), tp])
                stats.append({"eos_prob": eos_probs[i], "tokens": row[num_prompt_tokens:]})
    return captions, parameters, stats

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


@torch.no_grad()
def generate(model, inputs: Optional[torch.Tensor], encoder_hidden_states, encoder_attention_mask, eos_token_id, top_p, top_k,
             typ_p, min_length, max_length, repetition_penalty: Optional[float] = None, min_alternate_prob=0,
             force_eos_log_prob=math.log(0.9), engine: Engine = None,
             noise_fn: Optional[Callable[[int, int], torch.Tensor]] = None, check_every: int = 8):
    """sampling.py:165-268: the batched sampling loop with per-row min / max length, top-p / top-k / typical budgets, EOS
    masking before the minimum length, forced completion on a high EOS probability, and the alternate-sample fallback
    (`torch.multinomial(p, 2)`).  `model.forward(input_ids=..., encoder_hidden_states=..., ...)` is called as in the
    reference (any decoder with that surface).

    The loop state lives on the device and no step reads it back: rows are never compacted (a finished row keeps its slot
    and is masked), completion, the alternate continuation and the per-row budgets of the rows still active are elementwise
    device operations, everything between the logits and the two draws without replacement is ONE launch of the fused
    sampler kernel, and the reference's result groups (rows in completion order, a row that continues on its alternate
    sample is reported at that step too, sampling.py:230-246) are assembled after the last step from a [step, row] record.
    Every `check_every` steps one flag is read to stop early once every row has finished.
    `noise_fn(rows, V)` replays the reference's RNG stream (Exp(1) noise of the multinomial contract for the rows still
    active, in row order); it needs the active count on the host, i.e. one read per step -- a test aid.  Default: the
    device RNG for all rows.  Returns the reference's list of [tokens, min_length, max_length, top_p, eos_log_probs]."""
    eng = _engine_of(inputs, engine)
    dev = eng.device
    total = int(max_length.max())
    inputs = inputs.to(dev)
    B, L0 = inputs.shape
    tokens = torch.zeros(B, L0 + total, dtype=inputs.dtype, device=dev)
    tokens[:, :L0] = inputs
    min_length, max_length = min_length.to(dev), max_length.to(dev)
    tp_rows = top_p.to(dev).reshape(-1).float() if torch.is_tensor(top_p) else None
    ty_rows = typ_p.to(dev).reshape(-1).float() if torch.is_tensor(typ_p) else None
    tk_rows = top_k.to(dev).reshape(-1).float() if torch.is_tensor(top_k) else None
    active = torch.ones(B, dtype=torch.bool, device=dev)
    reported = torch.zeros(total, B, dtype=torch.bool, device=dev)
    eos_probs = torch.zeros(B, total, device=dev)
    cfg = getattr(model, "config", None)
    use_pen = repetition_penalty is not None and repetition_penalty > 0
    want_alt = min_alternate_prob > 0
    steps = 0
    for i in range(total):
        if check_every and i and i % check_every == 0 and not bool(active.any()):
            break
        outputs = model.forward(input_ids=tokens[:, :L0 + i], encoder_hidden_states=encoder_hidden_states,
                                encoder_attention_mask=encoder_attention_mask, return_dict=True,
                                output_attentions=getattr(cfg, "output_attentions", False),
                                output_hidden_states=getattr(cfg, "output_hidden_states", False))
        last = outputs["logits"][:, -1, :].to(dev, torch.float32).clone()
        V = last.shape[1]
        eos_prob = torch.log_softmax(last, dim=-1)[:, eos_token_id]          # raw_p[:, eos].log()
        last[:, eos_token_id].masked_fill_(i < min_length, float("-inf"))
        if noise_fn is not None:
            q = torch.ones(B, V, device=dev)
            q[active] = noise_fn(int(active.sum()), V).to(dev, torch.float32)
        else:
            q = torch.empty(B, V, device=dev).exponential_(1)
        kw = {}
        if tk_rows is not None:          # per row: k <= 0 leaves the row alone, a fraction means that share of V (sampling.py:140-148)
            k = torch.where((tk_rows > 0) & (tk_rows < 1), torch.clamp((tk_rows * V).floor(), min=1), tk_rows)
            kw["top_k_rows"], kw["top_k"] = k.clamp(min=0, max=V).to(torch.int32), 1
        else:
            kw["top_k"] = min(max(1, int(top_k * V)) if isinstance(top_k, float) and 0 < top_k < 1 else int(top_k), V)
        if tp_rows is not None:          # sampling.py:149: the filter runs on every row as soon as one ACTIVE budget is positive
            some = ((tp_rows > 0) & active).any()
            kw["top_p_rows"] = torch.where(tp_rows > 0, tp_rows, torch.where(some, 1e-30, 0.0).expand(B))
            kw["top_p"] = 1.0
        else:
            kw["top_p"] = float(top_p)
        if ty_rows is not None:          # sampling.py:77: likewise; a negative budget tells the kernel to skip the row
            some = ((ty_rows > 0) & active).any()
            kw["typ_p_rows"] = torch.where(some, ty_rows.clamp(min=0.0), torch.full_like(ty_rows, -1.0))
        elif typ_p > 0.0:
            kw["typ_p"] = float(typ_p)
        p = eng.gen_params("sample", 1, q_noise=q, repetition_penalty=float(repetition_penalty) if use_pen else 1.0,
                           normalized=True, **kw)
        nxt, filtered, alt = eng.sample(last, p, history=tokens[:, :L0 + i] if use_pen else None, return_filtered=want_alt,
                                        return_alt=want_alt)
        nxt = nxt.long()
        completed = (nxt == eos_token_id) | (max_length <= i)
        if force_eos_log_prob < 0:
            completed |= eos_prob > force_eos_log_prob
        completed &= active
        reported[i] = completed
        if want_alt:
            alt = alt.long()
            alt_p = torch.softmax(filtered, dim=-1).gather(-1, alt.unsqueeze(-1)).squeeze(-1)
            go_on = completed & (max_length > i) & (alt != eos_token_id) & (alt_p > min_alternate_prob)
            nxt = torch.where(go_on, alt, nxt)
            completed = completed & ~go_on
        active = active & ~completed
        tokens[:, L0 + i] = nxt
        eos_probs[:, i] = eos_prob
        steps = i + 1
    # ---- the reference's groups (one read of the record)
    results = []
    rep = reported[:steps].cpu()
    tp_out = top_p.to(dev) if torch.is_tensor(top_p) else top_p
    for i in range(steps):
        if bool(rep[i].any()):
            rows = rep[i].to(dev)
            results.append([tokens[rows, :L0 + i], min_length[rows], max_length[rows],
                            tp_out[rows] if torch.is_tensor(tp_out) else tp_out, eos_probs[rows, :i]])
    if bool(active.any()):
        results.append([tokens[active, :L0 + steps], min_length[active], max_length[active],
                        tp_out[active] if torch.is_tensor(tp_out) else tp_out, eos_probs[active, :steps]])
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

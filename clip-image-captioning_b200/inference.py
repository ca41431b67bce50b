"""Drop-in generation entry points with the reference's signatures (inference.py:70-148, 219-331), running the
whole loop on the device.

The reference loops are batch-1 (`assert logits.shape[0] == 1`, inference.py:253); here `embeds` may hold N images
and the result is the per-image application of the same rules.  For embeds [1, P, d] the return value has the
reference's shape (`generate_beam` -> [best caption]; `generate_no_beam` -> one caption per top_p value); for N > 1 a
list of those per image.
"""
from typing import List, Optional, Sequence

import torch

NO_BEAM_TOP_P = (0.1, 0.2, 0.3, 0.4, 0.5, 0.6, 0.7, 0.8, 0.9)   # inference.py:244


def _with_text_prefix(model, embeds, text_prefix_tokens):
    if text_prefix_tokens is not None:  # inference.py:91-93
        tp = model.language_model.get_embedding_text(text_prefix_tokens)
        if tp.dim() == 2:
            tp = tp.unsqueeze(0)
        if tp.shape[0] != embeds.shape[0]:
            tp = tp.expand(embeds.shape[0], *tp.shape[1:])
        embeds = torch.cat((embeds.to(tp.device), tp), dim=1)
    return embeds


def _stop_id(tokenizer, stop_token):
    if isinstance(stop_token, int):
        return stop_token
    return tokenizer.encode_text(stop_token)[0]


def generate_beam_ids(model, embeds, beam_size=5, entry_length=67, temperature=1.0, stop_id=13):
    """Token-level result of generate_beam for N images: list of (best_tokens, tokens [beam, T], lengths, scores)."""
    eng = model.engine
    p = eng.gen_params("beam", entry_length, stop_token=stop_id, beam_size=beam_size, temperature=temperature)
    tokens, lengths, scores = eng.generate(embeds, p)
    tokens, lengths, scores = tokens.cpu(), lengths.cpu(), scores.cpu()
    out = []
    for i in range(tokens.shape[0]):
        best = int(scores[i].argsort(descending=True)[0])  # inference.py:143-144
        out.append((tokens[i, best, :int(lengths[i, best])].tolist(), tokens[i], lengths[i], scores[i]))
    return out


def generate_beam(model, tokenizer, embeds: torch.Tensor, number_to_generate: int = 1,
                  text_prefix_tokens: Optional[torch.Tensor] = None, beam_size: int = 5, entry_length: int = 67,
                  temperature: float = 1.0, stop_token='.'):
    """inference.py:70-148.  `number_to_generate` > 1 re-enters the reference loop with stale state
    (SURVEY appendix A.4); only the default 1 is supported."""
    if number_to_generate != 1:
        raise ValueError("number_to_generate != 1 is not supported")
    embeds = _with_text_prefix(model, embeds, text_prefix_tokens)
    res = generate_beam_ids(model, embeds, beam_size, entry_length, temperature, _stop_id(tokenizer, stop_token))
    texts = [tokenizer.decode_tokens(r[0]) for r in res]
    return texts if embeds.shape[0] > 1 else [texts[0]]


def generate_no_beam_ids(model, embeds, top_p_values: Sequence[float], entry_length=67, temperature=1.0, stop_id=13,
                         repetition_penalty=1.2, max_stops=1, eos_token=-1, seed=0, q_noise=None, row_ids=None):
    """Nucleus sampling of len(top_p_values) captions per image in ONE batched device loop.
    Row layout: row = ci * N + i (caption ci of image i).  q_noise: optional [T, len(top_p)*N, V] Exp(1) draws (the
    torch.multinomial contract); otherwise an in-kernel Philox stream keyed by (seed, row id, step).
    Returns ids[i][ci] = list of token ids (stop token included, like the reference)."""
    eng = model.engine
    N, C_ = embeds.shape[0], len(top_p_values)
    rows = embeds.repeat(C_, 1, 1) if C_ > 1 else embeds
    tp_rows = torch.tensor([float(tp) for tp in top_p_values for _ in range(N)], dtype=torch.float32)
    if row_ids is None:
        row_ids = torch.arange(C_ * N, dtype=torch.int64)
    p = eng.gen_params("sample", entry_length, stop_token=stop_id, max_stops=max_stops, eos_token=eos_token,
                       temperature=temperature, top_p=1.0, top_p_rows=tp_rows, repetition_penalty=repetition_penalty,
                       seed=seed, q_noise=q_noise, row_ids=row_ids)
    tokens, lengths, _ = eng.generate(rows, p)
    tokens, lengths = tokens.cpu(), lengths.cpu()
    return [[tokens[ci * N + i, :int(lengths[ci * N + i])].tolist() for ci in range(C_)] for i in range(N)]


def generate_no_beam(model, tokenizer, embeds: torch.Tensor, number_to_generate: int = 1,
                     text_prefix_tokens: Optional[torch.Tensor] = None, entry_length: int = 67,
                     temperature: float = 1.0, stop_token='.', repetition_penalty: float = 1.2,
                     desired_sentence_length: int = 50, sentence_length_factor: float = 1.0, seed: int = 0,
                     q_noise: Optional[torch.Tensor] = None):
    """inference.py:219-292: one caption per top_p in 0.1..0.9 (`number_to_generate` is ignored there too).
    `sentence_length_penalty_apply` (inference.py:59-68) compares logit values with the stop-token id and never
    fires; it is kept a no-op."""
    embeds = _with_text_prefix(model, embeds, text_prefix_tokens)
    ids = generate_no_beam_ids(model, embeds, NO_BEAM_TOP_P, entry_length, temperature,
                               _stop_id(tokenizer, stop_token), repetition_penalty, 1, -1, seed, q_noise)
    texts = [[tokenizer.decode_tokens(t) for t in per_image] for per_image in ids]
    return texts if embeds.shape[0] > 1 else texts[0]


def demo_generate_captions(model, tokenizer, clip_model, clip_preproc, image, number_to_generate: int = 1,
                           text_prefix: Optional[str] = None, use_beam_search: bool = False, device="cuda:0",
                           **generation_kwargs):
    """inference.py:295-331.  `clip_model` needs `.encode_image` (model.visual_encoder works); `clip_preproc` maps
    the input to a [3, H, W] tensor (pass `lambda x: x` for tensors that are already preprocessed)."""
    image = clip_preproc(image)
    if image.dim() == 3:
        image = image.unsqueeze(0)
    prefix = clip_model.encode_image(image).to(dtype=torch.float32)
    prefix_embed = model.clip_project(prefix)
    text_prefix_tokens = None
    if text_prefix is not None:
        text_prefix_tokens = torch.tensor(tokenizer.encode_text(text_prefix)).unsqueeze(0)
    fn = generate_beam if use_beam_search else generate_no_beam
    captions = fn(model, tokenizer, prefix_embed, number_to_generate=number_to_generate,
                  text_prefix_tokens=text_prefix_tokens, **generation_kwargs)
    if text_prefix is not None and captions and isinstance(captions[0], str):
        captions = [text_prefix + c for c in captions]
    return captions, prefix

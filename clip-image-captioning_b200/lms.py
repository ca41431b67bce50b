"""Drop-in language-model wrappers: the attribute surface of the reference's `lms` package
(lms/GPT2.py:6-19, lms/GPTJ.py:5-18) served by the CUDA engine.

    lm.get_embedding_size()                       -> int
    lm.get_embedding_text(tokens)                 -> [..., d] f32
    lm.call(inputs_embeds=, labels=, attention_mask=) -> object with .logits [B, S, V] (and .loss with labels)
"""
from types import SimpleNamespace
from typing import Optional

import torch

from .engine import Engine


class _EngineLM:
    arch = "gpt2"

    def __init__(self, engine: Engine):
        if engine.cfg.lm_arch != self.arch:
            raise ValueError("engine was created for %s, not %s" % (engine.cfg.lm_arch, self.arch))
        self.engine = engine

    @property
    def device(self):
        return self.engine.device

    def get_embedding_size(self) -> int:  # lms/GPT2.py:11-12
        return self.engine.cfg.lm_d

    def get_embedding_text(self, tokens: torch.Tensor) -> torch.Tensor:  # lms/GPT2.py:14-15
        return self.engine.embed_tokens(tokens)

    def call(self, inputs_embeds: Optional[torch.Tensor] = None, labels: Optional[torch.Tensor] = None,
             attention_mask: Optional[torch.Tensor] = None):  # lms/GPT2.py:17-19
        if inputs_embeds is None:
            raise ValueError("inputs_embeds is required (the reference never calls the LM with input_ids)")
        logits = self.engine.lm_forward(inputs_embeds, attention_mask)
        loss = None
        if labels is not None:  # HF causal-LM loss: position t predicts label t + 1, ignore_index -100, mean
            B, S, V = logits.shape
            dev = logits.device
            flat = logits.as_strided((B * S, V), (logits.stride(1), 1), logits.storage_offset())
            row_map = (torch.arange(B, device=dev)[:, None] * S + torch.arange(S - 1, device=dev)[None, :]).reshape(-1)
            loss, _, _ = self.engine.cross_entropy(flat, labels.to(dev)[:, 1:].reshape(-1), ignore_index=-100, row_map=row_map)
        return SimpleNamespace(logits=logits, loss=loss)

    __call__ = call

    def eval(self):
        return self

    def to(self, *a, **k):
        return self


class GPT2(_EngineLM):
    """lms.GPT2 (HF GPT2LMHeadModel subclass) replacement."""
    arch = "gpt2"


class GPTJ(_EngineLM):
    """lms.GPTJ (HF GPTJForCausalLM subclass) replacement."""
    arch = "gptj"


class IdTokenizer:
    """Tokenizer stand-in for offline use (no vocab files in the image): ids in, ids out, with the constants of the
    GPT-2 tokenizer the loops rely on: '.' -> 13, bos = eos = 50256."""

    def __init__(self, stop_id: int = 13, bos_token_id: int = 50256, special_ids=(50256,)):
        self.stop_id = stop_id
        self.bos_token_id = bos_token_id
        self.eos_token_id = bos_token_id
        self.all_special_ids = list(special_ids)

    def encode_text(self, text, *a, **k):
        if isinstance(text, str):
            return [self.stop_id]
        return list(text)

    def decode_tokens(self, tokens):
        return [int(t) for t in tokens]

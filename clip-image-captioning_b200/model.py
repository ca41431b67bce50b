"""Drop-in `CLIPCaptionModel` (reference model.py:25-149): the attribute surface the generation loops and
validators use -- `language_model`, `clip_project`, `visual_encoder`, `tokenizer`, `forward` -- backed by one
Engine.  Lightning training hooks are out of scope."""
from typing import Dict, Optional

import torch

from .engine import Engine, EngineConfig
from .lms import GPT2, GPTJ


class _ClipProject:
    """`model.clip_project(prefix)` -> [B, P, d] (layers/Transformer.py:153-161; with use_all_vit_features the
    TransformerMapperAllFeatures of layers/Transformer.py:164-203 over [B, tokens, 512] features)."""

    def __init__(self, engine: Engine):
        self.engine = engine

    def __call__(self, prefix: torch.Tensor) -> torch.Tensor:
        return self.engine.map_prefix(prefix)

    forward = __call__


class _VisualEncoder:
    """`clip_model.encode_image(img)` / `model.visual_encoder(img)` -> [B, 512] f32 (inference.py:311), or, when the
    model was built with use_all_vit_features (inference.py:421-444), every projected token [B, 50, 512]."""

    def __init__(self, engine: Engine):
        self.engine = engine

    def __call__(self, images: torch.Tensor) -> torch.Tensor:
        if images.dim() == 3:
            images = images.unsqueeze(0)
        return self.engine.vit_encode(images)

    forward = encode_image = __call__


class CLIPCaptionModel:
    def __init__(self, engine: Engine, tokenizer=None, validator=None):
        self.engine = engine
        self.language_model = (GPTJ if engine.cfg.lm_arch == "gptj" else GPT2)(engine)
        self.tokenizer = tokenizer
        self.validator = validator
        self.lm_embedding_size = self.language_model.get_embedding_size()
        self.clip_project = _ClipProject(engine)
        self.visual_encoder = _VisualEncoder(engine) if engine.cfg.vit else None
        self.prefix_length = engine.cfg.map_prefix_len

    @classmethod
    def from_state_dict(cls, sd: Dict[str, torch.Tensor], cfg: EngineConfig, device: int = 0, tokenizer=None,
                        language_model_sd: Optional[Dict[str, torch.Tensor]] = None, visual_sd=None):
        """`sd`: a reference checkpoint state_dict (`clip_project.*`, optionally `language_model.*` /
        `visual_encoder.*`).  The reference re-supplies the LM at load time (model.py:38, inference.py:458-469):
        pass it as `language_model_sd` (HF names) and the CLIP visual tower as `visual_sd` (OpenAI names)."""
        eng = Engine(cfg, device)
        eng.load_state_dict(sd)
        if language_model_sd is not None:
            eng.load_state_dict(language_model_sd, prefix="language_model.")
        if visual_sd is not None:
            eng.load_state_dict(visual_sd, prefix="visual.")
        eng.check_weights()
        return cls(eng, tokenizer)

    @property
    def device(self):
        return self.engine.device

    def eval(self):
        return self

    def to(self, *a, **k):
        return self

    def forward(self, tokens: torch.Tensor, prefix: torch.Tensor, mask: Optional[torch.Tensor] = None,
                labels: Optional[torch.Tensor] = None):  # model.py:132-149
        embedding_text = self.language_model.get_embedding_text(tokens)
        prefix_projections = self.clip_project(prefix)
        embedding_cat = torch.cat((prefix_projections, embedding_text), dim=1)
        if mask is not None:
            ones = torch.ones(prefix_projections.shape[:-1], dtype=torch.bool, device=self.device)
            mask = torch.cat((ones, mask.to(self.device).bool()), dim=1)
        if labels is not None:
            dummy = torch.zeros(tokens.shape[0], self.prefix_length, dtype=torch.int64, device=self.device)
            labels = torch.cat((dummy, tokens.to(self.device).long()), dim=1)
        return self.language_model.call(inputs_embeds=embedding_cat, labels=labels, attention_mask=mask)

    __call__ = forward

    def caption_loss(self, tokens: torch.Tensor, prefix: torch.Tensor, mask: Optional[torch.Tensor] = None,
                     ignore_index: int = 0) -> torch.Tensor:
        """The teacher-forced loss of evaluate_model.py:497-516 (validation) and model.py:204-211 (training_step), forward
        only: `cross_entropy(forward(tokens, prefix, mask).logits[:, P-1:-1].reshape(-1, V), tokens.flatten(),
        ignore_index=0)`.  The logits slice is addressed through a row map instead of being copied."""
        tokens = tokens.to(self.device)
        if mask is None:
            mask = tokens.ge(0)                      # model.py:204-205
            tokens = tokens.masked_fill(~mask, 0)
        out = self.forward(tokens, prefix, mask)
        logits = out.logits                          # [B, P + L, V] view of a [B * (P + L), ldv] buffer
        B, S, V = logits.shape
        L, P = tokens.shape[1], self.prefix_length
        flat = logits.as_strided((B * S, V), (logits.stride(1), 1), logits.storage_offset())
        row_map = (torch.arange(B, device=self.device)[:, None] * S + (P - 1) + torch.arange(L, device=self.device)[None, :])
        loss, _, _ = self.engine.cross_entropy(flat, tokens.reshape(-1), ignore_index=ignore_index, row_map=row_map.reshape(-1))
        return loss


class CLIPCaptionPrefixOnly(CLIPCaptionModel):  # model.py:219-225 (training-time distinction only)
    pass

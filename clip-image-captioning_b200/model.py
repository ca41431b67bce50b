"""Drop-in `CLIPCaptionModel` (reference model.py:25-149): the attribute surface the generation loops and
validators use -- `language_model`, `clip_project`, `visual_encoder`, `tokenizer`, `forward` -- backed by one
Engine.  Lightning training hooks are out of scope."""
from types import SimpleNamespace
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

    @staticmethod
    def config_from_checkpoint(state_dict: Dict[str, torch.Tensor], hparams: Dict, language_model=None, visual_encoder=None,
                               **overrides) -> EngineConfig:
        """EngineConfig of a reference checkpoint: the mapper from the Lightning `hyper_parameters` (the kwargs of
        model.py:25-78 / train.py), the language model from `language_model.config` (HF GPT2Config / GPTJConfig) or, without
        one, from the shapes of the checkpoint's `language_model.*` tensors (head count then comes from `lm_heads=`), the
        ViT from the `visual_encoder.*` / `visual.*` tensors.  `overrides` win (max_images, max_beam, max_ctx, ...)."""
        kw = {}
        hp = dict(hparams or {})
        lm_cfg = getattr(language_model, "config", None)
        lm_sd = {k[len("language_model."):]: v for k, v in state_dict.items() if k.startswith("language_model.")}
        if not lm_sd and language_model is not None and hasattr(language_model, "state_dict"):
            lm_sd = language_model.state_dict()
        if lm_cfg is not None:
            arch = "gptj" if "gptj" in type(lm_cfg).__name__.lower() else "gpt2"
            kw.update(lm_arch=arch, lm_d=lm_cfg.n_embd, lm_layers=lm_cfg.n_layer, lm_heads=lm_cfg.n_head,
                      lm_vocab=lm_cfg.vocab_size, lm_n_pos=lm_cfg.n_positions)
            if arch == "gptj":
                kw.update(lm_rotary_dim=lm_cfg.rotary_dim)
        elif lm_sd:
            arch = "gptj" if any(".attn.q_proj." in k for k in lm_sd) else "gpt2"
            wte = lm_sd["transformer.wte.weight"]
            layers = 1 + max(int(k.split(".")[2]) for k in lm_sd if k.startswith("transformer.h."))
            kw.update(lm_arch=arch, lm_d=wte.shape[1], lm_vocab=wte.shape[0], lm_layers=layers)
            if "transformer.wpe.weight" in lm_sd:
                kw.update(lm_n_pos=lm_sd["transformer.wpe.weight"].shape[0])
            if "lm_heads" not in overrides:
                raise ValueError("the head count is not stored in the tensors: pass language_model (with .config) or lm_heads=")
        if hp:
            all_feats = bool(hp.get("use_all_vit_features", False))
            kw.update(map_kind="transformer_all" if all_feats else "transformer",
                      map_dim_clip=int(hp["prefix_size"]), map_prefix_len=int(hp["prefix_length"]),
                      map_clip_len=int(hp["clip_prefix_length"]), map_heads=int(hp.get("num_attention_heads", 8)),
                      map_layers=int(hp.get("num_layers", 8)), map_mlp_ratio=float(hp.get("mlp_ratio", 4.0)),
                      map_act=str(hp.get("act_fn_name", "relu")))
        vis = {k.split(".", 1)[1]: v for k, v in state_dict.items() if k.startswith(("visual_encoder.", "visual."))}
        if not vis and visual_encoder is not None and hasattr(visual_encoder, "state_dict"):
            vis = visual_encoder.state_dict()
        if "conv1.weight" in vis:
            w = vis["conv1.weight"]
            n_tok = vis["positional_embedding"].shape[0]
            grid = int(round((n_tok - 1) ** 0.5))
            kw.update(vit=True, vit_width=w.shape[0], vit_patch=w.shape[-1], vit_image=grid * w.shape[-1],
                      vit_layers=1 + max(int(k.split(".")[2]) for k in vis if k.startswith("transformer.resblocks.")),
                      vit_out=vis["proj"].shape[1], vit_heads=w.shape[0] // 64)
        else:
            kw.update(vit=False)
        if kw.get("map_kind") == "transformer_all":
            # one mapper token per ViT token (layers/Transformer.py:190-201): `clip_length` only sizes pos_embeddings there
            if "clip_project.pos_embeddings" in state_dict:
                kw["map_clip_len"] = state_dict["clip_project.pos_embeddings"].shape[0]
            elif kw.get("vit"):
                kw["map_clip_len"] = (kw["vit_image"] // kw["vit_patch"]) ** 2 + 1
        # capacities the reference-signature defaults need: generate_beam(beam_size=5, entry_length=67), generate_no_beam (nine
        # captions per image, entry_length=67), evaluate_model.generate_no_beam (BOS + max_decode_length <= 77)
        kw.setdefault("max_beam", 5)
        n_pos = kw.get("lm_n_pos", EngineConfig.lm_n_pos)
        want_ctx = kw.get("map_prefix_len", EngineConfig.map_prefix_len) + 1 + 77
        kw.setdefault("max_ctx", min(want_ctx, n_pos) if kw.get("lm_arch", "gpt2") == "gpt2" else want_ctx)
        kw.update(overrides)
        fields = set(EngineConfig.__dataclass_fields__)
        unknown = set(kw) - fields
        if unknown:
            raise TypeError("unknown EngineConfig fields: %s" % sorted(unknown))
        return EngineConfig(**kw)

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path: str, language_model=None, tokenizer=None, visual_encoder=None,
                             validator=None, strict: bool = False, device: int = 0, hparams: Optional[Dict] = None,
                             **overrides):
        """The reference's `CLIPCaptionModel.load_from_checkpoint(checkpoint_path=..., language_model=..., tokenizer=...,
        visual_encoder=..., validator=None, strict=False)` (inference.py:458-462, evaluate_model.py:596-599): reads a
        Lightning `.ckpt` (`state_dict` + `hyper_parameters`) -- or the bare `state_dict` file of inference.py:469, with the
        model kwargs in `hparams=` -- sizes an engine for it and loads the weights.  `language_model` / `visual_encoder` are the
        modules the reference re-supplies at load time (`save_hyperparameters(ignore=["language_model"])`, model.py:38):
        anything with `.state_dict()` (and `.config` for the LM) -- their tensors fill what the checkpoint does not hold
        (prefix-only checkpoints carry no language model).  With strict=True unused checkpoint keys raise."""
        ckpt = torch.load(checkpoint_path, map_location="cpu", weights_only=False)
        if isinstance(ckpt, dict) and "state_dict" in ckpt:
            sd, hp = ckpt["state_dict"], dict(ckpt.get("hyper_parameters", {}))
        else:
            sd, hp = ckpt, {}
        hp.update(hparams or {})
        cfg = cls.config_from_checkpoint(sd, hp, language_model, visual_encoder, **overrides)
        eng = Engine(cfg, device)
        try:
            # the re-supplied modules first, the checkpoint's own tensors last (they win, as in load_state_dict)
            if language_model is not None and hasattr(language_model, "state_dict"):
                eng.load_state_dict(language_model.state_dict(), prefix="language_model.")
            if visual_encoder is not None and hasattr(visual_encoder, "state_dict") and cfg.vit:
                eng.load_state_dict(visual_encoder.state_dict(), prefix="visual.")
            unused = eng.load_state_dict(sd)
            if strict and unused:
                raise RuntimeError("unexpected keys in the checkpoint: %s" % unused[:8])
            eng.check_weights()
        except Exception:
            eng.close()
            raise
        model = cls(eng, tokenizer, validator)
        model.hparams = SimpleNamespace(**hp)
        return model

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

"""Drop-ins for the sampler plug-in path of evaluate_model.py (:104-179 generate_no_beam, :182-312 generate_clip_guided,
:355-419 CaptionSamplerBase / NoBeamCaptionSampler / ClipGuidedCaptionSampler)."""
from typing import Callable, List, Optional, Sequence

import torch

from . import inference, sampling


def generate_no_beam(model, embeds: torch.Tensor, top_p_values: Sequence[float] = inference.NO_BEAM_TOP_P,
                     text_prefix_tokens: Optional[torch.Tensor] = None, max_decode_length: int = 75,
                     temperature: float = 1.0, stop_token='.', repetition_penalty: float = 1.2, max_stops: int = 3,
                     seed: int = 0, q_noise: Optional[torch.Tensor] = None):
    """evaluate_model.py:104-179: BOS embedding appended after the prefix (:124-133), stop after `max_stops` stop
    tokens or on any special id (:169-172), special ids stripped from the output (:174)."""
    assert max_decode_length <= 77, "maximum context length for CLIP models is 77"
    tokenizer = model.tokenizer
    special = list(tokenizer.all_special_ids)
    stop_id = stop_token if isinstance(stop_token, int) else tokenizer.encode_text(stop_token)[0]
    bos = torch.full((embeds.shape[0], 1), tokenizer.bos_token_id, dtype=torch.int64)
    tp = bos if text_prefix_tokens is None else torch.cat((bos, text_prefix_tokens.cpu().long()), dim=1)
    embeds = inference._with_text_prefix(model, embeds, tp)
    if len(special) > 1:
        raise ValueError("only one special (EOS) id is supported")
    ids = inference.generate_no_beam_ids(model, embeds, top_p_values, max_decode_length, temperature, stop_id,
                                         repetition_penalty, max_stops, special[0] if special else -1, seed, q_noise)
    texts = [[tokenizer.decode_tokens([t for t in cap if t not in special]) for cap in per_image] for per_image in ids]
    return texts if embeds.shape[0] > 1 else texts[0]


class CaptionSamplerBase:  # evaluate_model.py:355-367
    def sample(self, model, image_tensor, image=None):
        if image_tensor.dim() == 3:
            image_tensor = image_tensor.unsqueeze(0)
        image_embedding = model.visual_encoder(image_tensor)
        prefix = model.clip_project(image_embedding)
        return self.generate_captions(model, prefix, image_embedding, image)

    def get_description(self):
        raise NotImplementedError()

    def generate_captions(self, model, prefix, image_embedding, image):
        raise NotImplementedError()


class NoBeamCaptionSampler(CaptionSamplerBase):  # evaluate_model.py:370-385
    def __init__(self, top_p_values=(0.1,), temperature: float = 1.0, repetition_penalty: float = 1.2, seed: int = 0,
                 max_decode_length: int = 75):
        self.top_p_values = list(top_p_values)
        self.temperature = temperature
        self.repetition_penalty = repetition_penalty
        self.seed = seed
        self.max_decode_length = max_decode_length   # (the reference takes generate_no_beam's default, 75)

    def get_description(self):
        return f'NoBeam(rep_p={self.repetition_penalty}, temp={self.temperature}, top_p={self.top_p_values})'

    def generate_captions(self, model, prefix, image_embedding, image):
        return generate_no_beam(model, prefix, top_p_values=self.top_p_values, temperature=self.temperature,
                                repetition_penalty=self.repetition_penalty, seed=self.seed,
                                max_decode_length=self.max_decode_length)


class BeamCaptionSampler(CaptionSamplerBase):
    """generate_beam (inference.py:70-148) behind the same plug-in interface."""

    def __init__(self, beam_size: int = 5, entry_length: int = 67, temperature: float = 1.0):
        self.beam_size, self.entry_length, self.temperature = beam_size, entry_length, temperature

    def get_description(self):
        return f'Beam(beam_size={self.beam_size}, temp={self.temperature})'

    def generate_captions(self, model, prefix, image_embedding, image):
        return inference.generate_beam(model, model.tokenizer, prefix, beam_size=self.beam_size,
                                       entry_length=self.entry_length, temperature=self.temperature)


class EngineClipModel:
    """`clip_model` for the text side of CLIP scoring (evaluate_model.py:276-279, sampling.py:30-31): `.encode_text(tokens)`
    on the engine's CLIP text tower (EngineConfig(text=True)), in chunks of the context's `max_texts`."""

    def __init__(self, engine):
        self.engine = engine

    def encode_text(self, tokens: torch.Tensor) -> torch.Tensor:
        n = self.engine.cfg.max_texts
        return torch.cat([self.engine.clip_encode_text(tokens[i:i + n]) for i in range(0, tokens.shape[0], n)], dim=0)


def _next_token_logits(model, base: torch.Tensor, new_tokens: List[List[int]]) -> torch.Tensor:
    """Logits of the last position for every node of one tree level: rows = base embeddings || the node's new tokens, through
    the batched LM forward (last position only), in chunks that fit the context's `max_lm_tokens`."""
    eng = model.engine
    n, t = len(new_tokens), len(new_tokens[0])
    S = base.shape[1] + t
    step = max(1, min(n, eng.cfg.max_lm_tokens // S if eng.cfg.max_lm_tokens else eng.cfg.max_images))
    base = base.to(eng.device)
    out = []
    for i in range(0, n, step):
        rows = new_tokens[i:i + step]
        emb = base.expand(len(rows), -1, -1)
        if t:
            emb = torch.cat((emb, model.language_model.get_embedding_text(torch.tensor(rows, dtype=torch.int64))), dim=1)
        out.append(eng.lm_forward(emb.contiguous(), last_only=True).clone())
    return torch.cat(out, dim=0)


def generate_clip_guided(device, clip_image_embedding: torch.Tensor, model, clip_model, embeds: torch.Tensor,
                         text_prefix_tokens: Optional[torch.Tensor] = None, max_decode_length: int = 75, temperature: float = 1.0,
                         repetition_penalty: float = 1.2, look_ahead=5, branching_factor=3, step_by_step=False,
                         clip_tokenize: Optional[Callable] = None):
    """evaluate_model.py:182-312, batch 1 like the reference (`assert embeds.shape[0] == 1`).  The reference expands the
    search tree depth first with one full LM forward per node; here every LEVEL of the tree is one batched forward
    (branching_factor ** level rows), the leaves are put back into the reference's depth-first order (the arg-max over the
    similarities takes the first maximum) and scored by one batched `clip_model.encode_text`.  `clip_tokenize(list of
    decoded captions) -> [n, ctx] ids` is clip.tokenize (with truncate=True); it has to be supplied because no BPE
    vocabulary is available offline.  `temperature` is unused by the reference's greedy branch and here.
    step_by_step=True reads an undefined name in the reference (NameError): not reproduced."""
    assert max_decode_length <= 77, "maximum context length for CLIP models is 77"
    if step_by_step:
        raise NotImplementedError("step_by_step=True is broken in the reference (undefined stop_token)")
    if embeds.shape[0] != 1:
        raise ValueError("generate_clip_guided works on one image (the reference asserts batch size 1)")
    if clip_tokenize is None:
        raise RuntimeError("pass clip_tokenize= (clip.tokenize): no BPE vocabulary is bundled")
    if clip_image_embedding.dim() == 3 and clip_image_embedding.shape[-2] > 1:
        clip_image_embedding = clip_image_embedding[:, 0, :]
    tokenizer = model.tokenizer
    special = set(tokenizer.all_special_ids)
    eng = model.engine
    bos = torch.full((1, 1), tokenizer.bos_token_id, dtype=torch.int64)
    tp = bos if text_prefix_tokens is None else torch.cat((bos, text_prefix_tokens.cpu().long()), dim=1)
    base = inference._with_text_prefix(model, embeds, tp).to(eng.device)
    img = clip_image_embedding.to(eng.device).float()
    img = img / torch.norm(img)
    tokens: List[int] = []
    while True:
        depth = min(look_ahead, max_decode_length - len(tokens))
        frontier = [((), list(tokens))]            # (choice path, token ids accepted + chosen so far)
        leaves = []
        for level in range(depth + 1):
            logits = _next_token_logits(model, base, [f[1] for f in frontier])
            if repetition_penalty != 1.0 and frontier[0][1]:
                hist = torch.tensor([f[1] for f in frontier], dtype=torch.int64)
                logits = sampling.repetition_penalty_apply(logits, hist, repetition_penalty, engine=eng)
            idx = logits.topk(branching_factor, dim=-1).indices.tolist()
            nxt = []
            for (path, toks), row in zip(frontier, idx):
                for j, t in enumerate(row):
                    node = (path + (j,), toks + [t], t in special)
                    (leaves if level == depth or node[2] else nxt).append(node)
            frontier = [(p_, t_) for p_, t_, _ in nxt]
            if not frontier:
                break
        leaves.sort(key=lambda c: c[0])             # depth-first order of the reference's recursion
        texts = [tokenizer.decode_tokens(c[1]) for c in leaves]
        feats = clip_model.encode_text(clip_tokenize(texts)).float().to(eng.device)
        feats = feats / torch.norm(feats, dim=-1, keepdim=True)
        best = int((img @ feats.T).reshape(-1).argmax())
        _, tokens, stop = leaves[best]
        if stop or len(tokens) >= max_decode_length:
            break
    return tokenizer.decode_tokens([t for t in tokens if t not in special])


class ClipGuidedCaptionSampler(CaptionSamplerBase):  # evaluate_model.py:388-419
    """`clip_scoring` needs `.clip_model` (with encode_text) and `.embed_image(image tensor)`; with an engine that holds both
    CLIP towers, `EngineClipScoring` below provides them."""

    def __init__(self, clip_scoring, branching_factor: int = 3, look_ahead: int = 4, repetition_penalty: float = 1.2,
                 clip_tokenize: Optional[Callable] = None, max_decode_length: int = 75):
        self.clip_scoring = clip_scoring
        self.branching_factor = branching_factor
        self.look_ahead = look_ahead
        self.repetition_penalty = repetition_penalty
        self.clip_tokenize = clip_tokenize
        self.max_decode_length = max_decode_length

    def get_description(self):
        return f'ClipGuided(branching={self.branching_factor}, look_ahead={self.look_ahead}, rep_p={self.repetition_penalty})'

    def generate_captions(self, model, prefix, image_embedding, image):
        # (the reference re-embeds the raw image with clip_scoring; the embedding of the sampler's own ViT pass is the same
        # tensor when both use the engine's tower, and is taken when no raw image is handed in)
        emb = self.clip_scoring.embed_image(image) if image is not None else image_embedding
        caption = generate_clip_guided(model.device, emb, model, self.clip_scoring.clip_model, prefix,
                                       branching_factor=self.branching_factor, look_ahead=self.look_ahead,
                                       repetition_penalty=self.repetition_penalty, clip_tokenize=self.clip_tokenize,
                                       max_decode_length=self.max_decode_length)
        return [caption]


class EngineClipScoring:
    """The two calls of the reference's ClipScoring that ClipGuidedCaptionSampler uses (evaluate_model.py:404-412), on the
    engine's CLIP towers: `clip_model.encode_text` and `embed_image` (preprocessed image tensor -> [1, dim])."""

    def __init__(self, engine):
        self.engine = engine
        self.clip_model = EngineClipModel(engine)

    def embed_image(self, image: torch.Tensor) -> torch.Tensor:
        return self.engine.vit_encode(image.unsqueeze(0) if image.dim() == 3 else image)

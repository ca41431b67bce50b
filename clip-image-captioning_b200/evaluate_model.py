"""Drop-ins for the sampler plug-in path of evaluate_model.py (:104-179 generate_no_beam, :355-385
CaptionSamplerBase / NoBeamCaptionSampler)."""
from typing import Optional, Sequence

import torch

from . import inference


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

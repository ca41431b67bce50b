"""Checkpoint ingestion: `CLIPCaptionModel.load_from_checkpoint` (reference inference.py:458-469, evaluate_model.py:596-599).

pytorch_lightning is not installed here, so no genuine `.ckpt` can be produced: the files below have the documented layout
of a Lightning checkpoint (`state_dict` + `hyper_parameters` = the kwargs `save_hyperparameters(ignore=["language_model"])`
keeps, model.py:38) around the tensors of the reference-generated fixtures, i.e. the format is restated, not pinned.
"""
import os
import sys
from types import SimpleNamespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
TOL = 2e-2


def fixture(arch):
    return torch.load(os.path.join(GOLDEN, "tiny_%s.pt" % arch), weights_only=False)


def hparams_of(fx, **extra):
    hp = dict(prefix_size=fx["dim_clip"], prefix_length=fx["P"], clip_prefix_length=fx["CL"], num_attention_heads=fx["map_heads"],
              num_layers=2, mlp_ratio=4.0, prefix_init_std=1.0, act_fn_name="relu", use_all_vit_features=False,
              pos_embeddings=False, train_visual_encoder=False, autoclip_p=10, max_log_samples=64)
    hp.update(extra)
    return hp


def full_state_dict(fx, with_lm=True, with_vit=True):
    sd = {"clip_project." + k: v for k, v in fx["sd_mapper"].items()}
    if with_lm:
        sd.update({"language_model." + k: v for k, v in fx["sd_lm"].items()})
    if with_vit:
        sd.update({"visual_encoder." + k: v for k, v in fx["sd_vit"].items()})
    return sd


class Holder:   # stands in for the nn.Module the reference re-supplies (anything with state_dict() [+ config])
    def __init__(self, sd, config=None):
        self._sd, self.config = sd, config

    def state_dict(self):
        return self._sd


def test_config_from_checkpoint_tensors_only():
    import clipcap_b200 as cc
    fx = fixture("gpt2")
    cfg = cc.model.CLIPCaptionModel.config_from_checkpoint(full_state_dict(fx), hparams_of(fx), lm_heads=2, vit_heads=2, max_images=8)
    assert (cfg.lm_arch, cfg.lm_d, cfg.lm_layers, cfg.lm_vocab, cfg.lm_n_pos, cfg.lm_heads) == ("gpt2", 128, 2, 503, 64, 2)
    assert (cfg.map_kind, cfg.map_dim_clip, cfg.map_prefix_len, cfg.map_clip_len, cfg.map_heads, cfg.map_layers) == ("transformer", 64, 4, 4, 8, 2)
    assert (cfg.vit, cfg.vit_image, cfg.vit_patch, cfg.vit_width, cfg.vit_layers, cfg.vit_out) == (True, 64, 32, 64, 2, 64)
    with pytest.raises(ValueError):   # heads are not recoverable from tensor shapes
        cc.model.CLIPCaptionModel.config_from_checkpoint(full_state_dict(fx), hparams_of(fx))


def test_config_from_checkpoint_with_hf_config_and_all_features():
    from transformers import GPTJConfig
    import clipcap_b200 as cc
    fx = fixture("gptj")
    lm = Holder(fx["sd_lm"], GPTJConfig(vocab_size=fx["V"], n_positions=64, n_embd=128, n_layer=2, n_head=2, rotary_dim=16))
    sd = full_state_dict(fx, with_lm=False)
    cfg = cc.model.CLIPCaptionModel.config_from_checkpoint(sd, hparams_of(fx), lm, vit_heads=2)
    assert (cfg.lm_arch, cfg.lm_d, cfg.lm_layers, cfg.lm_heads, cfg.lm_vocab, cfg.lm_rotary_dim) == ("gptj", 128, 2, 2, fx["V"], 16)
    # use_all_vit_features: one mapper token per ViT token (5 for a 64-pixel image with 32-pixel patches)
    cfg = cc.model.CLIPCaptionModel.config_from_checkpoint(sd, hparams_of(fx, use_all_vit_features=True), lm, vit_heads=2)
    assert (cfg.map_kind, cfg.map_clip_len) == ("transformer_all", 5)
    sd["clip_project.pos_embeddings"] = torch.zeros(7, 128)
    cfg = cc.model.CLIPCaptionModel.config_from_checkpoint(sd, hparams_of(fx, use_all_vit_features=True, pos_embeddings=True), lm, vit_heads=2)
    assert cfg.map_clip_len == 7
    with pytest.raises(TypeError):
        cc.model.CLIPCaptionModel.config_from_checkpoint(sd, hparams_of(fx), lm, no_such_field=1)


@pytest.mark.gpu
@pytest.mark.parametrize("arch", ["gpt2", "gptj"])
def test_load_lightning_checkpoint_full_model(tmp_path, arch):
    """A checkpoint that carries everything (train_visual_encoder runs save the ViT too): forward logits and greedy captions
    equal the reference's."""
    import clipcap_b200 as cc
    from clipcap_b200 import inference
    fx = fixture(arch)
    path = str(tmp_path / "full.ckpt")
    torch.save({"state_dict": full_state_dict(fx), "hyper_parameters": hparams_of(fx), "epoch": 3, "global_step": 17}, path)
    extra = dict(lm_rotary_dim=16) if arch == "gptj" else {}
    model = cc.model.CLIPCaptionModel.load_from_checkpoint(checkpoint_path=path, tokenizer=None, validator=None, strict=False,
                                                           lm_heads=2, vit_heads=2, max_images=8, max_ctx=32, page_tokens=4, **extra)
    assert model.hparams.prefix_length == fx["P"] and model.prefix_length == fx["P"]
    feat = model.visual_encoder(fx["images"])
    logits = model(fx["tokens"], feat, fx["mask"]).logits
    ref = fx["logits_tf"]
    assert (logits.float().cpu() - ref).abs().max().item() <= TOL * ref.abs().max().item()
    prefix = model.clip_project(feat)
    for i, want in enumerate(fx["greedy"]):
        got = inference.generate_beam_ids(model, prefix[i:i + 1], beam_size=1, entry_length=10, stop_id=fx["stop_id"])[0][0]
        assert got == want
    model.engine.close()


@pytest.mark.gpu
def test_load_prefix_only_checkpoint_with_resupplied_modules(tmp_path):
    """CLIPCaptionPrefixOnly checkpoints hold `clip_project.*` only: the language model and the visual tower come from the
    modules passed at load time (inference.py:458-460); also the bare state_dict file of inference.py:469."""
    from transformers import GPT2Config
    import clipcap_b200 as cc
    fx = fixture("gpt2")
    lm = Holder(fx["sd_lm"], GPT2Config(vocab_size=fx["V"], n_positions=64, n_embd=128, n_layer=2, n_head=2))
    vis = Holder(fx["sd_vit"])
    sd = full_state_dict(fx, with_lm=False, with_vit=False)
    p1, p2 = str(tmp_path / "prefix_only.ckpt"), str(tmp_path / "bare.pt")
    torch.save({"state_dict": sd, "hyper_parameters": hparams_of(fx)}, p1)
    torch.save(sd, p2)
    ref = fx["logits_tf"]
    for path, hp in ((p1, None), (p2, hparams_of(fx))):
        model = cc.model.CLIPCaptionPrefixOnly.load_from_checkpoint(checkpoint_path=path, language_model=lm, visual_encoder=vis,
                                                                    strict=True, hparams=hp, vit_heads=2, max_images=8, max_ctx=32,
                                                                    page_tokens=4)
        logits = model(fx["tokens"], model.visual_encoder(fx["images"]), fx["mask"]).logits
        assert (logits.float().cpu() - ref).abs().max().item() <= TOL * ref.abs().max().item()
        model.engine.close()
    with pytest.raises(RuntimeError):   # a checkpoint without the language model and no module to take it from
        cc.model.CLIPCaptionModel.load_from_checkpoint(checkpoint_path=p1, lm_heads=2, lm_d=128, lm_layers=2, lm_vocab=fx["V"],
                                                       lm_n_pos=64, vit=False, max_images=8, max_ctx=32)

"""Error behaviour of the C ABI (a non-zero return + ccb_last_error, surfaced as RuntimeError by the ctypes
host; no exception crosses the ABI, nothing falls back, and the context stays usable after a rejected call)."""
import ctypes as C
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def loaded():
    import clipcap_b200 as cc
    fx = torch.load(os.path.join(GOLDEN, "tiny_gpt2.pt"), weights_only=False)
    cfg = cc.EngineConfig(
        lm_arch="gpt2", lm_d=fx["d"], lm_layers=2, lm_heads=fx["heads"], lm_vocab=fx["V"], lm_n_pos=64,
        map_dim_clip=fx["dim_clip"], map_clip_len=fx["CL"], map_prefix_len=fx["P"], map_heads=fx["map_heads"], map_layers=2,
        vit_image=fx["vit_image"], vit_patch=fx["vit_patch"], vit_width=fx["vit_width"], vit_layers=fx["vit_layers"],
        vit_heads=fx["vit_heads"], vit_out=fx["dim_clip"], max_images=4, max_beam=3, max_ctx=16, page_tokens=4)
    eng = cc.Engine(cfg)
    eng.load_state_dict(fx["sd_lm"], prefix="language_model.")
    eng.load_state_dict(fx["sd_mapper"], prefix="clip_project.")
    eng.load_state_dict(fx["sd_vit"], prefix="visual.")
    eng.check_weights()
    yield eng, fx
    eng.close()


def still_works(eng, fx):
    p = eng.gen_params("greedy", 10, stop_token=fx["stop_id"], max_stops=1)
    tokens, lengths, _ = eng.generate(fx["prefix"], p)
    tokens, lengths = tokens.cpu(), lengths.cpu()
    for i, want in enumerate(fx["greedy"]):
        assert tokens[i, :int(lengths[i])].tolist() == want


def test_generate_rejects_what_the_context_was_not_sized_for(loaded):
    eng, fx = loaded
    prefix = fx["prefix"]
    with pytest.raises(RuntimeError, match="max_images"):
        eng.generate(prefix.repeat(2, 1, 1)[:5], eng.gen_params("greedy", 4))
    with pytest.raises(RuntimeError, match="max_ctx"):
        eng.generate(prefix, eng.gen_params("greedy", 13))            # 4 prefix tokens + 13 > 16
    with pytest.raises(RuntimeError, match="beam"):
        eng.generate(prefix, eng.gen_params("beam", 4, beam_size=5))    # max_beam = 3
    with pytest.raises(RuntimeError):
        eng.generate(prefix, eng.gen_params("greedy", 0))
    still_works(eng, fx)


def test_bad_weights_are_reported_not_ignored(loaded):
    import clipcap_b200 as cc
    eng, fx = loaded
    fresh = cc.Engine(eng.cfg)
    try:
        with pytest.raises(RuntimeError, match="missing|not loaded|weight"):
            fresh.check_weights()                                         # nothing loaded yet
        bad = {"transformer.wte.weight": torch.zeros(7, fx["d"])}
        with pytest.raises(RuntimeError, match="expected"):
            fresh.load_state_dict(bad, prefix="language_model.")          # wrong shape
        unused = fresh.load_state_dict({"some.other.tensor": torch.zeros(3)}, prefix="language_model.")
        assert unused == ["some.other.tensor"]                            # unknown names are returned, not an error
    finally:
        fresh.close()


def test_null_and_out_of_range_arguments_at_the_abi(loaded):
    eng, fx = loaded
    lib = eng.lib
    out = torch.empty(4, dtype=torch.int32, device="cuda")
    logits = torch.randn(4, fx["V"], device="cuda")
    assert lib.ccb_argmax(eng._h, None, logits.stride(0), 4, fx["V"], C.c_void_p(out.data_ptr()), None) < 0
    assert b"null" in lib.ccb_last_error(eng._h)
    two = torch.empty(2, device="cuda")
    rl = torch.empty(4, device="cuda")
    tg = torch.zeros(4, dtype=torch.int32, device="cuda")
    assert lib.ccb_cross_entropy(eng._h, C.c_void_p(logits.data_ptr()), 3, 4, fx["V"], C.c_void_p(tg.data_ptr()), None, 0,
                                 C.c_void_p(rl.data_ptr()), C.c_void_p(two.data_ptr()), None) < 0          # ld < V
    with pytest.raises(ValueError):
        eng.cross_entropy(logits, torch.zeros(3, dtype=torch.int64))                                        # 4 rows, 3 targets
    assert lib.ccb_argmax(eng._h, C.c_void_p(logits.data_ptr()), logits.stride(0), 4, fx["V"], C.c_void_p(out.data_ptr()), None) == 0
    torch.cuda.synchronize()
    assert out.cpu().tolist() == logits.argmax(-1).cpu().tolist()
    still_works(eng, fx)


def test_lm_forward_rejects_too_many_tokens(loaded):
    eng, fx = loaded
    emb = torch.randn(4, 17, fx["d"])          # 17 positions > max_ctx = 16
    with pytest.raises(RuntimeError):
        eng.lm_forward(emb)
    still_works(eng, fx)

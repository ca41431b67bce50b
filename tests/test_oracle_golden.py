"""Pins oracle/clipcap_oracle.py against the fixtures produced by running the reference itself
(tools/make_golden.py -> tests/golden/*.pt).  CPU only."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import clipcap_oracle as orc  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def load(name):
    return torch.load(os.path.join(GOLDEN, name), weights_only=False)


def f32(sd):
    return {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}


def close(a, b, rel=2e-5):
    return (a - b).abs().max().item() <= rel * max(b.abs().max().item(), 1e-6)


@pytest.fixture(scope="module", params=["gpt2", "gptj"])
def fx(request):
    f = load("tiny_%s.pt" % request.param)
    f["lm"] = orc.OracleLM(f32(f["sd_lm"]), f["arch"], f["heads"], f["rotary_dim"])
    f["sd_mapper32"] = f32(f["sd_mapper"])
    f["sd_vit32"] = f32(f["sd_vit"])
    return f


def test_vit_matches_reference(fx):
    feat = orc.vit_forward(fx["sd_vit32"], fx["images"], fx["vit_heads"], fx["vit_patch"])
    assert close(feat, fx["feat"])


def test_mapper_matches_reference(fx):
    prefix = orc.mapper_forward(fx["sd_mapper32"], fx["feat"], fx["CL"], fx["map_heads"])
    assert close(prefix, fx["prefix"])


@pytest.mark.parametrize("act", ["elu", "gelu", "selu", "geglu"])
def test_mapper_activations_match_reference(act):
    """parse_act_fn's other activations (layers/Transformer.py:117-130), geglu with its 2 x hidden fc1 (:74, :112-114)."""
    ax = load("tiny_mapper_acts.pt")
    prefix = orc.mapper_forward(f32(ax["sd_" + act]), ax["feat"], ax["CL"], ax["map_heads"], act)
    assert close(prefix, ax["prefix_" + act])


def test_lm_call_matches_reference(fx):
    assert close(fx["lm"].logits(fx["prefix"]), fx["logits_prefix"], 5e-5)


def test_lm_cached_equals_full(fx):
    lm = fx["lm"]
    full = lm.logits(fx["prefix"])
    first, past = lm.forward(fx["prefix"][:, :3])
    rest, _ = lm.forward(fx["prefix"][:, 3:], past=past)
    assert close(torch.cat((first, rest), 1), full, 5e-5)


def test_caption_model_forward_matches_reference(fx):
    mapper = lambda feat: orc.mapper_forward(fx["sd_mapper32"], feat, fx["CL"], fx["map_heads"])
    logits = orc.caption_model_forward(fx["lm"], mapper, fx["tokens"], fx["feat"], fx["mask"])
    assert close(logits, fx["logits_tf"], 5e-5)


def test_caption_loss_matches_reference(fx):
    """model.py:204-211 / evaluate_model.py:505-514: tiny_loss.pt holds the reference's expression on the reference's logits."""
    want = load("tiny_loss.pt")[fx["arch"]]
    mapper = lambda feat: orc.mapper_forward(fx["sd_mapper32"], feat, fx["CL"], fx["map_heads"])
    loss, rows = orc.caption_loss(fx["lm"], mapper, fx["tokens"], fx["feat"], fx["mask"], fx["P"])
    assert abs(float(loss) - float(want["loss"])) <= 1e-5 * abs(float(want["loss"]))
    assert close(rows.reshape(-1), want["row_loss"], 5e-5)
    assert int((want["tokens"] == 0).sum()) >= 6        # the fixture exercises ignore_index


@pytest.mark.parametrize("use_cache", [False, True])
@pytest.mark.parametrize("key,beam,T,temp", [("greedy", 1, 10, 1.0), ("beam5", 5, 10, 1.0), ("beam3_T2", 3, 8, 2.0)])
def test_generate_beam_matches_reference(fx, key, beam, T, temp, use_cache):
    for i, want in enumerate(fx[key]):
        tokens, lens, scores, order = orc.generate_beam(fx["lm"], fx["prefix"][i:i + 1], beam, T, temp, fx["stop_id"], use_cache)
        best = int(order[0])
        assert tokens[best, :int(lens[best])].tolist() == want


def _replay(seed, V):
    g = torch.Generator().manual_seed(seed)
    # torch.multinomial draws q = empty_like(p).exponential_(1) from the same stream, one [V] tensor per call
    return lambda ci, step: torch.empty(V).exponential_(1, generator=g)


def test_generate_no_beam_inference_matches_reference(fx):
    for i, want in enumerate(fx["nobeam_inference"]):
        got = orc.generate_no_beam(fx["lm"], fx["prefix"][i:i + 1], [0.1 * k for k in range(1, 10)],
                                   _replay(fx["nobeam_inference_seeds"][i], fx["V"]), entry_length=8, stop_token=fx["stop_id"],
                                   repetition_penalty=1.2)
        assert got == want


def test_generate_no_beam_evaluate_matches_reference(fx):
    for i, want in enumerate(fx["nobeam_evaluate"]):
        got = orc.generate_no_beam(fx["lm"], fx["prefix"][i:i + 1], [0.3, 0.9], _replay(fx["nobeam_seed"] + 100 + i, fx["V"]),
                                   entry_length=8, stop_token=fx["stop_id"], repetition_penalty=1.2, max_stops=2,
                                   special_ids=[fx["V"] - 1], bos_token=fx["V"] - 1, use_cache=True)
        assert got == want


def clip_tokenize_ids(texts, ctx, sot, eot):
    """tools/make_golden.py's stand-in for clip.tokenize on id lists: [sot] + ids + [eot], truncated, zero-padded."""
    out = torch.zeros(len(texts), ctx, dtype=torch.int64)
    for i, t in enumerate(texts):
        ids = [sot] + [int(x) for x in t] + [eot]
        if len(ids) > ctx:
            ids = ids[:ctx]
            ids[-1] = eot
        out[i, :len(ids)] = torch.tensor(ids)
    return out


def test_generate_clip_guided_matches_reference():
    """evaluate_model.generate_clip_guided (:182-312) of the unmodified reference vs the oracle's restatement."""
    base, gx = load("tiny_gpt2.pt"), load("tiny_clip_guided.pt")
    lm = orc.OracleLM(f32(base["sd_lm"]), "gpt2", base["heads"], 0)
    sd_text = f32(gx["sd_text"])
    tokenize = lambda texts: clip_tokenize_ids(texts, gx["ctx"], gx["sot"], gx["eot"])
    encode = lambda tok: orc.clip_text_forward(sd_text, tok, gx["heads"])
    V = base["V"]
    for run in gx["runs"]:
        for i, want in enumerate(run["captions"]):
            got = orc.generate_clip_guided(lm, base["prefix"][i:i + 1], base["feat"][i:i + 1], tokenize, encode, V - 1, [V - 1], **run["kw"])
            assert got == want, (run["kw"], i, got, want)


def test_greedy_batched_equals_per_image(fx):
    toks, lens = orc.generate_greedy(fx["lm"], fx["prefix"], 10, fx["stop_id"])
    for i, want in enumerate(fx["greedy"]):
        assert toks[i, :int(lens[i])].tolist() == want


# ---------------------------------------------------------------------------------------------- logit processors
@pytest.fixture(scope="module")
def sx():
    return load("sampler.pt")


def same(a, b):
    """Equal kept sets and values, except that rows whose nucleus boundary falls inside a run of EXACTLY tied logits
    may keep different members of that run (torch.sort leaves their order unspecified): for such rows the number
    kept, every element above the boundary value and every element below it must still agree."""
    if a.dim() == 1:
        a, b = a[None], b[None]
    for ra, rb in zip(a, b):
        ka, kb = ~torch.isinf(ra), ~torch.isinf(rb)
        if torch.equal(ka, kb):
            if not torch.equal(ra[ka], rb[kb]):
                return False
            continue
        if ka.sum() != kb.sum():
            return False
        diff = ka ^ kb
        vals = torch.where(ka, ra, rb)[diff]
        if not bool((vals == vals[0]).all()):
            return False          # disagreement outside a single tied value
        boundary = vals[0]
        src = torch.where(ka, ra, torch.where(kb, rb, torch.full_like(ra, float("-inf"))))
        if not torch.equal(ka & (src > boundary), kb & (src > boundary)) or bool((ka & (src < boundary)).any()) \
                or bool((kb & (src < boundary)).any()):
            return False
    return True


def test_batch_filters_match_reference(sx):
    L = sx["logits"]
    assert same(orc.top_k_top_p_filtering_batch(L, 0, 0.9), sx["topp_0.9"])
    assert same(orc.top_k_top_p_filtering_batch(L, 0, 0.1), sx["topp_0.1"])
    assert same(orc.top_k_top_p_filtering_batch(L, 40, 0.0), sx["topk_40"])
    assert same(orc.top_k_top_p_filtering_batch(L, 0.05, 0.0), sx["topk_0.05"])
    assert same(orc.top_k_top_p_filtering_batch(L, 40, 0.5), sx["topk_40_topp_0.5"])
    assert same(orc.top_k_top_p_filtering_batch(L, sx["top_k_rows"].clone(), sx["top_p_rows"].clone()), sx["rows"])


def test_1d_filters_and_penalty_match_reference(sx):
    L = sx["logits"]
    for i in range(L.shape[0]):
        assert same(orc.top_k_top_p_filtering(L[i], 0, 0.8), sx["topp1d_0.8"][i])
        assert same(orc.top_k_top_p_filtering(L[i], 7, 0.0), sx["topk1d_7"][i])
    assert torch.equal(orc.repetition_penalty_apply(L, sx["history"], 1.2), sx["rep_1.2"])
    assert torch.equal(orc.repetition_penalty_apply(L[0], sx["history"][0], 1.2), sx["rep_1.2_inference_row0"])


def test_multinomial_identity(sx):
    probs = torch.softmax(sx["topp_0.9"], -1)
    g = torch.Generator().manual_seed(sx["multinomial_seed"])
    q1 = torch.empty_like(probs).exponential_(1, generator=g)
    q2 = torch.empty_like(probs).exponential_(1, generator=g)
    assert torch.equal(orc.multinomial_from_noise(probs, q1, 1), sx["multinomial_1"])
    assert torch.equal(orc.multinomial_from_noise(probs, q2, 2), sx["multinomial_2"])


def test_typical_filtering_matches_reference():
    """sampling.py:72-102 on raw logits, per-row budgets and behind the nucleus filter (order of sampling.generate)."""
    tx = torch.load(os.path.join(GOLDEN, "typical.pt"), weights_only=False)
    L = tx["logits"]
    for tp in (0.2, 0.5, 0.9):
        assert same(orc.typical_filtering(L, tp), tx["typ_%s" % tp]), tp
    assert same(orc.typical_filtering(L, tx["typ_rows_p"].clone()), tx["typ_rows"])
    assert same(orc.typical_filtering(orc.top_k_top_p_filtering_batch(L, 0, 0.9), 0.5), tx["topp_0.9_typ_0.5"])
    assert torch.equal(orc.typical_filtering(L, 0.0), L)          # disabled


def test_all_vit_features_path_matches_reference():
    """The fork's default configuration (use_all_vit_features): patched ViT forward (inference.py:421-444) ->
    TransformerMapperAllFeatures (layers/Transformer.py:164-203) -> CLIPCaptionModel.forward, against the unmodified
    reference modules."""
    fx = torch.load(os.path.join(GOLDEN, "tiny_allfeatures.pt"), weights_only=False)
    f32 = lambda sd: {k: (v.float() if v.is_floating_point() else v) for k, v in sd.items()}
    toks = orc.vit_forward(f32(fx["sd_vit"]), fx["images"], fx["vit_heads"], fx["vit_patch"], all_tokens=True)
    assert toks.shape == fx["vit_tokens"].shape == (3, fx["T"], fx["dim_clip"])
    assert (toks - fx["vit_tokens"]).abs().max().item() <= 1e-4 * fx["vit_tokens"].abs().max().item()
    lm = orc.OracleLM(f32(fx["sd_lm"]), "gpt2", fx["heads"])
    for key in ("pos", "nopos"):
        sd = f32(fx["sd_mapper_" + key])
        assert ("pos_embeddings" in sd) == (key == "pos")
        prefix = orc.mapper_all_forward(sd, fx["vit_tokens"], fx["map_heads"], "relu")
        assert (prefix - fx["prefix_" + key]).abs().max().item() <= 1e-4 * fx["prefix_" + key].abs().max().item()
        emb = torch.cat((prefix, lm.get_embedding_text(fx["cap_tokens"])), dim=1)
        logits = lm.logits(emb)
        ref = fx["forward_logits_" + key]
        assert (logits - ref).abs().max().item() <= 1e-3 * ref.abs().max().item()
        for i, want in enumerate(fx["beam5_" + key]):
            t, l, sc, order = orc.generate_beam(lm, fx["prefix_" + key][i:i + 1], beam_size=5, entry_length=10, stop_token=fx["stop_id"])
            best = t[order[0]][:int(l[order[0]])].tolist()
            assert best == want, (key, i, best, want)


def test_clip_text_tower_and_cos_sim_match_reference():
    """encode_text (HF CLIPTextModelWithProjection standing in for the un-installed OpenAI clip) + sampling.cos_sim."""
    fx = torch.load(os.path.join(GOLDEN, "tiny_clip_text.pt"), weights_only=False)
    sd = {k: v.float() for k, v in fx["sd_text"].items()}
    feats = orc.clip_text_forward(sd, fx["tokens"], fx["heads"])
    assert (feats - fx["text_features"]).abs().max().item() <= 1e-4 * fx["text_features"].abs().max().item()
    sims = orc.cos_sim(feats, fx["image_features"])
    assert (sims - fx["sims"]).abs().max().item() <= 1e-5


def test_unique_captions_collects_first_occurrences():
    """The id-level restatement of sampling.py:311-323 (`sample`'s dedupe of the generate groups)."""
    import clipcap_b200.sampling as S
    g1 = [torch.tensor([[7, 8, 1, 2, 99], [7, 8, 3, 4, 99]]), torch.tensor([2, 3]), torch.tensor([9, 9]), torch.tensor([0.5, 0.7]),
          torch.zeros(2, 3)]
    g2 = [torch.tensor([[7, 8, 1, 2, 99, 99], [7, 8, 5, 99, 99, 99]]), torch.tensor([2, 2]), torch.tensor([8, 8]), 0.9, torch.ones(2, 4)]
    caps, params, stats = S.unique_captions([g1, g2], num_prompt_tokens=2, special_ids=[99])
    assert caps == [(1, 2), (3, 4), (5,)]
    assert params == [[2, 9, 0.5], [3, 9, pytest.approx(0.7)], [2, 8, 0.9]]
    assert [s["tokens"] for s in stats] == [[1, 2, 99], [3, 4, 99], [5, 99, 99, 99]]
    assert S.unique_captions([g1, g2], 2, [99], unique=False) == ([], [], [])

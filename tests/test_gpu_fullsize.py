"""Full-size (BASELINE.json configs[1..2]: ViT-B/32 + 8-layer mapper + GPT2-XL) parity through size-independent
properties: the fp32 oracle cannot re-run a 1.5 B-parameter model in seconds, so at full size the CUDA path is checked
against itself across code paths that share no kernels, and against invariances the reference has by construction.

  * KV-cached decode (persistent decode kernel, paged cache) == teacher-forced full forward of [prefix || tokens]
    (persistent tcgen05 GEMMs + tensor-core prefill attention): the reference computes every token with the full
    forward (inference.py:97, :249), so both must select the same tokens.
  * captions do not depend on how the image batch is split (the reference loops image by image; this is also what makes
    the multi-GPU sharding of SURVEY section 8e exact), for greedy, beam and Philox-keyed sampling.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def xl():
    import clipcap_b200 as cc
    from clipcap_b200 import synthetic
    cfg = cc.EngineConfig(max_images=16, max_beam=5, max_ctx=80)   # defaults = config 2 (GPT2-XL, P = 40, clip_len = 40)
    eng = cc.Engine(cfg)
    sds = synthetic.load_synthetic(eng)
    del sds
    torch.cuda.empty_cache()
    images = synthetic.synthetic_images(16, cfg, device="cuda")
    yield eng, cfg, images
    eng.close()


def test_cached_decode_agrees_with_teacher_forced_forward(xl):
    eng, cfg, images = xl
    T = 12
    feat = eng.vit_encode(images)
    prefix = eng.map_prefix(feat)                                   # [16, 40, 1600]
    p = eng.gen_params("greedy", T, stop_token=-1, max_stops=0)
    tokens, lengths, _ = eng.generate(prefix, p)
    torch.cuda.synchronize()
    tokens = tokens.long()
    # full forward over [prefix || first T-1 generated tokens]: position P-1+t predicts token t
    emb = torch.cat([prefix, eng.embed_tokens(tokens[:, :T - 1])], dim=1)
    logits = eng.lm_forward(emb)
    P = prefix.shape[1]
    pred = logits[:, P - 1:P - 1 + T].float()
    top = pred.argmax(-1)
    agree = (top == tokens)
    # a disagreement must be a near-tie of the two candidates in the teacher-forced logits (bf16 noise of the
    # activations, BASELINE.md: random-init margins can be below it); everything after the first flip of a row differs
    # legitimately, so rows are compared up to their first disagreement
    for r in range(tokens.shape[0]):
        bad = (~agree[r]).nonzero()
        if bad.numel() == 0:
            continue
        t = int(bad[0])
        lg = pred[r, t]
        margin = (lg.max() - lg[tokens[r, t]]).item()
        scale = (lg.max() - lg.min()).item()
        assert margin <= 2e-2 * scale, (r, t, margin, scale)
    # every row is identical under the tie rule (asserted above: tests/test_gpu_config_parity.py states the rule); strictly,
    # >= 99 % of the individual decisions agree, and a free-running row survives without any near-tie flip 3 times out of 4
    assert agree.float().mean().item() >= 0.99 - 1e-9 or (~agree).sum().item() <= 2, agree.float().mean().item()
    first_flip_free = sum(int(agree[r].all()) for r in range(tokens.shape[0]))
    assert first_flip_free >= 0.75 * tokens.shape[0], first_flip_free
    assert agree[:, 0].float().mean().item() >= 0.99          # the first token comes from the same prefill in both


@pytest.mark.parametrize("mode,kw", [("greedy", {}), ("beam", {"beam_size": 5}), ("sample", {"top_p": 0.9, "seed": 7})])
def test_captions_do_not_depend_on_the_batch_split(xl, mode, kw):
    eng, cfg, images = xl
    T = 8
    ids = torch.arange(16, dtype=torch.int64)

    def run(sl):
        p = eng.gen_params(mode, T, stop_token=-1, max_stops=0, row_ids=ids[sl] if mode == "sample" else None, **kw)
        tok, ln, sc = eng.caption_images(images[sl], p)
        torch.cuda.synchronize()
        return tok.cpu(), (sc.cpu() if sc is not None else None)

    whole, whole_sc = run(slice(0, 16))
    parts = [run(slice(0, 5)), run(slice(5, 6)), run(slice(6, 16))]
    split = torch.cat([t for t, _ in parts], dim=0)
    same = (whole == split).flatten(1).all(dim=1)
    # GEMM tiles see different row counts (5 / 1 / 10 vs 16 rows: other MMA shapes and split-K orders), so fp32 sums
    # may differ in the last bits and flip a near-tie; identical rows must dominate
    if whole_sc is None:
        assert same.float().mean().item() >= 0.8, same
    else:
        # beam search: five hypotheses per image, any near-tie among 5 x 50 257 candidates per step may resolve differently
        # (measured 12-13 of 16 images identical with either prefill-attention kernel); what must hold is that a different
        # outcome is a TIE: the winning hypotheses' length-normalised scores agree within the bf16 tolerance
        split_sc = torch.cat([s for _, s in parts], dim=0)
        assert same.float().mean().item() >= 0.6, same
        assert (whole_sc[same] - split_sc[same]).abs().max().item() <= 1e-2
        best_w, best_s = whole_sc.max(-1).values, split_sc.max(-1).values
        assert ((best_w - best_s).abs() <= 2e-2 * best_w.abs()).all(), (best_w, best_s)


def test_sampling_is_keyed_by_global_image_id(xl):
    """Same images, same seed, different row ids -> different samples; same ids in another order -> same captions."""
    eng, cfg, images = xl
    T = 8
    ids = torch.arange(8, dtype=torch.int64)
    p0 = eng.gen_params("sample", T, stop_token=-1, max_stops=0, top_p=0.9, seed=11, row_ids=ids)
    a, _, _ = eng.caption_images(images[:8], p0)
    torch.cuda.synchronize()
    a = a.cpu()
    perm = torch.tensor([3, 1, 7, 0, 2, 6, 5, 4])
    p1 = eng.gen_params("sample", T, stop_token=-1, max_stops=0, top_p=0.9, seed=11, row_ids=ids[perm])
    b, _, _ = eng.caption_images(images[:8][perm.cuda()], p1)
    torch.cuda.synchronize()
    b = b.cpu()
    same = (a[perm] == b).all(dim=1)
    assert same.float().mean().item() >= 0.75, same
    p2 = eng.gen_params("sample", T, stop_token=-1, max_stops=0, top_p=0.9, seed=11, row_ids=ids + 1000)
    c, _, _ = eng.caption_images(images[:8], p2)
    torch.cuda.synchronize()
    assert (c.cpu() != a).any()


def test_caption_dataset_micro_batches(xl):
    """Engine.caption_dataset: beam search runs 256 // beam images per call (the persistent decode kernel's row limit) and
    the concatenated result equals the per-chunk calls; sampling is keyed by global image id across micro-batches."""
    eng, cfg, images = xl
    pb = eng.gen_params("beam", 6, stop_token=-1, max_stops=0, beam_size=5)
    assert eng.micro_batch_for(pb) == min(cfg.max_images, 51)
    tok, ln, sc = eng.caption_dataset(images, pb, micro_batch=7)
    torch.cuda.synchronize()
    assert tok.shape == (16, 5, 6) and sc.shape == (16, 5)
    ref = [eng.caption_images(images[lo:lo + 7], pb) for lo in (0, 7, 14)]
    torch.cuda.synchronize()
    assert torch.equal(tok.cpu(), torch.cat([r[0] for r in ref]).cpu())
    ps = eng.gen_params("sample", 6, stop_token=-1, max_stops=0, top_p=0.9, seed=5)
    a, _, _ = eng.caption_dataset(images, ps, micro_batch=16, first_row_id=100)
    b, _, _ = eng.caption_dataset(images, ps, micro_batch=4, first_row_id=100)
    torch.cuda.synchronize()
    assert ((a == b).all(dim=1)).float().mean().item() >= 0.8


def test_large_row_counts_are_deterministic():
    """Regression test of a shared-memory hazard in the persistent decode kernel: above ~170 rows the helper warps' early
    K/V prefetch landed inside the live activation ring (three 32 KB tiles) and corrupted c_attn inputs now and then, seen
    as run-to-run differences of beam / sampled captions at 200+ rows.  Same inputs must give the same captions."""
    import clipcap_b200 as cc
    from clipcap_b200 import synthetic
    cfg = cc.EngineConfig(max_images=255, max_beam=5, max_ctx=64)
    eng = cc.Engine(cfg)
    synthetic.load_synthetic(eng)
    torch.cuda.empty_cache()
    images = synthetic.synthetic_images(255, cfg, device="cuda")
    for mode, n, kw in (("beam", 51, {"beam_size": 5}), ("beam", 40, {"beam_size": 5}), ("sample", 255, {"top_p": 0.9, "seed": 1})):
        runs = []
        for _ in range(3):
            p = eng.gen_params(mode, 12, stop_token=-1, max_stops=0, **kw)
            tok, _, _ = eng.caption_images(images[:n], p)
            torch.cuda.synchronize()
            runs.append(tok.cpu())
        assert torch.equal(runs[0], runs[1]) and torch.equal(runs[0], runs[2]), (mode, n)
    eng.close()

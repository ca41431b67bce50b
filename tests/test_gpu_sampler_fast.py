"""The sampler's boundary-free nucleus path (sampler.cu nucleus_sample_fast: rank by exp(x - max) / q, test the best few
for membership in the nucleus) must return exactly the token of the radix-select path it short-cuts.  Asking for the
filtered logits forces the exact path, so the same call with and without `return_filtered` compares the two; the exact
path itself is pinned against the reference by tests/test_gpu_parity.py (test_filters_bit_exact, test_multinomial_bit_exact)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

V = 50257


def _logits(rows, scale, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(rows, V, generator=g) * scale).cuda()


@pytest.mark.parametrize("scale", [1.0, 3.0, 8.0])          # flat (nucleus ~ 90 % of the vocabulary) ... peaked (a handful)
@pytest.mark.parametrize("top_p", [0.9, 0.5, 0.999])
def test_fast_path_equals_exact_path_philox(tiny_engine, scale, top_p):
    eng = tiny_engine
    L = _logits(96, scale, int(scale * 10) + int(top_p * 1000))
    for step in (0, 5):
        p = eng.gen_params("sample", 1, top_p=top_p, temperature=1.0, seed=1234 + step)
        fast, _, _ = eng.sample(L, p, step=step)
        exact, filt, _ = eng.sample(L, p, step=step, return_filtered=True)
        assert torch.equal(fast, exact)
        # the sampled token lies inside the kept set
        assert bool(torch.isfinite(filt.gather(1, exact.long().view(-1, 1))).all())


def test_fast_path_equals_exact_path_given_noise_rows_and_temperature(tiny_engine):
    eng = tiny_engine
    rows = 64
    L = _logits(rows, 4.0, 99)
    g = torch.Generator().manual_seed(5)
    q = torch.empty(1, rows, V).exponential_(1, generator=g)
    tp = torch.rand(rows, generator=g) * 0.9 + 0.05
    hist = torch.randint(0, V, (rows, 12), generator=g, dtype=torch.int32)
    p = eng.gen_params("sample", 1, top_p=1.0, top_p_rows=tp, temperature=0.7, repetition_penalty=1.3, q_noise=q)
    fast, _, _ = eng.sample(L, p, history=hist, step=0)
    exact, _, _ = eng.sample(L, p, history=hist, step=0, return_filtered=True)
    assert torch.equal(fast, exact)


def test_fast_path_with_top_k_and_ties(tiny_engine):
    """top-k before the nucleus (the fast path runs on the top-k filtered row) and rows full of exactly equal logits
    (candidates tie: the fast path must hand over to the exact path, whose tie rule is lowest index first)."""
    eng = tiny_engine
    L = _logits(32, 2.0, 3)
    L[16:] = torch.round(L[16:])            # many exact ties, also at the top
    L[31] = 0.0                             # a constant row
    for k in (0, 50):
        p = eng.gen_params("sample", 1, top_p=0.8, top_k=k, seed=77)
        fast, _, _ = eng.sample(L, p, step=3)
        exact, _, _ = eng.sample(L, p, step=3, return_filtered=True)
        assert torch.equal(fast, exact)


def test_tiny_vocabulary(tiny_engine):
    eng = tiny_engine
    g = torch.Generator().manual_seed(11)
    for v in (1, 2, 3, 5, 37):
        L = torch.randn(8, v, generator=g).cuda()
        p = eng.gen_params("sample", 1, top_p=0.6, seed=v)
        fast, _, _ = eng.sample(L, p, step=1)
        exact, _, _ = eng.sample(L, p, step=1, return_filtered=True)
        assert torch.equal(fast, exact)

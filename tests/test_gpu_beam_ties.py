"""Tie rule of the beam step (inference.py:98-131 through torch.topk, made deterministic here: equal scores -> lowest flat
index r * V + token first) at full vocabulary, on rows built from a few well-separated logit levels so that the ranking does
not depend on the last bits of exp / log.  Covers both selection paths of beam_rows_kernel: the candidate list (few tokens at
the top) and the whole-row fallback (more than 128 equal tokens at the top: constant rows)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

V = 50257


def _expected(logits, scores, lens, stopped, beam, first):
    """(src row, token) of the `beam` best (value desc, flat index asc) per image, computed on the CPU in float64."""
    out = []
    N = scores.shape[0]
    for n in range(N):
        rows = 1 if first else beam
        vals = []
        for r in range(rows):
            lg = logits[n if first else n * beam + r].double()
            lp = torch.log_softmax(lg, -1)
            if first:
                vals.append(lp)
            elif stopped[n, r]:
                v = torch.full((V,), float("-inf"), dtype=torch.float64)
                v[0] = scores[n, r].double() / lens[n, r].double()
                vals.append(v)
            else:
                vals.append((scores[n, r].double() + lp) / (lens[n, r].double() + 1))
        flat = torch.cat(vals)
        order = torch.sort(-flat, stable=True).indices[:beam]
        out.append([(int(i) // V, int(i) % V) for i in order])
    return out


def _levels(rows, seed, top_count):
    """logits in {-4, -2, 0, 2} with exactly `top_count` tokens at level 4 (placed pseudo-randomly)"""
    g = torch.Generator().manual_seed(seed)
    lg = torch.randint(-2, 2, (rows, V), generator=g).float() * 2.0
    for r in range(rows):
        pos = torch.randperm(V, generator=g)[:top_count]
        lg[r, pos] = 4.0
    return lg


@pytest.mark.parametrize("top_count", [1, 3, 100, 200, V])
def test_first_step_ties(tiny_engine, top_count):
    eng = tiny_engine
    N, beam, T = 4, 5, 4
    lg = _levels(N, top_count, top_count)
    scores = torch.zeros(N, beam, device="cuda")
    seq = torch.ones(N, beam, device="cuda")
    stopped = torch.zeros(N, beam, dtype=torch.uint8, device="cuda")
    tokens = torch.zeros(N, beam, T, dtype=torch.int32, device="cuda")
    nxt, src = eng.beam_step(lg.cuda(), scores, seq, stopped, tokens, 0, beam, 1.0, -1)
    want = _expected(lg, scores.cpu(), seq.cpu(), stopped.cpu(), beam, True)
    got = nxt.view(N, beam).cpu().tolist()
    assert got == [[t for _, t in w] for w in want]


@pytest.mark.parametrize("top_count", [2, 100, 200])
def test_later_step_ties_across_rows_and_stopped_rows(tiny_engine, top_count):
    eng = tiny_engine
    N, beam, T = 2, 4, 6
    lg = _levels(N * beam, 1000 + top_count, top_count)
    lg[1] = lg[0]                                    # image 0: rows 0 and 1 are identical -> ties across rows
    scores = torch.tensor([[-1.0, -1.0, -3.0, -2.0], [-2.0, -0.5, -2.0, -4.0]], device="cuda")
    seq = torch.tensor([[2.0, 2.0, 2.0, 2.0], [2.0, 1.0, 2.0, 2.0]], device="cuda")
    stopped = torch.tensor([[0, 0, 0, 0], [0, 1, 0, 0]], dtype=torch.uint8, device="cuda")
    tokens = torch.zeros(N, beam, T, dtype=torch.int32, device="cuda")
    want = _expected(lg, scores.cpu(), seq.cpu(), stopped.cpu(), beam, False)
    nxt, src = eng.beam_step(lg.cuda(), scores, seq, stopped, tokens, 2, beam, 1.0, -1)
    got_tok = nxt.view(N, beam).cpu().tolist()
    got_src = (src.view(N, beam).cpu() - torch.arange(N).view(N, 1) * beam).tolist()
    assert got_tok == [[t for _, t in w] for w in want]
    assert got_src == [[r for r, _ in w] for w in want]

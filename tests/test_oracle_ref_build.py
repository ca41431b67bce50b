"""oracle/_ref (the reference byte-compiled by oracle/build_ref.py) is the reference: imported sourceless in a child process
(no /root/reference on sys.path), its own `inference.generate_beam` must reproduce the golden captions that
tools/make_golden.py recorded from the source tree.  Needs the reference tree to compile from -- skipped on the GPU box."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHILD = r"""
import os, sys, torch
sys.path.insert(0, os.path.join(%(root)r, "oracle"))
import ref_harness
assert ref_harness.kind() == "compiled" and ref_harness.available()
ref = ref_harness.load_reference()
assert ref.inference.__file__.endswith(".pyc") and not any(p.rstrip("/") == "/root/reference" for p in sys.path)
from transformers import GPT2Config
fx = torch.load(os.path.join(%(root)r, "tests", "golden", "tiny_gpt2.pt"), weights_only=False)
lm = ref.lms.GPT2(GPT2Config(vocab_size=fx["V"], n_positions=64, n_embd=fx["d"], n_layer=2, n_head=fx["heads"]))
missing, unexpected = lm.load_state_dict({k: v.float() for k, v in fx["sd_lm"].items()}, strict=False)
assert not unexpected, unexpected
lm.tie_weights()
lm.eval()
class Tok:
    def encode_text(self, text, *a, **k): return [fx["stop_id"]]
    def decode_tokens(self, tokens): return [int(t) for t in tokens]
class M: pass
m = M(); m.language_model = lm
for i, want in enumerate(fx["greedy"]):
    got = ref.inference.generate_beam(m, Tok(), fx["prefix"][i:i + 1].float(), beam_size=1, entry_length=10)[0]
    assert got == want, (i, got, want)
for i, want in enumerate(fx["beam5"]):
    got = ref.inference.generate_beam(m, Tok(), fx["prefix"][i:i + 1].float(), beam_size=5, entry_length=10)[0]
    assert got == want, (i, got, want)
print("ok")
"""


@pytest.mark.skipif(not os.path.isdir("/root/reference/layers"), reason="no reference tree to compile from")
def test_compiled_reference_reproduces_the_golden_captions(tmp_path):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    out = str(tmp_path / "_ref")
    n = build_ref.build_ref(out=out)
    assert n >= 10 and os.path.exists(os.path.join(out, "inference.pyc")) and os.path.exists(os.path.join(out, "lms", "GPT2.pyc"))
    assert not [f for _, _, fs in os.walk(out) for f in fs if f.endswith(".py")]      # bytecode only: no source text leaves the tree
    env = dict(os.environ, CLIPCAP_REFERENCE_ROOT=out, PYTHONDONTWRITEBYTECODE="1")
    # COMPILED_ROOT is what kind() compares with: point the harness at the temporary build
    child = CHILD % {"root": ROOT}
    child = child.replace('import ref_harness\n', 'import ref_harness\nref_harness.COMPILED_ROOT = os.path.abspath(os.environ["CLIPCAP_REFERENCE_ROOT"])\n', 1)
    r = subprocess.run([sys.executable, "-c", child], env=env, cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout[-2000:] + r.stderr[-3000:]

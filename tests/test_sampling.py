"""rwkvtts_b200.sampling against the reference's own functions (third_party/cosyvoice/utils/common.py:106-137) on CPU:
same generator state in, same token ids out, including the repetition branch of ras_sampling.  Build container only."""
import importlib.util
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/third_party/cosyvoice/utils/common.py"
pytestmark = pytest.mark.skipif(not os.path.exists(REF), reason="reference tree not mounted")
sys.path.insert(0, ROOT)


def _reference():
    import ast
    tree = ast.parse(open(REF).read())
    want = ("ras_sampling", "nucleus_sampling", "random_sampling")
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in want]
    ns = {"torch": torch}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


def test_same_tokens_as_reference_for_the_same_generator_state():
    from rwkvtts_b200 import sampling as S
    ref = _reference()
    g = torch.Generator().manual_seed(5)
    n_rep = 0
    for trial in range(300):
        V = int(torch.randint(3, 400, (1,), generator=g))
        sharp = float(torch.rand(1, generator=g)) * 6 + 0.2           # from nearly flat to very peaked distributions
        logp = (torch.randn(V, generator=g) * sharp).log_softmax(0)
        top_p, top_k = float(torch.rand(1, generator=g)) * 0.95 + 0.04, int(torch.randint(1, 40, (1,), generator=g))
        decoded = torch.randint(0, max(2, V // 8), (int(torch.randint(0, 15, (1,), generator=g)),), generator=g).tolist()
        if trial % 3 == 0 and decoded:
            decoded[-1] = int(logp.argmax())                           # make the repetition branch likely
        for name, args, kw in (("nucleus_sampling", (logp,), dict(top_p=top_p, top_k=top_k)),
                               ("random_sampling", (logp, decoded, 25), {}),
                               ("ras_sampling", (logp, decoded, 25), dict(top_p=top_p, top_k=top_k))):
            torch.manual_seed(1000 + trial)
            want = ref[name](*args, **kw)
            state_ref = torch.get_rng_state()
            torch.manual_seed(1000 + trial)
            got = getattr(S, name)(*args, **kw)
            assert got.shape == want.shape and got.dtype == want.dtype and torch.equal(got, want), (trial, name)
            assert torch.equal(torch.get_rng_state(), state_ref), (trial, name, "random stream consumed differently")
        if decoded and int((torch.tensor(decoded[-10:]) == int(logp.argmax())).sum()) >= 1:
            n_rep += 1
    assert n_rep > 20                                                  # the repetition branch was exercised


def test_sync_free_variant_picks_the_same_token():
    from rwkvtts_b200 import sampling as S
    g = torch.Generator().manual_seed(9)
    for trial in range(100):
        logp = (torch.randn(int(torch.randint(3, 300, (1,), generator=g)), generator=g) * 3).log_softmax(0)
        torch.manual_seed(trial)
        a = S.nucleus_sampling(logp, top_p=0.8, top_k=25)
        torch.manual_seed(trial)
        b = S.nucleus_sampling(logp, top_p=0.8, top_k=25, exact_stream=False)
        assert torch.equal(a, b), trial

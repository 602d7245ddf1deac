"""CPU timing of the batch builders against the reference's own functions, at the shapes
of config c2 / c5.  Needs /root/reference (build container only); on a GPU the gap is larger (every per-sample
embedding lookup / cat / .item() of the reference is a kernel launch or a sync there).  Output: profiles/r01_batch_builders_cpu.txt
"""
import os
import sys
import time
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch

import test_batch_builder as tb
from rwkvtts_b200 import batch as mine


def best(fn, n=3):
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return min(ts)


class Tok:
    def encode(self, text, add_special_tokens=False):
        return [(ord(c) * 7 + i) % 65536 for i, c in enumerate(text)]


def model(D):
    torch.manual_seed(0)
    m = types.SimpleNamespace(device=torch.device("cpu"))
    m.text_embedder = torch.nn.Embedding(65536, D)
    m.global_embedder = torch.nn.Embedding(4096, D)
    m.tts_tag_embedder = torch.nn.Embedding(3, D)
    m.model = types.SimpleNamespace(embeddings=torch.nn.Embedding(8193, D))
    return m


def main():
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    g = torch.Generator().manual_seed(42)
    rows = []
    # Spark layout, config c2: B = 8, text 128, 32 global, semantic fills T = 4096
    B, D = 8, 1024
    m = model(D)
    batch = {"text": ["x" * 128] * B,
             "global_tokens": [torch.randint(0, 4096, (32,), generator=g).tolist() for _ in range(B)],
             "semantic_tokens": [torch.randint(0, 8192, (3933 - 17 * i,), generator=g).tolist() for i in range(B)]}
    ref = tb._reference_fn()
    with torch.no_grad():
        a = best(lambda: ref(batch, Tok(), m, 8192, "cpu"))
        b = best(lambda: mine.create_inputs_and_labels(batch, Tok(), m, 8192, "cpu"))
    rows.append(("create_inputs_and_labels  B=8 T=4096 D=1024", a, b))
    # XY layout, config c5 per GPU: B = 2, T2 = 8192 - text, 8 channels
    feats = [{"json": {"text": "y" * 120}, "audio": {"array": np.zeros(4 * 8000, dtype=np.float32)}} for _ in range(2)]
    args = (tb._XYTextTok(), tb._XYCodec(), 8, 256, 40, "cpu")
    refxy = tb._reference_process_batch()
    a = best(lambda: refxy(feats, *args), n=2)
    b = best(lambda: mine.process_batch(feats, *args), n=2)
    rows.append(("process_batch (XY)  B=2 T2=8000 8 channels", a, b))
    # left-padded collator, B = 8
    lens = [(128, 32, 3933 - 17 * i) for i in range(B)]
    def left(L, ns, hi):
        ids = torch.randint(0, hi, (B, L), generator=g)
        mk = torch.zeros(B, L, dtype=torch.long)
        for i, n in enumerate(ns):
            mk[i, L - n:] = 1
        return ids * mk, mk
    it, mt = left(128, [l[0] for l in lens], 65536)
    ig, mg = left(32, [l[1] for l in lens], 4096)
    is_, ms = left(3933, [l[2] for l in lens], 8192)
    pb = {"input_ids": it, "attention_mask_input_ids": mt, "global_tokens_ids": ig, "global_tokens_attention_mask": mg,
          "semantic_tokens_ids": is_, "semantic_tokens_attention_mask": ms}
    refp = tb._reference_process_single_batch()
    with torch.no_grad():
        a = best(lambda: refp(pb, m, eos_token_id=8192))
        b = best(lambda: mine.process_single_batch(pb, m, eos_token_id=8192))
    rows.append(("process_single_batch  B=8 T=4096 D=1024", a, b))
    out = ["batch builders on CPU (%d threads), best of 3, seconds: reference function | this repo | ratio" % torch.get_num_threads()]
    for name, a, b in rows:
        out.append("%-48s %8.4f | %8.4f | %6.1fx" % (name, a, b, a / b))
    txt = "\n".join(out)
    print(txt)
    open(os.path.join(ROOT, "profiles", "r01_batch_builders_cpu.txt"), "w").write(txt + "\n")


if __name__ == "__main__":
    main()

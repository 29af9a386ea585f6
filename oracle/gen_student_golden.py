"""Golden logits of the reference's student model (models/lstm_seq2seq.py:LSTMSeq2SeqModel, imported
unmodified from /root/reference) for psketch_b200.students.Seq2SeqPolicy.  TEST INFRASTRUCTURE,
build container only.

    python -m oracle.gen_student_golden

Weights are a deterministic function of each parameter's reference name (no file of weights has to be
committed): tests/test_students.py rebuilds the same state dict, loads it through
Seq2SeqPolicy.load_reference_state_dict and must reproduce the logits stored here, for
(a) the ImitationStudent call pattern (2 reversed task tokens, no mask, time feature 0) and
(b) the PrimitiveLanguageStudent pattern (padded instruction words with a source mask, time = t).
"""
import os
import sys
import zlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def formula_state_dict(shapes):
    """{reference parameter name: shape} -> tensors, each from a generator seeded by its name."""
    sd = {}
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed(zlib.crc32(name.encode()))
        sd[name] = (torch.rand(tuple(shape), generator=g) - 0.5) * 0.2
    return sd


def main():
    ref_shim._install_shims()
    sys.path.insert(0, ref_shim.REF_ROOT)
    with ref_shim.reference_cwd():
        from misc.util import Struct
        from models.lstm_seq2seq import LSTMSeq2SeqModel
    cfg = Struct(vocab_size=28, word_embed_size=128, enc_hidden_size=256, dec_hidden_size=256,
                 hidden_size=256, pad_idx=2, dropout_ratio=0.0, device=torch.device("cpu"),
                 input_size=404, n_actions=6)
    model = LSTMSeq2SeqModel(cfg).eval()
    shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
    model.load_state_dict(formula_state_dict(shapes))
    S = np.load(os.path.join(OUT, "craft_medium_states.npz"))
    N, T = 16, 5
    feats = torch.from_numpy(S["features"][:N * T].astype(np.float32)).view(T, N, 404)
    rng = np.random.RandomState(3)
    out = {"shapes_names": np.asarray(list(shapes.keys())),
           "shapes": np.asarray([list(s) + [0] * (2 - len(s)) for s in shapes.values()], np.int64)}
    with torch.no_grad():
        # (a) students/imitation.py:54-84
        tok = torch.from_numpy(rng.randint(3, 28, size=(N, 2))).long()
        model.init(N, tok)
        la = [model.decode(feats[t], torch.zeros(N, dtype=torch.long)) for t in range(T)]
        out["a_tokens"], out["a_logits"] = tok.numpy(), torch.stack(la).numpy()
        # (b) students/primitive_language.py:44-93,149-168
        L = 9
        lens = rng.randint(2, L + 1, size=N)
        tokb = torch.full((N, L), 2, dtype=torch.long)
        mask = torch.ones((N, L), dtype=torch.bool)
        for i, n in enumerate(lens):
            tokb[i, :n] = torch.from_numpy(rng.randint(3, 28, size=n))
            mask[i, :n] = False
        model.init(N, tokb, src_mask=mask)
        lb = [model.decode(feats[t], torch.full((N,), t, dtype=torch.long)) for t in range(T)]
        out["b_tokens"], out["b_mask"], out["b_logits"] = tokb.numpy(), mask.numpy(), torch.stack(lb).numpy()
    path = os.path.join(OUT, "student_logits.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes", out["a_logits"].shape, float(np.abs(out["a_logits"]).mean()))


if __name__ == "__main__":
    main()

"""psketch_b200.students: the batched policy against the reference's own student
model, the sequence form against the stepwise form (values and gradients), the loss against the
reference's per-timestep CrossEntropyLoss(ignore_index=-1), and the graphed rollout on the GPU."""
import os
import zlib

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from conftest import GOLDEN  # noqa: E402


def _formula_state_dict(names, shapes):
    sd = {}
    for name, shape in zip(names, shapes):
        shape = tuple(int(v) for v in shape if v)
        g = torch.Generator().manual_seed(zlib.crc32(str(name).encode()))
        sd[str(name)] = (torch.rand(shape, generator=g) - 0.5) * 0.2
    return sd


def _policy():
    from psketch_b200.students import Seq2SeqPolicy
    return Seq2SeqPolicy(404, 6, vocab_size=28, pad_idx=2, hidden=256, word_embed=128)


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(GOLDEN, "student_logits.npz"))


def test_policy_reproduces_the_reference_model(golden, medium_states):
    """Logits of models/lstm_seq2seq.py:LSTMSeq2SeqModel (oracle/gen_student_golden.py) on the same
    weights: ImitationStudent's call pattern and PrimitiveLanguageStudent's (mask, time = t).
    Floating point: fp32 on the CPU, same operations in the same order — tolerance 1e-5."""
    pol = _policy().eval()
    sd = _formula_state_dict(golden["shapes_names"], golden["shapes"])
    pol.load_reference_state_dict(sd)
    back = pol.to_reference_state_dict()
    assert set(back) == set(sd) and all(torch.equal(back[k], sd[k]) for k in sd if k != "encoder.embedding.weight")
    N, T = 16, 5
    feats = torch.from_numpy(medium_states["features"][:N * T].astype(np.float32)).view(T, N, 404)
    with torch.no_grad():
        mem = pol.encode(torch.from_numpy(golden["a_tokens"]))
        state, got = (mem["h0"], mem["c0"]), []
        for t in range(T):
            lg, state = pol.decode_step(feats[t], torch.zeros(N, dtype=torch.long), state, mem)
            got.append(lg)
        assert np.allclose(torch.stack(got).numpy(), golden["a_logits"], atol=1e-5, rtol=0)
        seq = pol.decode_sequence(feats, torch.zeros((T, N), dtype=torch.long), mem)
        assert np.allclose(seq.numpy(), golden["a_logits"], atol=1e-5, rtol=0)
        mem = pol.encode(torch.from_numpy(golden["b_tokens"]), mask=torch.from_numpy(golden["b_mask"]))
        tt = torch.arange(T).unsqueeze(1).expand(T, N)
        seq = pol.decode_sequence(feats, tt, mem)
        assert np.allclose(seq.numpy(), golden["b_logits"], atol=1e-5, rtol=0)


def test_sequence_form_has_the_stepwise_gradients(medium_states):
    """The learner decodes a recorded window with one LSTM call (decode_sequence).  Its gradients must
    be those of the stepwise decode on per-step copies of the features — in particular the gradient of
    the input weights, which silently goes wrong when the steps alias one feature buffer that is
    overwritten in place (examples/train_dagger.py)."""
    from psketch_b200.students import imitation_loss
    torch.manual_seed(0)
    pol = _policy()
    N, T = 12, 7
    feats = torch.from_numpy(medium_states["features"][:N * T].astype(np.float32)).view(T, N, 404)
    tok = torch.randint(3, 28, (N, 2))
    refs = torch.randint(-1, 6, (T, N))
    refs[T - 1] = -1                                       # a timestep where every env is done

    def grads(fn):
        pol.zero_grad()
        loss, shown = fn()
        loss.backward()
        return float(loss), {k: v.grad.clone() for k, v in pol.named_parameters() if v.grad is not None}

    def sequence():
        mem = pol.encode(tok)
        return imitation_loss(pol.decode_sequence(feats, torch.zeros((T, N), dtype=torch.long), mem), refs)

    def stepwise():
        mem = pol.encode(tok)
        state, total, used = (mem["h0"], mem["c0"]), 0.0, 0
        buf = torch.empty((N, 404))
        for t in range(T):
            buf.copy_(feats[t])                            # one buffer, rewritten every step ...
            lg, state = pol.decode_step(buf.clone(), torch.zeros(N, dtype=torch.long), state, mem)   # ... cloned
            if (refs[t] >= 0).any():                       # students/imitation.py:86-98
                total = total + torch.nn.functional.cross_entropy(lg, refs[t], ignore_index=-1)
                used += 1
        return total, total / used

    la, ga = grads(sequence)
    lb, gb = grads(stepwise)
    assert abs(la - lb) < 1e-4 * max(1.0, abs(lb))
    assert set(ga) == set(gb)
    for k in ga:
        assert torch.allclose(ga[k], gb[k], atol=2e-5, rtol=1e-4), k
    assert float(ga["dec.weight_ih_l0"].abs().sum()) > 0


@pytest.mark.gpu
def test_graphed_rollout_matches_the_trainer_protocol(splits, medium_tables, medium_oracle):
    """GraphedRollout (one CUDA graph: features -> decode -> sample -> teacher -> step, 40 timesteps)
    against the trainers' loop run step by step on the CPU oracle with the SAME sampled actions:
    recorded features, teacher labels, executed actions, success and counters."""
    from psketch_b200.students import GraphedRollout, Seq2SeqPolicy, task_tokens
    from psketch_b200.vec import VecCraft
    torch.manual_seed(5)
    n = 1500
    rng = np.random.RandomState(2)
    idx = rng.randint(0, 2200, size=n)
    grids = splits["dev_grids"]
    ienv, ipos, itask = splits["dev_inst_env"][idx], splits["dev_inst_pos"][idx], splits["dev_inst_task"][idx]
    env = VecCraft.from_instances(medium_tables, grids, ienv, ipos, itask, max_timesteps=255)
    pol = Seq2SeqPolicy(404, 6, len(medium_tables.task_manager.vocab) + 1, 2).to(env.device)
    for greedy in (False, True):
        roll = GraphedRollout(env, pol, max_timesteps=40, greedy=greedy)
        with torch.no_grad():
            mem = pol.encode(task_tokens(medium_tables, env.task))
        for rep in range(2):                               # second run = a pure graph replay
            roll.run(mem)
            torch.cuda.synchronize()
            acts, refs = roll.acts.cpu().numpy(), roll.refs.cpu().numpy()
            feats = roll.feats.cpu().numpy()
            o = medium_oracle
            grid = grids[ienv.astype(np.int64)].copy()
            inv = np.zeros((n, 21), np.int32)
            pos, dirs = ipos.astype(np.int32).copy(), np.zeros(n, np.int32)
            task = itask.astype(np.int32)
            done = np.zeros(n, bool)
            success = np.zeros(n, bool)
            steps = inter = 0
            for t in range(40):
                live = ~done
                # (finished envs idle at their start state: what they show is masked by refs == -1)
                assert np.array_equal(feats[t][live], o.features(grid, inv, pos, dirs)[live]), t
                ref = o.expert(grid, inv, pos, dirs, task)[0]
                assert np.array_equal(refs[t][live], ref[live]) and (refs[t][done] == -1).all(), t
                assert (acts[t][done] == 255).all() and (acts[t][live] < 6).all(), t
                inter += int(live.sum())
                a = np.where(live, acts[t], 5).astype(np.int32)
                ends = live & ((a == 5) | (t == 39))
                success |= ends & (o.satisfies(grid, inv, pos, dirs, task) == 1)
                done |= ends
                g2, i2, p2, d2, _ = o.step(grid, inv, pos, dirs, a)
                act_mask = ~done
                grid = np.where(act_mask[:, None], g2, grid)
                inv = np.where(act_mask[:, None], i2, inv)
                pos = np.where(act_mask[:, None], p2, pos)
                dirs = np.where(act_mask, d2, dirs)
                steps += int(act_mask.sum())
            assert done.all()
            assert np.array_equal(roll.success.cpu().numpy(), success)
            assert int(roll.steps) == steps and int(roll.interactions) == inter
            if greedy and rep == 1:
                first = acts.copy()
        if greedy:                                         # greedy decoding is deterministic
            roll.run(mem)
            assert np.array_equal(roll.acts.cpu().numpy(), first)
    env.check_errors()

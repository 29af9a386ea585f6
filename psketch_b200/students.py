"""A PyTorch LSTM student that consumes the device-resident feature tensor directly
(no list-of-ndarrays -> torch.tensor -> H2D copy per timestep as in
students/imitation.py:71-84).  Architecture in the spirit of models/lstm_seq2seq.py: the two task
tokens are embedded and encoded by an LSTM; the decoder LSTM cell reads the 404 state features
and attends over the encoder outputs; a linear layer gives the 6 action logits."""
import torch
import torch.nn as nn
import torch.nn.functional as F


class LSTMStudent(nn.Module):
    def __init__(self, n_features, n_actions, vocab_size, hidden=256, embed=128):
        super().__init__()
        self.embed = nn.Embedding(vocab_size, embed)
        self.encoder = nn.LSTM(embed, hidden, batch_first=True)
        self.inp = nn.Linear(n_features, hidden)
        self.cell = nn.LSTMCell(hidden, hidden)
        self.attn = nn.Linear(hidden, hidden, bias=False)
        self.out = nn.Linear(2 * hidden, n_actions)
        self.hidden = hidden

    def encode(self, task_tokens):
        """task_tokens i64[N,2] (TaskManager encodings) -> (memory [N,2,H], (h0, c0))."""
        mem, (h, c) = self.encoder(self.embed(task_tokens))
        return mem, (h[0], c[0])

    def initial_state(self, task_tokens):
        mem, (h, c) = self.encode(task_tokens)
        return {"mem": mem, "h": h, "c": c, "h0": h, "c0": c}

    def step(self, state, features, reset=None):
        """One decoding step; ``reset`` (bool[N]) restarts the decoder of envs whose episode ended."""
        h, c = state["h"], state["c"]
        if reset is not None:
            r = reset.unsqueeze(1)
            h = torch.where(r, state["h0"], h)
            c = torch.where(r, state["c0"], c)
        h, c = self.cell(F.relu(self.inp(features)), (h, c))
        score = torch.bmm(state["mem"], self.attn(h).unsqueeze(2)).squeeze(2)
        ctx = torch.bmm(F.softmax(score, dim=1).unsqueeze(1), state["mem"]).squeeze(1)
        logits = self.out(torch.cat([h, ctx], dim=1))
        state = dict(state, h=h, c=c)
        return logits, state


def task_tokens(tables, task_ids):
    """u8/int tensor of task ids -> i64[N,2] vocabulary encodings (data/task.py:65-66)."""
    tm = tables.task_manager
    table = torch.zeros((len(tm.tasks), 2), dtype=torch.long)
    for t in tm.tasks:
        table[t.task_id] = torch.tensor(t.encoding)
    return table.to(task_ids.device)[task_ids.long()]

"""CPU, build container only: the oracle restatement against the UNMODIFIED reference modules run live from
/root/reference on seeds and sizes the committed fixtures do not contain (the fixtures pin it where the reference
cannot travel; this pins it where it can).  Skipped when /root/reference is absent (the GPU box)."""
import importlib.util
import os
import sys

import pytest
import torch
import torch.nn as nn

from helpers import rel_err
from oracle import showtell_oracle as O

REF = os.environ.get("SHOWTELL_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "rnn.py")), reason="reference sources not present")


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.path.insert(0, REF)          # rnn.py does `from cnn import ResNet` (class import only)
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.path.pop(0)
    return mod


def _captions(gen, B, T, V, lengths):
    cap = torch.zeros(B, T, dtype=torch.int64)
    for b, l in enumerate(lengths):
        cap[b, 0] = 1
        if l > 2:
            cap[b, 1:l - 1] = torch.randint(4, V, (l - 2,), generator=gen)
        if l > 1:
            cap[b, l - 1] = 2
    return cap


@pytest.mark.parametrize("kind,rel,L,seed", [("gru", "rnn.py", 1, 101), ("gru", "rnn.py", 3, 102),
                                             ("lstm", "LSTM/rnn_lstm.py", 2, 103)])
def test_base_decoder_live(kind, rel, L, seed):
    mod = _load(f"ref_{kind}_{seed}", rel)
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    E, H, V, B, T = 20, 28, 53, 7, 8
    lengths = [8, 8, 6, 5, 5, 3, 2]
    net = mod.RNN(E, H, V, L)
    feat = torch.randn(B, E, generator=gen)
    cap = _captions(gen, B, T, V, lengths)
    logits = net(feat, cap, lengths)                                                       # main.py:148
    target = nn.utils.rnn.pack_padded_sequence(cap, lengths, batch_first=True)[0]         # main.py:145
    loss = nn.CrossEntropyLoss()(logits, target)                                           # main.py:149
    loss.backward()
    p = {k: v.detach().clone() for k, v in net.state_dict().items()}
    loss_o, grads_o, ex = O.train_step(p, kind, feat, cap, lengths)
    assert rel_err(ex["logits"], logits.detach()) < 2e-6
    assert abs(float(loss_o) - float(loss)) < 1e-5
    for n, q in net.named_parameters():
        assert rel_err(grads_o[n], q.grad) < 2e-5, n
    with torch.no_grad():
        assert torch.equal(O.rnn_greedy(p, kind, feat), net.sentence_index(feat))          # rnn.py:44-58
        if kind == "gru":                                                                  # rnn.py:60-108, batch of one
            for K in (2, 4):
                want = net.sentence_index(feat[:1], beam_size=K)
                assert torch.equal(O.rnn_beam_chain(p, feat[:1], K, 25).reshape(-1), torch.as_tensor(want).reshape(-1))


@pytest.mark.parametrize("kind,rel,L,seed,alpha_c", [("attn_gru", "Attention/rnn_attn.py", 1, 201, 1.0),
                                                     ("attn_lstm", "Attention/rnn_attn_LSTM.py", 2, 202, 0.3)])
def test_attention_decoder_live(kind, rel, L, seed, alpha_c, monkeypatch):
    monkeypatch.setattr(torch.Tensor, "cuda", lambda self, *a, **k: self)     # the files hard-code .cuda() (rnn_attn.py:64)
    mod = _load(f"ref_{kind}_{seed}", rel)
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    E, C, A, H, V, P, B, T = 12, 18, 10, 20, 41, 6, 5, 7
    lengths = [7, 6, 6, 4, 2]
    net = mod.RNN_Attn(E, C, A, H, V, L)
    feat = torch.relu(torch.randn(B, C, P, generator=gen))
    cap = _captions(gen, B, T, V, lengths)
    logits, alphas = net(feat, cap, lengths)                                               # main_attn.py:129
    target = nn.utils.rnn.pack_padded_sequence(cap, lengths, batch_first=True)[0]         # main_attn.py:126
    loss = nn.CrossEntropyLoss()(logits, target) + alpha_c * ((1 - alphas.sum(dim=1)) ** 2).mean()   # :130-131
    loss.backward()
    p = {k: v.detach().clone() for k, v in net.state_dict().items()}
    loss_o, grads_o, ex = O.train_step(p, kind, feat, cap, lengths, alpha_c=alpha_c)
    assert abs(float(loss_o) - float(loss)) < 1e-5
    for n, q in net.named_parameters():
        if q.grad is None:
            continue
        if n == "attn.full_att.bias":       # identically zero up to rounding in both (softmax shift invariance)
            assert float((grads_o[n] - q.grad).abs().max()) < 1e-6
        else:
            assert rel_err(grads_o[n], q.grad) < 5e-5, n

"""CPU: pin the oracle restatement against fixtures produced by the unmodified reference
(tests/golden/make_golden.py).  This is the 'is the checker right' gate."""
import numpy as np
import pytest
import torch

from helpers import ATTN_CASES, BASE_CASES, golden_grads, golden_params, load_golden, rel_err
from oracle import showtell_oracle as O

TOL = 2e-6   # fp32 re-association noise between the explicit equations and torch's fused CPU cells


@pytest.mark.parametrize("name", BASE_CASES)
def test_base_forward_loss_grads(name):
    g = load_golden(name)
    kind = str(g["kind"])
    p = golden_params(g)
    feat, cap = torch.from_numpy(g["cnn_feature"]), torch.from_numpy(g["caption"])
    lengths = g["lengths"].tolist()
    loss, grads, ex = O.train_step(p, kind, feat, cap, lengths)
    assert ex["logits"].shape == g["logits"].shape
    assert rel_err(ex["logits"], g["logits"]) < TOL
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    gg = golden_grads(g)
    for k, v in gg.items():
        assert rel_err(grads[k], v) < 2e-5, k


@pytest.mark.parametrize("name", BASE_CASES)
def test_base_greedy(name):
    g = load_golden(name)
    p = golden_params(g)
    feat = torch.from_numpy(g["cnn_feature"])
    with torch.no_grad():
        tok = O.rnn_greedy(p, str(g["kind"]), feat)
        tok1 = O.rnn_greedy(p, str(g["kind"]), feat[:1])
    assert tok.shape == (feat.shape[0], 25) and tok.dtype == torch.int64
    assert np.array_equal(tok.numpy(), g["greedy"])
    assert tok1.shape == (25,)                       # squeeze quirk at B=1 (rnn.py:56)
    assert np.array_equal(tok1.numpy(), g["greedy_b1"])


@pytest.mark.parametrize("name,K", [("gru_l1", 1), ("gru_l1", 3), ("gru_l1", 5), ("gru_tiny", 2),
                                    ("gru_l2", 3), ("gru_med", 3)])
def test_beam_chain(name, K):
    """Exact where the ranking is well separated; rows whose ranking is decided by rounding noise
    (see rnn_beam_chain's docstring) are excluded and counted."""
    g = load_golden(name)
    p = golden_params(g)
    feat = torch.from_numpy(g["cnn_feature"])
    same = 0
    with torch.no_grad():
        for i in range(feat.shape[0]):
            seq, trace = O.rnn_beam_chain(p, feat[i:i + 1], K, return_trace=True)
            ok = np.array_equal(seq.numpy(), g[f"beam_chain_k{K}"][i])
            same += int(ok)
            # a mismatch is only tolerated when some round's ranking hung on < 1e-5 of logit
            assert ok or O.beam_chain_margin(trace, K, 1e-5, ties_ok=True) < len(trace), i
            if K == 1:                               # rnn.py:43 comment: beam 1 == greedy
                assert np.array_equal(seq.numpy(), g["greedy"][i])
    assert same >= feat.shape[0] - 1


def test_beam_chain_rejects_batches():
    g = load_golden("gru_l1")
    with pytest.raises(ValueError):
        O.rnn_beam_chain(golden_params(g), torch.from_numpy(g["cnn_feature"]), 3)


def test_beam_tree():
    g = load_golden("gru_tiny")
    t = load_golden("tree_beam_gru_tiny")
    K, max_length, end_id, rows, _ = t["meta"].tolist()
    p = golden_params(g)
    feat = torch.from_numpy(g["cnn_feature"])
    total = 0
    for i in range(rows):
        init_fn, gen_fn = O.gru_tree_callbacks(p, feat[i])
        hyps = O.beam_search_tree(init_fn, gen_fn, [0], 1, end_id, beam_width=K, num_hypotheses=K,
                                  max_length=max_length)
        assert len(hyps) == int(t[f"row{i}.n"])
        for j, h in enumerate(hyps):
            assert h.to_sequence_of_values() == t[f"row{i}.hyp{j}.seq"].tolist()
            assert abs(h.cum_cost - float(t[f"row{i}.hyp{j}.cost"])) < 1e-4
            total += 1
    assert total > 0


@pytest.mark.parametrize("name", ATTN_CASES)
def test_attn_forward_loss_grads(name):
    g = load_golden(name)
    kind = str(g["kind"])
    p = golden_params(g)
    feat, cap = torch.from_numpy(g["cnn_feature"]), torch.from_numpy(g["caption"])
    lengths = g["lengths"].tolist()
    loss, grads, ex = O.train_step(p, kind, feat, cap, lengths, alpha_c=float(g["alpha_c"]))
    assert rel_err(ex["logits"], g["logits"]) < TOL
    assert rel_err(ex["alphas"], g["alphas"]) < TOL
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    for k, v in golden_grads(g).items():
        if k == "attn.full_att.bias":                # == 0 up to rounding (softmax shift invariance)
            assert float(grads[k].abs().max()) < 1e-6 and float(v.abs().max()) < 1e-6
            continue
        assert rel_err(grads[k], v) < 5e-5, k


@pytest.mark.parametrize("name", ATTN_CASES)
def test_attn_greedy(name):
    g = load_golden(name)
    p = golden_params(g)
    with torch.no_grad():
        tok = O.attn_greedy(p, str(g["kind"]).split("_")[1], torch.from_numpy(g["cnn_feature"]))
    assert np.array_equal(tok.numpy(), g["greedy"])


def test_float64_oracle_matches_float32_goldens():
    """The fp64 instantiation is what the CUDA backward is checked against at tight tolerance."""
    g = load_golden("attn_lstm_l2")
    p = golden_params(g, torch.float64)
    loss, grads, _ = O.train_step(p, "attn_lstm", torch.from_numpy(g["cnn_feature"]).double(),
                                  torch.from_numpy(g["caption"]), g["lengths"].tolist(),
                                  alpha_c=float(g["alpha_c"]))
    assert abs(float(loss) - float(g["loss"])) < 1e-5
    assert rel_err(grads["embed.weight"], g["grad.embed.weight"]) < 5e-5


def test_unsorted_lengths_raise():
    with pytest.raises(RuntimeError):
        O.batch_sizes_from_lengths([3, 5, 2])

"""Generate tests/golden/*.npz by running the UNMODIFIED reference modules.

Run in the build container only (needs /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

It imports rnn.py, LSTM/rnn_lstm.py, Attention/rnn_attn.py, Attention/rnn_attn_LSTM.py and
beam_search.py from /root/reference, instantiates them under a fixed seed at small odd sizes,
drives them exactly like main.py:145-151 / main_attn.py:126-133 / utils.py:194 do, and stores
weights (state_dict), inputs and outputs.  Nothing from the reference is copied into the repo;
only these numeric fixtures are committed.  The attention files hard-code ``.cuda()``
(rnn_attn.py:64,65,128), so on this CPU-only box ``Tensor.cuda`` is shimmed to the identity.
"""
import importlib.util
import os
import sys

import numpy as np
import torch
import torch.nn as nn

REF = os.environ.get("SHOWTELL_REFERENCE", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.path.insert(0, REF)          # rnn.py does `from cnn import ResNet` (class import only)
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.path.pop(0)
    return mod


def _np(sd):
    return {k: v.detach().cpu().numpy() for k, v in sd.items()}


def _captions(gen, B, T, V, lengths):
    cap = torch.zeros(B, T, dtype=torch.int64)
    for b, l in enumerate(lengths):
        cap[b, 0] = 1
        if l > 2:
            cap[b, 1:l - 1] = torch.randint(4, V, (l - 2,), generator=gen)
        if l > 1:
            cap[b, l - 1] = 2
    return cap


def base_case(mod, kind, name, E, H, V, L, B, T, lengths, seed, beams=()):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    net = mod.RNN(E, H, V, L)
    feat = torch.randn(B, E, generator=gen)
    cap = _captions(gen, B, T, V, lengths)
    featg = feat.clone().requires_grad_(True)
    logits = net(featg, cap, lengths)                                     # main.py:148
    target = nn.utils.rnn.pack_padded_sequence(cap, lengths, batch_first=True)[0]  # main.py:145
    loss = nn.CrossEntropyLoss()(logits, target)                          # main.py:149
    loss.backward()
    out = {"kind": kind, "dims": np.array([E, H, V, L, B, T]), "lengths": np.array(lengths),
           "cnn_feature": feat.numpy(), "caption": cap.numpy(), "logits": logits.detach().numpy(),
           "loss": loss.detach().numpy(), "grad.cnn_feature": featg.grad.numpy()}
    for k, v in _np(net.state_dict()).items():
        out["param." + k] = v
    for k, v in net.named_parameters():
        out["grad." + k] = v.grad.numpy()
    with torch.no_grad():
        out["greedy"] = net.sentence_index(feat).numpy()                  # utils.py:194, beam 0
        out["greedy_b1"] = net.sentence_index(feat[:1]).numpy()           # squeeze quirk -> (25,)
        for K in beams:
            rows = [net.sentence_index(feat[i:i + 1], beam_size=K).numpy() for i in range(B)]
            out[f"beam_chain_k{K}"] = np.stack(rows)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", float(loss.detach()))
    return net, feat


def attn_case(mod, kind, name, E, C, A, H, V, L, P, B, T, lengths, alpha_c, seed):
    torch.manual_seed(seed)
    gen = torch.Generator().manual_seed(seed + 1)
    net = mod.RNN_Attn(E, C, A, H, V, L)
    feat = torch.relu(torch.randn(B, C, P, generator=gen))                # post-ReLU grid, cnn_attn.py:49
    cap = _captions(gen, B, T, V, lengths)
    logits, alphas = net(feat, cap, lengths)                              # main_attn.py:129
    target = nn.utils.rnn.pack_padded_sequence(cap, lengths, batch_first=True)[0]
    loss_ce = nn.CrossEntropyLoss()(logits, target)                       # main_attn.py:130
    loss = loss_ce + alpha_c * ((1. - alphas.sum(dim=1)) ** 2).mean()     # main_attn.py:131
    loss.backward()
    out = {"kind": kind, "dims": np.array([E, C, A, H, V, L, P, B, T]), "lengths": np.array(lengths),
           "alpha_c": np.array(alpha_c), "cnn_feature": feat.numpy(), "caption": cap.numpy(),
           "logits": logits.detach().numpy(), "alphas": alphas.detach().numpy(),
           "loss": loss.detach().numpy(), "loss_ce": loss_ce.detach().numpy()}
    for k, v in _np(net.state_dict()).items():
        out["param." + k] = v
    for k, v in net.named_parameters():
        out["grad." + k] = v.grad.numpy()
    with torch.no_grad():
        vocab = lambda w: {"<pad>": 0, "<start>": 1, "<end>": 2, "<unk>": 3}[w]   # vocab_builder.py:68-69
        out["greedy"] = net.sentence_index(feat, vocab).numpy()           # main_attn.py:189
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(name, "loss", float(loss.detach()))


def tree_beam_case(bs_mod, net, feat, name, K, max_length, end_id):
    """Drive the reference beam_search() (beam_search.py:45) with a GRU adapter: state after the
    image step is the initial state; generate = embedding -> GRU step -> softmax(linear)."""
    H = net.unit.hidden_size
    res = {}
    with torch.no_grad():
        for i in range(feat.shape[0]):
            def init_fn(_X, i=i):
                _, h = net.unit(feat[i:i + 1].unsqueeze(1), None)
                return h[0].numpy()

            def gen_fn(_X, y_prev, s_prev):
                x = net.embeddings(torch.from_numpy(y_prev.astype(np.int64))).unsqueeze(1)
                o, h = net.unit(x, torch.from_numpy(s_prev).unsqueeze(0))
                prob = torch.softmax(net.linear(o.squeeze(1)), dim=1)
                return h[0].numpy(), prob.numpy(), [None] * len(y_prev)

            hyps = bs_mod.beam_search(init_fn, gen_fn, [0], 1, end_id, beam_width=K,
                                      num_hypotheses=K, max_length=max_length)
            res[f"row{i}.n"] = np.array(len(hyps))
            for j, h in enumerate(hyps):
                res[f"row{i}.hyp{j}.seq"] = np.array(h.to_sequence_of_values(), dtype=np.int64)
                res[f"row{i}.hyp{j}.cost"] = np.array(h.cum_cost, dtype=np.float64)
    res["meta"] = np.array([K, max_length, end_id, feat.shape[0], H])
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **res)
    print(name, {k: int(v) for k, v in res.items() if k.endswith(".n")})


def main():
    torch.set_num_threads(1)
    torch.Tensor.cuda = lambda self, *a, **k: self        # CPU shim for hard-coded .cuda()
    rnn = _load("ref_rnn", "rnn.py")
    rnn_lstm = _load("ref_rnn_lstm", "LSTM/rnn_lstm.py")
    rnn_attn = _load("ref_rnn_attn", "Attention/rnn_attn.py")
    rnn_attn_lstm = _load("ref_rnn_attn_lstm", "Attention/rnn_attn_LSTM.py")
    bs = _load("ref_beam_search", "beam_search.py")

    small_len = [7, 6, 4, 4, 2]
    base_case(rnn, "gru", "gru_l1", 16, 24, 37, 1, 5, 7, small_len, 11, beams=(1, 3, 5))
    # tiny vocabulary so that <end>=2 is actually reached and hypotheses finish (beam_search.py:72-76)
    net, feat = base_case(rnn, "gru", "gru_tiny", 8, 12, 9, 1, 5, 7, small_len, 17, beams=(2,))
    tree_beam_case(bs, net, feat, "tree_beam_gru_tiny", 3, 15, 2)
    base_case(rnn, "gru", "gru_l2", 16, 24, 37, 2, 5, 7, small_len, 12, beams=(3,))
    base_case(rnn_lstm, "lstm", "lstm_l1", 16, 24, 37, 1, 5, 7, small_len, 13)
    base_case(rnn_lstm, "lstm", "lstm_l3", 16, 24, 37, 3, 5, 7, small_len, 14)
    med_len = [9, 9, 8, 6, 5, 3]
    base_case(rnn, "gru", "gru_med", 64, 64, 203, 1, 6, 9, med_len, 15, beams=(3,))
    base_case(rnn_lstm, "lstm", "lstm_med", 64, 64, 203, 1, 6, 9, med_len, 16)

    attn_case(rnn_attn, "attn_gru", "attn_gru_l1", 16, 20, 12, 24, 37, 1, 9, 5, 7, small_len, 1.0, 21)
    attn_case(rnn_attn, "attn_gru", "attn_gru_l2", 16, 20, 12, 24, 37, 2, 9, 5, 7, small_len, 1.0, 22)
    attn_case(rnn_attn_lstm, "attn_lstm", "attn_lstm_l1", 16, 20, 12, 24, 37, 1, 9, 5, 7, small_len, 1.0, 23)
    attn_case(rnn_attn_lstm, "attn_lstm", "attn_lstm_l2", 16, 20, 12, 24, 37, 2, 9, 5, 7, small_len, 0.5, 24)
    attn_case(rnn_attn, "attn_gru", "attn_gru_med", 64, 96, 64, 64, 203, 1, 49, 6, 9, med_len, 1.0, 25)
    attn_case(rnn_attn_lstm, "attn_lstm", "attn_lstm_med", 64, 96, 64, 64, 203, 1, 49, 6, 9, med_len, 1.0, 26)


if __name__ == "__main__":
    main()

"""CPU oracle for the show-tell caption-decoder hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``showtell_b200/`` imports it, and the product raises if its CUDA library is missing.

What it restates
----------------
The reference (guptakhil/show-tell) is pure Python on top of PyTorch; the arithmetic of the path
lives in a third-party dependency that is *not* under /root/reference: ``torch.nn.GRU``,
``torch.nn.LSTM``, ``nn.Linear``, ``nn.Embedding``, ``nn.Softmax``, ``nn.LeakyReLU``,
``nn.CrossEntropyLoss``, ``pack_padded_sequence`` (reference pins pytorch 1.1.0, README.md:14-25;
this image has torch 2.11.0 -- same published GRU/LSTM equations and gate order).  Every function
below writes those published equations out explicitly with matmul / sigmoid / tanh on CPU tensors
and cites the reference call site it follows.  Parameters are passed as a ``dict`` keyed by the
reference's ``state_dict`` names (``embeddings.weight``, ``unit.weight_ih_l0`` ... ``linear.bias``,
``init_h.*``, ``attn.encoder_att.*`` ...), so reference checkpoints plug in unchanged.

Pinning
-------
The reference ships no tests and no golden vectors, so by its own tests the
path is "parity unpinned".  The pin used here instead: ``tests/golden/make_golden.py`` imports the
*unmodified* reference modules from /root/reference in the build container, runs them on seeded
inputs and commits inputs + weights + outputs as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks every function in this file against those fixtures.

All functions are differentiable through torch autograd (gradient oracle) and dtype-generic
(float32 for parity with the reference, float64 for tight checks of the CUDA backward math).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

Params = Dict[str, torch.Tensor]

GRU, LSTM = "gru", "lstm"
_GATES = {GRU: 3, LSTM: 4}


# --------------------------------------------------------------------------------------------
# packing helpers (torch.nn.utils.rnn.pack_padded_sequence semantics, rnn.py:31, main.py:145)
# --------------------------------------------------------------------------------------------
def batch_sizes_from_lengths(lengths: Sequence[int]) -> List[int]:
    """batch_sizes[t] = #sequences with length > t.  Lengths must be sorted descending
    (pack_padded_sequence(enforce_sorted=True) raises otherwise; collate sorts, utils.py:66)."""
    lengths = [int(l) for l in lengths]
    if any(lengths[i] < lengths[i + 1] for i in range(len(lengths) - 1)):
        raise RuntimeError("lengths must be sorted in decreasing order")
    if len(lengths) == 0 or lengths[-1] <= 0:
        raise RuntimeError("every length must be > 0")
    return [sum(1 for l in lengths if l > t) for t in range(lengths[0])]


def pack_time_major(padded: torch.Tensor, lengths: Sequence[int]) -> torch.Tensor:
    """(B, T, ...) -> (N, ...) in packed time-major order: all live rows of t=0, then t=1, ..."""
    bs = batch_sizes_from_lengths(lengths)
    return torch.cat([padded[:b, t] for t, b in enumerate(bs)], dim=0)


# --------------------------------------------------------------------------------------------
# recurrent cells -- published torch.nn.GRU / torch.nn.LSTM equations
# --------------------------------------------------------------------------------------------
def gru_cell(gi: torch.Tensor, h: torch.Tensor, w_hh: torch.Tensor, b_hh: torch.Tensor) -> torch.Tensor:
    """gi = W_ih x + b_ih (b, 3H), gate order [r|z|n] (rnn.py:24 -> nn.GRU).
    r = s(gi_r + gh_r); z = s(gi_z + gh_z); n = tanh(gi_n + r*gh_n); h' = (1-z)*n + z*h."""
    H = h.shape[1]
    gh = h @ w_hh.t() + b_hh
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1.0 - z) * n + z * h


def lstm_cell(gi: torch.Tensor, h: torch.Tensor, c: torch.Tensor, w_hh: torch.Tensor,
              b_hh: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """gate order [i|f|g|o] (LSTM/rnn_lstm.py:22 -> nn.LSTM).
    c' = s(f)*c + s(i)*tanh(g); h' = s(o)*tanh(c')."""
    H = h.shape[1]
    a = gi + h @ w_hh.t() + b_hh
    i = torch.sigmoid(a[:, :H])
    f = torch.sigmoid(a[:, H:2 * H])
    g = torch.tanh(a[:, 2 * H:3 * H])
    o = torch.sigmoid(a[:, 3 * H:])
    c2 = f * c + i * g
    return o * torch.tanh(c2), c2


def _layer_weights(p: Params, layer: int):
    return (p[f"unit.weight_ih_l{layer}"], p[f"unit.weight_hh_l{layer}"],
            p[f"unit.bias_ih_l{layer}"], p[f"unit.bias_hh_l{layer}"])


def _num_layers(p: Params) -> int:
    n = 0
    while f"unit.weight_ih_l{n}" in p:
        n += 1
    return n


def _stack_step(p: Params, kind: str, x: torch.Tensor, hs: List[torch.Tensor],
                cs: Optional[List[torch.Tensor]]):
    """One time step through all L stacked layers (no dropout: reference uses the default 0).
    x (b, in); hs/cs: per-layer (b, H).  Returns top-layer output and the new state lists."""
    new_h, new_c = [], []
    inp = x
    for l in range(len(hs)):
        w_ih, w_hh, b_ih, b_hh = _layer_weights(p, l)
        gi = inp @ w_ih.t() + b_ih
        if kind == GRU:
            h2 = gru_cell(gi, hs[l], w_hh, b_hh)
        else:
            h2, c2 = lstm_cell(gi, hs[l], cs[l], w_hh, b_hh)
            new_c.append(c2)
        new_h.append(h2)
        inp = h2
    return inp, new_h, (new_c if kind == LSTM else None)


# --------------------------------------------------------------------------------------------
# F2: RNN.forward  (rnn.py:27-35, LSTM/rnn_lstm.py:25-33)
# --------------------------------------------------------------------------------------------
def rnn_forward(p: Params, kind: str, cnn_feature: torch.Tensor, caption: torch.Tensor,
                lengths: Sequence[int]) -> torch.Tensor:
    """Teacher-forced packed forward.  Input sequence of row b is
    [cnn_feature[b], emb(caption[b,0]), ..., emb(caption[b,len_b-2])] (rnn.py:29-31: the feature
    is prepended and only the first len_b of the T+1 steps survive packing); h0 = c0 = 0.
    Returns logits (N, V) in packed time-major order (rnn.py:32-33)."""
    L = _num_layers(p)
    H = p["unit.weight_hh_l0"].shape[1]
    B = cnn_feature.shape[0]
    bs = batch_sizes_from_lengths(lengths)
    emb = p["embeddings.weight"][caption]                       # rnn.py:29
    seq = torch.cat([cnn_feature.unsqueeze(1), emb], dim=1)     # rnn.py:30
    zeros = cnn_feature.new_zeros(B, H)
    hs = [zeros for _ in range(L)]
    cs = [zeros for _ in range(L)] if kind == LSTM else None
    tops = []
    for t, b in enumerate(bs):
        hs = [h[:b] for h in hs]
        if cs is not None:
            cs = [c[:b] for c in cs]
        top, hs, cs = _stack_step(p, kind, seq[:b, t], hs, cs)
        tops.append(top)
    packed_h = torch.cat(tops, dim=0)
    return packed_h @ p["linear.weight"].t() + p["linear.bias"]  # rnn.py:33


def ce_loss(logits: torch.Tensor, caption: torch.Tensor, lengths: Sequence[int]) -> torch.Tensor:
    """F3: nn.CrossEntropyLoss() (mean over the N packed rows) against
    pack_padded_sequence(caption, lengths)[0]  (main.py:94,145,149)."""
    target = pack_time_major(caption, lengths)
    m = logits.max(dim=1, keepdim=True).values
    lse = m.squeeze(1) + torch.log(torch.exp(logits - m).sum(dim=1))
    picked = logits.gather(1, target.view(-1, 1)).squeeze(1)
    return (lse - picked).mean()


# --------------------------------------------------------------------------------------------
# F4: greedy decode  (rnn.py:44-58, LSTM/rnn_lstm.py:35-57)
# --------------------------------------------------------------------------------------------
def _argmax_first(x: torch.Tensor) -> torch.Tensor:
    """Tensor.max(1)[1] on CPU returns the first maximal index."""
    return x.max(dim=1).indices


def rnn_greedy(p: Params, kind: str, cnn_feature: torch.Tensor, max_len: int = 25) -> torch.Tensor:
    """25 steps of {stacked step, Linear, arg-max, Embedding}; step 0 consumes the image feature
    with zero state.  Returns (B, max_len) int64, squeezed like rnn.py:56 (so (max_len,) at B=1)."""
    L = _num_layers(p)
    H = p["unit.weight_hh_l0"].shape[1]
    B = cnn_feature.shape[0]
    hs = [cnn_feature.new_zeros(B, H) for _ in range(L)]
    cs = [cnn_feature.new_zeros(B, H) for _ in range(L)] if kind == LSTM else None
    x = cnn_feature
    out = []
    for _ in range(max_len):
        top, hs, cs = _stack_step(p, kind, x, hs, cs)
        logits = top @ p["linear.weight"].t() + p["linear.bias"]
        tok = _argmax_first(logits)
        out.append(tok)
        x = p["embeddings.weight"][tok]
    return torch.stack(out, dim=1).squeeze()


# --------------------------------------------------------------------------------------------
# B1: RNN.sentence_index(beam_size=K>0)  (rnn.py:60-108) -- "chain" beam, batch 1 only
# --------------------------------------------------------------------------------------------
def rnn_beam_chain(p: Params, cnn_feature: torch.Tensor, beam_size: int, max_len: int = 25,
                   return_trace: bool = False):
    """Restatement of the inline beam search of rnn.py (GRU only; rnn_lstm.py has none).

    Quirks kept on purpose:
      * ONE hidden state is threaded through every GRU call in program order
        (rnn.py:61, 85-87): step 0, then for each of the max_len-1 rounds the K beams in turn.
        It is never re-ordered with the beams.
      * Candidates are ranked by that call's raw top-K logit only (rnn.py:90-100), not by a
        cumulative score and without a softmax.
      * Survivors: TWO independent descending sorts, one of (score, sentence) tuples and one
        of (score, word) tuples (rnn.py:102-103), first K of each.  On an exact score tie the
        tuple comparison falls through to the payload -- the sentence token list compared
        lexicographically, resp. the word id -- larger first; fully equal tuples keep
        generation order (python's sort is stable also under reverse=True).  Because the two
        sorts break ties differently, sentence k and word k may stop belonging together; that
        is reproduced, not repaired.
      * No <end> stop: always max_len tokens.  Output (max_len,) int64 (rnn.py:106-108).

    Conditioning: once every beam repeats one word the chained state converges and the K*K
    scores of a round differ by a few ulps only, so the ranking -- and with it the returned
    sentence -- is decided by rounding noise of whichever GEMM produced the logits (the
    reference run on CPU, on CUDA, or this restatement).  ``return_trace`` therefore also yields,
    per round, the descending candidate scores and the survivors, and ``beam_chain_margin`` turns
    that into "first round whose ranking is not separated by more than eps"; bit-exact
    comparison of token sequences is only defined up to that round.
    """
    if cnn_feature.shape[0] != 1:
        raise ValueError("rnn.py beam search only works with batch_size=1 (main.py:81-82)")
    K = int(beam_size)
    L = _num_layers(p)
    H = p["unit.weight_hh_l0"].shape[1]
    hs = [cnn_feature.new_zeros(1, H) for _ in range(L)]
    W, bias, E = p["linear.weight"], p["linear.bias"], p["embeddings.weight"]
    trace = []

    top, hs, _ = _stack_step(p, GRU, cnn_feature, hs, None)           # rnn.py:61
    logits = top @ W.t() + bias
    tk = torch.topk(logits[0], K + 1 if K < logits.shape[1] else K)
    words = tk.indices[:K].tolist()                                   # rnn.py:63
    sentences = [[w] for w in words]
    trace.append({"scores": tk.values.tolist(), "words": list(words), "sentences": [list(x) for x in sentences]})
    for _ in range(1, max_len):                                       # rnn.py:77-79
        cand = []
        for k in range(K):
            top, hs, _ = _stack_step(p, GRU, E[torch.tensor([words[k]])], hs, None)  # rnn.py:85-87
            logits = top @ W.t() + bias
            tk = torch.topk(logits[0], K)                             # rnn.py:90-91
            for j in range(K):
                cand.append((float(tk.values[j]), sentences[k] + [int(tk.indices[j])],
                             int(tk.indices[j])))
        by_sentence = sorted(cand, key=lambda c: (c[0], tuple(c[1])), reverse=True)   # rnn.py:102
        by_word = sorted(cand, key=lambda c: (c[0], c[2]), reverse=True)              # rnn.py:103
        sentences = [c[1] for c in by_sentence[:K]]
        words = [c[2] for c in by_word[:K]]
        trace.append({"scores": [c[0] for c in by_sentence], "words": list(words),
                      "sentences": [list(x) for x in sentences]})
    result = torch.tensor(sentences[0], dtype=torch.int64)            # rnn.py:106-108
    return (result, trace) if return_trace else result


def beam_chain_margin(trace, K: int, eps: float, ties_ok: bool = False) -> int:
    """First round (0-based) whose top-K ranking is not separated by more than ``eps``: some gap
    between consecutive scores among ranks 0..K (the K survivors and the best loser) is <= eps.
    ``ties_ok`` lets exact ties (gap == 0) pass: their order is defined by the payload rule, which
    two runs of the *same* arithmetic reproduce.  Returns len(trace) when every round is clean."""
    for r, rec in enumerate(trace):
        sc = rec["scores"][:K + 1]
        gaps = [sc[i] - sc[i + 1] for i in range(len(sc) - 1)]
        if any((g <= eps) and not (ties_ok and g == 0.0) for g in gaps):
            return r
    return len(trace)


# --------------------------------------------------------------------------------------------
# A1: Attention_Net.forward  (Attention/rnn_attn.py:21-31)
# --------------------------------------------------------------------------------------------
def leaky_relu02(x: torch.Tensor) -> torch.Tensor:
    return torch.where(x > 0, x, 0.2 * x)                             # nn.LeakyReLU(0.2), rnn_attn.py:18


def attention(p: Params, feat_bpc: torch.Tensor, h: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """feat_bpc (b, P, C) is the transposed view of the channels-first grid (rnn_attn.py:69).
    att1 = f W_e^T + b_e; att2 = h W_d^T + b_d; e = w_f . lrelu(att1 + att2) + b_f;
    alpha = softmax_P(e); ctx = sum_p alpha_p f_p.  Returns (ctx (b, C), alpha (b, P))."""
    att1 = feat_bpc @ p["attn.encoder_att.weight"].t() + p["attn.encoder_att.bias"]
    att2 = h @ p["attn.decoder_att.weight"].t() + p["attn.decoder_att.bias"]
    u = leaky_relu02(att1 + att2.unsqueeze(1))
    e = (u @ p["attn.full_att.weight"].t()).squeeze(2) + p["attn.full_att.bias"]
    e = e - e.max(dim=1, keepdim=True).values
    w = torch.exp(e)
    alpha = w / w.sum(dim=1, keepdim=True)
    ctx = (feat_bpc * alpha.unsqueeze(2)).sum(dim=1)
    return ctx, alpha


def _attn_init_state(p: Params, kind: str, cnn_feature: torch.Tensor, L: int):
    """rnn_attn.py:62 / rnn_attn_LSTM.py:63: every layer starts from init_h(mean_P f)
    (and init_c(mean_P f) for the LSTM)."""
    mean_f = cnn_feature.mean(dim=2)
    h0 = mean_f @ p["init_h.weight"].t() + p["init_h.bias"]
    hs = [h0 for _ in range(L)]
    cs = None
    if kind == LSTM:
        c0 = mean_f @ p["init_c.weight"].t() + p["init_c.bias"]
        cs = [c0 for _ in range(L)]
    return hs, cs


def _attn_step(p: Params, kind: str, feat_bpc, emb_t, hs, cs):
    """rnn_attn.py:69-70: query = pre-step top-layer hidden; layer-0 input = [emb_t | embed(ctx)]."""
    ctx, alpha = attention(p, feat_bpc, hs[-1])
    ctx_e = ctx @ p["embed.weight"].t() + p["embed.bias"]
    top, hs, cs = _stack_step(p, kind, torch.cat([emb_t, ctx_e], dim=1), hs, cs)
    return top, alpha, hs, cs


# --------------------------------------------------------------------------------------------
# A3/A4: RNN_Attn.rnn_iterator(is_train=True) + forward  (rnn_attn.py:60-76, 98-118)
# --------------------------------------------------------------------------------------------
def attn_forward(p: Params, kind: str, cnn_feature: torch.Tensor, caption: torch.Tensor,
                 lengths: Sequence[int]) -> Tuple[torch.Tensor, torch.Tensor]:
    """cnn_feature (B, C, P) channels-first.  No input/target shift: step t consumes
    emb(caption[:, t]) (rnn_attn.py:70).  Rows >= b_t are dropped from the carried state
    (rnn_attn.py:68-70).  Returns (logits (N, V) packed time-major, alphas (B, T, P) with
    zeros at padded steps) -- rnn_attn.py:64-65,72-73,115."""
    L = _num_layers(p)
    B, _, P = cnn_feature.shape
    T = caption.shape[1]
    lengths = [int(l) for l in lengths]
    emb = p["embeddings.weight"][caption]
    feat_bpc = cnn_feature.transpose(1, 2)
    hs, cs = _attn_init_state(p, kind, cnn_feature, L)
    tops, alpha_rows = [], []
    for t in range(T):
        b = sum(1 for l in lengths if l > t)                          # rnn_attn.py:68
        if b == 0:
            alpha_rows.append(cnn_feature.new_zeros(B, P))
            continue
        hs = [h[:b] for h in hs]
        if cs is not None:
            cs = [c[:b] for c in cs]
        top, alpha, hs, cs = _attn_step(p, kind, feat_bpc[:b], emb[:b, t], hs, cs)
        tops.append(top)
        alpha_rows.append(torch.cat([alpha, alpha.new_zeros(B - b, P)], dim=0))
    # pack_padded_sequence(predictions, caption_size)[0] keeps exactly the live (t, b<b_t) rows
    batch_sizes_from_lengths(lengths)
    logits = torch.cat(tops, dim=0) @ p["linear.weight"].t() + p["linear.bias"]
    return logits, torch.stack(alpha_rows, dim=1)


def attn_loss(logits: torch.Tensor, alphas: torch.Tensor, caption: torch.Tensor,
              lengths: Sequence[int], alpha_c: float) -> torch.Tensor:
    """A5: CE (same-index targets, main_attn.py:126,130) +
    alpha_c * mean_{B,P}((1 - sum_t alpha)^2)  (main_attn.py:131)."""
    return ce_loss(logits, caption, lengths) + alpha_c * ((1.0 - alphas.sum(dim=1)) ** 2).mean()


def attn_greedy(p: Params, kind: str, cnn_feature: torch.Tensor, start_id: int = 1,
                max_len: int = 25) -> torch.Tensor:
    """A6: rnn_attn.py:120-145 + 77-94: greedy from <start> (vocab('<start>') = 1,
    vocab_builder.py:68-69), attention recomputed on the full batch each step."""
    L = _num_layers(p)
    B = cnn_feature.shape[0]
    feat_bpc = cnn_feature.transpose(1, 2)
    hs, cs = _attn_init_state(p, kind, cnn_feature, L)
    x = p["embeddings.weight"][torch.full((B,), int(start_id), dtype=torch.int64)]
    out = []
    for _ in range(max_len):
        top, _, hs, cs = _attn_step(p, kind, feat_bpc, x, hs, cs)
        tok = _argmax_first(top @ p["linear.weight"].t() + p["linear.bias"])
        out.append(tok)
        x = p["embeddings.weight"][tok]
    return torch.stack(out, dim=1).squeeze()


# --------------------------------------------------------------------------------------------
# B2: beam_search.py:18-97 -- generic "tree" beam over callbacks (numpy)
# --------------------------------------------------------------------------------------------
class Hyp:
    """Counterpart of beam_search.Node (beam_search.py:18-43): value, parent link, flattened
    recurrent state, additive cum_cost, length, extras."""
    __slots__ = ("value", "parent", "state", "cum_cost", "length", "extras")

    def __init__(self, parent, state, value, cost, extras=None):
        self.value = value
        self.parent = parent
        self.state = None if state is None else np.asarray(state).reshape(-1)
        self.cum_cost = cost if parent is None else parent.cum_cost + cost
        self.length = 1 if parent is None else parent.length + 1
        self.extras = extras

    def to_sequence_of_values(self):
        seq, node = [], self
        while node is not None:
            seq.append(node.value)
            node = node.parent
        return seq[::-1]


def beam_search_tree(initial_state_function: Callable, generate_function: Callable, X, start_id: int,
                     end_id: int, beam_width: int = 4, num_hypotheses: int = 1,
                     max_length: int = 50) -> List[Hyp]:
    """Same contract as beam_search.beam_search (beam_search.py:45-97):
    finished nodes (value == end_id) leave the fringe first (:72-76); stop when the fringe is empty
    (:78); one batched generate call (:81-83) returning *probabilities*; per row the beam_width
    largest by ascending np.argsort (:84); cost = -log p (:88); keep the beam_width cheapest by a
    stable sort on cum_cost (:94); nodes alive after max_length rounds are dropped; result =
    hypotheses sorted by cum_cost, first num_hypotheses (:96-97)."""
    X = np.asarray(X)
    if X.ndim == 1:
        X = X.astype(np.int32).reshape(-1, 1)                         # beam_search.py:62-63
    assert X.ndim == 2 and X.shape[1] == 1
    live = [Hyp(None, initial_state_function(X), start_id, 0.0)]
    done: List[Hyp] = []
    for _ in range(max_length):
        fringe = []
        for n in live:
            (done if n.value == end_id else fringe).append(n)
        if not fringe:
            break
        y_prev = np.array([n.value for n in fringe], dtype=np.int32)
        s_prev = np.array([n.state for n in fringe], dtype=np.float32)
        s_t, p_t, extras_t = generate_function(X, y_prev, s_prev)
        best = np.argsort(p_t, axis=1)[:, -beam_width:]
        live = []
        for row, n in enumerate(fringe):
            for y in best[row]:
                ex = None if extras_t is None else extras_t[row]
                live.append(Hyp(n, s_t[row], int(y), float(-np.log(p_t[row][y])), ex))
        live = sorted(live, key=lambda n: n.cum_cost)[:beam_width]
    done.sort(key=lambda n: n.cum_cost)
    return done[:num_hypotheses]


def gru_tree_callbacks(p: Params, cnn_feature_row: torch.Tensor):
    """Adapter that drives beam_search_tree with the rnn.py GRU decoder (single layer, as
    beam_search.py:23 flattens one state): the initial state is the GRU state after consuming the
    image feature (step 0 of rnn.py:47-49); generate = embed, one GRU step, softmax(linear)."""
    assert _num_layers(p) == 1
    H = p["unit.weight_hh_l0"].shape[1]

    def init_fn(_X):
        with torch.no_grad():
            _, hs, _ = _stack_step(p, GRU, cnn_feature_row.view(1, -1), [cnn_feature_row.new_zeros(1, H)], None)
        return hs[0].numpy()

    def gen_fn(_X, y_prev, s_prev):
        with torch.no_grad():
            x = p["embeddings.weight"][torch.from_numpy(y_prev.astype(np.int64))]
            h = torch.from_numpy(np.asarray(s_prev, dtype=np.float32)).to(x.dtype)
            _, hs, _ = _stack_step(p, GRU, x, [h], None)
            logits = hs[0] @ p["linear.weight"].t() + p["linear.bias"]
            prob = torch.softmax(logits, dim=1)
        return hs[0].numpy(), prob.numpy(), None

    return init_fn, gen_fn


# --------------------------------------------------------------------------------------------
# convenience: one training step (loss + gradients) as the mains do it
# --------------------------------------------------------------------------------------------
def train_step(p: Params, model: str, cnn_feature, caption, lengths, alpha_c: float = 1.0):
    """model in {'gru','lstm','attn_gru','attn_lstm'}.  Mirrors main.py:145-151 /
    main_attn.py:126-133 without the optimizer.  Returns (loss, grads dict, extras)."""
    q = {k: v.detach().clone().requires_grad_(True) for k, v in p.items()}
    feat = cnn_feature.detach().clone().requires_grad_(not model.startswith("attn"))
    if model in (GRU, LSTM):
        logits = rnn_forward(q, model, feat, caption, lengths)
        loss = ce_loss(logits, caption, lengths)
        extras = {"logits": logits.detach()}
    else:
        kind = model.split("_")[1]
        logits, alphas = attn_forward(q, kind, feat, caption, lengths)
        loss = attn_loss(logits, alphas, caption, lengths, alpha_c)
        extras = {"logits": logits.detach(), "alphas": alphas.detach()}
    loss.backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in q.items()}
    if feat.requires_grad:
        grads["cnn_feature"] = feat.grad
    return loss.detach(), grads, extras


def flops_forward(model: str, B: int, T: int, E=512, H=512, V=10000, C=2048, P=196, A=512, L=1) -> float:
    """Algorithmic forward FLOPs for fixed-length captions, L=1."""
    N = B * T
    g = 3 if "gru" in model else 4
    if not model.startswith("attn"):
        return 2.0 * N * (g * H * (E + H) + H * V)
    step = 2.0 * B * (H * A + P * A + P * C + C * E + g * H * (E + H))
    hoist = 2.0 * B * P * C * A + 2.0 * N * g * H * E + 2.0 * B * C * H * (2 if g == 4 else 1)
    return hoist + T * step + 2.0 * N * H * V

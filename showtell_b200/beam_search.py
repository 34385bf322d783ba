"""Counterpart of the reference's beam_search.py (generic beam search, `beam_search.py:45-97`).

The reference function drives host callbacks (`initial_state_function`, `generate_function`) that
exchange numpy arrays once per step; nothing in the reference ever calls it.  A device
implementation cannot call back into Python per step, so the decoder plays both callbacks here:
the initial state is the GRU state after consuming the image feature (step 0 of rnn.py:47-49) and
one generate step is {embedding, GRU step, softmax(linear)} -- the adapter SURVEY.md section 8(a) row
B2 describes.  Search semantics are the reference's, line for line: finished nodes leave the
fringe first, cost = -log p accumulated in float32, the `beam_width` most probable tokens per node
in ascending order of probability, a stable sort on cum_cost, nodes alive after `max_length`
rounds dropped, hypotheses sorted by cum_cost and cut to `num_hypotheses`.  All images of the
batch are searched at once (`st_decode_beam_tree`).
"""
import torch

from . import decode


class Node:
    """Result record with the reference Node's read API (`beam_search.py:18-43`)."""

    def __init__(self, values, cum_cost):
        self._values = list(values)
        self.value = self._values[-1]
        self.cum_cost = float(cum_cost)
        self.length = len(self._values)

    def to_sequence_of_values(self):
        return list(self._values)

    def __repr__(self):
        return f"Node(values={self._values}, cum_cost={self.cum_cost:.6f})"


def beam_search(model, cnn_feature, start_id, end_id, beam_width=4, num_hypotheses=1, max_length=50):
    """model: showtell_b200.rnn.RNN (single-layer GRU, as `beam_search.py:23` keeps one flattened
    state); cnn_feature (B, E) CUDA tensor.  Returns, per image, the list of up to `num_hypotheses`
    finished hypotheses (possibly empty, as in the reference when none reaches `end_id`)."""
    with torch.no_grad():
        tok, ln, cost = decode.beam_tree(model, cnn_feature, start_id, end_id, beam_width, num_hypotheses,
                                         max_length)
    tok, ln, cost = tok.cpu(), ln.cpu(), cost.cpu()
    out = []
    for i in range(tok.shape[0]):
        hyps = []
        for j in range(tok.shape[1]):
            n = int(ln[i, j])
            if n > 0:
                hyps.append(Node(tok[i, j, :n].tolist(), cost[i, j]))
        out.append(hyps)
    return out

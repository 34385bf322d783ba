"""Drop-in for the reference's rnn.py: class RNN (GRU caption decoder).

    from showtell_b200.rnn import RNN        # instead of `from rnn import RNN` (main.py:20)

Same constructor, parameter names / shapes / initialisation (so reference checkpoints load),
forward(cnn_feature, image_caption, caption_size) and sentence_index(cnn_feature, beam_size) as
rnn.py:12-108; every tensor operation runs in libshowtell_b200.so.  `self.unit` is an nn.GRU used
ONLY as the parameter container that gives the reference's state_dict keys; its forward is never
called.
"""
import torch
import torch.nn as nn

from . import _lib, decode, engine


class RNN(nn.Module):
    _kind = _lib.ST_GRU
    _unit_cls = nn.GRU

    def __init__(self, embed_dim, num_hidden_units, vocab_size, num_layers, *, dtype="fp32"):
        super().__init__()
        if dtype not in ("fp32", "bf16"):
            raise ValueError('dtype must be "fp32" or "bf16"')
        self.embed_dim, self.num_hidden_units = int(embed_dim), int(num_hidden_units)
        self.vocab_size, self.num_layers = int(vocab_size), int(num_layers)
        self.compute_dtype = dtype
        self.caption_max_size = 25                                         # rnn.py:39
        self.embeddings = nn.Embedding(vocab_size, embed_dim)              # rnn.py:23
        self.unit = self._unit_cls(embed_dim, num_hidden_units, num_layers, batch_first=True)  # rnn.py:24
        self.linear = nn.Linear(num_hidden_units, vocab_size)              # rnn.py:25

    def _params(self):
        return [p for _, p in self.named_parameters()]

    def forward(self, cnn_feature, image_caption, caption_size):
        """rnn.py:27-35: (B,E), (B,T) int64, list of lengths sorted descending -> logits (N, V) in
        packed time-major order."""
        return engine.BaseLogitsFn.apply(self, cnn_feature, image_caption, list(caption_size), *self._params())

    def forward_loss(self, cnn_feature, image_caption, caption_size, global_tokens=None):
        """Fused training entry point: CrossEntropyLoss(forward(...), pack(caption)) of
        main.py:145-149 as one scalar; gradients are produced in the same pass."""
        return engine.BaseLossFn.apply(self, cnn_feature, image_caption, list(caption_size), global_tokens,
                                       *self._params())

    def forward_backward(self, cnn_feature, image_caption, caption_size, global_tokens=None):
        """The training iteration's `loss = loss_fn(rnn(...), target); loss.backward()` (main.py:148-151) as ONE call:
        returns the (detached) loss and leaves every parameter's .grad set to this step's gradient -- the same
        kernels as forward_loss(...).backward() without the autograd round trip (no chain-rule pass over the
        gradients with grad_output = 1).  Gradients are SET, not accumulated (the reference zeroes them every
        iteration, main.py:146); if cnn_feature requires grad its gradient flows on into the encoder head."""
        params = self._params()
        ctx = engine.DirectCtx(cnn_feature.requires_grad, 5, len(params))
        with torch.no_grad():
            loss = engine.BaseLossFn.forward(ctx, self, cnn_feature, image_caption, list(caption_size), global_tokens,
                                             *params)
        engine.assign_grads(self, ctx, cnn_feature)
        return loss

    def sentence_index(self, cnn_feature, beam_size=0, max_len=None, beam_mode="chain", **kw):
        """rnn.py:37-108.  beam_size=0: greedy.  beam_size=K>0: the reference's inline beam search
        (beam_mode="chain"); unlike the reference it is batched over images.  beam_mode="tree" runs
        beam_search.py semantics and returns the best hypothesis per image."""
        max_len = self.caption_max_size if max_len is None else int(max_len)
        with torch.no_grad():
            if beam_size == 0:
                tok = decode.greedy(self, cnn_feature, max_len)
            elif beam_mode == "chain":
                tok = decode.beam_chain(self, cnn_feature, beam_size, max_len)
            elif beam_mode == "tree":
                return decode.beam_tree(self, cnn_feature, kw.get("start_id", 1), kw.get("end_id", 2),
                                        beam_size, kw.get("num_hypotheses", 1), max_len)
            else:
                raise ValueError('beam_mode must be "chain" or "tree"')
        return tok.squeeze()                                               # rnn.py:56,107

    sample = sentence_index          # north-star alias


DecoderRNN = RNN                     # north-star alias

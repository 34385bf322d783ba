"""Drop-in for the reference's Attention/rnn_attn.py: class RNN_Attn (GRU + soft attention).

    from showtell_b200.rnn_attn import RNN_Attn as RNN   # instead of `from rnn_attn import ...` (main_attn.py:19)

Same constructor, parameter names (embeddings, unit, linear, init_h, attn.encoder_att /
decoder_att / full_att, embed), forward(cnn_feature, image_caption, caption_size) ->
(logits, alphas) and sentence_index(cnn_feature, vocab) as rnn_attn.py:35-145.  The submodules are
parameter containers only; all arithmetic runs in libshowtell_b200.so.
"""
import torch
import torch.nn as nn

from . import _lib, attn_decode, attn_engine


class Attention_Net(nn.Module):
    """Parameter container with the reference's names (rnn_attn.py:8-19)."""

    def __init__(self, nos_filters, num_hidden_units, attention_dim=512):
        super().__init__()
        self.encoder_att = nn.Linear(nos_filters, attention_dim)
        self.decoder_att = nn.Linear(num_hidden_units, attention_dim)
        self.full_att = nn.Linear(attention_dim, 1)


class RNN_Attn(nn.Module):
    _kind = _lib.ST_GRU
    _unit_cls = nn.GRU

    def __init__(self, embed_dim, nos_filters, attention_dim, num_hidden_units, vocab_size, num_layers, *,
                 dtype="fp32"):
        super().__init__()
        if dtype not in ("fp32", "bf16"):
            raise ValueError('dtype must be "fp32" or "bf16"')
        self.embed_dim, self.nos_filters, self.attention_dim = int(embed_dim), int(nos_filters), int(attention_dim)
        self.num_hidden_units, self.vocab_size, self.num_layers = int(num_hidden_units), int(vocab_size), int(num_layers)
        self.compute_dtype = dtype
        self.embeddings = nn.Embedding(vocab_size, embed_dim)                                 # rnn_attn.py:49
        self.unit = self._unit_cls(2 * embed_dim, num_hidden_units, num_layers, batch_first=True)  # :50
        self.linear = nn.Linear(num_hidden_units, vocab_size)                                 # :51
        self.cap_max_size = 25                                                                # :53
        self.init_h = nn.Linear(nos_filters, num_hidden_units)                                # :54
        if self._kind == _lib.ST_LSTM:
            self.init_c = nn.Linear(nos_filters, num_hidden_units)                            # rnn_attn_LSTM.py:55
        self.attn = Attention_Net(nos_filters, num_hidden_units, attention_dim)               # :56
        self.embed = nn.Linear(nos_filters, embed_dim)                                        # :58

    def _params(self):
        return [p for _, p in self.named_parameters()]

    def forward(self, cnn_feature, image_caption, caption_size):
        """rnn_attn.py:98-118: (B,C,P) channels-first grid, (B,T) int64, lengths sorted descending ->
        (logits (N,V) packed time-major, alphas (B,T,P) with zeros at padded steps)."""
        return attn_engine.AttnLogitsFn.apply(self, cnn_feature, image_caption, list(caption_size),
                                              *self._params())

    def forward_loss(self, cnn_feature, image_caption, caption_size, alpha_c=1.0, global_tokens=None,
                     global_batch=None):
        """Fused training entry point: the loss of main_attn.py:130-131 (CE + doubly-stochastic
        penalty) and all gradients in one pass.  Returns (loss, alphas)."""
        return attn_engine.AttnLossFn.apply(self, cnn_feature, image_caption, list(caption_size), alpha_c,
                                            global_tokens, global_batch, *self._params())

    def forward_backward(self, cnn_feature, image_caption, caption_size, alpha_c=1.0, global_tokens=None, global_batch=None):
        """`loss = CE + alpha_c * penalty; loss.backward()` (main_attn.py:129-133) as ONE call: returns (loss, alphas),
        both detached, and leaves every parameter's .grad set to this step's gradient (set, not accumulated; the
        grid features get no gradient, as in the reference: cnn_attn.py:47)."""
        from . import engine
        params = self._params()
        ctx = engine.DirectCtx(False, 7, len(params))
        with torch.no_grad():
            loss, alphas = attn_engine.AttnLossFn.forward(ctx, self, cnn_feature, image_caption, list(caption_size), alpha_c,
                                                          global_tokens, global_batch, *params)
        engine.assign_grads(self, ctx, cnn_feature.detach())
        return loss, alphas

    def sentence_index(self, cnn_feature, vocab, max_len=None):
        """rnn_attn.py:120-145: greedy decoding from vocab('<start>')."""
        max_len = self.cap_max_size if max_len is None else int(max_len)
        start = vocab("<start>") if callable(vocab) else int(vocab)
        with torch.no_grad():
            return attn_decode.greedy(self, cnn_feature, start, max_len).squeeze()

    sample = sentence_index


DecoderRNN = RNN_Attn

"""Drop-in for the reference's LSTM/rnn_lstm.py: class RNN (LSTM caption decoder).

    from showtell_b200.rnn_lstm import RNN   # instead of `from rnn_lstm import RNN` (main_lstm.py:19)
"""
import torch
import torch.nn as nn

from . import _lib, decode
from .rnn import RNN as _GruRNN


class RNN(_GruRNN):
    _kind = _lib.ST_LSTM
    _unit_cls = nn.LSTM                                                    # rnn_lstm.py:22

    def sentence_index(self, cnn_feature, max_len=None):
        """rnn_lstm.py:35-57: greedy only (the reference LSTM decoder has no beam search)."""
        max_len = self.caption_max_size if max_len is None else int(max_len)
        with torch.no_grad():
            return decode.greedy(self, cnn_feature, max_len).squeeze()

    sample = sentence_index


DecoderRNN = RNN

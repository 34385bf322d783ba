"""Drop-in for the reference's Attention/rnn_attn_LSTM.py: class RNN_Attn (LSTM + soft attention).

    from showtell_b200.rnn_attn_LSTM import RNN_Attn as RNN   # main_attn_LSTM.py:19
"""
import torch.nn as nn

from . import _lib
from .rnn_attn import Attention_Net, RNN_Attn as _GruAttn  # noqa: F401


class RNN_Attn(_GruAttn):
    _kind = _lib.ST_LSTM
    _unit_cls = nn.LSTM                                                    # rnn_attn_LSTM.py:50


DecoderRNN = RNN_Attn

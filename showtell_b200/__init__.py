"""showtell_b200: B200-native (sm_100a) implementation of the show-tell caption-decoder hot path.

Modules mirror the reference's file names so that only the import line of main*.py changes:
    showtell_b200.rnn.RNN            <- rnn.py
    showtell_b200.rnn_lstm.RNN       <- LSTM/rnn_lstm.py
    showtell_b200.rnn_attn.RNN_Attn  <- Attention/rnn_attn.py
    showtell_b200.rnn_attn_LSTM.RNN_Attn <- Attention/rnn_attn_LSTM.py
    showtell_b200.beam_search        <- beam_search.py
All arithmetic runs in libshowtell_b200.so (include/showtell_b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"

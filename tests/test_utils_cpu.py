"""CPU: the host-side helpers next to the decoder path: ids -> words and the checkpoint files."""
import os

import torch

from showtell_b200 import utils


class Vocab:
    """vocab_builder.DatasetVocabulary's interface (vocab_builder.py:11-44)."""

    def __init__(self, words):
        self.word_to_index, self.index_to_word = {}, {}
        for w in words:
            self.word_to_index[w] = len(self.word_to_index)
            self.index_to_word[self.word_to_index[w]] = w

    def start_token(self):
        return "<start>"

    def end_token(self):
        return "<end>"


def reference_words(tokenized, vocab, flag_blue=False):
    """utils.py:105-123 restated line by line (the reference module imports pycocotools / nltk at import time)."""
    caption_words = []
    for token in tokenized:
        curr_word = []
        for idx in token:
            if vocab.index_to_word[idx] == vocab.end_token():
                break
            if idx != vocab.word_to_index[vocab.start_token()]:
                curr_word.append(vocab.index_to_word[idx])
        caption_words.append([curr_word] if flag_blue else curr_word)
    return caption_words


def test_ids_to_words():
    v = Vocab(["<pad>", "<start>", "<end>", "<unk>", "a", "dog", "runs", "fast"])
    g = torch.Generator().manual_seed(0)
    tok = torch.randint(0, 8, (16, 25), generator=g)
    tok[0, :] = torch.tensor([1, 4, 5, 6, 2] + [7] * 20)           # <start> a dog runs <end> fast ...
    tok[1, :] = 4                                                  # no end token at all
    tok[2, 0] = 2                                                  # end token first: empty caption
    tok[3, :5] = torch.tensor([4, 1, 5, 1, 2])                     # start tokens in the middle are dropped too
    for flag in (False, True):
        assert utils.create_caption_word_format(tok, v, flag) == reference_words(tok.tolist(), v, flag)
        assert utils.create_caption_word_format(tok.tolist(), v, flag) == reference_words(tok.tolist(), v, flag)
    assert utils.create_caption_word_format(tok, v)[0] == ["a", "dog", "runs"]
    assert utils.create_caption_word_format(tok, v)[2] == []
    assert utils.create_caption_word_format(tok[0], v) == [["a", "dog", "runs"]]      # squeezed single caption (rnn.py:56)


def test_checkpoint_round_trip(tmp_path):
    """utils.py:125-145: file names, keys, and a state_dict that loads into torch's own modules."""
    torch.manual_seed(0)
    cnn, rnn = torch.nn.Linear(4, 3), torch.nn.GRU(3, 5)
    opt = torch.optim.Adam(list(cnn.parameters()) + list(rnn.parameters()), lr=1e-3)
    rnn(cnn(torch.randn(2, 1, 4)))[0].sum().backward()
    opt.step()
    params = {"output_dir": str(tmp_path)}
    utils.create_checkpoint(cnn, rnn, opt, 3, 17, [1.5, 1.25], params)
    assert sorted(os.listdir(tmp_path)) == ["model_3.ckpt", "model_3_metrics.ckpt"]
    ck = torch.load(os.path.join(tmp_path, "model_3.ckpt"))
    assert sorted(ck) == ["decoder_state_dict", "encoder_state_dict", "epoch", "optimizer_state_dict", "step"]
    assert torch.load(os.path.join(tmp_path, "model_3_metrics.ckpt")) == {"train_loss": [1.5, 1.25]}
    cnn2, rnn2 = torch.nn.Linear(4, 3), torch.nn.GRU(3, 5)
    opt2 = torch.optim.Adam(list(cnn2.parameters()) + list(rnn2.parameters()), lr=1e-3)
    assert utils.load_checkpoint(cnn2, rnn2, opt2, os.path.join(tmp_path, "model_3.ckpt")) == (3, 17)
    for a, b in zip(list(cnn.parameters()) + list(rnn.parameters()), list(cnn2.parameters()) + list(rnn2.parameters())):
        assert torch.equal(a, b)
    assert opt2.state_dict()["state"][0]["step"] == opt.state_dict()["state"][0]["step"]

"""Host-side helpers either side of the decoder path: token ids -> words (utils.py:105-123) and the
checkpoint files of a training run (utils.py:125-145), so that a training script built on the reference's `utils`
finds the same functions with the same results here.  Pure Python / torch.save: nothing runs on the GPU.
"""
import os

import torch


def create_caption_word_format(tokenized, vocab, flag_blue=False):
    """utils.py:105-123: every row of token ids (a tensor, array or list; the output of `sentence_index`) becomes its
    list of words -- the start token is dropped wherever it occurs, everything from the first end token on is cut.
    `vocab` needs `index_to_word`, `word_to_index`, `start_token()`, `end_token()` (vocab_builder.Vocabulary).
    flag_blue=True wraps each caption in a list (the reference-list shape nltk's corpus BLEU expects)."""
    if torch.is_tensor(tokenized):
        tokenized = tokenized.detach().cpu().tolist()
        if tokenized and not isinstance(tokenized[0], list):
            tokenized = [tokenized]                      # a single caption (sentence_index squeezes B = 1, rnn.py:56)
    end, start_id = vocab.end_token(), vocab.word_to_index[vocab.start_token()]
    caption_words = []
    for token in tokenized:
        curr_word = []
        for idx in token:
            idx = int(idx)
            if vocab.index_to_word[idx] == end:
                break
            if idx != start_id:
                curr_word.append(vocab.index_to_word[idx])
        caption_words.append([curr_word] if flag_blue else curr_word)
    return caption_words


def create_checkpoint(cnn, rnn, optimizer, epoch, step, train_loss, params):
    """utils.py:125-145: `model_<epoch>.ckpt` = {encoder_state_dict, decoder_state_dict, optimizer_state_dict, epoch,
    step} and `model_<epoch>_metrics.ckpt` = {train_loss} under params['output_dir'].  The decoders and the fused
    optimizers of this package keep the reference's state_dict keys and layouts, so files written here load into the
    reference's modules / torch.optim and vice versa."""
    out = params["output_dir"]
    torch.save({"encoder_state_dict": cnn.state_dict(), "decoder_state_dict": rnn.state_dict(),
                "optimizer_state_dict": optimizer.state_dict(), "epoch": epoch, "step": step},
               os.path.join(out, "model_" + str(epoch) + ".ckpt"))
    torch.save({"train_loss": train_loss}, os.path.join(out, "model_" + str(epoch) + "_metrics.ckpt"))
    print("Checkpoint created for Epoch %d (Step %d)." % (epoch, step))


def load_checkpoint(cnn, rnn, optimizer, path, map_location=None):
    """The inverse (main.py:117-122 / utils.py:151-153 do it inline): restores the three state dicts, returns (epoch, step)."""
    ck = torch.load(path, map_location=map_location)
    if cnn is not None:
        cnn.load_state_dict(ck["encoder_state_dict"])
    rnn.load_state_dict(ck["decoder_state_dict"])
    if optimizer is not None:
        optimizer.load_state_dict(ck["optimizer_state_dict"])
    return ck["epoch"], ck["step"]

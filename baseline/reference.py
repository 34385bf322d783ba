"""The UNMODIFIED reference decoder modules as a measurable arm (bench.py `--impl reference`, the
`cpu_baseline` leg and the torch-eager-on-B200 leg) and as a live checker in tests.

Test / measurement infrastructure only: nothing under showtell_b200/ imports this package
(tests/test_host_cpu.py enforces it).

The reference is a flat script collection (no setup.py), so "installing" it means placing its
decoder files, byte for byte, under baseline/_ref/ (git-ignored, shipped to the GPU box with the
repo snapshot): rnn.py (+ cnn.py, which rnn.py imports a class from), LSTM/rnn_lstm.py,
Attention/rnn_attn.py, Attention/rnn_attn_LSTM.py, beam_search.py.  `install()` does that whenever
/root/reference is present (the build container); on the GPU box the files are already there.
The modules are imported from those files and driven exactly as the reference's own mains drive
them: main.py:145-151, LSTM/main_lstm.py:123-130, Attention/main_attn.py:126-133.
"""
import contextlib
import importlib.util
import os
import shutil
import sys

import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.environ.get("SHOWTELL_REFERENCE", "/root/reference")
REF_DIR = os.path.join(HERE, "_ref")
FILES = ["rnn.py", "cnn.py", "LSTM/rnn_lstm.py", "Attention/rnn_attn.py", "Attention/rnn_attn_LSTM.py",
         "beam_search.py"]
# model name -> (file, class)
MODULES = {"gru": ("rnn.py", "RNN"), "beam": ("rnn.py", "RNN"), "lstm": ("LSTM/rnn_lstm.py", "RNN"),
           "attn_gru": ("Attention/rnn_attn.py", "RNN_Attn"), "attn_lstm": ("Attention/rnn_attn_LSTM.py", "RNN_Attn")}


def install():
    """Copy the reference's decoder files verbatim into baseline/_ref/.  Returns the directory, or None
    when the reference sources are not on this machine."""
    if not os.path.exists(os.path.join(REF_SRC, "rnn.py")):
        return REF_DIR if available() else None
    for rel in FILES:
        dst = os.path.join(REF_DIR, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(REF_SRC, rel), dst)
    return REF_DIR


def root():
    """Directory the reference modules are imported from."""
    if os.path.exists(os.path.join(REF_DIR, "rnn.py")):
        return REF_DIR
    if os.path.exists(os.path.join(REF_SRC, "rnn.py")):
        return REF_SRC
    return None


def available():
    return root() is not None


_loaded = {}


def load(rel):
    """Import one reference file as a module (cached)."""
    base = root()
    if base is None:
        raise RuntimeError("reference modules not installed (baseline/_ref missing and /root/reference absent)")
    key = (base, rel)
    if key in _loaded:
        return _loaded[key]
    name = "showtell_ref_" + rel.replace("/", "_").replace(".py", "")
    spec = importlib.util.spec_from_file_location(name, os.path.join(base, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.path.insert(0, base)                    # rnn.py does `from cnn import ResNet` (class import only)
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.path.pop(0)
    _loaded[key] = mod
    return mod


@contextlib.contextmanager
def cpu_cuda_shim():
    """The attention files hard-code `.cuda()` (Attention/rnn_attn.py:64,65,128).  On the host-CPU arm that call
    is made the identity for the duration of the run; the module source stays untouched."""
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.Tensor.cuda = orig


def make_module(model, E, H, V, L=1, C=2048, A=512):
    rel, cls = MODULES[model]
    ctor = getattr(load(rel), cls)
    if model.startswith("attn"):
        return ctor(E, C, A, H, V, L)            # main_attn.py:87
    return ctor(E, H, V, L)                      # main.py:93


def train_step(net, model, feature, caption, lengths, alpha_c=1.0):
    """One training iteration minus the optimizer, as the reference's mains write it.  Returns the loss."""
    for p in net.parameters():
        p.grad = None                                                                       # optimizer.zero_grad()
    target = nn.utils.rnn.pack_padded_sequence(caption, lengths, batch_first=True)[0]       # main.py:145
    if model.startswith("attn"):
        logits, alphas = net(feature, caption, lengths)                                     # main_attn.py:129
        loss = nn.CrossEntropyLoss()(logits, target)                                        # main_attn.py:130
        loss = loss + alpha_c * ((1. - alphas.sum(dim=1)) ** 2).mean()                      # main_attn.py:131
    else:
        logits = net(feature, caption, lengths)                                             # main.py:148
        loss = nn.CrossEntropyLoss()(logits, target)                                        # main.py:149
    loss.backward()                                                                         # main.py:151
    return loss


def beam_captions(net, features, K):
    """rnn.py beam search, one image per call as the reference requires (main.py:81-82); the reference
    hard-codes 25 steps (rnn.py:39).  No torch.no_grad(): the reference has none (utils.py:194)."""
    out = []
    for i in range(features.shape[0]):
        out.append(net.sentence_index(features[i:i + 1], beam_size=K))                      # utils.py:194
    return out

"""CPU: host-side logic and the C-ABI surface (no compute calls without a GPU)."""
import os
import re

import pytest
import torch

from conftest import ROOT


def test_library_exports_every_header_symbol():
    from showtell_b200 import _lib, build
    build.build()
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "showtell_b200.h")).read()
    declared = set(re.findall(r"\b(st_[a-z0-9_]+)\s*\(", header))
    declared -= {"st_status"}
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/showtell_b200.h but not exported"
    assert declared == set(_lib.exported_symbols()), declared ^ set(_lib.exported_symbols())
    assert lib.st_version() >= 100


def test_batch_sizes_matches_pack_padded_sequence():
    from showtell_b200 import _lib
    for lengths in ([7, 6, 4, 4, 2], [5], [3, 3, 3], [20] * 32, [9, 1]):
        x = torch.zeros(len(lengths), max(lengths), 1)
        ref = torch.nn.utils.rnn.pack_padded_sequence(x, lengths, batch_first=True).batch_sizes.tolist()
        assert _lib.batch_sizes(lengths) == ref
    for bad in ([2, 3], [], [3, 0]):
        with pytest.raises(RuntimeError):
            _lib.batch_sizes(bad)


def test_state_dict_keys_match_reference_checkpoints():
    from helpers import golden_params, load_golden
    from showtell_b200.rnn import RNN as GRU
    from showtell_b200.rnn_lstm import RNN as LSTM
    for name, cls in (("gru_l2", GRU), ("lstm_l3", LSTM)):
        g = load_golden(name)
        E, H, V, L, _, _ = g["dims"].tolist()
        m = cls(E, H, V, L)
        p = golden_params(g)
        assert set(m.state_dict().keys()) == set(p.keys())
        m.load_state_dict(p)
        for k, v in m.state_dict().items():
            assert v.shape == p[k].shape


def test_no_cpu_fallback():
    from showtell_b200.rnn import RNN
    m = RNN(8, 8, 11, 1)
    with pytest.raises(RuntimeError):
        m(torch.randn(2, 8), torch.zeros(2, 3, dtype=torch.int64), [3, 2])
    with pytest.raises(RuntimeError):
        m.sentence_index(torch.randn(2, 8))


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "showtell_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("no oracle", ""), f
                assert not re.search(r"^\s*(from|import)\s+baseline", src, re.M), f     # the reference arm is bench / test only


def test_reference_install_is_verbatim():
    """baseline/_ref holds the reference's decoder files byte for byte (bench.py's reference arm runs the stock code)."""
    import filecmp
    from baseline import reference as R
    if not os.path.exists(os.path.join(R.REF_SRC, "rnn.py")):
        pytest.skip("reference sources not on this machine")
    assert R.install() == R.REF_DIR
    for rel in R.FILES:
        assert filecmp.cmp(os.path.join(R.REF_SRC, rel), os.path.join(R.REF_DIR, rel), shallow=False), rel
    net = R.make_module("lstm", 8, 8, 20, 1)
    assert type(net).__module__.startswith("showtell_ref_") and isinstance(net.unit, torch.nn.LSTM)


def test_host_modules_reference_only_defined_ops():
    """Every `ops.<name>` / `_lib.<name>` the host-side modules use exists (a truncated ops.py once
    shipped with the attention wrappers missing: only the GPU tests would have noticed)."""
    import importlib
    pkg = os.path.join(ROOT, "showtell_b200")
    mods = {"ops": importlib.import_module("showtell_b200.ops"),
            "_lib": importlib.import_module("showtell_b200._lib"),
            "graphs": importlib.import_module("showtell_b200.graphs"),
            "parallel": importlib.import_module("showtell_b200.parallel")}
    for f in sorted(os.listdir(pkg)):
        if not f.endswith(".py"):
            continue
        src = open(os.path.join(pkg, f)).read()
        for modname, mod in mods.items():
            for name in set(re.findall(r"(?<![\w.])" + modname + r"\.([A-Za-z_]\w*)", src)):
                assert hasattr(mod, name), f"{f} uses {modname}.{name}, which does not exist"


def test_optimizers_mirror_torch_and_refuse_cpu_tensors():
    """showtell_b200.optim: same param-group keys as torch.optim (checkpoint interchange, utils.py:125-145), the
    reference's bad-option errors, and no CPU path."""
    from showtell_b200 import optim
    p = [torch.nn.Parameter(torch.randn(3, 4))]
    assert set(optim.Adam(p, lr=1e-3).param_groups[0]) == set(torch.optim.Adam(p, lr=1e-3).param_groups[0])
    assert set(optim.SGD(p, lr=1e-3, momentum=0.9).param_groups[0]) == \
        set(torch.optim.SGD(p, lr=1e-3, momentum=0.9).param_groups[0])
    for bad in (lambda: optim.SGD(p, lr=-1.0), lambda: optim.SGD(p, lr=0.1, momentum=-0.5),
                lambda: optim.Adam(p, lr=-1.0), lambda: optim.Adam(p, lr=1e-3, betas=(1.0, 0.9)),
                lambda: optim.Adam(p, lr=1e-3, eps=-1.0)):
        with pytest.raises(ValueError):
            bad()
    p[0].grad = torch.randn(3, 4)
    for o in (optim.Adam(p, lr=1e-3), optim.SGD(p, lr=1e-3, momentum=0.9)):
        with pytest.raises(RuntimeError):
            o.step()
    # a torch checkpoint loads
    t = torch.optim.Adam(p, lr=1e-3)
    t.step()
    o = optim.Adam(p, lr=5e-4)
    o.load_state_dict(t.state_dict())
    assert o.param_groups[0]["lr"] == 1e-3 and set(o.state[p[0]]) == {"step", "exp_avg", "exp_avg_sq"}


def test_encoder_head_names_init_and_cpu_refusal():
    """cnn.py:37-42: sub-module names / shapes of the two trained ResNet layers; CPU input raises."""
    from showtell_b200.cnn_head import EncoderHead
    h = EncoderHead(2048, 512)
    assert set(h.state_dict()) == {"linear_secondlast_layer.weight", "linear_secondlast_layer.bias", "last_layer.weight",
                                   "last_layer.bias", "last_layer.running_mean", "last_layer.running_var",
                                   "last_layer.num_batches_tracked"}
    assert h.last_layer.momentum == 0.01 and float(h.last_layer.bias.abs().max()) == 0.0
    assert 0.04 < float(h.linear_secondlast_layer.weight.std()) < 0.06
    with pytest.raises(RuntimeError):
        h(torch.randn(4, 2048))
    with pytest.raises(ValueError):
        EncoderHead(8, 4, dtype="fp16")


def test_grad_reducer_single_process_is_identity():
    from showtell_b200 import parallel
    red = parallel.GradReducer()
    ts = [torch.randn(3, 3), torch.randn(5)]
    out = red.reduce(ts)
    assert all(a is b for a, b in zip(out, ts)) and red.slots([t.shape for t in ts]) is None
    red.finish()


def test_collate_contract_matches_create_batch():
    """utils.py:61-77: length-descending stable order, zero padding, lengths list; sort_batch gives the same batch
    from unsorted tensors and its permutation restores the caller's order."""
    from showtell_b200 import _lib, collate
    g = torch.Generator().manual_seed(0)
    lens = [4, 7, 4, 2, 7, 5, 1]
    data = [(f"img{i}.jpg", torch.randn(3, 4, 4, generator=g), torch.randint(1, 50, (n,), generator=g))
            for i, n in enumerate(lens)]
    # the reference, restated verbatim (it sorts the list in place)
    ref = list(data)
    ref.sort(key=lambda x: len(x[2]), reverse=True)
    paths, images, target, caption_len = collate.create_batch(data)
    assert list(paths) == [r[0] for r in ref] and caption_len == [len(r[2]) for r in ref] == [7, 7, 5, 4, 4, 2, 1]
    assert paths[:2] == ("img1.jpg", "img4.jpg") and paths[3:5] == ("img0.jpg", "img2.jpg")      # ties keep their order
    assert torch.equal(images, torch.stack([r[1] for r in ref]))
    for i, r in enumerate(ref):
        assert torch.equal(target[i, :len(r[2])], r[2]) and int(target[i, len(r[2]):].abs().sum()) == 0
    assert _lib.batch_sizes(caption_len) == [7, 6, 5, 5, 3, 2, 2]
    # unsorted padded tensors -> the same batch
    padded = torch.zeros(len(lens), 9, dtype=torch.long)
    for i, d in enumerate(data):
        padded[i, :lens[i]] = d[2]
    feats = torch.stack([d[1] for d in data])
    f2, c2, l2, perm = collate.sort_batch(feats, padded, torch.tensor(lens))
    assert l2 == caption_len and torch.equal(f2, images) and torch.equal(c2, target)
    restored = torch.empty_like(f2)
    restored[perm] = f2
    assert torch.equal(restored, feats)
    with pytest.raises(RuntimeError):
        collate.sort_batch(feats, padded, [4, 7, 4, 0, 7, 5, 1])
    with pytest.raises(ValueError):
        collate.sort_batch(feats, padded, [4, 7, 4, 2, 7, 5])

"""GPU parity of the base decoders (rnn.py / LSTM/rnn_lstm.py) against
  (a) the golden fixtures produced by the unmodified reference, and
  (b) the CPU oracle on seeded inputs at larger sizes.
fp32 mode; the bar is north_star's 1e-4 relative for loss / logits / gradients and bit-exact token
ids (outside rounding-noise-decided rankings, see oracle.rnn_beam_chain)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from helpers import BASE_CASES, golden_grads, golden_params, load_golden, rel_err
from oracle import showtell_oracle as O

TOL = 1e-4


def _module(g, dev):
    E, H, V, L, _, _ = g["dims"].tolist()
    kind = str(g["kind"])
    if kind == "gru":
        from showtell_b200.rnn import RNN
    else:
        from showtell_b200.rnn_lstm import RNN
    m = RNN(E, H, V, L)
    m.load_state_dict(golden_params(g))           # reference checkpoint keys load unchanged
    return m.to(dev)


@pytest.mark.parametrize("name", BASE_CASES)
def test_golden_forward_backward(name):
    dev = torch.device("cuda:0")
    g = load_golden(name)
    m = _module(g, dev)
    feat = torch.from_numpy(g["cnn_feature"]).to(dev).requires_grad_(True)
    cap = torch.from_numpy(g["caption"]).to(dev)
    lengths = g["lengths"].tolist()
    logits = m(feat, cap, lengths)                                        # main.py:148
    assert logits.shape == g["logits"].shape
    assert rel_err(logits, g["logits"]) < TOL
    target = torch.nn.utils.rnn.pack_padded_sequence(cap, lengths, batch_first=True)[0]
    loss = torch.nn.CrossEntropyLoss()(logits, target)                    # main.py:149
    assert abs(float(loss) - float(g["loss"])) < TOL * abs(float(g["loss"]))
    loss.backward()
    gg = golden_grads(g)
    for n, p in m.named_parameters():
        assert rel_err(p.grad, gg[n]) < TOL, n
    assert rel_err(feat.grad, gg["cnn_feature"]) < TOL


@pytest.mark.parametrize("name", BASE_CASES)
def test_golden_fused_loss(name):
    dev = torch.device("cuda:0")
    g = load_golden(name)
    m = _module(g, dev)
    feat = torch.from_numpy(g["cnn_feature"]).to(dev).requires_grad_(True)
    cap = torch.from_numpy(g["caption"]).to(dev)
    loss = m.forward_loss(feat, cap, g["lengths"].tolist())
    assert abs(float(loss) - float(g["loss"])) < TOL * abs(float(g["loss"]))
    (loss * 3.0).backward()
    gg = golden_grads(g)
    for n, p in m.named_parameters():
        assert rel_err(p.grad / 3.0, gg[n]) < TOL, n
    assert rel_err(feat.grad / 3.0, gg["cnn_feature"]) < TOL


@pytest.mark.parametrize("name", BASE_CASES)
def test_golden_greedy(name):
    dev = torch.device("cuda:0")
    g = load_golden(name)
    m = _module(g, dev)
    feat = torch.from_numpy(g["cnn_feature"]).to(dev)
    tok = m.sentence_index(feat)
    assert tok.shape == (feat.shape[0], 25) and tok.dtype == torch.int64
    assert np.array_equal(tok.cpu().numpy(), g["greedy"])
    tok1 = m.sentence_index(feat[:1])
    assert tok1.shape == (25,)                                            # squeeze quirk (rnn.py:56)
    assert np.array_equal(tok1.cpu().numpy(), g["greedy_b1"])


@pytest.fixture
def table_mode():
    """st_debug_decode_table: pin how the tensor-core decoding loops obtain the input projection of a fed-back word
    (1 = rows of the projected-embedding table, -1 = a GEMM per step); restored to "by size" afterwards."""
    from showtell_b200 import _lib
    lib = _lib.load()
    yield lambda mode: lib.st_debug_decode_table(int(mode))
    lib.st_debug_decode_table(0)


@pytest.mark.parametrize("table", [-1, 1])
@pytest.mark.parametrize("name", BASE_CASES)
def test_golden_greedy_tensor_cores(name, table, table_mode):
    """decode_gemm="tf32x3": every product of the loop on the tensor cores, arg-max fused into the vocabulary
    projection (no logits buffer) -- the same tokens as the reference run."""
    dev = torch.device("cuda:0")
    g = load_golden(name)
    m = _module(g, dev)
    m.decode_gemm = "tf32x3"
    table_mode(table)
    feat = torch.from_numpy(g["cnn_feature"]).to(dev)
    assert np.array_equal(m.sentence_index(feat).cpu().numpy(), g["greedy"])
    assert np.array_equal(m.sentence_index(feat[:1]).cpu().numpy(), g["greedy_b1"])


# ranking margins (in logit units) below which a chain-beam round is decided by the rounding noise of the
# products: 1e-5 for fp32 CUDA-core GEMMs, 4e-5 for the 3xTF32 tensor-core GEMMs (test_gemm_tf32x3_is_fp32_accurate)
EPS = {"fp32": 1e-5, "tf32x3": 4e-5}


def _check_chain(m, p, feat, K, max_len=25, eps=1e-5):
    """CUDA chain beam vs the oracle, row by row: identical survivors (words + scores) in every
    round up to the first round whose ranking hangs on <= eps of logit; identical final sentence
    when no round is that close.  Returns (#rows compared to the end, #rows)."""
    from showtell_b200 import decode
    tok, ts, tw = decode.beam_chain(m, feat, K, max_len, trace=True)
    tok, ts, tw = tok.cpu(), ts.cpu(), tw.cpu()
    full = 0
    pc = {k: v.cpu() for k, v in p.items()}
    for i in range(feat.shape[0]):
        with torch.no_grad():
            seq, trace = O.rnn_beam_chain(pc, feat[i:i + 1].cpu(), K, max_len, return_trace=True)
        stop = O.beam_chain_margin(trace, K, eps)
        for r in range(min(stop, len(trace))):
            assert tw[r, i].tolist() == trace[r]["words"], (i, r)
            assert np.allclose(ts[r, i].numpy(), np.array(trace[r]["scores"][:K]), atol=2 * eps), (i, r)
        if stop == len(trace):
            full += 1
            assert tok[i].tolist() == seq.tolist(), i
    return full, feat.shape[0]


# decode_gemm: the nn.Linear products of the decoding loops on the CUDA cores (fp32) or as 3xTF32 on the
# tensor cores (fp32-accurate) -- the same parity criteria hold for both
@pytest.mark.parametrize("gemm", ["fp32", "tf32x3"])
@pytest.mark.parametrize("name,K", [("gru_l1", 1), ("gru_l1", 3), ("gru_l1", 5), ("gru_tiny", 2),
                                    ("gru_l2", 3), ("gru_med", 3)])
def test_golden_beam_chain(name, K, gemm):
    dev = torch.device("cuda:0")
    g = load_golden(name)
    m = _module(g, dev)
    m.decode_gemm = gemm
    feat = torch.from_numpy(g["cnn_feature"]).to(dev)
    full, n = _check_chain(m, golden_params(g), feat, K, eps=EPS[gemm])
    tok = m.sentence_index(feat, beam_size=K)
    same = (tok.cpu().numpy() == g[f"beam_chain_k{K}"]).all(axis=1).sum()
    print(f"{name} K={K}: {same}/{n} rows identical to the reference run, {full}/{n} with separated rankings")
    # rows whose rankings are separated equal the oracle (asserted above), which equals the reference run on
    # those rows (tests/test_oracle_golden.py): so at least that many rows equal the reference's own output
    assert same >= full
    if K == 1:
        assert np.array_equal(tok.cpu().numpy(), g["greedy"])             # rnn.py:43
    one = m.sentence_index(feat[:1], beam_size=K)
    assert one.shape == (25,)                                             # rnn.py:107 squeeze


def test_golden_beam_tree():
    dev = torch.device("cuda:0")
    g = load_golden("gru_tiny")
    t = load_golden("tree_beam_gru_tiny")
    K, max_length, end_id, rows, _ = t["meta"].tolist()
    m = _module(g, dev)
    feat = torch.from_numpy(g["cnn_feature"]).to(dev)
    tok, ln, cost = m.sentence_index(feat, beam_size=K, beam_mode="tree", max_len=max_length,
                                     start_id=1, end_id=end_id, num_hypotheses=K)
    tok, ln, cost = tok.cpu(), ln.cpu(), cost.cpu()
    for i in range(rows):
        n = int(t[f"row{i}.n"])
        assert int((ln[i] > 0).sum()) == n
        for j in range(n):
            ref = t[f"row{i}.hyp{j}.seq"].tolist()
            assert tok[i, j, :int(ln[i, j])].tolist() == ref, (i, j)
            assert abs(float(cost[i, j]) - float(t[f"row{i}.hyp{j}.cost"])) < 1e-4


def test_beam_search_module_api():
    """showtell_b200.beam_search.beam_search mirrors beam_search.py's results (golden fixture)."""
    from showtell_b200.beam_search import beam_search
    g = load_golden("gru_tiny")
    t = load_golden("tree_beam_gru_tiny")
    K, max_length, end_id, rows, _ = t["meta"].tolist()
    m = _module(g, torch.device("cuda:0"))
    res = beam_search(m, torch.from_numpy(g["cnn_feature"]).cuda(), 1, end_id, beam_width=K, num_hypotheses=K,
                      max_length=max_length)
    for i in range(rows):
        assert len(res[i]) == int(t[f"row{i}.n"])
        for j, h in enumerate(res[i]):
            assert h.to_sequence_of_values() == t[f"row{i}.hyp{j}.seq"].tolist()
            assert abs(h.cum_cost - float(t[f"row{i}.hyp{j}.cost"])) < 1e-4


def _random_case(kind, E, H, V, L, B, T, seed, ragged, dtype="fp32"):
    torch.manual_seed(seed)
    g = torch.Generator().manual_seed(seed)
    if kind == "gru":
        from showtell_b200.rnn import RNN
    else:
        from showtell_b200.rnn_lstm import RNN
    m = RNN(E, H, V, L, dtype=dtype)
    lengths = sorted(torch.randint(max(2, T // 3), T + 1, (B,), generator=g).tolist(), reverse=True) \
        if ragged else [T] * B
    lengths[0] = T
    cap = torch.zeros(B, T, dtype=torch.int64)
    for b, l in enumerate(lengths):
        cap[b, 0] = 1
        cap[b, 1:l - 1] = torch.randint(4, V, (l - 2,), generator=g)
        cap[b, l - 1] = 2
    feat = torch.randn(B, E, generator=g)
    return m, feat, cap, lengths


@pytest.mark.parametrize("kind,E,H,V,L,B,T,ragged", [
    ("gru", 512, 512, 10000, 1, 32, 20, False),      # BASELINE config 1
    ("lstm", 512, 512, 10000, 1, 48, 20, True),
    ("gru", 96, 160, 1000, 3, 150, 12, True),        # >128 rows: two batch tiles, 3 layers
    ("lstm", 64, 128, 777, 2, 9, 5, True),
])
def test_oracle_parity_train(kind, E, H, V, L, B, T, ragged):
    _parity_train(kind, E, H, V, L, B, T, ragged, "fp32", TOL)


@pytest.mark.parametrize("kind,E,H,V,L,B,T,ragged", [
    ("lstm", 512, 512, 10000, 1, 64, 20, False),     # BASELINE config 2 shapes (smaller batch)
    ("gru", 512, 512, 10000, 1, 32, 20, True),
    ("lstm", 64, 128, 777, 2, 150, 7, True),         # odd vocabulary, two layers, two batch tiles
])
def test_oracle_parity_train_bf16(kind, E, H, V, L, B, T, ragged):
    """bf16 mode (tensor-core GEMMs, fused vocabulary CE): north_star's 2e-2 relative bar."""
    _parity_train(kind, E, H, V, L, B, T, ragged, "bf16", 2e-2)


def _parity_train(kind, E, H, V, L, B, T, ragged, dtype, TOL):
    dev = torch.device("cuda:0")
    m, feat, cap, lengths = _random_case(kind, E, H, V, L, B, T, 7, ragged, dtype)
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    loss_ref, grads_ref, ex = O.train_step(p, kind, feat, cap, lengths)
    m = m.to(dev)
    featg = feat.to(dev).requires_grad_(True)
    loss = m.forward_loss(featg, cap.to(dev), lengths)
    loss.backward()
    assert abs(float(loss) - float(loss_ref)) < TOL * float(loss_ref)
    for n, q in m.named_parameters():
        assert rel_err(q.grad, grads_ref[n]) < TOL, n
    assert rel_err(featg.grad, grads_ref["cnn_feature"]) < TOL
    with torch.no_grad():
        logits = m(feat.to(dev), cap.to(dev), lengths)
    assert rel_err(logits, ex["logits"]) < TOL


@pytest.mark.parametrize("gemm,table", [("fp32", 0), ("tf32x3", -1), ("tf32x3", 1), ("tf32x3", 3)])
@pytest.mark.parametrize("kind,L", [("gru", 1), ("lstm", 1), ("gru", 2), ("lstm", 2)])
def test_oracle_parity_greedy_full_size(kind, L, gemm, table, table_mode):
    dev = torch.device("cuda:0")
    m, feat, _, _ = _random_case(kind, 512, 512, 10000, L, 16, 20, 11, False)
    m.decode_gemm = gemm
    from showtell_b200 import _lib
    _lib.load().st_debug_decode_screen(-1 if table == 3 else 0)     # 3: table + the 3xTF32 fused arg-max instead of screening
    table_mode(1 if table == 3 else table)
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = O.rnn_greedy(p, kind, feat)
    tok = m.to(dev).sentence_index(feat.to(dev))
    _lib.load().st_debug_decode_screen(0)
    assert np.array_equal(tok.cpu().numpy(), ref.numpy())


@pytest.mark.parametrize("gemm", ["fp32", "tf32x3"])
def test_oracle_parity_beam_tree_full_size(gemm):
    """beam_search.py:45-97 at E=H=512, V=10000, beam 3, 20 rounds, 24 images in one call against the oracle's
    per-image run.  The <end> logit is lifted so hypotheses finish at various lengths.  tf32x3: top-K and the
    soft-max normaliser come out of the vocabulary GEMM's epilogue (no logits buffer)."""
    dev = torch.device("cuda:0")
    n, K, NH, T, end_id = 24, 3, 3, 20, 2
    m, feat, _, _ = _random_case("gru", 512, 512, 10000, 1, n, 20, 23, False)
    with torch.no_grad():
        m.linear.bias[end_id] += 0.45
    m.decode_gemm = gemm
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    tok, ln, cost = m.to(dev).sentence_index(feat.to(dev), beam_size=K, beam_mode="tree", max_len=T, start_id=1,
                                            end_id=end_id, num_hypotheses=NH)
    tok, ln, cost = tok.cpu(), ln.cpu(), cost.cpu()
    same = finished = 0
    for i in range(n):
        init_fn, gen_fn = O.gru_tree_callbacks(p, feat[i])
        ref = O.beam_search_tree(init_fn, gen_fn, np.zeros((1, 1), np.int32), 1, end_id, K, NH, T)
        finished += len(ref)
        ours = [tok[i, j, :int(ln[i, j])].tolist() for j in range(NH) if int(ln[i, j]) > 0]
        if ours == [h.to_sequence_of_values() for h in ref]:
            same += 1
            for j, h in enumerate(ref):
                assert abs(float(cost[i, j]) - h.cum_cost) < 1e-3 * max(1.0, abs(h.cum_cost)), (i, j)
    print(f"tree beam-{K} ({gemm}), full size: {same}/{n} images with identical hypothesis lists, {finished} finished hypotheses")
    assert finished >= n and same >= 0.9 * n


@pytest.mark.parametrize("gemm", ["fp32", "tf32x3"])
@pytest.mark.parametrize("K", [3, 5])
def test_oracle_parity_beam_chain_full_size(K, gemm):
    dev = torch.device("cuda:0")
    m, feat, _, _ = _random_case("gru", 512, 512, 10000, 1, 6, 20, 13, False)
    m.decode_gemm = gemm
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    full, n = _check_chain(m.to(dev), p, feat.to(dev), K, max_len=20, eps=EPS[gemm])
    print(f"beam-{K} chain, full size: {full}/{n} rows separated in every round and bit-exact")
    assert full >= n // 2


@pytest.mark.parametrize("K", [0, 3])
def test_decode_tensor_core_gemm_agrees_with_fp32(K):
    """Bench-size check of decode_gemm="tf32x3": greedy / beam-3 captions of 512 images against the fp32
    CUDA-core path.  Differences can only come from rankings decided inside fp32 rounding noise."""
    dev = torch.device("cuda:0")
    m, feat, _, _ = _random_case("gru", 512, 512, 10000, 1, 512, 20, 17, False)
    m, feat = m.to(dev), feat.to(dev)
    a = m.sentence_index(feat, beam_size=K, max_len=20)
    m.decode_gemm = "tf32x3"
    b = m.sentence_index(feat, beam_size=K, max_len=20)
    same = int((a == b).all(dim=1).sum())
    print(f"decode_gemm tf32x3 vs fp32, K={K}: {same}/512 captions identical")
    assert same >= 500


def test_cpu_tensors_raise():
    from showtell_b200.rnn import RNN
    m = RNN(8, 8, 11, 1).cuda()
    with pytest.raises(RuntimeError):
        m(torch.randn(2, 8), torch.zeros(2, 3, dtype=torch.int64), [3, 2])
    with pytest.raises(RuntimeError):
        m(torch.randn(2, 8).cuda(), torch.zeros(2, 3, dtype=torch.int64).cuda(), [2, 3])   # unsorted


@pytest.mark.parametrize("kind,dtype", [("lstm", "bf16"), ("gru", "fp32")])
def test_cuda_graph_replay_matches_eager(kind, dtype):
    """forward_loss captures the step into a CUDA graph on the third identical call; replays must
    reproduce the eager gradients and follow changing inputs."""
    dev = torch.device("cuda:0")
    m, feat, cap, lengths = _random_case(kind, 64, 128, 500, 2, 20, 6, 3, True, dtype)
    m = m.to(dev)
    feat, cap = feat.to(dev), cap.to(dev)

    def run(f):
        m.zero_grad()
        loss = m.forward_loss(f, cap, lengths)
        loss.backward()
        return float(loss), {n: p.grad.clone() for n, p in m.named_parameters()}

    m.use_cuda_graphs = False
    l_ref, g_ref = run(feat)
    l_ref2, g_ref2 = run(feat * 0.5)
    m.use_cuda_graphs = True
    for i in range(5):                      # 2 eager, capture on the 3rd, then replays
        l, g = run(feat)
        assert l == pytest.approx(l_ref, rel=1e-6), i
        for n in g_ref:
            assert torch.allclose(g[n], g_ref[n], rtol=1e-5, atol=1e-8), (i, n)
    assert any(e.graph is not None for e in m._step_graphs.values()), "step was never captured"
    l, g = run(feat * 0.5)                  # replay with new input values
    assert l == pytest.approx(l_ref2, rel=1e-6)
    for n in g_ref2:
        assert torch.allclose(g[n], g_ref2[n], rtol=1e-5, atol=1e-8), n


@pytest.mark.parametrize("kind,dtype", [("lstm", "bf16"), ("gru", "fp32")])
def test_forward_backward_equals_forward_loss_backward(kind, dtype):
    """RNN.forward_backward: the fused iteration (loss + .grad of every parameter, no autograd round trip) gives the
    gradients of forward_loss(...).backward(), eagerly and when the step is replayed as a CUDA graph; the gradient
    w.r.t. cnn_feature continues into its producer."""
    dev = torch.device("cuda:0")
    m, feat, cap, lengths = _random_case(kind, 64, 128, 500, 2, 20, 6, 3, True, dtype)
    m = m.to(dev)
    cap = cap.to(dev)
    w = torch.randn(64, 64, device=dev, requires_grad=True)
    x = feat.to(dev)
    m.zero_grad()
    f1 = x @ w
    loss = m.forward_loss(f1, cap, lengths)
    loss.backward()
    ref = {n: p.grad.clone() for n, p in m.named_parameters()}
    ref_w, ref_loss = w.grad.clone(), float(loss)
    for i in range(5):                      # eager, eager, capture, replay, replay
        m.zero_grad()
        w.grad = None
        f2 = x @ w
        l2 = m.forward_backward(f2, cap, lengths)
        assert not l2.requires_grad and float(l2) == pytest.approx(ref_loss, rel=1e-6), i
        for n, p in m.named_parameters():
            assert torch.allclose(p.grad, ref[n], rtol=1e-5, atol=1e-8), (i, n)
        assert torch.allclose(w.grad, ref_w, rtol=1e-5, atol=1e-8), i


def test_stale_loss_raises_instead_of_returning_newer_gradients():
    """One loss at a time: once the step is replayed from a CUDA graph its gradients live in static
    storage; backward() of a loss whose gradients a later forward_loss overwrote raises."""
    dev = torch.device("cuda:0")
    m, feat, cap, lengths = _random_case("lstm", 64, 128, 500, 1, 12, 6, 4, True, "bf16")
    m, feat, cap = m.to(dev), feat.to(dev), cap.to(dev)
    for _ in range(4):
        m.zero_grad()
        m.forward_loss(feat, cap, lengths).backward()
    l1 = m.forward_loss(feat, cap, lengths)
    l2 = m.forward_loss(feat * 0.5, cap, lengths)
    with pytest.raises(RuntimeError, match="overwritten"):
        l1.backward()
    l2.backward()                           # the latest loss is fine

"""Optimizer step (main.py:97-100,152): showtell_b200.optim.SGD / Adam against torch.optim on the same
parameters and gradient sequences, and checkpoint (state_dict) interchange in both directions."""
import pytest
import torch

pytestmark = pytest.mark.gpu
SHAPES = [(10000, 512), (2048, 512), (2048,), (3, 5, 7), (1,), (4097,), (513, 3)]


def _params(seed, dev):
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(s, generator=g).to(dev).requires_grad_(True) for s in SHAPES]


def _set_grads(ps, qs, it, skip=()):
    g = torch.Generator().manual_seed(100 + it)
    for i, (p, q) in enumerate(zip(ps, qs)):
        gr = torch.randn(p.shape, generator=g).to(p.device) * (0.1 + i)
        if i in skip:
            p.grad = q.grad = None
        else:
            p.grad, q.grad = gr.clone(), gr.clone()


def _err(a, b):
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


@pytest.mark.parametrize("kind,kw", [("sgd", dict(lr=0.05, momentum=0.9)), ("sgd", dict(lr=0.01, momentum=0.0)),
                                     ("adam", dict(lr=1e-3)), ("adam", dict(lr=3e-2, betas=(0.8, 0.95), eps=1e-6))])
def test_matches_torch_optim(kind, kw):
    from showtell_b200 import optim
    dev = torch.device("cuda:0")
    ps, qs = _params(1, dev), _params(1, dev)
    ours = (optim.SGD if kind == "sgd" else optim.Adam)(ps, **kw)
    ref = (torch.optim.SGD if kind == "sgd" else torch.optim.Adam)(qs, **kw)
    for it in range(6):
        _set_grads(ps, qs, it, skip=(2,) if it == 0 else ())      # one tensor gets its first gradient a step late
        ours.step()
        ref.step()
        for p, q in zip(ps, qs):
            assert _err(p.detach(), q.detach()) < 2e-6, (kind, it, tuple(p.shape))
    so, sr = ours.state_dict(), ref.state_dict()
    assert set(so["state"]) == set(sr["state"])
    if not sr["state"]:
        return                                                # SGD without momentum keeps no state
    assert set(so["state"][0]) == set(sr["state"][0])
    for k in sr["state"][0]:
        assert _err(torch.as_tensor(so["state"][0][k]).float().cpu(), torch.as_tensor(sr["state"][0][k]).float().cpu()) < 2e-6, k


@pytest.mark.parametrize("kind", ["sgd", "adam"])
def test_checkpoint_interchange(kind):
    """utils.py:125-145 stores optimizer.state_dict(): a checkpoint of torch's optimizer resumes in ours and back."""
    from showtell_b200 import optim
    dev = torch.device("cuda:0")
    kw = dict(lr=0.02, momentum=0.8) if kind == "sgd" else dict(lr=2e-3)
    T = torch.optim.SGD if kind == "sgd" else torch.optim.Adam
    O = optim.SGD if kind == "sgd" else optim.Adam
    ps, qs = _params(2, dev), _params(2, dev)
    a, b = T(ps, **kw), T(qs, **kw)
    for it in range(2):
        _set_grads(ps, qs, it)
        a.step(); b.step()
    ours = O(ps, **kw)
    ours.load_state_dict(a.state_dict())                      # torch checkpoint -> ours
    for it in range(2, 4):
        _set_grads(ps, qs, it)
        ours.step(); b.step()
    for p, q in zip(ps, qs):
        assert _err(p.detach(), q.detach()) < 2e-6
    back = T(ps, **kw)
    back.load_state_dict(ours.state_dict())                   # ours -> torch
    for it in range(4, 6):
        _set_grads(ps, qs, it)
        back.step(); b.step()
    for p, q in zip(ps, qs):
        assert _err(p.detach(), q.detach()) < 2e-6


def test_training_loop_with_fused_step():
    """main_lstm.py's loop body with the drop-in decoder and optimizer: the loss goes down."""
    from showtell_b200 import optim
    from showtell_b200.rnn_lstm import RNN
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    rnn = RNN(64, 96, 211, 1).to(dev)
    opt = optim.Adam(rnn.parameters(), lr=1e-2)
    g = torch.Generator().manual_seed(4)
    feat = torch.randn(16, 64, generator=g).to(dev)
    cap = torch.randint(4, 211, (16, 9), generator=g).to(dev)
    lengths = [9] * 8 + [6] * 8
    losses = []
    for _ in range(25):
        opt.zero_grad()
        loss = rnn.forward_loss(feat, cap, lengths)
        loss.backward()
        opt.step()
        losses.append(float(loss))
    assert losses[-1] < 0.5 * losses[0], losses

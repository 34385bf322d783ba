"""Host logic of showtell_b200.graphs that needs no GPU: the eager path of run() with and without a prologue, the
default (copy) prologue and the one-loss-at-a-time ticket (stale static gradients must raise)."""
import pytest
import torch

from showtell_b200 import graphs


class _Mod:
    use_cuda_graphs = False


def test_run_eager_default_prologue_passes_the_inputs_through():
    m = _Mod()
    a, b = torch.arange(4.0), torch.ones(2)
    out = graphs.run(m, ("k",), lambda x, y: (x * 2, y + 1), (a, b))
    assert torch.equal(out[0], a * 2) and torch.equal(out[1], b + 1)
    assert m.__dict__["_last_run_static"] is False


def test_run_eager_custom_prologue_feeds_the_body_its_operands():
    m = _Mod()
    calls = []

    def prologue(inputs, out):
        calls.append(out)
        ops_ = [inputs[0] + 10, inputs[0].sum().reshape(1), inputs[1]]       # e.g. re-laid grid, its mean, the captions
        if out is None:
            return ops_
        for d, s in zip(out, ops_):
            d.copy_(s)
        return out

    got = graphs.run(m, ("k",), lambda F, mean, cap: (F, mean, cap), (torch.zeros(3), torch.tensor([7])), prologue=prologue)
    assert torch.equal(got[0], torch.full((3,), 10.0)) and float(got[1]) == 0.0 and int(got[2]) == 7
    assert calls == [None]                                                   # eager: fresh operands, no static buffers


def test_copy_prologue_writes_into_static_operands():
    src = (torch.arange(3.0), torch.tensor([5]))
    dst = [torch.zeros(3), torch.tensor([0])]
    out = graphs._copy_prologue(src, dst)
    assert out is dst and torch.equal(dst[0], src[0]) and int(dst[1]) == 5
    assert graphs._copy_prologue(src, None) == list(src)


def test_ticket_raises_when_static_gradients_were_overwritten():
    m = _Mod()
    m.__dict__["_last_run_static"] = True
    t1 = graphs.ticket(m)
    graphs.check_ticket(m, t1)                       # still current
    t2 = graphs.ticket(m)                            # a later forward_loss reuses the static storage
    graphs.check_ticket(m, t2)
    with pytest.raises(RuntimeError, match="overwritten"):
        graphs.check_ticket(m, t1)
    e = _Mod()                                       # eager single-GPU steps own fresh tensors: never stale
    e.__dict__["_last_run_static"] = False
    k1 = graphs.ticket(e)
    graphs.ticket(e)
    graphs.check_ticket(e, k1)

"""Encoder head (cnn.py:37-38,49): EncoderHead against nn.Linear + nn.BatchNorm1d(momentum=0.01) with the same
parameters -- outputs, running statistics, every gradient, eval mode, and end to end through the decoder."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _err(a, b):
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def _ref(head, dev):
    lin = nn.Linear(head.linear_secondlast_layer.in_features, head.linear_secondlast_layer.out_features)
    bn = nn.BatchNorm1d(head.last_layer.num_features, momentum=0.01)
    m = nn.Sequential()
    m.linear_secondlast_layer, m.last_layer = lin, bn
    m.load_state_dict(head.state_dict())                     # same names as the reference's ResNet sub-modules
    return m.to(dev)


@pytest.mark.parametrize("B,K,E,dtype,tol", [(32, 2048, 512, "fp32", 1e-4), (256, 2048, 512, "fp32", 1e-4), (7, 96, 40, "fp32", 1e-4),
                                             (256, 2048, 512, "bf16", 2e-2)])
def test_head_matches_torch(B, K, E, dtype, tol):
    from showtell_b200.cnn_head import EncoderHead
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    head = EncoderHead(K, E, dtype=dtype).to(dev)
    head.last_layer.weight.data.uniform_(0.5, 1.5)
    head.last_layer.bias.data.normal_(0, 0.1)
    ref = _ref(head, dev)
    g = torch.Generator().manual_seed(5)
    for it in range(3):                                      # running statistics evolve identically
        x = torch.relu(torch.randn(B, K, generator=g)).to(dev)
        w = torch.randn(B, E, generator=g).to(dev)
        for m in (head, ref):
            m.zero_grad()
        out = head(x)
        out_r = ref.last_layer(ref.linear_secondlast_layer(x))
        (out * w).sum().backward()
        (out_r * w).sum().backward()
        assert _err(out, out_r) < tol, (it, _err(out, out_r))
        pairs = [(head.linear_secondlast_layer.weight, ref.linear_secondlast_layer.weight),
                 (head.last_layer.weight, ref.last_layer.weight), (head.last_layer.bias, ref.last_layer.bias)]
        for a, b in pairs:
            assert _err(a.grad, b.grad) < tol, (it, tuple(a.shape), _err(a.grad, b.grad))
        # the Linear bias feeds a batch norm: its gradient is identically zero up to rounding in both
        assert float(head.linear_secondlast_layer.bias.grad.abs().max()) < 1e-3 * float(w.abs().max())
        assert _err(head.last_layer.running_mean, ref.last_layer.running_mean) < tol
        assert _err(head.last_layer.running_var, ref.last_layer.running_var) < tol
        assert int(head.last_layer.num_batches_tracked) == int(ref.last_layer.num_batches_tracked) == it + 1
    head.eval(); ref.eval()
    x = torch.relu(torch.randn(5, K, generator=g)).to(dev)
    with torch.no_grad():
        assert _err(head(x), ref.last_layer(ref.linear_secondlast_layer(x))) < tol


def test_head_feeds_decoder_and_single_row_raises():
    """main.py:146-150: cnn_feature = head(pooled); loss through the decoder reaches the head's parameters."""
    from showtell_b200.cnn_head import EncoderHead
    from showtell_b200.rnn import RNN
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    head, rnn = EncoderHead(128, 64).to(dev), RNN(64, 96, 101, 1).to(dev)
    ref = _ref(head, dev)
    g = torch.Generator().manual_seed(2)
    x = torch.relu(torch.randn(8, 128, generator=g)).to(dev)
    cap = torch.randint(4, 101, (8, 6), generator=g).to(dev)
    lengths = [6, 6, 5, 5, 4, 3, 3, 2]
    loss = rnn.forward_loss(head(x), cap, lengths)
    loss.backward()
    gw = head.linear_secondlast_layer.weight.grad.clone()
    rnn.zero_grad()
    loss_r = rnn.forward_loss(ref.last_layer(ref.linear_secondlast_layer(x)), cap, lengths)
    loss_r.backward()
    assert abs(float(loss) - float(loss_r)) < 1e-5 * abs(float(loss_r))
    assert _err(gw, ref.linear_secondlast_layer.weight.grad) < 1e-4
    with pytest.raises(ValueError):
        head(x[:1])

"""GPU: the collate contract of the reference's data loader (utils.py:61-77) on device tensors in any order --
stable length-descending order, re-padded captions, gathered features, batch_sizes -- against `create_batch`."""
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("B,T,feat_shape,dtype", [(1, 5, (8,), torch.float32), (7, 12, (16,), torch.float32),
                                                  (256, 20, (512,), torch.float32), (128, 25, (64, 49), torch.bfloat16),
                                                  (1000, 30, (3,), torch.float32)])
def test_sort_batch_matches_create_batch(B, T, feat_shape, dtype):
    from showtell_b200 import collate
    g = torch.Generator().manual_seed(B + T)
    lengths = torch.randint(1, T + 1, (B,), generator=g)           # many ties: the order among equals must be the input order
    if B > 2:
        lengths[B // 2] = T
    feats = torch.randn(B, *feat_shape, generator=g).to(dtype)
    caps = torch.zeros(B, T, dtype=torch.int64)
    for i in range(B):
        caps[i, :lengths[i]] = torch.randint(1, 1000, (int(lengths[i]),), generator=g)
    junk = caps.clone()
    for i in range(B):
        junk[i, lengths[i]:] = 999                                  # whatever sits behind a caption's length is replaced by the zero padding
    data = [("img%d" % i, feats[i], caps[i, :lengths[i]]) for i in range(B)]
    _, ref_f, ref_c, ref_len = collate.create_batch(data)           # utils.py:61-77 on the host
    f, c, lens, perm = collate.sort_batch(feats.to(DEV), junk.to(DEV), lengths.to(DEV))
    assert lens == ref_len
    assert torch.equal(c.cpu(), ref_c) and c.dtype == torch.int64
    assert torch.equal(f.cpu(), ref_f)
    assert torch.equal(perm.cpu(), torch.sort(lengths, descending=True, stable=True)[1])
    bs = collate.device_batch_sizes(lengths.to(DEV), T).cpu().tolist()
    assert bs == [int((lengths > t).sum()) for t in range(T)]
    # and the CPU path of the same function agrees
    f2, c2, lens2, perm2 = collate.sort_batch(feats, junk, lengths)
    assert lens2 == lens and torch.equal(perm2, perm.cpu()) and torch.equal(f2, ref_f)


def test_sort_batch_rejects_bad_lengths():
    from showtell_b200 import collate
    f = torch.zeros(3, 4, device=DEV)
    c = torch.zeros(3, 5, dtype=torch.int64, device=DEV)
    with pytest.raises(RuntimeError):
        collate.sort_batch(f, c, [3, 0, 2])
    with pytest.raises(ValueError):
        collate.sort_batch(f, c, [3, 6, 2])

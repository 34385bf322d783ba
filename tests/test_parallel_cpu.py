"""CPU, world_size 2 over gloo: the data-parallel arithmetic (global denominators + sum all-reduce
of gradient buckets) reproduces the single-process gradients of the concatenated batch.  The CPU
oracle plays the gradient producer here (the CUDA engine cannot run without a GPU); the code under
test is showtell_b200.parallel."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, kind, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import showtell_oracle as O
        from showtell_b200 import parallel
        torch.manual_seed(0)
        torch.set_num_threads(1)
        B, T, V, E, H = 6, 5, 23, 8, 12
        g = torch.Generator().manual_seed(3)
        lengths = [5, 5, 4, 3, 3, 2]
        cap = torch.randint(0, V, (B, T), generator=g)
        if kind == "gru":
            p = {"embeddings.weight": torch.randn(V, E, generator=g),
                 "unit.weight_ih_l0": torch.randn(3 * H, E, generator=g) * 0.3,
                 "unit.weight_hh_l0": torch.randn(3 * H, H, generator=g) * 0.3,
                 "unit.bias_ih_l0": torch.randn(3 * H, generator=g) * 0.1,
                 "unit.bias_hh_l0": torch.randn(3 * H, generator=g) * 0.1,
                 "linear.weight": torch.randn(V, H, generator=g) * 0.3, "linear.bias": torch.randn(V, generator=g) * 0.1}
            feat = torch.randn(B, E, generator=g)
        else:
            C, A, P = 10, 7, 4
            p = {"embeddings.weight": torch.randn(V, E, generator=g),
                 "unit.weight_ih_l0": torch.randn(3 * H, 2 * E, generator=g) * 0.3,
                 "unit.weight_hh_l0": torch.randn(3 * H, H, generator=g) * 0.3,
                 "unit.bias_ih_l0": torch.randn(3 * H, generator=g) * 0.1,
                 "unit.bias_hh_l0": torch.randn(3 * H, generator=g) * 0.1,
                 "linear.weight": torch.randn(V, H, generator=g) * 0.3, "linear.bias": torch.randn(V, generator=g) * 0.1,
                 "init_h.weight": torch.randn(H, C, generator=g) * 0.3, "init_h.bias": torch.randn(H, generator=g) * 0.1,
                 "attn.encoder_att.weight": torch.randn(A, C, generator=g) * 0.3,
                 "attn.encoder_att.bias": torch.randn(A, generator=g) * 0.1,
                 "attn.decoder_att.weight": torch.randn(A, H, generator=g) * 0.3,
                 "attn.decoder_att.bias": torch.randn(A, generator=g) * 0.1,
                 "attn.full_att.weight": torch.randn(1, A, generator=g) * 0.3,
                 "attn.full_att.bias": torch.randn(1, generator=g) * 0.1,
                 "embed.weight": torch.randn(E, C, generator=g) * 0.3, "embed.bias": torch.randn(E, generator=g) * 0.1}
            feat = torch.relu(torch.randn(B, C, P, generator=g))
        p = {k: v.double() for k, v in p.items()}
        feat = feat.double()
        model = kind if kind == "gru" else "attn_gru"
        full_loss, full_grads, _ = O.train_step(p, model, feat, cap, lengths, alpha_c=1.0)
        lo, hi = parallel.shard_rows(B, rank, world)
        n_loc = sum(lengths[lo:hi])
        n_glob, b_glob = parallel.global_counts(n_loc, hi - lo)
        assert (n_glob, b_glob) == (sum(lengths), B)
        # local sums scaled by the GLOBAL denominators: mean-loss(local) * n_loc / n_glob (+ penalty * b_loc / B)
        q_ = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        if kind == "gru":
            logits = O.rnn_forward(q_, "gru", feat[lo:hi], cap[lo:hi], lengths[lo:hi])
            loss = O.ce_loss(logits, cap[lo:hi], lengths[lo:hi]) * n_loc / n_glob
        else:
            logits, alphas = O.attn_forward(q_, "gru", feat[lo:hi], cap[lo:hi], lengths[lo:hi])
            loss = O.ce_loss(logits, cap[lo:hi], lengths[lo:hi]) * n_loc / n_glob \
                + ((1.0 - alphas.sum(dim=1)) ** 2).mean() * (hi - lo) / b_glob
        loss.backward()
        names = sorted(q_)
        grads = [q_[n].grad if q_[n].grad is not None else torch.zeros_like(q_[n]) for n in names]
        red = parallel.GradReducer()
        # three exchanges, as the engine issues them; reduce() hands back the reduced tensors (views of a
        # symmetric bucket on CUDA, the inputs themselves on the CPU / gloo path) and the engine uses THOSE
        out = red.reduce(grads[:2]) + red.reduce(grads[2:3]) + red.reduce(grads[3:])
        red.finish()
        assert len(out) == len(grads) and all(o.shape == g_.shape for o, g_ in zip(out, grads))
        assert red.slots([g_.shape for g_ in grads[:2]]) is None          # no symmetric buckets off the GPU
        grads = out
        lt = loss.detach().clone()
        dist.all_reduce(lt)
        err = max(float((a - full_grads[n]).abs().max() / (full_grads[n].abs().max() + 1e-30))
                  for a, n in zip(grads, names) if n != "attn.full_att.bias")
        q.put((rank, err, abs(float(lt) - float(full_loss))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("kind", ["gru", "attn"])
def test_dp_equals_single_process(kind):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, kind, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    for rank, err, dl in res:
        assert err < 1e-10, (rank, err)
        assert dl < 1e-10, (rank, dl)


def test_shard_rows_partition():
    from showtell_b200 import parallel
    for n, w in ((40000, 8), (10, 4), (7, 8), (256, 2)):
        spans = [parallel.shard_rows(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
        assert max(b - a for a, b in spans) - min(b - a for a, b in spans) <= 1

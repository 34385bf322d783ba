#!/usr/bin/env python
"""bench.py -- headline benchmark of the show-tell caption-decoder hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload lstm_train|gru_train|attn_gru_train|attn_lstm_train|beam3|beam5]
                    [--dtype fp32|bf16]

Default workload (BASELINE.json configs[1], the configuration the metric is quoted on that fits one
GPU): LSTM/rnn_lstm.py decoder training step -- forward + cross-entropy + backward -- with
E = H = 512, V = 10 000, batch 256 per GPU, caption length 20, L = 1, bf16 tensor-core arithmetic
with fp32 state/accumulation, synthetic N(0,1) features of the ResNet head's output shape and
random-init weights.  One "step" = one such iteration on one batch (5120 tokens per GPU).
Metric: training tokens/s, whole job.  The other workloads are the remaining BASELINE configs
(attention decoders on a 14x14x2048 grid, chain beam search), selectable for reporting.

Our arm times the repo's public API (forward_loss + backward, or sentence_index) with CUDA events
on the launching stream; `value` has the batch resident in HBM, `e2e` copies the step's inputs
from pinned host memory and reads the result back every step.  Between timed steps a 256 MiB
buffer is rewritten to flush the 126 MB L2 (outside the per-step event pairs).  N > 1: one process
per GPU (torchrun), batch-sharded (weak scaling), gradients all-reduced over NCCL inside the timed
step on a side stream overlapped with backward; time = max over ranks.

`--impl reference` times the reference algorithm on the host CPU cores (the oracle port: the
reference is pure Python over torch CPU kernels and cannot travel to the GPU box).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

E = H = A = 512
C = 2048
V = 10000
T = 20
WORKLOADS = {
    # name: (model, per-GPU batch, grid locations / beam width, description)
    "lstm_train": ("lstm", 256, 0, "rnn_lstm.py LSTM decoder train step fwd+CE+bwd, E=H=512 V=10000 B=256/GPU T=20 L=1"),
    "gru_train": ("gru", 32, 0, "rnn.py GRU decoder train step fwd+CE+bwd, E=H=512 V=10000 B=32/GPU T=20 L=1"),
    "attn_gru_train": ("attn_gru", 128, 196, "rnn_attn.py GRU+soft attention train step (CE + alpha_c=1 penalty) "
                       "fwd+bwd, 14x14x2048 grid, E=H=A=512 V=10000 B=128/GPU T=20 L=1"),
    "attn_lstm_train": ("attn_lstm", 512, 196, "rnn_attn_LSTM.py LSTM+soft attention train step fwd+bwd, "
                        "14x14x2048 grid, E=H=A=512 V=10000 B=512/GPU T=20 L=1"),
    "beam3": ("beam", 4096, 3, "rnn.py sentence_index(beam_size=3) chain beam search, GRU E=H=512 V=10000 L=1, "
              "max_len 20, 4096 images/GPU per step"),
    "beam5": ("beam", 4096, 5, "rnn.py sentence_index(beam_size=5) chain beam search, GRU E=H=512 V=10000 L=1, "
              "max_len 20, 4096 images/GPU per step"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler(threading.Thread):
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        self.t0 = self.t1 = None          # wall-clock window of the timed region

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append((time.time(), [x.strip() for x in line.split(",")]))
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        try:
            self.proc.terminate()
        except Exception:
            pass
        rows = [r for t, r in self.rows if self.t0 is None or (self.t0 - 0.05 <= t <= (self.t1 or t) + 0.05)]
        rows = rows or [r for _, r in self.rows]
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def make_batch(model, B, Pn, seed):
    g = torch.Generator().manual_seed(seed)
    if model == "beam":
        return torch.randn(B, E, generator=g), None, None
    feat = torch.relu(torch.randn(B, C, Pn, generator=g)) if model.startswith("attn") else torch.randn(B, E, generator=g)
    cap = torch.randint(4, V, (B, T), generator=g)
    cap[:, 0] = 1
    cap[:, -1] = 2
    return feat, cap, [T] * B


def build_model(model, dtype):
    from showtell_b200.rnn import RNN as GruRNN
    from showtell_b200.rnn_attn import RNN_Attn as AttnGru
    from showtell_b200.rnn_attn_LSTM import RNN_Attn as AttnLstm
    from showtell_b200.rnn_lstm import RNN as LstmRNN
    torch.manual_seed(1)
    if model == "gru" or model == "beam":
        return GruRNN(E, H, V, 1, dtype="fp32" if model == "beam" else dtype)
    if model == "lstm":
        return LstmRNN(E, H, V, 1, dtype=dtype)
    return (AttnGru if model == "attn_gru" else AttnLstm)(E, C, A, H, V, 1, dtype=dtype)


def train_flops(model, B, Pn):
    """Algorithmic FLOPs of one training step: 3 x forward."""
    g = 3 if "gru" in model else 4
    N = B * T
    if not model.startswith("attn"):
        return 3.0 * 2.0 * N * (g * H * (E + H) + H * V)
    step = 2.0 * B * (H * A + Pn * A + Pn * C + C * E + g * H * (E + H))
    hoist = 2.0 * B * Pn * C * A + 2.0 * N * g * H * E + 2.0 * B * C * H * (2 if g == 4 else 1)
    return 3.0 * (hoist + T * step + 2.0 * N * H * V) - 2.0 * B * Pn * C * A


def cpu_reference(model, B, Pn, steps, warmup, budget_s=150.0):
    """Reference algorithm on the host cores (oracle port, explicit-equation torch CPU ops, all
    threads).  Training: forward + loss + backward; beam: rnn.py chain beam, batch 1 per call as the
    reference requires.  Bounded: the per-step sample shrinks until K+W steps fit in budget_s."""
    from oracle import showtell_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = build_model(model, "fp32")
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    feat, cap, lengths = make_batch(model, B if model != "beam" else 64, Pn, 1)
    if model == "beam":
        K = Pn
        t0 = time.perf_counter()
        with torch.no_grad():
            O.rnn_beam_chain(p, feat[:1], K, T)
        per = time.perf_counter() - t0
        n = max(1, min(64, int(budget_s / max(per, 1e-3) / max(steps + warmup, 1))))
        for _ in range(warmup):
            with torch.no_grad():
                O.rnn_beam_chain(p, feat[:1], K, T)
        t0 = time.perf_counter()
        for s in range(steps):
            with torch.no_grad():
                for i in range(n):
                    O.rnn_beam_chain(p, feat[i:i + 1], K, T)
        dt = (time.perf_counter() - t0) / steps
        return {"value": n / dt, "unit": "captions/s", "cores": cores, "kind": "port", "ms_per_step": dt * 1e3,
                "sample": f"{steps} steps of {n} images, beam {K}, max_len {T}, one image per call (torch CPU {torch.__version__})"}
    probe = min(B, 16)
    t0 = time.perf_counter()
    O.train_step(p, model, feat[:probe], cap[:probe], lengths[:probe])
    per = (time.perf_counter() - t0) / probe
    bs = B
    while bs > probe and per * bs * (steps + warmup) > budget_s:
        bs //= 2
    f, c, l = feat[:bs], cap[:bs], lengths[:bs]
    for _ in range(warmup):
        O.train_step(p, model, f, c, l)
    t0 = time.perf_counter()
    for _ in range(steps):
        O.train_step(p, model, f, c, l)
    dt = (time.perf_counter() - t0) / steps
    return {"value": bs * T / dt, "unit": "tokens/s", "cores": cores, "kind": "port", "ms_per_step": dt * 1e3,
            "sample": f"{steps} steps of {bs}/{B} rows x {T} tokens (fwd+loss+bwd, fp32, torch CPU {torch.__version__})"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lstm_train", choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--decode-gemm", default="tf32x3", choices=["fp32", "tf32x3"],
                    help="beam workloads: nn.Linear products on CUDA cores (fp32) or fp32-accurate 3xTF32 tensor cores")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--features", default="fp32", choices=["fp32", "bf16"],
                    help="attention workloads: dtype of the (B, 2048, P) grid handed to the decoder (reference: fp32)")
    ap.add_argument("--optimizer", default="none", choices=["none", "sgd", "adam"],
                    help="training workloads: also run the fused optimizer step (main.py:152) inside the timed step")
    args = ap.parse_args()
    model, B, Pn, desc = WORKLOADS[args.workload]
    is_beam = model == "beam"
    metric, unit = ("beam_captions_per_s", "captions/s") if is_beam else ("train_tokens_per_s", "tokens/s")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)

    if args.impl == "reference":
        if rank != 0:
            return 0
        w = min(args.warmup, 2)
        r = cpu_reference(model, B, Pn, args.steps, w)
        line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": w, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc + " [reference algorithm on host CPU]"},
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import torch.distributed as dist
    from showtell_b200 import _lib, ops, parallel

    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    dev = torch.device(f"cuda:{local}")
    lib = _lib.load()
    dtype = "fp32" if is_beam else args.dtype
    net = build_model(model, dtype).to(dev)
    if is_beam:
        net.decode_gemm = args.decode_gemm
    params = [p for p in net.parameters()]
    opt = None
    if args.optimizer != "none" and not is_beam:
        from showtell_b200 import optim
        opt = optim.Adam(params, lr=1e-4) if args.optimizer == "adam" else optim.SGD(params, lr=1e-3, momentum=0.9)
    if world > 1 and not is_beam:
        net.grad_reducer = parallel.GradReducer()          # NCCL all-reduce on a side stream
    feat_h, cap_h, lengths = make_batch(model, B, Pn, 1 + rank)
    if args.features == "bf16" and model.startswith("attn"):
        feat_h = feat_h.bfloat16()
    feat_p = feat_h.pin_memory()
    cap_p = cap_h.pin_memory() if cap_h is not None else None
    feat_d = feat_h.to(dev)
    cap_d = cap_h.to(dev) if cap_h is not None else None
    units_per_step = (B if is_beam else B * T) * world
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out_host = (torch.empty(B, T, dtype=torch.int64) if is_beam else torch.empty((), dtype=torch.float32)).pin_memory()

    def step(f, c):
        if is_beam:
            return net.sentence_index(f, beam_size=Pn, max_len=T)                # utils.py:194
        for p in params:
            p.grad = None
        if model.startswith("attn"):
            loss, _ = net.forward_loss(f, c, lengths, alpha_c=1.0, global_tokens=units_per_step,
                                       global_batch=B * world)
        else:
            loss = net.forward_loss(f, c, lengths, global_tokens=units_per_step)
        loss.backward()
        if opt is not None:
            opt.step()                                                            # main.py:152
        return loss

    def step_e2e():
        f = feat_p.to(dev, non_blocking=True)
        c = cap_p.to(dev, non_blocking=True) if cap_p is not None else None
        out = step(f, c)
        out_host.copy_(out.detach(), non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K):
        evs = []
        barrier()
        for _ in range(K):
            flush.fill_(1)                       # L2 flush, outside the event pair
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            evs.append((e0, e1))
        barrier()
        total = sum(a.elapsed_time(b) for a, b in evs)
        if world > 1:
            t = torch.tensor([total], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t)
        return total / K

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(warmup):
        step(feat_d, cap_d)
    if sampler:
        sampler.t0 = time.time()
    ms = timed(lambda: step(feat_d, cap_d), args.steps)     # CUDA-graph replay after the warm-up steps
    # per-kernel event timing + launch count: the same step issued eagerly (graphs are bypassed while
    # ops.TIMER is set), right after the timed region, same process, same buffers
    ops.TIMER = ops.KernelTimer()
    l0 = lib.st_launch_count()
    ksteps = max(3, min(args.steps, 10))
    timed(lambda: step(feat_d, cap_d), ksteps)
    launches = (lib.st_launch_count() - l0) // ksteps
    ksum = ops.TIMER.summary()
    ops.TIMER = None
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    if sampler:
        sampler.t1 = time.time()
    clocks = sampler.finish() if sampler else None

    if rank == 0:
        pk = peaks()
        kms = {k: round(v[1], 4) for k, v in ksum.items()}
        if model.startswith("attn"):
            # dominant per-step kernel: fused attention (HBM/L2-bound): reads att1 (B,P,A) + Fe (B,P,E)
            esz = 2 if dtype == "bf16" else 4
            bytes_step = B * Pn * (A + E) * esz
            tags = [t for t in ("attn_fwd", "attn_bwd") if t in ksum]
            k_ms = sum(ksum[t][1] for t in tags) / max(len(tags), 1)
            ach = bytes_step / (k_ms * 1e-3) / 1e9 if tags else None
            roof = {"bound": "hbm", "kernel": "fused attention step (fwd / bwd mean), algorithmic bytes = B*P*(A+E)*sizeof",
                    "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"] if ach else None,
                    "traffic": None, "peak_source": pk["src"] + " (copy bandwidth)"}
        elif is_beam:
            # whole decode loop: (1 + K (T-1)) dependent steps per caption, each {GRU gates 2*3H(E+H), vocabulary
            # projection 2HV} FLOPs.  The products run as 3xTF32 (three tf32 MMAs per fp32-accurate
            # product, tf32 at half the bf16 rate), so 1/6 of the bf16 peak is the most this arithmetic can reach.
            flops_caption = (1 + Pn * (T - 1)) * (2.0 * 3 * H * (E + H) + 2.0 * H * V)
            ach = flops_caption * units_per_step / world / (ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "decode loop (gate + vocabulary products per dependent step, "
                    + ("3xTF32 tensor-core" if args.decode_gemm == "tf32x3" else "fp32 CUDA-core") + " GEMMs), algorithmic "
                    "FLOPs = (1 + K(T-1)) * (2*3H(E+H) + 2HV) per caption; fp32-accurate 3xTF32 can reach at most peak/6",
                    "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                    "traffic": None, "peak_source": pk["src"] + " (bf16 cuBLAS sustained)"}
        else:
            # dominant kernels: the vocabulary-projection GEMMs (each 2*N*H*V FLOPs)
            vocab_flops = 2.0 * B * T * H * V
            tags = [t for t in ("vocab_fwd", "vocab_dlogits", "vocab_dw", "vocab_dx") if t in ksum]
            k_ms = sum(ksum[t][1] for t in tags) / max(len(tags), 1)
            ach = vocab_flops / (k_ms * 1e-3) / 1e12 if tags else None
            roof = {"bound": "tensor", "kernel": "vocabulary projection GEMMs (" + " / ".join(tags) + ", mean; each 2*N*H*V FLOPs)",
                    "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                    "frac": ach / pk["tf_sustained"] if ach else None, "traffic": None,
                    "peak_source": pk["src"] + " (bf16 cuBLAS sustained)"}
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):                         # measured once per round under ncu --set full
            ent = json.load(open(tp)).get(args.workload)
            if ent and dtype == "bf16":
                roof["traffic"] = ent["traffic_bytes_per_launch"]
                roof["traffic_source"] = ent["source"]
        roof["kernels_ms"] = kms
        if not is_beam:
            roof["step_flops"] = train_flops(model, B, Pn)
            roof["step_tflops"] = roof["step_flops"] / (ms * 1e-3) / 1e12
        h2d = feat_p.numel() * feat_p.element_size() + (cap_p.numel() * 8 if cap_p is not None else 0)
        line = {"metric": metric, "value": units_per_step / (ms * 1e-3), "unit": unit, "n_gpus": world,
                "steps": args.steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32" if dtype == "fp32" else "bf16",
                "data": "synthetic",
                "config": {"workload": desc, "global_batch": B * world, "seq_len": T, "parallelism": f"dp{world}",
                           "l2": "flushed between timed steps (256 MiB fill, outside the event pairs)",
                           "launch": "whole step replayed as one CUDA graph (captured on the 3rd identical step)",
                           "features": ("bf16 grid (autocast trunk)" if feat_h.dtype == torch.bfloat16 else "fp32 (as the reference's encoder emits them)"),
                           "optimizer": "excluded" if opt is None else args.optimizer + " step (fused, one launch) included"},
                "e2e": {"value": units_per_step / (ms_e2e * 1e-3), "unit": unit, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": out_host.numel() * out_host.element_size(), "ms_per_step": ms_e2e},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof}
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference(model, B, Pn, 5, 1, budget_s=25.0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        # Tear down in dependency order: captured step graphs hold NCCL work, so they go first; a
        # communicator teardown that still stalls must not keep torchrun alive (watchdog exit).
        sys.stdout.flush()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        net.__dict__.pop("_step_graphs", None)
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""bench.py -- headline benchmark of the show-tell caption-decoder hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload lstm_train|gru_train|attn_gru_train|attn_lstm_train|beam3|beam5]
                    [--dtype fp32|bf16] [--no-extras]

Default workload (BASELINE.json configs[1], the configuration the metric is quoted on that fits one
GPU): LSTM/rnn_lstm.py decoder training step -- forward + cross-entropy + backward -- with
E = H = 512, V = 10 000, batch 256 per GPU, caption length 20, L = 1, bf16 tensor-core arithmetic
with fp32 state/accumulation, synthetic N(0,1) features of the ResNet head's output shape and
random-init weights.  One "step" = one such iteration on one batch (5120 tokens per GPU).
Metric: training tokens/s, whole job.  Without an explicit --workload the same run also measures
the north-star's other two headline workloads at the same N (attention-GRU training, configs[2];
beam-3 decoding, configs[4]) and reports them under "others" in the one JSON line.

Our arm times the repo's public API (forward_backward = forward + loss + backward, or sentence_index) with CUDA events
on the launching stream; `value` has the batch resident in HBM, `e2e` copies every step's inputs
from pinned host memory (double-buffered on a copy stream, so step i+1's copy runs under step i's
kernels) and reads the result back every step.  Between timed steps a 256 MiB buffer is rewritten
to flush the 126 MB L2 (outside the per-step event pairs).  N > 1: one process per GPU (torchrun),
batch-sharded (weak scaling), gradients all-reduced inside the timed step on a side stream
overlapped with backward; time = max over ranks.

Reference bars, all the UNMODIFIED reference modules from baseline/_ref (baseline/reference.py):
  * `--impl reference`: on the host CPU cores, same batch / lengths / steps as our arm (a config that
    is too slow runs fewer STEPS, never fewer rows);
  * `cpu_baseline` (our arm, N = 1): the same, a bounded number of steps;
  * `gpu_reference` (our arm, N = 1): the same modules in torch-eager on the same B200 (cuDNN RNN +
    cuBLAS + ATen), fp32 and under bf16 autocast -- the existing Blackwell kernels to beat.
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
              "4096 images/GPU per step"),
    "beam5": ("beam", 4096, 5, "rnn.py sentence_index(beam_size=5) chain beam search, GRU E=H=512 V=10000 L=1, "
              "4096 images/GPU per step"),
}
EXTRAS = ["attn_gru_train", "beam3"]      # measured beside the default workload


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
        self.proc = None

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


# ------------------------------------------------------------------------------------------------ reference arms
def _reference_module(model):
    """Unmodified reference module (baseline/_ref), reference default init under seed 1 (main.py:26-27)."""
    from baseline import reference as R
    torch.manual_seed(1)
    return R, R.make_module(model, E, H, V, 1, C, A)


def cpu_reference(model, B, Pn, steps, warmup, budget_s=150.0, max_len=25):
    """The reference's own modules on the host cores, all threads, same batch as our arm.  Training:
    forward + loss + backward as main.py:145-151 / main_attn.py:126-133; beam: rnn.py beam search, one image
    per call as the reference requires (main.py:81-82), 25 steps as it hard-codes (rnn.py:39).  A slow
    configuration runs fewer timed STEPS (never fewer rows): timing stops once budget_s is spent (>= 2 steps)."""
    from baseline import reference as R
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if not R.available():
        return cpu_port(model, B, Pn, steps, warmup, budget_s)
    R, net = _reference_module(model)
    n_img = 16
    feat, cap, lengths = make_batch(model, B if model != "beam" else n_img, Pn, 1)
    with R.cpu_cuda_shim():
        if model == "beam":
            run = lambda: R.beam_captions(net, feat, Pn)
            units = n_img
        else:
            run = lambda: R.train_step(net, model, feat, cap, lengths, 1.0)
            units = B * T
        t_start = time.perf_counter()
        for _ in range(warmup):
            run()
            if time.perf_counter() - t_start > budget_s / 3:
                break
        done, t0 = 0, time.perf_counter()
        while done < steps:
            run()
            done += 1
            if done >= 2 and time.perf_counter() - t_start > budget_s:
                break
        dt = (time.perf_counter() - t0) / done
    what = (f"{done} steps of {n_img} images, beam {Pn}, 25 tokens (rnn.py:39), one image per call" if model == "beam"
            else f"{done} steps of {B}/{B} rows x {T} tokens (fwd+loss+bwd)")
    return {"value": units / dt, "unit": "captions/s" if model == "beam" else "tokens/s", "cores": cores,
            "kind": "reference", "ms_per_step": dt * 1e3, "steps_done": done,
            "sample": what + f", unmodified reference modules (baseline/_ref), fp32, torch CPU {torch.__version__}"}


def cpu_port(model, B, Pn, steps, warmup, budget_s):
    """Fallback when baseline/_ref is absent (a checkout that never saw /root/reference): the oracle port."""
    from oracle import showtell_oracle as O
    cores = os.cpu_count() or 1
    m = build_model(model, "fp32")
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    n_img = 16
    feat, cap, lengths = make_batch(model, B if model != "beam" else n_img, Pn, 1)
    if model == "beam":
        def run():
            with torch.no_grad():
                for i in range(n_img):
                    O.rnn_beam_chain(p, feat[i:i + 1], Pn, 25)
        units = n_img
    else:
        run = lambda: O.train_step(p, model, feat, cap, lengths)
        units = B * T
    t_start = time.perf_counter()
    for _ in range(min(warmup, 1)):
        run()
    done, t0 = 0, time.perf_counter()
    while done < steps:
        run()
        done += 1
        if done >= 2 and time.perf_counter() - t_start > budget_s:
            break
    dt = (time.perf_counter() - t0) / done
    return {"value": units / dt, "unit": "captions/s" if model == "beam" else "tokens/s", "cores": cores,
            "kind": "port", "ms_per_step": dt * 1e3, "steps_done": done,
            "sample": f"{done} steps, oracle port (baseline/_ref not installed), fp32, torch CPU {torch.__version__}"}


def gpu_reference(model, B, Pn, dev, steps=10, warmup=3):
    """The same unmodified reference modules in torch-eager on this B200 (cuDNN RNN + cuBLAS + ATen kernels:
    the existing Blackwell path), inputs resident, CUDA-event timed.  fp32 = torch defaults (what the reference's
    mains run); bf16 = the same modules under torch.autocast(bfloat16).  Beam: batch 1 per call, 25 steps."""
    from baseline import reference as R
    if not R.available():
        return {"unavailable": "baseline/_ref not installed"}
    out = {"what": "unmodified reference modules, torch-eager on the same GPU (torch %s, cuDNN %s)"
           % (torch.__version__, torch.backends.cudnn.version())}
    n_img = 8
    feat, cap, lengths = make_batch(model, B if model != "beam" else n_img, Pn, 1)
    feat = feat.to(dev)
    cap = cap.to(dev) if cap is not None else None
    for mode in ("fp32", "bf16_autocast"):
        try:
            R_, net = _reference_module(model)
            net = net.to(dev)
            ctx = torch.autocast("cuda", dtype=torch.bfloat16) if mode != "fp32" else torch.autocast("cuda", enabled=False)
            if model == "beam":
                run = lambda: R.beam_captions(net, feat, Pn)
                units = n_img
            else:
                run = lambda: R.train_step(net, model, feat, cap, lengths, 1.0)
                units = B * T
            with ctx:
                for _ in range(warmup):
                    run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(steps):
                    run()
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[mode] = {"value": units / (ms * 1e-3), "ms_per_step": ms, "steps": steps,
                         "units_per_step": units}
            del net
            torch.cuda.empty_cache()
        except Exception as exc:                      # a library path that does not exist for this dtype / shape
            out[mode] = {"error": f"{type(exc).__name__}: {str(exc)[:160]}"}
    out["unit"] = "captions/s" if model == "beam" else "tokens/s"
    return out


# ------------------------------------------------------------------------------------------------------ our arm
class Prefetcher:
    """Double-buffered host->device input copies on a copy stream: while step i computes, step i+1's inputs
    travel.  Every step's inputs are copied from pinned host memory inside the timed region."""

    def __init__(self, host_tensors, dev):
        self.host = [t.pin_memory() if t is not None else None for t in host_tensors]
        self.dev_bufs = [[torch.empty_like(t, device=dev) if t is not None else None for t in self.host] for _ in range(2)]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [None, None]
        self.free = [None, None]
        self.i = 0

    def issue(self):
        k = self.i & 1
        main = torch.cuda.current_stream()
        with torch.cuda.stream(self.copy_stream):
            if self.free[k] is not None:
                self.copy_stream.wait_event(self.free[k])          # the step that read this buffer is done with it
            for h, d in zip(self.host, self.dev_bufs[k]):
                if h is not None:
                    d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.copy_stream)
        self.ready[k] = ev
        self.i += 1
        return k

    def take(self, k):
        torch.cuda.current_stream().wait_event(self.ready[k])
        return self.dev_bufs[k]

    def release(self, k):
        ev = torch.cuda.Event()
        ev.record()
        self.free[k] = ev

    def bytes_per_step(self):
        return sum(t.numel() * t.element_size() for t in self.host if t is not None)


def run_ours(name, args, steps, warmup, rank, world, local, want_refs):
    import torch.distributed as dist
    from showtell_b200 import _lib, ops, parallel
    model, B, Pn, desc = WORKLOADS[name]
    is_beam = model == "beam"
    max_len = args.max_len if is_beam else T
    if is_beam:
        desc += f", max_len {max_len}"
    metric, unit = ("beam_captions_per_s", "captions/s") if is_beam else ("train_tokens_per_s", "tokens/s")
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
        net.grad_reducer = parallel.GradReducer()          # gradient exchange on a side stream
    feat_h, cap_h, lengths = make_batch(model, B, Pn, 1 + rank)
    if args.features == "bf16" and model.startswith("attn"):
        feat_h = feat_h.bfloat16()
    feat_d = feat_h.to(dev)
    cap_d = cap_h.to(dev) if cap_h is not None else None
    units_per_step = (B if is_beam else B * T) * world
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out_host = (torch.empty(B, max_len, dtype=torch.int64) if is_beam else torch.empty((), dtype=torch.float32)).pin_memory()

    def step(f, c):
        if is_beam:
            return net.sentence_index(f, beam_size=Pn, max_len=max_len)          # utils.py:194
        # the training iteration of main.py:146-151 / main_attn.py:127-133 (zero_grad, forward, loss, backward) through
        # the library's fused entry point: loss + .grad of every parameter in one call
        if model.startswith("attn"):
            loss, _ = net.forward_backward(f, c, lengths, alpha_c=1.0, global_tokens=units_per_step,
                                           global_batch=B * world)
        else:
            loss = net.forward_backward(f, c, lengths, global_tokens=units_per_step)
        if opt is not None:
            opt.step()                                                            # main.py:152
        return loss

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

    def timed_e2e(K, feats=None):
        """K steps, each with its own H2D copy (prefetched one step ahead on the copy stream) and D2H read of the
        result; one event pair around the whole pipelined region (the first copy is not hidden)."""
        pf = Prefetcher([feat_h if feats is None else feats, cap_h], dev)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        k = pf.issue()
        for i in range(K):
            f, c = pf.take(k)
            k_next = pf.issue() if i + 1 < K else None
            out = step(f, c)
            pf.release(k)
            out_host.copy_(out.detach(), non_blocking=True)
            k = k_next
        e1.record()
        barrier()
        total = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([total], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            total = float(t)
        return total / K, pf.bytes_per_step()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(warmup):
        step(feat_d, cap_d)
    if sampler:
        sampler.t0 = time.time()
    ms = timed(lambda: step(feat_d, cap_d), steps)     # CUDA-graph replay after the warm-up steps
    # per-kernel event timing + launch count: the same step issued eagerly (graphs are bypassed while
    # ops.TIMER is set), right after the timed region, same process, same buffers
    ops.TIMER = ops.KernelTimer()
    l0 = lib.st_launch_count()
    ksteps = max(3, min(steps, 10))
    timed(lambda: step(feat_d, cap_d), ksteps)
    launches = (lib.st_launch_count() - l0) // ksteps
    ksum = ops.TIMER.summary()
    ops.TIMER = None
    timed_e2e(2)
    ms_e2e, h2d = timed_e2e(steps)
    e2e_bf16 = None
    if model.startswith("attn") and feat_h.dtype == torch.float32:
        # the same end-to-end step with the grid handed over as bf16 (a trunk run under autocast emits it): half the
        # PCIe bytes of the step's largest input; st_attn_relayout_bf16in consumes it directly
        f16 = feat_h.bfloat16().pin_memory()
        timed_e2e(4, f16)                       # a new step signature: eager runs + graph capture happen here
        ms16, h2d16 = timed_e2e(steps, f16)
        e2e_bf16 = {"value": units_per_step / (ms16 * 1e-3), "unit": unit, "h2d_bytes_per_step": h2d16,
                    "d2h_bytes_per_step": out_host.numel() * out_host.element_size(), "ms_per_step": ms16}
    if sampler:
        sampler.t1 = time.time()
    clocks = sampler.finish() if sampler else None

    line = None
    if rank == 0:
        pk = peaks()
        kms = {k: round(v[1] * v[0] / ksteps, 4) for k, v in ksum.items()}    # ms per STEP spent under each tag
        kcalls = {k: v[0] // ksteps for k, v in ksum.items()}
        esz = 2 if dtype == "bf16" else 4
        vocab_flops = 2.0 * B * T * H * V
        g = 3 if "gru" in model else 4
        # algorithmic work per launch of each tagged kernel: (flops or bytes, bound)
        work = {"vocab_fwd": (vocab_flops, "tensor"), "vocab_bwd": (2 * vocab_flops, "tensor"),
                "vocab_dlogits": (vocab_flops, "tensor"), "vocab_dw": (vocab_flops, "tensor"),
                "vocab_dx": (vocab_flops, "tensor"),
                "ih_fwd": (2.0 * B * T * g * H * E, "tensor"), "ih_dx": (2.0 * B * T * g * H * E, "tensor"),
                "ih_dw": (2.0 * B * T * g * H * E, "tensor"), "hh_dw": (2.0 * B * T * g * H * H, "tensor"),
                "seq_fwd": (2.0 * B * T * g * H * H, "tensor"), "seq_bwd": (2.0 * B * T * g * H * H, "tensor"),
                "att1_fwd": (2.0 * B * Pn * C * A, "tensor"), "fe_fwd": (2.0 * B * Pn * C * E, "tensor"),
                "att1_dw": (2.0 * B * Pn * C * A, "tensor"), "embed_dw": (2.0 * B * Pn * C * E, "tensor"),
                "attn_fwd": (float(B * Pn * (A + E) * esz), "hbm"), "attn_bwd": (float(B * Pn * (A + E) * esz), "hbm")}
        table = {}
        for k, (n, mean_ms) in ksum.items():
            if k in work and mean_ms > 0:
                w, bound = work[k]
                ach = w / (mean_ms * 1e-3) / (1e12 if bound == "tensor" else 1e9)
                peak = pk["tf_burst"] if bound == "tensor" else pk["hbm_gbs"]
                table[k] = {"ms": round(mean_ms, 4), "launches_per_step": n // ksteps, "bound": bound,
                            "achieved": round(ach, 1), "frac": round(ach / peak, 3)}
        if is_beam:
            # whole decode loop: (1 + K (T-1)) dependent steps per caption, each {GRU gates 2*3H(E+H), vocabulary
            # projection 2HV} FLOPs.  tf32x3 mode: the recurrent product runs as 3xTF32 (three tf32 MMAs per
            # fp32-accurate product), the input projection is a row of a table built once per call, and the vocabulary
            # projection is a bf16 screening GEMM + exact fp32 re-scoring of the surviving columns (decode.cu) -- the
            # ALGORITHMIC FLOPs below are the reference's, whatever arithmetic produced the bit-exact tokens.
            flops_caption = (1 + Pn * (max_len - 1)) * (2.0 * 3 * H * (E + H) + 2.0 * H * V)
            ach = flops_caption * B / (ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "decode loop (gate + vocabulary products per dependent step; "
                    + ("tensor cores: 3xTF32 recurrent product, bf16-screened + fp32-re-scored vocabulary top-K"
                       if args.decode_gemm == "tf32x3" else "fp32 CUDA-core GEMMs") + "), algorithmic "
                    "FLOPs = (1 + K(T-1)) * (2*3H(E+H) + 2HV) per caption",
                    "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                    "traffic": None, "peak_source": pk["src"] + " (bf16 cuBLAS sustained: kernel timed inside a long step)"}
        else:
            # the dominant kernel = the tag with the largest share of the step; its own algorithmic work per launch
            # over its event-timed mean duration, against the BURST peak (a kernel timed alone between events)
            dom = max((k for k in kms if k in table), key=lambda k: kms[k], default=None)
            if dom is not None:
                d = table[dom]
                roof = {"bound": d["bound"], "kernel": f"{dom} (largest share of the step: {kms[dom]:.3f} of {ms:.3f} ms, "
                        f"{d['launches_per_step']} launch(es) per step)", "achieved": d["achieved"],
                        "peak": pk["tf_burst"] if d["bound"] == "tensor" else pk["hbm_gbs"],
                        "unit": "TFLOP/s" if d["bound"] == "tensor" else "GB/s", "frac": d["frac"], "traffic": None,
                        "peak_source": pk["src"] + (" (bf16 cuBLAS burst: kernel timed alone)" if d["bound"] == "tensor"
                                                    else " (copy bandwidth)")}
            else:
                roof = {"bound": "tensor", "kernel": None, "achieved": None, "peak": pk["tf_burst"], "unit": "TFLOP/s",
                        "frac": None, "traffic": None, "peak_source": pk["src"]}
            roof["step_flops"] = train_flops(model, B, Pn)
            roof["step_tflops"] = roof["step_flops"] / (ms * 1e-3) / 1e12
            roof["step_frac_of_sustained_peak"] = roof["step_tflops"] / pk["tf_sustained"]
        tp = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tp):                         # measured once per round under ncu --set full
            ent = json.load(open(tp)).get(name)
            if ent and dtype == "bf16":
                roof["traffic"] = ent["traffic_bytes_per_launch"]
                roof["traffic_source"] = ent["source"]
        roof["kernels_ms_per_step"] = kms
        roof["kernels"] = table
        line = {"metric": metric, "value": units_per_step / (ms * 1e-3), "unit": unit, "n_gpus": world,
                "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32" if dtype == "fp32" else "bf16",
                "data": "synthetic",
                "config": {"workload": desc, "global_batch": B * world, "seq_len": max_len, "parallelism": f"dp{world}",
                           "l2": "flushed between timed steps (256 MiB fill, outside the event pairs)",
                           "launch": "whole step replayed as one CUDA graph (captured on the 3rd identical step)",
                           "features": ("bf16 grid (autocast trunk)" if feat_h.dtype == torch.bfloat16 else "fp32 (as the reference's encoder emits them)"),
                           "optimizer": "excluded" if opt is None else args.optimizer + " step (fused, one launch) included"},
                "e2e": {"value": units_per_step / (ms_e2e * 1e-3), "unit": unit, "h2d_bytes_per_step": h2d,
                        "d2h_bytes_per_step": out_host.numel() * out_host.element_size(), "ms_per_step": ms_e2e,
                        "how": "double-buffered H2D on a copy stream, one event pair around all steps"},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roof}
        if e2e_bf16 is not None:
            line["e2e_bf16_features"] = e2e_bf16
    # free this workload's device memory before the reference legs / the next workload
    net.__dict__.pop("_step_graphs", None)
    del net, params, flush, feat_d, cap_d
    import gc
    gc.collect()
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and want_refs:
        line["gpu_reference"] = gpu_reference(model, B, Pn, dev)
        g32 = line["gpu_reference"].get("fp32", {}).get("value")
        g16 = line["gpu_reference"].get("bf16_autocast", {}).get("value")
        best = max([x for x in (g32, g16) if x], default=None)
        line["gpu_reference"]["ours_over_best_eager"] = (line["value"] / best) if best else None
        if not args.no_cpu_baseline:
            r = cpu_reference(model, B, Pn, 5, 1, budget_s=25.0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--dtype", default="bf16", choices=["fp32", "bf16"])
    ap.add_argument("--decode-gemm", default="tf32x3", choices=["fp32", "tf32x3"],
                    help="beam workloads: nn.Linear products on CUDA cores (fp32) or fp32-accurate 3xTF32 tensor cores")
    ap.add_argument("--max-len", type=int, default=20, help="beam workloads: caption length (BASELINE configs[4]: 20; "
                    "the reference hard-codes 25)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-reference", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="measure only the main workload")
    ap.add_argument("--features", default="fp32", choices=["fp32", "bf16"],
                    help="attention workloads: dtype of the (B, 2048, P) grid handed to the decoder (reference: fp32)")
    ap.add_argument("--optimizer", default="none", choices=["none", "sgd", "adam"],
                    help="training workloads: also run the fused optimizer step (main.py:152) inside the timed step")
    args = ap.parse_args()
    main_wl = args.workload or "lstm_train"
    extras = [] if (args.workload is not None or args.no_extras) else EXTRAS
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    warmup = max(args.warmup, 3)

    if args.impl == "reference":
        if rank != 0:
            return 0
        out = None
        for name in [main_wl] + extras:
            model, B, Pn, desc = WORKLOADS[name]
            is_beam = model == "beam"
            first = out is None
            # same batch, same steps, same warm-up as our arm for the headline workload; the slow extras are
            # bounded by time (fewer steps, never fewer rows)
            r = cpu_reference(model, B, Pn, args.steps if first else 3, warmup if first else 1,
                              budget_s=240.0 if first else 60.0)
            metric, unit = ("beam_captions_per_s", "captions/s") if is_beam else ("train_tokens_per_s", "tokens/s")
            line = {"impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": args.gpus,
                    "steps": r["steps_done"], "warmup": warmup if first else 1, "ms_per_step": r["ms_per_step"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": {"workload": desc + (", max_len 25 (hard-coded, rnn.py:39), one image per call"
                                                   if is_beam else "") + " [reference modules on host CPU]",
                               "global_batch": B if not is_beam else 16, "seq_len": 25 if is_beam else T,
                               "parallelism": "host cpu"},
                    "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
                    "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                    "gpu_launches": 0}
            if first:
                out = line
            else:
                out.setdefault("others", {})[name] = line
        print(json.dumps(out))
        return 0

    import torch.distributed as dist
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    torch.cuda.set_device(local)
    want_refs = not args.no_gpu_reference
    line = run_ours(main_wl, args, args.steps, warmup, rank, world, local, want_refs)
    for name in extras:
        # fewer timed steps for the extras (a beam "step" is 58 dependent decode steps over 4096 images)
        k = args.steps if not name.startswith("beam") else max(3, min(args.steps, 5))
        try:
            extra = run_ours(name, args, k, warmup, rank, world, local, want_refs)
        except Exception as exc:
            extra = {"error": f"{type(exc).__name__}: {str(exc)[:300]}"}
        if rank == 0:
            line.setdefault("others", {})[name] = extra
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        # A communicator teardown that stalls must not keep torchrun alive (watchdog exit).
        sys.stdout.flush()
        threading.Timer(20.0, lambda: os._exit(0)).start()
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        dist.destroy_process_group()
        os._exit(0)
    return 0


if __name__ == "__main__":
    sys.exit(main())

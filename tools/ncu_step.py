"""One eager training step inside a cudaProfilerStart/Stop range, for `ncu --profile-from-start off`:
    ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/X \
        python tools/ncu_step.py attn_gru 128 196 bf16
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch


def main():
    kind, B, P, dtype = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4]
    T = 20
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    if kind.startswith("attn"):
        from showtell_b200.rnn_attn import RNN_Attn as G
        from showtell_b200.rnn_attn_LSTM import RNN_Attn as Lm
        m = (G if kind == "attn_gru" else Lm)(512, 2048, 512, 512, 10000, 1, dtype=dtype).to(dev)
        feat = torch.relu(torch.randn(B, 2048, P, device=dev))
    else:
        from showtell_b200.rnn import RNN as G
        from showtell_b200.rnn_lstm import RNN as Lm
        m = (G if kind == "gru" else Lm)(512, 512, 10000, 1, dtype=dtype).to(dev)
        feat = torch.randn(B, 512, device=dev)
    m.use_cuda_graphs = False
    cap = torch.randint(4, 10000, (B, T), device=dev)
    lengths = [T] * B
    for _ in range(3):
        m.forward_backward(feat, cap, lengths)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    m.forward_backward(feat, cap, lengths)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()


if __name__ == "__main__":
    main()

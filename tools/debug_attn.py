import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from test_gpu_attn import _random_case, rel_err, O
for cfg in [("attn_gru", 512, 2048, 512, 512, 10000, 1, 49, 12, 20, "fp32"),
            ("attn_gru", 512, 2048, 512, 512, 1000, 1, 49, 12, 20, "fp32"),
            ("attn_gru", 512, 2048, 512, 512, 1000, 1, 48, 12, 20, "fp32"),
            ("attn_gru", 512, 2048, 512, 512, 1000, 1, 49, 12, 6, "fp32"),
            ("attn_lstm", 512, 2048, 512, 512, 1000, 1, 49, 16, 20, "bf16")]:
    kind, E, C, A, H, V, L, P, B, T, dtype = cfg
    m, feat, cap, lengths = _random_case(kind, E, C, A, H, V, L, P, B, T, 5, dtype)
    p = {k: v.detach().clone() for k, v in m.state_dict().items()}
    loss_ref, grads_ref, ex = O.train_step(p, kind, feat, cap, lengths, alpha_c=1.0)
    m = m.to("cuda:0")
    loss, alphas = m.forward_loss(feat.cuda(), cap.cuda(), lengths, alpha_c=1.0)
    loss.backward()
    print(cfg, "lengths", lengths, "loss", float(loss), float(loss_ref))
    print("  alphas", rel_err(alphas, ex["alphas"]))
    for n, q in m.named_parameters():
        d = (q.grad.cpu().double() - grads_ref[n].double())
        print(f"  {n:28s} max-rel {rel_err(q.grad, grads_ref[n]):.2e}  L2 {float(d.norm()/grads_ref[n].double().norm()):.2e}  argmax {int(d.abs().argmax())} of {d.numel()}")

"""Small end-to-end exercise of every bf16 / fp32 kernel for compute-sanitizer (memcheck / racecheck / synccheck):
one un-graphed train step per model kind and precision, one inference batch, one Grad-CAM batch.
    compute-sanitizer --tool memcheck --error-exitcode 1 python scratch/sanitize.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200.step import TrainStep

dev = "cuda:0"
which = sys.argv[1] if len(sys.argv) > 1 else "all"
g = torch.Generator().manual_seed(0)
for kind, nl, B, T in (("cnn", 5, 5, 1000), ("mm", 5, 3, 1000), ("cnn", 1, 2, 5000)):
    x = torch.randn(B, 12, T, generator=g).to(dev)
    y = (torch.rand(B, nl, generator=g) < 0.3).float().to(dev)
    demo = torch.rand(B, 5, generator=g).to(dev) if kind == "mm" else None
    for prec in (("bf16",) if which == "bf16" else ("bf16", "fp32")):
        torch.manual_seed(1)
        m = (P.ECGMultimodal() if kind == "mm" else P.ECGCNN(12, 256, nl)).to(dev).train()
        o = P.FusedAdamW(m.parameters(), lr=1e-3, weight_decay=1e-4)
        e = TrainStep(m, o, B, T, precision=prec, use_graph=False)
        for _ in range(2):
            loss = e(x, y, demo)
        torch.cuda.synchronize()
        print(kind, nl, B, T, prec, "train loss", float(loss))
    m.eval()
    inf = P.InferStep(m, B, T, use_graph=False)
    lg = inf(x, demo)
    part = inf(x[:B - 1], None if demo is None else demo[:B - 1])
    torch.cuda.synchronize()
    print(kind, "infer", float(lg.abs().max()), tuple(part.shape))
    cam, arg = P.gradcam_batch(m, x, demo, signal_length=T, engine=inf)
    cam2, arg2 = P.gradcam_batch(m, x, demo, signal_length=T)
    torch.cuda.synchronize()
    print(kind, "gradcam", float((cam - cam2).abs().max()))
print("sanitize run complete")

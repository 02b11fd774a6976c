"""Stock PyTorch (cuDNN/ATen) train step on this GPU: the library path to beat."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
dev = torch.device('cuda', 0)
for kind, B, T, nl, lr in [('cnn', 256, 1000, 5, 1.5e-3), ('mm', 128, 1000, 5, 1e-4), ('cnn', 64, 5000, 1, 1e-3)]:
    r = bench.gpu_reference(kind, B, T, nl, lr, 1e-4, dev)
    print(json.dumps({'kind': kind, 'B': B, 'T': T, **r}))

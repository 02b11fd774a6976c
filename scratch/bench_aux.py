"""Secondary workloads of BASELINE.json (configs 3-5) + eval forward, device-timed (CUDA events)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200.step import TrainStep

def timed(fn, iters):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
out = {}
# config 4: AF binary CNN, 12 x 5000, batch 512, bf16 (one GPU's view of it: per-rank 64 at 8 GPUs and the full 512)
for B in (64, 512):
    torch.manual_seed(42)
    m = P.ECGCNN(12, 256, 1).cuda().train()
    o = P.FusedAdamW(m.parameters(), lr=1e-3, weight_decay=1e-4)
    e = TrainStep(m, o, B, 5000, precision='bf16')
    e.x.normal_(); e.y.bernoulli_(0.07)
    ms = timed(e.run, 30)
    out[f'af_train_12x5000_bf16_B{B}'] = {'ms_per_step': ms, 'windows_per_s': B / ms * 1e3, 'roofline_frac': 3463e-9 * B / (ms * 1e-3), 'loss': float(e.loss)}
    del e, m, o
# config 3: multimodal (FiLM) train, per-rank shard of global batch 1024 at 8 GPUs (128) and at 1 GPU (1024)
for B in (128, 1024):
    torch.manual_seed(42)
    m = P.ECGMultimodal().cuda().train()
    o = P.FusedAdamW(m.parameters(), lr=1e-4, weight_decay=1e-4)
    e = TrainStep(m, o, B, 1000, precision='bf16')
    e.x.normal_(); e.y.bernoulli_(0.3); e.demo.uniform_()
    ms = timed(e.run, 50)
    out[f'mm_train_12x1000_bf16_B{B}'] = {'ms_per_step': ms, 'windows_per_s': B / ms * 1e3, 'roofline_frac': 693e-9 * B / (ms * 1e-3)}
    del e, m, o
# config 2 at larger batches (how far the fixed cost is amortised)
for B in (512, 1024):
    torch.manual_seed(42)
    m = P.ECGCNN(12, 256, 5).cuda().train()
    o = P.FusedAdamW(m.parameters(), lr=1.5e-3, weight_decay=1e-4)
    e = TrainStep(m, o, B, 1000, precision='bf16')
    e.x.normal_(); e.y.bernoulli_(0.3)
    ms = timed(e.run, 50)
    out[f'cnn_train_12x1000_bf16_B{B}'] = {'ms_per_step': ms, 'windows_per_s': B / ms * 1e3, 'roofline_frac': 693e-9 * B / (ms * 1e-3)}
    del e, m, o
# config 5: batched Grad-CAM over 10k windows, all 5 classes (fp32 exact path), in chunks of 1000
torch.manual_seed(42)
m = P.ECGCNN(12, 256, 5).cuda().eval()
x = torch.randn(10000, 12, 1000, device='cuda')
def cam_all():
    for i in range(0, 10000, 1000):
        P.gradcam_batch(m, x[i:i + 1000], signal_length=1000)
ms = timed(cam_all, 3)
out['gradcam_10k_x5classes_12x1000_fp32'] = {'ms_total': ms, 'windows_per_s': 10000 / ms * 1e3}
# eval forward (module path, fp32 exact kernels)
def fwd():
    with torch.no_grad():
        for i in range(0, 10000, 1000):
            m(x[i:i + 1000])
ms = timed(fwd, 3)
out['eval_forward_10k_12x1000_fp32'] = {'ms_total': ms, 'windows_per_s': 10000 / ms * 1e3}
print(json.dumps(out, indent=1))

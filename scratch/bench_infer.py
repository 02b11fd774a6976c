"""bf16 inference engine (InferStep): device-timed windows/s (CUDA events around graph replays) and the
fraction of the fused-inference roofline (max(t_HBM, t_MMA) per layer; 108*T bf16 elements +
12*T fp32 input per window, 226.56 MFLOP per 12x1000 window)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ptbxl_multimodal_b200 as P

PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
HBM = float(PEAKS.get("hbm_gbs", 6496.8)) * 1e9
TF = float(PEAKS.get("bf16_tflops_sustained", 1391.5)) * 1e12


def model_ns(T):
    """Per-layer roofline of the fused bf16 inference pass, ns per window."""
    chan = [16, 32, 64, 128, 256]
    L = [T, T // 2, T // 4, T // 8]
    ns = 12 * T * 4 / HBM * 1e9 + 16 * T * 2 / HBM * 1e9          # pack: fp32 in, bf16 out
    for l in range(4):
        fl = 2.0 * [12, 32, 64, 128][l] * chan[l + 1] * 15 * L[l]
        by = chan[l] * L[l] * 2 + (chan[l + 1] * (L[l] // 2) * 2 if l < 3 else 0)
        ns += max(fl / TF, by / HBM) * 1e9
    return ns


def timed(fn, iters):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

out = {}
for kind, nl, B, T in (("cnn", 5, 256, 1000), ("cnn", 5, 1024, 1000), ("cnn", 5, 4096, 1000), ("mm", 5, 1024, 1000),
                       ("cnn", 1, 512, 5000)):
    torch.manual_seed(42)
    m = (P.ECGMultimodal() if kind == "mm" else P.ECGCNN(12, 256, nl)).cuda().eval()
    e = P.InferStep(m, B, T)
    for s in (0, 1):
        e.xs[s].normal_()
        if kind == "mm": e.demos[s].uniform_()
    e.capture()
    slot = [0]
    def step():
        slot[0] ^= 1
        e.run(slot=slot[0])
    ms = timed(step, 50)
    mns = model_ns(T)
    out[f"{kind}_infer_12x{T}_bf16_B{B}"] = {"ms_per_batch": ms, "windows_per_s": B / ms * 1e3, "model_ns_per_window": mns,
                                           "roofline_frac": mns * 1e-9 * B / (ms * 1e-3)}
    del e, m
# config 5: batched Grad-CAM over 10k windows x 5 classes with the forward on the bf16 engine (chunks of 2000)
torch.manual_seed(42)
m = P.ECGCNN(12, 256, 5).cuda().eval()
e = P.InferStep(m, 2000, 1000)
x = torch.randn(10000, 12, 1000, device="cuda")
def cam_all():
    for i in range(0, 10000, 2000):
        P.gradcam_batch(m, x[i:i + 2000], signal_length=1000, engine=e)
ms = timed(cam_all, 5)
out["gradcam_10k_x5classes_12x1000_bf16_engine"] = {"ms_total": ms, "windows_per_s": 10000 / ms * 1e3}
def fwd_all():
    for i in range(0, 10000, 2000):
        e(x[i:i + 2000])
ms = timed(fwd_all, 5)
out["eval_forward_10k_12x1000_bf16_engine_incl_d2d_copy"] = {"ms_total": ms, "windows_per_s": 10000 / ms * 1e3}
print(json.dumps(out, indent=1))

cd /root/repo
export ECGB200_SPIN_TIMEOUT_MS=3000,60000
timeout 900 python -m pytest tests/test_gpu_infer.py -m gpu -x -q -s 2>&1 | grep -E "fp32x3|passed|failed|Error|error|assert" | head -20
python - <<'PY'
import torch, json, sys
sys.path.insert(0, '.')
import ptbxl_multimodal_b200 as P
torch.manual_seed(42)
m = P.ECGCNN(12, 256, 5).cuda().eval()
x = torch.randn(10000, 12, 1000, device='cuda')
out = {}
for prec in ('bf16', 'fp32x3'):
    e = P.InferStep(m, 1000, 1000, precision=prec)
    def run():
        for i in range(0, 10000, 1000): P.gradcam_batch(m, x[i:i+1000], signal_length=1000, engine=e)
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); run(); run(); run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    out[prec] = {'ms_10k_x5': round(ms, 3), 'windows_per_s': round(10000 / ms * 1e3)}
def run32():
    for i in range(0, 10000, 1000): P.gradcam_batch(m, x[i:i+1000], signal_length=1000)
run32(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); run32(); e1.record(); torch.cuda.synchronize()
out['fp32_module'] = {'ms_10k_x5': round(e0.elapsed_time(e1), 3), 'windows_per_s': round(10000 / e0.elapsed_time(e1) * 1e3)}
print(json.dumps(out))
PY

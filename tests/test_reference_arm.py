"""The reference arm of bench.py: oracle/_ref (the unmodified reference files, staged by oracle/make_ref.py) runs the
path through the reference's OWN loop functions, and the oracle port used everywhere else as the checker is
bit-identical to it -- same loss, same post-AdamW parameters, same running statistics, same CAMs (CPU, fp32).

oracle/_ref is git-ignored; it is staged by `__graft_entry__.build()` whenever /root/reference is present.  Where neither
exists (a checkout without the reference) these tests skip: the committed golden vectors (tests/test_oracle_golden.py)
pin the port independently."""
import json
import os
import subprocess
import sys
import warnings

import pytest
import torch

from conftest import ROOT
from oracle import ecg_oracle as O
from oracle import make_ref


@pytest.fixture(scope="module")
def ref():
    make_ref.build()                       # no-op without /root/reference
    R = make_ref.load()
    if R is None:
        pytest.skip("oracle/_ref not staged and /root/reference absent")
    os.environ.setdefault("TQDM_DISABLE", "1")
    return R


def _bench():
    sys.path.insert(0, ROOT)
    import bench
    return bench


def test_staged_files_are_the_reference_byte_for_byte(ref):
    assert make_ref.verify()
    for rel in make_ref.FILES:
        src = os.path.join("/root/reference", rel)
        if os.path.exists(src):
            with open(src, "rb") as a, open(os.path.join(make_ref.DEST, rel), "rb") as b:
                assert a.read() == b.read(), rel
    # git never sees them
    out = subprocess.run(["git", "-C", ROOT, "check-ignore", "oracle/_ref/src/models/ecg_cnn.py"], capture_output=True, text=True)
    assert out.returncode == 0
    with open(os.path.join(ROOT, ".gpurunignore")) as f:
        assert "oracle" not in f.read()      # ... but the directory travels to the GPU box


def test_reference_train_loop_equals_the_port_bit_for_bit(ref):
    """train_one_epoch (src/training/loop.py:14-38) over three batches == three O.train_step calls."""
    bench = _bench()
    torch.set_num_threads(4)
    torch.manual_seed(42)
    model = ref.ecg_cnn.ECGCNN(12, 256, 5)
    opt = torch.optim.AdamW(model.parameters(), lr=1.5e-3, weight_decay=1e-4)
    sd = O.init_state_dict("cnn", 5, seed=42)
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]), k                       # same initial weights from the same seed
    st = O.AdamWState(sd, 1.5e-3, 1e-4)
    batches = [O.synth_batch(8, 1000, 5, seed=s) for s in range(3)]
    mean_loss = ref.loop.train_one_epoch(model, bench._Batches(batches), opt, torch.device("cpu"))
    losses = [float(O.train_step(sd, x, y, st)["loss"]) for x, y in batches]
    assert mean_loss == sum(l * 8 for l in losses) / 24
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_reference_demo_loop_equals_the_port_bit_for_bit(ref):
    """train_one_epoch_demo (src/training/loop_demo.py:13-45) on the FiLM model."""
    bench = _bench()
    torch.set_num_threads(4)
    torch.manual_seed(42)
    model = ref.ecg_multimodal.ECGMultimodal(num_labels=5)
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
    sd = O.init_state_dict("mm", 5, seed=42)
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]), k
    st = O.AdamWState(sd, 1e-4, 1e-4)
    batches = [O.synth_batch(8, 1000, 5, seed=10 + s, with_demo=True) for s in range(2)]
    mean_loss = ref.loop_demo.train_one_epoch_demo(model, bench._Batches(batches), opt, torch.device("cpu"))
    losses = [float(O.train_step(sd, x, y, st, demo=d)["loss"]) for x, d, y in batches]
    assert mean_loss == sum(losses) / 2
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_reference_gradcam_equals_the_port_bit_for_bit(ref):
    torch.set_num_threads(4)
    torch.manual_seed(42)
    model = ref.ecg_cnn.ECGCNN(12, 256, 5).eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    x, _ = O.synth_batch(2, 1000, 5, seed=3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cam = ref.grad_cam_1d.GradCAM1D(model, model.backbone[-1].net[0])
        for i in range(2):
            for c in (0, 4):
                got = cam.generate_cam(x[i:i + 1], c, signal_length=1000)
                assert torch.equal(got, O.gradcam_v1(sd, x[i:i + 1], c, 1000)), (i, c)


def test_bench_reference_arm_line(ref):
    """`bench.py --impl reference` prints the contract's line, runs the staged reference (kind "reference"), and names the
    workload with exactly the `config` object the b200 arm of the same command line prints."""
    bench = _bench()
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--batch", "8", "--steps", "2",
                          "--warmup", "3"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "reference" and d["steps"] == 2
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0

    class A:                                                  # the same command line, parsed for the b200 arm
        config, batch, seq_len, steps, precision = 1, 8, None, 2, "auto"
    w = bench.resolve(A, 1)
    assert d["config"] == bench.workload_config(w, 1, "bf16")
    assert d["metric"] == w["metric"] == "ECG samples/sec train step (12x1000)"


def test_both_arms_build_config_with_one_function():
    """`config` carries no run-specific extras: both arms call workload_config(w, world, precision) and nothing else, so the
    driver's same-config check compares equal objects."""
    import inspect
    bench = _bench()
    assert list(inspect.signature(bench.workload_config).parameters) == ["w", "world", "precision"]
    src = inspect.getsource(bench)
    assert src.count('"config": workload_config(') == 3       # reference arm, train arm, Grad-CAM arm
    for n in (1, 2, 8):
        for cfg in (1, 2, 3, 4):
            class A:
                config, batch, seq_len, steps, precision = cfg, None, None, None, "auto"
            w = bench.resolve(A, n)
            c = bench.workload_config(w, n, "bf16")
            assert set(c) == {"workload", "batch_per_gpu", "global_batch", "seq_len", "parallelism", "precision", "l2"}
            assert c["workload"].startswith(f"configs[{cfg}]") and c["parallelism"] == f"dp{n}"


def test_demo_script_harness_on_the_unswapped_reference(ref, tmp_path, expected_probs, demo_inputs):
    """tests/ref_scripts.py drives scripts/00_demo_inference.py (unmodified, staged) end to end.  Here WITHOUT the import swap
    and on the CPU: the script's printed probabilities are the shipped known-answer row and the CAM it hands to its plot is
    the oracle's script-order Grad-CAM bit for bit -- so the GPU test that runs the same harness WITH the swap
    (tests/test_gpu_reference_scripts.py) measures the product, not the harness."""
    if torch.cuda.is_available():
        pytest.skip("the script picks cuda when it is there; the CPU leg of the harness check runs in the CPU suite")
    from conftest import load_ckpt
    import ref_scripts
    x, _ = demo_inputs
    sd = load_ckpt("ecg_baseline_best.pth")
    for row, c in ((3, 0), (0, 4)):
        got = ref_scripts.run_demo_inference(tmp_path, row, c, swap=False)
        assert got["device"] == "cpu" and got["model_module"] == "src.models.ecg_cnn"
        want = torch.tensor(expected_probs["baseline_prob"][row])
        assert float((torch.from_numpy(got["probs"]) - want).abs().max()) <= 5.1e-4        # printed with 3 decimals
        assert torch.equal(got["cam"], O.gradcam_v2(sd, x[row:row + 1], c, x.shape[-1]))


def test_script_local_gradcam_classes_equal_the_port_on_the_reference_models(ref, demo_inputs):
    """scripts/12_grad_cam_ecg_demo.py (GradCAM1D_ECGMultimodal, compute_demo_importance) and scripts/13_grad_cam_af.py
    (GradCAM1D_AF), imported as written from oracle/_ref, on the REFERENCE models with the shipped checkpoints (CPU): bit-equal
    to the oracle's V2 / V3 restatements.  The GPU twin (tests/test_gpu_zz_reference_scripts.py) hands the same classes the
    product models."""
    from conftest import load_ckpt
    import numpy as np
    import ref_scripts
    torch.set_num_threads(4)
    x, d = demo_inputs
    T = x.shape[-1]
    m12 = ref_scripts.load_script("12_grad_cam_ecg_demo.py")
    m13 = ref_scripts.load_script("13_grad_cam_af.py")
    assert "src.datasets" not in sys.modules and os.environ.get("TORCH_CUDA_ARCH_LIST") != "native"

    sd = load_ckpt("ecg_multimodal_best.pth")
    mm = ref.ecg_multimodal.ECGMultimodal()
    mm.load_state_dict(sd)
    mm.eval()
    g = m12.GradCAM1D_ECGMultimodal(mm, mm.ecg_backbone.backbone[-1].net[0])
    for j, c in ((0, 0), (2, 3)):
        xe, xd = x[3 + j:4 + j], d[j:j + 1]
        assert torch.equal(g.generate_cam(xe, xd, c, T), O.gradcam_v2(sd, xe, c, T, demo=xd, eps=1e-8)), (j, c)
        imp = m12.compute_demo_importance(mm, xe, xd, c)
        assert np.array_equal(imp, O.demo_importance(sd, xe, xd, c).numpy()), (j, c)
    g.remove_hooks()

    sa = load_ckpt("af_binary_best.pth")
    af = ref.ecg_cnn.ECGCNN(12, 256, 1)
    af.load_state_dict(sa)
    af.eval()
    g = m13.GradCAM1D_AF(af, af.backbone[-1].net[0])
    for row in (6, 8):
        assert torch.equal(g.generate_cam(x[row:row + 1], T), O.gradcam_v2(sa, x[row:row + 1], 0, T)), row
    g.remove_hooks()


def test_bench_reference_arm_under_torchrun_prints_one_line(ref):
    """The driver launches the reference arm like the b200 arm (torchrun, one process per GPU): rank 0 alone runs the
    reference and prints the line, the other ranks exit 0 without work."""
    import socket
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "bench.py"),
                          "--impl", "reference", "--gpus", "2", "--batch", "8", "--steps", "1", "--warmup", "3"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["config"]["parallelism"] == "dp2"
    assert d["config"]["batch_per_gpu"] == 8 and d["config"]["global_batch"] == 16 and d["scaling"] == "weak"

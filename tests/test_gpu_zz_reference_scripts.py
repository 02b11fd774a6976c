"""Drop-in boundary, end to end: the reference's OWN callers of the path -- staged unmodified in oracle/_ref by
oracle/make_ref.py -- run on the B200 with only the model import swapped:

  * scripts/00_demo_inference.py: load the shipped checkpoint with load_state_dict(strict=False), .to(cuda), eval forward,
    its script-local Grad-CAM class with forward / full-backward hooks on the last nn.Conv1d found by walking model.modules();
  * src/training/loop.py train_one_epoch / eval_one_epoch and loop_demo.py train_one_epoch_demo: the reference's loop bodies
    (torch's own BCE on the product's logits, loss.backward() through the product's autograd Functions) with
    torch.optim.AdamW and with FusedAdamW;
  * src/interpretability/grad_cam_1d.py GradCAM1D (legacy register_backward_hook) on the product model.

Results against the CPU oracle at the fp32 tolerance of tests/test_gpu_parity.py.  The harness itself is validated on the CPU
against the unswapped reference (tests/test_reference_arm.py).  Skipped when oracle/_ref did not travel."""
import os
import sys
import warnings

import pytest
import torch

import ptbxl_multimodal_b200 as P
from oracle import ecg_oracle as O
from oracle import make_ref
from conftest import ROOT, load_ckpt

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


@pytest.fixture(scope="module")
def ref():
    os.environ.setdefault("TQDM_DISABLE", "1")
    R = make_ref.load()
    if R is None:
        pytest.skip("oracle/_ref not staged (run __graft_entry__.build() where /root/reference exists)")
    return R


def _batches():
    sys.path.insert(0, ROOT)
    import bench
    return bench._Batches


def _peak_ok(cam, cam_ref):
    """argmax equal, or the product's peak sits on an oracle value within 1e-4 of the oracle's maximum (flat top)."""
    i = int(cam.argmax())
    return i == int(cam_ref.argmax()) or float(cam_ref.max() - cam_ref[i]) <= 1e-4


def test_demo_inference_script_runs_on_the_product_after_the_import_swap(ref, tmp_path, expected_probs, demo_inputs):
    import ref_scripts
    x, _ = demo_inputs
    sd = load_ckpt("ecg_baseline_best.pth")
    for row, c in ((3, 0), (0, 4), (4, 2)):
        got = ref_scripts.run_demo_inference(tmp_path, row, c, swap=True)
        assert got["device"] == "cuda" and got["model_module"] == "ptbxl_multimodal_b200.ecg_cnn"
        want = torch.tensor(expected_probs["baseline_prob"][row])
        assert float((torch.from_numpy(got["probs"]) - want).abs().max()) <= 5.1e-4 + 1e-4      # printed with 3 decimals
        cam_ref = O.gradcam_v2(sd, x[row:row + 1], c, x.shape[-1])
        assert got["cam"].shape == cam_ref.shape
        assert float((got["cam"] - cam_ref).abs().max()) < 1e-3, (row, c)
        assert _peak_ok(got["cam"], cam_ref), (row, c)
    # the swap is scoped: afterwards `src.models.ecg_cnn` is the reference's module again
    assert sys.modules["src.models.ecg_cnn"].__file__.startswith(make_ref.DEST)


@pytest.mark.parametrize("optimizer", ["torch.optim.AdamW", "FusedAdamW"])
def test_reference_train_and_eval_loops_drive_the_product_model(ref, optimizer):
    """src/training/loop.py:14-73 unmodified; model = P.ECGCNN on cuda."""
    Batches = _batches()
    torch.manual_seed(42)
    model = P.ECGCNN(12, 256, 5).to(DEV)
    opt = (torch.optim.AdamW if optimizer == "torch.optim.AdamW" else P.FusedAdamW)(model.parameters(), lr=1.5e-3, weight_decay=1e-4)
    data = [O.synth_batch(8, 250, 5, seed=3 + s) for s in range(3)]
    got = ref.loop.train_one_epoch(model, Batches(data), opt, torch.device(DEV))
    sd = O.init_state_dict("cnn", 5, seed=42)
    st = O.AdamWState(sd, 1.5e-3, 1e-4)
    want = sum(float(O.train_step(sd, x, y, st)["loss"]) * 8 for x, y in data) / 24
    assert abs(got - want) < TOL * abs(want), (got, want)

    m = ref.loop.eval_one_epoch(model, Batches(data), torch.device(DEV))
    assert set(m) == {"auroc_macro", "auprc_macro", "f1_macro", "bce_loss"}
    with torch.no_grad():
        want_eval = sum(float(O.bce_with_logits(O.ecgcnn_forward(dict(sd), x, train=False), y)) * 8 for x, y in data) / 24
    assert abs(m["bce_loss"] - want_eval) < 2e-3 * abs(want_eval), (m["bce_loss"], want_eval)
    mine = P.eval_one_epoch(model, Batches(data), DEV)        # the product's loop on the same model: same metrics
    assert abs(mine["bce_loss"] - m["bce_loss"]) < 1e-5 * abs(m["bce_loss"])
    assert abs(mine["auroc_macro"] - m["auroc_macro"]) < 1e-6 and abs(mine["f1_macro"] - m["f1_macro"]) < 1e-12


def test_reference_demo_loop_drives_the_product_multimodal_model(ref):
    """src/training/loop_demo.py:13-45 unmodified; model = P.ECGMultimodal on cuda."""
    Batches = _batches()
    torch.manual_seed(42)
    model = P.ECGMultimodal(num_labels=5).to(DEV)
    opt = P.FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
    data = [O.synth_batch(8, 250, 5, seed=14 + s, with_demo=True) for s in range(2)]
    got = ref.loop_demo.train_one_epoch_demo(model, Batches(data), opt, torch.device(DEV))
    sd = O.init_state_dict("mm", 5, seed=42)
    st = O.AdamWState(sd, 1e-4, 1e-4)
    want = sum(float(O.train_step(sd, x, y, st, demo=d)["loss"]) for x, d, y in data) / 2
    assert abs(got - want) < TOL * abs(want), (got, want)
    assert "bce_loss" in ref.loop_demo.eval_one_epoch_demo(model, Batches(data), torch.device(DEV))


def test_reference_gradcam_class_on_the_product_model(ref, demo_inputs):
    """src/interpretability/grad_cam_1d.py unmodified (legacy backward hook, batch-1 calls) on P.ECGCNN."""
    x, _ = demo_inputs
    sd = load_ckpt("ecg_baseline_best.pth")
    model = P.ECGCNN(12, 256, 5)
    model.load_state_dict(sd)
    model = model.to(DEV)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cam = ref.grad_cam_1d.GradCAM1D(model, model.backbone[-1].net[0])
        for row, c in ((3, 0), (5, 1)):
            got = cam.generate_cam(x[row:row + 1].to(DEV), c, signal_length=x.shape[-1]).cpu()
            want = O.gradcam_v1(sd, x[row:row + 1], c, x.shape[-1])
            assert got.shape == want.shape
            assert float((got - want).abs().max()) < 1e-3, (row, c)
            assert _peak_ok(got, want), (row, c)


def test_script_local_gradcam_classes_on_the_product_models(ref, demo_inputs):
    """scripts/12_grad_cam_ecg_demo.py:17-97 (GradCAM1D_ECGMultimodal: full backward hook on
    model.ecg_backbone.backbone[-1].net[0], two-input forward; compute_demo_importance: gradient w.r.t. x_demo) and
    scripts/13_grad_cam_af.py:22-76 (GradCAM1D_AF), imported as written, on P.ECGMultimodal / P.ECGCNN with the shipped
    checkpoints."""
    import numpy as np
    import ref_scripts
    x, d = demo_inputs
    T = x.shape[-1]
    m12 = ref_scripts.load_script("12_grad_cam_ecg_demo.py")
    m13 = ref_scripts.load_script("13_grad_cam_af.py")

    sd = load_ckpt("ecg_multimodal_best.pth")
    mm = P.ECGMultimodal()
    mm.load_state_dict(sd)
    mm = mm.to(DEV).eval()
    g = m12.GradCAM1D_ECGMultimodal(mm, mm.ecg_backbone.backbone[-1].net[0])
    for j, c in ((0, 0), (2, 3), (5, 1)):
        xe, xd = x[3 + j:4 + j], d[j:j + 1]
        got = g.generate_cam(xe.to(DEV), xd.to(DEV), c, T)
        want = O.gradcam_v2(sd, xe, c, T, demo=xd, eps=1e-8)
        assert got.shape == want.shape and float((got - want).abs().max()) < 1e-3, (j, c)
        assert _peak_ok(got, want), (j, c)
        imp = m12.compute_demo_importance(mm, xe.to(DEV), xd.to(DEV), c)
        np.testing.assert_allclose(imp, O.demo_importance(sd, xe, xd, c).numpy(), atol=TOL)
    g.remove_hooks()

    sa = load_ckpt("af_binary_best.pth")
    af = P.ECGCNN(12, 256, 1)
    af.load_state_dict(sa)
    af = af.to(DEV).eval()
    g = m13.GradCAM1D_AF(af, af.backbone[-1].net[0])
    for row in (6, 8):
        got = g.generate_cam(x[row:row + 1].to(DEV), T)
        want = O.gradcam_v2(sa, x[row:row + 1], 0, T)
        assert got.shape == want.shape and float((got - want).abs().max()) < 1e-3, row
        assert _peak_ok(got, want), row
    g.remove_hooks()

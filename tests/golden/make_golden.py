"""Generate tests/golden/* from the UNMODIFIED reference (build container only).

Run once here:  python tests/golden/make_golden.py
Needs /root/reference (read-only).  It
  1. copies the reference's shipped known-answer DATA artefacts (3 checkpoints,
     10 demo ECGs, 7 demo vectors, the CSV rows they map to, one CAM vector);
  2. runs the live reference modules (src.models.*, src.training.loop*,
     src.interpretability.grad_cam_1d, and the script-local Grad-CAM classes of
     scripts/00, 12, 13 imported with matplotlib/wfdb stubbed out) on seeded
     synthetic inputs, asserts that oracle/ecg_oracle.py reproduces them
     BIT-EXACTLY on CPU, and stores the results as fixtures.
No reference source code is copied; only data and computed vectors.
"""
import csv
import importlib.util
import json
import os
import shutil
import sys
import types
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from oracle import ecg_oracle as O                      # noqa: E402
from src.models.ecg_cnn import ECGCNN                    # noqa: E402
from src.models.ecg_multimodal import ECGMultimodal      # noqa: E402
from src.interpretability.grad_cam_1d import GradCAM1D   # noqa: E402
import torch.nn.functional as F                          # noqa: E402

torch.set_num_threads(8)
CLASSES = ["MI", "STTC", "HYP", "CD", "NORM"]


def stub_and_import(path, name):
    for mod in ["matplotlib", "matplotlib.pyplot", "wfdb", "seaborn"]:
        if mod not in sys.modules:
            sys.modules[mod] = types.ModuleType(mod)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def sig(t, stride=61):
    """Compact signature of a tensor: full if small, else strided sample + norms."""
    t = t.detach().double().flatten()
    full = t.numel() <= 4096
    return {"full": full,
            "vals": (t if full else t[::stride]).float().numpy(),
            "l2": float(t.norm()), "sum": float(t.sum()), "absmax": float(t.abs().max())}


def pack(prefix, d, out):
    for k, s in d.items():
        out[f"{prefix}/{k}/vals"] = s["vals"]
        out[f"{prefix}/{k}/meta"] = np.array([s["l2"], s["sum"], s["absmax"], float(s["full"])])


def main():
    # ---------------- 1. shipped artefacts ----------------
    os.makedirs(os.path.join(HERE, "ckpts"), exist_ok=True)
    ck = {"ecg_baseline_best.pth": "outputs/ecg_baseline/ckpts/ecg_baseline_best.pth",
          "ecg_multimodal_best.pth": "outputs/ecg_multimodal/ckpts/ecg_multimodal_best.pth",
          "af_binary_best.pth": "outputs/af_binary/ckpts/af_binary_best.pth"}
    for dst, src in ck.items():
        shutil.copyfile(os.path.join(REF, src), os.path.join(HERE, "ckpts", dst))
        os.chmod(os.path.join(HERE, "ckpts", dst), 0o644)

    rows = [0, 1, 2, 245, 1440, 1909, 75, 789, 334, 1259]     # data/demo/meta.csv
    ecgs = [np.load(f"{REF}/data/demo/demo_ecg_{i}.npy") for i in range(3)]
    demos = []
    for i in range(7):
        s = np.load(f"{REF}/data/demo/single/single_sample_0{i}.npz", allow_pickle=True)
        m = np.load(f"{REF}/data/demo/multimodal/mm_sample_0{i}.npz", allow_pickle=True)
        assert np.array_equal(s["ecg"], m["ecg"])
        ecgs.append(s["ecg"])
        demos.append(m["demo"])
    ecgs = np.stack(ecgs).astype(np.float32)                  # (10, 12, 5000)
    demos = np.stack(demos).astype(np.float32)                # (7, 5) -> rows[3:]
    np.savez_compressed(os.path.join(HERE, "demo_inputs.npz"), ecg=ecgs, demo=demos,
                        rows=np.array(rows))

    def csv_rows(path, cols):
        with open(os.path.join(REF, path)) as f:
            r = list(csv.DictReader(f))
        return [[float(r[i][c]) for c in cols] for i in rows]

    expected = {
        "rows": rows,
        "baseline_prob": csv_rows("outputs/ecg_baseline/preds/ecg_baseline_test_preds.csv",
                                  [f"y_prob_{c}" for c in CLASSES]),
        "baseline_pred": csv_rows("outputs/ecg_baseline/preds/ecg_baseline_test_preds.csv",
                                  [f"y_pred_{c}" for c in CLASSES]),
        "mm_prob": csv_rows("outputs/ecg_multimodal/preds/ecg_multimodal_test_preds.csv",
                            [f"y_prob_{c}_mm" for c in CLASSES])[3:],
        "mm_pred": csv_rows("outputs/ecg_multimodal/preds/ecg_multimodal_test_preds.csv",
                            [f"y_pred_{c}_mm" for c in CLASSES])[3:],
        "af_prob": csv_rows("outputs/af_binary/preds/af_binary_test_preds.csv", ["y_prob_AF"]),
        "af_pred": csv_rows("outputs/af_binary/preds/af_binary_test_preds.csv", ["y_pred_AF"]),
    }
    with open(os.path.join(HERE, "expected_probs.json"), "w") as f:
        json.dump(expected, f, indent=1)
    shutil.copyfile(f"{REF}/outputs/gradcam/sample_0_MI_cam.npy",
                    os.path.join(HERE, "sample_0_MI_cam.npy"))
    os.chmod(os.path.join(HERE, "sample_0_MI_cam.npy"), 0o644)

    # ---------------- 2. live reference vs oracle ----------------
    out = {}
    x10 = torch.from_numpy(ecgs)
    d7 = torch.from_numpy(demos)

    # 2a. eval forward with shipped checkpoints (reference modules, CPU fp32)
    def load(model, name):
        sd = torch.load(os.path.join(HERE, "ckpts", name), map_location="cpu")["model_state"]
        model.load_state_dict(sd)
        return model.eval(), sd

    base, sd_base = load(ECGCNN(12, 256, 5), "ecg_baseline_best.pth")
    mm, sd_mm = load(ECGMultimodal(), "ecg_multimodal_best.pth")
    af, sd_af = load(ECGCNN(12, 256, 1), "af_binary_best.pth")
    with torch.no_grad():
        lb = base(x10); lm = mm(x10[3:], d7); la = af(x10)
        assert torch.equal(lb, O.ecgcnn_forward(O.clone_sd(sd_base), x10))
        assert torch.equal(lm, O.multimodal_forward(O.clone_sd(sd_mm), x10[3:], d7))
        assert torch.equal(la, O.ecgcnn_forward(O.clone_sd(sd_af), x10))
    out["eval/baseline_logits"] = lb.numpy(); out["eval/mm_logits"] = lm.numpy()
    out["eval/af_logits"] = la.numpy()
    print("eval max|p-csv| baseline", np.abs(torch.sigmoid(lb).numpy() - np.array(expected["baseline_prob"])).max(),
          "mm", np.abs(torch.sigmoid(lm).numpy() - np.array(expected["mm_prob"])).max())

    # 2b. train steps on synthetic batches: reference loop body vs oracle
    def run_train(kind, num_labels, B, T, lr, wd, steps, tag):
        torch.manual_seed(42)
        model = ECGCNN(12, 256, num_labels) if kind == "cnn" else ECGMultimodal()
        sd0 = O.init_state_dict(kind, num_labels, seed=42)
        for k, v in model.state_dict().items():
            assert torch.equal(v, sd0[k]), k
        sdo = O.clone_sd(sd0)
        # pick the first batch seed whose ReLU / MaxPool decision margin is > 1e-6, so that
        # two correct fp32 implementations take identical routing decisions at step 0
        for bseed in range(400):
            batch = O.synth_batch(B, T, num_labels, seed=bseed, with_demo=(kind == "mm"))
            mg = O.decision_margin(sd0, batch[0], kind)
            if mg > 1.5e-6:
                break
        else:
            raise RuntimeError("no margin-safe seed found")
        out[f"{tag}/margin"] = np.array([mg, bseed])
        print(tag, "batch seed", bseed, "decision margin", mg)
        x, y = batch[0], batch[-1]
        demo = batch[1] if kind == "mm" else None
        opt = torch.optim.AdamW(model.parameters(), lr=lr, weight_decay=wd)
        st = O.AdamWState(sdo, lr, wd)
        model.train()
        out[f"{tag}/x"] = x.numpy(); out[f"{tag}/y"] = y.numpy()
        if demo is not None:
            out[f"{tag}/demo"] = demo.numpy()
        out[f"{tag}/cfg"] = np.array([B, T, num_labels, lr, wd, steps], dtype=np.float64)
        for s in range(steps):
            opt.zero_grad()
            logits = model(x) if demo is None else model(x, demo)
            loss = F.binary_cross_entropy_with_logits(logits, y)
            loss.backward()
            ref_grads = {k: p.grad.clone() for k, p in model.named_parameters()}
            opt.step()
            o = O.train_step(sdo, x, y, st, demo=demo)
            assert torch.equal(o["logits"], logits.detach()), (tag, s)
            assert torch.equal(o["loss"], loss.detach())
            for k in ref_grads:
                assert torch.equal(o["grads"][k], ref_grads[k]), (tag, s, k)
            for k, v in model.state_dict().items():
                assert torch.equal(v, sdo[k]), (tag, s, k, (v.float() - sdo[k].float()).abs().max())
            out[f"{tag}/step{s}/loss"] = np.array(loss.item())
            out[f"{tag}/step{s}/logits"] = logits.detach().numpy()
            if s == 0:
                pack(f"{tag}/step0/grad", {k: sig(g) for k, g in ref_grads.items()}, out)
        pack(f"{tag}/final", {k: sig(v) for k, v in model.state_dict().items()
                              if not k.endswith("num_batches_tracked")}, out)
        print(tag, "ok: loss", [float(out[f"{tag}/step{s}/loss"]) for s in range(steps)])

    run_train("cnn", 5, 3, 1000, 1.5e-3, 1e-4, 3, "train_cnn")
    run_train("mm", 5, 3, 1000, 1e-4, 1e-4, 3, "train_mm")
    run_train("cnn", 1, 1, 5000, 1e-3, 1e-4, 2, "train_af")
    run_train("cnn", 5, 5, 250, 1.5e-3, 1e-4, 2, "train_cnn_t250")   # L4=31 (odd), Lp=15

    # 2c. Grad-CAM variants, live
    gc = GradCAM1D(base, base.backbone[-1].net[0])
    x0 = x10[0:1]
    for c in range(5):
        cam_hi = gc.generate_cam(x0, c, signal_length=5000).detach()
        cam_lo = gc.generate_cam(x0, c, signal_length=None).detach()
        assert torch.equal(cam_hi, O.gradcam_v1(sd_base, x0, c, 5000)), c
        assert torch.equal(cam_lo, O.gradcam_v1(sd_base, x0, c, None)), c
        out[f"cam/v1_base_s0_c{c}_T"] = cam_hi.numpy()
        out[f"cam/v1_base_s0_c{c}_lo"] = cam_lo.numpy()
    shipped = np.load(os.path.join(HERE, "sample_0_MI_cam.npy"))
    print("shipped CAM: argmax", shipped.argmax(), out["cam/v1_base_s0_c0_T"].argmax(),
          "max abs diff", np.abs(shipped - out["cam/v1_base_s0_c0_T"]).max())

    s00 = stub_and_import(f"{REF}/scripts/00_demo_inference.py", "ref_s00")
    s12 = stub_and_import(f"{REF}/scripts/12_grad_cam_ecg_demo.py", "ref_s12")
    s13 = stub_and_import(f"{REF}/scripts/13_grad_cam_af.py", "ref_s13")
    base2, _ = load(ECGCNN(12, 256, 5), "ecg_baseline_best.pth")
    g2 = s00.GradCAM1D_ECG(base2, s00.find_last_conv1d(base2))
    for smp in (0, 4):
        for c in range(5):
            cam = g2.generate_cam(x10[smp:smp + 1], c, 5000)
            assert torch.equal(cam, O.gradcam_v2(sd_base, x10[smp:smp + 1], c, 5000, eps=1e-9)), (smp, c)
            out[f"cam/v2_base_s{smp}_c{c}"] = cam.numpy()
    af2, _ = load(ECGCNN(12, 256, 1), "af_binary_best.pth")
    g3 = s13.GradCAM1D_AF(af2, af2.backbone[-1].net[0])
    for smp in (0, 7):
        cam = g3.generate_cam(x10[smp:smp + 1], 5000)
        assert torch.equal(cam, O.gradcam_v2(sd_af, x10[smp:smp + 1], 0, 5000, eps=1e-9))
        out[f"cam/v2_af_s{smp}"] = cam.numpy()
    mm2, _ = load(ECGMultimodal(), "ecg_multimodal_best.pth")
    g4 = s12.GradCAM1D_ECGMultimodal(mm2, mm2.ecg_backbone.backbone[-1].net[0])
    for j in (0, 5):
        for c in (0, 3):
            cam = g4.generate_cam(x10[3 + j:4 + j], d7[j:j + 1], c, 5000)
            assert torch.equal(cam, O.gradcam_v2(sd_mm, x10[3 + j:4 + j], c, 5000,
                                                 demo=d7[j:j + 1], eps=1e-8)), (j, c)
            out[f"cam/v3_mm_j{j}_c{c}"] = cam.numpy()
            imp = s12.compute_demo_importance(mm2, x10[3 + j:4 + j], d7[j:j + 1], c)
            oi = O.demo_importance(sd_mm, x10[3 + j:4 + j], d7[j:j + 1], c).numpy()
            assert np.array_equal(imp, oi), (imp, oi)
            out[f"imp/mm_j{j}_c{c}"] = imp

    # 2d. closed-form batched Grad-CAM vs per-sample reference calls (not bit-exact: tolerance)
    cams = O.gradcam_batched(sd_base, x10[:4], signal_length=5000, variant="v1")
    worst = 0.0
    for n in range(4):
        for c in range(5):
            ref = gc.generate_cam(x10[n:n + 1], c, 5000).detach()
            worst = max(worst, float((cams[n, c] - ref).abs().max()))
            assert int(cams[n, c].argmax()) == int(ref.argmax())
    print("closed-form batched CAM vs reference: max abs", worst)
    assert worst < 1e-5
    cams3 = O.gradcam_batched(sd_mm, x10[3:6], 5000, demo=d7[:3], variant="v2", eps=1e-8)
    for j in range(3):
        ref = g4.generate_cam(x10[3 + j:4 + j], d7[j:j + 1], 2, 5000)
        assert float((cams3[j, 2] - ref).abs().max()) < 1e-5

    np.savez_compressed(os.path.join(HERE, "ref_vectors.npz"), **out)
    print("wrote", len(out), "arrays")


def legacy_concat_manifest():
    """Key / shape / dtype manifest of the legacy concat-fusion checkpoint (the model's source is not
    in the reference, so the manifest is the only pin): tests/golden/legacy_concat_manifest.json."""
    import json
    sd = torch.load(f"{REF}/outputs/ecg_demo/ckpts/ecg_demo_best.pth", map_location="cpu")["model_state"]
    with open(os.path.join(HERE, "legacy_concat_manifest.json"), "w") as f:
        json.dump({k: [list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in sd.items()}, f, indent=0)


if __name__ == "__main__":
    main()
    legacy_concat_manifest()

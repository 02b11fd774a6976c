"""Harness that runs the reference's OWN script scripts/00_demo_inference.py (staged unmodified in oracle/_ref by
oracle/make_ref.py) end to end, optionally with the import swap of INTEGRATION.md section 2 applied WITHOUT editing the
script: `from src.models.ecg_cnn import ECGCNN` is made to resolve to ptbxl_multimodal_b200.ecg_cnn.

Test infrastructure only.  matplotlib is not in this image: the script's `import matplotlib.pyplot as plt` gets an empty
stub, and its plotting function (the only user of plt) is replaced by a recorder that keeps the CAM it was handed."""
import argparse
import contextlib
import importlib.util
import io
import os
import re
import sys
import types

import numpy as np
import torch

from conftest import GOLDEN
from oracle import make_ref

CLASSES = ["MI", "STTC", "HYP", "CD", "NORM"]


@contextlib.contextmanager
def _swapped_model_import(swap: bool):
    """INTEGRATION.md section 2, first row, as an import alias: while active, `src.models.ecg_cnn` IS the product module."""
    key = "src.models.ecg_cnn"
    saved = sys.modules.get(key)
    try:
        if swap:
            import ptbxl_multimodal_b200.ecg_cnn as product
            sys.modules[key] = product
        yield
    finally:
        if saved is not None:
            sys.modules[key] = saved
        else:
            sys.modules.pop(key, None)


def run_demo_inference(tmp_path, row: int, class_idx: int, swap: bool):
    """Runs main(args) of scripts/00_demo_inference.py on demo ECG `row` of tests/golden/demo_inputs.npz with the shipped
    baseline checkpoint.  Returns dict(probs=(5,) as printed (3 decimals), cam=(T,) tensor handed to the plot, device=str,
    model_module=str, stdout=str)."""
    R = make_ref.load()
    assert R is not None, "oracle/_ref not staged"
    for name in ("matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

    d = np.load(os.path.join(GOLDEN, "demo_inputs.npz"))
    demo_path = os.path.join(str(tmp_path), f"demo_row{row}.npz")
    np.savez(demo_path, ecg=d["ecg"][row], classes=np.array(CLASSES))
    args = argparse.Namespace(demo_path=demo_path, ckpt=os.path.join(GOLDEN, "ckpts", "ecg_baseline_best.pth"),
                              class_idx=class_idx, lead=0)

    seen = {}
    with _swapped_model_import(swap):
        spec = importlib.util.spec_from_file_location(f"_ref_demo_inference_{int(swap)}",
                                                      os.path.join(R.root, "scripts", "00_demo_inference.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        seen["model_module"] = mod.ECGCNN.__module__

        def record(ecg, cam, lead_idx, title, save_path):
            seen["cam"] = torch.as_tensor(cam).detach().cpu().clone()
            seen["title"] = title
        mod.plot_ecg_with_cam = record

        cwd = os.getcwd()
        out = io.StringIO()
        flags = (torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark)
        os.chdir(str(tmp_path))                         # the script writes under ./outputs/demo
        try:
            with contextlib.redirect_stdout(out):
                mod.main(args)
        finally:
            os.chdir(cwd)
            torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = flags   # set_seed() changes them
    text = out.getvalue()
    probs = [float(m.group(1)) for m in (re.search(rf"^\s+{c}: ([0-9.]+)$", text, re.M) for c in CLASSES)]
    dev = re.search(r"\[INFO\] Device: (\S+)", text).group(1)
    return {"probs": np.array(probs, dtype=np.float32), "cam": seen["cam"], "device": dev,
            "model_module": seen["model_module"], "stdout": text}


def load_script(name: str):
    """Import a staged reference script as a module WITHOUT running its main(): scripts/12_grad_cam_ecg_demo.py and
    13_grad_cam_af.py define their Grad-CAM classes (and compute_demo_importance) at module level.  What they import but
    this image lacks (matplotlib, and the Dataset classes whose modules need wfdb) is stubbed with empty modules: none of it
    is touched by the classes under test."""
    R = make_ref.load()
    assert R is not None, "oracle/_ref not staged"
    stubs = {"matplotlib": {}, "matplotlib.pyplot": {}, "seaborn": {},
             "src.datasets": {"__path__": []},
             "src.datasets.ptbxl_ecg_multimodal": {"PTBXLECGMultimodalDataset": object},
             "src.datasets.ptbxl_af": {"PTBXLAFDataset": object}}
    added = []
    for mod, attrs in stubs.items():
        if mod in sys.modules:
            continue
        try:
            if not mod.startswith("src."):
                importlib.import_module(mod)
                continue
        except ImportError:
            pass
        m = types.ModuleType(mod)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[mod] = m
        added.append(mod)
    if not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    arch = os.environ.get("TORCH_CUDA_ARCH_LIST")               # scripts/13:15 sets it at import; keep the test process as it was
    try:
        spec = importlib.util.spec_from_file_location("_ref_script_" + re.sub(r"\W", "_", name),
                                                      os.path.join(R.root, "scripts", name))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for m in added:
            if m.startswith("src."):
                sys.modules.pop(m, None)
        if arch is None:
            os.environ.pop("TORCH_CUDA_ARCH_LIST", None)
        else:
            os.environ["TORCH_CUDA_ARCH_LIST"] = arch
    return mod

"""Recipe for oracle/_ref: the UNMODIFIED reference modules of the hot path, staged for the reference arm.

TEST / MEASUREMENT INFRASTRUCTURE ONLY (never imported by ptbxl_multimodal_b200/).

The reference (cyu0330/ptbxl-multimodal) is pure Python over PyTorch: there is nothing to compile.  "Building"
the reference for this path therefore means staging the few files the path consists of -- byte for byte, from
where they lie under /root/reference -- into the git-ignored directory oracle/_ref/ so that they travel to the
GPU box with the snapshot (where /root/reference does not exist) and `bench.py --impl reference` /
`cpu_baseline` can time the reference's OWN code (`cpu_baseline.kind = "reference"`) instead of the oracle port.
No reference source enters the repository's history: oracle/_ref/ is listed in .gitignore (not in .gpurunignore).
(Staging sourceless bytecode instead -- py_compile output only, no .py at all -- was tried and works here, but *.pyc
files do not travel with the gpurun snapshot: on the GPU box verify() failed and the reference arm fell back to the
port, profiles/r02_gpu_reference_scripts.log vs the skipped run.  So the unmodified .py files are what is staged.)

Files staged:
    src/models/ecg_cnn.py, src/models/ecg_multimodal.py           -- the models
    src/training/loop.py, loop_demo.py, metrics.py                -- train_one_epoch[_demo], eval_one_epoch[_demo]
    src/interpretability/grad_cam_1d.py                           -- GradCAM1D
    scripts/00_demo_inference.py (+ src/utils/seed.py)            -- the one reference script that runs on the shipped demo
                                                                     data: driven end to end with the model import swapped
    scripts/12_grad_cam_ecg_demo.py, scripts/13_grad_cam_af.py    -- their script-local Grad-CAM classes / compute_demo_importance

    python oracle/make_ref.py [--reference /root/reference]

`load()` is what tests/ and bench.py call: it returns the staged modules (or None when oracle/_ref is absent, in
which case the caller uses the oracle port, which tests/test_reference_arm.py shows to be bit-identical).
"""
from __future__ import annotations

import hashlib
import importlib
import json
import os
import shutil
import sys
import types
from typing import Optional

HERE = os.path.dirname(os.path.abspath(__file__))
DEST = os.path.join(HERE, "_ref")
FILES = (
    "src/models/ecg_cnn.py",
    "src/models/ecg_multimodal.py",
    "src/training/loop.py",
    "src/training/loop_demo.py",
    "src/training/metrics.py",
    "src/interpretability/grad_cam_1d.py",
    # one runnable caller of the path, for the end-to-end import-swap test (tests/test_gpu_reference_scripts.py):
    "src/utils/seed.py",
    "scripts/00_demo_inference.py",
    # the script-local Grad-CAM classes V2 / V3 and compute_demo_importance, run as written on the product models:
    "scripts/12_grad_cam_ecg_demo.py",
    "scripts/13_grad_cam_af.py",
)


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build(reference: str = "/root/reference") -> Optional[str]:
    """Stage the reference files into oracle/_ref/ (idempotent).  Returns the directory, or None when the
    reference tree is not present (GPU box: the prebuilt directory is used as it travelled)."""
    if not os.path.isdir(os.path.join(reference, "src")):
        return DEST if os.path.exists(os.path.join(DEST, "MANIFEST.json")) else None
    manifest = {}
    for rel in FILES:
        src = os.path.join(reference, rel)
        dst = os.path.join(DEST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = _sha(dst)
    with open(os.path.join(DEST, "MANIFEST.json"), "w") as f:
        json.dump({"source": "cyu0330/ptbxl-multimodal (unmodified files)", "sha256": manifest}, f, indent=1)
    return DEST


def available() -> bool:
    return os.path.exists(os.path.join(DEST, "MANIFEST.json"))


def verify() -> bool:
    """The staged files still have the hashes recorded when they were copied (nobody edited the reference arm)."""
    if not available():
        return False
    with open(os.path.join(DEST, "MANIFEST.json")) as f:
        manifest = json.load(f)["sha256"]
    return all(os.path.exists(os.path.join(DEST, rel)) and _sha(os.path.join(DEST, rel)) == h
               for rel, h in manifest.items()) and set(manifest) == set(FILES)


def load() -> Optional[types.SimpleNamespace]:
    """Import the staged reference modules (`src.*` resolved from oracle/_ref only).  None if not staged."""
    if not verify():
        return None
    if DEST not in sys.path:
        sys.path.insert(0, DEST)
    importlib.invalidate_caches()
    ns = types.SimpleNamespace(
        ecg_cnn=importlib.import_module("src.models.ecg_cnn"),
        ecg_multimodal=importlib.import_module("src.models.ecg_multimodal"),
        loop=importlib.import_module("src.training.loop"),
        loop_demo=importlib.import_module("src.training.loop_demo"),
        grad_cam_1d=importlib.import_module("src.interpretability.grad_cam_1d"),
        root=DEST,
    )
    for m in (ns.ecg_cnn, ns.ecg_multimodal, ns.loop, ns.loop_demo, ns.grad_cam_1d):
        if not os.path.abspath(m.__file__).startswith(DEST + os.sep):
            raise RuntimeError(f"{m.__name__} resolved outside oracle/_ref: {m.__file__}")
    return ns


if __name__ == "__main__":
    ref = sys.argv[sys.argv.index("--reference") + 1] if "--reference" in sys.argv else "/root/reference"
    out = build(ref)
    print("oracle/_ref:", out, "verified" if verify() else "NOT AVAILABLE")

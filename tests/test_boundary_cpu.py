"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol
include/ecgb200.h declares; the nn.Module tree / state_dict / init order match the
reference (checked through the shipped checkpoints and the oracle's seeded init);
there is no CPU fallback."""
import ctypes
import os
import re

import pytest
import torch

import ptbxl_multimodal_b200 as P
from oracle import ecg_oracle as O
from conftest import ROOT, load_ckpt


def test_header_symbols_exported():
    hdr = open(os.path.join(ROOT, "include", "ecgb200.h")).read()
    declared = set(re.findall(r"\b(ecgb200_\w+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = ctypes.CDLL(os.path.join(ROOT, "ptbxl_multimodal_b200", "libecgb200.so"))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in ecgb200.h but not exported"
    assert declared == set(P.EXPORTED), declared ^ set(P.EXPORTED)
    assert lib.ecgb200_arch() == 1000


def test_argument_errors_without_gpu():
    from ptbxl_multimodal_b200._lib import lib
    assert lib.ecgb200_conv1d_fwd_f32(None, None, None, None, None, 1, 12, 32, 100, None) == -1
    assert lib.ecgb200_gradcam_f32(None, None, None, 0, None, None, None, 1, 256, 125, 5, 0, 1, 0.0, None) == -1
    assert lib.ecgb200_conv1d_stat_tiles(256, 1000) == 256 * 8
    assert lib.ecgb200_conv1d_wgrad_ws_bytes(256, 128, 256, 125) > 0


@pytest.mark.parametrize("ckpt,ctor", [
    ("ecg_baseline_best.pth", lambda: P.ECGCNN(12, 256, 5)),
    ("af_binary_best.pth", lambda: P.ECGCNN(12, 256, 1)),
    ("ecg_multimodal_best.pth", lambda: P.ECGMultimodal()),
])
def test_shipped_checkpoints_load_strict(ckpt, ctor):
    sd = load_ckpt(ckpt)
    model = ctor()
    assert list(model.state_dict().keys()) == list(sd.keys())
    model.load_state_dict(sd, strict=True)
    for k, v in model.state_dict().items():
        assert v.dtype == sd[k].dtype and v.shape == sd[k].shape and torch.equal(v, sd[k])


def test_backbone_submodule_load():
    # scripts/04_train_multimodal_prototype.py:149-156: baseline ckpt into model.ecg_backbone, strict=False
    mm = P.ECGMultimodal()
    res = mm.ecg_backbone.load_state_dict(load_ckpt("ecg_baseline_best.pth"), strict=False)
    assert set(res.unexpected_keys) == {"head.weight", "head.bias"} and not res.missing_keys


@pytest.mark.parametrize("kind,nl", [("cnn", 5), ("cnn", 1), ("mm", 5)])
def test_seeded_init_matches_reference_order(kind, nl):
    torch.manual_seed(42)
    model = P.ECGCNN(12, 256, nl) if kind == "cnn" else P.ECGMultimodal(num_labels=nl)
    sd = O.init_state_dict(kind, nl, seed=42)
    for k, v in model.state_dict().items():
        assert torch.equal(v, sd[k]), k


def test_attribute_tree():
    m = P.ECGCNN(12, 256, 5)
    assert isinstance(m.backbone[-1].net[0], torch.nn.Conv1d)
    last = None
    for mod in m.modules():                       # scripts/00_demo_inference.py:64-71
        if isinstance(mod, torch.nn.Conv1d):
            last = mod
    assert last is m.backbone[-1].net[0]
    mm = P.ECGMultimodal(ecg_feat_dim=128, demo_hidden_dim=32)
    assert mm.head.in_features == 128 and mm.film_gen.out_features == 256
    assert isinstance(mm.ecg_backbone.backbone[-1].net[0], torch.nn.Conv1d)


def test_no_cpu_fallback():
    m = P.ECGCNN(12, 256, 5)
    with pytest.raises(P.EcgB200Error):
        m(torch.randn(2, 12, 256))
    with pytest.raises(P.EcgB200Error):
        P.functional.binary_cross_entropy_with_logits(torch.zeros(2, 5), torch.zeros(2, 5))
    opt = P.FusedAdamW(m.parameters(), lr=1e-3)
    for p in m.parameters():
        p.grad = torch.zeros_like(p)
    with pytest.raises(P.EcgB200Error):
        opt.step()


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "ptbxl_multimodal_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("# oracle", ""), fn


def test_legacy_concat_model_matches_the_checkpoint_manifest():
    """the reconstructed concat-fusion module has exactly the legacy checkpoints' state_dict
    (tests/golden/legacy_concat_manifest.json, generated from outputs/ecg_demo/ckpts/ecg_demo_best.pth) and loads
    such a state_dict strictly."""
    import json
    import os
    import torch
    import ptbxl_multimodal_b200 as P
    from conftest import GOLDEN
    with open(os.path.join(GOLDEN, "legacy_concat_manifest.json")) as f:
        man = json.load(f)
    model = P.ECGDemoConcat()
    sd = model.state_dict()
    assert list(sd.keys()) == list(man.keys())
    for k, (shape, dtype) in man.items():
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == "torch." + dtype, k
    g = torch.Generator().manual_seed(0)
    fake = {k: (torch.randn(shape, generator=g) if dtype == "float32" else torch.zeros(shape, dtype=torch.int64))
            for k, (shape, dtype) in man.items()}
    model.load_state_dict(fake, strict=True)


def test_engines_refuse_cpu_models():
    """No CPU fallback: the CUDA-graph engines raise on a CPU model before touching the library."""
    import pytest
    import ptbxl_multimodal_b200 as P
    from ptbxl_multimodal_b200.step import TrainStep
    model = P.ECGCNN(12, 256, 5)
    with pytest.raises(P.EcgB200Error):
        P.InferStep(model, 4, 1000)
    with pytest.raises(P.EcgB200Error):
        P.InferStep(object(), 4, 1000)
    opt = P.FusedAdamW(model.parameters(), lr=1e-3)
    with pytest.raises(P.EcgB200Error):
        TrainStep(model, opt, 4, 1000)


def test_committed_bench_lines_follow_the_driver_contract():
    """The bench lines kept under profiles/ carry every key the driver / judge read (bench.py's JSON contract)."""
    import glob
    import json
    import os
    from conftest import ROOT
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r02_bench_n[128].json")) +
                   glob.glob(os.path.join(ROOT, "profiles", "r02_bench_c[234]_n[128].json")))
    assert len(files) >= 6
    configs = set()
    for f in files:
        with open(f) as fh:
            d = json.loads(fh.read().strip().splitlines()[-1])
        for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                  "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline", "step_roofline"):
            assert k in d, (f, k)
        assert d["higher_is_better"] is True and d["scaling"] in ("weak", "strong") and d["vs_baseline"] is None
        assert "workload" in d["config"] and d["dtype"] == "bf16" and d["data"] == "synthetic"
        assert {"batch_per_gpu", "global_batch", "seq_len", "parallelism", "precision"} <= set(d["config"])
        assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
        assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["gpu_launches"] > 0
        assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
        assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        configs.add(d["config"]["workload"].split(":")[0])
        if d["n_gpus"] == 1:
            assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
            assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
            assert d["gpu_reference"]["best"] > 0 and d["speedup_vs_gpu_reference"]["device_timed"] > 1
            if "layers" in d:                       # every C-ABI call of the step, with its roofline fraction
                calls = {r["call"] for r in d["layers"]}
                assert {"conv_fwd_L1", "conv_fwd_L4", "bn_relu_pool_L1", "bn_bwd_L4", "dgrad_L2", "wgrad_L1", "wgrad_L4"} <= calls
                assert all({"call", "us", "model_us", "frac", "bound"} <= set(r) for r in d["layers"])
                assert len(d["layers_worst3"]) == 3
    assert configs == {"configs[1]", "configs[2]", "configs[3]", "configs[4]"}
    with open(os.path.join(ROOT, "profiles", "r02_bench_n1.json")) as fh:
        head = json.loads(fh.read().strip().splitlines()[-1])
    assert head["roofline"]["traffic"] is not None and head["roofline"]["traffic_source"].startswith("profiles/")
    with open(os.path.join(ROOT, "profiles", "r02_bench_reference_arm.json")) as fh:
        r = json.loads(fh.read().strip().splitlines()[-1])
    assert r["impl"] == "reference" and r["e2e"]["h2d_bytes_per_step"] == 0 and r["cpu_baseline"]["kind"] == "port"
    assert set(r["config"]) <= set(head["config"])          # the two arms name the workload with the same keys
    # the last lines of the round: the reference arm runs the UNMODIFIED reference staged in oracle/_ref, and both arms of the
    # same command line print one and the same `config` object (run details moved under `method`)
    with open(os.path.join(ROOT, "profiles", "r02_bench_n1_final.json")) as fh:
        head = json.loads(fh.read().strip().splitlines()[-1])
    with open(os.path.join(ROOT, "profiles", "r02_bench_reference_arm_final.json")) as fh:
        r = json.loads(fh.read().strip().splitlines()[-1])
    assert r["impl"] == "reference" and r["cpu_baseline"]["kind"] == "reference" == head["cpu_baseline"]["kind"]
    assert r["config"] == head["config"] and r["metric"] == head["metric"] and r["unit"] == head["unit"]
    assert "l2" in head["config"] and {"repeats", "timing", "e2e_input"} <= set(head["method"])
    assert head["gpu_launches"] > 0 and head["e2e"]["h2d_bytes_per_step"] > 0 and len(head["layers"]) >= 20


def test_pair_kernel_selection_rules_on_the_host():
    """conv_pair_cfg / wgrad_pair_cfg run on the host (148 SMs assumed without a device): which shapes of the network take the
    CTA-pair kernels.  A pair grid is even and never larger than the SM count; small problems stay with the one-SM kernels."""
    from ptbxl_multimodal_b200._lib import lib
    try:
        def parts(mask, *shape):
            lib.ecgb200_debug_set_conv_pair(mask)
            return lib.ecgb200_conv1d_stat_parts_bf16(*shape)
        # benchmark batch: blocks 3 / 4 forward (B, Ci, Co, L) as pairs; their grids are even
        for shape in [(256, 64, 128, 250), (256, 128, 256, 125), (256, 256, 128, 125), (256, 128, 64, 250)]:
            p = parts(3, *shape)
            assert p % 2 == 0 and 0 < p <= 148, (shape, p)
        # stem layers (weights fit in shared memory) and one-tile-per-SM batches: the same grid with and without the switch
        for shape in [(256, 16, 32, 1000), (256, 32, 64, 500), (256, 64, 32, 500), (64, 128, 256, 125), (64, 64, 128, 250)]:
            assert parts(3, *shape) == parts(0, *shape), shape
        # the 256 -> 128 channel dgrad with two input buffers (multi-round) keeps the one-SM kernel: its ring would be too short
        assert parts(3, 1024, 256, 128, 125) == parts(0, 1024, 256, 128, 125)
        assert parts(3, 1024, 128, 256, 125) % 2 == 0
        # forced (tests): pairs wherever the kernel can run
        assert parts(7, 64, 128, 256, 125) % 2 == 0 and parts(7, 64, 128, 256, 125) <= 148
        # the split-K workspace never shrinks below what either weight-gradient kernel needs
        for mask in (0, 3, 7):
            lib.ecgb200_debug_set_conv_pair(mask)
            assert lib.ecgb200_conv1d_wgrad_bf16_ws_bytes(256, 128, 256, 125) >= 18 * 256 * 128 * 16 * 4
    finally:
        lib.ecgb200_debug_set_conv_pair(3)

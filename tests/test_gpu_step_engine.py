"""GPU: the CUDA-graph TrainStep engine against (a) the nn.Module/autograd path built from the
same kernels (must agree bit-for-bit) and (b) the CPU oracle (tolerances as in test_gpu_parity)."""
import pytest
import torch

import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200 import functional as Fn
from ptbxl_multimodal_b200.step import TrainStep
from oracle import ecg_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel_inf(a, b):
    a = a.detach().double().cpu(); b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _mk(kind, nl):
    torch.manual_seed(42)
    m = P.ECGCNN(12, 256, nl) if kind == "cnn" else P.ECGMultimodal(num_labels=nl)
    return m.to(DEV).train()


@pytest.mark.parametrize("kind,nl,B,T,lr", [("cnn", 5, 8, 1000, 1.5e-3), ("mm", 5, 6, 1000, 1e-4), ("cnn", 1, 2, 5000, 1e-3)])
@pytest.mark.parametrize("graph", [True, False])
def test_engine_equals_module_path(kind, nl, B, T, lr, graph):
    batch = O.synth_batch(B, T, nl, seed=5, with_demo=(kind == "mm"))
    x, y = batch[0].to(DEV), batch[-1].to(DEV)
    demo = batch[1].to(DEV) if kind == "mm" else None
    ma, mb = _mk(kind, nl), _mk(kind, nl)
    oa = P.FusedAdamW(ma.parameters(), lr=lr, weight_decay=1e-4)
    ob = P.FusedAdamW(mb.parameters(), lr=lr, weight_decay=1e-4)
    eng = TrainStep(mb, ob, B, T, use_graph=graph)
    for s in range(3):
        oa.zero_grad()
        la = Fn.binary_cross_entropy_with_logits(ma(x) if demo is None else ma(x, demo), y)
        la.backward()
        oa.step()
        lb = eng(x, y, demo).clone()
        assert float(la.detach()) == float(lb), (s, float(la.detach()), float(lb))
    sa, sb = ma.state_dict(), mb.state_dict()
    for k in sa:
        assert torch.equal(sa[k], sb[k]), k
    assert ob.param_groups[0]["step"] == 3 and int(eng.step_dev) == 3
    assert eng.launches_per_step > 30


def test_engine_matches_oracle_and_checkpoint_roundtrip(tmp_path):
    B, T = 4, 1000
    x, y = O.synth_batch(B, T, 5, seed=1)
    sd = O.init_state_dict("cnn", 5, seed=42)
    st = O.AdamWState(sd, 1.5e-3, 1e-4)
    model = _mk("cnn", 5)
    opt = P.FusedAdamW(model.parameters(), lr=1.5e-3, weight_decay=1e-4)
    eng = TrainStep(model, opt, B, T)
    ref = O.train_step(sd, x, y, st)
    loss = eng(x.pin_memory(), y.pin_memory())           # host (pinned) inputs: the public entry
    assert abs(float(loss) - float(ref["loss"])) < 1e-5
    assert rel_inf(eng.logits, ref["logits"]) < 1e-4
    # the reference's checkpoint format round-trips through the flat parameter buffer
    path = tmp_path / "ck.pth"
    torch.save({"model_state": model.state_dict(), "classes": ["MI", "STTC", "HYP", "CD", "NORM"]}, path)
    fresh = P.ECGCNN(12, 256, 5)
    fresh.load_state_dict(torch.load(path, map_location="cpu")["model_state"], strict=True)
    for k, v in fresh.state_dict().items():
        assert torch.equal(v, model.state_dict()[k].cpu()), k
    # loading a checkpoint into the engine's model keeps the flat views alive
    model.load_state_dict(sd_to := {k: v.clone() for k, v in fresh.state_dict().items()})
    assert model.proj.weight.data_ptr() == eng.P.data_ptr() + 4 * eng.seg["proj.weight"].off
    del sd_to


def test_engine_rejects_wrong_shapes_and_cpu():
    model = _mk("cnn", 5)
    opt = P.FusedAdamW(model.parameters(), lr=1e-3)
    eng = TrainStep(model, opt, 4, 256)
    with pytest.raises(P.EcgB200Error):
        eng(torch.zeros(3, 12, 256, device=DEV), torch.zeros(3, 5, device=DEV))
    with pytest.raises(P.EcgB200Error):
        TrainStep(P.ECGCNN(12, 256, 5), P.FusedAdamW(P.ECGCNN(12, 256, 5).parameters()), 4, 256)

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


def _cos(a, b):
    a = a.detach().double().cpu().flatten(); b = b.detach().double().cpu().flatten()
    return float((a @ b) / (a.norm() * b.norm()).clamp_min(1e-300))


@pytest.mark.parametrize("kind,nl,B,T,lr", [("cnn", 5, 8, 1000, 1.5e-3), ("mm", 5, 6, 1000, 1e-4), ("cnn", 1, 2, 5000, 1e-3)])
def test_bf16_engine(kind, nl, B, T, lr):
    """bf16 tensor-core mode (tcgen05 convs, bf16 activations/gradients in HBM, fp32 master
    weights, fp32 BN statistics / head / loss / optimizer).  The reference has no bf16 mode
    (SURVEY D5); its definition is oracle.bf16_train_step (fp32 algorithm + bf16 rounding at the
    storage points).  Stated tolerances:
      vs the bf16 definition : logits rel_inf <= 5e-3, loss <= 1e-3 rel,
                               every gradient tensor 1-cos <= 3e-3 and rel_inf <= 0.3
      vs the fp32 oracle     : logits rel_inf <= 3e-2, loss <= 2e-2 rel, gradient 1-cos <= 3e-2
      vs the definition replayed on the engine's own stored activations (routing forced): every gradient tensor
                               rel_inf <= 2e-2 (the SURVEY 8c bar), logits <= 1e-3
    Element-wise agreement of conv gradients cannot be tighter: a conv output that lands within
    fp32 accumulation error of a bf16 rounding boundary rounds differently in two correct
    implementations, and the 1-ulp (0.4 %) difference can flip ReLU / MaxPool routing downstream.
    Measured on B200: 1-cos ~8e-4 vs the definition, while the definition itself sits at
    1-cos ~1e-2 from fp32.  (conv-bias grads are ~0 on both sides: absolute bound only)."""
    batch = O.synth_batch(B, T, nl, seed=5, with_demo=(kind == "mm"))
    x, y = batch[0], batch[-1]
    demo = batch[1] if kind == "mm" else None
    sd = O.init_state_dict(kind, nl, seed=42)
    ref32 = O.train_step(O.clone_sd(sd), x, y, None, demo=demo)
    ref16 = O.bf16_train_step(sd, x, y, demo=demo)
    model = _mk(kind, nl)
    opt = P.FusedAdamW(model.parameters(), lr=lr, weight_decay=1e-4)
    eng = TrainStep(model, opt, B, T, precision="bf16", use_graph=False)
    loss = float(eng(x.to(DEV), y.to(DEV), demo.to(DEV) if demo is not None else None))
    assert rel_inf(eng.logits, ref16["logits"]) < 5e-3, rel_inf(eng.logits, ref16["logits"])
    assert abs(loss - float(ref16["loss"])) < 1e-3 * float(ref16["loss"])
    assert rel_inf(eng.logits, ref32["logits"]) < 3e-2
    assert abs(loss - float(ref32["loss"])) < 2e-2 * float(ref32["loss"])
    gmax = max(float(g.abs().max()) for g in ref32["grads"].values())
    w16 = wcos = 0.0
    for k, p in model.named_parameters():
        if k.endswith("net.0.bias"):
            assert float(p.grad.abs().max()) < 1e-2 * gmax, k
            continue
        r = rel_inf(p.grad, ref16["grads"][k])
        c = _cos(p.grad, ref32["grads"][k])
        c16 = _cos(p.grad, ref16["grads"][k])
        print(f"   {k:40s} rel_inf(def) {r:.2e}  1-cos(def) {1 - c16:.2e}  1-cos(fp32) {1 - c:.2e}  "
              f"[def vs fp32: 1-cos {1 - _cos(ref16['grads'][k], ref32['grads'][k]):.2e}]")
        w16, wcos = max(w16, r), max(wcos, 1 - c)
        assert 1 - c16 < 3e-3 and r < 0.3, (k, r, 1 - c16)
        assert 1 - c < 3e-2, (k, 1 - c)
    print(f"bf16 {kind}: worst grad rel_inf vs bf16 definition {w16:.2e}; worst 1-cos vs fp32 oracle {wcos:.2e}")
    # ---- rounding isolated from routing (SURVEY 8c asks rel_inf <= 2e-2 on gradients): replay the definition's backward
    # with the values the engine itself STORED (bf16 conv outputs and pooled activations), so that every ReLU / MaxPool
    # decision is the engine's own.  What is left is bf16 rounding of dy / dp and fp32 summation order.
    def unblock(t, c):                                   # [B][C/8][L][8] bf16 -> (B, C, L) fp32
        return t.float().permute(0, 1, 3, 2).reshape(t.shape[0], -1, t.shape[2])[:, :c].contiguous().cpu()
    fc = [unblock(eng.ybuf[l], eng.chan[l + 1]) for l in range(4)]
    fp = [unblock(eng.acts[l + 1], eng.chan[l + 1]) for l in range(3)]
    refF = O.bf16_train_step(sd, x, y, demo=demo, forced_conv=fc, forced_pool=fp)
    wF = 0.0
    for k, p in model.named_parameters():
        if k.endswith("net.0.bias"):
            continue
        r = rel_inf(p.grad, refF["grads"][k])
        wF = max(wF, r)
        assert r < 2e-2, (k, r)
    assert rel_inf(eng.logits, refF["logits"]) < 1e-3
    print(f"bf16 {kind}: worst grad rel_inf vs the definition replayed on the engine's own activations {wF:.2e} (<= 2e-2)")
    # ---- the other candidate oracle of SURVEY 8c, reported beside it: the stock modules under torch.autocast(bf16)
    # (cuDNN bf16 conv, fp32 BatchNorm) on this GPU -- rounding points differ from the storage-rounding definition
    # (autocast rounds the BN output before ReLU/pool and keeps bf16 through the pool), so it sits further away.
    from oracle import torch_stock as S
    stock = S.build(kind, sd, nl).to(DEV).train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = stock(x.to(DEV)) if demo is None else stock(x.to(DEV), demo.to(DEV))
    la = torch.nn.functional.binary_cross_entropy_with_logits(out.float(), y.to(DEV))
    la.backward()
    wa = wd = 0.0
    for k, p in stock.named_parameters():
        if k.endswith("net.0.bias"):
            continue
        wa = max(wa, 1 - _cos(model.get_parameter(k).grad, p.grad))
        wd = max(wd, 1 - _cos(ref16["grads"][k], p.grad))
    print(f"bf16 {kind}: autocast(bf16) stock modules: worst gradient 1-cos engine vs autocast {wa:.2e}, definition vs autocast "
          f"{wd:.2e}; logits rel_inf engine vs autocast {rel_inf(eng.logits, out.float()):.2e}")
    assert wa < 6e-2 and rel_inf(eng.logits, out.float()) < 6e-2
    # graph replay of the bf16 step is bit-identical to the eager enqueue
    m2 = _mk(kind, nl)
    o2 = P.FusedAdamW(m2.parameters(), lr=lr, weight_decay=1e-4)
    e2 = TrainStep(m2, o2, B, T, precision="bf16", use_graph=True)
    l2 = float(e2(x.to(DEV), y.to(DEV), demo.to(DEV) if demo is not None else None))
    assert l2 == loss
    for (k, a), (_, b) in zip(model.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k


@pytest.mark.parametrize("nl,B,T", [(5, 256, 1000), (1, 64, 5000)])
def test_bf16_engine_at_benchmark_sizes(nl, B, T):
    """The benchmarked configurations themselves (BASELINE.json configs[1]: batch 256 x 12 x 1000; configs[3]: one
    rank's 64 x 12 x 5000 AF windows): the bf16 tcgen05 step against the fp32 exact engine (itself held to the CPU
    oracle at 1e-4 on small cases) from the same weights and batch.  Same stated bf16 tolerances as above:
    logits rel_inf <= 3e-2, loss <= 2e-2 relative, every gradient tensor 1-cos <= 3e-2; plus the size-independent
    property that one step leaves every parameter finite and the graph replay is deterministic."""
    x, y = O.synth_batch(B, T, nl, seed=11)
    xd, yd = x.to(DEV), y.to(DEV)
    res = {}
    for prec in ("fp32", "bf16"):
        model = _mk("cnn", nl)
        opt = P.FusedAdamW(model.parameters(), lr=1.5e-3, weight_decay=1e-4)
        eng = TrainStep(model, opt, B, T, precision=prec)
        loss = float(eng(xd, yd))
        torch.cuda.synchronize()
        grads = {k: eng.G[s.off:s.off + s.n].clone() for k, s in eng.seg.items()}
        res[prec] = (loss, eng.logits.clone(), grads, eng.P[:eng.total].clone())
        assert torch.isfinite(eng.P).all()
        if prec == "bf16":
            m2 = _mk("cnn", nl)
            e2 = TrainStep(m2, P.FusedAdamW(m2.parameters(), lr=1.5e-3, weight_decay=1e-4), B, T, precision="bf16")
            l2 = float(e2(xd, yd))
            assert l2 == loss and torch.equal(e2.P, eng.P)           # run-to-run deterministic (fixed-order reductions)
        del eng
    (l32, lg32, g32, _), (l16, lg16, g16, _) = res["fp32"], res["bf16"]
    assert rel_inf(lg16, lg32) < 3e-2, rel_inf(lg16, lg32)
    assert abs(l16 - l32) < 2e-2 * l32
    gmax = max(float(g.abs().max()) for g in g32.values())
    worst = 0.0
    for k in g32:
        if k.endswith("net.0.bias"):                                 # ~0 on both sides (BatchNorm follows the conv)
            assert float(g16[k].abs().max()) < 1e-2 * gmax, k
            continue
        c = 1 - _cos(g16[k], g32[k])
        worst = max(worst, c)
        assert c < 3e-2, (k, c)
    print(f"bf16 vs fp32 engine at B={B}, T={T}: logits rel_inf {rel_inf(lg16, lg32):.2e}, worst gradient 1-cos {worst:.2e}")


def test_reference_shaped_loops_run_on_the_engines():
    """train_one_epoch(..., engine=TrainStep) / eval_one_epoch(..., engine=InferStep): same return contracts as the
    reference loops (loop.py:14-38,41-73); the engine epoch equals the same steps driven by hand."""
    B, T, n = 8, 1000, 24
    xs = torch.randn(n, 12, T, generator=torch.Generator().manual_seed(3))
    ys = (torch.rand(n, 5, generator=torch.Generator().manual_seed(4)) < 0.3).float()
    loader = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xs, ys), batch_size=B, drop_last=True)
    model = _mk("cnn", 5)
    opt = P.FusedAdamW(model.parameters(), lr=1e-3, weight_decay=1e-4)
    eng = TrainStep(model, opt, B, T, precision="bf16")
    loss = P.train_one_epoch(model, loader, opt, DEV, engine=eng)
    m2 = _mk("cnn", 5)
    o2 = P.FusedAdamW(m2.parameters(), lr=1e-3, weight_decay=1e-4)
    e2 = TrainStep(m2, o2, B, T, precision="bf16")
    tot = 0.0
    for x, y in loader:
        tot += float(e2(x.to(DEV), y.to(DEV))) * B
    assert abs(loss - tot / n) < 1e-6 * max(1.0, abs(loss))
    for (k, a), (_, b) in zip(model.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    metrics = P.eval_one_epoch(model, loader, DEV, engine=P.InferStep(model, B, T))
    assert set(metrics) == {"auroc_macro", "auprc_macro", "f1_macro", "bce_loss"}
    ragged = torch.utils.data.DataLoader(torch.utils.data.TensorDataset(xs[:12], ys[:12]), batch_size=B)
    with pytest.raises(ValueError):
        P.train_one_epoch(model, ragged, opt, DEV, engine=eng)

"""GPU parity tests (fp32 exact path): every ecgb200 kernel and the assembled models,
called through the C ABI, against the CPU oracle on the same seeded inputs.

Tolerance (north_star): fp32 logits / gradients / Grad-CAM within 1e-4 relative, measured
as rel_inf = max|a-b| / max|b| per tensor against the CPU fp32 oracle; thresholded
predictions and Grad-CAM argmax indices bit-exact (predictions whose oracle logit lies
within 1e-5 of the decision boundary are reported, not asserted)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

import ptbxl_multimodal_b200 as P
from ptbxl_multimodal_b200 import functional as Fn
from oracle import ecg_oracle as O
from conftest import load_ckpt

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL = 1e-4


def rel_inf(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def gen(*shape, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g) * scale


# ------------------------------------------------------------------ single kernels
@pytest.mark.parametrize("B,Ci,Co,L", [(3, 12, 32, 1000), (2, 32, 64, 500), (2, 64, 128, 250),
                                       (2, 128, 256, 125), (1, 12, 32, 77), (2, 5, 32, 16),
                                       (1, 128, 256, 625), (2, 3, 20, 130)])
def test_conv1d_fwd_bwd(B, Ci, Co, L):
    x = gen(B, Ci, L, seed=1).requires_grad_(True)
    w = gen(Co, Ci, 15, seed=2, scale=0.1).requires_grad_(True)
    b = gen(Co, seed=3, scale=0.1).requires_grad_(True)
    dy = gen(B, Co, L, seed=4)
    y_ref = F.conv1d(x, w, b, padding=7)
    y_ref.backward(dy)
    xg, wg, bg = (t.detach().to(DEV).requires_grad_(True) for t in (x, w, b))
    y, stat = Fn.Conv1dK15Fn.apply(xg, wg, bg, True)
    y.backward(dy.to(DEV))
    assert rel_inf(y, y_ref) < 1e-5
    assert rel_inf(wg.grad, w.grad) < 2e-5
    assert rel_inf(bg.grad, b.grad) < 2e-5
    assert rel_inf(xg.grad, x.grad) < 2e-5
    # epilogue statistics reproduce mean / biased variance of y
    ntiles = stat.shape[2]
    tpr = ntiles // B
    cnt = torch.tensor([min(128, L - i * 128) for i in range(tpr)] * B, dtype=torch.float64)
    s = stat[0].double().cpu(); m2 = stat[1].double().cpu()
    mean = s.sum(1) / (B * L)
    var = (m2 + cnt * (s / cnt - mean[:, None]) ** 2).sum(1) / (B * L)
    yr = y_ref.detach().double()
    assert torch.allclose(mean, yr.mean(dim=(0, 2)), atol=1e-5)
    assert torch.allclose(var, yr.var(dim=(0, 2), unbiased=False), rtol=1e-4)


@pytest.mark.parametrize("B,C,L", [(4, 32, 1000), (3, 64, 250), (5, 256, 125), (2, 128, 31), (2, 32, 2)])
@pytest.mark.parametrize("train", [True, False])
@pytest.mark.parametrize("use_gap", [False, True])
def test_bn_relu_pool(B, C, L, train, use_gap):
    y = (gen(B, C, L, seed=5) * 1.7 + 0.3).requires_grad_(True)
    gamma = (1 + 0.2 * gen(C, seed=6)).requires_grad_(True)
    beta = (0.1 * gen(C, seed=7)).requires_grad_(True)
    rm0 = 0.1 * gen(C, seed=8); rv0 = 1 + 0.3 * torch.rand(C, generator=torch.Generator().manual_seed(9))
    rm, rv = rm0.clone(), rv0.clone()
    h = F.batch_norm(y, rm, rv, gamma, beta, training=train, momentum=0.1, eps=1e-5)
    p_ref = F.max_pool1d(F.relu(h), 2)
    out_ref = p_ref.mean(dim=2) if use_gap else p_ref
    dout = gen(*out_ref.shape, seed=10)
    out_ref.backward(dout)

    yg, gg, bg = (t.detach().to(DEV).requires_grad_(True) for t in (y, gamma, beta))
    rmg, rvg = rm0.to(DEV), rv0.to(DEV)
    nbt = torch.zeros((), dtype=torch.int64, device=DEV)
    p, gap = Fn.BnReluPoolFn.apply(yg, gg, bg, rmg, rvg, nbt, None, train, 0.1, 1e-5, use_gap)
    out = gap if use_gap else p
    out.backward(dout.to(DEV))
    assert rel_inf(out, out_ref) < 1e-5
    assert rel_inf(p, p_ref) < 1e-5
    assert rel_inf(yg.grad, y.grad) < TOL
    assert rel_inf(gg.grad, gamma.grad) < TOL
    assert rel_inf(bg.grad, beta.grad) < TOL
    if train:
        assert rel_inf(rmg, rm) < 1e-5 and rel_inf(rvg, rv) < 1e-5 and int(nbt) == 1
    else:
        assert torch.equal(rmg.cpu(), rm0) and int(nbt) == 0


def test_maxpool_tie_goes_to_first_index():
    # ties inside a pool pair survive the (monotone) BN affine map: gradient must go to the first index
    y = torch.tensor([[[1.0, 1.0, 0.0, 0.0, 5.0], [2.0, 3.0, 3.0, 3.0, -1.0]]]).requires_grad_(True)
    C = 2
    one, zero = torch.ones(C), torch.zeros(C)
    h = F.max_pool1d(F.relu(F.batch_norm(y, zero.clone(), one.clone(), one, zero, False, 0.1, 1e-5)), 2)
    h.backward(torch.ones_like(h))
    yg = y.detach().to(DEV).requires_grad_(True)
    p, _ = Fn.BnReluPoolFn.apply(yg, one.to(DEV), zero.to(DEV), zero.to(DEV), one.to(DEV),
                                 torch.zeros((), dtype=torch.int64, device=DEV), None, False, 0.1, 1e-5, False)
    p.backward(torch.ones_like(p))
    assert torch.allclose(p.cpu(), h.detach(), rtol=1e-6)
    assert torch.equal(yg.grad.cpu() != 0, y.grad != 0)          # same routing: [1,0,0,0,0], [0,1,1,0,0]
    assert torch.allclose(yg.grad.cpu(), y.grad, rtol=1e-6)


@pytest.mark.parametrize("M,K,N,act", [(256, 256, 256, 0), (7, 256, 5, 0), (33, 5, 64, 1), (64, 64, 512, 0), (9, 64, 64, 1)])
def test_linear(M, K, N, act):
    x = gen(M, K, seed=11).requires_grad_(True)
    w = gen(N, K, seed=12, scale=0.1).requires_grad_(True)
    b = gen(N, seed=13, scale=0.1).requires_grad_(True)
    y_ref = F.linear(x, w, b)
    if act:
        y_ref = F.relu(y_ref)
    dy = gen(M, N, seed=14)
    y_ref.backward(dy)
    xg, wg, bg = (t.detach().to(DEV).requires_grad_(True) for t in (x, w, b))
    y = Fn.linear(xg, wg, bg, act)
    y.backward(dy.to(DEV))
    for a, r in ((y, y_ref), (xg.grad, x.grad), (wg.grad, w.grad), (bg.grad, b.grad)):
        assert rel_inf(a, r) < 1e-5


def test_film_and_bce():
    z = gen(9, 256, seed=15).requires_grad_(True)
    film = gen(9, 512, seed=16).requires_grad_(True)
    gmm, bta = torch.chunk(film, 2, dim=-1)
    ref = (1.0 + torch.tanh(gmm)) * z + bta
    d = gen(9, 256, seed=17)
    ref.backward(d)
    zg, fg = z.detach().to(DEV).requires_grad_(True), film.detach().to(DEV).requires_grad_(True)
    out = Fn.FilmFn.apply(zg, fg)
    out.backward(d.to(DEV))
    assert rel_inf(out, ref) < 1e-5 and rel_inf(zg.grad, z.grad) < 1e-5 and rel_inf(fg.grad, film.grad) < 1e-5

    lo = (gen(64, 5, seed=18) * 6).requires_grad_(True)
    y = (torch.rand(64, 5, generator=torch.Generator().manual_seed(19)) < 0.3).float()
    l_ref = F.binary_cross_entropy_with_logits(lo, y)
    (3.0 * l_ref).backward()
    lg = lo.detach().to(DEV).requires_grad_(True)
    l = Fn.binary_cross_entropy_with_logits(lg, y.to(DEV))
    (3.0 * l).backward()
    assert abs(float(l) - float(l_ref)) < 1e-6
    assert rel_inf(lg.grad, lo.grad) < 1e-5
    assert rel_inf(Fn.sigmoid(lg.detach()), torch.sigmoid(lo.detach())) < 1e-6


def test_fused_adamw_matches_torch():
    ps = [gen(491520 // 8, seed=20), gen(37, seed=21), gen(256, 5, seed=22)]
    ref = [p.clone().requires_grad_(True) for p in ps]
    mine = [p.clone().to(DEV).requires_grad_(True) for p in ps]
    o_ref = torch.optim.AdamW(ref, lr=1.5e-3, weight_decay=1e-4)
    o_my = P.FusedAdamW(mine, lr=1.5e-3, weight_decay=1e-4)
    for step in range(4):
        for i, (r, m) in enumerate(zip(ref, mine)):
            g = gen(*r.shape, seed=100 + 10 * step + i) * (10.0 ** (-step))
            r.grad = g.clone(); m.grad = g.to(DEV)
        o_ref.step(); o_my.step()
    for r, m in zip(ref, mine):
        assert rel_inf(m, r) < 1e-6


def test_zscore():
    x = gen(3, 12, 5000, seed=23) * 3 + 1.5
    ref = (x - x.mean(dim=-1, keepdim=True)) / (x.std(dim=-1, unbiased=False, keepdim=True) + 1e-6)
    assert rel_inf(Fn.zscore(x.to(DEV)), ref) < 1e-5


# ------------------------------------------------------------------ assembled models
def _model(kind, nl):
    torch.manual_seed(42)
    m = P.ECGCNN(12, 256, nl) if kind == "cnn" else P.ECGMultimodal(num_labels=nl)
    return m.to(DEV)


@pytest.mark.parametrize("tag,kind", [("train_cnn", "cnn"), ("train_mm", "mm"),
                                      ("train_af", "cnn"), ("train_cnn_t250", "cnn")])
def test_train_steps_match_oracle_and_golden(golden, tag, kind):
    """Reference loop body (src/training/loop.py:22-36 / loop_demo.py:25-41) for a few steps.
    The golden batches were picked so that every ReLU / MaxPool routing decision of step 0 has
    a margin > 1e-6 (stored in the fixture), i.e. beyond fp32 rounding differences; with that,
    all step-0 gradients must agree to 1e-4.  Later steps start from parameters that already
    differ by Adam's normalisation of ~0 gradients (m/(sqrt(v)+eps) maps a 1e-9 rounding-noise
    gradient to an O(lr) step), so they are held to looser, stated bounds."""
    B, T, nl, lr, wd, steps = golden[f"{tag}/cfg"]
    nl, steps, lr, wd = int(nl), int(steps), float(lr), float(wd)
    assert golden[f"{tag}/margin"][0] > 1e-6
    x = torch.from_numpy(golden[f"{tag}/x"]); y = torch.from_numpy(golden[f"{tag}/y"])
    demo = torch.from_numpy(golden[f"{tag}/demo"]) if kind == "mm" else None
    sd = O.init_state_dict(kind, nl, seed=42)
    st = O.AdamWState(sd, lr, wd)
    model = _model(kind, nl)
    opt = P.FusedAdamW(model.parameters(), lr=lr, weight_decay=wd)
    model.train()
    xg, yg = x.to(DEV), y.to(DEV)
    dg = demo.to(DEV) if demo is not None else None
    for s in range(steps):
        o = O.train_step(sd, x, y, st, demo=demo)
        opt.zero_grad()
        logits = model(xg) if dg is None else model(xg, dg)
        loss = Fn.binary_cross_entropy_with_logits(logits, yg)
        loss.backward()
        tol = TOL if s == 0 else 2e-2
        assert rel_inf(logits, o["logits"]) < tol, (tag, s, rel_inf(logits, o["logits"]))
        assert rel_inf(logits, torch.from_numpy(golden[f"{tag}/step{s}/logits"])) < tol
        assert abs(float(loss.detach()) - float(o["loss"])) < tol * max(1.0, abs(float(o["loss"])))
        if s == 0:
            gmax = max(float(g.abs().max()) for g in o["grads"].values())
            for k, p in model.named_parameters():
                ref = o["grads"][k]
                if k.endswith("net.0.bias"):
                    # conv bias feeding a train-mode BN: the exact gradient is 0, both sides hold
                    # pure rounding noise -> absolute bound only
                    assert float(p.grad.abs().max()) < 1e-5 * gmax, k
                else:
                    assert rel_inf(p.grad, ref) < TOL, (tag, k, rel_inf(p.grad, ref))
        opt.step()
    for k, v in model.state_dict().items():
        if k.endswith("num_batches_tracked"):
            assert int(v) == steps and v.dtype == torch.int64
        elif k.endswith("running_mean") or k.endswith("running_var"):
            assert rel_inf(v, sd[k]) < 2e-2, (tag, k, rel_inf(v, sd[k]))
        else:
            d = (v.cpu() - sd[k]).abs()
            # no element can drift further than Adam's per-step bound, and the bulk agrees tightly
            assert float(d.max()) <= 2.5 * lr * steps, (tag, k, float(d.max()))
            if not k.endswith("net.0.bias"):
                assert float(d.flatten().kthvalue(max(1, int(0.99 * d.numel()))).values) < 0.25 * lr, (tag, k)


def test_adamw_step_on_oracle_gradients(golden):
    """Optimizer parity isolated from the backward pass: oracle gradients in, one fused AdamW
    launch, parameters must match torch.optim.AdamW semantics to 1e-6."""
    tag = "train_cnn"
    B, T, nl, lr, wd, steps = golden[f"{tag}/cfg"]
    x = torch.from_numpy(golden[f"{tag}/x"]); y = torch.from_numpy(golden[f"{tag}/y"])
    sd = O.init_state_dict("cnn", int(nl), seed=42)
    st = O.AdamWState(sd, float(lr), float(wd))
    model = _model("cnn", int(nl))
    opt = P.FusedAdamW(model.parameters(), lr=float(lr), weight_decay=float(wd))
    for s in range(3):
        o = O.train_step(sd, x, y, st)
        for k, p in model.named_parameters():
            p.grad = o["grads"][k].to(DEV)
        opt.step()
    for k, p in model.named_parameters():
        assert rel_inf(p, sd[k]) < 1e-6, k


def test_full_size_config2_step_properties():
    """BASELINE config 2 at full size (B=256, 12x1000): too many activations for every routing
    decision to be margin-safe, so check size-independent properties instead: loss / logits
    against the oracle, gradient linearity in the loss scale, sum(d loss / d conv-bias) ~ 0,
    BN running statistics, and run-to-run bit reproducibility."""
    x, y = O.synth_batch(256, 1000, 5, seed=0)
    sd = O.init_state_dict("cnn", 5, seed=42)
    ref = O.train_step(sd, x, y, None)
    xg, yg = x.to(DEV), y.to(DEV)

    def run(scale):
        model = _model("cnn", 5).train()
        logits = model(xg)
        loss = Fn.binary_cross_entropy_with_logits(logits, yg)
        (loss * scale).backward()
        return model, logits.detach(), loss.detach()

    m1, lg1, l1 = run(1.0)
    m2, lg2, l2 = run(1.0)
    m3, _, _ = run(4.0)
    assert rel_inf(lg1, ref["logits"]) < TOL and abs(float(l1) - float(ref["loss"])) < 1e-5
    for (k, p1), (_, p2), (_, p3) in zip(m1.named_parameters(), m2.named_parameters(), m3.named_parameters()):
        assert torch.equal(p1.grad, p2.grad), k                      # deterministic
        assert rel_inf(p3.grad, 4.0 * p1.grad) < 1e-5 or k.endswith("net.0.bias"), k   # linear in dloss
        r = rel_inf(p1.grad, ref["grads"][k])
        if k.endswith("net.0.bias"):
            assert float(p1.grad.abs().max()) < 1e-5
        else:
            assert r < 2e-2, (k, r)     # a handful of sub-1e-6-margin routing flips are expected at this size
    for k, v in m1.state_dict().items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            assert rel_inf(v, sd[k]) < 1e-5, k


def test_shipped_checkpoints_known_answers(demo_inputs, expected_probs, golden):
    x, d = demo_inputs
    xg, dg = x.to(DEV), d.to(DEV)
    cases = [("ecg_baseline_best.pth", P.ECGCNN(12, 256, 5), None, "baseline_prob", "eval/baseline_logits", slice(None)),
             ("af_binary_best.pth", P.ECGCNN(12, 256, 1), None, "af_prob", "eval/af_logits", slice(None)),
             ("ecg_multimodal_best.pth", P.ECGMultimodal(), dg, "mm_prob", "eval/mm_logits", slice(3, None))]
    for ckpt, model, demo, pk, lk, sl in cases:
        sd = load_ckpt(ckpt)
        model.load_state_dict(sd, strict=True)
        model = model.to(DEV).eval()
        with torch.no_grad():
            logits = model(xg[sl]) if demo is None else model(xg[sl], demo)
            prob = Fn.sigmoid(logits)
        ref_logits = torch.from_numpy(golden[lk])
        assert rel_inf(logits, ref_logits) < TOL, (ckpt, rel_inf(logits, ref_logits))
        exp = np.array(expected_probs[pk])
        assert np.abs(prob.cpu().numpy() - exp).max() < 2e-4          # shipped (author's TF32 GPU) rows
        # thresholded multi-label predictions: bit-exact against the oracle
        oracle_prob = torch.sigmoid(ref_logits)
        pred, ref_pred = O.predict(prob.cpu()), O.predict(oracle_prob)
        safe = ref_logits.abs() > 1e-5
        assert torch.equal(pred[safe], ref_pred[safe])
        # eval forward must not touch the running statistics
        for k, v in model.state_dict().items():
            assert torch.equal(v.cpu(), sd[k]), k


def test_return_features_and_eval_determinism():
    model = _model("cnn", 5).eval()
    x = gen(4, 12, 1000, seed=30).to(DEV)
    with torch.no_grad():
        logits, z = model(x, return_features=True)
        again = model(x)
    assert logits.shape == (4, 5) and z.shape == (4, 256)
    assert torch.equal(logits, again)                         # run-to-run bit reproducible


# ------------------------------------------------------------------ Grad-CAM
def test_gradcam1d_dropin_matches_reference_vectors(demo_inputs, golden):
    x, _ = demo_inputs
    model = P.ECGCNN(12, 256, 5)
    model.load_state_dict(load_ckpt("ecg_baseline_best.pth"))
    model = model.to(DEV)
    gc = P.GradCAM1D(model, model.backbone[-1].net[0])
    assert not model.training
    x0 = x[0:1].to(DEV)
    for c in range(5):
        cam = gc.generate_cam(x0, c, signal_length=5000)
        lo = gc.generate_cam(x0, c, signal_length=None)
        ref, ref_lo = golden[f"cam/v1_base_s0_c{c}_T"], golden[f"cam/v1_base_s0_c{c}_lo"]
        assert cam.shape == (5000,) and lo.shape == (625,)
        assert int(cam.argmax()) == int(ref.argmax()) and int(lo.argmax()) == int(ref_lo.argmax())
        assert np.abs(cam.cpu().numpy() - ref).max() < TOL and np.abs(lo.cpu().numpy() - ref_lo).max() < TOL
    assert gc.activations.shape == (1, 256, 625) and gc.gradients.shape == (1, 256, 625)
    sd = load_ckpt("ecg_baseline_best.pth")
    a_ref, g_ref = O._conv4_and_grad(sd, x[0:1], 4, None, sum_batch=False)
    assert rel_inf(gc.activations, a_ref) < TOL and rel_inf(gc.gradients, g_ref) < TOL
    shipped = np.load(__import__("os").path.join(__import__("conftest").GOLDEN, "sample_0_MI_cam.npy"))
    cam0 = gc.generate_cam(x0, 0, signal_length=5000).cpu().numpy()
    assert cam0.argmax() == shipped.argmax() == 620 and np.abs(cam0 - shipped).max() < 6e-4


def test_legacy_and_full_backward_hooks_like_the_scripts(demo_inputs):
    """scripts/00_demo_inference.py:19-61 style: user hooks on the last Conv1d + logits[:,c].sum()."""
    x, _ = demo_inputs
    sd = load_ckpt("ecg_baseline_best.pth")
    model = P.ECGCNN(12, 256, 5); model.load_state_dict(sd); model = model.to(DEV).eval()
    target = [m for m in model.modules() if isinstance(m, torch.nn.Conv1d)][-1]
    store = {}
    h1 = target.register_forward_hook(lambda m, i, o: store.__setitem__("a", o.detach()))
    h2 = target.register_full_backward_hook(lambda m, gi, go: store.__setitem__("g", go[0].detach()))
    model.zero_grad()
    logits = model(x[4:5].to(DEV))
    logits[:, 2].sum().backward()
    a_ref, g_ref = O._conv4_and_grad(sd, x[4:5], 2, None, sum_batch=True)
    assert rel_inf(store["a"], a_ref) < TOL and rel_inf(store["g"], g_ref) < TOL
    h1.remove(); h2.remove()
    # legacy register_backward_hook (src/interpretability/grad_cam_1d.py:36) on a fresh model
    import warnings
    model = P.ECGCNN(12, 256, 5); model.load_state_dict(sd); model = model.to(DEV).eval()
    target = model.backbone[-1].net[0]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        h3 = target.register_backward_hook(lambda m, gi, go: store.__setitem__("g_legacy", go[0].detach()))
        model.zero_grad()
        model(x[4:5].to(DEV))[0, 2].backward()
    assert rel_inf(store["g_legacy"], g_ref) < TOL
    h3.remove()


def test_gradcam_batch_all_variants(demo_inputs):
    x, d = demo_inputs
    sb, sa, sm = (load_ckpt(n) for n in ("ecg_baseline_best.pth", "af_binary_best.pth", "ecg_multimodal_best.pth"))
    base = P.ECGCNN(12, 256, 5); base.load_state_dict(sb); base = base.to(DEV).eval()
    af = P.ECGCNN(12, 256, 1); af.load_state_dict(sa); af = af.to(DEV).eval()
    mm = P.ECGMultimodal(); mm.load_state_dict(sm); mm = mm.to(DEV).eval()
    xg, dg = x.to(DEV), d.to(DEV)
    # V1 (library order), upsampled and low-res
    cam, arg, lo = P.gradcam_batch(base, xg[:6], signal_length=5000, variant="v1", return_lowres=True)
    ref = O.gradcam_batched(sb, x[:6], 5000, variant="v1")
    ref_lo = O.gradcam_batched(sb, x[:6], None, variant="v1")
    assert cam.shape == (6, 5, 5000) and lo.shape == (6, 5, 625) and arg.dtype == torch.int32
    assert float((cam.cpu() - ref).abs().max()) < TOL and float((lo.cpu() - ref_lo).abs().max()) < TOL
    assert torch.equal(arg.cpu().long(), ref.argmax(dim=2))
    assert torch.equal(cam.argmax(dim=2).cpu(), ref.argmax(dim=2))
    # V2 (scripts 00 / 13), eps 1e-9
    cam2, arg2 = P.gradcam_batch(base, xg[:6], signal_length=5000, variant="v2", eps=1e-9)
    ref2 = O.gradcam_batched(sb, x[:6], 5000, variant="v2", eps=1e-9)
    assert float((cam2.cpu() - ref2).abs().max()) < TOL and torch.equal(arg2.cpu().long(), ref2.argmax(dim=2))
    cam_af, arg_af = P.gradcam_batch(af, xg, signal_length=5000, variant="v2", eps=1e-9)
    ref_af = O.gradcam_batched(sa, x, 5000, variant="v2", eps=1e-9)
    assert float((cam_af.cpu() - ref_af).abs().max()) < TOL and torch.equal(arg_af.cpu().long(), ref_af.argmax(dim=2))
    # V3 (script 12, FiLM, per-sample class vectors), eps 1e-8
    cam3, arg3 = P.gradcam_batch(mm, xg[3:], x_demo=dg, signal_length=5000, variant="v2", eps=1e-8)
    ref3 = O.gradcam_batched(sm, x[3:], 5000, demo=d, variant="v2", eps=1e-8)
    assert float((cam3.cpu() - ref3).abs().max()) < TOL and torch.equal(arg3.cpu().long(), ref3.argmax(dim=2))
    # single-sample reference calls agree with the batched rows
    for n, c in ((0, 0), (4, 3)):
        one = O.gradcam_v2(sb, x[n:n + 1], c, 5000)
        assert float((cam2[n, c].cpu() - one).abs().max()) < TOL


def test_gradcam_synthetic_config5_slice():
    """config 5 shape (12x1000, random-init seed 42) on a 64-sample slice vs the oracle."""
    model = _model("cnn", 5).eval()
    sd = O.init_state_dict("cnn", 5, seed=42)
    x, _ = O.synth_batch(64, 1000, 5, seed=0)
    cam, arg, lo = P.gradcam_batch(model, x.to(DEV), signal_length=1000, variant="v1", return_lowres=True)
    ref = O.gradcam_batched(sd, x, 1000, variant="v1")
    ref_lo = O.gradcam_batched(sd, x, None, variant="v1")
    assert cam.shape == (64, 5, 1000) and lo.shape == (64, 5, 125)
    assert float((cam.cpu() - ref).abs().max()) < TOL and float((lo.cpu() - ref_lo).abs().max()) < TOL
    # argmax bit-exact wherever the oracle's top-2 gap exceeds fp32 noise
    top2 = ref.topk(2, dim=2).values
    clear = (top2[..., 0] - top2[..., 1]) > 1e-5
    assert torch.equal(arg.cpu().long()[clear], ref.argmax(dim=2)[clear])
    dead = ref.max(dim=2).values == 0                     # ReLU killed the whole map: argmax is index 0
    assert torch.equal(arg.cpu().long()[dead], torch.zeros_like(arg.cpu().long()[dead]))
    # remaining maps have a (near-)tie at the top: our value at the oracle's peak must equal our max
    ours_at_ref = cam.cpu().gather(2, ref.argmax(dim=2, keepdim=True)).squeeze(2)
    assert float((cam.cpu().max(dim=2).values - ours_at_ref).abs().max()) < 1e-5
    assert (clear | dead).float().mean() > 0.6


def test_demo_importance(demo_inputs, golden):
    x, d = demo_inputs
    mm = P.ECGMultimodal(); mm.load_state_dict(load_ckpt("ecg_multimodal_best.pth")); mm = mm.to(DEV).eval()
    for j in (0, 5):
        for c in (0, 3):
            imp = P.compute_demo_importance(mm, x[3 + j:4 + j].to(DEV), d[j:j + 1].to(DEV), c)
            np.testing.assert_allclose(imp, golden[f"imp/mm_j{j}_c{c}"], atol=TOL)


# ------------------------------------------------------------------ loops
def test_train_and_eval_loops_match_reference_loop_semantics():
    from torch.utils.data import DataLoader, TensorDataset
    x, y = O.synth_batch(24, 250, 5, seed=3)
    loader = DataLoader(TensorDataset(x, y), batch_size=8, shuffle=False)
    model = _model("cnn", 5)
    opt = P.FusedAdamW(model.parameters(), lr=1.5e-3, weight_decay=1e-4)
    got = P.train_one_epoch(model, loader, opt, DEV)
    sd = O.init_state_dict("cnn", 5, seed=42)
    st = O.AdamWState(sd, 1.5e-3, 1e-4)
    tot = 0.0
    for xb, yb in loader:                                     # src/training/loop.py:22-38
        tot += float(O.train_step(sd, xb, yb, st)["loss"]) * xb.size(0)
    assert abs(got - tot / 24) < 1e-4 * abs(tot / 24)
    m = P.eval_one_epoch(model, loader, DEV)
    with torch.no_grad():
        lo = O.ecgcnn_forward(dict(sd), x, train=False)
    ref_loss = sum(float(O.bce_with_logits(lo[i:i + 8], y[i:i + 8])) * 8 for i in range(0, 24, 8)) / 24
    assert set(m) == {"auroc_macro", "auprc_macro", "f1_macro", "bce_loss"}
    assert abs(m["bce_loss"] - ref_loss) < 2e-3 * abs(ref_loss)

    x, demo, y = O.synth_batch(16, 250, 5, seed=4, with_demo=True)
    loader = DataLoader(TensorDataset(x, demo, y), batch_size=8, shuffle=False)
    model = _model("mm", 5)
    opt = P.FusedAdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)
    got = P.train_one_epoch_demo(model, loader, opt, DEV)
    sd = O.init_state_dict("mm", 5, seed=42); st = O.AdamWState(sd, 1e-4, 1e-4)
    losses = [float(O.train_step(sd, a, c, st, demo=b_)["loss"]) for a, b_, c in loader]
    assert abs(got - sum(losses) / 2) < 1e-4 * abs(sum(losses) / 2)      # mean of batch means, loop_demo.py:38-43
    assert "bce_loss" in P.eval_one_epoch_demo(model, loader, DEV)


def test_eval_counts_match_sklearn():
    """Device-side evaluation epilogue (N3): sigmoid + threshold + confusion counts == numpy/sklearn on the same
    logits, including exact-threshold ties (prob >= 0.5 at logit 0) and an all-negative label."""
    from sklearn.metrics import f1_score
    from ptbxl_multimodal_b200.metrics import f1_macro_from_counts
    g = torch.Generator().manual_seed(9)
    logits = torch.randn(1000, 5, generator=g) * 2
    logits[::7, 2] = 0.0                                   # sigmoid(0) = 0.5 exactly -> predicted positive
    y = (torch.rand(1000, 5, generator=g) < torch.tensor([0.25, 0.24, 0.12, 0.23, 0.0])).float()
    counts = torch.zeros(5, 4, dtype=torch.int32, device=DEV)
    prob, pred = Fn.eval_counts(logits[:600].to(DEV), y[:600].to(DEV), counts)
    prob2, pred2 = Fn.eval_counts(logits[600:].to(DEV), y[600:].to(DEV), counts)      # accumulates across batches
    p = torch.cat([prob, prob2]).cpu()
    ref_p = torch.sigmoid(logits)
    assert float((p - ref_p).abs().max()) < 2e-7
    y_pred = (p.numpy() >= 0.5).astype(int)
    assert (torch.cat([pred, pred2]).cpu().numpy() == y_pred).all()
    yt = y.numpy().astype(int)
    c = counts.cpu().numpy()
    assert (c[:, 0] == (y_pred & yt).sum(0)).all() and (c[:, 1] == (y_pred & (1 - yt)).sum(0)).all()
    assert (c[:, 2] == ((1 - y_pred) & yt).sum(0)).all() and (c.sum(1) == 1000).all()
    assert abs(f1_macro_from_counts(c) - f1_score(yt, y_pred, average="macro", zero_division=0)) < 1e-12


def test_legacy_concat_fusion_model_equals_its_reconstruction():
    """SURVEY 8f N4 (parity unpinned: the reference ships no source for this model): the ecgb200 module equals
    the oracle's torch restatement of the same reconstruction, forward and w.r.t. the demographic input."""
    torch.manual_seed(11)
    model = P.ECGDemoConcat().to(DEV).eval()
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    x = gen(5, 12, 1000, seed=61)
    d = gen(5, 5, seed=62).abs()
    ref = O.concat_forward(sd, x, d)
    dg = d.to(DEV).requires_grad_(True)
    logits = model(x.to(DEV), dg)
    assert rel_inf(logits, ref) < TOL, rel_inf(logits, ref)
    logits[:, 2].sum().backward()
    dr = d.clone().requires_grad_(True)
    O.concat_forward(sd, x, dr)[:, 2].sum().backward()
    assert rel_inf(dg.grad, dr.grad) < TOL
    assert model.classifier[0].weight.grad is not None and model.ecg_encoder.proj.weight.grad is not None

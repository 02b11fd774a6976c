"""numpy restatement of the eval-mode forward of the ECG models (TEST INFRASTRUCTURE ONLY -- see ecg_oracle.py's header).

ecg_oracle.py calls the same ATen ops the reference calls; this file restates those ops themselves in plain numpy
(float64 accumulation), so that the definition of the path does not rest on PyTorch alone:
  Conv1d(k=15, padding=7, stride=1)          y[b,o,t] = bias[o] + sum_{c,k} w[o,c,k] * x[b,c,t+k-7]   (zero padded)   ecg_cnn.py:13
  BatchNorm1d (eval)                         (y - running_mean) / sqrt(running_var + 1e-5) * weight + bias            ecg_cnn.py:14
  ReLU, MaxPool1d(2) (floor mode)            max over (2j, 2j+1), j < L // 2                                          ecg_cnn.py:15-16
  AdaptiveAvgPool1d(1) + squeeze             mean over time                                                           ecg_cnn.py:61-62
  proj, head (Linear)                        x @ W.T + b                                                              ecg_cnn.py:63-64
  DemoEncoder, film_gen, FiLM                relu(Linear) x2; gamma, beta = chunk(film, 2); (1 + tanh(gamma)) * z + beta   ecg_multimodal.py:44-59,88-99
  Grad-CAM (V1 / V2 / V3 orderings)          closed-form gradient through eval BN / ReLU / MaxPool (first-index ties) / GAP,
                                             no autograd; F.interpolate(linear, align_corners=False) restated      grad_cam_1d.py:75-101, scripts/00, 12, 13
  one training step (ECGCNN, ECGMultimodal)  train-mode BN, BCE, the backward pass (incl. FiLM / demo encoder), AdamW    loop.py:22-36, loop_demo.py:25-41
Pinned in tests/test_oracle_golden.py against the golden logits / gradients / CAM curves produced by the unmodified
reference, the shipped prediction CSV rows and the shipped CAM file (argmax 620)."""
import numpy as np

EPS = 1e-5


def conv1d_k15(x, w, b):
    bsz, ci, L = x.shape
    co, _, k = w.shape
    pad = k // 2
    xp = np.zeros((bsz, ci, L + 2 * pad), dtype=np.float64)
    xp[:, :, pad:pad + L] = x
    y = np.zeros((bsz, co, L), dtype=np.float64)
    w64 = w.astype(np.float64)
    for kk in range(k):                                   # y += W_k @ x shifted by tap kk
        y += np.einsum("oc,bct->bot", w64[:, :, kk], xp[:, :, kk:kk + L], optimize=True)
    return y + b.astype(np.float64)[None, :, None]


def conv_block(sd, prefix, x):
    a = conv1d_k15(x, sd[prefix + "net.0.weight"], sd[prefix + "net.0.bias"])
    g, be = sd[prefix + "net.1.weight"].astype(np.float64), sd[prefix + "net.1.bias"].astype(np.float64)
    m, v = sd[prefix + "net.1.running_mean"].astype(np.float64), sd[prefix + "net.1.running_var"].astype(np.float64)
    h = (a - m[None, :, None]) / np.sqrt(v[None, :, None] + EPS) * g[None, :, None] + be[None, :, None]
    h = np.maximum(h, 0.0)
    lp = h.shape[2] // 2
    return np.maximum(h[:, :, 0:2 * lp:2], h[:, :, 1:2 * lp:2])


def linear(x, w, b):
    return x @ w.astype(np.float64).T + b.astype(np.float64)


def backbone(sd, prefix, x):
    h = x.astype(np.float64)
    for i in range(4):
        h = conv_block(sd, f"{prefix}backbone.{i}.", h)
    return linear(h.mean(axis=2), sd[prefix + "proj.weight"], sd[prefix + "proj.bias"])


def ecgcnn_logits(sd, x):
    return linear(backbone(sd, "", x), sd["head.weight"], sd["head.bias"])


def multimodal_logits(sd, x, d):
    z = backbone(sd, "ecg_backbone.", x)
    h = np.maximum(linear(d.astype(np.float64), sd["demo_encoder.mlp.0.weight"], sd["demo_encoder.mlp.0.bias"]), 0.0)
    h = np.maximum(linear(h, sd["demo_encoder.mlp.2.weight"], sd["demo_encoder.mlp.2.bias"]), 0.0)
    film = linear(h, sd["film_gen.weight"], sd["film_gen.bias"])
    f = z.shape[1]
    return linear((1.0 + np.tanh(film[:, :f])) * z + film[:, f:], sd["head.weight"], sd["head.bias"])


def sigmoid(x):
    return 1.0 / (1.0 + np.exp(-x))


# ------------------------------------------------------------------ Grad-CAM without autograd (closed form)
# In eval mode the gradient of logit c w.r.t. the raw 4th-conv output A (256 x L') is
#     G_c[ch, t] = v_c[ch] * s[ch] * mask[ch, t] / L_p,      s = gamma_bn / sqrt(running_var + eps),
#     mask[ch, t] = [t < 2 L_p] and [BN(A)[ch, t] > 0] and [t is the FIRST argmax of its pool pair]   (MaxPool1d ties -> first),
#     v_c = W_head[c] @ W_proj   (FiLM: (W_head[c] * (1 + tanh gamma_film(d))) @ W_proj),
# so the channel weights of grad_cam_1d.py:85 are mean_t G_c and the CAM needs no backward pass.
def conv4_raw_and_mask(sd, prefix, x):
    """Raw 4th-conv output A (B,256,L') and the routing mask of the BN -> ReLU -> MaxPool(2) -> mean that follows it."""
    h = x.astype(np.float64)
    for i in range(3):
        h = conv_block(sd, f"{prefix}backbone.{i}.", h)
    p4 = f"{prefix}backbone.3."
    a = conv1d_k15(h, sd[p4 + "net.0.weight"], sd[p4 + "net.0.bias"])
    g, be = sd[p4 + "net.1.weight"].astype(np.float64), sd[p4 + "net.1.bias"].astype(np.float64)
    m, v = sd[p4 + "net.1.running_mean"].astype(np.float64), sd[p4 + "net.1.running_var"].astype(np.float64)
    s = g / np.sqrt(v + EPS)
    r = (a - m[None, :, None]) * s[None, :, None] + be[None, :, None]
    lp = a.shape[2] // 2
    even, odd = r[:, :, 0:2 * lp:2], r[:, :, 1:2 * lp:2]
    mask = np.zeros_like(a)
    mask[:, :, 0:2 * lp:2] = (even >= odd) & (even > 0)            # ties -> first index; ReLU backward mask is out > 0
    mask[:, :, 1:2 * lp:2] = (odd > even) & (odd > 0)
    return a, mask, s, lp


def linear_upsample(cam, out_len):
    """F.interpolate(mode='linear', align_corners=False) along the last axis (grad_cam_1d.py:96-101)."""
    L = cam.shape[-1]
    src = np.maximum((np.arange(out_len) + 0.5) * (L / out_len) - 0.5, 0.0)
    i0 = np.floor(src).astype(np.int64)
    i1 = np.minimum(i0 + 1, L - 1)
    lam = src - i0
    return (1.0 - lam) * cam[..., i0] + lam * cam[..., i1]


def gradcam(sd, x, class_idx, signal_length=None, variant="v1", eps=1e-9, demo=None):
    """One window x (1,12,T) -> CAM.  variant "v1": GradCAM1D.generate_cam (grad_cam_1d.py:75-101: normalise at L', /max only
    if max > 0, THEN upsample); "v2": the script classes (scripts/00:39-61, 13:51-76; 12:44-75 with eps 1e-8 and `demo`):
    upsample THEN (cam - min) / (max + eps)."""
    prefix = "" if demo is None else "ecg_backbone."
    a, mask, s, lp = conv4_raw_and_mask(sd, prefix, x)
    wh = sd["head.weight"].astype(np.float64)[class_idx]
    if demo is not None:
        h = np.maximum(linear(demo.astype(np.float64), sd["demo_encoder.mlp.0.weight"], sd["demo_encoder.mlp.0.bias"]), 0.0)
        h = np.maximum(linear(h, sd["demo_encoder.mlp.2.weight"], sd["demo_encoder.mlp.2.bias"]), 0.0)
        film = linear(h, sd["film_gen.weight"], sd["film_gen.bias"])[0]
        wh = wh * (1.0 + np.tanh(film[:wh.shape[0]]))
    v = wh @ sd[prefix + "proj.weight"].astype(np.float64)                     # (256,)
    w = v * s * mask[0].sum(axis=1) / (lp * a.shape[2])                       # mean over t of G_c
    cam = np.maximum((w[:, None] * a[0]).sum(axis=0), 0.0)                    # relu(sum_ch w * A)
    if variant == "v1":
        cam = cam - cam.min()
        if cam.max() > 0:
            cam = cam / cam.max()
        if signal_length is not None and cam.shape[-1] != signal_length:
            cam = linear_upsample(cam, signal_length)
        return cam
    cam = linear_upsample(cam, signal_length)
    cam = cam - cam.min()
    return cam / (cam.max() + eps)


# ------------------------------------------------------------------ one training step of ECGCNN in numpy (float64)
# forward in train mode (batch statistics, biased variance; BatchNorm1d at ecg_cnn.py:14), mean BCE-with-logits
# (loop.py:32), the full backward pass and one AdamW update (loop.py:33-34; torch.optim.AdamW defaults betas (0.9, 0.999),
# eps 1e-8, decoupled weight decay).  Pinned to the reference's golden step-0 loss / logits / gradients / updated weights.
def conv1d_k15_backward(x, w, dy):
    bsz, ci, L = x.shape
    co, _, k = w.shape
    pad = k // 2
    xp = np.zeros((bsz, ci, L + 2 * pad)); xp[:, :, pad:pad + L] = x
    dxp = np.zeros_like(xp)
    dw = np.zeros(w.shape)
    w64 = w.astype(np.float64)
    for kk in range(k):
        dw[:, :, kk] = np.einsum("bot,bct->oc", dy, xp[:, :, kk:kk + L], optimize=True)
        dxp[:, :, kk:kk + L] += np.einsum("oc,bot->bct", w64[:, :, kk], dy, optimize=True)
    return dxp[:, :, pad:pad + L], dw, dy.sum(axis=(0, 2))


def train_step(sd, x, y, lr, wd, step=1, m=None, v=None, demo=None):
    """Returns (loss, logits, grads dict, updated parameter dict) for one step from state `sd` (numpy arrays).
    demo=None: ECGCNN (loop.py:22-36).  demo (B,5): ECGMultimodal with FiLM conditioning (loop_demo.py:25-41;
    ecg_multimodal.py:88-99) -- the demo encoder, film_gen and the FiLM product in the forward and the backward pass."""
    x = x.astype(np.float64); y = y.astype(np.float64)
    pre = "" if demo is None else "ecg_backbone."
    cache = []
    h = x
    for i in range(4):
        p = f"{pre}backbone.{i}."
        w, b = sd[p + "net.0.weight"], sd[p + "net.0.bias"]
        a = conv1d_k15(h, w, b)
        mean = a.mean(axis=(0, 2)); var = a.var(axis=(0, 2))                       # biased, as batch_norm normalises
        rstd = 1.0 / np.sqrt(var + EPS)
        xhat = (a - mean[None, :, None]) * rstd[None, :, None]
        g = sd[p + "net.1.weight"].astype(np.float64); be = sd[p + "net.1.bias"].astype(np.float64)
        r = np.maximum(xhat * g[None, :, None] + be[None, :, None], 0.0)
        lp = r.shape[2] // 2
        r0, r1 = r[:, :, 0:2 * lp:2], r[:, :, 1:2 * lp:2]
        pooled = np.maximum(r0, r1)
        cache.append((h, w, xhat, rstd, g, r, r0 >= r1, lp))                       # first index wins ties
        h = pooled
    gap = h.mean(axis=2)
    z = linear(gap, sd[pre + "proj.weight"], sd[pre + "proj.bias"])
    if demo is not None:
        d = demo.astype(np.float64)
        h1 = np.maximum(linear(d, sd["demo_encoder.mlp.0.weight"], sd["demo_encoder.mlp.0.bias"]), 0.0)
        h2 = np.maximum(linear(h1, sd["demo_encoder.mlp.2.weight"], sd["demo_encoder.mlp.2.bias"]), 0.0)
        film = linear(h2, sd["film_gen.weight"], sd["film_gen.bias"])
        f = z.shape[1]
        tg = np.tanh(film[:, :f])                                                  # gamma = first half of the chunk
        feat = (1.0 + tg) * z + film[:, f:]
    else:
        feat = z
    logits = linear(feat, sd["head.weight"], sd["head.bias"])
    loss = np.mean(np.maximum(logits, 0) - logits * y + np.log1p(np.exp(-np.abs(logits))))
    grads = {}
    dlog = (sigmoid(logits) - y) / logits.size
    grads["head.weight"] = dlog.T @ feat; grads["head.bias"] = dlog.sum(0)
    dfeat = dlog @ sd["head.weight"].astype(np.float64)
    if demo is not None:
        dz = dfeat * (1.0 + tg)
        dfilm = np.concatenate([dfeat * z * (1.0 - tg * tg), dfeat], axis=1)
        grads["film_gen.weight"] = dfilm.T @ h2; grads["film_gen.bias"] = dfilm.sum(0)
        dh2 = (dfilm @ sd["film_gen.weight"].astype(np.float64)) * (h2 > 0)
        grads["demo_encoder.mlp.2.weight"] = dh2.T @ h1; grads["demo_encoder.mlp.2.bias"] = dh2.sum(0)
        dh1 = (dh2 @ sd["demo_encoder.mlp.2.weight"].astype(np.float64)) * (h1 > 0)
        grads["demo_encoder.mlp.0.weight"] = dh1.T @ d; grads["demo_encoder.mlp.0.bias"] = dh1.sum(0)
    else:
        dz = dfeat
    grads[pre + "proj.weight"] = dz.T @ gap; grads[pre + "proj.bias"] = dz.sum(0)
    dgap = dz @ sd[pre + "proj.weight"].astype(np.float64)
    dp = np.repeat(dgap[:, :, None], h.shape[2], axis=2) / h.shape[2]
    for i in (3, 2, 1, 0):
        p = f"{pre}backbone.{i}."
        hin, w, xhat, rstd, g, r, first, lp = cache[i]
        dr = np.zeros_like(r)
        dr[:, :, 0:2 * lp:2] = np.where(first, dp, 0.0)
        dr[:, :, 1:2 * lp:2] = np.where(first, 0.0, dp)
        dr = dr * (r > 0)
        grads[p + "net.1.weight"] = (dr * xhat).sum(axis=(0, 2)); grads[p + "net.1.bias"] = dr.sum(axis=(0, 2))
        n = dr.shape[0] * dr.shape[2]
        dxhat = dr * g[None, :, None]
        da = (dxhat - dxhat.sum(axis=(0, 2))[None, :, None] / n
              - xhat * (dxhat * xhat).sum(axis=(0, 2))[None, :, None] / n) * rstd[None, :, None]
        dp, grads[p + "net.0.weight"], grads[p + "net.0.bias"] = conv1d_k15_backward(hin, w, da)
    new = {}
    b1, b2, eps = 0.9, 0.999, 1e-8
    for k, gk in grads.items():
        pk = sd[k].astype(np.float64) * (1.0 - lr * wd)
        mk = (1 - b1) * gk if m is None else b1 * m[k] + (1 - b1) * gk
        vk = (1 - b2) * gk * gk if v is None else b2 * v[k] + (1 - b2) * gk * gk
        new[k] = pk - lr / (1 - b1 ** step) * mk / (np.sqrt(vk) / np.sqrt(1 - b2 ** step) + eps)
    return loss, logits, grads, new


def train_step_cnn(sd, x, y, lr, wd, step=1, m=None, v=None):
    return train_step(sd, x, y, lr, wd, step=step, m=m, v=v)

"""numpy restatement of the eval-mode forward of the ECG models (TEST INFRASTRUCTURE ONLY -- see ecg_oracle.py's header).

ecg_oracle.py calls the same ATen ops the reference calls; this file restates those ops themselves in plain numpy
(float64 accumulation), so that the definition of the path does not rest on PyTorch alone:
  Conv1d(k=15, padding=7, stride=1)          y[b,o,t] = bias[o] + sum_{c,k} w[o,c,k] * x[b,c,t+k-7]   (zero padded)   ecg_cnn.py:13
  BatchNorm1d (eval)                         (y - running_mean) / sqrt(running_var + 1e-5) * weight + bias            ecg_cnn.py:14
  ReLU, MaxPool1d(2) (floor mode)            max over (2j, 2j+1), j < L // 2                                          ecg_cnn.py:15-16
  AdaptiveAvgPool1d(1) + squeeze             mean over time                                                           ecg_cnn.py:61-62
  proj, head (Linear)                        x @ W.T + b                                                              ecg_cnn.py:63-64
  DemoEncoder, film_gen, FiLM                relu(Linear) x2; gamma, beta = chunk(film, 2); (1 + tanh(gamma)) * z + beta   ecg_multimodal.py:44-59,88-99
Pinned in tests/test_oracle_golden.py against the golden logits produced by the unmodified reference and against the
shipped prediction CSV rows."""
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

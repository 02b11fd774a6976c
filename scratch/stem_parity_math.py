"""Index math of the planned stem-layer kernel, checked in numpy: fold the two time steps of a pool pair into N.

out[2i]   = sum_k X[2i   + k - 7] W_k        out[2i+1] = sum_k X[2i+1 + k - 7] W_k        (k = 0..14, zero padding)
With the input de-interleaved by time parity, Xe[j] = X[2j], Xo[j] = X[2j+1]:
  Xo[i-4+m], m = 0..7 : contributes W_{2m}   to out[2i]   and W_{2m-1} to out[2i+1]  (W_{-1} = 0)
  Xe[i-3+m], m = 0..7 : contributes W_{2m+1} to out[2i]   and W_{2m}   to out[2i+1]  (W_{15} = 0)
i.e. 16 MMAs per K-step with B = [W_a | W_b] (N = 2*Co) instead of 30 with N = Co, and the pool pair ends up in one accumulator row."""
import numpy as np
rng = np.random.default_rng(0)
Ci, Co, L = 16, 32, 64                        # L even
X = rng.standard_normal((L, Ci)); W = rng.standard_normal((15, Ci, Co))
Xp = np.zeros((L + 14, Ci)); Xp[7:7 + L] = X
ref = np.stack([sum(Xp[t + k] @ W[k] for k in range(15)) for t in range(L)])          # out[t]


def plane(par, j):                            # Xe[j] / Xo[j] with zero fill outside [0, L/2)
    return X[2 * j + par] if 0 <= j < L // 2 else np.zeros(Ci)


Wz = lambda k: W[k] if 0 <= k < 15 else np.zeros((Ci, Co))                             # noqa: E731
out = np.zeros((L // 2, 2 * Co))
for i in range(L // 2):
    acc = np.zeros(2 * Co)
    for m in range(8):
        acc += plane(1, i - 4 + m) @ np.concatenate([Wz(2 * m), Wz(2 * m - 1)], axis=1)
        acc += plane(0, i - 3 + m) @ np.concatenate([Wz(2 * m + 1), Wz(2 * m)], axis=1)
    out[i] = acc
assert np.allclose(out[:, :Co], ref[0::2]) and np.allclose(out[:, Co:], ref[1::2])
print("parity decomposition verified: max err", np.abs(out[:, :Co] - ref[0::2]).max(), np.abs(out[:, Co:] - ref[1::2]).max())

"""Shared helpers of the parity tests (synthetic inputs of SURVEY.md section 8d, error metrics)."""
import numpy as np

from oracle import head_oracle as ho

TOL_F32 = 1e-5   # BASELINE.json: loss and gradients within 1e-5 relative in fp32 mode
TOL_BF16 = 2e-2  # ... and within 2e-2 relative in bf16-GEMM mode


def rel_err(a, b):
    """max |a - b| / max |b|  (relative to the largest reference magnitude)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    if b.size == 0:
        return 0.0
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def lt_counts(C, n_max=1280, ratio=100.0):
    """n_c = floor(n_max * ratio^(-c/(C-1)))  (cls/imbalanced_dataset.py:27-29)."""
    if C == 1:
        return np.array([n_max], np.int64)
    return np.array([max(int(n_max * (1.0 / ratio) ** (c / (C - 1.0))), 1) for c in range(C)], np.int64)


def lt_labels(counts, n, rng):
    p = counts / counts.sum()
    return rng.choice(len(counts), size=n, p=p).astype(np.int64)


def head_inputs(B, D, C, seed=0, relu=False, bias=0.01):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((B, D)).astype(np.float32)
    if relu:
        x = np.maximum(x, 0)
    k = 1.0 / np.sqrt(D)
    w = rng.uniform(-k, k, size=(C, D)).astype(np.float32)
    b = np.full(C, bias, np.float32)
    counts = lt_counts(C)
    y = lt_labels(counts, B, rng)
    return x, w, b, counts, y


def iif_row(counts, variant="raw"):
    return ho.to_f32_row(ho.iif_weights_from_counts(counts)[variant])


def bf16_round(a):
    """fp32 -> bf16 (round to nearest even) -> fp32, in numpy."""
    a = np.ascontiguousarray(a, np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    r = ((u + 0x7FFF + ((u >> 16) & 1)) >> 16) << 16
    return r.astype(np.uint32).view(np.float32).reshape(a.shape)

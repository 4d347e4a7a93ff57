"""GPU parity: label histograms (bit-exact) and the IIF weight vectors (one fp64->fp32 rounding)
against the oracle, the CIFAR-LT closed-form profile and the reference's CSV weight tables."""
import numpy as np
import pytest
import torch

from oracle import head_oracle as ho
from _common import lt_counts, lt_labels

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CSV_TO_VARIANT = {"smooth": "smooth", "raw": "raw", "prob": "rel", "normit": "normit",
                  "gombit": "gombit", "base2": "base2", "base10": "base10"}


@pytest.fixture(scope="module")
def hist():
    from iif_b200 import histogram as h
    return h


def ulp_diff_f32(a, b):
    a = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    b = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b).max()


@pytest.mark.parametrize("n,C", [(0, 10), (1, 10), (12406, 10), (115846, 1000), (62500, 365), (1000003, 1204),
                                 (50000, 12288), (50000, 20000)])
def test_label_hist_bitexact(hist, n, C):
    rng = np.random.default_rng(n + C)
    y = lt_labels(lt_counts(C), n, rng) if n else np.zeros(0, np.int64)
    if n > 10:
        y[:5] = [-1, C, C + 7, -100, 2 ** 40]        # out-of-range labels are counted by nobody
    got = hist.class_counts(torch.from_numpy(y).to(DEV), C).cpu().numpy()
    assert got.dtype == np.int64
    assert np.array_equal(got, ho.label_hist(y, C))


def test_label_hist_unaligned_and_cifar_profile(hist):
    """CIFAR-10-LT r=100 profile (cls/imbalanced_dataset.py:23-29) recovered exactly from its labels;
    an odd-offset view exercises the non-128-bit path."""
    prof = ho.cifar_lt_profile(5000, 10, 0.01)
    y = np.repeat(np.arange(10), prof).astype(np.int64)
    np.random.default_rng(0).shuffle(y)
    t = torch.from_numpy(np.concatenate([[3], y])).to(DEV)
    assert hist.class_counts(t[1:], 10).cpu().tolist() == prof
    assert hist.class_counts(t[1:-1], 10).cpu().numpy().tolist() == ho.label_hist(y[:-1], 10).tolist()
    cmap = hist.lt_class_map(hist.class_counts(t[1:], 10))
    assert np.array_equal(cmap, ho.lt_class_map(np.array(prof)))


def test_image_dedup_hist(hist):
    rng = np.random.default_rng(7)
    n_img, C, n = 5000, 1203, 200000
    img = rng.integers(0, n_img, n).astype(np.int64)
    cat = lt_labels(lt_counts(C), n, rng)
    a, b = hist.image_instance_freq(torch.from_numpy(img).to(DEV), torch.from_numpy(cat).to(DEV), n_img, C)
    ei, en = ho.image_dedup_hist(img, cat, C)
    assert np.array_equal(a.cpu().numpy(), ei) and np.array_equal(b.cpu().numpy(), en)


@pytest.mark.parametrize("v", ho.VARIANTS)
def test_cls_weights_vs_golden(hist, golden, v):
    """Bit-exact against classification/custom.py:14-26 run unmodified (tests/golden/cls_iif.npz)."""
    g = golden("cls_iif")
    counts = torch.from_numpy(g["counts"]).to(DEV)
    w = hist.iif_weights(counts, v).cpu().numpy()
    assert w.shape == (1, 10)
    assert ulp_diff_f32(w, g[f"iif_{v}"]) <= 1, v      # fp64 libm differences can flip the final fp32 rounding
    wn = hist.iif_weights(counts, v, iif_norm=2.0).cpu().numpy()
    np.testing.assert_allclose(wn, g[f"iifn2_{v}"], rtol=1e-6)


@pytest.mark.parametrize("table", ["idf_1204", "idf_1231", "idf_91"])
def test_csv_tables(hist, golden, table):
    """(img_freq, instance_freq) -> the 14 weight columns of the reference CSVs, fp64 and fp32."""
    from iif_b200 import ops
    g = golden("weight_tables")
    img = torch.from_numpy(g[f"{table}_img_freq"]).to(DEV)
    inst = torch.from_numpy(g[f"{table}_instance_freq"]).to(DEV)
    n_img = int(g[f"{table}_n_img"])
    for col, var in CSV_TO_VARIANT.items():
        for suf, cnt, total in (("", img, n_img), ("_obj", inst, 0)):
            ref = g[f"{table}_{col}{suf}"][1:]
            w32, w64 = ops.weights_from_counts(cnt, var, total=total, return_f64=True)
            np.testing.assert_allclose(w64.cpu().numpy(), ref, rtol=0, atol=2e-12 * max(1, np.abs(ref).max()))
            assert ulp_diff_f32(w32.cpu().numpy(), ref.astype(np.float32)) <= 1
    tab = hist.detection_weight_table(img, inst, n_img)
    assert sorted(tab) == sorted([c + s for c in CSV_TO_VARIANT for s in ("", "_obj")])


def test_weights_edge_cases(hist):
    """zero-count class -> inf (cls/custom.py:16); f > N/2 -> negative rel / normit."""
    counts = torch.tensor([900, 0, 100], dtype=torch.int64, device=DEV)
    ref = ho.iif_weights_from_counts(np.array([900, 0, 100]))
    for v in ho.VARIANTS:
        got = hist.iif_weights(counts, v).cpu().numpy()[0].astype(np.float64)
        exp = ref[v].astype(np.float32).astype(np.float64)
        assert np.array_equal(np.isinf(got), np.isinf(exp)) and np.array_equal(np.isnan(got), np.isnan(exp)), v
        ok = np.isfinite(exp)
        np.testing.assert_allclose(got[ok], exp[ok], rtol=2e-7)
    assert hist.iif_weights(counts, "rel").cpu().numpy()[0, 0] < 0

"""Import the reference's own Python, unmodified, from /root/reference (TEST INFRASTRUCTURE ONLY).

Only usable in the build container (the reference tree does not travel to the GPU box); used by
``tests/golden/make_golden.py`` to freeze golden vectors and by the optional
``tests/test_oracle_vs_reference.py`` (skipped when the tree is absent).

* classification: ``custom.py`` imports directly (deps: torch / numpy / scipy).
* mmdet: ``losses/{utils,accuracy,cross_entropy_loss,iif_loss,fasa_iif_loss}.py`` are loaded by
  path behind a stub ``mmcv`` (``jit`` -> identity decorator) and a stub registry
  (``LOSSES.register_module()`` -> identity); no reference file is edited or copied.
* The reference hard-codes ``device='cuda'`` (iif_loss.py:50) and ``torch.cuda.FloatTensor``
  (custom.py:61); ``cpu_shims()`` remaps those two constructors to the CPU for the duration of a
  ``with`` block so the unmodified code runs on a GPU-less host.
"""
from __future__ import annotations

import contextlib
import importlib.util
import os
import sys
import types

REF = os.environ.get("IIF_REFERENCE_ROOT", "/root/reference")
_LOSS_DIR = os.path.join(REF, "instance_segmentation", "mmdet", "models", "losses")


def available() -> bool:
    return os.path.isfile(os.path.join(REF, "classification", "custom.py"))


def load_classification():
    p = os.path.join(REF, "classification")
    if p not in sys.path:
        sys.path.insert(0, p)
    import custom  # noqa: the reference module
    return custom


def load_mmdet_losses():
    """Returns a namespace with utils, accuracy, cross_entropy_loss, iif_loss, fasa_iif_loss."""
    if "mmdet.models.losses.iif_loss" in sys.modules:
        return types.SimpleNamespace(**{m: sys.modules["mmdet.models.losses." + m] for m in
                                        ("utils", "accuracy", "cross_entropy_loss", "iif_loss", "fasa_iif_loss")})
    mmcv = types.ModuleType("mmcv")
    mmcv.jit = lambda *a, **k: (lambda f: f)
    sys.modules.setdefault("mmcv", mmcv)
    for n in ("mmdet", "mmdet.models", "mmdet.models.losses"):
        if n not in sys.modules:
            m = types.ModuleType(n)
            m.__path__ = []
            sys.modules[n] = m
    b = types.ModuleType("mmdet.models.builder")

    class _Reg:
        def register_module(self, *a, **k):
            return lambda c: c

    b.LOSSES = _Reg()
    sys.modules["mmdet.models.builder"] = b
    out = {}
    for mod in ("utils", "accuracy", "cross_entropy_loss", "iif_loss", "fasa_iif_loss"):
        if mod == "fasa_iif_loss":
            pkg = sys.modules["mmdet.models.losses"]
            pkg.binary_cross_entropy = out["cross_entropy_loss"].binary_cross_entropy
            pkg.mask_cross_entropy = out["cross_entropy_loss"].mask_cross_entropy
        spec = importlib.util.spec_from_file_location("mmdet.models.losses." + mod,
                                                      os.path.join(_LOSS_DIR, mod + ".py"))
        m = importlib.util.module_from_spec(spec)
        sys.modules[spec.name] = m
        spec.loader.exec_module(m)
        out[mod] = m
    return types.SimpleNamespace(**out)


def load_normed_predictors():
    """mmdet/models/utils/normed_predictor.py behind stub registries (mmcv.cnn.CONV_LAYERS, .builder.LINEAR_LAYERS)."""
    name = "mmdet.models.utils.normed_predictor"
    if name in sys.modules:
        return sys.modules[name]
    load_mmdet_losses()

    class _Reg:
        def register_module(self, *a, **k):
            return lambda c: c

    cnn = types.ModuleType("mmcv.cnn")
    cnn.CONV_LAYERS = _Reg()
    sys.modules["mmcv.cnn"] = cnn
    sys.modules["mmcv"].cnn = cnn
    pkg = types.ModuleType("mmdet.models.utils")
    pkg.__path__ = []
    sys.modules.setdefault("mmdet.models.utils", pkg)
    b = types.ModuleType("mmdet.models.utils.builder")
    b.LINEAR_LAYERS = _Reg()
    sys.modules["mmdet.models.utils.builder"] = b
    spec = importlib.util.spec_from_file_location(
        name, os.path.join(REF, "instance_segmentation", "mmdet", "models", "utils", "normed_predictor.py"))
    m = importlib.util.module_from_spec(spec)
    sys.modules[name] = m
    spec.loader.exec_module(m)
    return m


def load_resnet_cifar():
    """classification/resnet_cifar.py (CosNorm_Classifier); its constructor calls .cuda(): use cpu_shims()."""
    p = os.path.join(REF, "classification")
    if p not in sys.path:
        sys.path.insert(0, p)
    import resnet_cifar  # noqa: the reference module
    return resnet_cifar


@contextlib.contextmanager
def cpu_shims():
    """Run reference code that names CUDA on a CPU-only host (device remap only)."""
    import torch

    real_tensor, real_zeros, real_ones = torch.tensor, torch.zeros, torch.ones
    real_cuda_ft = getattr(torch.cuda, "FloatTensor", None)
    real_t_cuda = torch.Tensor.cuda

    def tensor(*a, **k):
        if str(k.get("device", "")).startswith("cuda"):
            k["device"] = "cpu"
        return real_tensor(*a, **k)

    def ones(*a, **k):
        if str(k.get("device", "")).startswith("cuda"):
            k["device"] = "cpu"
        return real_ones(*a, **k)

    torch.tensor = tensor
    torch.ones = ones
    torch.cuda.FloatTensor = torch.FloatTensor
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        yield
    finally:
        torch.tensor = real_tensor
        torch.zeros = real_zeros
        torch.ones = real_ones
        torch.Tensor.cuda = real_t_cuda
        if real_cuda_ft is not None:
            torch.cuda.FloatTensor = real_cuda_ft


def load_functions(rel_path: str, names, class_name: str | None = None, extra_globals=None):
    """The UNMODIFIED source of selected functions / methods of a reference file whose module cannot be imported
    here (apex / mmcv / catalyst at import time): the file is parsed, the named FunctionDef nodes (module level, or
    inside `class_name`) are compiled as they stand and returned in a dict.  Nothing is edited or copied."""
    import ast

    import numpy as np
    import torch

    path = os.path.join(REF, rel_path)
    with open(path) as fh:
        tree = ast.parse(fh.read(), filename=path)
    body = tree.body
    if class_name is not None:
        body = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == class_name).body
    picked = [n for n in body if isinstance(n, ast.FunctionDef) and n.name in set(names)]
    missing = set(names) - {n.name for n in picked}
    if missing:
        raise KeyError(f"{rel_path}: no function(s) {sorted(missing)}")
    mod = ast.Module(body=picked, type_ignores=[])
    ns = {"torch": torch, "np": np, "nn": torch.nn}
    ns.update(extra_globals or {})
    exec(compile(mod, path, "exec"), ns)
    return {n: ns[n] for n in names}


def csv_path(name: str) -> str:
    sub = "coco_files" if name == "idf_91.csv" else "lvis_files"
    return os.path.join(REF, "instance_segmentation", sub, name)

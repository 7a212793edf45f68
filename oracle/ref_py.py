"""TEST INFRASTRUCTURE ONLY.  Import the reference's UNCHANGED Python hot-path files (staged byte-for-byte by
oracle/build_ref.py:stage_py into git-ignored oracle/_ref/py/) so that they can be *executed*:

  * on the GPU box on top of the drop-in shims (`vren`, `tinycudann`, `torch_scatter`, `faiss` -> ncn_b200): the
    reference's own models/rendering.py:render, models/ngp_mt.py:NGPMT (forward, update_density_grid, mark_invisible_cells),
    models/custom_functions.py autograd classes and losses.py:NeRFMTLoss then run against libncn.so - the drop-in claim of
    SURVEY.md section 8b, and the oracle for the mirrors in ncn_b200.{rendering,ngp,losses} and for FusedStep;
  * on CPU tensors (faiss="cpu": faiss.Kmeans bound to oracle.cluster_loss.spherical_kmeans, vren = inert stub): the
    reference's own losses.py as the CPU baseline of BASELINE.json config 1.

Nothing here is imported by the product.  Modules the reference imports at module level but never uses on the hot path
(h5py, imgviz) are satisfied by inert stubs when absent.
"""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if HERE not in sys.path:
    sys.path.insert(0, HERE)

_REF_TOP = ("models", "losses", "datasets")
_CACHE = {}


def staged_dir():
    import build_ref
    if build_ref.available():
        build_ref.stage_py(verbose=False)
    return build_ref.staged_py()


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _cpu_faiss():
    """faiss.Kmeans -> the numpy stand-in of SURVEY.md Appendix C (oracle/cluster_loss.py); index.search by dot product"""
    import numpy as np
    from oracle import cluster_loss as cl
    faiss = types.ModuleType("faiss")

    class _Index:
        def __init__(self, o):
            self.o = o

        def search(self, x, k):
            sim = x @ self.o.centroids.T
            i = sim.argmax(1)
            return sim[np.arange(len(x)), i][:, None], i[:, None].astype(np.int64)

    class Kmeans:
        last = None

        def __init__(self, d, k, niter=25, gpu=False, spherical=False, verbose=False):
            self.k, self.niter = k, niter
            self.index = _Index(self)

        def train(self, x):
            self.centroids, a = cl.spherical_kmeans(x, self.k, self.niter)
            Kmeans.last = dict(centroids=self.centroids.copy(), assign=a.copy(), x=x.copy())

    faiss.Kmeans = Kmeans
    contrib = types.ModuleType("faiss.contrib")
    tu = types.ModuleType("faiss.contrib.torch_utils")
    faiss.contrib = contrib
    contrib.torch_utils = tu
    return {"faiss": faiss, "faiss.contrib": contrib, "faiss.contrib.torch_utils": tu}


def load(faiss="shim", want_models=True, vren="shim"):
    """returns a namespace with .rendering, .ngp_mt, .custom_functions (want_models) and .losses, .hypersim_utils:
    the reference's own modules.  faiss="shim": k-means = libncn's GPU kernel through ncn_b200/shims/faiss (what a user of
    the drop-in gets); faiss="cpu": numpy stand-in, and `vren` an inert stub unless the shims were requested via want_models.
    vren="ref": the reference's models bind to the reference's OWN csrc kernels (oracle/_ref/vren_ref.so) instead of libncn -
    the GPU reference baseline of BASELINE.md section 3b (tiny-cuda-nn stays the shim: its source is not available)."""
    key = (faiss, want_models, vren)
    if key in _CACHE:
        return _CACHE[key]
    py = staged_dir()
    if py is None:
        return None
    import ncn_b200
    saved = {k: sys.modules.get(k) for k in list(sys.modules) if k.split(".")[0] in _REF_TOP + ("faiss", "vren", "h5py", "imgviz")}
    for k in saved:
        del sys.modules[k]
    injected = {}
    for name, attrs in (("h5py", {}), ("imgviz", dict(label_colormap=lambda *a, **k: None, depth2rgb=lambda *a, **k: None))):
        try:
            importlib.import_module(name)
        except Exception:  # noqa: BLE001
            injected[name] = _stub(name, **attrs)
    if faiss == "cpu":
        injected.update(_cpu_faiss())
        if not want_models:
            injected["vren"] = _stub("vren")
    if vren == "ref":
        import build_ref
        mod = build_ref.load()
        if mod is None:
            return None
        injected["vren"] = mod
    sys.modules.update(injected)
    path_before = list(sys.path)
    if want_models or faiss == "shim":
        ncn_b200.install_shims()
    sys.path.insert(0, py)
    try:
        ns = types.SimpleNamespace()
        if want_models:
            ns.custom_functions = importlib.import_module("models.custom_functions")
            ns.rendering = importlib.import_module("models.rendering")
            ns.ngp_mt = importlib.import_module("models.ngp_mt")
        ns.losses = importlib.import_module("losses")
        ns.hypersim_utils = importlib.import_module("datasets.hypersim_src.utils")
        ns.faiss = sys.modules["faiss"]
        ns.dir = py
        for mod in (getattr(ns, "custom_functions", None), getattr(ns, "rendering", None), getattr(ns, "ngp_mt", None), ns.losses, ns.hypersim_utils):
            if mod is not None:
                assert os.path.realpath(mod.__file__).startswith(os.path.realpath(py)), mod.__file__
    finally:
        # the reference's top-level names must not leak into the session (tests import `losses` etc. from ncn_b200)
        for k in [k for k in sys.modules if k.split(".")[0] in _REF_TOP or k in injected]:
            del sys.modules[k]
        for k, v in saved.items():
            if v is not None and k.split(".")[0] not in ("faiss", "vren"):
                sys.modules[k] = v
        sys.path[:] = [p for p in path_before]
    _CACHE[key] = ns
    return ns

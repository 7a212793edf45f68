"""CPU: the drop-in boundary.  When the reference tree is present (this container, not the GPU box)
its UNCHANGED files models/custom_functions.py, models/rendering.py and models/ngp_mt.py are imported with
ncn_b200.install_shims() providing `vren`, `tinycudann`, `torch_scatter`; the reference's NGPMT then builds on OUR
tcnn-style modules and must expose the same parameter names / sizes (the optimizer split and checkpoints key on them,
train_nerf.py:264-274), the same buffers, and autograd classes with the same argument lists as ours.  Runs in a
subprocess so the reference's top-level package names never leak into the test session."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"

PROBE = r'''
import inspect, json, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, %(root)r)
import ncn_b200
ncn_b200.install_shims()
sys.path.insert(0, %(ref)r)
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):
    import models.custom_functions as rcf
    import models.rendering as rr
    import models.ngp_mt as rng
    from ncn_b200 import custom_functions as ocf, rendering as orr, vren as ovren
    from ncn_b200.ngp import NGPMT
    kw = dict(scale=0.5, grid_size=128, rgb_act="Sigmoid", pred_sem=True, pred_norm=True, n_sem_cls=3)
    m_ref, m_our = rng.NGPMT(**kw), NGPMT(**kw)
import vren, tinycudann, torch_scatter
out = {
    "ref_params": {k: list(v.shape) for k, v in m_ref.named_parameters()},
    "our_params": {k: list(v.shape) for k, v in m_our.named_parameters()},
    "ref_buffers": {k: list(v.shape) for k, v in m_ref.named_buffers()},
    "our_buffers": {k: list(v.shape) for k, v in m_our.named_buffers()},
    "cascades": [m_ref.cascades, m_our.cascades],
    "classes": {n: [list(inspect.signature(getattr(rcf, n).forward).parameters), list(inspect.signature(getattr(ocf, n).forward).parameters)]
                for n in ("RayAABBIntersector", "RaySphereIntersector", "RayMarcher", "VolumeRenderer", "TruncExp")},
    "render_sig": [str(inspect.signature(rr.render)), str(inspect.signature(orr.render))],
    "vren_is_ours": vren.ray_aabb_intersect is ovren.ray_aabb_intersect,
    "vren_names": sorted(n for n in dir(vren) if not n.startswith("_") and callable(getattr(vren, n))),
    "encoder_type": type(m_ref.xyz_encoder).__module__,
    "segment_csr": callable(torch_scatter.segment_csr),
}
print("PROBE" + json.dumps(out))
'''


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="reference tree not present on this machine")
def test_unchanged_reference_files_bind_to_the_shims():
    r = subprocess.run([sys.executable, "-c", PROBE % dict(root=ROOT, ref=REF)], capture_output=True, text=True, cwd="/tmp", timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("PROBE")][0][5:])
    assert out["ref_params"] == out["our_params"]
    assert list(out["ref_params"]) == ["xyz_encoder.params", "sigma_net.params", "dir_encoder.params", "rgb_net.params",
                                       "sem_net.params", "norm_net.params"]
    assert out["ref_params"]["dir_encoder.params"] == [0]          # SphericalHarmonics: built, never evaluated (ngp_mt.py:94-101, 208)
    for k, v in out["ref_buffers"].items():                        # ours additionally registers density_grid / grid_coords itself
        assert out["our_buffers"][k] == v
    assert out["cascades"] == [1, 1]
    for name, (ref_args, our_args) in out["classes"].items():
        assert ref_args == our_args, name
    assert out["render_sig"][0] == out["render_sig"][1]
    assert out["vren_is_ours"] and out["segment_csr"]
    assert out["encoder_type"].endswith("tinycudann")
    binding = {"ray_aabb_intersect", "ray_sphere_intersect", "packbits", "morton3D", "morton3D_invert", "raymarching_train",
               "raymarching_test", "composite_train_fw", "composite_train_bw", "composite_train_multi_fw", "composite_train_multi_bw",
               "composite_test_fw", "composite_test_multi_fw", "distortion_loss_fw", "distortion_loss_bw"}      # binding.cpp:330-350
    assert binding <= set(out["vren_names"])


CKPT_PROBE = r'''
import io, contextlib, json, os, sys, tempfile, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, %(root)r)
import torch
import ncn_b200
from ncn_b200.ngp import NGPMT
sys.path.insert(0, %(ref)r)
import utils as ref_utils                      # the reference's checkpoint helpers (utils.py:4-39), unmodified
kw = dict(scale=0.5, grid_size=32, rgb_act="Sigmoid", pred_sem=True, pred_norm=True, n_sem_cls=3, log2_T=14)
with contextlib.redirect_stdout(io.StringIO()):
    a, b = NGPMT(**kw), NGPMT(**kw)
with torch.no_grad():
    for p in a.parameters():
        p.copy_(torch.randn_like(p))
    a.density_grid.uniform_(); a.density_bitfield.random_(0, 255)
# what pytorch-lightning writes: {'state_dict': {'model.<key>': tensor, ... other modules ...}}
sd = {"model." + k: v.clone() for k, v in a.state_dict().items()}
sd["val_lpips.net.weight"] = torch.zeros(3); sd["directions"] = torch.zeros(4, 3); sd["poses"] = torch.zeros(2, 3, 4)
path = os.path.join(tempfile.mkdtemp(), "last.ckpt")
torch.save({"state_dict": sd, "epoch": 3}, path)
ref_utils.load_ckpt(b, path)
same = all(torch.equal(x, y) for x, y in zip(a.state_dict().values(), b.state_dict().values()))
slim = ref_utils.slim_ckpt(path)
c = NGPMT(**kw)
torch.save({"state_dict": slim}, path)
ref_utils.load_ckpt(c, path, prefixes_to_ignore=["density_grid", "grid_coords"])
same_params = all(torch.equal(x, y) for x, y in zip(a.parameters(), c.parameters()))
print("PROBE" + json.dumps({"same": same, "same_params_after_slim": same_params, "keys": sorted(a.state_dict().keys()),
                            "slim_dropped": sorted(set(sd) - set(slim)), "bitfield_kept": bool(torch.equal(a.density_bitfield, c.density_bitfield))}))
'''


@pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "utils.py")), reason="reference tree not present on this machine")
def test_reference_checkpoint_helpers_round_trip_our_model():
    """SURVEY section 8 row f4 (checkpoint key compatibility): a Lightning-style checkpoint of our NGPMT goes through the
    reference's own load_ckpt / slim_ckpt (utils.py:4-39) and restores parameters and occupancy state key for key."""
    r = subprocess.run([sys.executable, "-c", CKPT_PROBE % dict(root=ROOT, ref=REF)], capture_output=True, text=True, cwd="/tmp", timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    out = json.loads([l for l in r.stdout.splitlines() if l.startswith("PROBE")][0][5:])
    assert out["same"] and out["same_params_after_slim"] and out["bitfield_kept"]
    assert out["keys"] == sorted(["center", "xyz_min", "xyz_max", "half_size", "density_bitfield", "density_grid", "grid_coords",
                                  "xyz_encoder.params", "sigma_net.params", "dir_encoder.params", "rgb_net.params", "sem_net.params",
                                  "norm_net.params"])
    assert out["slim_dropped"] == sorted(["directions", "model.density_grid", "model.grid_coords", "poses", "val_lpips.net.weight"])

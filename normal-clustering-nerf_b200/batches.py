"""Host side of the training batch either side of the path: what the reference's dataset object
prepares ONCE before training and what a DataLoader worker hands over per step, in the form the fused step consumes.

  generate_random_poses   datasets/base.py:235-263 (+ its helpers :186-232): the 10 000 generated camera poses of
                          --random_tr_poses - positions uniform in the inner 80 % of the training cameras' bounding box, every
                          camera looking AWAY from the point nearest to all training optical axes, up = mean training up vector
  sample_batch_indices    datasets/base.py:94-173 on the host with numpy (the DataLoader form; the device form is
                          ncn_sample_ray_batch[_ex] + ncn_sample_random_pose_half inside the step graph)

Everything here is set-up / host code by nature (numpy, float64 like the reference); nothing in the step calls it.
"""
import numpy as np
import torch


def _unit(v):
    return v / np.linalg.norm(v, axis=-1, keepdims=True)


def _frames(forward, up):
    """camera frames (n,3,3) with columns [right, true up, forward] from forward directions (n,3) and one up hint (3,)"""
    f = _unit(forward)
    r = _unit(np.cross(up[None, :], f))
    u = _unit(np.cross(f, r))
    return np.stack([r, u, f], axis=-1)


def focus_point(poses):
    """least-squares point nearest to all optical axes (datasets/base.py:205-215): with P_i = I - d_i d_i^T the projector off the
    axis of camera i (the reference negates d first - P is even in d - and squares P, which is idempotent only for unit d; both
    are kept: A = mean(P_i^T P_i), b = mean(P_i^T P_i o_i), x = A^-1 b)"""
    d = -poses[:, :3, 2]             # the outer product stays in the poses' own precision (fp32 in the reference), the rest is fp64
    o = poses[:, :3, 3].astype(np.float64)
    P = np.eye(3)[None] - d[:, :, None] * d[:, None, :]
    PtP = np.einsum("nji,njk->nik", P, P)
    return np.linalg.inv(PtP.mean(0)) @ np.einsum("nij,nj->ni", PtP, o).mean(0)


def average_pose(poses):
    """(3,4) pose at the mean position with the mean viewing direction and up vector (datasets/base.py:196-202)"""
    pos = poses[:, :3, 3].mean(0)
    frame = _frames(poses[:, :3, 2].mean(0)[None], poses[:, :3, 1].mean(0))[0]
    return np.concatenate([frame, pos[:, None]], axis=1)


def generate_random_poses(poses, xyz_cam_min, xyz_cam_max, n_poses=10000, focus_jitter=False, rng=None):
    """-> (random_poses (n_poses,3,4) float32 torch tensor, average pose (3,4) numpy).  `rng`: a numpy RandomState (default: the
    global numpy stream, which is what the reference draws from - with the same seed the same poses come out: per pose three
    uniforms, plus three normals when focus_jitter (`random_pose_focusptjitter`))."""
    rng = np.random if rng is None else rng
    poses = np.asarray(torch.as_tensor(poses).detach().cpu().numpy() if torch.is_tensor(poses) else poses)[:, :3, :]
    lo = np.asarray(torch.as_tensor(xyz_cam_min).cpu().numpy() if torch.is_tensor(xyz_cam_min) else xyz_cam_min)
    hi = np.asarray(torch.as_tensor(xyz_cam_max).cpu().numpy() if torch.is_tensor(xyz_cam_max) else xyz_cam_max)
    up = poses[:, :3, 1].mean(0)
    target = focus_point(poses)
    if focus_jitter:                 # draw order of the reference's loop: rand(3) then randn(3), pose by pose
        u = np.empty((n_poses, 3)); jit = np.empty((n_poses, 3))
        for i in range(n_poses):
            u[i] = rng.rand(3)
            jit[i] = rng.randn(3)
        targets = target[None] + jit * 0.125
    else:                            # one (n,3) draw consumes the stream exactly like n draws of 3
        u = rng.rand(n_poses, 3)
        targets = np.broadcast_to(target[None], (n_poses, 3))
    pos = lo + (hi - lo) * (u * 0.8 + 0.1)
    frames = _frames(-(targets - pos), up)          # z axis faces away from the focus point, as the training cameras' does
    out = np.concatenate([frames, pos[:, :, None]], axis=2)
    return torch.as_tensor(out).to(torch.float32), average_pose(poses)


STRATEGIES = ("all_images", "same_image", "all_images_triang", "same_image_triang", "all_images_triang_patch", "same_image_triang_patch")


def sample_batch_indices(strategy, batch_size, n_poses, height, width, patch_size=8, max_expand=0, random_tr_poses=False,
                         n_random_poses=0, rng=None):
    """index half of BaseDataset.__getitem__ (datasets/base.py:94-173) -> dict(img_idxs, pix_idxs[, rnd_img_idxs]) of int64 numpy
    arrays.  Same draws in the same order as the reference (np.random.choice(n, size) == randint(0, n, size) on the legacy
    stream), same index arithmetic, incl. the patch-corner quirk (the corner is an INDEX into valid_idx['patch_corners'] and that
    index is what the pixel offsets are added to, :164-166) and the triangle expansion (:130-141)."""
    rng = np.random if rng is None else rng
    H, W = int(height), int(width)
    choice = lambda n, size: rng.choice(n, size)
    out = {}
    if strategy == "all_images":
        out["img_idxs"] = choice(n_poses, batch_size)
        out["pix_idxs"] = choice(H * W, batch_size)
        return out
    if strategy == "same_image":
        out["img_idxs"] = np.full(batch_size, choice(n_poses, 1)[0])
        out["pix_idxs"] = choice(H * W, batch_size)
        return out
    patches = strategy.endswith("_patch")
    group = patch_size * patch_size if patches else 3
    n_groups = batch_size // group
    if random_tr_poses:
        n_groups //= 2
    same = strategy.startswith("same_image")
    if random_tr_poses:              # the generated poses are drawn BEFORE the training images (base.py:109-113, 148-152)
        r = choice(n_random_poses, 1)[0] if same else choice(n_random_poses, n_groups)
        out["rnd_img_idxs"] = np.full(group * n_groups, r) if same else np.repeat(r, group)
    i = choice(n_poses, 1)[0] if same else choice(n_poses, n_groups)
    out["img_idxs"] = np.full(group * n_groups, i) if same else np.repeat(i, group)
    if patches:
        corner = choice((H - patch_size + 1) * (W - patch_size + 1), n_groups)
        dy, dx = np.divmod(np.arange(group), patch_size)
        out["pix_idxs"] = (corner[:, None] + (dy * W + dx)[None, :]).reshape(-1)
        return out
    t = choice((H - 2) * (W - 2), n_groups)          # valid_idx['x1']: pixels with a row above and a column to the left
    y, x = 1 + t // (W - 2), 1 + t % (W - 2)
    x1 = y * W + x
    x2, x3 = x1 - W, x1 - 1
    if max_expand > 0:
        e = int(max_expand)
        x1 = np.where(x1 + e * W < H * W, x1 + e * W, x1)
        x2 = np.where(x2 - e * W >= 0, x2 - e * W, x2)
        x3 = np.where((x3 - e) // W == x3 // W, x3 - e, x3)
    out["pix_idxs"] = np.stack([x1, x2, x3], axis=1).reshape(-1)
    return out


class HostBatcher:
    """What a DataLoader worker of the reference does per step (BaseDataset.__getitem__, datasets/base.py:94-183, training split),
    in the form the fused step consumes: draw the batch indices (sample_batch_indices), gather the target colours (and labels)
    from HOST-resident images, and write ONE packed record - [img_idx i64 (R) | pix_idx i64 (R) | rgb f32 (R,3)], 28 bytes per ray,
    FusedStep.step_pixels / pack_pixel_batch - into a ring of pinned buffers, so the step needs a single host->device copy and the
    host can prepare record i+1 while the copy of record i is in flight (ring >= 2).

    images (P, H*W, 3) float32 (the reference's `self.rays[..., :3]`); labels: optional (P, H*W) integer map gathered for the
    training-view rays (e.g. `semantics`).  With random_tr_poses the record holds [training views | the same pixels with image index
    n_poses + rnd_img_idx] (the generated poses follow the training poses in trainer.set_cameras(random_poses=...))."""

    def __init__(self, images, height, width, strategy, batch_size, patch_size=8, max_expand=0, random_tr_poses=False, n_random_poses=0,
                 labels=None, ring=4, rng=None, pin=None):
        self.images = torch.as_tensor(images, dtype=torch.float32)
        P = self.images.shape[0]
        if self.images.shape != (P, height * width, 3):
            raise ValueError("HostBatcher: images must be (P, H*W, 3)")
        if strategy not in STRATEGIES:
            raise ValueError(f"HostBatcher: unknown ray_sampling_strategy {strategy!r}")
        if random_tr_poses and ("triang" not in strategy or n_random_poses < 1):      # asserted by the reference (base.py:98-102)
            raise ValueError("HostBatcher: random_tr_poses needs a triangle / patch strategy and generated poses")
        self.labels = None if labels is None else torch.as_tensor(labels)
        self.args = dict(strategy=strategy, batch_size=int(batch_size), n_poses=P, height=int(height), width=int(width),
                         patch_size=int(patch_size), max_expand=int(max_expand), random_tr_poses=bool(random_tr_poses),
                         n_random_poses=int(n_random_poses))
        self.rng = np.random if rng is None else rng
        probe = sample_batch_indices(rng=np.random.RandomState(0), **self.args)
        self.n_gt = len(probe["pix_idxs"])
        self.n_rays = self.n_gt * (2 if random_tr_poses else 1)          # rows of the record = FusedStep's batch_size
        pin = torch.cuda.is_available() if pin is None else pin
        self.ring = [torch.zeros(self.n_rays * 28, dtype=torch.uint8, pin_memory=pin) for _ in range(max(1, int(ring)))]
        self.pos = 0
        self.last = None

    def views(self, rec):
        """(img_idx (R) i64, pix_idx (R) i64, rgb (R,3) f32) views into a record"""
        R = self.n_rays
        return rec[:8 * R].view(torch.int64), rec[8 * R:16 * R].view(torch.int64), rec[16 * R:].view(torch.float32).view(R, 3)

    def next(self):
        """-> (record, labels of the training-view rays or None).  The record is a slot of the ring: it is overwritten `ring` calls
        later, so at most ring - 1 copies may be in flight."""
        s = sample_batch_indices(rng=self.rng, **self.args)
        self.last = s
        rec = self.ring[self.pos % len(self.ring)]
        self.pos += 1
        b_img, b_pix, b_rgb = self.views(rec)
        n = self.n_gt
        img = torch.from_numpy(np.ascontiguousarray(np.broadcast_to(s["img_idxs"], (n,)).astype(np.int64)))
        pix = torch.from_numpy(np.ascontiguousarray(s["pix_idxs"].astype(np.int64)))
        b_img[:n] = img; b_pix[:n] = pix
        torch.index_select(self.images.view(-1, 3), 0, img * self.images.shape[1] + pix, out=b_rgb[:n])
        if self.args["random_tr_poses"]:
            b_img[n:] = torch.from_numpy(s["rnd_img_idxs"].astype(np.int64)) + self.args["n_poses"]
            b_pix[n:] = pix
            b_rgb[n:] = 0.0
        lab = None if self.labels is None else self.labels[img, pix]
        return rec, lab

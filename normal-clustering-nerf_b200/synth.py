"""Synthetic Hypersim-/ScanNet-shaped scenes and ray batches.

No dataset exists on the box, so tests and bench.py use a procedurally generated room:
  * model box [-scale, scale]^3 (scale 0.5 -> cascades 1), occupancy grid G^3 in Morton order;
  * occupied cells = a 2-voxel shell around the walls of the room [-0.4,0.4]^3 plus six
    axis-aligned cuboids ("furniture");
  * pinhole cameras (Hypersim 1024x768 fx=fy=886.81, datasets/hypersim_src/scene.py:92-93;
    ScanNet 640x480 fx=fy=577.87, datasets/scannet_manhattan_src/scene.py:54) on a closed path
    inside [-0.25,0.25]^3 looking at jittered wall targets, unit-norm ray directions
    (datasets/hypersim_src/cam_model.py:192-194);
  * ray batches follow the `all_images_triang_patch` recipe (datasets/base.py:142-171):
    128 patches of 8x8 pixels = 8192 rays, 49 triangles per patch -> 6272 normals.
Everything is seeded and generated with numpy on the host (device-independent).
"""
import numpy as np

CAMERAS = {
    "hypersim": dict(W=1024, H=768, fx=886.81, fy=886.81, cx=512.0, cy=384.0),
    "scannet": dict(W=640, H=480, fx=577.87, fy=577.87, cx=319.5, cy=239.5),
}


def _spread3(v):
    v = v.astype(np.uint32)
    v = (v * np.uint32(0x00010001)) & np.uint32(0xFF0000FF)
    v = (v * np.uint32(0x00000101)) & np.uint32(0x0F00F00F)
    v = (v * np.uint32(0x00000011)) & np.uint32(0xC30C30C3)
    v = (v * np.uint32(0x00000005)) & np.uint32(0x49249249)
    return v


def morton3d_np(x, y, z):
    return _spread3(x) | (_spread3(y) << np.uint32(1)) | (_spread3(z) << np.uint32(2))


def room_occupancy(grid_size=128, scale=0.5, seed=0):
    """bool (G,G,G) occupancy in xyz order: wall shell + 6 cuboids."""
    G = grid_size
    rng = np.random.RandomState(seed)
    c = (np.arange(G) + 0.5) / G * 2 * scale - scale          # cell centres
    X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
    vox = 2 * scale / G
    wall = 0.4 * scale / 0.5
    d_wall = np.minimum.reduce([np.abs(np.abs(X) - wall), np.abs(np.abs(Y) - wall), np.abs(np.abs(Z) - wall)])
    inside = (np.abs(X) <= wall + 2 * vox) & (np.abs(Y) <= wall + 2 * vox) & (np.abs(Z) <= wall + 2 * vox)
    occ = (d_wall <= 2 * vox) & inside
    for _ in range(6):
        ctr = rng.uniform(-0.3, 0.3, 3) * scale / 0.5
        ctr[2] = -wall + rng.uniform(0.02, 0.15)            # standing on the floor
        half = rng.uniform(0.03, 0.10, 3) * scale / 0.5
        occ |= (np.abs(X - ctr[0]) <= half[0]) & (np.abs(Y - ctr[1]) <= half[1]) & (np.abs(Z - ctr[2]) <= half[2])
    return occ


def density_grid_from_occupancy(occ, value=10.0):
    """float32 (1, G^3) density grid in MORTON order (models/ngp_mt.py:153-156 layout)."""
    G = occ.shape[0]
    idx = np.arange(G, dtype=np.uint32)
    X, Y, Z = np.meshgrid(idx, idx, idx, indexing="ij")
    m = morton3d_np(X.ravel(), Y.ravel(), Z.ravel()).astype(np.int64)
    grid = np.zeros(G ** 3, dtype=np.float32)
    grid[m] = np.where(occ.ravel(), np.float32(value), np.float32(0))
    return grid[None, :]


def packbits_np(density_grid, thr):
    """numpy restatement of vren.packbits (raymarching.cu:122-141): LSB-first."""
    bits = (density_grid.reshape(-1) > np.float32(thr))
    return np.packbits(bits, bitorder="little")


def camera_poses(n_poses=50, seed=0):
    """(n,3,4) camera-to-world matrices on a closed path inside [-0.25,0.25]^3."""
    rng = np.random.RandomState(seed + 1)
    poses = np.zeros((n_poses, 3, 4), dtype=np.float32)
    for i in range(n_poses):
        a = 2 * np.pi * i / n_poses
        pos = np.array([0.2 * np.cos(a), 0.2 * np.sin(a), 0.05 * np.sin(2 * a)])
        target = rng.uniform(-0.4, 0.4, 3)
        target[rng.randint(3)] = 0.4 * rng.choice([-1, 1])       # on a wall
        fwd = target - pos
        fwd /= np.linalg.norm(fwd)
        up = np.array([0.0, 0.0, 1.0])
        right = np.cross(fwd, up)
        if np.linalg.norm(right) < 1e-3:
            right = np.array([1.0, 0.0, 0.0])
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        poses[i, :, 0], poses[i, :, 1], poses[i, :, 2], poses[i, :, 3] = right, down, fwd, pos
    return poses


def pixel_directions(cam="hypersim"):
    """(H*W,3) unit-norm camera-space directions."""
    k = CAMERAS[cam]
    u, v = np.meshgrid(np.arange(k["W"], dtype=np.float32), np.arange(k["H"], dtype=np.float32), indexing="xy")
    d = np.stack([(u - k["cx"] + 0.5) / k["fx"], (v - k["cy"] + 0.5) / k["fy"], np.ones_like(u)], -1)
    d /= np.linalg.norm(d, axis=-1, keepdims=True)
    return d.reshape(-1, 3).astype(np.float32)


def rays_from(poses, dirs, img_idx, pix_idx):
    """rays_o, rays_d (n,3) float32 (datasets/ray_utils.py:46-71 get_rays)."""
    R = poses[img_idx, :, :3]                                 # (n,3,3)
    rays_d = np.einsum("nij,nj->ni", R, dirs[pix_idx]).astype(np.float32)
    rays_o = poses[img_idx, :, 3].astype(np.float32)
    return np.ascontiguousarray(rays_o), np.ascontiguousarray(rays_d)


def patch_batch(n_rays=8192, cam="hypersim", n_poses=50, seed=0, patch=8):
    """`all_images_triang_patch` batch: returns dict(rays_o, rays_d, img_idx, pix_idx, tri (3,M) index triplets)."""
    k = CAMERAS[cam]
    rng = np.random.RandomState(seed + 2)
    n_patches = n_rays // (patch * patch)
    img = rng.randint(0, n_poses, n_patches)
    x0 = rng.randint(0, k["W"] - patch, n_patches)
    y0 = rng.randint(0, k["H"] - patch, n_patches)
    dy, dx = np.meshgrid(np.arange(patch), np.arange(patch), indexing="ij")
    px = (x0[:, None, None] + dx[None]).reshape(n_patches, -1)
    py = (y0[:, None, None] + dy[None]).reshape(n_patches, -1)
    pix_idx = (py * k["W"] + px).reshape(-1)
    img_idx = np.repeat(img, patch * patch)
    # triangles: (i,j),(i+1,j),(i,j+1) inside each patch -> (patch-1)^2 per patch
    ii, jj = np.meshgrid(np.arange(patch - 1), np.arange(patch - 1), indexing="ij")
    base = (ii * patch + jj).reshape(-1)
    off = (np.arange(n_patches) * patch * patch)[:, None]
    x1 = (off + base[None]).reshape(-1)
    x2 = (off + base[None] + patch).reshape(-1)
    x3 = (off + base[None] + 1).reshape(-1)
    poses = camera_poses(n_poses, seed)
    dirs = pixel_directions(cam)
    rays_o, rays_d = rays_from(poses, dirs, img_idx, pix_idx)
    return dict(rays_o=rays_o, rays_d=rays_d, img_idx=img_idx, pix_idx=pix_idx,
                tri=np.stack([x1, x2, x3]).astype(np.int64))


def random_batch(n_rays=8192, cam="hypersim", n_poses=50, seed=0):
    k = CAMERAS[cam]
    rng = np.random.RandomState(seed + 3)
    img_idx = rng.randint(0, n_poses, n_rays)
    pix_idx = rng.randint(0, k["W"] * k["H"], n_rays)
    rays_o, rays_d = rays_from(camera_poses(n_poses, seed), pixel_directions(cam), img_idx, pix_idx)
    return dict(rays_o=rays_o, rays_d=rays_d, img_idx=img_idx, pix_idx=pix_idx)


def full_image(pose_idx=0, cam="hypersim", n_poses=50, seed=0):
    k = CAMERAS[cam]
    pix_idx = np.arange(k["W"] * k["H"])
    img_idx = np.full_like(pix_idx, pose_idx)
    rays_o, rays_d = rays_from(camera_poses(n_poses, seed), pixel_directions(cam), img_idx, pix_idx)
    return dict(rays_o=rays_o, rays_d=rays_d, img_idx=img_idx, pix_idx=pix_idx)


def manhattan_normals(n=8192, seed=0, noise=0.05, frac_axes=0.7, frac_zero=0.01):
    """Config-1 input: unit normals, 70 % around +-3 axes of a random rotation, 30 % uniform, 1 % zeros."""
    rng = np.random.RandomState(seed + 4)
    q, _ = np.linalg.qr(rng.randn(3, 3))
    if np.linalg.det(q) < 0:
        q[:, 0] = -q[:, 0]
    n_ax = int(n * frac_axes)
    axes = q[:, rng.randint(0, 3, n_ax)].T * rng.choice([-1.0, 1.0], (n_ax, 1))
    a = axes + noise * rng.randn(n_ax, 3)
    u = rng.randn(n - n_ax, 3)
    x = np.concatenate([a, u])
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    x = x[rng.permutation(n)]
    x[rng.choice(n, int(n * frac_zero), replace=False)] = 0
    return x.astype(np.float32), q.astype(np.float32)

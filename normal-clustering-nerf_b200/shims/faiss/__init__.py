"""`import faiss` (losses.py:8) -> the single-CTA spherical k-means kernel of libncn.so.

Only the surface the reference touches (losses.py:86-92, train_nerf.py:495-502):
    km = faiss.Kmeans(d, k=K, niter=niter, gpu=False, spherical=True, verbose=False)
    km.train(x_np); D, I = km.index.search(x_np, 1); km.centroids
faiss itself is absent from the reference tree and unpinned (parity unpinned).
"""
import numpy as np
import torch

import ncn_b200  # noqa: F401
from ncn_b200.clustering import kmeans_spherical


class _Index:
    def __init__(self, owner):
        self._o = owner

    def search(self, x, k):
        assert k == 1
        c = self._o.centroids
        x = np.ascontiguousarray(x, dtype=np.float32)
        sim = x @ c.T if self._o.spherical else -((x[:, None, :] - c[None]) ** 2).sum(-1)
        idx = sim.argmax(1)
        return sim[np.arange(len(x)), idx][:, None].astype(np.float32), idx[:, None].astype(np.int64)


class Kmeans:
    def __init__(self, d, k, niter=25, gpu=False, spherical=False, verbose=False, seed=1234,
                 max_points_per_centroid=256, **kw):
        if d != 3:
            raise NotImplementedError("ncn faiss shim: d must be 3 (surface normals)")
        self.d, self.k, self.niter, self.spherical, self.seed = d, k, niter, spherical, seed
        self.max_points_per_centroid = max_points_per_centroid
        self.centroids = None
        self.index = _Index(self)

    def train(self, x):
        xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32)).cuda()
        cent, _assign, _nv = kmeans_spherical(xt, self.k, self.niter, seed=self.seed,
                                              max_points_per_centroid=self.max_points_per_centroid,
                                              spherical=self.spherical)
        self.centroids = cent.cpu().numpy()
        return 0.0

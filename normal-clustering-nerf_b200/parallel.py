"""Data-parallel sharding helpers: rays are independent, so rank g takes a contiguous
block of whole patches / triangles (normals never straddle ranks), clustering stays rank-local like the reference's
DDP (train_nerf.py:949-952), and the only collective is one sum all-reduce of the flat gradient followed by a
division by world_size inside the Adam kernel (FlatAdam.grad_div)."""


def shard_rays(n_rays_global, rank, world_size, unit=64):
    """[start, stop) of this rank's rays; `unit` = rays that must stay together (64 = one 8x8 patch, 3 = one triangle)."""
    if n_rays_global % unit:
        raise ValueError("global ray count must be a multiple of the sampling unit")
    units = n_rays_global // unit
    per, rem = divmod(units, world_size)
    start = rank * per + min(rank, rem)
    stop = start + per + (1 if rank < rem else 0)
    return start * unit, stop * unit


def shard_tiles(n_pixels, rank, world_size):
    """contiguous image tile of a full-image evaluation render (config 4: 786 432 pixels -> 98 304 per GPU at 8 ranks)"""
    per, rem = divmod(n_pixels, world_size)
    start = rank * per + min(rank, rem)
    return start, start + per + (1 if rank < rem else 0)


def dp_mean_gradient(summed_grad, world_size):
    """what the Adam kernel applies after the sum all-reduce (per-rank losses are means over equal shards)"""
    return summed_grad / float(world_size)

"""`from torch_scatter import segment_csr` (models/custom_functions.py:4) -> warp-per-segment sum kernel."""
import ncn_b200  # noqa: F401
from ncn_b200.custom_functions import segment_sum


def segment_csr(src, indptr, out=None, reduce="sum"):
    if reduce != "sum":
        raise NotImplementedError("ncn torch_scatter shim: only reduce='sum' (the one the reference uses)")
    res = segment_sum(src, indptr)
    if out is not None:
        out.copy_(res)
        return out
    return res.to(src.dtype)

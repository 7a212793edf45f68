"""`import tinycudann as tcnn` resolves here when ncn_b200.install_shims() is active."""
import ncn_b200  # noqa: F401
from ncn_b200.tinycudann import Encoding, Network, NetworkWithInputEncoding  # noqa: F401

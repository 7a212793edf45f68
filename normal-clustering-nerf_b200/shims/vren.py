"""`import vren` resolves here when ncn_b200.install_shims() is active: the reference's
models/custom_functions.py, models/rendering.py, models/ngp_mt.py and losses.py then call
the sm_100a kernels without modification."""
import ncn_b200  # noqa: F401
from ncn_b200.vren import *  # noqa: F401,F403
from ncn_b200.vren import (ray_aabb_intersect, ray_sphere_intersect, packbits, morton3D, morton3D_invert,  # noqa: F401
                           raymarching_train, raymarching_test, composite_train_fw, composite_train_multi_fw,
                           composite_train_bw, composite_train_multi_bw, composite_test_fw,
                           composite_test_multi_fw, distortion_loss_fw, distortion_loss_bw)

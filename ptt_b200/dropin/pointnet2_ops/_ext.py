"""`pointnet2_ops._ext` backed by the sm_100a kernels of libptt_b200.so.

Put `ptt_b200/dropin` on sys.path ahead of any other `pointnet2_ops` (or call
`ptt_b200.install_dropin()`), and the reference's `import pointnet2_ops._ext as _ext`
(pointnet2_utils.py:24) resolves here with the reference tree untouched.  Same function names,
positional arguments and return types as upstream's pybind module; errors are RuntimeError (upstream
`TORCH_CHECK`s, and calls exit() on launch failures -- this never does).  CUDA tensors only, as upstream.
"""
from ptt_b200.ops import (  # noqa: F401
    ball_query,
    furthest_point_sampling,
    furthest_point_sampling_with_dist,
    gather_points,
    gather_points_grad,
    group_points,
    group_points_grad,
    three_interpolate,
    three_interpolate_grad,
    three_nn,
)

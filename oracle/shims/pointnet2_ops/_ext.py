"""CPU `pointnet2_ops._ext` stand-in so the reference's own ptt.models can run in this container.

TEST INFRASTRUCTURE.  Upstream has no CPU path; every call lands in oracle/pointnet2_ref.c.
"""
from oracle.cops import (  # noqa: F401
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

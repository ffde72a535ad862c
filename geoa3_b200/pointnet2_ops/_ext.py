"""Stand-in for the reference's pybind module `pointnet2_ops._ext`
(_ext-src/src/bindings.cpp:6-19): the same nine function names and argument orders, taking and
returning torch CUDA tensors, implemented by the C ABI of libgeoa3_b200.so."""
from ..ops import (ball_query, furthest_point_sampling, gather_points, gather_points_grad, group_points,  # noqa: F401
                   group_points_grad, three_interpolate, three_interpolate_grad, three_nn)

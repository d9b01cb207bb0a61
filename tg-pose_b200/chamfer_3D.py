"""Drop-in for the reference's pybind extension module `chamfer_3D`
(losses/chamfer3D/chamfer_cuda.cpp:30-33): forward(...) / backward(...) on caller-allocated CUDA
tensors, returning 1 on success and 0 on failure like chamfer3D.cu:145-151,187-193.
Kernels run on PyTorch's current stream (the reference used the legacy default stream)."""
import sys

from . import ops


def forward(xyz1, xyz2, dist1, dist2, idx1, idx2):
    try:
        ops.chamfer_forward(xyz1, xyz2, dist1, dist2, idx1, idx2)
        return 1
    except RuntimeError as e:  # the reference printf()s and returns 0
        print(f"error in nnd updateOutput: {e}", file=sys.stderr)
        return 0


def backward(xyz1, xyz2, gradxyz1, gradxyz2, graddist1, graddist2, idx1, idx2):
    try:
        ops.chamfer_backward(xyz1, xyz2, graddist1, graddist2, idx1, idx2, gradxyz1, gradxyz2)
        return 1
    except RuntimeError as e:
        print(f"error in nnd get grad: {e}", file=sys.stderr)
        return 0

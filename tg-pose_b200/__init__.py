"""tg-pose_b200: sm_100a implementation of TG-Pose's 3D-GCN + chamfer3D hot path.

Host side (this package) mirrors the reference's Python interfaces for the path:
  gcn3d            -- drop-in for network/fs_net_repo/gcn3d.py (same names, signatures, parameter names)
  chamfer_3D       -- drop-in for the pybind module built from losses/chamfer3D/chamfer_cuda.cpp
  dist_chamfer_3D  -- drop-in for losses/chamfer3D/dist_chamfer_3D.py (chamfer_3DFunction / chamfer_3DDist)
  face_enc         -- Face_Enc (FaceRecon.py:12-86) on the fused kernels
All compute goes through the C-ABI library libtgpose_b200.so (include/tgpose_b200.h); there is no
CPU or PyTorch fallback -- ops raise if the library or a CUDA device is missing.
"""
__version__ = "0.1.0"

"""Stage / compile the UNMODIFIED reference into oracle/_ref/ (git-ignored, travels to the GPU box with gpurun).

TEST / BASELINE INFRASTRUCTURE ONLY.  Nothing under tg-pose_b200/ imports this.  Run here (the build container, where
/root/reference exists) by __graft_entry__.build(); the GPU box only uses the files it finds.

  oracle/_ref/pyref/      the reference's own CPU path for BASELINE configs[0-1]: config/config.py and
                          network/fs_net_repo/{gcn3d,FaceRecon,PoseNet9D,PoseR,PoseTs}.py, byte-for-byte (sha256 recorded
                          in oracle/_ref/MANIFEST.json); imported by bench.py --impl reference / cpu_baseline (kind "reference")
  oracle/_ref/chamfer3D/  losses/chamfer3D/{chamfer_cuda.cpp,chamfer3D.cu} compiled for sm_100a from where they lie
                          (torch.utils.cpp_extension, pybind module `chamfer_3D`): the GPU oracle of SURVEY 8c and the
                          kernel to beat (tests/test_gpu_ref_chamfer.py, bench.py kernel_rooflines.chamfer_fwd)
No reference source is committed: _ref/ is listed in .gitignore.
"""
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("TGPOSE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
PYREF = os.path.join(OUT, "pyref")
CHAMFER = os.path.join(OUT, "chamfer3D")

PY_FILES = [
    "config/__init__.py", "config/config.py",
    "network/__init__.py", "network/fs_net_repo/__init__.py",
    "network/fs_net_repo/gcn3d.py", "network/fs_net_repo/FaceRecon.py", "network/fs_net_repo/PoseNet9D.py",
    "network/fs_net_repo/PoseR.py", "network/fs_net_repo/PoseTs.py",
    # CPU chamfer the reference's own unit test compares its kernel with (losses/metrics/CD/unit_test.py:14-35)
    "losses/metrics/CD/chamfer_python.py",
]
CHAMFER_SOURCES = ["losses/chamfer3D/chamfer_cuda.cpp", "losses/chamfer3D/chamfer3D.cu"]


def _sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def have_reference():
    return os.path.isdir(os.path.join(REF, "network", "fs_net_repo"))


def stage_python():
    man = {}
    for rel in PY_FILES:
        src, dst = os.path.join(REF, rel), os.path.join(PYREF, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        if os.path.exists(src):
            shutil.copyfile(src, dst)
        else:                       # an absent __init__.py is an implicit namespace package in the reference too
            open(dst, "w").close()
        man[rel] = _sha(dst)
    return man


def chamfer_so():
    if not os.path.isdir(CHAMFER):
        return None
    for f in os.listdir(CHAMFER):
        if f.startswith("chamfer_3D") and f.endswith(".so"):
            return os.path.join(CHAMFER, f)
    return None


def build_chamfer(verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a on the reference's two files, in place."""
    srcs = [os.path.join(REF, s) for s in CHAMFER_SOURCES]
    so = chamfer_so()
    if so and all(os.path.getmtime(so) >= os.path.getmtime(s) for s in srcs):
        return so
    os.makedirs(CHAMFER, exist_ok=True)
    os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0")
    os.environ.pop("CC", None)
    os.environ.pop("CXX", None)
    from torch.utils import cpp_extension
    cpp_extension.load(name="chamfer_3D", sources=srcs, build_directory=CHAMFER, verbose=verbose,
                       extra_cuda_cflags=["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo"])
    return chamfer_so()


def build(verbose=False):
    if not have_reference():
        return None
    os.makedirs(OUT, exist_ok=True)
    man = {"reference": REF, "python": stage_python()}
    try:
        so = build_chamfer(verbose)
        man["chamfer_3D"] = {"so": os.path.relpath(so, OUT) if so else None,
                             "sources": {s: _sha(os.path.join(REF, s)) for s in CHAMFER_SOURCES}}
    except Exception as e:          # recorded, not fatal: the CPU arm does not need it
        man["chamfer_3D"] = {"so": None, "error": f"{type(e).__name__}: {e}"[:500]}
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(man, f, indent=1)
    return man


# ----------------------------------------------------------------------------- loaders (tests / bench only)
def load_chamfer():
    """the reference's pybind module `chamfer_3D` (forward/backward, chamfer_cuda.cpp:30-33), or None."""
    so = chamfer_so()
    if so is None:
        return None
    import importlib.util
    import torch  # noqa: F401  (libtorch must be loaded before the extension)
    spec = importlib.util.spec_from_file_location("chamfer_3D", so)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def import_pyref():
    """import the staged reference modules; returns (gcn3d, FaceRecon, PoseNet9D) modules or None if not staged."""
    if not os.path.exists(os.path.join(PYREF, "network", "fs_net_repo", "PoseNet9D.py")):
        return None
    if PYREF not in sys.path:
        sys.path.insert(0, PYREF)
    import importlib
    importlib.import_module("config.config")
    from absl import flags
    if not flags.FLAGS.is_parsed():
        flags.FLAGS(["tgpose_ref"])
    gcn3d = importlib.import_module("network.fs_net_repo.gcn3d")
    face = importlib.import_module("network.fs_net_repo.FaceRecon")
    pose = importlib.import_module("network.fs_net_repo.PoseNet9D")
    return gcn3d, face, pose


if __name__ == "__main__":
    print(json.dumps(build(verbose="-v" in sys.argv), indent=1))

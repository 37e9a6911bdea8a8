"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by ``deepsvc_b200``.

Stages the three UNMODIFIED reference sources the hot path lives in
(``modules.py``, ``image_model.py``, ``video_model.py``) from ``/root/reference`` into
``oracle/_ref/`` so that the GPU box -- where ``/root/reference`` does not exist -- can run
the reference's own ``DeepSVC`` both stock (torch CUDA + ``oracle/shim``) and through
``deepsvc_b200.patch_reference()`` (``tests/test_gpu_reference_dropin.py``, ``bench.py
--workload dropin``).  ``oracle/_ref/`` is git-ignored (reference sources are never
committed) but not gpurun-ignored, so it travels with the snapshot like the built ``.so``.

The copies are byte-identical; their sha256 is written next to them so that a test can show
which reference revision it ran.  The reference is pure Python: there is nothing to compile.
"""
import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference"
REF_DST = os.path.join(HERE, "_ref")
FILES = ("modules.py", "image_model.py", "video_model.py")


def stage(src: str = REF_SRC, dst: str = REF_DST) -> bool:
    """Copy the files if the reference is mounted; returns whether ``dst`` is usable."""
    if os.path.isdir(src) and all(os.path.isfile(os.path.join(src, f)) for f in FILES):
        os.makedirs(dst, exist_ok=True)
        manifest = {}
        for f in FILES:
            shutil.copyfile(os.path.join(src, f), os.path.join(dst, f))
            with open(os.path.join(dst, f), "rb") as fh:
                manifest[f] = hashlib.sha256(fh.read()).hexdigest()
        with open(os.path.join(dst, "MANIFEST.json"), "w") as fh:
            json.dump({"source": src, "sha256": manifest}, fh, indent=1)
    return available(dst)


def available(dst: str = REF_DST) -> bool:
    return all(os.path.isfile(os.path.join(dst, f)) for f in FILES)


def import_reference(dst: str = REF_DST):
    """(modules, image_model, video_model) of the staged, unmodified reference with the
    oracle's compressai / pytorch_msssim shims on ``sys.path``.  Raises if not staged."""
    import sys
    if not available(dst):
        raise FileNotFoundError(f"reference not staged under {dst} (run oracle/stage_reference.py "
                                "in the build container)")
    shim = os.path.join(HERE, "shim")
    for p in (dst, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    import modules
    import image_model
    import video_model
    return modules, image_model, video_model


if __name__ == "__main__":
    print("staged" if stage() else "reference not available")

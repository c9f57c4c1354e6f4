"""Stage the reference's own CPU implementation of the calibration path under oracle/_ref/ (TEST INFRASTRUCTURE).

    python -m oracle.stage_ref

quantize/distribution_calibrate.py of the reference imports only numpy and tqdm, so it runs verbatim
wherever NumPy does.  /root/reference does not exist on the GPU box; oracle/_ref/ is git-ignored (reference sources
never enter the history) but not gpurun-ignored, so the staged copy travels with the tree like a built .so and
``bench.py --impl reference`` / ``cpu_baseline`` can time the reference itself (kind "reference") instead of the
oracle port.  The file is copied byte for byte; its SHA-256 is written next to it.  Nothing under
quantization/ ever reads it.
"""
import hashlib
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/quantize/distribution_calibrate.py"
DST_DIR = os.path.join(HERE, "_ref")
DST = os.path.join(DST_DIR, "distribution_calibrate.py")


def stage():
    """Copy when the reference is mounted (the build container); keep whatever is staged otherwise.  Returns the
    staged path or None."""
    if os.path.exists(SRC):
        os.makedirs(DST_DIR, exist_ok=True)
        shutil.copyfile(SRC, DST)
        with open(DST, "rb") as f:
            digest = hashlib.sha256(f.read()).hexdigest()
        with open(os.path.join(DST_DIR, "SHA256"), "w") as f:
            f.write("%s  distribution_calibrate.py  (copied from %s)\n" % (digest, SRC))
    return DST if os.path.exists(DST) else None


if __name__ == "__main__":
    print(stage())

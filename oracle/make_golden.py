#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING THE REFERENCE ITSELF.

TEST INFRASTRUCTURE ONLY.  Runs only in the build container, where
/root/reference exists; the GPU box uses the committed fixtures.

The reference's quantize/distribution_calibrate.py imports nothing but numpy and
tqdm, so it is loaded verbatim by file path (a package import would pull in
MXNet through quantize/__init__.py:3).  Its scalar math therefore runs under this
container's NumPy (2.x => NEP 50 promotion); the fixtures are labelled with the
regime.

    python -m oracle.make_golden
"""
import importlib.util
import os
import sys
import time

import numpy as np

from . import golden_recipes as R

REF = "/root/reference/quantize/distribution_calibrate.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def load_reference():
    spec = importlib.util.spec_from_file_location("ref_distribution_calibrate", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    ref = load_reference()
    os.makedirs(OUT, exist_ok=True)
    regime = "nep50" if int(np.__version__.split(".")[0]) >= 2 else "legacy"

    # ---- histograms: _discrete_histogram + the accumulate logic of collect_feature_maps:95-104
    out = {"regime": np.array(regime), "numpy": np.array(np.__version__)}
    for name, batches in R.hist_cases().items():
        fm_max = None
        hist = 0
        err = ""
        for b, fm in enumerate(batches):
            h, m = ref._discrete_histogram(fm, R.BINS, fm_max)
            out["hist/%s/batch%d" % (name, b)] = h
            if fm_max is None:
                fm_max = m
            try:
                hist = hist + h
            except ValueError as e:       # (2048,) + (2049,) -- reference quirk
                err = type(e).__name__
                break
        out["hist/%s/max" % name] = np.float32(fm_max)
        out["hist/%s/error" % name] = np.array(err)
        if not err:
            out["hist/%s/acc" % name] = hist
        print("hist", name, "len", len(h), "max", fm_max, err)
    np.savez_compressed(os.path.join(OUT, "hist_%s.npz" % regime), **out)

    # ---- KL search
    out = {"regime": np.array(regime), "numpy": np.array(np.__version__)}
    hists = R.kl_hist_cases()
    for name, h in hists.items():
        out["kl/%s/hist" % name] = h
        for levels in R.KL_LEVELS[name]:
            t0 = time.time()
            best = ref.kl_calibrate(h, levels=levels, min_bins=levels, bins=R.BINS)
            out["kl/%s/L%d/best" % (name, levels)] = np.int32(best)
            print("kl", name, levels, "->", best, "%.1fs" % (time.time() - t0))
        levels = R.KL_LEVELS[name][0]
        for lo, hi in R.KL_WINDOWS:
            best = ref.kl_calibrate(h, levels=levels, min_bins=max(lo, levels), bins=hi)
            out["kl/%s/L%d/win_%d_%d" % (name, levels, lo, hi)] = np.int32(best)
    np.savez_compressed(os.path.join(OUT, "kl_%s.npz" % regime), **out)
    print("wrote", OUT)


if __name__ == "__main__":
    sys.exit(main())

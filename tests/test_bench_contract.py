"""bench.py contract pieces that run without a GPU: the reference arm (the oracle port of the reference's CPU
path on the host cores) prints ONE JSON line with the keys the driver reads, and the workload bookkeeping of
the GPU arm (layer shapes, element counts) is consistent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    env = dict(os.environ, OMP_NUM_THREADS="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0"], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "kl_calibration_images_per_sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["steps"] == 1 and d["warmup"] == 0 and d["n_gpus"] == 1
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("mobilenet1.0 KL calibration")
    cb = d["cpu_baseline"]
    # the reference's own file when build() has staged it under oracle/_ref/, else the oracle port
    staged = os.path.exists(os.path.join(ROOT, "oracle", "_ref", "distribution_calibrate.py"))
    assert cb["kind"] == ("reference" if staged else "port")
    assert cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    # like for like with the GPU arm: same config block (no extra keys), one KL search amortised over K batches
    sys.path.insert(0, ROOT)
    import bench
    assert d["config"] == bench.workload_config(1) and "amortised over K=1 batches of 128" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0


def test_reference_arm_is_silent_on_other_ranks():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env, cwd=ROOT)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_workload_bookkeeping():
    sys.path.insert(0, ROOT)
    import bench
    shapes = bench.layer_shapes()
    assert len(shapes) == bench.N_LAYERS == 27
    assert sum(c * h * w for c, h, w in shapes) == bench.ELEMS_PER_IMAGE == 4_993_536
    cfg = bench.workload_config(4)
    assert cfg["elements_per_step_per_gpu"] == 639_172_608 and "4 GPU(s)" in cfg["parallelism"]
    assert str(bench.RING) in cfg["parallelism"]

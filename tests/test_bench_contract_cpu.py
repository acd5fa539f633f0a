"""bench.py contract that can be checked without a GPU: the reference arm (the reference's own CPU path: the unmodified
modules from baseline/_ref; the numpy oracle port only when that copy is absent) prints ONE JSON line with the agreed keys,
also under torchrun's OMP_NUM_THREADS=1, and the GPU arm refuses to run silently on CPU."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    env = dict(os.environ, OMP_NUM_THREADS="1")       # what torch.distributed.run exports to every rank
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-batch", "1"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "vit_b16_224_train_images_per_sec" and d["unit"] == "images/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["dtype"] == "f32"
    sys.path.insert(0, ROOT)
    from baseline import loader
    assert d["cpu_baseline"]["kind"] == ("reference" if loader.available() else "port")
    assert d["cpu_baseline"]["cores"] == (os.cpu_count() if loader.available() else d["cpu_baseline"]["cores"])   # not 1 under torchrun
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_train_flops_formula():
    sys.path.insert(0, ROOT)
    import bench
    # ViT-B/16 224 -> 96.786 GFLOP per trained image
    assert abs(bench.train_flops_per_image() / 1e9 - 96.786) < 0.01
    t = bench.load_gemm_traffic()
    assert t is None or t > 1e6


def test_workload_flop_formulas_and_descriptors():
    """bench_workloads: the algorithmic work per unit follows SURVEY.md §8(d) (ViT-L/16: 339.274 GFLOP per trained image) and every
    workload name the parser accepts has a descriptor (no GPU needed: descriptors are built lazily, models are not)."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "vit-is-all-you-need_b200"))
    import torch

    import bench_workloads as W
    from b200vit import modules as M
    w = W._workload("vit_l", torch, M, None)
    assert abs(w["flops_per_unit"] / 1e9 - 339.274) < 0.01 and w["batch"] == 256 and w["reference"] is None
    for name in ("vit_ti", "titok_s", "tatitok_s", "videogpt_b"):
        d = W._workload(name, torch, M, None)
        assert d["flops_per_unit"] > 0 and d["unit"] and d["metric"] and callable(d["ours"]) and callable(d["inputs"])
    x = W._workload("videogpt_b", torch, M, None)["inputs"](2, torch.Generator().manual_seed(0))
    assert x[0].shape == (2, 16, 64) and x[0].dtype == torch.int64

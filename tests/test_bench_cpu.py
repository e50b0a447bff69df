"""The reference arm of bench.py runs on CPU only: check its JSON line against the contract keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--points-per-gpu", "300000"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "Mpoints/s" and line["higher_is_better"] is True
    assert line["metric"] == "RCB Mpoints/s (3D f64, 2^10 parts)" and line["scaling"] == "weak"
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "Mpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and line["vs_baseline"] is None
    assert line["config"]["name"] == "C4" and line["config"]["points_total"] == 300000


def test_reference_arm_other_configs():
    for cfg, n in (("C2", "200000"), ("C3", "400000"), ("C5", "100000"), ("C1", "50000")):
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                              "--warmup", "0", "--config", cfg, "--points-per-gpu", n, "--scaling", "strong"],
                             capture_output=True, text=True, timeout=600, cwd=ROOT)
        assert out.returncode == 0, out.stderr
        line = json.loads(out.stdout.strip().splitlines()[-1])
        assert line["config"]["name"] == cfg and line["scaling"] == "strong" and line["value"] > 0


def test_generator_does_not_depend_on_the_rank_count():
    import torch

    sys.path.insert(0, ROOT)
    import bench

    dev = torch.device("cpu")
    for name in ("C2", "C4", "C5"):
        n = bench.CHUNK + 12_345
        whole_p, whole_w = bench.gen_range(torch, name, 0, n, dev)
        for world in (2, 3):
            ps, ws = [], []
            for r in range(world):
                b, e = bench.shard_of(n, r, world)
                p, w = bench.gen_range(torch, name, b, e, dev)
                ps.append(p)
                ws.append(w)
            assert torch.equal(torch.cat(ps), whole_p)
            if not isinstance(whole_w, float):
                assert torch.equal(torch.cat(ws), whole_w)


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""

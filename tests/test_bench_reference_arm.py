"""CPU check of `bench.py --impl reference`: the arm the driver runs beside the GPU one must print ONE JSON line with
the contract's keys, from rank 0 only, without touching a device (it times oracle/oracle.c), on the SAME `config` dict
as the GPU arm, and with every host core whatever OMP_NUM_THREADS says (torchrun exports OMP_NUM_THREADS=1)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _run(env_extra=None, extra=()):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")
    env.update(env_extra or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                           "--ref-vars", "14", *extra], env=env, capture_output=True, text=True, timeout=300, cwd=ROOT)


def test_reference_arm_prints_the_contract_line():
    import bench

    out = _run({"OMP_NUM_THREADS": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sumcheck_prover_throughput" and d["unit"] == "Melem/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] == 1 and d["warmup"] == 1
    cores = len(os.sched_getaffinity(0))
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] == cores == d["host_threads"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["sample_vars"] == 14
    assert d["e2e"] == {"value": d["value"], "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the config is the GPU arm's, key for key
    assert d["config"] == bench.config_dict(28, 1, bench.MODULUS, 1, 0)
    assert "2^14-entry tables" in d["cpu_baseline"]["sample"]


def test_reference_arm_config_follows_n_gpus():
    import bench

    out = _run({"RANK": "0", "WORLD_SIZE": "4", "LOCAL_RANK": "0", "OMP_NUM_THREADS": "1"}, extra=("--gpus", "4"))
    assert out.returncode == 0, out.stderr[-2000:]
    d = json.loads(out.stdout.strip())
    assert d["n_gpus"] == 4 and d["config"] == bench.config_dict(28, 4, bench.MODULUS, 1, 0)
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))  # not 1


def test_reference_arm_other_ranks_stay_silent():
    out = _run({"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == ""

"""CPU tests of bench.py's host logic, including the N > 1 path on gloo with world_size 2
(replicas: max-over-ranks time, sum-over-ranks units; no data-path collective)."""
import json
import os
import subprocess
import sys

import pytest

import bench

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys, json
sys.path.insert(0, {root!r})
import torch, torch.distributed as dist
import bench
dist.init_process_group("gloo")
rank, local_rank, world = bench.dist_env()
assert world == 2 and rank == dist.get_rank()
# rank 1 is the slow one: the job's time is ITS time, the job's units are the sum
ms, units = bench.reduce_over_ranks(100.0 + 50.0 * rank, 5, world)
dist.barrier()
if rank == 0:
    print(json.dumps({{"ms": ms, "units": units}}))
dist.destroy_process_group()
"""


def test_reduce_over_ranks_gloo_world2(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                         capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    res = json.loads(line)
    assert res == {"ms": 150.0, "units": 10.0}


def test_reduce_single_rank_is_identity():
    assert bench.reduce_over_ranks(12.5, 3, 1) == (12.5, 3)


def test_algorithmic_bytes_match_survey():
    # MSM 2^20 BN254 = 100 663 296 B, BLS12-381 = 134 217 728 B, 2^17 BN254 = 12 582 912 B
    assert bench.msm_algorithmic_bytes(1 << 20, "BN254") == 100663296
    assert bench.msm_algorithmic_bytes(1 << 20, "BLS12_381") == 134217728
    assert bench.msm_algorithmic_bytes(1 << 17, "BN254") == 12582912


def test_peaks_come_from_measured_file():
    peak, src = bench.peaks()
    assert peak > 1000 and ("measured" in src or "fallback" in src)


def test_reference_arm_prints_contract_line():
    """--impl reference at a tiny size: one JSON line with the contract's keys, CPU only."""
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--log2", "8",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == bench.METRIC and line["unit"] == bench.UNIT
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"] > 0
    assert line["steps"] == 2 and line["warmup"] == 1 and line["higher_is_better"] is True


def test_reference_arm_uses_all_host_cores_under_torchrun_env():
    """torch.distributed.run exports OMP_NUM_THREADS=1 to every rank; the CPU arm must ignore it (round 1's N > 1
    reference lines ran on ONE core), prove the workload it names at full size and never scale a smaller one."""
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", LOCAL_RANK="0", WORLD_SIZE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--log2", "9",
                          "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["cpu_baseline"]["cores"] == bench.host_cores()
    assert "scaled x" not in line["cpu_baseline"]["sample"] and "FULL 2^9-row" in line["cpu_baseline"]["sample"]
    assert abs(line["ms_per_step"] * line["value"] - 1e3) < 1e-6


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                          "--log2", "8", "--steps", "1", "--warmup", "0"], capture_output=True, text=True,
                         timeout=600, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_reference_step_plan_never_extrapolates():
    # fits: run exactly what was asked
    assert bench.plan_reference_steps(8.5, 20, 5, 240.0) == (5, 20)
    # does not fit: one warm-up proof, as many timed proofs as the budget holds, at least one
    assert bench.plan_reference_steps(40.0, 20, 5, 240.0) == (1, 5)
    assert bench.plan_reference_steps(400.0, 20, 5, 240.0) == (1, 1)
    assert bench.plan_reference_steps(40.0, 3, 0, 240.0) == (0, 3)


def test_b200_arm_fails_loudly_without_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("GPU present")
    except ImportError:
        pass
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--log2", "4", "--steps", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode != 0
    assert "{" not in out.stdout        # no JSON line from a CPU fallback

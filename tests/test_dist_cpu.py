"""N > 1 host logic on CPU (gloo, world size 2): pose sharding, the single all-reduce of the metric sums, the
max-over-ranks timing rule, and the `--impl reference` arm under torchrun (rank 0 alone works and prints)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from openmpl_b200 import dist as mdist, synth
from oracle import mpl_oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _torchrun(nproc, script_args, timeout=300):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port())] + script_args
    env = dict(os.environ, OMP_NUM_THREADS="2")
    return subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=timeout)


@pytest.mark.parametrize("total,world", [(10, 1), (10, 3), (7, 8), (65536 * 8, 8), (0, 2)])
def test_shard_range_partitions_the_global_pose_index(total, world):
    spans = [mdist.shard_range(total, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == total
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [b - a for a, b in spans]
    assert max(sizes) - min(sizes) <= 1


def test_generator_is_keyed_by_global_pose_index():
    rig = synth.make_rig(4)
    whole = synth.make_batch(12, rig, seed=5)
    part = synth.make_batch(5, rig, seed=5, start=7)
    for k in ("poses", "rays", "centers", "target"):
        np.testing.assert_array_equal(whole[k][7:], part[k])


def test_two_rank_all_reduce_equals_single_process_metric(tmp_path):
    total = 37                                     # odd: the two shards differ in size
    out = tmp_path / "dist.json"
    r = _torchrun(2, [os.path.join(ROOT, "tests", "dist_worker.py"), str(out), str(total)])
    assert r.returncode == 0, r.stdout + r.stderr
    d = json.load(open(out))
    assert d["world"] == 2 and d["n"] == total and d["slowest"] == 2.0
    assert d["shards"] == [[0, 19], [19, 37]]
    pred = np.asarray(d["pred"], dtype=np.float32)
    gt = synth.make_batch(total, synth.make_rig(4), seed=5)["target"]
    ref_abs = mpl_oracle.evaluate(pred, gt, output_in_meter=True, relative=False)
    ref_rel = mpl_oracle.evaluate(pred, gt, output_in_meter=True, relative=True)
    np.testing.assert_allclose(d["pjpe_abs"], ref_abs["pjpe"], rtol=1e-12)
    assert abs(d["mpjpe_abs"] - ref_abs["mpjpe"]) < 1e-10
    assert abs(d["mpjpe_rel"] - ref_rel["mpjpe"]) < 1e-10
    whole = mpl_oracle.pmpjpe_sums(pred, gt)                  # Procrustes-aligned sums reduce the same way
    assert abs(d["p_mpjpe"] - whole[:17].mean() / total) < 1e-10


def test_reference_arm_under_torchrun_prints_one_line_from_rank0():
    r = _torchrun(2, [os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1",
                      "--depth", "1", "--cpu-batch", "32"])
    assert r.returncode == 0, r.stdout + r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "poses/s" and d["value"] > 0
    # "reference": the unmodified module (checkout mounted, or staged under oracle/_ref by build()); "port" only without it
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["e2e"]["h2d_bytes_per_step"] == 0


def test_reference_gpu_comparator_says_unavailable_without_a_gpu():
    """`bench.py --impl reference-gpu` (the unmodified reference module as PyTorch-eager kernels on the B200, SURVEY §8d's second
    comparator) needs a GPU: on a CPU-only host it prints one JSON line saying so and exits 0."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the comparator would run")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference-gpu"], capture_output=True, text=True,
                       timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert d["impl"] == "reference-gpu" and "unavailable" in d

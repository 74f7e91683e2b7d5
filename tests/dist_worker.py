"""Worker of tests/test_dist_cpu.py: one process per rank (gloo, CPU).  Each rank owns a contiguous shard of the global
pose range, fills the metric accumulator for its shard (sums computed by the oracle — no GPU here), all-reduces once,
and rank 0 writes what the N-rank job reports."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from openmpl_b200 import dist as mdist, metric, synth  # noqa: E402
from oracle import mpl_oracle  # noqa: E402  (checker: supplies the per-shard sums the CUDA kernel would produce)


def main():
    out_path, total = sys.argv[1], int(sys.argv[2])
    rank, world, _ = mdist.init_from_env("gloo")
    start, stop = mdist.shard_range(total, rank, world)
    # the generator is keyed by the global pose index: any sharding sees the same data
    batch = synth.make_batch(stop - start, synth.make_rig(4), seed=5, start=start)
    rng = np.random.default_rng(100 + rank)
    pred = (batch["target"] + 0.01 * rng.standard_normal(batch["target"].shape)).astype(np.float32)
    acc = metric.MpjpeAccumulator(17, output_in_meter=True, device="cpu")
    acc.acc += torch.from_numpy(mpl_oracle.metric_sums(pred, batch["target"]))
    acc.all_reduce()
    pacc = metric.PmpjpeAccumulator(17, output_in_meter=True, device="cpu")
    pacc.acc += torch.from_numpy(mpl_oracle.pmpjpe_sums(pred, batch["target"]))
    pacc.all_reduce()
    slowest = mdist.max_over_ranks(float(rank + 1))
    gathered = [None] * world
    dist.all_gather_object(gathered, (start, stop, pred.tolist()))
    if rank == 0:
        res = acc.result()
        json.dump({"world": world, "n": res["n"], "mpjpe_abs": res["mpjpe_abs"], "mpjpe_rel": res["mpjpe_rel"],
                   "pjpe_abs": res["pjpe_abs"].tolist(), "slowest": slowest, "p_mpjpe": pacc.result()["p_mpjpe"],
                   "shards": [(g[0], g[1]) for g in gathered],
                   "pred": [p for g in gathered for p in g[2]]}, open(out_path, "w"))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

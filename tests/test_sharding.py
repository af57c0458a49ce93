"""N>1 host logic on CPU: 2-rank gloo run of the sharding / gather / reduction helpers."""
import os
import subprocess
import sys

import numpy as np

from _harness import ROOT
from simbody_b200.sharding import shard_range

WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from simbody_b200.sharding import shard_range, gather_final_states, reduce_stats
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
N, ny = 1001, 7
lo, hi = shard_range(N, rank, world)
full = np.arange(ny * N, dtype=np.float64).reshape(ny, N)
got = gather_final_states(full[:, lo:hi], N, dist)
assert got.shape == (ny, N) and np.array_equal(got, full), "gather mismatch"
mx, bad, ms = reduce_stats(0.5 + rank, rank + 1, 10.0 * (rank + 1), dist)
assert (mx, bad, ms) == (0.5 + world - 1, world * (world + 1) // 2, 10.0 * world), (mx, bad, ms)
dist.barrier()
if rank == 0:
    print("SHARDING_OK", lo, hi)
dist.destroy_process_group()
'''


def test_shard_ranges_cover_the_batch():
    for n, w in [(1, 1), (7, 2), (1048576, 8), (1001, 8), (5, 8)]:
        ranges = [shard_range(n, r, w) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        sizes = [hi - lo for lo, hi in ranges]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_gather_and_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script), ROOT],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "SHARDING_OK" in r.stdout

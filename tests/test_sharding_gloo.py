"""CPU test of the multi-GPU host logic (world_size 2, gloo): contiguous guide shards cover the guide list exactly
once, and the timing reductions bench.py uses (max / sum over ranks, barrier) work across processes."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

from conftest import ROOT


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


@pytest.mark.timeout(300)
def test_two_rank_sharding_and_reductions(tmp_path):
    script = os.path.join(tmp_path, "w.py")
    open(script, "w").write(textwrap.dedent("""
        import os, sys, json
        sys.path.insert(0, %r)
        import torch, torch.distributed as dist
        import bench
        dist.init_process_group(backend="gloo")
        rank, world = dist.get_rank(), dist.get_world_size()
        n = 1001
        lo, hi = bench.shard_guides(n, rank, world)
        mine = torch.zeros(n, dtype=torch.int64); mine[lo:hi] = 1
        dist.all_reduce(mine)
        assert int(mine.min()) == 1 and int(mine.max()) == 1, "shards must partition the guides"
        bench.barrier_sync(dist, 0)
        mx = bench.reduce_max(dist, 0, float(rank + 1))
        sm = bench.reduce_sum(dist, 0, float(hi - lo))
        assert mx == float(world) and sm == float(n)
        if rank == 0:
            print(json.dumps({"ok": True, "world": world}))
        dist.destroy_process_group()
    """ % ROOT))
    port = _free_port()
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), script], capture_output=True, text=True, timeout=280)
    assert r.returncode == 0, r.stderr[-2000:]
    assert '"ok": true' in r.stdout

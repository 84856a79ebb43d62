"""The N>1 path on CPU: world_size-2 gloo run of the contig sharding + filter broadcast + ordered merge (shard.py)."""
import os
import re
import socket
import subprocess
import sys

from ntedit_b200 import shard

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_lpt_assignment_balances_and_is_deterministic():
    lens = [250, 240, 200, 180, 150, 100, 90, 50, 50, 10] + [1] * 40
    for world in (1, 2, 4, 8):
        owner = shard.assign_contigs(lens, world)
        assert owner == shard.assign_contigs(lens, world)
        loads = [sum(lens[i] for i in shard.my_contigs(owner, r)) for r in range(world)]
        assert sum(loads) == sum(lens)
        assert max(loads) - min(loads) <= max(lens)
        assert sorted(i for r in range(world) for i in shard.my_contigs(owner, r)) == list(range(len(lens)))


def test_merge_keeps_input_order_and_drops_missing():
    parts = [{0: (b"a", b"A", b""), 3: (b"d", b"D", b"v")}, {1: (b"b", b"B", b"")}]
    assert shard.merge_in_input_order(4, parts) == (b"abd", b"ABD", b"v")


def test_two_rank_gloo_run_equals_single_process():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    env = dict(os.environ, OMP_NUM_THREADS="1")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "dist_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, timeout=600, env=env, cwd=ROOT)
    out = r.stdout.decode(errors="replace")
    assert r.returncode == 0, out[-3000:]
    m = re.search(r"SHARD_RESULT ok=(\d) loads=\[(\d+), (\d+)\] edits=(\d+)", out)
    assert m, out[-3000:]
    assert m.group(1) == "1"
    assert int(m.group(2)) > 0 and int(m.group(3)) > 0 and int(m.group(4)) > 10

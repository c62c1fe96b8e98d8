"""Host side of the multi-GPU path on CPU: two processes, gloo backend.  The device side (NCCL halo
exchange) is covered on GPUs by tests/mgpu_check.py under torchrun."""
import json
import os
import socket
import subprocess
import sys

import numpy as np

from pyapi import demb200, scenes

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_two_rank_partition_over_gloo(built, tmp_path):
    port = _free_port()
    out = str(tmp_path / "result")
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE="2", MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), GLOO_SOCKET_IFNAME="lo")
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, "gloo_worker.py"), out], env=env,
                                      stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    logs = []
    for p in procs:
        try:
            o, _ = p.communicate(timeout=240)
        except subprocess.TimeoutExpired:
            p.kill()
            o, _ = p.communicate()
        logs.append(o)
    assert all(p.returncode == 0 for p in procs), "\n".join(logs)
    res = [json.load(open(out + ".%d" % r)) for r in range(2)]
    # slabs tile the axis: rank 0's upper cut is rank 1's lower cut, the ends are open
    assert res[0]["hi"] == res[1]["lo"] and res[0]["lo"] < -1e37 and res[1]["hi"] > 1e37
    counts = np.asarray(res[0]["counts"])
    assert np.array_equal(counts, np.asarray(res[1]["counts"]))          # both ranks gathered the same table
    assert counts[:, 0].sum() == res[0]["n"]                               # owners are partitioned
    assert counts[0, 3] == counts[1, 1] and counts[1, 2] == counts[0, 1]  # sent right == neighbour's ghosts, and back
    assert counts[0, 2] == 0 and counts[1, 3] == 0                         # nothing leaves through the outer faces
    assert counts[0, 3] > 0 and counts[1, 2] > 0


def test_slab_bounds_and_partition_single_process(built):
    f = scenes.flatten(scenes.config2_clumps(16, 4, 3, spacing=2.7))
    p = demb200.params_from_flat(f)
    for world in (1, 2, 4, 8):
        cuts = [demb200.host_slab_bounds(p, world, r) for r in range(world)]
        assert cuts[0][0] < -1e37 and cuts[-1][1] > 1e37
        for r in range(world - 1):
            assert cuts[r][1] == cuts[r + 1][0]
        own = np.zeros(f.nClumps, "i4")
        for r in range(world):
            role, send = demb200.host_partition_owners(p, world, r, 0.01, f.voxelID[: f.nClumps], f.locX[: f.nClumps])
            own += (role == 1)
            assert not (send[role != 1]).any()
            if r == 0:
                assert not (send & 1).any()
            if r == world - 1:
                assert not (send & 2).any()
        assert (own == 1).all()

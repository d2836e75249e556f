"""N>1 host logic on the CPU: two real processes (torch.distributed, gloo, 127.0.0.1) each build their halo plan
through the C ABI's host-only entry points, exchange packed boundary values exactly as Halo::begin does on the
device (send lists in plan order, receives into contiguous ranges of the sorted ghost list), and every ghost value
must equal its owner's value.  No GPU involved."""
import os
import socket
import subprocess
import sys
import textwrap

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = textwrap.dedent('''
    import os, sys
    import numpy as np
    import torch.distributed as dist
    sys.path.insert(0, %r)
    import saddle_point_petsc_b200 as sp
    dist.init_process_group("gloo", init_method="env://")
    rank, size = dist.get_rank(), dist.get_world_size()
    for (M, N, dof) in ((9, 7, 2), (12, 12, 1), (5, 30, 2)):
        xs, ys, xm, ym = sp.dmda_corners(M, N, size, rank)
        plan = sp.dmda_halo_plan(M, N, size, rank)
        # owned field: value of dof c at node = 10 * global_petsc_node + c  (known from the id alone)
        own = np.array([[10.0 * sp.dmda_global_node(M, N, size, xs + i, ys + j)[0] + c for c in range(dof)]
                        for j in range(ym) for i in range(xm)])
        send = {}
        for q in sorted(set(plan["send_rank"].tolist())):
            send[q] = own[plan["send_lnode"][plan["send_rank"] == q]].copy()      # pack: plan order
        box = [None] * size
        dist.all_gather_object(box, send)                                         # the "wire"
        ghost = np.full((len(plan["ghost_gnode"]), dof), np.nan)
        for q in range(size):
            if q == rank or rank not in box[q]:
                continue
            sel = np.flatnonzero(plan["ghost_owner"] == q)
            assert len(sel) == len(box[q][rank]) and np.all(np.diff(sel) == 1)    # one contiguous range per owner
            ghost[sel] = box[q][rank]
        want = np.array([[10.0 * g + c for c in range(dof)] for g in plan["ghost_gnode"]])
        assert np.array_equal(ghost, want), (rank, M, N)
        # every rank reports its box; together they tile the grid exactly once
        boxes = [None] * size
        dist.all_gather_object(boxes, (xs, ys, xm, ym))
        cover = np.zeros((N, M), dtype=int)
        for (a, b, c, d) in boxes:
            cover[b:b + d, a:a + c] += 1
        assert np.all(cover == 1)
    dist.barrier()
    print("rank", rank, "ok")
''')


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4])
def test_halo_plan_across_real_processes(tmp_path, world):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(r), WORLD_SIZE=str(world), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and ("rank %d ok" % r) in o, o

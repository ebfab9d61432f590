"""CPU test of the multi-GPU host logic with 2 gloo ranks: j-slab decomposition + halo exchange plan,
checked by running the oracle per slab and comparing with the oracle on the undivided domain."""
import json
import os
import subprocess
import sys

import numpy as np

from util import HALO, ROOT
from dawn_b200 import slabs

WORKER = r'''
import os, sys, json
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.environ["REPO"]); sys.path.insert(0, os.path.join(os.environ["REPO"], "tests"))
from util import HALO, hori_diff_type2_inputs, run_oracle
from dawn_b200 import slabs
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
I, J, K = 20, 22, 4
f = hori_diff_type2_inputs(I, J, K)                      # global fields (every rank builds them)
dom = (I + 2 * HALO, J + 2 * HALO, K)
want = run_oracle("hori_diff_type2_stencil", dom, f)["out"]
sl = slabs.decompose(J, world, rank, HALO)
rows = slice(sl.j0, sl.j0 + sl.nj_local + 2 * HALO)      # local storage rows in global numbering
loc = {k: (v[:, rows, :].copy() if v.shape[1] > 1 else v.copy()) for k, v in f.items()}
# wipe the halo rows that belong to a neighbour, then exchange them back in
info = json.load(open(os.path.join(os.environ["REPO"], "tests", "golden", "hd2_fields.json")))
widths = slabs.exchange_widths(info["fields"], set(info["written"]))
assert widths == {"in": (2, 2)}, widths
for name, (wm, wp) in widths.items():
    t = torch.from_numpy(loc[name])
    if sl.has_lower: t[:, :HALO, :] = float("nan")
    if sl.has_upper: t[:, -HALO:, :] = float("nan")
    slabs.exchange_torch(t, sl, wm, wp, dist)
got = run_oracle("hori_diff_type2_stencil", (dom[0], sl.nj_local + 2 * HALO, K), loc)["out"]
ok = np.array_equal(got[:, HALO:-HALO, :], want[:, HALO + sl.j0:HALO + sl.j0 + sl.nj_local, :], equal_nan=True)
res = torch.tensor([1 if ok else 0]); dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0: print("SLABS_OK" if int(res) == 1 else "SLABS_MISMATCH")
dist.destroy_process_group()
'''


def test_decompose_covers_domain():
    for J, P in ((1024, 8), (10, 3), (7, 7), (4096, 4)):
        parts = [slabs.decompose(J, P, r) for r in range(P)]
        assert parts[0].j0 == 0 and sum(p.nj_local for p in parts) == J
        for a, b in zip(parts, parts[1:]):
            assert a.j0 + a.nj_local == b.j0
        assert max(p.nj_local for p in parts) - min(p.nj_local for p in parts) <= 1


def test_row_ranges_mirror_runtime_plan():
    sl = slabs.decompose(64, 4, 1)
    r = slabs.row_ranges(sl, sl.nj_local + 2 * HALO, 2, 1)
    assert r["send_down"] == (3, 4) and r["recv_low"] == (1, 3)
    assert r["send_up"] == (17, 19) and r["recv_high"] == (19, 20)


def test_two_rank_halo_exchange_reproduces_global_result(tmp_path):
    fields = [{"name": "out", "dims": "ijk", "extent": [0] * 6},
              {"name": "in", "dims": "ijk", "extent": [-2, 2, -2, 2, 0, 0]},
              {"name": "crlato", "dims": "j", "extent": [-1, 1, -1, 1, 0, 0]},
              {"name": "crlatu", "dims": "j", "extent": [-1, 1, -1, 1, 0, 0]},
              {"name": "hdmask", "dims": "ijk", "extent": [0] * 6}]
    with open(os.path.join(ROOT, "tests", "golden", "hd2_fields.json"), "w") as f:
        json.dump({"fields": fields, "written": ["out"]}, f)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, REPO=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                        "--nproc-per-node=2", "--master-addr", "127.0.0.1", "--master-port", "29611",
                        str(script)], capture_output=True, text=True, env=env, timeout=300)
    assert "SLABS_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]

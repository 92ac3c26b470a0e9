"""CPU tier: the N>1 path (GOP sharding, no data-path collective) with world_size 2 over gloo."""
import hashlib
import io
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import hashlib, io, json, os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
import numpy as np
import torch.distributed as dist
import hwang_b200 as hw
from hwang_b200 import _lib, build, shard
import hwb_testutil as util
_lib.use_library(build.EMU)
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
kw = dict(width=96, height=80, frames=48, gop=6, profile=1, bframes=1, seed=88, qp=30)
mp4, index, samples, kf = util.make_clip(**kw)
rows = list(range(0, 48, 5)) + [46, 47]
items = shard.gop_work_items(index, 0, rows)
mine = shard.merge_adjacent(shard.partition(items, world)[rank])
dec = hw.Decoder(io.BytesIO(mp4), video_index=index)
got = {}
for (_, a, b, cost, want) in mine:
    for r, f in zip(want, dec.retrieve(want)):
        got[r] = hashlib.md5(np.asarray(f).tobytes()).hexdigest()
allgot = [None] * world
dist.all_gather_object(allgot, got)   # control plane only: checking the result, not part of the decode path
if rank == 0:
    merged = {}
    for g in allgot:
        assert not (set(g) & set(merged)), 'a row was decoded by two ranks'
        merged.update(g)
    print('RESULT ' + json.dumps({'rows': sorted(merged), 'md5': [merged[r] for r in sorted(merged)], 'per_rank': [len(g) for g in allgot]}))
dist.destroy_process_group()
'''


def test_gop_sharding_world_size_2(emu, built, tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=2', '--master-addr', '127.0.0.1',
                          '--master-port', '29653', str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith('RESULT ')][0]
    res = json.loads(line[7:])
    rows = sorted(set(list(range(0, 48, 5)) + [46, 47]))
    assert res['rows'] == rows
    assert all(n > 0 for n in res['per_rank'])
    # single-process result for the same rows
    import hwang_b200 as hw
    import hwb_testutil as util
    kw = dict(width=96, height=80, frames=48, gop=6, profile=1, bframes=1, seed=88, qp=30)
    mp4, index, samples, kf = util.make_clip(**kw)
    frames = hw.Decoder(io.BytesIO(mp4), video_index=index).retrieve(rows)
    assert [hashlib.md5(np.asarray(f).tobytes()).hexdigest() for f in frames] == res['md5']


def test_partition_properties():
    from hwang_b200 import shard
    rng = np.random.default_rng(1)
    for _ in range(50):
        n = int(rng.integers(1, 60))
        items = [(int(rng.integers(0, 4)), i * 10, i * 10 + 10, int(rng.integers(1, 1000)), [i * 10]) for i in range(n)]
        for w in (1, 2, 4, 8):
            parts = shard.partition(items, w)
            flat = sorted(it for p in parts for it in p)
            assert flat == sorted(items)
            loads = [sum(it[3] for it in p) for p in parts]
            assert max(loads) - min(loads) <= max(it[3] for it in items)  # LPT bound

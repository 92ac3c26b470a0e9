"""The BASELINE.json workloads at their stated size, shared by the GPU parity tests, bench.py and
the tools: clip parameters for the in-repo generator, the wanted rows, and a file cache under tests/_cache (clips are
deterministic functions of their parameters, so a cached file is as good as a fresh one)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CACHE = os.path.join(ROOT, 'tests', '_cache')

CONFIG2 = dict(name='bench_1080p_main_cabac_gop30_3000',
               kw=dict(width=1920, height=1080, frames=3000, gop=30, profile=1, bframes=0, num_ref=2, qp=27, seed=2, slices=1, intra_in_p_pct=2))
CONFIG3 = dict(name='sparse3_1080p_high_bpyr_3000',
               kw=dict(width=1920, height=1080, frames=3000, gop=30, profile=2, bframes=3, b_pyramid=1, num_ref=3, weighted=2, seed=3, qp=27))
CONFIG4 = dict(name='sparse4_4k_high_gop250_1000',
               kw=dict(width=3840, height=2160, frames=1000, gop=250, profile=2, bframes=1, num_ref=2, seed=4, qp=30))


def config3_rows(n=3000):
    return list(range(0, n, 17))


def config4_rows(n=1000):
    rng = np.random.default_rng(0)
    return sorted(set(int(x) for x in rng.integers(0, n, 64)))


def config5_clips(frames=300):
    """64 clips, seeded mix of {640x480, 1280x720, 1920x1080, 3840x2160}, profiles cycling CBP / Main / High, GOP 30-60,
    300 frames each (about 700 MB: more than a gpurun snapshot may carry, so the GPU box only gets them in dedicated
    runs).  frames=60: the same 64 clips cut to 60 frames ('config5s_*', about 140 MB, shipped with every snapshot)."""
    rng = np.random.default_rng(5)
    sizes = [(640, 480), (1280, 720), (1920, 1080), (3840, 2160)]
    out = []
    for i in range(64):
        w, h = sizes[int(rng.integers(0, 4))]
        profile = i % 3
        kw = dict(width=w, height=h, frames=frames, gop=int(rng.integers(30, 61)), profile=profile, seed=500 + i, qp=28 + (2 if w >= 3840 else 0),
                  num_ref=2 + profile, bframes=[0, 1, 2][profile], weighted=[0, 0, 2][profile], slices=1 + (i % 5 == 0))
        out.append(dict(name='%s_%02d_%dx%d_p%d' % ('config5' if frames == 300 else 'config5s', i, w, h, profile), kw=kw))
    return out


def path(spec):
    return os.path.join(CACHE, spec['name'] + '.mp4')


def available(spec):
    return os.path.exists(path(spec))


def load(spec, generate=True):
    """mp4 bytes of the workload clip; generated (minutes of CPU for the big ones) when not cached and `generate`."""
    p = path(spec)
    if not os.path.exists(p):
        if not generate:
            return None
        from hwang_b200 import build
        from hwang_b200.testing import streamgen
        build.build_gen()
        os.makedirs(CACHE, exist_ok=True)
        t = time.time()
        data = streamgen.generate(**spec['kw'])
        tmp = p + '.tmp%d' % os.getpid()
        with open(tmp, 'wb') as f:
            f.write(data)
        os.replace(tmp, p)
        sys.stderr.write('[workloads] generated %s: %.1f MB in %.1fs\n' % (os.path.basename(p), len(data) / 1e6, time.time() - t))
    with open(p, 'rb') as f:
        return f.read()

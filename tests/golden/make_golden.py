#!/usr/bin/env python3
"""Regenerates the committed golden fixtures (run in the build container, where /root/reference exists):
  *.mp4   small clips from the in-repo generator (fixed seeds)
  *.json  - per-frame MD5 of the decoded planar YUV as produced by libavcodec (oracle/ffmpeg_oracle.py) and of the
            RGB24 frames as produced by the real libswscale (the reference's get_frame path)
          - the VideoIndex the REFERENCE's own MP4IndexCreator computes for the file (oracle/_ref/ref_tool, compiled
            from /root/reference/hwang/mp4_index_creator.cpp) and its slice_into_video_intervals for several row sets
"""
import hashlib
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from hwang_b200.testing import streamgen  # noqa: E402
from oracle import ffmpeg_oracle as fo, mp4_simple  # noqa: E402

CLIPS = {
    'cbp_cavlc_96x80': dict(width=96, height=80, frames=24, gop=8, profile=0, num_ref=2, slices=2, seed=101, qp=27, ipcm_per_100k=3000, intra_in_p_pct=6),
    'main_cabac_b_96x80': dict(width=96, height=80, frames=24, gop=12, profile=1, bframes=2, num_ref=2, seed=102, qp=28, weighted=2, deblock=3),
    'high_8x8_b_temporal_112x64': dict(width=112, height=64, frames=20, gop=10, profile=2, bframes=1, num_ref=3, seed=103, qp=26, direct_spatial=0,
                                       scaling_lists=1, qp_jitter=2, slices=2),
    'main_fragmented_64x48': dict(width=64, height=48, frames=18, gop=6, profile=1, seed=104, fragmented=1),
}
ROWSETS = [[0], [5], [0, 1, 2, 3], [3, 9, 17], list(range(0, 18, 5)), list(range(18)), [7, 8], [17]]
REF_TOOL = os.path.join(ROOT, 'oracle', '_ref', 'ref_tool')


def main():
    subprocess.check_call([sys.executable, os.path.join(ROOT, 'oracle', 'build_ref.py')])
    for name, kw in CLIPS.items():
        mp4 = streamgen.generate(**kw)
        path = os.path.join(HERE, name + '.mp4')
        with open(path, 'wb') as f:
            f.write(mp4)
        idx = mp4_simple.index_mp4(mp4)
        samples = [mp4[o:o + s] for o, s in zip(idx['offsets'], idx['sizes'])]
        kf = set(idx['keyframes'])
        frames = fo.decode_samples(idx['avcc'], samples, [i in kf for i in range(len(samples))])
        sws = fo.SwsRgb24(kw['width'], kw['height'])
        g = {'params': kw, 'ffmpeg': fo.ffmpeg_version(), 'yuv_md5': [], 'rgb_md5': []}
        for y, u, v in frames:
            g['yuv_md5'].append(hashlib.md5(y.tobytes() + u.tobytes() + v.tobytes()).hexdigest())
            g['rgb_md5'].append(hashlib.md5(sws(y, u, v).tobytes()).hexdigest())
        g['reference_index'] = json.loads(subprocess.check_output([REF_TOOL, 'index', path]))
        g['reference_intervals'] = []
        for rows in ROWSETS:
            rows = [r for r in rows if r < kw['frames']]
            out = json.loads(subprocess.check_output([REF_TOOL, 'slice', path] + [str(r) for r in rows]))
            g['reference_intervals'].append({'rows': rows, 'intervals': out})
        with open(os.path.join(HERE, name + '.json'), 'w') as f:
            json.dump(g, f, indent=1)
        print(name, len(mp4), 'bytes', len(frames), 'frames')


if __name__ == '__main__':
    main()

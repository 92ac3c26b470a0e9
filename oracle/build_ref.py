#!/usr/bin/env python3
"""ORACLE build recipe: compile the parts of the reference that build from their own few sources
(hwang/mp4_index_creator.cpp + util/mp4.h + util/bits.h, hwang/video_index.cpp against a protobuf stub)
straight from /root/reference into oracle/_ref/ref_tool.  Nothing is copied out of the reference.
The H.264 decode arithmetic itself cannot be built this way: it lives in FFmpeg n3.3.1 (deps.sh:148),
which is not vendored; for that row the oracle is oracle/ffmpeg_oracle.py."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = '/root/reference'
OUT = os.path.join(HERE, '_ref')


def main():
    if not os.path.isdir(REF):
        print('reference tree not present; keeping any prebuilt oracle/_ref')
        return 0
    os.makedirs(OUT, exist_ok=True)
    cmd = ['g++', '-std=c++14', '-O1', '-w', '-include', 'cstdint', '-include', 'functional', '-include', 'cstddef',
           '-I', os.path.join(HERE, 'ref_stub'), '-I', REF,
           os.path.join(HERE, 'ref_main.cpp'), os.path.join(REF, 'hwang', 'mp4_index_creator.cpp'), os.path.join(REF, 'hwang', 'video_index.cpp'),
           '-o', os.path.join(OUT, 'ref_tool')]
    print('+', ' '.join(cmd))
    subprocess.check_call(cmd)
    return 0


if __name__ == '__main__':
    sys.exit(main())

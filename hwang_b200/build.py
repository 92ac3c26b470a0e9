"""In-tree build of the native code (no cmake needed: nvcc + g++ directly).

Targets
  hwang_b200/libhwang_b200.so  the product: C++ host side + CUDA kernels for sm_100a, exports include/hwang_b200.h
  build/libh264gen.so          synthetic stream generator (test / bench tooling)
  tests/emu/libhwb_emu.so      TEST ONLY: same host side linked against a host emulation of the device
                               code, so the CPU-only test tier can run the decode core bit-exactly
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST_SRCS = ['b200_video_decoder.cpp', 'capi.cpp', 'decoder_automata.cpp', 'h264_stream.cpp', 'mp4_index_creator.cpp',
             'video_decoder_factory.cpp', 'video_index.cpp']
PRODUCT = os.path.join(ROOT, 'hwang_b200', 'libhwang_b200.so')
GEN = os.path.join(ROOT, 'build', 'libh264gen.so')
EMU = os.path.join(ROOT, 'tests', 'emu', 'libhwb_emu.so')
CXXFLAGS = ['-std=c++17', '-O2', '-g', '-fPIC', '-Wall', '-Wno-unused', '-Wno-unknown-pragmas', '-fno-strict-aliasing', '-pthread']


def _run(cmd):
    print('+', ' '.join(cmd), flush=True)
    subprocess.check_call(cmd, cwd=ROOT)


def _newer(target, srcs):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in srcs)


def _all_sources():
    out = []
    for d in ('hwang_b200/csrc/dev', 'hwang_b200/csrc/host', 'hwang_b200/csrc/cuda', 'tools/h264gen', 'tests/emu', 'include'):
        p = os.path.join(ROOT, d)
        for f in os.listdir(p):
            if f.endswith(('.h', '.cpp', '.cu')):
                out.append(os.path.join(p, f))
    return out


def build_gen(force=False):
    if not force and not _newer(GEN, _all_sources()):
        return GEN
    os.makedirs(os.path.dirname(GEN), exist_ok=True)
    tmp = GEN + '.tmp%d' % os.getpid()  # link beside the target, then rename: a process that has the old library mapped keeps it
    _run(['g++'] + CXXFLAGS + ['-O3', '-shared', 'tools/h264gen/h264gen.cpp', '-o', tmp])
    os.replace(tmp, GEN)
    return GEN


def build_emu(force=False):
    if not force and not _newer(EMU, _all_sources()):
        return EMU
    srcs = ['hwang_b200/csrc/host/' + s for s in HOST_SRCS] + ['tests/emu/devapi_emu.cpp']
    tmp = EMU + '.tmp%d' % os.getpid()
    _run(['g++'] + CXXFLAGS + ['-shared'] + srcs + ['-o', tmp])
    os.replace(tmp, EMU)
    return EMU


def build_product(force=False, out=None, nvcc_extra=()):
    """out / nvcc_extra: an experimental build beside the product (A/B runs on one GPU box: HWB_PRODUCT_LIB=out)."""
    PRODUCT = out or globals()['PRODUCT']
    if not force and not _newer(PRODUCT, _all_sources()):
        return PRODUCT
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    obj = os.path.join(ROOT, 'build', 'obj' if out is None else 'obj_' + os.path.basename(out))
    os.makedirs(obj, exist_ok=True)
    objs = []
    for s in HOST_SRCS:
        o = os.path.join(obj, s.replace('.cpp', '.o'))
        _run(['g++'] + CXXFLAGS + ['-c', 'hwang_b200/csrc/host/' + s, '-o', o])
        objs.append(o)
    ko = os.path.join(obj, 'kernels.o')
    _run([nvcc, '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-Xcompiler', '-fPIC',
          '-Xptxas', '-v'] + list(nvcc_extra) + ['-c', 'hwang_b200/csrc/cuda/kernels.cu', '-o', ko])
    objs.append(ko)
    tmp = PRODUCT + '.tmp%d' % os.getpid()
    _run([nvcc, '-shared', '-o', tmp] + objs + ['-lcudart_static', '-lpthread', '-ldl', '-lrt'])
    os.replace(tmp, PRODUCT)
    return PRODUCT


def build_all(force=False):
    build_gen(force)
    build_emu(force)
    build_product(force)


if __name__ == '__main__':
    if '--variant' in sys.argv:  # python -m hwang_b200.build --variant build/libX.so -DFLAG=1 ...
        i = sys.argv.index('--variant')
        build_product(True, os.path.join(ROOT, sys.argv[i + 1]), sys.argv[i + 2:])
    else:
        build_all('--force' in sys.argv)

"""Fuzz the host side and the decode core on the CPU (no GPU): corrupted MP4 boxes, avcC, slice headers, slice
payloads, truncated and dropped samples, through the emulation build of the library -- meant to run with the emulation
compiled with -fsanitize=address,undefined (tools/README.md).  The decoder must return an error or frames: never
crash, hang or touch memory it does not own (on the GPU an out-of-bounds access is a sticky illegal address for the
whole context, triggered by untrusted media).

  LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 HWB_FUZZ_LIB=/tmp/asan/libhwb_emu_asan.so \\
      python tools/fuzz_host.py [iterations] [seed]
"""
import io, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
from hwang_b200 import _lib, build
_lib.use_library(os.environ.get('HWB_FUZZ_LIB') or build.EMU)
import hwang_b200 as hw
import hwb_testutil as util

ITER = int(sys.argv[1]) if len(sys.argv) > 1 else 200
SEED = int(sys.argv[2]) if len(sys.argv) > 2 else 1
CLIPS = [dict(width=64, height=48, frames=8, gop=4, profile=0, seed=31, num_ref=2, slices=2, ipcm_per_100k=3000),
         dict(width=64, height=48, frames=8, gop=4, profile=1, seed=32, num_ref=3, rplm_pct=50, mmco=1),
         dict(width=80, height=48, frames=9, gop=9, profile=2, seed=33, num_ref=3, bframes=2, b_pyramid=1, weighted=2, direct_spatial=0),
         dict(width=64, height=64, frames=8, gop=4, profile=2, seed=34, num_ref=2, bframes=1, scaling_lists=1, slices=3, pad_refs=1)]


def decode(width, height, fmt, avcc, samples, kf, limit_s=20):
    dec = hw.VideoDecoder(0)
    try:
        dec.configure(width, height, fmt, avcc)
        for s, k in zip(samples, kf):
            dec.feed(s, k)
        dec.feed(None)
        dec.flush()
        got, deadline = 0, time.time() + limit_s
        while got < len(samples):
            n = dec.frames_ready()
            if n != 0:
                dec.get_frame_yuv(); got += 1
            elif time.time() > deadline:
                return 'TIMEOUT'
            else:
                time.sleep(0.0002)
        return 'frames'
    except RuntimeError as e:
        return 'error: ' + str(e)[:70]


def retrieve(mp4, rows, limit_s=30):
    """The python API (index + DecoderAutomata.get_frames) on a damaged file, in a thread so that a hang is seen as one."""
    import threading
    res = {}
    def run():
        try:
            frames = hw.Decoder(io.BytesIO(mp4)).retrieve(rows)
            res['out'] = 'frames' if len(frames) == len(rows) else 'error: %d of %d frames' % (len(frames), len(rows))
        except Exception as e:
            res['out'] = 'error: ' + str(e)[:70]
    t = threading.Thread(target=run, daemon=True)
    t.start(); t.join(limit_s)
    return res.get('out', 'TIMEOUT')


def flip(b, rng, lo, hi, n):
    b = bytearray(b)
    hi = min(hi, len(b))
    for _ in range(n):
        if hi <= lo: break
        pos = rng.randrange(lo, hi)
        b[pos] = rng.randrange(256) if rng.random() < 0.5 else b[pos] ^ (1 << rng.randrange(8))
    return bytes(b)


rng = random.Random(SEED)
made = [util.make_clip(**kw) for kw in CLIPS]
counts = {}
t0 = time.time()
for it in range(ITER):
    ci = rng.randrange(len(CLIPS))
    kw = CLIPS[ci]
    mp4, index, samples, kf = made[ci]
    mode = rng.choice(['mp4', 'avcc', 'header', 'payload', 'truncate', 'drop', 'lenfield', 'swap', 'retrieve', 'retrieve'])
    W, H, fmt, avcc = kw['width'], kw['height'], index.format(), index.metadata_bytes()
    samples = list(samples); kf = list(kf)
    if mode == 'retrieve':
        # the whole python path on a file whose media data (and sometimes boxes) are damaged: sparse or dense rows
        offs = index.sample_offsets()
        lo = offs[0] if rng.random() < 0.8 else 0
        bad = flip(mp4, rng, lo, len(mp4), rng.randrange(1, 12))
        rows = sorted(rng.sample(range(kw['frames']), rng.randrange(1, kw['frames'] + 1)))
        out = retrieve(bad, rows)
    elif mode == 'mp4':
        # container: corrupt the boxes (not the media data), then index and decode whatever the index says
        bad = flip(mp4, rng, 0, min(len(mp4), 4096), rng.randrange(1, 6))
        try:
            idx = hw.index_video(io.BytesIO(bad))
            offs, sizes = idx.sample_offsets(), idx.sample_sizes()
            kfs = set(idx.keyframe_indices())
            ss = [bad[o:o + s] for o, s in zip(offs, sizes)][:16]
            if not ss: out = 'no samples'
            elif any(len(x) != s for x, s in zip(ss, sizes)): out = 'short read'  # the index points outside the file (an empty feed would mean "drain")
            else: out = decode(idx.frame_width(), idx.frame_height(), idx.format(), idx.metadata_bytes(), ss, [i in kfs for i in range(len(ss))])
        except Exception as e:  # index_video raises with the indexer's message
            out = 'index error'
    else:
        if mode == 'avcc':
            avcc = flip(avcc, rng, 0, len(avcc), rng.randrange(1, 4))
        elif mode == 'header':
            v = rng.randrange(len(samples)); samples[v] = flip(samples[v], rng, 4, 14, rng.randrange(1, 4))
        elif mode == 'payload':
            for _ in range(rng.randrange(1, 4)):
                v = rng.randrange(len(samples)); samples[v] = flip(samples[v], rng, 8, len(samples[v]), rng.randrange(1, 10))
        elif mode == 'truncate':
            v = rng.randrange(len(samples)); samples[v] = samples[v][:rng.randrange(1, len(samples[v]))]  # at least one byte: an empty feed means "drain" (end of interval), not a picture
        elif mode == 'drop':
            v = rng.randrange(1, len(samples)); del samples[v]; del kf[v]
        elif mode == 'lenfield':
            v = rng.randrange(len(samples)); samples[v] = flip(samples[v], rng, 0, 4, 1)
        elif mode == 'swap':
            a, b = rng.randrange(len(samples)), rng.randrange(len(samples)); samples[a], samples[b] = samples[b], samples[a]
        out = decode(W, H, fmt, avcc, samples, kf)
    key = mode + ': ' + out.split(':')[0]
    counts[key] = counts.get(key, 0) + 1
    if out == 'TIMEOUT':
        print('TIMEOUT at iteration', it, mode, kw, flush=True)
        if os.environ.get('HWB_FUZZ_DUMP') and mode != 'mp4':  # the failing input, for a replay outside the loop
            import pickle
            pickle.dump(dict(mode=mode, kw=kw, width=W, height=H, format=fmt, avcc=avcc, samples=samples, keyframes=kf), open(os.environ['HWB_FUZZ_DUMP'], 'wb'))
        sys.exit(2)
for k in sorted(counts): print('%-24s %d' % (k, counts[k]))
# the library still decodes a clean clip bit-exactly afterwards
util.assert_yuv_parity(CLIPS[1])
print('fuzz ok: %d iterations, seed %d, %.0f s' % (ITER, SEED, time.time() - t0))

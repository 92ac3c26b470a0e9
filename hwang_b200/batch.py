"""Batch retrieval across clips and GPUs.

The reference retrieves one video, one interval at a time (python/hwang/decoder.py:30-69), re-configuring the decoder
in between.  A B200 wants thousands of slices in flight, and a box has eight of them: `retrieve_many` takes many
(clip, rows) requests, cuts them into keyframe-delimited GOP work items (hwang_b200.shard), assigns the items
longest-first to the devices, and runs one worker thread per device (the native calls release the GIL).  Inside a worker
the intervals of ALL its clips of equal geometry go through ONE decoder with deferred submission
(hwb_decoder_set_defer_submit): their pictures are collected into common GPU batches, so a single entropy launch sees
the slices of many clips -- sparse requests are bound by the latency of one slice, not by throughput, and the GPU
works on all of them at once.  GOPs are independent, so nothing crosses between devices."""
import io
import threading

from . import shard
from .api import PinnedBuffer, VideoDecoder, DeviceFrames, index_video, _Owned
from . import _lib

REORDER_MARGIN = 16  # samples fed after the last wanted frame of an interval (DecoderAutomata::fed_samples)


def _open(src):
    if isinstance(src, (bytes, bytearray)):
        return (lambda b: (lambda: io.BytesIO(b)))(bytes(src))
    if isinstance(src, str):
        return (lambda p: (lambda: open(p, 'rb')))(src)
    data = src.read() if hasattr(src, 'read') else bytes(src)
    return (lambda b: (lambda: io.BytesIO(b)))(data)


def _decode_group(dev, work, srcs, results, device_output):
    """work: [(clip, start_kf, end_kf, cost, rows)] of clips that share width x height, on device `dev`."""
    dec = VideoDecoder(dev)
    dec.set_defer_submit(True)
    plan = []
    for (clip, a, b, _, rows) in work:
        opener, index, _ = srcs[clip]
        offs, sizes = index.sample_offsets(), index.sample_sizes()
        kfs = set(index.keyframe_indices())
        dec.configure(index.frame_width(), index.frame_height(), index.format(), index.metadata_bytes())
        fed = min(b - a, rows[-1] - a + 1 + REORDER_MARGIN)
        with opener() as f:
            f.seek(offs[a], 0)
            blob = f.read(offs[a + fed - 1] + sizes[a + fed - 1] - offs[a])
        dec.set_interval_hint(a, rows)
        for i in range(a, a + fed):
            o = offs[i] - offs[a]
            dec.feed(blob[o:o + sizes[i]], i in kfs)
        dec.feed(None)
        dec.flush()
        plan.append((clip, a, rows, fed, index.frame_width(), index.frame_height()))
    dec.submit_pending()
    use_pinned = _lib.library_path() == _lib.PRODUCT_LIB
    for (clip, a, rows, fed, w, h) in plan:
        fs = w * h * 3
        if device_output:
            buf = DeviceFrames(dev, len(rows), h, w)
            base = buf.ptr
        elif use_pinned:
            buf = PinnedBuffer(fs * len(rows))
            base = buf.ptr
        else:
            import numpy as np
            buf = np.empty(fs * len(rows), np.uint8)
            base = buf.ctypes.data
        k = 0
        for j in range(fed):
            if k < len(rows) and rows[k] == a + j:
                dec.get_frame_into(base + k * fs, fs)
                k += 1
            else:
                dec.discard_frame()
        dec.wait_until_frames_copied()
        if device_output:
            results[clip].append((rows, buf))
        else:
            arr = buf.array if use_pinned else buf
            frames = arr.reshape(len(rows), h, w, 3)
            for i, r in enumerate(rows):
                results[clip].append((r, _Owned(frames[i], buf) if use_pinned else frames[i]))


def retrieve_many(requests, devices=None, indexes=None, device_output=False):
    """requests: list of (path | bytes | file object, rows).  devices: list of device ids (default: every CUDA device).
    -> list (one per request) of lists of (H, W, 3) uint8 frames in the order of each request's sorted rows; with
    device_output=True a list of (rows, DeviceFrames) pairs per request instead (one pair per decoded interval)."""
    from .api import device_count
    if devices is None:
        devices = list(range(max(1, device_count())))
    srcs = []
    for i, (src, rows) in enumerate(requests):
        opener = _open(src)
        index = indexes[i] if indexes is not None else index_video(opener())
        srcs.append((opener, index, sorted(set(rows))))
    items = []
    for i, (_, index, rows) in enumerate(srcs):
        items += shard.gop_work_items(index, i, rows)
    parts = shard.partition(items, len(devices))
    results = [list() for _ in requests]
    errors = []

    def worker(dev, mine):
        try:
            groups = {}
            for it in shard.merge_adjacent(mine):
                index = srcs[it[0]][1]
                groups.setdefault((index.frame_width(), index.frame_height()), []).append(it)
            for key in sorted(groups):
                _decode_group(dev, groups[key], srcs, results, device_output)
        except Exception as e:  # surfaced on the calling thread
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(dev, part)) for dev, part in zip(devices, parts) if part]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    if device_output:
        return [sorted(r, key=lambda p: p[0][0]) for r in results]
    out = []
    for i in range(len(requests)):
        by_row = dict(results[i])
        out.append([by_row[r] for r in srcs[i][2]])
    return out

"""Batch retrieval across clips and GPUs.

The reference retrieves one video, one interval at a time (python/hwang/decoder.py:30-69).  A B200 wants thousands of
slices in flight, and a box has eight of them: `retrieve_many` takes many (clip, rows) requests, cuts them into
keyframe-delimited GOP work items (hwang_b200.shard), assigns the items longest-first to the devices, and runs one
worker thread per device (the native calls release the GIL), each with its own DecoderAutomata per clip.  GOPs are
independent, so nothing crosses between devices."""
import io
import threading

from . import shard
from .api import Decoder, DeviceType, index_video


def retrieve_many(requests, devices=None, indexes=None):
    """requests: list of (path | bytes | file object, rows).  devices: list of device ids (default: every CUDA device).
    -> list (one per request) of lists of (H, W, 3) uint8 frames, in the order of each request's sorted rows."""
    from .api import device_count
    if devices is None:
        devices = list(range(max(1, device_count())))
    srcs = []
    for i, (src, rows) in enumerate(requests):
        if isinstance(src, (bytes, bytearray)):
            opener = (lambda b: (lambda: io.BytesIO(b)))(bytes(src))
        elif isinstance(src, str):
            opener = (lambda p: (lambda: open(p, 'rb')))(src)
        else:
            data = src.read() if hasattr(src, 'read') else bytes(src)
            opener = (lambda b: (lambda: io.BytesIO(b)))(data)
        index = indexes[i] if indexes is not None else index_video(opener())
        srcs.append((opener, index, sorted(set(rows))))
    items = []
    for i, (_, index, rows) in enumerate(srcs):
        items += shard.gop_work_items(index, i, rows)
    parts = shard.partition(items, len(devices))
    results = [dict() for _ in requests]
    errors = []

    def worker(dev, mine):
        try:
            decoders = {}
            for (clip, a, b, cost, want) in shard.merge_adjacent(mine):
                if clip not in decoders:
                    opener, index, _ = srcs[clip]
                    decoders[clip] = Decoder(opener(), video_index=index, device_type=DeviceType.GPU, device_id=dev)
                for r, f in zip(want, decoders[clip].retrieve(want)):
                    results[clip][r] = f
        except Exception as e:  # surfaced on the calling thread
            errors.append(e)

    threads = [threading.Thread(target=worker, args=(dev, part)) for dev, part in zip(devices, parts) if part]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return [[results[i][r] for r in srcs[i][2]] for i in range(len(requests))]

"""BASELINE.json configs[4]: the batch of 64 mixed-resolution clips, GOP-sharded over the GPUs of one box.

  python tools/config5_bench.py [--frames 60|300] [--steps K]                                   one GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/config5_bench.py ...   N GPUs

Every rank indexes the 64 clips, cuts them into keyframe-delimited GOP work items, takes its share from
hwang_b200.shard.partition (cost-weighted, longest first; no collective on the data path) and decodes it densely through
DecoderAutomata.get_frames into page-locked host memory.  Prints one JSON line (rank 0): aggregate frames/s (strong
scaling: the workload is the same whatever N), per-rank frames and seconds, load imbalance."""
import argparse, io, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import hwang_b200 as hw
from hwang_b200 import _lib, shard
from hwang_b200.testing import workloads as wl

ap = argparse.ArgumentParser()
ap.add_argument('--frames', type=int, default=60)
ap.add_argument('--steps', type=int, default=3)
ap.add_argument('--workers', type=int, default=6, help='clips in flight per GPU (each clip alone is bound by the latency of its GOP chains)')
args = ap.parse_args()
rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
_lib.use_library(_lib.PRODUCT_LIB)
L = _lib.lib()
clips = []
for spec in wl.config5_clips(args.frames):
    mp4 = wl.load(spec, generate=False)
    if mp4 is None:
        raise SystemExit('clip %s.mp4 not in tests/_cache' % spec['name'])
    clips.append((mp4, hw.index_video(io.BytesIO(mp4))))
items = []
for ci, (mp4, index) in enumerate(clips):
    items += shard.gop_work_items(index, ci)
mine = shard.merge_adjacent(shard.partition(items, world)[rank])
by_clip = {}
for it in mine:
    by_clip.setdefault(it[0], []).append(it)
work = []
for ci, its in by_clip.items():
    mp4, index = clips[ci]
    rows = [r for it in its for r in it[4]]
    work.append((index, hw.api.encoded_intervals(io.BytesIO(mp4), index, rows), len(rows)))
import threading
from concurrent.futures import ThreadPoolExecutor
os.environ.setdefault('HWB_MEMORY_BUDGET_MB', str(140000 // args.workers))  # the workers share one GPU
tls = threading.local()


def decode_clip(job):
    """one clip's share of this rank, through an automaton of the worker thread (one per geometry)"""
    index, intervals, total = job
    if not hasattr(tls, 'autos'):
        tls.autos, tls.pinned = {}, hw.api.PinnedBuffer(3840 * 2160 * 3 * 4)
    key = (index.frame_width(), index.frame_height())
    if key not in tls.autos:
        tls.autos[key] = hw.DecoderAutomata(hw.DeviceHandle(hw.DeviceType.GPU, local), 1, hw.VideoDecoderType.B200)
    a = tls.autos[key]
    a.initialize(intervals, index.metadata_bytes())
    fs = key[0] * key[1] * 3
    batch = max(1, tls.pinned.nbytes // fs)
    done = 0
    while done < total:
        k = min(batch, total - done)
        if L.hwb_automata_get_frames(a._h, tls.pinned.ptr, k) != 0:
            raise RuntimeError(L.hwb_automata_last_error(a._h).decode())
        done += k
    return total


pool = ThreadPoolExecutor(max_workers=args.workers)
# longest first: the 4K clips start at once, the small ones fill in
work.sort(key=lambda w: -w[2] * w[0].frame_width() * w[0].frame_height())


def step():
    return sum(pool.map(decode_clip, work))


def barrier():
    torch.cuda.synchronize(local)
    if dist:
        dist.barrier()

step()
barrier()
t0 = time.perf_counter()
frames = 0
for _ in range(args.steps):
    frames += step()
t_mine = time.perf_counter() - t0
barrier()
t_all = time.perf_counter() - t0
v = torch.tensor([float(frames), t_mine, t_all], dtype=torch.float64, device='cuda')
allv = [torch.zeros_like(v) for _ in range(world)] if dist else [v]
if dist:
    dist.all_gather(allv, v)
if rank == 0:
    tot = sum(float(x[0]) for x in allv)
    wall = max(float(x[2]) for x in allv)
    print(json.dumps({'workload': 'config 5: 64 mixed-resolution clips x %d frames, dense, GOP-sharded (shard.partition)' % args.frames, 'n_gpus': world, 'clips_in_flight_per_gpu': args.workers,
                      'steps': args.steps, 'frames_per_s': round(tot / wall, 1), 'frames_per_step': int(tot / args.steps),
                      'per_rank_frames': [int(float(x[0]) / args.steps) for x in allv], 'per_rank_busy_s': [round(float(x[1]) / args.steps, 3) for x in allv],
                      'imbalance_max_over_mean': round(max(float(x[1]) for x in allv) / (sum(float(x[1]) for x in allv) / world), 3)}), flush=True)
if dist:
    dist.barrier()
    dist.destroy_process_group()

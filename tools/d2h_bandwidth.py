"""Pinned device-to-host copy bandwidth per GPU, alone and with all ranks copying at once (what bounds dense frame
delivery on a multi-GPU box: every 1080p frame is 6.2 MB of RGB24 that must cross PCIe and land in host DRAM).

  python tools/d2h_bandwidth.py                                   one GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/d2h_bandwidth.py   N GPUs

Prints one JSON line (rank 0): GB/s of every rank copying alone (one after the other) and of all ranks at once, plus the
NUMA / CPU-affinity information the driver exposes."""
import json, os, subprocess, time
import torch

rank = int(os.environ.get('RANK', '0')); world = int(os.environ.get('WORLD_SIZE', '1')); local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dist = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
nbytes = 1 << 30
dev = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
frame = 1920 * 1080 * 3


def measure(seconds=1.5):
    """copies of one frame's size back to back (as get_frame issues them), GB/s"""
    s = torch.cuda.Stream()
    n = 0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s):
        while time.perf_counter() - t0 < seconds:
            for k in range(32):
                off = ((n + k) * frame) % (nbytes - frame)
                host[off:off + frame].copy_(dev[off:off + frame], non_blocking=True)
            n += 32
            s.synchronize()
    dt = time.perf_counter() - t0
    return n * frame / dt / 1e9


def barrier():
    if dist:
        dist.barrier()

measure(0.3)
alone = [0.0] * world
for r in range(world):
    barrier()
    if r == rank:
        alone[r] = measure()
    barrier()
barrier()
together = measure()
vals = torch.tensor([alone[rank], together], dtype=torch.float64, device='cuda')
if dist:
    allv = [torch.zeros_like(vals) for _ in range(world)]
    dist.all_gather(allv, vals)
else:
    allv = [vals]
if rank == 0:
    topo = subprocess.run(['nvidia-smi', 'topo', '-m'], capture_output=True, text=True).stdout
    aff = [l for l in topo.splitlines() if l.startswith('GPU')]
    print(json.dumps({'gpus': world, 'alone_GBps': [round(float(v[0]), 1) for v in allv], 'together_GBps': [round(float(v[1]), 1) for v in allv],
                      'together_sum_GBps': round(sum(float(v[1]) for v in allv), 1), 'host_cores': os.cpu_count(),
                      'frames_per_s_ceiling_together': round(sum(float(v[1]) for v in allv) * 1e9 / frame), 'nvidia_smi_topo': aff}), flush=True)
if dist:
    dist.barrier()
    dist.destroy_process_group()

"""End-to-end timing helper: the bench clip through DecoderAutomata.get_frames into pinned host memory."""
import io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import hwang_b200 as hw
from hwang_b200 import _lib
_lib.use_library(_lib.PRODUCT_LIB)
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 3000
mp4 = bench.get_clip(frames)
index = hw.index_video(io.BytesIO(mp4))
offs, sizes, kf = index.sample_offsets(), index.sample_sizes(), sorted(index.keyframe_indices())
n = len(offs)
L = _lib.lib()
auto = hw.DecoderAutomata(hw.DeviceHandle(hw.DeviceType.GPU, 0), 1, hw.VideoDecoderType.B200)
ed = hw.EncodedData()
ed.width, ed.height, ed.format = bench.W, bench.H, index.format()
ed.start_keyframe, ed.end_keyframe = 0, n
ed.sample_offsets = [o - offs[0] for o in offs]; ed.sample_sizes = sizes; ed.keyframes = kf; ed.valid_frames = list(range(n))
ed.encoded_video = mp4[offs[0]:offs[-1] + sizes[-1]]
batch = 64
fs = bench.W * bench.H * 3
pinned = hw.api.PinnedBuffer(fs * batch)
def step():
    t0 = time.perf_counter()
    auto.initialize([ed], index.metadata_bytes())
    t1 = time.perf_counter()
    done = 0; first = None
    while done < n:
        k = min(batch, n - done)
        if L.hwb_automata_get_frames(auto._h, pinned.ptr, k) != 0: raise RuntimeError(L.hwb_automata_last_error(auto._h).decode())
        if first is None: first = time.perf_counter()
        done += k
    t2 = time.perf_counter()
    return t2 - t0, t1 - t0, first - t0
for r in range(4):
    tot, ini, first = step()
    print('chunk=%s frames=%d total %.3fs (%.0f fps) initialize %.3fs first batch at %.3fs' % (os.environ.get('HWB_CHUNK_PICTURES', 'default'), n, tot, n / tot, ini, first), flush=True)

"""GOP-level work partitioning across GPUs.

Units are keyframe-delimited intervals of (clip, GOP range); they are independent, so sharding needs no data-path
collective.  Cost model: bytes of the interval's samples (available from the VideoIndex) -- entropy decoding is the
dominant stage and scales with bits -- assigned longest-first to the least-loaded rank."""
import heapq


def gop_work_items(index, clip_id=0, rows=None):
    """-> [(clip_id, start_keyframe, end_keyframe, cost_bytes, wanted_rows)] one item per GOP that holds wanted rows."""
    kfs = list(index.keyframe_indices()) + [index.frames()]
    sizes = index.sample_sizes()
    want = None if rows is None else sorted(rows)
    items = []
    j = 0
    for a, b in zip(kfs[:-1], kfs[1:]):
        if want is None:
            r = list(range(a, b))
        else:
            r = []
            while j < len(want) and want[j] < b:
                if want[j] >= a:
                    r.append(want[j])
                j += 1
        if r:
            items.append((clip_id, a, b, int(sum(sizes[a:b])), r))
    return items


def partition(items, world_size):
    """Longest-processing-time-first assignment.  Returns world_size lists; deterministic on every rank."""
    order = sorted(range(len(items)), key=lambda i: (-items[i][3], items[i][0], items[i][1]))
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    out = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        out[r].append(items[i])
        heapq.heappush(heap, (load + items[i][3], r))
    for lst in out:
        lst.sort(key=lambda it: (it[0], it[1]))
    return out


def merge_adjacent(items):
    """Consecutive GOPs of one clip owned by the same rank become one interval (fewer initialize calls)."""
    out = []
    for it in items:
        if out and out[-1][0] == it[0] and out[-1][2] == it[1]:
            p = out[-1]
            out[-1] = (p[0], p[1], it[2], p[3] + it[3], p[4] + it[4])
        else:
            out.append(it)
    return out

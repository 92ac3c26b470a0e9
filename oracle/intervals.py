"""ORACLE (test infrastructure only): Python restatement of slice_into_video_intervals,
hwang/video_index.cpp:62-109, pinned against the reference's own compiled function (oracle/_ref/ref_tool)
through the committed fixtures in tests/golden/*.json."""


def slice_into_video_intervals(sample_offsets, sample_sizes, keyframe_indices, num_frames, rows):
    kf = list(keyframe_indices) + [num_frames]          # :64-65
    intervals, valid_lists = [], []
    start, end = 0, 1                                   # :68-69
    next_keyframe = kf[end]
    valid = []
    for row in rows:                                    # :73
        if row >= next_keyframe:                        # :74
            last_endpoint = sample_offsets[next_keyframe - 1] + sample_sizes[next_keyframe - 1]   # :76-77
            adjacent = last_endpoint == sample_offsets[next_keyframe]                             # :78-79
            end += 1
            next_keyframe = kf[end]                     # :82
            if row >= next_keyframe or not adjacent:    # :84
                if valid:                               # :86-91
                    intervals.append((kf[start], kf[end - 1]))
                    valid_lists.append(valid)
                while row >= kf[end]:                   # :93-96
                    end += 1
                valid = []
                start = end - 1
                next_keyframe = kf[end]
        valid.append(row)                               # :102
    intervals.append((kf[start], kf[end]))              # :104-107
    valid_lists.append(valid)
    return list(zip(intervals, valid_lists))

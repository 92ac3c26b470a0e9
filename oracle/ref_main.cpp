// ORACLE (test infrastructure): drives the REFERENCE's own MP4IndexCreator and slice_into_video_intervals,
// compiled from the sources under /root/reference (never copied), and prints what they compute as JSON.
//   ref_tool index <file.mp4>              1 KiB pull loop of hwang/mp4_index_creator_test.cpp:36-41
//   ref_tool slice <file.mp4> r0 r1 ...    hwang/video_index.cpp:62-109 on that file's index
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <vector>
#include "hwang/mp4_index_creator.h"
#include "hwang/video_index.h"
using namespace hwang;
static void arr(const char *k, const std::vector<uint64_t> &v) { printf("\"%s\": [", k); for (size_t i = 0; i < v.size(); ++i) printf("%s%llu", i ? "," : "", (unsigned long long)v[i]); printf("]"); }
int main(int argc, char **argv) {
  if (argc < 3) return 2;
  std::ifstream f(argv[2], std::ios::binary);
  std::vector<uint8_t> data((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  MP4IndexCreator ic(data.size());
  // First read is 64 bytes, not the 1 KiB of the reference test: when 'ftyp'+'moov'+the first 'moof' all fit in one
  // buffer the reference double-counts the buffer position of that 'moof' (mp4_index_creator.cpp:590,601 add offset_,
  // which already includes it, to bs.offset/8), which only tiny synthetic files can trigger.
  uint64_t off = 0, size = std::min<uint64_t>(64, data.size());
  while (!ic.is_done()) { ic.feed(data.data() + off, size, off, size); }
  if (ic.is_error()) { printf("{\"error\": \"%s\"}\n", ic.error_message().c_str()); return 0; }
  VideoIndex vi = ic.get_video_index();
  if (!strcmp(argv[1], "index")) {
    printf("{\"timescale\": %u, \"duration\": %llu, \"width\": %u, \"height\": %u, \"format\": \"%s\", \"frames\": %llu, ", vi.timescale(), (unsigned long long)vi.duration(),
           vi.frame_width(), vi.frame_height(), vi.format().c_str(), (unsigned long long)vi.frames());
    arr("offsets", vi.sample_offsets()); printf(", "); arr("sizes", vi.sample_sizes()); printf(", "); arr("keyframes", vi.keyframe_indices());
    printf(", \"metadata_hex\": \""); for (uint8_t b : vi.metadata_bytes()) printf("%02x", b); printf("\"}\n");
  } else {
    std::vector<uint64_t> rows;
    for (int i = 3; i < argc; ++i) rows.push_back(strtoull(argv[i], 0, 10));
    VideoIntervals v = slice_into_video_intervals(vi, rows);
    printf("[");
    for (size_t i = 0; i < v.sample_index_intervals.size(); ++i) {
      printf("%s{\"start\": %zu, \"end\": %zu, ", i ? "," : "", std::get<0>(v.sample_index_intervals[i]), std::get<1>(v.sample_index_intervals[i]));
      arr("rows", v.valid_frames[i]); printf("}");
    }
    printf("]\n");
  }
  return 0;
}

// Host/device portability layer for the decode core.
//
// Every routine of the H.264 decode core (entropy decode, MV derivation, reconstruction,
// deblocking, colour conversion) is written once, as warp-cooperative code:
//   * under nvcc the functions are __device__ and HWB_LANES(l) binds `l` to the calling
//     thread's lane id; consecutive HWB_LANES blocks are separated by __syncwarp();
//   * under a plain C++ compiler (stream generator's closed-loop reconstruction and the
//     tests' host emulation of the kernels) HWB_LANES(l) is a 32-iteration loop.
// The discipline that makes both legal: inside a lane block lanes communicate only through
// the explicit "shared" scratch struct handed to the function, never through registers.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define HWB_DEVICE_BUILD 1
#define HWB_HD __device__ __forceinline__
#define HWB_FN __device__ __noinline__  /* big routines stay out of line: the entropy kernel must fit the instruction cache */
#define HWB_TABLE static __device__ const
#define HWB_CTABLE static __constant__ const  /* small tables read by one lane per warp: constant cache */
#define HWB_LANES(l) { const int l = (int)(threadIdx.x & 31);
#define HWB_LANES_END } __syncwarp();
#define HWB_LANE0 if ((threadIdx.x & 31) == 0)
#else
#define HWB_DEVICE_BUILD 0
#define HWB_HD static inline
#define HWB_FN static
#define HWB_TABLE static const
#define HWB_CTABLE static const
#define HWB_LANES(l) for (int l = 0; l < 32; ++l) {
#define HWB_LANES_END }
#define HWB_LANE0
#endif

namespace hwb {

HWB_HD int clip3(int lo, int hi, int v) { return v < lo ? lo : (v > hi ? hi : v); }
HWB_HD int clip8(int v) { return v < 0 ? 0 : (v > 255 ? 255 : v); }
HWB_HD int iabs(int v) { return v < 0 ? -v : v; }
HWB_HD int imin(int a, int b) { return a < b ? a : b; }
HWB_HD int imax(int a, int b) { return a > b ? a : b; }
HWB_HD int median3(int a, int b, int c) { return imax(imin(a, b), imin(imax(a, b), c)); }
HWB_HD int popc32(uint32_t v) {
#if HWB_DEVICE_BUILD
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}
HWB_HD int clz32(uint32_t v) {
#if HWB_DEVICE_BUILD
  return __clz(v);
#else
  return v ? __builtin_clz(v) : 32;
#endif
}

// Loads of data produced earlier in the SAME kernel by another SM (wavefront neighbours)
// must not be served from the non-coherent L1: use ld.global.cg on the device.
HWB_HD uint8_t ld_u8_cg(const uint8_t *p) {
#if HWB_DEVICE_BUILD
  return __ldcg(p);
#else
  return *p;
#endif
}
HWB_HD uint32_t ld_u32_cg(const uint32_t *p) {
#if HWB_DEVICE_BUILD
  return __ldcg(p);
#else
  return *p;
#endif
}

HWB_HD int ld_i8_cg(const int8_t *p) {
#if HWB_DEVICE_BUILD
  return (int)(int8_t)__ldcg((const signed char *)p);
#else
  return *p;
#endif
}
HWB_HD int ld_i16_cg(const int16_t *p) {
#if HWB_DEVICE_BUILD
  return (int)__ldcg((const short *)p);
#else
  return *p;
#endif
}

// 4x4 block index conventions. "z" = luma4x4BlkIdx of the standard (quadrant order),
// "r" = raster index y*4+x inside the macroblock.
HWB_HD int z2x(int z) { return ((z & 1) | ((z >> 1) & 2)); }
HWB_HD int z2y(int z) { return (((z >> 1) & 1) | ((z >> 2) & 2)); }
HWB_HD int xy2z(int x, int y) { return (x & 1) | ((y & 1) << 1) | ((x & 2) << 1) | ((y & 2) << 2); }

}  // namespace hwb

// bf16 hi/lo activation planes (include/clslam_b200.h, clslam_planes): device-side view + the
// epilogue helper that writes a pixel's channels into the padded buffer (reflection halo and
// nearest 2x up-sampling on write).
#pragma once
#include "tc_prims.cuh"

namespace clslam {

struct PlaneView {
  __nv_bfloat16* base;     // element (hi, image 0, padded row 0, padded col 0, first channel of the slice)
  long long plane;         // hi -> lo, elements
  int hp, wp, c_total;     // padded dims, pixel stride
  int h_lo, w_lo;          // dims of the tensor being written (before up-sampling)
  int up, halo;
};

static inline PlaneView make_plane_view(const clslam_planes* p, int up, int halo) {
  PlaneView v{};
  if (!p) return v;
  v.base = reinterpret_cast<__nv_bfloat16*>(p->base) + p->c_off;
  v.plane = p->plane_stride;
  v.hp = p->h + 2; v.wp = p->w + 2; v.c_total = p->c_total;
  v.h_lo = p->h / up; v.w_lo = p->w / up;
  v.up = up; v.halo = halo;
  return v;
}

// padded coordinates written by logical coordinate `o` of a tensor of size `size_lo`:
//   up = 1: o + 1, and the halo cell that reflects onto o (padded 0 <- o == 1, padded size + 1 <- o == size - 2)
//   up = 2: 2o + 1, 2o + 2 (nearest), and the hi-res halo cells (padded 0 mirrors hi-res 1 -> o == 0, ...)
// (three fixed slots, -1 = unused: static indexing keeps them in registers)
__device__ __forceinline__ void plane_coords(int o, int size_lo, int up, int halo, int (&out)[3]) {
  if (up == 1) {
    out[0] = o + 1;
    out[1] = (halo && o == 1) ? 0 : -1;
    out[2] = (halo && o == size_lo - 2) ? size_lo + 1 : -1;
  } else {
    out[0] = 2 * o + 1;
    out[1] = 2 * o + 2;
    out[2] = !halo ? -1 : (o == 0 ? 0 : (o == size_lo - 1 ? 2 * size_lo + 1 : -1));
  }
}

// NV consecutive channels (NV = 4, 8 or 16) of logical pixel (img, oh, ow); ch is relative to the slice
template <int NV>
__device__ __forceinline__ void planes_emit(const PlaneView& po, int img, int oh, int ow, int ch, const float (&v)[NV]) {
  static_assert(NV == 4 || NV == 8 || NV == 16, "NV");
  uint32_t hi[NV / 2], lo[NV / 2];
#pragma unroll
  for (int j = 0; j < NV / 2; ++j) split_bf16x2(v[2 * j], v[2 * j + 1], hi[j], lo[j]);
  int rows[3], cols[3];
  plane_coords(oh, po.h_lo, po.up, po.halo, rows);
  plane_coords(ow, po.w_lo, po.up, po.halo, cols);
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      if (rows[a] < 0 || cols[b] < 0) continue;
      __nv_bfloat16* p = po.base + (((size_t)img * po.hp + rows[a]) * po.wp + cols[b]) * po.c_total + ch;
      if constexpr (NV == 4) {
        *reinterpret_cast<uint2*>(p) = make_uint2(hi[0], hi[1]);
        *reinterpret_cast<uint2*>(p + po.plane) = make_uint2(lo[0], lo[1]);
      } else {
#pragma unroll
        for (int j = 0; j < NV / 8; ++j) {
          *reinterpret_cast<uint4*>(p + 8 * j) = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
          *reinterpret_cast<uint4*>(p + po.plane + 8 * j) = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
        }
      }
    }
  }
}

// hi + lo of 8 consecutive channels at padded position (row, col) of a planes tensor
__device__ __forceinline__ void planes_load8(const PlaneView& pv, int img, int row, int col, int ch, float (&v)[8]) {
  const __nv_bfloat16* p = pv.base + (((size_t)img * pv.hp + row) * pv.wp + col) * pv.c_total + ch;
  const uint4 h = *reinterpret_cast<const uint4*>(p);
  const uint4 l = *reinterpret_cast<const uint4*>(p + pv.plane);
  const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float2 f = join_bf16x2(hh[j], ll[j]);
    v[2 * j] = f.x; v[2 * j + 1] = f.y;
  }
}

}  // namespace clslam

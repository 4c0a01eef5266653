// ilf_sao.cu -- sample adaptive offset (sm_100a).
//
// Replaces SampleAdaptiveOffset::SAOProcess after parameter resolution (SampleAdaptiveOffset.cpp:585-601,
// offsetCTU :510-562, offsetBlock :292-508).  The source is the whole deblocked picture (the reference copies
// it to m_tempBuf, :587), so every sample is independent: one thread produces one int16x8 vector of one row
// from the three rows around it.  Horizontal neighbours come from the adjacent lanes by warp shuffle (the two
// border lanes of a warp fetch one extra sample each); the rows above/below are re-read through L1/L2, HBM
// sees each sample once.  All three planes go in one launch.
//
// The row/column special cases of offsetBlock (:308-487) are the statement "an edge-offset sample is modified
// iff both neighbours along the class direction are inside the CTU block or inside a neighbouring CTU whose
// availability flag (deriveLoopFilterBoundaryAvailibility, :685-760) is set" -- evaluated here per vector.
#include "ilf_common.cuh"

namespace ilf {
namespace {

constexpr int ROWS_PER_CTA = 8;

__device__ __forceinline__ int sgn(int v) { return (v > 0) - (v < 0); }

__device__ __forceinline__ void unpack8(const uint4& r, int v[8]) {
  v[0] = (int)(int16_t)(r.x & 0xFFFF); v[1] = (int)(int16_t)(r.x >> 16);
  v[2] = (int)(int16_t)(r.y & 0xFFFF); v[3] = (int)(int16_t)(r.y >> 16);
  v[4] = (int)(int16_t)(r.z & 0xFFFF); v[5] = (int)(int16_t)(r.z >> 16);
  v[6] = (int)(int16_t)(r.w & 0xFFFF); v[7] = (int)(int16_t)(r.w >> 16);
}
__device__ __forceinline__ uint4 pack8(const int v[8]) {
  uint4 o;
  o.x = (uint32_t)(uint16_t)v[0] | ((uint32_t)(uint16_t)v[1] << 16);
  o.y = (uint32_t)(uint16_t)v[2] | ((uint32_t)(uint16_t)v[3] << 16);
  o.z = (uint32_t)(uint16_t)v[4] | ((uint32_t)(uint16_t)v[5] << 16);
  o.w = (uint32_t)(uint16_t)v[6] | ((uint32_t)(uint16_t)v[7] << 16);
  return o;
}

// Availability of the region a neighbour falls in.  dxr/dyr in {-1,0,1}: left/inside/right, above/inside/below.
__device__ __forceinline__ bool region_ok(unsigned m9, int dxr, int dyr) { return (m9 >> ((dyr + 1) * 3 + dxr + 1)) & 1u; }

__global__ void __launch_bounds__(32 * ROWS_PER_CTA) sao_kernel(Geom g, const SlotDev* __restrict__ slots, int first_slot,
                                                                int src_b, int dst_b,
                                                                int groups_y, int groups_c) {
  const SlotDev& sd = slots[first_slot + blockIdx.z];
  const int lane = threadIdx.x & 31, wrow = threadIdx.x >> 5;
  // blockIdx.y enumerates row groups of Y, then Cb, then Cr.
  int plane, grp = blockIdx.y;
  if (grp < groups_y) plane = 0;
  else if (grp < groups_y + groups_c) { plane = 1; grp -= groups_y; }
  else { plane = 2; grp -= groups_y + groups_c; }
  const int sh = plane ? 1 : 0;
  const int pw = g.width >> sh, ph_local = g.rows >> sh, ph_global = g.height >> sh;
  const int pitch = plane ? g.pitch_c : g.pitch_y;
  const int y = grp * ROWS_PER_CTA + wrow;                 // local row
  const int x0 = (blockIdx.x * 32 + lane) * 8;
  const bool in = x0 < pw && y < ph_local;
  if (blockIdx.x * 256 >= pw || grp * ROWS_PER_CTA >= ph_local) return;  // whole CTA outside (chroma grids are padded)
  const unsigned full = 0xffffffffu;
  if (__all_sync(full, !in)) return;

  const int16_t* __restrict__ src = sd.buf[src_b][plane];
  int16_t* __restrict__ dst = sd.buf[dst_b][plane];
  const int gy = y + (g.row0 >> sh);                         // picture row
  const int ctu_sz = 1 << (g.ctu_log2 - sh);
  int type = ILF_SAO_OFF, band = 0, o0 = 0, o1 = 0, o2 = 0, o3 = 0;
  unsigned avail = 0;
  int ctu_x0 = 0, ctu_y0 = 0;
  if (in) {
    const int cx = x0 >> (g.ctu_log2 - sh), cy = gy >> (g.ctu_log2 - sh);
    const ilf_sao_ctu* __restrict__ p = sd.sao + (size_t)cy * g.ctus_w + cx;
    type = p->type[plane];
    if (type != ILF_SAO_OFF) {
      band = p->band_pos[plane];
      o0 = p->offset[plane][0]; o1 = p->offset[plane][1]; o2 = p->offset[plane][2]; o3 = p->offset[plane][3];
      avail = p->avail;
      ctu_x0 = cx * ctu_sz; ctu_y0 = cy * ctu_sz;
    }
  }
  const size_t row_off = (size_t)y * pitch;
  uint4 rc = make_uint4(0, 0, 0, 0);
  if (in) rc = ldg_u4(src + row_off + x0);

  const bool vert = type == ILF_SAO_EO_90 || type == ILF_SAO_EO_135 || type == ILF_SAO_EO_45;
  const bool any_vert = __any_sync(full, vert);
  const bool any_on = __any_sync(full, type != ILF_SAO_OFF);
  if (!any_on) {
    if (in) *reinterpret_cast<uint4*>(dst + row_off + x0) = rc;
    return;
  }

  // rows above / below (clamped inside the held rows; unavailable neighbours are never used)
  uint4 ra = rc, rb = rc;
  const int ya = max(y - 1, 0), yb = min(y + 1, ph_local - 1);
  if (any_vert && in) {
    ra = ldg_u4(src + (size_t)ya * pitch + x0);
    rb = ldg_u4(src + (size_t)yb * pitch + x0);
  }
  // horizontal halo: last sample of the lane to the left, first sample of the lane to the right
  int cl = __shfl_up_sync(full, (int)(rc.w >> 16), 1), cr = __shfl_down_sync(full, (int)(rc.x & 0xFFFF), 1);
  int al = __shfl_up_sync(full, (int)(ra.w >> 16), 1), ar = __shfl_down_sync(full, (int)(ra.x & 0xFFFF), 1);
  int bl = __shfl_up_sync(full, (int)(rb.w >> 16), 1), br = __shfl_down_sync(full, (int)(rb.x & 0xFFFF), 1);
  if (in && type != ILF_SAO_OFF && type != ILF_SAO_BO && type != ILF_SAO_EO_90) {
    if (lane == 0 && x0 > 0) {
      cl = (uint16_t)src[row_off + x0 - 1];
      if (vert) { al = (uint16_t)src[(size_t)ya * pitch + x0 - 1]; bl = (uint16_t)src[(size_t)yb * pitch + x0 - 1]; }
    }
    if (lane == 31 && x0 + 8 < pw) {
      cr = (uint16_t)src[row_off + x0 + 8];
      if (vert) { ar = (uint16_t)src[(size_t)ya * pitch + x0 + 8]; br = (uint16_t)src[(size_t)yb * pitch + x0 + 8]; }
    }
  }
  if (!in) return;
  if (type == ILF_SAO_OFF) {
    *reinterpret_cast<uint4*>(dst + row_off + x0) = rc;
    return;
  }

  const int max_val = (1 << (plane ? g.bd_chroma : g.bd_luma)) - 1;
  int c[8], out[8];
  unpack8(rc, c);
  if (type == ILF_SAO_BO) {
    const int shift = (plane ? g.bd_chroma : g.bd_luma) - 5;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      const int k = ((c[j] >> shift) - band) & 31;
      const int off = k == 0 ? o0 : k == 1 ? o1 : k == 2 ? o2 : k == 3 ? o3 : 0;
      out[j] = clip3i(0, max_val, c[j] + off);
    }
  } else {
    // neighbour rows n1 (first neighbour) and n2 (second neighbour), 10 samples each: index j+1 <-> column x0 + j
    int n1[10], n2[10];
    int sx;  // x step of the SECOND neighbour; the first one is the mirror image
    {
      int a[8], b[8];
      unpack8(ra, a);
      unpack8(rb, b);
      const int16_t al16 = (int16_t)al, ar16 = (int16_t)ar, bl16 = (int16_t)bl, br16 = (int16_t)br, cl16 = (int16_t)cl, cr16 = (int16_t)cr;
      if (type == ILF_SAO_EO_0) {
        sx = 1;
        n1[0] = cl16; n2[0] = cl16; n1[9] = cr16; n2[9] = cr16;
#pragma unroll
        for (int j = 0; j < 8; j++) { n1[j + 1] = c[j]; n2[j + 1] = c[j]; }
      } else {
        sx = type == ILF_SAO_EO_90 ? 0 : type == ILF_SAO_EO_135 ? 1 : -1;
        n1[0] = al16; n1[9] = ar16; n2[0] = bl16; n2[9] = br16;
#pragma unroll
        for (int j = 0; j < 8; j++) { n1[j + 1] = a[j]; n2[j + 1] = b[j]; }
      }
    }
    const int sy = type == ILF_SAO_EO_0 ? 0 : 1;  // second neighbour is below (first above) for the other classes
    // 3x3 availability mask of the CTU: bit (dyr+1)*3 + (dxr+1)
    const unsigned m9 = ((avail & ILF_AVAIL_AL) ? 1u : 0u) | ((avail & ILF_AVAIL_A) ? 2u : 0u) | ((avail & ILF_AVAIL_AR) ? 4u : 0u) |
                        ((avail & ILF_AVAIL_L) ? 8u : 0u) | 16u | ((avail & ILF_AVAIL_R) ? 32u : 0u) |
                        ((avail & ILF_AVAIL_BL) ? 64u : 0u) | ((avail & ILF_AVAIL_B) ? 128u : 0u) | ((avail & ILF_AVAIL_BR) ? 256u : 0u);
    // rj = index of the vector's sample that is the last column of the CTU block (>= 8: none).  Chroma widths are
    // multiples of 4 only, so the block may end in the middle of the picture's last vector.
    const bool at_left = x0 == ctu_x0;
    const int rj = min(ctu_x0 + ctu_sz, pw) - x0 - 1;
    const bool at_right = rj < 8;
    const bool at_top = gy == ctu_y0, at_bottom = (gy == min(ctu_y0 + ctu_sz, ph_global) - 1);
    const int dy1 = (at_top && sy) ? -1 : 0, dy2 = (at_bottom && sy) ? 1 : 0;  // first neighbour is above, second below
    // first neighbour x step = -sx, second = +sx
    const bool ok_mid = region_ok(m9, 0, dy1) && region_ok(m9, 0, dy2);
    const bool ok_first = region_ok(m9, sx > 0 ? -1 : 0, dy1) && region_ok(m9, sx < 0 ? -1 : 0, dy2);
    const bool ok_last = region_ok(m9, (at_right && sx < 0) ? 1 : 0, dy1) && region_ok(m9, (at_right && sx > 0) ? 1 : 0, dy2);
#pragma unroll
    for (int j = 0; j < 8; j++) {
      // first neighbour at column j - sx, second at j + sx  (arrays are offset by one)
      const int v1 = sx == 0 ? n1[j + 1] : (sx > 0 ? n1[j] : n1[j + 2]);
      const int v2 = sx == 0 ? n2[j + 1] : (sx > 0 ? n2[j + 2] : n2[j]);
      const int e = sgn(c[j] - v1) + sgn(c[j] - v2);
      const int off = e == -2 ? o0 : e == -1 ? o1 : e == 1 ? o2 : e == 2 ? o3 : 0;
      const bool ok = j == rj ? ok_last : ((j == 0 && at_left) ? ok_first : ok_mid);
      out[j] = ok ? clip3i(0, max_val, c[j] + off) : c[j];
    }
  }
  *reinterpret_cast<uint4*>(dst + row_off + x0) = pack8(out);
}

}  // namespace

void launch_sao(const Geom& g, const SlotDev* slots, int first_slot, int num_slots, int src_b, int dst_b, cudaStream_t st) {
  const int groups_y = (g.rows + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  const int groups_c = (g.rows / 2 + ROWS_PER_CTA - 1) / ROWS_PER_CTA;
  dim3 grid((g.width + 255) / 256, groups_y + 2 * groups_c, num_slots);
  sao_kernel<<<grid, 32 * ROWS_PER_CTA, 0, st>>>(g, slots, first_slot, src_b, dst_b, groups_y, groups_c);
}

}  // namespace ilf

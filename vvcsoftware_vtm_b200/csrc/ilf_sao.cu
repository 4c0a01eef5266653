// ilf_sao.cu -- sample adaptive offset (sm_100a).
//
// Replaces SampleAdaptiveOffset::SAOProcess after parameter resolution (SampleAdaptiveOffset.cpp:585-601,
// offsetCTU :510-562, offsetBlock :292-508).  The source is the whole deblocked picture (the reference copies it to
// m_tempBuf, :587), so every sample is independent.
//
// Data movement: band walking over a TMA ring (ilf_ring.cuh).  A CTA owns a band of 32 rows of one plane (or a
// horizontal segment of it when few pictures are in the batch) and walks it in tiles of 128 samples; the TMA unit
// fetches each tile with one halo row above and below (box 128 x 34, always 256-byte aligned rows) into a ring of
// shared-memory stages, several tiles ahead of the arithmetic.  The left / right neighbours of a tile's first / last
// column are in the neighbouring tiles of the ring.  Nothing is read twice horizontally, 6 % vertically.
//
// Arithmetic: a thread owns 8 samples x 4 rows of the tile: it reads its rows and the two neighbouring rows as int16x8
// vectors from shared memory, filters two samples per instruction (ilf_packed.cuh) and stores int16x8 vectors (16-byte
// aligned) to the destination plane.  A strip lies inside one CTU, so the CTU's parameters are fetched once per
// thread and tile; a warp spans one 128-sample CTU width, so it does not diverge on the SAO type at CTU size 128.
// CTUs with SAO off are copied through; planes whose SAO is off for the whole picture are not touched at all
// (BatchCtl skip bit).
//
// The row/column special cases of offsetBlock (:308-487) are the statement "an edge-offset sample is modified iff both
// neighbours along the class direction are inside the CTU block or inside a neighbouring CTU whose availability flag
// (deriveLoopFilterBoundaryAvailibility, :685-760) is set"; it is evaluated per row for the first column, the last
// column and the columns in between, and only for CTUs that have an unavailable neighbour.
#include "ilf_common.cuh"
#include "ilf_packed.cuh"
#include "ilf_ring.cuh"

namespace ilf {
namespace {

constexpr int R = 4;                       // rows per thread and tile
constexpr int TW = SAO_TILE;               // tile width (samples)
constexpr int BR = SAO_BAND_ROWS;          // rows of a band
constexpr int SR = BR + 2;                 // staged rows: one halo row above and below
#ifndef SAO_STAGES
#define SAO_STAGES 6
#endif
constexpr int STAGES = SAO_STAGES;                  // ring depth: previous, current, next tile + 3 tiles in flight
constexpr int STAGE_BYTES = TW * SR * 2;   // 8704
#ifndef SAO_CTAS
#define SAO_CTAS (SAO_TILE_W == 128 ? 4 : 8)
#endif
constexpr int KPR = TW / 8;                // eight-sample strips per tile row
constexpr int NTHREADS = KPR * (BR / R);   // one strip of R rows per thread
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + STAGES * 8;

struct Row { uint32_t v[4]; };  // 8 samples

__device__ __forceinline__ Row ld_row(const int16_t* p) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  Row o; o.v[0] = r.x; o.v[1] = r.y; o.v[2] = r.z; o.v[3] = r.w;
  return o;
}
__device__ __forceinline__ void st_row(int16_t* p, const Row& r) { *reinterpret_cast<uint4*>(p) = make_uint4(r.v[0], r.v[1], r.v[2], r.v[3]); }

// The row shifted by one sample: s[0] = [x-1, x0], s[1..3] = [x1,x2] [x3,x4] [x5,x6], s[4] = [x7, x+8].
// left neighbours of the 8 samples = s[0..3], right neighbours = s[1..4].
struct Shifted { uint32_t s[5]; };
__device__ __forceinline__ Shifted shift_row(const Row& r, uint32_t left, uint32_t right) {
  Shifted o;
  o.s[0] = (r.v[0] << 16) | (left & 0xFFFFu);
  o.s[1] = pk::funnel16(r.v[0], r.v[1]);
  o.s[2] = pk::funnel16(r.v[1], r.v[2]);
  o.s[3] = pk::funnel16(r.v[2], r.v[3]);
  o.s[4] = (r.v[3] >> 16) | (right << 16);
  return o;
}

// Availability of the region a neighbour falls in.  dxr/dyr in {-1,0,1}: left/inside/right, above/inside/below.
__device__ __forceinline__ bool region_ok(unsigned m9, int dxr, int dyr) { return (m9 >> ((dyr + 1) * 3 + dxr + 1)) & 1u; }

struct EoCtx {
  uint32_t lut_lo, lut_hi, maxv;
  unsigned m9;
  int ctu_y0, last_row, rj, gy0, nrows;
  bool at_left, need_mask;
};

// Edge offset of the strip's rows for one class: first neighbour a = (x - SX, y - SY), second b = (x + SX, y + SY).
// v[0] = row above the strip, v[1 + r] = row r, v[R + 1] = row below; sft[] = the same rows shifted by one sample.
template <int SX, int SY>
__device__ __forceinline__ void eo_rows(const EoCtx& e, const Row (&v)[R + 2], const Shifted (&sft)[R + 2], int16_t* __restrict__ out, int pitch) {
#pragma unroll
  for (int r = 0; r < R; r++) {
    if (r >= e.nrows) break;
    const Row& c = v[r + 1];
    uint32_t a[4], b[4];
    if (SX == 0) {
#pragma unroll
      for (int k = 0; k < 4; k++) { a[k] = v[r].v[k]; b[k] = v[r + 2].v[k]; }
    } else {
      constexpr int ia_off = SY ? 0 : 1, ib_off = SY ? 2 : 1;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        a[k] = sft[r + ia_off].s[SX > 0 ? k : k + 1];
        b[k] = sft[r + ib_off].s[SX > 0 ? k + 1 : k];
      }
    }
    Row o;
#pragma unroll
    for (int k = 0; k < 4; k++) o.v[k] = pk::sao_apply2(c.v[k], pk::sao_eo_index2(c.v[k], a[k], b[k]), e.lut_lo, e.lut_hi, e.maxv);
    if (e.need_mask) {
      const int gy = e.gy0 + r;
      const bool at_top = gy == e.ctu_y0, at_bottom = gy == e.last_row;
      const int dy1 = (at_top && SY) ? -1 : 0, dy2 = (at_bottom && SY) ? 1 : 0;
      const bool ok_mid = region_ok(e.m9, 0, dy1) && region_ok(e.m9, 0, dy2);
      const bool ok_first = region_ok(e.m9, SX > 0 ? -1 : 0, dy1) && region_ok(e.m9, SX < 0 ? -1 : 0, dy2);
      const bool ok_last = region_ok(e.m9, SX < 0 ? 1 : 0, dy1) && region_ok(e.m9, SX > 0 ? 1 : 0, dy2);
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const bool k0 = (k == 0 && e.at_left) ? ok_first : ok_mid;
        const bool l0 = e.rj == 2 * k ? ok_last : k0;
        const bool l1 = e.rj == 2 * k + 1 ? ok_last : ok_mid;
        const uint32_t m = (l0 ? 0xFFFFu : 0u) | (l1 ? 0xFFFF0000u : 0u);
        o.v[k] = (o.v[k] & m) | (c.v[k] & ~m);
      }
    }
    st_row(out + (size_t)r * pitch, o);
  }
}

// One tile of the walk: `cur` is the tile's stage (SR rows x TW samples), `prev` / `next` the neighbouring tiles' stages (or
// nullptr at the picture border, where the availability flags already exclude the missing neighbour).
// The strip's CTU parameters (two 16-byte words of ilf_sao_ctu): loaded one tile ahead by the walk, so that their latency is
// not exposed at the start of every tile (everything below branches on the type).
struct CtuPrm { uint4 a, b; };
__device__ __forceinline__ CtuPrm load_ctu_prm(const Geom& g, const SlotDev& sd, int plane, int tx, int by0) {
  const int k = threadIdx.x % KPR, rg = threadIdx.x / KPR;
  const int sh = plane ? 1 : 0;
  const int ctu_log2 = g.ctu_log2 - sh;
  const int cx = min((tx * TW + 8 * k) >> ctu_log2, g.ctus_w - 1), cy = min((by0 + R * rg + (g.row0 >> sh)) >> ctu_log2, g.ctus_h - 1);
  const uint4* __restrict__ prm = reinterpret_cast<const uint4*>(sd.sao + (size_t)cy * g.ctus_w + cx);
  CtuPrm p;
  p.a = __ldg(prm); p.b = __ldg(prm + 1);
  return p;
}

__device__ __forceinline__ void sao_tile(const Geom& g, const SlotDev& sd, int plane, int16_t* __restrict__ dst, int tx, int by0, const int16_t* cur,
                                         const int16_t* prev, const int16_t* next, const CtuPrm& cp) {
  const int k = threadIdx.x % KPR, rg = threadIdx.x / KPR;
  const int sh = plane ? 1 : 0;
  const int pw = g.width >> sh, ph_local = g.rows >> sh, ph_global = g.height >> sh;
  const int pitch = plane ? g.pitch_c : g.pitch_y;
  const int x0 = tx * TW + 8 * k, y0 = by0 + R * rg;  // local row of the strip's first row
  if (x0 >= pw || y0 >= ph_local) return;
  const int nrows = min(R, ph_local - y0);
  const int gy0 = y0 + (g.row0 >> sh);  // picture row
  const int ctu_log2 = g.ctu_log2 - sh, ctu_sz = 1 << ctu_log2;
  const int cx = x0 >> ctu_log2, cy = gy0 >> ctu_log2;
  const uint4 pa = cp.a, pb = cp.b;
  // ilf_sao_ctu: offset[3][4] int16 (24 bytes), type[3] int8 at byte 24, band_pos[3] at 27, avail at 30
  const int type = (int)(int8_t)((plane == 0 ? pb.z : plane == 1 ? pb.z >> 8 : pb.z >> 16) & 0xFF);
  const int16_t* base = cur + (R * rg) * TW + 8 * k;  // staged row 0 of the strip = the row above it
  int16_t* out = dst + (size_t)y0 * pitch + x0;

  if (type == ILF_SAO_OFF) {  // plain copy
#pragma unroll
    for (int r = 0; r < R; r++) if (r < nrows) st_row(out + (size_t)r * pitch, ld_row(base + (r + 1) * TW));
    return;
  }

  // ---- parameters of this component ----
  const uint32_t w0 = plane == 0 ? pa.x : plane == 1 ? pa.z : pb.x;  // offset[plane][0..1]
  const uint32_t w1 = plane == 0 ? pa.y : plane == 1 ? pa.w : pb.y;  // offset[plane][2..3]
  const uint32_t o0 = w0 & 0xFF, o1 = (w0 >> 16) & 0xFF, o2 = w1 & 0xFF, o3 = (w1 >> 16) & 0xFF;  // int8 range checked on the host
  const int bd = plane ? g.bd_chroma : g.bd_luma;
  const uint32_t maxv = pk::splat((1 << bd) - 1);
  const unsigned avail = (pb.w >> 16) & 0xFF;

  if (type == ILF_SAO_BO) {
    const uint32_t lut_lo = o0 | (o1 << 8) | (o2 << 16) | (o3 << 24), lut_hi = 0;
    const int band = (plane == 0 ? pb.z >> 24 : plane == 1 ? pb.w : pb.w >> 8) & 0xFF;
    const uint32_t nband = pk::splat((32 - band) & 31);
#pragma unroll
    for (int r = 0; r < R; r++) {
      if (r >= nrows) break;
      const Row c = ld_row(base + (r + 1) * TW);
      Row o;
#pragma unroll
      for (int q = 0; q < 4; q++) o.v[q] = pk::sao_apply2(c.v[q], pk::sao_bo_index2(c.v[q], bd - 5, nband), lut_lo, lut_hi, maxv);
      st_row(out + (size_t)r * pitch, o);
    }
    return;
  }

  // ---- edge offset: the strip's rows, the rows above and below, and (except for the vertical class) their shifted copies ----
  Row v[R + 2];
  Shifted sft[R + 2];
#pragma unroll
  for (int i = 0; i < R + 2; i++) v[i] = ld_row(base + i * TW);
  if (type != ILF_SAO_EO_90) {
    const int16_t* lp = k > 0 ? base - 1 : (prev ? prev + (R * rg) * TW + TW - 1 : nullptr);
    const int16_t* rp = k < KPR - 1 ? base + 8 : (next ? next + (R * rg) * TW : nullptr);
#pragma unroll
    for (int i = 0; i < R + 2; i++) {
      const uint32_t lft = lp ? (uint16_t)lp[i * TW] : 0u, rgt = rp ? (uint16_t)rp[i * TW] : 0u;
      sft[i] = shift_row(v[i], lft, rgt);
    }
  }

  EoCtx e;
  e.lut_lo = o0 | (o1 << 8) | (o2 << 24); e.lut_hi = o3;  // index 2 (flat) adds 0
  e.maxv = maxv;
  // 3x3 availability mask of the CTU: bit (dyr+1)*3 + (dxr+1)
  e.m9 = ((avail & ILF_AVAIL_AL) ? 1u : 0u) | ((avail & ILF_AVAIL_A) ? 2u : 0u) | ((avail & ILF_AVAIL_AR) ? 4u : 0u) |
         ((avail & ILF_AVAIL_L) ? 8u : 0u) | 16u | ((avail & ILF_AVAIL_R) ? 32u : 0u) |
         ((avail & ILF_AVAIL_BL) ? 64u : 0u) | ((avail & ILF_AVAIL_B) ? 128u : 0u) | ((avail & ILF_AVAIL_BR) ? 256u : 0u);
  const int ctu_x0 = cx << ctu_log2;
  e.ctu_y0 = cy << ctu_log2;
  e.at_left = x0 == ctu_x0;
  e.rj = min(ctu_x0 + ctu_sz, pw) - x0 - 1;  // index of the block's last column inside this vector (>= 8: not here)
  e.last_row = min(e.ctu_y0 + ctu_sz, ph_global) - 1;
  e.need_mask = avail != 0xFFu;
  e.gy0 = gy0; e.nrows = nrows;
  if (type == ILF_SAO_EO_0) eo_rows<1, 0>(e, v, sft, out, pitch);
  else if (type == ILF_SAO_EO_90) eo_rows<0, 1>(e, v, sft, out, pitch);
  else if (type == ILF_SAO_EO_135) eo_rows<1, 1>(e, v, sft, out, pitch);
  else eo_rows<-1, 1>(e, v, sft, out, pitch);
}

__global__ void __launch_bounds__(NTHREADS, SAO_CTAS) sao_kernel(Geom g, const SlotDev* __restrict__ slots, int first_slot, BatchCtl bc, int bands_y, int bands_c, int nseg) {
  extern __shared__ __align__(128) unsigned char smem[];
  pdl_launch_dependents();
  const unsigned ctl = bc.v[blockIdx.z];
  const SlotDev& sd = slots[first_slot + bc.slot[blockIdx.z]];
  // blockIdx.y enumerates the bands of Y, then Cb, then Cr; blockIdx.x the horizontal segments of a band
  int band = blockIdx.y, plane = 0;
  if (band >= bands_y) { band -= bands_y; plane = 1; if (band >= bands_c) { band -= bands_c; plane = 2; } }
  if (ctl_skip(ctl, plane)) return;
  const int pw = plane ? g.width >> 1 : g.width;
  const int ntx = (pw + TW - 1) / TW;
  const int ta = (int)blockIdx.x * ntx / nseg, tb = ((int)blockIdx.x + 1) * ntx / nseg;  // this CTA's tiles [ta, tb)
  if (ta >= tb) return;
  ring::Walk<STAGES> walk;
  walk.first = max(ta - 1, 0); walk.last = min(tb, ntx - 1);  // tiles the walk loads: its own and one neighbour on each side
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  const int tid = threadIdx.x;
  const int by0 = band * BR;
  const int src_buf = ctl_src(ctl, plane);
  int16_t* __restrict__ dst = sd.buf[ctl_dst(ctl, plane)][plane];
  const CUtensorMap* map = &sd.tm_sao[plane];
  auto stage_ptr = [&](int t) { return reinterpret_cast<int16_t*>(smem + walk.stage(t) * STAGE_BYTES); };
  auto issue = [&](int t) {
    uint64_t* bar = &full[walk.stage(t)];
    ring::mbar_expect_tx(bar, STAGE_BYTES);
    ring::tma_load_3d(stage_ptr(t), map, bar, t * TW, by0 - 1, src_buf);
  };
  if (tid == 0) {
    for (int i = 0; i < STAGES; i++) ring::mbar_init(&full[i], 1);
    ring::mbar_init_fence();
  }
  __syncthreads();
  pdl_wait();  // deblocking has written the planes read from here on
  if (tid == 0)
    for (int t = walk.first; t <= walk.last && t < walk.first + STAGES; t++) issue(t);
  // stage index and barrier phase of tiles tx - 1, tx, tx + 1, advanced by one per step (no division by the ring depth)
  struct Pos { int st; uint32_t par; __device__ __forceinline__ void next() { if (++st == STAGES) { st = 0; par ^= 1u; } } };
  Pos pp = {walk.stage(ta) - 1, 0u}, pc = {walk.stage(ta), 0u}, pn = {walk.stage(ta), 0u};   // ta - first is 0 or 1: all in the first lap
  pn.next();
  auto sptr = [&](const Pos& p) { return reinterpret_cast<int16_t*>(smem + p.st * STAGE_BYTES); };
  CtuPrm cp = load_ctu_prm(g, sd, plane, ta, by0);
  for (int tx = ta; tx < tb; tx++) {
    const CtuPrm cp_next = load_ctu_prm(g, sd, plane, min(tx + 1, tb - 1), by0);   // consumed by the next step
    if (tx == ta) {
      if (tx > walk.first) ring::mbar_wait(&full[pp.st], pp.par);
      ring::mbar_wait(&full[pc.st], pc.par);
    }
    if (tx + 1 <= walk.last) ring::mbar_wait(&full[pn.st], pn.par);
    sao_tile(g, sd, plane, dst, tx, by0, sptr(pc), tx > 0 ? sptr(pp) : nullptr, tx + 1 < ntx ? sptr(pn) : nullptr, cp);
    cp = cp_next;
    __syncthreads();  // every thread is done with tile tx - 1: its stage can be refilled
    if (tid == 0 && tx - 1 >= walk.first && tx - 1 + STAGES <= walk.last) issue(tx - 1 + STAGES);
    pp = pc; pc = pn; pn.next();
  }
}

}  // namespace

void launch_sao(const Geom& g, const SlotDev* slots, int first_slot, int num_slots, const BatchCtl& ctl, cudaStream_t st) {
  static bool attr_set[64] = {};
  static const int smem = SMEM_BYTES + env_int("ILF_SAO_SMEM_PAD");
  once_per_device(attr_set, [&] { cudaFuncSetAttribute(sao_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
  const int bands_y = (g.rows + BR - 1) / BR, bands_c = (g.rows / 2 + BR - 1) / BR;
  // enough CTAs to fill the machine when the batch is small: split bands into horizontal segments
  const int ntx = (g.width + TW - 1) / TW;
  // bands that really run: a plane whose SAO is off in the whole picture is skipped (its CTAs leave at once)
  int bands = 0;
  for (int z = 0; z < num_slots; z++)
    bands += (ctl_skip(ctl.v[z], 0) ? 0 : bands_y) + (ctl_skip(ctl.v[z], 1) ? 0 : bands_c) + (ctl_skip(ctl.v[z], 2) ? 0 : bands_c);
  if (bands < 1) bands = 1;
  int nseg = (148 * SAO_CTAS + bands - 1) / bands;
  nseg = nseg < 1 ? 1 : (nseg > ntx ? ntx : nseg);
  static const int force = env_int("ILF_SAO_NSEG");   // experiment knob
  if (force > 0) nseg = force < ntx ? force : ntx;
  dim3 grid(nseg, bands_y + 2 * bands_c, num_slots);
  launch_pdl(sao_kernel, grid, dim3(NTHREADS), smem, st, g, slots, first_slot, ctl, bands_y, bands_c, nseg);
}

}  // namespace ilf

// ilf_sao_stats.cu -- encoder-side SAO statistics on the device-resident deblocked picture (sm_100a).
//
// Replaces EncSampleAdaptiveOffset::getStatistics (EncoderLib/EncSampleAdaptiveOffset.cpp:278-331) and its per-block worker
// getBlkStats (:1122-1487) for the path the encoder takes without SaoCtuBoundary: for every CTU, component and SAO type
// (EO 0/90/135/45 degrees, BO) the number of samples per class and the sum of (original - deblocked) per class.  The
// samples of a CTU block that contribute depend on the type, on the neighbour-CTU availability and on the skip lines of
// createEncData (:122-128): the last 5 (chroma 3) columns and 4 (2) rows of a CTU are left out when the right / below CTU
// exists; the first row of the two diagonal types has its own column range (:1278-1287, :1360-1372).
//
// One CTA per (CTU, component); a thread walks the block's samples with stride 256.  Every sample's four edge classes are
// formed from its 3x3 neighbourhood (coordinates clamped to the picture: a class that would look outside the picture is
// never inside its type's region).  Edge-offset sums stay in registers (4 types x 5 classes, selected without indexing),
// band-offset sums go to a per-warp histogram in shared memory; a warp reduction and one shared-memory atomic per class and
// warp later the CTA writes the 5 x 64 int64 words of SAOStatData.  Sums of one CTU fit 32 bits (16384 samples x 4095).
// Bound: HBM reads of two pictures (6 B per luma pixel); the output is 7.7 KB per CTU.
#include "ilf_common.cuh"

namespace ilf {
namespace {

constexpr int NT = 256;

__device__ __forceinline__ int sgn3(int a, int b) { return (a > b) - (a < b); }

__global__ void __launch_bounds__(NT) sao_stats_kernel(Geom g, const SlotDev* __restrict__ slots, int first_slot, BatchCtl bc) {
  __shared__ int bo_cnt[NT / 32][32], bo_dif[NT / 32][32];
  __shared__ int eo_cnt[4][5], eo_dif[4][5];
  pdl_launch_dependents();
  const SlotDev& sd = slots[first_slot + bc.slot[blockIdx.z]];
  const unsigned ctl = bc.v[blockIdx.z];
  const int comp = blockIdx.y, ctu = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int cx = ctu % g.ctus_w, cy = ctu / g.ctus_w;
  const int sh = comp ? 1 : 0, csz = 1 << g.ctu_log2;
  const int lx0 = cx << g.ctu_log2, ly0 = cy << g.ctu_log2;
  const int pw = g.width >> sh, ph = g.height >> sh, pitch = comp ? g.pitch_c : g.pitch_y;
  const int x0 = lx0 >> sh, y0 = ly0 >> sh;
  const int w = min(csz, g.width - lx0) >> sh, h = min(csz, g.height - ly0) >> sh;
  const unsigned av = sd.stats_avail[ctu];
  const bool L = av & ILF_AVAIL_L, A = av & ILF_AVAIL_A, AL = av & ILF_AVAIL_AL;
  const bool R = lx0 + csz < g.width, B = ly0 + csz < g.height, AR = ly0 > 0 && R;
  const int skip_r = comp ? 3 : 5, skip_b = comp ? 2 : 4;
  // column / row ranges per type (getBlkStats, isCalculatePreDeblockSamples == false)
  const int sx_e = L ? 0 : 1, ex_e = R ? w - skip_r : w - 1;       // EO 0 / 135 / 45 columns
  const int ey_0 = B ? h - skip_b : h;                              // EO 0 and BO rows
  const int ex_90 = R ? w - skip_r : w, sy_90 = A ? 0 : 1, ey_d = B ? h - skip_b : h - 1;  // EO 90 columns; rows of 90 / 135 / 45
  const int fs_135 = AL ? 0 : 1, fe_135 = A ? ex_e : 1;             // first row of EO 135
  const int fs_45 = A ? sx_e : ex_e, fe_45 = (!R && AR) ? w : ex_e;  // first row of EO 45
  const int bo_shift = (comp ? g.bd_chroma : g.bd_luma) - 5;

  for (int i = tid; i < (NT / 32) * 32; i += NT) { (&bo_cnt[0][0])[i] = 0; (&bo_dif[0][0])[i] = 0; }
  if (tid < 20) { (&eo_cnt[0][0])[tid] = 0; (&eo_dif[0][0])[tid] = 0; }
  __syncthreads();
  pdl_wait();  // the deblocking kernel has written the picture read from here on

  const int16_t* __restrict__ src = sd.buf[ctl_src(ctl, comp)][comp];
  const int16_t* __restrict__ org = sd.org[comp];
  int cnt[4][5], dif[4][5];
#pragma unroll
  for (int t = 0; t < 4; t++)
#pragma unroll
    for (int c = 0; c < 5; c++) cnt[t][c] = dif[t][c] = 0;

  for (int i = tid; i < w * h; i += NT) {
    const int y = i / w, x = i - y * w;
    const int gx = x0 + x, gy = y0 + y;
    const int xl = max(gx - 1, 0), xr = min(gx + 1, pw - 1);
    const int16_t* rm = src + (size_t)gy * pitch;
    const int16_t* ru = src + (size_t)max(gy - 1, 0) * pitch;
    const int16_t* rd = src + (size_t)min(gy + 1, ph - 1) * pitch;
    const int c = rm[gx];
    const int d = (int)org[(size_t)gy * pitch + gx] - c;
    int cls[4];
    cls[0] = 2 + sgn3(c, rm[xl]) + sgn3(c, rm[xr]);
    cls[1] = 2 + sgn3(c, ru[gx]) + sgn3(c, rd[gx]);
    cls[2] = 2 + sgn3(c, ru[xl]) + sgn3(c, rd[xr]);
    cls[3] = 2 + sgn3(c, ru[xr]) + sgn3(c, rd[xl]);
    bool in[4];
    in[0] = y < ey_0 && x >= sx_e && x < ex_e;
    in[1] = y >= sy_90 && y < ey_d && x < ex_90;
    in[2] = y < ey_d && (y == 0 ? (x >= fs_135 && x < fe_135) : (x >= sx_e && x < ex_e));
    in[3] = y < ey_d && (y == 0 ? (x >= fs_45 && x < fe_45) : (x >= sx_e && x < ex_e));
#pragma unroll
    for (int t = 0; t < 4; t++)
#pragma unroll
      for (int k = 0; k < 5; k++) {
        const bool hit = in[t] && cls[t] == k;
        cnt[t][k] += hit;
        dif[t][k] += hit ? d : 0;
      }
    if (y < ey_0 && x < ex_90) {  // BO: same columns as EO 90, same rows as EO 0
      const int b = c >> bo_shift;
      atomicAdd(&bo_cnt[warp][b], 1);
      atomicAdd(&bo_dif[warp][b], d);
    }
  }
#pragma unroll
  for (int t = 0; t < 4; t++)
#pragma unroll
    for (int k = 0; k < 5; k++) {
      int a = cnt[t][k], b = dif[t][k];
#pragma unroll
      for (int o = 16; o; o >>= 1) { a += __shfl_xor_sync(0xFFFFFFFFu, a, o); b += __shfl_xor_sync(0xFFFFFFFFu, b, o); }
      if (lane == 0) { atomicAdd(&eo_cnt[t][k], a); atomicAdd(&eo_dif[t][k], b); }
    }
  __syncthreads();
  // SAOStatData layout: [type][diff[32], count[32]] int64
  long long* out = sd.stats + ((size_t)ctu * 3 + comp) * (5 * 64);
  for (int i = tid; i < 5 * 64; i += NT) {
    const int t = i >> 6, k = i & 31, is_cnt = (i >> 5) & 1;
    int v = 0;
    if (t < 4) { if (k < 5) v = is_cnt ? eo_cnt[t][k] : eo_dif[t][k]; }
    else {
#pragma unroll
      for (int wv = 0; wv < NT / 32; wv++) v += is_cnt ? bo_cnt[wv][k] : bo_dif[wv][k];
    }
    out[i] = v;
  }
}

}  // namespace

void launch_sao_stats(const Geom& g, const SlotDev* slots, int first_slot, int num_slots, const BatchCtl& ctl, cudaStream_t st) {
  launch_pdl(sao_stats_kernel, dim3(g.ctus_w * g.ctus_h, 3, num_slots), dim3(NT), 0, st, g, slots, first_slot, ctl);
}

}  // namespace ilf

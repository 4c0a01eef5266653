// ilf_alf_stats.cu -- encoder ALF statistics on the device-resident picture (sm_100a).
//
// Replaces EncAdaptiveLoopFilter::deriveStatsForFiltering / getBlkStats / calcCovariance
// (source/Lib/EncoderLib/EncAdaptiveLoopFilter.cpp:1317-1514), the largest in-loop cost of the encoder: per CTU and class the
// covariance E[k][l] of the 13 symmetric tap sums of the reconstruction around every sample (7x7 shape), their correlation y[k]
// with (original - reconstruction) and its energy.  The reference accumulates doubles that hold exact integers below 2^53;
// int64 here is bit-equivalent.  Only the 7x7 shape is computed for luma: the 5x5 shape's taps are a subset of it under every
// transposition, so its statistics are rows / columns {2, 5, 6, 7, 10, 11, 12} of the 7x7 record (the host shim slices them).
//
// One thread = one 4x4 block (one class, one transposition): the picture strip is staged in shared memory with its border
// replicated (recYuv.extendBorderPel, :252), the thread walks its 16 samples in the block's TRANSPOSED coordinate frame, so the
// same code serves all four transpositions, forms the 13 tap sums of a sample and adds their 91 + 13 + 1 products to int32
// registers (16 samples of a 12-bit picture stay below 2^31); the 105 sums of the block then go to the CTU's record of its
// class with 64-bit reductions in L2 (no ordering issue: integer adds).  The multiplies bound it: 105 IMAD per sample.
#include "ilf_common.cuh"

namespace ilf {
namespace {

constexpr int ST_W = 128, ST_H = 32;   // luma samples a CTA covers (chroma: the same numbers on a chroma plane)
constexpr int NT = (ST_W / 4) * (ST_H / 4);

template <int HALF>
struct Shape {
  static constexpr int N = HALF == 3 ? 13 : 7;
  static constexpr int WORDS = N * (N + 1) / 2 + N + 1;
};

// canonical tap offsets (a, b) of coefficient k, transposeIdx 0: rows -HALF .. -1 left to right, then the left half of the centre row
template <int HALF>
__device__ __forceinline__ void tap_ab(int k, int& a, int& b) {
  int i = 0;
#pragma unroll
  for (int bb = -HALF; bb <= 0; bb++)
#pragma unroll
    for (int aa = -(HALF + bb); aa <= (bb < 0 ? HALF + bb : -1); aa++, i++)
      if (i == k) { a = aa; b = bb; }
}

template <int HALF, bool CLASSES>
__global__ void __launch_bounds__(NT) alf_stats_kernel(Geom g, const SlotDev* __restrict__ slots, int first_slot, BatchCtl bc, int plane) {
  constexpr int N = Shape<HALF>::N, WORDS = Shape<HALF>::WORDS;
  constexpr int TP = ST_W + 2 * HALF + 2;  // tile pitch (samples)
  __shared__ int16_t tile[(ST_H + 2 * HALF) * TP];
  const SlotDev& sd = slots[first_slot + bc.slot[blockIdx.z]];
  const unsigned ctl = bc.v[blockIdx.z];
  const int sh = plane ? 1 : 0;
  const int w = g.width >> sh, h = g.height >> sh, pitch = plane ? g.pitch_c : g.pitch_y;
  const int16_t* __restrict__ rec = sd.buf[ctl_src(ctl, plane)][plane];
  const int16_t* __restrict__ org = sd.org[plane];
  const int x0 = blockIdx.x * ST_W, y0 = blockIdx.y * ST_H;
  // strip with HALF replicated samples around it (clamped reads = the reference's border extension)
  for (int i = threadIdx.x; i < (ST_H + 2 * HALF) * (ST_W + 2 * HALF); i += NT) {
    const int ty = i / (ST_W + 2 * HALF), tx = i - ty * (ST_W + 2 * HALF);
    const int y = min(max(y0 + ty - HALF, 0), h - 1), x = min(max(x0 + tx - HALF, 0), w - 1);
    tile[ty * TP + tx] = rec[(size_t)y * pitch + x];
  }
  __syncthreads();
  const int bj = threadIdx.x % (ST_W / 4), bi = threadIdx.x / (ST_W / 4);
  const int bx = x0 + 4 * bj, by = y0 + 4 * bi;
  if (bx >= w || by >= h) return;
  int cls = 0;
  if (CLASSES) cls = sd.alf_class[(size_t)(by >> 2) * g.units_w + (bx >> 2)];
  const int t = cls >> 5;
  // the block's transposed frame: sample (u, v) sits at P0 + u ex + v ey; transposeIdx 0: ex = (1, 0), ey = (0, 1); 1: (0, 1), (1, 0);
  // 2: (-1, 0), (0, 1) from the block's right edge; 3: (0, -1), (1, 0) from its bottom edge (EncAdaptiveLoopFilter.cpp:1463-1510)
  const int exx = t == 0 ? 1 : (t == 2 ? -1 : 0), exy = t == 1 ? 1 : (t == 3 ? -1 : 0);
  const int eyx = (t & 1) ? 1 : 0, eyy = (t & 1) ? 0 : 1;
  const int px0 = 4 * bj + (t == 2 ? 3 : 0), py0 = 4 * bi + (t == 3 ? 3 : 0);   // block-relative origin inside the strip
  const int sx = exx + exy * TP, sy = eyx + eyy * TP;
  const int c00 = (py0 + HALF) * TP + px0 + HALF;
  int off[N - 1];
#pragma unroll
  for (int k = 0; k < N - 1; k++) { int a = 0, b = 0; tap_ab<HALF>(k, a, b); off[k] = a * sx + b * sy; }
  const int16_t* org0 = org + (size_t)(y0 + py0) * pitch + x0 + px0;
  const int osx = exx + exy * pitch, osy = eyx + eyy * pitch;

  int acc[WORDS];
#pragma unroll
  for (int i = 0; i < WORDS; i++) acc[i] = 0;
#pragma unroll 1
  for (int s = 0; s < 16; s++) {
    const int u = s & 3, v = s >> 2;
    const int c = c00 + u * sx + v * sy;
    int e[N];
#pragma unroll
    for (int k = 0; k < N - 1; k++) e[k] = tile[c + off[k]] + tile[c - off[k]];
    e[N - 1] = tile[c];
    const int yl = org0[u * osx + v * osy] - e[N - 1];
    int i = 0;
#pragma unroll
    for (int k = 0; k < N; k++)
#pragma unroll
      for (int l = k; l < N; l++) acc[i++] += e[k] * e[l];
#pragma unroll
    for (int k = 0; k < N; k++) acc[i++] += e[k] * yl;
    acc[i] += yl * yl;
  }
  const int ctu_log2 = g.ctu_log2 - sh;
  long long* out = sd.alf_stats + (size_t)((by >> ctu_log2) * g.ctus_w + (bx >> ctu_log2)) * ILF_ALF_STATS_WORDS +
                   (plane == 0 ? (size_t)(cls & 31) * WORDS : (size_t)(25 * 105 + (plane - 1) * 36));
#pragma unroll
  for (int i = 0; i < WORDS; i++) atomicAdd(reinterpret_cast<unsigned long long*>(out + i), (unsigned long long)(long long)acc[i]);
}

}  // namespace

// The slots' class maps (SlotDev::alf_class) must be current: launch_alf_classify first.  The records must be zero.
void launch_alf_stats(const Geom& g, const SlotDev* slots, int first_slot, int num_slots, const BatchCtl& ctl, cudaStream_t st) {
  alf_stats_kernel<3, true><<<dim3((g.width + ST_W - 1) / ST_W, (g.height + ST_H - 1) / ST_H, num_slots), NT, 0, st>>>(g, slots, first_slot, ctl, 0);
  const dim3 gc((g.width / 2 + ST_W - 1) / ST_W, (g.height / 2 + ST_H - 1) / ST_H, num_slots);
  alf_stats_kernel<2, false><<<gc, NT, 0, st>>>(g, slots, first_slot, ctl, 1);
  alf_stats_kernel<2, false><<<gc, NT, 0, st>>>(g, slots, first_slot, ctl, 2);
}

}  // namespace ilf

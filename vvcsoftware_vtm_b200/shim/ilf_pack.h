// ilf_pack.h -- host-side packer: VTM 2.1 data model -> plain side-information arrays of include/ilf_b200.h.
//
// Product host code (C++11, compiled against the reference's headers, which stay untouched).  It is the
// only place that touches CodingStructure / CodingUnit / TransformUnit / MotionInfo / Slice; everything
// after it (C ABI, kernels, the oracle restatement) sees flat arrays only.
//
// What is replayed here, by POSITION (not by CU list order, SURVEY.md section 7 "hard parts"):
//   * LoopFilter::xDeblockCU edge marking      LoopFilter.cpp:243-284, 313-354  (edge / TU-edge / walked bits)
//   * LoopFilter::xSetLoopfilterParam          LoopFilter.cpp:394-417           (slice boundary, disable flag)
//   * inputs of xGetBoundaryStrengthSingle     LoopFilter.cpp:419-541           (intra, cbf, ref picture ids, MVs)
//   * SampleAdaptiveOffset::xReconstructBlkSAOParams / getMergeList / invertQuantOffsets
//                                              SampleAdaptiveOffset.cpp:147-289
//   * deriveLoopFilterBoundaryAvailibility     SampleAdaptiveOffset.cpp:685-760
//   * AdaptiveLoopFilter::reconstructCoeff     AdaptiveLoopFilter.cpp:141-194
#ifndef ILF_PACK_H
#define ILF_PACK_H

#include <cstddef>
#include <cstdint>
#include <new>
#include <vector>

#include "ilf_b200.h"

class CodingStructure;
struct SAOBlkParam;
struct AlfSliceParam;

// Where the big side-information arrays live.  The packer itself is plain C++ (it is also linked into tools that have no CUDA
// library); the host shim points these at ilf_host_alloc / ilf_host_free before the first picture, so that the arrays sit in
// page-locked memory the packer owns for its whole life: the library then copies them to the device straight from there
// (ilf_b200.h "transfer pipeline"), nothing is registered or unregistered per picture, and a vector that grows simply gets a
// new page-locked block.  Set once, before the first ilfPackDeblock call.
struct IlfPackMemory
{
  static void* ( *alloc )( size_t );
  static void ( *release )( void* );
};
template<class T> struct IlfPackAlloc
{
  typedef T value_type;
  IlfPackAlloc() {}
  template<class U> IlfPackAlloc( const IlfPackAlloc<U>& ) {}
  T* allocate( size_t n )
  {
    void* p = IlfPackMemory::alloc( n * sizeof( T ) );
    if( !p ) throw std::bad_alloc();
    return static_cast<T*>( p );
  }
  void deallocate( T* p, size_t ) { IlfPackMemory::release( p ); }
  template<class U> bool operator==( const IlfPackAlloc<U>& ) const { return true; }
  template<class U> bool operator!=( const IlfPackAlloc<U>& ) const { return false; }
};

struct IlfPackedDeblock
{
  int                   unitsW = 0, unitsH = 0, ctusW = 0, ctusH = 0;
  ilf_deblock_params    params;
  std::vector<uint32_t, IlfPackAlloc<uint32_t> > info;        // luma-tree layer
  std::vector<uint32_t, IlfPackAlloc<uint32_t> > infoChroma;  // chroma-tree layer, empty when the picture has no dual-tree slice
  bool                  wantMv32 = true;  // in: also fill mv32 (the shim asks for it only after a picture did not fit 16 bits)
  std::vector<int32_t, IlfPackAlloc<int32_t> >  mv32;        // 4 per unit (when wantMv32)
  std::vector<int16_t, IlfPackAlloc<int16_t> >  mv16;        // same values as int16, valid while mvFits16
  bool                  mvFits16 = true;
  bool                  anyInter = false;
  std::vector<uint8_t>  ctuSlice;
};

struct IlfPackedSao
{
  std::vector<ilf_sao_ctu> ctus;
  bool                     anyEnabled = false;  // false -> SAOProcess returns early (SampleAdaptiveOffset.cpp:572-583)
};

struct IlfPackedAlf
{
  ilf_alf_params       params;
  std::vector<uint8_t> ctuEnable;  // [3][numCtus]
  bool                 enabled = false;  // false -> ALFProcess returns early (AdaptiveLoopFilter.cpp:70-73)
};

// Deblocking grid of the whole picture.
void ilfPackDeblock( CodingStructure& cs, IlfPackedDeblock& out );

// SAO: resolves merge candidates and scales the offsets IN `saoBlkParams` exactly as the reference does
// (the array is left in the same "reconstructed" state SAOProcess leaves it in), then flattens it.
void ilfPackSao( CodingStructure& cs, SAOBlkParam* saoBlkParams, const uint32_t offsetStepLog2[3], IlfPackedSao& out );

// ALF: reconstructs the coefficients (mutating alfSliceParam like the reference) and flattens.
void ilfReconstructAlfCoeff( AlfSliceParam& alfSliceParam, bool isLuma, short* coeffFinal /* [25*13], luma only */, bool redo );
// encoder = true: called after EncAdaptiveLoopFilter::alfEncoder, which has already run reconstructCoeff with bRedo on the enabled
// channels (EncAdaptiveLoopFilter.cpp:431): the luma deltas are put back the same way and a disabled channel is left untouched,
// so alfSliceParam reaches the bitstream writer exactly as the reference leaves it.
void ilfPackAlf( CodingStructure& cs, AlfSliceParam& alfSliceParam, IlfPackedAlf& out, bool encoder = false );


// Encoder SAO statistics of cs's picture through the shim's context (ilf_shim.cpp): `src` is the deblocked picture the
// encoder measures (uploaded only when the device does not already hold it from loopFilterPic), `org` the source picture,
// ctuAvail the ILF_AVAIL_L/_A/_AL flags per CTU; out[numCtus][3][5][64] in SAOStatData layout (diff[32], count[32]).
struct ilfPlanes { const int16_t* p[3]; ptrdiff_t stride[3]; };
void ilfShimSaoStatistics( CodingStructure& cs, const ilfPlanes& org, const ilfPlanes& src, const uint8_t* ctuAvail, int64_t* out );
// Encoder ALF statistics and block classification of `rec` (the SAO'd picture) against `org`: out[numCtus][ILF_ALF_STATS_WORDS],
// classMap[unitsH][unitsW] = classIdx | transposeIdx << 5.
void ilfShimAlfStatistics( CodingStructure& cs, const ilfPlanes& org, const ilfPlanes& rec, int64_t* out, uint8_t* classMap );
// Encoder ALF application: filters the picture ilfShimAlfStatistics left in the slot with the parameters alfEncoder chose and
// writes the result into cs.getRecoBuf() (what the m_filter7x7Blk / m_filter5x5Blk loops of EncAdaptiveLoopFilter.cpp:433-462 do).
void ilfShimAlfApply( CodingStructure& cs, AlfSliceParam& alfSliceParam );

#endif

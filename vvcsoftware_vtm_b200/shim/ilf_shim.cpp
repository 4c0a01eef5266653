// ilf_shim.cpp -- host shim: the three picture-level entry points of VTM 2.1's in-loop filter classes, re-implemented
// on top of the C ABI of libilf_b200.so (include/ilf_b200.h).
//
//   LoopFilter::loopFilterPic            replaces source/Lib/CommonLib/LoopFilter.cpp:149-230
//   SampleAdaptiveOffset::SAOProcess     replaces source/Lib/CommonLib/SampleAdaptiveOffset.cpp:564-612
//   AdaptiveLoopFilter::ALFProcess       replaces source/Lib/CommonLib/AdaptiveLoopFilter.cpp:68-139
//
// Product host code, compiled against the reference's UNMODIFIED headers.  The class declarations, create()/destroy(),
// the static tables and every protected helper the encoder subclasses use (offsetCTU, getMergeList, m_filter7x7Blk ...)
// stay the reference's own code; only these three function bodies are swapped (INTEGRATION.md shows the CMake option
// and the three #if blocks a maintainer adds; the demonstration build in oracle/Makefile obtains the same effect
// without touching the reference sources by compiling its three .cpp files with the entry points renamed).
//
// Data flow per picture (DecLib::executeLoopFilters, DecLib.cpp:506-533): loopFilterPic packs the CU/TU/motion grid,
// uploads the reconstructed picture once and runs the deblocking kernel; SAOProcess and ALFProcess continue on the
// device-resident picture; the last stage that the SPS enables downloads the result into cs.getRecoBuf().  On the
// encoder (cs.pcv->isEncoder) every stage downloads, because the RDO code reads the intermediate pictures on the host.
// There is no CPU fallback: a CUDA/library error is THROWn like any other reference error (TypeDef.h:1187-1209).
//
// Environment: ILF_B200_DEVICE=<ordinal> (default 0), ILF_TIMING=1 prints "[ILFTIME] ..." per picture on stderr.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <utility>
#include <mutex>
#include <vector>

#include "CommonLib/AdaptiveLoopFilter.h"
#include "CommonLib/CodingStructure.h"
#include "CommonLib/LoopFilter.h"
#include "CommonLib/Picture.h"
#include "CommonLib/SampleAdaptiveOffset.h"
#include "CommonLib/UnitTools.h"
#include "ilf_b200.h"
#include "ilf_pack.h"

namespace
{
using clk = std::chrono::steady_clock;

struct ShimState
{
  ilf_ctx*       ctx = nullptr;
  ilf_config     cfg;
  const Picture* resident = nullptr;  // picture whose current state lives on the device and not (yet) in the host reco buffer
  const Picture* mirrored = nullptr;  // picture whose host reco buffer equals the slot's current device state (just downloaded) ...
  int            mirroredPoc = -1;    // ... and its POC (Picture objects are recycled)
  bool           timing   = false;
  long long      usDeblock = 0, usSao = 0, usAlf = 0;
  int            picCount = 0;
  ~ShimState()
  {
    if( ctx ) ilf_destroy( ctx );
  }
};

// The shim keeps one context and one resident picture per process; its entry points are serialised, so that several decoder or
// encoder objects in one process cannot interleave inside them (they take turns).
std::recursive_mutex& shimMutex()
{
  static std::recursive_mutex m;
  return m;
}

ShimState& state()
{
  static ShimState s;
  static bool      init = false;
  if( !init )
  {
    init     = true;
    s.timing = getenv( "ILF_TIMING" ) && atoi( getenv( "ILF_TIMING" ) ) != 0;
    IlfPackMemory::alloc   = ilf_host_alloc;   // the packer's arrays in page-locked memory (ilf_pack.h)
    IlfPackMemory::release = ilf_host_free;
  }
  return s;
}

void ck( ShimState& s, int rc, const char* what )
{
  if( rc != ILF_OK ) THROW( "libilf_b200: " << what << " failed (" << rc << "): " << ilf_last_error( s.ctx ) );
}

// One context per geometry; the decoder re-creates the filter objects for every picture (DecLib.cpp:747-748,789), the
// context survives that.
ShimState& contextFor( const CodingStructure& cs )
{
  ShimState&           s   = state();
  const PreCalcValues& pcv = *cs.pcv;
  CHECK( pcv.chrFormat != CHROMA_420, "libilf_b200 supports 4:2:0 only" );
  ilf_config c;
  c.width            = int( pcv.lumaWidth );
  c.height           = int( pcv.lumaHeight );
  c.bit_depth_luma   = cs.sps->getBitDepth( CHANNEL_TYPE_LUMA );
  c.bit_depth_chroma = cs.sps->getBitDepth( CHANNEL_TYPE_CHROMA );
  c.ctu_log2         = int( pcv.maxCUWidthLog2 );
  c.chroma_format    = 1;
  c.device           = getenv( "ILF_B200_DEVICE" ) ? atoi( getenv( "ILF_B200_DEVICE" ) ) : 0;
  c.num_slots        = 1;
  CHECK( pcv.maxCUWidth != pcv.maxCUHeight, "square CTUs expected" );
  if( s.ctx && ( c.width != s.cfg.width || c.height != s.cfg.height || c.bit_depth_luma != s.cfg.bit_depth_luma || c.bit_depth_chroma != s.cfg.bit_depth_chroma ||
                 c.ctu_log2 != s.cfg.ctu_log2 ) )
  {
    ilf_destroy( s.ctx );
    s.ctx      = nullptr;
    s.resident = nullptr;
  }
  if( !s.ctx )
  {
    const int rc = ilf_create( &s.ctx, &c );
    if( rc != ILF_OK ) THROW( "libilf_b200: ilf_create failed (" << rc << "): " << ilf_last_error( nullptr ) );
    s.cfg = c;
  }
  return s;
}

// The reference's picture planes are pageable and are handed to the library as they are: ilf_upload / ilf_download stage them
// through the library's own page-locked buffers with a few copy threads (ilf_api.cu CopyPool).  Page-locking the planes of the
// decoded-picture buffer instead (cudaHostRegister) was tried: 20 - 300 ms per plane set on the boxes measured, and a
// registration outlives the Picture it was made for (Picture::destroy, a new CVS), so the shim registers nothing.
void upload( ShimState& s, CodingStructure& cs )
{
  const CPelUnitBuf reco = cs.getRecoBuf();
  const CPelBuf     y = reco.get( COMPONENT_Y ), cb = reco.get( COMPONENT_Cb ), cr = reco.get( COMPONENT_Cr );
  ck( s, ilf_upload( s.ctx, 0, y.buf, y.stride, cb.buf, cb.stride, cr.buf, cr.stride ), "ilf_upload" );
  s.resident = cs.picture;
  s.mirrored = nullptr;
}

// In the decoder the filtered picture is final, and its next use is as a reference picture with replicated borders
// (Slice.cpp:389-413 call Picture::extendPicBorder): the download writes the margins too and marks the picture as extended, so
// that the reference's own pass over the picture is skipped.  The encoder keeps modifying / re-filtering the picture: plain download.
void download( ShimState& s, CodingStructure& cs )
{
  PelUnitBuf reco = cs.getRecoBuf();
  PelBuf     y = reco.get( COMPONENT_Y ), cb = reco.get( COMPONENT_Cb ), cr = reco.get( COMPONENT_Cr );
  static const bool extend = !( getenv( "ILF_SHIM_EXTEND" ) && atoi( getenv( "ILF_SHIM_EXTEND" ) ) == 0 );
  if( extend && !cs.pcv->isEncoder && ( cs.picture->margin & 1 ) == 0 )
  {
    ck( s, ilf_download_extended( s.ctx, 0, y.buf, y.stride, cb.buf, cb.stride, cr.buf, cr.stride, int( cs.picture->margin ) ), "ilf_download_extended" );
    cs.picture->m_bIsBorderExtended = true;
  }
  else
    ck( s, ilf_download( s.ctx, 0, y.buf, y.stride, cb.buf, cb.stride, cr.buf, cr.stride ), "ilf_download" );
  s.resident    = nullptr;
  s.mirrored    = cs.picture;
  s.mirroredPoc = cs.slice->getPOC();
}

long long usSince( clk::time_point t0 ) { return std::chrono::duration_cast<std::chrono::microseconds>( clk::now() - t0 ).count(); }

void report( ShimState& s, const CodingStructure& cs )
{
  if( s.timing )
    fprintf( stderr, "[ILFTIME] pic=%d poc=%d w=%u h=%u deblock_us=%lld sao_us=%lld alf_us=%lld impl=b200\n", s.picCount, cs.slice->getPOC(), cs.pcv->lumaWidth,
             cs.pcv->lumaHeight, s.usDeblock, s.usSao, s.usAlf );
  s.picCount++;
  s.usDeblock = s.usSao = s.usAlf = 0;
}
}  // namespace

// ------------------------------------------------------------------------------------------------------------
void LoopFilter::loopFilterPic( CodingStructure& cs )
{
  std::lock_guard<std::recursive_mutex> lock( shimMutex() );
  const auto t0 = clk::now();
  ShimState& s  = contextFor( cs );
  const auto tc = clk::now();
  // the picture goes up on a helper thread while this one packs (the library is not entered by anybody else meanwhile)
  int        upRc = ILF_OK;
  struct Joiner { std::thread t; ~Joiner() { if( t.joinable() ) t.join(); } } up;   // (the packer may THROW)
  up.t = std::thread( [&] {
    const CPelUnitBuf reco = cs.getRecoBuf();
    const CPelBuf     y = reco.get( COMPONENT_Y ), cb = reco.get( COMPONENT_Cb ), cr = reco.get( COMPONENT_Cr );
    upRc = ilf_upload( s.ctx, 0, y.buf, y.stride, cb.buf, cb.stride, cr.buf, cr.stride );
  } );
  static IlfPackedDeblock db;  // kept between pictures: the arrays keep their (page-locked) storage
  db.wantMv32 = false;
  ilfPackDeblock( cs, db );  // the walk over CUs/TUs/motion of LoopFilter.cpp:167-222, 243-541, flattened
  if( db.anyInter && !db.mvFits16 )
  {
    db.wantMv32 = true;  // a motion vector beyond 16 bits: walk again for the 32-bit array
    ilfPackDeblock( cs, db );
  }
  const auto tp = clk::now();
  up.t.join();
  ck( s, upRc, "ilf_upload" );
  s.resident = cs.picture;
  s.mirrored = nullptr;
  const auto tu = clk::now();
  ck( s, ilf_set_deblock_info( s.ctx, 0, &db.params, db.info.data(), db.infoChroma.empty() ? nullptr : db.infoChroma.data(),
                               ( db.anyInter && db.mvFits16 ) ? db.mv16.data() : nullptr, ( db.anyInter && !db.mvFits16 ) ? db.mv32.data() : nullptr, db.ctuSlice.data() ),
      "ilf_set_deblock_info" );
  ck( s, ilf_deblock( s.ctx, 0 ), "ilf_deblock" );
  const auto tk = clk::now();
  const bool laterStage = !cs.pcv->isEncoder && ( cs.sps->getUseSAO() || cs.sps->getUseALF() );
  if( !laterStage ) download( s, cs );
  s.usDeblock = usSince( t0 );
  if( s.timing && getenv( "ILF_TIMING" ) && atoi( getenv( "ILF_TIMING" ) ) > 1 )
    fprintf( stderr, "[ILFTIME2] context_us=%lld pack_us=%lld upload_us=%lld set+launch_us=%lld rest_us=%lld\n",
             (long long) std::chrono::duration_cast<std::chrono::microseconds>( tc - t0 ).count(), (long long) std::chrono::duration_cast<std::chrono::microseconds>( tp - tc ).count(),
             (long long) std::chrono::duration_cast<std::chrono::microseconds>( tu - tp ).count(), (long long) std::chrono::duration_cast<std::chrono::microseconds>( tk - tu ).count(), usSince( tk ) );
  if( !laterStage ) report( s, cs );
}

// ------------------------------------------------------------------------------------------------------------
void SampleAdaptiveOffset::SAOProcess( CodingStructure& cs, SAOBlkParam* saoBlkParams )
{
  std::lock_guard<std::recursive_mutex> lock( shimMutex() );
  CHECK( !saoBlkParams, "No parameters present" );
  const auto t0 = clk::now();
  ShimState& s  = contextFor( cs );
  CHECK( cs.sps->getPCMFilterDisableFlag() || cs.pps->getTransquantBypassEnabledFlag(),
         "libilf_b200: PCM loop-filter-disable / transquant-bypass sample restoration (SampleAdaptiveOffset.cpp:614-683) is not supported" );
  IlfPackedSao ps;
  ilfPackSao( cs, saoBlkParams, m_offsetStepLog2, ps );  // xReconstructBlkSAOParams: merge resolution + offset scaling, in place like the reference
  const bool laterStage = !cs.pcv->isEncoder && cs.sps->getUseALF();
  if( ps.anyEnabled )
  {
    if( s.resident != cs.picture ) upload( s, cs );
    ck( s, ilf_set_sao_params( s.ctx, 0, ps.ctus.data() ), "ilf_set_sao_params" );
    ck( s, ilf_sao( s.ctx, 0 ), "ilf_sao" );
  }
  if( !laterStage && s.resident == cs.picture ) download( s, cs );
  s.usSao = usSince( t0 );
  if( !laterStage ) report( s, cs );
}

// ------------------------------------------------------------------------------------------------------------
void AdaptiveLoopFilter::ALFProcess( CodingStructure& cs, AlfSliceParam& alfSliceParam )
{
  std::lock_guard<std::recursive_mutex> lock( shimMutex() );
  const auto t0 = clk::now();
  ShimState& s  = contextFor( cs );
  if( alfSliceParam.enabledFlag[COMPONENT_Y] || alfSliceParam.enabledFlag[COMPONENT_Cb] || alfSliceParam.enabledFlag[COMPONENT_Cr] )
  {
    alfSliceParam.filterShapes = m_filterShapes;
    m_clpRngs                  = cs.slice->getClpRngs();
    IlfPackedAlf pa;
    ilfPackAlf( cs, alfSliceParam, pa );  // reconstructCoeff luma + chroma (mutates alfSliceParam like the reference) + CTU flags
    if( s.resident != cs.picture ) upload( s, cs );
    ck( s, ilf_set_alf_params( s.ctx, 0, &pa.params, pa.ctuEnable.data() ), "ilf_set_alf_params" );
    ck( s, ilf_alf( s.ctx, 0 ), "ilf_alf" );
  }
  if( s.resident == cs.picture ) download( s, cs );
  s.usAlf = usSince( t0 );
  report( s, cs );
}

// ------------------------------------------------------------------------------------------------------------
// Encoder SAO statistics (called by EncSampleAdaptiveOffset::SAOProcess in ilf_shim_enc.cpp).  EncGOP::compressGOP calls
// loopFilterPic and then the SAO search on the same picture (EncGOP.cpp:2122-2137): the deblocked picture is still in the
// slot, so only the source picture crosses PCIe here.
void ilfShimSaoStatistics( CodingStructure& cs, const ilfPlanes& org, const ilfPlanes& src, const uint8_t* ctuAvail, int64_t* out )
{
  std::lock_guard<std::recursive_mutex> lock( shimMutex() );
  const auto t0 = clk::now();
  ShimState& s  = contextFor( cs );
  if( !( s.mirrored == cs.picture && s.mirroredPoc == cs.slice->getPOC() ) )
  {
    ck( s, ilf_upload( s.ctx, 0, src.p[0], src.stride[0], src.p[1], src.stride[1], src.p[2], src.stride[2] ), "ilf_upload" );
    s.resident = nullptr;
    s.mirrored = nullptr;
  }
  ck( s, ilf_set_original( s.ctx, 0, org.p[0], org.stride[0], org.p[1], org.stride[1], org.p[2], org.stride[2], ctuAvail ), "ilf_set_original" );
  ck( s, ilf_sao_stats( s.ctx, 0, 1 ), "ilf_sao_stats" );
  ck( s, ilf_get_sao_stats( s.ctx, 0, out ), "ilf_get_sao_stats" );
  if( s.timing ) fprintf( stderr, "[ILFTIME] poc=%d sao_stats_us=%lld impl=b200\n", cs.slice->getPOC(), usSince( t0 ) );
}

// Encoder ALF statistics (called by EncAdaptiveLoopFilter::ALFProcess in ilf_shim_enc.cpp): the SAO'd picture and the source
// picture go up, 21 KB of integers per CTU and one byte per 4x4 block come back.
void ilfShimAlfStatistics( CodingStructure& cs, const ilfPlanes& org, const ilfPlanes& rec, int64_t* out, uint8_t* classMap )
{
  std::lock_guard<std::recursive_mutex> lock( shimMutex() );
  const auto t0 = clk::now();
  ShimState& s  = contextFor( cs );
  ck( s, ilf_upload( s.ctx, 0, rec.p[0], rec.stride[0], rec.p[1], rec.stride[1], rec.p[2], rec.stride[2] ), "ilf_upload" );
  s.resident = nullptr;
  s.mirrored = nullptr;
  std::vector<uint8_t> avail( cs.pcv->sizeInCtus, 0 );
  ck( s, ilf_set_original( s.ctx, 0, org.p[0], org.stride[0], org.p[1], org.stride[1], org.p[2], org.stride[2], avail.data() ), "ilf_set_original" );
  ck( s, ilf_alf_stats( s.ctx, 0, 1 ), "ilf_alf_stats" );
  ck( s, ilf_alf_classify( s.ctx, 0, classMap ), "ilf_alf_classify" );
  ck( s, ilf_get_alf_stats( s.ctx, 0, out ), "ilf_get_alf_stats" );
  if( s.timing ) fprintf( stderr, "[ILFTIME] poc=%d alf_stats_us=%lld impl=b200\n", cs.slice->getPOC(), usSince( t0 ) );
}

// Encoder ALF application (called by EncAdaptiveLoopFilter::ALFProcess in ilf_shim_enc.cpp after the filter search): the SAO'd
// picture is still in the slot from the statistics pass, so nothing goes up but the parameters; the filtered picture comes down
// into the reconstruction buffer.  A plane whose ALF is off keeps the uploaded samples, as in the reference (the reconstruction
// buffer already holds them).
void ilfShimAlfApply( CodingStructure& cs, AlfSliceParam& alfSliceParam )
{
  std::lock_guard<std::recursive_mutex> lock( shimMutex() );
  const auto t0 = clk::now();
  ShimState& s  = contextFor( cs );
  IlfPackedAlf pa;
  ilfPackAlf( cs, alfSliceParam, pa, true );
  if( !pa.enabled ) return;
  ck( s, ilf_set_alf_params( s.ctx, 0, &pa.params, pa.ctuEnable.data() ), "ilf_set_alf_params" );
  ck( s, ilf_alf( s.ctx, 0 ), "ilf_alf" );
  PelUnitBuf reco = cs.getRecoBuf();
  PelBuf     y = reco.get( COMPONENT_Y ), cb = reco.get( COMPONENT_Cb ), cr = reco.get( COMPONENT_Cr );
  ck( s, ilf_download( s.ctx, 0, y.buf, y.stride, cb.buf, cb.stride, cr.buf, cr.stride ), "ilf_download" );
  if( s.timing ) fprintf( stderr, "[ILFTIME] poc=%d alf_apply_us=%lld impl=b200\n", cs.slice->getPOC(), usSince( t0 ) );
}

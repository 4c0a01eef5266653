// oracle/capture_hook.cpp -- TEST INFRASTRUCTURE (golden-vector capture + CPU baseline timer), not product.
//
// Replaces exactly ONE function of the unmodified reference decoder: DecLib::executeLoopFilters
// (source/Lib/DecoderLib/DecLib.cpp:506-533).  oracle/Makefile compiles the reference's DecLib.cpp a second
// time with -DexecuteLoopFilters=executeLoopFilters_stock (so the stock body keeps existing under another
// name) and links this file's definition in its place.  The three filter classes that run are the
// REFERENCE's own (LoopFilter, SampleAdaptiveOffset, AdaptiveLoopFilter from libvtm.a); this hook only
//   * packs the side information with the product packer (vvcsoftware_vtm_b200/shim/ilf_pack.cpp),
//   * dumps the picture before deblocking and after each stage,
//   * times the three reference calls with steady_clock (SURVEY.md 8d "CPU baseline timing").
//
// Environment:
//   ILF_CAPTURE_DIR=<dir>   write <dir>/pic_%04d.ilfcap per picture (container format: tools/ilfcap.py)
//   ILF_CAPTURE_MAX=<n>     stop dumping after n pictures (default: all)
//   ILF_CAPTURE_PLANES=0    side information only (no sample planes)
//   ILF_TIMING=1            print "[ILFTIME] ..." per picture on stderr
//   ILF_EXIT_AFTER=<n>      exit(0) after the filters of the n-th picture (bounded CPU-baseline samples)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "DecoderLib/DecLib.h"
#include "CommonLib/UnitTools.h"
#include "ilf_pack.h"
#include "capture_common.h"

namespace
{
bool sameSao( const SAOBlkParam& a, const SAOBlkParam& b )
{
  for( int c = 0; c < 3; c++ )
  {
    if( a[c].modeIdc != b[c].modeIdc ) return false;
    if( a[c].modeIdc == SAO_MODE_OFF ) continue;
    if( a[c].typeIdc != b[c].typeIdc || a[c].typeAuxInfo != b[c].typeAuxInfo ) return false;
    if( memcmp( a[c].offset, b[c].offset, sizeof( a[c].offset ) ) ) return false;
  }
  return true;
}
}  // namespace

void DecLib::executeLoopFilters()
{
  if( !m_pcPic ) return;
  CodingStructure& cs = *m_pcPic->cs;

  static int         picCount   = 0;
  static const char* capDir     = getenv( "ILF_CAPTURE_DIR" );
  static const int   capMax     = getenv( "ILF_CAPTURE_MAX" ) ? atoi( getenv( "ILF_CAPTURE_MAX" ) ) : ( 1 << 30 );
  static const bool  capPlanes  = !( getenv( "ILF_CAPTURE_PLANES" ) && atoi( getenv( "ILF_CAPTURE_PLANES" ) ) == 0 );
  static const bool  timing     = getenv( "ILF_TIMING" ) && atoi( getenv( "ILF_TIMING" ) ) != 0;
  const bool         dump       = capDir && picCount < capMax;
  const bool         useSAO     = cs.sps->getUseSAO();
  const bool         useALF     = cs.sps->getUseALF();
  const PreCalcValues& pcv      = *cs.pcv;

  CapWriter w;
  if( dump )
  {
    char name[64];
    snprintf( name, sizeof( name ), "/pic_%04d.ilfcap", picCount );
    if( !w.open( std::string( capDir ) + name ) ) { fprintf( stderr, "capture: cannot open output in %s\n", capDir ); exit( 2 ); }

    IlfPackedDeblock db;
    ilfPackDeblock( cs, db );
    capWriteDeblockInfo( w, db );
    if( capPlanes ) w.planes( "pre", cs.getRecoBuf() );
  }

  using clk = std::chrono::steady_clock;
  if( getenv( "ILF_PACK_TIMING" ) )  // how long does the product packer take on this picture? (host-side cost of the drop-in)
  {
    static IlfPackedDeblock db;
    db.wantMv32 = false;
    const auto p0 = clk::now();
    ilfPackDeblock( cs, db );
    fprintf( stderr, "[PACKTIME] poc=%d pack_us=%lld\n", cs.slice->getPOC(), (long long) std::chrono::duration_cast<std::chrono::microseconds>( clk::now() - p0 ).count() );
  }
  const auto t0 = clk::now();
  m_cLoopFilter.loopFilterPic( cs );
  const auto t1 = clk::now();
  if( dump && capPlanes ) w.planes( "dbk", cs.getRecoBuf() );

  long long saoUs = 0, alfUs = 0;
  if( useSAO )
  {
    IlfPackedSao ps;
    if( dump )
    {
      std::vector<SAOBlkParam> copy( cs.picture->getSAO(), cs.picture->getSAO() + pcv.sizeInCtus );
      const uint32_t           steps[3] = { cs.pps->getPpsRangeExtension().getLog2SaoOffsetScale( CHANNEL_TYPE_LUMA ),
                                            cs.pps->getPpsRangeExtension().getLog2SaoOffsetScale( CHANNEL_TYPE_CHROMA ),
                                            cs.pps->getPpsRangeExtension().getLog2SaoOffsetScale( CHANNEL_TYPE_CHROMA ) };
      ilfPackSao( cs, copy.data(), steps, ps );
      const auto ts0 = clk::now();
      m_cSAO.SAOProcess( cs, cs.picture->getSAO() );
      saoUs = std::chrono::duration_cast<std::chrono::microseconds>( clk::now() - ts0 ).count();
      for( unsigned i = 0; i < pcv.sizeInCtus; i++ )
        if( !sameSao( copy[i], cs.picture->getSAO()[i] ) ) { fprintf( stderr, "capture: SAO parameter resolution differs from the reference at CTU %u\n", i ); exit( 3 ); }
      const uint32_t sd[3] = { uint32_t( ps.ctus.size() ), uint32_t( sizeof( ilf_sao_ctu ) ), 1 };
      w.rec( "sao_ctus", 0, 2, sd, ps.ctus.data(), ps.ctus.size() * sizeof( ilf_sao_ctu ) );
      const int32_t  any   = ps.anyEnabled;
      const uint32_t od[3] = { 1, 1, 1 };
      w.rec( "sao_any", 2, 1, od, &any, 4 );
      if( capPlanes ) w.planes( "sao", cs.getRecoBuf() );
    }
    else
    {
      const auto ts0 = clk::now();
      m_cSAO.SAOProcess( cs, cs.picture->getSAO() );
      saoUs = std::chrono::duration_cast<std::chrono::microseconds>( clk::now() - ts0 ).count();
    }
  }
  if( useALF )
  {
    if( dump )
    {
      AlfSliceParam copy = cs.slice->getAlfSliceParam();
      IlfPackedAlf  pa;
      ilfPackAlf( cs, copy, pa );
      const uint32_t ad[3] = { uint32_t( sizeof( pa.params ) ), 1, 1 };
      w.rec( "alf_params", 0, 1, ad, &pa.params, sizeof( pa.params ) );
      const uint32_t ed[3] = { 3, uint32_t( pcv.sizeInCtus ), 1 };
      w.rec( "alf_ctu_enable", 0, 2, ed, pa.ctuEnable.data(), pa.ctuEnable.size() );
      const int32_t  en    = pa.enabled;
      const uint32_t od[3] = { 1, 1, 1 };
      w.rec( "alf_enabled", 2, 1, od, &en, 4 );
    }
    const auto ta0 = clk::now();
    m_cALF.ALFProcess( cs, cs.slice->getAlfSliceParam() );
    alfUs = std::chrono::duration_cast<std::chrono::microseconds>( clk::now() - ta0 ).count();
    if( dump && capPlanes ) w.planes( "alf", cs.getRecoBuf() );
  }
  const long long dbUs = std::chrono::duration_cast<std::chrono::microseconds>( t1 - t0 ).count();

  if( dump )
  {
    const int32_t geom[16] = { int32_t( pcv.lumaWidth ), int32_t( pcv.lumaHeight ), cs.sps->getBitDepth( CHANNEL_TYPE_LUMA ), cs.sps->getBitDepth( CHANNEL_TYPE_CHROMA ),
                               int32_t( pcv.maxCUWidthLog2 ), cs.slice->getPOC(), int32_t( cs.slice->getSliceType() ), useSAO, useALF, CS::isDualITree( cs ),
                               int32_t( dbUs ), int32_t( saoUs ), int32_t( alfUs ), int32_t( cs.picture->slices.size() ), 0, 0 };
    const uint32_t gd[3]   = { 16, 1, 1 };
    w.rec( "geom", 2, 1, gd, geom, sizeof( geom ) );
    w.close();
  }
  if( timing )
    fprintf( stderr, "[ILFTIME] pic=%d poc=%d w=%u h=%u deblock_us=%lld sao_us=%lld alf_us=%lld\n", picCount, cs.slice->getPOC(), pcv.lumaWidth, pcv.lumaHeight, dbUs, saoUs, alfUs );
  picCount++;
  static const int exitAfter = getenv( "ILF_EXIT_AFTER" ) ? atoi( getenv( "ILF_EXIT_AFTER" ) ) : 0;
  if( exitAfter > 0 && picCount >= exitAfter ) { fflush( stderr ); _Exit( 0 ); }
}

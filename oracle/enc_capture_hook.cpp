// oracle/enc_capture_hook.cpp -- TEST INFRASTRUCTURE (golden-vector capture on the ENCODER side), not product.
//
// The reference decoder cannot parse its own encoder's multi-slice streams (BitStream.cpp:399), so side information that only
// such pictures carry -- several slices with their own beta / tc offsets, LFCrossSliceBoundaryFlag = 0, chroma QP offsets --
// is pinned where the reference ENCODER filters its reconstruction: EncGOP::compressGOP calls LoopFilter::loopFilterPic
// (EncGOP.cpp:2122).  oracle/Makefile compiles the reference's LoopFilter.cpp with the entry point renamed to
// loopFilterPic_reference (the same object the drop-in build uses) and links this definition of LoopFilter::loopFilterPic in
// its place: it packs the side information with the product packer, dumps the picture, runs the REFERENCE's own deblocking
// and dumps the result.  Everything else of EncoderApp is the unmodified reference.
//
// Environment: ILF_CAPTURE_DIR=<dir> (pic_%04d.ilfcap per loopFilterPic call), ILF_CAPTURE_MAX=<n>.
#include <cstdlib>
#include <string>

// declare the renamed stock entry point next to the real one (the class layout does not change)
#define loopFilterPic( x ) loopFilterPic_reference( x ); void loopFilterPic( x )
#include "CommonLib/LoopFilter.h"
#undef loopFilterPic
#include "CommonLib/CodingStructure.h"
#include "CommonLib/Picture.h"
#include "CommonLib/UnitTools.h"
#include "capture_common.h"

void LoopFilter::loopFilterPic( CodingStructure& cs )
{
  static int         picCount = 0;
  static const char* capDir   = getenv( "ILF_CAPTURE_DIR" );
  static const int   capMax   = getenv( "ILF_CAPTURE_MAX" ) ? atoi( getenv( "ILF_CAPTURE_MAX" ) ) : ( 1 << 30 );
  const bool         dump     = capDir && picCount < capMax;
  CapWriter          w;
  if( dump )
  {
    char name[64];
    snprintf( name, sizeof( name ), "/pic_%04d.ilfcap", picCount );
    if( !w.open( std::string( capDir ) + name ) ) { fprintf( stderr, "capture: cannot open output in %s\n", capDir ); exit( 2 ); }
    IlfPackedDeblock db;
    ilfPackDeblock( cs, db );
    capWriteDeblockInfo( w, db );
    w.planes( "pre", cs.getRecoBuf() );
  }
  loopFilterPic_reference( cs );
  if( dump )
  {
    w.planes( "dbk", cs.getRecoBuf() );
    const PreCalcValues& pcv = *cs.pcv;
    const int32_t geom[16] = { int32_t( pcv.lumaWidth ), int32_t( pcv.lumaHeight ), cs.sps->getBitDepth( CHANNEL_TYPE_LUMA ), cs.sps->getBitDepth( CHANNEL_TYPE_CHROMA ),
                               int32_t( pcv.maxCUWidthLog2 ), cs.slice->getPOC(), int32_t( cs.slice->getSliceType() ), 0, 0, CS::isDualITree( cs ),
                               0, 0, 0, int32_t( cs.picture->slices.size() ), 0, 0 };
    const uint32_t gd[3]   = { 16, 1, 1 };
    w.rec( "geom", 2, 1, gd, geom, sizeof( geom ) );
    w.close();
  }
  picCount++;
}

// oracle/ref_units.cpp -- TEST INFRASTRUCTURE: extern "C" doors onto the REFERENCE's own block functions so that
// tests/test_oracle_units.py can drive them with random blocks, offsets and availability flags that real
// bitstreams rarely produce (multi-slice availability patterns, extreme sample values):
//   SampleAdaptiveOffset::offsetBlock                 SampleAdaptiveOffset.cpp:292-508
//   AdaptiveLoopFilter::deriveClassificationBlk       AdaptiveLoopFilter.cpp:292-463 (scalar) / x86 SIMD
//   AdaptiveLoopFilter::filterBlk<5|7>                AdaptiveLoopFilter.cpp:465-650 (scalar) / x86 SIMD
//   EncSampleAdaptiveOffset::getBlkStats              EncoderLib/EncSampleAdaptiveOffset.cpp:1122-1487
// Linked against oracle/_ref/libvtm.a (unmodified reference objects).  Nothing here is product code.
#include <cstring>
#include <vector>

#include "CommonLib/AdaptiveLoopFilter.h"
#include "CommonLib/SampleAdaptiveOffset.h"
#include "CommonLib/CodingStructure.h"
#include "CommonLib/Picture.h"
#include "CommonLib/SEI.h"
#include "EncoderLib/CABACWriter.h"
// getBlkStats and the skip-line tables are private members of the reference class: open them for this test door only
// (every header the class needs is already included above, so nothing else is parsed under the define)
#define private public
#include "EncoderLib/EncSampleAdaptiveOffset.h"
#include "EncoderLib/EncAdaptiveLoopFilter.h"
#undef private

namespace
{
struct SaoDoor : public SampleAdaptiveOffset
{
  void run( int bd, int type, int* offset, const Pel* src, Pel* dst, int ss, int ds, int w, int h, unsigned av )
  {
    if( m_signLineBuf1.size() < size_t( w + 2 ) ) { m_signLineBuf1.resize( w + 2 ); m_signLineBuf2.resize( w + 2 ); }
    ClpRng clp; clp.min = 0; clp.max = ( 1 << bd ) - 1; clp.bd = bd; clp.n = 0;
    offsetBlock( bd, clp, type, offset, src, dst, ss, ds, w, h, av & 1, av & 2, av & 4, av & 8, av & 16, av & 32, av & 64, av & 128 );
  }
};
}  // namespace

extern "C" {

// offset32: the 32-entry offset array of SAOOffset (EO: entries 0..4, BO: per band).  avail: ILF_AVAIL_* bits
// (L=1, R=2, A=4, B=8, AL=16, AR=32, BL=64, BR=128).  src/dst point at the block's first sample inside a larger
// picture (the function reads one sample beyond the block on every side).
int ref_sao_offset_block( int bit_depth, int type, const int* offset32, const int16_t* src, int16_t* dst, int src_stride, int dst_stride, int w, int h, unsigned avail )
{
  static SaoDoor door;
  int off[MAX_NUM_SAO_CLASSES];
  memcpy( off, offset32, sizeof( off ) );
  door.run( bit_depth, type, off, src, dst, src_stride, dst_stride, w, h, avail );
  return 0;
}

// luma points at sample (0,0) of a plane that is readable 3 samples beyond [0,w)x[0,h) (caller pads).  out[h/4][w/4]
// = classIdx | transposeIdx << 5.  simd != 0 uses the x86 function the reference installs at start-up.
int ref_alf_classify( int simd, const int16_t* luma, int stride, int w, int h, int bit_depth, uint8_t* out )
{
  static AdaptiveLoopFilter simdAlf;  // constructor installs the SIMD pointers (AdaptiveLoopFilter.cpp:57-65)
  std::vector<AlfClassifier>  store( size_t( w ) * h );
  std::vector<AlfClassifier*> rows( h );
  for( int y = 0; y < h; y++ ) rows[y] = &store[size_t( y ) * w];
  int  lapStore[NUM_DIRECTIONS][37][37];
  int* lapRows[NUM_DIRECTIONS][37];
  int** lap[NUM_DIRECTIONS];
  for( int d = 0; d < NUM_DIRECTIONS; d++ ) { for( int y = 0; y < 37; y++ ) lapRows[d][y] = lapStore[d][y]; lap[d] = lapRows[d]; }
  CPelBuf src( luma, stride, w, h );
  for( int y = 0; y < h; y += 32 )
    for( int x = 0; x < w; x += 32 )
    {
      const Area blk( x, y, std::min( 32, w - x ), std::min( 32, h - y ) );
      if( simd ) simdAlf.m_deriveClassificationBlk( rows.data(), lap, src, blk, bit_depth + 4 );
      else AdaptiveLoopFilter::deriveClassificationBlk( rows.data(), lap, src, blk, bit_depth + 4 );
    }
  for( int y = 0; y < h; y += 4 )
    for( int x = 0; x < w; x += 4 ) out[size_t( y / 4 ) * ( w / 4 ) + x / 4] = uint8_t( rows[y][x].classIdx | ( rows[y][x].transposeIdx << 5 ) );
  return 0;
}

// One plane through filterBlk.  cls (luma only) = output of ref_alf_classify; coeff = 25x13 (luma) or 7 (chroma) shorts.
int ref_alf_filter( int simd, int is7, int chroma, const int16_t* src, int src_stride, int16_t* dst, int dst_stride, int w, int h, int bit_depth, const uint8_t* cls, const short* coeff )
{
  static AdaptiveLoopFilter simdAlf;
  std::vector<AlfClassifier>  store( chroma ? 1 : size_t( w ) * h );
  std::vector<AlfClassifier*> rows( chroma ? 1 : h );
  if( !chroma )
    for( int y = 0; y < h; y++ )
    {
      rows[y] = &store[size_t( y ) * w];
      for( int x = 0; x < w; x++ ) { const uint8_t c = cls[size_t( y / 4 ) * ( w / 4 ) + x / 4]; rows[y][x] = AlfClassifier( c & 31, c >> 5 ); }
    }
  std::vector<short> cf( coeff, coeff + ( chroma ? 7 : 25 * 13 ) );
  ClpRng clp; clp.min = 0; clp.max = ( 1 << bit_depth ) - 1; clp.bd = bit_depth; clp.n = 0;
  // the function indexes recDst/recSrc by compId; give it the same plane in every slot
  PelBuf  d( dst, dst_stride, w, h );
  CPelBuf s( src, src_stride, w, h );
  PelUnitBuf  dstU( CHROMA_420, d, d, d );
  CPelUnitBuf srcU( CHROMA_420, s, s, s );
  const Area blk( 0, 0, w, h );
  const ComponentID comp = chroma ? COMPONENT_Cb : COMPONENT_Y;
  if( simd ) { if( is7 ) simdAlf.m_filter7x7Blk( rows.data(), dstU, srcU, blk, comp, cf.data(), clp ); else simdAlf.m_filter5x5Blk( rows.data(), dstU, srcU, blk, comp, cf.data(), clp ); }
  else { if( is7 ) AdaptiveLoopFilter::filterBlk<ALF_FILTER_7>( rows.data(), dstU, srcU, blk, comp, cf.data(), clp ); else AdaptiveLoopFilter::filterBlk<ALF_FILTER_5>( rows.data(), dstU, srcU, blk, comp, cf.data(), clp ); }
  return 0;
}

// One CTU block of one component through the reference's own statistics function, skip lines as createEncData sets them
// without SaoCtuBoundary (EncSampleAdaptiveOffset.cpp:122-128).  avail6: bit0 L, bit1 R, bit2 A, bit3 B, bit4 AL, bit5 AR.
// src/org point at the block's first sample inside larger pictures.  out[5][64]: diff[32] then count[32] per type.
int ref_sao_blk_stats( int is_chroma, int bit_depth, const int16_t* src, const int16_t* org, int src_stride, int org_stride, int w, int h, unsigned avail6, int64_t* out )
{
  struct EncDoor : public EncSampleAdaptiveOffset
  {
    void lineBufs( int w ) { if( m_signLineBuf1.size() < size_t( w + 2 ) ) { m_signLineBuf1.resize( w + 2 ); m_signLineBuf2.resize( w + 2 ); } }
  };
  static EncDoor* enc = nullptr;
  if( !enc )
  {
    enc = new EncDoor;
    enc->create( 256, 256, CHROMA_420, 128, 128, 4, 0, 0 );
    enc->createEncData( false, 4 );
  }
  enc->lineBufs( w );
  SAOStatData stats[NUM_SAO_NEW_TYPES];
  enc->getBlkStats( is_chroma ? COMPONENT_Cb : COMPONENT_Y, bit_depth, stats, const_cast<Pel*>( src ), const_cast<Pel*>( org ), src_stride, org_stride, w, h,
                    avail6 & 1, ( avail6 >> 1 ) & 1, ( avail6 >> 2 ) & 1, ( avail6 >> 3 ) & 1, ( avail6 >> 4 ) & 1, ( avail6 >> 5 ) & 1, false );
  for( int t = 0; t < NUM_SAO_NEW_TYPES; t++ )
  {
    memcpy( out + t * 64, stats[t].diff, sizeof( int64_t ) * 32 );
    memcpy( out + t * 64 + 32, stats[t].count, sizeof( int64_t ) * 32 );
  }
  return 0;
}
}

// The decoded-picture hashes the reference compares with the SEI (PicYuvMD5.cpp:91-175; the component functions have external linkage
// but no header) and Picture::extendPicBorder (Picture.cpp:996-1040) on a real Picture object.
uint32_t compCRC( int bitdepth, const Pel* plane, uint32_t width, uint32_t height, uint32_t stride, PictureHash& digest );
uint32_t compChecksum( int bitdepth, const Pel* plane, uint32_t width, uint32_t height, uint32_t stride, PictureHash& digest, const BitDepths& bitDepths );
extern "C" {
unsigned ref_plane_crc( const int16_t* plane, int stride, int w, int h, int bit_depth )
{
  PictureHash d;
  compCRC( bit_depth, plane, w, h, stride, d );
  return ( unsigned( d.hash[0] ) << 8 ) | d.hash[1];
}
unsigned ref_plane_checksum( const int16_t* plane, int stride, int w, int h, int bit_depth )
{
  PictureHash d;
  BitDepths   bd;
  bd.recon[0] = bd.recon[1] = bit_depth;
  compChecksum( bit_depth, plane, w, h, stride, d, bd );
  return ( unsigned( d.hash[0] ) << 24 ) | ( unsigned( d.hash[1] ) << 16 ) | ( unsigned( d.hash[2] ) << 8 ) | d.hash[3];
}
// in: three dense planes (4:2:0); out: the three planes WITH their margins (luma margin `margin`, chroma margin / 2), dense
int ref_extend_pic_border( const int16_t* y, const int16_t* cb, const int16_t* cr, int w, int h, int margin, int16_t* oy, int16_t* ocb, int16_t* ocr )
{
  Picture pic;
  pic.create( CHROMA_420, Size( w, h ), 128, margin, true );
  CodingStructure cs( g_globalUnitCache.cuCache, g_globalUnitCache.puCache, g_globalUnitCache.tuCache );
  cs.area = UnitArea( CHROMA_420, Area( 0, 0, w, h ) );
  pic.cs  = &cs;
  const int16_t* in[3]  = { y, cb, cr };
  int16_t*       out[3] = { oy, ocb, ocr };
  for( int c = 0; c < 3; c++ )
  {
    PelBuf b = pic.getRecoBuf().get( ComponentID( c ) );
    for( unsigned r = 0; r < b.height; r++ ) memcpy( b.buf + ptrdiff_t( r ) * b.stride, in[c] + size_t( r ) * b.width, b.width * sizeof( Pel ) );
  }
  pic.m_bIsBorderExtended = false;
  pic.extendPicBorder();
  for( int c = 0; c < 3; c++ )
  {
    PelBuf    b  = pic.getRecoBuf().get( ComponentID( c ) );
    const int mx = margin >> ( c ? 1 : 0 ), my = mx;
    const int ow = int( b.width ) + 2 * mx;
    for( int r = -my; r < int( b.height ) + my; r++ ) memcpy( out[c] + size_t( r + my ) * ow, b.buf + ptrdiff_t( r ) * b.stride - mx, ow * sizeof( Pel ) );
  }
  pic.cs = nullptr;
  pic.destroy();
  return 0;
}
}

// EncAdaptiveLoopFilter::getBlkStats (EncAdaptiveLoopFilter.cpp:1394-1440) of one block of a plane: rec / org point at sample (0, 0)
// of planes that have at least 3 samples of margin; cls = classIdx | transposeIdx << 5 per SAMPLE row-major over (w, h), or NULL
// (chroma: one class).  shape7 = 1: 7x7 (13 coefficients), 0: 5x5 (7).  out[classes][n n + n + 1] doubles as the reference holds them:
// E row-major (full symmetric matrix), y, pixAcc.
extern "C" int ref_alf_blk_stats( int shape7, const int16_t* org, int org_stride, const int16_t* rec, int rec_stride, int x0, int y0, int w, int h, const uint8_t* cls, int cls_stride,
                                  double* out )
{
  static EncAdaptiveLoopFilter* enc = new EncAdaptiveLoopFilter;
  const AlfFilterShape shape( shape7 ? 7 : 5 );
  const int            n       = shape.numCoeff;
  const int            classes = cls ? MAX_NUM_ALF_CLASSES : 1;
  std::vector<AlfCovariance> cov( classes );
  for( auto& c : cov ) { c.create( n ); c.reset(); }
  std::vector<std::vector<AlfClassifier>> rows;
  std::vector<AlfClassifier*>             ptrs;
  if( cls )
  {
    rows.resize( y0 + h );
    for( int y = 0; y < y0 + h; y++ )
    {
      rows[y].resize( x0 + w );
      for( int x = 0; x < x0 + w; x++ ) rows[y][x] = AlfClassifier( cls[y * cls_stride + x] & 31, cls[y * cls_stride + x] >> 5 );
      ptrs.push_back( rows[y].data() );
    }
  }
  const CompArea area( COMPONENT_Y, CHROMA_420, Area( x0, y0, w, h ) );
  enc->getBlkStats( cov.data(), shape, cls ? ptrs.data() : nullptr, const_cast<Pel*>( org ) + y0 * org_stride + x0, org_stride, const_cast<Pel*>( rec ) + y0 * rec_stride + x0,
                    rec_stride, area );
  for( int c = 0; c < classes; c++ )
  {
    double* o = out + size_t( c ) * ( n * n + n + 1 );
    for( int k = 0; k < n; k++ )
      for( int l = 0; l < n; l++ ) *o++ = cov[c].E[k][l];
    for( int k = 0; k < n; k++ ) *o++ = cov[c].y[k];
    *o = cov[c].pixAcc;
    cov[c].destroy();
  }
  return 0;
}

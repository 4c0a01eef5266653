// oracle/capture_common.h -- TEST INFRASTRUCTURE: the ILFCAP container writer shared by the decoder and encoder capture hooks
// (format: tools/ilfcap.py).
#ifndef ILF_CAPTURE_COMMON_H
#define ILF_CAPTURE_COMMON_H
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "CommonLib/Buffer.h"
#include "ilf_pack.h"

struct CapWriter
{
  FILE* f = nullptr;
  bool  open( const std::string& path )
  {
    f = fopen( path.c_str(), "wb" );
    if( f ) fwrite( "ILFCAP1\0", 1, 8, f );
    return f != nullptr;
  }
  // dtype: 0 u8, 1 i16, 2 i32, 3 u32
  void rec( const char* name, int dtype, int ndim, const uint32_t dims[3], const void* data, size_t bytes )
  {
    char nm[24];
    memset( nm, 0, sizeof( nm ) );
    strncpy( nm, name, 23 );
    fwrite( nm, 1, 24, f );
    uint8_t  hdr[8] = { uint8_t( dtype ), uint8_t( ndim ), 0, 0, 0, 0, 0, 0 };
    uint32_t d[3]   = { dims[0], ndim > 1 ? dims[1] : 1, ndim > 2 ? dims[2] : 1 };
    uint64_t nb     = bytes;
    fwrite( hdr, 1, 8, f );
    fwrite( d, 4, 3, f );
    fwrite( &nb, 8, 1, f );
    fwrite( data, 1, bytes, f );
  }
  void plane( const char* name, const CPelBuf& b )
  {
    std::vector<int16_t> tmp( size_t( b.width ) * b.height );
    for( unsigned y = 0; y < b.height; y++ ) memcpy( &tmp[size_t( y ) * b.width], b.buf + ptrdiff_t( y ) * b.stride, b.width * sizeof( int16_t ) );
    const uint32_t dims[3] = { b.height, b.width, 1 };
    rec( name, 1, 2, dims, tmp.data(), tmp.size() * 2 );
  }
  void planes( const char* prefix, const CPelUnitBuf& u )
  {
    static const char* sfx[3] = { "_y", "_cb", "_cr" };
    for( int c = 0; c < 3; c++ ) plane( ( std::string( prefix ) + sfx[c] ).c_str(), u.get( ComponentID( c ) ) );
  }
  void close()
  {
    if( f ) fclose( f );
    f = nullptr;
  }
};


// side information of the deblocking stage as the product packer produces it
inline void capWriteDeblockInfo( CapWriter& w, const IlfPackedDeblock& db )
{
  const uint32_t gdims[3] = { uint32_t( db.unitsH ), uint32_t( db.unitsW ), 4 };
  const uint32_t pd[3]    = { uint32_t( sizeof( db.params ) ), 1, 1 };
  w.rec( "db_params", 0, 1, pd, &db.params, sizeof( db.params ) );
  w.rec( "db_info", 3, 2, gdims, db.info.data(), db.info.size() * 4 );
  if( !db.infoChroma.empty() ) w.rec( "db_info_c", 3, 2, gdims, db.infoChroma.data(), db.infoChroma.size() * 4 );
  w.rec( "db_mv32", 2, 3, gdims, db.mv32.data(), db.mv32.size() * 4 );
  const uint32_t cd[3] = { uint32_t( db.ctusH ), uint32_t( db.ctusW ), 1 };
  w.rec( "ctu_slice", 0, 2, cd, db.ctuSlice.data(), db.ctuSlice.size() );
}
#endif

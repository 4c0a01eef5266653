// ilf_shim_enc.cpp -- host shim, encoder side: EncSampleAdaptiveOffset::SAOProcess with its statistics pass on the GPU.
//
//   EncSampleAdaptiveOffset::SAOProcess     replaces source/Lib/EncoderLib/EncSampleAdaptiveOffset.cpp:213-253
//     its call of getStatistics (:227, :278-331, getBlkStats :1122-1487) becomes ilf_set_original / ilf_sao_stats /
//     ilf_get_sao_stats on the deblocked picture that loopFilterPic left on the device.
//
// Everything else of the encoder's SAO search (decidePicParams, decideBlkParams with its RDO and the offsetCTU calls
// that apply the chosen offsets, xPCMLFDisableProcess) stays the reference's own code and runs on the host, as do the
// private helpers this function calls.  Linked into EncoderApp only (oracle/Makefile: EncoderApp_ilf_b200; the reference's
// EncSampleAdaptiveOffset.cpp is compiled with the entry point renamed, like the three decoder-side classes).
// SaoCtuBoundary (statistics on pre-deblocking samples, off in every reference cfg) is refused, not emulated.
#include <cstring>
#include <vector>

#include "CommonLib/CodingStructure.h"
#include "CommonLib/Picture.h"
#include "EncoderLib/EncSampleAdaptiveOffset.h"
#include "ilf_b200.h"
#include "ilf_pack.h"

#if K0238_SAO_GREEDY_MERGE_ENCODING
void EncSampleAdaptiveOffset::SAOProcess( CodingStructure& cs, bool* sliceEnabled, const double* lambdas, const bool bTestSAODisableAtPictureLevel, const double saoEncodingRate,
                                          const double saoEncodingRateChroma, bool isPreDBFSamplesUsed, bool isGreedymergeEncoding )
#else
void EncSampleAdaptiveOffset::SAOProcess( CodingStructure& cs, bool* sliceEnabled, const double* lambdas, const bool bTestSAODisableAtPictureLevel, const double saoEncodingRate,
                                          const double saoEncodingRateChroma, bool isPreDBFSamplesUsed )
#endif
{
  CHECK( isPreDBFSamplesUsed, "libilf_b200: SaoCtuBoundary (statistics on pre-deblocking samples) is not supported" );
  CHECK( cs.pcv->chrFormat != CHROMA_420, "libilf_b200 supports 4:2:0 only" );
  PelUnitBuf org = cs.getOrgBuf();
  PelUnitBuf res = cs.getRecoBuf();
  PelUnitBuf src = m_tempBuf;
  memcpy( m_lambda, lambdas, sizeof( m_lambda ) );
  src.copyFrom( res );  // the search below reads the deblocked picture from m_tempBuf and writes the SAO result into the reco buffer

  // ---- statistics: per CTU the three availability flags the reference derives (:305), everything else on the device ----
  const PreCalcValues& pcv = *cs.pcv;
  std::vector<uint8_t> avail( pcv.sizeInCtus );
  int ctuRsAddr = 0;
  for( uint32_t yPos = 0; yPos < pcv.lumaHeight; yPos += pcv.maxCUHeight )
    for( uint32_t xPos = 0; xPos < pcv.lumaWidth; xPos += pcv.maxCUWidth )
    {
      bool l, a, al;
      deriveLoopFilterBoundaryAvailibility( cs, Position( xPos, yPos ), l, a, al );
      avail[ctuRsAddr++] = uint8_t( ( l ? ILF_AVAIL_L : 0 ) | ( a ? ILF_AVAIL_A : 0 ) | ( al ? ILF_AVAIL_AL : 0 ) );
    }
  ilfPlanes po, ps;
  for( int c = 0; c < 3; c++ )
  {
    po.p[c] = org.get( ComponentID( c ) ).buf; po.stride[c] = org.get( ComponentID( c ) ).stride;
    ps.p[c] = src.get( ComponentID( c ) ).buf; ps.stride[c] = src.get( ComponentID( c ) ).stride;
  }
  std::vector<int64_t> words( size_t( pcv.sizeInCtus ) * 3 * ILF_SAO_STATS_WORDS );
  ilfShimSaoStatistics( cs, po, ps, avail.data(), words.data() );
  static_assert( sizeof( SAOStatData ) == 64 * sizeof( int64_t ), "SAOStatData = diff[32], count[32]" );
  for( uint32_t ctu = 0; ctu < pcv.sizeInCtus; ctu++ )
    for( int c = 0; c < 3; c++ )
      for( int t = 0; t < NUM_SAO_NEW_TYPES; t++ )
      {
        const int64_t* w = &words[( ( size_t( ctu ) * 3 + c ) * NUM_SAO_NEW_TYPES + t ) * 64];
        memcpy( m_statData[ctu][c][t].diff, w, sizeof( int64_t ) * 32 );
        memcpy( m_statData[ctu][c][t].count, w + 32, sizeof( int64_t ) * 32 );
      }

  // ---- decisions: the reference's own search ----
  decidePicParams( *cs.slice, sliceEnabled, saoEncodingRate, saoEncodingRateChroma );
  std::vector<SAOBlkParam> reconParams( cs.pcv->sizeInCtus );
#if K0238_SAO_GREEDY_MERGE_ENCODING
  decideBlkParams( cs, sliceEnabled, m_statData, src, res, &reconParams[0], cs.picture->getSAO(), bTestSAODisableAtPictureLevel, saoEncodingRate, saoEncodingRateChroma,
                   isGreedymergeEncoding );
#else
  decideBlkParams( cs, sliceEnabled, m_statData, src, res, &reconParams[0], cs.picture->getSAO(), bTestSAODisableAtPictureLevel, saoEncodingRate, saoEncodingRateChroma );
#endif
  xPCMLFDisableProcess( cs );
}

// ilf_shim_enc.cpp -- host shim, encoder side: the statistics passes of the encoder's SAO and ALF searches on the GPU.
//
//   EncAdaptiveLoopFilter::ALFProcess       replaces source/Lib/EncoderLib/EncAdaptiveLoopFilter.cpp:220-268
//     its deriveClassification (:257) and deriveStatsForFiltering (:260, :1317-1514) become ilf_upload / ilf_set_original /
//     ilf_alf_stats / ilf_alf_classify / ilf_get_alf_stats on the SAO'd picture; the filter derivation (alfEncoder) stays the
//     reference's own code and reads the class map and the covariances exactly as its own passes would have left them; the
//     per-CTU filter application at its end (:433-462) becomes ilf_set_alf_params / ilf_alf / ilf_download on the resident picture.
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
#include "EncoderLib/EncAdaptiveLoopFilter.h"
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

// ------------------------------------------------------------------------------------------------------------
void EncAdaptiveLoopFilter::ALFProcess( CodingStructure& cs, const double* lambdas, AlfSliceParam& alfSliceParam )
{
  CHECK( cs.pcv->chrFormat != CHROMA_420, "libilf_b200 supports 4:2:0 only" );
  // ---- set-up: as the reference (:222-252) ----
  alfSliceParam.filterShapes = m_filterShapes;
  m_clpRngs                  = cs.slice->getClpRngs();
  for( int compIdx = 0; compIdx < MAX_NUM_COMPONENT; compIdx++ ) m_ctuEnableFlag[compIdx] = cs.picture->getAlfCtuEnableFlag( compIdx );
  alfSliceParam.reset();
#if DISTORTION_LAMBDA_BUGFIX
  const int shiftLuma   = 2 * DISTORTION_PRECISION_ADJUSTMENT( m_inputBitDepth[CHANNEL_TYPE_LUMA] );
  const int shiftChroma = 2 * DISTORTION_PRECISION_ADJUSTMENT( m_inputBitDepth[CHANNEL_TYPE_CHROMA] );
#else
  const int shiftLuma   = 2 * DISTORTION_PRECISION_ADJUSTMENT( m_inputBitDepth[CHANNEL_TYPE_LUMA] - 8 );
  const int shiftChroma = 2 * DISTORTION_PRECISION_ADJUSTMENT( m_inputBitDepth[CHANNEL_TYPE_CHROMA] - 8 );
#endif
  m_lambda[COMPONENT_Y]  = lambdas[COMPONENT_Y] * double( 1 << shiftLuma );
  m_lambda[COMPONENT_Cb] = lambdas[COMPONENT_Cb] * double( 1 << shiftChroma );
  m_lambda[COMPONENT_Cr] = lambdas[COMPONENT_Cr] * double( 1 << shiftChroma );
  PelUnitBuf orgYuv = cs.getOrgBuf();
  // the reference copies the picture into m_tempBuf and pads it (:250-252) as the source of its CPU filters; here the source is
  // the copy in the slot (the device kernels pad on the fly), so the reconstruction buffer itself goes up
  PelUnitBuf recYuv = cs.getRecoBuf();

  // ---- classification + statistics on the device ----
  const PreCalcValues& pcv = *cs.pcv;
  ilfPlanes po, pr;
  for( int c = 0; c < 3; c++ )
  {
    po.p[c] = orgYuv.get( ComponentID( c ) ).buf; po.stride[c] = orgYuv.get( ComponentID( c ) ).stride;
    pr.p[c] = recYuv.get( ComponentID( c ) ).buf; pr.stride[c] = recYuv.get( ComponentID( c ) ).stride;
  }
  const int            unitsW = int( pcv.lumaWidth ) / 4, unitsH = int( pcv.lumaHeight ) / 4;
  std::vector<int64_t> words( size_t( pcv.sizeInCtus ) * ILF_ALF_STATS_WORDS );
  std::vector<uint8_t> cls( size_t( unitsW ) * unitsH );
  ilfShimAlfStatistics( cs, po, pr, words.data(), cls.data() );
  for( int by = 0; by < unitsH; by++ )
    for( int bx = 0; bx < unitsW; bx++ )
    {
      const AlfClassifier c( cls[size_t( by ) * unitsW + bx] & 31, cls[size_t( by ) * unitsW + bx] >> 5 );
      for( int y = 0; y < 4; y++ )
        for( int x = 0; x < 4; x++ ) m_classifier[4 * by + y][4 * bx + x] = c;
    }
  // a record = E upper triangle row-major, y, pixAcc of the 7x7 (luma) / 5x5 (chroma) shape; the luma 5x5 shape is a slice of it
  static const int in7[7] = { 2, 5, 6, 7, 10, 11, 12 };
  auto fill = [&]( AlfCovariance& cov, const int64_t* rec, int nRec, const int* map )
  {
    const int n = cov.numCoeff;
    for( int k = 0; k < n; k++ )
    {
      const int rk = map ? map[k] : k;
      for( int l = k; l < n; l++ )
      {
        const int rl = map ? map[l] : l;
        const int a = std::min( rk, rl ), b = std::max( rk, rl );
        cov.E[k][l] = cov.E[l][k] = double( rec[a * nRec - a * ( a - 1 ) / 2 + ( b - a )] );
      }
      cov.y[k] = double( rec[nRec * ( nRec + 1 ) / 2 + rk] );
    }
    cov.pixAcc = double( rec[nRec * ( nRec + 1 ) / 2 + nRec] );
  };
  for( int ch = 0; ch < 2; ch++ )
    for( size_t shape = 0; shape < m_filterShapes[ch].size(); shape++ )
      for( int classIdx = 0; classIdx < ( ch == 0 ? MAX_NUM_ALF_CLASSES : 1 ); classIdx++ ) m_alfCovarianceFrame[ch][shape][classIdx].reset();
  for( int ctu = 0; ctu < m_numCTUsInPic; ctu++ )
  {
    const int64_t* w = &words[size_t( ctu ) * ILF_ALF_STATS_WORDS];
    for( size_t shape = 0; shape < m_filterShapes[CHANNEL_TYPE_LUMA].size(); shape++ )
      for( int classIdx = 0; classIdx < MAX_NUM_ALF_CLASSES; classIdx++ )
      {
        AlfCovariance& cov = m_alfCovariance[COMPONENT_Y][shape][ctu][classIdx];
        fill( cov, w + classIdx * 105, 13, m_filterShapes[CHANNEL_TYPE_LUMA][shape].filterLength == 7 ? nullptr : in7 );
        m_alfCovarianceFrame[CHANNEL_TYPE_LUMA][shape][classIdx] += cov;
      }
    for( int c = 1; c < 3; c++ )
      for( size_t shape = 0; shape < m_filterShapes[CHANNEL_TYPE_CHROMA].size(); shape++ )
      {
        AlfCovariance& cov = m_alfCovariance[c][shape][ctu][0];
        fill( cov, w + 25 * 105 + ( c - 1 ) * 36, 7, nullptr );
        m_alfCovarianceFrame[CHANNEL_TYPE_CHROMA][shape][0] += cov;
      }
  }

  // ---- filter derivation: the reference's own search (:262-268).  alfEncoder ends with the reconstruction loops that call
  //      m_filter7x7Blk / m_filter5x5Blk per CTU (:433-462); those two are function pointers of the base class, pointed at a
  //      function that does nothing while the search runs -- the picture is filtered on the device afterwards ----
  const auto filter5 = m_filter5x5Blk, filter7 = m_filter7x7Blk;
  m_filter5x5Blk = m_filter7x7Blk = []( AlfClassifier**, const PelUnitBuf&, const CPelUnitBuf&, const Area&, const ComponentID, short*, const ClpRng& ) {};
  alfEncoder( cs, alfSliceParam, orgYuv, recYuv, cs.getRecoBuf(), CHANNEL_TYPE_LUMA );
  if( alfSliceParam.enabledFlag[COMPONENT_Y] ) alfEncoder( cs, alfSliceParam, orgYuv, recYuv, cs.getRecoBuf(), CHANNEL_TYPE_CHROMA );
  m_filter5x5Blk = filter5;
  m_filter7x7Blk = filter7;

  // ---- filter application on the device: the SAO'd picture is in the slot since the statistics pass ----
  ilfShimAlfApply( cs, alfSliceParam );
}

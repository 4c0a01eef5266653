// ilf_pack.cpp -- see ilf_pack.h.  Host C++11 against the reference's data model.
#include "ilf_pack.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <map>
#include <mutex>
#include <thread>

#include "CommonLib/CodingStructure.h"
#include "CommonLib/Picture.h"
#include "CommonLib/Slice.h"
#include "CommonLib/Unit.h"
#include "CommonLib/UnitTools.h"

void* ( *IlfPackMemory::alloc )( size_t ) = malloc;
void ( *IlfPackMemory::release )( void* )  = free;

namespace
{
// Per-picture scratch that plays the role of the reference's per-CTU arrays m_aapbEdgeFilter / m_aapucBS
// (LoopFilter.h:58-59).  CU areas are disjoint, so one picture-sized array per tree layer is equivalent to
// the reference's "memset per CTU, fill per CU" (LoopFilter.cpp:171-172).
struct EdgeScratch
{
  int                  unitsW, unitsH;
  std::vector<uint8_t> edge[2];  // [dir] m_aapbEdgeFilter
  std::vector<uint8_t> pre[2];   // [dir] value left in m_aapucBS by xSetEdgefilterMultiple
  void                 init( int w, int h )
  {
    unitsW = w;
    unitsH = h;
    for( int d = 0; d < 2; d++ )
    {
      edge[d].assign( size_t( w ) * h, 0 );
      pre[d].assign( size_t( w ) * h, 0 );
    }
  }
  // xSetEdgefilterMultiple (LoopFilter.cpp:372-393): first column (VER) / first row (HOR) of `a`.
  void mark( int dir, const Area& a, bool value, bool edgeIdx )
  {
    const int n  = dir == 0 ? a.height / 4 : a.width / 4;
    const int ux = a.x / 4, uy = a.y / 4;
    for( int i = 0; i < n; i++ )
    {
      const int x = dir == 0 ? ux : ux + i;
      const int y = dir == 0 ? uy + i : uy;
      if( x >= unitsW || y >= unitsH ) continue;
      edge[dir][size_t( y ) * unitsW + x] = value;
      if( !edgeIdx ) pre[dir][size_t( y ) * unitsW + x] = value;
    }
  }
};

// Read-only while the walk runs: the picture's slices with their dense ids and reference-picture ids [list][refIdx]
// (one map lookup per reference and slice, not per unit), filled serially before the CTU rows are dealt to the workers.
struct PackShared
{
  std::vector<const Slice*> slices;
  std::vector<uint32_t>     refTabs;  // [slice][2][MAX_NUM_REF]
  int find( const Slice* s ) const
  {
    for( size_t i = 0; i < slices.size(); i++ ) if( slices[i] == s ) return int( i );
    THROW( "ilf_b200: a coding unit refers to a slice that is not in Picture::slices" );
  }
};

struct PackCtx  // one per worker
{
  CodingStructure*  cs;
  IlfPackedDeblock* out;
  EdgeScratch*      scratch;     // [tree layer]; kept between pictures so that the arrays keep their pages
  const PackShared* sh;
  bool              pcmFilter, tqBypass, highPrecMv;
  bool              anyInter = false, mvFits16 = true;
  const Slice*      lastSlice = nullptr;
  int               lastSliceId = 0;
  int sliceId( const Slice* s )
  {
    if( s != lastSlice ) { lastSlice = s; lastSliceId = sh->find( s ); }
    return lastSliceId;
  }
  const uint32_t* refsOf( const Slice& s ) { return &sh->refTabs[size_t( sliceId( &s ) ) * 2 * MAX_NUM_REF]; }
};

// One call of LoopFilter::xDeblockCU for both edge directions, reduced to its effect on the grid.
void packCU( PackCtx& pc, const CodingUnit& cu, int layer )
{
  const PreCalcValues& pcv = *cu.cs->pcv;
  auto& info = layer == 0 ? pc.out->info : pc.out->infoChroma;
  EdgeScratch&          es   = pc.scratch[layer];
  const int             unitsW = pc.out->unitsW, unitsH = pc.out->unitsH;
  const bool            lumaValid = cu.Y().valid();
  // LoopFilter.cpp:246
  const Area area = lumaValid ? Area( cu.Y() )
                              : Area( recalcPosition( cu.chromaFormat, cu.chType, CHANNEL_TYPE_LUMA, cu.blocks[cu.chType].pos() ),
                                      recalcSize( cu.chromaFormat, cu.chType, CHANNEL_TYPE_LUMA, cu.blocks[cu.chType].size() ) );

  // xSetLoopfilterParam (LoopFilter.cpp:394-417)
  const Slice& slice = *cu.slice;
  bool         leftEdge = false, topEdge = false, internalEdge = false;
  if( !slice.getDeblockingFilterDisable() )
  {
    const Position& pos = cu.blocks[cu.chType].pos();
    internalEdge        = true;
    // the neighbour is looked up only when the answer depends on it (the flag is 1 in every reference configuration)
    const bool across = slice.getLFCrossSliceBoundaryFlag();
    if( pos.x > 0 ) leftEdge = across || CU::isSameSlice( cu, *cu.cs->getCU( pos.offset( -1, 0 ), cu.chType ) );
    if( pos.y > 0 ) topEdge = across || CU::isSameSlice( cu, *cu.cs->getCU( pos.offset( 0, -1 ), cu.chType ) );
  }

  // TU edges (:250-255), PU edges (:257-266), affine sub-block edges (:268-284)
  for( auto& tu : CU::traverseTUs( cu ) )
  {
    const Area areaTu = lumaValid ? Area( tu.block( COMPONENT_Y ) ) : area;
    es.mark( 0, areaTu, internalEdge, false );
    es.mark( 1, areaTu, internalEdge, false );
  }
  for( auto& pu : CU::traversePUs( cu ) )
  {
    const Area areaPu = lumaValid ? Area( pu.block( COMPONENT_Y ) ) : area;
    const bool xOff   = pu.blocks[cu.chType].x != cu.blocks[cu.chType].x;
    const bool yOff   = pu.blocks[cu.chType].y != cu.blocks[cu.chType].y;
    es.mark( 0, areaPu, xOff ? internalEdge : leftEdge, xOff );
    es.mark( 1, areaPu, yOff ? internalEdge : topEdge, yOff );
  }
  if( cu.affine )
  {
    for( int e = 1; e < int( cu.Y().width ) / 4; e++ ) es.mark( 0, Area( cu.Y().x + e * 4, cu.Y().y, 4, cu.Y().height ), internalEdge, true );
    for( int e = 1; e < int( cu.Y().height ) / 4; e++ ) es.mark( 1, Area( cu.Y().x, cu.Y().y + e * 4, cu.Y().width, 4 ), internalEdge, true );
  }
  // Per-unit CU / TU / motion fields (looked up by position in xGetBoundaryStrengthSingle / xEdgeFilter*).
  const bool     noFilt = ( pc.pcmFilter && cu.ipcm ) || ( pc.tqBypass && cu.transQuantBypass );
  const uint32_t cuBits = ( cu.predMode == MODE_INTRA ? ILF_BI_INTRA : 0u ) | ( noFilt ? ILF_BI_NOFILT : 0u ) |
                          ( slice.isInterB() ? ILF_BI_BSLICE : 0u ) | ( uint32_t( uint8_t( cu.qp ) ) << 8 ) | 0xFFFF0000u;
  const int ux0 = area.x / 4, uy0 = area.y / 4;
  const int ux1 = std::min<int>( unitsW, ( area.x + area.width ) / 4 ), uy1 = std::min<int>( unitsH, ( area.y + area.height ) / 4 );
  for( int y = uy0; y < uy1; y++ )
    for( int x = ux0; x < ux1; x++ ) info[size_t( y ) * unitsW + x] = cuBits;

  if( lumaValid )
  {
    for( auto& tu : CU::traverseTUs( cu ) )
    {
      if( !TU::getCbf( tu, COMPONENT_Y ) ) continue;
      const Area a = tu.block( COMPONENT_Y );
      for( int y = a.y / 4; y < std::min<int>( unitsH, ( a.y + a.height ) / 4 ); y++ )
        for( int x = a.x / 4; x < std::min<int>( unitsW, ( a.x + a.width ) / 4 ); x++ ) info[size_t( y ) * unitsW + x] |= ILF_BI_CBF;
    }
    if( cu.predMode != MODE_INTRA && layer == 0 )
    {
      const uint32_t*  refTab = pc.refsOf( slice );
      const CMotionBuf mb     = cu.cs->getMotionBuf( Area( ux0 * 4, uy0 * 4, ( ux1 - ux0 ) * 4, ( uy1 - uy0 ) * 4 ) );  // one MotionInfo per 4x4 unit
      int32_t*         mv32   = pc.out->wantMv32 ? pc.out->mv32.data() : nullptr;
      int16_t*         mv16   = pc.out->mv16.data();
      bool             fits   = true;
      for( int y = uy0; y < uy1; y++ )
      {
        const MotionInfo* row = mb.buf + size_t( y - uy0 ) * mb.stride;
        for( int x = ux0; x < ux1; x++ )
        {
          const MotionInfo& mi  = row[x - ux0];
          const size_t      idx = size_t( y ) * unitsW + x;
          uint32_t          refs[2] = { ILF_REF_NONE, ILF_REF_NONE };
          int               v[4] = { 0, 0, 0, 0 };
          for( int l = 0; l < 2; l++ )
          {
            if( mi.refIdx[l] >= 0 )  // LoopFilter.cpp:456-466
            {
              refs[l] = refTab[l * MAX_NUM_REF + mi.refIdx[l]];
              // Mv::setHighPrec (:470-477, Mv.h:259-265) multiplies a low-precision vector by 4, sign-symmetrically
              const int sc = ( pc.highPrecMv && !mi.mv[l].highPrec ) ? ( 1 << VCEG_AZ07_MV_ADD_PRECISION_BIT_FOR_STORE ) : 1;
              v[2 * l]     = mi.mv[l].hor * sc;
              v[2 * l + 1] = mi.mv[l].ver * sc;
            }
          }
          if( mv32 ) { mv32[idx * 4] = v[0]; mv32[idx * 4 + 1] = v[1]; mv32[idx * 4 + 2] = v[2]; mv32[idx * 4 + 3] = v[3]; }
          mv16[idx * 4] = int16_t( v[0] ); mv16[idx * 4 + 1] = int16_t( v[1] ); mv16[idx * 4 + 2] = int16_t( v[2] ); mv16[idx * 4 + 3] = int16_t( v[3] );
          fits &= ( unsigned( v[0] + 32768 ) | unsigned( v[1] + 32768 ) | unsigned( v[2] + 32768 ) | unsigned( v[3] + 32768 ) ) < 65536u;
          info[idx] = ( info[idx] & 0x0000FFFFu ) | ( refs[0] << 16 ) | ( refs[1] << 24 );
        }
      }
      pc.anyInter = true;
      if( !fits ) pc.mvFits16 = false;
    }
  }
  // Which columns / rows of the CU does the reference actually filter (LoopFilter.cpp:313-354)?
  for( int dir = 0; dir < 2; dir++ )
  {
    if( lumaValid && ( ( dir == 0 ? cu.Y().x : cu.Y().y ) % 8 ) != 0 ) continue;  // DEBLOCKING_GRID_8x8 early return
    int len = 1, inc = 1;                                                         // DB_TU_FIX
    if( lumaValid )
    {
      if( dir == 1 && cu.Y().height > 64 ) { inc = 16; len = cu.Y().height / 4; }
      if( dir == 0 && cu.Y().width > 64 )  { inc = 16; len = cu.Y().width / 4; }
    }
    const int      n     = dir == 0 ? area.height / 4 : area.width / 4;
    const uint32_t eBit  = dir == 0 ? ILF_BI_EDGE_V : ILF_BI_EDGE_H;
    const uint32_t tuBit = dir == 0 ? ILF_BI_TU_V : ILF_BI_TU_H;
    for( int e = 0; e < len; e += inc )
      for( int i = 0; i < n; i++ )
      {
        const int x = dir == 0 ? ux0 + e : ux0 + i;
        const int y = dir == 0 ? uy0 + i : uy0 + e;
        if( x >= unitsW || y >= unitsH ) continue;
        const size_t idx = size_t( y ) * unitsW + x;
        if( es.edge[dir][idx] ) info[idx] |= eBit;
        if( es.pre[dir][idx] ) info[idx] |= tuBit;
      }
  }
  pc.out->ctuSlice[size_t( area.y >> pcv.maxCUHeightLog2 ) * pc.out->ctusW + ( area.x >> pcv.maxCUWidthLog2 )] = uint8_t( pc.sliceId( cu.slice ) );
}
}  // namespace

void ilfPackDeblock( CodingStructure& cs, IlfPackedDeblock& out )
{
  const PreCalcValues& pcv = *cs.pcv;
  CHECK( pcv.minCUWidth != 4 || pcv.minCUHeight != 4, "ilf_b200: 4x4 minimum unit expected" );
  CHECK( pcv.chrFormat != CHROMA_420, "ilf_b200: 4:2:0 only" );
  out.unitsW = pcv.lumaWidth / 4;
  out.unitsH = pcv.lumaHeight / 4;
  out.ctusW  = pcv.widthInCtus;
  out.ctusH  = pcv.heightInCtus;
  const size_t n = size_t( out.unitsW ) * out.unitsH;
  out.info.assign( n, 0xFFFF0000u );
  if( out.wantMv32 ) out.mv32.assign( n * 4, 0 ); else out.mv32.clear();
  out.mv16.assign( n * 4, 0 );   // written in place by the walk; meaningful only while mvFits16
  out.mvFits16 = true;
  out.anyInter = false;
  out.ctuSlice.assign( size_t( out.ctusW ) * out.ctusH, 0 );
  std::memset( &out.params, 0, sizeof( out.params ) );

  // Dual tree is a slice property (CS::isDualITree uses cs.slice, UnitTools.cpp:59-62; loopFilterPic asks once
  // per CTU but always with the picture-level cs, LoopFilter.cpp:182).
  const bool dual = CS::isDualITree( cs );
  if( dual ) out.infoChroma.assign( n, 0xFFFF0000u ); else out.infoChroma.clear();

  static EdgeScratch scratchStore[2];
  scratchStore[0].init( out.unitsW, out.unitsH );
  if( dual ) scratchStore[1].init( out.unitsW, out.unitsH );

  const bool highPrecMv   = cs.sps->getSpsNext().getUseHighPrecMv();
  out.params.cb_qp_offset = cs.pps->getQpOffset( COMPONENT_Cb );
  out.params.cr_qp_offset = cs.pps->getQpOffset( COMPONENT_Cr );
  out.params.mv_threshold = highPrecMv ? ( 4 << VCEG_AZ07_MV_ADD_PRECISION_BIT_FOR_STORE ) : 4;

  // slices and their reference pictures (dense ids: only identity matters, xGetBoundaryStrengthSingle :456-466)
  PackShared sh;
  std::map<const Picture*, int> refIds;
  for( const Slice* sl : cs.picture->slices ) sh.slices.push_back( sl );
  if( sh.slices.empty() ) sh.slices.push_back( cs.slice );
  CHECK( sh.slices.size() > ILF_MAX_SLICES, "ilf_b200: more slices per picture than ILF_MAX_SLICES" );
  out.params.num_slices = int( sh.slices.size() );
  sh.refTabs.assign( sh.slices.size() * 2 * MAX_NUM_REF, ILF_REF_NONE );
  for( size_t i = 0; i < sh.slices.size(); i++ )
  {
    const Slice& sl = *sh.slices[i];
    out.params.slices[i].beta_offset_div2 = int8_t( sl.getDeblockingFilterBetaOffsetDiv2() );
    out.params.slices[i].tc_offset_div2   = int8_t( sl.getDeblockingFilterTcOffsetDiv2() );
    if( sl.isIntra() ) continue;
    for( int l = 0; l < 2; l++ )
      for( int r = 0; r < MAX_NUM_REF && r < sl.getNumRefIdx( RefPicList( l ) ); r++ )
      {
        const Picture* rp = sl.getRefPic( RefPicList( l ), r );
        if( !rp ) continue;
        auto it = refIds.find( rp );
        if( it == refIds.end() ) { CHECK( refIds.size() >= 255, "ilf_b200: too many distinct reference pictures" ); it = refIds.emplace( rp, int( refIds.size() ) ).first; }
        sh.refTabs[( i * 2 + l ) * MAX_NUM_REF + r] = uint32_t( it->second );
      }
  }

  // The CTU rows are dealt to a few workers: coding units are disjoint, so every worker writes its own units of the grids.
  static const int maxThreads = getenv( "ILF_PACK_THREADS" ) ? std::max( 1, atoi( getenv( "ILF_PACK_THREADS" ) ) ) : 4;
  const int nThreads = ( pcv.sizeInCtus >= 64 ) ? std::min<int>( maxThreads, int( pcv.heightInCtus ) ) : 1;
  std::vector<PackCtx> ctx( nThreads );
  auto work = [&]( int t )
  {
    PackCtx& pc   = ctx[t];
    pc.cs         = &cs;
    pc.out        = &out;
    pc.scratch    = scratchStore;
    pc.sh         = &sh;
    pc.pcmFilter  = cs.sps->getUsePCM() && cs.sps->getPCMFilterDisableFlag();
    pc.tqBypass   = cs.pps->getTransquantBypassEnabledFlag();
    pc.highPrecMv = highPrecMv;
    for( unsigned y = t; y < pcv.heightInCtus; y += nThreads )
      for( unsigned x = 0; x < pcv.widthInCtus; x++ )
      {
        const UnitArea ctuArea( pcv.chrFormat, Area( x << pcv.maxCUWidthLog2, y << pcv.maxCUHeightLog2, pcv.maxCUWidth, pcv.maxCUWidth ) );
        for( auto& cu : cs.traverseCUs( CS::getArea( cs, ctuArea, CH_L ), CH_L ) ) packCU( pc, cu, 0 );
        if( dual )
          for( auto& cu : cs.traverseCUs( CS::getArea( cs, ctuArea, CH_C ), CH_C ) ) packCU( pc, cu, 1 );
      }
  };
  if( nThreads == 1 ) work( 0 );
  else
  {
    std::vector<std::thread> pool;
    std::exception_ptr       err;
    std::mutex               errLock;
    for( int t = 0; t < nThreads; t++ )
      pool.emplace_back( [&, t]() { try { work( t ); } catch( ... ) { std::lock_guard<std::mutex> g( errLock ); err = std::current_exception(); } } );
    for( auto& th : pool ) th.join();
    if( err ) std::rethrow_exception( err );
  }
  for( const PackCtx& pc : ctx ) { out.anyInter |= pc.anyInter; out.mvFits16 &= pc.mvFits16; }
}

// ------------------------------------------------------------------------------------------------------------
// SAO
// ------------------------------------------------------------------------------------------------------------
namespace
{
// deriveLoopFilterBoundaryAvailibility (SampleAdaptiveOffset.cpp:685-760), no tiles in this reference version.
uint8_t saoAvail( CodingStructure& cs, const Position& pos )
{
  const int w = cs.pcv->maxCUWidth, h = cs.pcv->maxCUHeight;
  const CodingUnit* cur = cs.getCU( pos, CH_L );
  auto same = [&]( const CodingUnit* o, bool useOtherFlag, bool diag ) -> bool
  {
    if( !o ) return false;
    if( CU::isSameSlice( *cur, *o ) ) return true;
    if( diag )  // above-right / below-left: the flag of whichever slice starts later (:732, :740)
      return ( cur->slice->getSliceCurStartCtuTsAddr() > o->slice->getSliceCurStartCtuTsAddr() ) ? cur->slice->getLFCrossSliceBoundaryFlag()
                                                                                               : o->slice->getLFCrossSliceBoundaryFlag();
    return useOtherFlag ? o->slice->getLFCrossSliceBoundaryFlag() : cur->slice->getLFCrossSliceBoundaryFlag();
  };
  uint8_t a = 0;
  if( same( cs.getCU( pos.offset( -w, 0 ), CH_L ), false, false ) ) a |= ILF_AVAIL_L;
  if( same( cs.getCU( pos.offset( 0, -h ), CH_L ), false, false ) ) a |= ILF_AVAIL_A;
  if( same( cs.getCU( pos.offset( w, 0 ), CH_L ), true, false ) ) a |= ILF_AVAIL_R;
  if( same( cs.getCU( pos.offset( 0, h ), CH_L ), true, false ) ) a |= ILF_AVAIL_B;
  if( same( cs.getCU( pos.offset( -w, -h ), CH_L ), false, false ) ) a |= ILF_AVAIL_AL;
  if( same( cs.getCU( pos.offset( w, h ), CH_L ), true, false ) ) a |= ILF_AVAIL_BR;
  if( same( cs.getCU( pos.offset( w, -h ), CH_L ), false, true ) ) a |= ILF_AVAIL_AR;
  if( same( cs.getCU( pos.offset( -w, h ), CH_L ), false, true ) ) a |= ILF_AVAIL_BL;
  return a;
}
}  // namespace

void ilfPackSao( CodingStructure& cs, SAOBlkParam* blk, const uint32_t offsetStepLog2[3], IlfPackedSao& out )
{
  const PreCalcValues& pcv   = *cs.pcv;
  const int            nComp = getNumberValidComponents( pcv.chrFormat );
  out.ctus.assign( pcv.sizeInCtus, ilf_sao_ctu() );
  out.anyEnabled = false;

  for( int ctu = 0; ctu < int( pcv.sizeInCtus ); ctu++ )
  {
    const int         cx = ctu % pcv.widthInCtus, cy = ctu / pcv.widthInCtus;
    const Position    pos( cx * pcv.maxCUWidth, cy * pcv.maxCUHeight );
    const CodingUnit& cu = *cs.getCU( pos, CH_L );
    // merge candidates (getMergeList, :172-226): the neighbour CTU must be in the same slice and earlier in
    // coding order, which is what getCURestricted tests.
    SAOBlkParam* mergeAbove = ( cy > 0 && cs.getCURestricted( pos.offset( 0, -int( pcv.maxCUHeight ) ), cu, cu.chType ) ) ? &blk[ctu - pcv.widthInCtus] : nullptr;
    SAOBlkParam* mergeLeft  = ( cx > 0 && cs.getCURestricted( pos.offset( -int( pcv.maxCUWidth ), 0 ), cu, cu.chType ) ) ? &blk[ctu - 1] : nullptr;

    ilf_sao_ctu& o = out.ctus[ctu];
    std::memset( &o, 0, sizeof( o ) );
    for( int c = 0; c < 3; c++ ) o.type[c] = ILF_SAO_OFF;
    for( int c = 0; c < nComp; c++ )
    {
      SAOOffset& p = blk[ctu][c];
      if( p.modeIdc == SAO_MODE_MERGE )  // reconstructBlkSAOParam (:229-263)
      {
        SAOBlkParam* t = p.typeIdc == SAO_MERGE_LEFT ? mergeLeft : mergeAbove;
        CHECK( t == nullptr, "Merge target does not exist" );
        p = ( *t )[c];
      }
      else if( p.modeIdc == SAO_MODE_NEW )  // invertQuantOffsets (:147-170)
      {
        int coded[MAX_NUM_SAO_CLASSES];
        std::memcpy( coded, p.offset, sizeof( coded ) );
        std::memset( p.offset, 0, sizeof( p.offset ) );
        if( p.typeIdc == SAO_TYPE_START_BO )
          for( int i = 0; i < 4; i++ ) p.offset[( p.typeAuxInfo + i ) % NUM_SAO_BO_CLASSES] = coded[( p.typeAuxInfo + i ) % NUM_SAO_BO_CLASSES] * ( 1 << offsetStepLog2[c] );
        else
        {
          for( int i = 0; i < NUM_SAO_EO_CLASSES; i++ ) p.offset[i] = coded[i] * ( 1 << offsetStepLog2[c] );
          CHECK( p.offset[SAO_CLASS_EO_PLAIN] != 0, "EO offset is not '0'" );
        }
      }
      if( p.modeIdc == SAO_MODE_OFF ) continue;
      out.anyEnabled = true;
      o.type[c]      = int8_t( p.typeIdc );
      if( p.typeIdc == SAO_TYPE_START_BO )
      {
        o.band_pos[c] = uint8_t( p.typeAuxInfo );
        for( int i = 0; i < 4; i++ ) o.offset[c][i] = int16_t( p.offset[( p.typeAuxInfo + i ) % NUM_SAO_BO_CLASSES] );
      }
      else
      {
        o.offset[c][0] = int16_t( p.offset[0] );
        o.offset[c][1] = int16_t( p.offset[1] );
        o.offset[c][2] = int16_t( p.offset[3] );
        o.offset[c][3] = int16_t( p.offset[4] );
      }
    }
    o.avail = saoAvail( cs, pos );
  }
}

// ------------------------------------------------------------------------------------------------------------
// ALF
// ------------------------------------------------------------------------------------------------------------
void ilfReconstructAlfCoeff( AlfSliceParam& p, bool luma, short* coeffFinal, bool redo )
{
  // AdaptiveLoopFilter::reconstructCoeff (AdaptiveLoopFilter.cpp:141-194); fixed point with 1.0 == 1 << 9.
  const int  nCoef    = ( luma && p.lumaFilterType == ALF_FILTER_7 ) ? 13 : 7;
  const int  nFilters = luma ? p.numLumaFilters : 1;
  short*     c        = luma ? p.lumaCoeff : p.chromaCoeff;
  const int  rowLen   = MAX_NUM_ALF_LUMA_COEFF;
  const bool pred     = luma && p.coeffDeltaPredModeFlag;
  if( pred )
    for( int f = 1; f < nFilters; f++ )
      for( int j = 0; j < nCoef - 1; j++ ) c[f * rowLen + j] += c[( f - 1 ) * rowLen + j];
  for( int f = 0; f < nFilters; f++ )
  {
    int twice = 0;
    for( int j = 0; j < nCoef - 1; j++ ) twice += c[f * rowLen + j] << 1;
    c[f * rowLen + nCoef - 1] = short( ( 1 << 9 ) - twice );
  }
  if( !luma ) return;
  for( int cls = 0; cls < MAX_NUM_ALF_CLASSES; cls++ )
    std::memcpy( coeffFinal + cls * rowLen, c + p.filterCoeffDeltaIdx[cls] * rowLen, sizeof( short ) * nCoef );
  if( redo && pred )
    for( int f = nFilters - 1; f > 0; f-- )
      for( int j = 0; j < nCoef - 1; j++ ) c[f * rowLen + j] = c[f * rowLen + j] - c[( f - 1 ) * rowLen + j];
}

void ilfPackAlf( CodingStructure& cs, AlfSliceParam& p, IlfPackedAlf& out, bool encoder )
{
  std::memset( &out.params, 0, sizeof( out.params ) );
  out.enabled = p.enabledFlag[COMPONENT_Y] || p.enabledFlag[COMPONENT_Cb] || p.enabledFlag[COMPONENT_Cr];
  const int n = cs.pcv->sizeInCtus;
  out.ctuEnable.assign( size_t( 3 ) * n, 0 );
  if( !out.enabled ) return;
  short coeffFinal[MAX_NUM_ALF_CLASSES * MAX_NUM_ALF_LUMA_COEFF];
  std::memset( coeffFinal, 0, sizeof( coeffFinal ) );
  if( !encoder || p.enabledFlag[COMPONENT_Y] ) ilfReconstructAlfCoeff( p, true, coeffFinal, encoder );
  if( !encoder || p.enabledFlag[COMPONENT_Cb] || p.enabledFlag[COMPONENT_Cr] ) ilfReconstructAlfCoeff( p, false, nullptr, false );
  for( int cls = 0; cls < 25; cls++ )
    for( int j = 0; j < 13; j++ ) out.params.luma_coeff[cls][j] = coeffFinal[cls * MAX_NUM_ALF_LUMA_COEFF + j];
  for( int j = 0; j < 7; j++ ) out.params.chroma_coeff[j] = p.chromaCoeff[j];
  out.params.luma_filter_7x7 = p.lumaFilterType == ALF_FILTER_7 ? 1 : 0;
  for( int c = 0; c < 3; c++ ) std::memcpy( &out.ctuEnable[size_t( c ) * n], cs.picture->getAlfCtuEnableFlag( c ), n );
}

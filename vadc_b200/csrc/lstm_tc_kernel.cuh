// vadc_b200/csrc/lstm_tc_kernel.cuh -- the decoder LSTM on the 5th-gen tensor cores (tcgen05).
//
// Same function as lstm_kernel.cuh (lstm.c:31-341 + decoder silero_v3.c:231-303), used when the
// stream batch is wide enough to make the gate contraction a dense GEMM. Per time step and tile of
// LTC_N = 32 streams:   Z[256 gates][32 streams] = W[256][128] * [x_t ; h_{t-1}][128][32]
// issued as tcgen05.mma (M = 128 x 2 tiles, N = 32, K = 16 x 8) by one elected thread, accumulated
// in TMEM, with the bf16x2 split (3 partial products: Whi*Xhi + Wlo*Xhi + Whi*Xlo, ~16 significant
// bits + fp32 accumulation; measured effect on the speech probability: ~5e-6).
//
// Roles: warp 8 = MMA issuer (and TMEM owner); warps 0..7 = cell update ("epilogue"): they read the
// gate pre-activations from TMEM (tcgen05.ld), apply the gate nonlinearities and the cell update
// (lstm.c:64-88) with c in registers, write h_t as the bf16 hi/lo B operand of the next step, and
// stage x_{t+1} next to it. mbarriers carry the two hand-offs (operand ready -> MMA, MMA done ->
// cell update). The weights never change, so they are the A operand IN TENSOR MEMORY (tcgen05.mma with [a_tmem]): 256 of the
// 512 TMEM columns hold W as bf16 hi/lo (row = lane, two K elements per 32-bit column; rows permuted so that the four gates of
// a hidden unit land in one warp), written once per CTA with tcgen05.st. With W in shared memory every one of the 48 MMAs of a
// step re-read 4 KB of it (32 cycles at 128 B/clk against 16 cycles of math for N = 32): ncu showed the issuing warp
// back-pressured on every UTCHMMA and the cell-update warps waiting 41 % of the step for the contraction.
//
// Row permutation: M-tile m (0,1), TMEM lane L = 32*wq + q holds gate row  gate*64 + u  with
//   u = 16*wq + (q & 15),  gate = m == 0 ? (q < 16 ? i : f) : (q < 16 ? g : o).
// A warp (wq = warp & 3, stream half hf = warp >> 2) therefore owns units 16*wq..16*wq+15 for 16
// streams; lanes q and q^16 swap halves by shuffle so that each lane ends up with all four gates of
// one unit for 8 streams.
//
// Two chains per CTA. A step is a strict dependency chain (operand -> MMA -> TMEM load -> gates -> operand). The 32 streams
// of a tile are split into two independent chains of LTC_NC = 16 streams (N = 16 MMAs), each with its own operand buffers,
// TMEM columns, mbarriers and four cell-update warps (chain = warp >> 2): while one chain's gates are evaluated the tensor
// core works on the other's contraction.
//
// Layer 0 reads the encoder output a4 (fp32 [S][steps][64]) and writes its h sequence packed as the
// next layer's operand: hp [tile][step][split][8 chunks][32 streams][8] bf16 (8 KB per tile-step).
// Layer 1 reads hp and folds the decoder head in (relu -> 64->2 -> mean over 7 frames -> sigmoid).
#pragma once
#include "common.cuh"
#include "tc_common.cuh"

#define LTC_N 32
#define LTC_EPI_WARPS 8
#define LTC_EPI_THREADS ( LTC_EPI_WARPS * 32 )
#define LTC_THREADS ( LTC_EPI_THREADS + 32 )
#define LTC_W_LBO ( 256 * 16 )                  // bytes between K chunks of the weight operand
#define LTC_W_SPLIT_BYTES ( 16 * LTC_W_LBO )    // 64 KB per split
#define LTC_W_BYTES ( 2 * LTC_W_SPLIT_BYTES )   // hi, lo
#define LTC_NC 16                               // streams per chain (two chains per tile)
#define LTC_CHAIN_THREADS 128
#define LTC_X_LBO ( LTC_NC * 16 )               // operand buffers are per chain: [16 chunks][16 streams][8]
#define LTC_X_SPLIT_BYTES ( 16 * LTC_X_LBO )    // 4 KB
#define LTC_X_BUF_BYTES ( 2 * LTC_X_SPLIT_BYTES )
#define LTC_HP_LBO ( LTC_N * 16 )
#define LTC_HP_BYTES ( 2 * 8 * LTC_HP_LBO )     // packed h of one tile-step: [split][8 chunks][32 streams][8] = 8 KB
#define LTC_TMEM_COLS 512                       // accumulators: 2 chains x 2 buffers x 2 M-tiles x 16 columns; weights: see LTC_TMEM_W
#define LTC_TMEM_W 128                          // first weight column; (M-tile m, split p) at LTC_TMEM_W + (2 m + p) * 64, K = 128 -> 64 columns
#define LTC_DEC_FLOATS ( 2 * 4 * LTC_N * 2 )         // decoder partial sums: [chunk parity][unit quarter][stream][head]
#define LTC_SMEM_BYTES ( 4 * LTC_X_BUF_BYTES + 256 * 4 + 128 * 4 + LTC_DEC_FLOATS * 4 + 128 )

// host-side image of one layer's weights in shared-memory order: [split][chunk][row'][8] bf16
// (engine.cu: pack_lstm_tc)

// Gate nonlinearities on the special-function unit (ex2.approx + rcp.approx, ~2^-22 relative each): the pre-activations
// they are applied to already carry the ~1e-5 relative error of the bf16x2 contraction, so libm-accurate expf/tanhf
// (3x the instructions) buy nothing here. The FP32 kernel (lstm_kernel.cuh) keeps the accurate forms.
__device__ __forceinline__ float ltc_sigmoid( float v ) { return __fdividef( 1.0f, 1.0f + __expf( -v ) ); }
__device__ __forceinline__ float ltc_tanh( float v )
{
   // 1 - 2/(1 + e^{2v}); saturates correctly for large |v| (e^{2v} -> inf gives 1, -> 0 gives -1)
   return 1.0f - __fdividef( 2.0f, 1.0f + __expf( 2.0f * v ) );
}
// 2^min(x, 40) and 1/x on the special-function unit
__device__ __forceinline__ float ltc_ex2c( float x )
{
   float r;
   asm( "ex2.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( fminf( x, 40.0f ) ) );
   return r;
}
__device__ __forceinline__ float ltc_rcp( float x )
{
   float r;
   asm( "rcp.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( x ) );
   return r;
}
// One LSTM cell (lstm.c:64-88) with the gate nonlinearities over common denominators: the cell-update warps are bound by the
// special-function unit (ncu: XU pipe 63 %), and sigma(zi) tanh(zg) + sigma(zf) c = [c Di Dg + (1 - eg) Df] / (Df Di Dg) with
// e* = 2^(-z* log2 e), D* = 1 + e* needs one reciprocal instead of three, o tanh(c') = (1 - ec) / ((1 + eo)(1 + ec)) one instead of
// two: 7 MUFU per cell instead of 10. Exponents are clamped at 2^40 so that a product of three denominators stays finite
// (sigma(-27.7) = 9e-13 is already below fp32 resolution next to 1).
__device__ __forceinline__ void ltc_cell( float zi, float zf, float zg, float zo, float &c, float &h )
{
   constexpr float L = 1.4426950408889634f;
   const float ei = ltc_ex2c( -L * zi ), ef = ltc_ex2c( -L * zf ), eg = ltc_ex2c( -2.0f * L * zg ), eo = ltc_ex2c( -L * zo );
   const float di = 1.0f + ei, df = 1.0f + ef, dg = 1.0f + eg;
   const float didg = di * dg;
   const float cn = fmaf( c, didg, ( 1.0f - eg ) * df ) * ltc_rcp( df * didg );
   const float ec = ltc_ex2c( -2.0f * L * cn );
   c = cn;
   h = ( 1.0f - ec ) * ltc_rcp( ( 1.0f + eo ) * ( 1.0f + ec ) );
}

template <int LAYER>
__global__ void __launch_bounds__( LTC_THREADS, 1 )
lstm_tc_kernel( const float *__restrict__ x_f32,            // LAYER 0: a4 [S][steps][64]
                const unsigned char *__restrict__ x_packed, // LAYER 1: hp
                unsigned char *__restrict__ hp_out,         // LAYER 0: hp
                float *__restrict__ state_h, float *__restrict__ state_c, const unsigned char *__restrict__ wimg /*[2 layers][LTC_W_BYTES]*/,
                const float *__restrict__ bias, const float *__restrict__ dec_w, const float *__restrict__ dec_b, int nstreams, int nw,
                float *__restrict__ out2, float *__restrict__ probs, long long out_stride, long long out_off )
{
   extern __shared__ __align__( 128 ) unsigned char ltc_smem[];
   unsigned char *smem = ltc_smem;
   unsigned char *sX = smem;                                                // [2 chains][2 bufs][2 splits][16][16][8]
   float *sBias = reinterpret_cast<float *>( sX + 4 * LTC_X_BUF_BYTES );    // [256]
   float *sDw = sBias + 256;                                                // [2][64]
   float *sDec = sDw + 128;                                                 // [2][4][32 streams][2 heads]
   uint64_t *bars = reinterpret_cast<uint64_t *>( sDec + LTC_DEC_FLOATS );  // xh[chain][buf], d[chain][buf]
   uint32_t *tmem_slot = reinterpret_cast<uint32_t *>( bars + 8 );
   uint64_t *bar_xh = bars, *bar_d = bars + 4;

   const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
   const int steps = nw * 7;
   const int ntiles = ( nstreams + LTC_N - 1 ) / LTC_N;

   // ---- one-time setup ---------------------------------------------------------------------------
   {
      for ( int i = tid; i < 256; i += LTC_THREADS ) sBias[i] = bias[LAYER * 256 + i];
      for ( int i = tid; i < 128; i += LTC_THREADS ) sDw[i] = dec_w[i];
   }
   if ( warp == LTC_EPI_WARPS )
   {
      tc::tmem_alloc( tmem_slot, LTC_TMEM_COLS );
      if ( lane == 0 )
      {
         for ( int i = 0; i < 4; ++i )
         {
            tc::mbar_init( &bar_xh[i], LTC_EPI_WARPS / 2 );
            tc::mbar_init( &bar_d[i], 1 );
         }
         tc::mbar_fence_init();
      }
   }
   tc::fence_before_sync();
   __syncthreads();
   tc::fence_after_sync();
   const uint32_t tmem = *tmem_slot;
   if ( warp < LTC_EPI_WARPS )
   {
      // weights -> tensor memory: warp (wq, m) writes rows 32 wq .. 32 wq + 31 of M-tile m; the host image is
      // [split][16 K-chunks][256 rows'][8 bf16], i.e. 16 bytes = 4 columns per (chunk, row)
      const int wq = warp & 3, m = warp >> 2;
      const int4 *src = reinterpret_cast<const int4 *>( wimg + (size_t)LAYER * LTC_W_BYTES );
      const uint32_t trow = tmem + ( (uint32_t)( wq * 32 ) << 16 ) + LTC_TMEM_W + m * 128;
#pragma unroll 1
      for ( int p = 0; p < 2; ++p )
#pragma unroll 4
         for ( int c = 0; c < 16; ++c )
         {
            const int4 v = __ldg( src + ( (size_t)p * 16 + c ) * 256 + m * 128 + wq * 32 + lane );
            tc::tmem_st4( trow + p * 64 + c * 4, (uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w );
         }
      tc::tmem_wait_st();
   }
   tc::fence_before_sync();
   __syncthreads();
   tc::fence_after_sync();

   // `it` counts tile-steps processed by this CTA; buffer = it & 1, mbarrier phase = (it >> 1) & 1
   if ( warp == LTC_EPI_WARPS )
   {
      // ================================ MMA issuer ================================================
      const uint32_t idesc = tc::idesc_bf16_f32( 128, LTC_NC );
      const uint64_t dX = tc::smem_desc( tc::smem_u32( sX ), LTC_X_LBO, 128 );
      uint32_t it = 0;
      for ( int tile = blockIdx.x; tile < ntiles; tile += gridDim.x )
         for ( int step = 0; step < steps; ++step, ++it )
#pragma unroll 1
         for ( int chain = 0; chain < 2; ++chain )
         {
            const uint32_t buf = it & 1, ph = ( it >> 1 ) & 1;
            tc::mbar_wait( &bar_xh[chain * 2 + buf], ph );
            tc::fence_after_sync();
            if ( tc::elect_one() )
            {
               const uint64_t dXb = dX + (uint64_t)( ( chain * 2 + buf ) * ( LTC_X_BUF_BYTES >> 4 ) );
#pragma unroll
               for ( int m = 0; m < 2; ++m )
               {
                  const uint32_t d_tmem = tmem + chain * 64 + buf * 32 + m * 16;
                  const uint32_t aWm = tmem + LTC_TMEM_W + m * 128;
                  // (W split, X split): (hi,hi) (lo,hi) (hi,lo)
#pragma unroll
                  for ( int p = 0; p < 3; ++p )
                  {
                     const uint32_t ta = aWm + ( p == 1 ? 64 : 0 );
                     const uint64_t db = dXb + (uint64_t)( ( p == 2 ? 1 : 0 ) * ( LTC_X_SPLIT_BYTES >> 4 ) );
#pragma unroll
                     for ( int kk = 0; kk < 8; ++kk )
                        tc::mma_bf16_ts( d_tmem, ta + kk * 8, db + (uint64_t)( kk * ( 2 * LTC_X_LBO >> 4 ) ), idesc, ( p | kk ) ? 1u : 0u );
                  }
               }
               tc::mma_commit( &bar_d[chain * 2 + buf] );
            }
            __syncwarp();
         }
   }
   else
   {
      // ================================ cell update ===============================================
      const int wq = warp & 3, chain = warp >> 2;
      const int q = lane & 15, up = lane >> 4; // up = 0: holds i,g and keeps streams 0..7 ; 1: holds f,o and keeps 8..15
      const int u = 16 * wq + q;
      const float bi = sBias[u], bf = sBias[64 + u], bg = sBias[128 + u], bo = sBias[192 + u];
      const float dw0 = sDw[u], dw1 = sDw[64 + u];
      const float db0 = __ldg( dec_b ), db1 = __ldg( dec_b + 1 );
      const int sl0 = chain * LTC_NC + up * 8; // first of this lane's 8 streams inside the tile
      const int rl0 = up * 8;                  // ... and inside the chain's operand buffer
      unsigned char *sXc = sX + chain * 2 * LTC_X_BUF_BYTES;
      uint64_t *bar_xh_c = bar_xh + chain * 2, *bar_d_c = bar_d + chain * 2;
      const int ctid = tid & ( LTC_CHAIN_THREADS - 1 );
      // byte offset of this lane's h element (k = 64 + u) inside one split of an X buffer, for stream row 0
      const uint32_t hoff = (uint32_t)( 8 + ( u >> 3 ) ) * LTC_X_LBO + (uint32_t)( u & 7 ) * 2u;
      // staging role for x inside the chain: thread e -> stream row e & 15, chunk e >> 4
      const int xs = ctid & 15, xc = ctid >> 4;

      uint32_t it = 0;
      for ( int tile = blockIdx.x; tile < ntiles; tile += gridDim.x )
      {
         const int s0 = tile * LTC_N;
         float c[8], hlast[8], d0[8], d1[8];
         // ---- tile prologue: state -> registers / operand buffer, x_0 -> operand buffer -----------------
         {
            unsigned char *xb = sXc + ( it & 1 ) * LTC_X_BUF_BYTES;
#pragma unroll
            for ( int j = 0; j < 8; ++j )
            {
               const int s = s0 + sl0 + j;
               const bool ok = s < nstreams;
               c[j] = ok ? state_c[( (size_t)s * 2 + LAYER ) * 64 + u] : 0.0f;
               hlast[j] = ok ? state_h[( (size_t)s * 2 + LAYER ) * 64 + u] : 0.0f;
               d0[j] = d1[j] = 0.0f;
               tc::Split2 sp = tc::split2( hlast[j] );
               *reinterpret_cast<__nv_bfloat16 *>( xb + hoff + ( rl0 + j ) * 16 ) = sp.hi;
               *reinterpret_cast<__nv_bfloat16 *>( xb + LTC_X_SPLIT_BYTES + hoff + ( rl0 + j ) * 16 ) = sp.lo;
            }
         }
         // x of step `st` -> registers (raw), then -> operand buffer
         float4 xa, xb4;   // LAYER 0: 8 fp32
         int4 xp0, xp1;    // LAYER 1: hi chunk row, lo chunk row
         auto x_fetch = [&]( int st ) {
            if ( LAYER == 0 )
            {
               const int s = s0 + chain * LTC_NC + xs;
               if ( s < nstreams && st < steps )
               {
                  const float4 *p = reinterpret_cast<const float4 *>( x_f32 + ( (size_t)s * steps + st ) * 64 + xc * 8 );
                  xa = __ldg( p );
                  xb4 = __ldg( p + 1 );
               }
               else
                  xa = xb4 = make_float4( 0.f, 0.f, 0.f, 0.f );
            }
            else
            {
               if ( st < steps )
               {
                  const int4 *p = reinterpret_cast<const int4 *>( x_packed + ( (size_t)tile * steps + st ) * LTC_HP_BYTES );
                  xp0 = __ldg( p + xc * LTC_N + chain * LTC_NC + xs );
                  xp1 = __ldg( p + ( 8 + xc ) * LTC_N + chain * LTC_NC + xs );
               }
               else
                  xp0 = xp1 = make_int4( 0, 0, 0, 0 );
            }
         };
         auto x_stage = [&]( unsigned char *xbuf ) {
            int4 hi, lo;
            if ( LAYER == 0 )
            {
               const float v[8] = { xa.x, xa.y, xa.z, xa.w, xb4.x, xb4.y, xb4.z, xb4.w };
               __nv_bfloat16 *ph = reinterpret_cast<__nv_bfloat16 *>( &hi ), *pl = reinterpret_cast<__nv_bfloat16 *>( &lo );
#pragma unroll
               for ( int e = 0; e < 8; ++e )
               {
                  tc::Split2 sp = tc::split2( v[e] );
                  ph[e] = sp.hi;
                  pl[e] = sp.lo;
               }
            }
            else
            {
               hi = xp0;
               lo = xp1;
            }
            *reinterpret_cast<int4 *>( xbuf + xc * LTC_X_LBO + xs * 16 ) = hi;
            *reinterpret_cast<int4 *>( xbuf + LTC_X_SPLIT_BYTES + xc * LTC_X_LBO + xs * 16 ) = lo;
         };
         x_fetch( 0 );
         x_stage( sXc + ( it & 1 ) * LTC_X_BUF_BYTES );
         x_fetch( 1 );
         tc::fence_async_smem();
         __syncwarp();
         if ( lane == 0 ) tc::mbar_arrive( &bar_xh_c[it & 1] );

         for ( int step = 0; step < steps; ++step, ++it )
         {
            const uint32_t buf = it & 1, ph = ( it >> 1 ) & 1;
            unsigned char *xnext = sXc + ( buf ^ 1 ) * LTC_X_BUF_BYTES;
            // x_{t+1} can be staged while the MMAs of step t run: buffer buf^1 was last read by step t-1
            x_stage( xnext );
            x_fetch( step + 2 );

            tc::mbar_wait( &bar_d_c[buf], ph );
            tc::fence_after_sync();
            float a[16], b[16];
            const uint32_t taddr = tmem + ( (uint32_t)( wq * 32 ) << 16 ) + chain * 64 + buf * 32;
            tc::tmem_ld16( taddr, a );
            tc::tmem_ld16( taddr + 16, b );
            tc::tmem_wait_ld();

#pragma unroll
            for ( int j = 0; j < 8; ++j )
            {
               // lanes q and q^16 swap: lower gives (i,g) of streams 8..15, upper gives (f,o) of streams 0..7
               const float ra = __shfl_xor_sync( 0xffffffffu, up ? a[j] : a[8 + j], 16 );
               const float rb = __shfl_xor_sync( 0xffffffffu, up ? b[j] : b[8 + j], 16 );
               const float zi = ( up ? ra : a[j] ) + bi;
               const float zf = ( up ? a[8 + j] : ra ) + bf;
               const float zg = ( up ? rb : b[j] ) + bg;
               const float zo = ( up ? b[8 + j] : rb ) + bo;
               float hn;
               ltc_cell( zi, zf, zg, zo, c[j], hn );
               hlast[j] = hn;
               tc::Split2 sp = tc::split2( hn );
               *reinterpret_cast<__nv_bfloat16 *>( xnext + hoff + ( rl0 + j ) * 16 ) = sp.hi;
               *reinterpret_cast<__nv_bfloat16 *>( xnext + LTC_X_SPLIT_BYTES + hoff + ( rl0 + j ) * 16 ) = sp.lo;
               if ( LAYER == 1 )
               {
                  const float r = fmaxf( hn, 0.0f );
                  d0[j] = fmaf( dw0, r, d0[j] );
                  d1[j] = fmaf( dw1, r, d1[j] );
               }
            }
            tc::fence_async_smem();
            tc::fence_before_sync();
            __syncwarp();
            if ( lane == 0 && step + 1 < steps ) tc::mbar_arrive( &bar_xh_c[buf ^ 1] );

            if ( LAYER == 0 )
            {
               // h_t of the chain's 16 streams is complete in xnext (chunks 8..15 of both splits) once its four cell-update
               // warps are here; write it out as the next layer's packed operand ([split][8 chunks][32 streams][8])
               bar_sync( 1 + chain, LTC_CHAIN_THREADS );
               int4 *dst = reinterpret_cast<int4 *>( hp_out + ( (size_t)tile * steps + step ) * LTC_HP_BYTES ) + xc * LTC_N + chain * LTC_NC + xs;
               const int4 *s_hi = reinterpret_cast<const int4 *>( xnext + 8 * LTC_X_LBO );
               const int4 *s_lo = reinterpret_cast<const int4 *>( xnext + LTC_X_SPLIT_BYTES + 8 * LTC_X_LBO );
               dst[0] = s_hi[ctid];
               dst[8 * LTC_N] = s_lo[ctid];
               // the copy must be done before step t+1's cell update overwrites... it writes the OTHER buffer; the
               // buffer read here is next written at step t+2, after the barrier of step t+1
            }
            else if ( ( step % 7 ) == 6 )
            {
               // decoder head: sum over the 64 units (16 lanes, then the 4 unit quarters in a fixed order so that the
               // result does not depend on warp timing), mean over 7 frames, sigmoid
               float *part = sDec + ( ( step / 7 ) & 1 ) * ( 4 * LTC_N * 2 );
#pragma unroll
               for ( int j = 0; j < 8; ++j )
               {
                  float v0 = d0[j], v1 = d1[j];
#pragma unroll
                  for ( int off = 8; off > 0; off >>= 1 )
                  {
                     v0 += __shfl_xor_sync( 0xffffffffu, v0, off );
                     v1 += __shfl_xor_sync( 0xffffffffu, v1, off );
                  }
                  if ( q == 0 )
                  {
                     part[( wq * LTC_N + sl0 + j ) * 2 + 0] = v0;
                     part[( wq * LTC_N + sl0 + j ) * 2 + 1] = v1;
                  }
                  d0[j] = d1[j] = 0.0f;
               }
               bar_sync( 1 + chain, LTC_CHAIN_THREADS );
               // (the other parity is written 7 steps from now; every step in between needs the chain's four warps to
               // arrive before its MMAs run, so these reads are long done by then)
               if ( ctid < LTC_NC * 2 )
               {
                  const int sl = chain * LTC_NC + ( ctid >> 1 ), head = ctid & 1, s = s0 + sl, pi = sl * 2 + head;
                  const float sum = ( ( part[pi] + part[LTC_N * 2 + pi] ) + part[2 * LTC_N * 2 + pi] ) + part[3 * LTC_N * 2 + pi];
                  const float mean = sum / 7.0f + ( head ? db1 : db0 );
                  if ( s < nstreams )
                  {
                     const float p = 1.0f / ( 1.0f + expf( -mean ) );
                     const long long n = out_off + step / 7;
                     if ( out2 ) out2[( (long long)s * out_stride + n ) * 2 + head] = p;
                     if ( probs && head == 1 ) probs[(long long)s * out_stride + n] = p;
                  }
               }
            }
         }
         // ---- tile epilogue: state back ---------------------------------------------------------------
#pragma unroll
         for ( int j = 0; j < 8; ++j )
         {
            const int s = s0 + sl0 + j;
            if ( s < nstreams )
            {
               state_c[( (size_t)s * 2 + LAYER ) * 64 + u] = c[j];
               state_h[( (size_t)s * 2 + LAYER ) * 64 + u] = hlast[j];
            }
         }
         // the chain's cell-update warps must be done with this tile's buffers before the next prologue writes them
         bar_sync( 1 + chain, LTC_CHAIN_THREADS );
      }
   }
   tc::fence_before_sync();
   __syncthreads();
   if ( warp == LTC_EPI_WARPS ) tc::tmem_dealloc( tmem, LTC_TMEM_COLS );
}

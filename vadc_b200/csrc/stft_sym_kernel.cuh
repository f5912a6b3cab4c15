// vadc_b200/csrc/stft_sym_kernel.cuh -- the bit-faithful conv-basis STFT at half the arithmetic.
//
// Same result, bit for bit, as stft_kernel.cuh (my_stft stft.c:15-229 in the AVX2 reduction tree of stft.c:108-184, magnitude
// stft.c:194-213, log1p of misc.c:40-46). What changes is how much of the tree is evaluated:
//
//   the reference combines the 256 rounded products of an output as  lane l = k % 8:  R[l] = tree over the 32 taps with k = l mod 8,
//   y = ((R0+R1)+(R2+R3)) + ((R4+R5)+(R6+R7)).  The stored basis is hann[k] cos(2 pi f k / 256) | -hann[k] sin(2 pi f k / 256), and for
//   this table  B_re[128-f][k] == (-1)^k B_re[f][k]  and  B_im[128-f][k] == -(-1)^k B_im[f][k]  hold exactly, value for value
//   (checked on the host when the engine is created; a table without the property takes stft_kernel.cuh). The sign of a product
//   follows the sign of its factor exactly, a sum of negated terms is the negated sum exactly, and all taps of a lane l share the
//   parity of k: so  R[l] of bin 128-f  ==  (-1)^l R[l] of bin f  (up to the overall sign of the imaginary part, which the magnitude
//   squares away). Bins f and 128-f therefore share the whole tree below the last seven additions:
//       y(f)     = ((R0+R1)+(R2+R3)) + ((R4+R5)+(R6+R7))
//       y(128-f) = ((R0-R1)+(R2-R3)) + ((R4-R5)+(R6-R7))
//   Bin 64 is its own partner: its real row is zero at odd k, its imaginary row at even k (exactly), products with an exact zero add
//   nothing, so ONE merged row (real taps at even k, imaginary taps at odd k) yields re = (R0+R2)+(R4+R6), im = (R1+R3)+(R5+R7).
//   Bins 0 and 128 have all-zero imaginary rows. 128 rows instead of 256: bins 1..63 (re, im), bin 0 (re), the merged row of bin 64.
//   (Zero signs: where both rows hold +0 the shared product has the sign of the sample for both bins while the rule above would flip
//   it for odd k -- a difference between +0 and -0 that vanishes at the first non-zero addend, and at the squares otherwise.)
//
// Mapping: the 128 rows (128 KB) are resident in shared memory for the whole launch, one CTA per SM, one chunk at a time per CTA.
// thread = (unit u = bin pair (u, 128-u), half of the eight lanes, group of five frames): 64 x 2 x 5 = 640 threads. A thread walks
// its four lanes exactly like stft_kernel.cuh walks all eight (tree state for 2 rows x 5 frames in registers), keeps the plain and
// the alternating pair sums, and meets its partner (lane ^ 16: same unit, other half) in one shuffle per value: the half-0 thread
// finishes bin u, the half-1 thread bin 128-u -- magnitude and log1p are spread over all 640 threads.
// Bin 64 (the second row of unit 0) would give the five warps that hold unit 0 ten magnitudes + log1p per thread instead of five,
// and warp w runs on scheduler w % 4: all five sit on scheduler 0, which made every other warp wait ~17 % of the kernel at the
// chunk barrier (r02c profile). Its 25 (re, im) pairs go through shared memory instead and ONE warp finishes them after the next
// barrier, 25 lanes wide.
#pragma once
#include "common.cuh"
#include "libm_exact.cuh"
#include "stft_kernel.cuh"

#define SSYM_THREADS 640
#define SSYM_BS_FLOATS ( 64 * 128 * 4 )
#define SSYM_B64_FLOATS ( 2 * VB_FRAMES * 2 ) // bin 64: [2 buffers][25 frames](re, im)
#define SSYM_SMEM_BYTES ( ( SSYM_BS_FLOATS + 2 * STFT_XS_FLOATS + SSYM_B64_FLOATS ) * 4 )
#define SSYM_B64_WARP 19 // the warp that finishes bin 64 (no staging work, a scheduler without the two-staging-warp load)

// Packed arithmetic (sm_100 f32x2: two independently rounded IEEE operations in one issue slot, two cycles of the FP32 pipe).
// Measured on this kernel, 131 072 chunks: all scalar 10.28 ms; additions packed over frame pairs (FADD2 fed by scalar FMULs) 10.04 ms;
// PRODUCTS packed over adjacent taps -- FMUL2 on the register pairs an LDS.128 delivers, no moves -- with scalar additions 9.71 ms.
// (scripts/microbench/f32x2_throughput.cu: a stream of 2 FMUL + FADD2 reaches 81 % of the FP32 pipe, FMUL2 + 2 FADD 90 %, FMUL + FADD
// 98 % but at one issue slot per operation.) Never both: ptxas contracts a packed product feeding a packed add into FFMA2 (common.cuh).
#ifndef SSYM_PACKED_MUL
#define SSYM_PACKED_MUL 1
#endif
// a frame pair's two values travel together
struct ssym2
{
   float lo, hi;
};
__device__ __forceinline__ ssym2 ssym_pack( float lo, float hi ) { return ssym2{ lo, hi }; }
__device__ __forceinline__ void ssym_unpack( ssym2 v, float &lo, float &hi ) { lo = v.lo; hi = v.hi; }
__device__ __forceinline__ ssym2 ssym_add2( ssym2 a, ssym2 b ) { return ssym2{ __fadd_rn( a.lo, b.lo ), __fadd_rn( a.hi, b.hi ) }; }
__device__ __forceinline__ ssym2 ssym_sub2( ssym2 a, ssym2 b ) { return ssym2{ __fsub_rn( a.lo, b.lo ), __fsub_rn( a.hi, b.hi ) }; }
// stft_tree8 with packed multiplies: (p0,p1) (p2,p3) (p4,p5) (p6,p7) are four FMUL2, the tree's seven additions scalar
__device__ __forceinline__ float ssym_tree8( const float4 xa, const float4 xb, const float4 b0, const float4 b1 )
{
#if SSYM_PACKED_MUL
   float p0, p1, p2, p3, p4, p5, p6, p7;
   unpk2( mul2( pk2( xa.x, xa.y ), pk2( b0.x, b0.y ) ), p0, p1 );
   unpk2( mul2( pk2( xa.z, xa.w ), pk2( b0.z, b0.w ) ), p2, p3 );
   unpk2( mul2( pk2( xb.x, xb.y ), pk2( b1.x, b1.y ) ), p4, p5 );
   unpk2( mul2( pk2( xb.z, xb.w ), pk2( b1.z, b1.w ) ), p6, p7 );
   return __fadd_rn( __fadd_rn( __fadd_rn( p0, p1 ), __fadd_rn( p2, p3 ) ), __fadd_rn( __fadd_rn( p4, p5 ), __fadd_rn( p6, p7 ) ) );
#else
   return stft_tree8( xa, xb, b0, b1 );
#endif
}
// stft_tree8 for two frames (x*, y*) that share the basis quads
__device__ __forceinline__ ssym2 ssym_tree8x2( const float4 xa, const float4 xb, const float4 ya, const float4 yb, const float4 b0, const float4 b1 )
{
   return ssym_pack( ssym_tree8( xa, xb, b0, b1 ), ssym_tree8( ya, yb, b0, b1 ) );
}

// out_mode 0: log1p(mag * 2^20) (production); 1: raw magnitude (parity tap for stft.c alone)
//
// One __syncthreads per chunk (at its end) hands the double-buffered input tile over.
template <bool F32>
__global__ void __launch_bounds__( SSYM_THREADS, 1 )
stft_sym_kernel( const void *__restrict__ in, long long stream_stride, int nw, int nchunks, const float *__restrict__ basis_sym /*[64][128][4]*/,
                 float *__restrict__ spec, int out_mode )
{
   extern __shared__ __align__( 16 ) float smem[];
   float *Bs = smem;
   float *Xs_all = smem + SSYM_BS_FLOATS; // [2 buffers][1792]
   float *B64 = Xs_all + 2 * STFT_XS_FLOATS; // [2 buffers][25][2]
   const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
   const int tg = w >> 2;                              // frames 5 tg .. 5 tg + 4
   const int u = ( ( w & 3 ) << 4 ) | ( lane & 15 );   // unit: bins u and 128 - u
   const int half = lane >> 4;                         // lanes 4 half .. 4 half + 3 of the tree
   const bool special = u == 0;                        // rows: re of bin 0, merged row of bin 64
   {
      const float4 *src = reinterpret_cast<const float4 *>( basis_sym );
      float4 *dst = reinterpret_cast<float4 *>( Bs );
      for ( int i = tid; i < SSYM_BS_FLOATS / 4; i += SSYM_THREADS ) dst[i] = __ldg( src + i );
   }
   // Staging: a chunk is NV 16-byte vectors; lanes 0..23 of the first NV / 24 warps take one each (8 or 16 warps: the same number on
   // every scheduler). The tile of chunk n + 1 is written at the START of iteration n (from registers filled an iteration earlier; the
   // vectors of chunk n + 2 are requested right after) while the other warps are already in the main loop. One barrier per chunk, at
   // its end, hands the tile over and frees the other one.
   // Tried and measured slower (r02): letting warps run apart so that one group's epilogue (sqrt, log1p: mostly the other pipes) overlaps
   // the other group's main loop (all FP32 pipe) -- a ring of four tiles with mbarrier hand-over: 15.8 / 14.2 ms against 13.0 ms; two
   // groups skewed by the epilogue around the same barrier: 10.4 ms against 9.7 ms, with the epilogue rolled or unrolled. The main loop
   // only reaches its 86 % of the FP32 pipe with all five warps of a scheduler inside it.
   constexpr int NV = F32 ? 384 : 192; // 16-byte vectors per chunk
   const int sv = w * 24 + lane;       // this lane's vector
   const bool stager = lane < 24 && w < NV / 24;
   int4 raw = make_int4( 0, 0, 0, 0 );
   auto stage = [&]( float *xs ) {
      if ( !stager ) return;
      if ( F32 )
      {
         const float *f = reinterpret_cast<const float *>( &raw );
#pragma unroll
         for ( int e = 0; e < 4; ++e ) stft_put( xs, 4 * sv + e, f[e] );
      }
      else
      {
         const short *hh = reinterpret_cast<const short *>( &raw );
         // (float)s16 / 32768.0f (vadc.c:884,898); the division by a power of two is exact
#pragma unroll
         for ( int e = 0; e < 8; ++e ) stft_put( xs, 8 * sv + e, (float)hh[e] * ( 1.0f / 32768.0f ) );
      }
   };
   // (a predicated load straight into `raw`: a conditional assignment made the compiler wait for the load at a register move)
   auto request = [&]( int c ) {
      const bool on = c < nchunks && stager;
      const int4 *ptr = on ? (const int4 *)stft_chunk_ptr<F32>( in, stream_stride, nw, c ) + sv : (const int4 *)in;
      asm volatile( "{\n .reg .pred p;\n setp.ne.s32 p, %4, 0;\n @p ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%5];\n}"
                    : "+r"( raw.x ), "+r"( raw.y ), "+r"( raw.z ), "+r"( raw.w )
                    : "r"( (int)on ), "l"( ptr ) );
   };
   int ci = blockIdx.x;
   request( ci );
   if ( ci < nchunks ) stage( Xs_all );
   request( ci + gridDim.x );
   __syncthreads();

   // bin 64 of chunk `c` from the pairs the unit-0 lanes left in B64[pbuf] (a barrier after the last of them has been passed)
   auto finish_bin64 = [&]( int c, int pbuf ) {
      if ( w == SSYM_B64_WARP && lane < VB_FRAMES )
      {
         const float2 v = *reinterpret_cast<const float2 *>( B64 + ( pbuf * VB_FRAMES + lane ) * 2 );
         const float m2 = sqrtf( __fadd_rn( __fmul_rn( v.x, v.x ), __fmul_rn( v.y, v.y ) ) );
         spec[(size_t)c * ( VB_BINS * VB_FRAMES ) + 64 * VB_FRAMES + lane] = out_mode ? m2 : lme::log1pf_ref( __fmul_rn( m2, 1048576.0f ) );
      }
   };

   int buf = 0, it = 0, cprev = -1;
   for ( ; ci < nchunks; cprev = ci, ci += gridDim.x, buf ^= 1, ++it )
   {
      const float *xs = Xs_all + buf * STFT_XS_FLOATS;
      if ( ci + gridDim.x < nchunks )
      {
         stage( Xs_all + ( buf ^ 1 ) * STFT_XS_FLOATS );
         request( ci + 2 * gridDim.x );
      }
      if ( cprev >= 0 ) finish_bin64( cprev, ( it - 1 ) & 1 );

      // Frames go through in pairs (0,1), (2,3) + frame 4: the two frames of a pair share every basis value. Per 8-tap group and
      // row: four FMUL2 + seven FADD; the FP32 pipe paces the loop (155 pipe cycles per 123 issue slots), it is 86 % busy inside it.
      ssym2 Sp[2][2], Sm[2][2]; // [row][pair]
      float Sp4[2], Sm4[2];     // frame 4
#pragma unroll 1
      for ( int lp2 = 0; lp2 < 2; ++lp2 )
      {
         ssym2 TL[2][2];
         float TL4[2];
#pragma unroll
         for ( int lo = 0; lo < 2; ++lo )
         {
            const int l = 4 * half + 2 * lp2 + lo;
            ssym2 A[2][2], Bv[2][2];
            float A4[2], B4[2];
#pragma unroll
            for ( int g = 0; g < 4; ++g )
            {
               const float *bq = Bs + ( ( l * 8 + g * 2 ) * 128 + u ) * 4;
               const float4 re0 = ld4( bq ), re1 = ld4( bq + 128 * 4 );
               const float4 im0 = ld4( bq + 64 * 4 ), im1 = ld4( bq + 64 * 4 + 128 * 4 );
#pragma unroll
               for ( int pr = 0; pr < 2; ++pr )
               {
                  const float *xp = xs + ( 5 * tg + 2 * pr + g ) * 64 + l * 8;
                  const float4 xa = ld4( xp ), xb = ld4( xp + 4 ), ya = ld4( xp + 64 ), yb = ld4( xp + 68 );
                  const ssym2 rr = ssym_tree8x2( xa, xb, ya, yb, re0, re1 );
                  const ssym2 ri = ssym_tree8x2( xa, xb, ya, yb, im0, im1 );
                  if ( g == 0 ) { A[0][pr] = rr; A[1][pr] = ri; }
                  else if ( g == 1 ) { A[0][pr] = ssym_add2( A[0][pr], rr ); A[1][pr] = ssym_add2( A[1][pr], ri ); }
                  else if ( g == 2 ) { Bv[0][pr] = rr; Bv[1][pr] = ri; }
                  else { Bv[0][pr] = ssym_add2( Bv[0][pr], rr ); Bv[1][pr] = ssym_add2( Bv[1][pr], ri ); }
               }
               {
                  const float *xp = xs + ( 5 * tg + 4 + g ) * 64 + l * 8;
                  const float4 xa = ld4( xp ), xb = ld4( xp + 4 );
                  const float rr = ssym_tree8( xa, xb, re0, re1 );
                  const float ri = ssym_tree8( xa, xb, im0, im1 );
                  if ( g == 0 ) { A4[0] = rr; A4[1] = ri; }
                  else if ( g == 1 ) { A4[0] = __fadd_rn( A4[0], rr ); A4[1] = __fadd_rn( A4[1], ri ); }
                  else if ( g == 2 ) { B4[0] = rr; B4[1] = ri; }
                  else { B4[0] = __fadd_rn( B4[0], rr ); B4[1] = __fadd_rn( B4[1], ri ); }
               }
            }
#pragma unroll
            for ( int a = 0; a < 2; ++a )
            {
               // plain and alternating sum of the lane pair; the merged row of bin 64 keeps its even and odd lanes apart
               const bool split = special && a == 1;
#pragma unroll
               for ( int pr = 0; pr < 2; ++pr )
               {
                  const ssym2 R = ssym_add2( A[a][pr], Bv[a][pr] );
                  if ( lo == 0 )
                     TL[a][pr] = R;
                  else
                  {
                     const ssym2 sum = ssym_add2( TL[a][pr], R ), dif = ssym_sub2( TL[a][pr], R );
                     const ssym2 vp = split ? TL[a][pr] : sum, vm = split ? R : dif;
                     if ( lp2 == 0 ) { Sp[a][pr] = vp; Sm[a][pr] = vm; }
                     else { Sp[a][pr] = ssym_add2( Sp[a][pr], vp ); Sm[a][pr] = ssym_add2( Sm[a][pr], vm ); }
                  }
               }
               const float R4 = __fadd_rn( A4[a], B4[a] );
               if ( lo == 0 )
                  TL4[a] = R4;
               else
               {
                  const float vp = split ? TL4[a] : __fadd_rn( TL4[a], R4 ), vm = split ? R4 : __fsub_rn( TL4[a], R4 );
                  if ( lp2 == 0 ) { Sp4[a] = vp; Sm4[a] = vm; }
                  else { Sp4[a] = __fadd_rn( Sp4[a], vp ); Sm4[a] = __fadd_rn( Sm4[a], vm ); }
               }
            }
         }
      }
      float Sps[2][5], Sms[2][5]; // back to one value per frame
#pragma unroll
      for ( int a = 0; a < 2; ++a )
      {
#pragma unroll
         for ( int pr = 0; pr < 2; ++pr )
         {
            ssym_unpack( Sp[a][pr], Sps[a][2 * pr], Sps[a][2 * pr + 1] );
            ssym_unpack( Sm[a][pr], Sms[a][2 * pr], Sms[a][2 * pr + 1] );
         }
         Sps[a][4] = Sp4[a];
         Sms[a][4] = Sm4[a];
      }

      // lanes 0..3 (half 0) + lanes 4..7 (half 1): the half-0 thread takes the plain sums (bin u), the half-1 thread the alternating ones (bin 128-u)
      float *o = spec + (size_t)ci * ( VB_BINS * VB_FRAMES ) + 5 * tg + ( half ? 128 - u : u ) * VB_FRAMES;
      float yr[5], yi[5];
#pragma unroll
      for ( int i = 0; i < 5; ++i )
      {
         const float g0 = __shfl_xor_sync( 0xffffffffu, half ? Sps[0][i] : Sms[0][i], 16 );
         const float g1 = __shfl_xor_sync( 0xffffffffu, half ? Sps[1][i] : Sms[1][i], 16 );
         yr[i] = half ? __fadd_rn( g0, Sms[0][i] ) : __fadd_rn( Sps[0][i], g0 );
         yi[i] = half ? __fadd_rn( g1, Sms[1][i] ) : __fadd_rn( Sps[1][i], g1 );
      }
      // magnitude + log1p per frame in a ROLLED loop (the values rotate through yr[0], yi[0]): a fifth of the code -- the kernel's hot
      // code stays inside the 32 KB instruction cache of the SM
      float *b64 = B64 + ( ( it & 1 ) * VB_FRAMES + 5 * tg ) * 2 + half;
#pragma unroll 1
      for ( int i = 0; i < 5; ++i )
      {
         const float re = yr[0], y1 = yi[0];
#pragma unroll
         for ( int k = 0; k < 4; ++k )
         {
            yr[k] = yr[k + 1];
            yi[k] = yi[k + 1];
         }
         // unit 0: half 0 holds re(bin 0), re(bin 64); half 1 holds re(bin 128), im(bin 64)
         if ( special ) b64[2 * i] = y1;
         const float im = special ? 0.0f : y1;
         const float m = sqrtf( __fadd_rn( __fmul_rn( re, re ), __fmul_rn( im, im ) ) );
         o[i] = out_mode ? m : lme::log1pf_ref( __fmul_rn( m, 1048576.0f ) );
      }
      __syncthreads(); // the next tile is complete, this one and the other B64 buffer are free
   }
   if ( cprev >= 0 ) finish_bin64( cprev, ( it - 1 ) & 1 );
}

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
// compute thread = (unit u = bin pair (u, 128-u), half of the eight lanes, group of five frames): 64 x 2 x 5 = 640 threads (+ 32 that
// stage the input, see the kernel). A thread walks
// its four lanes exactly like stft_kernel.cuh walks all eight (tree state for 2 rows x 5 frames in registers), keeps the plain and
// the alternating pair sums, and meets its partner (lane ^ 16: same unit, other half) in one shuffle per value: the half-0 thread
// finishes bin u, the half-1 thread bin 128-u -- magnitude and log1p are spread over all 640 threads.
#pragma once
#include "common.cuh"
#include "libm_exact.cuh"
#include "stft_kernel.cuh"

#define SSYM_COMPUTE 640
#define SSYM_THREADS ( SSYM_COMPUTE + 32 )     // + one producer warp
#define SSYM_NBUF 3                             // ring of chunk tiles
#define SSYM_BS_FLOATS ( 64 * 128 * 4 )
#define SSYM_SMEM_BYTES ( ( SSYM_BS_FLOATS + SSYM_NBUF * STFT_XS_FLOATS ) * 4 )

__device__ __forceinline__ uint32_t ssym_smem_u32( const void *p ) { return (uint32_t)__cvta_generic_to_shared( p ); }
__device__ __forceinline__ void ssym_mbar_init( uint64_t *bar, uint32_t count )
{
   asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( ssym_smem_u32( bar ) ), "r"( count ) : "memory" );
}
__device__ __forceinline__ void ssym_mbar_arrive( uint64_t *bar )
{
   asm volatile( "{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"( ssym_smem_u32( bar ) ) : "memory" );
}
__device__ __forceinline__ void ssym_mbar_wait( uint64_t *bar, uint32_t parity )
{
   uint32_t ok;
   do
   {
      asm volatile( "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"( ok )
                    : "r"( ssym_smem_u32( bar ) ), "r"( parity )
                    : "memory" );
   } while ( !ok );
}

// out_mode 0: log1p(mag * 2^20) (production); 1: raw magnitude (parity tap for stft.c alone)
//
// Warps 0..19 compute, warp 20 produces: it converts the PCM of chunk n+1, n+2 into padded, permuted tiles (a ring of SSYM_NBUF) while the
// compute warps work on chunk n. Tiles change hands through mbarriers (full[b]: the 32 producer lanes arrive; empty[b]: the 640 compute
// threads arrive after their last read), so no warp ever waits for the slowest of the other nineteen: with one __syncthreads per chunk
// 16 % of all stall samples sat on that barrier (profiles/ncu_summary_exact_r02a.md).
// (672 threads x 96 registers = 64 512 of the SM's 65 536: __launch_bounds__ would round the block up to 768 threads and cap at 80)
template <bool F32>
__global__ void __maxnreg__( 96 )
stft_sym_kernel( const void *__restrict__ in, long long stream_stride, int nw, int nchunks, const float *__restrict__ basis_sym /*[64][128][4]*/,
                 float *__restrict__ spec, int out_mode )
{
   extern __shared__ __align__( 16 ) float smem[];
   __shared__ __align__( 8 ) uint64_t bar_full[SSYM_NBUF], bar_empty[SSYM_NBUF];
   float *Bs = smem;
   float *Xs_all = smem + SSYM_BS_FLOATS; // [SSYM_NBUF][1792]
   const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
   {
      const float4 *src = reinterpret_cast<const float4 *>( basis_sym );
      float4 *dst = reinterpret_cast<float4 *>( Bs );
      for ( int i = tid; i < SSYM_BS_FLOATS / 4; i += SSYM_THREADS ) dst[i] = __ldg( src + i );
   }
   if ( tid == 0 )
   {
      for ( int b = 0; b < SSYM_NBUF; ++b )
      {
         ssym_mbar_init( &bar_full[b], 32 );
         ssym_mbar_init( &bar_empty[b], SSYM_COMPUTE );
      }
      asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" );
   }
   __syncthreads();

   if ( w == SSYM_COMPUTE / 32 )
   {
      // ---- producer warp: PCM -> tile ring ------------------------------------------------------------------------------
      constexpr int NV = F32 ? 384 : 192; // 16-byte vectors per chunk
      constexpr int PER = NV / 32;
      int4 raw[PER];
      int ci = blockIdx.x;
      if ( ci < nchunks )
      {
         const int4 *src = (const int4 *)stft_chunk_ptr<F32>( in, stream_stride, nw, ci );
#pragma unroll
         for ( int i = 0; i < PER; ++i ) raw[i] = __ldg( src + lane + 32 * i );
      }
      for ( int n = 0; ci < nchunks; ci += gridDim.x, ++n )
      {
         const int b = n % SSYM_NBUF;
         if ( n >= SSYM_NBUF ) ssym_mbar_wait( &bar_empty[b], ( ( n / SSYM_NBUF ) - 1 ) & 1 );
         float *xs = Xs_all + b * STFT_XS_FLOATS;
#pragma unroll
         for ( int i = 0; i < PER; ++i )
         {
            const int q = lane + 32 * i;
            if ( F32 )
            {
               const float *f = reinterpret_cast<const float *>( &raw[i] );
#pragma unroll
               for ( int e = 0; e < 4; ++e ) stft_put( xs, 4 * q + e, f[e] );
            }
            else
            {
               const short *hh = reinterpret_cast<const short *>( &raw[i] );
               // (float)s16 / 32768.0f (vadc.c:884,898); the division by a power of two is exact
#pragma unroll
               for ( int e = 0; e < 8; ++e ) stft_put( xs, 8 * q + e, (float)hh[e] * ( 1.0f / 32768.0f ) );
            }
         }
         ssym_mbar_arrive( &bar_full[b] );
         const int cn = ci + gridDim.x;
         if ( cn < nchunks )
         {
            const int4 *src = (const int4 *)stft_chunk_ptr<F32>( in, stream_stride, nw, cn );
#pragma unroll
            for ( int i = 0; i < PER; ++i ) raw[i] = __ldg( src + lane + 32 * i );
         }
      }
      return;
   }

   // ---- compute warps ---------------------------------------------------------------------------------------------------
   const int tg = w >> 2;                              // frames 5 tg .. 5 tg + 4
   const int u = ( ( w & 3 ) << 4 ) | ( lane & 15 );   // unit: bins u and 128 - u
   const int half = lane >> 4;                         // lanes 4 half .. 4 half + 3 of the tree
   const bool special = u == 0;                        // rows: re of bin 0, merged row of bin 64
   int n = 0;
   for ( int ci = blockIdx.x; ci < nchunks; ci += gridDim.x, ++n )
   {
      const int b = n % SSYM_NBUF;
      const float *xs = Xs_all + b * STFT_XS_FLOATS;
      ssym_mbar_wait( &bar_full[b], ( n / SSYM_NBUF ) & 1 );

      float Sp[2][5], Sm[2][5];
#pragma unroll 1
      for ( int lp2 = 0; lp2 < 2; ++lp2 )
      {
         float TL[2][5];
#pragma unroll
         for ( int lo = 0; lo < 2; ++lo )
         {
            const int l = 4 * half + 2 * lp2 + lo;
            float A[2][5], Bv[2][5];
#pragma unroll
            for ( int g = 0; g < 4; ++g )
            {
               const float *bq = Bs + ( ( l * 8 + g * 2 ) * 128 + u ) * 4;
               const float4 re0 = ld4( bq ), re1 = ld4( bq + 128 * 4 );
               const float4 im0 = ld4( bq + 64 * 4 ), im1 = ld4( bq + 64 * 4 + 128 * 4 );
#pragma unroll
               for ( int i = 0; i < 5; ++i )
               {
                  const float *xp = xs + ( 5 * tg + i + g ) * 64 + l * 8;
                  const float4 xa = ld4( xp ), xb = ld4( xp + 4 );
                  const float rr = stft_tree8( xa, xb, re0, re1 );
                  const float ri = stft_tree8( xa, xb, im0, im1 );
                  if ( g == 0 ) { A[0][i] = rr; A[1][i] = ri; }
                  else if ( g == 1 ) { A[0][i] = __fadd_rn( A[0][i], rr ); A[1][i] = __fadd_rn( A[1][i], ri ); }
                  else if ( g == 2 ) { Bv[0][i] = rr; Bv[1][i] = ri; }
                  else { Bv[0][i] = __fadd_rn( Bv[0][i], rr ); Bv[1][i] = __fadd_rn( Bv[1][i], ri ); }
               }
            }
#pragma unroll
            for ( int a = 0; a < 2; ++a )
#pragma unroll
               for ( int i = 0; i < 5; ++i )
               {
                  const float R = __fadd_rn( A[a][i], Bv[a][i] );
                  if ( lo == 0 )
                     TL[a][i] = R;
                  else
                  {
                     // plain and alternating sum of the lane pair; the merged row of bin 64 keeps its even and odd lanes apart
                     const bool split = special && a == 1;
                     const float pr = split ? TL[a][i] : __fadd_rn( TL[a][i], R );
                     const float pm = split ? R : __fsub_rn( TL[a][i], R );
                     if ( lp2 == 0 ) { Sp[a][i] = pr; Sm[a][i] = pm; }
                     else { Sp[a][i] = __fadd_rn( Sp[a][i], pr ); Sm[a][i] = __fadd_rn( Sm[a][i], pm ); }
                  }
               }
         }
      }
      ssym_mbar_arrive( &bar_empty[b] ); // (every load of the tile has been consumed by the arithmetic above)

      // lanes 0..3 (half 0) + lanes 4..7 (half 1): the half-0 thread takes the plain sums (bin u), the half-1 thread the alternating ones (bin 128-u)
      float *o = spec + (size_t)ci * ( VB_BINS * VB_FRAMES ) + 5 * tg;
      const int bin = half ? 128 - u : u;
#pragma unroll
      for ( int i = 0; i < 5; ++i )
      {
         float y[2];
#pragma unroll
         for ( int a = 0; a < 2; ++a )
         {
            const float got = __shfl_xor_sync( 0xffffffffu, half ? Sp[a][i] : Sm[a][i], 16 );
            y[a] = half ? __fadd_rn( got, Sm[a][i] ) : __fadd_rn( Sp[a][i], got );
         }
         // unit 0: half 0 holds re(bin 0), re(bin 64); half 1 holds re(bin 128), im(bin 64)
         const float im64 = __shfl_xor_sync( 0xffffffffu, y[1], 16 );
         const float re = y[0], im = special ? 0.0f : y[1];
         const float m = sqrtf( __fadd_rn( __fmul_rn( re, re ), __fmul_rn( im, im ) ) );
         o[bin * VB_FRAMES + i] = out_mode ? m : lme::log1pf_ref( __fmul_rn( m, 1048576.0f ) );
         if ( special && half == 0 )
         {
            const float m2 = sqrtf( __fadd_rn( __fmul_rn( y[1], y[1] ), __fmul_rn( im64, im64 ) ) );
            o[64 * VB_FRAMES + i] = out_mode ? m2 : lme::log1pf_ref( __fmul_rn( m2, 1048576.0f ) );
         }
      }
   }
}

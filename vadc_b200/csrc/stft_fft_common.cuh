// vadc_b200/csrc/stft_fft_common.cuh -- helpers of the FFT-hybrid STFT kernel (stft_fft8_kernel.cuh): input tile staging, the
// warp-cooperative exact row of flagged bins, fast magnitude / log. (The first FFT kernel, one warp per frame, lived here in round 1;
// stft_fft8_kernel.cuh superseded it.)
// the reference's exact reduction tree only where it matters.
//
// Replaces my_stft (stft.c:15-229) + the log1p of adaptive_audio_normalization_inplace
// (misc.c:40-46), like stft_kernel.cuh, but ~10x cheaper.
//
// Why this is allowed: the parity bar is on probabilities (1e-4). A magnitude m
// computed with absolute error d enters the network as log1p(m*2^20), i.e. with error ~d/m. For
// any fp32 evaluation of the 256-tap correlation d ~ eps*||frame||; it only matters at bins with
// m << ||frame||. So every frame is transformed with a 256-point real FFT in fp32 (one warp per
// frame, 4 complex points per lane, radix-4 in registers + 5 shuffle stages), and every bin whose
// magnitude is below tau = K*||windowed frame||_2 is re-evaluated with the reference's own rounding
// sequence (stft.c:108-184: 256 rounded products, AVX2 tree order, no FMA) by the whole warp:
// lane = (l, g) owns one 8-tap leaf, the g- and l-combines are xor-shuffle butterflies, which is the
// same tree because fp32 addition is commutative. Those bins are bit-identical to the reference.
// Measured on the CPU restatement (K = 3e-3): 0.3 % of bins take the exact path and probabilities
// stay within 2e-5 of the reference (pure FFT without the fix-up: 4e-4, i.e. outside the bar).
// Degenerate inputs (pure tones, DC) flag most bins and degrade towards the cost of the exact
// kernel, never in accuracy.
#pragma once
#include "common.cuh"

#define HYB_WARPS 5
#define HYB_THREADS ( HYB_WARPS * 32 )
#define HYB_XS_FLOATS 1792
#define HYB_OUT_FLOATS ( VB_BINS * VB_FRAMES )
#define HYB_SMEM_BYTES ( ( 2 * HYB_XS_FLOATS + HYB_OUT_FLOATS + 32 ) * 4 )

// fast paths for the bins that are NOT re-evaluated exactly: their magnitude already carries ~1e-4 relative error
// (eps * ||frame|| / m), so a 1-ulp square root and a 2^-22-relative logarithm change nothing measurable.
__device__ __forceinline__ float hyb_sqrt_fast( float v )
{
   float r;
   asm( "sqrt.approx.ftz.f32 %0, %1;" : "=f"( r ) : "f"( v ) );
   return r;
}
// log1p(m * 2^20) (misc.c:40-46) as log(1 + x): the absolute error of forming 1 + x (<= 6e-8 * (1 + x)) is what the
// network sees (the value enters linearly), and the bins where log1p's relative accuracy at tiny x would matter are
// exactly zero or ~1e-6, i.e. 1e-6 absolute either way
__device__ __forceinline__ float hyb_log1p_scaled( float m ) { return __logf( fmaf( m, 1048576.0f, 1.0f ) ); }

__device__ __forceinline__ int brev5( int j ) { return (int)( __brev( (unsigned)j ) >> 27 ); }

struct cpx
{
   float re, im;
};
__device__ __forceinline__ cpx cmul( cpx a, cpx w ) { return cpx{ fmaf( a.re, w.re, -a.im * w.im ), fmaf( a.re, w.im, a.im * w.re ) }; }

// the reference's 256-tap tree for basis row `row` at frame t, evaluated by one warp
// (lane = l*4 + g). Returns the same value in every lane. xs: padded chunk, natural order.
__device__ __forceinline__ float hyb_exact_row( const float *__restrict__ xs, const float *__restrict__ basis, int row, int t, int lane )
{
   const int l = lane >> 2, g = lane & 3;
   const float *xp = xs + 64 * t + 64 * g + l;
   const float *bp = basis + (size_t)row * 256 + 64 * g + l;
   float p[8];
#pragma unroll
   for ( int v = 0; v < 8; ++v ) p[v] = __fmul_rn( xp[8 * v], __ldg( bp + 8 * v ) );
   float s01 = __fadd_rn( p[0], p[1] ), s23 = __fadd_rn( p[2], p[3] ), s45 = __fadd_rn( p[4], p[5] ), s67 = __fadd_rn( p[6], p[7] );
   float r = __fadd_rn( __fadd_rn( s01, s23 ), __fadd_rn( s45, s67 ) );
   r = __fadd_rn( r, __shfl_xor_sync( 0xffffffffu, r, 1 ) );  // r0+r1 | r2+r3
   r = __fadd_rn( r, __shfl_xor_sync( 0xffffffffu, r, 2 ) );  // R_l
   r = __fadd_rn( r, __shfl_xor_sync( 0xffffffffu, r, 4 ) );  // R0+R1, R2+R3, ...
   r = __fadd_rn( r, __shfl_xor_sync( 0xffffffffu, r, 8 ) );  // (R0+R1)+(R2+R3), (R4+R5)+(R6+R7)
   r = __fadd_rn( r, __shfl_xor_sync( 0xffffffffu, r, 16 ) ); // y
   return r;
}

// exact magnitude of bin f at frame t (stft.c:194-213)
__device__ __forceinline__ float hyb_exact_mag( const float *xs, const float *basis, int f, int t, int lane )
{
   float re = hyb_exact_row( xs, basis, f, t, lane );
   float im = hyb_exact_row( xs, basis, 129 + f, t, lane );
   return sqrtf( __fadd_rn( __fmul_rn( re, re ), __fmul_rn( im, im ) ) );
}

// sample m (0..1535) -> padded tile (natural order) incl. its reflect-padding images (tensor.h:942-953)
__device__ __forceinline__ void hyb_put( float *xs, int m, float v )
{
   xs[128 + m] = v;
   if ( m >= 1 && m <= 128 ) xs[128 - m] = v;
   if ( m >= 1407 && m <= 1534 ) xs[3198 - m] = v;
}

template <bool F32>
struct HybRaw
{
   int4 v[F32 ? 3 : 2];
};

template <bool F32>
__device__ __forceinline__ void hyb_load_raw( HybRaw<F32> &raw, const void *chunk, int tid )
{
   constexpr int NV = F32 ? 384 : 192;
   constexpr int PER = F32 ? 3 : 2;
#pragma unroll
   for ( int i = 0; i < PER; ++i )
   {
      int q = tid + i * HYB_THREADS;
      if ( q < NV ) raw.v[i] = __ldg( (const int4 *)chunk + q );
   }
}

template <bool F32>
__device__ __forceinline__ void hyb_store_x( float *xs, const HybRaw<F32> &raw, int tid )
{
   constexpr int NV = F32 ? 384 : 192;
   constexpr int PER = F32 ? 3 : 2;
#pragma unroll
   for ( int i = 0; i < PER; ++i )
   {
      int q = tid + i * HYB_THREADS;
      if ( q < NV )
      {
         if ( F32 )
         {
            const float *f = reinterpret_cast<const float *>( &raw.v[i] );
#pragma unroll
            for ( int e = 0; e < 4; ++e ) hyb_put( xs, 4 * q + e, f[e] );
         }
         else
         {
            const short *h = reinterpret_cast<const short *>( &raw.v[i] );
#pragma unroll
            for ( int e = 0; e < 8; ++e ) hyb_put( xs, 8 * q + e, (float)h[e] * ( 1.0f / 32768.0f ) ); // vadc.c:884,898
         }
      }
   }
}

template <bool F32>
__device__ __forceinline__ const void *hyb_chunk_ptr( const void *in, long long stream_stride, int nw, int ci )
{
   int s = ci / nw, n = ci - s * nw;
   long long off = (long long)s * stream_stride + (long long)n * VB_CHUNK;
   return F32 ? (const void *)( (const float *)in + off ) : (const void *)( (const int16_t *)in + off );
}

// vadc_b200/csrc/stft_hybrid_kernel.cuh -- STFT + magnitude + log1p: fast transform everywhere,
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

// basis: the reference's forward_basis_buffer [258][256] (row 0 is the periodic Hann window).
// k_rel: fix-up threshold relative to ||windowed frame||_2. out_mode 0: log1p(m*2^20); 1: m.
// flagged: optional global counter of bins that took the exact path (statistics).
template <bool F32>
__global__ void __launch_bounds__( HYB_THREADS )
stft_hybrid_kernel( const void *__restrict__ in, long long stream_stride, int nw, int nchunks, const float *__restrict__ basis,
                    float *__restrict__ spec, float *__restrict__ mu_out, float k_rel, int out_mode, unsigned long long *__restrict__ flagged )
{
   extern __shared__ __align__( 16 ) float smem[];
   float *Xs = smem;                         // [2][1792]
   float *Os = smem + 2 * HYB_XS_FLOATS;     // [129][25]
   float *Ms = Os + HYB_OUT_FLOATS;          // [25] per-frame mean of the log spectrum

   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int j = lane;

   // ---- per-lane constants ------------------------------------------------------------------
   float win[8];
#pragma unroll
   for ( int r = 0; r < 4; ++r )
   {
      win[2 * r] = __ldg( basis + 2 * j + 64 * r );
      win[2 * r + 1] = __ldg( basis + 2 * j + 64 * r + 1 );
   }
   cpx tw4[3]; // W128^(j*q), q = 1..3
#pragma unroll
   for ( int q = 1; q <= 3; ++q )
   {
      float s, c;
      sincospif( -(float)( j * q ) / 64.0f, &s, &c );
      tw4[q - 1] = cpx{ c, s };
   }
   cpx tws[3]; // shuffle stages h = 16, 8, 4: W_(2h)^(j mod h) on the upper half, 1 on the lower
#pragma unroll
   for ( int si = 0; si < 3; ++si )
   {
      const int h = 16 >> si;
      float s, c;
      sincospif( -(float)( j & ( h - 1 ) ) / (float)h, &s, &c );
      tws[si] = ( j & h ) ? cpx{ c, s } : cpx{ 1.0f, 0.0f };
   }
   const int kp = brev5( j );                // after the 5 DIF stages lane j holds index k' = brev5(j)
   const int p0 = brev5( ( 32 - kp ) & 31 ); // lane holding Z_0[(32-k') mod 32]
   float pc[4], ps[4];                       // cos, sin of 2*pi*k/256 for k = 4k'+q
#pragma unroll
   for ( int q = 0; q < 4; ++q ) sincospif( (float)( 4 * kp + q ) / 128.0f, &ps[q], &pc[q] );

   HybRaw<F32> raw;
   int ci = blockIdx.x;
   if ( ci < nchunks ) hyb_load_raw<F32>( raw, hyb_chunk_ptr<F32>( in, stream_stride, nw, ci ), tid );
   unsigned nflag = 0;

   int buf = 0;
   for ( ; ci < nchunks; ci += gridDim.x, buf ^= 1 )
   {
      float *xs = Xs + buf * HYB_XS_FLOATS;
      hyb_store_x<F32>( xs, raw, tid );
      __syncthreads(); // tile complete; also: everyone finished copying the previous Os
      int cn = ci + gridDim.x;
      if ( cn < nchunks ) hyb_load_raw<F32>( raw, hyb_chunk_ptr<F32>( in, stream_stride, nw, cn ), tid );

#pragma unroll 1
      for ( int t = warp; t < VB_FRAMES; t += HYB_WARPS )
      {
         // ---- windowed frame, packed as z[n] = y[2n] + i y[2n+1], n = j + 32 r -------------------
         const float *xf = xs + 64 * t;
         cpx z[4];
         float e2 = 0.0f;
#pragma unroll
         for ( int r = 0; r < 4; ++r )
         {
            float2 v = *reinterpret_cast<const float2 *>( xf + 2 * j + 64 * r );
            z[r].re = v.x * win[2 * r];
            z[r].im = v.y * win[2 * r + 1];
            e2 = fmaf( z[r].re, z[r].re, e2 );
            e2 = fmaf( z[r].im, z[r].im, e2 );
         }
#pragma unroll
         for ( int off = 16; off > 0; off >>= 1 ) e2 += __shfl_xor_sync( 0xffffffffu, e2, off );
         const float tau = k_rel * sqrtf( e2 );

         // ---- radix-4 DIF over r, then twiddle by W128^(j q) ------------------------------------
         cpx a[4];
         {
            cpx t0{ z[0].re + z[2].re, z[0].im + z[2].im }, t1{ z[0].re - z[2].re, z[0].im - z[2].im };
            cpx t2{ z[1].re + z[3].re, z[1].im + z[3].im }, t3{ z[1].re - z[3].re, z[1].im - z[3].im };
            a[0] = cpx{ t0.re + t2.re, t0.im + t2.im };
            a[2] = cmul( cpx{ t0.re - t2.re, t0.im - t2.im }, tw4[1] );
            a[1] = cmul( cpx{ t1.re + t3.im, t1.im - t3.re }, tw4[0] );
            a[3] = cmul( cpx{ t1.re - t3.im, t1.im + t3.re }, tw4[2] );
         }
         // ---- four 32-point DIF FFTs across the lanes --------------------------------------------
#pragma unroll
         for ( int si = 0; si < 5; ++si )
         {
            const int h = 16 >> si;
            const float sgn = ( j & h ) ? -1.0f : 1.0f;
#pragma unroll
            for ( int q = 0; q < 4; ++q )
            {
               float orr = __shfl_xor_sync( 0xffffffffu, a[q].re, h );
               float oi = __shfl_xor_sync( 0xffffffffu, a[q].im, h );
               cpx d{ fmaf( sgn, a[q].re, orr ), fmaf( sgn, a[q].im, oi ) }; // lower: a+o ; upper: o-a
               if ( si < 3 )
                  a[q] = cmul( d, tws[si] );
               else if ( si == 3 )
                  a[q] = ( ( j & 3 ) == 3 ) ? cpx{ d.im, -d.re } : d; // W4^1 = -i on (upper, odd)
               else
                  a[q] = d;
            }
         }
         // ---- real-input post-processing: Y[k] from Z[k] and conj(Z[128-k]) ----------------------
         float mag[4], nyq = 0.0f;
         {
            cpx zp[4];
            zp[0].re = __shfl_sync( 0xffffffffu, a[0].re, p0 );
            zp[0].im = __shfl_sync( 0xffffffffu, a[0].im, p0 );
            zp[1].re = __shfl_sync( 0xffffffffu, a[3].re, 31 - j );
            zp[1].im = __shfl_sync( 0xffffffffu, a[3].im, 31 - j );
            zp[2].re = __shfl_sync( 0xffffffffu, a[2].re, 31 - j );
            zp[2].im = __shfl_sync( 0xffffffffu, a[2].im, 31 - j );
            zp[3].re = __shfl_sync( 0xffffffffu, a[1].re, 31 - j );
            zp[3].im = __shfl_sync( 0xffffffffu, a[1].im, 31 - j );
#pragma unroll
            for ( int q = 0; q < 4; ++q )
            {
               float er = 0.5f * ( a[q].re + zp[q].re ), ei = 0.5f * ( a[q].im - zp[q].im );
               float orr = 0.5f * ( a[q].re - zp[q].re ), oi = 0.5f * ( a[q].im + zp[q].im );
               float yr = er - ps[q] * orr + pc[q] * oi;
               float yi = ei - ps[q] * oi - pc[q] * orr;
               mag[q] = hyb_sqrt_fast( fmaf( yr, yr, yi * yi ) );
            }
            if ( j == 0 ) nyq = fabsf( a[0].re - a[0].im ); // Y[128] = Re Z0 - Im Z0
         }
         // ---- exact re-evaluation of small bins ---------------------------------------------------
#pragma unroll
         for ( int q = 0; q < 4; ++q )
         {
            unsigned m = __ballot_sync( 0xffffffffu, mag[q] < tau );
            while ( m )
            {
               int src = __ffs( m ) - 1;
               m &= m - 1;
               float ex = hyb_exact_mag( xs, basis, 4 * brev5( src ) + q, t, lane );
               if ( lane == src ) mag[q] = ex;
               ++nflag;
            }
         }
         if ( __shfl_sync( 0xffffffffu, nyq, 0 ) < tau )
         {
            float ex = hyb_exact_mag( xs, basis, 128, t, lane );
            if ( lane == 0 ) nyq = ex;
            ++nflag;
         }
         // ---- log1p(m * 2^20) (misc.c:40-46) into the chunk's output tile --------------------------
         float fsum = 0.0f;
#pragma unroll
         for ( int q = 0; q < 4; ++q )
         {
            const float lv = out_mode ? mag[q] : hyb_log1p_scaled( mag[q] );
            Os[( 4 * kp + q ) * VB_FRAMES + t] = lv;
            fsum += lv;
         }
         if ( j == 0 )
         {
            const float lv = out_mode ? nyq : hyb_log1p_scaled( nyq );
            Os[128 * VB_FRAMES + t] = lv;
            fsum += lv;
         }
         // per-frame mean over the 129 bins (misc.c:48-62) as a warp reduction: the values are already in registers.
         // (The summation order differs from the reference's sequential loop by ~1e-7 relative, far inside the budget
         // of everything downstream of the log)
#pragma unroll
         for ( int off = 16; off > 0; off >>= 1 ) fsum += __shfl_xor_sync( 0xffffffffu, fsum, off );
         if ( j == 0 ) Ms[t] = fsum / (float)VB_BINS;
      }
      __syncthreads();
      if ( warp == 0 && mu_out )
      {
         // the scalar of adaptive_audio_normalization_inplace (misc.c:48-82): per-frame means (computed by the frame
         // warps above), reflect-pad 3 + 7-tap smoothing, mean over the 25 frames
         const float gk[7] = { 0.03663284704089164733887f, 0.11128076165914535522461f, 0.21674531698226928710938f, 0.27068215608596801757812f,
                               0.21674531698226928710938f, 0.11128076165914535522461f, 0.03663284704089164733887f };
         const int tt = lane < VB_FRAMES ? lane : VB_FRAMES - 1;
         const float m = Ms[tt];
         float v = 0.0f;
#pragma unroll
         for ( int k = 0; k < 7; ++k )
         {
            int idx = tt + k - 3;
            if ( idx < 0 ) idx = -idx;
            if ( idx >= VB_FRAMES ) idx = 2 * ( VB_FRAMES - 1 ) - idx;
            v = __fadd_rn( v, __fmul_rn( __shfl_sync( 0xffffffffu, m, idx ), gk[k] ) );
         }
         float a = 0.0f;
         for ( int i = 0; i < VB_FRAMES; ++i ) a = __fadd_rn( a, __shfl_sync( 0xffffffffu, v, i ) );
         if ( lane == 0 ) mu_out[ci] = a / (float)VB_FRAMES;
      }
      float *o = spec + (size_t)ci * HYB_OUT_FLOATS;
      for ( int i = tid; i < HYB_OUT_FLOATS; i += HYB_THREADS ) o[i] = Os[i];
   }
   if ( flagged && lane == 0 && nflag ) atomicAdd( flagged, (unsigned long long)nflag );
}

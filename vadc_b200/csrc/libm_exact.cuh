// vadc_b200/csrc/libm_exact.cuh -- expf, tanhf and log1pf with the bits of the reference's C library.
//
// Why: the decoder LSTM (lstm.c:64-88) integrates c = f c + i g over thousands of steps. In long silences the gates sit at almost
// constant values with f within 1e-3 of 1, so a nonlinearity that differs from the reference's by a single ulp differs the SAME way
// on every step, and the cell state drifts away by ulp / (1 - f): measured 1.2e-3 in the speech probability at the next onset of a
// 3000-chunk stream with CUDA's own expf / tanhf (each within 2 ulp of the truth, like glibc's, but not the same 2 ulp), although
// every stage matched to 1e-5 when restarted from the reference's state (scripts/gpu_outlier_probe.py). The reference build links
// glibc (the pinned Linux build the parity tests compare against); these are its two algorithms, restated:
//   expf  -- the table method of glibc >= 2.27 (sysdeps/ieee754/flt-32/e_expf.c): k = round(32 x / ln 2) in double, 2^(k/32) from a
//            32-entry table, a cubic in the reduced argument, one rounding to float at the end;
//   tanhf -- fdlibm's float tanhf over expm1f (s_tanhf.c, s_expm1f.c), every operation a separately rounded float operation.
// Checked on the host against the C library itself over ALL 2.2e9 floats with |x| < 87 (expf) / < 30 (tanhf), and in round 2 on the
// GPU over the underflow range -104.5 .. -86 (tests/test_gpu_exact.py::test_expf_underflow_range): tanhf identical
// everywhere, expf identical except at two arguments (|x| = 32.56.. and 63.1.., where the library's FMA build rounds the double
// polynomial the other way). No fused multiply-add may be formed here: everything goes through the _rn intrinsics.
#pragma once
#include <stdint.h>
#if defined( __CUDACC__ )
#include <cuda_runtime.h>
#define LME_FN __device__ __forceinline__
#define LME_TAB __device__ __constant__ const
#else
// Host build (tests/hostcheck/libm_exhaustive.cpp: every function below swept over ALL floats of its domain against the C library):
// the intrinsics become plain IEEE operations; compile with -ffp-contract=off so that nothing is fused.
#include <math.h>
#include <string.h>
#define LME_FN static inline
#define LME_TAB static const
static inline float __fmul_rn( float a, float b ) { return a * b; }
static inline float __fadd_rn( float a, float b ) { return a + b; }
static inline float __fsub_rn( float a, float b ) { return a - b; }
static inline float __fdiv_rn( float a, float b ) { return a / b; }
static inline double __dmul_rn( double a, double b ) { return a * b; }
static inline double __dadd_rn( double a, double b ) { return a + b; }
static inline float __double2float_rn( double a ) { return (float)a; }
static inline uint32_t __float_as_uint( float a ) { uint32_t u; memcpy( &u, &a, 4 ); return u; }
static inline int __float_as_int( float a ) { int u; memcpy( &u, &a, 4 ); return u; }
static inline float __uint_as_float( uint32_t u ) { float a; memcpy( &a, &u, 4 ); return a; }
static inline float __int_as_float( int u ) { float a; memcpy( &a, &u, 4 ); return a; }
static inline long long __double_as_longlong( double a ) { long long u; memcpy( &u, &a, 8 ); return u; }
static inline double __longlong_as_double( long long u ) { double a; memcpy( &a, &u, 8 ); return a; }
#endif

namespace lme
{
// asuint64( 2^(i/32) ) - ( i << 47 ), i < 32
LME_TAB unsigned long long EXP2F_TAB[32] = {
   0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull, 0x3fef72b83c7d517bull, 0x3fef54873168b9aaull,
   0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull, 0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
   0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull, 0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull,
   0x3feea11473eb0187ull, 0x3feea589994cce13ull, 0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
   0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull, 0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full,
   0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull };

// tab: EXP2F_TAB, or a copy of it in shared memory (32 x 8 bytes): lanes index the table with different k, which the constant cache
// serves one address at a time (measured: 5 % of the LSTM kernel's stall samples sat on this load) while shared memory serves the
// whole warp in one or two wavefronts
LME_FN float expf_ref( float x, const unsigned long long *tab = EXP2F_TAB )
{
   const double InvLn2N = 0x1.71547652b82fep+0 * 32.0, SHIFT = 0x1.8p+52;
   const double C0 = 0x1.c6af84b912394p-5 / 32.0 / 32.0 / 32.0, C1 = 0x1.ebfce50fac4f3p-3 / 32.0 / 32.0, C2 = 0x1.62e42ff0c52d6p-1 / 32.0;
   // e_expf.c: x > log(2^128) overflows; x < log(2^-150) = -0x1.9fe368p6 underflows to 0; x < log(2^-149) = -0x1.9d1d9ep6 takes
   // __math_may_uflowf, whose 0x1.4p-75f squared rounds to the smallest denormal; in between the double result rounds to a denormal
   // (the three special ranges are SELECTED at the end, not branched to: the polynomial below is straight-line code that the compiler
   // can interleave with a neighbour's -- sigmoid3_ref -- and whatever it computes for an out-of-range argument is discarded)
   double z = __dmul_rn( InvLn2N, (double)x );
   double kd = __dadd_rn( z, SHIFT );
   const unsigned long long ki = (unsigned long long)__double_as_longlong( kd );
   kd = __dadd_rn( kd, -SHIFT );
   const double r = __dadd_rn( z, -kd );
   const double s = __longlong_as_double( (long long)( tab[ki & 31] + ( ki << 47 ) ) );
   z = __dadd_rn( __dmul_rn( C0, r ), C1 );
   const double r2 = __dmul_rn( r, r );
   double y = __dadd_rn( __dmul_rn( C2, r ), 1.0 );
   y = __dadd_rn( __dmul_rn( z, r2 ), y );
   y = __dmul_rn( y, s );
   float res = __double2float_rn( y );
   res = x < -0x1.9d1d9ep6f ? __int_as_float( 1 ) : res;
   res = x < -0x1.9fe368p6f ? 0.0f : res;
   return x > 88.72283f ? __int_as_float( 0x7f800000 ) : res;
}

LME_FN float fmul( float a, float b ) { return __fmul_rn( a, b ); }
LME_FN float fadd( float a, float b ) { return __fadd_rn( a, b ); }
LME_FN float fsub( float a, float b ) { return __fsub_rn( a, b ); }
LME_FN float fdiv( float a, float b ) { return __fdiv_rn( a, b ); }
LME_FN float word( uint32_t u ) { return __uint_as_float( u ); }

// fdlibm expm1f (finite arguments; callers pass |x| <= 44)
LME_FN float expm1f_ref( float x )
{
   const float one = 1.0f, huge = 1.0e+30f, tiny = 1.0e-30f, ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f, invln2 = 1.4426950216e+00f;
   const float Q1 = -3.3333335072e-02f, Q2 = 1.5873016091e-03f, Q3 = -7.9365076090e-05f, Q4 = 4.0082177293e-06f, Q5 = -2.0109921195e-07f;
   float y, hi, lo, c = 0.0f, t, e, hxs, hfx, r1;
   int k;
   uint32_t hx = __float_as_uint( x );
   const uint32_t xsb = hx & 0x80000000u;
   hx &= 0x7fffffffu;
   // fdlibm's two early returns, selected at the end like everything else
   const float x0 = x;
   const bool sat = hx >= 0x4195b844u && xsb != 0 && fadd( x, tiny ) < 0.0f; // x <= -27 ln 2: tiny - one
   const bool small = hx < 0x33000000u;                                       // |x| < 2^-25: x - ((huge + x) - (huge + x))
   // Argument reduction without divergence (the lanes of a warp hold pre-activations of all sizes, and every divergent path is
   // executed by the whole warp). fdlibm's three cases collapse into one formula: for 0.5 ln2 < |x| < 1.5 ln2 it sets k = +-1,
   // hi = x -+ ln2_hi, lo = +-ln2_lo -- exactly what the general case computes from t = (float)k = +-1 (1 * ln2_hi and 1 * ln2_lo are
   // exact), so only k itself needs the select; |x| <= 0.5 ln2 keeps x and c = 0 (k = 0).
   {
      const bool reduce = hx > 0x3eb17218u, near = hx < 0x3F851592u; // |x| > 0.5 ln 2, |x| < 1.5 ln 2
      const int kg = (int)fadd( fmul( invln2, x ), xsb == 0 ? 0.5f : -0.5f );
      k = reduce ? ( near ? ( xsb == 0 ? 1 : -1 ) : kg ) : 0;
      t = (float)k;
      hi = fsub( x, fmul( t, ln2_hi ) );
      lo = fmul( t, ln2_lo );
      const float xr = fsub( hi, lo );
      const float cr = fsub( fsub( hi, xr ), lo );
      x = reduce ? xr : x;
      c = reduce ? cr : 0.0f;
   }
   hfx = fmul( 0.5f, x );
   hxs = fmul( x, hfx );
   r1 = fadd( one, fmul( hxs, fadd( Q1, fmul( hxs, fadd( Q2, fmul( hxs, fadd( Q3, fmul( hxs, fadd( Q4, fmul( hxs, Q5 ) ) ) ) ) ) ) ) ) );
   t = fsub( 3.0f, fmul( r1, hfx ) );
   e = fmul( hxs, fdiv( fsub( r1, t ), fsub( 6.0f, fmul( x, t ) ) ) );
   // fdlibm finishes by cases of k (0, -1, 1, k <= -2 or k > 56, k < 23, else). The lanes of a warp fall into most of them at once, and
   // a warp executes every path some lane takes: all cases are evaluated here (each one's own arithmetic, a handful of operations,
   // independent of each other) and one is selected. Results of the cases not selected may be anything; shift counts are masked.
   const float r0 = fsub( x, fsub( fmul( x, e ), hxs ) ); // k == 0
   e = fsub( fmul( x, fsub( e, c ) ), c );
   e = fsub( e, hxs );
   const float rm1 = fsub( fmul( 0.5f, fsub( x, e ) ), 0.5f ); // k == -1
   const float rp1 = x < -0.25f ? fmul( -2.0f, fsub( e, fadd( x, 0.5f ) ) ) : fadd( one, fmul( 2.0f, fsub( x, e ) ) ); // k == 1
   const uint32_t kbits = (uint32_t)k << 23;
   const float emx = fsub( e, x );
   const float ya = fsub( one, emx ); // k <= -2 || k > 56
   const float ra = fsub( word( __float_as_uint( ya ) + kbits ), one );
   const float yb = fsub( word( 0x3f800000u - ( 0x1000000u >> ( k & 31 ) ) ), emx ); // 2 <= k < 23: t = 1 - 2^-k
   const float rb = word( __float_as_uint( yb ) + kbits );
   const float yc = fadd( fsub( x, fadd( e, word( (uint32_t)( 0x7f - k ) << 23 ) ) ), one ); // 23 <= k <= 56: t = 2^-k
   const float rc = word( __float_as_uint( yc ) + kbits );
   y = k < 23 ? rb : rc;
   y = ( k <= -2 || k > 56 ) ? ra : y;
   y = k == 1 ? rp1 : y;
   y = k == -1 ? rm1 : y;
   y = k == 0 ? r0 : y;
   t = fadd( huge, x0 );
   y = small ? fsub( x0, fsub( t, fadd( huge, x0 ) ) ) : y;
   return sat ? fsub( tiny, one ) : y;
}

LME_FN float tanhf_ref( float x )
{
   const float one = 1.0f, two = 2.0f, tiny = 1.0e-30f;
   const uint32_t jx = __float_as_uint( x ), ix = jx & 0x7fffffffu;
   // One expm1f for both ranges, and selects instead of fdlibm's early returns (the cell state of a silent stream sits beyond 22
   // while its neighbours' are small: no warp may take two paths):
   // |x| >= 1: t = expm1f(2|x|), z = 1 - 2/(t+2);  |x| < 1: t = expm1f(-2|x|), z = -t/(t+2);  |x| >= 22: 1 - tiny;  |x| < 2^-55: x (1 + x)
   const bool big = ix >= 0x3f800000u;
   const float t = expm1f_ref( fmul( big ? two : -two, fminf( fabsf( x ), 22.0f ) ) );
   const float q = fdiv( big ? two : -t, fadd( t, two ) );
   float z = big ? fsub( one, q ) : q;
   z = ix < 0x41b00000u ? z : fsub( one, tiny );
   z = ( (int)jx >= 0 ) ? z : -z;
   z = ix < 0x24000000u ? fmul( x, fadd( one, x ) ) : z;
   z = ix == 0 ? x : z;
   // infinity -> +-1 (fdlibm: one / x +- one), NaN -> NaN
   z = ix == 0x7f800000u ? ( (int)jx >= 0 ? one : -one ) : z;
   return ix > 0x7f800000u ? fadd( x, x ) : z;
}

// fdlibm log1pf (s_log1pf.c) for x >= 0 (misc.c:40-46 applies it to magnitude * 2^20); checked on the host against the C library over
// ALL 2 139 095 040 non-negative floats: identical
LME_FN float log1pf_ref( float x )
{
   const float ln2_hi = 6.9313812256e-01f, ln2_lo = 9.0580006145e-06f;
   const float Lp1 = 6.6666668653e-01f, Lp2 = 4.0000000596e-01f, Lp3 = 2.8571429849e-01f, Lp4 = 2.2222198546e-01f, Lp5 = 1.8183572590e-01f,
               Lp6 = 1.5313838422e-01f, Lp7 = 1.4798198640e-01f;
   float hfsq, f = 0.0f, c = 0.0f, s, z, R, u;
   int k = 1, hu = 0;
   const int hx = __float_as_int( x ), ax = hx & 0x7fffffff;
   if ( hx < 0x3ed413d7 ) // x < 0.41422
   {
      if ( ax < 0x31000000 ) // |x| < 2^-29
      {
         if ( ax < 0x24800000 ) return x;
         return fsub( x, fmul( fmul( x, x ), 0.5f ) );
      }
      if ( hx > 0 )
      {
         k = 0;
         f = x;
         hu = 1;
      }
   }
   if ( hx >= 0x7f800000 ) return fadd( x, x );
   if ( k != 0 )
   {
      if ( hx < 0x5a000000 )
      {
         u = fadd( 1.0f, x );
         hu = __float_as_int( u );
         k = ( hu >> 23 ) - 127;
         c = ( k > 0 ) ? fsub( 1.0f, fsub( u, x ) ) : fsub( x, fsub( u, 1.0f ) );
         // 1 + x is exact for most x in [1, 2^24): c is then +0, and (+0) / u == +0 for the u > 0 of this domain. Skipping the division
         // there is not only cheaper on the host: on the device a zero numerator sends the division down its slow path.
         if ( c != 0.0f ) c = fdiv( c, u );
      }
      else
      {
         u = x;
         hu = __float_as_int( u );
         k = ( hu >> 23 ) - 127;
         c = 0.0f;
      }
      hu &= 0x007fffff;
      if ( hu < 0x3504f7 )
         u = __int_as_float( hu | 0x3f800000 );
      else
      {
         k += 1;
         u = __int_as_float( hu | 0x3f000000 );
         hu = ( 0x00800000 - hu ) >> 2;
      }
      f = fsub( u, 1.0f );
   }
   const float kf = (float)k;
   hfsq = fmul( fmul( 0.5f, f ), f );
   if ( hu == 0 ) // |f| < 2^-20
   {
      if ( f == 0.0f )
      {
         if ( k == 0 ) return 0.0f;
         c = fadd( c, fmul( kf, ln2_lo ) );
         return fadd( fmul( kf, ln2_hi ), c );
      }
      R = fmul( hfsq, fsub( 1.0f, fmul( 0.66666666666666666f, f ) ) );
      if ( k == 0 ) return fsub( f, R );
      return fsub( fmul( kf, ln2_hi ), fsub( fsub( R, fadd( fmul( kf, ln2_lo ), c ) ), f ) );
   }
   s = fdiv( f, fadd( 2.0f, f ) );
   z = fmul( s, s );
   R = fmul( z, fadd( Lp1, fmul( z, fadd( Lp2, fmul( z, fadd( Lp3, fmul( z, fadd( Lp4, fmul( z, fadd( Lp5, fmul( z, fadd( Lp6, fmul( z, Lp7 ) ) ) ) ) ) ) ) ) ) ) ) );
   if ( k == 0 ) return fsub( f, fsub( hfsq, fmul( s, fadd( hfsq, R ) ) ) );
   return fsub( fmul( kf, ln2_hi ), fsub( fsub( hfsq, fadd( fmul( s, fadd( hfsq, R ) ), fadd( fmul( kf, ln2_lo ), c ) ) ), f ) );
}

// maths.h:327-334: 1 / (1 + expf(-x))
LME_FN float sigmoid_ref( float v, const unsigned long long *tab = EXP2F_TAB ) { return fdiv( 1.0f, fadd( 1.0f, expf_ref( -v, tab ) ) ); }
// three at once (the input, forget and output gate of a cell): the three exponentials are independent straight-line chains
LME_FN void sigmoid3_ref( float &a, float &b, float &c, const unsigned long long *tab = EXP2F_TAB )
{
   const float ea = expf_ref( -a, tab ), eb = expf_ref( -b, tab ), ec = expf_ref( -c, tab );
   const float da = fadd( 1.0f, ea ), db = fadd( 1.0f, eb ), dc = fadd( 1.0f, ec );
   a = fdiv( 1.0f, da );
   b = fdiv( 1.0f, db );
   c = fdiv( 1.0f, dc );
}
#if defined( __CUDACC__ )
// copy of the table for expf_ref( x, tab ); call with all threads of the CTA, then synchronize
LME_FN void stage_exp2f_tab( unsigned long long *dst, int tid )
{
   if ( tid < 32 ) dst[tid] = EXP2F_TAB[tid];
}
#endif
} // namespace lme

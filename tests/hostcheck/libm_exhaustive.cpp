// tests/hostcheck/libm_exhaustive.cpp -- vadc_b200/csrc/libm_exact.cuh compiled for the HOST (same source as the kernels use; the
// intrinsics become plain IEEE operations, -ffp-contract=off) and swept against the C library the pinned reference build links:
//   tanhf_ref  : every float with |x| <= 23 (beyond, both saturate to +-(1 - tiny)) + all exponents above, inf
//   expf_ref   : every float in [-105, 89]
//   log1pf_ref : every non-negative float (the STFT applies it to magnitude * 2^20)
// usage: libm_exhaustive [stride]   (stride 1 = all values; the pytest run uses a stride to stay within seconds)
// prints one line per function: values checked, mismatches, first mismatching argument. Exit code = 0 iff tanhf and log1pf match everywhere
// and expf mismatches only where the library's own FMA and non-FMA builds differ (at most 2 arguments, printed).
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "libm_exact.cuh"

static float f_of( uint32_t u ) { float f; memcpy( &f, &u, 4 ); return f; }
static uint32_t u_of( float f ) { uint32_t u; memcpy( &u, &f, 4 ); return u; }

template <typename F, typename G>
static long long sweep( const char *name, uint32_t lo, uint32_t hi, uint32_t stride, F mine, G libc, int sign_both, long long *checked, float *first_bad )
{
   long long bad = 0, n = 0;
   float first = 0.0f;
   int have = 0;
#pragma omp parallel for reduction( + : bad, n ) schedule( static )
   for ( long long i = lo; i <= (long long)hi; i += stride )
   {
      for ( int s = 0; s <= sign_both; ++s )
      {
         const float x = f_of( (uint32_t)i | ( s ? 0x80000000u : 0u ) );
         const uint32_t a = u_of( mine( x ) ), b = u_of( libc( x ) );
         ++n;
         if ( a != b && !( ( a & 0x7fffffffu ) > 0x7f800000u && ( b & 0x7fffffffu ) > 0x7f800000u ) )
         {
            ++bad;
#pragma omp critical
            if ( !have ) { have = 1; first = x; }
         }
      }
   }
   printf( "%-10s checked %lld values, %lld mismatches", name, n, bad );
   if ( bad ) printf( " (first at x = %.9g = 0x%08x: mine 0x%08x, libc 0x%08x)", first, u_of( first ), u_of( mine( first ) ), u_of( libc( first ) ) );
   printf( "\n" );
   *checked = n;
   *first_bad = first;
   return bad;
}

int main( int argc, char **argv )
{
   const uint32_t stride = argc > 1 ? (uint32_t)atoi( argv[1] ) : 1u;
   long long n;
   float fb;
   // |x| from 0 to 23.0 (0x41b80000), then one value per exponent step above, and infinity
   long long bad_t = sweep( "tanhf", 0u, 0x41b80000u, stride, []( float x ) { return lme::tanhf_ref( x ); }, []( float x ) { return tanhf( x ); }, 1, &n, &fb );
   bad_t += sweep( "tanhf>23", 0x41b80000u, 0x7f800000u, 4099u, []( float x ) { return lme::tanhf_ref( x ); }, []( float x ) { return tanhf( x ); }, 1, &n, &fb );
   // expf: positive arguments up to 89 (0x42b20000), negative down to -105 (0xc2d20000)
   long long bad_e = sweep( "expf+", 0u, 0x42b20000u, stride, []( float x ) { return lme::expf_ref( x ); }, []( float x ) { return expf( x ); }, 0, &n, &fb );
   bad_e += sweep( "expf-", 0x80000000u, 0xc2d20000u, stride, []( float x ) { return lme::expf_ref( x ); }, []( float x ) { return expf( x ); }, 0, &n, &fb );
   long long bad_l = sweep( "log1pf", 0u, 0x7f800000u, stride, []( float x ) { return lme::log1pf_ref( x ); }, []( float x ) { return log1pf( x ); }, 0, &n, &fb );
   printf( "RESULT tanhf %lld expf %lld log1pf %lld\n", bad_t, bad_e, bad_l );
   return ( bad_t == 0 && bad_l == 0 && bad_e <= 2 ) ? 0 : 1;
}

/* vadc_b200/csrc/synth.c -- deterministic synthetic 16 kHz s16le audio for tests and bench.py
 * (SURVEY.md section 8d). Not part of the reference; declared in include/vadc_segmenter.h. */
#include "vadc_segmenter.h"

#include <math.h>
#include <stdint.h>

typedef struct rng64 { uint64_t s; } rng64;

static uint64_t rng_next( rng64 *r ) /* splitmix64 */
{
   uint64_t z = ( r->s += 0x9E3779B97F4A7C15ull );
   z = ( z ^ ( z >> 30 ) ) * 0xBF58476D1CE4E5B9ull;
   z = ( z ^ ( z >> 27 ) ) * 0x94D049BB133111EBull;
   return z ^ ( z >> 31 );
}

static double rng_uniform( rng64 *r, double a, double b )
{
   return a + ( b - a ) * ( (double)( rng_next( r ) >> 11 ) * ( 1.0 / 9007199254740992.0 ) );
}

static double rng_gauss( rng64 *r )
{
   double u1 = rng_uniform( r, 1e-12, 1.0 ), u2 = rng_uniform( r, 0.0, 1.0 );
   return sqrt( -2.0 * log( u1 ) ) * cos( 6.283185307179586 * u2 );
}

static short to_s16( double v )
{
   double s = floor( v * 32768.0 + 0.5 );
   if ( s > 32767.0 ) s = 32767.0;
   if ( s < -32768.0 ) s = -32768.0;
   return (short)s;
}

void vadc_synth_pcm( unsigned long long seed, int kind, long long nsamples, short *out )
{
   rng64 r = { seed * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull };
   if ( kind == 1 )
   {
      for ( long long i = 0; i < nsamples; ++i ) out[i] = 0;
      return;
   }
   if ( kind == 2 )
   {
      for ( long long i = 0; i < nsamples; ++i ) out[i] = (short)(int)( ( rng_next( &r ) >> 48 ) - 32768 );
      return;
   }
   const double sr = 16000.0, two_pi = 6.283185307179586;
   /* noise floor first, bursts added on top in double, rounded once */
   long long pos = 0;
   long long burst_begin = -1, burst_end = -1;
   double f0 = 0, vib = 0, am_rate = 0, am_phase = 0, amp = 0, phase = 0;
   /* next burst schedule */
   long long next_start = (long long)( rng_uniform( &r, 0.3, 3.0 ) * sr );
   for ( pos = 0; pos < nsamples; ++pos )
   {
      if ( burst_begin < 0 && pos >= next_start )
      {
         burst_begin = pos;
         burst_end = pos + (long long)( rng_uniform( &r, 0.3, 2.5 ) * sr );
         f0 = rng_uniform( &r, 90.0, 220.0 );
         vib = rng_uniform( &r, 4.0, 7.0 );
         am_rate = rng_uniform( &r, 3.0, 6.0 );
         am_phase = rng_uniform( &r, 0.0, 6.0 );
         amp = rng_uniform( &r, 0.05, 0.4 );
         phase = 0.0;
      }
      double v = 0.0;
      if ( burst_begin >= 0 )
      {
         double tt = (double)( pos - burst_begin ) / sr;
         double len = (double)( burst_end - burst_begin ) / sr;
         phase += two_pi * f0 * ( 1.0 + 0.03 * sin( two_pi * vib * tt ) ) / sr;
         if ( phase > two_pi ) phase -= two_pi;
         /* harmonic stack sum_{k=1..11} sin(k*phase)/k via the Chebyshev recurrence */
         double s1 = sin( phase ), c2 = 2.0 * cos( phase );
         double sk_1 = 0.0, sk = s1, acc = s1;
         for ( int k = 2; k <= 11; ++k )
         {
            double sn = c2 * sk - sk_1;
            sk_1 = sk;
            sk = sn;
            acc += sn / k;
         }
         double am = 0.55 + 0.45 * sin( two_pi * am_rate * tt + am_phase );
         double env = tt / 0.03;
         if ( ( len - tt ) / 0.03 < env ) env = ( len - tt ) / 0.03;
         if ( env > 1.0 ) env = 1.0;
         v = amp * ( acc / 1.9 ) * am * env;
         if ( pos + 1 >= burst_end )
         {
            burst_begin = -1;
            next_start = pos + 1 + (long long)( rng_uniform( &r, 0.3, 3.0 ) * sr );
         }
      }
      v += 0.003 * rng_gauss( &r );
      out[pos] = to_s16( v );
   }
}

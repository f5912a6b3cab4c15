// tests/hostcheck/faithful_host_race.cpp -- TEST INFRASTRUCTURE ONLY.
//
// Race check of the faithful encoder's orchestration without a GPU: vadc_b200/csrc/faithful_kernel.cuh compiled for the host with
// fq::barrier() mapped to a pthread barrier, fq::encoder_chunk run by NT host threads (tid 0..NT-1 of NT, exactly the kernel's
// "for ( e = tid; e < n; e += nt )" stage loops) under ThreadSanitizer (-fsanitize=thread). A stage that reads what another thread
// writes in the same stage, or a buffer reused one barrier too early, is a data race TSan reports; and the parallel result must
// equal the serial one bit for bit. usage: faithful_host_race <weights.testtensor> [threads]   (exit 0 = no race, same bits)
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define FQ_HOST_BARRIER 1
#include "../../vadc_b200/csrc/faithful_kernel.cuh"
#include "../../oracle/silero_oracle.h"

enum { MAX_NT = 16, CHUNKS = 3 };
static int NT = 8; // host threads standing in for the CTA's threads (argv[2])
static pthread_barrier_t g_bar;
static bool g_parallel = false;
extern "C" void fq_host_barrier( void )
{
   if ( g_parallel ) pthread_barrier_wait( &g_bar );
}

struct Job
{
   const fq::Weights *W;
   const float *logspec;
   float *a4, *sm;
   int tid;
};

static void *worker( void *p )
{
   Job *j = (Job *)p;
   for ( int c = 0; c < CHUNKS; ++c )
   {
      fq::encoder_chunk( *j->W, j->logspec + (size_t)c * 3225, j->a4 + (size_t)c * 448, j->sm, j->tid, NT );
      pthread_barrier_wait( &g_bar ); // the kernel's chunk loop: the next chunk reuses the shared buffers (encoder_chunk ends on a barrier too)
   }
   return 0;
}

int main( int argc, char **argv )
{
   if ( argc < 2 ) return 2;
   if ( argc > 2 ) NT = atoi( argv[2] );
   if ( NT < 2 || NT > MAX_NT ) return 2;
   so_model *m = so_model_load_file( argv[1] );
   if ( !m ) return 3;
   fq::Weights W;
   for ( int i = 0; i < 99; ++i ) W.t[i] = so_model_tensor( m, i, 0, 0 );
   for ( int k = 0; k < fq::N_TRANSPOSED; ++k )
   {
      int idx, n_out, n_in;
      fq::transposed_slot( k, &idx, &n_out, &n_in );
      float *t = (float *)malloc( sizeof( float ) * (size_t)n_out * n_in );
      for ( int r = 0; r < n_out; ++r )
         for ( int c = 0; c < n_in; ++c ) t[(size_t)c * n_out + r] = W.t[idx][(size_t)r * n_in + c];
      W.tt[k] = t;
   }
   // log spectrogram-like input: log1p of positive pseudo-random magnitudes over six decades
   float *logspec = (float *)malloc( sizeof( float ) * CHUNKS * 3225 );
   unsigned long long lcg = 12345;
   for ( int i = 0; i < CHUNKS * 3225; ++i )
   {
      lcg = lcg * 6364136223846793005ull + 1442695040888963407ull;
      const float u = (float)( ( lcg >> 40 ) & 0xffffff ) / 16777216.0f;
      logspec[i] = log1pf( expf( 14.0f * u - 7.0f ) * 1048576.0f * 1e-3f );
   }
   float *serial = (float *)calloc( CHUNKS * 448, sizeof( float ) ), *par = (float *)calloc( CHUNKS * 448, sizeof( float ) );
   float *sm = (float *)malloc( sizeof( float ) * fq::SM_FLOATS );
   for ( int c = 0; c < CHUNKS; ++c ) fq::encoder_chunk( W, logspec + (size_t)c * 3225, serial + (size_t)c * 448, sm, 0, 1 );

   pthread_barrier_init( &g_bar, 0, NT );
   g_parallel = true;
   pthread_t th[MAX_NT];
   Job jobs[MAX_NT];
   for ( int t = 0; t < NT; ++t )
   {
      jobs[t] = Job{ &W, logspec, par, sm, t };
      pthread_create( &th[t], 0, worker, &jobs[t] );
   }
   for ( int t = 0; t < NT; ++t ) pthread_join( th[t], 0 );
   g_parallel = false;
   const int same = memcmp( serial, par, sizeof( float ) * CHUNKS * 448 ) == 0;
   double sum = 0;
   for ( int i = 0; i < CHUNKS * 448; ++i ) sum += serial[i];
   printf( "threads %d, chunks %d, checksum %.6f, parallel == serial: %s\n", NT, CHUNKS, sum, same ? "yes" : "NO" );
   return same ? 0 : 1;
}

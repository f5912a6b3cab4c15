/* vadc_b200/csrc/group.c -- one host process, several GPUs: the multi-stream chunk scheduler across the devices of a box.
 *
 * The reference is a single-threaded, single-stream program (vadc.c); its chunk -> probability -> segment loop shards naturally by
 * STREAM: streams share nothing but the read-only weights. A group owns one engine (silero_b200 handle) per
 * device and one host thread per engine; global stream s lives on device s / streams_per_device for its whole life, with its LSTM
 * and segmenter state resident there. A group call fans the caller's stream range out to the devices that own a part of it, every
 * device works on its slice of the caller's HOST buffers in place (pinned or not), and the call returns when all are done: the
 * "final gather of per-stream segments" is the concatenation the caller's own arrays already are. There is no collective and no
 * device-to-device traffic on this path.
 *
 * Plain C over the public ABI of silero_b200.h (nothing here touches CUDA directly), POSIX threads.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "silero_b200.h"

typedef struct group_job
{
   int kind; /* 1 run_streams_segments, 2 reset, 3 segments_configure, 4 segments_reset, 0 quit */
   const int16_t *pcm;
   long long stream_stride;
   int first_stream, nstreams, nchunks, end_of_stream, cap;
   vadc_segment *segs;
   int *counts;
   float *probs;
   const vadc_seg_params *params;
} group_job;

typedef struct group_worker
{
   struct silero_b200_group *g;
   int index;
   silero_b200 *h;
   pthread_t thread;
   pthread_mutex_t mu;
   pthread_cond_t cv;
   int has_job, done, rc;
   group_job job;
   char err[256];
} group_worker;

struct silero_b200_group
{
   int ndev, per_dev, max_streams;
   int devices[SILERO_B200_GROUP_MAX_DEVICES];
   group_worker w[SILERO_B200_GROUP_MAX_DEVICES];
};

static _Thread_local char g_group_err[320] = "";
const char *silero_b200_group_last_error( void ) { return g_group_err; }

static void *worker_main( void *arg )
{
   group_worker *w = (group_worker *)arg;
   for ( ;; )
   {
      pthread_mutex_lock( &w->mu );
      while ( !w->has_job ) pthread_cond_wait( &w->cv, &w->mu );
      group_job j = w->job;
      pthread_mutex_unlock( &w->mu );
      int rc = 0;
      if ( j.kind == 0 ) break;
      if ( j.kind == 1 )
         rc = silero_b200_run_streams_segments( w->h, j.pcm, j.stream_stride, j.first_stream, j.nstreams, j.nchunks, j.end_of_stream, j.segs, j.cap, j.counts, j.probs );
      else if ( j.kind == 2 )
         rc = silero_b200_reset( w->h, j.first_stream, j.nstreams );
      else if ( j.kind == 3 )
         rc = silero_b200_segments_configure( w->h, j.params );
      else if ( j.kind == 4 )
         rc = silero_b200_segments_reset( w->h, j.first_stream, j.nstreams );
      if ( rc ) snprintf( w->err, sizeof( w->err ), "device %d: %s", w->g->devices[w->index], silero_b200_last_error() );
      pthread_mutex_lock( &w->mu );
      w->has_job = 0;
      w->rc = rc;
      w->done = 1;
      pthread_cond_broadcast( &w->cv );
      pthread_mutex_unlock( &w->mu );
   }
   return 0;
}

static void post( group_worker *w, const group_job *j )
{
   pthread_mutex_lock( &w->mu );
   w->job = *j;
   w->done = 0;
   w->has_job = 1;
   pthread_cond_broadcast( &w->cv );
   pthread_mutex_unlock( &w->mu );
}

static int collect( group_worker *w )
{
   pthread_mutex_lock( &w->mu );
   while ( !w->done ) pthread_cond_wait( &w->cv, &w->mu );
   const int rc = w->rc;
   pthread_mutex_unlock( &w->mu );
   if ( rc ) snprintf( g_group_err, sizeof( g_group_err ), "%s", w->err );
   return rc;
}

void silero_b200_group_destroy( silero_b200_group *g )
{
   if ( !g ) return;
   for ( int d = 0; d < g->ndev; ++d )
   {
      group_worker *w = &g->w[d];
      if ( w->g )
      {
         group_job j;
         memset( &j, 0, sizeof( j ) );
         post( w, &j );
         pthread_join( w->thread, 0 );
         pthread_mutex_destroy( &w->mu );
         pthread_cond_destroy( &w->cv );
      }
      if ( w->h ) silero_b200_destroy( w->h );
   }
   free( g );
}

int silero_b200_group_create( const void *testtensor_bytes, size_t nbytes, const int *devices, int ndevices, const silero_b200_opts *opts, silero_b200_group **out )
{
   if ( !testtensor_bytes || !devices || !out || ndevices < 1 || ndevices > SILERO_B200_GROUP_MAX_DEVICES )
   {
      snprintf( g_group_err, sizeof( g_group_err ), "bad argument (1..%d devices)", SILERO_B200_GROUP_MAX_DEVICES );
      return SILERO_B200_ERR_ARG;
   }
   *out = 0;
   silero_b200_group *g = (silero_b200_group *)calloc( 1, sizeof( *g ) );
   if ( !g )
   {
      snprintf( g_group_err, sizeof( g_group_err ), "out of host memory" );
      return SILERO_B200_ERR_NOMEM;
   }
   silero_b200_opts o;
   if ( opts )
      o = *opts;
   else
      silero_b200_default_opts( &o );
   if ( o.max_streams < 1 ) o.max_streams = 1;
   g->ndev = ndevices;
   g->max_streams = o.max_streams;
   g->per_dev = ( o.max_streams + ndevices - 1 ) / ndevices;
   for ( int d = 0; d < ndevices; ++d )
   {
      g->devices[d] = devices[d];
      silero_b200_opts od = o;
      od.device = devices[d];
      od.max_streams = g->per_dev;
      const int rc = silero_b200_create( testtensor_bytes, nbytes, &od, &g->w[d].h );
      if ( rc )
      {
         snprintf( g_group_err, sizeof( g_group_err ), "device %d: %s", devices[d], silero_b200_last_error() );
         silero_b200_group_destroy( g );
         return rc;
      }
   }
   for ( int d = 0; d < ndevices; ++d )
   {
      group_worker *w = &g->w[d];
      w->index = d;
      pthread_mutex_init( &w->mu, 0 );
      pthread_cond_init( &w->cv, 0 );
      w->g = g;
      if ( pthread_create( &w->thread, 0, worker_main, w ) )
      {
         w->g = 0;
         pthread_mutex_destroy( &w->mu );
         pthread_cond_destroy( &w->cv );
         snprintf( g_group_err, sizeof( g_group_err ), "cannot start the host thread of device %d", devices[d] );
         silero_b200_group_destroy( g );
         return SILERO_B200_ERR_NOMEM;
      }
   }
   *out = g;
   return SILERO_B200_OK;
}

int silero_b200_group_create_from_file( const char *path, const int *devices, int ndevices, const silero_b200_opts *opts, silero_b200_group **out )
{
   if ( !path || !out )
   {
      snprintf( g_group_err, sizeof( g_group_err ), "null argument" );
      return SILERO_B200_ERR_ARG;
   }
   FILE *f = fopen( path, "rb" );
   if ( !f )
   {
      snprintf( g_group_err, sizeof( g_group_err ), "cannot open %s", path );
      return SILERO_B200_ERR_WEIGHTS;
   }
   fseek( f, 0, SEEK_END );
   const long n = ftell( f );
   fseek( f, 0, SEEK_SET );
   void *buf = n > 0 ? malloc( (size_t)n ) : 0;
   if ( !buf || fread( buf, 1, (size_t)n, f ) != (size_t)n )
   {
      fclose( f );
      free( buf );
      snprintf( g_group_err, sizeof( g_group_err ), "cannot read %s", path );
      return SILERO_B200_ERR_WEIGHTS;
   }
   fclose( f );
   const int rc = silero_b200_group_create( buf, (size_t)n, devices, ndevices, opts, out );
   free( buf );
   return rc;
}

int silero_b200_group_get_info( const silero_b200_group *g, int *ndevices, int *streams_per_device, int *max_streams )
{
   if ( !g ) return SILERO_B200_ERR_ARG;
   if ( ndevices ) *ndevices = g->ndev;
   if ( streams_per_device ) *streams_per_device = g->per_dev;
   if ( max_streams ) *max_streams = g->max_streams;
   return SILERO_B200_OK;
}

/* fan a stream range out: fill(job, device, local_first, count, offset of the device's first stream inside the caller's range) */
static int fan_out( silero_b200_group *g, int first_stream, int nstreams, const group_job *proto )
{
   if ( !g || first_stream < 0 || nstreams < 0 || first_stream + nstreams > g->max_streams )
   {
      snprintf( g_group_err, sizeof( g_group_err ), "streams [%d,%d) outside [0,%d)", first_stream, first_stream + nstreams, g ? g->max_streams : 0 );
      return SILERO_B200_ERR_ARG;
   }
   int posted[SILERO_B200_GROUP_MAX_DEVICES] = { 0 };
   for ( int d = 0; d < g->ndev; ++d )
   {
      const int lo = d * g->per_dev, hi = lo + g->per_dev;
      const int a = first_stream > lo ? first_stream : lo, b = first_stream + nstreams < hi ? first_stream + nstreams : hi;
      if ( b <= a ) continue;
      group_job j = *proto;
      const long long off = a - first_stream; /* streams of the caller's range before this device's part */
      j.first_stream = a - lo;
      j.nstreams = b - a;
      if ( j.pcm ) j.pcm += off * j.stream_stride;
      if ( j.segs ) j.segs += off * j.cap;
      if ( j.counts ) j.counts += off;
      if ( j.probs ) j.probs += off * j.nchunks;
      post( &g->w[d], &j );
      posted[d] = 1;
   }
   int rc = 0;
   for ( int d = 0; d < g->ndev; ++d )
      if ( posted[d] )
      {
         const int r = collect( &g->w[d] );
         if ( r && !rc ) rc = r;
      }
   return rc;
}

int silero_b200_group_run_streams_segments( silero_b200_group *g, const int16_t *pcm, long long stream_stride, int first_stream, int nstreams, int nchunks,
                                            int end_of_stream, vadc_segment *segs, int cap, int *counts, float *probs )
{
   group_job j;
   memset( &j, 0, sizeof( j ) );
   j.kind = 1;
   j.pcm = pcm;
   j.stream_stride = stream_stride;
   j.nchunks = nchunks;
   j.end_of_stream = end_of_stream;
   j.segs = segs;
   j.cap = cap;
   j.counts = counts;
   j.probs = probs;
   return fan_out( g, first_stream, nstreams, &j );
}

int silero_b200_group_reset( silero_b200_group *g, int first_stream, int nstreams )
{
   group_job j;
   memset( &j, 0, sizeof( j ) );
   j.kind = 2;
   int rc = fan_out( g, first_stream, nstreams, &j );
   if ( rc ) return rc;
   j.kind = 4;
   return fan_out( g, first_stream, nstreams, &j );
}

int silero_b200_group_segments_configure( silero_b200_group *g, const vadc_seg_params *params )
{
   if ( !g ) return SILERO_B200_ERR_ARG;
   group_job j;
   memset( &j, 0, sizeof( j ) );
   j.kind = 3;
   j.params = params;
   return fan_out( g, 0, g->max_streams, &j );
}

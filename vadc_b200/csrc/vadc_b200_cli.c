/* vadc_b200/csrc/vadc_b200_cli.c -- native Linux command line over libsilero_b200.so.
 *
 * Keeps the reference CLI's contract (README.md:62-84, vadc.c:1084-1276): s16le 16 kHz mono PCM on
 * stdin, one "start,end" line per speech segment on stdout (flushed as soon as it is final), all
 * diagnostics on stderr; same option names, defaults and "a value <= 0 is ignored" rule
 * (vadc.c:1110-1124, 1215); --raw_probabilities prints "%f\n" per chunk (vadc.c:992-997);
 * --output_centi_seconds (vadc.c:251-256); --stats line on stderr (vadc.c:1069-1075).
 * What replaces the Win32 shell (vadc.c:401-667): plain read(2) on stdin instead of ReadFile; a named input
 * is decoded by an ffmpeg child exactly as the reference does (init_buffered_stream_ffmpeg, vadc.c:531-626:
 * same command line, -ss from --start_seconds, -map 0:a:N from --audio_source, ffmpeg's stdin closed, its
 * stdout on a pipe) but with fork/execvp instead of CreatePipe/CreateProcessW. Files named *.s16le, *.raw or
 * *.pcm are taken as already-decoded PCM and read directly (no child). $VADC_FFMPEG overrides the program name.
 *
 * New: any number of raw s16le FILES may be given; each file is an independent stream, all of them
 * are pushed through the multi-stream scheduler at once (per-stream LSTM and segmenter state on the
 * GPU, segments produced by the on-device segmenter) and the per-file results are printed in
 * argument order ("# path" header lines when there is more than one file). --devices 0,1,2,... spreads
 * the files over several GPUs of the box (silero_b200_group_*: one host thread per device, streams
 * sharded in contiguous blocks, results gathered in argument order); --device N picks one.
 */
#define _POSIX_C_SOURCE 200809L
#include <errno.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#include <pthread.h>
#include <sys/stat.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#include "silero_b200.h"

typedef struct cli_opts
{
   vadc_seg_params seg;
   int batch;
   int batch_given; /* --batch on the command line */
   float start_seconds;
   int raw_probabilities, stats, device, audio_source;
   int devices[SILERO_B200_GROUP_MAX_DEVICES], ndevices;
   const char *model;
   const char **files;
   int nfiles;
} cli_opts;

static double now_s( void )
{
   struct timespec ts;
   clock_gettime( CLOCK_MONOTONIC, &ts );
   return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

static int parse_args( int argc, char **argv, cli_opts *o )
{
   vadc_seg_params_default( &o->seg );
   o->batch = 96; /* vadc.c:1116 */
   o->files = (const char **)calloc( (size_t)argc + 1, sizeof( char * ) );
   if ( !o->files ) return -1;
   for ( int i = 1; i < argc; ++i )
   {
      const char *a = argv[i];
      float *fdst = 0;
      if ( !strcmp( a, "--raw_probabilities" ) ) o->raw_probabilities = 1;
      else if ( !strcmp( a, "--stats" ) ) o->stats = 1;
      else if ( !strcmp( a, "--output_centi_seconds" ) ) o->seg.centiseconds = 1;
      else if ( !strcmp( a, "--model" ) ) { if ( i + 1 < argc ) o->model = argv[++i]; }
      else if ( !strcmp( a, "--devices" ) )
      {
         /* comma-separated CUDA ordinals: the multi-file mode shards its streams over them */
         if ( i + 1 < argc )
            for ( const char *p = argv[++i]; *p && o->ndevices < SILERO_B200_GROUP_MAX_DEVICES; )
            {
               char *end;
               long v = strtol( p, &end, 10 );
               if ( end == p ) break;
               if ( v >= 0 ) o->devices[o->ndevices++] = (int)v;
               p = *end == ',' ? end + 1 : end;
            }
      }
      else if ( !strcmp( a, "--min_silence" ) ) fdst = &o->seg.min_silence_ms;
      else if ( !strcmp( a, "--min_speech" ) ) fdst = &o->seg.min_speech_ms;
      else if ( !strcmp( a, "--threshold" ) ) fdst = &o->seg.threshold;
      else if ( !strcmp( a, "--neg_threshold_relative" ) ) fdst = &o->seg.neg_threshold_relative;
      else if ( !strcmp( a, "--speech_pad" ) ) fdst = &o->seg.speech_pad_ms;
      else if ( !strcmp( a, "--start_seconds" ) ) fdst = &o->start_seconds;
      else if ( !strcmp( a, "--batch" ) || !strcmp( a, "--sequence_count" ) || !strcmp( a, "--audio_source" ) || !strcmp( a, "--device" ) )
      {
         /* numeric options that are not floats of the segmenter: --sequence_count is clamped to 1536 by the C backend
            (silero.h:41-42, vadc.c:742-754); --audio_source selects the ffmpeg audio stream (-map 0:a:N, vadc.c:537) */
         if ( i + 1 < argc )
         {
            float v = (float)atof( argv[++i] );
            if ( !strcmp( a, "--batch" ) && v > 0.0f )
            {
               o->batch = (int)v;
               o->batch_given = 1;
            }
            if ( !strcmp( a, "--device" ) && v >= 0.0f ) o->device = (int)v;
            if ( !strcmp( a, "--audio_source" ) && v > 0.0f ) o->audio_source = (int)v; /* vadc.c:1122, 1215 */
         }
      }
      else
         o->files[o->nfiles++] = a; /* vadc.c:1226-1229: anything else is the input */
      if ( fdst && i + 1 < argc )
      {
         float v = (float)atof( argv[++i] );
         if ( v > 0.0f ) *fdst = v; /* vadc.c:1215 */
      }
   }
   if ( o->batch < 1 ) o->batch = 1;
   return 0;
}

static void print_segment( const vadc_segmenter *s, vadc_segment seg, double *total_speech )
{
   char line[96];
   float b, e;
   vadc_segment_format( s, seg, line, sizeof( line ) );
   fputs( line, stdout );
   fflush( stdout ); /* vadc.c:258 */
   vadc_segment_times( s, seg, &b, &e );
   *total_speech += (double)e - (double)b;
}

static void print_stats( double total_speech, long long total_samples, double t0 )
{
   /* vadc.c:1037-1075 */
   double total_duration = (double)total_samples / SILERO_B200_SAMPLE_RATE;
   double elapsed = now_s() - t0;
   int hours = (int)( total_duration / 3600.0 );
   int minutes = (int)( ( total_duration - hours * 3600.0 ) / 60.0 );
   int seconds = (int)( total_duration - hours * 3600.0 - minutes * 60.0 );
   int milliseconds = (int)( ( total_duration - hours * 3600.0 - minutes * 60.0 - seconds ) * 1000.0 );
   fprintf( stderr, "time=%02d:%02d:%02d.%04d", hours, minutes, seconds, milliseconds );
   fprintf( stderr, " %7.2f speech (%5.1f%%), %5.1f / %5.1f (%5.1fx)\r", total_speech, total_duration > 0 ? total_speech / total_duration * 100.0 : 0.0,
            total_duration, elapsed, elapsed > 0 ? total_duration / elapsed : 0.0 );
}

/* read exactly n bytes unless EOF comes first; returns bytes read, -1 on error */
static long long read_full( int fd, void *buf, size_t n )
{
   size_t got = 0;
   while ( got < n )
   {
      ssize_t r = read( fd, (char *)buf + got, n - got );
      if ( r < 0 )
      {
         if ( errno == EINTR ) continue;
         return -1;
      }
      if ( r == 0 ) break;
      got += (size_t)r;
   }
   return (long long)got;
}

/* ---- ingest: raw PCM file or ffmpeg child (vadc.c:531-626) ---- */
static int is_raw_pcm_name( const char *path )
{
   const char *dot = strrchr( path, '.' );
   return dot && ( !strcmp( dot, ".s16le" ) || !strcmp( dot, ".raw" ) || !strcmp( dot, ".pcm" ) );
}

/* Launches `ffmpeg -hide_banner -loglevel error -nostats -ss <start> -i <path> -map 0:a:<n> -vn -sn -dn -ac 1 -ar 16k -f s16le -`
   (the reference's command line, vadc.c:537) with stdin closed (vadc.c:560) and stdout on a pipe; returns the read end, or -1. */
static int spawn_ffmpeg( const char *path, float start_seconds, int audio_source, pid_t *pid_out )
{
   int fds[2];
   if ( pipe( fds ) )
   {
      fprintf( stderr, "Error creating ffmpeg pipe\n" ); /* vadc.c:552 */
      return -1;
   }
   char ss[64], map[64];
   snprintf( ss, sizeof( ss ), "%f", (double)start_seconds );
   snprintf( map, sizeof( map ), "0:a:%d", audio_source );
   const char *prog = getenv( "VADC_FFMPEG" );
   if ( !prog || !*prog ) prog = "ffmpeg";
   pid_t pid = fork();
   if ( pid < 0 )
   {
      fprintf( stderr, "Error launching ffmpeg\n" ); /* vadc.c:570 */
      close( fds[0] );
      close( fds[1] );
      return -1;
   }
   if ( pid == 0 )
   {
      close( fds[0] );
      dup2( fds[1], 1 );
      close( fds[1] );
      close( 0 );
      char *const argv[] = { (char *)prog, (char *)"-hide_banner", (char *)"-loglevel", (char *)"error", (char *)"-nostats", (char *)"-ss", ss,
                             (char *)"-i", (char *)path, (char *)"-map", map, (char *)"-vn", (char *)"-sn", (char *)"-dn", (char *)"-ac", (char *)"1",
                             (char *)"-ar", (char *)"16k", (char *)"-f", (char *)"s16le", (char *)"-", 0 };
      execvp( prog, argv );
      fprintf( stderr, "Error launching ffmpeg\n" );
      _exit( 127 );
   }
   close( fds[1] );
   *pid_out = pid;
   return fds[0];
}

/* whole decoded stream into memory (multi-file mode needs every stream's length up front) */
static int16_t *slurp_fd( int fd, long long *samples_out )
{
   size_t cap = 1u << 22, got = 0;
   char *buf = (char *)malloc( cap );
   while ( buf )
   {
      if ( got == cap )
      {
         char *nb = (char *)realloc( buf, cap * 2 );
         if ( !nb )
         {
            free( buf );
            return 0;
         }
         buf = nb;
         cap *= 2;
      }
      ssize_t r = read( fd, buf + got, cap - got );
      if ( r < 0 && errno == EINTR ) continue;
      if ( r <= 0 ) break;
      got += (size_t)r;
   }
   *samples_out = (long long)( got / 2 );
   return (int16_t *)buf;
}

#define CLI_FILE_BATCH 1536 /* chunks per call for regular-file input without --batch (147 s of audio, 4.7 MB) */

/* The next batch is read while the GPU works on this one: a reader thread fills the other of two buffers (a 10-hour recording is
   1.15 GB of reads, a sixth of its wall time when they alternate with the inference calls). Order, batch boundaries and therefore
   the output are those of the sequential loop; on a live pipe a batch is still processed as soon as it is complete. */
typedef struct prefetch
{
   int fd;
   size_t cap_bytes;
   int16_t *buf[2];
   long long bytes[2]; /* what read_full returned for the buffer */
   int filled[2];
   int err;            /* errno of a failed read (errno itself is per thread) */
   pthread_mutex_t m;
   pthread_cond_t cv;
} prefetch;

static void *prefetch_main( void *arg )
{
   prefetch *p = (prefetch *)arg;
   for ( int i = 0;; i ^= 1 )
   {
      pthread_mutex_lock( &p->m );
      while ( p->filled[i] ) pthread_cond_wait( &p->cv, &p->m );
      pthread_mutex_unlock( &p->m );
      const long long r = read_full( p->fd, p->buf[i], p->cap_bytes );
      if ( r < 0 ) p->err = errno;
      pthread_mutex_lock( &p->m );
      p->bytes[i] = r;
      p->filled[i] = 1;
      pthread_cond_broadcast( &p->cv );
      pthread_mutex_unlock( &p->m );
      if ( r < (long long)p->cap_bytes ) return 0; /* error or end of stream: the consumer stops after this buffer */
   }
}

/* ---- one stream from a descriptor (stdin, or the ffmpeg pipe): the reference's own loop (vadc.c:852-1027) ---- */
static int run_fd( silero_b200 *h, const cli_opts *o, int in_fd, int skip_on_read )
{
   /* Chunks per inference call: --batch (default 96, vadc.c:1116). The batch does not change a single output bit, only how often the
      host and the GPU take turns; when the input is a regular FILE (nothing to wait for, no listener to keep up with) and the user
      has not chosen, calls of CLI_FILE_BATCH chunks are used -- 10 % off a 10-hour recording's wall time. Pipes keep the reference's batch. */
   int batch = o->batch;
   struct stat st;
   if ( !o->batch_given && fstat( in_fd, &st ) == 0 && S_ISREG( st.st_mode ) && batch < CLI_FILE_BATCH ) batch = CLI_FILE_BATCH;
   const size_t cap_samples = (size_t)batch * SILERO_B200_CHUNK_SAMPLES;
   int16_t *pcm = (int16_t *)malloc( 2 * cap_samples * sizeof( int16_t ) );
   float *probs = (float *)malloc( (size_t)batch * sizeof( float ) );
   vadc_segment segs[64];
   if ( !pcm || !probs ) return 1;
   vadc_segmenter seg;
   vadc_segmenter_init( &seg, &o->seg );
   double total_speech = 0.0, t0 = now_s();
   long long total_samples = 0;
   /* --start_seconds: the reference hands it to ffmpeg only (-ss, vadc.c:537) and ignores it for stdin (init_buffered_stream_stdin,
      vadc.c:610-627, 818): so does this program. skip_on_read is for callers that decode without a seeking child. */
   long long skip = skip_on_read ? (long long)( (double)o->start_seconds * SILERO_B200_SAMPLE_RATE ) * 2 : 0;
   while ( skip > 0 )
   {
      size_t n = skip < (long long)( cap_samples * 2 ) ? (size_t)skip : cap_samples * 2;
      long long r = read_full( in_fd, pcm, n );
      if ( r <= 0 ) break;
      skip -= r;
   }
   prefetch pf;
   memset( &pf, 0, sizeof pf );
   pf.fd = in_fd;
   pf.cap_bytes = cap_samples * sizeof( int16_t );
   pf.buf[0] = pcm;
   pf.buf[1] = pcm + cap_samples;
   pthread_mutex_init( &pf.m, 0 );
   pthread_cond_init( &pf.cv, 0 );
   pthread_t reader;
   if ( pthread_create( &reader, 0, prefetch_main, &pf ) )
   {
      fprintf( stderr, "Error: cannot start the reader thread\n" );
      free( pcm );
      free( probs );
      return 1;
   }
   int rc = 0;
   for ( int cur = 0;; cur ^= 1 )
   {
      pthread_mutex_lock( &pf.m );
      while ( !pf.filled[cur] ) pthread_cond_wait( &pf.cv, &pf.m );
      pthread_mutex_unlock( &pf.m );
      const int16_t *batch = pf.buf[cur];
      long long bytes = pf.bytes[cur];
      if ( bytes < 0 )
      {
         fprintf( stderr, "Error: read failed: %s\n", strerror( pf.err ) );
         break;
      }
      const long long values_read = bytes / 2;
      const int nchunks = (int)( values_read / SILERO_B200_CHUNK_SAMPLES ); /* vadc.c:964: the trailing partial chunk is dropped */
      if ( nchunks > 0 )
      {
         if ( silero_b200_run_streams( h, batch, (long long)cap_samples, 0, 1, nchunks, probs, 0 ) )
         {
            fprintf( stderr, "Error: %s\n", silero_b200_last_error() );
            rc = 1;
            break; /* (the reader is joined below; it ends at the end of its input) */
         }
         total_samples += (long long)nchunks * SILERO_B200_CHUNK_SAMPLES;
         if ( o->raw_probabilities )
            for ( int i = 0; i < nchunks; ++i ) printf( "%f\n", probs[i] ); /* vadc.c:995 */
         else
            for ( int i = 0; i < nchunks; i += 64 )
            {
               int n = nchunks - i < 64 ? nchunks - i : 64;
               long long k = vadc_segmenter_feed( &seg, probs + i, n, segs, 64 );
               for ( long long j = 0; j < k; ++j ) print_segment( &seg, segs[j], &total_speech );
            }
         if ( o->stats ) print_stats( total_speech, total_samples, t0 );
      }
      if ( (size_t)bytes < cap_samples * sizeof( int16_t ) ) break; /* end of stream */
      pthread_mutex_lock( &pf.m );
      pf.filled[cur] = 0; /* the reader may overwrite it */
      pthread_cond_broadcast( &pf.cv );
      pthread_mutex_unlock( &pf.m );
   }
   if ( rc )
   {
      pthread_cancel( reader ); /* blocked in read() or on the condition: both are cancellation points */
      pthread_join( reader, 0 );
      free( pcm );
      free( probs );
      return rc;
   }
   pthread_join( reader, 0 ); /* it has returned: the loop above ends with the reader's last buffer */
   if ( !o->raw_probabilities )
   {
      long long k = vadc_segmenter_finish( &seg, segs, 64 );
      for ( long long j = 0; j < k; ++j ) print_segment( &seg, segs[j], &total_speech );
   }
   else
      fflush( stdout );
   if ( o->stats )
   {
      print_stats( total_speech, total_samples, t0 );
      fputc( '\n', stderr );
   }
   free( pcm );
   free( probs );
   return 0;
}

/* ---- many files = many concurrent streams ---- */
typedef struct file_stream
{
   const char *path;
   long long nchunks;
   int order;        /* position on the command line */
   int16_t *decoded; /* ffmpeg inputs: the whole decoded stream; raw PCM files are read straight into the pinned buffer */
} file_stream;

static int by_length_desc( const void *a, const void *b )
{
   const file_stream *x = (const file_stream *)a, *y = (const file_stream *)b;
   if ( x->nchunks != y->nchunks ) return x->nchunks > y->nchunks ? -1 : 1;
   return x->order - y->order;
}

static int run_files( silero_b200_group *h, const cli_opts *o )
{
   const int S = o->nfiles;
   int rc = 1; /* every exit goes through `done`: buffers are released on errors too */
   int16_t *pcm = 0;
   vadc_segment *segs = 0, *tmp = 0;
   int *nseg = 0, *cnt = 0;
   float *probs = 0, *ptmp = 0;
   FILE *f = 0;
   file_stream *fs = (file_stream *)calloc( (size_t)S, sizeof( file_stream ) );
   if ( !fs ) return 1;
   const long long skip_samples = (long long)( (double)o->start_seconds * SILERO_B200_SAMPLE_RATE );
   long long maxchunks = 0;
   for ( int i = 0; i < S; ++i )
   {
      long long samples = 0;
      if ( is_raw_pcm_name( o->files[i] ) )
      {
         f = fopen( o->files[i], "rb" );
         if ( !f )
         {
            fprintf( stderr, "Error: cannot open %s\n", o->files[i] );
            goto done;
         }
         fseek( f, 0, SEEK_END );
         samples = ftell( f ) / 2 - skip_samples;
         fclose( f );
         f = 0;
      }
      else
      {
         /* decoded by an ffmpeg child, which also does the --start_seconds seek (-ss) */
         pid_t pid = 0;
         int fd = spawn_ffmpeg( o->files[i], o->start_seconds, o->audio_source, &pid );
         if ( fd < 0 ) goto done;
         fs[i].decoded = slurp_fd( fd, &samples );
         close( fd );
         waitpid( pid, 0, 0 );
         if ( !fs[i].decoded ) goto done;
      }
      fs[i].path = o->files[i];
      fs[i].order = i;
      fs[i].nchunks = samples > 0 ? samples / SILERO_B200_CHUNK_SAMPLES : 0;
      if ( fs[i].nchunks > maxchunks ) maxchunks = fs[i].nchunks;
   }
   /* longest first: at any time the streams that still have audio are a prefix of the stream numbering */
   qsort( fs, (size_t)S, sizeof( file_stream ), by_length_desc );
   const long long stride = ( maxchunks > 0 ? maxchunks : 1 ) * SILERO_B200_CHUNK_SAMPLES;
   if ( silero_b200_host_alloc_pinned( (size_t)S * (size_t)stride * sizeof( int16_t ), (void **)&pcm ) )
   {
      fprintf( stderr, "Error: %s\n", silero_b200_last_error() );
      pcm = 0;
      goto done;
   }
   for ( int s = 0; s < S; ++s )
   {
      size_t want = (size_t)fs[s].nchunks * SILERO_B200_CHUNK_SAMPLES;
      if ( fs[s].decoded )
      {
         memcpy( pcm + (size_t)s * stride, fs[s].decoded, want * 2 );
         free( fs[s].decoded );
         fs[s].decoded = 0;
         continue;
      }
      f = fopen( fs[s].path, "rb" );
      if ( !f ) goto done;
      fseek( f, skip_samples * 2, SEEK_SET );
      if ( fread( pcm + (size_t)s * stride, 2, want, f ) != want )
      {
         fprintf( stderr, "Error: short read on %s\n", fs[s].path );
         goto done;
      }
      fclose( f );
      f = 0;
   }
   /* results per stream */
   const int cap = (int)( maxchunks / 2 + 2 );
   segs = (vadc_segment *)malloc( (size_t)S * cap * sizeof( vadc_segment ) );
   nseg = (int *)calloc( (size_t)S, sizeof( int ) );
   probs = o->raw_probabilities ? (float *)malloc( (size_t)S * ( maxchunks > 0 ? maxchunks : 1 ) * sizeof( float ) ) : 0;
   const int B = o->batch * 16; /* chunks per stream and call: large calls keep the GPU busy */
   const int tcap = B / 2 + 2;
   tmp = (vadc_segment *)malloc( (size_t)S * tcap * sizeof( vadc_segment ) );
   cnt = (int *)malloc( (size_t)S * sizeof( int ) );
   ptmp = o->raw_probabilities ? (float *)malloc( (size_t)S * B * sizeof( float ) ) : 0;
   if ( !segs || !nseg || !tmp || !cnt || ( o->raw_probabilities && ( !probs || !ptmp ) ) ) goto done;
   if ( silero_b200_group_segments_configure( h, &o->seg ) )
   {
      fprintf( stderr, "Error: %s\n", silero_b200_group_last_error() );
      goto done;
   }
   vadc_segmenter fmt;
   vadc_segmenter_init( &fmt, &o->seg );
   double t0 = now_s(), total_speech = 0.0;
   long long total_samples = 0;
   for ( long long n0 = 0; n0 < maxchunks; n0 += B )
   {
      /* streams [0, full) have a full slice left, [full, active) end inside this slice */
      int active = 0, full = 0;
      while ( active < S && fs[active].nchunks > n0 ) ++active;
      while ( full < active && fs[full].nchunks >= n0 + B ) ++full;
      for ( int pass = 0; pass < 2; ++pass )
      {
         int s_begin = pass == 0 ? 0 : full, s_end = pass == 0 ? full : active;
         while ( s_begin < s_end )
         {
            /* pass 0: one call for all full streams; pass 1: runs of streams with the same remaining length, closing them */
            int s_stop = s_end;
            int n = B;
            if ( pass == 1 )
            {
               n = (int)( fs[s_begin].nchunks - n0 );
               s_stop = s_begin + 1;
               while ( s_stop < s_end && fs[s_stop].nchunks == fs[s_begin].nchunks ) ++s_stop;
            }
            const int ns = s_stop - s_begin;
            if ( silero_b200_group_run_streams_segments( h, pcm + (size_t)s_begin * stride + (size_t)n0 * SILERO_B200_CHUNK_SAMPLES, stride, s_begin, ns, n,
                                                         pass == 1 ? 1 : 0, tmp, tcap, cnt, ptmp ) )
            {
               fprintf( stderr, "Error: %s\n", silero_b200_group_last_error() );
               goto done;
            }
            for ( int i = 0; i < ns; ++i )
            {
               const int s = s_begin + i;
               for ( int j = 0; j < cnt[i] && j < tcap && nseg[s] < cap; ++j )
               {
                  float tb, te;
                  vadc_segment_times( &fmt, tmp[(size_t)i * tcap + j], &tb, &te );
                  total_speech += (double)te - (double)tb; /* the --stats line counts the speech of all files */
                  segs[(size_t)s * cap + nseg[s]++] = tmp[(size_t)i * tcap + j];
               }
               if ( probs ) memcpy( probs + (size_t)s * maxchunks + n0, ptmp + (size_t)i * n, (size_t)n * sizeof( float ) );
            }
            total_samples += (long long)ns * n * SILERO_B200_CHUNK_SAMPLES;
            s_begin = s_stop;
         }
      }
      if ( o->stats ) print_stats( total_speech, total_samples, t0 );
   }
   /* streams whose length is an exact multiple of the slice were never closed: flush them (no audio, end of stream) */
   for ( int s = 0; s < S; )
   {
      if ( fs[s].nchunks == 0 || fs[s].nchunks % B != 0 )
      {
         ++s;
         continue;
      }
      int e = s + 1;
      while ( e < S && fs[e].nchunks == fs[s].nchunks ) ++e;
      if ( silero_b200_group_run_streams_segments( h, 0, 0, s, e - s, 0, 1, tmp, tcap, cnt, 0 ) )
      {
         fprintf( stderr, "Error: %s\n", silero_b200_group_last_error() );
         goto done;
      }
      for ( int i = 0; i < e - s; ++i )
         for ( int j = 0; j < cnt[i] && j < tcap && nseg[s + i] < cap; ++j )
         {
            float tb, te;
            vadc_segment_times( &fmt, tmp[(size_t)i * tcap + j], &tb, &te );
            total_speech += (double)te - (double)tb;
            segs[(size_t)( s + i ) * cap + nseg[s + i]++] = tmp[(size_t)i * tcap + j];
         }
      s = e;
   }
   /* print in argument order */
   for ( int order = 0; order < S; ++order )
   {
      int s = 0;
      while ( fs[s].order != order ) ++s;
      if ( S > 1 ) printf( "# %s\n", fs[s].path );
      if ( o->raw_probabilities )
         for ( long long i = 0; i < fs[s].nchunks; ++i ) printf( "%f\n", probs[(size_t)s * maxchunks + i] );
      else
         for ( int j = 0; j < nseg[s]; ++j )
         {
            char line[96];
            vadc_segment_format( &fmt, segs[(size_t)s * cap + j], line, sizeof( line ) );
            fputs( line, stdout );
         }
   }
   fflush( stdout );
   if ( o->stats )
   {
      print_stats( total_speech, total_samples, t0 );
      fputc( '\n', stderr );
   }
   rc = 0;
done:
   if ( f ) fclose( f );
   if ( pcm ) silero_b200_host_free_pinned( pcm );
   for ( int i = 0; i < S; ++i ) free( fs[i].decoded );
   free( segs ); free( nseg ); free( probs ); free( tmp ); free( cnt ); free( ptmp ); free( fs );
   return rc;
}

int main( int argc, char **argv )
{
   cli_opts o;
   memset( &o, 0, sizeof( o ) );
   if ( parse_args( argc, argv, &o ) ) return 1;
   /* weights: --model, else $VADC_B200_WEIGHTS, else weights/silero_v31_16k.testtensor next to the executable
      (the reference embeds the same file at build time, build_msvc.bat:52-54) */
   const char *weights = o.model ? o.model : getenv( "VADC_B200_WEIGHTS" );
   char exe[4096];
   if ( !weights )
   {
      ssize_t n = readlink( "/proc/self/exe", exe, sizeof( exe ) - 64 );
      if ( n > 0 )
      {
         exe[n] = 0;
         char *slash = strrchr( exe, '/' );
         if ( slash ) strcpy( slash + 1, "weights/silero_v31_16k.testtensor" );
         weights = exe;
      }
      else
         weights = "silero_v31_16k.testtensor";
   }
   silero_b200_opts eo;
   silero_b200_default_opts( &eo );
   eo.device = o.ndevices > 0 ? o.devices[0] : o.device;
   eo.max_streams = o.nfiles > 1 ? o.nfiles : 1;
   const int files_mode = !( o.nfiles == 0 || ( o.nfiles == 1 && !is_raw_pcm_name( o.files[0] ) ) );
   if ( files_mode )
   {
      /* many files = many streams, over one or several GPUs (--devices): the group scheduler */
      if ( o.ndevices == 0 ) o.devices[o.ndevices++] = o.device;
      if ( o.ndevices > o.nfiles ) o.ndevices = o.nfiles;
      silero_b200_group *g = 0;
      if ( silero_b200_group_create_from_file( weights, o.devices, o.ndevices, &eo, &g ) != SILERO_B200_OK )
      {
         fprintf( stderr, "Error: %s\n", silero_b200_group_last_error() );
         free( o.files );
         return 1;
      }
      fprintf( stderr, "batch size: %d\n", o.batch ); /* vadc.c:716 */
      const int rcf = run_files( g, &o );
      silero_b200_group_destroy( g );
      free( o.files );
      return rcf;
   }
   silero_b200 *h = 0;
   if ( silero_b200_create_from_file( weights, &eo, &h ) != SILERO_B200_OK )
   {
      fprintf( stderr, "Error: %s\n", silero_b200_last_error() ); /* vadc.c:692-695: init failure ends the run */
      return 1;
   }
   fprintf( stderr, "batch size: %d\n", o.batch ); /* vadc.c:716 */
   int rc;
   if ( o.nfiles == 0 )
      rc = run_fd( h, &o, 0, 0 );
   else if ( o.nfiles == 1 && !is_raw_pcm_name( o.files[0] ) )
   {
      /* the reference's named-input mode: stream the ffmpeg child's output through the same loop as stdin */
      pid_t pid = 0;
      int fd = spawn_ffmpeg( o.files[0], o.start_seconds, o.audio_source, &pid );
      if ( fd < 0 )
         rc = 1;
      else
      {
         rc = run_fd( h, &o, fd, 0 );
         close( fd );
         waitpid( pid, 0, 0 );
      }
   }
   else
      rc = 1; /* (files mode returned above) */
   silero_b200_destroy( h );
   free( o.files );
   return rc;
}

/* vadc_b200/csrc/segmenter.c -- streaming probability -> segment state machine (host C).
 * Interface and reference citations: include/vadc_segmenter.h. */
#include "vadc_segmenter.h"

#include <stdio.h>
#include <string.h>

void vadc_seg_params_default( vadc_seg_params *p )
{
   /* option table defaults, vadc.c:1110-1124 */
   p->min_silence_ms = 200.0f;
   p->min_speech_ms = 250.0f;
   p->threshold = 0.5f;
   p->neg_threshold_relative = 0.15f;
   p->speech_pad_ms = 30.0f;
   p->chunk_samples = 1536;
   p->centiseconds = 0;
}

static int ms_to_chunks( float ms, float chunk_ms )
{
   /* vadc.c:758-768: round to nearest, at least one chunk */
   int n = (int)( ms / chunk_ms + 0.5f );
   return n < 1 ? 1 : n;
}

void vadc_segmenter_init( vadc_segmenter *s, const vadc_seg_params *p )
{
   memset( s, 0, sizeof( *s ) );
   s->p = *p;
   const float sample_rate = (float)16000;
   float chunk_ms = p->chunk_samples / sample_rate * 1000.0f; /* vadc.c:756 */
   s->min_speech_chunks = ms_to_chunks( p->min_speech_ms, chunk_ms );
   s->min_silence_chunks = ms_to_chunks( p->min_silence_ms, chunk_ms );
   s->neg_threshold = p->threshold - p->neg_threshold_relative; /* vadc.c:1244 */
   s->seconds_per_chunk = (float)p->chunk_samples / 16000;      /* vadc.c:846 */
}

void vadc_segment_times( const vadc_segmenter *s, vadc_segment seg, float *start_s, float *end_s )
{
   /* vadc.c:229-240 */
   const float pad = s->p.speech_pad_ms / 1000.0f;
   float e = ( seg.end_chunk * s->seconds_per_chunk ) + pad;
   float b = ( seg.start_chunk * s->seconds_per_chunk ) - pad;
   if ( b < 0.0f ) b = 0.0f;
   *start_s = b;
   *end_s = e;
}

int vadc_segment_format( const vadc_segmenter *s, vadc_segment seg, char *buf, size_t cap )
{
   float b, e;
   vadc_segment_times( s, seg, &b, &e );
   if ( s->p.centiseconds ) /* vadc.c:251-256 */
      return snprintf( buf, cap, "%lld,%lld\n", (long long)( (double)b * 100.0 + 0.5 ), (long long)( (double)e * 100.0 + 0.5 ) );
   return snprintf( buf, cap, "%.2f,%.2f\n", b, e ); /* vadc.c:246 */
}

typedef struct seg_out
{
   vadc_segment *out;
   long long cap, count;
} seg_out;

static void push_out( seg_out *o, vadc_segment seg )
{
   if ( o->out && o->count < o->cap ) o->out[o->count] = seg;
   o->count++;
}

/* vadc.c:262-299: a candidate either extends the buffered segment (their padded spans touch)
   or pushes it out and takes its place */
static void offer_candidate( vadc_segmenter *s, vadc_segment cand, seg_out *o )
{
   if ( !s->buffered_valid )
   {
      s->buffered = cand;
      s->buffered_valid = 1;
      return;
   }
   const float pad = s->p.speech_pad_ms / 1000.0f;
   float cand_start = ( cand.start_chunk * s->seconds_per_chunk ) - pad;
   if ( cand_start < 0.0f ) cand_start = 0.0f;
   float buffered_end = ( s->buffered.end_chunk * s->seconds_per_chunk ) + pad;
   if ( buffered_end >= cand_start )
      s->buffered.end_chunk = cand.end_chunk;
   else
   {
      push_out( o, s->buffered );
      s->buffered = cand;
   }
}

long long vadc_segmenter_feed( vadc_segmenter *s, const float *prob, long long n, vadc_segment *out, long long cap )
{
   seg_out o = { out, cap, 0 };
   const float thr = s->p.threshold;
   const float neg = s->neg_threshold;
   for ( long long i = 0; i < n; ++i )
   {
      const float p = prob[i];
      const int g = s->global_chunk_index;
      /* vadc.c:176-218 */
      if ( p >= thr && s->temp_end > 0 ) s->temp_end = 0;
      if ( !s->triggered )
      {
         if ( p >= thr )
         {
            s->triggered = 1;
            s->current_speech_start = g;
         }
      }
      else if ( p < neg )
      {
         if ( s->temp_end == 0 ) s->temp_end = g; /* 0 doubles as "unset", as in the reference */
         if ( g - s->temp_end >= s->min_silence_chunks )
         {
            if ( s->temp_end - s->current_speech_start >= s->min_speech_chunks )
            {
               vadc_segment cand = { s->current_speech_start, s->temp_end };
               offer_candidate( s, cand, &o );
            }
            s->current_speech_start = 0;
            s->temp_end = 0;
            s->triggered = 0;
         }
      }
      s->global_chunk_index = g + 1;
   }
   return o.count;
}

long long vadc_segmenter_finish( vadc_segmenter *s, vadc_segment *out, long long cap )
{
   seg_out o = { out, cap, 0 };
   if ( s->triggered ) /* vadc.c:1008-1021 */
   {
      const int cs = s->p.chunk_samples;
      int audio_length_samples = ( s->global_chunk_index - 1 ) * cs;
      if ( audio_length_samples - s->current_speech_start * cs > s->min_speech_chunks * cs )
      {
         vadc_segment cand = { s->current_speech_start, audio_length_samples / cs };
         offer_candidate( s, cand, &o );
      }
      s->triggered = 0;
   }
   if ( s->buffered_valid ) /* vadc.c:1023-1026 */
   {
      push_out( &o, s->buffered );
      s->buffered_valid = 0;
   }
   return o.count;
}

size_t vadc_segments_text( const float *prob, long long nchunks, const vadc_seg_params *p, char *text, size_t cap )
{
   vadc_segmenter s;
   vadc_segmenter_init( &s, p );
   size_t len = 0;
   if ( cap ) text[0] = 0;
   vadc_segment tmp[64];
   long long done = 0;
   int finishing = 0;
   for ( ;; )
   {
      long long produced;
      if ( done < nchunks )
      {
         /* small slices so that tmp[] can never overflow: a chunk yields at most one segment */
         long long n = nchunks - done < 64 ? nchunks - done : 64;
         produced = vadc_segmenter_feed( &s, prob + done, n, tmp, 64 );
         done += n;
      }
      else
      {
         produced = vadc_segmenter_finish( &s, tmp, 64 );
         finishing = 1;
      }
      for ( long long i = 0; i < produced && i < 64; ++i )
      {
         char line[96];
         int m = vadc_segment_format( &s, tmp[i], line, sizeof( line ) );
         for ( int j = 0; j < m && len + 1 < cap; ++j ) text[len++] = line[j];
         if ( cap ) text[len] = 0;
      }
      if ( finishing ) break;
   }
   return len;
}

// vadc_b200/csrc/segment_kernel.cuh -- probability -> speech-segment state machine on the device.
//
// Same contract as the host segmenter (segmenter.c; include/vadc_segmenter.h), i.e. vadc's
// feed_probability (vadc.c:165-221), combine_or_emit_speech_segment (vadc.c:262-299) and the
// end-of-stream logic (vadc.c:1005-1027), evaluated as a per-stream scan: one thread per stream walks
// its probabilities in chunk order and carries the FeedState + the one buffered candidate in a
// 32-byte per-stream record that persists across calls. Only finished (start_chunk, end_chunk) pairs
// and a per-stream count leave the device, so the [streams][chunks] probability matrix never has to
// be copied to the host ("final gather of per-stream segments").
//
// The merge test compares two fp32 expressions (vadc.c:275-283); they are evaluated with explicit
// __fmul_rn/__fadd_rn so that no FMA contraction can change a rounding, and the emitted pairs are
// bit-identical to the host state machine's.
#pragma once
#include "common.cuh"

struct SegStateDev // per stream, 32 bytes
{
   int temp_end, current_speech_start, triggered;
   int buffered_start, buffered_end, buffered_valid;
   int global_chunk_index;
   int pad_;
};

struct SegParamsDev
{
   float threshold, neg_threshold, seconds_per_chunk, pad_seconds;
   int min_speech_chunks, min_silence_chunks, chunk_samples;
};

struct SegPair
{
   int start_chunk, end_chunk;
};

// probs: stream s, chunk n at probs[s * stride + off + n], n < nchunks
// segs:  [nstreams][cap] pairs finished by THIS call; counts[s] = how many (may exceed cap: overflow is counted, not stored)
__global__ void __launch_bounds__( 128 )
segment_scan_kernel( const float *__restrict__ probs, long long stride, long long off, int nchunks, SegStateDev *__restrict__ state, int nstreams,
                     SegParamsDev p, int finish, SegPair *__restrict__ segs, int cap, int *__restrict__ counts )
{
   const int s = blockIdx.x * blockDim.x + threadIdx.x;
   if ( s >= nstreams ) return;
   SegStateDev st = state[s];
   SegPair *out = segs + (size_t)s * cap;
   int count = 0;

   auto offer = [&]( int cs, int ce ) {
      // vadc.c:262-299
      if ( !st.buffered_valid )
      {
         st.buffered_start = cs;
         st.buffered_end = ce;
         st.buffered_valid = 1;
         return;
      }
      float cand_start = __fsub_rn( __fmul_rn( (float)cs, p.seconds_per_chunk ), p.pad_seconds );
      if ( cand_start < 0.0f ) cand_start = 0.0f;
      const float buffered_end = __fadd_rn( __fmul_rn( (float)st.buffered_end, p.seconds_per_chunk ), p.pad_seconds );
      if ( buffered_end >= cand_start )
         st.buffered_end = ce;
      else
      {
         if ( count < cap ) out[count] = SegPair{ st.buffered_start, st.buffered_end };
         ++count;
         st.buffered_start = cs;
         st.buffered_end = ce;
      }
   };

   const float *pr = probs + (long long)s * stride + off;
   for ( int n = 0; n < nchunks; ++n )
   {
      const float v = __ldg( pr + n );
      const int g = st.global_chunk_index;
      // vadc.c:176-218
      if ( v >= p.threshold && st.temp_end > 0 ) st.temp_end = 0;
      if ( !st.triggered )
      {
         if ( v >= p.threshold )
         {
            st.triggered = 1;
            st.current_speech_start = g;
         }
      }
      else if ( v < p.neg_threshold )
      {
         if ( st.temp_end == 0 ) st.temp_end = g;
         if ( g - st.temp_end >= p.min_silence_chunks )
         {
            if ( st.temp_end - st.current_speech_start >= p.min_speech_chunks ) offer( st.current_speech_start, st.temp_end );
            st.current_speech_start = 0;
            st.temp_end = 0;
            st.triggered = 0;
         }
      }
      st.global_chunk_index = g + 1;
   }
   if ( finish )
   {
      if ( st.triggered ) // vadc.c:1008-1021
      {
         const int cs = p.chunk_samples;
         const int audio_length_samples = ( st.global_chunk_index - 1 ) * cs;
         if ( audio_length_samples - st.current_speech_start * cs > p.min_speech_chunks * cs ) offer( st.current_speech_start, audio_length_samples / cs );
         st.triggered = 0;
      }
      if ( st.buffered_valid ) // vadc.c:1023-1026
      {
         if ( count < cap ) out[count] = SegPair{ st.buffered_start, st.buffered_end };
         ++count;
         st.buffered_valid = 0;
      }
   }
   state[s] = st;
   counts[s] = count;
}

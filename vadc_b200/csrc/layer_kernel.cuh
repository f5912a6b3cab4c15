// vadc_b200/csrc/layer_kernel.cuh -- one encoder transformer_layer per launch, batched over chunks.
//
// Replaces transformer_layer (transformer.c:237-295) = conv_block (conv.c:761-814: dw_conv_tensor
// :60, pw_conv_tensor :726) -> transformer_block (transformer.c:160-234: dual_head_attention :13,
// layer_norm misc.c:143-210, tensor_linear tensor.h:675-723, softmax tensor.h:751-784)
// -> conv 1x1 with stride (conv.c:715) -> batch_norm1d (misc.c:221-258) -> ReLU, and, for the first
// layer, the mean part of adaptive_audio_normalization_inplace (misc.c:48-121).
//
// Mapping ("lane = token"): the T frames of a chunk are padded to TP = 32/16/8 lanes so that a
// chunk never straddles a warp: 1, 2 or 4 chunks per warp. A lane owns token (chunk, frame) and
// carries its activation row through the whole layer in a private shared-memory row of RS floats
// (RS*4 B = odd multiple of 16 B, so float4 row accesses of a warp are bank-conflict free). All
// cross-token traffic (depthwise taps, attention) stays inside the warp, so the only barriers are
// __syncwarp (layer 4 additionally stages its 134 KB of weights through shared memory per sub-step).
// Every linear layer is an in-lane GEMV whose weight operand is a warp-wide shared-memory
// broadcast; because a broadcast float4 costs a shared-memory wavefront per 4 FMAs, each lane
// processes U = 2 tokens (two different chunks) so every weight load feeds 8 FMAs. The first layer
// reads the spectrogram straight from global memory (one coalesced load per bin and chunk) and gets
// its +-2 depthwise taps by warp shuffle. Attention, layer norm and softmax are sequential in-lane,
// in the reference's order. Tiny irregular dims (T = 25/13/7, C = 16..64) make this CUDA-core work:
// the layers are 6.7 / 4.2 / 2.3 / 8.8 % of the model's FLOPs.
//
// Activations between layers are token-major: [chunk][T][C].
#pragma once
#include "common.cuh"


template <int L>
struct LayerCfg
{
   using P = LayerPack<L>;
   static constexpr int CIN = P::CIN, C = P::C, T = P::T, D = P::D, STRIDE = P::STRIDE, TOUT = P::TOUT;
   static constexpr bool PROJ = P::PROJ != 0;
   static constexpr int TP = T <= 8 ? 8 : ( T <= 16 ? 16 : 32 ); // lanes per chunk
   static constexpr int CPW = 32 / TP;                            // chunks per warp (per token set)
   static constexpr int U = ( C < 64 ) ? 2 : 1;                   // tokens per lane
   // warps per CTA: sized so that rows + weights fill the SM with 8..12 warps
   static constexpr int NWARPS = ( C == 16 ) ? 4 : ( C == 32 ? 8 : 4 );
   static constexpr int THREADS = NWARPS * 32;
   static constexpr int G = NWARPS * CPW * U;                     // chunks per CTA tile
   static constexpr int TOK = THREADS * U;                        // token rows in shared memory
   // row: U[C] | Q[3*D]. U carries y -> u1 -> u2; Q holds the block input X (CIN <= 3*D), then q|k|v of
   // one head, then the FFN hidden row (C <= 3*D). The attention output stays in registers.
   static constexpr int OFF_U = 0, OFF_Q = C;
   static constexpr int RS_RAW = C + 3 * D;
   // smallest RS >= RS_RAW with RS % 8 == 4 (RS*4 bytes = odd multiple of 16)
   static constexpr int RS = RS_RAW + ( ( 4 - ( RS_RAW % 8 ) + 8 ) % 8 );
   static constexpr bool FIRST = ( CIN == VB_BINS ); // consumes the [chunk][129][T] spectrogram layout
   static constexpr bool RESIDENT = ( C < 64 );
   static constexpr int WS = RESIDENT ? P::TOTAL : ( 3 * D * C + 3 * D ); // staged: largest stage (one QKV head)
   static constexpr int SMEM_FLOATS = TOK * RS + WS;
   static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
};

// acc[u][o] += sum_k W[o*ldw + k] * x[u][k], k < K (K % 4 == 0); x[u]: own rows, W: warp broadcast
template <int K, int NB, int U>
__device__ __forceinline__ void lin_acc( const float *const ( &x )[U], const float *__restrict__ W, int ldw, float ( &acc )[U][NB] )
{
#pragma unroll 2
   for ( int k = 0; k < K; k += 4 )
   {
      float4 xv[U];
#pragma unroll
      for ( int u = 0; u < U; ++u ) xv[u] = ld4( x[u] + k );
#pragma unroll
      for ( int o = 0; o < NB; ++o )
      {
         float4 w = ld4( W + o * ldw + k );
#pragma unroll
         for ( int u = 0; u < U; ++u )
         {
            acc[u][o] = fmaf( w.x, xv[u].x, acc[u][o] );
            acc[u][o] = fmaf( w.y, xv[u].y, acc[u][o] );
            acc[u][o] = fmaf( w.z, xv[u].z, acc[u][o] );
            acc[u][o] = fmaf( w.w, xv[u].w, acc[u][o] );
         }
      }
   }
}

// same with the input rows held in registers
template <int K, int NB, int U>
__device__ __forceinline__ void lin_acc_reg( const float ( &x )[U][K], const float *__restrict__ W, int ldw, float ( &acc )[U][NB] )
{
#pragma unroll
   for ( int k = 0; k < K; k += 4 )
   {
#pragma unroll
      for ( int o = 0; o < NB; ++o )
      {
         float4 w = ld4( W + o * ldw + k );
#pragma unroll
         for ( int u = 0; u < U; ++u )
         {
            acc[u][o] = fmaf( w.x, x[u][k], acc[u][o] );
            acc[u][o] = fmaf( w.y, x[u][k + 1], acc[u][o] );
            acc[u][o] = fmaf( w.z, x[u][k + 2], acc[u][o] );
            acc[u][o] = fmaf( w.w, x[u][k + 3], acc[u][o] );
         }
      }
   }
}

// layer_norm (misc.c:143-210) over the row x[0..C) in place
template <int C>
__device__ __forceinline__ void layer_norm_row( float *x, const float *__restrict__ w, const float *__restrict__ b )
{
   float sum = 0.0f;
#pragma unroll
   for ( int i = 0; i < C; i += 4 )
   {
      float4 v = ld4( x + i );
      sum += v.x; sum += v.y; sum += v.z; sum += v.w;
   }
   const float inv = 1.0f / C;
   float mean = sum * inv;
   float vs = 0.0f;
#pragma unroll
   for ( int i = 0; i < C; i += 4 )
   {
      float4 v = ld4( x + i );
      float d0 = v.x - mean, d1 = v.y - mean, d2 = v.z - mean, d3 = v.w - mean;
      vs += d0 * d0; vs += d1 * d1; vs += d2 * d2; vs += d3 * d3;
   }
   float var = vs * inv;
   float rstd = 1.0f / sqrtf( var + 1e-5f );
   float mr = mean * rstd;
#pragma unroll
   for ( int i = 0; i < C; i += 4 )
   {
      float4 v = ld4( x + i ), ww = ld4( w + i ), bb = ld4( b + i );
      v.x = ( v.x * rstd - mr ) * ww.x + bb.x;
      v.y = ( v.y * rstd - mr ) * ww.y + bb.y;
      v.z = ( v.z * rstd - mr ) * ww.z + bb.z;
      v.w = ( v.w * rstd - mr ) * ww.w + bb.w;
      st4( x + i, v );
   }
}

// one head of dual_head_attention (transformer.c:72-143) for token (chunk rows at crows, frame t):
// A = softmax_rows((K Q^T) / sqrt(D)) with rows = K positions; O = A V.
// Not inlined: the fully unrolled body (T x D) is the largest piece of code in the layer and is
// called 2*U times per tile; one copy keeps the kernel inside the instruction cache. The result is
// written over the lane's own k (consumed into registers first; no other lane reads it).
template <int T, int D, int RS, int OFF_Q>
__device__ __noinline__ void attention_head( const float *crows, int t )
{
   const float *mine = crows + t * RS + OFF_Q;
   float *o_out = const_cast<float *>( mine ) + D;
   float kreg[D];
#pragma unroll
   for ( int j = 0; j < D; j += 4 )
   {
      float4 v = ld4( mine + D + j );
      kreg[j] = v.x; kreg[j + 1] = v.y; kreg[j + 2] = v.z; kreg[j + 3] = v.w;
   }
   const float scale = 1.0f / sqrtf( (float)D );
   float s[T];
#pragma unroll
   for ( int tq = 0; tq < T; ++tq )
   {
      const float *q = crows + tq * RS + OFF_Q;
      float acc = 0.0f;
#pragma unroll
      for ( int j = 0; j < D; j += 4 )
      {
         float4 v = ld4( q + j );
         acc = fmaf( kreg[j], v.x, acc );
         acc = fmaf( kreg[j + 1], v.y, acc );
         acc = fmaf( kreg[j + 2], v.z, acc );
         acc = fmaf( kreg[j + 3], v.w, acc );
      }
      s[tq] = acc * scale;
   }
   float mx = s[0];
#pragma unroll
   for ( int tq = 1; tq < T; ++tq ) mx = fmaxf( mx, s[tq] );
   float sum = 0.0f;
#pragma unroll
   for ( int tq = 0; tq < T; ++tq )
   {
      s[tq] = expf( s[tq] - mx );
      sum += s[tq];
   }
   float inv = 1.0f / sum;
   float o[D];
#pragma unroll
   for ( int j = 0; j < D; ++j ) o[j] = 0.0f;
#pragma unroll
   for ( int tq = 0; tq < T; ++tq )
   {
      const float *v = crows + tq * RS + OFF_Q + 2 * D;
      float a = s[tq] * inv;
#pragma unroll
      for ( int j = 0; j < D; j += 4 )
      {
         float4 vv = ld4( v + j );
         o[j] = fmaf( a, vv.x, o[j] );
         o[j + 1] = fmaf( a, vv.y, o[j + 1] );
         o[j + 2] = fmaf( a, vv.z, o[j + 2] );
         o[j + 3] = fmaf( a, vv.w, o[j + 3] );
      }
   }
#pragma unroll
   for ( int j = 0; j < D; j += 4 ) st4( o_out + j, make_float4( o[j], o[j + 1], o[j + 2], o[j + 3] ) );
}

// NORM (first layer only): input is log1p(mag*2^20) and the adaptive-normalization mean is computed
// and subtracted here; otherwise the per-chunk mean comes from mu_in (computed by the STFT kernel),
// or the input is taken as already normalized when mu_in is NULL (parity tap).
//
// Parity taps for the reference's op/block-level fixtures (production launches pass 0, 0):
//   entry 0: `in` is the layer input.  1: `in` is the conv_block output y [chunk][T][C] (skips the
//            conv block).  2: `in` is the transformer_block output [chunk][T][C] (only step 6 runs).
//   tap   0: layer output [chunk][TOUT][C].  Otherwise `out` is [chunk][T][C] holding:
//         1: conv_block output (conv.c:761)   2: dual_head_attention output incl. out-proj (transformer.c:13)
//         3: after the first layer_norm       4: transformer_block output (transformer.c:160)
enum { TAP_LAYER = 0, TAP_CONV_BLOCK = 1, TAP_ATTENTION = 2, TAP_NORM1 = 3, TAP_BLOCK = 4 };
enum { ENTRY_LAYER = 0, ENTRY_BLOCK = 1, ENTRY_CONV = 2 };

template <int L, bool NORM>
__global__ void __launch_bounds__( LayerCfg<L>::THREADS )
layer_kernel( const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ wblob, int nchunks, int entry, int tap,
              const float *__restrict__ mu_in )
{
   using Cfg = LayerCfg<L>;
   using P = LayerPack<L>;
   constexpr int CIN = Cfg::CIN, C = Cfg::C, T = Cfg::T, D = Cfg::D, G = Cfg::G, RS = Cfg::RS;
   constexpr int TP = Cfg::TP, CPW = Cfg::CPW, U = Cfg::U;
   constexpr int OFF_U = Cfg::OFF_U, OFF_Q = Cfg::OFF_Q;
   constexpr int LAYER_THREADS = Cfg::THREADS;
   constexpr bool RES = Cfg::RESIDENT;
   constexpr bool FIRST = Cfg::FIRST;
   constexpr unsigned FULL = 0xffffffffu;

   extern __shared__ __align__( 16 ) float smem[];
   float *rows = smem;
   float *wbuf = smem + Cfg::TOK * RS;

   const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
   const int slot = lane / TP, t = lane - slot * TP;
   float *myrow[U];
   const float *crows[U];
   int gidx[U];
#pragma unroll
   for ( int u = 0; u < U; ++u )
   {
      const int set = warp * U + u; // token set: 32 rows
      myrow[u] = rows + ( set * 32 + lane ) * RS;
      crows[u] = rows + ( set * 32 + slot * TP ) * RS;
      gidx[u] = set * CPW + slot;
   }

   // weight staging: resident layers load the whole blob once; layer 4 stages per sub-step
   auto stage = [&]( int base, int n ) -> const float * {
      if ( RES ) return wbuf + base;
      __syncthreads();
      for ( int i = tid * 4; i < n; i += LAYER_THREADS * 4 ) st4( wbuf + i, __ldg( reinterpret_cast<const float4 *>( wblob + base + i ) ) );
      __syncthreads();
      return wbuf;
   };
   if ( RES )
   {
      for ( int i = tid * 4; i < P::TOTAL; i += LAYER_THREADS * 4 ) st4( wbuf + i, __ldg( reinterpret_cast<const float4 *>( wblob + i ) ) );
      __syncthreads();
   }

   // warp-level copy of this warp's chunks, token-major [chunk][T][W] in global -> row slot `off`
   auto load_rows = [&]( int chunk0, int width, int off ) {
#pragma unroll
      for ( int u = 0; u < U; ++u )
      {
         const int cb = chunk0 + ( warp * U + u ) * CPW;
         const int nvalid = max( 0, min( CPW, nchunks - cb ) );
         const int total = nvalid * T * width;
         const float *src = in + (size_t)cb * ( T * width );
         float *dst = rows + ( ( warp * U + u ) * 32 ) * RS + off;
         for ( int i = lane * 4; i < total; i += 128 )
         {
            int cs = i / ( T * width ), rem = i - cs * ( T * width );
            int tt = rem / width, c = rem - tt * width;
            st4( dst + ( cs * TP + tt ) * RS + c, __ldg( reinterpret_cast<const float4 *>( src + i ) ) );
         }
      }
      __syncwarp();
   };

   const int ntiles = ( nchunks + G - 1 ) / G;
   for ( int tile = blockIdx.x; tile < ntiles; tile += gridDim.x )
   {
      const int chunk0 = tile * G;
      bool live[U];
#pragma unroll
      for ( int u = 0; u < U; ++u ) live[u] = ( t < T ) && ( chunk0 + gidx[u] < nchunks );
      __syncwarp(); // this warp is done with the rows of its previous tile

      auto tap_row = [&]( int off ) {
#pragma unroll
         for ( int u = 0; u < U; ++u )
            if ( live[u] )
            {
               float *dst = out + ( (size_t)( chunk0 + gidx[u] ) * T + t ) * C;
#pragma unroll
               for ( int c = 0; c < C; c += 4 ) st4( dst + c, ld4( myrow[u] + off + c ) );
            }
      };

      if ( entry != ENTRY_LAYER ) load_rows( chunk0, C, OFF_U );
      if ( entry == ENTRY_LAYER )
      {
         // ---- 1+2. input -> conv_block -> U ---------------------------------------------------
         const float *wa = RES ? wbuf : ( stage( P::DW, P::QKV - P::DW ) - P::DW );
         if ( FIRST )
         {
            // one chunk per warp per token set: lane = frame; the spectrogram is read from global
            const float *sp[U];
#pragma unroll
            for ( int u = 0; u < U; ++u ) sp[u] = in + (size_t)min( chunk0 + gidx[u], nchunks - 1 ) * ( VB_BINS * T ) + min( t, T - 1 );
            float mu[U];
#pragma unroll
            for ( int u = 0; u < U; ++u ) mu[u] = ( !NORM && mu_in ) ? __ldg( mu_in + min( chunk0 + gidx[u], nchunks - 1 ) ) : 0.0f;
            if ( NORM )
            {
               const float gk[7] = { 0.03663284704089164733887f, 0.11128076165914535522461f,
                                     0.21674531698226928710938f, 0.27068215608596801757812f,
                                     0.21674531698226928710938f, 0.11128076165914535522461f,
                                     0.03663284704089164733887f };
#pragma unroll
               for ( int u = 0; u < U; ++u )
               {
                  // misc.c:48-62: per-frame mean over the 129 bins, sequential
                  float s = 0.0f;
#pragma unroll 8
                  for ( int f = 0; f < VB_BINS; ++f ) s = __fadd_rn( s, __ldg( sp[u] + f * T ) );
                  const float m = s / (float)VB_BINS;
                  // misc.c:64-66: reflect pad 3 + 7-tap smoothing (generic conv path: 0 + left-to-right)
                  float v = 0.0f;
#pragma unroll
                  for ( int k = 0; k < 7; ++k )
                  {
                     int idx = t + k - 3;
                     if ( idx < 0 ) idx = -idx;
                     if ( idx >= T ) idx = 2 * ( T - 1 ) - idx;
                     v = __fadd_rn( v, __fmul_rn( __shfl_sync( FULL, m, idx & 31 ), gk[k] ) );
                  }
                  // misc.c:68-82: mean over the 25 smoothed values, sequential
                  float a = 0.0f;
                  for ( int i = 0; i < T; ++i ) a = __fadd_rn( a, __shfl_sync( FULL, v, i ) );
                  mu[u] = a / (float)T;
               }
            }
            float acc[U][2 * C];
#pragma unroll
            for ( int u = 0; u < U; ++u )
#pragma unroll
               for ( int o = 0; o < 2 * C; ++o ) acc[u][o] = 0.0f;
            const float *dw = wa + P::DW;
            const float *pw = wa + P::PW;
            constexpr int FB = 4; // bins per batch of loads; the next batch is in flight while this one is used
            float xnext[FB][U];
#pragma unroll
            for ( int k = 0; k < FB; ++k )
#pragma unroll
               for ( int u = 0; u < U; ++u ) xnext[k][u] = __ldg( sp[u] + k * T );
#pragma unroll 1
            for ( int f0 = 0; f0 < VB_BINS; f0 += FB )
            {
               float xin[FB][U];
#pragma unroll
               for ( int k = 0; k < FB; ++k )
#pragma unroll
                  for ( int u = 0; u < U; ++u )
                  {
                     xin[k][u] = xnext[k][u];
                     xnext[k][u] = ( f0 + FB + k < VB_BINS ) ? __ldg( sp[u] + ( f0 + FB + k ) * T ) : 0.0f;
                  }
#pragma unroll
               for ( int k = 0; k < FB; ++k )
               {
                  const int f = f0 + k;
                  if ( f < VB_BINS )
                  {
                     const float4 w0 = ld4( dw + f * 8 ), w1 = ld4( dw + f * 8 + 4 );
                     float dv[U], x0[U];
#pragma unroll
                     for ( int u = 0; u < U; ++u )
                     {
                        // zero padding applies to the normalized signal: absent taps contribute nothing
                        x0[u] = live[u] ? xin[k][u] - mu[u] : 0.0f;
                        float xm1 = __shfl_up_sync( FULL, x0[u], 1 ), xm2 = __shfl_up_sync( FULL, x0[u], 2 );
                        float xp1 = __shfl_down_sync( FULL, x0[u], 1 ), xp2 = __shfl_down_sync( FULL, x0[u], 2 );
                        if ( t < 1 ) xm1 = 0.0f;
                        if ( t < 2 ) xm2 = 0.0f;
                        if ( t + 1 >= T ) xp1 = 0.0f;
                        if ( t + 2 >= T ) xp2 = 0.0f;
                        float d = w1.y; // bias
                        d = fmaf( xm2, w0.x, d );
                        d = fmaf( xm1, w0.y, d );
                        d = fmaf( x0[u], w0.z, d );
                        d = fmaf( xp1, w0.w, d );
                        d = fmaf( xp2, w1.x, d );
                        dv[u] = fmaxf( d, 0.0f );
                     }
                     const float *wf = pw + f * ( 2 * C );
#pragma unroll
                     for ( int o = 0; o < C; o += 4 )
                     {
                        const float4 a = ld4( wf + o ), b = ld4( wf + C + o );
#pragma unroll
                        for ( int u = 0; u < U; ++u )
                        {
                           acc[u][o] = fmaf( a.x, dv[u], acc[u][o] );
                           acc[u][o + 1] = fmaf( a.y, dv[u], acc[u][o + 1] );
                           acc[u][o + 2] = fmaf( a.z, dv[u], acc[u][o + 2] );
                           acc[u][o + 3] = fmaf( a.w, dv[u], acc[u][o + 3] );
                           acc[u][C + o] = fmaf( b.x, x0[u], acc[u][C + o] );
                           acc[u][C + o + 1] = fmaf( b.y, x0[u], acc[u][C + o + 1] );
                           acc[u][C + o + 2] = fmaf( b.z, x0[u], acc[u][C + o + 2] );
                           acc[u][C + o + 3] = fmaf( b.w, x0[u], acc[u][C + o + 3] );
                        }
                     }
                  }
               }
            }
            const float *pb = wa + P::PWB;
#pragma unroll
            for ( int u = 0; u < U; ++u )
#pragma unroll
               for ( int o = 0; o < C; o += 4 )
                  st4( myrow[u] + OFF_U + o, make_float4( fmaxf( acc[u][o] + acc[u][C + o] + pb[o], 0.0f ),
                                                          fmaxf( acc[u][o + 1] + acc[u][C + o + 1] + pb[o + 1], 0.0f ),
                                                          fmaxf( acc[u][o + 2] + acc[u][C + o + 2] + pb[o + 2], 0.0f ),
                                                          fmaxf( acc[u][o + 3] + acc[u][C + o + 3] + pb[o + 3], 0.0f ) ) );
         }
         else
         {
            load_rows( chunk0, CIN, OFF_Q );
            // depthwise k=5 zero-pad 2 + bias + ReLU, 4 channels at a time, into registers
            float dreg[U][CIN];
            const float *dw = wa + P::DW;
#pragma unroll
            for ( int u = 0; u < U; ++u )
               if ( live[u] )
               {
#pragma unroll
                  for ( int c = 0; c < CIN; c += 4 )
                  {
                     float4 x[5];
#pragma unroll
                     for ( int k = 0; k < 5; ++k )
                     {
                        int tt = t + k - 2;
                        x[k] = ( tt >= 0 && tt < T ) ? ld4( crows[u] + tt * RS + OFF_Q + c ) : make_float4( 0.f, 0.f, 0.f, 0.f );
                     }
#pragma unroll
                     for ( int e = 0; e < 4; ++e )
                     {
                        float4 w0 = ld4( dw + ( c + e ) * 8 ), w1 = ld4( dw + ( c + e ) * 8 + 4 );
                        float dv = w1.y;
                        dv = fmaf( reinterpret_cast<const float *>( &x[0] )[e], w0.x, dv );
                        dv = fmaf( reinterpret_cast<const float *>( &x[1] )[e], w0.y, dv );
                        dv = fmaf( reinterpret_cast<const float *>( &x[2] )[e], w0.z, dv );
                        dv = fmaf( reinterpret_cast<const float *>( &x[3] )[e], w0.w, dv );
                        dv = fmaf( reinterpret_cast<const float *>( &x[4] )[e], w1.x, dv );
                        dreg[u][c + e] = fmaxf( dv, 0.0f );
                     }
                  }
               }
               else
               {
#pragma unroll
                  for ( int c = 0; c < CIN; ++c ) dreg[u][c] = 0.0f;
               }
            // pointwise (+ projection of the block input, or identity residual) + ReLU
            const float *pw = wa + P::PW;
            const float *pb = wa + P::PWB;
            const float *xrow[U];
#pragma unroll
            for ( int u = 0; u < U; ++u ) xrow[u] = myrow[u] + OFF_Q;
            constexpr int NB = 8;
#pragma unroll 1
            for ( int ob = 0; ob < C; ob += NB )
            {
               float acc[U][NB];
#pragma unroll
               for ( int u = 0; u < U; ++u )
#pragma unroll
                  for ( int o = 0; o < NB; ++o ) acc[u][o] = pb[ob + o];
#pragma unroll
               for ( int k = 0; k < CIN; k += 4 )
               {
#pragma unroll
                  for ( int o = 0; o < NB; ++o )
                  {
                     float4 w = ld4( pw + ( ob + o ) * P::KP + k );
#pragma unroll
                     for ( int u = 0; u < U; ++u )
                     {
                        acc[u][o] = fmaf( w.x, dreg[u][k], acc[u][o] );
                        acc[u][o] = fmaf( w.y, dreg[u][k + 1], acc[u][o] );
                        acc[u][o] = fmaf( w.z, dreg[u][k + 2], acc[u][o] );
                        acc[u][o] = fmaf( w.w, dreg[u][k + 3], acc[u][o] );
                     }
                  }
               }
               if ( Cfg::PROJ )
                  lin_acc<CIN, NB, U>( xrow, pw + ob * P::KP + CIN, P::KP, acc );
               else
               {
#pragma unroll
                  for ( int u = 0; u < U; ++u )
#pragma unroll
                     for ( int o = 0; o < NB; ++o ) acc[u][o] += xrow[u][ob + o];
               }
#pragma unroll
               for ( int u = 0; u < U; ++u )
#pragma unroll
                  for ( int o = 0; o < NB; o += 4 )
                     st4( myrow[u] + OFF_U + ob + o, make_float4( fmaxf( acc[u][o], 0.f ), fmaxf( acc[u][o + 1], 0.f ),
                                                                  fmaxf( acc[u][o + 2], 0.f ), fmaxf( acc[u][o + 3], 0.f ) ) );
            }
         }
      } // entry == ENTRY_LAYER
      if ( tap == TAP_CONV_BLOCK )
      {
         tap_row( OFF_U );
         continue;
      }

      const float *urow[U], *hrow[U];
#pragma unroll
      for ( int u = 0; u < U; ++u )
      {
         urow[u] = myrow[u] + OFF_U;
         hrow[u] = myrow[u] + OFF_Q; // FFN hidden row
      }

      if ( entry != ENTRY_CONV )
      {
         // ---- 3. attention, one head at a time: QKV_h -> Q, then A V -> registers ------------------
         float att[U][C];
#pragma unroll
         for ( int u = 0; u < U; ++u )
#pragma unroll
            for ( int c = 0; c < C; ++c ) att[u][c] = 0.0f;
#pragma unroll
         for ( int h = 0; h < 2; ++h )
         {
            const float *wq = RES ? ( wbuf + P::QKV + h * P::QH ) : stage( P::QKV + h * P::QH, P::QH );
            __syncwarp(); // Q of the previous head / X of this tile no longer read by other lanes
            {
               constexpr int NB = 12;
#pragma unroll 1
               for ( int ob = 0; ob < 3 * D; ob += NB )
               {
                  float acc[U][NB];
#pragma unroll
                  for ( int u = 0; u < U; ++u )
#pragma unroll
                     for ( int o = 0; o < NB; ++o ) acc[u][o] = wq[3 * D * C + ob + o];
                  lin_acc<C, NB, U>( urow, wq + ob * C, C, acc );
#pragma unroll
                  for ( int u = 0; u < U; ++u )
#pragma unroll
                     for ( int o = 0; o < NB; o += 4 )
                        st4( myrow[u] + OFF_Q + ob + o, make_float4( acc[u][o], acc[u][o + 1], acc[u][o + 2], acc[u][o + 3] ) );
               }
            }
            __syncwarp();
#pragma unroll
            for ( int u = 0; u < U; ++u )
               if ( live[u] )
               {
                  attention_head<T, D, RS, OFF_Q>( crows[u], t );
#pragma unroll
                  for ( int jj = 0; jj < D; jj += 4 )
                  {
                     float4 v = ld4( myrow[u] + OFF_Q + D + jj );
                     att[u][h * D + jj] = v.x; att[u][h * D + jj + 1] = v.y; att[u][h * D + jj + 2] = v.z; att[u][h * D + jj + 3] = v.w;
                  }
               }
         }

         // ---- 4. out-proj + residual + LayerNorm1 -> U -------------------------------------------
         {
            const float *w = RES ? wbuf : ( stage( P::AO, P::F1 - P::AO ) - P::AO );
            constexpr int NB = 8;
#pragma unroll 1
            for ( int ob = 0; ob < C; ob += NB )
            {
               float acc[U][NB];
#pragma unroll
               for ( int u = 0; u < U; ++u )
#pragma unroll
                  for ( int o = 0; o < NB; ++o ) acc[u][o] = w[P::AOB + ob + o];
               lin_acc_reg<C, NB, U>( att, w + P::AO + ob * C, C, acc );
#pragma unroll
               for ( int u = 0; u < U; ++u )
#pragma unroll
                  for ( int o = 0; o < NB; ++o )
                     myrow[u][OFF_U + ob + o] = ( tap == TAP_ATTENTION ) ? acc[u][o] : myrow[u][OFF_U + ob + o] + acc[u][o];
            }
            if ( tap != TAP_ATTENTION )
            {
#pragma unroll
               for ( int u = 0; u < U; ++u ) layer_norm_row<C>( myrow[u] + OFF_U, w + P::LN1W, w + P::LN1B );
            }
         }
         if ( tap == TAP_ATTENTION || tap == TAP_NORM1 )
         {
            tap_row( OFF_U );
            continue;
         }
         // ---- 5. FFN: linear1 + ReLU -> Q slot ; linear2 + residual + LayerNorm2 -> U -------------
         {
            const float *w = RES ? wbuf : ( stage( P::F1, P::F2 - P::F1 ) - P::F1 );
            __syncwarp(); // the other lanes are done reading this row's q|k|v
            constexpr int NB = 16;
#pragma unroll 1
            for ( int ob = 0; ob < C; ob += NB )
            {
               float acc[U][NB];
#pragma unroll
               for ( int u = 0; u < U; ++u )
#pragma unroll
                  for ( int o = 0; o < NB; ++o ) acc[u][o] = w[P::F1B + ob + o];
               lin_acc<C, NB, U>( urow, w + P::F1 + ob * C, C, acc );
#pragma unroll
               for ( int u = 0; u < U; ++u )
#pragma unroll
                  for ( int o = 0; o < NB; o += 4 )
                     st4( myrow[u] + OFF_Q + ob + o, make_float4( fmaxf( acc[u][o], 0.f ), fmaxf( acc[u][o + 1], 0.f ),
                                                                  fmaxf( acc[u][o + 2], 0.f ), fmaxf( acc[u][o + 3], 0.f ) ) );
            }
         }
         {
            const float *w = RES ? wbuf : ( stage( P::F2, P::CV - P::F2 ) - P::F2 );
            constexpr int NB = 16;
#pragma unroll 1
            for ( int ob = 0; ob < C; ob += NB )
            {
               float acc[U][NB];
#pragma unroll
               for ( int u = 0; u < U; ++u )
#pragma unroll
                  for ( int o = 0; o < NB; ++o ) acc[u][o] = w[P::F2B + ob + o];
               lin_acc<C, NB, U>( hrow, w + P::F2 + ob * C, C, acc );
#pragma unroll
               for ( int u = 0; u < U; ++u )
#pragma unroll
                  for ( int o = 0; o < NB; ++o ) myrow[u][OFF_U + ob + o] += acc[u][o];
            }
#pragma unroll
            for ( int u = 0; u < U; ++u ) layer_norm_row<C>( myrow[u] + OFF_U, w + P::LN2W, w + P::LN2B );
         }
      } // entry != ENTRY_CONV
      if ( tap == TAP_BLOCK )
      {
         tap_row( OFF_U );
         continue;
      }
      // ---- 6. conv 1x1 (stride) + BatchNorm(eval) + ReLU -> global -----------------------------
      {
         const float *w = RES ? wbuf : ( stage( P::CV, P::TOTAL - P::CV ) - P::CV );
         constexpr int NB = 16;
#pragma unroll 1
         for ( int ob = 0; ob < C; ob += NB )
         {
            float acc[U][NB];
#pragma unroll
            for ( int u = 0; u < U; ++u )
#pragma unroll
               for ( int o = 0; o < NB; ++o ) acc[u][o] = 0.0f;
            lin_acc<C, NB, U>( urow, w + P::CV + ob * C, C, acc );
#pragma unroll
            for ( int u = 0; u < U; ++u )
               if ( live[u] && ( t % Cfg::STRIDE ) == 0 )
               {
                  float *o_row = out + ( (size_t)( chunk0 + gidx[u] ) * Cfg::TOUT + t / Cfg::STRIDE ) * C;
                  float r[NB];
#pragma unroll
                  for ( int o = 0; o < NB; ++o )
                  {
                     float z = acc[u][o] + w[P::CVB + ob + o];
                     float nv = ( z - w[P::BNM + ob + o] ) / w[P::BNS + ob + o]; // misc.c:251 true division
                     r[o] = fmaxf( nv * w[P::BNW + ob + o] + w[P::BNB + ob + o], 0.0f );
                  }
#pragma unroll
                  for ( int o = 0; o < NB; o += 4 ) st4( o_row + ob + o, make_float4( r[o], r[o + 1], r[o + 2], r[o + 3] ) );
               }
         }
      }
   }
}

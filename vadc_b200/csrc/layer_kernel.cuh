// vadc_b200/csrc/layer_kernel.cuh -- one encoder transformer_layer per launch, batched over chunks.
//
// Replaces transformer_layer (transformer.c:237-295) = conv_block (conv.c:761-814: dw_conv_tensor
// :60, pw_conv_tensor :726) -> transformer_block (transformer.c:160-234: dual_head_attention :13,
// layer_norm misc.c:143-210, tensor_linear tensor.h:675-723, softmax tensor.h:751-784)
// -> conv 1x1 with stride (conv.c:715) -> batch_norm1d (misc.c:221-258) -> ReLU, and, for the first
// layer, the mean part of adaptive_audio_normalization_inplace (misc.c:48-121).
//
// Mapping ("thread = token row"): a CTA of 128 threads owns a tile of G = 128/T chunks; thread r
// owns token (chunk g, frame t) and carries its activation row through the whole layer in a
// private shared-memory row of RS floats (RS*4 is an odd multiple of 16 bytes, so float4 row
// accesses of a warp are bank-conflict free). Every linear layer is an in-thread GEMV whose weight
// operand is a warp-wide shared-memory broadcast (all lanes read the same float4), so the FP32
// pipe sees K*N FMAs per (K/4)*(N+1) shared loads. Attention reads the other rows of the same
// chunk from shared memory; layer norm and softmax are sequential in-thread, in the reference's
// order. Tiny irregular dims (T = 25/13/7, C = 16..64) make this CUDA-core work, not tensor-core
// work: the layer is 0.3..8.8 % of the model's FLOPs each.
//
// Activations between layers are token-major: [chunk][T][C].
#pragma once
#include "common.cuh"

#define LAYER_THREADS 128

template <int L>
struct LayerCfg
{
   using P = LayerPack<L>;
   static constexpr int CIN = P::CIN, C = P::C, T = P::T, D = P::D, STRIDE = P::STRIDE, TOUT = P::TOUT;
   static constexpr bool PROJ = P::PROJ != 0;
   static constexpr int G = LAYER_THREADS / T; // chunks per tile
   static constexpr int R = G * T;             // live rows
   // row: U[C] | Q[3*D] | O[C]   (O doubles as the input row X for L>0: CIN <= C)
   static constexpr int OFF_U = 0, OFF_Q = C, OFF_O = C + 3 * D;
   static constexpr int RS_RAW = 2 * C + 3 * D;
   // smallest RS >= RS_RAW with RS % 8 == 4 (RS*4 bytes = odd multiple of 16)
   static constexpr int RS = RS_RAW + ( ( 4 - ( RS_RAW % 8 ) + 8 ) % 8 );
   static constexpr bool RESIDENT = ( C < 64 );
   static constexpr int WS = RESIDENT ? P::TOTAL : ( 3 * D * C + 3 * D ); // staged: largest stage (one QKV head)
   static constexpr bool FIRST = ( CIN == VB_BINS ); // consumes the [chunk][129][T] spectrogram layout
   static constexpr int SPEC = FIRST ? G * VB_BINS * T + 2 * LAYER_THREADS : 0;
   static constexpr int SMEM_FLOATS = LAYER_THREADS * RS + WS + SPEC;
   static constexpr int SMEM_BYTES = SMEM_FLOATS * 4;
};

// acc[o] += sum_k W[o*ldw + k] * x[k], k < K (K % 4 == 0); x: own row (float4 aligned), W: broadcast
template <int K, int NB>
__device__ __forceinline__ void lin_acc( const float *__restrict__ x, const float *__restrict__ W, int ldw, float ( &acc )[NB] )
{
#pragma unroll 4
   for ( int k = 0; k < K; k += 4 )
   {
      float4 xv = ld4( x + k );
#pragma unroll
      for ( int o = 0; o < NB; ++o )
      {
         float4 w = ld4( W + o * ldw + k );
         acc[o] = fmaf( w.x, xv.x, acc[o] );
         acc[o] = fmaf( w.y, xv.y, acc[o] );
         acc[o] = fmaf( w.z, xv.z, acc[o] );
         acc[o] = fmaf( w.w, xv.w, acc[o] );
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
template <int T, int D, int RS, int OFF_Q>
__device__ __forceinline__ void attention_head( const float *__restrict__ crows, int t, float *__restrict__ o_out )
{
   const float *mine = crows + t * RS + OFF_Q;
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
// and subtracted here; otherwise the input is taken as already normalized (parity tap).
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
__global__ void __launch_bounds__( LAYER_THREADS )
layer_kernel( const float *__restrict__ in, float *__restrict__ out, const float *__restrict__ wblob, int nchunks, int entry, int tap )
{
   using Cfg = LayerCfg<L>;
   using P = LayerPack<L>;
   constexpr int CIN = Cfg::CIN, C = Cfg::C, T = Cfg::T, D = Cfg::D, G = Cfg::G, R = Cfg::R, RS = Cfg::RS;
   constexpr int OFF_U = Cfg::OFF_U, OFF_Q = Cfg::OFF_Q, OFF_O = Cfg::OFF_O;
   constexpr bool RES = Cfg::RESIDENT;
   constexpr bool FIRST = Cfg::FIRST;

   extern __shared__ __align__( 16 ) float smem[];
   float *rows = smem;
   float *wbuf = smem + LAYER_THREADS * RS;
   float *spec = wbuf + Cfg::WS; // first layer only

   const int tid = threadIdx.x;
   const int g = tid / T, t = tid - g * T;
   float *myrow = rows + tid * RS;
   const float *crows = rows + g * T * RS; // rows of my chunk

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
   }

   const int ntiles = ( nchunks + G - 1 ) / G;
   for ( int tile = blockIdx.x; tile < ntiles; tile += gridDim.x )
   {
      const int chunk0 = tile * G;
      const int gvalid = min( G, nchunks - chunk0 );
      const bool live = ( tid < R ) && ( g < gvalid );
      __syncthreads(); // previous tile fully consumed (and resident weights visible)

      // copies row-private values [chunk][T][C] <-> the U slots (taps and alternate entries)
      auto load_rows_u = [&]() {
         const float *src = in + (size_t)chunk0 * ( T * C );
         const int n = gvalid * T * C;
         for ( int i = tid * 4; i < n; i += LAYER_THREADS * 4 )
         {
            int r = i / C, c = i - r * C;
            st4( rows + r * RS + OFF_U + c, __ldg( reinterpret_cast<const float4 *>( src + i ) ) );
         }
      };
      auto tap_row = [&]( const float *v ) {
         if ( live )
         {
            float *dst = out + ( (size_t)( chunk0 + g ) * T + t ) * C;
#pragma unroll
            for ( int c = 0; c < C; c += 4 ) st4( dst + c, ld4( v + c ) );
         }
      };

      if ( entry != ENTRY_LAYER )
      {
         load_rows_u();
         __syncthreads();
      }
      if ( entry == ENTRY_LAYER )
      {
      // ---- 1. input tile -> shared ---------------------------------------------------------
      if ( FIRST )
      {
         const float *src = in + (size_t)chunk0 * ( VB_BINS * T );
         const int n = gvalid * VB_BINS * T;
         for ( int i = tid; i < n; i += LAYER_THREADS ) spec[i] = __ldg( src + i );
      }
      else
      {
         const float *src = in + (size_t)chunk0 * ( T * CIN );
         const int n = gvalid * T * CIN;
         for ( int i = tid * 4; i < n; i += LAYER_THREADS * 4 )
         {
            int r = i / CIN, c = i - r * CIN;
            st4( rows + r * RS + OFF_O + c, __ldg( reinterpret_cast<const float4 *>( src + i ) ) );
         }
      }
      __syncthreads();

      // ---- 2. conv_block -> U ---------------------------------------------------------------
      const float *wa = RES ? wbuf : ( stage( P::DW, P::QKV - P::DW ) - P::DW );
      if ( FIRST )
      {
         float *mbuf = spec + G * VB_BINS * T;
         float *sbuf = mbuf + LAYER_THREADS;
         const float *sp = spec + g * ( VB_BINS * T );
         float mu = 0.0f;
         if ( NORM )
         {
            // misc.c:48-62: per-frame mean over the 129 bins, sequential
            if ( live )
            {
               float s = 0.0f;
               for ( int f = 0; f < VB_BINS; ++f ) s = __fadd_rn( s, sp[f * T + t] );
               mbuf[tid] = s / (float)VB_BINS;
            }
            __syncthreads();
            // misc.c:64-66: reflect pad 3 + 7-tap smoothing (generic conv path: 0 + left-to-right)
            if ( live )
            {
               const float gk[7] = { 0.03663284704089164733887f, 0.11128076165914535522461f,
                                     0.21674531698226928710938f, 0.27068215608596801757812f,
                                     0.21674531698226928710938f, 0.11128076165914535522461f,
                                     0.03663284704089164733887f };
               float v = 0.0f;
#pragma unroll
               for ( int k = 0; k < 7; ++k )
               {
                  int idx = t + k - 3;
                  if ( idx < 0 ) idx = -idx;
                  if ( idx >= T ) idx = 2 * ( T - 1 ) - idx;
                  v = __fadd_rn( v, __fmul_rn( mbuf[g * T + idx], gk[k] ) );
               }
               sbuf[tid] = v;
            }
            __syncthreads();
            // misc.c:68-82: mean over the 25 smoothed values
            if ( live )
            {
               float s = 0.0f;
               for ( int i = 0; i < T; ++i ) s = __fadd_rn( s, sbuf[g * T + i] );
               mu = s / (float)T;
            }
         }
         if ( live )
         {
            float acc[2 * C];
#pragma unroll
            for ( int o = 0; o < 2 * C; ++o ) acc[o] = 0.0f;
            const float *dw = wa + P::DW;
            const float *pw = wa + P::PW;
            for ( int f = 0; f < VB_BINS; ++f )
            {
               const float *xf = sp + f * T;
               float4 w0 = ld4( dw + f * 8 ), w1 = ld4( dw + f * 8 + 4 );
               // zero padding applies to the normalized signal: absent taps contribute nothing
               float xm2 = ( t >= 2 ) ? xf[t - 2] - mu : 0.0f;
               float xm1 = ( t >= 1 ) ? xf[t - 1] - mu : 0.0f;
               float x0 = xf[t] - mu;
               float xp1 = ( t + 1 < T ) ? xf[t + 1] - mu : 0.0f;
               float xp2 = ( t + 2 < T ) ? xf[t + 2] - mu : 0.0f;
               float dv = w1.y; // bias
               dv = fmaf( xm2, w0.x, dv );
               dv = fmaf( xm1, w0.y, dv );
               dv = fmaf( x0, w0.z, dv );
               dv = fmaf( xp1, w0.w, dv );
               dv = fmaf( xp2, w1.x, dv );
               dv = fmaxf( dv, 0.0f );
               const float *wf = pw + f * ( 2 * C );
#pragma unroll
               for ( int o = 0; o < C; o += 4 )
               {
                  float4 a = ld4( wf + o ), b = ld4( wf + C + o );
                  acc[o] = fmaf( a.x, dv, acc[o] );
                  acc[o + 1] = fmaf( a.y, dv, acc[o + 1] );
                  acc[o + 2] = fmaf( a.z, dv, acc[o + 2] );
                  acc[o + 3] = fmaf( a.w, dv, acc[o + 3] );
                  acc[C + o] = fmaf( b.x, x0, acc[C + o] );
                  acc[C + o + 1] = fmaf( b.y, x0, acc[C + o + 1] );
                  acc[C + o + 2] = fmaf( b.z, x0, acc[C + o + 2] );
                  acc[C + o + 3] = fmaf( b.w, x0, acc[C + o + 3] );
               }
            }
            const float *pb = wa + P::PWB;
#pragma unroll
            for ( int o = 0; o < C; ++o ) myrow[OFF_U + o] = fmaxf( acc[o] + acc[C + o] + pb[o], 0.0f );
         }
      }
      else
      {
         if ( live )
         {
            // depthwise k=5 zero-pad 2 + bias + ReLU, 4 channels at a time, into registers
            float dreg[CIN];
            const float *dw = wa + P::DW;
#pragma unroll
            for ( int c = 0; c < CIN; c += 4 )
            {
               float4 x[5];
#pragma unroll
               for ( int k = 0; k < 5; ++k )
               {
                  int tt = t + k - 2;
                  x[k] = ( tt >= 0 && tt < T ) ? ld4( crows + tt * RS + OFF_O + c ) : make_float4( 0.f, 0.f, 0.f, 0.f );
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
                  dreg[c + e] = fmaxf( dv, 0.0f );
               }
            }
            // pointwise (+ projection of the block input, or identity residual) + ReLU
            const float *pw = wa + P::PW;
            const float *pb = wa + P::PWB;
            const float *xrow = myrow + OFF_O;
            constexpr int NB = 16;
#pragma unroll 1
            for ( int ob = 0; ob < C; ob += NB )
            {
               float acc[NB];
#pragma unroll
               for ( int o = 0; o < NB; ++o ) acc[o] = pb[ob + o];
#pragma unroll
               for ( int k = 0; k < CIN; k += 4 )
               {
#pragma unroll
                  for ( int o = 0; o < NB; ++o )
                  {
                     float4 w = ld4( pw + ( ob + o ) * P::KP + k );
                     acc[o] = fmaf( w.x, dreg[k], acc[o] );
                     acc[o] = fmaf( w.y, dreg[k + 1], acc[o] );
                     acc[o] = fmaf( w.z, dreg[k + 2], acc[o] );
                     acc[o] = fmaf( w.w, dreg[k + 3], acc[o] );
                  }
               }
               if ( Cfg::PROJ )
                  lin_acc<CIN, NB>( xrow, pw + ob * P::KP + CIN, P::KP, acc );
               else
               {
#pragma unroll
                  for ( int o = 0; o < NB; ++o ) acc[o] += xrow[ob + o];
               }
#pragma unroll
               for ( int o = 0; o < NB; o += 4 )
                  st4( myrow + OFF_U + ob + o, make_float4( fmaxf( acc[o], 0.f ), fmaxf( acc[o + 1], 0.f ),
                                                            fmaxf( acc[o + 2], 0.f ), fmaxf( acc[o + 3], 0.f ) ) );
            }
         }
      }

      } // entry == ENTRY_LAYER
      if ( tap == TAP_CONV_BLOCK )
      {
         tap_row( myrow + OFF_U );
         continue;
      }

      if ( entry != ENTRY_CONV )
      {
      // ---- 3. attention, one head at a time: QKV_h -> Q, then A V -> O[h*D..] -----------------
#pragma unroll 1
      for ( int h = 0; h < 2; ++h )
      {
         const float *wq = RES ? ( wbuf + P::QKV + h * P::QH ) : stage( P::QKV + h * P::QH, P::QH );
         if ( RES ) __syncthreads(); // Q of the previous head / X of this tile no longer read by others
         if ( live )
         {
            constexpr int NB = 12;
#pragma unroll 1
            for ( int ob = 0; ob < 3 * D; ob += NB )
            {
               float acc[NB];
#pragma unroll
               for ( int o = 0; o < NB; ++o ) acc[o] = wq[3 * D * C + ob + o];
               lin_acc<C, NB>( myrow + OFF_U, wq + ob * C, C, acc );
#pragma unroll
               for ( int o = 0; o < NB; o += 4 ) st4( myrow + OFF_Q + ob + o, make_float4( acc[o], acc[o + 1], acc[o + 2], acc[o + 3] ) );
            }
         }
         __syncthreads();
         if ( live ) attention_head<T, D, RS, OFF_Q>( crows, t, myrow + OFF_O + h * D );
      }

      // ---- 4. out-proj + residual + LayerNorm1 -> U -------------------------------------------
      {
         const float *w = RES ? wbuf : ( stage( P::AO, P::F1 - P::AO ) - P::AO );
         if ( live )
         {
            constexpr int NB = 16;
#pragma unroll 1
            for ( int ob = 0; ob < C; ob += NB )
            {
               float acc[NB];
#pragma unroll
               for ( int o = 0; o < NB; ++o ) acc[o] = w[P::AOB + ob + o];
               lin_acc<C, NB>( myrow + OFF_O, w + P::AO + ob * C, C, acc );
               if ( tap == TAP_ATTENTION )
               {
#pragma unroll
                  for ( int o = 0; o < NB; ++o ) myrow[OFF_U + ob + o] = acc[o];
               }
               else
               {
#pragma unroll
                  for ( int o = 0; o < NB; ++o ) myrow[OFF_U + ob + o] += acc[o];
               }
            }
            if ( tap != TAP_ATTENTION ) layer_norm_row<C>( myrow + OFF_U, w + P::LN1W, w + P::LN1B );
         }
      }
      if ( tap == TAP_ATTENTION || tap == TAP_NORM1 )
      {
         tap_row( myrow + OFF_U );
         continue;
      }
      // ---- 5. FFN: linear1 + ReLU -> O ; linear2 + residual + LayerNorm2 -> U ------------------
      {
         const float *w = RES ? wbuf : ( stage( P::F1, P::F2 - P::F1 ) - P::F1 );
         if ( live )
         {
            constexpr int NB = 16;
#pragma unroll 1
            for ( int ob = 0; ob < C; ob += NB )
            {
               float acc[NB];
#pragma unroll
               for ( int o = 0; o < NB; ++o ) acc[o] = w[P::F1B + ob + o];
               lin_acc<C, NB>( myrow + OFF_U, w + P::F1 + ob * C, C, acc );
#pragma unroll
               for ( int o = 0; o < NB; ++o ) myrow[OFF_O + ob + o] = fmaxf( acc[o], 0.0f );
            }
         }
      }
      {
         const float *w = RES ? wbuf : ( stage( P::F2, P::CV - P::F2 ) - P::F2 );
         if ( live )
         {
            constexpr int NB = 16;
#pragma unroll 1
            for ( int ob = 0; ob < C; ob += NB )
            {
               float acc[NB];
#pragma unroll
               for ( int o = 0; o < NB; ++o ) acc[o] = w[P::F2B + ob + o];
               lin_acc<C, NB>( myrow + OFF_O, w + P::F2 + ob * C, C, acc );
#pragma unroll
               for ( int o = 0; o < NB; ++o ) myrow[OFF_U + ob + o] += acc[o];
            }
            layer_norm_row<C>( myrow + OFF_U, w + P::LN2W, w + P::LN2B );
         }
      }
      } // entry != ENTRY_CONV
      if ( tap == TAP_BLOCK )
      {
         tap_row( myrow + OFF_U );
         continue;
      }
      // ---- 6. conv 1x1 (stride) + BatchNorm(eval) + ReLU -> global -----------------------------
      {
         const float *w = RES ? wbuf : ( stage( P::CV, P::TOTAL - P::CV ) - P::CV );
         if ( live && ( t % Cfg::STRIDE ) == 0 )
         {
            float *o_row = out + ( (size_t)( chunk0 + g ) * Cfg::TOUT + t / Cfg::STRIDE ) * C;
            constexpr int NB = 16;
#pragma unroll 1
            for ( int ob = 0; ob < C; ob += NB )
            {
               float acc[NB];
#pragma unroll
               for ( int o = 0; o < NB; ++o ) acc[o] = 0.0f;
               lin_acc<C, NB>( myrow + OFF_U, w + P::CV + ob * C, C, acc );
               float r[NB];
#pragma unroll
               for ( int o = 0; o < NB; ++o )
               {
                  float z = acc[o] + w[P::CVB + ob + o];
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

// vadc_b200/csrc/exact_lstm_kernel.cuh -- the decoder LSTM in the reference's rounding sequence, for ANY number of streams.
//
// Replaces lstm_tensor_minibatched / lstm_seq / lstm / lstm_cell (lstm.c:31-341): per step and layer z = W [x;h] + b through
// dotproduct_simd (maths.h:123-158: 16 taps per block into eight lanes in _mm256_hadd_ps order, lanes summed left to right),
// i,f,o = 1/(1+expf(-z)), g = tanhf(z), c = f c + i g, h = o tanhf(c) (lstm.c:64-88) -- separate multiplies and adds, glibc's
// expf / tanhf bit for bit (libm_exact.cuh). The result is bit-identical to the reference for streams of any length.
//
// The recurrence is serial per stream, but streams are independent: a CTA owns a SET of streams (stream s -> CTA s mod grid) and walks
// them step by step together:
//   * the layer's weight matrix lives in REGISTERS for the whole launch: 512 threads, thread = (gate row r, half): the eight taps
//     16b + 8 half .. + 7 of every 16-tap block b of row r, i.e. exactly the taps that feed four of dotproduct_simd's eight lanes
//     (half 0: r0 r1 r4 r5, half 1: r2 r3 r6 r7) -- 64 weights per thread, loaded once;
//   * [x|h] of every stream of the CTA lives in shared memory; a thread reads its 64 taps of one stream as 16 LDS.128 that the
//     warp shares as two broadcasts, and owes 64 FMUL + 64 FADD for them: no weight traffic at all in the step loop;
//   * streams go through in pairs: the two lanes of a row each contract their half for both streams, then swap four partial sums
//     by one shuffle each, so that the even lane finishes the row for the first stream and the odd lane for the second (the lane
//     sums in the reference's order, bias, the gate's own nonlinearity): no lane idles in the libm restatements;
//   * the gate nonlinearities and the cell update run on all 512 threads over (unit, stream) pairs of a sub-batch of 8 streams
//     through a double-buffered tile of pre-activations in shared memory: every thread evaluates the same three sigmoids and two
//     tanh (no warp is the slow one at the barrier, as the tanh rows were when each row's owner applied its own nonlinearity);
//     one barrier per sub-batch and step.
// Shared memory per stream: two x slots + h 768 B + c 256 B, so one CTA carries up to XL_MAX_STREAMS streams per pass (more go in passes).
//   x: [S][steps][64] layer input (stream-major), hseq: [S][steps][64] layer output, state_h/state_c: [S][2][64].
#pragma once
#include "common.cuh"
#include "libm_exact.cuh"

#define XL_THREADS 512
#define XL_SUB 8                         // streams per sub-batch (one (unit, stream) pair per thread in the cell update)
#define XL_MAX_STREAMS 192               // streams a CTA carries in one pass
#define XL_ACT_FLOATS ( 2 * XL_SUB * 256 ) // double-buffered activation tile
#define XL_ROW 192                       // per stream: x of this step | x of the next step (cp.async target) | h -- the two x slots alternate
#define XL_SMEM_BYTES ( ( XL_ACT_FLOATS + XL_MAX_STREAMS * ( XL_ROW + 64 ) ) * 4 )

#ifndef XL_PACKED_MUL
#define XL_PACKED_MUL 1 // 1: FMUL2 products + scalar adds per stream; 0: scalar products + FADD2 across the stream pair
#endif
// four of dotproduct_simd's eight lanes over this thread's taps of one stream: acc[j] += x[2j] w[2j] + x[2j+1] w[2j+1] per block
__device__ __forceinline__ void xl_half_row( const float *__restrict__ xp /* x: this thread's first tap */, const float *__restrict__ hp /* h: likewise */,
                                             const float ( &w )[64], float ( &acc )[4] )
{
   acc[0] = acc[1] = acc[2] = acc[3] = 0.0f;
#pragma unroll
   for ( int b = 0; b < 8; ++b )
   {
      const float *src = b < 4 ? xp + 16 * b : hp + 16 * ( b - 4 );
      const float4 u = ld4( src ), v = ld4( src + 4 );
#if XL_PACKED_MUL
      // the two products of a lane as one FMUL2 (mul.rn.f32x2 on the register pairs of the LDS.128 and of the weights), scalar adds
      float p[8];
      unpk2( mul2( pk2( u.x, u.y ), pk2( w[8 * b + 0], w[8 * b + 1] ) ), p[0], p[1] );
      unpk2( mul2( pk2( u.z, u.w ), pk2( w[8 * b + 2], w[8 * b + 3] ) ), p[2], p[3] );
      unpk2( mul2( pk2( v.x, v.y ), pk2( w[8 * b + 4], w[8 * b + 5] ) ), p[4], p[5] );
      unpk2( mul2( pk2( v.z, v.w ), pk2( w[8 * b + 6], w[8 * b + 7] ) ), p[6], p[7] );
      acc[0] = __fadd_rn( acc[0], __fadd_rn( p[0], p[1] ) );
      acc[1] = __fadd_rn( acc[1], __fadd_rn( p[2], p[3] ) );
      acc[2] = __fadd_rn( acc[2], __fadd_rn( p[4], p[5] ) );
      acc[3] = __fadd_rn( acc[3], __fadd_rn( p[6], p[7] ) );
#else
      acc[0] = __fadd_rn( acc[0], __fadd_rn( __fmul_rn( u.x, w[8 * b + 0] ), __fmul_rn( u.y, w[8 * b + 1] ) ) );
      acc[1] = __fadd_rn( acc[1], __fadd_rn( __fmul_rn( u.z, w[8 * b + 2] ), __fmul_rn( u.w, w[8 * b + 3] ) ) );
      acc[2] = __fadd_rn( acc[2], __fadd_rn( __fmul_rn( v.x, w[8 * b + 4] ), __fmul_rn( v.y, w[8 * b + 5] ) ) );
      acc[3] = __fadd_rn( acc[3], __fadd_rn( __fmul_rn( v.z, w[8 * b + 6] ), __fmul_rn( v.w, w[8 * b + 7] ) ) );
#endif
   }
}

// the same for two streams that share the weights: scalar products into adjacent registers, every addition one packed FADD2
// (add.rn.f32x2: two independently rounded IEEE additions in one issue slot) -- 28 instead of 36 issue slots per 16-tap block and stream pair
__device__ __forceinline__ void xl_half_row_pair( const float *__restrict__ xa, const float *__restrict__ ha, const float *__restrict__ xb, const float *__restrict__ hb,
                                                  const float ( &w )[64], float ( &acc_a )[4], float ( &acc_b )[4] )
{
   f32x2 acc[4];
   acc[0] = acc[1] = acc[2] = acc[3] = pk2( 0.0f, 0.0f );
#pragma unroll
   for ( int b = 0; b < 8; ++b )
   {
      const float *sa = b < 4 ? xa + 16 * b : ha + 16 * ( b - 4 ), *sb = b < 4 ? xb + 16 * b : hb + 16 * ( b - 4 );
      const float4 ua = ld4( sa ), va = ld4( sa + 4 ), ub = ld4( sb ), vb = ld4( sb + 4 );
      acc[0] = add2( acc[0], add2( pk2( __fmul_rn( ua.x, w[8 * b + 0] ), __fmul_rn( ub.x, w[8 * b + 0] ) ), pk2( __fmul_rn( ua.y, w[8 * b + 1] ), __fmul_rn( ub.y, w[8 * b + 1] ) ) ) );
      acc[1] = add2( acc[1], add2( pk2( __fmul_rn( ua.z, w[8 * b + 2] ), __fmul_rn( ub.z, w[8 * b + 2] ) ), pk2( __fmul_rn( ua.w, w[8 * b + 3] ), __fmul_rn( ub.w, w[8 * b + 3] ) ) ) );
      acc[2] = add2( acc[2], add2( pk2( __fmul_rn( va.x, w[8 * b + 4] ), __fmul_rn( vb.x, w[8 * b + 4] ) ), pk2( __fmul_rn( va.y, w[8 * b + 5] ), __fmul_rn( vb.y, w[8 * b + 5] ) ) ) );
      acc[3] = add2( acc[3], add2( pk2( __fmul_rn( va.z, w[8 * b + 6] ), __fmul_rn( vb.z, w[8 * b + 6] ) ), pk2( __fmul_rn( va.w, w[8 * b + 7] ), __fmul_rn( vb.w, w[8 * b + 7] ) ) ) );
   }
#pragma unroll
   for ( int i = 0; i < 4; ++i ) unpk2( acc[i], acc_a[i], acc_b[i] );
}

__device__ __forceinline__ int xl_ld_acquire( const int *p )
{
   int v;
   asm volatile( "ld.acquire.gpu.global.s32 %0, [%1];" : "=r"( v ) : "l"( p ) : "memory" );
   return v;
}
__device__ __forceinline__ void xl_st_release( int *p, int v )
{
   asm volatile( "st.release.gpu.global.s32 [%0], %1;" ::"l"( p ), "r"( v ) : "memory" );
}

// WAVE = false: one layer per launch (layer_arg), CTA b owns streams b, b + grid, ...  -- any number of streams.
// WAVE = true:  BOTH layers in one launch as a wavefront, for stream counts that leave SMs free (2 x groups <= SMs): the grid is
//   2 G CTAs; a CTA draws a ticket (2 g = layer 0 of stream group g, 2 g + 1 = layer 1 of the same group: whoever holds a layer-1
//   ticket knows its layer-0 partner has already been STARTED -- no co-residency assumption, no deadlock whatever else shares the
//   GPU); the layer-1 CTA consumes the layer-0 CTA's output sequence while it is being produced, a step or two behind, so a step of
//   both layers costs what a step of one does. Hand-off per group through progress[g] = steps the layer-0 CTA has completed
//   (st.release.gpu after a CTA barrier / ld.acquire.gpu + ld.global.cg). The consumer's wait is bounded: if the producer is lost the
//   consumer raises err_word (mapped host memory; the host fails the call at its next synchronization) and leaves the state untouched.
//   Layer 1 may write its output over the encoder output (x0 == out1): it writes row t after layer 0 has completed step t, and by
//   then layer 0 has read rows 0..t+1 for good (rows are 256 bytes, 256-byte aligned).
//   sync: [0] ticket counter, [1 + g] progress of group g; zeroed by the host before the launch.
template <bool WAVE>
__global__ void __launch_bounds__( XL_THREADS, 1 )
exact_lstm_kernel( const float *x0, float *h0seq, float *out1, float *__restrict__ state_h, float *__restrict__ state_c,
                   const float *__restrict__ wpack /*[2][32][256][4]*/, const float *__restrict__ bias /*[2][256]*/, int nstreams, int nw, int layer_arg,
                   int *sync, int *err_word, int spin_limit, int debug_stall_producer )
{
   extern __shared__ __align__( 16 ) float xsm[];
   __shared__ unsigned long long exp_tab[32];
   __shared__ int s_ticket;
   lme::stage_exp2f_tab( exp_tab, threadIdx.x );
   float *act = xsm;                                 // [2][XL_SUB][256] pre-activations z = W [x;h] + b
   float *xh = xsm + XL_ACT_FLOATS;                  // [K][XL_ROW]: x slot 0 | x slot 1 | h
   float *cst = xh + XL_MAX_STREAMS * XL_ROW;        // [K][64]
   const int tid = threadIdx.x, half = tid & 1, row = tid >> 1;
   const int steps = nw * 7;
   int grp = blockIdx.x, ngrp = gridDim.x, layer = layer_arg;
   if ( WAVE )
   {
      if ( tid == 0 ) s_ticket = atomicAdd( sync, 1 );
      __syncthreads();
      grp = s_ticket >> 1;
      layer = s_ticket & 1;
      ngrp = gridDim.x >> 1;
   }
   const bool consumer = WAVE && layer == 1;
   const float *x = layer == 0 ? x0 : h0seq;
   float *hseq = layer == 0 ? h0seq : out1;
   int *progress = WAVE ? sync + 1 + grp : 0;

   // this thread's 64 weights: quads 4b + 2 half, 4b + 2 half + 1 of row `row`
   float w[64];
   {
      const float4 *src = reinterpret_cast<const float4 *>( wpack ) + (size_t)layer * ( 32 * 256 );
#pragma unroll
      for ( int b = 0; b < 8; ++b )
      {
         const float4 q0 = __ldg( src + ( 4 * b + 2 * half ) * 256 + row ), q1 = __ldg( src + ( 4 * b + 2 * half + 1 ) * 256 + row );
         w[8 * b + 0] = q0.x; w[8 * b + 1] = q0.y; w[8 * b + 2] = q0.z; w[8 * b + 3] = q0.w;
         w[8 * b + 4] = q1.x; w[8 * b + 5] = q1.y; w[8 * b + 6] = q1.z; w[8 * b + 7] = q1.w;
      }
   }
   const float b_row = bias[layer * 256 + row];
   // cell-update role: (unit uj, stream uk of the sub-batch)
   const int uk = tid >> 6, uj = tid & 63;
   int seen = 0;
   bool lost = false;
   // the consumer of a wavefront waits until its producer has completed step t before it touches input row t
   auto wait_row = [&]( size_t s, int t ) {
      if ( !consumer || seen > t || lost ) return;
      int spins = 0;
      do
      {
         seen = xl_ld_acquire( progress );
      } while ( seen <= t && ++spins < spin_limit );
      if ( seen <= t )
      {
         lost = true;
         atomicExch( err_word, 1 + (int)s );
      }
   };
   // input row t of stream s (four floats from column uj on) -> shared memory, asynchronously: the row is needed a whole step later, and
   // a register-held load was sunk by the compiler to its use, where its full latency sat on the critical path of the step
   // (measured on one stream: 2 800 of 4 500 cycles per step). cp.async.cg: L2 only, as the wavefront's hand-over needs.
   auto copy_row = [&]( float *dst, size_t s, int t ) {
      const float *p = x + ( s * steps + t ) * 64 + uj;
      asm volatile( "cp.async.cg.shared.global [%0], [%1], 16;" ::"r"( (unsigned)__cvta_generic_to_shared( dst ) ), "l"( p ) : "memory" );
   };

   // streams of this CTA: grp + i * ngrp, in passes of at most XL_MAX_STREAMS (a wavefront launch always fits one pass)
   const int mine = ( nstreams - grp + ngrp - 1 ) / ngrp;
   for ( int pass0 = 0; pass0 < mine; pass0 += XL_MAX_STREAMS )
   {
      const int K = min( XL_MAX_STREAMS, mine - pass0 );
      __syncthreads(); // the previous pass is done with the shared buffers
      for ( int e = tid; e < K * 64; e += XL_THREADS )
      {
         const int k = e >> 6, j = e & 63; // (j == uj: XL_THREADS is a multiple of 64)
         const size_t s = (size_t)grp + (size_t)( pass0 + k ) * ngrp;
         cst[k * 64 + j] = state_c[( s * 2 + layer ) * 64 + j];
         xh[k * XL_ROW + 128 + j] = state_h[( s * 2 + layer ) * 64 + j];
         if ( ( j & 3 ) == 0 )
         {
            wait_row( s, 0 );
            copy_row( xh + k * XL_ROW + j, s, 0 ); // step 0 reads x slot 0
         }
      }
      asm volatile( "cp.async.wait_all;" ::: "memory" );
      __syncthreads();
      const int nsub = ( K + XL_SUB - 1 ) / XL_SUB;
      int buf = 0;
#ifdef XL_TIMING // -DXL_TIMING: where the cycles of a step go, seen from the first row owner of the tanh gate (debug builds only)
      long long tq[6] = { 0, 0, 0, 0, 0, 0 }, tl = clock64();
#define XL_T( i ) { const long long now = clock64(); tq[i] += now - tl; tl = now; }
#else
#define XL_T( i )
#endif
      for ( int step = 0; step < steps; ++step )
      {
         for ( int sb = 0; sb < nsub; ++sb, buf ^= 1 )
         {
            const int k0 = sb * XL_SUB, kn = min( XL_SUB, K - k0 );
            // next step's input of my (unit, stream) goes to the other x slot while this step computes
            const bool upd = uk < kn;
            const size_t us = (size_t)grp + (size_t)( pass0 + k0 + uk ) * ngrp;
            const int xs = ( step & 1 ) * 64, xn = 64 - xs; // x slot of this step / of the next
            const bool copying = upd && ( uj & 3 ) == 0 && step + 1 < steps;
            if ( copying )
            {
               wait_row( us, step + 1 );
               copy_row( xh + ( k0 + uk ) * XL_ROW + xn + uj, us, step + 1 );
            }
            XL_T( 0 ) // copy issue (+ the consumer's wait)
            float *a = act + buf * ( XL_SUB * 256 );
            const bool spread = kn <= 4; // (measured: up to four streams the row owners' 1-2 evaluations + the cell's tanh beat five evaluations in the cell-update phase)
#pragma unroll 1
            for ( int p = 0; p < kn; p += 2 )
            {
               const bool two = p + 1 < kn;
               float a0[4], a1[4];
               if ( two && !XL_PACKED_MUL )
                  xl_half_row_pair( xh + ( k0 + p ) * XL_ROW + xs + 8 * half, xh + ( k0 + p ) * XL_ROW + 128 + 8 * half,
                                    xh + ( k0 + p + 1 ) * XL_ROW + xs + 8 * half, xh + ( k0 + p + 1 ) * XL_ROW + 128 + 8 * half, w, a0, a1 );
               else
               {
                  xl_half_row( xh + ( k0 + p ) * XL_ROW + xs + 8 * half, xh + ( k0 + p ) * XL_ROW + 128 + 8 * half, w, a0 );
                  if ( two )
                     xl_half_row( xh + ( k0 + p + 1 ) * XL_ROW + xs + 8 * half, xh + ( k0 + p + 1 ) * XL_ROW + 128 + 8 * half, w, a1 );
                  else
                     a1[0] = a1[1] = a1[2] = a1[3] = 0.0f;
               }
               // the even lane finishes stream p, the odd lane stream p + 1: swap the other stream's partial sums
               float lo[4], hi[4];
#pragma unroll
               for ( int i = 0; i < 4; ++i )
               {
                  const float got = __shfl_xor_sync( 0xffffffffu, half ? a0[i] : a1[i], 1 );
                  lo[i] = half ? got : a0[i];   // lanes r0 r1 r4 r5 of my stream
                  hi[i] = half ? a1[i] : got;   // lanes r2 r3 r6 r7
               }
               XL_T( 1 ) // contraction + exchange
               float z = 0.0f;
               z = __fadd_rn( z, lo[0] );
               z = __fadd_rn( z, lo[1] );
               z = __fadd_rn( z, hi[0] );
               z = __fadd_rn( z, hi[1] );
               z = __fadd_rn( z, lo[2] );
               z = __fadd_rn( z, lo[3] );
               z = __fadd_rn( z, hi[2] );
               z = __fadd_rn( z, hi[3] );
               z = __fadd_rn( z, b_row );
               // Up to four streams in the sub-batch (few streams per CTA: the latency-bound shapes, down to ONE stream): the lane that
               // finished a row applies the row's own gate nonlinearity right here (a warp's 16 rows belong to one gate: no divergence),
               // so that the serial chain of a step is one nonlinearity + the cell's tanh. With more streams the cell-update phase takes
               // all five evaluations: there every thread does the same work and no warp is the slow one at the barrier.
               if ( !half || two ) a[( p + half ) * 256 + row] = !spread ? z : ( ( row >> 6 ) == 2 ? lme::tanhf_ref( z ) : lme::sigmoid_ref( z, exp_tab ) );
            }
            XL_T( 2 ) // lane sum + the row's nonlinearity
            __syncthreads();
            XL_T( 3 ) // barrier
            // gate nonlinearities and cell update (lstm.c:64-88) of (unit uj, stream k0 + uk)
            if ( upd )
            {
               const float *av = a + uk * 256;
               float ig = av[uj], fg = av[64 + uj], gg = av[128 + uj], og = av[192 + uj];
               if ( !spread )
               {
                  lme::sigmoid3_ref( ig, fg, og, exp_tab );
                  gg = lme::tanhf_ref( gg );
               }
               const int k = k0 + uk;
               const float cn = __fadd_rn( __fmul_rn( fg, cst[k * 64 + uj] ), __fmul_rn( ig, gg ) );
               const float hn = __fmul_rn( lme::tanhf_ref( cn ), og );
               cst[k * 64 + uj] = cn;
               xh[k * XL_ROW + 128 + uj] = hn;
               hseq[( us * steps + step ) * 64 + uj] = hn;
            }
            if ( copying ) asm volatile( "cp.async.wait_all;" ::: "memory" ); // my copy has landed before the barrier that releases the next step
            // with a single sub-batch the next contraction reads what this update wrote; a wavefront producer publishes the step
            // after every thread's h of it has been stored (the barrier orders the stores before thread 0's release)
            const bool publish = WAVE && layer == 0 && sb == nsub - 1;
            XL_T( 4 ) // cell update (threads of the first two warps only)
            // (Tried, r02: with one stream per CTA, the x half of every row contracted a step ahead -- by the update warps after their
            // sigmoid, by everybody else during the cell update, the same additions in the same order: 18.9 ms per 10 752 steps against
            // 16.6 ms; the early work competes with the update, which IS the critical chain.)
            // (Tried, r02: the release and the consumer's acquire moved to warps that idle through the cell update, or into the slack
            // before the first barrier -- one stream 17.0 / 18.7 ms per 10 752 steps against 16.6 ms here: the release is a ~1000-cycle
            // fence wherever it sits, and a later publication only makes the consumer wait for it.)
            if ( nsub == 1 || publish ) __syncthreads();
            XL_T( 5 ) // barrier
            if ( publish && tid == 0 && !debug_stall_producer ) xl_st_release( progress, step + 1 );
         }
      }
#ifdef XL_TIMING
      if ( ( tid == 256 || tid == 0 ) && blockIdx.x < 2 )
         printf( "XL_TIMING cta %d layer %d tid %d steps %d: copy/wait %lld contraction %lld nonlinearity %lld barrier %lld update %lld barrier %lld cycles per step\n", blockIdx.x, layer, tid,
                 steps, tq[0] / steps, tq[1] / steps, tq[2] / steps, tq[3] / steps, tq[4] / steps, tq[5] / steps );
#endif
      const int any_lost = __syncthreads_or( lost ? 1 : 0 );
      if ( !any_lost )
         for ( int e = tid; e < K * 64; e += XL_THREADS )
         {
            const int k = e >> 6, j = e & 63;
            const size_t s = (size_t)grp + (size_t)( pass0 + k ) * ngrp;
            state_c[( s * 2 + layer ) * 64 + j] = cst[k * 64 + j];
            state_h[( s * 2 + layer ) * 64 + j] = xh[k * XL_ROW + 128 + j];
         }
   }
}

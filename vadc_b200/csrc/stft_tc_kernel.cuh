// vadc_b200/csrc/stft_tc_kernel.cuh -- STFT + magnitude + log on the 5th-gen tensor cores.
//
// Same function and the same hybrid rule as stft_hybrid_kernel.cuh (my_stft stft.c:15-229 + the log1p of
// adaptive_audio_normalization_inplace misc.c:40-46): every bin is evaluated fast,
// and every bin whose magnitude is below k_rel * ||windowed frame|| is re-evaluated with the reference's
// own rounding sequence (stft.c:108-184) and is bit-identical to it. Here "fast" is the reference's
// own formulation -- the 258x256 conv-basis correlation -- as a tensor-core GEMM:
//     Y[frame][n] = sum_k xpad[64*frame + k] * B[n][k],   M = 128 frames (4 chunks x 32 rows, 25 live),
//                                                         N = 256 (rows 129 and 257 of the basis are zero), K = 256
// with the fp16x2 split (3 products, fp32 accumulation in TMEM). s16 samples / 32768 split EXACTLY into
// fp16 hi + lo (15 significant bits), so only the basis is rounded (22 bits).
//
// im2col-free A operand: frame t needs padded samples [64t, 64t+256) = 64-sample blocks t..t+3. The chunk is
// stored once per tile as [k-chunk c of a block (8)][block slot (32 per chunk)][8 samples]; in the K-major
// no-swizzle operand layout rows are 16 bytes apart inside a k-chunk, so "row r of K-slice ks" is simply
// block slot r + ks/2: the eight K-slices (32 samples each) of ALL 128 frames are eight descriptor offsets
// into the same 34 KB image -- every sample is converted and stored once instead of four times.
//
// The basis image (256 KB as fp16 hi/lo) does not fit in shared memory: a producer thread streams it in
// eight 32 KB K-slices per tile with cp.async.bulk (TMA engine, mbarrier complete_tx) through a 4-stage ring
// (the image stays L2-resident). Roles: warps 0..7 convert PCM -> A image of the NEXT tile, then run the
// epilogue of the current one (tcgen05.ld -> magnitude -> Parseval energy -> threshold -> log -> coalesced
// stores, exact fix-up of flagged bins by the whole warp); warp 8 issues the MMAs; warp 9 is the producer.
// Two 256-column TMEM accumulators alternate, so the epilogue of tile i overlaps the MMAs of tile i+1.
// Column order: n = 2f -> Re Y_f (basis row f), n = 2f+1 -> Im Y_f (row 129+f) for f = 1..127; n = 0 -> Re Y_0,
// n = 1 -> Re Y_128 (Im of both is identically zero).
#pragma once
#include "common.cuh"
#include "stft_hybrid_kernel.cuh" // hyb_sqrt_fast, hyb_log1p_scaled
#include "tc_common.cuh"

#define STC_WORKER_WARPS 16                      // 4 TMEM lane quarters x 4 column groups of 64 (32 bins per thread)
#define STC_NCG ( STC_WORKER_WARPS / 4 )
#define STC_BPT ( 128 / STC_NCG )                 // bins per thread
#define STC_THREADS ( ( STC_WORKER_WARPS + 2 ) * 32 )
#define STC_NSTAGE 4
#define STC_B_LBO ( 256 * 16 )                    // bytes between k-chunks of a basis slice
#define STC_B_SPLIT ( 4 * STC_B_LBO )             // one split of a K=32 slice
#define STC_B_SLICE ( 2 * STC_B_SPLIT )           // 32 KB
#define STC_B_IMAGE ( 8 * STC_B_SLICE )           // 256 KB in global memory
#define STC_A_SLOTS 136                           // 4 chunks x 32 block slots + 3 (K-slice shift) rounded up to 8
#define STC_A_LBO ( STC_A_SLOTS * 16 )
#define STC_A_SPLIT ( 8 * STC_A_LBO )
#define STC_A_IMAGE ( 2 * STC_A_SPLIT )           // 34 816 B
#define STC_SMEM_BYTES ( STC_NSTAGE * STC_B_SLICE + 2 * STC_A_IMAGE + 2 * STC_NCG * 128 * 4 + 256 )

namespace tc
{
__device__ __forceinline__ void mbar_arrive_expect_tx( uint64_t *bar, uint32_t bytes )
{
   asm volatile( "mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"( smem_u32( bar ) ), "r"( bytes ) : "memory" );
}
// 1-D bulk copy global -> shared through the TMA engine; completion is counted in bytes on `bar`
__device__ __forceinline__ void bulk_g2s( void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar )
{
   asm volatile( "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"( smem_u32( smem_dst ) ), "l"( gmem_src ),
                 "r"( bytes ), "r"( smem_u32( bar ) )
                 : "memory" );
}
} // namespace tc

// raw sample index of padded position p (tensor.h:942-953: reflect pad 128 without repeating the edge sample)
__device__ __forceinline__ int stc_raw_index( int p ) { return p < 128 ? 128 - p : ( p < 1664 ? p - 128 : 3198 - p ); }

template <bool F32>
__device__ __forceinline__ float stc_sample( const void *chunk, int m )
{
   if ( F32 ) return __ldg( reinterpret_cast<const float *>( chunk ) + m );
   return (float)__ldg( reinterpret_cast<const short *>( chunk ) + m ) * ( 1.0f / 32768.0f ); // vadc.c:884,898
}

// the reference's 256-tap tree (stft.c:108-184) for basis row `row` at frame t, evaluated by one warp from the
// caller's input (lane = l*4 + g owns one 8-tap leaf). Same value in every lane.
template <bool F32>
__device__ __forceinline__ float stc_exact_row( const void *chunk, const float *__restrict__ basis, int row, int t, int lane )
{
   const int l = lane >> 2, g = lane & 3;
   const int p0 = 64 * t + 64 * g + l;
   const float *bp = basis + (size_t)row * 256 + 64 * g + l;
   float p[8];
#pragma unroll
   for ( int v = 0; v < 8; ++v ) p[v] = __fmul_rn( stc_sample<F32>( chunk, stc_raw_index( p0 + 8 * v ) ), __ldg( bp + 8 * v ) );
   float s01 = __fadd_rn( p[0], p[1] ), s23 = __fadd_rn( p[2], p[3] ), s45 = __fadd_rn( p[4], p[5] ), s67 = __fadd_rn( p[6], p[7] );
   float r = __fadd_rn( __fadd_rn( s01, s23 ), __fadd_rn( s45, s67 ) );
   r = __fadd_rn( r, __shfl_xor_sync( 0xffffffffu, r, 1 ) );
   r = __fadd_rn( r, __shfl_xor_sync( 0xffffffffu, r, 2 ) );
   r = __fadd_rn( r, __shfl_xor_sync( 0xffffffffu, r, 4 ) );
   r = __fadd_rn( r, __shfl_xor_sync( 0xffffffffu, r, 8 ) );
   r = __fadd_rn( r, __shfl_xor_sync( 0xffffffffu, r, 16 ) );
   return r;
}

// exact magnitude of bin f at frame t (stft.c:108-213): both basis rows share the gathered samples
template <bool F32>
__device__ __forceinline__ float stc_exact_mag( const void *chunk, const float *__restrict__ basis, int f, int t, int lane )
{
   const int l = lane >> 2, g = lane & 3;
   const int p0 = 64 * t + 64 * g + l;
   const float *br = basis + (size_t)f * 256 + 64 * g + l, *bi = br + (size_t)129 * 256;
   float pr[8], pi[8];
#pragma unroll
   for ( int v = 0; v < 8; ++v )
   {
      const float x = stc_sample<F32>( chunk, stc_raw_index( p0 + 8 * v ) );
      pr[v] = __fmul_rn( x, __ldg( br + 8 * v ) );
      pi[v] = __fmul_rn( x, __ldg( bi + 8 * v ) );
   }
   float re = __fadd_rn( __fadd_rn( __fadd_rn( pr[0], pr[1] ), __fadd_rn( pr[2], pr[3] ) ), __fadd_rn( __fadd_rn( pr[4], pr[5] ), __fadd_rn( pr[6], pr[7] ) ) );
   float im = __fadd_rn( __fadd_rn( __fadd_rn( pi[0], pi[1] ), __fadd_rn( pi[2], pi[3] ) ), __fadd_rn( __fadd_rn( pi[4], pi[5] ), __fadd_rn( pi[6], pi[7] ) ) );
#pragma unroll
   for ( int off = 1; off < 32; off <<= 1 )
   {
      re = __fadd_rn( re, __shfl_xor_sync( 0xffffffffu, re, off ) );
      im = __fadd_rn( im, __shfl_xor_sync( 0xffffffffu, im, off ) );
   }
   return sqrtf( __fadd_rn( __fmul_rn( re, re ), __fmul_rn( im, im ) ) );
}

template <bool F32>
__global__ void __launch_bounds__( STC_THREADS, 1 )
stft_tc_kernel( const void *__restrict__ in, long long stream_stride, int nw, int nchunks, const unsigned char *__restrict__ bimg /*STC_B_IMAGE*/,
                const float *__restrict__ basis /*[258][256] fp32, exact path*/, float *__restrict__ spec, float k_rel, int out_mode,
                unsigned long long *__restrict__ flagged, unsigned long long *__restrict__ fix_list, unsigned int *__restrict__ fix_count, unsigned int fix_cap )
{
   extern __shared__ __align__( 128 ) unsigned char stc_smem[];
   unsigned char *sB = stc_smem;                                        // [STC_NSTAGE][STC_B_SLICE]
   unsigned char *sA = sB + STC_NSTAGE * STC_B_SLICE;                   // [2][STC_A_IMAGE]
   float *sE = reinterpret_cast<float *>( sA + 2 * STC_A_IMAGE );       // [2 tiles][STC_NCG column groups][128 rows] partial frame energies
   uint64_t *bars = reinterpret_cast<uint64_t *>( sE + 2 * STC_NCG * 128 );
   uint64_t *full_b = bars, *empty_b = bars + STC_NSTAGE, *a_ready = bars + 2 * STC_NSTAGE, *acc_full = a_ready + 2, *acc_empty = acc_full + 2;
   uint32_t *tmem_slot = reinterpret_cast<uint32_t *>( acc_empty + 2 );

   const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
   const int ntiles = ( nchunks + 3 ) / 4;

   if ( warp == STC_WORKER_WARPS )
   {
      tc::tmem_alloc( tmem_slot, 512 );
      if ( lane == 0 )
      {
         for ( int i = 0; i < STC_NSTAGE; ++i )
         {
            tc::mbar_init( &full_b[i], 1 );
            tc::mbar_init( &empty_b[i], 1 );
         }
         for ( int i = 0; i < 2; ++i )
         {
            tc::mbar_init( &a_ready[i], STC_WORKER_WARPS );
            tc::mbar_init( &acc_full[i], 1 );
            tc::mbar_init( &acc_empty[i], STC_WORKER_WARPS );
         }
         tc::mbar_fence_init();
      }
   }
   tc::fence_before_sync();
   __syncthreads();
   tc::fence_after_sync();
   const uint32_t tmem = *tmem_slot;

   auto chunk_ptr = [&]( int ci ) -> const void * {
      const int s = ci / nw, n = ci - s * nw;
      const long long off = (long long)s * stream_stride + (long long)n * VB_CHUNK;
      return F32 ? (const void *)( reinterpret_cast<const float *>( in ) + off ) : (const void *)( reinterpret_cast<const int16_t *>( in ) + off );
   };

   if ( warp == STC_WORKER_WARPS + 1 )
   {
      // ================================ basis producer ===============================================
      if ( lane == 0 )
      {
         uint32_t cnt = 0;
         for ( int tile = blockIdx.x; tile < ntiles; tile += gridDim.x )
            for ( int ks = 0; ks < 8; ++ks, ++cnt )
            {
               const uint32_t st = cnt % STC_NSTAGE, use = cnt / STC_NSTAGE;
               if ( use > 0 ) tc::mbar_wait( &empty_b[st], ( use - 1u ) & 1u );
               tc::mbar_arrive_expect_tx( &full_b[st], STC_B_SLICE );
               tc::bulk_g2s( sB + st * STC_B_SLICE, bimg + (size_t)ks * STC_B_SLICE, STC_B_SLICE, &full_b[st] );
            }
      }
   }
   else if ( warp == STC_WORKER_WARPS )
   {
      // ================================ MMA issuer ===================================================
      constexpr uint32_t idesc = tc::idesc_f16_f32( 128, 256 );
      const uint32_t a_base = tc::smem_u32( sA ), b_base = tc::smem_u32( sB );
      uint32_t cnt = 0, it = 0;
      for ( int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it )
      {
         const uint32_t buf = it & 1u, ph = ( it >> 1 ) & 1u;
         tc::mbar_wait( &a_ready[buf], ph );
         if ( it >= 2 ) tc::mbar_wait( &acc_empty[buf], ph ^ 1u );
         tc::fence_after_sync();
         const uint32_t d_tmem = tmem + buf * 256u;
         for ( int ks = 0; ks < 8; ++ks, ++cnt )
         {
            const uint32_t st = cnt % STC_NSTAGE, use = cnt / STC_NSTAGE;
            tc::mbar_wait( &full_b[st], use & 1u );
            tc::fence_after_sync();
            if ( tc::elect_one() )
            {
               // row r of K-slice ks lives at block slot r + ks/2, k-chunks 4*(ks&1) .. +3 of the block
               const uint64_t dA = tc::smem_desc( a_base + buf * STC_A_IMAGE + ( 4 * ( ks & 1 ) ) * STC_A_LBO + ( ks >> 1 ) * 16, STC_A_LBO, 128 );
               const uint64_t dB = tc::smem_desc( b_base + st * STC_B_SLICE, STC_B_LBO, 128 );
#pragma unroll
               for ( int p = 0; p < 3; ++p ) // (A split, B split): (hi,hi) (lo,hi) (hi,lo)
               {
                  const uint64_t da = dA + (uint64_t)( ( p == 1 ? STC_A_SPLIT : 0 ) >> 4 );
                  const uint64_t db = dB + (uint64_t)( ( p == 2 ? STC_B_SPLIT : 0 ) >> 4 );
#pragma unroll
                  for ( int kk = 0; kk < 2; ++kk )
                     tc::mma_bf16( d_tmem, da + (uint64_t)( ( kk * 2 * STC_A_LBO ) >> 4 ), db + (uint64_t)( ( kk * 2 * STC_B_LBO ) >> 4 ), idesc,
                                   ( ks | p | kk ) ? 1u : 0u );
               }
               tc::mma_commit( &empty_b[st] );                 // the stage is free once these MMAs have read it
               if ( ks == 7 ) tc::mma_commit( &acc_full[buf] ); // ... and the accumulator is complete
            }
            __syncwarp();
         }
      }
   }
   else
   {
      // ================================ workers: PCM -> A image, epilogue ================================
      const int wq = warp & 3, cg = warp >> 2, t = lane; // TMEM lane quarter (= chunk of the tile), column group
      unsigned nflag = 0;

      // PCM -> operand image of one tile. Per chunk 224 octets of 8 padded samples (octet po -> k-chunk po%8 of block po/8):
      // the 192 interior octets are aligned 16-byte loads (all of a thread's loads are issued before the first use), the 32
      // reflect-padded ones (tensor.h:942-953) are gathered sample by sample by the first 128 threads.
      auto write_image = [&]( int tile, unsigned char *img ) {
         constexpr int NT = STC_WORKER_WARPS * 32, NO = ( 768 + NT - 1 ) / NT; // interior octets per thread
         constexpr int NL = F32 ? 2 * NO : NO;
         int4 raw[NL];
         bool okc[NO];
#pragma unroll
         for ( int i = 0; i < NO; ++i )
         {
            const int o = tid + NT * i, q = o / 192, po = 16 + ( o - q * 192 );
            const int ci = tile * 4 + q;
            okc[i] = o < 768 && ci < nchunks;
            if ( okc[i] )
            {
               const void *cp = chunk_ptr( ci );
               if ( F32 )
               {
                  const int4 *src = reinterpret_cast<const int4 *>( reinterpret_cast<const float *>( cp ) + ( 8 * po - 128 ) );
                  raw[2 * i] = __ldg( src );
                  raw[2 * i + 1] = __ldg( src + 1 );
               }
               else
                  raw[i] = __ldg( reinterpret_cast<const int4 *>( reinterpret_cast<const short *>( cp ) + ( 8 * po - 128 ) ) );
            }
         }
         float ev[8];
         int eq = 0, epo = 0;
         if ( tid < 128 )
         {
            eq = tid >> 5;
            const int e = tid & 31;
            epo = e < 16 ? e : 192 + e; // octets 0..15 and 208..223
            const int ci = tile * 4 + eq;
#pragma unroll
            for ( int k = 0; k < 8; ++k ) ev[k] = ci < nchunks ? stc_sample<F32>( chunk_ptr( ci ), stc_raw_index( 8 * epo + k ) ) : 0.0f;
         }
#pragma unroll
         for ( int i = 0; i < NO; ++i )
         {
            const int o = tid + NT * i, q = o / 192, po = 16 + ( o - q * 192 );
            if ( o >= 768 ) break;
            float v[8];
            if ( !okc[i] )
            {
#pragma unroll
               for ( int e = 0; e < 8; ++e ) v[e] = 0.0f;
            }
            else if ( F32 )
            {
               const float *f = reinterpret_cast<const float *>( &raw[F32 ? 2 * i : 0] );
#pragma unroll
               for ( int e = 0; e < 8; ++e ) v[e] = f[e];
            }
            else
            {
               const short *h = reinterpret_cast<const short *>( &raw[i] );
#pragma unroll
               for ( int e = 0; e < 8; ++e ) v[e] = (float)h[e] * ( 1.0f / 32768.0f ); // vadc.c:884,898
            }
            unsigned char *dst = img + ( po & 7 ) * STC_A_LBO + ( 32 * q + ( po >> 3 ) ) * 16;
            tc::split_store8_f16( v, dst, dst + STC_A_SPLIT );
         }
         if ( tid < 128 )
         {
            unsigned char *dst = img + ( epo & 7 ) * STC_A_LBO + ( 32 * eq + ( epo >> 3 ) ) * 16;
            tc::split_store8_f16( ev, dst, dst + STC_A_SPLIT );
         }
         tc::fence_async_smem();
         __syncwarp();
         if ( lane == 0 ) tc::mbar_arrive( &a_ready[( img == sA ) ? 0 : 1] );
      };

      if ( (int)blockIdx.x < ntiles ) write_image( blockIdx.x, sA );
      uint32_t it = 0;
      for ( int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it )
      {
         const uint32_t buf = it & 1u, ph = ( it >> 1 ) & 1u;
         // the other image was last read by the MMAs of the previous tile, whose completion (acc_full) this warp has seen
         if ( tile + (int)gridDim.x < ntiles ) write_image( tile + gridDim.x, sA + ( buf ^ 1u ) * STC_A_IMAGE );

         tc::mbar_wait( &acc_full[buf], ph );
         tc::fence_after_sync();
         const int ci = tile * 4 + wq;
         const bool live = ( t < VB_FRAMES ) && ( ci < nchunks );
         const uint32_t taddr = tmem + ( (uint32_t)( wq * 32 ) << 16 ) + buf * 256u + cg * ( 2 * STC_BPT );

         // pass 1: magnitudes of this thread's bins (+ Nyquist in column group 0) and the frame's energy (Parseval)
         float mag[STC_BPT], mnyq = 0.0f, e2 = 0.0f;
#pragma unroll
         for ( int c0 = 0; c0 < 2 * STC_BPT; c0 += 32 )
         {
            float y[32];
            tc::tmem_ld32( taddr + c0, y );
            tc::tmem_wait_ld();
#pragma unroll
            for ( int j = 0; j < 16; ++j )
            {
               const int b = c0 / 2 + j; // bin inside the column group
               if ( b == 0 )
               {
                  if ( cg == 0 )
                  {
                     // columns 0, 1 = Re Y_0, Re Y_128
                     mnyq = fabsf( y[1] );
                     e2 += 0.5f * fmaf( y[0], y[0], y[1] * y[1] );
                     mag[0] = fabsf( y[0] );
                  }
                  else
                  {
                     const float m2 = fmaf( y[0], y[0], y[1] * y[1] );
                     e2 += m2;
                     mag[0] = hyb_sqrt_fast( m2 );
                  }
               }
               else
               {
                  const float m2 = fmaf( y[2 * j], y[2 * j], y[2 * j + 1] * y[2 * j + 1] );
                  e2 += m2;
                  mag[b] = hyb_sqrt_fast( m2 );
               }
            }
         }
         // the accumulator has been read: hand it back to the MMA warp
         tc::fence_before_sync();
         __syncwarp();
         if ( lane == 0 ) tc::mbar_arrive( &acc_empty[buf] );

         // ||windowed frame||^2 = ( |Y0|^2 + |Y128|^2 + 2 sum_{1..127} |Yf|^2 ) / 256: exchange the column groups' partial sums
         float *ex = sE + buf * ( STC_NCG * 128 );
         ex[cg * 128 + wq * 32 + lane] = e2;
         bar_sync( 1 + wq, 32 * STC_NCG );
         float etot = 0.0f;
#pragma unroll
         for ( int i = 0; i < STC_NCG; ++i ) etot += ex[i * 128 + wq * 32 + lane]; // same order in every thread of the row
         const float tau = k_rel * sqrtf( etot * ( 2.0f / 256.0f ) );

         // pass 2: threshold, log, coalesced stores ([chunk][bin][frame]: the 25 live lanes write 100 contiguous bytes)
         float *o = spec + (size_t)min( ci, nchunks - 1 ) * ( VB_BINS * VB_FRAMES ) + min( t, VB_FRAMES - 1 ) + cg * STC_BPT * VB_FRAMES;
         unsigned long long fl = 0ull;
#pragma unroll
         for ( int b = 0; b < STC_BPT; ++b )
         {
            const bool small = mag[b] < tau;
            if ( small ) fl |= 1ull << b;
            const float lv = out_mode ? mag[b] : hyb_log1p_scaled( mag[b] );
            if ( live && !small ) o[b * VB_FRAMES] = lv;
         }
         bool nyq_small = false;
         if ( cg == 0 )
         {
            nyq_small = mnyq < tau;
            const float lv = out_mode ? mnyq : hyb_log1p_scaled( mnyq );
            if ( live && !nyq_small ) o[128 * VB_FRAMES] = lv;
         }
         // flagged bins go to a global work list and are re-evaluated exactly by stft_fixup_kernel (thousands of warps in
         // flight hide the latency of the gather; in this kernel, with two worker warps per scheduler, it would be exposed).
         // Entry = (chunk << 13) | (bin << 5) | frame. One atomicAdd per warp; lanes that do not fit fall back to the
         // in-kernel evaluation below (correct, just slow -- only degenerate inputs such as pure tones get there).
         {
            const unsigned mycnt = live ? (unsigned)__popcll( fl ) + ( nyq_small ? 1u : 0u ) : 0u;
            unsigned incl = mycnt;
#pragma unroll
            for ( int off = 1; off < 32; off <<= 1 )
            {
               const unsigned nb = __shfl_up_sync( 0xffffffffu, incl, off );
               if ( lane >= off ) incl += nb;
            }
            const unsigned total = __shfl_sync( 0xffffffffu, incl, 31 );
            unsigned base = 0;
            if ( total )
            {
               if ( lane == 31 ) base = atomicAdd( fix_count, total );
               base = __shfl_sync( 0xffffffffu, base, 31 );
            }
            unsigned pos = base + incl - mycnt;
            const bool fits = fix_list && ( base + total <= fix_cap );
            if ( fits && mycnt )
            {
               const unsigned long long key = ( (unsigned long long)ci << 13 ) | (unsigned long long)t;
               unsigned long long m = fl;
               while ( m )
               {
                  const int b = __ffsll( (long long)m ) - 1;
                  m &= m - 1;
                  fix_list[pos++] = key | ( (unsigned long long)( cg * STC_BPT + b ) << 5 );
               }
               if ( nyq_small ) fix_list[pos++] = key | ( 128ull << 5 );
            }
            const unsigned lanes0 = __ballot_sync( 0xffffffffu, !fits && mycnt != 0u );
            if ( lanes0 && ci < nchunks )
            {
               const void *cp = chunk_ptr( ci );
               float *oc = spec + (size_t)ci * ( VB_BINS * VB_FRAMES );
               unsigned lanes = lanes0;
               while ( lanes )
               {
                  const int src = __ffs( lanes ) - 1;
                  lanes &= lanes - 1;
                  unsigned long long m = ( (unsigned long long)__shfl_sync( 0xffffffffu, (unsigned)( fl >> 32 ), src ) << 32 ) | __shfl_sync( 0xffffffffu, (unsigned)fl, src );
                  bool ny = __shfl_sync( 0xffffffffu, nyq_small ? 1 : 0, src ) != 0;
                  while ( m || ny )
                  {
                     int f;
                     if ( m )
                     {
                        const int b = __ffsll( (long long)m ) - 1;
                        m &= m - 1;
                        f = cg * STC_BPT + b;
                     }
                     else
                     {
                        ny = false;
                        f = 128;
                     }
                     const float ex_m = stc_exact_mag<F32>( cp, basis, f, src, lane );
                     if ( lane == 0 ) oc[f * VB_FRAMES + src] = out_mode ? ex_m : hyb_log1p_scaled( ex_m );
                     ++nflag;
                  }
               }
            }
         }
      }
      if ( flagged && lane == 0 && nflag ) atomicAdd( flagged, (unsigned long long)nflag );
   }
   tc::fence_before_sync();
   __syncthreads();
   if ( warp == STC_WORKER_WARPS ) tc::tmem_dealloc( tmem, 512 );
}

// exact re-evaluation of the bins flagged by stft_tc_kernel: one warp per work-list entry, the reference's rounding
// sequence (stft.c:108-184, 194-213) from the caller's input samples; bit-identical magnitudes
template <bool F32>
__global__ void __launch_bounds__( 256 )
stft_fixup_kernel( const void *__restrict__ in, long long stream_stride, int nw, const float *__restrict__ basis, float *__restrict__ spec,
                   const unsigned long long *__restrict__ fix_list, const unsigned int *__restrict__ fix_count, unsigned int fix_cap, int out_mode,
                   unsigned long long *__restrict__ flagged, int nchunks )
{
   const unsigned n = min( *fix_count, fix_cap );
   const int lane = threadIdx.x & 31;
   const unsigned warp = ( blockIdx.x * blockDim.x + threadIdx.x ) >> 5, nwarps = ( gridDim.x * blockDim.x ) >> 5;
   for ( unsigned e = warp; e < n; e += nwarps )
   {
      const unsigned long long key = __ldg( fix_list + e );
      const int t = (int)( key & 31ull ), f = (int)( ( key >> 5 ) & 255ull );
      // A warp of stft_tc_kernel whose entries did not fit reserved its range without writing it (it evaluated in-kernel),
      // so below the cap the list can hold stale entries of an earlier launch: skip anything outside this launch. A stale
      // entry that is in range only replaces an approximate magnitude by the exact one.
      if ( ( key >> 13 ) >= (unsigned long long)nchunks || f >= VB_BINS || t >= VB_FRAMES ) continue;
      const int ci = (int)( key >> 13 );
      const int s = ci / nw, c = ci - s * nw;
      const long long off = (long long)s * stream_stride + (long long)c * VB_CHUNK;
      const void *cp = F32 ? (const void *)( reinterpret_cast<const float *>( in ) + off ) : (const void *)( reinterpret_cast<const int16_t *>( in ) + off );
      const float ex_m = stc_exact_mag<F32>( cp, basis, f, t, lane );
      if ( lane == 0 ) spec[(size_t)ci * ( VB_BINS * VB_FRAMES ) + f * VB_FRAMES + t] = out_mode ? ex_m : hyb_log1p_scaled( ex_m );
   }
   if ( flagged && blockIdx.x == 0 && threadIdx.x == 0 && n ) atomicAdd( flagged, (unsigned long long)n );
}

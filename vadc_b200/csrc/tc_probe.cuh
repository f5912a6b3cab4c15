// vadc_b200/csrc/tc_probe.cuh -- one-tile tcgen05 GEMM used as a parity tap for the tensor-core
// plumbing of tc_common.cuh: D[128][N] = A[128][K] * B[N][K]^T with the bf16xS split scheme.
// Exercised by tests/test_gpu_tensorcore.py against an fp64 host product; also reports the cycle
// count of the MMA phase so the per-instruction cost at small N can be read off.
#pragma once
#include "tc_common.cuh"

#define TCP_THREADS 160 // 4 epilogue/loader warps + 1 MMA warp

// nsplit: 1, 2 or 3.  N % 16 == 0, 16 <= N <= 256.  K % 16 == 0, K <= 128.
// smem: nsplit * (K/8) * (128 + N) * 16 bytes + 16
__global__ void __launch_bounds__( TCP_THREADS, 1 )
tc_probe_kernel( const float *__restrict__ A, const float *__restrict__ B, float *__restrict__ D, int N, int K, int nsplit, int reps,
                 long long *__restrict__ cycles )
{
   extern __shared__ __align__( 128 ) unsigned char smem_raw[];
   __shared__ __align__( 8 ) uint64_t bar_mma;
   __shared__ uint32_t tmem_base_slot;

   const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
   const int nchunk = K / 8;
   const uint32_t lboA = 128 * 16, lboB = (uint32_t)N * 16;
   const uint32_t szA = (uint32_t)nchunk * lboA, szB = (uint32_t)nchunk * lboB;
   unsigned char *sA = smem_raw;                 // [nsplit][K/8][128][8] bf16
   unsigned char *sB = smem_raw + nsplit * szA;      // [nsplit][K/8][N][8]   bf16

   if ( warp == 4 )
   {
      tc::tmem_alloc( &tmem_base_slot, 256 );
      if ( lane == 0 )
      {
         tc::mbar_init( &bar_mma, 1 );
         tc::mbar_fence_init();
      }
   }
   // operands: fp32 global -> bf16 splits in the canonical layout
   for ( int i = tid; i < 128 * K; i += TCP_THREADS )
   {
      int r = i / K, k = i - r * K;
      float v = A[i];
#pragma unroll 1
      for ( int s = 0; s < nsplit; ++s )
      {
         __nv_bfloat16 b = __float2bfloat16_rn( v );
         *reinterpret_cast<__nv_bfloat16 *>( sA + s * szA + tc::op_off( r, k, lboA ) ) = b;
         v -= __bfloat162float( b );
      }
   }
   for ( int i = tid; i < N * K; i += TCP_THREADS )
   {
      int r = i / K, k = i - r * K;
      float v = B[i];
#pragma unroll 1
      for ( int s = 0; s < nsplit; ++s )
      {
         __nv_bfloat16 b = __float2bfloat16_rn( v );
         *reinterpret_cast<__nv_bfloat16 *>( sB + s * szB + tc::op_off( r, k, lboB ) ) = b;
         v -= __bfloat162float( b );
      }
   }
   tc::fence_async_smem();
   tc::fence_before_sync();
   __syncthreads();
   tc::fence_after_sync();
   const uint32_t tmem = tmem_base_slot;

   if ( warp == 4 )
   {
      {
         const uint32_t idesc = tc::idesc_bf16_f32( 128, N );
         long long t0 = clock64();
         if ( K == 128 )
         {
            // issue-optimised path (what the production kernels do): descriptors advance by constants
            const uint32_t stepA = ( 2 * lboA ) >> 4, stepB = ( 2 * lboB ) >> 4;
            const uint64_t dA = tc::smem_desc( tc::smem_u32( sA ), lboA, 128 ), dB = tc::smem_desc( tc::smem_u32( sB ), lboB, 128 );
            for ( int rep = 0; rep < reps; ++rep )
            {
               if ( tc::elect_one() )
               {
                  uint32_t acc = 0;
                  for ( int sa = 0; sa < nsplit; ++sa )
                     for ( int sb = 0; sa + sb < nsplit; ++sb )
                     {
                        const uint64_t da0 = dA + (uint64_t)( sa * ( szA >> 4 ) ), db0 = dB + (uint64_t)( sb * ( szB >> 4 ) );
#pragma unroll
                        for ( int kk = 0; kk < 8; ++kk )
                        {
                           tc::mma_bf16( tmem, da0 + (uint64_t)( kk * stepA ), db0 + (uint64_t)( kk * stepB ), idesc, acc );
                           acc = 1;
                        }
                     }
                  tc::mma_commit( &bar_mma );
               }
               __syncwarp();
               tc::mbar_wait( &bar_mma, rep & 1 );
            }
         }
         else
         for ( int rep = 0; rep < reps; ++rep )
         {
            uint32_t acc = 0;
            for ( int sa = 0; sa < nsplit; ++sa )
               for ( int sb = 0; sa + sb < nsplit; ++sb )
                  for ( int kk = 0; kk < K / 16; ++kk )
                  {
                     uint64_t da = tc::smem_desc( tc::smem_u32( sA + sa * szA ) + kk * 2 * lboA, lboA, 128 );
                     uint64_t db = tc::smem_desc( tc::smem_u32( sB + sb * szB ) + kk * 2 * lboB, lboB, 128 );
                     if ( tc::elect_one() ) tc::mma_bf16( tmem, da, db, idesc, acc );
                     acc = 1;
                  }
            if ( tc::elect_one() ) tc::mma_commit( &bar_mma );
            tc::mbar_wait( &bar_mma, rep & 1 );
         }
         long long t1 = clock64();
         if ( cycles && lane == 0 ) *cycles = t1 - t0;
      }
      __syncwarp();
   }
   else
   {
      // the last commit completes phase (reps-1)&1
      tc::mbar_wait( &bar_mma, ( reps - 1 ) & 1 );
      tc::fence_after_sync();
      const int row = warp * 32 + lane;
      for ( int c0 = 0; c0 < N; c0 += 16 )
      {
         float v[16];
         tc::tmem_ld16( tmem + ( (uint32_t)( warp * 32 ) << 16 ) + (uint32_t)c0, v );
         tc::tmem_wait_ld();
#pragma unroll
         for ( int j = 0; j < 16; ++j ) D[(size_t)row * N + c0 + j] = v[j];
      }
      tc::fence_before_sync();
   }
   __syncthreads();
   if ( warp == 4 ) tc::tmem_dealloc( tmem, 256 );
}

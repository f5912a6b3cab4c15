// vadc_b200/csrc/tc_common.cuh -- sm_100a tensor-core plumbing: tcgen05.mma / TMEM / mbarrier PTX
// wrappers and the shared-memory operand layout used by the tensor-core kernels.
//
// Precision scheme ("bf16xS"): an fp32 value v is split into S bf16 terms v = v0 + v1 (+ v2),
// v0 = bf16(v), v1 = bf16(v - v0), ...; a product a*b is evaluated as the sum of the partial
// products a_i*b_j with i + j < S (S = 2: 3 MMAs, ~16 significant bits; S = 3: 6 MMAs, ~fp32),
// accumulated in fp32 by the tensor core. DESIGN.md section 2 has the measurements that justify S.
//
// Operand layout (both A and B are K-major, no swizzle = UMMA "interleave" canonical layout):
// a [R rows][K] bf16 operand is stored as [K/8 chunks][R][8]: the 8x(16 B) core matrices of the
// canonical layout are 8 consecutive rows of one chunk (128 contiguous bytes), so
//   SBO (stride between 8-row groups)    = 128 B
//   LBO (stride between K chunks of 8)   = chunk stride (R*16 B, or padded)
// and one MMA (K = 16) consumes two consecutive chunks; advancing K by 16 adds 2*LBO to the address.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace tc
{

__device__ __forceinline__ uint32_t smem_u32( const void *p ) { return (uint32_t)__cvta_generic_to_shared( p ); }

// ---- mbarrier ------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init( uint64_t *bar, uint32_t count )
{
   asm volatile( "mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"( smem_u32( bar ) ), "r"( count ) : "memory" );
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile( "fence.mbarrier_init.release.cluster;" ::: "memory" ); }
__device__ __forceinline__ void mbar_arrive( uint64_t *bar )
{
   asm volatile( "{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"( smem_u32( bar ) ) : "memory" );
}
__device__ __forceinline__ bool mbar_try_wait( uint64_t *bar, uint32_t parity )
{
   uint32_t ok;
   asm volatile( "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"( ok )
                 : "r"( smem_u32( bar ) ), "r"( parity )
                 : "memory" );
   return ok != 0;
}
__device__ __forceinline__ void mbar_wait( uint64_t *bar, uint32_t parity )
{
   while ( !mbar_try_wait( bar, parity ) ) {}
}

// generic-proxy writes (st.shared) -> visible to the async proxy (tensor core operand reads)
__device__ __forceinline__ void fence_async_smem() { asm volatile( "fence.proxy.async.shared::cta;" ::: "memory" ); }

// ---- TMEM ----------------------------------------------------------------------------------------
// one full warp; ncols power of two in [32, 512]; the base address lands in *slot (shared memory)
__device__ __forceinline__ void tmem_alloc( uint32_t *slot, uint32_t ncols )
{
   asm volatile( "tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"( smem_u32( slot ) ), "r"( ncols ) : "memory" );
   asm volatile( "tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory" );
}
__device__ __forceinline__ void tmem_dealloc( uint32_t taddr, uint32_t ncols )
{
   asm volatile( "tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"( taddr ), "r"( ncols ) : "memory" );
}
__device__ __forceinline__ void fence_before_sync() { asm volatile( "tcgen05.fence::before_thread_sync;" ::: "memory" ); }
__device__ __forceinline__ void fence_after_sync() { asm volatile( "tcgen05.fence::after_thread_sync;" ::: "memory" ); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile( "tcgen05.wait::ld.sync.aligned;" ::: "memory" ); }

// warp-collective: lane l of warp w reads TMEM lane 32*(w%4)+l, 16 / 8 consecutive 32-bit columns
__device__ __forceinline__ void tmem_ld16( uint32_t taddr, float ( &v )[16] )
{
   uint32_t r[16];
   asm volatile( "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"( r[0] ), "=r"( r[1] ), "=r"( r[2] ), "=r"( r[3] ), "=r"( r[4] ), "=r"( r[5] ), "=r"( r[6] ), "=r"( r[7] ), "=r"( r[8] ),
                   "=r"( r[9] ), "=r"( r[10] ), "=r"( r[11] ), "=r"( r[12] ), "=r"( r[13] ), "=r"( r[14] ), "=r"( r[15] )
                 : "r"( taddr )
                 : "memory" );
#pragma unroll
   for ( int i = 0; i < 16; ++i ) v[i] = __uint_as_float( r[i] );
}
__device__ __forceinline__ void tmem_ld8( uint32_t taddr, float ( &v )[8] )
{
   uint32_t r[8];
   asm volatile( "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"( r[0] ), "=r"( r[1] ), "=r"( r[2] ), "=r"( r[3] ), "=r"( r[4] ), "=r"( r[5] ), "=r"( r[6] ), "=r"( r[7] )
                 : "r"( taddr )
                 : "memory" );
#pragma unroll
   for ( int i = 0; i < 8; ++i ) v[i] = __uint_as_float( r[i] );
}

__device__ __forceinline__ void tmem_ld32( uint32_t taddr, float ( &v )[32] )
{
   uint32_t r[32];
   asm volatile( "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                 "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"( r[0] ), "=r"( r[1] ), "=r"( r[2] ), "=r"( r[3] ), "=r"( r[4] ), "=r"( r[5] ), "=r"( r[6] ), "=r"( r[7] ), "=r"( r[8] ),
                   "=r"( r[9] ), "=r"( r[10] ), "=r"( r[11] ), "=r"( r[12] ), "=r"( r[13] ), "=r"( r[14] ), "=r"( r[15] ), "=r"( r[16] ), "=r"( r[17] ),
                   "=r"( r[18] ), "=r"( r[19] ), "=r"( r[20] ), "=r"( r[21] ), "=r"( r[22] ), "=r"( r[23] ), "=r"( r[24] ), "=r"( r[25] ), "=r"( r[26] ),
                   "=r"( r[27] ), "=r"( r[28] ), "=r"( r[29] ), "=r"( r[30] ), "=r"( r[31] )
                 : "r"( taddr )
                 : "memory" );
#pragma unroll
   for ( int i = 0; i < 32; ++i ) v[i] = __uint_as_float( r[i] );
}
// NC consecutive columns (NC = 16 or a multiple of 32) of this thread's TMEM lane; the caller waits (tmem_wait_ld)
template <int NC>
__device__ __forceinline__ void tmem_ld_cols( uint32_t taddr, float ( &v )[NC] )
{
   static_assert( NC == 16 || NC % 32 == 0, "column count" );
   if constexpr ( NC == 16 )
      tmem_ld16( taddr, v );
   else
   {
#pragma unroll
      for ( int c = 0; c < NC; c += 32 ) tmem_ld32( taddr + c, *reinterpret_cast<float( * )[32]>( &v[c] ) );
   }
}

// ---- descriptors ---------------------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_NONE, version 1 (sm_100)
__device__ __forceinline__ uint64_t smem_desc( uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes )
{
   return (uint64_t)( ( saddr & 0x3FFFFu ) >> 4 ) | ( (uint64_t)( ( lbo_bytes >> 4 ) & 0x3FFFu ) << 16 ) | ( (uint64_t)( ( sbo_bytes >> 4 ) & 0x3FFFu ) << 32 ) |
          ( 1ull << 46 );
}
// instruction descriptor for kind::f16: A, B = bf16 (K-major), D = fp32, dense
__host__ __device__ constexpr uint32_t idesc_bf16_f32( int M, int N )
{
   return ( 1u << 4 ) | ( 1u << 7 ) | ( 1u << 10 ) | ( (uint32_t)( N >> 3 ) << 17 ) | ( (uint32_t)( M >> 4 ) << 24 );
}

// same for A, B = fp16
__host__ __device__ constexpr uint32_t idesc_f16_f32( int M, int N )
{
   return ( 1u << 4 ) | ( (uint32_t)( N >> 3 ) << 17 ) | ( (uint32_t)( M >> 4 ) << 24 );
}

// one lane of a converged warp (keeps the surrounding code warp-uniform, so descriptors stay in
// uniform registers instead of going through a per-MMA R2UR election loop)
__device__ __forceinline__ bool elect_one()
{
   uint32_t pred;
   asm volatile( "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"( pred ) );
   return pred != 0;
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues
__device__ __forceinline__ void mma_bf16( uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate )
{
   asm volatile( "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"( d_tmem ),
                 "l"( a_desc ), "l"( b_desc ), "r"( idesc ), "r"( accumulate )
                 : "memory" );
}
// D[tmem] (+)= A[tmem] * B[smem]^T ; A sits in tensor memory (row m of the M-tile in lane m, K elements packed two per 32-bit
// column, K-major): the operand that never changes (a weight matrix) is read from TMEM instead of 4 KB of shared memory per MMA
__device__ __forceinline__ void mma_bf16_ts( uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate )
{
   asm volatile( "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"( d_tmem ),
                 "r"( a_tmem ), "l"( b_desc ), "r"( idesc ), "r"( accumulate )
                 : "memory" );
}
// four consecutive 32-bit columns of this thread's TMEM lane (warp w owns lanes 32 (w % 4) ..); the caller waits (tmem_wait_st)
__device__ __forceinline__ void tmem_st4( uint32_t taddr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3 )
{
   asm volatile( "tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"( taddr ), "r"( r0 ), "r"( r1 ), "r"( r2 ), "r"( r3 ) : "memory" );
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile( "tcgen05.wait::st.sync.aligned;" ::: "memory" ); }

// all previously issued MMAs of this thread arrive on `bar` when complete (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit( uint64_t *bar )
{
   asm volatile( "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"( smem_u32( bar ) ) : "memory" );
}

// ---- bf16 splitting ------------------------------------------------------------------------------
struct Split2
{
   __nv_bfloat16 hi, lo;
};
__device__ __forceinline__ Split2 split2( float v )
{
   Split2 s;
   s.hi = __float2bfloat16_rn( v );
   s.lo = __float2bfloat16_rn( v - __bfloat162float( s.hi ) );
   return s;
}

// fp16x2 split of 8 consecutive K elements of one operand row: hi = fp16(v), lo = fp16(v - hi) (22 significant
// bits together; products hi*hi + lo*hi + hi*lo). One 16-byte store per split into the [K/8][R][8] layout.
__device__ __forceinline__ void split_store8_f16( const float *v, unsigned char *hi_ptr, unsigned char *lo_ptr )
{
   __half2 hi[4], lo[4];
#pragma unroll
   for ( int i = 0; i < 4; ++i )
   {
      hi[i] = __floats2half2_rn( v[2 * i], v[2 * i + 1] );
      const float2 f = __half22float2( hi[i] );
      lo[i] = __floats2half2_rn( v[2 * i] - f.x, v[2 * i + 1] - f.y );
   }
   *reinterpret_cast<int4 *>( hi_ptr ) = *reinterpret_cast<const int4 *>( hi );
   *reinterpret_cast<int4 *>( lo_ptr ) = *reinterpret_cast<const int4 *>( lo );
}

// the same split written to TENSOR memory as an A operand: 8 K elements = 4 packed 32-bit columns per split of this thread's
// lane (row). The caller issues tmem_wait_st() before the hand-off to the MMA-issuing thread.
__device__ __forceinline__ void split_st8_f16_tmem( const float *v, uint32_t taddr_hi, uint32_t taddr_lo )
{
   __half2 hi[4], lo[4];
#pragma unroll
   for ( int i = 0; i < 4; ++i )
   {
      hi[i] = __floats2half2_rn( v[2 * i], v[2 * i + 1] );
      const float2 f = __half22float2( hi[i] );
      lo[i] = __floats2half2_rn( v[2 * i] - f.x, v[2 * i + 1] - f.y );
   }
   const uint32_t *h = reinterpret_cast<const uint32_t *>( hi ), *l = reinterpret_cast<const uint32_t *>( lo );
   tmem_st4( taddr_hi, h[0], h[1], h[2], h[3] );
   tmem_st4( taddr_lo, l[0], l[1], l[2], l[3] );
}

// byte offset of element (row r, k) in the [K/8][R][8] bf16 operand layout with chunk stride `lbo` bytes
__device__ __forceinline__ uint32_t op_off( int r, int k, uint32_t lbo ) { return (uint32_t)( k >> 3 ) * lbo + (uint32_t)r * 16u + (uint32_t)( k & 7 ) * 2u; }

} // namespace tc

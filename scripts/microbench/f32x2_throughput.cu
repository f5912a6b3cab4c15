// scripts/microbench/f32x2_throughput.cu -- how many FP32 lane-operations per clock and SM does a B200 retire for
// unfused multiplies and adds, scalar (FMUL/FADD) against packed (FMUL2/FADD2, sm_100 add/mul.rn.f32x2)?
// The bit-faithful kernels of this engine may not contract a multiply with an add, so this, not the FFMA peak, is their roofline.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_throughput f32x2_throughput.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk( float lo, float hi ) { u64 r; asm( "mov.b64 %0, {%1,%2};" : "=l"( r ) : "f"( lo ), "f"( hi ) ); return r; }
__device__ __forceinline__ void upk( u64 v, float &lo, float &hi ) { asm( "mov.b64 {%0,%1}, %2;" : "=f"( lo ), "=f"( hi ) : "l"( v ) ); }
__device__ __forceinline__ u64 mul2( u64 a, u64 b ) { u64 c; asm volatile( "mul.rn.f32x2 %0, %1, %2;" : "=l"( c ) : "l"( a ), "l"( b ) ); return c; }
__device__ __forceinline__ u64 add2( u64 a, u64 b ) { u64 c; asm volatile( "add.rn.f32x2 %0, %1, %2;" : "=l"( c ) : "l"( a ), "l"( b ) ); return c; }
__device__ __forceinline__ u64 fma2( u64 a, u64 b, u64 c ) { u64 d; asm volatile( "fma.rn.f32x2 %0, %1, %2, %3;" : "=l"( d ) : "l"( a ), "l"( b ), "l"( c ) ); return d; }

#define NCH 8
// mode 0: FADD x16 chains; 1: FADD2 x8; 2: FMUL x16; 3: FMUL2 x8; 4: FFMA x16; 5: FFMA2 x8
// 6: tree scalar: p=a*b (FMUL), acc+=p (FADD)  [2 lane-ops per pair, unfused]
// 7: FMUL2 + 2 scalar FADD; 8: 2 FMUL + FADD2; 9: FADD2 + LDS.128 broadcast per 2 FADD2; 10: FADD + LDS.128 per 4 FADD
template <int MODE>
__global__ void __launch_bounds__( 1024 ) k( float *out, const float *in, int iters, long long *cycles )
{
   __shared__ float4 sh[256];
   if ( threadIdx.x < 256 ) sh[threadIdx.x] = make_float4( in[threadIdx.x], 1.0f, 2.0f, 3.0f );
   __syncthreads();
   float a[16], m = in[threadIdx.x & 255], q = in[( threadIdx.x + 7 ) & 255];
   for ( int i = 0; i < 16; ++i ) a[i] = in[( threadIdx.x + i ) & 255];
   u64 A[8], M = pk( m, q ), Q = pk( q, m );
   for ( int i = 0; i < 8; ++i ) A[i] = pk( a[2 * i], a[2 * i + 1] );
   float4 ld = make_float4( 0, 0, 0, 0 );
   long long t0 = clock64();
   for ( int it = 0; it < iters; ++it )
   {
      if ( MODE == 0 ) { _Pragma( "unroll" ) for ( int r = 0; r < 4; ++r ) _Pragma( "unroll" ) for ( int i = 0; i < 16; ++i ) a[i] = __fadd_rn( a[i], m ); }
      if ( MODE == 1 ) { _Pragma( "unroll" ) for ( int r = 0; r < 4; ++r ) _Pragma( "unroll" ) for ( int i = 0; i < 8; ++i ) A[i] = add2( A[i], M ); }
      if ( MODE == 2 ) { _Pragma( "unroll" ) for ( int r = 0; r < 4; ++r ) _Pragma( "unroll" ) for ( int i = 0; i < 16; ++i ) a[i] = __fmul_rn( a[i], m ); }
      if ( MODE == 3 ) { _Pragma( "unroll" ) for ( int r = 0; r < 4; ++r ) _Pragma( "unroll" ) for ( int i = 0; i < 8; ++i ) A[i] = mul2( A[i], M ); }
      if ( MODE == 4 ) { _Pragma( "unroll" ) for ( int r = 0; r < 4; ++r ) _Pragma( "unroll" ) for ( int i = 0; i < 16; ++i ) a[i] = __fmaf_rn( a[i], m, q ); }
      if ( MODE == 5 ) { _Pragma( "unroll" ) for ( int r = 0; r < 4; ++r ) _Pragma( "unroll" ) for ( int i = 0; i < 8; ++i ) A[i] = fma2( A[i], M, Q ); }
      if ( MODE == 6 ) { _Pragma( "unroll" ) for ( int r = 0; r < 2; ++r ) _Pragma( "unroll" ) for ( int i = 0; i < 16; ++i ) a[i] = __fadd_rn( a[i], __fmul_rn( a[( i + 1 ) & 15], m ) ); }
      if ( MODE == 7 )
      {
         _Pragma( "unroll" ) for ( int r = 0; r < 2; ++r ) _Pragma( "unroll" ) for ( int i = 0; i < 8; ++i )
         {
            u64 p = mul2( pk( a[2 * i], a[2 * i + 1] ), M );
            float lo, hi;
            upk( p, lo, hi );
            a[( 2 * i + 2 ) & 15] = __fadd_rn( a[( 2 * i + 2 ) & 15], lo );
            a[( 2 * i + 3 ) & 15] = __fadd_rn( a[( 2 * i + 3 ) & 15], hi );
         }
      }
      if ( MODE == 8 )
      {
         _Pragma( "unroll" ) for ( int r = 0; r < 2; ++r ) _Pragma( "unroll" ) for ( int i = 0; i < 8; ++i )
         {
            float lo, hi;
            upk( A[( i + 1 ) & 7], lo, hi );
            A[i] = add2( A[i], pk( __fmul_rn( lo, m ), __fmul_rn( hi, q ) ) );
         }
      }
      if ( MODE == 9 )
      {
         _Pragma( "unroll" ) for ( int r = 0; r < 4; ++r ) _Pragma( "unroll" ) for ( int i = 0; i < 8; ++i )
         {
            A[i] = add2( A[i], M );
            if ( ( i & 1 ) == 0 )
            {
               float4 v = sh[( it + r * 8 + i ) & 255];
               ld.x += v.x;
            }
         }
      }
      if ( MODE == 10 )
      {
         _Pragma( "unroll" ) for ( int r = 0; r < 4; ++r ) _Pragma( "unroll" ) for ( int i = 0; i < 16; ++i )
         {
            a[i] = __fadd_rn( a[i], m );
            if ( ( i & 3 ) == 0 )
            {
               float4 v = sh[( it + r * 16 + i ) & 255];
               ld.x += v.x;
            }
         }
      }
   }
   long long t1 = clock64();
   float s = ld.x;
   for ( int i = 0; i < 16; ++i ) s += a[i];
   for ( int i = 0; i < 8; ++i ) { float lo, hi; upk( A[i], lo, hi ); s += lo + hi; }
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
   if ( threadIdx.x == 0 ) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run( const char *name, double laneops_per_iter_thread, int threads )
{
   float *out, *in;
   long long *cyc;
   cudaMalloc( &out, 148 * 1024 * 4 );
   cudaMalloc( &in, 1024 );
   cudaMalloc( &cyc, 148 * 8 );
   float h[256];
   for ( int i = 0; i < 256; ++i ) h[i] = 1.0f + i * 1e-3f;
   cudaMemcpy( in, h, 1024, cudaMemcpyHostToDevice );
   const int iters = 4096;
   k<MODE><<<148, threads>>>( out, in, 16, cyc );
   k<MODE><<<148, threads>>>( out, in, iters, cyc );
   cudaDeviceSynchronize();
   long long hc[148];
   cudaMemcpy( hc, cyc, sizeof hc, cudaMemcpyDeviceToHost );
   double mx = 0;
   for ( int i = 0; i < 148; ++i ) mx = hc[i] > mx ? hc[i] : mx;
   printf( "%-44s threads/SM %4d: %7.1f lane-ops/clk/SM\n", name, threads, laneops_per_iter_thread * iters * threads / mx );
   cudaFree( out ); cudaFree( in ); cudaFree( cyc );
}

int main()
{
   for ( int threads : { 256, 512, 1024 } )
   {
      run<0>( "FADD scalar", 64, threads );
      run<1>( "FADD2 packed", 64, threads );
      run<2>( "FMUL scalar", 64, threads );
      run<3>( "FMUL2 packed", 64, threads );
      run<4>( "FFMA scalar (1 lane-op each)", 64, threads );
      run<5>( "FFMA2 packed (1 lane-op per lane)", 64, threads );
      run<6>( "FMUL + FADD unfused", 64, threads );
      run<7>( "FMUL2 + 2 FADD", 64, threads );
      run<8>( "2 FMUL + FADD2", 64, threads );
      run<9>( "FADD2 + LDS.128 per 2 (math lane-ops only)", 64, threads );
      run<10>( "FADD + LDS.128 per 4 (math lane-ops only)", 64, threads );
   }
   return 0;
}

// scripts/microbench/lds_throughput.cu -- shared-memory load throughput next to the FP32 pipe on a B200:
// how many LDS (32/64/128 bit; broadcast or one word per lane) per clock and SM, alone and mixed with independent FADDs?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds_throughput lds_throughput.cu
#include <cstdio>
#include <cuda_runtime.h>

// MODE: 0 LDS.32 broadcast, 1 LDS.128 broadcast, 2 LDS.32 distinct (lane-consecutive), 3 LDS.128 distinct (lane-consecutive 16 B),
//       4 LDS.64 broadcast, 5 LDS.128 two addresses per warp (even / odd lanes)
//       6..9 LDS.128 with 8 (lane>>2), 8 (lane&7), 4 (lane>>3), 4 (lane&3) distinct 16-byte addresses per warp; 10,11 LDS.64 with 8 (lane>>2), 4 (lane>>3)
// NF = independent FADDs issued per load (0, 4, 8, 16)
template <int MODE, int NF>
__global__ void __launch_bounds__( 1024 ) k( float *out, const float *in, int iters, long long *cycles )
{
   extern __shared__ float4 sh[]; // 2048 float4 = 32 KB
   for ( int i = threadIdx.x; i < 2048; i += blockDim.x ) sh[i] = make_float4( in[i & 255], 1.0f, 2.0f, 3.0f );
   __syncthreads();
   const int lane = threadIdx.x & 31;
   const float *base;
   if ( MODE == 0 || MODE == 1 || MODE == 4 ) base = (const float *)sh;
   else if ( MODE == 2 ) base = (const float *)sh + lane;
   else if ( MODE == 3 ) base = (const float *)sh + lane * 4;
   else if ( MODE == 5 ) base = (const float *)sh + ( lane & 1 ) * 4;
   else if ( MODE == 6 ) base = (const float *)sh + ( lane >> 2 ) * 4;
   else if ( MODE == 7 ) base = (const float *)sh + ( lane & 7 ) * 4;
   else if ( MODE == 8 ) base = (const float *)sh + ( lane >> 3 ) * 4;
   else if ( MODE == 9 ) base = (const float *)sh + ( lane & 3 ) * 4;
   else if ( MODE == 10 ) base = (const float *)sh + ( lane >> 2 ) * 2;
   else base = (const float *)sh + ( lane >> 3 ) * 2;
   float a[16], m = in[threadIdx.x & 255];
   for ( int i = 0; i < 16; ++i ) a[i] = in[( threadIdx.x + i ) & 255];
   float acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
   long long t0 = clock64();
   for ( int it = 0; it < iters; ++it )
   {
      const float *p = base + ( it & 1 ) * 2048; // loop-carried address so the loads cannot be hoisted
#pragma unroll
      for ( int j = 0; j < 16; ++j )
      {
         if ( MODE == 0 || MODE == 2 )
         {
            float v;
            asm volatile( "ld.shared.f32 %0, [%1];" : "=f"( v ) : "r"( (unsigned)__cvta_generic_to_shared( p + j * 128 ) ) );
            if ( j & 1 ) acc0 = __uint_as_float( __float_as_uint( acc0 ) ^ __float_as_uint( v ) ); else acc1 = __uint_as_float( __float_as_uint( acc1 ) ^ __float_as_uint( v ) );
         }
         else if ( MODE == 4 || MODE >= 10 )
         {
            float v, w;
            asm volatile( "ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"( v ), "=f"( w ) : "r"( (unsigned)__cvta_generic_to_shared( p + j * 128 ) ) );
            acc0 = __uint_as_float( __float_as_uint( acc0 ) ^ __float_as_uint( v ) );
            acc1 = __uint_as_float( __float_as_uint( acc1 ) ^ __float_as_uint( w ) );
         }
         else
         {
            float v, w, x, y;
            asm volatile( "ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"( v ), "=f"( w ), "=f"( x ), "=f"( y ) : "r"( (unsigned)__cvta_generic_to_shared( p + j * 128 ) ) );
            acc0 = __uint_as_float( __float_as_uint( acc0 ) ^ __float_as_uint( v ) );
            acc1 = __uint_as_float( __float_as_uint( acc1 ) ^ __float_as_uint( w ) );
            acc2 = __uint_as_float( __float_as_uint( acc2 ) ^ __float_as_uint( x ) );
            acc3 = __uint_as_float( __float_as_uint( acc3 ) ^ __float_as_uint( y ) );
         }
#pragma unroll
         for ( int f = 0; f < NF; ++f ) a[( j * NF + f ) & 15] = __fadd_rn( a[( j * NF + f ) & 15], m );
      }
   }
   long long t1 = clock64();
   float s = acc0 + acc1 + acc2 + acc3;
   for ( int i = 0; i < 16; ++i ) s += a[i];
   out[blockIdx.x * blockDim.x + threadIdx.x] = s;
   if ( threadIdx.x == 0 ) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE, int NF>
void run( const char *name, int threads )
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
   cudaFuncSetAttribute( k<MODE, NF>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536 );
   k<MODE, NF><<<148, threads, 65536>>>( out, in, 16, cyc );
   k<MODE, NF><<<148, threads, 65536>>>( out, in, iters, cyc );
   cudaDeviceSynchronize();
   long long hc[148];
   cudaMemcpy( hc, cyc, sizeof hc, cudaMemcpyDeviceToHost );
   double mx = 0;
   for ( int i = 0; i < 148; ++i ) mx = hc[i] > mx ? hc[i] : mx;
   const double warps = threads / 32.0;
   printf( "%-34s FADD per load %2d, threads/SM %4d: %6.3f LDS/clk/SM, %6.1f FADD lane-ops/clk/SM\n", name, NF, threads, 16.0 * iters * warps / mx,
           16.0 * NF * iters * threads / mx );
   cudaFree( out ); cudaFree( in ); cudaFree( cyc );
}

#define ALLNF( MODE, NAME, T ) run<MODE, 0>( NAME, T ); run<MODE, 4>( NAME, T ); run<MODE, 8>( NAME, T ); run<MODE, 16>( NAME, T );
int main()
{
   
   for ( int threads : { 1024 } )
   {
      ALLNF( 6, "LDS.128 8 addr (lane>>2)", threads )
      ALLNF( 7, "LDS.128 8 addr (lane&7)", threads )
      ALLNF( 8, "LDS.128 4 addr (lane>>3)", threads )
      ALLNF( 9, "LDS.128 4 addr (lane&3)", threads )
      ALLNF( 10, "LDS.64 8 addr (lane>>2)", threads )
      ALLNF( 11, "LDS.64 4 addr (lane>>3)", threads )
      ALLNF( 0, "LDS.32 broadcast", threads )
      ALLNF( 4, "LDS.64 broadcast", threads )
      ALLNF( 1, "LDS.128 broadcast", threads )
      ALLNF( 5, "LDS.128 two addresses per warp", threads )
      ALLNF( 2, "LDS.32 one word per lane", threads )
      ALLNF( 3, "LDS.128 16 B per lane", threads )
   }
   return 0;
}

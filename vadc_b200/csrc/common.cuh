// vadc_b200/csrc/common.cuh -- shared device helpers and the packed-weight layout.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define VB_CHUNK 1536
#define VB_FRAMES 25
#define VB_BINS 129

// ---- model dimensions (tensor.h:154-170) -----------------------------
struct LayerDims
{
   int cin, c, t, stride, proj;
};
__host__ __device__ constexpr LayerDims layer_dims( int l )
{
   return l == 0 ? LayerDims{129, 16, 25, 2, 1}
        : l == 1 ? LayerDims{16, 32, 13, 2, 1}
        : l == 2 ? LayerDims{32, 32, 7, 1, 0}
                 : LayerDims{32, 64, 7, 1, 1};
}

// ---- packed per-layer weights (device global memory, floats) ---------------------------------
// All offsets are in floats and multiples of 4 (16-byte aligned for float4 access).
//   DW     [CIN][8]      w0..w4, bias, 0, 0
//   PW     first layer: [CIN][2*C]  (pw_w[o][f] for o<C, then proj_w[o][f])    (f-major)
//          other layers: [C][KP] with KP = CIN*(1+proj): row o = [pw_w[o][:], proj_w[o][:]]
//   PWB    [C]           pw_b (+ proj_b)
//   QKV    [2 heads] x { W[3*D][C] rows q_h (D), k_h (D), v_h (D);  b[3*D] }   with D = C/2
//   AO     [C][C], AOB [C]        attention out-proj
//   LN1W, LN1B [C]
//   F1 [C][C], F1B [C], F2 [C][C], F2B [C]
//   LN2W, LN2B [C]
//   CV [C][C], CVB [C]
//   BNM [C] running mean, BNS [C] sqrtf(var+eps), BNW [C], BNB [C]
template <int L>
struct LayerPack
{
   static constexpr LayerDims d = layer_dims( L );
   static constexpr int CIN = d.cin, C = d.c, T = d.t, STRIDE = d.stride, PROJ = d.proj;
   static constexpr int D = C / 2;
   static constexpr int TOUT = 1 + ( T - 1 ) / STRIDE;
   static constexpr int KP = CIN * ( 1 + PROJ );
   static constexpr int al4( int x ) { return ( x + 3 ) & ~3; }
   static constexpr int DW = 0;
   static constexpr int PW = DW + CIN * 8;
   static constexpr int PWB = PW + al4( CIN == VB_BINS ? CIN * 2 * C : C * KP );
   static constexpr int QKV = PWB + al4( C );
   static constexpr int QH = 3 * D * C + 3 * D; // per-head stride inside QKV
   static constexpr int AO = QKV + 2 * QH;
   static constexpr int AOB = AO + C * C;
   static constexpr int LN1W = AOB + C;
   static constexpr int LN1B = LN1W + C;
   static constexpr int F1 = LN1B + C;
   static constexpr int F1B = F1 + C * C;
   static constexpr int F2 = F1B + C;
   static constexpr int F2B = F2 + C * C;
   static constexpr int LN2W = F2B + C;
   static constexpr int LN2B = LN2W + C;
   static constexpr int CV = LN2B + C;
   static constexpr int CVB = CV + C * C;
   static constexpr int BNM = CVB + C;
   static constexpr int BNS = BNM + C;
   static constexpr int BNW = BNS + C;
   static constexpr int BNB = BNW + C;
   static constexpr int TOTAL = BNB + C;
};

struct DeviceWeights
{
   const float *basis_pack; // [2 halves][64 kquads][128 rows][4]  (stft_kernel.cuh)
   const float *basis_raw;  // [258][256] as stored in the container (stft_hybrid_kernel.cuh)
   const float *layer[4];   // LayerPack<L> blobs
   const float *lstm_w;     // [2 layers][32 kquads][256 rows][4]
   const float *lstm_b;     // [2][256]
   const float *dec_w;      // [2][64]
   const float *dec_b;      // [2]
};

__device__ __forceinline__ float4 ld4( const float *p ) { return *reinterpret_cast<const float4 *>( p ); }
__device__ __forceinline__ void st4( float *p, float4 v ) { *reinterpret_cast<float4 *>( p ) = v; }

// Two fp32 values in a 64-bit register pair for the packed arithmetic of sm_100 (mul.rn.f32x2 -> FMUL2, add.rn.f32x2 -> FADD2): each
// half is an independently rounded IEEE operation, one issue slot for both. ONE rule: a packed product never feeds a packed add --
// ptxas 12.9 contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (even with --fmad=false), which changes the rounding. The exact path
// packs the PRODUCTS (operands are the register pairs an LDS.128 delivers) and adds scalar; the alternative -- scalar products,
// packed adds -- measured slower and survives behind XL_PACKED_MUL 0. tests/test_host_logic.py reads the SASS: no FFMA2 in an
// exact-path kernel, no kernel with both FMUL2 and FADD2.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2( float lo, float hi )
{
   f32x2 r;
   asm( "mov.b64 %0, {%1,%2};" : "=l"( r ) : "f"( lo ), "f"( hi ) );
   return r;
}
__device__ __forceinline__ void unpk2( f32x2 v, float &lo, float &hi ) { asm( "mov.b64 {%0,%1}, %2;" : "=f"( lo ), "=f"( hi ) : "l"( v ) ); }
__device__ __forceinline__ f32x2 add2( f32x2 a, f32x2 b )
{
   f32x2 c;
   asm( "add.rn.f32x2 %0, %1, %2;" : "=l"( c ) : "l"( a ), "l"( b ) );
   return c;
}
// (only where the products feed SCALAR additions)
__device__ __forceinline__ f32x2 mul2( f32x2 a, f32x2 b )
{
   f32x2 c;
   asm( "mul.rn.f32x2 %0, %1, %2;" : "=l"( c ) : "l"( a ), "l"( b ) );
   return c;
}
__device__ __forceinline__ f32x2 sub2( f32x2 a, f32x2 b )
{
   f32x2 c;
   asm( "sub.rn.f32x2 %0, %1, %2;" : "=l"( c ) : "l"( a ), "l"( b ) );
   return c;
}

// named barrier over `nthreads` threads (multiple of 32); id 1..15 (0 is __syncthreads)
__device__ __forceinline__ void bar_sync( int id, int nthreads )
{
   asm volatile( "bar.sync %0, %1;" ::"r"( id ), "r"( nthreads ) : "memory" );
}

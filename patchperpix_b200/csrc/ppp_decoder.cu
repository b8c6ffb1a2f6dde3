// ppp+dec: code -> patch decoder (experiments/flylight/setups/setup01/decode.py:16-65,
// torch_model.py:452-544) on the 5th-generation tensor cores.
//
// Network (flylight values, default_train_code.toml:61-85; layer definitions of the
// un-vendored funlib.learn.torch fork are ASSUMED as stated in DESIGN.md §8):
//   code[176] -> [22][2^3] -> 1x1x1 conv 22->128 + ReLU          (dec_from_code_kernel)
//   nearest x2 -> 3^3 conv 128->64 + ReLU                        (dec_conv_tc_kernel<128>)
//   3^3 conv 64->64 + ReLU, twice                                (dec_conv_tc_kernel<64>)
//   nearest x2 -> 3^3 conv 64->1 + ReLU -> 3^3 conv 1->1, twice  (dec_tail_kernel)
//   crop 8^3 -> 7^3 (offset 0)
// 97 % of the 58.6 MFLOP per code are the three 3^3 convolutions with 64 output
// channels.  They run as implicit GEMMs on tcgen05: activations live in HBM as
// fp16 [B][4][4][4][C]; for every filter tap the A operand of two codes (128
// output positions x 64 channels) is ONE TMA box {64,4,4,4,2} whose start is
// shifted by the tap - the TMA unit zero-fills what falls outside the 4^3 cube,
// which IS the 'same' padding.  Weights are [tap][cin/64][cout][64 cin] fp16.
// The last up-sampling layer (nearest x2 + 3^3 conv 64->1) is folded into the
// same kernel: each of the 8 output parities is a conv over the 4^3 input with
// pre-summed taps, i.e. a GEMM with N = 8 (padded to 16).
// Per k-block (tap, cin chunk): TMA -> 128B-swizzled smem ring -> 4 x tcgen05.mma
// (M128 N64 K16, fp32 accumulator in TMEM) issued by one thread; the epilogue
// warps read TMEM (tcgen05.ld), add bias, ReLU, and store fp16 into the interior
// of the next layer's padded buffer.
#include <cuda.h>
#include <cuda_fp16.h>
#include "ppp_api.cuh"
#include "../../include/ppp_b200.h"

#define DEC_CF 22          // code fmaps
#define DEC_C1 128
#define DEC_C0 64
#define DEC_STAGES 4
#define DEC_A_BYTES (128 * 128)     // 128 rows x 64 fp16
#define DEC_W_BYTES (64 * 128)      // 64 rows x 64 fp16

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void d_mbar_init(uint64_t* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" :: "r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void d_mbar_expect_tx(uint64_t* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n"
                 :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void d_mbar_wait(uint64_t* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "DWAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DDONE;\n"
        "bra DWAIT;\n"
        "DDONE:\n"
        "}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* smem, const CUtensorMap* tm, uint64_t* bar,
                                            int c0, int c1, int c2, int c3, int c4)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6, %7}], [%2];\n"
        :: "r"(smem_u32(smem)), "l"((uint64_t)tm), "r"(smem_u32(bar)),
           "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem, const CUtensorMap* tm, uint64_t* bar,
                                            int c0, int c1)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];\n"
        :: "r"(smem_u32(smem)), "l"((uint64_t)tm), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// K-major, 128-byte swizzled operand tile: rows of 128 B, 8-row atoms of 1024 B
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr)
{
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3ffff) >> 4);          // start address
    d |= (uint64_t)1 << 16;                               // leading byte offset (unused, swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                     // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                               // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                               // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" :: "r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n"
                 :: "r"(smem_u32(bar)) : "memory");
}
#define TMEM_LD_32(taddr, r)                                                                   \
    asm volatile(                                                                              \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                              \
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"                              \
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"           \
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),  \
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]),           \
          "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),        \
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),        \
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),        \
          "=r"(r[31])                                                                          \
        : "r"(taddr))

// ---------------------------------------------------------------------------
// from_code (1x1x1 conv 22->128 + ReLU) fused with the nearest x2 up-sampling:
// writes act [B][4^3][128] fp16.  One block per code.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dec_from_code_kernel(const float* __restrict__ codes, int64_t B, const float* __restrict__ w,
                     const float* __restrict__ bias, __half* __restrict__ act)
{
    const int64_t b = blockIdx.x;
    __shared__ float s_code[DEC_CF * 8];
    __shared__ __half s_out[8][DEC_C1];
    for (int i = threadIdx.x; i < DEC_CF * 8; i += 256) s_code[i] = codes[b * (DEC_CF * 8) + i];
    __syncthreads();
    for (int o = threadIdx.x; o < 8 * DEC_C1; o += 256) {
        const int src = o / DEC_C1, co = o % DEC_C1;
        float acc = bias[co];
#pragma unroll
        for (int ci = 0; ci < DEC_CF; ci++) acc = fmaf(w[co * DEC_CF + ci], s_code[ci * 8 + src], acc);
        s_out[src][co] = __float2half_rn(fmaxf(acc, 0.0f));
    }
    __syncthreads();
    // 64 target positions x 128 channels, 16 bytes per thread-iteration
    uint4* dst = reinterpret_cast<uint4*>(act + b * 64 * DEC_C1);
    for (int i = threadIdx.x; i < 64 * DEC_C1 / 8; i += 256) {
        const int pos = i / (DEC_C1 / 8), c8 = i % (DEC_C1 / 8);
        const int z = pos >> 4, y = (pos >> 2) & 3, x = pos & 3;
        const int src = ((z >> 1) * 2 + (y >> 1)) * 2 + (x >> 1);
        dst[i] = reinterpret_cast<const uint4*>(&s_out[src][0])[c8];
    }
}

// ---------------------------------------------------------------------------
// 3^3 convolution CIN -> COUT (+bias, ReLU) as an implicit GEMM on tcgen05.
// CTA = two codes (M = 128 output positions).  Warp 0: TMA producer, warp 1:
// TMEM allocation + MMA issue, warps 2-5: epilogue (one TMEM lane quarter each).
// COUT = 64: fp16 output [B][4^3][64].  COUT = 16: the folded up-sampling layer,
// f32 output [B][8^3] (column = output parity, only 8 of the 16 are real).
// ---------------------------------------------------------------------------
template <int CIN, int COUT>
__global__ void __launch_bounds__(192, 1)
dec_conv_tc_kernel(const __grid_constant__ CUtensorMap tm_act, const __grid_constant__ CUtensorMap tm_w,
                   const float* __restrict__ bias, void* __restrict__ out_, int64_t B)
{
    constexpr int NKB = 27 * (CIN / 64);
    constexpr int W_BYTES = COUT * 128;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 128-byte swizzled tiles need 1024-byte alignment (the launch adds 1 KB of slack)
    unsigned char* sbase = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* sA = sbase;                                      // [STAGES][16 KB]
    unsigned char* sW = sbase + DEC_STAGES * DEC_A_BYTES;           // [STAGES][COUT x 128 B]
    __shared__ __align__(8) uint64_t s_full[DEC_STAGES], s_empty[DEC_STAGES], s_acc;
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b0 = blockIdx.x * 2;

    if (threadIdx.x == 0) {
        for (int s = 0; s < DEC_STAGES; s++) { d_mbar_init(&s_full[s], 1); d_mbar_init(&s_empty[s], 1); }
        d_mbar_init(&s_acc, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;\n"
                     :: "r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < NKB; kb++) {
                const int s = kb % DEC_STAGES, it = kb / DEC_STAGES;
                if (it > 0) d_mbar_wait(&s_empty[s], (it - 1) & 1);
                const int tap = kb / (CIN / 64), cc = kb % (CIN / 64);
                const int dz = tap / 9, dy = (tap / 3) % 3, dx = tap % 3;
                d_mbar_expect_tx(&s_full[s], DEC_A_BYTES + W_BYTES);
                // box start shifted by the tap; outside [0,4) the TMA unit writes zeros
                tma_load_5d(sA + s * DEC_A_BYTES, &tm_act, &s_full[s], cc * 64, dx - 1, dy - 1,
                            dz - 1, b0);
                tma_load_2d(sW + s * W_BYTES, &tm_w, &s_full[s], 0, kb * COUT);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // instruction descriptor: D=F32, A=B=F16, both K-major, N=COUT, M=128
            const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) |
                                   (((uint32_t)COUT >> 3) << 17) | ((128u >> 4) << 24);
            for (int kb = 0; kb < NKB; kb++) {
                const int s = kb % DEC_STAGES, it = kb / DEC_STAGES;
                d_mbar_wait(&s_full[s], it & 1);
                asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
                const uint32_t a0 = smem_u32(sA + s * DEC_A_BYTES), w0 = smem_u32(sW + s * W_BYTES);
#pragma unroll
                for (int k = 0; k < 4; k++)                           // 4 x K16 = 64 channels
                    umma_f16(tmem, umma_desc_sw128(a0 + k * 32), umma_desc_sw128(w0 + k * 32), idesc,
                             (kb | k) != 0);
                umma_commit(&s_empty[s]);                             // frees the smem slot
            }
            umma_commit(&s_acc);                                      // accumulator complete
        }
    } else {
        // ---- epilogue: TMEM -> registers -> bias, ReLU, fp16 -> padded output --------
        const int q = warp & 3;                                       // TMEM lane quarter of this warp
        const int m = q * 32 + lane;                                  // output row = (code, position)
        d_mbar_wait(&s_acc, 0);
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16);
        const int64_t b = b0 + (m >> 6);
        const int pos = m & 63, z = pos >> 4, y = (pos >> 2) & 3, x = pos & 3;
        if constexpr (COUT == 64) {
            uint32_t r0[32], r1[32];
            TMEM_LD_32(taddr, r0);
            TMEM_LD_32(taddr + 32, r1);
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            if (b < B) {
                __half* dst = reinterpret_cast<__half*>(out_) + (b * 64 + pos) * DEC_C0;
                uint4 pk[8];
                __half2* h2 = reinterpret_cast<__half2*>(pk);
#pragma unroll
                for (int c = 0; c < 32; c += 2) {
                    float v0 = fmaxf(__uint_as_float(r0[c]) + bias[c], 0.0f);
                    float v1 = fmaxf(__uint_as_float(r0[c + 1]) + bias[c + 1], 0.0f);
                    h2[c >> 1] = __floats2half2_rn(v0, v1);
                    float u0 = fmaxf(__uint_as_float(r1[c]) + bias[32 + c], 0.0f);
                    float u1 = fmaxf(__uint_as_float(r1[c + 1]) + bias[32 + c + 1], 0.0f);
                    h2[16 + (c >> 1)] = __floats2half2_rn(u0, u1);
                }
#pragma unroll
                for (int i = 0; i < 8; i++) reinterpret_cast<uint4*>(dst)[i] = pk[i];
            }
        } else {
            // folded up-sampling layer: column p = output parity (a,b,c) of position (z,y,x)
            uint32_t r0[32];
            TMEM_LD_32(taddr, r0);          // only the first 16 columns were written
            asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
            if (b < B) {
                float* dst = reinterpret_cast<float*>(out_) + b * 512;
#pragma unroll
                for (int p = 0; p < 8; p++) {
                    const int oz = 2 * z + (p >> 2), oy = 2 * y + ((p >> 1) & 1), ox = 2 * x + (p & 1);
                    dst[(oz * 8 + oy) * 8 + ox] = fmaxf(__uint_as_float(r0[p]) + bias[0], 0.0f);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;\n" :: "r"(tmem));
    }
}

// ---------------------------------------------------------------------------
// tail: two 3^3 convs 1->1 (no activation) on the 8^3 map, crop to 7^3.
// u: f32 [B][8^3]; out: f32 [B][343] (logits, or probabilities with
// apply_sigmoid).  One warp-sized CTA group per code, SIMT (0.1 % of the flops).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dec_tail_kernel(const float* __restrict__ u, const float* __restrict__ w_a,
                const float* __restrict__ b_a, const float* __restrict__ w_b,
                const float* __restrict__ b_b, int apply_sigmoid, float* __restrict__ out)
{
    __shared__ float s_u[10][10][10];       // 8^3 + zero halo
    __shared__ float s_v[10][10][10];
    const int64_t b = blockIdx.x;
    const int tid = threadIdx.x;
    for (int i = tid; i < 1000; i += 256) { (&s_u[0][0][0])[i] = 0.0f; (&s_v[0][0][0])[i] = 0.0f; }
    __syncthreads();
    for (int o = tid; o < 512; o += 256)
        s_u[(o >> 6) + 1][((o >> 3) & 7) + 1][(o & 7) + 1] = u[b * 512 + o];
    __syncthreads();
    for (int o = tid; o < 512; o += 256) {
        const int z = o >> 6, y = (o >> 3) & 7, x = o & 7;
        float acc = b_a[0];
#pragma unroll
        for (int t = 0; t < 27; t++)
            acc = fmaf(w_a[t], s_u[z + t / 9][y + (t / 3) % 3][x + t % 3], acc);
        s_v[z + 1][y + 1][x + 1] = acc;
    }
    __syncthreads();
    for (int o = tid; o < 343; o += 256) {
        const int z = o / 49, y = (o / 7) % 7, x = o % 7;
        float acc = b_b[0];
#pragma unroll
        for (int t = 0; t < 27; t++)
            acc = fmaf(w_b[t], s_v[z + t / 9][y + (t / 3) % 3][x + t % 3], acc);
        if (apply_sigmoid) acc = 1.0f / (1.0f + __expf(-acc));
        out[b * 343 + o] = acc;
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode()
{
    static PFN_encodeTiled enc = nullptr;
    if (enc == nullptr) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) ==
                cudaSuccess && fn != nullptr)
            enc = (PFN_encodeTiled)fn;
    }
    return enc;
}

static int make_act_map(CUtensorMap* tm, const __half* act, int64_t B, int C)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) return ppp_fail(-1, "ppp_decode: cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[5] = {(cuuint64_t)C, 4, 4, 4, (cuuint64_t)B};
    cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)C * 8, (cuuint64_t)C * 32,
                             (cuuint64_t)C * 128};
    cuuint32_t box[5] = {64, 4, 4, 4, 2};
    cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, (void*)act, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : ppp_fail(-1, "ppp_decode: activation tensor map failed");
}

static int make_w_map(CUtensorMap* tm, const __half* w, int nkb, int cout)
{
    PFN_encodeTiled enc = get_encode();
    if (!enc) return ppp_fail(-1, "ppp_decode: cuTensorMapEncodeTiled unavailable");
    cuuint64_t dims[2] = {64, (cuuint64_t)nkb * cout};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)cout};
    cuuint32_t es[2] = {1, 1};
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, (void*)w, dims, strides, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : ppp_fail(-1, "ppp_decode: weight tensor map failed");
}

extern "C" int64_t ppp_decode_scratch_bytes(int64_t B)
{
    // act0 [B][64][128] fp16, two ping-pong [B][64][64] fp16, u f32 [B][512]
    return B * 64 * (128 + 64 + 64) * 2 + B * 512 * 4 + 1024;
}

// weights (device): w_fc f32 [128][22], b_fc f32 [128];
//   w_up0 fp16 [27][2][64][64], w_c0a / w_c0b fp16 [27][1][64][64]  ([tap][cin chunk][cout][cin]);
//   b_up0, b_c0a, b_c0b f32 [64];
//   w_up1 fp16 [27][1][16][64]: the folded up-sampling layer, row p < 8 = output parity,
//   rows 8..15 zero (b_up1 f32 [1]); w_c1a, w_c1b f32 [27] (+ scalar biases).
extern "C" int ppp_decode(const float* codes, int64_t B, const float* w_fc, const float* b_fc,
                          const void* w_up0, const float* b_up0, const void* w_c0a,
                          const float* b_c0a, const void* w_c0b, const float* b_c0b,
                          const void* w_up1, const float* b_up1, const float* w_c1a,
                          const float* b_c1a, const float* w_c1b, const float* b_c1b,
                          int32_t apply_sigmoid, float* patches, void* scratch, void* stream)
{
    if (B <= 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    __half* act0 = (__half*)scratch;
    __half* act1 = act0 + B * 64 * 128;
    __half* act2 = act1 + B * 64 * 64;
    float* u = (float*)(act2 + B * 64 * 64);
    dec_from_code_kernel<<<(unsigned)B, 256, 0, s>>>(codes, B, w_fc, b_fc, act0);
    CUtensorMap ta0, ta1, ta2, tw0, tw1, tw2, tw3;
    int rc;
    if ((rc = make_act_map(&ta0, act0, B, 128))) return rc;
    if ((rc = make_act_map(&ta1, act1, B, 64))) return rc;
    if ((rc = make_act_map(&ta2, act2, B, 64))) return rc;
    if ((rc = make_w_map(&tw0, (const __half*)w_up0, 54, 64))) return rc;
    if ((rc = make_w_map(&tw1, (const __half*)w_c0a, 27, 64))) return rc;
    if ((rc = make_w_map(&tw2, (const __half*)w_c0b, 27, 64))) return rc;
    if ((rc = make_w_map(&tw3, (const __half*)w_up1, 27, 16))) return rc;
    const size_t smem = DEC_STAGES * (DEC_A_BYTES + DEC_W_BYTES) + 1024;
    cudaFuncSetAttribute(dec_conv_tc_kernel<128, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(dec_conv_tc_kernel<64, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(dec_conv_tc_kernel<64, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    const unsigned grid = (unsigned)((B + 1) / 2);
    dec_conv_tc_kernel<128, 64><<<grid, 192, smem, s>>>(ta0, tw0, b_up0, act1, B);
    dec_conv_tc_kernel<64, 64><<<grid, 192, smem, s>>>(ta1, tw1, b_c0a, act2, B);
    dec_conv_tc_kernel<64, 64><<<grid, 192, smem, s>>>(ta2, tw2, b_c0b, act1, B);
    dec_conv_tc_kernel<64, 16><<<grid, 192, smem, s>>>(ta1, tw3, b_up1, u, B);
    dec_tail_kernel<<<(unsigned)B, 256, 0, s>>>(u, w_c1a, b_c1a, w_c1b, b_c1b, apply_sigmoid,
                                                 patches);
    return ppp_check("ppp_decode");
}

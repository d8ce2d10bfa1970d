// fcn_conv.cu -- implicit-GEMM convolution / transposed convolution for FCN-LectureNet on sm_100a:
// TMA (cp.async.bulk.tensor) -> 128B-swizzled shared memory -> tcgen05.mma (bf16 x bf16 -> fp32 in TMEM)
// -> tcgen05.ld epilogue with fused bias (BatchNorm folded) + exact-erf GELU + bf16/fp32 NHWC store.
//
// Replaces the nn.Conv2d / nn.ConvTranspose2d / BatchNorm2d / GELU stacks of
//   R/AccessMath/lecturenet_v1/FCN_lecturenet.py:26-201 (layers) and :260-323, :364-403 (forward).
//
// GEMM view ("row-run" implicit GEMM).  Activations are NHWC bf16 with physical zero padding in x:
//   buffer[n][y][xp][c], xp in [0, W + 2*pad).  For one vertical tap dy, output group r (S output pixels
//   x = S*r .. S*r+S-1) reads the contiguous run  buffer[n][y+dy-padY][S*r + j][c],  j in [0, KW+S-1), c in [0,C)
//   = (KW+S-1)*C contiguous elements.  So A[m=(y,r)][k] is a 2-D view with row stride S*C elements (rows
//   overlap) and the GEMM is   D[(y,r)][(sx,co)] = sum_dy sum_k A_dy[(y,r)][k] * Wp_dy[(sx,co)][k]
//   with Wp the filter re-packed (zero where the tap j-sx falls outside 0..KW-1).  S>1 packs S output
//   pixels into the N dimension, which is what makes the Cout=16/32 full-resolution 7x7 layers fill a UMMA
//   tile.  Concatenated inputs (skip connections, the `diff` image) are extra K segments with their own
//   tensor map -- torch.cat is never materialised.  Transposed 2x2/s2 conv = the same GEMM with KH=KW=1 and
//   the (sy,sx) sub-pixel in N (pixel shuffle in the epilogue).
//   The vertical halo is loaded once per K chunk: the A box has YT+KH-1 rows and tap dy just offsets the
//   UMMA descriptor by dy*RT rows (RT multiple of 8 keeps the 1024-byte swizzle atom aligned).
//
// CTA = 6 warps: warp0 TMA producer, warp1 TMEM owner + MMA issuer, warps2-5 epilogue (one TMEM lane
// quarter each).  Two smem rings: A boxes (per K chunk) and B tiles (per K chunk x dy).
#include "am_common.cuh"
#include "../../include/accessmath_b200.h"
#include <cuda.h>
#include <cuda_bf16.h>

struct alignas(64) ConvParams {
    CUtensorMap tmA[2];
    CUtensorMap tmB;
    int nseg;
    int seg_nkx[2];      // horizontal taps enumerated by separate loads (1 in row-run mode)
    int seg_nck[2];      // 64-element K chunks per (segment, kx)
    int seg_klast[2];    // valid K (multiple of 16) of the last chunk
    int seg_c1step[2];   // coordinate-1 step per kx (0 in row-run mode)
    int seg_c1off[2];    // coordinate-1 offset
    int KH, RT, YT, padY;
    int nRT, nYT;        // tiles per row / per frame column
    int nR, Hin;         // valid groups per row, valid rows
    int NT;              // UMMA N of this launch
    int Ntot_pad;        // rows per (chunk,dy) block of the packed weights
    int tmem_cols;
    int stagesA, stagesB;
    // epilogue
    void* out; int out_f32;
    int out_H, out_W;
    long long out_sn, out_sy; int out_sx, out_padx, out_coff;
    int Cout, Sy, Sx, Ntot, act;
    const float* bias;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t it = 0; !done; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (it > (1u << 26)) { printf("[accessmath_b200] mbarrier timeout (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x); __trap(); }
    }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"((unsigned long long)tm), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((unsigned long long)tm), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// K-major, 128-byte swizzle: rows of 128 B, 8-row atoms of 1024 B (cute::UMMA::SmemDescriptor, mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);          // start address, 16-byte units
    d |= (uint64_t)1 << 16;                          // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: 8 rows * 128 B
    d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }

__global__ void __launch_bounds__(192, 1) k_conv_gemm(const __grid_constant__ ConvParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [A ring][B ring][barriers][tmem ptr]
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bytesA = (uint32_t)(p.YT + p.KH - 1) * p.RT * 128u;
    const uint32_t bytesB = (uint32_t)p.NT * 128u;
    const uint32_t sA0 = smem_base;
    const uint32_t sB0 = sA0 + bytesA * p.stagesA;             // bytesA multiple of 1024 (RT % 8 == 0)
    const uint32_t bar0 = sB0 + ((bytesB * p.stagesB + 1023u) & ~1023u);
    // barriers: fullA[sA], emptyA[sA], fullB[sB], emptyB[sB], accum
    const uint32_t fullA = bar0, emptyA = fullA + 8 * p.stagesA, fullB = emptyA + 8 * p.stagesA, emptyB = fullB + 8 * p.stagesB;
    const uint32_t accum = emptyB + 8 * p.stagesB;
    const uint32_t tmem_slot = accum + 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // tile coordinates
    int t = blockIdx.x;
    const int rt = t % p.nRT; t /= p.nRT;
    const int yt = t % p.nYT; const int frame = t / p.nYT;
    const int r0 = rt * p.RT, y0 = yt * p.YT, n0 = blockIdx.y * p.NT;

    int total_chunks = 0;
    for (int s = 0; s < p.nseg; ++s) total_chunks += p.seg_nkx[s] * p.seg_nck[s];

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stagesA; ++i) { mbar_init(fullA + 8 * i, 1); mbar_init(emptyA + 8 * i, 1); }
        for (int i = 0; i < p.stagesB; ++i) { mbar_init(fullB + 8 * i, 1); mbar_init(emptyB + 8 * i, 1); }
        mbar_init(accum, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            int chunk = 0;
            for (int s = 0; s < p.nseg; ++s) {
                const CUtensorMap* tm = &p.tmA[s];
                for (int kx = 0; kx < p.seg_nkx[s]; ++kx) {
                    for (int ck = 0; ck < p.seg_nck[s]; ++ck, ++chunk) {
                        mbar_wait(emptyA + 8 * sa, pa ^ 1);
                        mbar_expect_tx(fullA + 8 * sa, bytesA);
                        tma_load_4d(sA0 + bytesA * sa, tm, fullA + 8 * sa, ck * 64, r0 + p.seg_c1off[s] + kx * p.seg_c1step[s],
                                    y0 - p.padY, frame);
                        if (++sa == p.stagesA) { sa = 0; pa ^= 1; }
                        for (int dy = 0; dy < p.KH; ++dy) {
                            mbar_wait(emptyB + 8 * sb, pb ^ 1);
                            mbar_expect_tx(fullB + 8 * sb, bytesB);
                            tma_load_2d(sB0 + bytesB * sb, &p.tmB, fullB + 8 * sb, 0, (chunk * p.KH + dy) * p.Ntot_pad + n0);
                            if (++sb == p.stagesB) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((128u >> 4) << 24);
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            uint32_t acc = 0;
            for (int s = 0; s < p.nseg; ++s) {
                for (int kx = 0; kx < p.seg_nkx[s]; ++kx) {
                    for (int ck = 0; ck < p.seg_nck[s]; ++ck) {
                        const int ksteps = ((ck == p.seg_nck[s] - 1) ? p.seg_klast[s] : 64) >> 4;
                        mbar_wait(fullA + 8 * sa, pa);
                        for (int dy = 0; dy < p.KH; ++dy) {
                            mbar_wait(fullB + 8 * sb, pb);
                            tc_fence_after();
                            const uint32_t a_addr = sA0 + bytesA * sa + (uint32_t)dy * p.RT * 128u;
                            const uint32_t b_addr = sB0 + bytesB * sb;
                            for (int k = 0; k < ksteps; ++k) {
                                tc_mma_bf16(tmem_base, make_desc_sw128(a_addr + k * 32), make_desc_sw128(b_addr + k * 32), idesc, acc);
                                acc = 1;
                            }
                            tc_commit(emptyB + 8 * sb);
                            if (++sb == p.stagesB) { sb = 0; pb ^= 1; }
                        }
                        tc_commit(emptyA + 8 * sa);
                        if (++sa == p.stagesA) { sa = 0; pa ^= 1; }
                    }
                }
            }
            tc_commit(accum);
        }
    } else {
        // ===================== epilogue (warps 2..5) =====================
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int m = q * 32 + lane;
        const int yy = m / p.RT, rr = m % p.RT;
        const int y = y0 + yy, r = r0 + rr;
        const bool row_ok = (y < p.Hin) && (r < p.nR);
        mbar_wait(accum, 0);
        tc_fence_after();
        const bool vec8 = (p.Cout % 8) == 0;
        for (int j0 = 0; j0 < p.NT; j0 += 16) {
            uint32_t v[16];
            tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)j0, v);
            if (!row_ok) continue;
            if (vec8) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const int n = n0 + j0 + g * 8;
                    if (n >= p.Ntot) continue;
                    const int grp = n / p.Cout, co = n - grp * p.Cout;
                    const int sy = grp / p.Sx, sx = grp - sy * p.Sx;
                    const int oy = p.Sy * y + sy, ox = p.Sx * r + sx;
                    if (oy >= p.out_H || ox >= p.out_W) continue;
                    const long long off = (long long)frame * p.out_sn + (long long)oy * p.out_sy + (long long)(ox + p.out_padx) * p.out_sx + p.out_coff + co;
                    float f[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float x = __uint_as_float(v[g * 8 + i]) + __ldg(p.bias + n + i);
                        f[i] = p.act == 1 ? gelu_erf(x) : x;
                    }
                    if (p.out_f32) {
                        float4* o = (float4*)((float*)p.out + off);
                        o[0] = make_float4(f[0], f[1], f[2], f[3]); o[1] = make_float4(f[4], f[5], f[6], f[7]);
                    } else {
                        __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
                        __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
                        uint4 u;
                        u.x = *(uint32_t*)&h0; u.y = *(uint32_t*)&h1; u.z = *(uint32_t*)&h2; u.w = *(uint32_t*)&h3;
                        *(uint4*)((__nv_bfloat16*)p.out + off) = u;
                    }
                }
            } else {
                for (int i = 0; i < 16; ++i) {
                    const int n = n0 + j0 + i;
                    if (n >= p.Ntot) break;
                    const int grp = n / p.Cout, co = n - grp * p.Cout;
                    const int sy = grp / p.Sx, sx = grp - sy * p.Sx;
                    const int oy = p.Sy * y + sy, ox = p.Sx * r + sx;
                    if (oy >= p.out_H || ox >= p.out_W) continue;
                    const long long off = (long long)frame * p.out_sn + (long long)oy * p.out_sy + (long long)(ox + p.out_padx) * p.out_sx + p.out_coff + co;
                    float x = __uint_as_float(v[i]) + __ldg(p.bias + n);
                    x = p.act == 1 ? gelu_erf(x) : x;
                    if (p.out_f32) ((float*)p.out)[off] = x;
                    else ((__nv_bfloat16*)p.out)[off] = __float2bfloat16_rn(x);
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// host side
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = (PFN_encodeTiled)sym;
    }
    return fn;
}

static int encode_map(CUtensorMap* tm, void* base, int rank, const unsigned long long* dims, const unsigned long long* strides_bytes,
                      const unsigned* box) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) { fprintf(stderr, "[accessmath_b200] cuTensorMapEncodeTiled unavailable\n"); return AM_ERR_CUDA; }
    cuuint64_t gd[5]; cuuint64_t gs[4]; cuuint32_t bx[5]; cuuint32_t es[5];
    for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i < rank - 1; ++i) gs[i] = strides_bytes[i];
    CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, base, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        fprintf(stderr, "[accessmath_b200] cuTensorMapEncodeTiled failed (%d): rank %d dims %llu %llu %llu %llu strides %llu %llu %llu box %u %u %u %u\n",
                (int)r, rank, dims[0], dims[1], rank > 2 ? dims[2] : 0ull, rank > 3 ? dims[3] : 0ull, strides_bytes[0],
                rank > 2 ? strides_bytes[1] : 0ull, rank > 3 ? strides_bytes[2] : 0ull, box[0], box[1], rank > 2 ? box[2] : 0u, rank > 3 ? box[3] : 0u);
        return AM_ERR_CUDA;
    }
    return AM_OK;
}

// One convolution-as-GEMM launch.  Everything the kernel needs is in this plain-C descriptor (see header).
extern "C" int am_conv_gemm(const am_conv_desc* d, void* stream) {
    if (!d || d->nseg < 1 || d->nseg > 2 || d->RT % 8 != 0 || d->RT * d->YT != 128 || d->NT % 16 != 0 || d->NT < 16 || d->NT > 256)
        return AM_ERR_ARG;
    ConvParams p;
    memset(&p, 0, sizeof(p));
    p.nseg = d->nseg;
    const int box_rows = d->YT + d->KH - 1;
    int total_chunks = 0;
    for (int s = 0; s < d->nseg; ++s) {
        const am_conv_seg* g = &d->seg[s];
        unsigned long long dims[4], strides[3]; unsigned box[4];
        if (g->rowrun) {        // dim0 = run elements, dim1 = output group r (row stride S*C elements: rows overlap)
            dims[0] = (unsigned long long)g->run_len; dims[1] = (unsigned long long)d->nR; dims[2] = (unsigned long long)d->Hin; dims[3] = (unsigned long long)d->batch;
            strides[0] = (unsigned long long)g->S * g->C * 2ull;
            p.seg_nkx[s] = 1; p.seg_c1step[s] = 0; p.seg_c1off[s] = 0;
            p.seg_nck[s] = (g->run_len + 63) / 64;
            p.seg_klast[s] = ((g->run_len - (p.seg_nck[s] - 1) * 64) + 15) & ~15;
        } else {                // dim0 = channels, dim1 = padded x; one load per horizontal tap
            dims[0] = (unsigned long long)g->C; dims[1] = (unsigned long long)g->Wp; dims[2] = (unsigned long long)d->Hin; dims[3] = (unsigned long long)d->batch;
            strides[0] = (unsigned long long)g->C * 2ull;
            p.seg_nkx[s] = g->KW; p.seg_c1step[s] = 1; p.seg_c1off[s] = 0;
            p.seg_nck[s] = (g->C + 63) / 64;
            p.seg_klast[s] = ((g->C - (p.seg_nck[s] - 1) * 64) + 15) & ~15;
        }
        strides[1] = (unsigned long long)g->Wp * g->C * 2ull;
        strides[2] = strides[1] * (unsigned long long)g->Hbuf;
        box[0] = 64; box[1] = (unsigned)d->RT; box[2] = (unsigned)box_rows; box[3] = 1;
        // base: first element of the window of output x = 0 (buffer pad minus conv pad)
        char* base = (char*)g->ptr + (long long)g->x_off * g->C * 2ll;
        int rc = encode_map(&p.tmA[s], base, 4, dims, strides, box);
        if (rc) return rc;
        total_chunks += p.seg_nkx[s] * p.seg_nck[s];
    }
    {
        unsigned long long dims[2] = {64ull, (unsigned long long)total_chunks * d->KH * d->Ntot_pad};
        unsigned long long strides[1] = {128ull};
        unsigned box[2] = {64u, (unsigned)d->NT};
        int rc = encode_map(&p.tmB, (void*)d->weights, 2, dims, strides, box);
        if (rc) return rc;
    }
    p.KH = d->KH; p.RT = d->RT; p.YT = d->YT; p.padY = d->padY;
    p.nRT = (d->nR + d->RT - 1) / d->RT; p.nYT = (d->Hin + d->YT - 1) / d->YT;
    p.nR = d->nR; p.Hin = d->Hin; p.NT = d->NT; p.Ntot_pad = d->Ntot_pad;
    p.tmem_cols = 32; while (p.tmem_cols < d->NT) p.tmem_cols <<= 1;
    p.out = d->out; p.out_f32 = d->out_f32; p.out_H = d->out_H; p.out_W = d->out_W;
    p.out_sn = d->out_sn; p.out_sy = d->out_sy; p.out_sx = d->out_sx; p.out_padx = d->out_padx; p.out_coff = d->out_coff;
    p.Cout = d->Cout; p.Sy = d->Sy; p.Sx = d->Sx; p.Ntot = d->Ntot; p.act = d->act; p.bias = d->bias;
    // shared memory budget: A ring + B ring + barriers
    const size_t bytesA = (size_t)box_rows * d->RT * 128, bytesB = (size_t)d->NT * 128;
    const size_t budget = 200 * 1024;
    int sa = 2, sb = 2;
    while (true) {      // grow the rings alternately while they fit; B stages are consumed KH times faster
        bool grew = false;
        if (sb < 3 * d->KH && sb < 12 && bytesA * sa + bytesB * (sb + 1) + 2048 <= budget) { ++sb; grew = true; }
        if (sa < 4 && bytesA * (sa + 1) + bytesB * sb + 2048 <= budget) { ++sa; grew = true; }
        if (!grew) break;
    }
    if (bytesA * sa + bytesB * sb + 2048 > 227 * 1024) return AM_ERR_ARG;
    p.stagesA = sa; p.stagesB = sb;
    const size_t smem = 1024 + bytesA * sa + ((bytesB * sb + 1023) & ~(size_t)1023) + 1024;
    static size_t smem_set = 0;
    if (smem > smem_set) {
        AM_CUDA(cudaFuncSetAttribute(k_conv_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        smem_set = 227 * 1024;
    }
    dim3 grid((unsigned)(p.nRT * p.nYT * d->batch), (unsigned)(d->Ntot_pad / d->NT));
    k_conv_gemm<<<grid, 192, smem, (cudaStream_t)stream>>>(p);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

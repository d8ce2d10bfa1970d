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
// Execution model (v2): PERSISTENT CTAs (grid = min(work items, SM count)), 10 warps:
//   warp 0      TMA producer (A boxes per K chunk; B tiles per (chunk, dy) unless the weights are resident)
//   warp 1      TMEM owner + tcgen05.mma issuer of M-tile 0
//   warp 2      tcgen05.mma issuer of M-tile 1 (kMT = 2): a single issuing warp is latency bound (~100 clk per MMA
//               measured with N = 64), two issuers on two independent accumulators double the issue rate
//   warp 3      idle
//   warps 4..19 epilogue: four warps per TMEM lane quarter, interleaved 16-column groups
// A work item = MT (1 or 2) M-tiles of 128 GEMM rows x one N block of NT columns.  MT = 2 shares every B tile
// between two accumulators (halves the L2->smem weight traffic of the streamed mode).  Weights that fit in
// shared memory are loaded ONCE per CTA ("resident B") and the B ring disappears.  Accumulators are double
// buffered in TMEM when 2*MT*NTc <= 512 columns so the epilogue of tile i overlaps the main loop of tile i+1.
#include "am_common.cuh"
#include "../../include/accessmath_b200.h"
#include <cuda.h>
#include <cuda_bf16.h>

#ifndef EPI_WARPS
#define EPI_WARPS 16                      // 4 epilogue warps per TMEM lane quarter (8 measured ~3 % slower per step: the
#endif                                    // epilogue is latency bound -- TMEM load, MUFU -- so more warps hide more of it)
#define EPI_PER_Q (EPI_WARPS / 4)
#define CONV_THREADS (128 + 32 * EPI_WARPS)
#define SCHED_DEPTH 4                     // work-item ring between the producer (fetches with atomicAdd) and its consumers

struct alignas(64) ConvParams {
    CUtensorMap tmA[2];
    CUtensorMap tmB;
    CUtensorMap tmBq;    // CTA pairs with 2-D packing: box of NT/4 weight rows (one CTA's share of an edge tap, see edge_half)
    int edge_half;       // 1: the first / last Toeplitz tap in y only feeds the sy = 0 / sy = 1 half of the columns -> N/2 MMAs
    int nseg;
    int seg_nkx[2];      // horizontal taps enumerated by separate loads (1 in row-run mode)
    int seg_nck[2];      // 64-element K chunks per (segment, kx)
    int seg_klast[2];    // valid K (multiple of 16) of the last chunk
    int seg_c1step[2];   // coordinate-1 step per kx (0 in row-run mode)
    int seg_c1off[2];    // coordinate-1 offset
    int KH, RT, YT, padY, logRT;
    int ystep;           // input rows per GEMM row step (2 = 2-D packing: UMMA M-atoms skip every other image row)
    int nRT, nYT, batch; // tiles per row / per frame column, frames
    int nR, Hin;         // valid groups per row, valid rows
    int NT, NTc;         // UMMA N of this launch, TMEM columns per accumulator (power of two >= NT)
    int Ntot_pad, nNB;   // rows per (chunk,dy) block of the packed weights, N blocks
    int MT, acc_stages, residentB, total_chunks;
    int n_mtiles, n_work;
    int tmem_cols;
    int stagesA, stagesB;
    // epilogue
    void* out; int out_f32;
    int out_H, out_W;
    long long out_sn, out_sy; int out_sx, out_padx, out_coff;
    int Cout, Sy, Sx, Ntot, act;
    unsigned cout_magic;  // floor(2^32 / Cout) + 1 (Cout >= 2)
    void* pool_out; int pool_H, pool_W, pool_sx, pool_padx; long long pool_sn, pool_sy;   // fused MaxPool2d(2) destination (NULL = off)
    int pool2;            // fused MaxPool2d(2) with Sx = Sy = 2 packing: the 2x2 block of a pooled pixel is ONE GEMM row (see kPOOL2)
    int poolx;            // fused MaxPool2d(2) with Sx even, Sy = 1: x pairs in two column units of one warp, y pairs in lanes l, l ^ RT (kPOOLX)
    int epi_per_q;        // epilogue warps per TMEM lane quarter that take part in this launch (<= compiled EPI_WARPS / 4)
    unsigned nrt_magic, nyt_magic, nnb_magic;   // floor(2^32 / d) + 1 for d = nRT, nYT, nNB (0 when d == 1): exact for n * d < 2^32
    const float* bias;
    int* work_counter;    // [0] next work item (dynamic tile scheduler), [1] CTAs finished (the last one resets both)
    // fused epilogues of the fp32 layers (am_conv_desc.epi_mode)
    int epi_mode;
    const uint8_t* frames; int img_H, img_W;
    __nv_bfloat16* diff_out; int diff_C, diff_pad;
    float* text_out; float* rec_out;
    uint16_t* bits_out; int bits_hpr, threshold;      // halfwords per mask row
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done;
}
// try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (wake-up is immediate) or the hint
// expires, instead of re-issuing the poll loop every few hundred cycles -- on the epilogue-bound layers the waiting producer /
// issuer warps were spending 12 % of the SM's issue slots in that loop (profile r01_n, op 0)
__device__ __forceinline__ uint32_t mbar_try_sleep(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
    return done;
}
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity) {
    for (uint32_t it = 0; !mbar_try_sleep(bar, parity); ++it)
        if (it > (1u << 20)) {                       // >= 0.2 s of polling, ~20 s when every poll sleeps its 20 us hint
            printf("[accessmath_b200] mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    if (!mbar_try(bar, parity)) mbar_wait_slow(bar, parity);
}
// for the warp-uniform producer / issuer loops: the vote makes the branch provably uniform, so ptxas keeps the loop
// state in uniform registers instead of moving it R->UR before every UTMALDG / UTCHMMA
__device__ __forceinline__ void mbar_wait_uniform(uint32_t bar, uint32_t parity) {
    if (!__all_sync(0xffffffffu, mbar_try(bar, parity))) mbar_wait_slow(bar, parity);
}
__device__ __forceinline__ void mbar_wait_old(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (uint32_t it = 0; !done; ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (it > (1u << 20)) {                       // >= 0.2 s of polling, ~20 s when every poll sleeps its 20 us hint
            printf("[accessmath_b200] mbarrier timeout (block %d thread %d)\n", blockIdx.x, threadIdx.x); __trap(); }
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
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((unsigned long long)tm) : "memory");
}
// Programmatic dependent launch: a conv kernel launched with programmaticStreamSerialization may start its prologue (barrier init,
// TMEM allocation, tensor-map prefetch, resident weights) while the tail of its predecessor still runs on other SMs; nothing the
// predecessor wrote is read and nothing is written before pdl_wait() returns (= predecessor complete and flushed).  Both are no-ops
// in a launch without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
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
// One leader lane of a fully active warp (same lane on every call).  Producer and MMA loops run WARP-UNIFORM (all 32
// lanes execute the control flow, so addresses/descriptors live in uniform registers) and only the asynchronous
// instruction itself is predicated on the elected lane -- a `lane == 0` region would make ptxas wrap every
// UTMALDG/UTCHMMA in an R2UR + ELECT waterfall loop (seen in profiles/ r01: ~220 issue cycles per MMA).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, 0xffffffff;\n\t@px mov.s32 %0, 1;\n\t}" : "+r"(pred));
    return pred != 0;
}
// descriptor bits 32..63: stride byte offset between 8-row M/N atoms (1024 B; 2048 B for A when every other image row is
// skipped), version, SWIZZLE_128B
__device__ __forceinline__ uint64_t desc_hi_sw128(uint32_t sbo_bytes) { return ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr >> 4) & 0x3FFFu) | (1u << 16); }

// 16 consecutive fp32 accumulator columns of this thread's TMEM lane (asynchronous: tmem_wait before use)
__device__ __forceinline__ void tmem_ld16_async(uint32_t taddr, uint32_t* v) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
// wait for outstanding tcgen05.ld; the registers are in/out operands so no use of them can be hoisted above the wait
__device__ __forceinline__ void tmem_wait16(uint32_t* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                   "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
                 :: "memory");
}

// nn.GELU() (exact-erf form) = x * Phi(x) = 0.5 x (1 + tanh z(x)) with z(x) = atanh(erf(x / sqrt 2)), fitted as
// z = x (a + b s + c s^2), s = min(x^2, 36) (minimax fit, tools/fit_gelu.py: |error| <= 2.6e-5 absolute for every fp32
// x; beyond |x| = 6 the clamp keeps z monotone and the tanh saturated).  tanh is ONE MUFU op (tanh.approx.f32, relative
// error 2^-11 => |error| <= 2.5e-4 |x| on top of the fit, an order of magnitude below the bf16 rounding of the stored
// activation); measured end to end (oracle/check_fcn_accuracy.py, 1080p, fp32 oracle): max probability error 5.4e-4 and mask
// disagreement 2.4e-5 - 4.4e-5, identical to the Abramowitz-Stegun 7.1.28 erf (13 FP32 ops + 1 MUFU) used before.
// The epilogue warps are issue bound on the K <= 96 layers, so the arithmetic runs on PACKED fp32 pairs
// (fma/mul/add.rn.f32x2: two elements per issue slot on sm_100): 3.5 FMA-pipe + 1 ALU (min) + 1 MUFU per element.
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) { uint64_t d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
#define GELU_A 0.79750788f
#define GELU_B 0.037005651f
#define GELU_C (-0.00035151753f)
__device__ __forceinline__ float tanh_approx(float z) { float t; asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(z)); return t; }
__device__ __forceinline__ float gelu_erf(float x) {
#ifdef GELU_SIGMOID   // x * sigmoid(2z): no MUFU.TANH error, two MUFU ops (ex2, rcp); coefficients carry the -2 log2(e)
    const float s2 = fminf(x * x, 36.0f);
    float q2 = fmaf(0.0010142652f, s2, -0.10677574f);
    q2 = fmaf(q2, s2, -2.3011212f);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * q2));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
    return x * r;
#else
    const float s = fminf(x * x, 36.0f);
    float q = fmaf(GELU_C, s, GELU_B);
    q = fmaf(q, s, GELU_A);
    const float hx = 0.5f * x;
    return fmaf(hx, tanh_approx(x * q), hx);
#endif
}
// the same on a packed pair
__device__ __forceinline__ uint64_t gelu_erf2(uint64_t x) {
#ifdef GELU_2TERM     // experiment: z = x (a + b x^2), no clamp: |fit error| <= 2.7e-4
    float y0, y1;
    const uint64_t qq = f2_fma(f2_mul(x, x), f2_pack(0.03470092f, 0.03470092f), f2_pack(0.80015702f, 0.80015702f));
    f2_unpack(f2_mul(x, qq), y0, y1);
    const uint64_t hh = f2_mul(x, f2_pack(0.5f, 0.5f));
    return f2_fma(hh, f2_pack(tanh_approx(y0), tanh_approx(y1)), hh);
#endif
    float s0, s1, z0, z1;
    f2_unpack(f2_mul(x, x), s0, s1);
    const uint64_t s = f2_pack(fminf(s0, 36.0f), fminf(s1, 36.0f));
    uint64_t q = f2_fma(f2_pack(GELU_C, GELU_C), s, f2_pack(GELU_B, GELU_B));
    q = f2_fma(q, s, f2_pack(GELU_A, GELU_A));
    f2_unpack(f2_mul(x, q), z0, z1);
    const uint64_t hx = f2_mul(x, f2_pack(0.5f, 0.5f));
    return f2_fma(hx, f2_pack(tanh_approx(z0), tanh_approx(z1)), hx);
}

// n / d through the precomputed reciprocal (magic = 0 encodes d == 1); exact for n * d < 2^32 (work-item counts are < 2^20)
__device__ __forceinline__ int fast_div(int n, unsigned magic) { return magic ? (int)__umulhi((unsigned)n, magic) : n; }
struct TileCoord { int frame, y0, r0; bool valid; };
__device__ __forceinline__ TileCoord decode_tile(const ConvParams& p, int t) {
    TileCoord c;
    c.valid = t < p.n_mtiles;
    if (!c.valid) t = p.n_mtiles - 1;                // duplicate the last tile: loads stay in bounds, stores are skipped
    const int t1 = fast_div(t, p.nrt_magic), rt = t - t1 * p.nRT;      // t / nRT, t % nRT without the integer-division sequence
    const int fr = fast_div(t1, p.nyt_magic), yt = t1 - fr * p.nYT;
    c.frame = fr; c.y0 = yt * p.YT; c.r0 = rt * p.RT;
    return c;
}

// Fused epilogues of the two fp32 layers.  They live in their own kernel instantiations (template parameter kFUSED): compiled into the
// common kernels their registers (expf / tanhf, 16 live logits) spilled the GELU path of k_conv_gemm<1, true>, which is issue bound on
// the K <= 640 layers (conv_down_block_1 and the transposed convs ran 20 % slower, profile r02_b).
__device__ __forceinline__ void epi_heads(const ConvParams& p, const uint32_t (&v)[16], const float* __restrict__ sbias, int n0, int j0,
                                          long long base, long long pbase, bool sy1_ok) {
    // Cout = 4 columns per pixel: (text logit, reconstruction pre-tanh R, G, B).  diff = (x0 - tanh(rec)) * sigmoid(text)
    // (FCN_lecturenet.py:370-377) goes straight to the bf16 `diff` buffer; the fp32 heads never reach HBM.
    // `base` = pixel index (frame*H + Sy*y) * W + Sx*r of the unit's row (out_sx = 1 in this mode), `pbase` = frame*H + Sy*y.
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int n = n0 + j0 + 4 * i;
        if (n >= p.Ntot) break;
        const int grp = n >> 2;
        const int sy = (p.Sy == 2 && grp >= p.Sx) ? 1 : 0, sx = grp - sy * p.Sx;
        if (sy != 0 && !sy1_ok) continue;
        const long long row = pbase + sy;                                          // frame*H + y
        const int x = (int)(base - pbase * p.img_W) + sx;
        const long long pix = row * p.img_W + x;
        const float t = __uint_as_float(v[4 * i]) + sbias[j0 + 4 * i];
        const float r0 = tanhf(__uint_as_float(v[4 * i + 1]) + sbias[j0 + 4 * i + 1]);
        const float r1 = tanhf(__uint_as_float(v[4 * i + 2]) + sbias[j0 + 4 * i + 2]);
        const float r2 = tanhf(__uint_as_float(v[4 * i + 3]) + sbias[j0 + 4 * i + 3]);
        const float sg = 1.0f / (1.0f + expf(-t));                                   // torch.sigmoid(text_mask), :372
        const uint8_t* px = p.frames + pix * 3;
        const float x0r = ((float)px[2] / 255.0f - 0.5f) / 0.5f, x0g = ((float)px[1] / 255.0f - 0.5f) / 0.5f,
                    x0b = ((float)px[0] / 255.0f - 0.5f) / 0.5f;                    // BGR -> RGB, prepare_image :607-618
        const __nv_bfloat162 h0 = __floats2bfloat162_rn((x0r - r0) * sg, (x0g - r1) * sg);
        const __nv_bfloat162 h1 = __floats2bfloat162_rn((x0b - r2) * sg, 0.0f);
        __nv_bfloat16* o = p.diff_out + (row * (p.img_W + 2 * p.diff_pad) + x + p.diff_pad) * p.diff_C;
        if (p.diff_C == 4) *(uint2*)o = make_uint2(*(const uint32_t*)&h0, *(const uint32_t*)&h1);
        else *(uint4*)o = make_uint4(*(const uint32_t*)&h0, *(const uint32_t*)&h1, 0u, 0u);
        if (p.text_out) p.text_out[pix] = t;
        if (p.rec_out) { p.rec_out[pix * 3] = r0; p.rec_out[pix * 3 + 1] = r1; p.rec_out[pix * 3 + 2] = r2; }
    }
}
__device__ __forceinline__ void epi_threshold(const ConvParams& p, const uint32_t (&v)[16], const float* __restrict__ sbias, int n0, int j0,
                                              long long base, long long pbase, bool sy1_ok) {
    // Cout = 1, Sx % 16 == 0: the unit's 16 columns are 16 consecutive pixels of one image row -> one 16-bit store of the
    // bit-packed ink mask: ink <=> (uint8)(sigmoid(z) * 255) < threshold (FCN_lecturenet.py:461-467 + `255 - binary`)
    const int n = n0 + j0;
    if (n >= p.Ntot) return;
    const int sy = (p.Sy == 2 && n >= p.Sx) ? 1 : 0, sx = n - sy * p.Sx;
    if (sy != 0 && !sy1_ok) return;
    const long long row = pbase + sy;                                              // frame*H + y; base = pixel index of (frame, Sy*y, Sx*r)
    const int x = (int)(base - pbase * p.img_W) + sx;
    const long long pix = row * p.img_W + x;
    unsigned m = 0;
    float* o = p.out != nullptr ? (float*)p.out + pix : nullptr;
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
        float z[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            z[k] = __uint_as_float(v[i + k]) + sbias[j0 + i + k];
            const float sg = 1.0f / (1.0f + expf(-z[k]));
            m |= ((int)(sg * 255.0f) < p.threshold ? 1u : 0u) << (i + k);
        }
        if (o) *(float4*)(o + i) = make_float4(z[0], z[1], z[2], z[3]);
    }
    p.bits_out[row * p.bits_hpr + (x >> 4)] = (uint16_t)m;
}

// Epilogue of one 16-column unit of one accumulator row: bias + activation + NHWC store.
// With a fused max-pool (p.pool_out) EVERY lane of the warp calls this (row_ok only predicates the stores): the 2x2 block of an
// output pixel sits in lanes l, l^1 (x) and l^RT (y), so the pooled value is two shuffle + max rounds on the packed bf16 pairs.
// kPX (kPOOLX kernels: Sx even, Sy = 1): the two x neighbours of a pooled pixel are the same 16-channel slice of two column units that
// one warp visits back to back -- kPX = 1 keeps the first unit's packed bf16 results in `carry`, kPX = 2 maxes them with the second
// unit's, then with the row below (lane ^ RT) and stores the pooled pixel.
template <bool kFUSED, bool kPOOL2 = false, int kPX = 0>
__device__ __forceinline__ void epi_unit(const ConvParams& p, const uint32_t (&v)[16], const float* __restrict__ sbias, int n0, int j0,
                                         long long base, bool vec16, bool vec8, bool f32fast, bool sy1_ok, bool row_ok = true,
                                         long long pbase = 0, bool pool_ok = false, uint4* spool = nullptr, uint32_t* carry = nullptr) {
    if (kFUSED) {                                          // the instantiations the two fp32 layers are launched with
        if (p.epi_mode == AM_EPI_HEADS) epi_heads(p, v, sbias, n0, j0, base, pbase, sy1_ok);
        else epi_threshold(p, v, sbias, n0, j0, base, pbase, sy1_ok);
        return;
    }
    if (vec16) {
        // Cout % 16 == 0: the unit's 16 columns are 16 consecutive channels of ONE output pixel -> one address
        // computation and one 32-byte store per unit (every GELU layer of the network takes this path)
        const int n = n0 + j0;
        const int grp = (int)__umulhi((unsigned)n, p.cout_magic), co = n - grp * p.Cout;
        const int sy = (p.Sy == 2 && grp >= p.Sx) ? 1 : 0, sx = grp - sy * p.Sx;
        if (n >= p.Ntot || (sy != 0 && !sy1_ok)) return;
        const long long off = base + (long long)sy * p.out_sy + (long long)sx * p.out_sx + co;
        uint32_t pk[8];
        uint64_t a[8];
#pragma unroll
        for (int i = 0; i < 8; i += 2) {
            const float4 b = *(const float4*)(sbias + j0 + 2 * i);
            a[i] = f2_add(f2_pack(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1])), f2_pack(b.x, b.y));
            a[i + 1] = f2_add(f2_pack(__uint_as_float(v[2 * i + 2]), __uint_as_float(v[2 * i + 3])), f2_pack(b.z, b.w));
        }
        if (p.act == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = gelu_erf2(a[i]);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float f0, f1;
            f2_unpack(a[i], f0, f1);
            const __nv_bfloat162 h = __floats2bfloat162_rn(f0, f1);
            pk[i] = *(const uint32_t*)&h;
        }
        __nv_bfloat16* o = (__nv_bfloat16*)p.out + off;
        if (row_ok) {
            if ((off & 15) == 0) {
                asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]),
                             "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
            } else {
                *(uint4*)o = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                *(uint4*)(o + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            }
        }
        if (kPX == 1) {
#pragma unroll
            for (int i = 0; i < 8; ++i) carry[i] = pk[i];
        } else if (kPX == 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                __nv_bfloat162 m = __hmax2(*(const __nv_bfloat162*)&pk[i], *(const __nv_bfloat162*)&carry[i]);
                const uint32_t t = __shfl_xor_sync(0xffffffffu, *(const uint32_t*)&m, p.RT);
                m = __hmax2(m, *(const __nv_bfloat162*)&t);
                pk[i] = *(const uint32_t*)&m;
            }
            if (pool_ok) {                                 // even rows: pooled pixel (y / 2, Sx / 2 * r + sx / 2)
                __nv_bfloat16* po = (__nv_bfloat16*)p.pool_out + pbase + (long long)(sx >> 1) * p.pool_sx + co;
                if ((((uintptr_t)po) & 31) == 0) {
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(po), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]),
                                 "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                } else {
                    *(uint4*)po = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *(uint4*)(po + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
        } else if (kPOOL2) {                               // this lane's 16 activated channels of one (sy, sx) group -> shared memory
            spool[0] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            spool[1] = make_uint4(pk[4], pk[5], pk[6], pk[7]);
        } else if (p.pool_out != nullptr) {                // warp-uniform
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                uint32_t t = __shfl_xor_sync(0xffffffffu, pk[i], 1);
                __nv_bfloat162 m = __hmax2(*(const __nv_bfloat162*)&pk[i], *(const __nv_bfloat162*)&t);
                uint32_t mu = *(const uint32_t*)&m;
                t = __shfl_xor_sync(0xffffffffu, mu, p.RT);
                m = __hmax2(m, *(const __nv_bfloat162*)&t);
                pk[i] = *(const uint32_t*)&m;
            }
            if (pool_ok) {
                __nv_bfloat16* po = (__nv_bfloat16*)p.pool_out + pbase + co;
                if ((((uintptr_t)po) & 31) == 0) {
                    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(po), "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]),
                                 "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7]) : "memory");
                } else {
                    *(uint4*)po = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    *(uint4*)(po + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
                }
            }
        }
    } else if (vec8) {
        // two 8-channel groups; a group never straddles a pixel because Cout % 8 == 0
        uint4 pk[2]; long long offs[2]; bool ok[2];
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
            const int n = n0 + j0 + gg * 8;
            const int grp = (int)__umulhi((unsigned)n, p.cout_magic), co = n - grp * p.Cout;
            const int sy = (p.Sy == 2 && grp >= p.Sx) ? 1 : 0, sx = grp - sy * p.Sx;
            ok[gg] = n < p.Ntot && (sy == 0 || sy1_ok);
            offs[gg] = base + (long long)sy * p.out_sy + (long long)sx * p.out_sx + co;
            const float4 b0 = *(const float4*)(sbias + j0 + gg * 8), b1 = *(const float4*)(sbias + j0 + gg * 8 + 4);
            float f[8] = {__uint_as_float(v[gg * 8 + 0]) + b0.x, __uint_as_float(v[gg * 8 + 1]) + b0.y, __uint_as_float(v[gg * 8 + 2]) + b0.z,
                          __uint_as_float(v[gg * 8 + 3]) + b0.w, __uint_as_float(v[gg * 8 + 4]) + b1.x, __uint_as_float(v[gg * 8 + 5]) + b1.y,
                          __uint_as_float(v[gg * 8 + 6]) + b1.z, __uint_as_float(v[gg * 8 + 7]) + b1.w};
            if (p.act == 1) {
#pragma unroll
                for (int i = 0; i < 8; ++i) f[i] = gelu_erf(f[i]);
            }
            __nv_bfloat162 h0 = __floats2bfloat162_rn(f[0], f[1]), h1 = __floats2bfloat162_rn(f[2], f[3]);
            __nv_bfloat162 h2 = __floats2bfloat162_rn(f[4], f[5]), h3 = __floats2bfloat162_rn(f[6], f[7]);
            pk[gg].x = *(uint32_t*)&h0; pk[gg].y = *(uint32_t*)&h1; pk[gg].z = *(uint32_t*)&h2; pk[gg].w = *(uint32_t*)&h3;
        }
        __nv_bfloat16* o = (__nv_bfloat16*)p.out;
        if (ok[0] && ok[1] && offs[1] == offs[0] + 8 && ((offs[0] & 15) == 0)) {
            // 32 contiguous, 32-byte aligned bytes: one full sector per thread (STG.256)
            asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(o + offs[0]), "r"(pk[0].x), "r"(pk[0].y), "r"(pk[0].z),
                         "r"(pk[0].w), "r"(pk[1].x), "r"(pk[1].y), "r"(pk[1].z), "r"(pk[1].w) : "memory");
        } else {
            if (ok[0]) *(uint4*)(o + offs[0]) = pk[0];
            if (ok[1]) *(uint4*)(o + offs[1]) = pk[1];
        }
    } else if (f32fast) {
        // fp32 heads / logits: the N columns of one output row are contiguous in memory (n = sx*Cout + co); with 2-D packing
        // the second half of the columns is the next image row (a 16-column unit never straddles: Sx*Cout % 16 == 0)
        const int n = n0 + j0, half = p.Sx * p.Cout;
        const int sy = (p.Sy == 2 && n >= half) ? 1 : 0;
        if (sy != 0 && !sy1_ok) return;
        float* o = (float*)p.out + base + (long long)sy * p.out_sy + (n - sy * half);
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
            if (n + i >= p.Ntot) break;
            float4 f;
            f.x = __uint_as_float(v[i]) + sbias[j0 + i]; f.y = __uint_as_float(v[i + 1]) + sbias[j0 + i + 1];
            f.z = __uint_as_float(v[i + 2]) + sbias[j0 + i + 2]; f.w = __uint_as_float(v[i + 3]) + sbias[j0 + i + 3];
            if (p.act == 1) { f.x = gelu_erf(f.x); f.y = gelu_erf(f.y); f.z = gelu_erf(f.z); f.w = gelu_erf(f.w); }
            *(float4*)(o + i) = f;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const int n = n0 + j0 + i;
            if (n >= p.Ntot) break;
            const int grp = n / p.Cout, co = n - grp * p.Cout;
            const int sy = grp / p.Sx, sx = grp - sy * p.Sx;
            if (sy > 0 && !sy1_ok) continue;
            const long long off = base + (long long)sy * p.out_sy + (long long)sx * p.out_sx + co;
            float x = __uint_as_float(v[i]) + sbias[j0 + i];
            x = p.act == 1 ? gelu_erf(x) : x;
            if (p.out_f32) ((float*)p.out)[off] = x;
            else ((__nv_bfloat16*)p.out)[off] = __float2bfloat16_rn(x);
        }
    }
}

// Dynamic tile scheduler: the producer warp draws work items from a global counter and publishes them through a small
// shared-memory ring; MMA issuers and epilogue warps consume the ring (-1 = no more work).  Unlike a static
// blockIdx-strided loop this tolerates SMs that are busy with other streams' kernels (NCCL hand-off, temporal matching).
__device__ __forceinline__ int sched_fetch(int* counter, int lane) {
    int v = 0;
    if (lane == 0) v = atomicAdd(counter, 1);
    return __shfl_sync(0xffffffffu, v, 0);
}
__device__ __forceinline__ int sched_next(uint32_t schedFull, uint32_t schedEmpty, uint32_t sched_w, int& slot, uint32_t& phase, int lane) {
    mbar_wait(schedFull + 8 * slot, phase);
    int w;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(w) : "r"(sched_w + 4 * slot) : "memory");
    __syncwarp();
    if (lane == 0) mbar_arrive(schedEmpty + 8 * slot);
    if (++slot == SCHED_DEPTH) { slot = 0; phase ^= 1; }
    return w;
}

// kMT = M-tiles per work item, kRES = weights resident in shared memory (compile-time so the single-warp issue loops stay short)
// kMT = 4 (four issuer warps, warps 1..4) takes the epilogue down to 12 warps (8..19) so that the CTA stays at 640 threads
// (96 registers each: 20 warps is what the register file holds).
// kPOOL2 (kMT = 1 only): MaxPool2d(2) fused for Sx = Sy = 2 packing (conv_down_block_1).  A GEMM row holds the whole 2x2 block of one
// pooled pixel, but as four 16-column units per channel slice that the unit interleaving hands to four DIFFERENT warps of the lane
// quarter (same lane).  Every warp parks its packed bf16 results in shared memory (32 B per unit and lane), the quarter's warps meet at
// a named barrier, then each warp takes channel slices h, h + epiPerQ, ..., maxes the four groups and stores the pooled pixel: the
// 1.5 GB re-read of the separate k_maxpool2 pass disappears for ~5 % more epilogue instructions.
template <int kMT, bool kRES, bool kFUSED, bool kPOOL2 = false, bool kPOOLX = false>
__global__ void __launch_bounds__(CONV_THREADS, 1) k_conv_gemm(const __grid_constant__ ConvParams p) {
    constexpr int kEpiFirst = (kMT == 4) ? 8 : 4;                      // first epilogue warp (multiple of 4: TMEM lane quarters)
    constexpr int kEpiWarps = (CONV_THREADS / 32) - kEpiFirst;
    constexpr int kEpiPerQ = kEpiWarps / 4;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: [A ring][B ring | resident B][bias][barriers][tmem ptr]
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bytesA1 = (uint32_t)(p.ystep * (p.YT - 1) + p.KH) * p.RT * 128u;    // one M-tile box (multiple of 1024: RT % 8 == 0)
    const uint32_t bytesA = bytesA1 * kMT;
    const uint32_t bytesB = (uint32_t)p.NT * 128u;
    const uint32_t nB = kRES ? (uint32_t)(p.total_chunks * p.KH) : (uint32_t)p.stagesB;
    const uint32_t sA0 = smem_base;
    const uint32_t sB0 = sA0 + bytesA * p.stagesA;
    const uint32_t sBias = sB0 + ((bytesB * nB + 1023u) & ~1023u);
    const uint32_t bar0 = sBias + (((uint32_t)p.NT * 4u + 127u) & ~127u);
    // barriers: fullA[sA], emptyA[sA], fullB[sB], emptyB[sB], accFull[2], accEmpty[2]
    const uint32_t fullA = bar0, emptyA = fullA + 8 * p.stagesA, fullB = emptyA + 8 * p.stagesA, emptyB = fullB + 8 * p.stagesB;
    const uint32_t accFull = emptyB + 8 * p.stagesB, accEmpty = accFull + 16;
    const uint32_t schedFull = accEmpty + 16, schedEmpty = schedFull + 8 * SCHED_DEPTH;
    const uint32_t sched_w = schedEmpty + 8 * SCHED_DEPTH;             // SCHED_DEPTH ints
    const uint32_t tmem_slot = sched_w + 4 * SCHED_DEPTH;
    const uint32_t sPool = (tmem_slot + 4u + 15u) & ~15u;             // kPOOL2: [4 quarters][NT / 16 units][32 lanes] x 32 B
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // provably warp-uniform for ptxas
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        // every MMA issuer (one per M-tile) commits to the empty / accumulator-full barriers
        for (int i = 0; i < p.stagesA; ++i) { mbar_init(fullA + 8 * i, 1); mbar_init(emptyA + 8 * i, kMT); }
        for (int i = 0; i < p.stagesB; ++i) { mbar_init(fullB + 8 * i, 1); mbar_init(emptyB + 8 * i, kMT); }
        // only the first p.epi_per_q epilogue warps of every lane quarter work in this launch (the others idle at the final barrier)
        const int n_epi = min(4 * p.epi_per_q, kEpiWarps);
        for (int i = 0; i < 2; ++i) { mbar_init(accFull + 8 * i, kMT); mbar_init(accEmpty + 8 * i, n_epi); }
        for (int i = 0; i < SCHED_DEPTH; ++i) { mbar_init(schedFull + 8 * i, 1); mbar_init(schedEmpty + 8 * i, kMT + n_epi); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&p.tmA[0]);
        if (p.nseg > 1) tma_prefetch_desc(&p.tmA[1]);
        tma_prefetch_desc(&p.tmB);
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
    pdl_launch_dependents();                      // the next conv launch may stage its prologue as soon as SMs free up
    if (warp != 0) pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer (warp-uniform, elected lane issues) =====================
        int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
        if (kRES) {                               // all weight tiles once (single N block): fullB[0] collects them
            if (elect_one()) {                    // (weights are constants: loaded before the predecessor's results are waited for)
                mbar_expect_tx(fullB, bytesB * nB);
                for (uint32_t i = 0; i < nB; ++i) tma_load_2d(sB0 + bytesB * i, &p.tmB, fullB, 0, (int)i * p.Ntot_pad);
            }
            __syncwarp();
        }
        pdl_wait();
        int slot = 0; uint32_t ps = 0;
        int w_next = sched_fetch(p.work_counter, lane);
        while (true) {
            const int w = w_next < p.n_work ? w_next : -1;
            mbar_wait_uniform(schedEmpty + 8 * slot, ps ^ 1);
            if (elect_one()) {
                asm volatile("st.shared.b32 [%0], %1;" ::"r"(sched_w + 4 * slot), "r"(w) : "memory");
                mbar_arrive(schedFull + 8 * slot);
            }
            __syncwarp();
            if (++slot == SCHED_DEPTH) { slot = 0; ps ^= 1; }
            if (w < 0) break;
            w_next = sched_fetch(p.work_counter, lane);                  // latency hidden behind this item's loads
            const int st = fast_div(w, p.nnb_magic), nb = w - st * p.nNB;
            const int n0 = nb * p.NT;
            TileCoord tcs[kMT];
#pragma unroll
            for (int i = 0; i < kMT; ++i) tcs[i] = decode_tile(p, st * kMT + i);
            int chunk = 0;
            for (int s = 0; s < p.nseg; ++s) {
                const CUtensorMap* tm = &p.tmA[s];
                for (int kx = 0; kx < p.seg_nkx[s]; ++kx) {
                    const int c1 = p.seg_c1off[s] + kx * p.seg_c1step[s];
                    for (int ck = 0; ck < p.seg_nck[s]; ++ck, ++chunk) {
                        mbar_wait_uniform(emptyA + 8 * sa, pa ^ 1);
                        if (elect_one()) {
                            const uint32_t dst = sA0 + bytesA * sa, bar = fullA + 8 * sa;
                            mbar_expect_tx(bar, bytesA);
#pragma unroll
                            for (int i = 0; i < kMT; ++i)
                                tma_load_4d(dst + bytesA1 * i, tm, bar, ck * 64, tcs[i].r0 + c1, tcs[i].y0 * p.ystep - p.padY, tcs[i].frame);
                        }
                        __syncwarp();
                        if (++sa == p.stagesA) { sa = 0; pa ^= 1; }
                        if (!kRES) {
                            for (int dy = 0; dy < p.KH; ++dy) {
                                mbar_wait_uniform(emptyB + 8 * sb, pb ^ 1);
                                if (elect_one()) {
                                    mbar_expect_tx(fullB + 8 * sb, bytesB);
                                    tma_load_2d(sB0 + bytesB * sb, &p.tmB, fullB + 8 * sb, 0, (chunk * p.KH + dy) * p.Ntot_pad + n0);
                                }
                                __syncwarp();
                                if (++sb == p.stagesB) { sb = 0; pb ^= 1; }
                            }
                        }
                    }
                }
            }
        }
    } else if (warp >= 1 && warp <= kMT) {
        // ===================== MMA issuer of M-tile `mt` (warp-uniform, elected lane issues) =====================
        // Everything the loop needs is hoisted into (uniform) registers: an issuing warp is latency bound, so each
        // extra instruction per tcgen05.mma shows up directly when N is small (one MMA = N/2 tensor cycles).
        const int mt = warp - 1;
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((128u >> 4) << 24);
        const int KH = p.KH, stagesA = p.stagesA, stagesB = p.stagesB, acc_stages = p.acc_stages, nseg = p.nseg;
        const uint32_t NTc = (uint32_t)p.NTc;
        const uint32_t dy_step = ((uint32_t)p.RT * 128u) >> 4, b_step = bytesB >> 4, a_step = bytesA >> 4;
        const uint32_t a_lo0 = desc_lo(sA0 + bytesA1 * (uint32_t)mt), b_lo0 = desc_lo(sB0);
        const uint64_t hiA = desc_hi_sw128(1024u * (uint32_t)p.ystep), hiB = desc_hi_sw128(1024u);
        int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
        int as = 0; uint32_t pacc = 0;
        if (kRES) { mbar_wait_uniform(fullB, 0); tc_fence_after(); }
        int slot = 0; uint32_t ps = 0;
        while (sched_next(schedFull, schedEmpty, sched_w, slot, ps, lane) >= 0) {
            mbar_wait_uniform(accEmpty + 8 * as, pacc ^ 1);              // epilogue drained this accumulator stage
            tc_fence_after();
            const uint32_t td = tmem_base + (uint32_t)(as * kMT + mt) * NTc;
            uint32_t acc = 0;
            uint32_t b_res = b_lo0;                                      // resident mode: walks the weight tiles in order
            for (int s = 0; s < nseg; ++s) {
                const int nck = p.seg_nck[s], klast = p.seg_klast[s] >> 4;
                for (int kc = p.seg_nkx[s] * nck, ck = 0; kc > 0; --kc) {
                    const int ksteps = (ck == nck - 1) ? klast : 4;
                    if (++ck == nck) ck = 0;
                    mbar_wait_uniform(fullA + 8 * sa, pa);
                    tc_fence_after();
                    uint32_t alo = a_lo0 + a_step * (uint32_t)sa;
                    for (int dy = 0; dy < KH; ++dy) {
                        uint32_t blo;
                        if (kRES) { blo = b_res; b_res += b_step; }
                        else { mbar_wait_uniform(fullB + 8 * sb, pb); tc_fence_after(); blo = b_lo0 + b_step * (uint32_t)sb; }
                        if (elect_one()) {
                            tc_mma_bf16(td, hiA | alo, hiB | blo, idesc, acc);
                            if (ksteps > 1) tc_mma_bf16(td, hiA | (alo + 2), hiB | (blo + 2), idesc, 1);
                            if (ksteps > 2) tc_mma_bf16(td, hiA | (alo + 4), hiB | (blo + 4), idesc, 1);
                            if (ksteps > 3) tc_mma_bf16(td, hiA | (alo + 6), hiB | (blo + 6), idesc, 1);
                            if (!kRES) tc_commit(emptyB + 8 * sb);
                        }
                        __syncwarp();
                        acc = 1;
                        alo += dy_step;
                        if (!kRES) { if (++sb == stagesB) { sb = 0; pb ^= 1; } }
                    }
                    if (elect_one()) tc_commit(emptyA + 8 * sa);
                    __syncwarp();
                    if (++sa == stagesA) { sa = 0; pa ^= 1; }
                }
            }
            if (elect_one()) tc_commit(accFull + 8 * as);
            __syncwarp();
            if (++as == acc_stages) { as = 0; pacc ^= 1; }
        }
    } else if (warp < kEpiFirst || ((warp - kEpiFirst) >> 2) >= p.epi_per_q) {
        // idle warps (incl. the epilogue warps this launch does not use: tensor-bound layers run faster with 8 than with 16)
    } else {
        // ===================== epilogue (warps 4..) =====================
        const int q = warp & 3;                          // TMEM lane quarter this warp may access
        const int h = (warp - kEpiFirst) >> 2;           // which share of the 16-column groups (0 .. epi_per_q-1)
        const int epiPerQ = min(p.epi_per_q, kEpiPerQ), epiThreads = epiPerQ * 128;
        const int m = q * 32 + lane;
        const int yy = m >> p.logRT, rr = m & (p.RT - 1);
        // bias of this CTA's N block -> smem (reloaded per work item only when there are several N blocks)
        float* sbias = (float*)(smem_raw + (sBias - smem_u32(smem_raw)));
        int bias_nb = -1;
        int as = 0; uint32_t pacc = 0;
        const int units_per_tile = p.NT >> 4;
        const int units = units_per_tile * kMT;
        const bool vec8 = (p.Cout & 7) == 0 && !p.out_f32;
        // 16-byte alignment of every address the fast path forms (8-element granularity of base and strides)
        const bool vec16 = (p.Cout & 15) == 0 && !p.out_f32 && ((p.out_coff | p.out_sx | (int)(p.out_sy & 7) | (int)(p.out_sn & 7)) & 7) == 0;
        const bool f32fast = p.out_f32 && p.out_sx == p.Cout && (p.Sy == 1 || ((p.Sx * p.Cout) & 15) == 0) && p.out_coff == 0 &&
                             ((p.Ntot | (int)p.out_sy | (int)p.out_sn) & 3) == 0;
        int slot = 0; uint32_t ps = 0;
        int tile_no = 0;                                 // kPOOLX: rotates the unit-pair assignment
        while (true) {
            const int w = sched_next(schedFull, schedEmpty, sched_w, slot, ps, lane);
            if (w < 0) break;
            const int st = fast_div(w, p.nnb_magic), nb = w - st * p.nNB;
            const int n0 = nb * p.NT;
            if (nb != bias_nb) {
                asm volatile("bar.sync 1, %0;" ::"r"(epiThreads));     // everyone finished reading the previous bias
                for (int i = threadIdx.x - kEpiFirst * 32; i < p.NT; i += epiThreads) sbias[i] = __ldg(p.bias + n0 + i);
                asm volatile("bar.sync 1, %0;" ::"r"(epiThreads));
                bias_nb = nb;
            }
            mbar_wait(accFull + 8 * as, pacc);
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kMT * p.NTc);
            // software pipeline over this warp's 16-column units: the TMEM load of the next unit is in flight while the
            // current one is processed; two register sets ping-pong (no copies)
            uint32_t va[16], vb[16];
            int cur_mt = -1;
            bool row_ok = false, sy1_ok = true, pool_ok = false;
            long long base = 0, pbase = 0;
            auto tile_of = [&](int g) -> int {               // M-tile of unit g (units are laid out tile after tile)
                if (kMT == 1) return 0;
                if (kMT == 2) return g >= units_per_tile ? 1 : 0;
                return g / units_per_tile;
            };
            auto unit_addr = [&](int g) -> uint32_t {
                const int mt = tile_of(g);
                return tacc + (uint32_t)(mt * p.NTc + (g - mt * units_per_tile) * 16);
            };
            auto enter = [&](int g) -> int {                 // (re)compute the row's output base when the M-tile changes; returns j0
                const int mt = tile_of(g);
                if (mt != cur_mt) {
                    cur_mt = mt;
                    const TileCoord tc = decode_tile(p, st * kMT + mt);
                    const int y = tc.y0 + yy, r = tc.r0 + rr;
                    row_ok = tc.valid && (y < p.Hin) && (r < p.nR);
                    if (kPOOLX) {                            // Sx even, Sy = 1: GEMM row (y, r) holds Sx / 2 pooled pixels of pooled row y / 2
                        pool_ok = tc.valid && !(y & 1) && (y >> 1) < p.pool_H && r < p.nR;
                        pbase = (long long)tc.frame * p.pool_sn + (long long)(y >> 1) * p.pool_sy +
                                (long long)((p.Sx >> 1) * r + p.pool_padx) * p.pool_sx;
                    } else if (kPOOL2) {                     // Sx = Sy = 2: GEMM row (y, r) IS the pooled pixel
                        pool_ok = tc.valid && y < p.pool_H && r < p.pool_W;
                        pbase = (long long)tc.frame * p.pool_sn + (long long)y * p.pool_sy + (long long)(r + p.pool_padx) * p.pool_sx;
                    } else if (p.pool_out != nullptr) {      // Sx = Sy = 1: (y, r) is the output pixel; even lanes of even rows write the pooled pixel
                        pool_ok = tc.valid && !((y | r) & 1) && (y >> 1) < p.pool_H && (r >> 1) < p.pool_W;
                        pbase = (long long)tc.frame * p.pool_sn + (long long)(y >> 1) * p.pool_sy + (long long)((r >> 1) + p.pool_padx) * p.pool_sx;
                    }
                    base = (long long)tc.frame * p.out_sn + (long long)(p.Sy * y) * p.out_sy + (long long)(p.Sx * r + p.out_padx) * p.out_sx + p.out_coff;
                    if (p.epi_mode != AM_EPI_PLAIN) pbase = (long long)tc.frame * p.out_H + (long long)(p.Sy * y);
                    if (p.Sy * y >= p.out_H || p.Sx * r + p.Sx > p.out_W) row_ok = false;
                    sy1_ok = p.Sy * y + 1 < p.out_H;                               // odd image height: the last row pair has no second row
                }
                return (g - mt * units_per_tile) * 16;
            };
            uint4* spool_q = (uint4*)(smem_raw + (sPool - smem_u32(smem_raw))) + (size_t)q * units_per_tile * 64;     // 2 x uint4 per lane
            auto spool_of = [&](int g) -> uint4* { return kPOOL2 ? spool_q + ((size_t)g * 32 + lane) * 2 : nullptr; };
            if (kPOOLX) {
                // kMT = 1, one N block of Sx x Cout columns: unit (sx, slice) = sx * cpu + slice.  A warp takes PAIRS of units, the same
                // slice of pixels 2 sxp and 2 sxp + 1, back to back (same ping-pong of the two register sets); the pair -> warp assignment
                // rotates from tile to tile so that an uneven pair count (6 pairs on 4 warps) evens out over the two accumulator stages.
                const int cpu = p.Cout >> 4, npairs = units_per_tile >> 1;
                int hr = h + ((tile_no & 1) ? (epiPerQ >> 1) : 0);
                if (hr >= epiPerQ) hr -= epiPerQ;
                ++tile_no;
                auto first_of = [&](int pi) -> int {         // first unit of pair pi = (sxp, slice): pixel 2 sxp
                    const int sxp = (int)__umulhi((unsigned)(pi << 4), p.cout_magic);
                    return (2 * sxp) * cpu + (pi - sxp * cpu);
                };
                uint32_t carry[8];
                int pi = hr;
                int gA = pi < npairs ? first_of(pi) : 0;
                if (pi < npairs) tmem_ld16_async(unit_addr(gA), va);
                while (pi < npairs) {
                    tmem_wait16(va);
                    tmem_ld16_async(unit_addr(gA + cpu), vb);
                    { const int j0 = enter(gA); epi_unit<kFUSED, false, 1>(p, va, sbias, n0, j0, base, vec16, vec8, f32fast, sy1_ok, row_ok, pbase, pool_ok, nullptr, carry); }
                    const int pn = pi + epiPerQ;
                    const int gN = pn < npairs ? first_of(pn) : 0;
                    tmem_wait16(vb);
                    if (pn < npairs) tmem_ld16_async(unit_addr(gN), va);
                    { const int j0 = enter(gA + cpu); epi_unit<kFUSED, false, 2>(p, vb, sbias, n0, j0, base, vec16, vec8, f32fast, sy1_ok, row_ok, pbase, pool_ok, nullptr, carry); }
                    pi = pn; gA = gN;
                }
            } else {
            int g = h;
            if (g < units) tmem_ld16_async(unit_addr(g), va);
            while (g < units) {
                tmem_wait16(va);
                int g2 = g + epiPerQ;
                if (g2 < units) tmem_ld16_async(unit_addr(g2), vb);
                { const int j0 = enter(g); if (row_ok || p.pool_out != nullptr) epi_unit<kFUSED, kPOOL2>(p, va, sbias, n0, j0, base, vec16, vec8, f32fast, sy1_ok, row_ok, pbase, pool_ok, spool_of(g)); }
                g = g2;
                if (g >= units) break;
                tmem_wait16(vb);
                g2 = g + epiPerQ;
                if (g2 < units) tmem_ld16_async(unit_addr(g2), va);
                { const int j0 = enter(g); if (row_ok || p.pool_out != nullptr) epi_unit<kFUSED, kPOOL2>(p, vb, sbias, n0, j0, base, vec16, vec8, f32fast, sy1_ok, row_ok, pbase, pool_ok, spool_of(g)); }
                g = g2;
            }
            }
            // all tcgen05.ld of this stage have completed (tmem_wait16 in the last iteration): hand the stage back
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(accEmpty + 8 * as);
            if (++as == p.acc_stages) { as = 0; pacc ^= 1; }
            if (kPOOL2) {
                // the quarter's warps have parked all units of this tile: max over the four (sy, sx) groups per channel slice
                asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "r"(epiPerQ * 32) : "memory");
                const int slices = units_per_tile >> 2;                              // Cout / 16
                for (int sl = h; sl < slices; sl += epiPerQ) {
                    uint4 m0, m1;
                    {
                        const uint4* a = spool_q + ((size_t)sl * 32 + lane) * 2;
                        m0 = a[0]; m1 = a[1];
                    }
#pragma unroll
                    for (int grp = 1; grp < 4; ++grp) {
                        const uint4* a = spool_q + ((size_t)(grp * slices + sl) * 32 + lane) * 2;
                        const uint4 b0 = a[0], b1 = a[1];
                        __nv_bfloat162 t;
                        t = __hmax2(*(__nv_bfloat162*)&m0.x, *(const __nv_bfloat162*)&b0.x); m0.x = *(uint32_t*)&t;
                        t = __hmax2(*(__nv_bfloat162*)&m0.y, *(const __nv_bfloat162*)&b0.y); m0.y = *(uint32_t*)&t;
                        t = __hmax2(*(__nv_bfloat162*)&m0.z, *(const __nv_bfloat162*)&b0.z); m0.z = *(uint32_t*)&t;
                        t = __hmax2(*(__nv_bfloat162*)&m0.w, *(const __nv_bfloat162*)&b0.w); m0.w = *(uint32_t*)&t;
                        t = __hmax2(*(__nv_bfloat162*)&m1.x, *(const __nv_bfloat162*)&b1.x); m1.x = *(uint32_t*)&t;
                        t = __hmax2(*(__nv_bfloat162*)&m1.y, *(const __nv_bfloat162*)&b1.y); m1.y = *(uint32_t*)&t;
                        t = __hmax2(*(__nv_bfloat162*)&m1.z, *(const __nv_bfloat162*)&b1.z); m1.z = *(uint32_t*)&t;
                        t = __hmax2(*(__nv_bfloat162*)&m1.w, *(const __nv_bfloat162*)&b1.w); m1.w = *(uint32_t*)&t;
                    }
                    if (pool_ok) {
                        __nv_bfloat16* po = (__nv_bfloat16*)p.pool_out + pbase + sl * 16;
                        *(uint4*)po = m0;
                        *(uint4*)(po + 8) = m1;
                    }
                }
                asm volatile("bar.sync %0, %1;" ::"r"(2 + q), "r"(epiPerQ * 32) : "memory");      // parked values consumed: next tile may overwrite
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (threadIdx.x == 0) {                              // last CTA out re-arms the scheduler for the next launch of this plan
        __threadfence();
        if (atomicAdd(p.work_counter + 1, 1) == (int)gridDim.x - 1) { p.work_counter[0] = 0; p.work_counter[1] = 0; __threadfence(); }
    }
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
    }
}

// ================================================================================================
// CTA-pair variant (cta_group::2).  Two CTAs of a cluster (= the two SMs of a TPC) run ONE tcgen05.mma of M = 256:
// every CTA keeps its own A tiles (its own 128 GEMM rows) and accumulators, but only HALF of each weight tile
// (NT/2 rows of B) -- the tensor cores of both SMs read both halves.  That halves the shared-memory operand
// traffic of B per SM (the bound of the M = N = 128 layers: 4 KB of A + 4 KB of B per K = 16 step at 128 B/clk
// is exactly the tensor time; with a pair it is 4 + 2 KB) and the L2 -> shared-memory weight traffic.
// Structure = k_conv_gemm<2, false> per CTA (two M-tiles, two issuer warps, streamed weights), with
//   * a work item = 4 M-tiles (2 per CTA) x one N block, statically strided over the clusters;
//   * TMA loads of BOTH CTAs completing on the leader's (rank 0) full barriers (.cta_group::2 + mapa);
//   * the leader's issuer warps issuing tcgen05.mma.cta_group::2 and multicasting their commits to the
//     empty / accumulator-full barriers of both CTAs;
//   * the peer's epilogue warps arriving remotely on the leader's accumulator-empty barriers.
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_rank(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(dst), "l"((unsigned long long)tm), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* tm, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"((unsigned long long)tm), "r"(bar_cluster), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tc2_commit(uint32_t bar) {            // arrives on `bar` (same offset) in both CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void tc2_mma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

template <bool kFUSED>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(CONV_THREADS, 1) k_conv_gemm_pair(const __grid_constant__ ConvParams p) {
    constexpr int kMT = 2;                               // M-tiles per CTA (4 per work item)
    constexpr int kEpiFirst = 4, kEpiWarps = (CONV_THREADS / 32) - kEpiFirst, kEpiPerQ = kEpiWarps / 4;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t bytesA1 = (uint32_t)(p.ystep * (p.YT - 1) + p.KH) * p.RT * 128u;
    const uint32_t bytesA = bytesA1 * kMT;
    const uint32_t bytesBh = (uint32_t)(p.NT >> 1) * 128u;               // this CTA's half of a weight tile
    const uint32_t sA0 = smem_base;
    const uint32_t sB0 = sA0 + bytesA * p.stagesA;
    const uint32_t sBias = sB0 + ((bytesBh * (uint32_t)p.stagesB + 1023u) & ~1023u);
    const uint32_t bar0 = sBias + (((uint32_t)p.NT * 4u + 127u) & ~127u);
    const uint32_t fullA = bar0, emptyA = fullA + 8 * p.stagesA, fullB = emptyA + 8 * p.stagesA, emptyB = fullB + 8 * p.stagesB;
    const uint32_t accFull = emptyB + 8 * p.stagesB, accEmpty = accFull + 16;
    const uint32_t tmem_slot = accEmpty + 16;
    const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int n_work = ((p.n_mtiles + 2 * kMT - 1) / (2 * kMT)) * p.nNB;
    const int w_first = (int)cluster_id_x(), w_step = (int)cluster_count_x();

    if (threadIdx.x == 0) {
        for (int i = 0; i < p.stagesA; ++i) { mbar_init(fullA + 8 * i, 1); mbar_init(emptyA + 8 * i, kMT); }
        for (int i = 0; i < p.stagesB; ++i) { mbar_init(fullB + 8 * i, 1); mbar_init(emptyB + 8 * i, kMT); }
        const int n_epi = min(4 * p.epi_per_q, kEpiWarps);          // working epilogue warps per CTA of the pair
        for (int i = 0; i < 2; ++i) { mbar_init(accFull + 8 * i, kMT); mbar_init(accEmpty + 8 * i, 2 * n_epi); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        tma_prefetch_desc(&p.tmA[0]);
        if (p.nseg > 1) tma_prefetch_desc(&p.tmA[1]);
        tma_prefetch_desc(&p.tmB);
        if (p.edge_half) tma_prefetch_desc(&p.tmBq);
    }
    if (warp == 1) {                                     // the same warp of both CTAs allocates the pair's tensor memory
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)p.tmem_cols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                                  // barriers of both CTAs initialised before any remote signal
    tc_fence_after();
    uint32_t tmem_base;
    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ===================== TMA producer: own A tiles, own half of every weight tile; completes on the leader =====================
        int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
        const uint32_t fullA_l = mapa_rank(fullA, 0), fullB_l = mapa_rank(fullB, 0);
        for (int w = w_first; w < n_work; w += w_step) {
            const int st = fast_div(w, p.nnb_magic), nb = w - st * p.nNB;
            const int n0 = nb * p.NT + (int)rank * (p.NT >> 1);
            TileCoord tcs[kMT];
#pragma unroll
            for (int i = 0; i < kMT; ++i) tcs[i] = decode_tile(p, st * 2 * kMT + (int)rank * kMT + i);
            int chunk = 0;
            for (int s = 0; s < p.nseg; ++s) {
                const CUtensorMap* tm = &p.tmA[s];
                for (int kx = 0; kx < p.seg_nkx[s]; ++kx) {
                    const int c1 = p.seg_c1off[s] + kx * p.seg_c1step[s];
                    for (int ck = 0; ck < p.seg_nck[s]; ++ck, ++chunk) {
                        mbar_wait_uniform(emptyA + 8 * sa, pa ^ 1);
                        if (elect_one()) {
                            const uint32_t dst = sA0 + bytesA * sa;
                            if (rank == 0) mbar_expect_tx(fullA + 8 * sa, 2 * bytesA);       // both CTAs' boxes
#pragma unroll
                            for (int i = 0; i < kMT; ++i)
                                tma2_load_4d(dst + bytesA1 * i, tm, fullA_l + 8 * sa, ck * 64, tcs[i].r0 + c1, tcs[i].y0 * p.ystep - p.padY, tcs[i].frame);
                        }
                        __syncwarp();
                        if (++sa == p.stagesA) { sa = 0; pa ^= 1; }
                        for (int t = 0; t < p.KH; ++t) {
                            // edge_half: taps in the order 1 .. KH-2, 0, KH-1 (the item's first MMA must cover all columns); the two
                            // edge taps only carry the sy = 0 resp. sy = 1 half of the weight rows, a quarter per CTA
                            const int dy = !p.edge_half ? t : (t < p.KH - 2 ? t + 1 : (t == p.KH - 2 ? 0 : p.KH - 1));
                            const bool edge = p.edge_half && t >= p.KH - 2;
                            mbar_wait_uniform(emptyB + 8 * sb, pb ^ 1);
                            if (elect_one()) {
                                if (!edge) {
                                    if (rank == 0) mbar_expect_tx(fullB + 8 * sb, 2 * bytesBh);
                                    tma2_load_2d(sB0 + bytesBh * sb, &p.tmB, fullB_l + 8 * sb, 0, (chunk * p.KH + dy) * p.Ntot_pad + n0);
                                } else {
                                    if (rank == 0) mbar_expect_tx(fullB + 8 * sb, bytesBh);
                                    tma2_load_2d(sB0 + bytesBh * sb, &p.tmBq, fullB_l + 8 * sb, 0,
                                                 (chunk * p.KH + dy) * p.Ntot_pad + (dy == 0 ? 0 : (p.NT >> 1)) + (int)rank * (p.NT >> 2));
                                }
                            }
                            __syncwarp();
                            if (++sb == p.stagesB) { sb = 0; pb ^= 1; }
                        }
                    }
                }
            }
        }
    } else if (warp >= 1 && warp <= kMT) {
        if (rank == 0) {
            // ===================== MMA issuer of M-tile pair `mt` (leader CTA only) =====================
            const int mt = warp - 1;
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 3) << 17) | ((256u >> 4) << 24);
            const uint32_t idesc_half = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(p.NT >> 4) << 17) | ((256u >> 4) << 24);
            const int KH = p.KH, stagesA = p.stagesA, stagesB = p.stagesB, acc_stages = p.acc_stages, nseg = p.nseg;
            const int edge_half = p.edge_half;
            const uint32_t NTc = (uint32_t)p.NTc, NT = (uint32_t)p.NT;
            const uint32_t dy_step = ((uint32_t)p.RT * 128u) >> 4, b_step = bytesBh >> 4, a_step = bytesA >> 4;
            const uint32_t a_lo0 = desc_lo(sA0 + bytesA1 * (uint32_t)mt), b_lo0 = desc_lo(sB0);
            const uint64_t hiA = desc_hi_sw128(1024u * (uint32_t)p.ystep), hiB = desc_hi_sw128(1024u);
            int sa = 0, sb = 0; uint32_t pa = 0, pb = 0;
            int as = 0; uint32_t pacc = 0;
            for (int w = w_first; w < n_work; w += w_step) {
                mbar_wait_uniform(accEmpty + 8 * as, pacc ^ 1);          // both CTAs' epilogues drained this accumulator stage
                tc_fence_after();
                const uint32_t td = tmem_base + (uint32_t)(as * kMT + mt) * NTc;
                uint32_t acc = 0;
                for (int s = 0; s < nseg; ++s) {
                    const int nck = p.seg_nck[s], klast = p.seg_klast[s] >> 4;
                    for (int kc = p.seg_nkx[s] * nck, ck = 0; kc > 0; --kc) {
                        const int ksteps = (ck == nck - 1) ? klast : 4;
                        if (++ck == nck) ck = 0;
                        mbar_wait_uniform(fullA + 8 * sa, pa);
                        tc_fence_after();
                        const uint32_t alo0 = a_lo0 + a_step * (uint32_t)sa;
                        for (int t = 0; t < KH; ++t) {
                            const int dy = !edge_half ? t : (t < KH - 2 ? t + 1 : (t == KH - 2 ? 0 : KH - 1));      // see the producer
                            const bool edge = edge_half && t >= KH - 2;
                            const uint32_t alo = alo0 + dy_step * (uint32_t)dy;
                            const uint32_t idx = edge ? idesc_half : idesc;
                            const uint32_t tdd = (edge && dy != 0) ? td + (NT >> 1) : td;        // sy = 1 columns
                            mbar_wait_uniform(fullB + 8 * sb, pb);
                            tc_fence_after();
                            const uint32_t blo = b_lo0 + b_step * (uint32_t)sb;
                            if (elect_one()) {
                                tc2_mma_bf16(tdd, hiA | alo, hiB | blo, idx, acc);
                                if (ksteps > 1) tc2_mma_bf16(tdd, hiA | (alo + 2), hiB | (blo + 2), idx, 1);
                                if (ksteps > 2) tc2_mma_bf16(tdd, hiA | (alo + 4), hiB | (blo + 4), idx, 1);
                                if (ksteps > 3) tc2_mma_bf16(tdd, hiA | (alo + 6), hiB | (blo + 6), idx, 1);
                                tc2_commit(emptyB + 8 * sb);
                            }
                            __syncwarp();
                            acc = 1;
                            if (++sb == stagesB) { sb = 0; pb ^= 1; }
                        }
                        if (elect_one()) tc2_commit(emptyA + 8 * sa);
                        __syncwarp();
                        if (++sa == stagesA) { sa = 0; pa ^= 1; }
                    }
                }
                if (elect_one()) tc2_commit(accFull + 8 * as);
                __syncwarp();
                if (++as == acc_stages) { as = 0; pacc ^= 1; }
            }
        }
    } else if (warp < kEpiFirst || ((warp - kEpiFirst) >> 2) >= p.epi_per_q) {
        // idle
    } else {
        // ===================== epilogue: this CTA's two M-tiles =====================
        const int q = warp & 3;
        const int h = (warp - kEpiFirst) >> 2;
        const int epiPerQ = min(p.epi_per_q, kEpiPerQ), epiThreads = epiPerQ * 128;
        const int m = q * 32 + lane;
        const int yy = m >> p.logRT, rr = m & (p.RT - 1);
        float* sbias = (float*)(smem_raw + (sBias - smem_u32(smem_raw)));
        const uint32_t accEmpty_l = mapa_rank(accEmpty, 0);
        int bias_nb = -1;
        int as = 0; uint32_t pacc = 0;
        const int units_per_tile = p.NT >> 4;
        const int units = units_per_tile * kMT;
        const bool vec8 = (p.Cout & 7) == 0 && !p.out_f32;
        const bool vec16 = (p.Cout & 15) == 0 && !p.out_f32 && ((p.out_coff | p.out_sx | (int)(p.out_sy & 7) | (int)(p.out_sn & 7)) & 7) == 0;
        const bool f32fast = p.out_f32 && p.out_sx == p.Cout && (p.Sy == 1 || ((p.Sx * p.Cout) & 15) == 0) && p.out_coff == 0 &&
                             ((p.Ntot | (int)p.out_sy | (int)p.out_sn) & 3) == 0;
        for (int w = w_first; w < n_work; w += w_step) {
            const int st = fast_div(w, p.nnb_magic), nb = w - st * p.nNB;
            const int n0 = nb * p.NT;
            if (nb != bias_nb) {
                asm volatile("bar.sync 1, %0;" ::"r"(epiThreads));
                for (int i = threadIdx.x - kEpiFirst * 32; i < p.NT; i += epiThreads) sbias[i] = __ldg(p.bias + n0 + i);
                asm volatile("bar.sync 1, %0;" ::"r"(epiThreads));
                bias_nb = nb;
            }
            mbar_wait(accFull + 8 * as, pacc);
            tc_fence_after();
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(as * kMT * p.NTc);
            uint32_t va[16], vb[16];
            int cur_mt = -1;
            bool row_ok = false, sy1_ok = true, pool_ok = false;
            long long base = 0, pbase = 0;
            auto unit_addr = [&](int g) -> uint32_t {
                const int mt = g >= units_per_tile ? 1 : 0;
                return tacc + (uint32_t)(mt * p.NTc + (g - mt * units_per_tile) * 16);
            };
            auto enter = [&](int g) -> int {
                const int mt = g >= units_per_tile ? 1 : 0;
                if (mt != cur_mt) {
                    cur_mt = mt;
                    const TileCoord tc = decode_tile(p, st * 2 * kMT + (int)rank * kMT + mt);
                    const int y = tc.y0 + yy, r = tc.r0 + rr;
                    row_ok = tc.valid && (y < p.Hin) && (r < p.nR);
                    if (p.pool_out != nullptr) {             // Sx = Sy = 1: (y, r) is the output pixel; even lanes of even rows write the pooled pixel
                        pool_ok = tc.valid && !((y | r) & 1) && (y >> 1) < p.pool_H && (r >> 1) < p.pool_W;
                        pbase = (long long)tc.frame * p.pool_sn + (long long)(y >> 1) * p.pool_sy + (long long)((r >> 1) + p.pool_padx) * p.pool_sx;
                    }
                    base = (long long)tc.frame * p.out_sn + (long long)(p.Sy * y) * p.out_sy + (long long)(p.Sx * r + p.out_padx) * p.out_sx + p.out_coff;
                    if (p.epi_mode != AM_EPI_PLAIN) pbase = (long long)tc.frame * p.out_H + (long long)(p.Sy * y);
                    if (p.Sy * y >= p.out_H || p.Sx * r + p.Sx > p.out_W) row_ok = false;
                    sy1_ok = p.Sy * y + 1 < p.out_H;
                }
                return (g - mt * units_per_tile) * 16;
            };
            int g = h;
            if (g < units) tmem_ld16_async(unit_addr(g), va);
            while (g < units) {
                tmem_wait16(va);
                int g2 = g + epiPerQ;
                if (g2 < units) tmem_ld16_async(unit_addr(g2), vb);
                { const int j0 = enter(g); if (row_ok || p.pool_out != nullptr) epi_unit<kFUSED>(p, va, sbias, n0, j0, base, vec16, vec8, f32fast, sy1_ok, row_ok, pbase, pool_ok); }
                g = g2;
                if (g >= units) break;
                tmem_wait16(vb);
                g2 = g + epiPerQ;
                if (g2 < units) tmem_ld16_async(unit_addr(g2), va);
                { const int j0 = enter(g); if (row_ok || p.pool_out != nullptr) epi_unit<kFUSED>(p, vb, sbias, n0, j0, base, vec16, vec8, f32fast, sy1_ok, row_ok, pbase, pool_ok); }
                g = g2;
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) { if (rank == 0) mbar_arrive(accEmpty + 8 * as); else mbar_arrive_cluster(accEmpty_l + 8 * as); }
            if (++as == p.acc_stages) { as = 0; pacc ^= 1; }
        }
        tc_fence_before();
    }
    __syncthreads();
    cluster_sync_all();                                  // nobody leaves while the pair still reads its shared / tensor memory
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols) : "memory");
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

struct am_conv_plan {
    ConvParams p;
    size_t smem;
    int grid;
    int pair;            // 1: k_conv_gemm_pair (cta_group::2 clusters of two CTAs)
    int* d_counter;      // 2 ints, zero between launches
};

static int sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}

// Fills a launch plan from the plain-C descriptor: tensor maps, tiling, smem/TMEM budget.
static int conv_prepare(const am_conv_desc* d, am_conv_plan* plan) {
    if (!d || d->nseg < 1 || d->nseg > 2 || d->RT % 8 != 0 || (d->RT & (d->RT - 1)) != 0 || d->RT * d->YT != 128 || d->NT % 16 != 0 ||
        d->NT < 16 || d->NT > 256 || d->batch <= 0 || d->Ntot_pad % d->NT != 0)
        return AM_ERR_ARG;
    ConvParams& p = plan->p;
    memset(&p, 0, sizeof(p));
    p.nseg = d->nseg;
    const int ystep = d->in_ystep == 2 ? 2 : 1;
    if (ystep == 2 && d->RT != 8) return AM_ERR_ARG;      // M-atoms must be whole image rows for the descriptor stride trick
    const int box_rows = ystep * (d->YT - 1) + d->KH;
    int total_chunks = 0;
    for (int s = 0; s < d->nseg; ++s) {
        const am_conv_seg* g = &d->seg[s];
        unsigned long long dims[4], strides[3]; unsigned box[4];
        if (g->rowrun) {        // dim0 = run elements, dim1 = output group r (row stride S*C elements: rows overlap)
            dims[0] = (unsigned long long)g->run_len; dims[1] = (unsigned long long)d->nR; dims[2] = (unsigned long long)g->Hbuf; dims[3] = (unsigned long long)d->batch;
            strides[0] = (unsigned long long)g->S * g->C * 2ull;
            p.seg_nkx[s] = 1; p.seg_c1step[s] = 0; p.seg_c1off[s] = 0;
            p.seg_nck[s] = (g->run_len + 63) / 64;
            p.seg_klast[s] = ((g->run_len - (p.seg_nck[s] - 1) * 64) + 15) & ~15;
        } else {                // dim0 = channels, dim1 = padded x; one load per horizontal tap
            dims[0] = (unsigned long long)g->C; dims[1] = (unsigned long long)g->Wp; dims[2] = (unsigned long long)g->Hbuf; dims[3] = (unsigned long long)d->batch;
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
    const bool pair = (d->flags & AM_CONV_CTA_PAIR) != 0;
    if (pair && (d->NT % 16 != 0 || sm_count() < 2)) return AM_ERR_ARG;
    plan->pair = pair ? 1 : 0;
    {
        unsigned long long dims[2] = {64ull, (unsigned long long)total_chunks * d->KH * d->Ntot_pad};
        unsigned long long strides[1] = {128ull};
        unsigned box[2] = {64u, (unsigned)(pair ? d->NT / 2 : d->NT)};       // a CTA pair splits every weight tile in two
        int rc = encode_map(&p.tmB, (void*)d->weights, 2, dims, strides, box);
        if (rc) return rc;
    }
    // 2-D packing (Sy = 2): the Toeplitz extension in y leaves the first vertical tap with weights for the sy = 0 columns only and the
    // last one for the sy = 1 columns only.  A CTA pair issues those two taps as N/2 MMAs (half the tensor and weight-read time of
    // 2 of the KH taps) when the whole layer is one N block whose halves are exactly the two sy groups.  N >= 128 only: below that
    // an MMA is bound by the shared-memory read of A, and the narrower edge MMAs only add issue slots (conv_out + 11 %, r02_d).
    p.edge_half = (pair && ystep == 2 && d->Sy == 2 && d->KH >= 4 && d->Ntot_pad == d->NT && d->Ntot == d->NT && d->NT % 32 == 0 && d->NT >= 128 &&
                   !(d->flags & AM_CONV_NO_EDGE_HALF)) ? 1 : 0;
    if (p.edge_half) {
        unsigned long long dims[2] = {64ull, (unsigned long long)total_chunks * d->KH * d->Ntot_pad};
        unsigned long long strides[1] = {128ull};
        unsigned box[2] = {64u, (unsigned)(d->NT / 4)};
        int rc = encode_map(&p.tmBq, (void*)d->weights, 2, dims, strides, box);
        if (rc) return rc;
    }
    p.KH = d->KH; p.RT = d->RT; p.YT = d->YT; p.padY = d->padY; p.ystep = ystep;
    p.logRT = 0; while ((1 << p.logRT) < d->RT) ++p.logRT;
    p.nRT = (d->nR + d->RT - 1) / d->RT; p.nYT = (d->Hin + d->YT - 1) / d->YT; p.batch = d->batch;
    p.nR = d->nR; p.Hin = d->Hin; p.NT = d->NT; p.Ntot_pad = d->Ntot_pad; p.nNB = d->Ntot_pad / d->NT;
    p.NTc = 32; while (p.NTc < d->NT) p.NTc <<= 1;
    p.total_chunks = total_chunks;
    p.n_mtiles = p.nRT * p.nYT * d->batch;
    p.out = d->out; p.out_f32 = d->out_f32; p.out_H = d->out_H; p.out_W = d->out_W;
    p.out_sn = d->out_sn; p.out_sy = d->out_sy; p.out_sx = d->out_sx; p.out_padx = d->out_padx; p.out_coff = d->out_coff;
    p.Cout = d->Cout; p.Sy = d->Sy; p.Sx = d->Sx; p.Ntot = d->Ntot; p.act = d->act; p.bias = d->bias;
    p.pool_out = d->pool_out; p.pool_H = d->pool_H; p.pool_W = d->pool_W; p.pool_sn = d->pool_sn; p.pool_sy = d->pool_sy;
    p.pool_sx = d->pool_sx; p.pool_padx = d->pool_padx;
    p.pool2 = 0; p.poolx = 0;
    if (d->pool_out && d->Sx == 2 && d->Sy == 2) {        // the 2x2 block of a pooled pixel is one GEMM row (kPOOL2)
        const bool vec16 = (d->Cout & 15) == 0 && !d->out_f32 && ((d->out_coff | d->out_sx | (int)(d->out_sy & 7) | (int)(d->out_sn & 7)) & 7) == 0;
        if (ystep != 2 || !vec16 || d->Ntot != 4 * d->Cout || d->Ntot_pad != d->NT || d->NT != d->Ntot || (d->pool_sx & 7) || (d->pool_sy & 7) ||
            (d->pool_sn & 7) || d->pool_H != d->out_H / 2 || d->pool_W != d->out_W / 2 || pair || (d->flags & (AM_CONV_FORCE_MT2 | AM_CONV_FORCE_MT4))) {
            fprintf(stderr, "[accessmath_b200] am_conv: fused max-pool with Sx = Sy = 2 needs one N block of 4 x Cout columns, Cout %% 16 == 0, one M-tile per item\n");
            return AM_ERR_ARG;
        }
        p.pool2 = 1;
    } else if (d->pool_out && d->Sx >= 2 && d->Sy == 1) { // x neighbours in two column units, y neighbours in two lanes (kPOOLX)
        const bool vec16 = (d->Cout & 15) == 0 && !d->out_f32 && ((d->out_coff | d->out_sx | (int)(d->out_sy & 7) | (int)(d->out_sn & 7)) & 7) == 0;
        if ((d->Sx & 1) || d->RT > 16 || (d->YT & 1) || !vec16 || d->Ntot != d->Sx * d->Cout || d->Ntot_pad != d->NT || d->NT != d->Ntot ||
            (d->pool_sx & 7) || (d->pool_sy & 7) || (d->pool_sn & 7) || d->pool_H != d->out_H / 2 || d->pool_W != d->out_W / 2 ||
            d->pool_W * 2 != d->Sx * d->nR || pair || (d->flags & (AM_CONV_FORCE_MT2 | AM_CONV_FORCE_MT4))) {
            fprintf(stderr, "[accessmath_b200] am_conv: fused max-pool with Sx > 1 needs even Sx, Sy = 1, RT <= 16, one N block of Sx x Cout columns, "
                            "Cout %% 16 == 0, one M-tile per item\n");
            return AM_ERR_ARG;
        }
        p.poolx = 1;
    } else if (d->pool_out) {            // the fused pool rides on the single-address 16-channel epilogue path
        const bool vec16 = (d->Cout & 15) == 0 && !d->out_f32 && ((d->out_coff | d->out_sx | (int)(d->out_sy & 7) | (int)(d->out_sn & 7)) & 7) == 0;
        if (d->Sx != 1 || d->Sy != 1 || d->RT > 16 || !vec16 || d->Ntot != d->Cout || (d->pool_sx & 15) || (d->pool_sy & 15) || (d->pool_sn & 15) ||
            d->pool_H != d->out_H / 2 || d->pool_W != d->out_W / 2) {
            fprintf(stderr, "[accessmath_b200] am_conv: fused max-pool needs Sx = Sy = 1, RT <= 16, bf16 output with Cout %% 16 == 0\n");
            return AM_ERR_ARG;
        }
    }
    p.epi_mode = d->epi_mode; p.frames = d->frames; p.img_H = d->out_H; p.img_W = d->out_W;
    p.diff_out = (__nv_bfloat16*)d->diff_out; p.diff_C = d->diff_C; p.diff_pad = d->diff_pad;
    p.text_out = d->text_out; p.rec_out = d->rec_out;
    p.bits_out = (uint16_t*)d->bits_out; p.bits_hpr = 2 * d->bits_wpr; p.threshold = d->threshold;
    if (d->epi_mode == AM_EPI_HEADS) {
        if (d->Cout != 4 || !d->out_f32 || !d->diff_out || !d->frames || (d->diff_C != 4 && d->diff_C != 8) || d->out_sx != 1 ||
            d->out_sy != d->out_W || d->out_sn != (long long)d->out_H * d->out_W || d->out_padx != 0 || d->out_coff != 0 || d->pool_out) {
            fprintf(stderr, "[accessmath_b200] am_conv: AM_EPI_HEADS needs Cout = 4, fp32 geometry, a diff buffer with 4 or 8 channels\n");
            return AM_ERR_ARG;
        }
    } else if (d->epi_mode == AM_EPI_THRESHOLD) {
        if (d->Cout != 1 || !d->out_f32 || !d->bits_out || d->Sx % 16 != 0 || d->out_W % 16 != 0 || d->out_sx != 1 ||
            d->out_sy != d->out_W || d->out_sn != (long long)d->out_H * d->out_W || d->out_padx != 0 || d->out_coff != 0 || d->pool_out ||
            d->bits_wpr * 32 < d->out_W) {
            fprintf(stderr, "[accessmath_b200] am_conv: AM_EPI_THRESHOLD needs Cout = 1, Sx %% 16 == 0, width %% 16 == 0\n");
            return AM_ERR_ARG;
        }
    } else if (d->epi_mode != AM_EPI_PLAIN || !d->out) return AM_ERR_ARG;
    p.cout_magic = d->Cout >= 2 ? (unsigned)((1ull << 32) / (unsigned long long)d->Cout) + 1u : 0u;
    auto magic = [](int dv) -> unsigned { return dv >= 2 ? (unsigned)((1ull << 32) / (unsigned long long)dv) + 1u : 0u; };
    p.nrt_magic = magic(p.nRT); p.nyt_magic = magic(p.nYT); p.nnb_magic = magic(p.nNB);
    // epilogue warps per TMEM lane quarter: layers whose main loop is short (K <= 640 incl. padding) are bound by the epilogue's
    // issue rate and want all 16 warps; the tensor-bound layers run 2-5 % faster with 8 (fewer warps competing with the MMA
    // issuers / TMA producer for issue slots and power) -- measured per layer, profiles/README.md (r01_v)
    p.epi_per_q = (d->flags & AM_CONV_EPI16) ? 4 : (d->flags & AM_CONV_EPI8) ? 2 : ((long long)total_chunks * d->KH * 64 <= 640 ? 4 : 2);
    if ((unsigned long long)(p.n_mtiles + 8) * (unsigned long long)std::max(p.nRT, std::max(p.nYT, p.nNB)) * (unsigned long long)std::max(1, p.nNB) >= (1ull << 32)) {
        fprintf(stderr, "[accessmath_b200] am_conv: tile count too large for the reciprocal division\n");
        return AM_ERR_ARG;
    }
    if (d->Sy < 1 || d->Sy > 2 || d->Sx < 1) return AM_ERR_ARG;

    // ---- shared-memory / TMEM budget -------------------------------------------------------------------
    const size_t bytesA1 = (size_t)box_rows * d->RT * 128, bytesB = (size_t)d->NT * 128;
    const size_t fixed = 1024 /*align*/ + 1024 /*bias (<= 256 floats)*/ + 512 /*barriers*/;
    const size_t budget_total = 226 * 1024;
    if (pair) {
        const size_t budget = budget_total;          // two M-tiles per CTA, streamed half weight tiles
        if (2 * p.NTc > 512) return AM_ERR_ARG;
        const size_t bytesBh = bytesB / 2;
        int sa = 2, sb = 2;
        if (fixed + bytesA1 * 2 * sa + bytesBh * sb > budget) return AM_ERR_ARG;
        while (true) {
            bool grew = false;
            if (sb < 3 * d->KH && sb < 12 && fixed + bytesA1 * 2 * sa + bytesBh * (sb + 1) <= budget) { ++sb; grew = true; }
            if (sa < 4 && fixed + bytesA1 * 2 * (sa + 1) + bytesBh * sb <= budget) { ++sa; grew = true; }
            if (!grew) break;
        }
        p.residentB = 0; p.MT = 2; p.stagesA = sa; p.stagesB = sb;
        p.acc_stages = (2 * 2 * p.NTc <= 512) ? 2 : 1;
        p.tmem_cols = 32; while (p.tmem_cols < p.acc_stages * 2 * p.NTc) p.tmem_cols <<= 1;
        p.n_work = ((p.n_mtiles + 3) / 4) * p.nNB;
        plan->smem = 1024 + bytesA1 * 2 * sa + ((bytesBh * sb + 1023) & ~(size_t)1023) + 1024 + 512;
        if (plan->smem > 227 * 1024) return AM_ERR_ARG;
        const int clusters = p.n_work < sm_count() / 2 ? p.n_work : sm_count() / 2;
        plan->grid = 2 * clusters;
        return AM_OK;
    }
    const size_t pool_bytes = p.pool2 ? (size_t)4 * (d->NT / 16) * 32 * 32 + 16 : 0;      // kPOOL2 staging: [4 quarters][units][32 lanes] x 32 B
    const size_t budget = budget_total - pool_bytes;
    const size_t allB = bytesB * (size_t)total_chunks * d->KH;
    // Mode choice (mirrored by fcn_lecturenet.layer_cost):
    //   MT = 2 (two M-tiles per work item, one MMA issuer warp each) whenever there is enough work to keep every SM busy;
    //   resident weights when a single N block's whole packed filter fits next to >= 2 (MT = 2) / 3 (MT = 1) A stages.
    //   Narrow layers (N < 128) are issue bound, so for them two issuers beat resident weights if both do not fit.
    const bool tmem2 = 2 * p.NTc <= 512;
    const bool many = !(d->flags & AM_CONV_NO_MT2) && p.n_mtiles * p.nNB >= 4 * sm_count() && tmem2;
    const bool can_res = !(d->flags & AM_CONV_NO_RESIDENT) && p.nNB == 1;
    const bool res2 = can_res && fixed + allB + 2 * 2 * bytesA1 <= budget;
    const bool res1 = can_res && fixed + allB + (p.pool2 ? 2 : 3) * bytesA1 <= budget;
    int resident, MT;
    const bool res4 = can_res && fixed + allB + 2 * 4 * bytesA1 <= budget;
    if (p.pool2 || p.poolx) { MT = 1; resident = res1 ? 1 : 0; }
    else if ((d->flags & AM_CONV_FORCE_MT4) && 4 * p.NTc <= 512) { MT = 4; resident = res4 ? 1 : 0; }  // the planner decided
    else if ((d->flags & AM_CONV_FORCE_MT2) && tmem2) { MT = 2; resident = res2 ? 1 : 0; }
    else if (d->flags & AM_CONV_NO_MT2) { MT = 1; resident = res1 ? 1 : 0; }
    else if (res2 && many) { resident = 1; MT = 2; }
    else if (res1 && (d->NT >= 128 || !many)) { resident = 1; MT = 1; }
    else if (many) { resident = 0; MT = 2; }
    else { resident = res1 ? 1 : 0; MT = 1; }
    int sa, sb;
    if (resident) {
        sb = 1; sa = (MT >= 2 || p.pool2) ? 2 : 3;
        while (sa < 8 && fixed + allB + (size_t)(sa + 1) * MT * bytesA1 <= budget) ++sa;
    } else {
        sa = 2; sb = 2;
        while (true) {      // grow the rings alternately while they fit; B stages are consumed KH times faster
            bool grew = false;
            if (sb < 3 * d->KH && sb < 12 && fixed + bytesA1 * MT * sa + bytesB * (sb + 1) <= budget) { ++sb; grew = true; }
            if (sa < 4 && fixed + bytesA1 * MT * (sa + 1) + bytesB * sb <= budget) { ++sa; grew = true; }
            if (!grew) break;
        }
        if (fixed + bytesA1 * MT * sa + bytesB * sb > budget) {
            if (MT >= 2) { MT = 1; sa = 2; sb = 2; }
            if (fixed + bytesA1 * sa + bytesB * sb > budget) return AM_ERR_ARG;
        }
    }
    p.residentB = resident; p.MT = MT; p.stagesA = sa; p.stagesB = sb;
    p.acc_stages = (2 * MT * p.NTc <= 512) ? 2 : 1;
    p.tmem_cols = 32; while (p.tmem_cols < p.acc_stages * MT * p.NTc) p.tmem_cols <<= 1;
    p.n_work = ((p.n_mtiles + MT - 1) / MT) * p.nNB;
    const size_t nB = resident ? (size_t)total_chunks * d->KH : (size_t)sb;
    plan->smem = 1024 + bytesA1 * MT * sa + ((bytesB * nB + 1023) & ~(size_t)1023) + 1024 + 512 + pool_bytes;
    if (plan->smem > 227 * 1024) return AM_ERR_ARG;
    plan->grid = p.n_work < sm_count() ? p.n_work : sm_count();
    return AM_OK;
}

typedef void (*conv_kernel_t)(const ConvParams);
static int conv_launch(const am_conv_plan* plan, void* stream) {
    // [fused epilogue][MT 1 / 2 / 4][streamed / resident weights], then the two CTA-pair kernels
    static const conv_kernel_t kernels[18] = {
        k_conv_gemm<1, false, false>, k_conv_gemm<1, true, false>, k_conv_gemm<2, false, false>, k_conv_gemm<2, true, false>,
        k_conv_gemm<4, false, false>, k_conv_gemm<4, true, false>,
        k_conv_gemm<1, false, true>, k_conv_gemm<1, true, true>, k_conv_gemm<2, false, true>, k_conv_gemm<2, true, true>,
        k_conv_gemm<4, false, true>, k_conv_gemm<4, true, true>,
        k_conv_gemm_pair<false>, k_conv_gemm_pair<true>,
        k_conv_gemm<1, false, false, true>, k_conv_gemm<1, true, false, true>,
        k_conv_gemm<1, false, false, false, true>, k_conv_gemm<1, true, false, false, true>};
    static bool attr_set = false;
    if (!attr_set) {
        for (int i = 0; i < 18; ++i) AM_CUDA(cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024)));
        attr_set = true;
    }
    const int fused = plan->p.epi_mode != AM_EPI_PLAIN ? 1 : 0;
    const conv_kernel_t k = plan->pair ? kernels[12 + fused]                                               // __cluster_dims__(2,1,1)
                          : plan->p.pool2 ? kernels[14 + (plan->p.residentB ? 1 : 0)]
                          : plan->p.poolx ? kernels[16 + (plan->p.residentB ? 1 : 0)]
                                          : kernels[6 * fused + (plan->p.MT == 4 ? 4 : plan->p.MT == 2 ? 2 : 0) + (plan->p.residentB ? 1 : 0)];
    // programmatic dependent launch, see pdl_wait().  Opt-in (AM_B200_PDL=1): measured on a B200 it changes nothing (end to end 813 vs 812
    // frames/s, ABBA runs, tools/ab_pdl.sh): the persistent CTAs own every SM until they exit, so only the ~3 us prologue can overlap.
    static const int use_pdl = [] { const char* e = getenv("AM_B200_PDL"); return (e && e[0] == '1') ? 1 : 0; }();
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(plan->grid); cfg.blockDim = dim3(CONV_THREADS); cfg.dynamicSmemBytes = plan->smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = use_pdl ? 1 : 0;
    AM_CUDA(cudaLaunchKernelEx(&cfg, k, plan->p));
    return AM_OK;
}

// One convolution-as-GEMM launch.  Everything the kernel needs is in this plain-C descriptor (see header).
extern "C" int am_conv_gemm(const am_conv_desc* d, void* stream) {
    am_conv_plan plan;
    int rc = conv_prepare(d, &plan);
    if (rc) return rc;
    AM_CUDA(cudaMallocAsync(&plan.d_counter, 8, (cudaStream_t)stream));
    AM_CUDA(cudaMemsetAsync(plan.d_counter, 0, 8, (cudaStream_t)stream));
    plan.p.work_counter = plan.d_counter;
    rc = conv_launch(&plan, stream);
    AM_CUDA(cudaFreeAsync(plan.d_counter, (cudaStream_t)stream));
    return rc;
}

// Prepared form: tensor maps and tiling are computed once per layer, launches reuse them.
extern "C" am_conv_plan* am_conv_plan_create(const am_conv_desc* d) {
    am_conv_plan* plan = new am_conv_plan();
    if (conv_prepare(d, plan) != AM_OK) { delete plan; return nullptr; }
    if (cudaMalloc(&plan->d_counter, 8) != cudaSuccess || cudaMemset(plan->d_counter, 0, 8) != cudaSuccess) { delete plan; return nullptr; }
    plan->p.work_counter = plan->d_counter;
    return plan;
}
extern "C" void am_conv_plan_destroy(am_conv_plan* plan) {
    if (plan) { cudaFree(plan->d_counter); delete plan; }
}
extern "C" int am_conv_plan_launch(const am_conv_plan* plan, void* stream) {
    if (!plan) return AM_ERR_ARG;
    return conv_launch(plan, stream);
}
extern "C" int am_conv_plan_bind(am_conv_plan* plan, int which, void* ptr) {
    if (!plan) return AM_ERR_ARG;
    switch (which) {
    case AM_BIND_FRAMES: plan->p.frames = (const uint8_t*)ptr; break;
    case AM_BIND_TEXT_OUT: plan->p.text_out = (float*)ptr; break;
    case AM_BIND_REC_OUT: plan->p.rec_out = (float*)ptr; break;
    case AM_BIND_OUT:
        if (!ptr && plan->p.epi_mode == AM_EPI_PLAIN) return AM_ERR_ARG;
        plan->p.out = ptr; break;
    case AM_BIND_THRESHOLD: plan->p.threshold = (int)(intptr_t)ptr; break;
    default: return AM_ERR_ARG;
    }
    return AM_OK;
}
// info[8] = MT, residentB, acc_stages, stagesA, stagesB, grid, smem bytes, n_work
extern "C" int am_conv_plan_info(const am_conv_plan* plan, int* info) {
    if (!plan || !info) return AM_ERR_ARG;
    info[0] = plan->p.MT; info[1] = plan->p.residentB; info[2] = plan->p.acc_stages; info[3] = plan->p.stagesA;
    info[4] = plan->p.stagesB; info[5] = plan->grid; info[6] = (int)plan->smem; info[7] = plan->p.n_work;
    return AM_OK;
}

// fcn_misc.cu -- the memory-bound glue kernels of the FCN binarizer (sm_100a): input normalisation, 2x2 max-pool,
// transposed-conv border fill, heads post-processing (tanh / sigmoid / diff) and sigmoid-threshold bit-packing.
// All are single-pass, coalesced, 128-bit vectorised where the layout allows.
#include "am_common.cuh"
#include "../../include/accessmath_b200.h"
#include <cuda_bf16.h>

// x0 = (v/255 - 0.5)/0.5 in fp32, the op order of TF.to_tensor + TF.normalize (FCN_lecturenet.py:607-618)
__device__ __forceinline__ float norm_px(uint8_t v) { return ((float)v / 255.0f - 0.5f) / 0.5f; }

__global__ void k_prep_input(const uint8_t* __restrict__ bgr, int H, int W, __nv_bfloat16* __restrict__ out, int C, int pad) {
    const int f = blockIdx.z, y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const uint8_t* px = bgr + (((size_t)f * H + y) * W + x) * 3;
    __nv_bfloat16* o = out + (((size_t)f * H + y) * (W + 2 * pad) + x + pad) * C;
    float r = norm_px(px[2]), g = norm_px(px[1]), b = norm_px(px[0]);      // BGR -> RGB (FCN_lecturenet_binarizer.py:50)
    __nv_bfloat162 h0 = __floats2bfloat162_rn(r, g), h1 = __floats2bfloat162_rn(b, 0.0f);
    if (C == 8) {
        uint4 u; u.x = *(uint32_t*)&h0; u.y = *(uint32_t*)&h1; u.z = 0; u.w = 0;
        *(uint4*)o = u;
    } else {
        o[0] = __float2bfloat16_rn(r); o[1] = __float2bfloat16_rn(g); o[2] = __float2bfloat16_rn(b);
        for (int c = 3; c < C; ++c) o[c] = __float2bfloat16_rn(0.0f);
    }
}

__device__ __forceinline__ uint4 max8(uint4 a, uint4 b) {
    uint4 r;
    __nv_bfloat162 t;
    t = __hmax2(*(__nv_bfloat162*)&a.x, *(__nv_bfloat162*)&b.x); r.x = *(uint32_t*)&t;
    t = __hmax2(*(__nv_bfloat162*)&a.y, *(__nv_bfloat162*)&b.y); r.y = *(uint32_t*)&t;
    t = __hmax2(*(__nv_bfloat162*)&a.z, *(__nv_bfloat162*)&b.z); r.z = *(uint32_t*)&t;
    t = __hmax2(*(__nv_bfloat162*)&a.w, *(__nv_bfloat162*)&b.w); r.w = *(uint32_t*)&t;
    return r;
}
// one thread per (output pixel, 8-channel group)
__global__ void k_maxpool2(const __nv_bfloat16* __restrict__ in, int H, int W, int C, int pad_in, __nv_bfloat16* __restrict__ out,
                           int Ho, int Wo, int pad_out) {
    const int f = blockIdx.z, yo = blockIdx.y;
    const int cg = C >> 3;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Wo * cg) return;
    const int xo = i / cg, c8 = (i - xo * cg) * 8;
    const size_t rs = (size_t)(W + 2 * pad_in) * C;
    const __nv_bfloat16* p = in + ((size_t)f * H + 2 * yo) * rs + (size_t)(2 * xo + pad_in) * C + c8;
    uint4 a = *(const uint4*)p, b = *(const uint4*)(p + C), c = *(const uint4*)(p + rs), d = *(const uint4*)(p + rs + C);
    uint4 m = max8(max8(a, b), max8(c, d));
    *(uint4*)(out + (((size_t)f * Ho + yo) * (Wo + 2 * pad_out) + xo + pad_out) * C + c8) = m;
}

__global__ void k_fill_border(__nv_bfloat16* __restrict__ buf, int H, int W, int C, int pad, int y_from, int x_from,
                              const __nv_bfloat16* __restrict__ vals) {
    const int f = blockIdx.z, y = blockIdx.y;
    const int cg = C >> 3;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= W * cg) return;
    const int x = i / cg, c8 = (i - x * cg) * 8;
    if (y < y_from && x < x_from) return;
    *(uint4*)(buf + (((size_t)f * H + y) * (W + 2 * pad) + x + pad) * C + c8) = *(const uint4*)(vals + c8);
}

__global__ void k_heads_post(const float* __restrict__ heads, const uint8_t* __restrict__ bgr, int H, int W,
                             __nv_bfloat16* __restrict__ diff, int C, int pad, float* __restrict__ text_out, float* __restrict__ rec_out) {
    const int f = blockIdx.z, y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= W) return;
    const size_t pix = ((size_t)f * H + y) * W + x;
    const float4 h = *(const float4*)(heads + pix * 4);
    const uint8_t* px = bgr + pix * 3;
    const float s = 1.0f / (1.0f + expf(-h.x));                          // torch.sigmoid(text_mask), :372
    const float r0 = tanhf(h.y), r1 = tanhf(h.z), r2 = tanhf(h.w);       // conv_reconstruct ... nn.Tanh, :153-160
    const float d0 = (norm_px(px[2]) - r0) * s, d1 = (norm_px(px[1]) - r1) * s, d2 = (norm_px(px[0]) - r2) * s;   // :377
    __nv_bfloat16* o = diff + (((size_t)f * H + y) * (W + 2 * pad) + x + pad) * C;
    __nv_bfloat162 h0 = __floats2bfloat162_rn(d0, d1), h1 = __floats2bfloat162_rn(d2, 0.0f);
    if (C == 8) {
        uint4 u; u.x = *(uint32_t*)&h0; u.y = *(uint32_t*)&h1; u.z = 0; u.w = 0;
        *(uint4*)o = u;
    } else {
        o[0] = __float2bfloat16_rn(d0); o[1] = __float2bfloat16_rn(d1); o[2] = __float2bfloat16_rn(d2);
        for (int c = 3; c < C; ++c) o[c] = __float2bfloat16_rn(0.0f);
    }
    if (text_out) text_out[pix] = h.x;
    if (rec_out) { rec_out[pix * 3] = r0; rec_out[pix * 3 + 1] = r1; rec_out[pix * 3 + 2] = r2; }
}

// binary = (uint8)(sigmoid(z)*255) >= thr ? 255 : 0 (FCN_lecturenet.py:461-467); ink = 255 - binary
__global__ void k_threshold_pack(const float* __restrict__ logits, int H, int W, int WPR, int thr, uint32_t* __restrict__ bits) {
    const int f = blockIdx.z, y = blockIdx.y;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    bool ink = false;
    if (x < W) {
        float z = logits[((size_t)f * H + y) * W + x];
        float s = 1.0f / (1.0f + expf(-z));
        int u8 = (int)(s * 255.0f);
        ink = u8 < thr;
    }
    unsigned m = __ballot_sync(0xffffffffu, ink);
    if ((threadIdx.x & 31) == 0 && (x >> 5) < WPR) bits[((size_t)f * H + y) * WPR + (x >> 5)] = m;
}

static inline cudaStream_t S(void* s) { return (cudaStream_t)s; }

extern "C" int am_fcn_prep_input(const uint8_t* d_bgr, int batch, int height, int width, void* d_out, int C, int pad, void* stream) {
    if (!d_bgr || !d_out || C < 3 || batch <= 0) return AM_ERR_ARG;
    k_prep_input<<<dim3(am_div_up(width, 128), height, batch), 128, 0, S(stream)>>>(d_bgr, height, width, (__nv_bfloat16*)d_out, C, pad);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}
extern "C" int am_fcn_maxpool2(const void* d_in, int batch, int height, int width, int C, int pad_in, void* d_out, int pad_out, void* stream) {
    if (!d_in || !d_out || C % 8 != 0 || batch <= 0) return AM_ERR_ARG;
    const int Ho = height / 2, Wo = width / 2;
    if (Ho == 0 || Wo == 0) return AM_ERR_ARG;
    k_maxpool2<<<dim3(am_div_up((long long)Wo * (C / 8), 128), Ho, batch), 128, 0, S(stream)>>>((const __nv_bfloat16*)d_in, height, width, C, pad_in,
                                                                                            (__nv_bfloat16*)d_out, Ho, Wo, pad_out);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}
extern "C" int am_fcn_fill_border(void* d_buf, int batch, int height, int width, int C, int pad, int y_from, int x_from,
                                  const void* d_values, void* stream) {
    if (!d_buf || !d_values || C % 8 != 0) return AM_ERR_ARG;
    if (y_from >= height && x_from >= width) return AM_OK;
    k_fill_border<<<dim3(am_div_up((long long)width * (C / 8), 128), height, batch), 128, 0, S(stream)>>>((__nv_bfloat16*)d_buf, height, width, C, pad,
                                                                                                   y_from, x_from, (const __nv_bfloat16*)d_values);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}
extern "C" int am_fcn_heads_post(const float* d_heads, const uint8_t* d_bgr, int batch, int height, int width, void* d_diff, int C,
                                 int pad, float* d_text_logit, float* d_rec, void* stream) {
    if (!d_heads || !d_bgr || !d_diff || C < 3) return AM_ERR_ARG;
    k_heads_post<<<dim3(am_div_up(width, 128), height, batch), 128, 0, S(stream)>>>(d_heads, d_bgr, height, width, (__nv_bfloat16*)d_diff, C, pad,
                                                                              d_text_logit, d_rec);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}
extern "C" int am_fcn_threshold_pack(const float* d_logits, int batch, int height, int width, int threshold, uint32_t* d_bits, void* stream) {
    if (!d_logits || !d_bits) return AM_ERR_ARG;
    const int WPR = am_words_per_row_impl(width);
    k_threshold_pack<<<dim3(am_div_up((long long)WPR * 32, 256), height, batch), 256, 0, S(stream)>>>(d_logits, height, width, WPR, threshold, d_bits);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

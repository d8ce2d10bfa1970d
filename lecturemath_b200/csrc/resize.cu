// resize.cu -- the > 2.5 MP branch of FCN_LectureNet.binarize on the device (sm_100a), bit-identical to the third-party
// operators the reference calls there:
//   * am_lanczos_resize_u8   = PIL.Image.resize((w, h), LANCZOS) on uint8 interleaved images
//                              (R/AccessMath/lecturenet_v1/FCN_lecturenet.py:434-437).  Pillow's algorithm (libImaging/Resample.c):
//                              per output position a window [center - support, center + support) of Lanczos-3 weights computed in
//                              double, normalised and rounded to 22-bit fixed point; a horizontal pass rounded to uint8, then a
//                              vertical pass.  The tables are built on the host with the same libm calls, both passes are fused in
//                              one kernel (tile + halo staged in shared memory, horizontal result kept in shared memory).
//   * am_bits_resize_nearest = cv2.resize(mask, (w, h), interpolation=INTER_NEAREST) on bit-packed masks (FCN_lecturenet.py:481-486):
//                              source index = min(floor(dst * (1 / (dsize / (double) ssize))), ssize - 1) per axis.
// Both are single-pass HBM-bound kernels; at 3840x2160 they move 31 MB / 1.3 MB per frame next to ~1.3 ms of FCN tensor time.
#include "am_common.cuh"
#include "../../include/accessmath_b200.h"
#include <cmath>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

namespace {

constexpr int kPrecisionBits = 32 - 8 - 2;       // Pillow: PRECISION_BITS
constexpr int kTW = 64, kTH = 16;                // output tile of one CTA
constexpr int kThreads = 256;

struct AxisTable {                               // host copy + device copy of one axis' bounds / coefficients
    std::vector<int> bounds, coef;
    int ksize = 0, max_span = 0;                 // max_span: input positions touched by `tile` consecutive outputs
    int *d_bounds = nullptr, *d_coef = nullptr;
};

double sinc_filter(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return sin(x) / x;
}
double lanczos_filter(double x) {
    if (-3.0 <= x && x < 3.0) return sinc_filter(x) * sinc_filter(x / 3);
    return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for box = (0, in_size)
void build_axis(int in_size, int out_size, int tile, AxisTable& t) {
    const float in0 = 0.0f, in1 = (float)in_size;
    double scale = (double)(in1 - in0) / out_size, filterscale = scale;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 3.0 * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    t.ksize = ksize;
    t.bounds.assign((size_t)out_size * 2, 0);
    t.coef.assign((size_t)out_size * ksize, 0);
    std::vector<double> k(ksize);
    for (int xx = 0; xx < out_size; ++xx) {
        const double center = in0 + (xx + 0.5) * scale, ss = 1.0 / filterscale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            const double w = lanczos_filter((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) k[x] /= ww;
            t.coef[(size_t)xx * ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << kPrecisionBits)) : (int)(0.5 + k[x] * (1 << kPrecisionBits));
        }
        t.bounds[2 * xx] = xmin;
        t.bounds[2 * xx + 1] = xmax;
    }
    if (in_size == out_size) {                   // Pillow skips the pass: a single unit tap reproduces the input exactly
        for (int xx = 0; xx < out_size; ++xx) {
            t.bounds[2 * xx] = xx; t.bounds[2 * xx + 1] = 1;
            for (int x = 0; x < ksize; ++x) t.coef[(size_t)xx * ksize + x] = x == 0 ? (1 << kPrecisionBits) : 0;
        }
    }
    t.max_span = 0;
    for (int o0 = 0; o0 < out_size; o0 += tile) {
        const int o1 = (o0 + tile < out_size ? o0 + tile : out_size) - 1;
        int lo = t.bounds[2 * o0], hi = 0;
        for (int o = o0; o <= o1; ++o) {
            if (t.bounds[2 * o] < lo) lo = t.bounds[2 * o];
            if (t.bounds[2 * o] + t.bounds[2 * o + 1] > hi) hi = t.bounds[2 * o] + t.bounds[2 * o + 1];
        }
        if (hi - lo > t.max_span) t.max_span = hi - lo;
    }
}

int upload(AxisTable& t) {
    AM_CUDA(cudaMalloc(&t.d_bounds, t.bounds.size() * sizeof(int)));
    AM_CUDA(cudaMalloc(&t.d_coef, t.coef.size() * sizeof(int)));
    AM_CUDA(cudaMemcpy(t.d_bounds, t.bounds.data(), t.bounds.size() * sizeof(int), cudaMemcpyHostToDevice));
    AM_CUDA(cudaMemcpy(t.d_coef, t.coef.data(), t.coef.size() * sizeof(int), cudaMemcpyHostToDevice));
    return AM_OK;
}

std::mutex g_mu;
std::map<std::tuple<int, int, int, int>, AxisTable*> g_axes;          // (device, in, out, tile) -> table, lives until exit
std::map<std::tuple<int, int, int>, int*> g_nn;                       // (device, src, dst) -> nearest source offsets

int axis_table(int in_size, int out_size, int tile, AxisTable** out) {
    int dev = 0;
    AM_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(dev, in_size, out_size, tile);
    auto it = g_axes.find(key);
    if (it == g_axes.end()) {
        AxisTable* t = new AxisTable();
        build_axis(in_size, out_size, tile, *t);
        int rc = upload(*t);
        if (rc) { delete t; return rc; }
        it = g_axes.emplace(key, t).first;
    }
    *out = it->second;
    return AM_OK;
}

// resize.cpp resizeNN: x_ofs[x] = min(cvFloor(x * ifx), ssize - 1), ifx = 1 / (dsize / (double) ssize)
int nearest_table(int src, int dst, int** out) {
    int dev = 0;
    AM_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_tuple(dev, src, dst);
    auto it = g_nn.find(key);
    if (it == g_nn.end()) {
        std::vector<int> ofs(dst);
        const double inv_scale = dst / (double)src, ifx = 1.0 / inv_scale;
        for (int x = 0; x < dst; ++x) {
            int s = (int)floor(x * ifx);
            ofs[x] = s < src - 1 ? s : src - 1;
        }
        int* d = nullptr;
        AM_CUDA(cudaMalloc(&d, (size_t)dst * sizeof(int)));
        AM_CUDA(cudaMemcpy(d, ofs.data(), (size_t)dst * sizeof(int), cudaMemcpyHostToDevice));
        it = g_nn.emplace(key, d).first;
    }
    *out = it->second;
    return AM_OK;
}

__device__ __forceinline__ int clip8(int acc) {
    int v = acc >> kPrecisionBits;                                     // arithmetic shift, as Pillow's clip8()
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// One CTA = one kTH x kTW output tile of one frame.  Shared memory: the input window (rows x row_bytes, loaded as aligned
// 32-bit words), the horizontally resampled window (rows x kTW*C uint8) and the tile's coefficient rows of both axes.
// Horizontal pass: thread = (window row, output x), all C channels -- one coefficient load feeds C multiply-adds.
template <int C>
__global__ void __launch_bounds__(kThreads)
k_lanczos_resize(const uint8_t* __restrict__ in, int lead, long long total_bytes, int in_h, int in_w,
                 const int* __restrict__ bx, const int* __restrict__ kx, int ksx,
                 const int* __restrict__ by, const int* __restrict__ ky, int ksy,
                 int out_h, int out_w, uint8_t* __restrict__ out, int win_words, int win_rows) {
    extern __shared__ uint32_t smem[];
    uint32_t* win = smem;                                              // [win_rows][win_words]
    int* s_kx = (int*)(smem + (size_t)win_rows * win_words);           // [kTW][ksx]
    int* s_ky = s_kx + kTW * ksx;                                      // [kTH][ksy]
    int* s_bx = s_ky + kTH * ksy;                                      // [kTW][2]
    int* s_by = s_bx + 2 * kTW;                                        // [kTH][2]
    uint8_t* hbuf = (uint8_t*)(s_by + 2 * kTH);                        // [win_rows][kTW * C]
    const int f = blockIdx.z, oy0 = blockIdx.y * kTH, ox0 = blockIdx.x * kTW;
    const int oy1 = min(oy0 + kTH, out_h), ox1 = min(ox0 + kTW, out_w);
    const int tw = ox1 - ox0, th = oy1 - oy0;
    for (int i = threadIdx.x; i < tw * ksx; i += kThreads) s_kx[i] = kx[(size_t)ox0 * ksx + i];
    for (int i = threadIdx.x; i < th * ksy; i += kThreads) s_ky[i] = ky[(size_t)oy0 * ksy + i];
    for (int i = threadIdx.x; i < 2 * tw; i += kThreads) s_bx[i] = bx[2 * ox0 + i];
    for (int i = threadIdx.x; i < 2 * th; i += kThreads) s_by[i] = by[2 * oy0 + i];
    // input window of this tile (bounds are monotone in the output index)
    const int x_lo = bx[2 * ox0], x_hi = bx[2 * (ox1 - 1)] + bx[2 * (ox1 - 1) + 1];
    const int y_lo = by[2 * oy0], y_hi = by[2 * (oy1 - 1)] + by[2 * (oy1 - 1) + 1];
    const int rows = y_hi - y_lo;
    const long long frame0 = lead + (long long)f * in_h * in_w * C;    // `in` is 4-byte aligned, the images start `lead` bytes in
    // phase 0: stage rows [y_lo, y_hi) x bytes [x_lo*C, x_hi*C) through aligned word loads
    for (int r = threadIdx.x / 32; r < rows; r += kThreads / 32) {
        const long long b0 = frame0 + ((long long)(y_lo + r) * in_w + x_lo) * C, b1 = b0 + (long long)(x_hi - x_lo) * C;
        const long long a0 = b0 & ~3LL;
        const int nw = (int)((b1 - a0 + 3) >> 2);
        for (int w = threadIdx.x & 31; w < nw; w += 32) {
            const long long a = a0 + 4LL * w;
            uint32_t v;
            if (a + 4 <= total_bytes) v = *(const uint32_t*)(in + a);
            else {
                v = 0;
                for (int j = 0; j < 4; ++j) if (a + j < total_bytes) v |= (uint32_t)in[a + j] << (8 * j);
            }
            win[(size_t)r * win_words + w] = v;
        }
    }
    __syncthreads();
    // phase 1: horizontal pass -> hbuf (uint8, as ImagingResampleHorizontal_8bpc)
    for (int i = threadIdx.x; i < rows * tw; i += kThreads) {
        const int r = i / tw, ox = i - r * tw;
        const int xmin = s_bx[2 * ox], n = s_bx[2 * ox + 1];
        const int* k = s_kx + ox * ksx;
        const long long b0 = frame0 + ((long long)(y_lo + r) * in_w + x_lo) * C;
        const uint8_t* px = (const uint8_t*)(win + (size_t)r * win_words) + (int)(b0 & 3) + (xmin - x_lo) * C;
        int acc[C];
#pragma unroll
        for (int c = 0; c < C; ++c) acc[c] = 1 << (kPrecisionBits - 1);
#pragma unroll 4
        for (int t = 0; t < n; ++t) {
            const int kt = k[t];
#pragma unroll
            for (int c = 0; c < C; ++c) acc[c] += (int)px[t * C + c] * kt;
        }
        uint8_t* dst = hbuf + (size_t)r * (kTW * C) + ox * C;
#pragma unroll
        for (int c = 0; c < C; ++c) dst[c] = (uint8_t)clip8(acc[c]);
    }
    __syncthreads();
    // phase 2: vertical pass (ImagingResampleVertical_8bpc), coalesced row-segment stores
    const int twc = tw * C;
    for (int i = threadIdx.x; i < th * twc; i += kThreads) {
        const int oy = i / twc, j = i - oy * twc;
        const int ymin = s_by[2 * oy], n = s_by[2 * oy + 1];
        const int* k = s_ky + oy * ksy;
        const uint8_t* col = hbuf + (size_t)(ymin - y_lo) * (kTW * C) + j;
        int acc = 1 << (kPrecisionBits - 1);
#pragma unroll 4
        for (int t = 0; t < n; ++t) acc += (int)col[(size_t)t * (kTW * C)] * k[t];
        out[(((long long)f * out_h + oy0 + oy) * out_w + ox0) * C + j] = (uint8_t)clip8(acc);
    }
}

// Fast path for ksize <= kKMax on both axes (scale factors up to 2.33x: every halving of the 2.5 MP guard).  Same tiles and shared
// memory layout as k_lanczos_resize; the passes are re-mapped so that shared-memory LOAD count -- the bound of the generic kernel --
// drops ~5x: horizontal: thread = one output x for every 4th window row, coefficients in registers, pixels read as aligned 32-bit
// words + funnel shift; vertical: thread = one output row and C words (4 bytes each) of it, coefficients in registers.
constexpr int kKMax = 16;
template <int C>
__global__ void __launch_bounds__(kThreads)
k_lanczos_fast(const uint8_t* __restrict__ in, int lead, long long total_bytes, int in_h, int in_w,
               const int* __restrict__ bx, const int* __restrict__ kx, int ksx,
               const int* __restrict__ by, const int* __restrict__ ky, int ksy,
               int out_h, int out_w, uint8_t* __restrict__ out, int win_words, int win_rows) {
    extern __shared__ uint32_t smem[];
    uint32_t* win = smem;                                              // [win_rows][win_words]
    uint32_t* hbuf = smem + (size_t)win_rows * win_words;              // [win_rows][kTW * C / 4]
    constexpr int kRowWords = kTW * C / 4;
    const int f = blockIdx.z, oy0 = blockIdx.y * kTH, ox0 = blockIdx.x * kTW;
    const int oy1 = min(oy0 + kTH, out_h), ox1 = min(ox0 + kTW, out_w);
    const int tw = ox1 - ox0, th = oy1 - oy0;
    const int x_lo = bx[2 * ox0], x_hi = bx[2 * (ox1 - 1)] + bx[2 * (ox1 - 1) + 1];
    const int y_lo = by[2 * oy0], y_hi = by[2 * (oy1 - 1)] + by[2 * (oy1 - 1) + 1];
    const int rows = y_hi - y_lo;
    const long long frame0 = lead + (long long)f * in_h * in_w * C;
    // phase 0: window rows through aligned word loads (one spare word per row is cleared for the funnel shift)
    for (int r = threadIdx.x / 32; r < rows; r += kThreads / 32) {
        const long long b0 = frame0 + ((long long)(y_lo + r) * in_w + x_lo) * C, b1 = b0 + (long long)(x_hi - x_lo) * C;
        const long long a0 = b0 & ~3LL;
        const int nw = (int)((b1 - a0 + 3) >> 2);
        for (int w = threadIdx.x & 31; w < win_words; w += 32) {
            uint32_t v = 0;
            const long long a = a0 + 4LL * w;
            if (w < nw) {
                if (a + 4 <= total_bytes) v = *(const uint32_t*)(in + a);
                else
                    for (int j = 0; j < 4; ++j) if (a + j < total_bytes) v |= (uint32_t)in[a + j] << (8 * j);
            }
            win[(size_t)r * win_words + w] = v;
        }
    }
    __syncthreads();
    // phase 1: horizontal pass
    {
        const int ox = threadIdx.x % kTW, rl = threadIdx.x / kTW;     // 4 row lanes
        if (ox < tw) {
            const int xmin = bx[2 * (ox0 + ox)], n = bx[2 * (ox0 + ox) + 1];
            int kreg[kKMax];
#pragma unroll
            for (int t = 0; t < kKMax; ++t) kreg[t] = t < n ? kx[(size_t)(ox0 + ox) * ksx + t] : 0;
            constexpr int kNW = (kKMax * C + 3) / 4 + 1;                // words that can hold the taps at any byte phase
            for (int r = rl; r < rows; r += kThreads / kTW) {
                const long long b0 = frame0 + ((long long)(y_lo + r) * in_w + x_lo) * C;
                const int s = (int)(b0 & 3) + (xmin - x_lo) * C, sh = 8 * (s & 3);
                const uint32_t* src = win + (size_t)r * win_words + (s >> 2);
                const int lim = win_words - (s >> 2);                 // words left in this row of the window
                uint32_t w[kNW];
#pragma unroll
                for (int j = 0; j < kNW; ++j) w[j] = j < lim ? src[j] : 0u;
                int acc[C];
#pragma unroll
                for (int c = 0; c < C; ++c) acc[c] = 1 << (kPrecisionBits - 1);
#pragma unroll
                for (int t = 0; t < kKMax; ++t) {
#pragma unroll
                    for (int c = 0; c < C; ++c) {
                        const int i = t * C + c;                       // byte i of the tap stream (compile-time)
                        const uint32_t al = __funnelshift_r(w[i >> 2], w[(i >> 2) + 1], sh);
                        acc[c] += (int)((al >> (8 * (i & 3))) & 0xFFu) * kreg[t];
                    }
                }
                uint8_t* dst = (uint8_t*)(hbuf + (size_t)r * kRowWords) + ox * C;
#pragma unroll
                for (int c = 0; c < C; ++c) dst[c] = (uint8_t)clip8(acc[c]);
            }
        }
    }
    __syncthreads();
    // phase 2: vertical pass: thread = (output row, C words of it)
    {
        const int oy = threadIdx.x / 16, wl = threadIdx.x % 16;
        if (oy < th) {
            const int ymin = by[2 * (oy0 + oy)], n = by[2 * (oy0 + oy) + 1];
            int kreg[kKMax];
#pragma unroll
            for (int t = 0; t < kKMax; ++t) kreg[t] = t < n ? ky[(size_t)(oy0 + oy) * ksy + t] : 0;
            const int twc = tw * C;
            uint8_t* orow = out + (((long long)f * out_h + oy0 + oy) * out_w + ox0) * C;
            const bool aligned = (((uintptr_t)orow) & 3) == 0;
#pragma unroll
            for (int q = 0; q < C; ++q) {
                const int wi = wl + 16 * q;
                if (4 * wi >= twc) continue;
                const uint32_t* col = hbuf + (size_t)(ymin - y_lo) * kRowWords + wi;
                int a0 = 1 << (kPrecisionBits - 1), a1 = a0, a2 = a0, a3 = a0;
#pragma unroll
                for (int t = 0; t < kKMax; ++t) {
                    if (t < n) {
                        const uint32_t v = col[(size_t)t * kRowWords];
                        a0 += (int)(v & 0xFFu) * kreg[t]; a1 += (int)((v >> 8) & 0xFFu) * kreg[t];
                        a2 += (int)((v >> 16) & 0xFFu) * kreg[t]; a3 += (int)(v >> 24) * kreg[t];
                    }
                }
                const uint32_t pk = (uint32_t)clip8(a0) | ((uint32_t)clip8(a1) << 8) | ((uint32_t)clip8(a2) << 16) | ((uint32_t)clip8(a3) << 24);
                if (aligned && 4 * wi + 4 <= twc) *(uint32_t*)(orow + 4 * wi) = pk;
                else
                    for (int j = 0; j < 4; ++j) if (4 * wi + j < twc) orow[4 * wi + j] = (uint8_t)(pk >> (8 * j));
            }
        }
    }
}

__device__ __forceinline__ uint32_t spread16(uint32_t x) {            // abcd -> 0a0b0c0d
    x = (x | (x << 8)) & 0x00ff00ffu;
    x = (x | (x << 4)) & 0x0f0f0f0fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// one thread per output word; exact2x: every source bit is doubled in x (and every row in y through y_ofs)
__global__ void k_bits_resize_nearest(const uint32_t* __restrict__ src, int in_h, int wpr_in, int out_h, int out_w, int wpr_out,
                                      const int* __restrict__ x_ofs, const int* __restrict__ y_ofs, int exact2x,
                                      uint32_t* __restrict__ dst) {
    const int f = blockIdx.z, oy = blockIdx.y, ow = blockIdx.x * blockDim.x + threadIdx.x;
    if (ow >= wpr_out) return;
    const uint32_t* row = src + ((size_t)f * in_h + y_ofs[oy]) * wpr_in;
    uint32_t v = 0;
    if (exact2x) {
        if (ow * 32 < out_w) {
            const uint32_t h = (row[ow >> 1] >> (16 * (ow & 1))) & 0xffffu, s = spread16(h);
            v = s | (s << 1);
            const int rem = out_w - ow * 32;
            if (rem < 32) v &= (1u << rem) - 1u;
        }
    } else {
        for (int b = 0; b < 32; ++b) {
            const int ox = ow * 32 + b;
            if (ox < out_w) {
                const int sx = x_ofs[ox];
                v |= ((row[sx >> 5] >> (sx & 31)) & 1u) << b;
            }
        }
    }
    dst[((size_t)f * out_h + oy) * wpr_out + ow] = v;
}

}  // namespace

extern "C" int am_fcn_working_size(int width, int height, int* out_width, int* out_height) {
    int n = 0;
    while ((long long)width * height > 2500000LL) {                    // FCN_lecturenet.py:435-437: int(w / 2), int(h / 2)
        width = width / 2;
        height = height / 2;
        ++n;
    }
    if (out_width) *out_width = width;
    if (out_height) *out_height = height;
    return n;
}

extern "C" int am_lanczos_resize_u8(const uint8_t* d_in, int batch, int in_h, int in_w, int channels, int out_h, int out_w,
                                    uint8_t* d_out, void* stream) {
    if (!d_in || !d_out || batch <= 0 || in_h <= 0 || in_w <= 0 || out_h <= 0 || out_w <= 0 || channels < 1 || channels > 4)
        return AM_ERR_ARG;
    AxisTable *tx = nullptr, *ty = nullptr;
    int rc = axis_table(in_w, out_w, kTW, &tx);
    if (rc) return rc;
    rc = axis_table(in_h, out_h, kTH, &ty);
    if (rc) return rc;
    const int win_words = (tx->max_span * channels + 3 + 3) / 4 + 2, win_rows = ty->max_span;    // + spare words (funnel shift)
    const size_t smem = (size_t)win_rows * win_words * 4 + ((size_t)kTW * tx->ksize + (size_t)kTH * ty->ksize + 2 * kTW + 2 * kTH) * 4 +
                        (size_t)win_rows * kTW * channels;
    if (smem > 200 * 1024) {
        fprintf(stderr, "[accessmath_b200] am_lanczos_resize_u8: scale too large for one tile (%zu B of shared memory)\n", smem);
        return AM_ERR_ARG;
    }
    dim3 grid(am_div_up(out_w, kTW), am_div_up(out_h, kTH), batch);
    const int lead = (int)((uintptr_t)d_in & 3);
    const long long total = lead + (long long)batch * in_h * in_w * channels;
    const bool fast = tx->ksize <= kKMax && ty->ksize <= kKMax;
    const size_t smem_fast = (size_t)win_rows * win_words * 4 + (size_t)win_rows * kTW * channels;
#define AM_LANCZOS_LAUNCH(CH)                                                                                                       \
    do {                                                                                                                            \
        if (fast) {                                                                                                                 \
            if (smem_fast > 48 * 1024)                                                                                              \
                AM_CUDA(cudaFuncSetAttribute(k_lanczos_fast<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_fast));     \
            k_lanczos_fast<CH><<<grid, kThreads, smem_fast, (cudaStream_t)stream>>>(d_in - lead, lead, total, in_h, in_w, tx->d_bounds,     \
                                                                                   tx->d_coef, tx->ksize, ty->d_bounds, ty->d_coef,         \
                                                                                   ty->ksize, out_h, out_w, d_out, win_words, win_rows);    \
        } else {                                                                                                                    \
            if (smem > 48 * 1024)                                                                                                   \
                AM_CUDA(cudaFuncSetAttribute(k_lanczos_resize<CH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
            k_lanczos_resize<CH><<<grid, kThreads, smem, (cudaStream_t)stream>>>(d_in - lead, lead, total, in_h, in_w, tx->d_bounds,        \
                                                                               tx->d_coef, tx->ksize, ty->d_bounds, ty->d_coef, ty->ksize,  \
                                                                               out_h, out_w, d_out, win_words, win_rows);                   \
        }                                                                                                                           \
    } while (0)
    switch (channels) {
        case 1: AM_LANCZOS_LAUNCH(1); break;
        case 2: AM_LANCZOS_LAUNCH(2); break;
        case 3: AM_LANCZOS_LAUNCH(3); break;
        default: AM_LANCZOS_LAUNCH(4); break;
    }
#undef AM_LANCZOS_LAUNCH
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

extern "C" int am_bits_resize_nearest(const uint32_t* d_bits, int batch, int in_h, int in_w, int out_h, int out_w,
                                      uint32_t* d_out, void* stream) {
    if (!d_bits || !d_out || batch <= 0 || in_h <= 0 || in_w <= 0 || out_h <= 0 || out_w <= 0) return AM_ERR_ARG;
    int *xo = nullptr, *yo = nullptr;
    int rc = nearest_table(in_w, out_w, &xo);
    if (rc) return rc;
    rc = nearest_table(in_h, out_h, &yo);
    if (rc) return rc;
    const int wpr_in = am_words_per_row_impl(in_w), wpr_out = am_words_per_row_impl(out_w);
    dim3 grid(am_div_up(wpr_out, 128), out_h, batch);
    k_bits_resize_nearest<<<grid, 128, 0, (cudaStream_t)stream>>>(d_bits, in_h, wpr_in, out_h, out_w, wpr_out, xo, yo,
                                                                out_w == 2 * in_w ? 1 : 0, d_out);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}


// ------------------------------------------------------------------------------------------------------------------------------
// am_resize_linear_u8 = cv2.resize(frame, (w, h)) -- INTER_LINEAR, the forced-resolution resize of VideoProcessor.doProcessing
// (R/AccessMath/preprocessing/video_processor/video_processor.py:164-165) -- on uint8 interleaved frames.  OpenCV's 8-bit algorithm
// (imgproc/resize.cpp, resizeGeneric_ + HResizeLinear / VResizeLinear for uchar): per axis source index floor(f) and fraction of
// f = (float)((d + 0.5) * scale - 0.5), scale = 1 / (dsize / (double) ssize), clamped at the borders; coefficients
// cvRound((1 - frac) * 2048), cvRound(frac * 2048) as int16; horizontal pass in int32, vertical pass
// (((b0 * (S0 >> 4)) >> 16) + ((b1 * (S1 >> 4)) >> 16) + 2) >> 2.  Tables are built on the host with the same float operations.
// Bit-identical to OpenCV's own generic code path (exact against cv2.resize whenever it takes that path: every down-scale); the
// build in this image sends up-scales through IPP, whose result differs by 1 grey level on ~0.1 % of the pixels (tests state it).
namespace {
struct LinearAxis { int* d_tab = nullptr; };     // [3 * n]: source index, coefficient 0, coefficient 1 per output position
std::mutex g_lin_mu;
std::map<std::tuple<int, int, int>, LinearAxis> g_lin;   // (device, in, out)

int linear_axis(int in, int out, int** d_tab) {
    int dev = 0;
    AM_CUDA(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_lin_mu);
    auto key = std::make_tuple(dev, in, out);
    auto it = g_lin.find(key);
    if (it == g_lin.end()) {
        std::vector<int> t(3 * (size_t)out);
        const double inv_scale = (double)out / in, scale = 1.0 / inv_scale;
        for (int d = 0; d < out; ++d) {
            float f = (float)((d + 0.5) * scale - 0.5);
            int sidx = (int)std::floor(f);
            f -= sidx;
            if (sidx < 0) { f = 0; sidx = 0; }
            if (sidx >= in - 1) { f = 0; sidx = in - 1; }
            t[3 * d] = sidx;
            t[3 * d + 1] = (int)std::lrintf((1.f - f) * 2048.f);      // saturate_cast<short>(float) = cvRound (round half to even)
            t[3 * d + 2] = (int)std::lrintf(f * 2048.f);
        }
        LinearAxis a;
        AM_CUDA(cudaMalloc(&a.d_tab, t.size() * sizeof(int)));
        AM_CUDA(cudaMemcpy(a.d_tab, t.data(), t.size() * sizeof(int), cudaMemcpyHostToDevice));
        it = g_lin.emplace(key, a).first;
    }
    *d_tab = it->second.d_tab;
    return AM_OK;
}

template <int CH>
__global__ void k_resize_linear(const uint8_t* __restrict__ in, int in_h, int in_w, const int* __restrict__ tx, const int* __restrict__ ty,
                                int out_h, int out_w, uint8_t* __restrict__ out) {
    const int f = blockIdx.z, y = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= out_w) return;
    const int sx = tx[3 * x], a0 = tx[3 * x + 1], a1 = tx[3 * x + 2];
    const int sy = ty[3 * y], b0 = ty[3 * y + 1], b1 = ty[3 * y + 2];
    const int sx1 = min(sx + 1, in_w - 1), sy1 = min(sy + 1, in_h - 1);
    const uint8_t* r0 = in + ((size_t)f * in_h + sy) * (size_t)in_w * CH;
    const uint8_t* r1 = in + ((size_t)f * in_h + sy1) * (size_t)in_w * CH;
    uint8_t* o = out + (((size_t)f * out_h + y) * out_w + x) * CH;
#pragma unroll
    for (int c = 0; c < CH; ++c) {
        const int h0 = r0[sx * CH + c] * a0 + r0[sx1 * CH + c] * a1;      // (a1 = 0 wherever sx1 was clamped)
        const int h1 = r1[sx * CH + c] * a0 + r1[sx1 * CH + c] * a1;
        const int v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        o[c] = (uint8_t)min(max(v, 0), 255);
    }
}
}  // namespace

extern "C" int am_resize_linear_u8(const uint8_t* d_in, int batch, int in_h, int in_w, int channels, int out_h, int out_w,
                                   uint8_t* d_out, void* stream) {
    if (!d_in || !d_out || batch <= 0 || in_h <= 0 || in_w <= 0 || out_h <= 0 || out_w <= 0 || (channels != 1 && channels != 3))
        return AM_ERR_ARG;
    int *tx = nullptr, *ty = nullptr;
    int rc = linear_axis(in_w, out_w, &tx);
    if (rc) return rc;
    rc = linear_axis(in_h, out_h, &ty);
    if (rc) return rc;
    dim3 grid(am_div_up(out_w, 128), out_h, batch);
    if (channels == 3) k_resize_linear<3><<<grid, 128, 0, (cudaStream_t)stream>>>(d_in, in_h, in_w, tx, ty, out_h, out_w, d_out);
    else k_resize_linear<1><<<grid, 128, 0, (cudaStream_t)stream>>>(d_in, in_h, in_w, tx, ty, out_h, out_w, d_out);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

// legacy_ops.cu -- the four remaining exports of the reference's accessmath_lib (a drop-in
// .so must export all five).  Bandwidth-bound one-pass byte kernels; every fp64 expression is evaluated with
// explicit round-to-nearest intrinsics (no FMA contraction) in the reference's left-to-right order, so results are
// bit-identical to the gcc x86-64 build of R/accessmath_lib.c.
//
// Replaces (R/ = reference ACCESS2021_release/):
//   speaker_detection_handle_frame   R/accessmath_lib.c:7-111
//   regionCumulativeDistribution     R/accessmath_lib.c:113-173
//   adapthisteq                      R/accessmath_lib.c:175-329   (caller R/AccessMath/preprocessing/tools/adaptive_equalizer.py:273-291)
//   combine_results                  R/accessmath_lib.c:331-354   (caller R/AccessMath/preprocessing/content/binarizer.py:381-402)
#include <cmath>
#include <vector>

#include "am_common.cuh"
#include "../../include/accessmath_b200.h"

static inline cudaStream_t S(void* s) { return (cudaStream_t)s; }

// ------------------------------------------------------------------------------------------------
// C `round()` (halfway away from zero) followed by the x86-64 `(unsigned char)` conversion of a double:
// cvttsd2si to a 32-bit int, low byte kept (out of range / NaN -> 0x80000000 -> 0).
__device__ __forceinline__ unsigned char round_to_u8(double v) {
    double t = trunc(v);
    if (fabs(__dsub_rn(v, t)) >= 0.5) t = __dadd_rn(t, copysign(1.0, v));
    if (!(fabs(t) < 2147483648.0)) return 0;
    return (unsigned char)(((int)t) & 0xFF);
}

// ------------------------------------------------------------------------------------------------
// Contrast-limited, centred cumulative distribution of one rectangular cell (accessmath_lib.c:113-173).
// One block per cell.  cells[c] = (min_x, max_x, min_y, max_y) inclusive; dist[c][256].
__global__ void k_region_cdf(const uint8_t* __restrict__ gray, int W, const int4* __restrict__ cells, double slope_max,
                             double* __restrict__ dist) {
    __shared__ int hist[256];
    __shared__ int cum[256];
    __shared__ double out[256];
    const int4 c = cells[blockIdx.x];
    const int t = threadIdx.x;                           // blockDim.x == 256
    hist[t] = 0;
    __syncthreads();
    const int cw = c.y - c.x + 1, ch = c.w - c.z + 1;
    if (cw > 0 && ch > 0) {
        const int n = cw * ch;
        for (int i = t; i < n; i += 256) {
            const int ry = i / cw, rx = i - ry * cw;
            atomicAdd(&hist[gray[(size_t)(c.z + ry) * W + c.x + rx]], 1);     // :126-136
        }
    }
    __syncthreads();
    // inclusive integer scan (exact, so the order of the additions does not matter) :139-144
    int v = hist[t];
    cum[t] = v;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        int a = (t >= o) ? cum[t - o] : 0;
        __syncthreads();
        cum[t] += a;
        __syncthreads();
    }
    const int count = cum[255];
    out[t] = __ddiv_rn((double)cum[t], (double)count);    // :148-151
    __syncthreads();
    if (slope_max > 0.0) {                                // :154-172, inherently sequential
        if (t == 0) {
            double dh = 0.0;
            for (int i = 0; i < 255; ++i) {
                double diff = __dsub_rn(__dsub_rn(__dsub_rn(out[i + 1], out[i]), dh), slope_max);
                dh = __dadd_rn(dh, (diff < 0.0 ? 0.0 : diff));
                out[i + 1] = __dsub_rn(out[i + 1], dh);
            }
        }
        __syncthreads();
        const double add = __ddiv_rn(__dsub_rn(1.0, __dsub_rn(out[255], out[0])), 2.0);
        dist[(size_t)blockIdx.x * 256 + t] = __dadd_rn(out[t], add);
    } else {
        dist[(size_t)blockIdx.x * 256 + t] = out[t];
    }
}

// Per-pixel interpolation between the cell distributions (accessmath_lib.c:240-318).
// cellx[x] / celly[y] = grid cell of the column / row; xmid / ymid = rounded cell centres (:203, :211).
__global__ void k_adapthisteq_apply(const uint8_t* __restrict__ gray, int W, int H, int gx, int gy,
                                    const int* __restrict__ cellx, const int* __restrict__ celly, const int* __restrict__ xmid,
                                    const int* __restrict__ ymid, const double* __restrict__ dist, uint8_t* __restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    const int cx = cellx[x], cy = celly[y];
    const int tone = gray[(size_t)y * W + x];
    const bool edge_x = (cx == 0 && x <= xmid[cx]) || (cx == gx - 1 && x >= xmid[cx]);
    const bool edge_y = (cy == 0 && y <= ymid[cy]) || (cy == gy - 1 && y >= ymid[cy]);
    double v;
    if (edge_x && edge_y) {
        v = dist[((size_t)gx * cy + cx) * 256 + tone];                                          // :264
    } else if (edge_x) {
        const int y0 = cy - (y <= ymid[cy] ? 1 : 0), y1 = y0 + 1;
        const double wy1 = __ddiv_rn((double)(y - ymid[y0]), (double)(ymid[y1] - ymid[y0]));
        const double d00 = dist[((size_t)gx * y0 + cx) * 256 + tone], d01 = dist[((size_t)gx * y1 + cx) * 256 + tone];
        v = __dadd_rn(__dmul_rn(d00, __dsub_rn(1.0, wy1)), __dmul_rn(d01, wy1));                // :276
    } else if (edge_y) {
        const int x0 = cx - (x <= xmid[cx] ? 1 : 0), x1 = x0 + 1;
        const double wx1 = __ddiv_rn((double)(x - xmid[x0]), (double)(xmid[x1] - xmid[x0]));
        const double d00 = dist[((size_t)gx * cy + x0) * 256 + tone], d10 = dist[((size_t)gx * cy + x1) * 256 + tone];
        v = __dadd_rn(__dmul_rn(d00, __dsub_rn(1.0, wx1)), __dmul_rn(d10, wx1));                // :290
    } else {
        const int x0 = cx - (x <= xmid[cx] ? 1 : 0), x1 = x0 + 1;
        const double wx1 = __ddiv_rn((double)(x - xmid[x0]), (double)(xmid[x1] - xmid[x0]));
        const int y0 = cy - (y <= ymid[cy] ? 1 : 0), y1 = y0 + 1;
        const double wy1 = __ddiv_rn((double)(y - ymid[y0]), (double)(ymid[y1] - ymid[y0]));
        const double d00 = dist[((size_t)gx * y0 + x0) * 256 + tone], d01 = dist[((size_t)gx * y1 + x0) * 256 + tone];
        const double d10 = dist[((size_t)gx * y0 + x1) * 256 + tone], d11 = dist[((size_t)gx * y1 + x1) * 256 + tone];
        const double ax = __dsub_rn(1.0, wx1), ay = __dsub_rn(1.0, wy1);
        double s = __dmul_rn(__dmul_rn(d00, ax), ay);                                           // :306-309, left to right
        s = __dadd_rn(s, __dmul_rn(__dmul_rn(d01, ax), wy1));
        s = __dadd_rn(s, __dmul_rn(__dmul_rn(d10, wx1), ay));
        s = __dadd_rn(s, __dmul_rn(__dmul_rn(d11, wx1), wy1));
        v = s;
    }
    out[(size_t)y * W + x] = round_to_u8(__dmul_rn(v, 255.0));
}

// ------------------------------------------------------------------------------------------------
// combine_results (accessmath_lib.c:339-351): 16 pixels per thread when the three planes are 16-byte aligned
__global__ void k_combine_results(const uint8_t* __restrict__ only_board, const uint8_t* __restrict__ equalized, long long n,
                                  unsigned threshold, uint8_t* __restrict__ out, int vec) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (vec) {
        long long n16 = n >> 4;
        if (i < n16) {
            uint4 b = ((const uint4*)only_board)[i], e = ((const uint4*)equalized)[i], r;
            const uint32_t* bp = &b.x; const uint32_t* ep = &e.x; uint32_t* rp = &r.x;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                uint32_t o = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    unsigned bb = (bp[k] >> (8 * j)) & 0xFF, ee = (ep[k] >> (8 * j)) & 0xFF;
                    unsigned v = (bb > 128) ? 0u : (ee < threshold ? 255u : 0u);
                    o |= v << (8 * j);
                }
                rp[k] = o;
            }
            ((uint4*)out)[i] = r;
        }
        long long tail = (n16 << 4) + i;
        if (i < (n & 15)) out[tail] = (only_board[tail] > 128) ? 0 : (equalized[tail] < threshold ? 255 : 0);
    } else if (i < n) {
        out[i] = (only_board[i] > 128) ? 0 : (equalized[i] < threshold ? 255 : 0);
    }
}

// ------------------------------------------------------------------------------------------------
// speaker_detection_handle_frame (accessmath_lib.c:7-111).  Pass 1: every sampled pixel (row, col multiples of
// jump_cells) tests "some channel changed by more than threshold" and bumps the column / row histograms.  All other
// outputs derive from the two histograms: the sums of col / row are integers < 2^53, hence exact in any order.
// Each block owns 256 sampled columns x SD_ROWS sampled rows: per-thread column counts stay in a register.
#define SD_ROWS 16
__global__ void k_speaker_hist(const uint8_t* __restrict__ frame, const uint8_t* __restrict__ last, int W, int H, int C, int threshold,
                               int jump, int ncols, int nrows, int* __restrict__ hist_x, int* __restrict__ hist_y) {
    const int ci = blockIdx.x * blockDim.x + threadIdx.x;
    const int col = ci * jump;
    int mine = 0;
    for (int r = 0; r < SD_ROWS; ++r) {
        const int ri = blockIdx.y * SD_ROWS + r;
        if (ri >= nrows) break;
        const int row = ri * jump;
        int changed = 0;
        if (ci < ncols) {
            const size_t off = ((size_t)row * W + col) * C;
            for (int k = 0; k < C; ++k) {
                int d = (int)last[off + k] - (int)frame[off + k];
                if ((d < 0 ? -d : d) > threshold) { changed = 1; break; }
            }
        }
        mine += changed;
        int cnt = __syncthreads_count(changed);
        if (threadIdx.x == 0 && cnt) atomicAdd(&hist_y[row], cnt);
    }
    if (mine) atomicAdd(&hist_x[col], mine);
}
// Pass 2 (one warp): totals, bounds, means, then the two variance sums in the reference's sequential order (:92-98)
__global__ void k_speaker_finish(const int* __restrict__ hist_x, const int* __restrict__ hist_y, int W, int H, double* __restrict__ res) {
    const int lane = threadIdx.x;
    long long tot = 0, sx = 0, sy = 0;
    int mnx = W + 1, mxx = -1, mny = H + 1, mxy = -1;
    for (int c = lane; c < W; c += 32) {
        int h = hist_x[c];
        if (h) { tot += h; sx += (long long)h * c; mnx = min(mnx, c); mxx = max(mxx, c); }
    }
    for (int r = lane; r < H; r += 32) {
        int h = hist_y[r];
        if (h) { sy += (long long)h * r; mny = min(mny, r); mxy = max(mxy, r); }
    }
    for (int o = 16; o; o >>= 1) {
        tot += __shfl_xor_sync(0xffffffffu, tot, o); sx += __shfl_xor_sync(0xffffffffu, sx, o); sy += __shfl_xor_sync(0xffffffffu, sy, o);
        mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o)); mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
        mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o)); mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
    }
    if (lane != 0) return;
    res[0] = mnx; res[1] = mxx; res[2] = mny; res[3] = mxy;                  // :78-81
    double ax = 0.0, ay = 0.0, dx = 0.0, dy = 0.0;
    if (tot > 0) {
        ax = __ddiv_rn((double)sx, (double)tot); ay = __ddiv_rn((double)sy, (double)tot);
        for (int c = 0; c < W; ++c) {
            double d = __dsub_rn((double)c, ax);
            dx = __dadd_rn(dx, __dmul_rn(__dmul_rn(d, d), (double)hist_x[c]));
        }
        for (int r = 0; r < H; ++r) {
            double d = __dsub_rn((double)r, ay);
            dy = __dadd_rn(dy, __dmul_rn(__dmul_rn(d, d), (double)hist_y[r]));
        }
        dx = __dsqrt_rn(__ddiv_rn(dx, (double)tot)); dy = __dsqrt_rn(__ddiv_rn(dy, (double)tot));
    }
    res[4] = ax; res[5] = ay; res[6] = dx; res[7] = dy; res[8] = (double)tot;
}

// ================================================================================================
// Device-pointer entry points
// ================================================================================================
extern "C" int am_region_cdf_dev(const uint8_t* d_gray, int width, int height, int min_x, int max_x, int min_y, int max_y,
                                 double slope_max, double* d_out256, void* stream) {
    if (!d_gray || !d_out256 || width <= 0 || height <= 0 || min_x < 0 || min_y < 0 || max_x >= width || max_y >= height) return AM_ERR_ARG;
    int4* d_cell = nullptr;
    AM_CUDA(cudaMallocAsync(&d_cell, sizeof(int4), S(stream)));
    int4 c = make_int4(min_x, max_x, min_y, max_y);
    AM_CUDA(cudaMemcpyAsync(d_cell, &c, sizeof(c), cudaMemcpyHostToDevice, S(stream)));
    k_region_cdf<<<1, 256, 0, S(stream)>>>(d_gray, width, d_cell, slope_max, d_out256);
    AM_CUDA(cudaGetLastError());
    AM_CUDA(cudaFreeAsync(d_cell, S(stream)));
    AM_CUDA(cudaStreamSynchronize(S(stream)));            // `c` is a stack variable
    return AM_OK;
}

extern "C" int am_adapthisteq_dev(const uint8_t* d_gray, int width, int height, double slope, int grid_x, int grid_y,
                                  uint8_t* d_out, void* stream) {
    if (!d_gray || !d_out || width <= 0 || height <= 0 || grid_x <= 0 || grid_y <= 0 || grid_x > width || grid_y > height) return AM_ERR_ARG;
    // cell limits, exactly as accessmath_lib.c:178-220 (host side: O(grid) integers)
    std::vector<int> tab((size_t)width + height + grid_x + grid_y);
    int* cellx = tab.data(); int* celly = cellx + width; int* xmid = celly + height; int* ymid = xmid + grid_x;
    std::vector<int4> cells((size_t)grid_x * grid_y);
    std::vector<int> xmin(grid_x), xmax(grid_x), ymin(grid_y), ymax(grid_y);
    const int sx = width / grid_x, sy = height / grid_y, mx = width % grid_x, my = height % grid_y;
    int start = 0;
    for (int rx = 0; rx < grid_x; ++rx) {
        int end = start + sx + (rx < mx ? 1 : 0) - 1;
        xmin[rx] = start; xmax[rx] = end; xmid[rx] = (int)round((start + end) / 2.0);
        for (int x = start; x <= end; ++x) cellx[x] = rx;
        start = end + 1;
    }
    start = 0;
    for (int ry = 0; ry < grid_y; ++ry) {
        int end = start + sy + (ry < my ? 1 : 0) - 1;
        ymin[ry] = start; ymax[ry] = end; ymid[ry] = (int)round((start + end) / 2.0);
        for (int y = start; y <= end; ++y) celly[y] = ry;
        start = end + 1;
    }
    for (int ry = 0; ry < grid_y; ++ry)
        for (int rx = 0; rx < grid_x; ++rx) cells[(size_t)ry * grid_x + rx] = make_int4(xmin[rx], xmax[rx], ymin[ry], ymax[ry]);
    const size_t n_cells = cells.size();
    int* d_tab = nullptr; int4* d_cells = nullptr; double* d_dist = nullptr;
    AM_CUDA(cudaMallocAsync(&d_tab, tab.size() * 4, S(stream)));
    AM_CUDA(cudaMallocAsync(&d_cells, n_cells * sizeof(int4), S(stream)));
    AM_CUDA(cudaMallocAsync(&d_dist, n_cells * 256 * sizeof(double), S(stream)));
    AM_CUDA(cudaMemcpyAsync(d_tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice, S(stream)));
    AM_CUDA(cudaMemcpyAsync(d_cells, cells.data(), n_cells * sizeof(int4), cudaMemcpyHostToDevice, S(stream)));
    k_region_cdf<<<(unsigned)n_cells, 256, 0, S(stream)>>>(d_gray, width, d_cells, slope, d_dist);
    k_adapthisteq_apply<<<dim3(am_div_up(width, 256), height), 256, 0, S(stream)>>>(
        d_gray, width, height, grid_x, grid_y, d_tab, d_tab + width, d_tab + width + height, d_tab + width + height + grid_x, d_dist, d_out);
    AM_CUDA(cudaGetLastError());
    AM_CUDA(cudaFreeAsync(d_tab, S(stream))); AM_CUDA(cudaFreeAsync(d_cells, S(stream))); AM_CUDA(cudaFreeAsync(d_dist, S(stream)));
    AM_CUDA(cudaStreamSynchronize(S(stream)));            // the host tables above are pageable stack/heap memory
    return AM_OK;
}

extern "C" int am_combine_results_dev(const uint8_t* d_only_board, const uint8_t* d_equalized, int width, int height,
                                      unsigned char threshold, uint8_t* d_out, void* stream) {
    if (!d_only_board || !d_equalized || !d_out || width <= 0 || height <= 0) return AM_ERR_ARG;
    const long long n = (long long)width * height;
    const int vec = ((((uintptr_t)d_only_board | (uintptr_t)d_equalized | (uintptr_t)d_out) & 15) == 0) ? 1 : 0;
    const long long threads = vec ? ((n >> 4) > 16 ? (n >> 4) : 16) : n;
    k_combine_results<<<am_div_up(threads, 256), 256, 0, S(stream)>>>(d_only_board, d_equalized, n, threshold, d_out, vec);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

// d_result[9] = change_boundaries[4], change_avg[2], change_deviation[2], total_changes
extern "C" int am_speaker_detection_dev(const uint8_t* d_frame, const uint8_t* d_last_frame, int width, int height, int channels,
                                        int threshold, int jump_cells, double* d_result, void* stream) {
    if (!d_frame || !d_last_frame || !d_result || width <= 0 || height <= 0 || channels <= 0 || jump_cells <= 0) return AM_ERR_ARG;
    int* d_hist = nullptr;
    AM_CUDA(cudaMallocAsync(&d_hist, ((size_t)width + height) * 4, S(stream)));
    AM_CUDA(cudaMemsetAsync(d_hist, 0, ((size_t)width + height) * 4, S(stream)));
    const int ncols = (width + jump_cells - 1) / jump_cells, nrows = (height + jump_cells - 1) / jump_cells;
    k_speaker_hist<<<dim3(am_div_up(ncols, 256), am_div_up(nrows, SD_ROWS)), 256, 0, S(stream)>>>(
        d_frame, d_last_frame, width, height, channels, threshold, jump_cells, ncols, nrows, d_hist, d_hist + width);
    k_speaker_finish<<<1, 32, 0, S(stream)>>>(d_hist, d_hist + width, width, height, d_result);
    AM_CUDA(cudaGetLastError());
    AM_CUDA(cudaFreeAsync(d_hist, S(stream)));
    return AM_OK;
}

// ================================================================================================
// Legacy host-pointer exports: same names and signatures as the reference, staging H2D / D2H internally
// ================================================================================================
struct DevBuf {                                           // frees on scope exit
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, n ? n : 1); }
};

extern "C" void regionCumulativeDistribution(unsigned char* grayscale, int width, int height, int min_x, int max_x, int min_y,
                                             int max_y, double slope_max, double* output) {
    if (!grayscale || !output) return;
    DevBuf g, o;
    const size_t P = (size_t)width * height;
    if (g.alloc(P) != cudaSuccess || o.alloc(256 * sizeof(double)) != cudaSuccess ||
        cudaMemcpy(g.p, grayscale, P, cudaMemcpyHostToDevice) != cudaSuccess ||
        am_region_cdf_dev((const uint8_t*)g.p, width, height, min_x, max_x, min_y, max_y, slope_max, (double*)o.p, nullptr) != AM_OK ||
        cudaMemcpy(output, o.p, 256 * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) {
        fprintf(stderr, "[accessmath_b200] regionCumulativeDistribution failed: %s\n", cudaGetErrorString(cudaGetLastError()));
    }
}

extern "C" int adapthisteq(unsigned char* grayscale, int width, int height, double slope, int grid_x, int grid_y, unsigned char* output) {
    if (!grayscale || !output) return AM_ERR_ARG;
    DevBuf g, o;
    const size_t P = (size_t)width * height;
    AM_CUDA(g.alloc(P)); AM_CUDA(o.alloc(P));
    AM_CUDA(cudaMemcpy(g.p, grayscale, P, cudaMemcpyHostToDevice));
    int rc = am_adapthisteq_dev((const uint8_t*)g.p, width, height, slope, grid_x, grid_y, (uint8_t*)o.p, nullptr);
    if (rc != AM_OK) return rc;
    AM_CUDA(cudaMemcpy(output, o.p, P, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int combine_results(unsigned char* only_board, unsigned char* equalized, int width, int height, unsigned char threshold,
                               unsigned char* final_content) {
    if (!only_board || !equalized || !final_content) return AM_ERR_ARG;
    DevBuf b, e, o;
    const size_t P = (size_t)width * height;
    AM_CUDA(b.alloc(P)); AM_CUDA(e.alloc(P)); AM_CUDA(o.alloc(P));
    AM_CUDA(cudaMemcpy(b.p, only_board, P, cudaMemcpyHostToDevice));
    AM_CUDA(cudaMemcpy(e.p, equalized, P, cudaMemcpyHostToDevice));
    int rc = am_combine_results_dev((const uint8_t*)b.p, (const uint8_t*)e.p, width, height, threshold, (uint8_t*)o.p, nullptr);
    if (rc != AM_OK) return rc;
    AM_CUDA(cudaMemcpy(final_content, o.p, P, cudaMemcpyDeviceToHost));
    return 0;
}

// Returns total_changes like the reference (:110); -1 on a CUDA failure (the reference cannot fail).
extern "C" int speaker_detection_handle_frame(unsigned char* frame, unsigned char* last_frame, int width, int height, int channels,
                                              int threshold, int jump_cells, double* change_boundaries, double* change_avg,
                                              double* change_deviation) {
    if (!frame || !last_frame || !change_boundaries || !change_avg || !change_deviation) return -1;
    DevBuf f, l, r;
    const size_t n = (size_t)width * height * channels;
    double res[9];
    if (f.alloc(n) != cudaSuccess || l.alloc(n) != cudaSuccess || r.alloc(sizeof(res)) != cudaSuccess ||
        cudaMemcpy(f.p, frame, n, cudaMemcpyHostToDevice) != cudaSuccess ||
        cudaMemcpy(l.p, last_frame, n, cudaMemcpyHostToDevice) != cudaSuccess ||
        am_speaker_detection_dev((const uint8_t*)f.p, (const uint8_t*)l.p, width, height, channels, threshold, jump_cells, (double*)r.p, nullptr) != AM_OK ||
        cudaMemcpy(res, r.p, sizeof(res), cudaMemcpyDeviceToHost) != cudaSuccess) {
        fprintf(stderr, "[accessmath_b200] speaker_detection_handle_frame failed: %s\n", cudaGetErrorString(cudaGetLastError()));
        return -1;
    }
    for (int i = 0; i < 4; ++i) change_boundaries[i] = res[i];
    change_avg[0] = res[4]; change_avg[1] = res[5];
    change_deviation[0] = res[6]; change_deviation[1] = res[7];
    return (int)res[8];
}


// ------------------------------------------------------------------------------------------------------------------------------
// Small downstream reducer: VideoSegmenter.compute_binary_sums
// (R/AccessMath/preprocessing/content/video_segmenter.py:21-28: `binary.sum() / 255` per frame, the input of the stage-04
// regression tree).  Per frame the SUM OF THE PIXEL VALUES as an exact 64-bit integer (the caller divides by 255 in fp64 exactly as
// numpy does): from bit-packed frames 255 * popcount, from uint8 frames the byte sum.  One pass, 128-bit loads, block reduce + atomicAdd.
__global__ void k_frame_sums_bits(const uint32_t* __restrict__ bits, long long words_per_frame, unsigned long long* __restrict__ out) {
    const int f = blockIdx.y;
    const uint4* p = (const uint4*)(bits + (size_t)f * words_per_frame);          // words_per_frame % 4 == 0 (WPR is)
    unsigned long long acc = 0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < words_per_frame / 4; i += (long long)gridDim.x * blockDim.x) {
        const uint4 v = p[i];
        acc += __popc(v.x) + __popc(v.y) + __popc(v.z) + __popc(v.w);
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&out[f], 255ull * acc);
}
__global__ void k_frame_sums_u8(const uint8_t* __restrict__ px, long long bytes_per_frame, unsigned long long* __restrict__ out) {
    const int f = blockIdx.y;
    const uint8_t* p = px + (size_t)f * bytes_per_frame;
    unsigned long long acc = 0;
    const long long head = min(bytes_per_frame, (long long)((16 - ((uintptr_t)p & 15)) & 15));     // bytes in front of the first aligned uint4
    const long long nvec = (bytes_per_frame - head) / 16;
    const uint4* v4 = (const uint4*)(p + head);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (long long)gridDim.x * blockDim.x) {
        const uint4 v = v4[i];
        acc += __vsadu4(v.x, 0) + __vsadu4(v.y, 0) + __vsadu4(v.z, 0) + __vsadu4(v.w, 0);            // sum of the four bytes of a word
    }
    if (blockIdx.x == 0) {
        for (long long i = threadIdx.x; i < head; i += blockDim.x) acc += p[i];
        for (long long i = head + nvec * 16 + threadIdx.x; i < bytes_per_frame; i += blockDim.x) acc += p[i];
    }
    for (int o = 16; o; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0 && acc) atomicAdd(&out[f], acc);
}
extern "C" int am_frame_sums_bits(const uint32_t* d_bits, int batch, int height, int width, unsigned long long* d_sums, void* stream) {
    if (!d_bits || !d_sums || batch <= 0 || height <= 0 || width <= 0) return AM_ERR_ARG;
    const long long wpf = (long long)height * am_words_per_row_impl(width);
    AM_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(unsigned long long) * batch, S(stream)));
    k_frame_sums_bits<<<dim3(32, batch), 256, 0, S(stream)>>>(d_bits, wpf, d_sums);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}
extern "C" int am_frame_sums_u8(const uint8_t* d_frames, int batch, long long bytes_per_frame, unsigned long long* d_sums, void* stream) {
    if (!d_frames || !d_sums || batch <= 0 || bytes_per_frame <= 0) return AM_ERR_ARG;
    AM_CUDA(cudaMemsetAsync(d_sums, 0, sizeof(unsigned long long) * batch, S(stream)));
    k_frame_sums_u8<<<dim3(64, batch), 256, 0, S(stream)>>>(d_frames, bytes_per_frame, d_sums);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

// grouping.cu -- stage 03 (CC grouping) on the unique-CC tables that stage 02 left in HBM (sm_100a).
//
// The pixel work of R/AccessMath/preprocessing/content/cc_stability_estimator.py:166-681 as three HBM/latency-bound kernels
// families over bit-packed crops (word-aligned rows at their ABSOLUTE x position):
//   am_group_overlaps  compute_overlapping_stable_cc (:245-306): all-pairs inclusive bbox test among the stable uniques (the two
//                      IntervalIndex sweeps + set intersection) in ascending (idx1, idx2) order, then AND + popcount per pair
//                      (getOverlapFMeasure, connected_component.py:202-250); recall / precision are formed by the caller in fp64.
//   am_group_images    compute_group_images (:575-636): per (group, time segment) vote image = sum of member crops weighted by the
//                      number of frames each member is seen in the segment; pixel kept iff (double) votes / (double) max >= t.
//   am_paint_frames    rebuilt_binary_frame (:174-179) and the clean channel of frames_from_groups (:638-681): bit-packed images
//                      added into uint8 frames with the reference's uint8 wrap-around (`+= 255`; two overlapping groups give 254).
// The order-dependent list / dictionary logic between those steps (split_stable_cc_by_gaps, compute_groups, conflicts, ages)
// is host logic in lecturemath_b200/cc_grouping.py.
#include "am_common.cuh"
#include "../../include/accessmath_b200.h"

namespace {

constexpr int kTile = 256;

__device__ __forceinline__ bool boxes_touch(int4 a, int4 b) {         // .x = min_x .y = max_x .z = min_y .w = max_y, inclusive
    return a.x <= b.y && b.x <= a.y && a.z <= b.w && b.z <= a.w;
}

// ---- all-pairs bbox test --------------------------------------------------------------------------------------------
// Thread a walks b = a+1 .. n-1 through shared-memory tiles; pass 0 counts, pass 1 writes (a, b) at off[a] + running rank, so the
// pair list comes out sorted by (a, b) without a sort.
template <int kPass>
__global__ void __launch_bounds__(kTile)
k_group_pairs(const am_unique_view v, const int* __restrict__ ids, int n, int* __restrict__ counts, const long long* __restrict__ off,
              int* __restrict__ pairs, long long capacity) {
    __shared__ int4 s_box[kTile];
    const int a = blockIdx.x * kTile + threadIdx.x;
    int4 mine = make_int4(0, -1, 0, -1);
    if (a < n) { const int u = ids[a]; mine = make_int4(v.min_x[u], v.max_x[u], v.min_y[u], v.max_y[u]); }
    long long w = (kPass == 1 && a < n) ? off[a] : 0;
    int cnt = 0;
    for (int t0 = blockIdx.x * kTile; t0 < n; t0 += kTile) {          // tiles at or after this block's own
        __syncthreads();
        const int b = t0 + threadIdx.x;
        if (b < n) { const int u = ids[b]; s_box[threadIdx.x] = make_int4(v.min_x[u], v.max_x[u], v.min_y[u], v.max_y[u]); }
        __syncthreads();
        if (a < n) {
            const int lim = min(kTile, n - t0);
            for (int j = max(0, a + 1 - t0); j < lim; ++j) {
                if (boxes_touch(mine, s_box[j])) {
                    if (kPass == 1) {
                        if (w < capacity) { pairs[3 * w] = a; pairs[3 * w + 1] = t0 + j; pairs[3 * w + 2] = 0; }
                        ++w;
                    } else ++cnt;
                }
            }
        }
    }
    if (kPass == 0 && a < n) counts[a] = cnt;
}

// exclusive scan of n counts into 64-bit offsets, one block (n is the number of stable CCs: a one-shot, tiny pass)
__global__ void __launch_bounds__(1024) k_group_scan(const int* __restrict__ counts, int n, long long* __restrict__ off) {
    __shared__ int s_scan[33];
    __shared__ long long s_base;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const int c = i < n ? counts[i] : 0;
        int total;
        const int ex = block_excl_scan(c, s_scan, &total);
        if (i < n) off[i] = s_base + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_base += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) off[n] = s_base;
}

// one warp per pair: popcount(crop_a & crop_b) over the intersection box; rows of both crops start at a 32-pixel boundary of
// the frame, so words line up without shifting and pixels outside either bbox are zero in that crop
__global__ void __launch_bounds__(256)
k_group_pair_overlap(const am_unique_view v, const int* __restrict__ ids, int* __restrict__ pairs, long long n_pairs) {
    const long long p = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (p >= n_pairs) return;
    const int ua = ids[pairs[3 * p]], ub = ids[pairs[3 * p + 1]];
    const int x0 = max(v.min_x[ua], v.min_x[ub]), x1 = min(v.max_x[ua], v.max_x[ub]);
    const int y0 = max(v.min_y[ua], v.min_y[ub]), y1 = min(v.max_y[ua], v.max_y[ub]);
    const int wa0 = v.min_x[ua] >> 5, wb0 = v.min_x[ub] >> 5;
    const int cwa = (v.max_x[ua] >> 5) - wa0 + 1, cwb = (v.max_x[ub] >> 5) - wb0 + 1;
    const uint32_t* ca = v.arena + v.crop_off[ua] + (size_t)(y0 - v.min_y[ua]) * cwa;
    const uint32_t* cb = v.arena + v.crop_off[ub] + (size_t)(y0 - v.min_y[ub]) * cwb;
    const int w0 = x0 >> 5, nw = (x1 >> 5) - w0 + 1, rows = y1 - y0 + 1;
    int m = 0;
    for (int i = lane; i < nw * rows; i += 32) {
        const int r = i / nw, w = w0 + (i - r * nw);
        m += __popc(ca[(size_t)r * cwa + (w - wa0)] & cb[(size_t)r * cwb + (w - wb0)]);
    }
    m = __reduce_add_sync(0xffffffffu, m);
    if (lane == 0) pairs[3 * p + 2] = m;
}

// ---- group images ---------------------------------------------------------------------------------------------------
// votes of the 32 pixels of word (y, wx) of a segment: sum over members of crop bit * weight
__device__ __forceinline__ void segment_votes(const am_unique_view& v, const int* __restrict__ members, int m0, int m1, int y, int wx,
                                              int (&votes)[32]) {
#pragma unroll
    for (int i = 0; i < 32; ++i) votes[i] = 0;
    for (int m = m0; m < m1; ++m) {
        const int u = members[2 * m], weight = members[2 * m + 1];
        const int my0 = v.min_y[u], mw0 = v.min_x[u] >> 5, mw1 = v.max_x[u] >> 5;
        if (y < my0 || y > v.max_y[u] || wx < mw0 || wx > mw1) continue;
        const uint32_t bits = v.arena[v.crop_off[u] + (size_t)(y - my0) * (mw1 - mw0 + 1) + (wx - mw0)];
        if (!bits) continue;
#pragma unroll
        for (int i = 0; i < 32; ++i) votes[i] += ((bits >> i) & 1u) ? weight : 0;
    }
}

// one CTA per segment: pass 1 = maximum vote of the segment, pass 2 = threshold + bit-pack (votes are recomputed: the crops
// are small and L2-resident, a vote plane in HBM would cost more than it saves)
__global__ void __launch_bounds__(256)
k_group_images(const am_unique_view v, const int* __restrict__ seg, const int* __restrict__ members, double threshold,
               const unsigned long long* __restrict__ out_off, uint32_t* __restrict__ out) {
    __shared__ int s_max;
    const int s = blockIdx.x;
    const int x0 = seg[6 * s], x1 = seg[6 * s + 1], y0 = seg[6 * s + 2], y1 = seg[6 * s + 3], m0 = seg[6 * s + 4], m1 = seg[6 * s + 5];
    const int w0 = x0 >> 5, cw = (x1 >> 5) - w0 + 1, rows = y1 - y0 + 1;
    if (threadIdx.x == 0) s_max = 0;
    __syncthreads();
    int votes[32];
    int vmax = 0;
    for (int i = threadIdx.x; i < cw * rows; i += blockDim.x) {
        const int r = i / cw;
        segment_votes(v, members, m0, m1, y0 + r, w0 + (i - r * cw), votes);
#pragma unroll
        for (int k = 0; k < 32; ++k) vmax = max(vmax, votes[k]);
    }
    vmax = __reduce_max_sync(0xffffffffu, vmax);
    if ((threadIdx.x & 31) == 0) atomicMax(&s_max, vmax);
    __syncthreads();
    const double denom = (double)s_max;
    uint32_t* dst = out + out_off[s];
    for (int i = threadIdx.x; i < cw * rows; i += blockDim.x) {
        const int r = i / cw;
        segment_votes(v, members, m0, m1, y0 + r, w0 + (i - r * cw), votes);
        uint32_t bits = 0;
#pragma unroll
        for (int k = 0; k < 32; ++k) bits |= ((double)votes[k] / denom >= threshold ? 1u : 0u) << k;     // 0/0 = NaN -> false, as numpy
        dst[i] = bits;
    }
}

// ---- painting bit-packed images into uint8 frames --------------------------------------------------------------------
// per-byte wrap-around add (no carry between the four pixels of a word)
__device__ __forceinline__ uint32_t add_bytes(uint32_t a, uint32_t b) {
    return ((a & 0x7f7f7f7fu) + (b & 0x7f7f7f7fu)) ^ ((a ^ b) & 0x80808080u);
}

// one CTA per item (an image placed in a frame); thread = one image word = 32 pixels = up to 9 aligned 4-byte groups of the frame
__global__ void __launch_bounds__(256)
k_paint_items(const int* __restrict__ item_frame, const int* __restrict__ item_img, const int* __restrict__ boxes,
              const unsigned long long* __restrict__ img_off, const uint32_t* __restrict__ imgs, int H, int W, int frame0,
              uint32_t* __restrict__ out_words, long long out_lead) {
    const int k = blockIdx.x, im = item_img[k];
    const int x0 = boxes[4 * im], x1 = boxes[4 * im + 1], y0 = boxes[4 * im + 2], y1 = boxes[4 * im + 3];
    const int w0 = x0 >> 5, cw = (x1 >> 5) - w0 + 1, rows = y1 - y0 + 1;
    const uint32_t* src = imgs + img_off[im];
    const long long fbase = out_lead + (long long)(item_frame[k] - frame0) * H * W;      // byte index of the frame in the aligned view
    for (int i = threadIdx.x; i < cw * rows; i += blockDim.x) {
        const uint32_t bits = src[i];
        if (!bits) continue;
        const int r = i / cw, xw = (w0 + (i - r * cw)) << 5;
        const long long a0 = fbase + (long long)(y0 + r) * W + xw;                       // byte of pixel xw
        for (long long g = a0 & ~3LL; g < a0 + 32; g += 4) {
            uint32_t add = 0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const long long px = g + j - a0;                                          // pixel index inside the word
                if (px >= 0 && px < 32 && ((bits >> px) & 1u)) add |= 0xffu << (8 * j);
            }
            if (!add) continue;
            uint32_t* p = out_words + (g >> 2);
            uint32_t old = *p, assumed;
            do { assumed = old; old = atomicCAS(p, assumed, add_bytes(assumed, add)); } while (old != assumed);
        }
    }
}

}  // namespace

extern "C" int am_group_overlaps(const am_unique_view* v, const int* d_ids, int n, int* d_pairs, long long capacity,
                                 long long* h_n_pairs, void* stream) {
    if (!v || !h_n_pairs || n < 0 || (n > 0 && !d_ids) || capacity < 0 || (capacity > 0 && !d_pairs)) return AM_ERR_ARG;
    *h_n_pairs = 0;
    if (n < 2) return AM_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int* counts = nullptr;
    long long* off = nullptr;
    AM_CUDA(cudaMalloc(&counts, (size_t)n * sizeof(int)));
    if (cudaMalloc(&off, (size_t)(n + 1) * sizeof(long long)) != cudaSuccess) { cudaFree(counts); return AM_ERR_CUDA; }
    const int blocks = am_div_up(n, kTile);
    k_group_pairs<0><<<blocks, kTile, 0, st>>>(*v, d_ids, n, counts, nullptr, nullptr, 0);
    k_group_scan<<<1, 1024, 0, st>>>(counts, n, off);
    long long total = 0;
    cudaError_t e = cudaMemcpyAsync(&total, off + n, sizeof(long long), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    int rc = AM_OK;
    if (e != cudaSuccess) rc = AM_ERR_CUDA;
    *h_n_pairs = total;
    if (rc == AM_OK && total > capacity) rc = AM_ERR_CAPACITY;       // the caller re-allocates d_pairs and calls again
    if (rc == AM_OK && total > 0) {
        k_group_pairs<1><<<blocks, kTile, 0, st>>>(*v, d_ids, n, nullptr, off, d_pairs, capacity);
        k_group_pair_overlap<<<am_div_up(total * 32, 256), 256, 0, st>>>(*v, d_ids, d_pairs, total);
        e = cudaStreamSynchronize(st);
        if (e == cudaSuccess) e = cudaGetLastError();
        if (e != cudaSuccess) { fprintf(stderr, "[accessmath_b200] am_group_overlaps: %s\n", cudaGetErrorString(e)); rc = AM_ERR_CUDA; }
    }
    cudaFree(counts);
    cudaFree(off);
    return rc;
}

extern "C" int am_group_images(const am_unique_view* v, int n_seg, const int* d_seg, const int* d_members, double threshold,
                               const unsigned long long* d_out_off, uint32_t* d_out, void* stream) {
    if (!v || n_seg < 0 || (n_seg > 0 && (!d_seg || !d_members || !d_out_off || !d_out))) return AM_ERR_ARG;
    if (n_seg == 0) return AM_OK;
    k_group_images<<<n_seg, 256, 0, (cudaStream_t)stream>>>(*v, d_seg, d_members, threshold, d_out_off, d_out);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

extern "C" int am_paint_frames(int n_items, const int* d_item_frame, const int* d_item_img, const int* d_boxes,
                               const unsigned long long* d_img_off, const uint32_t* d_imgs, int frame0, int n_frames, int height,
                               int width, uint8_t* d_out, void* stream) {
    if (n_items < 0 || n_frames <= 0 || height <= 0 || width <= 0 || !d_out) return AM_ERR_ARG;
    if (n_items > 0 && (!d_item_frame || !d_item_img || !d_boxes || !d_img_off || !d_imgs)) return AM_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    AM_CUDA(cudaMemsetAsync(d_out, 0, (size_t)n_frames * height * width, st));
    if (n_items == 0) return AM_OK;
    const long long lead = (long long)((uintptr_t)d_out & 3);         // 32-bit atomics need an aligned view of the uint8 frames
    k_paint_items<<<n_items, 256, 0, st>>>(d_item_frame, d_item_img, d_boxes, d_img_off, d_imgs, height, width, frame0,
                                           (uint32_t*)(d_out - lead), lead);
    AM_CUDA(cudaGetLastError());
    return AM_OK;
}

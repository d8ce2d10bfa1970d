/*
 * accessmath_b200.h -- C ABI of libaccessmath_b200.so, the B200 (sm_100a) drop-in for the native part of
 * LectureMath's per-frame content-extraction hot path.
 *
 * Plain C: pointers and sizes only, no torch types.  "d_" arguments are DEVICE pointers, "h_" are HOST
 * pointers, `stream` is a cudaStream_t passed as void* (NULL = default stream).  Every function returns
 * 0 on success, non-zero on failure (1 = CUDA error, 2 = bad argument, 3 = capacity exceeded) and logs to
 * stderr; the reference's callers ignore return values.
 *
 * R/ = reference ACCESS2021_release/.
 */
#ifndef ACCESSMATH_B200_H
#define ACCESSMATH_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ===== 1. Legacy entry point (host pointers, same signature as the reference) ================
 * Replaces CC_AgeBoundaries, R/accessmath_lib.c:357-359 (argtypes R/AccessMath/preprocessing/content/
 * labeler.py:159-165).  Stages H2D/D2H internally.  The age rule `if (age < 0 || a < age) age = a` in raster order
 * (accessmath_lib.c:405-407) is reproduced for every input, negative ages included (it is the plain minimum for ages >= 0,
 * frame times); labels outside 1..count_labels are ignored instead of written out of bounds. */
int CC_AgeBoundaries(int* labels, float* ages, int width, int height, int count_labels,
                     int* out_mins_y, int* out_maxs_y, int* out_mins_x, int* out_maxs_x,
                     int* out_counts, float* output_age);
/* The same operator on DEVICE pointers, asynchronous on `stream` (no PCIe round trip of the 8*P label / age bytes):
 * d_out6[6][count_labels] = min_y, max_y, min_x, max_x, count (int32), age (fp32 bits); d_ages may be NULL (all zero);
 * d_scratch >= 10 * count_labels + 4 ints. */
int am_cc_age_boundaries_dev(const int* d_labels, const float* d_ages, int width, int height, int count_labels, int* d_out6,
                             int* d_scratch, void* stream);

/* ===== 1b. The other four exports of the reference's accessmath_lib (legacy 2013-15 binarizer / speaker detection).
 * Same names and signatures as R/accessmath_lib.c so that ctypes.CDLL callers resolve them unchanged;
 * host pointers, H2D/D2H staged internally, results bit-identical to the gcc x86-64 build of the reference (fp64
 * evaluated in the reference's order without FMA contraction).  am_*_dev = the same operators on device pointers. */
/* R/accessmath_lib.c:175-329; caller R/AccessMath/preprocessing/tools/adaptive_equalizer.py:273-291 */
int adapthisteq(unsigned char* grayscale, int width, int height, double slope, int grid_x, int grid_y,
                unsigned char* output);
/* R/accessmath_lib.c:113-173 (internal helper of adapthisteq, exported by the reference) */
void regionCumulativeDistribution(unsigned char* grayscale, int width, int height, int min_x, int max_x,
                                  int min_y, int max_y, double slope_max, double* output);
/* R/accessmath_lib.c:331-354; caller R/AccessMath/preprocessing/content/binarizer.py:381-402 */
int combine_results(unsigned char* only_board, unsigned char* equalized, int width, int height,
                    unsigned char threshold, unsigned char* final_content);
/* R/accessmath_lib.c:7-111 (no Python caller in this release); returns total_changes, -1 on a CUDA failure */
int speaker_detection_handle_frame(unsigned char* frame, unsigned char* last_frame, int width, int height,
                                   int channels, int threshold, int jump_cells, double* change_boundaries,
                                   double* change_avg, double* change_deviation);
int am_adapthisteq_dev(const uint8_t* d_gray, int width, int height, double slope, int grid_x, int grid_y,
                       uint8_t* d_out, void* stream);
int am_region_cdf_dev(const uint8_t* d_gray, int width, int height, int min_x, int max_x, int min_y, int max_y,
                      double slope_max, double* d_out256, void* stream);
int am_combine_results_dev(const uint8_t* d_only_board, const uint8_t* d_equalized, int width, int height,
                           unsigned char threshold, uint8_t* d_out, void* stream);
/* d_result[9] = change_boundaries[4], change_avg[2], change_deviation[2], total_changes */
int am_speaker_detection_dev(const uint8_t* d_frame, const uint8_t* d_last_frame, int width, int height, int channels,
                             int threshold, int jump_cells, double* d_result, void* stream);

/* ===== 2. Library / device ================================================================== */
int am_version(void);
int am_device_count(void);                 /* 0 when no CUDA device: callers must fail loudly */
int am_words_per_row(int width);           /* uint32 words per bit-packed mask row (padded to 4) */

/* ===== 3. Bit-packed masks ==================================================================
 * bits[f][y][w] bit b <-> pixel x = 32*w + b, ink = 1.  Replaces the uint8 0/255 frames exchanged between
 * stage 01 and 02 (R/AccessMath/preprocessing/video_worker/FCN_lecturenet_binarizer.py:54-64,
 * R/AccessMath/preprocessing/content/helper.py:27-34). */
int am_pack_mask_u8(const uint8_t* d_mask, int width, int height, int batch, uint32_t* d_bits, void* stream);
/* am_pack_mask_u8, and d_flags[f] = 1 when frame f holds a value other than 0 / 255 (no exact 1-bit form) */
int am_pack_mask_u8_exact(const uint8_t* d_mask, int width, int height, int batch, uint32_t* d_bits, int* d_flags, void* stream);
int am_unpack_mask_u8(const uint32_t* d_bits, int width, int height, int batch, uint8_t* d_mask, void* stream);

/* ===== 4. CC labeling + statistics + crops ==================================================
 * Replaces scipy.ndimage.label (labeler.py:126), CC_AgeBoundaries (accessmath_lib.c:357-413) and the
 * crop loop of Labeler.extractSpatioTemporalContent (labeler.py:171-189) for a batch of frames. */
typedef struct am_cc_ctx am_cc_ctx;
am_cc_ctx* am_cc_create(int width, int height, int max_batch, int max_runs, int max_labels, int max_kept,
                        int crop_words, int min_pixels);
void am_cc_destroy(am_cc_ctx* ctx);
/* labels (optional, may be NULL): int32 [batch][H][W], 0 = background, 1..n in raster order of first pixel */
int am_cc_label_batch(am_cc_ctx* ctx, const uint32_t* d_bits, int batch, int32_t* d_labels, void* stream);
/* h_counts[batch][4] = n_runs, n_labels, n_kept (count >= min_pixels), crop_words.  Synchronises `stream`. */
int am_cc_counts(am_cc_ctx* ctx, int batch, int* h_counts, void* stream);
/* per-label table of one frame (index = label-1), the six CC_AgeBoundaries outputs (age = 0 / -1) */
int am_cc_read_label_table(am_cc_ctx* ctx, int frame, int n_labels, int* h_min_y, int* h_max_y, int* h_min_x,
                           int* h_max_x, int* h_count, void* stream);
/* kept CCs of one frame, ascending label: h_rows[n_kept][8] =
 *   unique_idx (after am_est_add_frames, else -1), raw_label, min_x, max_x, min_y, max_y, size, crop_offset */
int am_cc_read_kept(am_cc_ctx* ctx, int frame, int n_kept, int* h_rows, void* stream);
/* bit-packed crops of one frame (see crop layout in DESIGN.md), crop_words uint32 */
int am_cc_read_crops(am_cc_ctx* ctx, int frame, int crop_words, uint32_t* h_crops, void* stream);
/* all kept rows of a batch, compacted: d_rows[sum n_kept][8] (device), d_row_offsets[batch+1] */
int am_cc_pack_rows(am_cc_ctx* ctx, int batch, int* d_rows, int row_capacity, int* d_row_offsets, void* stream);

/* ===== 5. Temporal matching ==================================================================
 * Replaces CCStabilityEstimator.__init__/add_frame (R/AccessMath/preprocessing/content/
 * cc_stability_estimator.py:11-31, 41-155), IntervalIndex.find_matches (R/AccessMath/preprocessing/tools/
 * interval_index.py:42-99) and ConnectedComponent.getOverlapFMeasure (R/AM_CommonTools/data/
 * connected_component.py:202-250). */
typedef struct am_estimator am_estimator;
am_estimator* am_est_create(int width, int height, double min_recall, double min_precision, int max_gap,
                            int max_uniques, int max_active, long long arena_words);
void am_est_destroy(am_estimator* est);
/* match frames [first, first+n) of ctx, in order, against the active unique CCs; results land in ctx
 * (am_cc_read_kept / am_cc_pack_rows column 0) */
int am_est_add_frames(am_estimator* est, am_cc_ctx* ctx, int first, int n, void* stream);
/* h_state[6] = n_unique, n_active, img_idx, status, tempo_count(lo), tempo_count(hi).  Synchronises. */
int am_est_state(am_estimator* est, int* h_state, void* stream);
/* unique CCs [first, first+n): h_rows[n][8] = first_frame, first_label, min_x, max_x, min_y, max_y, size, last_seen */
int am_est_read_uniques(am_estimator* est, int first, int n, int* h_rows, void* stream);
/* first-seen crop of one unique CC (bit-packed, word aligned) */
int am_est_read_unique_crop(am_estimator* est, int unique_idx, int words, uint32_t* h_crop, void* stream);
/* multi-GPU hand-off of the ACTIVE set between frame shards:
 *   export: h_sizes[2] = n_active, crop_words (synchronises); then fill d_meta[n_active][10] =
 *           unique_idx, min_x, max_x, min_y, max_y, size, last_seen, first_frame, first_label, crop_words
 *           and d_crops[crop_words].
 *   import: install that set (global unique indices kept) into a fresh estimator. */
int am_est_export_sizes(am_estimator* est, long long* h_sizes, void* stream);
int am_est_export(am_estimator* est, int* d_meta, uint32_t* d_crops, void* stream);
int am_est_import(am_estimator* est, int n_active, int n_unique, int img_idx, unsigned long long tempo_count,
                  const int* d_meta, const uint32_t* d_crops, long long crop_words, void* stream);
/* the same hand-off without any host synchronisation (sizes stay on the device): ONE fixed-capacity int32 buffer
 *   d_buf[0..16) = n_active, crop_words, n_unique, img_idx, tempo_count lo, hi, flags (non-zero = did not fit), 0...
 *   d_buf[16..)  = meta[n_active][10] as above, then the crops.  A failed hand-off surfaces in am_est_state(). */
int am_est_export_dev(am_estimator* est, int* d_buf, long long capacity_words, void* stream);
int am_est_import_dev(am_estimator* est, const int* d_buf, void* stream);

/* ===== 5b. Frame-shard hand-off over NVLink peer memory (one process per GPU) ===
 * Every rank owns a mailbox (cudaMalloc'ed: receive buffer + flags) that its ring neighbours map through CUDA IPC.  The
 * sender's am_est_export_dev stores straight into the successor's mailbox; stream memory operations publish / await the
 * chunk counter, so no SM spins while a rank waits. */
#define AM_P2P_HANDLE_BYTES 64
void* am_p2p_alloc(long long bytes);                         /* zero-filled device memory that can be exported */
int am_p2p_free(void* d_ptr);
int am_p2p_export_handle(void* d_ptr, void* h_handle);       /* h_handle[AM_P2P_HANDLE_BYTES] for the neighbours */
void* am_p2p_open_handle(const void* h_handle);              /* neighbour's mailbox mapped into this process */
int am_p2p_close_handle(void* d_peer_ptr);
int am_stream_write32(void* d_flag, unsigned value, void* stream);      /* stream-ordered store, local or peer address */
int am_stream_wait_geq32(void* d_flag, unsigned value, void* stream);   /* stream waits until the flag reached value */

/* ===== 6. FCN-LectureNet binarizer (tcgen05 implicit GEMM) ===================================
 * Replaces the PyTorch arithmetic of FCN_LectureNet.forward / binarize
 * (R/AccessMath/lecturenet_v1/FCN_lecturenet.py:260-323, 364-403, 430-467) and the frame pre/post
 * processing of FCN_LectureNet_Binarizer.handleFrame (R/AccessMath/preprocessing/video_worker/
 * FCN_lecturenet_binarizer.py:47-54).  Activations are NHWC bf16 with physical zero padding in x:
 * buf[n][y][xp][c], xp in [0, W + 2*pad).  See DESIGN.md for the GEMM formulation. */
typedef struct am_conv_seg {       /* one K segment = one input tensor of a (concatenated) convolution */
    const void* ptr;               /* bf16 activation buffer */
    int C;                         /* channels per pixel in the buffer (multiple of 8) */
    int Wp;                        /* padded row length in pixels */
    int Hbuf;                      /* rows per frame in the buffer */
    int x_off;                     /* buffer pad - conv pad (pixels) */
    int rowrun;                    /* 1: row-run mode (overlapping rows, S-packing); 0: one load per horizontal tap */
    int S;                         /* output pixels packed per GEMM row (row-run mode) */
    int run_len;                   /* (KW + S - 1) * C   (row-run mode) */
    int KW;                        /* horizontal taps */
} am_conv_seg;

typedef struct am_conv_desc {
    int nseg;
    am_conv_seg seg[2];
    const void* weights;           /* packed bf16 [chunks*KH*Ntot_pad][64] (fcn_lecturenet.pack_weights) */
    const float* bias;             /* fp32 [Ntot_pad], BatchNorm folded */
    int KH, padY;
    int RT, YT;                    /* tile: RT groups x YT rows = 128 GEMM rows, RT % 8 == 0 */
    int nR, Hin, batch;            /* GEMM rows per image row, GEMM row units per frame column (image rows / in_ystep), frames */
    int NT, Ntot, Ntot_pad;        /* UMMA N per CTA, valid N, padded N (multiple of NT) */
    void* out; int out_f32;        /* destination (bf16 or fp32), NHWC */
    int out_H, out_W;
    long long out_sn, out_sy;      /* element strides: frame, row */
    int out_sx, out_padx, out_coff;/* pixel stride (channels), left pad (pixels), channel offset */
    int Cout, Sy, Sx;              /* GEMM column n = ((sy*Sx)+sx)*Cout + co  ->  pixel (Sy*y+sy, Sx*r+sx), channel co */
    int act;                       /* 0 = none, 1 = nn.GELU() (erf form) evaluated as 0.5 x (1 + tanh z(x)), z = a three-term minimax fit of
                                      atanh(erf(x / sqrt 2)), one tanh.approx per element: |error| <= 2.6e-5 + 2.5e-4 |x| (csrc/fcn_conv.cu) */
    int flags;                     /* AM_CONV_* tuning overrides (0 = let the library choose) */
    int in_ystep;                  /* input rows per GEMM row step: 1, or 2 = "2-D packing": a GEMM row produces Sy = 2 output rows,
                                      KH is then the Toeplitz-extended tap count KH_conv + 1 and RT must be 8 (0 means 1) */
    /* optional fused MaxPool2d(2) (floor) of the activated output (FCN_lecturenet.py:264-276): the epilogue also writes the pooled
     * NHWC bf16 tensor, saving the separate pass over the un-pooled one.  NULL = off.  Needs Sx = Sy = 1, RT <= 16, bf16 output,
     * Cout % 16 == 0 (the 2x2 block of a pixel then lives in four lanes of one epilogue warp). */
    void* pool_out;
    int pool_H, pool_W;            /* out_H / 2, out_W / 2 */
    long long pool_sn, pool_sy;    /* element strides: frame, row */
    int pool_sx, pool_padx;        /* pixel stride (channels), left pad (pixels) */
    /* optional fused epilogues of the two fp32 layers (`out` may then be NULL: the fp32 intermediate never reaches HBM):
     * AM_EPI_HEADS (Cout = 4: text logit, reconstruction pre-tanh x3):  diff = (x0 - tanh(rec)) * sigmoid(text)
     *   (FCN_lecturenet.py:370-377) as bf16 into diff_out[B][H][W+2*diff_pad][diff_C] (diff_C = 4 or 8, channels >= 3 zero), x0 read
     *   from the uint8 BGR `frames` [B][H][W][3]; optional text_out fp32 [B][H][W], rec_out fp32 [B][H][W][3] (RGB, tanh applied).
     * AM_EPI_THRESHOLD (Cout = 1, Sx % 16 == 0): ink <=> (uint8)(sigmoid(z) * 255) < threshold (FCN_lecturenet.py:452-467 and
     *   `255 - binary`, FCN_lecturenet_binarizer.py:54) bit-packed into bits_out[B][H][bits_wpr] (bit b of word w = pixel 32w+b),
     *   16 pixels per thread and store; `out` (fp32 logits [B][H][W]) is written as well when it is not NULL. */
    int epi_mode;
    const uint8_t* frames;
    void* diff_out; int diff_C, diff_pad;
    float* text_out; float* rec_out;
    uint32_t* bits_out; int bits_wpr, threshold;
} am_conv_desc;
#define AM_EPI_PLAIN 0
#define AM_EPI_HEADS 1
#define AM_EPI_THRESHOLD 2
#define AM_CONV_NO_RESIDENT 1      /* always stream the weights through the B ring */
#define AM_CONV_NO_MT2 2           /* one M-tile per work item even when two would share the weight tiles */
#define AM_CONV_FORCE_MT2 4        /* two M-tiles per work item (two MMA issuer warps) whenever TMEM and smem allow */
#define AM_CONV_CTA_PAIR 16        /* cta_group::2: clusters of two CTAs run M = 256 MMAs, each CTA holds half of every weight tile */
#define AM_CONV_EPI8 32            /* 8 epilogue warps take part (tensor-bound layers); default: chosen from K */
#define AM_CONV_EPI16 64           /* all 16 epilogue warps (epilogue-bound layers: K <= 640) */
#define AM_CONV_NO_EDGE_HALF 128    /* CTA pairs with 2-D packing: issue the two edge taps in y as full-N MMAs (tuning / debugging) */
#define AM_CONV_FORCE_MT4 8        /* four M-tiles per work item (four MMA issuer warps, 12 epilogue warps): narrow N <= 128 layers */

/* one launch: encodes the tensor maps, picks the tiling and runs the persistent tcgen05 kernel */
int am_conv_gemm(const am_conv_desc* desc, void* stream);
/* prepared form: tensor maps / tiling computed once per layer (needs a CUDA device), then launched many times */
typedef struct am_conv_plan am_conv_plan;
am_conv_plan* am_conv_plan_create(const am_conv_desc* desc);
void am_conv_plan_destroy(am_conv_plan* plan);
int am_conv_plan_launch(const am_conv_plan* plan, void* stream);
/* re-bind the per-call pointers of a prepared launch (they are kernel arguments, not baked into the tensor maps): the frame batch
 * AM_EPI_HEADS reads, and the optional fp32 outputs (NULL = do not write).  which = AM_BIND_*; takes effect at the next launch. */
#define AM_BIND_FRAMES 0
#define AM_BIND_TEXT_OUT 1
#define AM_BIND_REC_OUT 2
#define AM_BIND_OUT 3
#define AM_BIND_THRESHOLD 4        /* ptr carries the integer threshold */
int am_conv_plan_bind(am_conv_plan* plan, int which, void* ptr);
/* info[8] = M-tiles per work item, resident weights (0/1), accumulator stages, A stages, B stages, grid, smem bytes, work items */
int am_conv_plan_info(const am_conv_plan* plan, int* info);

/* uint8 BGR frames [B][H][W][3] -> normalised bf16 RGB (x/255-0.5)/0.5 in buf[B][H][W+2*pad][C] (channels >= 3 zero).
 * Replaces cv2.cvtColor + TF.to_tensor + TF.normalize (FCN_lecturenet_binarizer.py:50, FCN_lecturenet.py:607-618). */
int am_fcn_prep_input(const uint8_t* d_bgr, int batch, int height, int width, void* d_out, int C, int pad, void* stream);
/* MaxPool2d(2) (floor) on padded NHWC bf16 (FCN_lecturenet.py:264-276) */
int am_fcn_maxpool2(const void* d_in, int batch, int height, int width, int C, int pad_in, void* d_out, int pad_out, void* stream);
/* fill the rows/columns a k=2,s=2 transposed conv with output_padding leaves untouched (they equal act(bias)):
 * pixels with y >= y_from or x >= x_from of buf[B][H][W+2*pad][C] get values[c] (bf16[C]) */
int am_fcn_fill_border(void* d_buf, int batch, int height, int width, int C, int pad, int y_from, int x_from,
                       const void* d_values, void* stream);
/* heads: d_heads fp32 [B][H][W][4] = (text logit, rec pre-tanh x3) + the uint8 BGR frame ->
 *   diff = (x0 - tanh(rec)) * sigmoid(text)  (FCN_lecturenet.py:370-377) into bf16 buf[B][H][W+2*pad][C];
 *   optional d_text_logit fp32 [B][H][W], d_rec fp32 [B][H][W][3] (RGB, tanh applied) */
int am_fcn_heads_post(const float* d_heads, const uint8_t* d_bgr, int batch, int height, int width, void* d_diff, int C,
                      int pad, float* d_text_logit, float* d_rec, void* stream);
/* final logits fp32 [B][H][W] -> bit-packed INK mask: ink <=> (uint8)(sigmoid(z)*255) < threshold
 * (FCN_lecturenet.py:452-467 followed by `255 - binary`, FCN_lecturenet_binarizer.py:54) */
int am_fcn_threshold_pack(const float* d_logits, int batch, int height, int width, int threshold, uint32_t* d_bits, void* stream);

/* ----- frames above 2.5 MP (4K video): FCN_LectureNet.binarize halves them until they fit and resizes the masks back -----
 * number of halvings and the FCN working size of a width x height frame: `while w*h > 2500000: w, h = int(w/2), int(h/2)`
 * (FCN_lecturenet.py:434-437).  Host-only helper, no device needed. */
int am_fcn_working_size(int width, int height, int* out_width, int* out_height);
/* PIL.Image.resize((out_w, out_h), PIL.Image.LANCZOS) on uint8 interleaved images [B][H][W][channels] (channels <= 4), bit-identical
 * to Pillow's fixed-point two-pass resampler (FCN_lecturenet.py:436; algorithm: Pillow libImaging/Resample.c) */
int am_lanczos_resize_u8(const uint8_t* d_in, int batch, int in_h, int in_w, int channels, int out_h, int out_w,
                         uint8_t* d_out, void* stream);
/* cv2.resize(frame, (out_w, out_h)) -- INTER_LINEAR, VideoProcessor's forced-resolution resize
 * (R/AccessMath/preprocessing/video_processor/video_processor.py:164-165) -- on uint8 interleaved frames [B][H][W][channels]
 * (channels 1 or 3): OpenCV's 8-bit fixed-point algorithm (11-bit coefficients, int32 rows), bit-identical to its generic code path */
int am_resize_linear_u8(const uint8_t* d_in, int batch, int in_h, int in_w, int channels, int out_h, int out_w,
                        uint8_t* d_out, void* stream);
/* cv2.resize(mask, (out_w, out_h), interpolation=cv2.INTER_NEAREST) on bit-packed masks (FCN_lecturenet.py:481-486; the ink
 * inversion `255 - binary` commutes with it) */
int am_bits_resize_nearest(const uint32_t* d_bits, int batch, int in_h, int in_w, int out_h, int out_w,
                           uint32_t* d_out, void* stream);

/* ===== 7. Stage 03: CC grouping on the device-resident unique-CC tables =====================
 * Pixel work of R/AccessMath/preprocessing/content/cc_stability_estimator.py:166-681 (called by
 * R/pre_ST3D_v3.0_03_cc_grouping.py:41-101).  The order-dependent list / dictionary logic between these steps
 * (split_stable_cc_by_gaps, compute_groups, compute_conflicting_groups ...) is host code in lecturemath_b200/cc_grouping.py. */
typedef struct am_unique_view {            /* DEVICE pointers, index = unique CC */
    const int *min_x, *max_x, *min_y, *max_y, *size;
    const unsigned long long* crop_off;    /* word offset of the unique's first-seen crop in `arena` */
    const uint32_t* arena;                 /* bit-packed crops: word-aligned rows at their absolute x */
    int n;
} am_unique_view;
/* the tables a live estimator already holds in HBM (zero copy; valid until am_est_destroy).  Synchronises to read n. */
int am_est_unique_view(am_estimator* est, am_unique_view* out, void* stream);
/* compute_overlapping_stable_cc (:245-306; IntervalIndex.find_matches interval_index.py:42-99 + getOverlapFMeasure
 * connected_component.py:202-250): all position pairs a < b of the n listed uniques (d_ids[a] = index into the view, listed in
 * ascending stage-03 index) whose inclusive bounding boxes intersect, in ascending (a, b) order: d_pairs[k] = (a, b, matched
 * pixels).  *h_n_pairs = number of pairs; returns 3 (capacity) without writing when it exceeds `capacity` -- call again with a
 * larger buffer.  Synchronises. */
int am_group_overlaps(const am_unique_view* v, const int* d_ids, int n, int* d_pairs, long long capacity,
                      long long* h_n_pairs, void* stream);
/* compute_group_images (:575-636) for n_seg (group, time segment) items: d_seg[s] = (min_x, max_x, min_y, max_y of the group,
 * member_begin, member_end), d_members[m] = (index into the view, frames the CC is seen in inside the segment, > 0);
 * image = ((double) votes / (double) max votes >= threshold), bit-packed like a crop at word offset d_out_off[s] of d_out. */
int am_group_images(const am_unique_view* v, int n_seg, const int* d_seg, const int* d_members, double threshold,
                    const unsigned long long* d_out_off, uint32_t* d_out, void* stream);
/* rebuilt_binary_frame (:174-179) and the clean channel of frames_from_groups (:638-681): zero d_out (uint8
 * [n_frames][height][width], frames frame0 .. frame0+n_frames-1), then for every item k add 255 (uint8 wrap-around, as numpy `+=`)
 * at the set pixels of bit-packed image d_item_img[k] (box d_boxes[img] = min_x, max_x, min_y, max_y; words at d_img_off[img] of
 * d_imgs) into frame d_item_frame[k].  d_out must be accessible up to the next 4-byte boundary past its end. */
int am_paint_frames(int n_items, const int* d_item_frame, const int* d_item_img, const int* d_boxes,
                    const unsigned long long* d_img_off, const uint32_t* d_imgs, int frame0, int n_frames, int height, int width,
                    uint8_t* d_out, void* stream);

/* ===== 7b. Small downstream reducer: VideoSegmenter.compute_binary_sums ========================
 * R/AccessMath/preprocessing/content/video_segmenter.py:21-28 (`binary.sum() / 255` per frame; caller
 * R/pre_ST3D_v3.0_04_vid_segmentation.py:39).  d_sums[f] = exact integer sum of the pixel VALUES of frame f: 255 * popcount for
 * bit-packed frames [batch][height][am_words_per_row(width)], the byte sum for uint8 frames [batch][bytes_per_frame]. */
int am_frame_sums_bits(const uint32_t* d_bits, int batch, int height, int width, unsigned long long* d_sums, void* stream);
int am_frame_sums_u8(const uint8_t* d_frames, int batch, long long bytes_per_frame, unsigned long long* d_sums, void* stream);

/* ===== 8. Wire format 01 -> 02 written on the device ========================================
 * Replaces cv2.imencode(".png", binary) (R/AccessMath/preprocessing/video_worker/FCN_lecturenet_binarizer.py:56); the reader stays
 * cv2.imdecode(raw, IMREAD_GRAYSCALE) (R/AccessMath/preprocessing/content/helper.py:31).  The file is a 1-bit grayscale PNG
 * (ink = white), filter 0, zlib stored blocks: it decodes to the same 0 / 255 pixels as the reference's file. */
long long am_png1_size(int width, int height);          /* bytes per frame (fixed for a frame size); host-only helper */
/* d_bits [batch][height][am_words_per_row(width)] -> d_out [batch][am_png1_size(width, height)] */
int am_png1_encode(const uint32_t* d_bits, int batch, int height, int width, uint8_t* d_out, void* stream);
/* the same file with a COMPRESSED zlib stream (RFC 1951 fixed Huffman codes + run-length matches, written in parallel: one deflate
 * block per 8 KB of scanline bytes; whiteboard masks shrink ~20-60x against the stored form).  d_out [batch][am_png1_capacity()]
 * (16-byte aligned), file sizes in d_sizes[batch]; frame f starts at d_out + f * capacity. */
long long am_png1_capacity(int width, int height);      /* host-only helper */
int am_png1_encode_deflate(const uint32_t* d_bits, int batch, int height, int width, uint8_t* d_out, long long* d_sizes, void* stream);
/* the same writer for 8-bit grayscale frames d_frames [batch][height][width] (the 03 -> 04 clean frames of
 * R/AccessMath/preprocessing/content/cc_stability_estimator.py:677-678 when overlapping groups wrapped a pixel to 254) */
long long am_png8_capacity(int width, int height);
int am_png8_encode_deflate(const uint8_t* d_frames, int batch, int height, int width, uint8_t* d_out, long long* d_sizes, void* stream);
/* decode half on the device: the scanline bytes a 1-bit filter-0 PNG inflates to (d_scan [batch][height][1 + ceil(width / 8)], what
 * Helper.decompress_binary_images keeps of such a file, R/AccessMath/preprocessing/content/helper.py:27-34) -> mask words */
int am_png1_scanlines_to_bits(const uint8_t* d_scan, int batch, int height, int width, uint32_t* d_bits, void* stream);

#ifdef __cplusplus
}
#endif
#endif

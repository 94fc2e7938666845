/* libcald_b200.so -- C ABI of the B200-native CALD unlabeled-pool scoring engine.
 *
 * The reference (we1pingyu/CALD) is pure Python and has no FFI; the boundary this
 * library replaces is the body of
 *     get_uncertainty(task_model, unlabeled_loader, augs, num_cls)      cald_train.py:91-231
 * i.e. for every image: 1 reference forward of the detector (detection/frcnn_la.py:237-275),
 * A augmented forwards (cald/cald_helper.py:23-243), and the paired-prediction reduction
 * (cald_train.py:189-228).  The host-side mirror of the reference signature lives in
 * cald_b200/api.py and binds these entry points with ctypes.
 *
 * Conventions: every call returns 0 on success and <0 on error with the
 * message available from cald_last_error(); no exception crosses the boundary; all buffers
 * are caller-owned HOST memory unless stated, and the engine copies what it keeps; one
 * handle per GPU; a handle is not thread-safe, independent handles are; calls are
 * synchronous at return.
 */
#ifndef CALD_B200_H
#define CALD_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct cald_engine cald_engine;

enum { CALD_ARCH_FRCNN = 0, CALD_ARCH_RETINANET = 1 };
/* F16X3: split-half operands (22-bit significand), three tensor-core passes per product -- fp32-faithful, the product's
 * only validated mode.  F16: one pass over the hi planes (11 bits) -- a speed-of-light reference point, not parity. */
enum { CALD_PREC_F16X3 = 0, CALD_PREC_F16 = 1 };
enum { CALD_CONV_TCGEN05 = 0, CALD_CONV_SIMT = 1 };
/* augmentation kinds (cald/cald_helper.py); `param` carries the reference's per-call argument */
enum {
  CALD_AUG_FLIP = 0,        /* HorizontalFlip                      cald_helper.py:23-30   */
  CALD_AUG_CUTOUT = 1,      /* cutout(cut_num = param)             cald_helper.py:88-132  */
  CALD_AUG_RESIZE = 2,      /* resize(ratio = param), PIL BILINEAR cald_helper.py:47-53   */
  CALD_AUG_ROTATION = 3,    /* rotate(angle = param degrees)       cald_helper.py:135-223 */
  CALD_AUG_GAUSS = 4,       /* GaussianNoise(std = param)          cald_helper.py:72-75   */
  CALD_AUG_SALTPEPPER = 5,  /* SaltPepperNoise(prob = param)       cald_helper.py:78-85   */
  CALD_AUG_COLOR_ADJUST = 6,/* ColorAdjust(factor = param): PIL brightness -> contrast -> saturation, cald_helper.py:65-69 */
  CALD_AUG_COLOR_SWAP = 7   /* ColorSwap: channel permutation perms[i] cald_helper.py:56-62; i from swap_perms, else param */
};
#define CALD_AUG_SMALLER_RESIZE CALD_AUG_RESIZE
typedef struct {
  int kind;
  double param;
} cald_aug;

typedef struct {
  int arch;                  /* CALD_ARCH_*: FRCNN_Feature (frcnn_la.py:146) | RetinaNet (retinanet_cal.py:243, 584-625) */
  int depth;                 /* 50 | 101 (resnet_fpn_backbone) */
  int num_classes;           /* incl. background: 21 VOC / 91 COCO (detection/train.py:43-46) */
  int min_size, max_size;    /* GeneralizedRCNNTransform: 600/1000 VOC, 800/1333 COCO (cald_train.py:340-347) */
  int rpn_pre_nms_top_n;     /* 1000 (frcnn_la.py:154) */
  int rpn_post_nms_top_n;    /* 1000 */
  float rpn_nms_thresh;      /* 0.7 */
  float box_score_thresh;    /* 0.05 (frcnn_la.py:161) */
  float box_nms_thresh;      /* 0.5 */
  int box_detections_per_img;/* FRCNN: 100 per image (frcnn_la.py:161); RetinaNet: 300 PER CLASS (retinanet_cal.py:333,463) */
  int retina_max_detections; /* RetinaNet: capacity of one image's concatenated detection list; default 300 * K =
                                the most the reference can emit (300 per class, retinanet_cal.py:333, 463), so the
                                default cannot overflow.  The number of candidates above the score threshold is
                                unlimited, like in the reference (classes with more than 4096 are scanned in windows) */
  int device;                /* CUDA ordinal */
  int precision;             /* CALD_PREC_* */
  int conv_impl;             /* CALD_CONV_* */
  int max_views_per_pass;    /* views batched through one forward pass (0 = auto: as many as the arena holds at the
                                largest padded input, up to 32) */
  size_t workspace_bytes;    /* device arena (0 = auto from free memory) */
  int debug;                 /* 1: keep host copies of stage tensors for cald_debug_fetch */
} cald_config;

/* Fill cfg with the reference defaults for (arch, depth, num_classes, min_size, max_size). */
int cald_config_default(cald_config* cfg, int arch, int depth, int num_classes, int min_size, int max_size);

int cald_create(const cald_config* cfg, cald_engine** out);
void cald_destroy(cald_engine* e);
const char* cald_last_error(const cald_engine* e); /* e may be NULL: last create() error */

/* Weights: the torchvision-keyed state_dict the reference saves / loads
 * (cald_train.py:351-356, 420-426), fp32 host tensors.  Both the torchvision 0.8.2 and the
 * current key spellings are accepted.  FrozenBatchNorm is folded into the preceding conv
 * (tv:ops/misc.py:54-63, eps 1e-5).  Must be called once with ALL tensors before scoring. */
int cald_load_weights(cald_engine* e, int n, const char* const* names, const float* const* data,
                      const int* ndim, const int64_t* shapes /* [n][4], unused dims = 1 */);

/* get_uncertainty over n images (cald_train.py:91-231).
 * images[i]: u8 RGB, HWC contiguous, heights[i] x widths[i]  (what F.to_tensor(PIL) reads, cald_train.py:107).
 * augs: the augmented views in the order cald_train.py:123-183 appends them; bp: args.bp (cald_train.py:220, 1.3).
 * noise: for GAUSS / SALTPEPPER views, host fp32 planes [3][H][W] drawn by the caller from torch's CPU generator
 *   (torch.randn / torch.rand of image.size(), cald_helper.py:74,80), one pointer per (image, noise view) in
 *   image-major order; NULL when no such view is requested.
 * rng_uniforms: raw random.random() doubles from python's generator (4 per cutout try, drawn in stream order);
 *   uniforms_consumed returns how many the scoring consumed so the caller can advance its generator exactly as
 *   cald_helper.cutout would have (cald_helper.py:106-114).
 * swap_perms: for COLOR_SWAP views, the index random.randint(0, 5) the caller drew for each (image, swap view), image-major;
 *   NULL = use the view's param for every image.
 * out_consistency[n]: np.mean(consistency_aug) per image; out_cls[n][num_classes-1]: mean class-max vector. */
int cald_score(cald_engine* e, int n_images, const uint8_t* const* images, const int* heights, const int* widths,
               int n_augs, const cald_aug* augs, double bp, const double* rng_uniforms, int n_uniforms,
               int* uniforms_consumed, const float* const* noise, const int* swap_perms, double* out_consistency,
               double* out_cls);

/* Pool ingest: the reference reads its pools with PIL on DataLoader workers --
 * Image.open(path).convert('RGB'), detection/voc_utils.py:52-58, detection/coco_utils.py:209-220.  These entry points
 * take the FILES instead: the host walks the marker segments and (on a few threads, while the previous chunk's forward
 * passes run) the serial entropy-coded segment; the coefficients travel to the device, which does everything parallel
 * (dequantisation, JDCT_ISLOW inverse DCT, fancy chroma upsampling, YCbCr -> RGB: bit for bit what Pillow's libjpeg
 * produces).  CALD_JPEG_WALK=device moves the entropy walk to the device too (the compressed scan travels; slower,
 * DESIGN.md section 5).  Supported: 8-bit sequential Huffman JPEG (SOF0 /
 * SOF1), grayscale or YCbCr 4:4:4 / 4:2:2 / 4:2:0, restart intervals; anything else fails the call with a message.
 *
 * cald_jpeg_info   : frame size of one file (host only, needs no engine).
 * cald_jpeg_coefficients : the host half of the ingest on its own (needs no engine or GPU): the quantised DCT
 *                    coefficients of one file as the entropy walk produces them, int16, component after component,
 *                    each [blocks_h][blocks_w][64] in natural (de-zigzagged) order with blocks_w / blocks_h covering
 *                    whole MCUs.  blocks_w / blocks_h: int[3], filled for the file's components.  Returns -1 with a
 *                    message if `capacity` (in int16 elements) is too small; *n_coef always receives the needed count.
 * cald_jpeg_decode : out_images[i] = caller's HOST buffer of height*width*3 bytes (u8 RGB, HWC).
 * cald_score_jpeg  : cald_score() over files; out_heights / out_widths (optional) return the decoded sizes.  Noise
 *                    views (GAUSS / SALTPEPPER) need caller-drawn planes of the image size and are refused here. */
int cald_jpeg_info(const uint8_t* file, size_t size, int* height, int* width, int* components);
int cald_jpeg_coefficients(const uint8_t* file, size_t size, int16_t* out, size_t capacity, size_t* n_coef, int* blocks_w,
                           int* blocks_h);
int cald_jpeg_decode(cald_engine* e, int n_files, const uint8_t* const* files, const size_t* file_sizes,
                     uint8_t* const* out_images);
int cald_score_jpeg(cald_engine* e, int n_files, const uint8_t* const* files, const size_t* file_sizes, int n_augs,
                    const cald_aug* augs, double bp, const double* rng_uniforms, int n_uniforms, int* uniforms_consumed,
                    const int* swap_perms, double* out_consistency, double* out_cls, int* out_heights, int* out_widths);

/* On-device selection: the end of an AL cycle, cald_train.py:439-448 + cls_kldiv (234-271).
 * uncertainty[n] and cls[n][c1] are what cald_score returned for the whole (gathered) pool; mean_hist[c1] is the mean of
 * the labeled set's per-image label histograms (cald_train.py:237-242, 253); n_cand = int(args.mr * budget).
 * out_positions receives pool positions p (loader order) such that new_labeled = subset[p], in the reference's pick
 * order: candidates with an all-zero class vector first (all of them, even beyond `budget`), then by descending
 * JS(softmax(mean_hist) || softmax(cls)) (ascending JS against the uniform distribution when `uniform`).  Equal
 * uncertainty values resolve to the lower pool position (numpy's argsort leaves their order unspecified).
 * out_capacity >= n_cand.  At most 8192 candidates. */
int cald_select(cald_engine* e, int n, const double* uncertainty, const double* cls, int c1, const double* mean_hist,
                int budget, int n_cand, int uniform, int* out_positions, int out_capacity, int* n_picked);

/* The two detection-only baseline scorers of the reference, on the same engine:
 * LS+C  ls_c_train.get_uncertainty (ls_c_train.py:108-155): stability of the 30 most confident reference boxes under
 *       6 Gaussian-noise views (std 8..48) minus max(1 - prob_max).  noise: torch.randn(image.size()) planes, 6 per
 *       image, image-major (as for CALD_AUG_GAUSS).  Ties of prob_max at the top-30 cut resolve to the lower index.
 * LT/C  lt_c_train.get_uncertainty (lt_c_train.py:105-121): min(1, min |calcu_iou(box, prop) + prob_max - 1|) of one
 *       Faster R-CNN forward per image. */
int cald_score_lsc(cald_engine* e, int n_images, const uint8_t* const* images, const int* heights, const int* widths,
                   const float* const* noise, double* out_stability);
int cald_score_ltc(cald_engine* e, int n_images, const uint8_t* const* images, const int* heights, const int* widths,
                   double* out_uncertainty);

/* One detector forward per image: task_model([F.to_tensor(img)])[0] (frcnn_la.py:131-141).
 * Outputs are fixed-capacity [n][cap] with counts[n]; cap = box_detections_per_img.
 * scores_cls is [n][cap][num_classes].  Any output pointer may be NULL. */
int cald_detect(cald_engine* e, int n_images, const uint8_t* const* images, const int* heights, const int* widths,
                int* counts, float* boxes, float* scores, int64_t* labels, float* props, float* prob_max,
                float* scores_cls);

/* Number of detections in each image's reference view during the last cald_score() / cald_score_lsc() call (after the
 * >40 -> 50 sub-sampling for cald_score).  0 marks the images the reference skips before drawing any augmentation
 * randomness (cald_train.py:118-121, ls_c_train.py:118-120); the host wrapper rewinds its generators with it. */
int cald_last_ref_counts(cald_engine* e, int* out, int capacity);

/* Per-view consistency of the last cald_score() call: out[n_images][n_augs] (cald_train.py:223). */
int cald_last_per_view(cald_engine* e, float* out, int capacity);

/* Stage tensors of the LAST forward pass (debug=1): returns element count, or <0.  With buf == NULL only the size
 * is returned.  Names: "input", "c2".."c5", "p2".."p6", "rpn0".."rpn4", "proposals", "proposal_count", "pooled", "head". */
long long cald_debug_fetch(cald_engine* e, const char* name, float* buf, long long capacity);

/* debug = 1: the detections of every view of the last cald_score() call, ragged: counts[n_images * (1 + n_augs)] in
 * (image, [reference view, augmented views in order]) order, then that many rows of boxes[.][4] (in the view's own image
 * coordinates), scores, labels, prob_max.  Returns the total number of rows (call with NULL buffers to size them). */
long long cald_debug_views(cald_engine* e, int* counts, int counts_capacity, float* boxes, float* scores, int* labels,
                           float* prob_max, long long rows_capacity);

/* High-water mark of the device arena in bytes since creation (sizing max_views_per_pass / workspace_bytes). */
long long cald_arena_peak(cald_engine* e);

/* Views (detector forwards) batched through one pass: cfg.max_views_per_pass, or what cald_create derived from the
 * arena size when that was 0.  A scoring call processes max(1, views / n_augs) images per chunk. */
int cald_views_per_pass(cald_engine* e);

/* Counters since creation: kernels launched by this engine, algorithmic conv/GEMM FLOPs (2*MAC),
 * and device milliseconds spent inside tcgen05 conv kernels when timing is enabled. */
int cald_counters(cald_engine* e, long long* kernel_launches, double* conv_flops);

/* Measurement hooks (bench.py).  cald_profile(e,1) brackets every tcgen05 conv/GEMM launch with CUDA events on
 * the engine's stream; cald_profile_read() synchronises and returns the summed kernel time, launch count and
 * algorithmic FLOPs since the last read.  cald_event_record/elapsed time whole regions on the same stream. */
int cald_profile(cald_engine* e, int enable);
int cald_profile_read(cald_engine* e, double* conv_ms, long long* conv_launches, double* conv_flops);
/* Per-layer breakdown of the last cald_profile_read(): TSV text "layer count ms gflop mbyte" (algorithmic FLOPs and
 * HBM bytes per distinct conv problem); returns the size needed incl. the terminator (buf may be NULL). */
long long cald_profile_layers(cald_engine* e, char* buf, long long capacity);
int cald_event_record(cald_engine* e, int slot /* 0..7 */);
int cald_event_elapsed_ms(cald_engine* e, int slot_a, int slot_b, float* ms);

/* Same as cald_score but the u8 images already live in device memory (device pointers). */
int cald_score_device(cald_engine* e, int n_images, const uint8_t* const* d_images, const int* heights,
                      const int* widths, int n_augs, const cald_aug* augs, double bp, const double* rng_uniforms,
                      int n_uniforms, int* uniforms_consumed, const float* const* d_noise, const int* swap_perms,
                      double* out_consistency, double* out_cls);

#ifdef __cplusplus
}
#endif
#endif

/* Stage-level entry points of libcald_b200.so.
 *
 * These expose single stages of the scoring pipeline on HOST buffers so that the
 * stage-wise parity tests (tests/test_gpu_*.py) can compare each hand-written
 * kernel with the CPU oracle in isolation ("parity must be
 * asserted stage-wise").  They are not needed by a caller of cald_score().
 *
 * All functions return 0 on success, <0 on error; cald_ops_last_error() returns
 * the message of the last failure on the calling thread.
 */
#ifndef CALD_B200_OPS_H
#define CALD_B200_OPS_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

const char* cald_ops_last_error(void);

/* conv2d / linear on the implicit-GEMM engine.
 * Replaces: torch.nn.functional.conv2d as used by torchvision resnet.py Bottleneck,
 *           ops/feature_pyramid_network.py:172-204, models/detection/rpn.py:71-78 and
 *           F.linear in models/detection/faster_rcnn.py:286-307 (a 1x1 conv on 1x1 maps).
 * x:      [n][h][w][cin]  fp32 NHWC (cin multiple of 64)
 * weight: [cout][cin][k][k] fp32 (torch layout), k in {1,3}; pad = k/2; stride in {1,2} (stride 2 reads the
 *         full-resolution input through an element-strided TMA tensor map)
 * bias:   [cout] or NULL
 * res:    optional residual added before ReLU, NHWC fp32 [n][res_h][res_w][cout];
 *         res_mode 0 none, 1 same shape, 2 nearest-upsampled to the output size
 * prec:   0 = split-bf16 x3 (fp32-faithful), 1 = single-pass bf16
 * impl:   0 = tcgen05 kernel, 1 = SIMT checker kernel
 * kc:     k-blocks per accumulation chunk in split mode (-1 = engine default, 0 = never chunk)
 * out:    [n][ho][wo][cout] fp32
 */
int cald_op_conv2d(const float* x, int n, int h, int w, int cin, const float* weight, const float* bias, int cout,
                   int k, int stride, int relu, const float* res, int res_mode, int res_h, int res_w, int prec,
                   int impl, int block_n, int kc, float* out);

/* conv2d with a second 1x1 contraction accumulated into the same output tile (one launch, one accumulator):
 *   out = act(conv_k(x, weight) + bias + conv_1x1(x2[::stride2, ::stride2], weight2) + bias2)
 * Replaces: the tail of torchvision resnet.py Bottleneck.forward for a stage's first block,
 *           out = relu(bn3(conv3(out)) + downsample(x)), with both FrozenBN layers folded (tv:models/resnet.py:143-163).
 * x:  [n][h][w][cin] fp32 NHWC, weight [cout][cin][k][k] (k in {1,3}, stride 1); x2: [n][h2][w2][cin2] with
 * ceil(h2 / stride2) == h and ceil(w2 / stride2) == w, weight2 [cout][cin2][1][1]; split-bf16 x3 arithmetic.
 * out: [n][h][w][cout] fp32. */
int cald_op_conv2d_dual(const float* x, int n, int h, int w, int cin, const float* weight, const float* bias, int cout,
                        int k, const float* x2, int h2, int w2, int cin2, const float* weight2, const float* bias2,
                        int stride2, int relu, float* out);

/* Number of launches of the CTA-pair (tcgen05 cta_group::2) conv kernel in this process so far; the parity tests use
 * it to assert which kernel a call exercised (CALD_CTA2=0/1 selects it, see cald_b200/csrc/conv_host.cuh). */
long long cald_ops_pair_launches(void);
/* Same for the transposed-role kernel of the 64-output-channel layers (CALD_TFORM=0/1, cald_b200/csrc/igemm_t.cuh). */
long long cald_ops_tform_launches(void);

/* Pillow-exact augmentation images on the device (cald/cald_helper.py:47-53 resize, 135-223 rotate).
 * kind 2 = img.resize((int(w*0.8), int(h*0.8)), BILINEAR); kind 3 = img.rotate(5, expand=True).resize((w, h)) (BICUBIC).
 * img: u8 [h][w][3].  out must hold h*w*3 bytes; the produced size is returned in out_h / out_w. */
int cald_op_aug_image(int kind, const uint8_t* img, int h, int w, uint8_t* out, int* out_h, int* out_w);

/* cald_helper.ColorAdjust(image, factor) (cald/cald_helper.py:65-69): PIL.ImageEnhance Brightness -> Contrast -> Color
 * with the same factor, bit-exact u8 arithmetic of libImaging/Blend.c / Convert.c.  img, out: u8 [h][w][3]. */
int cald_op_color_adjust(const uint8_t* img, int h, int w, double factor, uint8_t* out);

/* Measurement probe (tools/overlap_probe.py): times a train of `iters` conv launches alone (out[0], ms) and with a
 * 32-block one-warp kernel that spins for spin_ms (and declares spin_smem_bytes of dynamic shared memory) launched on a
 * second stream after the second conv (out[1]); out[2] / out[3] = the spin kernel's own start-to-end time alone / inside
 * the train.  Answers whether small kernels can co-reside with the persistent conv CTAs (one per SM, ~226 KB smem). */
int cald_op_overlap_probe(int n, int h, int w, int cin, int cout, int k, int iters, int spin_ms, int spin_smem_bytes,
                          double* out);

#ifdef __cplusplus
}
#endif
#endif

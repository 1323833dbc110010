/*
 * yomitoku_b200 C ABI (libytk_b200.so) - the drop-in boundary for the DBNet -> PARSeq hot path.
 *
 * The reference (kotaro-kinoshita/yomitoku) has no FFI: its seam is the Python object protocol
 * `self.model(tensor)` inside TextDetector / TextRecognizer (reference src/yomitoku/text_detector.py:127-131,
 * src/yomitoku/text_recognizer.py:247-256).  These entry points are what a ctypes binding
 * behind those two call sites binds to; INTEGRATION.md shows the stub.  Plain pointers and sizes only, no torch
 * types; all device pointers are caller-owned; every call returns 0 on success and a nonzero code on failure, with a
 * human-readable message from ytk_last_error() (thread-local).  No exceptions cross this boundary.
 */
#ifndef YOMITOKU_B200_H
#define YOMITOKU_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YTK_OK 0
#define YTK_ERR 1

/* activation codes for the op-level entry points */
#define YTK_ACT_NONE 0
#define YTK_ACT_RELU 1
#define YTK_ACT_GELU 2
#define YTK_ACT_SIGMOID 3

const char* ytk_last_error(void);
int ytk_version(void);
/* number of kernel launches issued by this library on the calling process since load (bench.py gpu_launches) */
long long ytk_launch_count(void);

/* ---- op level (kernel parity tests; replaces the cuDNN/cuBLAS call sites listed in SURVEY.md section 2.3) ----
 * Convolution as tcgen05 implicit GEMM.  in: NHWC bf16 [N,H,W,in_ld] (first Cin channels used), w: bf16
 * [Cout][kh][kw][Cin], bias fp32 [Cout] or NULL, resid: [N,Ho,Wo,ldr] bf16/fp32 or NULL, out: [N,Ho,Wo,ldc]
 * bf16/fp32.  Replaces torch.nn.Conv2d + BatchNorm2d(eval, folded) + ReLU (+ residual add) of
 * torchvision ResNet-50 bottlenecks (reference models/dbnet_plus.py:30-38) and the decoder convs (:56-116).
 * mode 1 = ConvTranspose2d(kernel 2, stride 2) written as a GEMM with a pixel-shuffle epilogue (:111,:114). */
int ytk_op_conv2d_bf16(const void* in, int N, int H, int W, int Cin, long long in_ld, const void* w, const float* bias,
                       int kh, int kw, int stride, int pad, int dil, int Cout, const void* resid, int resid_f32,
                       long long ldr, void* out, int out_f32, long long ldc, int act, int mode, void* cuda_stream);

/* Linear layer y = act(A W^T + b (+ resid)); A [M,lda] bf16, W [N,K] bf16 (torch nn.Linear layout), K % 64 == 0.
 * Replaces nn.Linear / timm Mlp / attention projections (reference models/layers/parseq_transformer.py:43-52,
 * models/parseq.py:72). */
int ytk_op_linear_bf16(const void* A, long long lda, int M, int K, const void* W, int N, const float* bias,
                       const void* resid, int resid_f32, long long ldr, void* out, int out_f32, long long ldc, int act,
                       void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* YOMITOKU_B200_H */

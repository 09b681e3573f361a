/*
 * tlxcv_b200.h — C ABI of the B200-native CNN-backbone forward path.
 *
 * The reference (tensorlayer/TLXCV) has no FFI of its own: its hot path is
 * Python model code over tensorlayerx layers that, with TL_BACKEND=torch, end in
 * torch.nn.functional calls — one per layer, 175 per ResNet-50 forward
 * (SURVEY.md §3.2).  This library replaces that chain, ONE call per forward:
 *
 *   reference interface replaced                                   entry point
 *   ------------------------------------------------------------   --------------------
 *   ResNet.forward      classification/resnet.py:286-300            tlxcv_plan_run
 *   ResNeXt.forward     classification/resnext.py:201-209           tlxcv_plan_run
 *   MobileNetV2.forward classification/mobilenetv2.py:102-109       tlxcv_plan_run
 *   MobileNetV1.forward classification/mobilenetv1.py:254-262       tlxcv_plan_run
 *   DarkNet.forward     detection/backbones/darknet.py:299-312      tlxcv_plan_run
 *   DarkNet.forward     classification/darknet53.py:100-133         tlxcv_plan_run
 *   ImageClassification.predict tasks/image_classification.py:20-23 tlxcv_plan_run (+ARGMAX op, fused into the
 *                                                                   Linear launch that produces the logits)
 *   nn.GroupConv2d + nn.BatchNorm2d + nn.ReLU/ReLU6/LeakyReLU + add TLXCV_OP_CONV
 *   nn.MaxPool2d                                                    TLXCV_OP_MAXPOOL
 *   nn.AdaptiveAvgPool2d(1)                                         TLXCV_OP_GAP
 *   nn.AvgPool2d(2, 2)   segmentation/backbones/resnet_vd.py:25-27  TLXCV_OP_AVGPOOL
 *   nn.AvgPool2d(3, 2, 1) classification/resnest.py:245-250         TLXCV_OP_AVGPOOL (pad = 1)
 *   SplatConv.forward    classification/resnest.py:146-166          TLXCV_OP_CONV (dense 3x3) + TLXCV_OP_GAP + 2 x
 *                                                                   TLXCV_OP_CONV (1x1) + TLXCV_OP_SPLAT_APPLY
 *   nn.Linear                                                       TLXCV_OP_LINEAR
 *   tlx.losses.softmax_cross_entropy_with_logits  tasks/image_classification.py:14  TLXCV_OP_SOFTMAX_CE
 *   Interpolater (nearest x2) + tlx.concat  detection/yolov3.py:244,252  TLXCV_OP_UPSAMPLE_CONCAT
 *   YOLOv3FPN / YOLOv3Head convs     detection/yolov3.py:122-258,306-378  TLXCV_OP_CONV (bias, any C_out)
 *   Module construction / set_eval (weights become static)          tlxcv_plan_build
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ / torch types cross this boundary.
 *   - every function returns TLXCV_OK (0) or a negative tlxcv_status and never
 *     throws; the message is kept per context (tlxcv_last_error).
 *   - all device buffers passed in are BORROWED.  Parameters (filters, BN
 *     statistics ...) are read once, inside tlxcv_plan_build, in the reference's
 *     own layout (OIHW fp32 conv filters, (in,out) fp32 Linear weights); the
 *     library packs its own bf16/KRSC copies and folded scale/shift and owns
 *     those, its activation workspace, TMA descriptors and CUDA graphs.
 *   - plan inputs are NCHW fp32 (the reference's tensors); 4-D outputs come back
 *     NCHW fp32, logits as (N, classes) fp32, ARGMAX results as (N) int64.
 *   - a context is bound to one device and used by one host thread at a time
 *     (one process per GPU).  All work is enqueued on the caller's stream; there
 *     is no hidden synchronisation except in the *_destroy calls and
 *     tlxcv_plan_profile.
 */
#ifndef TLXCV_B200_H
#define TLXCV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TLXCV_ABI_VERSION 1

typedef enum {
  TLXCV_OK = 0,
  TLXCV_ERR_INVALID = -1,      /* bad argument / inconsistent description      */
  TLXCV_ERR_UNSUPPORTED = -2,  /* valid, but outside the hot path this library covers */
  TLXCV_ERR_CUDA = -3,         /* a CUDA runtime / driver call failed           */
  TLXCV_ERR_NO_DEVICE = -4,    /* no sm_100 device; there is NO CPU fallback    */
  TLXCV_ERR_OOM = -5
} tlxcv_status;

typedef enum { TLXCV_F32 = 0, TLXCV_BF16 = 1, TLXCV_I64 = 2, TLXCV_ACT = 3, TLXCV_U8 = 4 } tlxcv_dtype;
/* TLXCV_ACT: the plan's activation type — bf16 (TLXCV_PREC_BF16) or fp32 (TLXCV_PREC_F32_VALIDATE). */

typedef enum { TLXCV_ROLE_INTERNAL = 0, TLXCV_ROLE_INPUT = 1, TLXCV_ROLE_OUTPUT = 2 } tlxcv_role;

typedef enum {
  TLXCV_OP_IMPORT_NCHW = 0,  /* external NCHW fp32 -> internal NHWC activation                     */
  TLXCV_OP_CONV = 1,         /* act2(act1(conv(x)*scale+shift) + residual), dense/grouped/depthwise */
  TLXCV_OP_MAXPOOL = 2,
  TLXCV_OP_GAP = 3,          /* global average pool -> (N, C)                                      */
  TLXCV_OP_LINEAR = 4,       /* (N, F) x (F, K) + bias -> fp32 logits                              */
  TLXCV_OP_ADD_ACT = 5,      /* act2(in0 [+ in1]) — only for adds / activations no conv could absorb */
  TLXCV_OP_ARGMAX = 6,       /* (N, K) fp32 -> (N) int64                                           */
  TLXCV_OP_EXPORT_NCHW = 7,  /* internal NHWC activation -> external NCHW fp32                     */
  TLXCV_OP_IMPORT_U8_NHWC = 8, /* external NHWC uint8 image batch (N, H, W, C<=4) -> internal activation
                                 (x - mean[c]) / std[c]: the reference's host-side Normalize + ToTensor
                                 (demo/image_classification/predict-resnet.py:50-54) fused into the layout
                                 pass; bn_mean / bn_var carry device pointers to float mean[C] / std[C]      */
  TLXCV_OP_UPSAMPLE_CONCAT = 9, /* out[..., :C0] = in0 nearest-up-sampled r times, out[..., C0:] = in1 up-sampled s
                                 times (in1 = -1: up-sampling alone; r = s = 1: channel concat): Interpolater +
                                 tlx.concat of YOLOv3FPN.forward (detection/yolov3.py:244,252-253) in one pass          */
  TLXCV_OP_SOFTMAX = 10,     /* (N, K) fp32 logits -> (N, K) fp32 probabilities (softmax over the class axis)           */
  TLXCV_OP_AVGPOOL = 12,     /* AvgPool2d(k, stride, pad), padding counted as zeros (ResNet_vd / ResNeSt shortcut, segmentation/backbones/resnet_vd.py:25-27;
                                ResNeSt avd pool 3x3 / 2 / 1, classification/resnest.py:245-250) */
  TLXCV_OP_SPLAT_APPLY = 13, /* ResNeSt split attention (classification/resnest.py:53-82,146-166): in0 = (N, radix*C, H, W) map, in1 = (N, radix*C, 1, 1)
                                attention logits, r = radix, groups = cardinality; out[n,c] = sum_r softmax_r(in1)[r,c] * in0[n, r*C + c] */
  TLXCV_OP_SOFTMAX_CE = 11   /* in0 = (N, K) fp32 logits, in1 = (N) int64 labels -> (1) fp32 mean cross-entropy:
                                 tlx.losses.softmax_cross_entropy_with_logits (tasks/image_classification.py:10-15)        */
} tlxcv_op_kind;

typedef enum { TLXCV_ACT_NONE = 0, TLXCV_ACT_RELU = 1, TLXCV_ACT_RELU6 = 2, TLXCV_ACT_LEAKY = 3 } tlxcv_act;

typedef enum {
  TLXCV_PREC_BF16 = 0,         /* bf16 activations/weights, fp32 accumulate in TMEM, fp32 epilogue  */
  TLXCV_PREC_F32_VALIDATE = 1  /* fp32 everywhere on CUDA cores: the <=1e-4 validation mode         */
} tlxcv_precision;

/* Logical tensor: NHWC dims; (N, F) tensors have h = w = 1, c = F; (N) has c = 1. */
typedef struct {
  int32_t n, h, w, c;
  int32_t dtype; /* tlxcv_dtype */
  int32_t role;  /* tlxcv_role  */
} tlxcv_tensor_desc;

typedef struct {
  int32_t kind;          /* tlxcv_op_kind */
  int32_t in0, in1, out; /* tensor indices; in1 = residual / second addend, -1 if none */
  int32_t r, s, stride, pad, dil, groups;
  int32_t act1;
  float alpha1; /* LeakyReLU slope */
  int32_t act2;
  float alpha2;
  /* reference-layout fp32 parameters on the device (borrowed; read during plan_build) */
  const float* filters;  /* conv: OIHW [K][C/g][R][S]; linear: (in,out) [F][K]            */
  const float* bias;     /* [K] or NULL (GroupConv2d b_init / Linear biases)               */
  const float* bn_gamma; /* [K] or NULL: BatchNorm folded as scale = gamma/sqrt(var+eps)   */
  const float* bn_beta;  /*               shift = beta + (bias - mean) * scale             */
  const float* bn_mean;
  const float* bn_var;
  float bn_eps;
  int32_t reserved;
} tlxcv_op_desc;

/* Per-op report (tlxcv_plan_op_info): what ran and the algorithmic work it did. */
typedef struct {
  char kernel[48];      /* kernel family that executes the op, e.g. "conv_tcgen05_im2col_n256" */
  int32_t launches;     /* kernel launches per plan run                                        */
  int32_t bound;        /* 0 = HBM, 1 = tensor                                                  */
  double flops;         /* 2*MACs per run (conv / linear), else 0                               */
  double bytes;         /* algorithmic bytes per run: inputs + weights read once, output written once */
  int32_t grid, block, smem_bytes;
  int32_t tile_n;       /* MMA N tile for tcgen05 convs, else 0                                 */
} tlxcv_op_info;

typedef struct tlxcv_ctx tlxcv_ctx;
typedef struct tlxcv_plan tlxcv_plan;

/* ---- context ----------------------------------------------------------------------------- */
int tlxcv_abi_version(void);
int tlxcv_create(int device, tlxcv_ctx** out);
int tlxcv_destroy(tlxcv_ctx* ctx);
const char* tlxcv_last_error(const tlxcv_ctx* ctx); /* ctx may be NULL: last error of a failed tlxcv_create */
int tlxcv_device_sm_count(const tlxcv_ctx* ctx);

/* ---- plan ---------------------------------------------------------------------------------- */
/* Build a static plan: allocate the activation workspace, pack weights, fold BN, choose kernels
 * and tiles, encode TMA descriptors.  `stream` (a cudaStream_t) orders the one-time packing
 * kernels after the caller's parameter uploads.  */
int tlxcv_plan_build(tlxcv_ctx* ctx, const tlxcv_tensor_desc* tensors, int n_tensors, const tlxcv_op_desc* ops,
                     int n_ops, int precision, void* stream, tlxcv_plan** out);
int tlxcv_plan_destroy(tlxcv_plan* plan);

/* Run the plan.  `inputs` / `outputs`: device pointers for the ROLE_INPUT / ROLE_OUTPUT tensors in
 * tensor-table order.  With use_graph != 0 the launch sequence is captured into a CUDA graph on the
 * first call for a given set of pointers and replayed afterwards. */
int tlxcv_plan_run(tlxcv_plan* plan, const void* const* inputs, void* const* outputs, void* stream, int use_graph);

/* Same, from/to pinned HOST buffers: H2D copies, run, D2H copies, all on `stream`
 * (the end-to-end call a host-side user of the reference's API makes). */
int tlxcv_plan_run_host(tlxcv_plan* plan, const void* const* host_inputs, void* const* host_outputs, void* stream,
                        int use_graph);

/* Synchronous profiling run: per-op device milliseconds (CUDA events around each op). */
int tlxcv_plan_profile(tlxcv_plan* plan, const void* const* inputs, void* const* outputs, void* stream,
                       float* per_op_ms, int n_ops);

int tlxcv_plan_num_ops(const tlxcv_plan* plan);
int tlxcv_plan_num_launches(const tlxcv_plan* plan); /* kernel launches per run */
int tlxcv_plan_op_info(const tlxcv_plan* plan, int op_index, tlxcv_op_info* info);
size_t tlxcv_plan_workspace_bytes(const tlxcv_plan* plan);
/* Debug/validation: copy internal tensor `tensor_index` (NHWC, plan activation dtype) to a device
 * buffer of at least n*h*w*c_storage elements after a run; returns the storage channel count. */
int tlxcv_plan_read_tensor(tlxcv_plan* plan, int tensor_index, void* dst_device, size_t dst_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* TLXCV_B200_H */

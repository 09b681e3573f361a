// C ABI of libtlxcv_b200.so (include/tlxcv_b200.h): context, plan build (workspace planning,
// weight packing, BN folding, kernel/tile selection, TMA descriptors), plan run (stream launch or
// CUDA-graph replay), host-buffer run, per-op profiling.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "kernels.h"

using namespace tlxcv;

namespace {

thread_local std::string g_create_error;

struct TensorRt {
  tlxcv_tensor_desc d;
  int cs = 0;          // storage channels (NHWC4 for C<=4 activations)
  size_t bytes = 0;
  size_t offset = 0;   // into the arena (internal tensors)
  int ext_index = -1;  // position among inputs / outputs (external tensors)
  int first_def = -1, last_use = -1;
  int wp = 0, pad_l = 0;  // C<=4 stem inputs: rows stored with zero pad columns, pitch wp pixels (0 = dense)
  bool elided = false;    // never materialised (the conv map between a stem and its fused max-pool)
};

enum Impl : int { kImplImport, kImplExport, kImplTcConv, kImplDwConv, kImplDirectF32, kImplMaxpool, kImplGap, kImplAddAct, kImplArgmax,
                  kImplStem, kImplSlab, kImplNop, kImplImportU8, kImplUpsampleConcat, kImplSoftmax, kImplSoftmaxCe, kImplAvgpool, kImplSplatApply };

struct OpRt {
  tlxcv_op_desc d;
  int impl = 0;
  TcConvLaunch tc;            // kImplTcConv
  StemLaunch stem;            // kImplStem
  SlabLaunch slab;            // kImplSlab
  StemGeometry geo;
  bool use_stem = false;      // row-ring stem kernel chosen in the pre-pass
  int pool_op = -1;           // index of the max-pool op fused into this stem (-1: none)
  bool nop = false;           // fused into another op: no launch, no tensors of its own
  bool chain_first = false;   // ... as the first GEMM of a chain kernel (report only)
  int dual_a = -1;            // index of the 1x1 conv whose output was this conv's residual and now runs as the
                              // first accumulator of this op's dual-GEMM kernel (-1: none)
  int chain_mid = -1;         // chain kernel: the (elided) tensor between the two convs; the op's in0 is rewired to the first conv's input
  int chain_a = -1;           // 1x1 conv: index of the conv that produced its input and now runs as the first GEMM of this op's
                              // chain kernel (-1: none)
  int fused_argmax = -1;      // Linear: index of the ARGMAX op over its logits that runs inside this launch (-1: none)
  bool skip_logits = false;   // ... and nothing else reads the logits: they are never written
  unsigned long long* amax_keys = nullptr;  // fused argmax scratch (owned, zero between launches)
  unsigned int* ticket = nullptr;           // fused argmax / softmax-CE "last block" counter (owned, zero between launches)
  float* row_loss = nullptr;                // softmax-CE per-row scratch (owned)
  int4* resize_tx = nullptr;                // import_u8 with a resize in front: per-column / per-row {i0, i1, c0, c1} (owned)
  int4* resize_ty = nullptr;
  float* scale2 = nullptr;    // folded BN of this op's own (second) GEMM in a dual launch (owned)
  float* shift2 = nullptr;
  void* weights = nullptr;    // packed weights (owned)
  float* scale = nullptr;     // folded BN / bias (owned)
  float* shift = nullptr;
  tlxcv_op_info info;
};

}  // namespace

struct tlxcv_ctx {
  int device = 0;
  int sm_count = 0;
  std::string error;
};

struct tlxcv_plan {
  tlxcv_ctx* ctx = nullptr;
  int precision = 0;
  bool f32 = false;
  size_t esize = 2;
  std::vector<TensorRt> tensors;
  std::vector<OpRt> ops;
  std::vector<int> input_ids, output_ids;
  uint8_t* arena = nullptr;
  size_t arena_bytes = 0;
  std::vector<void*> owned;
  // CUDA graph cache keyed by the external pointers (same pointers => replay); a few entries so that
  // double-buffered callers do not re-capture every step
  std::vector<std::pair<std::vector<const void*>, cudaGraphExec_t>> graphs;
  std::mutex graphs_mu;  // ctypes releases the GIL around plan_run: two host threads may look up / capture at once
  // A pointer set seen for the first time runs in SEGMENTS instead: the runs of ops that touch only the plan's own
  // workspace are captured ONCE (independent of the caller's pointers) and the few ops that read an external input or
  // write an external output are launched directly around them, in op order.  A caller that allocates fresh outputs on
  // every forward therefore never re-captures; a pointer set that comes back is promoted to a whole-forward graph.
  struct Segment {
    int begin, end;
    bool graphable;
    cudaGraphExec_t exec;
  };
  std::vector<Segment> segments;
  std::vector<std::vector<const void*>> seen_keys;  // pointer sets run once in segment mode (bounded)
  // host-run staging
  std::vector<void*> stage_in, stage_out;
};

namespace {

int fail(tlxcv_ctx* ctx, int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (ctx)
    ctx->error = buf;
  else
    g_create_error = buf;
  return code;
}

#define TLX_CUDA(ctx, expr)                                                                              \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      return fail(ctx, _e == cudaErrorMemoryAllocation ? TLXCV_ERR_OOM : TLXCV_ERR_CUDA, "%s failed: %s", \
                  #expr, cudaGetErrorString(_e));                                                        \
  } while (0)

size_t dtype_size(const tlxcv_plan* p, int dt) {
  switch (dt) {
    case TLXCV_F32: return 4;
    case TLXCV_BF16: return 2;
    case TLXCV_I64: return 8;
    case TLXCV_U8: return 1;
    default: return p->esize;
  }
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Makes `device` current for the scope and restores the caller's device afterwards: a process with several GPUs (or a
// plan destroyed from a garbage collector at an arbitrary moment) must not find its current device changed.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != device) switched = cudaSetDevice(device) == cudaSuccess;
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

// first-fit arena with coalescing free list
struct Arena {
  std::map<size_t, size_t> free_blocks;  // offset -> size
  size_t end = 0;
  size_t alloc(size_t bytes) {
    bytes = align_up(bytes, 1024);
    for (auto it = free_blocks.begin(); it != free_blocks.end(); ++it) {
      if (it->second >= bytes) {
        size_t off = it->first, rest = it->second - bytes;
        free_blocks.erase(it);
        if (rest) free_blocks[off + bytes] = rest;
        return off;
      }
    }
    // extend a trailing free block if there is one
    if (!free_blocks.empty()) {
      auto last = std::prev(free_blocks.end());
      if (last->first + last->second == end) {
        size_t off = last->first;
        free_blocks.erase(last);
        end = off + bytes;
        return off;
      }
    }
    size_t off = end;
    end += bytes;
    return off;
  }
  void release(size_t off, size_t bytes) {
    bytes = align_up(bytes, 1024);
    auto it = free_blocks.emplace(off, bytes).first;
    auto nxt = std::next(it);
    if (nxt != free_blocks.end() && it->first + it->second == nxt->first) {
      it->second += nxt->second;
      free_blocks.erase(nxt);
    }
    if (it != free_blocks.begin()) {
      auto prv = std::prev(it);
      if (prv->first + prv->second == it->first) {
        prv->second += it->second;
        free_blocks.erase(it);
      }
    }
  }
};

void* tensor_ptr(const tlxcv_plan* p, int t, const void* const* inputs, void* const* outputs) {
  const TensorRt& T = p->tensors[t];
  if (T.d.role == TLXCV_ROLE_INPUT) return const_cast<void*>(inputs[T.ext_index]);
  if (T.d.role == TLXCV_ROLE_OUTPUT) return outputs[T.ext_index];
  return p->arena + T.offset;
}

template <typename T>
int dev_alloc(tlxcv_plan* p, T** out, size_t count) {
  void* ptr = nullptr;
  cudaError_t e = cudaMalloc(&ptr, std::max<size_t>(count * sizeof(T), 256));
  if (e != cudaSuccess) return fail(p->ctx, TLXCV_ERR_OOM, "cudaMalloc(%zu) failed: %s", count * sizeof(T), cudaGetErrorString(e));
  p->owned.push_back(ptr);
  *out = static_cast<T*>(ptr);
  return TLXCV_OK;
}

void set_info(OpRt& op, const char* kernel, int launches, int bound, double flops, double bytes, int grid, int block,
              int smem, int tile_n) {
  memset(&op.info, 0, sizeof op.info);
  snprintf(op.info.kernel, sizeof op.info.kernel, "%s", kernel);
  op.info.launches = launches, op.info.bound = bound, op.info.flops = flops, op.info.bytes = bytes;
  op.info.grid = grid, op.info.block = block, op.info.smem_bytes = smem, op.info.tile_n = tile_n;
}

// OpenCV's INTER_LINEAR coefficient set-up for 8-bit images (imgproc/src/resize.cpp): source indices and 11-bit weights per
// destination index.  Columns clamp the FRACTION at the image border, rows clamp the source INDICES (oracle/cv_resize.py).
std::vector<int4> resize_table(int dst, int src, bool clamp_fraction) {
  std::vector<int4> t(dst);
  const double scale = 1.0 / (static_cast<double>(dst) / static_cast<double>(src));
  for (int d = 0; d < dst; ++d) {
    float f = static_cast<float>((d + 0.5) * scale - 0.5);
    int s = static_cast<int>(std::floor(f));
    f -= static_cast<float>(s);
    if (clamp_fraction) {
      if (s < 0) f = 0.0f, s = 0;
      if (s >= src - 1) f = 0.0f, s = src - 1;
    }
    const int c0 = static_cast<int>(std::nearbyintf((1.0f - f) * 2048.0f)), c1 = static_cast<int>(std::nearbyintf(f * 2048.0f));
    t[d] = make_int4(std::min(std::max(s, 0), src - 1), std::min(std::max(s + 1, 0), src - 1), c0, c1);
  }
  return t;
}

// C_in <= 4 stem on the row-ring kernel (stem_rowring.cu); `out` is the pooled map when a max-pool is fused
int compile_stem(tlxcv_plan* p, OpRt& op, cudaStream_t st) {
  tlxcv_ctx* ctx = p->ctx;
  const tlxcv_op_desc& d = op.d;
  const TensorRt& in = p->tensors[d.in0];
  const TensorRt& out = p->tensors[d.out];
  const StemGeometry& g = op.geo;
  const int N = in.d.n, H = in.d.h, C = in.d.c, K = g.pairs ? g.block_n / 2 : g.block_n;
  const bool pool = op.pool_op >= 0;
  if (!d.filters) return fail(ctx, TLXCV_ERR_INVALID, "op conv: filters pointer is NULL");
  if (pool ? (out.d.c != K || out.d.n != N) : (out.d.h != g.P || out.d.w != g.Q || out.d.c != K || out.d.n != N))
    return fail(ctx, TLXCV_ERR_INVALID, "stem conv: output tensor shape mismatch");
  if (out.d.role != TLXCV_ROLE_INTERNAL || out.d.dtype != TLXCV_ACT)
    return fail(ctx, TLXCV_ERR_UNSUPPORTED, "stem conv: output must be an internal activation tensor");
  if (pool) {
    const int Pp = (g.P + 2 - 3) / 2 + 1, Qp = (g.Q + 2 - 3) / 2 + 1;
    if (out.d.h != Pp || out.d.w != Qp) return fail(ctx, TLXCV_ERR_INVALID, "stem conv + max-pool: pooled shape mismatch");
  }
  int rc;
  if ((rc = dev_alloc(p, &op.scale, 256)) != TLXCV_OK) return rc;
  if ((rc = dev_alloc(p, &op.shift, 256)) != TLXCV_OK) return rc;
  TLX_CUDA(ctx, fold_bn(op.scale, op.shift, d.bn_gamma, d.bn_beta, d.bn_mean, d.bn_var, d.bias, d.bn_eps, K, 256, st));
  if (g.pairs) {  // GEMM column (parity, channel): the per-channel values twice
    TLX_CUDA(ctx, cudaMemcpyAsync(op.scale + K, op.scale, K * sizeof(float), cudaMemcpyDeviceToDevice, st));
    TLX_CUDA(ctx, cudaMemcpyAsync(op.shift + K, op.shift, K * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  __nv_bfloat16* w = nullptr;
  if ((rc = dev_alloc(p, &w, static_cast<size_t>(stem_rowring_weight_elems(g, d.r, d.stride)))) != TLXCV_OK) return rc;
  TLX_CUDA(ctx, pack_stem_weights(d.filters, w, K, C, d.r, d.s, d.stride, g, st));
  op.weights = w;
  std::string err = stem_rowring_prepare(op.stem, ctx->sm_count, g, reinterpret_cast<const __nv_bfloat16*>(p->arena + in.offset), N,
                                         H, d.r, d.stride, d.pad, w, p->arena + out.offset, pool ? 1 : 0, pool ? out.d.h : 0,
                                         pool ? out.d.w : 0);
  if (!err.empty()) return fail(ctx, TLXCV_ERR_UNSUPPORTED, "%s", err.c_str());
  StemParams& sp = op.stem.p;
  if (d.act1 == TLXCV_ACT_LEAKY && !(d.alpha1 >= 0.0f && d.alpha1 <= 1.0f))  // the epilogue evaluates max(v, alpha * v)
    return fail(ctx, TLXCV_ERR_UNSUPPORTED, "stem conv: LeakyReLU slope outside [0, 1]");
  sp.scale = op.scale, sp.shift = op.shift, sp.act = d.act1, sp.alpha = d.alpha1;
  op.impl = kImplStem;
  const double M = static_cast<double>(N) * g.P * g.Q;
  const double flops = 2.0 * M * K * C * d.r * d.s;
  const double bytes = static_cast<double>(N) * H * in.d.w * 4 * 2 + static_cast<double>(K) * C * d.r * d.s * 2 +
                       static_cast<double>(out.bytes);
  char name[48];
  snprintf(name, sizeof name, "stem_rowring_n%d%s%s", g.block_n, g.pairs ? "_pairs" : "", pool ? "_maxpool" : "");
  set_info(op, name, 1, 0, flops, bytes, op.stem.grid, op.stem.threads, op.stem.smem, g.block_n);
  return TLXCV_OK;
}

// B = the block's 1x1 (strided) downsample conv, A = the bottleneck's last 1x1 conv whose output was B's residual:
// one kernel, two TMEM accumulators, out = act2( A-branch + B-branch ) with both folded BNs applied in fp32
int compile_conv_dual(tlxcv_plan* p, OpRt& op, cudaStream_t st) {
  tlxcv_ctx* ctx = p->ctx;
  const tlxcv_op_desc& d = op.d;
  const tlxcv_op_desc& a = p->ops[op.dual_a].d;
  const TensorRt& x2 = p->tensors[a.in0];
  const TensorRt& xin = p->tensors[d.in0];
  const TensorRt& out = p->tensors[d.out];
  const int N = xin.d.n, H2 = xin.d.h, W2 = xin.d.w, C2 = xin.d.c, K1 = x2.d.c, K = out.d.c;
  const long long M = static_cast<long long>(out.d.n) * out.d.h * out.d.w;
  if (!d.filters || !a.filters) return fail(ctx, TLXCV_ERR_INVALID, "op conv: filters pointer is NULL");
  if (static_cast<long long>(x2.d.n) * x2.d.h * x2.d.w != M || p->tensors[a.out].d.c != K)
    return fail(ctx, TLXCV_ERR_INVALID, "dual conv: branch shapes differ");
  if (x2.d.role != TLXCV_ROLE_INTERNAL || xin.d.role != TLXCV_ROLE_INTERNAL || x2.cs != K1 || xin.cs != C2)
    return fail(ctx, TLXCV_ERR_UNSUPPORTED, "dual conv: inputs must be internal dense activation tensors");
  const int K_pad = static_cast<int>(align_up(K, 256));
  int rc;
  if ((rc = dev_alloc(p, &op.scale, K_pad)) != TLXCV_OK) return rc;
  if ((rc = dev_alloc(p, &op.shift, K_pad)) != TLXCV_OK) return rc;
  if ((rc = dev_alloc(p, &op.scale2, K_pad)) != TLXCV_OK) return rc;
  if ((rc = dev_alloc(p, &op.shift2, K_pad)) != TLXCV_OK) return rc;
  TLX_CUDA(ctx, fold_bn(op.scale, op.shift, a.bn_gamma, a.bn_beta, a.bn_mean, a.bn_var, a.bias, a.bn_eps, K, K_pad, st));
  TLX_CUDA(ctx, fold_bn(op.scale2, op.shift2, d.bn_gamma, d.bn_beta, d.bn_mean, d.bn_var, d.bias, d.bn_eps, K, K_pad, st));
  const int Kt1 = tc_conv_packed_k(K1, 1, 1, 1, kModeTiled), Kt2 = tc_conv_packed_k(C2, 1, 1, 1, kModeTiled);
  __nv_bfloat16 *w1 = nullptr, *w2 = nullptr;
  if ((rc = dev_alloc(p, &w1, static_cast<size_t>(K_pad) * Kt1)) != TLXCV_OK) return rc;
  if ((rc = dev_alloc(p, &w2, static_cast<size_t>(K_pad) * Kt2)) != TLXCV_OK) return rc;
  TLX_CUDA(ctx, pack_conv_weights(a.filters, w1, K, K_pad, K1, 1, 1, 1, kModeTiled, Kt1, st));
  TLX_CUDA(ctx, pack_conv_weights(d.filters, w2, K, K_pad, C2, 1, 1, 1, kModeTiled, Kt2, st));
  op.weights = w1;
  std::string err = tc_conv_prepare_dual(op.tc, ctx->sm_count, reinterpret_cast<const __nv_bfloat16*>(p->arena + x2.offset),
                                         static_cast<int>(M), K1, w1, reinterpret_cast<const __nv_bfloat16*>(p->arena + xin.offset),
                                         N, H2, W2, C2, d.stride, w2, K, p->arena + out.offset);
  if (!err.empty()) return fail(ctx, TLXCV_ERR_UNSUPPORTED, "%s", err.c_str());
  ConvKernelParams& kp = op.tc.p;
  kp.scale = op.scale, kp.shift = op.shift, kp.scale2 = op.scale2, kp.shift2 = op.shift2;
  kp.act1 = d.act2, kp.alpha1 = d.alpha2, kp.act2 = TLXCV_ACT_NONE, kp.alpha2 = 0.0f;
  kp.out_f32 = 0;
  op.impl = kImplTcConv;
  const double flops = 2.0 * M * K * (K1 + C2);
  const double bytes = (static_cast<double>(M) * (K1 + C2 + K) + static_cast<double>(K) * (K1 + C2)) * 2;
  set_info(op, "conv_tcgen05_dual_n128", 1, flops / bytes > 248.0 ? 1 : 0, flops, bytes, op.tc.grid, op.tc.threads, op.tc.smem,
           op.tc.block_n);
  return TLXCV_OK;
}

// B = a 1x1 conv (+ BN, + residual, + ReLU), A = the conv (+ BN + ReLU) whose N1 = 64 / 128-channel output is B's only input:
// ONE kernel, the intermediate stays in shared memory (conv_chain_kernel)
int compile_conv_chain(tlxcv_plan* p, OpRt& op, cudaStream_t st) {
  tlxcv_ctx* ctx = p->ctx;
  const tlxcv_op_desc& d = op.d;
  const tlxcv_op_desc& a = p->ops[op.chain_a].d;
  const TensorRt& xin = p->tensors[a.in0];
  const TensorRt& mid = p->tensors[op.chain_mid];
  const TensorRt& out = p->tensors[d.out];
  const int N = xin.d.n, H = xin.d.h, W = xin.d.w, C = xin.d.c, N1 = mid.d.c, N2 = out.d.c;
  if (!d.filters || !a.filters) return fail(ctx, TLXCV_ERR_INVALID, "op conv: filters pointer is NULL");
  const int N2_pad = static_cast<int>(align_up(N2, 256));
  int rc;
  if ((rc = dev_alloc(p, &op.scale, N2_pad)) != TLXCV_OK) return rc;
  if ((rc = dev_alloc(p, &op.shift, N2_pad)) != TLXCV_OK) return rc;
  if ((rc = dev_alloc(p, &op.scale2, 256)) != TLXCV_OK) return rc;
  if ((rc = dev_alloc(p, &op.shift2, 256)) != TLXCV_OK) return rc;
  TLX_CUDA(ctx, fold_bn(op.scale, op.shift, d.bn_gamma, d.bn_beta, d.bn_mean, d.bn_var, d.bias, d.bn_eps, N2, N2_pad, st));
  TLX_CUDA(ctx, fold_bn(op.scale2, op.shift2, a.bn_gamma, a.bn_beta, a.bn_mean, a.bn_var, a.bias, a.bn_eps, N1, 256, st));
  const int K1 = tc_conv_packed_k(C, a.r, a.s, 1, kModeIm2col);
  __nv_bfloat16 *w1 = nullptr, *w2 = nullptr;
  if ((rc = dev_alloc(p, &w1, static_cast<size_t>(256) * K1)) != TLXCV_OK) return rc;
  if ((rc = dev_alloc(p, &w2, static_cast<size_t>(N2_pad) * N1)) != TLXCV_OK) return rc;
  TLX_CUDA(ctx, pack_conv_weights(a.filters, w1, N1, 256, C, a.r, a.s, 1, kModeIm2col, K1, st));
  TLX_CUDA(ctx, pack_conv_weights(d.filters, w2, N2, N2_pad, N1, 1, 1, 1, kModeTiled, N1, st));
  op.weights = w1;
  const void* res = d.in1 >= 0 ? p->arena + p->tensors[d.in1].offset : nullptr;
  std::string err = tc_chain_prepare(op.tc, ctx->sm_count, reinterpret_cast<const __nv_bfloat16*>(p->arena + xin.offset), N, H, W, C, w1,
                                     K1, N1, a.r, a.s, a.stride, a.pad, a.dil, w2, N2, p->arena + out.offset, res);
  if (!err.empty()) return fail(ctx, TLXCV_ERR_UNSUPPORTED, "%s", err.c_str());
  ConvKernelParams& kp = op.tc.p;
  kp.scale = op.scale, kp.shift = op.shift, kp.scale2 = op.scale2, kp.shift2 = op.shift2;
  kp.act1 = d.act1, kp.alpha1 = d.alpha1, kp.act2 = d.act2, kp.alpha2 = d.alpha2;
  kp.out_f32 = 0;
  op.impl = kImplTcConv;
  const double M = static_cast<double>(out.d.n) * out.d.h * out.d.w;
  const double flops = 2.0 * M * (static_cast<double>(N1) * C * a.r * a.s + static_cast<double>(N2) * N1);
  const double th = std::min<double>(H, (out.d.h - 1.0) * a.stride + (a.r - 1) * a.dil + 1);
  const double tw = std::min<double>(W, (out.d.w - 1.0) * a.stride + (a.s - 1) * a.dil + 1);
  const double bytes = (static_cast<double>(N) * th * tw * C + static_cast<double>(N1) * C * a.r * a.s + static_cast<double>(N2) * N1 +
                        M * N2 * (d.in1 >= 0 ? 2 : 1)) * 2;
  char name[48];
  snprintf(name, sizeof name, "conv_chain_%dx%d_n%d_to_1x1", a.r, a.s, N1);
  set_info(op, name, 1, flops / bytes > 248.0 ? 1 : 0, flops, bytes, op.tc.grid, op.tc.threads, op.tc.smem, 128);
  return TLXCV_OK;
}

int compile_conv(tlxcv_plan* p, OpRt& op, cudaStream_t st, bool is_linear) {
  tlxcv_ctx* ctx = p->ctx;
  const tlxcv_op_desc& d = op.d;
  const TensorRt& in = p->tensors[d.in0];
  const TensorRt& out = p->tensors[d.out];
  const int N = in.d.n, H = in.d.h, W = in.d.w, C = in.d.c, K = out.d.c;
  const int R = is_linear ? 1 : d.r, S = is_linear ? 1 : d.s;
  const int stride = is_linear ? 1 : d.stride, pad = is_linear ? 0 : d.pad, dil = is_linear ? 1 : d.dil;
  const int groups = is_linear ? 1 : d.groups;
  if (!d.filters) return fail(ctx, TLXCV_ERR_INVALID, "op %s: filters pointer is NULL", is_linear ? "linear" : "conv");
  if (groups < 1 || C % groups || K % groups) return fail(ctx, TLXCV_ERR_INVALID, "conv: channels not divisible by groups");
  const int P = (H + 2 * pad - dil * (R - 1) - 1) / stride + 1, Q = (W + 2 * pad - dil * (S - 1) - 1) / stride + 1;
  if (P != out.d.h || Q != out.d.w || out.d.n != N)
    return fail(ctx, TLXCV_ERR_INVALID, "conv: output tensor is %dx%dx%d but geometry gives %dx%dx%d", out.d.n, out.d.h,
                out.d.w, N, P, Q);
  if (d.in1 >= 0) {
    const TensorRt& r = p->tensors[d.in1];
    if (r.d.n != out.d.n || r.d.h != out.d.h || r.d.w != out.d.w || r.d.c != out.d.c)
      return fail(ctx, TLXCV_ERR_INVALID, "conv: residual shape mismatch");
  }
  const int Cg = C / groups;
  const int K_pad = static_cast<int>(align_up(K, 256));
  int rc;
  if ((rc = dev_alloc(p, &op.scale, K_pad)) != TLXCV_OK) return rc;
  if ((rc = dev_alloc(p, &op.shift, K_pad)) != TLXCV_OK) return rc;
  TLX_CUDA(ctx, fold_bn(op.scale, op.shift, d.bn_gamma, d.bn_beta, d.bn_mean, d.bn_var, d.bias, d.bn_eps, K, K_pad, st));

  const double M = static_cast<double>(N) * P * Q;
  const double flops = 2.0 * M * K * Cg * R * S;
  const double wbytes = static_cast<double>(K) * Cg * R * S * (p->f32 ? 4 : 2);
  const double obytes = M * K * (out.d.dtype == TLXCV_F32 ? 4 : p->esize);
  // input pixels a filter actually touches (a 1x1 stride-2 conv reads a quarter of its input)
  const double th = R >= stride ? std::min<double>(H, (P - 1.0) * stride + (R - 1) * dil + 1) : static_cast<double>(P) * R;
  const double tw = S >= stride ? std::min<double>(W, (Q - 1.0) * stride + (S - 1) * dil + 1) : static_cast<double>(Q) * S;
  const double bytes = static_cast<double>(N) * th * tw * C * p->esize + wbytes + obytes + (d.in1 >= 0 ? M * K * p->esize : 0);

  if (p->f32) {
    float* w = nullptr;
    if (is_linear) {
      // (in,out) == [R=1][S=1][F][K] already: plain copy so that the plan owns its parameters
      if ((rc = dev_alloc(p, &w, static_cast<size_t>(K) * C)) != TLXCV_OK) return rc;
      TLX_CUDA(ctx, cudaMemcpyAsync(w, d.filters, static_cast<size_t>(K) * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
      op.weights = w;
    } else {
      if ((rc = dev_alloc(p, &w, static_cast<size_t>(K) * Cg * R * S)) != TLXCV_OK) return rc;
      TLX_CUDA(ctx, pack_conv_weights_f32(d.filters, w, K, Cg, R, S, st));
      op.weights = w;
    }
    op.impl = kImplDirectF32;
    set_info(op, "conv_direct_f32", 1, 1, flops, bytes, 0, 256, 0, 0);
    return TLXCV_OK;
  }
  // depthwise layers the slab kernel covers (64-channel blocks, stride 1, wide maps) run on the tensor cores with
  // diagonal weight taps: the CUDA-core kernel is bound by its instruction count (~40 per output)
  if (out.cs != K && d.in1 >= 0) return fail(ctx, TLXCV_ERR_UNSUPPORTED, "conv: a residual add needs a multiple of 8 output channels");
  const bool dw_on_slab = !is_linear && groups == C && K == C && groups > 1 && in.cs == C && d.in1 < 0 && !tuning_env("TLXCV_NO_DW_SLAB") &&
                          conv3x3_slab_supported(C, K, H, W, R, S, stride, pad, dil, groups);
  const bool depthwise = !is_linear && groups == C && K == C && groups > 1 && !dw_on_slab;
  if (depthwise) {
    if (dil != 1) return fail(ctx, TLXCV_ERR_UNSUPPORTED, "depthwise conv: dilation must be 1");
    __nv_bfloat16* w = nullptr;
    if ((rc = dev_alloc(p, &w, static_cast<size_t>(C) * R * S)) != TLXCV_OK) return rc;
    TLX_CUDA(ctx, pack_dw_weights(d.filters, w, C, R * S, 0, st));
    op.weights = w;
    op.impl = kImplDwConv;
    set_info(op, "dwconv_nhwc_bf16", 1, 0, flops, bytes, 0, 256, 0, 0);
    return TLXCV_OK;
  }
  // tensor-core path
  const int mode = tc_conv_mode(C, R, S, stride, pad, groups);
  const int layout = is_linear ? kLayoutPlain : tc_conv_layout(C, K, R, S, stride, pad, dil, groups, H, W);
  const int Ktot = tc_conv_packed_k(C, R, S, groups, mode, layout);
  __nv_bfloat16* w = nullptr;
  if ((rc = dev_alloc(p, &w, static_cast<size_t>(K_pad) * Ktot)) != TLXCV_OK) return rc;
  if (is_linear)
    TLX_CUDA(ctx, pack_linear_weights(d.filters, w, C, K, K_pad, st));
  else
    TLX_CUDA(ctx, pack_conv_weights(d.filters, w, K, K_pad, C, R, S, groups, layout == kLayoutPixelPairs ? kModePixelPairs : mode, Ktot, st));
  if (is_linear && Ktot != C)
    return fail(ctx, TLXCV_ERR_UNSUPPORTED, "linear: in_features (%d) must be a multiple of 64", C);
  op.weights = w;
  const __nv_bfloat16* act_in = reinterpret_cast<const __nv_bfloat16*>(p->arena + in.offset);
  int force_bn = 0;
  if (const char* e = tuning_env("TLXCV_FORCE_BLOCK_N")) force_bn = atoi(e);
  void* out_bf16 = nullptr;
  if (out.d.dtype != TLXCV_F32) {
    if (out.d.role != TLXCV_ROLE_INTERNAL) return fail(ctx, TLXCV_ERR_UNSUPPORTED, "conv: bf16 output must be an internal tensor");
    out_bf16 = p->arena + out.offset;
  }
  const void* res_bf16 = nullptr;
  if (d.in1 >= 0) {
    const TensorRt& r = p->tensors[d.in1];
    if (r.d.role != TLXCV_ROLE_INTERNAL || r.d.dtype != TLXCV_ACT)
      return fail(ctx, TLXCV_ERR_UNSUPPORTED, "conv: the residual must be an internal activation tensor");
    res_bf16 = p->arena + r.offset;
  }
  if (d.act2 != TLXCV_ACT_NONE && d.act2 != TLXCV_ACT_RELU)
    return fail(ctx, TLXCV_ERR_UNSUPPORTED, "conv: only ReLU (or nothing) may follow the residual add on the tensor-core path");
  if (!is_linear && out_bf16 && in.cs == C && out.cs == K && conv3x3_slab_supported(C, K, H, W, R, S, stride, pad, dil, groups)) {
    // 3x3 stride-1 layer with 64-channel work items: slab kernel (each input row fetched once, stationary weights)
    __nv_bfloat16* ws = w;
    int Kts = Ktot;
    if (groups == 1) {  // dense: K = 9 taps x exactly C channels
      Kts = 9 * C;
      if ((rc = dev_alloc(p, &ws, static_cast<size_t>(K_pad) * Kts)) != TLXCV_OK) return rc;
      TLX_CUDA(ctx, pack_conv_weights(d.filters, ws, K, K_pad, C, R, S, 1, kModeSlabDense, Kts, st));
    }
    std::string serr = conv3x3_slab_prepare(op.slab, ctx->sm_count, act_in, N, H, W, C, K, groups, ws, Kts, out_bf16, res_bf16);
    if (serr.empty()) {
      SlabParams& sp = op.slab.p;
      sp.scale = op.scale, sp.shift = op.shift, sp.act = d.act1, sp.alpha = d.alpha1, sp.act2 = d.act2, sp.alpha2 = d.alpha2;
      op.impl = kImplSlab;
      set_info(op, groups > 1 ? "conv3x3_slab_grouped" : (C == 32 ? "conv3x3_slab_c32" : "conv3x3_slab_n64"), 1,
               flops / bytes > 248.0 ? 1 : 0, flops, bytes, op.slab.grid, op.slab.threads, op.slab.smem, op.slab.block_n);
      return TLXCV_OK;
    }
  }
  if (d.act2 != TLXCV_ACT_NONE && d.in1 < 0)
    return fail(ctx, TLXCV_ERR_UNSUPPORTED, "conv: a second activation without a residual add");
  // the tcgen05 epilogue evaluates LeakyReLU as max(v, alpha * v)
  if ((d.act1 == TLXCV_ACT_LEAKY && !(d.alpha1 >= 0.0f && d.alpha1 <= 1.0f)) ||
      (d.act2 == TLXCV_ACT_LEAKY && !(d.alpha2 >= 0.0f && d.alpha2 <= 1.0f)))
    return fail(ctx, TLXCV_ERR_UNSUPPORTED, "conv: LeakyReLU slope outside [0, 1]");
  std::string err = tc_conv_prepare(op.tc, ctx->sm_count, act_in, N, H, W, C, in.cs, w, Ktot, K, R, S, stride, pad, dil,
                                    groups, force_bn, out_bf16, out.cs, res_bf16);
  if (!err.empty()) return fail(ctx, TLXCV_ERR_UNSUPPORTED, "%s", err.c_str());
  ConvKernelParams& kp = op.tc.p;
  kp.scale = op.scale, kp.shift = op.shift;
  kp.act1 = d.act1, kp.alpha1 = d.alpha1, kp.act2 = d.act2, kp.alpha2 = d.alpha2;
  kp.out_f32 = out.d.dtype == TLXCV_F32;
  op.impl = kImplTcConv;
  char name[48];
  static const char* mnames[] = {"tiled", "im2col", "gatherc4"};
  snprintf(name, sizeof name, "conv_tcgen05_%s_n%d%s%s%s", mnames[op.tc.mode], op.tc.block_n, groups > 1 ? "_grouped" : "",
           op.tc.two ? "_2sm" : "", op.tc.kblock == 32 ? "_kb32" : (op.tc.p.pair_taps ? "_pixelpairs" : ""));
  // tensor-bound when arithmetic intensity exceeds the ridge (~248 FLOP/B on the measured peaks)
  set_info(op, name, 1, flops / bytes > 248.0 ? 1 : 0, flops, bytes, op.tc.grid, op.tc.threads, op.tc.smem, op.tc.block_n);
  return TLXCV_OK;
}

int launch_op(tlxcv_plan* p, OpRt& op, const void* const* inputs, void* const* outputs, cudaStream_t st) {
  tlxcv_ctx* ctx = p->ctx;
  const tlxcv_op_desc& d = op.d;
  if (op.nop) return TLXCV_OK;
  const TensorRt& in = p->tensors[d.in0];
  const TensorRt& out = p->tensors[d.out];
  void* pin = tensor_ptr(p, d.in0, inputs, outputs);
  void* pout = tensor_ptr(p, d.out, inputs, outputs);
  void* pres = d.in1 >= 0 ? tensor_ptr(p, d.in1, inputs, outputs) : nullptr;
  if (!pin || !pout) return fail(ctx, TLXCV_ERR_INVALID, "NULL external tensor pointer");
  const int is_f32 = p->f32 ? 1 : 0;
  switch (op.impl) {
    case kImplNop:
      break;
    case kImplImport:
      if (out.wp > 0)
        TLX_CUDA(ctx, import_nchw_c4_padded(static_cast<const float*>(pin), pout, in.d.n, in.d.c, in.d.h, in.d.w, out.wp, out.pad_l, st));
      else
        TLX_CUDA(ctx, import_nchw(static_cast<const float*>(pin), pout, in.d.n, in.d.c, in.d.h, in.d.w, out.cs, is_f32, st));
      break;
    case kImplImportU8:
      if (op.resize_tx != nullptr) {
        TLX_CUDA(ctx, import_u8_resize(static_cast<const uint8_t*>(pin), pout, d.bn_mean, d.bn_var, op.resize_tx, op.resize_ty, in.d.n,
                                       in.d.c, in.d.h, in.d.w, out.d.h, out.d.w, out.wp > 0 ? out.wp : out.d.w, out.pad_l, is_f32, st));
        break;
      }
      TLX_CUDA(ctx, import_u8_nhwc(static_cast<const uint8_t*>(pin), pout, d.bn_mean, d.bn_var, in.d.n, in.d.c, in.d.h, in.d.w,
                                   out.wp > 0 ? out.wp : out.d.w, out.pad_l, is_f32, st));
      break;
    case kImplStem:
      TLX_CUDA(ctx, stem_rowring_launch(op.stem, st));
      break;
    case kImplSlab:
      TLX_CUDA(ctx, conv3x3_slab_launch(op.slab, st));
      break;
    case kImplExport:
      TLX_CUDA(ctx, export_nchw(pin, static_cast<float*>(pout), in.d.n, in.d.c, in.cs, in.d.h, in.d.w, is_f32, st));
      break;
    case kImplTcConv: {
      TcConvLaunch L = op.tc;
      L.p.out = op.skip_logits ? nullptr : pout;  // only read by fp32-output (logits) launches; bf16 tensors go through the baked TMA maps
      if (op.fused_argmax >= 0) {
        L.p.amax_keys = op.amax_keys, L.p.amax_ticket = op.ticket;
        L.p.amax_out = static_cast<long long*>(tensor_ptr(p, p->ops[op.fused_argmax].d.out, inputs, outputs));
        if (!L.p.amax_out) return fail(ctx, TLXCV_ERR_INVALID, "NULL external tensor pointer");
      }
      TLX_CUDA(ctx, tc_conv_launch(L, st));
      break;
    }
    case kImplDwConv:
      TLX_CUDA(ctx, dwconv_nhwc(pin, op.weights, pout, op.scale, op.shift, pres, in.d.n, in.d.h, in.d.w, in.d.c, out.d.h,
                                out.d.w, d.r, d.s, d.stride, d.pad, d.act1, d.alpha1, d.act2, d.alpha2, is_f32, st));
      break;
    case kImplDirectF32: {
      const bool lin = d.kind == TLXCV_OP_LINEAR;
      TLX_CUDA(ctx, conv_direct_f32(static_cast<const float*>(pin), static_cast<const float*>(op.weights),
                                    static_cast<float*>(pout), op.scale, op.shift, static_cast<const float*>(pres),
                                    in.d.n, in.d.h, in.d.w, in.d.c, out.d.h, out.d.w, out.d.c, lin ? 1 : d.r, lin ? 1 : d.s,
                                    lin ? 1 : d.stride, lin ? 0 : d.pad, lin ? 1 : d.dil, lin ? 1 : d.groups, d.act1,
                                    d.alpha1, d.act2, d.alpha2, st));
      break;
    }
    case kImplMaxpool:
      TLX_CUDA(ctx, maxpool_nhwc(pin, pout, in.d.n, in.d.h, in.d.w, in.d.c, out.d.h, out.d.w, d.r, d.stride, d.pad, is_f32, st));
      break;
    case kImplAvgpool:
      TLX_CUDA(ctx, avgpool_nhwc(pin, pout, in.d.n, in.d.h, in.d.w, in.d.c, out.d.h, out.d.w, d.r, d.stride, d.pad, is_f32, st));
      break;
    case kImplSplatApply:
      TLX_CUDA(ctx, splat_apply(pin, pres, pout, in.d.n, in.d.h * in.d.w, out.d.c, d.r, d.groups, is_f32, st));
      break;
    case kImplGap:
      TLX_CUDA(ctx, gap_nhwc(pin, pout, in.d.n, in.d.h * in.d.w, in.d.c, is_f32, st));
      break;
    case kImplAddAct:
      TLX_CUDA(ctx, add_act(pin, pres, pout, static_cast<size_t>(in.d.n) * in.d.h * in.d.w * in.d.c, d.act2, d.alpha2, is_f32, st));
      break;
    case kImplUpsampleConcat:
      TLX_CUDA(ctx, upsample_concat(pin, pres, pout, out.d.n, out.d.h, out.d.w, in.d.c, d.in1 >= 0 ? p->tensors[d.in1].d.c : 0, d.r,
                                    d.s, is_f32, st));
      break;
    case kImplSoftmax:
      TLX_CUDA(ctx, softmax_rows(static_cast<const float*>(pin), static_cast<float*>(pout), in.d.n, in.d.c, st));
      break;
    case kImplSoftmaxCe:
      TLX_CUDA(ctx, softmax_ce(static_cast<const float*>(pin), static_cast<const long long*>(pres), op.row_loss, op.ticket,
                               static_cast<float*>(pout), in.d.n, in.d.c, st));
      break;
    case kImplArgmax:
      TLX_CUDA(ctx, argmax_rows(static_cast<const float*>(pin), static_cast<long long*>(pout), in.d.n, in.d.c, st));
      break;
    default:
      return fail(ctx, TLXCV_ERR_INVALID, "unknown op implementation");
  }
  return TLXCV_OK;
}

int launch_range(tlxcv_plan* p, int begin, int end, const void* const* inputs, void* const* outputs, cudaStream_t st) {
  static const bool sync_each = tuning_env("TLXCV_SYNC_EACH_OP") != nullptr;  // debugging: attribute faults to an op
  for (int i = begin; i < end; ++i) {
    OpRt& op = p->ops[i];
    int rc = launch_op(p, op, inputs, outputs, st);
    if (rc != TLXCV_OK) return rc;
    if (sync_each) {
      cudaError_t e = cudaStreamSynchronize(st);
      if (e != cudaSuccess)
        return fail(p->ctx, TLXCV_ERR_CUDA, "op %d (%s) faulted: %s", i, op.info.kernel, cudaGetErrorString(e));
    }
  }
  return TLXCV_OK;
}

int launch_all(tlxcv_plan* p, const void* const* inputs, void* const* outputs, cudaStream_t st) {
  return launch_range(p, 0, static_cast<int>(p->ops.size()), inputs, outputs, st);
}

// Captures ops [begin, end) into an executable graph (on a private stream; the caller's stream is not touched).
int capture_range(tlxcv_plan* p, int begin, int end, const void* const* inputs, void* const* outputs, cudaGraphExec_t* exec) {
  tlxcv_ctx* ctx = p->ctx;
  cudaStream_t cap;
  TLX_CUDA(ctx, cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking));
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
  int rc = TLXCV_OK;
  if (e == cudaSuccess) {
    rc = launch_range(p, begin, end, inputs, outputs, cap);
    e = cudaStreamEndCapture(cap, &graph);
  }
  cudaStreamDestroy(cap);
  if (rc != TLXCV_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return fail(ctx, TLXCV_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail(ctx, TLXCV_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
  return TLXCV_OK;
}

bool touches_external(const tlxcv_plan* p, const OpRt& op) {
  if (op.nop) return false;
  const int extra_out = op.fused_argmax >= 0 ? p->ops[op.fused_argmax].d.out : -1;
  for (int t : {op.d.in0, op.d.in1, op.d.out, extra_out})
    if (t >= 0 && p->tensors[t].d.role != TLXCV_ROLE_INTERNAL) return true;
  return false;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int tlxcv_abi_version(void) { return TLXCV_ABI_VERSION; }

int tlxcv_create(int device, tlxcv_ctx** out) {
  if (!out) return fail(nullptr, TLXCV_ERR_INVALID, "tlxcv_create: out is NULL");
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
    return fail(nullptr, TLXCV_ERR_NO_DEVICE, "no CUDA device: tlxcv_b200 has no CPU fallback");
  if (device < 0 || device >= count) return fail(nullptr, TLXCV_ERR_INVALID, "device %d out of range (0..%d)", device, count - 1);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, TLXCV_ERR_CUDA, "cudaGetDeviceProperties failed");
  if (prop.major != 10)
    return fail(nullptr, TLXCV_ERR_NO_DEVICE, "device %d is sm_%d%d; this library contains sm_100a (Blackwell B200) code only",
                device, prop.major, prop.minor);
  DeviceGuard guard(device);
  cudaError_t e = tc_conv_set_attributes();
  if (e == cudaSuccess) e = stem_rowring_set_attributes();
  if (e == cudaSuccess) e = conv3x3_slab_set_attributes();
  if (e != cudaSuccess) return fail(nullptr, TLXCV_ERR_CUDA, "cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
  tlxcv_ctx* c = new tlxcv_ctx();
  c->device = device;
  c->sm_count = prop.multiProcessorCount;
  *out = c;
  return TLXCV_OK;
}

int tlxcv_destroy(tlxcv_ctx* ctx) {
  delete ctx;
  return TLXCV_OK;
}

const char* tlxcv_last_error(const tlxcv_ctx* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int tlxcv_device_sm_count(const tlxcv_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

int tlxcv_plan_build(tlxcv_ctx* ctx, const tlxcv_tensor_desc* tensors, int n_tensors, const tlxcv_op_desc* ops, int n_ops,
                     int precision, void* stream, tlxcv_plan** out) {
  if (!ctx) return TLXCV_ERR_INVALID;
  if (!tensors || !ops || !out || n_tensors <= 0 || n_ops <= 0) return fail(ctx, TLXCV_ERR_INVALID, "plan_build: bad arguments");
  if (precision != TLXCV_PREC_BF16 && precision != TLXCV_PREC_F32_VALIDATE)
    return fail(ctx, TLXCV_ERR_INVALID, "plan_build: unknown precision %d", precision);
  *out = nullptr;
  DeviceGuard dev_guard(ctx->device);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  tlxcv_plan* p = new tlxcv_plan();
  p->ctx = ctx;
  p->precision = precision;
  p->f32 = precision == TLXCV_PREC_F32_VALIDATE;
  p->esize = p->f32 ? 4 : 2;
  struct Guard {
    tlxcv_plan* p;
    ~Guard() {
      if (p) tlxcv_plan_destroy(p);
    }
  } guard{p};

  // ---- tensors ----
  p->tensors.resize(n_tensors);
  for (int i = 0; i < n_tensors; ++i) {
    TensorRt& T = p->tensors[i];
    T.d = tensors[i];
    if (T.d.n <= 0 || T.d.h <= 0 || T.d.w <= 0 || T.d.c <= 0) return fail(ctx, TLXCV_ERR_INVALID, "tensor %d: non-positive dimension", i);
    T.cs = T.d.c;
    if (T.d.dtype == TLXCV_ACT && T.d.role == TLXCV_ROLE_INTERNAL) {
      if (T.d.c <= 4)
        T.cs = 4;
      else if (T.d.c % 8 && !p->f32)
        T.cs = static_cast<int>(align_up(T.d.c, 8));  // rows padded to 16 bytes: head convs (YOLOv3 outputs: 3 x (classes + 5)
                                                      // channels); checked below: written by a conv, read by the export only
    }
    T.bytes = static_cast<size_t>(T.d.n) * T.d.h * T.d.w * T.cs * dtype_size(p, T.d.dtype);
    if (T.d.role == TLXCV_ROLE_INPUT) {
      T.ext_index = static_cast<int>(p->input_ids.size());
      p->input_ids.push_back(i);
    } else if (T.d.role == TLXCV_ROLE_OUTPUT) {
      T.ext_index = static_cast<int>(p->output_ids.size());
      p->output_ids.push_back(i);
    }
  }
  // ---- ops: validate indices, liveness ----
  p->ops.resize(n_ops);
  for (int i = 0; i < n_ops; ++i) {
    p->ops[i].d = ops[i];
    const tlxcv_op_desc& d = ops[i];
    if (d.in0 < 0 || d.in0 >= n_tensors || d.out < 0 || d.out >= n_tensors || d.in1 >= n_tensors)
      return fail(ctx, TLXCV_ERR_INVALID, "op %d: tensor index out of range", i);
  }
  for (int i = 0; i < n_ops; ++i) {
    const tlxcv_op_desc& d = ops[i];
    for (int t : {d.in0, d.in1}) {
      if (t < 0) continue;
      const TensorRt& T = p->tensors[t];
      if (T.d.role == TLXCV_ROLE_INTERNAL && T.d.dtype == TLXCV_ACT && T.d.c > 4 && T.d.c % 8 && d.kind != TLXCV_OP_EXPORT_NCHW)
        return fail(ctx, TLXCV_ERR_UNSUPPORTED,
                    "op %d reads tensor %d with %d channels: activations that feed another layer need a multiple of 8 channels "
                    "(other widths are only supported for maps that are returned)", i, t, T.d.c);
    }
    const TensorRt& O = p->tensors[d.out];
    if (O.d.role == TLXCV_ROLE_INTERNAL && O.d.dtype == TLXCV_ACT && O.d.c > 4 && O.d.c % 8 && d.kind != TLXCV_OP_CONV)
      return fail(ctx, TLXCV_ERR_UNSUPPORTED, "op %d: only a convolution may produce a map with %d channels (not a multiple of 8)", i, O.d.c);
  }
  // ---- pre-pass: C_in <= 4 stems go to the row-ring kernel (padded-column input written by the import
  //      op), and a 3x3/s2/p1 max-pool that is the stem's only consumer is fused into its epilogue ----
  if (!p->f32 && !tuning_env("TLXCV_NO_ROWRING")) {
    for (int i = 0; i < n_ops; ++i) {
      OpRt& op = p->ops[i];
      const tlxcv_op_desc& d = op.d;
      if (d.kind != TLXCV_OP_CONV || d.in1 >= 0 || d.act2 != TLXCV_ACT_NONE) continue;
      TensorRt& in = p->tensors[d.in0];
      if (in.d.c > 4 || in.d.role != TLXCV_ROLE_INTERNAL || in.d.dtype != TLXCV_ACT) continue;
      if (p->tensors[d.out].d.role != TLXCV_ROLE_INTERNAL || p->tensors[d.out].d.dtype != TLXCV_ACT) continue;
      int producer = -1, consumers = 0;
      for (int k = 0; k < n_ops; ++k) {
        if (p->ops[k].d.out == d.in0) producer = k;
        if (p->ops[k].d.in0 == d.in0 || p->ops[k].d.in1 == d.in0) ++consumers;
      }
      if (producer < 0 || consumers != 1) continue;
      if (p->ops[producer].d.kind != TLXCV_OP_IMPORT_NCHW && p->ops[producer].d.kind != TLXCV_OP_IMPORT_U8_NHWC) continue;
      StemGeometry g;
      if (!stem_rowring_geometry(g, in.d.c, p->tensors[d.out].d.c, in.d.h, in.d.w, d.r, d.s, d.stride, d.pad, d.dil, d.groups))
        continue;
      op.use_stem = true;
      op.geo = g;
      in.wp = g.Wp, in.pad_l = g.pad_l, in.cs = 4;
      in.bytes = static_cast<size_t>(in.d.n) * in.d.h * g.Wp * 4 * 2;
      if (g.pairs || g.Qw > 128 || tuning_env("TLXCV_NO_POOL_FUSION")) continue;
      int pool = -1, users = 0;
      for (int k = 0; k < n_ops; ++k)
        if (p->ops[k].d.in0 == d.out || p->ops[k].d.in1 == d.out) ++users, pool = k;
      if (users != 1 || pool < i) continue;
      const tlxcv_op_desc& m = p->ops[pool].d;
      if (m.kind != TLXCV_OP_MAXPOOL || m.r != 3 || m.s != 3 || m.stride != 2 || m.pad != 1) continue;
      if (p->tensors[m.out].d.role != TLXCV_ROLE_INTERNAL) continue;
      p->tensors[d.out].elided = true;
      op.pool_op = pool;
      op.d.out = m.out;
      p->ops[pool].nop = true;
    }
  }
  // ---- pre-pass: `relu(bn(conv1x1(a)) + bn(conv1x1_stride(b)))` (last conv of a bottleneck + the block's
  //      downsample conv) becomes ONE dual-accumulator kernel: the residual map never reaches HBM ----
  if (!p->f32 && !tuning_env("TLXCV_NO_DUAL")) {
    for (int i = 0; i < n_ops; ++i) {
      OpRt& B = p->ops[i];
      const tlxcv_op_desc& d = B.d;
      if (B.nop || d.kind != TLXCV_OP_CONV || d.in1 < 0) continue;
      if (d.r != 1 || d.s != 1 || d.pad != 0 || d.dil != 1 || d.groups != 1 || (d.stride != 1 && d.stride != 2)) continue;
      if (d.act1 != TLXCV_ACT_NONE || (d.act2 != TLXCV_ACT_NONE && d.act2 != TLXCV_ACT_RELU)) continue;
      const TensorRt& T = p->tensors[d.in1];
      const TensorRt& xin = p->tensors[d.in0];
      const TensorRt& o = p->tensors[d.out];
      if (T.d.role != TLXCV_ROLE_INTERNAL || T.d.dtype != TLXCV_ACT || o.d.role != TLXCV_ROLE_INTERNAL || o.d.dtype != TLXCV_ACT) continue;
      if (xin.d.dtype != TLXCV_ACT || xin.d.c % 8 || o.d.c % 8 || xin.d.c <= 4) continue;
      int producer = -1, users = 0;
      for (int k = 0; k < n_ops; ++k) {
        if (p->ops[k].nop) continue;
        if (p->ops[k].d.out == d.in1) producer = k;
        if (p->ops[k].d.in0 == d.in1 || p->ops[k].d.in1 == d.in1) ++users;
      }
      if (producer < 0 || producer >= i || users != 1) continue;
      OpRt& A = p->ops[producer];
      const tlxcv_op_desc& a = A.d;
      if (a.kind != TLXCV_OP_CONV || a.in1 >= 0 || a.r != 1 || a.s != 1 || a.pad != 0 || a.stride != 1 || a.dil != 1 || a.groups != 1)
        continue;
      if (a.act1 != TLXCV_ACT_NONE || a.act2 != TLXCV_ACT_NONE) continue;
      const TensorRt& x2 = p->tensors[a.in0];
      if (x2.d.dtype != TLXCV_ACT || x2.d.c % 8 || x2.d.c <= 4) continue;
      // worth it only while the pair is bound by HBM traffic: the dual kernel works on 128-wide tiles, which costs
      // the MMA / L2-bound deep stages more than the saved residual round trip brings
      if ((x2.d.c + 63) / 64 + (xin.d.c + 63) / 64 > 8 && !tuning_env("TLXCV_FORCE_DUAL")) continue;
      // a must not be rewritten between A and B (it is read later now): its producer precedes A by construction
      p->tensors[d.in1].elided = true;
      A.nop = true;
      B.dual_a = producer;
      B.d.in1 = -1;
    }
  }
  // ---- pre-pass: conv (+BN+ReLU) -> 1x1 conv (+BN, +residual, +ReLU) whose 64 / 128-channel intermediate has no other reader
  //      (conv2 -> conv3 of a ResNet bottleneck, classification/resnet.py:146-155): ONE chain kernel, the intermediate
  //      never reaches HBM and the MMA-bound 3x3 runs under the 1x1's output traffic ----
  if (!p->f32 && !tuning_env("TLXCV_NO_CHAIN")) {
    const int max_n1 = tuning_env("TLXCV_CHAIN_MAX_N1") ? atoi(tuning_env("TLXCV_CHAIN_MAX_N1")) : 128;
    for (int i = 0; i < n_ops; ++i) {
      OpRt& B = p->ops[i];
      const tlxcv_op_desc& d = B.d;
      if (B.nop || B.dual_a >= 0 || d.kind != TLXCV_OP_CONV) continue;
      if (d.r != 1 || d.s != 1 || d.stride != 1 || d.pad != 0 || d.dil != 1 || d.groups != 1) continue;
      const bool res = d.in1 >= 0;
      if (res ? (d.act1 != TLXCV_ACT_NONE || (d.act2 != TLXCV_ACT_NONE && d.act2 != TLXCV_ACT_RELU))
              : ((d.act1 != TLXCV_ACT_NONE && d.act1 != TLXCV_ACT_RELU) || d.act2 != TLXCV_ACT_NONE))
        continue;
      const TensorRt& T = p->tensors[d.in0];
      const TensorRt& o = p->tensors[d.out];
      if (T.d.role != TLXCV_ROLE_INTERNAL || T.d.dtype != TLXCV_ACT || o.d.role != TLXCV_ROLE_INTERNAL || o.d.dtype != TLXCV_ACT) continue;
      if (d.in1 == d.in0 || o.cs != o.d.c || T.d.c > max_n1) continue;
      if (res && (p->tensors[d.in1].d.role != TLXCV_ROLE_INTERNAL || p->tensors[d.in1].d.dtype != TLXCV_ACT)) continue;
      int producer = -1, users = 0;
      for (int k = 0; k < n_ops; ++k) {
        if (p->ops[k].nop) continue;
        if (p->ops[k].d.out == d.in0) producer = k;
        if (p->ops[k].d.in0 == d.in0 || p->ops[k].d.in1 == d.in0) ++users;
      }
      if (producer < 0 || producer >= i || users != 1) continue;
      OpRt& A = p->ops[producer];
      const tlxcv_op_desc& a = A.d;
      if (A.use_stem || A.dual_a >= 0 || A.chain_a >= 0 || a.kind != TLXCV_OP_CONV || a.in1 >= 0 || a.groups != 1) continue;
      if (a.act1 != TLXCV_ACT_RELU || a.act2 != TLXCV_ACT_NONE || a.r > 3 || a.s > 3) continue;
      const TensorRt& x = p->tensors[a.in0];
      if (x.d.role != TLXCV_ROLE_INTERNAL || x.d.dtype != TLXCV_ACT || x.cs != x.d.c) continue;
      if (!tc_chain_supported(x.d.c, T.d.c, o.d.c)) continue;
      // the chain kernel's first GEMM loads its operands by im2col TMA (every pixel 9 times from L2 for a 3x3): worth it
      // when the 1x1 that follows is wide enough to be the longer half (the epilogue / HBM traffic hides the rest)
      if (o.d.c < 2 * T.d.c && !tuning_env("TLXCV_FORCE_CHAIN")) continue;
      p->tensors[d.in0].elided = true;
      A.nop = true;
      A.chain_first = true;
      B.chain_a = producer;
      B.chain_mid = d.in0;
      B.d.in0 = a.in0;  // the chain kernel reads the first conv's input
    }
  }
  // ---- pre-pass: `argmax(linear(x))` (ImageClassification.predict, tasks/image_classification.py:20-23): the argmax runs
  //      inside the Linear launch (per-row 64-bit atomicMax keys, decoded by the last CTA); logits nobody else reads are
  //      never written ----
  if (!p->f32 && !tuning_env("TLXCV_NO_ARGMAX_FUSION")) {
    for (int i = 0; i < n_ops; ++i) {
      OpRt& A = p->ops[i];
      if (A.nop || A.d.kind != TLXCV_OP_ARGMAX) continue;
      const TensorRt& L = p->tensors[A.d.in0];
      if (L.d.dtype != TLXCV_F32 || p->tensors[A.d.out].d.dtype != TLXCV_I64) continue;
      int producer = -1, users = 0;
      for (int k = 0; k < n_ops; ++k) {
        if (p->ops[k].nop) continue;
        if (p->ops[k].d.out == A.d.in0) producer = k;
        if (p->ops[k].d.in0 == A.d.in0 || p->ops[k].d.in1 == A.d.in0) ++users;
      }
      if (producer < 0 || producer >= i || p->ops[producer].d.kind != TLXCV_OP_LINEAR || p->ops[producer].fused_argmax >= 0) continue;
      p->ops[producer].fused_argmax = i;
      p->ops[producer].skip_logits = users == 1 && L.d.role == TLXCV_ROLE_INTERNAL;
      if (p->ops[producer].skip_logits) p->tensors[A.d.in0].elided = true;
      A.nop = true;
    }
  }
  for (int i = 0; i < n_ops; ++i) {
    OpRt& op = p->ops[i];
    const tlxcv_op_desc& d = op.d;
    if (op.nop) continue;
    const int extra_in = op.dual_a >= 0 ? p->ops[op.dual_a].d.in0 : -1;
    if (op.fused_argmax >= 0) {  // the fused argmax's result is written by this op
      TensorRt& ao = p->tensors[p->ops[op.fused_argmax].d.out];
      if (ao.first_def >= 0 || ao.d.role == TLXCV_ROLE_INPUT) return fail(ctx, TLXCV_ERR_INVALID, "op %d: tensor is written twice", i);
      ao.first_def = i;
      ao.last_use = std::max(ao.last_use, i);
    }
    for (int t : {d.in0, d.in1, extra_in}) {
      if (t < 0) continue;
      if (p->tensors[t].d.role != TLXCV_ROLE_INPUT && p->tensors[t].first_def < 0)
        return fail(ctx, TLXCV_ERR_INVALID, "op %d reads tensor %d before it is produced", i, t);
      p->tensors[t].last_use = i;
    }
    if (p->tensors[d.out].first_def >= 0 || p->tensors[d.out].d.role == TLXCV_ROLE_INPUT)
      return fail(ctx, TLXCV_ERR_INVALID, "op %d: tensor %d is written twice", i, d.out);
    p->tensors[d.out].first_def = i;
    p->tensors[d.out].last_use = std::max(p->tensors[d.out].last_use, i);
  }
  // ---- workspace: first-fit arena over tensor lifetimes ----
  {
    Arena arena;
    const bool no_reuse = tuning_env("TLXCV_NO_REUSE") != nullptr;  // debugging: keep every intermediate readable
    for (int i = 0; i < n_ops; ++i) {
      if (p->ops[i].nop) continue;
      TensorRt& o = p->tensors[p->ops[i].d.out];
      if (o.d.role == TLXCV_ROLE_INTERNAL) o.offset = arena.alloc(o.bytes);
      if (p->ops[i].fused_argmax >= 0) {
        TensorRt& ao = p->tensors[p->ops[p->ops[i].fused_argmax].d.out];
        if (ao.d.role == TLXCV_ROLE_INTERNAL) ao.offset = arena.alloc(ao.bytes);
      }
      const int extra_in = p->ops[i].dual_a >= 0 ? p->ops[p->ops[i].dual_a].d.in0 : -1;
      for (int t : {p->ops[i].d.in0, p->ops[i].d.in1, extra_in, p->ops[i].d.out}) {
        if (t < 0) continue;
        TensorRt& T = p->tensors[t];
        if (!no_reuse && T.d.role == TLXCV_ROLE_INTERNAL && T.last_use == i && T.first_def >= 0) {
          arena.release(T.offset, T.bytes);
          T.last_use = -2;  // released
        }
      }
    }
    p->arena_bytes = std::max<size_t>(arena.end, 1024);
    void* ptr = nullptr;
    cudaError_t e = cudaMalloc(&ptr, p->arena_bytes);
    if (e != cudaSuccess) return fail(ctx, TLXCV_ERR_OOM, "workspace cudaMalloc(%zu) failed: %s", p->arena_bytes, cudaGetErrorString(e));
    p->arena = static_cast<uint8_t*>(ptr);
  }
  // ---- compile ops ----
  for (int i = 0; i < n_ops; ++i) {
    OpRt& op = p->ops[i];
    const tlxcv_op_desc& d = op.d;
    const TensorRt& in = p->tensors[d.in0];
    const TensorRt& o = p->tensors[d.out];
    const double in_bytes = static_cast<double>(in.bytes), out_bytes = static_cast<double>(o.bytes);
    int rc = TLXCV_OK;
    if (op.nop) {
      op.impl = kImplNop;
      set_info(op, d.kind == TLXCV_OP_CONV ? (op.chain_first ? "(first GEMM of the chain conv)" : "(first GEMM of the dual conv)")
                                           : (d.kind == TLXCV_OP_ARGMAX ? "(fused into the linear launch)" : "(fused into the stem conv)"),
               0, 0, 0, 0, 0, 0, 0, 0);
      continue;
    }
    switch (d.kind) {
      case TLXCV_OP_IMPORT_NCHW:
        if (in.d.role != TLXCV_ROLE_INPUT || in.d.dtype != TLXCV_F32 || o.d.dtype != TLXCV_ACT || o.d.role != TLXCV_ROLE_INTERNAL)
          return fail(ctx, TLXCV_ERR_INVALID, "op %d: import expects external f32 -> internal activation", i);
        op.impl = kImplImport;
        set_info(op, o.wp > 0 ? "import_nchw_c4_padded" : (o.cs == 4 ? "import_nchw_c4" : "import_nchw_tile"), 1, 0, 0,
                 in_bytes + out_bytes, 0, 256, 0, 0);
        break;
      case TLXCV_OP_IMPORT_U8_NHWC:
        if (in.d.role != TLXCV_ROLE_INPUT || in.d.dtype != TLXCV_U8 || o.d.dtype != TLXCV_ACT || o.d.role != TLXCV_ROLE_INTERNAL ||
            in.d.c > 4 || in.d.c != o.d.c || o.cs != 4 || !d.bn_mean || !d.bn_var)
          return fail(ctx, TLXCV_ERR_INVALID, "op %d: import_u8 expects external uint8 NHWC (C <= 4) + mean/std -> internal activation", i);
        op.impl = kImplImportU8;
        if (in.d.n != o.d.n) return fail(ctx, TLXCV_ERR_INVALID, "op %d: import_u8 batch size mismatch", i);
        if (in.d.h != o.d.h || in.d.w != o.d.w) {  // Resize((H, W)) in front of Normalize: fused into the same pass
          if (in.d.h < 1 || in.d.w < 1) return fail(ctx, TLXCV_ERR_INVALID, "op %d: empty source image", i);
          const std::vector<int4> tx = resize_table(o.d.w, in.d.w, true), ty = resize_table(o.d.h, in.d.h, false);
          if ((rc = dev_alloc(p, &op.resize_tx, tx.size())) != TLXCV_OK) break;
          if ((rc = dev_alloc(p, &op.resize_ty, ty.size())) != TLXCV_OK) break;
          TLX_CUDA(ctx, cudaMemcpyAsync(op.resize_tx, tx.data(), tx.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
          TLX_CUDA(ctx, cudaMemcpyAsync(op.resize_ty, ty.data(), ty.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
          TLX_CUDA(ctx, cudaStreamSynchronize(st));  // the host tables go out of scope here
          set_info(op, o.wp > 0 ? "import_u8_resize_padded" : "import_u8_resize", 1, 0, 0, in_bytes + out_bytes, 0, 256, 0, 0);
          break;
        }
        set_info(op, o.wp > 0 ? "import_u8_nhwc_padded" : "import_u8_nhwc", 1, 0, 0, in_bytes + out_bytes, 0, 256, 0, 0);
        break;
      case TLXCV_OP_EXPORT_NCHW:
        if (o.d.role != TLXCV_ROLE_OUTPUT || o.d.dtype != TLXCV_F32 || in.d.dtype != TLXCV_ACT || in.wp > 0)
          return fail(ctx, TLXCV_ERR_INVALID, "op %d: export expects internal activation -> external f32", i);
        op.impl = kImplExport;
        set_info(op, "export_nchw_tile", 1, 0, 0, in_bytes + out_bytes, 0, 256, 0, 0);
        break;
      case TLXCV_OP_CONV:
        if (in.d.dtype != TLXCV_ACT || o.d.dtype != TLXCV_ACT) return fail(ctx, TLXCV_ERR_INVALID, "op %d: conv tensors must be activations", i);
        rc = op.use_stem ? compile_stem(p, op, st)
                         : (op.dual_a >= 0 ? compile_conv_dual(p, op, st)
                                           : (op.chain_a >= 0 ? compile_conv_chain(p, op, st) : compile_conv(p, op, st, false)));
        break;
      case TLXCV_OP_LINEAR:
        if (in.d.dtype != TLXCV_ACT || o.d.dtype != TLXCV_F32 || in.d.h != 1 || in.d.w != 1)
          return fail(ctx, TLXCV_ERR_INVALID, "op %d: linear expects (N, F) activation -> (N, K) f32", i);
        rc = compile_conv(p, op, st, true);
        if (rc == TLXCV_OK && op.fused_argmax >= 0) {
          if (op.impl != kImplTcConv || op.tc.two || op.tc.dual) return fail(ctx, TLXCV_ERR_INVALID, "op %d: argmax fusion needs the single-CTA tcgen05 Linear", i);
          if ((rc = dev_alloc(p, &op.amax_keys, static_cast<size_t>(in.d.n))) != TLXCV_OK) break;
          if ((rc = dev_alloc(p, &op.ticket, 1)) != TLXCV_OK) break;
          TLX_CUDA(ctx, cudaMemsetAsync(op.amax_keys, 0, std::max<size_t>(in.d.n * sizeof(unsigned long long), 256), st));
          TLX_CUDA(ctx, cudaMemsetAsync(op.ticket, 0, 256, st));
          snprintf(op.info.kernel + strlen(op.info.kernel), sizeof op.info.kernel - strlen(op.info.kernel), "+argmax");
        }
        break;
      case TLXCV_OP_SOFTMAX:
        if (in.d.dtype != TLXCV_F32 || o.d.dtype != TLXCV_F32 || o.d.n != in.d.n || o.d.c != in.d.c || in.d.h != 1 || in.d.w != 1)
          return fail(ctx, TLXCV_ERR_INVALID, "op %d: softmax expects (N, K) f32 -> (N, K) f32", i);
        op.impl = kImplSoftmax;
        set_info(op, "softmax_rows", 1, 0, 0, in_bytes + out_bytes, 0, 256, 0, 0);
        break;
      case TLXCV_OP_SOFTMAX_CE:
        if (d.in1 < 0 || in.d.dtype != TLXCV_F32 || p->tensors[d.in1].d.dtype != TLXCV_I64 || o.d.dtype != TLXCV_F32 ||
            p->tensors[d.in1].d.n != in.d.n || o.d.n * o.d.c != 1 || in.d.h != 1 || in.d.w != 1)
          return fail(ctx, TLXCV_ERR_INVALID, "op %d: softmax_ce expects (N, K) f32 logits + (N) i64 labels -> (1) f32", i);
        if ((rc = dev_alloc(p, &op.row_loss, static_cast<size_t>(in.d.n))) != TLXCV_OK) break;
        if ((rc = dev_alloc(p, &op.ticket, 1)) != TLXCV_OK) break;
        TLX_CUDA(ctx, cudaMemsetAsync(op.ticket, 0, 256, st));
        op.impl = kImplSoftmaxCe;
        set_info(op, "softmax_ce", 1, 0, 0, in_bytes + out_bytes, 0, 256, 0, 0);
        break;
      case TLXCV_OP_MAXPOOL:
        if (d.r != d.s) return fail(ctx, TLXCV_ERR_UNSUPPORTED, "op %d: non-square pooling window", i);
        if (o.d.h != (in.d.h + 2 * d.pad - d.r) / d.stride + 1 || o.d.w != (in.d.w + 2 * d.pad - d.r) / d.stride + 1 || o.d.c != in.d.c)
          return fail(ctx, TLXCV_ERR_INVALID, "op %d: maxpool output shape mismatch", i);
        op.impl = kImplMaxpool;
        set_info(op, "maxpool_nhwc", 1, 0, 0, in_bytes + out_bytes, 0, 256, 0, 0);
        break;
      case TLXCV_OP_AVGPOOL:
        if (d.r != d.s || d.pad < 0 || 2 * d.pad >= d.r + 1 || d.r < 1 || d.stride < 1)
          return fail(ctx, TLXCV_ERR_UNSUPPORTED, "op %d: average pooling needs a square window with padding below half of it", i);
        if (o.d.h != (in.d.h + 2 * d.pad - d.r) / d.stride + 1 || o.d.w != (in.d.w + 2 * d.pad - d.r) / d.stride + 1 || o.d.c != in.d.c ||
            in.cs != in.d.c)
          return fail(ctx, TLXCV_ERR_INVALID, "op %d: avgpool output shape mismatch", i);
        op.impl = kImplAvgpool;
        set_info(op, "avgpool_nhwc", 1, 0, 0, in_bytes + out_bytes, 0, 256, 0, 0);
        break;
      case TLXCV_OP_SPLAT_APPLY: {
        if (d.in1 < 0 || d.r < 1 || d.groups < 1) return fail(ctx, TLXCV_ERR_INVALID, "op %d: split attention needs logits, radix and cardinality", i);
        const TensorRt& lg = p->tensors[d.in1];
        if (in.d.c != d.r * o.d.c || lg.d.c != in.d.c || lg.d.h != 1 || lg.d.w != 1 || lg.d.n != in.d.n || o.d.h != in.d.h || o.d.w != in.d.w ||
            o.d.n != in.d.n || o.d.c % 8 || o.d.c % d.groups || in.cs != in.d.c || lg.cs != lg.d.c)
          return fail(ctx, TLXCV_ERR_INVALID, "op %d: split attention shape mismatch", i);
        if (lg.d.dtype != in.d.dtype) return fail(ctx, TLXCV_ERR_INVALID, "op %d: split attention operands differ in type", i);
        op.impl = kImplSplatApply;
        set_info(op, "splat_apply", 1, 0, 0, in_bytes + out_bytes, 0, 256, 0, 0);
        break;
      }
      case TLXCV_OP_GAP:
        if (o.d.h != 1 || o.d.w != 1 || o.d.c != in.d.c) return fail(ctx, TLXCV_ERR_INVALID, "op %d: gap output shape mismatch", i);
        op.impl = kImplGap;
        set_info(op, "gap_nhwc", 1, 0, 0, in_bytes + out_bytes, in.d.n, 256, 0, 0);
        break;
      case TLXCV_OP_ADD_ACT:
        if (in.cs != in.d.c) return fail(ctx, TLXCV_ERR_UNSUPPORTED, "op %d: add/act on a padded-channel tensor", i);
        op.impl = kImplAddAct;
        set_info(op, "add_act", 1, 0, 0, in_bytes * (d.in1 >= 0 ? 2 : 1) + out_bytes, 0, 256, 0, 0);
        break;
      case TLXCV_OP_UPSAMPLE_CONCAT: {
        const TensorRt* b = d.in1 >= 0 ? &p->tensors[d.in1] : nullptr;
        const int c1 = b ? b->d.c : 0;
        if (d.r < 1 || d.s < 1 || in.d.dtype != TLXCV_ACT || o.d.dtype != TLXCV_ACT || o.d.role != TLXCV_ROLE_INTERNAL ||
            in.d.role != TLXCV_ROLE_INTERNAL || (b && (b->d.dtype != TLXCV_ACT || b->d.role != TLXCV_ROLE_INTERNAL)))
          return fail(ctx, TLXCV_ERR_INVALID, "op %d: upsample_concat works on internal activation tensors", i);
        if (in.d.c % 8 || c1 % 8 || in.cs != in.d.c || (b && b->cs != b->d.c))
          return fail(ctx, TLXCV_ERR_UNSUPPORTED, "op %d: upsample_concat needs channel counts that are multiples of 8", i);
        if (o.d.n != in.d.n || o.d.c != in.d.c + c1 || o.d.h != in.d.h * d.r || o.d.w != in.d.w * d.r ||
            (b && (b->d.n != in.d.n || o.d.h != b->d.h * d.s || o.d.w != b->d.w * d.s)))
          return fail(ctx, TLXCV_ERR_INVALID, "op %d: upsample_concat output shape mismatch", i);
        op.impl = kImplUpsampleConcat;
        set_info(op, "upsample_concat", 1, 0, 0, in_bytes + (b ? static_cast<double>(b->bytes) : 0.0) + out_bytes, 0, 256, 0, 0);
        break;
      }
      case TLXCV_OP_ARGMAX:
        if (in.d.dtype != TLXCV_F32 || o.d.dtype != TLXCV_I64) return fail(ctx, TLXCV_ERR_INVALID, "op %d: argmax expects f32 -> i64", i);
        op.impl = kImplArgmax;
        set_info(op, "argmax_rows", 1, 0, 0, in_bytes + out_bytes, 0, 256, 0, 0);
        break;
      default:
        return fail(ctx, TLXCV_ERR_INVALID, "op %d: unknown kind %d", i, d.kind);
    }
    if (rc != TLXCV_OK) {
      std::string msg = ctx->error;
      return fail(ctx, rc, "op %d: %s", i, msg.c_str());
    }
  }
  // The packing / folding kernels above were enqueued on `st`; the plan may be run on ANY stream afterwards (a pipeline
  // builds on one stream and runs on another), so the one-time preparation is complete before plan_build returns.
  TLX_CUDA(ctx, cudaStreamSynchronize(st));
  guard.p = nullptr;
  *out = p;
  return TLXCV_OK;
}

int tlxcv_plan_destroy(tlxcv_plan* p) {
  if (!p) return TLXCV_OK;
  DeviceGuard dev_guard(p->ctx->device);
  cudaDeviceSynchronize();
  for (auto& g : p->graphs) cudaGraphExecDestroy(g.second);
  for (auto& sgm : p->segments)
    if (sgm.exec) cudaGraphExecDestroy(sgm.exec);
  for (void* q : p->owned) cudaFree(q);
  for (void* q : p->stage_in) cudaFree(q);
  for (void* q : p->stage_out) cudaFree(q);
  if (p->arena) cudaFree(p->arena);
  delete p;
  return TLXCV_OK;
}

int tlxcv_plan_run(tlxcv_plan* p, const void* const* inputs, void* const* outputs, void* stream, int use_graph) {
  if (!p || !inputs || !outputs) return TLXCV_ERR_INVALID;
  tlxcv_ctx* ctx = p->ctx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(ctx->device);
  if (!use_graph) return launch_all(p, inputs, outputs, st);
  std::lock_guard<std::mutex> lock(p->graphs_mu);
  std::vector<const void*> key;
  for (size_t i = 0; i < p->input_ids.size(); ++i) key.push_back(inputs[i]);
  for (size_t i = 0; i < p->output_ids.size(); ++i) key.push_back(outputs[i]);
  for (auto& g : p->graphs)
    if (g.first == key) {
      TLX_CUDA(ctx, cudaGraphLaunch(g.second, st));
      return TLXCV_OK;
    }
  bool seen = false;
  for (auto& k : p->seen_keys) seen = seen || k == key;
  if (seen) {  // the pointer set came back: worth a whole-forward graph
    cudaGraphExec_t exec = nullptr;
    int rc = capture_range(p, 0, static_cast<int>(p->ops.size()), inputs, outputs, &exec);
    if (rc != TLXCV_OK) return rc;
    if (p->graphs.size() >= 8) {  // bounded cache: drop the oldest capture
      cudaGraphExecDestroy(p->graphs.front().second);
      p->graphs.erase(p->graphs.begin());
    }
    p->graphs.emplace_back(key, exec);
    TLX_CUDA(ctx, cudaGraphLaunch(exec, st));
    return TLXCV_OK;
  }
  if (p->seen_keys.size() >= 16) p->seen_keys.erase(p->seen_keys.begin());
  p->seen_keys.push_back(key);
  if (p->segments.empty()) {
    const int n = static_cast<int>(p->ops.size());
    for (int i = 0; i < n;) {
      const bool ext = touches_external(p, p->ops[i]);
      int j = i + 1;
      while (j < n && touches_external(p, p->ops[j]) == ext) ++j;
      p->segments.push_back({i, j, !ext, nullptr});
      i = j;
    }
  }
  for (auto& sgm : p->segments) {
    if (sgm.graphable && sgm.end - sgm.begin >= 2) {
      if (!sgm.exec) {
        int rc = capture_range(p, sgm.begin, sgm.end, inputs, outputs, &sgm.exec);
        if (rc != TLXCV_OK) return rc;
      }
      TLX_CUDA(ctx, cudaGraphLaunch(sgm.exec, st));
    } else {
      int rc = launch_range(p, sgm.begin, sgm.end, inputs, outputs, st);
      if (rc != TLXCV_OK) return rc;
    }
  }
  return TLXCV_OK;
}

int tlxcv_plan_run_host(tlxcv_plan* p, const void* const* host_inputs, void* const* host_outputs, void* stream, int use_graph) {
  if (!p || !host_inputs || !host_outputs) return TLXCV_ERR_INVALID;
  tlxcv_ctx* ctx = p->ctx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(ctx->device);
  if (p->stage_in.empty() && p->stage_out.empty()) {
    for (int t : p->input_ids) {
      void* q = nullptr;
      TLX_CUDA(ctx, cudaMalloc(&q, p->tensors[t].bytes));
      p->stage_in.push_back(q);
    }
    for (int t : p->output_ids) {
      void* q = nullptr;
      TLX_CUDA(ctx, cudaMalloc(&q, p->tensors[t].bytes));
      p->stage_out.push_back(q);
    }
  }
  for (size_t i = 0; i < p->input_ids.size(); ++i)
    TLX_CUDA(ctx, cudaMemcpyAsync(p->stage_in[i], host_inputs[i], p->tensors[p->input_ids[i]].bytes, cudaMemcpyHostToDevice, st));
  int rc = tlxcv_plan_run(p, p->stage_in.data(), p->stage_out.data(), stream, use_graph);
  if (rc != TLXCV_OK) return rc;
  for (size_t i = 0; i < p->output_ids.size(); ++i)
    TLX_CUDA(ctx, cudaMemcpyAsync(host_outputs[i], p->stage_out[i], p->tensors[p->output_ids[i]].bytes, cudaMemcpyDeviceToHost, st));
  return TLXCV_OK;
}

int tlxcv_plan_profile(tlxcv_plan* p, const void* const* inputs, void* const* outputs, void* stream, float* per_op_ms, int n_ops) {
  if (!p || !per_op_ms || n_ops != static_cast<int>(p->ops.size())) return TLXCV_ERR_INVALID;
  tlxcv_ctx* ctx = p->ctx;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DeviceGuard dev_guard(ctx->device);
  std::vector<cudaEvent_t> ev(p->ops.size() + 1);
  for (auto& e : ev) TLX_CUDA(ctx, cudaEventCreate(&e));
  TLX_CUDA(ctx, cudaEventRecord(ev[0], st));
  int rc = TLXCV_OK;
  for (size_t i = 0; i < p->ops.size() && rc == TLXCV_OK; ++i) {
    rc = launch_op(p, p->ops[i], inputs, outputs, st);
    if (rc == TLXCV_OK && cudaEventRecord(ev[i + 1], st) != cudaSuccess) rc = fail(ctx, TLXCV_ERR_CUDA, "cudaEventRecord failed");
  }
  if (rc == TLXCV_OK) {
    cudaError_t e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) rc = fail(ctx, TLXCV_ERR_CUDA, "profile run failed: %s", cudaGetErrorString(e));
  }
  if (rc == TLXCV_OK)
    for (size_t i = 0; i < p->ops.size(); ++i) cudaEventElapsedTime(&per_op_ms[i], ev[i], ev[i + 1]);
  for (auto& e : ev) cudaEventDestroy(e);
  return rc;
}

int tlxcv_plan_num_ops(const tlxcv_plan* p) { return p ? static_cast<int>(p->ops.size()) : 0; }

int tlxcv_plan_num_launches(const tlxcv_plan* p) {
  int n = 0;
  if (p)
    for (const OpRt& op : p->ops) n += op.info.launches;
  return n;
}

int tlxcv_plan_op_info(const tlxcv_plan* p, int i, tlxcv_op_info* info) {
  if (!p || !info || i < 0 || i >= static_cast<int>(p->ops.size())) return TLXCV_ERR_INVALID;
  *info = p->ops[i].info;
  return TLXCV_OK;
}

size_t tlxcv_plan_workspace_bytes(const tlxcv_plan* p) { return p ? p->arena_bytes : 0; }

int tlxcv_plan_read_tensor(tlxcv_plan* p, int t, void* dst, size_t dst_bytes, void* stream) {
  if (!p || t < 0 || t >= static_cast<int>(p->tensors.size()) || !dst) return TLXCV_ERR_INVALID;
  const TensorRt& T = p->tensors[t];
  if (T.d.role != TLXCV_ROLE_INTERNAL) return fail(p->ctx, TLXCV_ERR_INVALID, "read_tensor: tensor %d is external", t);
  if (T.elided || T.wp > 0) return fail(p->ctx, TLXCV_ERR_UNSUPPORTED, "read_tensor: tensor %d is fused away or stored padded", t);
  if (dst_bytes < T.bytes) return fail(p->ctx, TLXCV_ERR_INVALID, "read_tensor: need %zu bytes", T.bytes);
  TLX_CUDA(p->ctx, cudaMemcpyAsync(dst, p->arena + T.offset, T.bytes, cudaMemcpyDeviceToDevice, static_cast<cudaStream_t>(stream)));
  return T.cs;
}

}  // extern "C"

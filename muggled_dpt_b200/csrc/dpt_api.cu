// libdpt_b200.so - host side of the C ABI declared in include/dpt_b200.h: weight registry, workspace arena, launch
// plans (recorded once per (shape, buffers), replayed afterwards) and the stage builders that mirror the reference's
// five sub-models (muggled_dpt/dpt_model.py:61-83).
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "dpt_b200.h"
#include <dlfcn.h>
#include "attn_tc.cuh"
#include "attn64_tc.cuh"
#include "conv_halo.cuh"
#include "gemm_tc.cuh"
#include "kernels_misc.cuh"
#include "kernels_swin.cuh"

using namespace dpt;

namespace {

// ---------------------------------------------------------------------------------------------------------------
// small utilities

thread_local std::string g_err;  // errors without a handle (dpt_create, dpt_op_*)

struct Weight {
  const void* ptr = nullptr;
  int64_t shape[4] = {0, 0, 0, 0};
  int ndim = 0;
  int dtype = 0;
};

using LaunchFnRaw = std::function<cudaError_t(cudaStream_t)>;
struct LaunchFn {
  LaunchFnRaw fn;
  std::string label;   // kernel family + layer, e.g. "gemm256:blk3.fc1"
  double flops = 0;    // algorithmic FLOPs (2*MAC, unpadded)
  double bytes = 0;    // algorithmic HBM bytes (operands read once + outputs written once)
  cudaError_t operator()(cudaStream_t s) const { return fn(s); }
};

struct Plan {
  std::vector<LaunchFn> launches;
  // launches that fill per-grid constant tables (position table, BEiT / SwinV2 bias tables): they depend on the weights
  // and the grid only - the reference memoises them the same way (relative_positional_encoder.py:351-442,
  // windowed_attention.py:232-260) - so they run once, when the plan is first used, and are not replayed
  std::vector<LaunchFn> init;
  bool init_done = false;
  cudaGraphExec_t graph_exec = nullptr;  // the replayed launches as one CUDA graph (built on the second use)
  bool graph_tried = false;
  int uses = 0;
  unsigned long long last_used = 0;
  // cache key
  const void* img = nullptr;
  const void* out = nullptr;
  const void* ws = nullptr;
  int B = 0, H = 0, W = 0;
  bool valid = false;
};

struct Arena {
  char* base = nullptr;
  size_t cap = 0, off = 0, peak = 0;
  size_t top = 0;  // bytes handed out from the END of the workspace: persistent per-plan tables, never reused
  bool dry = false;
  bool overflow = false;
  void* alloc(size_t n) {
    off = (off + 1023) & ~size_t(1023);
    void* p = dry ? nullptr : base + off;
    off += n;
    if (off > peak) peak = off;
    if (!dry && off + top > cap) overflow = true;
    return p;
  }
  void* alloc_persistent(size_t n) {
    top += (n + 1023) & ~size_t(1023);
    if (!dry && (top > cap || peak + top > cap)) overflow = true;
    return dry ? nullptr : base + ((cap - top) & ~size_t(1023));
  }
  size_t need() const { return peak + top + 2048; }
  size_t mark() const { return off; }
  void reset(size_t m) { off = m; }
};

}  // namespace

struct dpt_model_s {
  dpt_config cfg;
  std::unordered_map<std::string, Weight> weights;
  std::unordered_map<std::string, std::vector<float>> host_copies;  // "*_host" weights (tiny, read at plan time)
  std::string err;
  // launch plans of dpt_forward, one per (image buffer, depth buffer, workspace, shape) - a double-buffered host loop
  // alternates between two of them; least recently used is replaced
  static constexpr int kMaxPlans = 4;
  Plan plans[kMaxPlans];
  unsigned long long plan_clock = 0;
  // dpt_forward_host_async: per device buffer, the event of its last reader / writer (see there)
  struct BufEvent { const void* ptr = nullptr; cudaEvent_t ev = nullptr; bool armed = false; };
  std::vector<BufEvent> img_read_done, depth_copied, h2d_done;
  int last_launches = 0;
  int num_sms = 148;
  int device = 0;
  // optional per-launch CUDA-event timing (bench.py's roofline leg)
  bool profiling = false;
  std::vector<cudaEvent_t> events;
  std::vector<std::string> prof_labels;
  std::vector<double> prof_flops, prof_bytes;
};

namespace {

struct Ctx {
  dpt_model_s* m;
  Arena ar;
  std::vector<LaunchFn>* launches;  // null in dry mode
  std::vector<LaunchFn>* init = nullptr;  // per-plan table builders (run once); null: they go to `launches`
  bool dry;
  std::string err;
  bool ok = true;
  int is_bf16;
  int num_sms;
  std::string scope;  // label prefix for launches recorded by the current stage builder
  // debug capture (dpt_encoder_capture): per encoder block, where to dump the attention probabilities / block output
  void* const* cap_probs = nullptr;
  void* const* cap_block_out = nullptr;
  int cap_blocks = 0;
  void* cap_probs_at(int i) const { return (cap_probs && i < cap_blocks) ? cap_probs[i] : nullptr; }
  void* cap_block_out_at(int i) const { return (cap_block_out && i < cap_blocks) ? cap_block_out[i] : nullptr; }
  bool fail(const std::string& s) {
    if (ok) err = s;
    ok = false;
    return false;
  }
  void add(const std::string& label, double flops, double bytes, LaunchFnRaw fn) {
    LaunchFn l;
    l.fn = std::move(fn);
    l.label = label;
    l.flops = flops;
    l.bytes = bytes;
    launches->push_back(std::move(l));
  }
  // a launch that fills a per-grid constant table in persistent workspace
  void add_init(const std::string& label, double bytes, LaunchFnRaw fn) {
    LaunchFn l;
    l.fn = std::move(fn);
    l.label = label;
    l.bytes = bytes;
    (init ? init : launches)->push_back(std::move(l));
  }
};

PFN_cuTensorMapEncodeTiled get_encode_fn() {
  static PFN_cuTensorMapEncodeTiled fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled>(p);
  }
  return fn;
}

// 16-bit tensor map with 128B swizzle. dims/box innermost first; strides in bytes for dims 1..rank-1.
bool make_tmap(CUtensorMap* tm, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides,
               const uint32_t* box, int is_bf16, std::string& err) {
  PFN_cuTensorMapEncodeTiled enc = get_encode_fn();
  if (!enc) {
    err = "cuTensorMapEncodeTiled not available (no CUDA driver?)";
    return false;
  }
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  cuuint64_t d[5], s[5];
  cuuint32_t bx[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides[i];
  CUresult r = enc(tm, is_bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, rank,
                   const_cast<void*>(ptr), d, s, bx, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[256];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled failed (%d): rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u] ptr=%p",
             (int)r, rank, (unsigned long long)d[0], (unsigned long long)(rank > 1 ? d[1] : 0),
             (unsigned long long)(rank > 2 ? d[2] : 0), (unsigned long long)(rank > 3 ? d[3] : 0), bx[0],
             rank > 1 ? bx[1] : 0, rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0, ptr);
    err = buf;
    return false;
  }
  return true;
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPT_PDL");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// cudaLaunchKernelEx with the programmatic-dependent-launch attribute (kernels launched through here call
// pdl_wait() before reading anything a previous kernel wrote) and, optionally, a cluster of two CTAs.
template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool cluster2,
                      Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int na = 0;
  if (cluster2) {
    attr[na].id = cudaLaunchAttributeClusterDimension;
    attr[na].val.clusterDim.x = 2;
    attr[na].val.clusterDim.y = 1;
    attr[na].val.clusterDim.z = 1;
    ++na;
  }
  if (pdl_enabled()) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) belongs to the device's primary context: a kernel has to be opted in
// once per DEVICE, not once per process (two models on two GPUs of one process share these statics).
cudaError_t ensure_dyn_smem(const void* kern, int bytes) {
  static std::mutex mu;
  static std::unordered_map<const void*, uint64_t> done;  // kernel -> bit mask of devices already opted in
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  std::lock_guard<std::mutex> lk(mu);
  uint64_t& mask = done[kern];
  if (dev >= 0 && dev < 64 && ((mask >> dev) & 1)) return cudaSuccess;
  e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e == cudaSuccess && dev >= 0 && dev < 64) mask |= uint64_t(1) << dev;
  return e;
}

// Entry points that take a handle run on the handle's device whatever the caller's current device is.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  cudaError_t err = cudaSuccess;
  explicit DeviceGuard(int dev) {
    err = cudaGetDevice(&prev);
    if (err == cudaSuccess && prev != dev) {
      err = cudaSetDevice(dev);
      switched = err == cudaSuccess;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

int pick_block_n(int N) { return N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256)); }

// the residual-prefetch variant of the fp32-output kernels (gemm_tc.cuh RESPF) is used for short K loops
bool use_respf(const GemmParams& p) { return p.out_kind == OUT_F32 && p.add1 != nullptr && p.num_taps * p.kchunks <= 32; }

template <int BN, int OUT_KIND, int ACT, bool BF16, bool RESPF = false>
cudaError_t launch_gemm_inst(const GemmParams& p, int grid, cudaStream_t s) {
  cudaError_t e = ensure_dyn_smem((const void*)gemm_tc_kernel<BN, OUT_KIND, ACT, BF16, false, RESPF>,
                                  GemmCfg<BN, false, RESPF>::SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return launch_ex(gemm_tc_kernel<BN, OUT_KIND, ACT, BF16, false, RESPF>, dim3((unsigned)grid), dim3(GEMM_THREADS),
                   GemmCfg<BN, false, RESPF>::SMEM_BYTES, s, false, p);
}

// 2-CTA (cta_group::2) variant: BLOCK_N = 256 (or 128 for the 16-bit-output convolutions), clusters of two CTAs
template <int OUT_KIND, int ACT, bool BF16, bool RESPF = false, int BN = 256>
cudaError_t launch_gemm2_inst(const GemmParams& p, int grid, cudaStream_t s) {
  auto kern = gemm_tc_kernel<BN, OUT_KIND, ACT, BF16, true, RESPF>;
  constexpr int smem = GemmCfg<BN, true, RESPF>::SMEM_BYTES;
  cudaError_t e = ensure_dyn_smem((const void*)kern, smem);
  if (e != cudaSuccess) return e;
  return launch_ex(kern, dim3((unsigned)grid), dim3(GEMM_THREADS), smem, s, true, p);
}

template <bool BF16>
cudaError_t launch_gemm2_dt(const GemmParams& p, int bn, int grid, cudaStream_t s) {
  if (bn == 384) {  // 256 x 384 pair tiles: fp32 residual outputs whose m-pairs x 3 n-tiles fit ONE round (small M)
    return use_respf(p) ? launch_gemm2_inst<OUT_F32, ACT_NONE, BF16, true, 384>(p, grid, s)
                        : launch_gemm2_inst<OUT_F32, ACT_NONE, BF16, false, 384>(p, grid, s);
  }
  if (bn == 128) {  // 256 x 128 pair tiles: 16-bit outputs without GELU (3x3 convolutions), fp32 residual outputs (small M)
    if (p.out_kind == OUT_F32)
      return use_respf(p) ? launch_gemm2_inst<OUT_F32, ACT_NONE, BF16, true, 128>(p, grid, s)
                          : launch_gemm2_inst<OUT_F32, ACT_NONE, BF16, false, 128>(p, grid, s);
    if (p.act == ACT_RELU) return launch_gemm2_inst<OUT_HALF, ACT_RELU, BF16, false, 128>(p, grid, s);
    return launch_gemm2_inst<OUT_HALF, ACT_NONE, BF16, false, 128>(p, grid, s);
  }
  if (p.out_kind == OUT_F32)
    return use_respf(p) ? launch_gemm2_inst<OUT_F32, ACT_NONE, BF16, true>(p, grid, s)
                        : launch_gemm2_inst<OUT_F32, ACT_NONE, BF16>(p, grid, s);
  if (p.act == ACT_GELU) return launch_gemm2_inst<OUT_HALF, ACT_GELU, BF16>(p, grid, s);
  if (p.act == ACT_RELU) return launch_gemm2_inst<OUT_HALF, ACT_RELU, BF16>(p, grid, s);
  if (p.act == ACT_SWIGLU) return launch_gemm2_inst<OUT_HALF, ACT_SWIGLU, BF16>(p, grid, s);
  return launch_gemm2_inst<OUT_HALF, ACT_NONE, BF16>(p, grid, s);
}

template <int BN, bool BF16>
cudaError_t launch_gemm_bn(const GemmParams& p, int grid, cudaStream_t s) {
  if (p.out_kind == OUT_F32)
    return use_respf(p) ? launch_gemm_inst<BN, OUT_F32, ACT_NONE, BF16, true>(p, grid, s)
                        : launch_gemm_inst<BN, OUT_F32, ACT_NONE, BF16>(p, grid, s);
  if (p.act == ACT_GELU) return launch_gemm_inst<BN, OUT_HALF, ACT_GELU, BF16>(p, grid, s);
  if (p.act == ACT_RELU) return launch_gemm_inst<BN, OUT_HALF, ACT_RELU, BF16>(p, grid, s);
  if constexpr (BN >= 128) {
    if (p.act == ACT_SWIGLU) return launch_gemm_inst<BN, OUT_HALF, ACT_SWIGLU, BF16>(p, grid, s);
  }
  return launch_gemm_inst<BN, OUT_HALF, ACT_NONE, BF16>(p, grid, s);
}

template <bool BF16>
cudaError_t launch_gemm_dt(const GemmParams& p, int bn, int grid, cudaStream_t s) {
  if (p.out_kind == OUT_HEAD) return launch_gemm_inst<32, OUT_HEAD, ACT_NONE, BF16>(p, grid, s);
  switch (bn) {
    case 32: return launch_gemm_bn<32, BF16>(p, grid, s);
    case 64: return launch_gemm_bn<64, BF16>(p, grid, s);
    case 128: return launch_gemm_bn<128, BF16>(p, grid, s);
    default: return launch_gemm_bn<256, BF16>(p, grid, s);
  }
}

cudaError_t launch_gemm(const GemmParams& p, int bn, int grid, bool two_cta, cudaStream_t s) {
  if (two_cta) return p.is_bf16 ? launch_gemm2_dt<true>(p, bn, grid, s) : launch_gemm2_dt<false>(p, bn, grid, s);
  return p.is_bf16 ? launch_gemm_dt<true>(p, bn, grid, s) : launch_gemm_dt<false>(p, bn, grid, s);
}

constexpr double kCostPair = 1.00, kCost256 = 1.08, kCost128 = 0.60;
// 256 x 128 pair tiles (fp32-output GEMMs only: proj / fc2 at small M), cost per pair tile; DPT_COST_PAIR128 overrides
double cost_pair128() {
  static double v = -1;
  if (v < 0) {
    const char* e = getenv("DPT_COST_PAIR128");
    v = e ? atof(e) : 0.60;  // measured at B=4 / 8 (profiles/r2_pair128_f32.txt): on par with 128 x 128 single-CTA tiles
  }
  return v;
}

// 256 x 384 pair tiles: cost of the single round they run in, in units of a 256 x 256 pair tile (1.5 by area);
// DPT_COST_PAIR384 overrides, DPT_GEMM_WIDE=0 disables them
double cost_pair384() {
  static double v = -1;
  if (v < 0) {
    const char* e = getenv("DPT_COST_PAIR384");
    v = e ? atof(e) : 1.50;
  }
  return v;
}
bool wide_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPT_GEMM_WIDE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// DPT_TMA_STORE=0: every 16-bit output goes through the shuffle + st.global path (A/B switch)
bool tma_store_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPT_TMA_STORE");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

bool two_cta_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPT_GEMM_2CTA");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

// DPT_GEMM_MODE (tuning aid): 0 = cost model (default), 1 = CTA pairs whenever they fill half the machine,
// 2 = 128x256 single-CTA tiles, 3 = 128x128 single-CTA tiles. Modes 1-3 only affect GEMMs with N > 128.
int gemm_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPT_GEMM_MODE");
    v = (e && e[0] >= '0' && e[0] <= '3') ? e[0] - '0' : 0;
  }
  return v;
}

// Tile choice for wide GEMMs (N > 128). Tiles are dealt round-robin to a persistent grid, so a launch costs
// rounds x time-per-tile; the relative tile costs below were fitted to per-launch timings on B200.
struct TileChoice { int bn; bool two_cta; };
TileChoice choose_tile(long long m_tiles, int N, int num_sms, bool f32_out) {
  const long long pairs = (m_tiles + 1) / 2;
  const int nt256 = (N + 255) / 256, nt128 = (N + 127) / 128;
  const bool pairs_ok = two_cta_enabled() && pairs * nt256 >= num_sms / 2;
  switch (gemm_mode()) {
    case 1: return {256, pairs_ok};
    case 2: return {256, false};
    case 3: return {128, false};
    default: break;
  }
  auto rounds = [](long long tiles, int slots) { return (double)((tiles + slots - 1) / slots); };
  const double c_pair = pairs_ok ? rounds(pairs * nt256, num_sms / 2) * kCostPair : 1e30;
  const double c_256 = rounds(m_tiles * nt256, num_sms) * kCost256;
  const double c_128 = rounds(m_tiles * nt128, num_sms) * kCost128;
  const double c_pair128 = (pairs_ok && f32_out) ? rounds(pairs * nt128, num_sms / 2) * cost_pair128() : 1e30;
  // 256 x 384 pair tiles (fp32 outputs): one accumulator stage, so only when everything fits a single round
  const int nt384 = (N + 383) / 384;
  const double c_pair384 = (two_cta_enabled() && f32_out && wide_enabled() && pairs * nt384 <= num_sms / 2 &&
                            pairs * nt384 >= num_sms / 4) ? cost_pair384() : 1e30;
  if (c_pair384 < c_pair && c_pair384 < c_256 && c_pair384 < c_128 && c_pair384 < c_pair128) return {384, true};
  if (c_pair128 < c_pair && c_pair128 < c_256 && c_pair128 < c_128) return {128, true};
  if (c_pair <= c_256 && c_pair <= c_128) return {256, true};
  return c_256 <= c_128 ? TileChoice{256, false} : TileChoice{128, false};
}

// Description of one spatial GEMM (see gemm_tc.cuh).
struct GemmOp {
  const void* A = nullptr;
  int B = 1, Ht = 1, Wt = 1, C = 0;   // A tensor [B, Ht, Wt, C] (C innermost, dense)
  int H = 0, W = 0;                   // output / tiling extent (defaults Ht, Wt - xoff)
  int xoff = 0;
  const void* Wt_ptr = nullptr;       // [N, taps*kpad] 16-bit
  int N = 0, taps = 1, kpad = 0;
  const float* bias = nullptr;
  long long bias_bstride = 0;         // per-image bias vectors (BEiT readout)
  int act = ACT_NONE;
  int out_kind = OUT_HALF;
  void* out = nullptr;
  long long ldo = 0;                  // default N
  int OH = 0, OW = 0, so = 1, oy = 0, ox = 0;  // default OH = H, OW = W
  int shuffle_n = 0;                  // merged ConvTranspose: N = so*so*shuffle_n (gemm_tc.cuh)
  int force_bn = 0;                   // 0 = tile width chosen by the cost model; else single-CTA tiles of this width
  const void* add1 = nullptr;
  long long ld_add1 = 0;
  const void* add2 = nullptr;
  long long ld_add2 = 0;
  void* out2_relu = nullptr;
  long long ld_out2 = 0;
  // folded LayerNorm (gemm_tc.cuh): consumer side reads row statistics, producer side (OUT_F32) writes them
  const float* ln_stats = nullptr;    // [rows, ln_parts, 2]
  int ln_parts = 0;
  const float* ln_colsum = nullptr;   // [N]
  float ln_eps = 0.f;
  float* stats_out = nullptr;         // [rows, *stats_parts, 2]; room for 2 * ceil(N / 64) parts per row
  int* stats_parts = nullptr;         // receives the number of parts this GEMM writes per row
  void* out16 = nullptr;              // 16-bit copy of the fp32 output
  const float* qk_logit = nullptr;    // SwinV2: cosine-normalise q / k heads (32 columns each) in the epilogue, q *= logit[h]
  int qk_features = 0;
  const float* head_w = nullptr;      // host pointer to 32 floats (OUT_HEAD)
  float head_b = 0.f;
  int head_act = ACT_RELU;
  const char* label = "";
};

bool add_gemm(Ctx& c, GemmOp op) {
  if (op.H == 0) op.H = op.Ht;
  if (op.W == 0) op.W = op.Wt - op.xoff;
  if (op.OH == 0) op.OH = op.H * op.so;
  if (op.OW == 0) op.OW = op.W * op.so;
  if (op.ldo == 0) op.ldo = op.act == ACT_SWIGLU ? op.N / 2 : op.N;
  if (op.kpad == 0) op.kpad = (op.C + 63) / 64 * 64;
  if (op.C % 8 != 0) return c.fail("gemm: channel count must be a multiple of 8");
  if (op.out_kind != OUT_HEAD && op.N % 8 != 0) return c.fail("gemm: N must be a multiple of 8");
  if (op.out_kind == OUT_F32 && op.act != ACT_NONE) return c.fail("gemm: fp32 output has no activation variant");
  if (op.act == ACT_SIGMOID) return c.fail("gemm: sigmoid is only available in head mode");
  if (op.act == ACT_SWIGLU) {  // output = N / 2 columns; tiles of 128 or 256 GEMM columns (gemm_tc.cuh)
    if (op.N % 128 != 0 || op.out_kind != OUT_HALF || op.shuffle_n > 0 || op.add1 || op.add2 || op.out2_relu || op.force_bn)
      return c.fail("gemm: the SwiGLU epilogue needs N % 128 == 0, a 16-bit output and no residual / shuffle");
    if (op.ldo == 0) op.ldo = op.N / 2;
  }
  if (c.dry) return true;

  GemmParams p;
  memset(&p, 0, sizeof p);
  // tile shape: minimise the number of 128-pixel tiles
  int best_log2 = 7;
  long long best_tiles = -1;
  for (int l2 = 7; l2 >= 3; --l2) {
    const int tw = 1 << l2, th = 128 >> l2;
    const long long t = (long long)((op.W + tw - 1) / tw) * ((op.H + th - 1) / th);
    if (best_tiles < 0 || t < best_tiles) { best_tiles = t; best_log2 = l2; }
  }
  const int TW = 1 << best_log2, TH = 128 >> best_log2;
  if (op.out_kind == OUT_HEAD && op.N != 32) return c.fail("gemm: head mode needs N == 32");
  const long long m_tiles_all = (long long)op.B * ((op.W + TW - 1) / TW) * ((op.H + TH - 1) / TH);
  int bn = op.out_kind == OUT_HEAD ? 32 : pick_block_n(op.N);
  bool two_cta = false;
  if (op.force_bn > 0) {
    bn = op.force_bn;
  } else if (bn == 256 && op.out_kind != OUT_HEAD) {
    const TileChoice tc = choose_tile(m_tiles_all, op.N, c.num_sms, op.out_kind == OUT_F32);
    bn = tc.bn;
    two_cta = tc.two_cta;
  } else if (bn == 128 && op.taps == 9 && op.out_kind == OUT_HALF && op.act != ACT_GELU && two_cta_enabled() &&
             gemm_mode() == 0 && (m_tiles_all + 1) / 2 >= 4 * (c.num_sms / 2)) {
    // 128-wide 3x3 convolutions (head c1) are bound by L2 -> SM operand traffic: CTA pairs on 256 x 128 tiles stage
    // half of the weight tile each, a quarter less traffic per FLOP
    two_cta = true;
  }

  {
    uint64_t dims[4] = {(uint64_t)op.C, (uint64_t)op.Wt, (uint64_t)op.Ht, (uint64_t)op.B};
    uint64_t str[3] = {(uint64_t)op.C * 2, (uint64_t)op.C * 2 * op.Wt, (uint64_t)op.C * 2 * op.Wt * op.Ht};
    uint32_t box[4] = {64, (uint32_t)TW, (uint32_t)TH, 1};
    if (!make_tmap(&p.tmA, op.A, 4, dims, str, box, c.is_bf16, c.err)) return c.fail(c.err);
  }
  // CTA pairs (cta_group::2, 256 x 256 tiles) when the problem is wide and tall enough to fill the machine with pairs
  const int n_tiles_all = (op.N + bn - 1) / bn;
  {
    const uint64_t ktot = (uint64_t)op.taps * op.kpad;
    uint64_t dims[2] = {ktot, (uint64_t)op.N};
    uint64_t str[1] = {ktot * 2};
    uint32_t box[2] = {64, (uint32_t)(bn == 384 ? 64 : (two_cta ? bn / 2 : bn))};  // 384: three 64-row boxes per CTA
    if (!make_tmap(&p.tmB, op.Wt_ptr, 2, dims, str, box, c.is_bf16, c.err)) return c.fail(c.err);
  }
  p.W = op.W; p.H = op.H; p.B = op.B;
  p.tw_log2 = best_log2;
  p.tiles_x = (op.W + TW - 1) / TW;
  p.tiles_y = (op.H + TH - 1) / TH;
  p.N = op.N;
  p.n_tiles = (op.N + bn - 1) / bn;
  // plain 16-bit outputs leave through the TMA (gemm_tc.cuh tma_store): box = the 32 rows of one epilogue warp
  if (op.out_kind == OUT_HALF && !op.add1 && !op.add2 && !op.out2_relu && op.shuffle_n == 0 && op.so == 1 && op.oy == 0 &&
      op.ox == 0 && bn >= 128 && tma_store_enabled() && (op.ldo * 2) % 16 == 0 && ((uintptr_t)op.out & 15) == 0) {
    const uint64_t n_out = (uint64_t)(op.act == ACT_SWIGLU ? op.N / 2 : op.N);
    uint64_t dims[4] = {n_out, (uint64_t)op.OW, (uint64_t)op.OH, (uint64_t)op.B};
    uint64_t str[3] = {(uint64_t)op.ldo * 2, (uint64_t)op.ldo * 2 * op.OW, (uint64_t)op.ldo * 2 * op.OW * op.OH};
    uint32_t box[4] = {64, (uint32_t)std::min(TW, 32), (uint32_t)std::max(1, 32 / TW), 1};
    if (!make_tmap(&p.tmOut, op.out, 4, dims, str, box, c.is_bf16, c.err)) return c.fail(c.err);
    p.tma_store = 1;
  }
  p.inv_n_tiles = 1.0f / (float)p.n_tiles;
  p.inv_tiles_x = 1.0f / (float)p.tiles_x;
  p.inv_tiles_y = 1.0f / (float)p.tiles_y;
  p.num_taps = op.taps;
  p.kchunks = op.kpad / 64;
  p.a_xoff = op.xoff;
  p.is_bf16 = c.is_bf16;
  p.bias = op.bias;
  p.bias_bstride = op.bias_bstride;
  p.act = op.act;
  p.out_kind = op.out_kind;
  p.out = op.out;
  p.ldo = op.ldo;
  p.OH = op.OH; p.OW = op.OW; p.so = op.so; p.oy = op.oy; p.ox = op.ox;
  if (op.shuffle_n > 0) {
    if (op.shuffle_n % bn != 0 || op.N != op.so * op.so * op.shuffle_n || op.out_kind != OUT_HALF)
      return c.fail("gemm: merged pixel-shuffle needs BLOCK_N | channels");
    p.shuffle_n = op.shuffle_n;
  }
  p.add1 = op.add1; p.ld_add1 = op.ld_add1 ? op.ld_add1 : op.ldo;
  p.add2 = op.add2; p.ld_add2 = op.ld_add2 ? op.ld_add2 : op.ldo;
  p.out2_relu = op.out2_relu; p.ld_out2 = op.ld_out2 ? op.ld_out2 : op.ldo;
  if (bn == 384 && (op.ln_stats != nullptr || op.out_kind != OUT_F32 || !two_cta))
    return c.fail("gemm: 384-wide tiles are CTA-pair fp32-output tiles without a folded LayerNorm");
  if (op.ln_stats != nullptr) {
    if (op.ln_colsum == nullptr || op.ln_parts <= 0 || (op.ln_parts & 1) || op.N < 32) return c.fail("gemm: folded LayerNorm needs colsum / parts");
    p.ln_stats = op.ln_stats; p.ln_parts = op.ln_parts; p.ln_colsum = op.ln_colsum;
    p.ln_inv_f = 1.0f / (float)op.C; p.ln_eps = op.ln_eps;
  }
  if (op.stats_out != nullptr) {
    if (op.out_kind != OUT_F32 || bn < 64 || op.stats_parts == nullptr) return c.fail("gemm: row statistics need an fp32 output and BLOCK_N >= 64");
    p.stats_out = op.stats_out; p.stats_parts = 2 * p.n_tiles; *op.stats_parts = p.stats_parts;
    p.out16 = op.out16; p.ld_out16 = op.ldo;
  }
  if (op.qk_logit != nullptr) {
    if (op.out_kind != OUT_HALF || op.act != ACT_NONE || op.qk_features % 32 != 0 || op.N != 3 * op.qk_features)
      return c.fail("gemm: q/k normalisation needs a 16-bit [q|k|v] output with 32 features per head");
    p.qk_logit = op.qk_logit;
    p.qk_features = op.qk_features;
  }
  if (op.head_w) memcpy(p.head_w, op.head_w, 32 * sizeof(float));
  p.head_b = op.head_b;
  p.head_act = op.head_act;

  const long long total = (long long)p.B * p.tiles_y * p.tiles_x * p.n_tiles;
  if (total >= (1ll << 23)) return c.fail("gemm: more than 2^23 tiles (the kernel's fast tile decomposition is exact below that)");
  int grid = (int)std::min<long long>(total, c.num_sms);
  if (two_cta) grid = (int)std::min<long long>(2 * ((m_tiles_all + 1) / 2) * n_tiles_all, c.num_sms & ~1);

  {
    const double pix = (double)op.B * op.H * op.W;
    const double flops = 2.0 * pix * op.N * op.C * op.taps;
    const double osz = op.out_kind == OUT_F32 ? 4.0 : 2.0;
    double bytes = pix * op.C * 2.0 + (double)op.N * op.taps * op.kpad * 2.0 +
                   (op.out_kind == OUT_HEAD ? pix * 2.0 : pix * op.N * osz);
    if (op.add1) bytes += pix * op.N * osz;
    if (op.add2) bytes += pix * op.N * 2.0;
    if (op.out2_relu) bytes += pix * op.N * 2.0;
    if (op.out16) bytes += pix * op.N * 2.0;
    c.add(std::string("gemm") + std::to_string(bn) + (two_cta ? "x2" : "") + ":" + c.scope + op.label, flops, bytes,
          [p, bn, grid, two_cta](cudaStream_t s) { return launch_gemm(p, bn, grid, two_cta, s); });
  }
  return true;
}

// Depth-head convolution (3x3 -> 32 channels -> ReLU -> 1x1 -> act) with the halo-tile kernel (conv_halo.cuh).
// DPT_HALO=0 falls back to the generic nine-load spatial GEMM (A/B switch for tools/, same results).
bool halo_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPT_HALO");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}

// DPT_HALO_FUSE=1: do not materialise the up-sampled map; the halo kernel's builder warps interpolate each tile from
// the source map (conv_halo.cuh FUSED_RESIZE). Same results (tests/test_model_gpu.py runs it in a subprocess), but
// measured slower on ViT-L B=32: 3.1 ms against 0.64 ms (resize) + 1.2 ms (TMA-fed halo conv) - the builders are bound by
// the latency of their dependent load -> interpolate -> st.shared rounds, not by DRAM (an L2 prefetch of the source
// footprint three tiles ahead changed nothing). Off by default.
bool halo_fuse_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPT_HALO_FUSE");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v != 0;
}

template <bool BF16, bool FUSED>
cudaError_t launch_halo_inst(const HaloParams& p, int grid, cudaStream_t s) {
  cudaError_t e = ensure_dyn_smem((const void*)conv3x3_halo_head_kernel<BF16, FUSED>, HALO_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return launch_ex(conv3x3_halo_head_kernel<BF16, FUSED>, dim3((unsigned)grid),
                   dim3(FUSED ? HALO_THREADS_FUSED : HALO_THREADS), HALO_SMEM_BYTES, s, false, p);
}

// in [B, H, W, C] 16-bit (C <= 128), Wt [32, 9 * kpad], out [B, H, W]. src != nullptr: `in` is not read; the input is
// the bilinear (align_corners=True) resize of src [B, IH, IW, C] to H x W, built tile by tile inside the kernel.
bool add_conv_halo_head(Ctx& c, const void* in, const void* Wt, int kpad, const float* bias, const float* head_w,
                        float head_b, int head_act, void* out, int B, int H, int W, int C, const char* label,
                        const void* src = nullptr, int IH = 0, int IW = 0) {
  if (c.dry) return true;
  HaloParams p;
  memset(&p, 0, sizeof p);
  const bool fused = src != nullptr;
  p.src = src; p.IH = IH; p.IW = IW; p.C = C;
  if (!fused) {
    uint64_t dims[4] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)B};
    uint64_t str[3] = {(uint64_t)C * 2, (uint64_t)C * 2 * W, (uint64_t)C * 2 * W * H};
    uint32_t box[4] = {64, HALO_PW, HALO_PH, 1};
    if (!make_tmap(&p.tmA, in, 4, dims, str, box, c.is_bf16, c.err)) return c.fail(c.err);
  }
  {
    uint64_t dims[2] = {(uint64_t)9 * kpad, (uint64_t)HALO_N};
    uint64_t str[1] = {(uint64_t)9 * kpad * 2};
    uint32_t box[2] = {64, HALO_N};
    if (!make_tmap(&p.tmB, Wt, 2, dims, str, box, c.is_bf16, c.err)) return c.fail(c.err);
  }
  p.W = W; p.H = H; p.B = B;
  p.tiles_x = (W + HALO_TW - 1) / HALO_TW;
  p.tiles_y = (H + HALO_TH - 1) / HALO_TH;
  p.kchunks = kpad / 64;
  p.is_bf16 = c.is_bf16;
  p.bias = bias;
  memcpy(p.head_w, head_w, 32 * sizeof(float));
  p.head_b = head_b;
  p.head_act = head_act;
  p.out = out;
  const long long total = (long long)B * p.tiles_x * p.tiles_y;
  const int grid = (int)std::min<long long>(total, c.num_sms);
  const double pix = (double)B * H * W;
  const int is_bf16 = c.is_bf16;
  const double in_bytes = fused ? (double)B * IH * IW * C * 2.0 : pix * C * 2.0;
  c.add(std::string(fused ? "resize_conv_halo32:" : "conv_halo32:") + c.scope + label, 2.0 * pix * HALO_N * C * 9,
        in_bytes + 9.0 * kpad * HALO_N * 2.0 + pix * 2.0, [p, grid, is_bf16, fused](cudaStream_t s) {
          if (fused) return is_bf16 ? launch_halo_inst<true, true>(p, grid, s) : launch_halo_inst<false, true>(p, grid, s);
          return is_bf16 ? launch_halo_inst<true, false>(p, grid, s) : launch_halo_inst<false, false>(p, grid, s);
        });
  return true;
}

template <bool HAS_BIAS, bool BF16, int HD>
cudaError_t launch_attn_inst(const AttnParams& p, dim3 grid, cudaStream_t s) {
  cudaError_t e = ensure_dyn_smem((const void*)attn_tc_kernel<HAS_BIAS, BF16, HD>, ATT_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return launch_ex(attn_tc_kernel<HAS_BIAS, BF16, HD>, grid, dim3(ATT_THREADS), ATT_SMEM_BYTES, s, false, p);
}
template <bool HAS_BIAS, int HD>
cudaError_t launch_attn(const AttnParams& p, dim3 grid, cudaStream_t s) {
  return p.is_bf16 ? launch_attn_inst<HAS_BIAS, true, HD>(p, grid, s) : launch_attn_inst<HAS_BIAS, false, HD>(p, grid, s);
}

// second-generation kernel (attn64_tc.cuh): 64-column kv steps, four CTAs per SM
template <bool HAS_BIAS, bool BF16, int HD>
cudaError_t launch_attn64_inst(const AttnParams& p, unsigned grid, cudaStream_t s) {
  cudaError_t e = ensure_dyn_smem((const void*)attn64_tc_kernel<HAS_BIAS, BF16, HD>, A2_SMEM_BYTES);
  if (e != cudaSuccess) return e;
  return launch_ex(attn64_tc_kernel<HAS_BIAS, BF16, HD>, dim3(grid), dim3(A2_THREADS), A2_SMEM_BYTES, s, false, p);
}
template <bool HAS_BIAS, int HD>
cudaError_t launch_attn64(const AttnParams& p, unsigned grid, cudaStream_t s) {
  return p.is_bf16 ? launch_attn64_inst<HAS_BIAS, true, HD>(p, grid, s) : launch_attn64_inst<HAS_BIAS, false, HD>(p, grid, s);
}

// DPT_ATTN_V1=1 selects the first-generation kernel (attn_tc.cuh: 128-column kv steps, two CTAs per SM) - an A/B switch
// for tools/ and the ncu before/after captures; same results within the 16-bit rounding of P.
bool attn_v1_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPT_ATTN_V1");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// qkv [B, N, 3F] with F = heads * hd (hd = 64 or 32); bias (optional): [wmod, heads, N, ldb] 16-bit, ldb % 128 == 0
bool add_attention(Ctx& c, const void* qkv, const void* bias, long long ldb, int wmod, void* out, int B, int N,
                   int heads, int hd, float scale) {
  if (hd != 64 && hd != 32) return c.fail("attention: head_dim must be 64 or 32");
  if (bias != nullptr && (ldb % ATT_BN != 0 || ldb < N)) return c.fail("attention: bias row stride must be a multiple of 128");
  if (c.dry) return true;  // (workspace sizing pass: buffers are null)
  if (hd == 32 && bias == nullptr) return c.fail("attention: the head_dim 32 kernel is built with bias only (SwinV2)");
  const bool v1 = attn_v1_enabled();
  AttnParams p;
  memset(&p, 0, sizeof p);
  const int F = heads * hd;
  uint64_t dims[3] = {(uint64_t)3 * F, (uint64_t)N, (uint64_t)B};
  uint64_t str[2] = {(uint64_t)3 * F * 2, (uint64_t)3 * F * 2 * N};
  uint32_t box[3] = {64, v1 ? 128u : (uint32_t)A2_BN, 1};
  if (!make_tmap(&p.tmQKV, qkv, 3, dims, str, box, c.is_bf16, c.err)) return c.fail(c.err);
  p.N = N; p.H = heads; p.B = B; p.F = F;
  p.is_bf16 = c.is_bf16;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.out = out;
  p.bias = bias;
  p.ldb = ldb;
  p.bias_wmod = wmod > 0 ? wmod : 1;
  const bool has_bias = bias != nullptr;
  if (has_bias && !v1) {  // the bias tables as a TMA map: tile (64 kv columns, 128 query rows) of table (b % wmod) * H + h
    uint64_t bdims[3] = {(uint64_t)ldb, (uint64_t)N, (uint64_t)p.bias_wmod * heads};
    uint64_t bstr[2] = {(uint64_t)ldb * 2, (uint64_t)ldb * 2 * N};
    uint32_t bbox[3] = {64, 128, 1};
    if (!make_tmap(&p.tmBias, bias, 3, bdims, bstr, bbox, c.is_bf16, c.err)) return c.fail(c.err);
  }
  const double flops = 4.0 * B * heads * (double)N * N * hd, bytes = 4.0 * (double)B * N * F * 2.0;
  if (v1) {
    dim3 grid((N + ATT_BM - 1) / ATT_BM, heads, B);
    c.add("attn_v1:" + c.scope, flops, bytes, [p, grid, has_bias, hd](cudaStream_t s) {
      if (hd == 32) return launch_attn<true, 32>(p, grid, s);
      return has_bias ? launch_attn<true, 64>(p, grid, s) : launch_attn<false, 64>(p, grid, s);
    });
    return true;
  }
  const unsigned grid = (unsigned)((N + A2_BM - 1) / A2_BM) * heads * B;
  c.add("attn:" + c.scope, flops, bytes, [p, grid, has_bias, hd](cudaStream_t s) {
    if (hd == 32) return launch_attn64<true, 32>(p, grid, s);
    return has_bias ? launch_attn64<true, 64>(p, grid, s) : launch_attn64<false, 64>(p, grid, s);
  });
  return true;
}

// Debug only (dpt_encoder_capture): attention probabilities of one block -> probs [Bt, heads, N, N] 16-bit
bool add_attn_probs(Ctx& c, const void* qkv, const void* bias, long long ldb, int wmod, void* probs, int Bt, int N,
                    int heads, int hd, float scale) {
  if (c.dry) return true;
  const int is_bf16 = c.is_bf16;
  const int wm = wmod > 0 ? wmod : 1;
  c.add("attn_probs(debug):" + c.scope, 6.0 * Bt * heads * (double)N * N * hd, (double)Bt * heads * N * N * 2.0, [=](cudaStream_t s) {
    const dim3 grid((unsigned)N, (unsigned)heads, (unsigned)Bt);
    if (is_bf16)
      attn_probs_kernel<__nv_bfloat16><<<grid, 128, hd * sizeof(float), s>>>((const __nv_bfloat16*)qkv, (const __nv_bfloat16*)bias, ldb, wm,
                                                                          (__nv_bfloat16*)probs, N, heads, hd, scale);
    else
      attn_probs_kernel<__half><<<grid, 128, hd * sizeof(float), s>>>((const __half*)qkv, (const __half*)bias, ldb, wm, (__half*)probs, N,
                                                                    heads, hd, scale);
    return cudaGetLastError();
  });
  return true;
}

int ew_grid(long long n, int block, int num_sms) {
  long long g = (n + block - 1) / block;
  const long long cap = (long long)num_sms * 16;
  return (int)std::max<long long>(1, std::min(g, cap));
}

#define DISPATCH_T(is_bf16, ...)                     \
  do {                                               \
    if (is_bf16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { using T = __half; __VA_ARGS__; }          \
  } while (0)

bool add_layernorm(Ctx& c, const float* x, const float* w, const float* b, void* y, long long M, int F, float eps) {
  if (F % 4 != 0 || F > 1536) return c.fail("layernorm: F must be a multiple of 4 and <= 1536");
  if (c.dry) return true;
  const int is_bf16 = c.is_bf16;
  c.add("layernorm:" + c.scope, 0.0, (double)M * F * 6.0, [=](cudaStream_t s) {
    const int rows_per_block = 8;
    const unsigned grid = (unsigned)((M + rows_per_block - 1) / rows_per_block);
    cudaError_t e;
    DISPATCH_T(is_bf16, (e = launch_ex(layernorm_kernel<T, float>, dim3(grid), dim3(rows_per_block * 32), 0, s, false, x,
                                       w, b, (T*)y, (long long)M, F, eps)));
    return e;
  });
  return true;
}

// LayerNorm on 16-bit input (SwinV2 patch embed: conv -> LayerNorm, v31_swinv2/patch_embed.py:59,92)
bool add_layernorm_h(Ctx& c, const void* x16, const float* w, const float* b, void* y, long long M, int F, float eps) {
  if (F % 4 != 0 || F > 1536) return c.fail("layernorm: F must be a multiple of 4 and <= 1536");
  if (c.dry) return true;
  const int is_bf16 = c.is_bf16;
  c.add("layernorm:" + c.scope, 0.0, (double)M * F * 4.0, [=](cudaStream_t s) {
    const int rows_per_block = 8;
    const unsigned grid = (unsigned)((M + rows_per_block - 1) / rows_per_block);
    cudaError_t e;
    DISPATCH_T(is_bf16, (e = launch_ex(layernorm_kernel<T, T>, dim3(grid), dim3(rows_per_block * 32), 0, s, false,
                                       (const T*)x16, w, b, (T*)y, (long long)M, F, eps)));
    return e;
  });
  return true;
}

bool add_cast_to_half(Ctx& c, const float* x, void* y, long long n, const char* label) {
  if (n % 4 != 0) return c.fail("cast: size must be a multiple of 4");
  if (c.dry) return true;
  const int is_bf16 = c.is_bf16, nsm = c.num_sms;
  c.add(std::string(label) + ":" + c.scope, 0.0, (double)n * 6.0, [=](cudaStream_t s) {
    const int grid = ew_grid(n / 4, 256, nsm);
    cudaError_t e;
    DISPATCH_T(is_bf16, (e = launch_ex(cast_f32_kernel<T>, dim3(grid), dim3(256), 0, s, false, x, (T*)y, (long long)(n / 4))));
    return e;
  });
  return true;
}

bool add_resize(Ctx& c, const void* in, void* out, int B, int IH, int IW, int OH, int OW, int C) {
  if (C % 8 != 0) return c.fail("resize: C must be a multiple of 8");
  if (c.dry) return true;
  const int is_bf16 = c.is_bf16;
  const int nsm = c.num_sms;
  c.add("resize:" + c.scope, 0.0, ((double)B * IH * IW + (double)B * OH * OW) * C * 2.0, [=](cudaStream_t s) {
    (void)nsm;
    const dim3 grid((unsigned)((OW * (C / 8) + 255) / 256), (unsigned)((OH + RESIZE_ROWS - 1) / RESIZE_ROWS), (unsigned)B);
    cudaError_t e;
    DISPATCH_T(is_bf16, (e = launch_ex(resize_bilinear_ac_kernel<T>, grid, dim3(256), 0, s, false, (const T*)in, (T*)out, B,
                                       IH, IW, OH, OW, C)));
    return e;
  });
  return true;
}

bool add_relu_copy(Ctx& c, const void* in, void* out, long long n) {
  if (n % 8 != 0) return c.fail("relu_copy: size must be a multiple of 8");
  if (c.dry) return true;
  const int is_bf16 = c.is_bf16;
  const int nsm = c.num_sms;
  c.add("relu_copy:" + c.scope, 0.0, (double)n * 4.0, [=](cudaStream_t s) {
    const int grid = ew_grid(n / 8, 256, nsm);
    DISPATCH_T(is_bf16, (relu_copy_kernel<T><<<grid, 256, 0, s>>>((const T*)in, (T*)out, n / 8)));
    return cudaGetLastError();
  });
  return true;
}

// ---------------------------------------------------------------------------------------------------------------
// weights

const Weight* get_w(Ctx& c, const std::string& name, int expect_dtype) {
  auto it = c.m->weights.find(name);
  if (it == c.m->weights.end()) {
    c.fail("missing weight: " + name);
    return nullptr;
  }
  if (it->second.dtype != expect_dtype) {
    c.fail("weight " + name + " has the wrong dtype");
    return nullptr;
  }
  return &it->second;
}
int half_dt(const Ctx& c) { return c.is_bf16 ? DPT_BF16 : DPT_F16; }

// ---------------------------------------------------------------------------------------------------------------
// stage builders (DINOv2 / Depth-Anything V2)

// PatchEmbed.forward - v2_depthanything/patch_embed.py:77-99
bool build_patch_embed(Ctx& c, const void* img, void* tokens, int B, int H, int W) {
  const dpt_config& cfg = c.m->cfg;
  const int P = cfg.patch_size_px, F = cfg.features_per_token;
  if (H % P || W % P) return c.fail("image size must be a multiple of the patch size");
  const int gh = H / P, gw = W / P;
  const Weight* w = get_w(c, "patch.w", half_dt(c));
  const Weight* b = get_w(c, "patch.b", DPT_F32);
  if (!w || !b) return false;
  const int kpad = (int)w->shape[1];
  const long long M = (long long)B * gh * gw;
  const size_t mk = c.ar.mark();
  void* A = c.ar.alloc((size_t)M * kpad * 2);
  if (!c.dry) {
    const int is_bf16 = c.is_bf16, nsm = c.num_sms;
    c.add("im2col_patch", 0.0, (double)B * 3 * H * W * 2.0 + (double)M * kpad * 2.0, [=](cudaStream_t s) {
      const int grid = ew_grid(M * kpad / 2, 256, nsm);
      DISPATCH_T(is_bf16, (im2col_patch_kernel<T><<<grid, 256, 0, s>>>((const T*)img, (T*)A, B, 3, H, W, P, gh, gw, kpad)));
      return cudaGetLastError();
    });
  }
  GemmOp op;
  op.A = A; op.B = 1; op.Ht = 1; op.Wt = (int)M; op.C = kpad;
  op.Wt_ptr = w->ptr; op.N = F; op.taps = 1; op.kpad = kpad;
  op.bias = (const float*)b->ptr;
  op.label = "patch_embed";
  c.scope = "";
  bool ok;
  if (cfg.variant == DPT_VARIANT_SWINV2) {
    const Weight *lw = get_w(c, "patch.ln.w", DPT_F32), *lb = get_w(c, "patch.ln.b", DPT_F32);
    if (!lw || !lb) return false;
    void* tmp = c.ar.alloc((size_t)M * F * 2);
    op.out = tmp;
    ok = add_gemm(c, op);
    c.scope = "patch_embed";
    ok = ok && add_layernorm_h(c, tmp, (const float*)lw->ptr, (const float*)lb->ptr, tokens, M, F, cfg.ln_eps);
  } else {
    op.out = tokens;
    ok = add_gemm(c, op);
  }
  c.ar.reset(mk);
  return ok;
}

// DinoV2Model4Stages.forward - v2_depthanything/image_encoder_model.py:80-94, transformer_block.py:53-65,154-170
bool build_encoder(Ctx& c, const void* tokens, void* const taps[4], int B, int gh, int gw) {
  const dpt_config& cfg = c.m->cfg;
  const int F = cfg.features_per_token, heads = cfg.num_heads, L = cfg.num_blocks;
  if (L % 4 != 0) return c.fail("num_blocks must be a multiple of 4");
  if (heads * 64 != F) return c.fail("features_per_token must be 64 * num_heads");
  const int per_stage = L / 4;
  const int N = gh * gw + 1;
  const long long M = (long long)B * N;
  const int hd = half_dt(c);
  const size_t mk = c.ar.mark();
  float* pos = (float*)c.ar.alloc_persistent((size_t)N * F * 4);  // per-grid position table (BEiT: unused)
  float* x = (float*)c.ar.alloc((size_t)M * F * 4);
  void* ln = c.ar.alloc((size_t)M * F * 2);  // 16-bit copy of the residual stream (A operand of qkv / fc1)
  float* stats = (float*)c.ar.alloc((size_t)M * 2 * ((F + 63) / 64) * 8);  // per-row partial (sum, sum sq)
  int stats_parts = 2;  // row_stats_cast_kernel writes two parts per row (the second one empty)
  void* qkv = c.ar.alloc((size_t)M * 3 * F * 2);
  void* att = c.ar.alloc((size_t)M * F * 2);
  void* hid = c.ar.alloc((size_t)M * 4 * F * 2);

  const bool is_beit = cfg.variant == DPT_VARIANT_BEIT;
  const Weight *on_w = nullptr, *on_b = nullptr;
  void* bias_buf = nullptr;
  const long long ldb = (N + ATT_BN - 1) / ATT_BN * ATT_BN;
  if (!is_beit) {
    const Weight* base = get_w(c, "pos.base", DPT_F32);
    const Weight* cls_tok = get_w(c, "pos.cls_tok", DPT_F32);
    const Weight* cls_emb = get_w(c, "pos.cls_emb", DPT_F32);
    on_w = get_w(c, "outnorm.w", DPT_F32);
    on_b = get_w(c, "outnorm.b", DPT_F32);
    if (!base || !cls_tok || !cls_emb || !on_w || !on_b) return false;
    if (!c.dry) {
      const int bh = cfg.base_grid_h, bw = cfg.base_grid_w;
      const int is_bf16 = c.is_bf16, nsm = c.num_sms;
      const float *bp = (const float*)base->ptr, *ct = (const float*)cls_tok->ptr, *ce = (const float*)cls_emb->ptr;
      c.add_init("pos_table", (double)N * F * 4.0, [=](cudaStream_t s) {
        pos_table_kernel<<<N, 128, 0, s>>>(bp, ct, ce, pos, bh, bw, gh, gw, F);
        return cudaGetLastError();
      });
      c.add("assemble_tokens", 0.0, (double)M * F * 6.0, [=](cudaStream_t s) {
        const int grid = ew_grid(M * F / 4, 256, nsm);
        DISPATCH_T(is_bf16, (assemble_tokens_kernel<T><<<grid, 256, 0, s>>>((const T*)tokens, pos, x, B, N, F, N)));
        return cudaGetLastError();
      });
    }
  } else {
    // BEiT: cls token only, no position embedding (v31_beit/image_encoder_model.py:77-79); per-layer bias tables
    const Weight* cls = get_w(c, "beit.cls", DPT_F32);
    if (!cls) return false;
    if (!c.dry) {
      const int is_bf16 = c.is_bf16, nsm = c.num_sms;
      const float* cp = (const float*)cls->ptr;
      c.add("assemble_tokens", 0.0, (double)M * F * 6.0, [=](cudaStream_t s) {
        const int grid = ew_grid(M * F / 4, 256, nsm);
        DISPATCH_T(is_bf16, (assemble_tokens_kernel<T><<<grid, 256, 0, s>>>((const T*)tokens, cp, x, B, N, F, 1)));
        return cudaGetLastError();
      });
    }
  }
  // Both LayerNorms of a block are folded into the GEMM that consumes them (weights.py: qkv.w / fc1.w carry the LN
  // scale, *.b the LN shift, *.s the column sums): the A operand is the raw 16-bit residual stream `ln` and the row
  // statistics `stats`, both written by the epilogue of the GEMM that last updated x (here: by one small kernel).
  if (!c.dry) {
    const int is_bf16 = c.is_bf16;
    c.add("row_stats", 0.0, (double)M * F * 6.0, [=](cudaStream_t s) {
      cudaError_t e;
      DISPATCH_T(is_bf16, (e = launch_ex(row_stats_cast_kernel<T>, dim3((unsigned)((M + 7) / 8)), dim3(256), 0, s, false,
                                         (const float*)x, (T*)ln, stats, M, F)));
      return e;
    });
  }
  const float scale = 1.0f / sqrtf(64.0f);
  for (int i = 0; i < L && c.ok; ++i) {
    const std::string pre = "blk" + std::to_string(i) + ".";
    const Weight *qs = get_w(c, pre + "qkv.s", DPT_F32), *f1s = get_w(c, pre + "fc1.s", DPT_F32);
    const Weight *qw = get_w(c, pre + "qkv.w", hd), *qb = get_w(c, pre + "qkv.b", DPT_F32);
    const Weight *pw = get_w(c, pre + "proj.w", hd), *pb = get_w(c, pre + "proj.b", DPT_F32);
    const Weight *f1w = get_w(c, pre + "fc1.w", hd), *f1b = get_w(c, pre + "fc1.b", DPT_F32);
    const Weight *f2w = get_w(c, pre + "fc2.w", hd), *f2b = get_w(c, pre + "fc2.b", DPT_F32);
    if (!c.ok) return false;
    // ViT-G: fc1 is the doubled inner Linear of the SwiGLU FFN packed [2 hp, F] - hp = h rounded up to 64, rows in
    // blocks of 32 gate rows + the 32 linear rows of the same features, zero rows beyond h - and the gate is applied in
    // fc1's epilogue (ACT_SWIGLU); fc2 is the outer Linear [F, hp]
    const bool swiglu = cfg.mlp_swiglu != 0;
    const int hidden_pad = (int)f2w->shape[1];  // fc2's K, a multiple of 64
    const int hidden = swiglu ? hidden_pad : (int)f1w->shape[0];
    if (swiglu && ((int)f1w->shape[0] != 2 * hidden_pad || hidden_pad > 4 * F))
      return c.fail("SwiGLU: fc1 must be packed as [2 x padded hidden, F] with the padded hidden width <= 4 F");
    c.scope = pre;
    {
      GemmOp op;  // qkv = LN1(x) Wqkv^T + b
      op.A = ln; op.Wt = (int)M; op.C = F; op.Wt_ptr = qw->ptr; op.N = 3 * F; op.kpad = (int)qw->shape[1];
      op.bias = (const float*)qb->ptr; op.out = qkv; op.label = "qkv";
      op.ln_stats = stats; op.ln_parts = stats_parts; op.ln_colsum = (const float*)qs->ptr; op.ln_eps = cfg.ln_eps;
      add_gemm(c, op);
    }
    if (is_beit) {
      // relative position bias of this layer -> [H, N, ldb] (relative_positional_encoder.py:242-309)
      const Weight* tb = get_w(c, pre + "relpos.table", DPT_F32);
      if (!tb) return false;
      bias_buf = c.ar.alloc_persistent((size_t)heads * N * ldb * 2);  // one table per layer, built once per plan
      if (!c.dry) {
        const int is_bf16 = c.is_bf16;
        const int bh = cfg.base_grid_h, bw = cfg.base_grid_w;
        const float* tp = (const float*)tb->ptr;
        void* bb = bias_buf;
        const int ldbi = (int)ldb;
        c.add_init("beit_bias_table:" + pre, (double)heads * N * ldb * 2.0, [=](cudaStream_t s) {
          DISPATCH_T(is_bf16, (beit_bias_table_kernel<T><<<dim3(N, heads), 128, 0, s>>>(tp, (T*)bb, heads, bh, bw, gh, gw, ldbi)));
          return cudaGetLastError();
        });
      }
      if (c.cap_probs_at(i)) add_attn_probs(c, qkv, bias_buf, ldb, 1, c.cap_probs_at(i), B, N, heads, 64, scale);
      add_attention(c, qkv, bias_buf, ldb, 1, att, B, N, heads, 64, scale);
    } else {
      if (c.cap_probs_at(i)) add_attn_probs(c, qkv, nullptr, 0, 1, c.cap_probs_at(i), B, N, heads, 64, scale);
      add_attention(c, qkv, nullptr, 0, 1, att, B, N, heads, 64, scale);
    }
    {
      GemmOp op;  // x += (gamma1 . proj)(att)   (LayerScale folded into the packed weights)
      op.A = att; op.Wt = (int)M; op.C = F; op.Wt_ptr = pw->ptr; op.N = F; op.kpad = (int)pw->shape[1];
      op.bias = (const float*)pb->ptr; op.out_kind = OUT_F32; op.out = x; op.add1 = x; op.label = "proj";
      op.stats_out = stats; op.stats_parts = &stats_parts; op.out16 = ln;
      add_gemm(c, op);
    }
    {
      GemmOp op;  // hid = GELU(LN2(x) W1^T + b1)
      op.A = ln; op.Wt = (int)M; op.C = F; op.Wt_ptr = f1w->ptr; op.N = swiglu ? 2 * hidden : hidden;
      op.kpad = (int)f1w->shape[1];
      op.bias = (const float*)f1b->ptr; op.act = swiglu ? ACT_SWIGLU : ACT_GELU; op.out = hid; op.label = "fc1";
      op.ln_stats = stats; op.ln_parts = stats_parts; op.ln_colsum = (const float*)f1s->ptr; op.ln_eps = cfg.ln_eps;
      add_gemm(c, op);
    }
    {
      GemmOp op;
      op.A = hid; op.Wt = (int)M; op.C = swiglu ? hidden_pad : hidden; op.Wt_ptr = f2w->ptr; op.N = F; op.kpad = hidden_pad;
      op.bias = (const float*)f2b->ptr; op.out_kind = OUT_F32; op.out = x; op.add1 = x; op.label = "fc2";
      op.stats_out = stats; op.stats_parts = &stats_parts; op.out16 = ln;
      add_gemm(c, op);
    }
    if (c.cap_block_out_at(i)) add_cast_to_half(c, x, c.cap_block_out_at(i), M * F, "block_out(debug)");
    // taps: last block of each quarter (DA-V2, BEiT) or the last four blocks (DA-V1, image_encoder_model.py:92-103)
    const bool is_tap = cfg.taps_last4 ? (i >= L - 4) : ((i + 1) % per_stage == 0);
    if (is_tap) {
      const int st = cfg.taps_last4 ? i - (L - 4) : (i + 1) / per_stage - 1;
      c.scope = "outnorm" + std::to_string(st);
      if (!is_beit) {
        add_layernorm(c, x, (const float*)on_w->ptr, (const float*)on_b->ptr, taps[st], M, F, cfg.ln_eps);
      } else if (!c.dry) {  // BEiT taps are the raw residual stream
        const int is_bf16 = c.is_bf16, nsm = c.num_sms;
        void* tp = taps[st];
        c.add("cast_tap:" + c.scope, 0.0, (double)M * F * 6.0, [=](cudaStream_t s) {
          const int grid = ew_grid(M * F / 4, 256, nsm);
          cudaError_t e;
          DISPATCH_T(is_bf16, (e = launch_ex(cast_f32_kernel<T>, dim3(grid), dim3(256), 0, s, false, (const float*)x, (T*)tp,
                                             (long long)(M * F / 4))));
          return e;
        });
      }
    }
  }
  c.ar.reset(mk);
  return c.ok;
}

// ---------------------------------------------------------------------------------------------------------------
// SwinV2 (v31_swinv2/*)

// adjust_window_and_shift_sizes - v31_swinv2/components/windowed_attention.py:345-388 (one axis)
void swin_window_and_shift(int patch, int targ, int& win, int& shift) {
  win = std::min(targ, patch);
  if (patch % win != 0) {
    int best = -1;
    for (int d = win / 2; d < 2 * win; ++d)
      if (d > 0 && patch % d == 0 && (best < 0 || std::abs(patch - d) < std::abs(patch - best))) best = d;
    win = best;
  }
  shift = patch <= win ? 0 : win / 2;
}

// Python slice semantics of make_shift_mask's h_slices / w_slices (windowed_attention.py:420-421)
void swin_mask_slices_axis(int n, int win, int shift, int* s0, int* s1) {
  s0[0] = 0;                          s1[0] = n - win;                  // slice(0, -win)
  s0[1] = n - win;                    s1[1] = shift > 0 ? n - shift : 0;  // slice(-win, -shift): -0 == 0 -> empty
  s0[2] = shift > 0 ? n - shift : 0;  s1[2] = n;                        // slice(-shift, None): -0 == 0 -> everything
}

// SwinV2Model4Stages.forward - v31_swinv2/image_encoder_model.py:77-98,164-169,213-225
bool build_encoder_swin(Ctx& c, const void* tokens, void* const taps[4], int B, int gh, int gw) {
  const dpt_config& cfg = c.m->cfg;
  const int hd = half_dt(c);
  const int F0 = cfg.features_per_token;
  if (gh % 8 || gw % 8) return c.fail("SwinV2 needs a patch grid divisible by 8 (image multiple of 32 px)");
  const size_t mk = c.ar.mark();
  const long long M0 = (long long)B * gh * gw;
  float* x = (float*)c.ar.alloc((size_t)M0 * F0 * 4);
  float* x_next = (float*)c.ar.alloc((size_t)M0 * F0 * 2);  // next stage: N/4 tokens, 2F features
  void* xw = c.ar.alloc((size_t)M0 * F0 * 2);
  void* qkv = c.ar.alloc((size_t)M0 * 3 * F0 * 2);
  void* att = c.ar.alloc((size_t)M0 * F0 * 2);
  void* y = c.ar.alloc((size_t)M0 * F0 * 2);
  void* hid = c.ar.alloc((size_t)M0 * 4 * F0 * 2);
  if (!c.dry) {
    const int is_bf16 = c.is_bf16, nsm = c.num_sms;
    const long long n = M0 * F0;
    c.add("cast_tokens", 0.0, (double)n * 6.0, [=](cudaStream_t s) {
      const int grid = ew_grid(n / 4, 256, nsm);
      DISPATCH_T(is_bf16, (cast_to_f32_kernel<T><<<grid, 256, 0, s>>>((const T*)tokens, x, n / 4)));
      return cudaGetLastError();
    });
  }
  int sgh = gh, sgw = gw;
  int blk_index = 0;  // running block number over the four stages (debug capture)
  for (int st = 0; st < 4 && c.ok; ++st) {
    const int F = F0 << st, heads = cfg.heads_per_stage[st];
    if (heads * 32 != F) return c.fail("SwinV2 stages must have 32 features per head");
    const std::string sp = "sw" + std::to_string(st) + ".";
    if (st > 0) {
      // PatchMerge (components/patch_merge.py:49-103): 2x2 gather -> Linear(4C, 2C, no bias) -> LayerNorm
      const std::string mp = "sw" + std::to_string(st - 1) + ".merge.";
      const Weight *mw = get_w(c, mp + "w", hd), *lw = get_w(c, mp + "ln.w", DPT_F32), *lb = get_w(c, mp + "ln.b", DPT_F32);
      if (!c.ok) return false;
      const int Cp = F / 2, pgh = sgh * 2, pgw = sgw * 2;
      const long long Mn = (long long)B * sgh * sgw;
      c.scope = mp;
      if (!c.dry) {
        const int is_bf16 = c.is_bf16, nsm = c.num_sms;
        const float* xin = x;
        void* mg = hid;
        c.add("patch_merge_gather:" + mp, 0.0, (double)Mn * 4 * Cp * 6.0, [=](cudaStream_t s) {
          const int grid = ew_grid(Mn * 4 * (Cp / 4), 256, nsm);
          DISPATCH_T(is_bf16, (swin_patch_merge_gather_kernel<T><<<grid, 256, 0, s>>>(xin, (T*)mg, B, pgh, pgw, Cp)));
          return cudaGetLastError();
        });
      }
      GemmOp op;
      op.A = hid; op.Wt = (int)Mn; op.C = 4 * Cp; op.Wt_ptr = mw->ptr; op.N = F; op.kpad = (int)mw->shape[1];
      op.out = y; op.label = "reduction";
      add_gemm(c, op);
      if (!c.dry) {
        const int is_bf16 = c.is_bf16;
        const float *lwp = (const float*)lw->ptr, *lbp = (const float*)lb->ptr;
        float* xo = x_next;
        const void* yin = y;
        const float eps = cfg.ln_eps;
        c.add("ln_merge:" + mp, 0.0, (double)Mn * F * 6.0, [=](cudaStream_t s) {
          SwinWin w0{};
          cudaError_t e;
          DISPATCH_T(is_bf16, (e = launch_swin_ln_residual<T, 0, false, 0>((const T*)yin, lwp, lbp, xo, (T*)nullptr, Mn, F, eps, w0, w0, s)));
          return e;
        });
      }
      std::swap(x, x_next);
    }
    const long long M = (long long)B * sgh * sgw;
    int wh, ww, sh, sw;
    swin_window_and_shift(sgh, cfg.window_h, wh, sh);
    swin_window_and_shift(sgw, cfg.window_w, ww, sw);
    const int A = wh * ww, nW = (sgh / wh) * (sgw / ww);
    const long long ldb = (A + ATT_BN - 1) / ATT_BN * ATT_BN;
    const size_t mk_stage = c.ar.mark();
    const int pre_w = cfg.pretrained_window[st];
    const float div_h = (float)std::max((pre_w > 0 ? pre_w : wh) - 1, 1), div_w = (float)std::max((pre_w > 0 ? pre_w : ww) - 1, 1);
    for (int bi = 0; bi < cfg.layers_per_stage[st] && c.ok; ++bi) {
      const std::string pre = sp + std::to_string(bi) + ".";
      c.scope = pre;
      const bool shifted = (bi % 2 == 1) && (sh > 0 || sw > 0);
      const Weight *qw = get_w(c, pre + "qkv.w", hd), *qb = get_w(c, pre + "qkv.b", DPT_F32);
      const Weight* ls = get_w(c, pre + "logit", DPT_F32);
      const Weight *c1 = get_w(c, pre + "cpb.w1", DPT_F32), *cb = get_w(c, pre + "cpb.b1", DPT_F32), *c2 = get_w(c, pre + "cpb.w2", DPT_F32);
      const Weight *pw = get_w(c, pre + "proj.w", hd), *pb = get_w(c, pre + "proj.b", DPT_F32);
      const Weight *n1w = get_w(c, pre + "ln1.w", DPT_F32), *n1b = get_w(c, pre + "ln1.b", DPT_F32);
      const Weight *n2w = get_w(c, pre + "ln2.w", DPT_F32), *n2b = get_w(c, pre + "ln2.b", DPT_F32);
      const Weight *f1w = get_w(c, pre + "fc1.w", hd), *f1b = get_w(c, pre + "fc1.b", DPT_F32);
      const Weight *f2w = get_w(c, pre + "fc2.w", hd), *f2b = get_w(c, pre + "fc2.b", DPT_F32);
      if (!c.ok) return false;
      // continuous-position-bias table and bias (+ shift mask) tables of this block: per-grid constants, built once per plan
      const int n_wm_alloc = shifted ? nW : 1;
      float* table = (float*)c.ar.alloc_persistent((size_t)(2 * wh - 1) * (2 * ww - 1) * heads * 4);
      void* bias = c.ar.alloc_persistent((size_t)n_wm_alloc * heads * A * ldb * 2);
      SwinWin w{sgh, sgw, wh, ww, shifted ? sh : 0, shifted ? sw : 0};
      // the NEXT block of this stage (its window partition / shift is applied by this block's last kernel)
      const bool has_next = bi + 1 < cfg.layers_per_stage[st];
      const bool next_shifted = ((bi + 1) % 2 == 1) && (sh > 0 || sw > 0);
      SwinWin wn{sgh, sgw, wh, ww, next_shifted ? sh : 0, next_shifted ? sw : 0};
      SwinMaskSlices ms{};
      swin_mask_slices_axis(sgh, wh, sh, ms.h0, ms.h1);
      swin_mask_slices_axis(sgw, ww, sw, ms.w0, ms.w1);
      const int n_wm = shifted ? nW : 1;
      if (!c.dry) {
        const int is_bf16 = c.is_bf16, nsm = c.num_sms;
        const float* xin = x;
        const float *w1p = (const float*)c1->ptr, *b1p = (const float*)cb->ptr, *w2p = (const float*)c2->ptr;
        const int ldbi = (int)ldb, sh_i = shifted ? 1 : 0;
        if (bi == 0) {
          // first block of a stage: window partition of the fp32 stream (later blocks receive their windowed 16-bit
          // input from the previous block's post-norm kernel)
          c.add("window_gather:" + pre, 0.0, (double)M * F * 6.0, [=](cudaStream_t s) {
            const int grid = ew_grid(M * (F / 4), 256, nsm);
            DISPATCH_T(is_bf16, (swin_window_gather_kernel<T><<<grid, 256, 0, s>>>(xin, (T*)xw, w, B, F)));
            return cudaGetLastError();
          });
        }
        c.add_init("cpb_table:" + pre, 0.0, [=](cudaStream_t s) {
          swin_cpb_table_kernel<<<(2 * wh - 1) * (2 * ww - 1), 256, 0, s>>>(w1p, b1p, w2p, table, wh, ww, heads, div_h, div_w);
          return cudaGetLastError();
        });
        c.add_init("swin_bias:" + pre, (double)n_wm * heads * A * ldb * 2.0, [=](cudaStream_t s) {
          DISPATCH_T(is_bf16, (swin_bias_kernel<T><<<dim3(A, heads, n_wm), 128, 0, s>>>(table, (T*)bias, w, ms, heads, sh_i, ldbi)));
          return cudaGetLastError();
        });
      }
      {
        GemmOp op;  // qkv = xw Wqkv^T + [q_bias, 0, v_bias]; q, k heads cosine-normalised (q * logit scale) in the epilogue
        op.A = xw; op.Wt = (int)M; op.C = F; op.Wt_ptr = qw->ptr; op.N = 3 * F; op.kpad = (int)qw->shape[1];
        op.bias = (const float*)qb->ptr; op.out = qkv; op.label = "qkv";
        op.qk_logit = (const float*)ls->ptr; op.qk_features = F;
        add_gemm(c, op);
      }
      if (c.cap_probs_at(blk_index)) add_attn_probs(c, qkv, bias, ldb, n_wm, c.cap_probs_at(blk_index), B * nW, A, heads, 32, 1.0f);
      add_attention(c, qkv, bias, ldb, n_wm, att, B * nW, A, heads, 32, 1.0f);
      {
        GemmOp op;
        op.A = att; op.Wt = (int)M; op.C = F; op.Wt_ptr = pw->ptr; op.N = F; op.kpad = (int)pw->shape[1];
        op.bias = (const float*)pb->ptr; op.out = y; op.label = "proj";
        add_gemm(c, op);
      }
      if (!c.dry) {
        // x[pixel(i)] += LN1(y[i]) (window-major -> image scatter) and the 16-bit copy of the new rows = fc1's input
        const int is_bf16 = c.is_bf16;
        const float *gp = (const float*)n1w->ptr, *bp = (const float*)n1b->ptr;
        float* xo = x;
        const void* yin = y;
        void* x16 = xw;
        const float eps = cfg.ln_eps;
        c.add("ln_residual:" + pre + "attn", 0.0, (double)M * F * 12.0, [=](cudaStream_t s) {
          cudaError_t e;
          DISPATCH_T(is_bf16, (e = launch_swin_ln_residual<T, 1, true, 1>((const T*)yin, gp, bp, xo, (T*)x16, M, F, eps, w, w, s)));
          return e;
        });
      }
      {
        GemmOp op;
        op.A = xw; op.Wt = (int)M; op.C = F; op.Wt_ptr = f1w->ptr; op.N = (int)f1w->shape[0]; op.kpad = (int)f1w->shape[1];
        op.bias = (const float*)f1b->ptr; op.act = ACT_GELU; op.out = hid; op.label = "fc1";
        add_gemm(c, op);
      }
      {
        GemmOp op;
        op.A = hid; op.Wt = (int)M; op.C = (int)f1w->shape[0]; op.Wt_ptr = f2w->ptr; op.N = F; op.kpad = (int)f2w->shape[1];
        op.bias = (const float*)f2b->ptr; op.out = y; op.label = "fc2";
        add_gemm(c, op);
      }
      if (!c.dry) {
        // x += LN2(y); the new rows also go, as 16 bits, to their window-major position for the next block's QKV GEMM
        const int is_bf16 = c.is_bf16;
        const float *gp = (const float*)n2w->ptr, *bp = (const float*)n2b->ptr;
        float* xo = x;
        const void* yin = y;
        void* x16 = xw;
        const float eps = cfg.ln_eps;
        c.add("ln_residual:" + pre + "mlp", 0.0, (double)M * F * (has_next ? 12.0 : 10.0), [=](cudaStream_t s) {
          SwinWin w0{};
          cudaError_t e;
          if (has_next) {
            DISPATCH_T(is_bf16, (e = launch_swin_ln_residual<T, 0, true, 2>((const T*)yin, gp, bp, xo, (T*)x16, M, F, eps, w0, wn, s)));
          } else {
            DISPATCH_T(is_bf16, (e = launch_swin_ln_residual<T, 0, true, 0>((const T*)yin, gp, bp, xo, (T*)nullptr, M, F, eps, w0, w0, s)));
          }
          return e;
        });
      }
      if (c.cap_block_out_at(blk_index)) add_cast_to_half(c, x, c.cap_block_out_at(blk_index), M * F, "block_out(debug)");
      ++blk_index;
    }
    c.ar.reset(mk_stage);
    c.scope = "tap" + std::to_string(st);
    add_cast_to_half(c, x, taps[st], M * F, "cast_tap");
    sgh /= 2;
    sgw /= 2;
  }
  c.ar.reset(mk);
  return c.ok;
}

// ReassembleModel.forward - v31_swinv2/reassembly_model.py:61-94,113-122: reshape + 3x3 projection only
bool build_reassemble_swin(Ctx& c, const void* const taps[4], void* const maps[4], void* const maps_relu[4], int B,
                           int gh, int gw) {
  const dpt_config& cfg = c.m->cfg;
  const int hd = half_dt(c);
  for (int k = 0; k < 4 && c.ok; ++k) {
    const std::string pre = "reasm" + std::to_string(k) + ".";
    const Weight* fw = get_w(c, pre + "fuse.w", hd);
    if (!fw) return false;
    c.scope = pre;
    GemmOp op;
    op.A = taps[k]; op.B = B; op.Ht = gh >> k; op.Wt = gw >> k; op.C = cfg.features_per_token << k;
    op.Wt_ptr = fw->ptr; op.N = cfg.fusion_channels; op.taps = 9; op.kpad = (int)fw->shape[1] / 9;
    op.out = maps[k];
    op.out2_relu = maps_relu ? maps_relu[k] : nullptr;
    op.label = "fuse3x3";
    add_gemm(c, op);
  }
  return c.ok;
}

// ReassembleModel.forward - v2_depthanything/reassembly_model.py:61-94,139-149
// maps_relu[i] (optional) receives relu(maps[i]) for the fusion stage's first convolutions.
bool build_reassemble(Ctx& c, const void* const taps[4], void* const maps[4], void* const maps_relu[4], int B, int gh,
                      int gw) {
  const dpt_config& cfg = c.m->cfg;
  const int F = cfg.features_per_token, C = cfg.fusion_channels;
  if (gh % 2 || gw % 2) return c.fail("patch grid must be even (reference: fusion size mismatch otherwise)");
  const int N = gh * gw + 1;
  const int hd = half_dt(c);
  const size_t mk = c.ar.mark();
  for (int k = 0; k < 4 && c.ok; ++k) {
    const int R = cfg.reassembly_features[k];
    const std::string pre = "reasm" + std::to_string(k) + ".";
    const Weight *pw = get_w(c, pre + "proj.w", hd), *pb = get_w(c, pre + "proj.b", DPT_F32);
    const Weight* fw = get_w(c, pre + "fuse.w", hd);
    if (!c.ok) return false;
    const size_t mk2 = c.ar.mark();
    c.scope = pre;
    void* proj = c.ar.alloc((size_t)B * gh * gw * R * 2);
    if (cfg.variant == DPT_VARIANT_BEIT) {
      // readout projection: GELU(Linear(2F,F)([patch, cls])) = GELU(W1 patch + (W2 cls + b)), the bracket being one
      // vector per image (readout_projection.py:71-81)
      const Weight *w1 = get_w(c, pre + "readout.w1", hd), *w2 = get_w(c, pre + "readout.w2", DPT_F32);
      const Weight* rb = get_w(c, pre + "readout.b", DPT_F32);
      if (!c.ok) return false;
      float* u = (float*)c.ar.alloc((size_t)B * F * 4);
      void* ro = c.ar.alloc((size_t)B * gh * gw * F * 2);
      if (!c.dry) {
        const int is_bf16 = c.is_bf16;
        const void* tp = taps[k];
        const float *w2p = (const float*)w2->ptr, *rbp = (const float*)rb->ptr;
        c.add("readout_vec:" + pre, 2.0 * B * F * F, (double)F * F * 4.0, [=](cudaStream_t s) {
          const int warps = B * F;
          DISPATCH_T(is_bf16, (readout_vec_kernel<T><<<(warps * 32 + 255) / 256, 256, 0, s>>>((const T*)tp, w2p, rbp, u, B, N, F)));
          return cudaGetLastError();
        });
      }
      {
        GemmOp op;
        op.A = taps[k]; op.B = B; op.Ht = 1; op.Wt = N; op.C = F; op.xoff = 1;
        op.Wt_ptr = w1->ptr; op.N = F; op.kpad = (int)w1->shape[1];
        op.bias = u; op.bias_bstride = F; op.act = ACT_GELU; op.out = ro; op.label = "readout";
        add_gemm(c, op);
      }
      GemmOp op;
      op.A = ro; op.B = 1; op.Ht = 1; op.Wt = B * gh * gw; op.C = F;
      op.Wt_ptr = pw->ptr; op.N = R; op.kpad = (int)pw->shape[1];
      op.bias = (const float*)pb->ptr; op.out = proj; op.label = "proj1x1";
      add_gemm(c, op);
    } else {
      // 1x1 projection on the patch tokens (cls row skipped through the TMA x offset)
      GemmOp op;
      op.A = taps[k]; op.B = B; op.Ht = 1; op.Wt = N; op.C = F; op.xoff = 1;
      op.Wt_ptr = pw->ptr; op.N = R; op.kpad = (int)pw->shape[1];
      op.bias = (const float*)pb->ptr; op.out = proj; op.label = "proj1x1";
      add_gemm(c, op);
    }
    void* res = proj;
    int rh = gh, rw = gw;
    int Rp = R;  // channels of `res` (ConvTranspose outputs are padded to 32)
    if (k == 0 || k == 1) {
      // ConvTranspose2d(k = s, stride = s) == s*s GEMMs with pixel-shuffled stores
      const int s = k == 0 ? 4 : 2;
      const Weight *uw = get_w(c, pre + "up.w", hd), *ub = get_w(c, pre + "up.b", DPT_F32);
      if (!c.ok) return false;
      rh = gh * s; rw = gw * s;
      // the packer pads the output channels to whole 32-column tiles (zero rows / zero bias), so a tile always lies
      // inside one sub-pixel block and all s*s blocks run as ONE launch with pixel-shuffled stores; `res` carries the
      // padded channels (zeros), which the fuse convolution's zero K-padding ignores
      const int kpad = (int)uw->shape[1];
      Rp = (int)(uw->shape[0] / (s * s));
      if (Rp < R || Rp % 32 != 0 || uw->shape[0] != (long long)s * s * Rp || ub->shape[0] != Rp)
        return c.fail("reassembly: ConvTranspose weights must be packed with the output channels padded to 32");
      res = c.ar.alloc((size_t)B * rh * rw * Rp * 2);
      // tile width: the cost model's choice for multiples of 256 channels, else the widest tile that divides the
      // (padded) channel count (ViT-B: 96 -> 32, 192 -> 64; ViT-S: 48 -> 64, 96 -> 32)
      int bn_forced = 0;
      if (Rp % 256 != 0) bn_forced = Rp % 128 == 0 ? 128 : (Rp % 64 == 0 ? 64 : 32);
      GemmOp op;
      op.A = proj; op.B = B; op.Ht = gh; op.Wt = gw; op.C = R;
      op.Wt_ptr = uw->ptr; op.N = s * s * Rp; op.kpad = kpad; op.ldo = Rp;
      op.bias = (const float*)ub->ptr; op.out = res;
      op.so = s; op.shuffle_n = Rp; op.OH = rh; op.OW = rw; op.label = "convT";
      op.force_bn = bn_forced;
      add_gemm(c, op);
    } else if (k == 3) {
      // Conv2d(k=3, s=2, p=1): im2col gather + GEMM
      const Weight *dw = get_w(c, pre + "down.w", hd), *db = get_w(c, pre + "down.b", DPT_F32);
      if (!c.ok) return false;
      rh = gh / 2; rw = gw / 2;
      const int cpad = (int)dw->shape[1] / 9;
      void* col = c.ar.alloc((size_t)B * rh * rw * 9 * cpad * 2);
      res = c.ar.alloc((size_t)B * rh * rw * R * 2);
      if (!c.dry) {
        const int is_bf16 = c.is_bf16, nsm = c.num_sms;
        const void* pin = proj;
        c.add("im2col_3x3s2", 0.0, (double)B * rh * rw * 9 * cpad * 2.0 * 2.0, [=](cudaStream_t s) {
          const long long total = (long long)B * rh * rw * 9 * (cpad / 8);
          const int grid = ew_grid(total, 256, nsm);
          DISPATCH_T(is_bf16, (im2col_3x3s2_kernel<T><<<grid, 256, 0, s>>>((const T*)pin, (T*)col, B, gh, gw, R, cpad)));
          return cudaGetLastError();
        });
      }
      GemmOp op;
      op.A = col; op.B = 1; op.Ht = 1; op.Wt = B * rh * rw; op.C = 9 * cpad;
      op.Wt_ptr = dw->ptr; op.N = R; op.kpad = 9 * cpad;
      op.bias = (const float*)db->ptr; op.out = res; op.label = "conv3x3s2";
      add_gemm(c, op);
    }
    {
      // fuse_proj: 3x3, R -> C, no bias
      GemmOp op;
      op.A = res; op.B = B; op.Ht = rh; op.Wt = rw; op.C = Rp;
      op.Wt_ptr = fw->ptr; op.N = C; op.taps = 9; op.kpad = (int)fw->shape[1] / 9;
      if (op.kpad < Rp) return c.fail("reassembly: fuse convolution K padding is narrower than its input");
      op.out = maps[k];
      op.out2_relu = maps_relu ? maps_relu[k] : nullptr;
      op.label = "fuse3x3";
      add_gemm(c, op);
    }
    c.ar.reset(mk2);
  }
  c.ar.reset(mk);
  return c.ok;
}

// FusionModel.forward - v2_depthanything/fusion_model.py:55-80,148-154,159-220
// The 1x1 output projection is applied before the x2 bilinear upsample (they commute: both are linear and the
// interpolation weights sum to one), a 4x FLOP saving - SURVEY.md §8a-bis.

// One fusion block (FusionBlock.forward fusion_model.py:148-154; lvl 3 = TopMostFusionBlock :113-114):
//   up = upsample2x(out1x1(RCU2(lvl < 3 ? RCU1(r) + f_prev : r)))      r [B,h,w,C], f_prev [B,h,w,C], up [B,2h,2w,C]
// r_relu (optional) = relu(r) already materialised by the producer of r.
bool build_fusion_level(Ctx& c, int lvl, const void* r, const void* r_relu, const void* f_prev, void* up, int B, int h,
                        int w) {
  const dpt_config& cfg = c.m->cfg;
  const int C = cfg.fusion_channels;
  const int hd = half_dt(c);
  if (lvl < 0 || lvl > 3) return c.fail("fusion: level must be 0..3");
  if (lvl < 3 && f_prev == nullptr && !c.dry) return c.fail("fusion: levels 0..2 need the previous fusion output");
  const size_t mk = c.ar.mark();
  const size_t big = (size_t)B * h * w * C * 2;
  void* t_buf = c.ar.alloc(big);
  void* t_relu = c.ar.alloc(big);
  void* a_relu = c.ar.alloc(big);
  void* v_buf = c.ar.alloc(big);
  void* w_buf = c.ar.alloc(big);
  void* r_relu_tmp = r_relu ? nullptr : c.ar.alloc(big);
  const std::string pre = "fus" + std::to_string(lvl) + ".";
  c.scope = pre;
  if (!r_relu) {
    add_relu_copy(c, r, r_relu_tmp, (long long)B * h * w * C);
    r_relu = r_relu_tmp;
  }
  auto conv3 = [&](const void* in, const std::string& wn, int act, void* out, const void* add1, const void* add2,
                   void* out_relu) {
    const Weight *ww = get_w(c, wn + ".w", hd), *bb = get_w(c, wn + ".b", DPT_F32);
    if (!c.ok) return;
    GemmOp op;
    op.A = in; op.B = B; op.Ht = h; op.Wt = w; op.C = C;
    op.Wt_ptr = ww->ptr; op.N = C; op.taps = 9; op.kpad = (int)ww->shape[1] / 9;
    op.bias = (const float*)bb->ptr; op.act = act; op.out = out; op.add1 = add1; op.add2 = add2;
    op.out2_relu = out_relu;
    const std::string lab = wn.substr(pre.size());
    op.label = lab.c_str();
    add_gemm(c, op);
  };
  const void* t = r;
  const void* tr = r_relu;
  if (lvl < 3) {
    // conv_reassembly (ResidualConv2D) + previous fusion
    conv3(r_relu, pre + "rcu1.c1", ACT_RELU, a_relu, nullptr, nullptr, nullptr);
    conv3(a_relu, pre + "rcu1.c2", ACT_NONE, t_buf, r, f_prev, t_relu);
    t = t_buf;
    tr = t_relu;
  }
  // scale_proj_seq: ResidualConv2D -> (1x1 projection, x2 upsample)
  conv3(tr, pre + "rcu2.c1", ACT_RELU, a_relu, nullptr, nullptr, nullptr);
  conv3(a_relu, pre + "rcu2.c2", ACT_NONE, v_buf, t, nullptr, nullptr);
  {
    const Weight *ww = get_w(c, pre + "out.w", hd), *bb = get_w(c, pre + "out.b", DPT_F32);
    if (!c.ok) return false;
    GemmOp op;
    op.A = v_buf; op.B = B; op.Ht = h; op.Wt = w; op.C = C;
    op.Wt_ptr = ww->ptr; op.N = C; op.taps = 1; op.kpad = (int)ww->shape[1];
    op.bias = (const float*)bb->ptr; op.out = w_buf; op.label = "out1x1";
    add_gemm(c, op);
  }
  add_resize(c, w_buf, up, B, h, w, 2 * h, 2 * w, C);
  c.ar.reset(mk);
  return c.ok;
}

bool build_fusion(Ctx& c, const void* const maps[4], const void* const maps_relu_in[4], void* fused, int B, int h0,
                  int w0) {
  // h0 x w0 = size of the finest reassembly map (4x the patch grid for ViT/BEiT, the patch grid for SwinV2)
  const dpt_config& cfg = c.m->cfg;
  const int C = cfg.fusion_channels;
  if (h0 % 8 || w0 % 8) return c.fail("reassembly map size must be divisible by 8");
  const size_t mk = c.ar.mark();
  // upsampled outputs ping-pong between two buffers sized for the largest consumer (level 0 input = 4g); the final
  // one goes to `fused`.
  void* f_bufs[2];
  f_bufs[0] = c.ar.alloc((size_t)B * h0 * w0 * C * 2);
  f_bufs[1] = c.ar.alloc((size_t)B * h0 * w0 * C * 2);
  const void* f_prev = nullptr;  // previous fusion output, already at this level's resolution
  for (int lvl = 3; lvl >= 0 && c.ok; --lvl) {
    void* up = lvl == 0 ? fused : f_bufs[lvl & 1];
    build_fusion_level(c, lvl, maps[lvl], maps_relu_in ? maps_relu_in[lvl] : nullptr, f_prev, up, B, h0 >> lvl, w0 >> lvl);
    f_prev = up;
  }
  c.ar.reset(mk);
  return c.ok;
}

// MonocularDepthHead.forward - v2_depthanything/head_model.py:61-106
bool build_head(Ctx& c, const void* fused, void* depth, int B, int h, int w) {
  // h x w = fused map size. Upsample factor: P / 8 for Depth-Anything (head_model.py:67), 2 for the MiDaS heads
  // (v31_beit/head_model.py:43, v31_swinv2/head_model.py:43); F.interpolate output size = floor(in * scale)
  const dpt_config& cfg = c.m->cfg;
  const int C = cfg.fusion_channels, P = cfg.patch_size_px;
  const int hd = half_dt(c);
  const double scale = cfg.variant == DPT_VARIANT_DINOV2 ? (double)P / 8.0 : 2.0;
  const int OH = (int)floor((double)h * scale), OW = (int)floor((double)w * scale);
  const Weight *w1 = get_w(c, "head.c1.w", hd), *b1 = get_w(c, "head.c1.b", DPT_F32);
  const Weight *w2 = get_w(c, "head.c2.w", hd), *b2 = get_w(c, "head.c2.b", DPT_F32);
  const Weight *w3 = get_w(c, "head.c3.w_host", DPT_F32), *b3 = get_w(c, "head.c3.b_host", DPT_F32);
  if (!c.ok) return false;
  const int C2 = (int)w1->shape[0];
  const size_t mk = c.ar.mark();
  const int kpad2 = (int)w2->shape[1] / 9;
  const bool use_halo = halo_enabled() && (int)w2->shape[0] == HALO_N && kpad2 <= 64 * HALO_MAX_KCHUNKS && C2 % 8 == 0;
  const bool fuse_resize = use_halo && halo_fuse_enabled();  // the up-sampled map is never materialised
  void* h1 = c.ar.alloc((size_t)B * h * w * C2 * 2);
  void* h2 = fuse_resize ? nullptr : c.ar.alloc((size_t)B * OH * OW * C2 * 2);
  {
    GemmOp op;
    op.A = fused; op.B = B; op.Ht = h; op.Wt = w; op.C = C;
    op.Wt_ptr = w1->ptr; op.N = C2; op.taps = 9; op.kpad = (int)w1->shape[1] / 9;
    op.bias = (const float*)b1->ptr; op.out = h1; op.label = "c1";
    c.scope = "head.";
    add_gemm(c, op);
  }
  if (!fuse_resize) add_resize(c, h1, h2, B, h, w, OH, OW, C2);
  if (use_halo) {
    c.scope = "head.";
    add_conv_halo_head(c, h2, w2->ptr, kpad2, (const float*)b2->ptr, (const float*)w3->ptr, *(const float*)b3->ptr,
                       cfg.is_metric ? ACT_SIGMOID : ACT_RELU, depth, B, OH, OW, C2, "c2c3",
                       fuse_resize ? h1 : nullptr, h, w);
  } else {
    GemmOp op;
    op.A = h2; op.B = B; op.Ht = OH; op.Wt = OW; op.C = C2;
    op.Wt_ptr = w2->ptr; op.N = 32; op.taps = 9; op.kpad = (int)w2->shape[1] / 9;
    op.bias = (const float*)b2->ptr; op.out_kind = OUT_HEAD; op.out = depth; op.ldo = 1;
    op.head_w = (const float*)w3->ptr;        // host memory (see dpt_set_weight: "*_host" names are copied)
    op.head_b = *(const float*)b3->ptr;
    op.head_act = cfg.is_metric ? ACT_SIGMOID : ACT_RELU;
    op.label = "c2c3";
    add_gemm(c, op);
  }
  c.ar.reset(mk);
  return c.ok;
}

struct FwdBuffers {
  void* tokens;
  void* taps[4];
  void* maps[4];
  void* maps_relu[4];
  void* fused;
};

// finest reassembly-map size for a patch grid: 4x the grid (ViT / BEiT, reassembly_model.py:61-94), the grid itself
// for SwinV2 (v31_swinv2/reassembly_model.py:61-94)
inline int map0_scale_num(const dpt_config& cfg) { return cfg.variant == DPT_VARIANT_SWINV2 ? 1 : 4; }

bool stage_encoder(Ctx& c, const void* tokens, void* const taps[4], int B, int gh, int gw) {
  return c.m->cfg.variant == DPT_VARIANT_SWINV2 ? build_encoder_swin(c, tokens, taps, B, gh, gw)
                                                 : build_encoder(c, tokens, taps, B, gh, gw);
}
bool stage_reassemble(Ctx& c, const void* const taps[4], void* const maps[4], void* const maps_relu[4], int B, int gh,
                      int gw) {
  return c.m->cfg.variant == DPT_VARIANT_SWINV2 ? build_reassemble_swin(c, taps, maps, maps_relu, B, gh, gw)
                                                 : build_reassemble(c, taps, maps, maps_relu, B, gh, gw);
}

bool build_forward(Ctx& c, const void* img, void* depth, int B, int H, int W) {
  const dpt_config& cfg = c.m->cfg;
  const int P = cfg.patch_size_px, F = cfg.features_per_token, C = cfg.fusion_channels;
  const bool swin = cfg.variant == DPT_VARIANT_SWINV2;
  if (H % P || W % P) return c.fail("image height/width must be multiples of the patch size");
  const int gh = H / P, gw = W / P;
  if (!swin && (gh % 2 || gw % 2))
    return c.fail("patch grid must be even in both directions (the reference raises inside fusion for odd grids)");
  if (swin && (gh % 8 || gw % 8)) return c.fail("SwinV2 needs image sizes that are multiples of 32 px");
  FwdBuffers fb;
  fb.tokens = c.ar.alloc((size_t)B * gh * gw * F * 2);
  for (int k = 0; k < 4; ++k) {
    const size_t n = swin ? (size_t)B * (gh >> k) * (gw >> k) * ((size_t)F << k) : (size_t)B * (gh * gw + 1) * F;
    fb.taps[k] = c.ar.alloc(n * 2);
  }
  const int h0 = gh * map0_scale_num(cfg), w0 = gw * map0_scale_num(cfg);
  for (int k = 0; k < 4; ++k) {
    fb.maps[k] = c.ar.alloc((size_t)B * (h0 >> k) * (w0 >> k) * C * 2);
    fb.maps_relu[k] = c.ar.alloc((size_t)B * (h0 >> k) * (w0 >> k) * C * 2);
  }
  fb.fused = c.ar.alloc((size_t)B * h0 * 2 * w0 * 2 * C * 2);
  if (!build_patch_embed(c, img, fb.tokens, B, H, W)) return false;
  if (!stage_encoder(c, fb.tokens, fb.taps, B, gh, gw)) return false;
  if (!stage_reassemble(c, fb.taps, fb.maps, fb.maps_relu, B, gh, gw)) return false;
  if (!build_fusion(c, fb.maps, fb.maps_relu, fb.fused, B, h0, w0)) return false;
  if (!build_head(c, fb.fused, depth, B, 2 * h0, 2 * w0)) return false;
  return c.ok;
}

Ctx make_ctx(dpt_model_s* m, void* ws, size_t ws_bytes, std::vector<LaunchFn>* launches, bool dry) {
  Ctx c;
  c.m = m;
  c.ar.base = (char*)ws;
  c.ar.cap = ws_bytes;
  c.ar.dry = dry;
  c.launches = launches;
  c.dry = dry;
  c.is_bf16 = m ? (m->cfg.dtype == DPT_BF16) : 1;
  c.num_sms = m ? m->num_sms : 148;
  return c;
}

int run_launches(dpt_model_s* m, std::vector<LaunchFn>& launches, cudaStream_t s, std::string& err) {
  int n = 0;
  const bool prof = m && m->profiling;
  if (prof) {
    // one event before every launch + one after the last; durations are read back by dpt_profile_get
    while (m->events.size() < launches.size() + 1) {
      cudaEvent_t ev;
      if (cudaEventCreate(&ev) != cudaSuccess) { err = "cudaEventCreate failed"; return DPT_ERR_CUDA; }
      m->events.push_back(ev);
    }
    m->prof_labels.clear();
    m->prof_flops.clear();
    m->prof_bytes.clear();
  }
  for (auto& f : launches) {
    if (prof) {
      cudaEventRecord(m->events[n], s);
      m->prof_labels.push_back(f.label);
      m->prof_flops.push_back(f.flops);
      m->prof_bytes.push_back(f.bytes);
    }
    cudaError_t e = f(s);
    if (e != cudaSuccess) {
      err = std::string("kernel launch failed (") + f.label + "): " + cudaGetErrorString(e);
      return DPT_ERR_CUDA;
    }
    ++n;
  }
  if (prof) cudaEventRecord(m->events[n], s);
  if (m) m->last_launches = n;
  return DPT_OK;
}

template <typename BuildFn>
int build_and_run(dpt_model_s* h, void* ws, size_t ws_bytes, void* stream, BuildFn&& fn) {
  if (!h) return DPT_ERR_INVALID;
  DeviceGuard guard(h->device);
  if (guard.err != cudaSuccess) { h->err = std::string("cudaSetDevice: ") + cudaGetErrorString(guard.err); return DPT_ERR_CUDA; }
  // a stage call lays its own tables / scratch out in the persistent end of the workspace: if the forward plan lives in
  // the same workspace its per-grid tables have to be rebuilt before its next use
  for (Plan& pl : h->plans) pl.init_done = false;
  std::vector<LaunchFn> launches;
  Ctx c = make_ctx(h, ws, ws_bytes, &launches, false);
  if (!fn(c)) {
    h->err = c.err;
    return c.err.rfind("missing weight", 0) == 0 ? DPT_ERR_MISSING : DPT_ERR_INVALID;
  }
  if (c.ar.overflow) {
    h->err = "workspace too small: need " + std::to_string(c.ar.need()) + " bytes";
    return DPT_ERR_WORKSPACE;
  }
  return run_launches(h, launches, (cudaStream_t)stream, h->err);  // (one-off plan: table builders run in line)
}

}  // namespace

// =================================================================================================================
// C ABI
// =================================================================================================================

extern "C" {

const char* dpt_version(void) { return "dpt_b200 0.2 (sm_100a)"; }
int dpt_config_size(void) { return (int)sizeof(dpt_config); }

int dpt_create(const dpt_config* cfg, dpt_handle* out) {
  if (!cfg || !out) { g_err = "null argument"; return DPT_ERR_INVALID; }
  if (cfg->variant != DPT_VARIANT_DINOV2 && cfg->variant != DPT_VARIANT_BEIT && cfg->variant != DPT_VARIANT_SWINV2) {
    g_err = "variant not supported by this build";
    return DPT_ERR_UNSUPPORTED;
  }
  if (cfg->dtype != DPT_BF16 && cfg->dtype != DPT_F16) { g_err = "dtype must be fp16 or bf16"; return DPT_ERR_INVALID; }
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) { g_err = std::string("cudaGetDevice: ") + cudaGetErrorString(e); return DPT_ERR_CUDA; }
  cudaDeviceProp prop;
  e = cudaGetDeviceProperties(&prop, dev);
  if (e != cudaSuccess) { g_err = std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e); return DPT_ERR_CUDA; }
  if (prop.major != 10) {
    g_err = "libdpt_b200 requires an sm_100 (Blackwell B200) device; found sm_" + std::to_string(prop.major) +
            std::to_string(prop.minor);
    return DPT_ERR_UNSUPPORTED;
  }
  dpt_model_s* m = new dpt_model_s();
  m->cfg = *cfg;
  m->num_sms = prop.multiProcessorCount;
  m->device = dev;
  *out = m;
  return DPT_OK;
}

void dpt_destroy(dpt_handle h) {
  if (!h) return;
  for (Plan& pl : h->plans)
    if (pl.graph_exec) cudaGraphExecDestroy(pl.graph_exec);
  for (cudaEvent_t ev : h->events) cudaEventDestroy(ev);
  for (auto* v : {&h->img_read_done, &h->depth_copied, &h->h2d_done})
    for (auto& b : *v)
      if (b.ev) cudaEventDestroy(b.ev);
  delete h;
}

int dpt_profile_enable(dpt_handle h, int on) {
  if (!h) return DPT_ERR_INVALID;
  h->profiling = on != 0;
  return DPT_OK;
}

int dpt_profile_count(dpt_handle h) { return h ? (int)h->prof_labels.size() : 0; }

int dpt_profile_get(dpt_handle h, int i, char* label, int label_len, double* ms, double* flops, double* bytes) {
  if (!h || i < 0 || i >= (int)h->prof_labels.size() || (size_t)i + 1 >= h->events.size() + 0 + 1) return DPT_ERR_INVALID;
  float t = 0.f;
  cudaError_t e = cudaEventElapsedTime(&t, h->events[i], h->events[i + 1]);
  if (e != cudaSuccess) { h->err = std::string("cudaEventElapsedTime: ") + cudaGetErrorString(e); return DPT_ERR_CUDA; }
  if (label && label_len > 0) {
    strncpy(label, h->prof_labels[i].c_str(), label_len - 1);
    label[label_len - 1] = 0;
  }
  if (ms) *ms = t;
  if (flops) *flops = h->prof_flops[i];
  if (bytes) *bytes = h->prof_bytes[i];
  return DPT_OK;
}

const char* dpt_last_error(dpt_handle h) { return h ? h->err.c_str() : g_err.c_str(); }
const char* dpt_op_last_error(void) { return g_err.c_str(); }
int dpt_last_launch_count(dpt_handle h) { return h ? h->last_launches : 0; }

int dpt_set_weight(dpt_handle h, const char* name, const void* dev_ptr, const int64_t* shape, int ndim, int dtype) {
  if (!h || !name || !dev_ptr || ndim < 0 || ndim > 4) return DPT_ERR_INVALID;
  Weight w;
  w.ptr = dev_ptr;
  w.ndim = ndim;
  w.dtype = dtype;
  for (int i = 0; i < ndim; ++i) w.shape[i] = shape[i];
  const std::string nm(name);
  if (nm.size() > 5 && nm.compare(nm.size() - 5, 5, "_host") == 0) {
    // host-resident fp32 vector: copied into the handle (used as kernel parameters, e.g. the head's 32->1 weights)
    if (dtype != DPT_F32) { h->err = "host weights must be f32"; return DPT_ERR_INVALID; }
    int64_t n = 1;
    for (int i = 0; i < ndim; ++i) n *= shape[i];
    std::vector<float>& v = h->host_copies[nm];
    v.assign((const float*)dev_ptr, (const float*)dev_ptr + n);
    w.ptr = v.data();
  }
  h->weights[name] = w;
  for (Plan& pl : h->plans) pl.valid = false;  // (the next dpt_forward rebuilds the plan, its tables and its graph)
  return DPT_OK;
}

int dpt_workspace_bytes(dpt_handle h, int B, int H, int W, size_t* bytes) {
  if (!h || !bytes) return DPT_ERR_INVALID;
  Ctx c = make_ctx(h, nullptr, 0, nullptr, true);
  if (!build_forward(c, nullptr, nullptr, B, H, W)) {
    h->err = c.err;
    return c.err.rfind("missing weight", 0) == 0 ? DPT_ERR_MISSING : DPT_ERR_INVALID;
  }
  *bytes = c.ar.need();
  return DPT_OK;
}

// DPT_GRAPH=0 replays a plan launch by launch instead of as one CUDA graph (A/B switch for tools/)
static bool graph_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DPT_GRAPH");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

static void destroy_plan_graph(Plan& pl) {
  if (pl.graph_exec) {
    cudaGraphExecDestroy(pl.graph_exec);
    pl.graph_exec = nullptr;
  }
}

// Capture the plan's launches (programmatic-dependent-launch attributes included: they become programmatic graph
// edges) into one executable graph. Runs on a private capture stream in thread-local mode, so nothing else the
// process does is recorded; a failure leaves the plan on the launch-by-launch path.
static bool capture_plan_graph(dpt_model_s* h, Plan& pl) {
  cudaStream_t cs = nullptr;
  if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) return false;
  bool ok = cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
  if (ok) {
    for (auto& f : pl.launches)
      if (f(cs) != cudaSuccess) { ok = false; break; }
    cudaGraph_t g = nullptr;
    const cudaError_t e = cudaStreamEndCapture(cs, &g);  // must be called even after a failed launch
    ok = ok && e == cudaSuccess && g != nullptr;
    if (ok) ok = cudaGraphInstantiate(&pl.graph_exec, g, 0) == cudaSuccess;
    if (g) cudaGraphDestroy(g);
  }
  cudaStreamDestroy(cs);
  if (!ok) {
    cudaGetLastError();  // clear the sticky-free error state of the failed capture
    pl.graph_exec = nullptr;
  }
  (void)h;
  return ok;
}

int dpt_forward(dpt_handle h, const void* img, void* depth, void* ws, size_t ws_bytes, int B, int H, int W,
                void* stream) {
  if (!h || !img || !depth || !ws) return DPT_ERR_INVALID;
  DeviceGuard guard(h->device);
  if (guard.err != cudaSuccess) { h->err = std::string("cudaSetDevice: ") + cudaGetErrorString(guard.err); return DPT_ERR_CUDA; }
  Plan* hit = nullptr;
  Plan* victim = nullptr;  // an unused slot, else the least recently used plan
  for (Plan& c : h->plans) {
    if (c.valid && c.img == img && c.out == depth && c.ws == ws && c.B == B && c.H == H && c.W == W) hit = &c;
    if (!c.valid) {
      if (!victim || victim->valid) victim = &c;
    } else if (!victim || (victim->valid && c.last_used < victim->last_used)) {
      victim = &c;
    }
  }
  Plan& pl = hit ? *hit : *victim;
  pl.last_used = ++h->plan_clock;
  if (!hit) {
    // plans that share this workspace keep their per-grid tables in the same place with the same contents, but a plan
    // for another shape in the same workspace would overwrite them: those rebuild their tables on their next use
    for (Plan& c : h->plans)
      if (&c != &pl && c.valid && c.ws == ws && !(c.B == B && c.H == H && c.W == W)) c.init_done = false;
    pl.valid = false;
    pl.launches.clear();
    pl.init.clear();
    pl.init_done = false;
    pl.uses = 0;
    pl.graph_tried = false;
    destroy_plan_graph(pl);
    Ctx c = make_ctx(h, ws, ws_bytes, &pl.launches, false);
    c.init = &pl.init;
    if (!build_forward(c, img, depth, B, H, W)) {
      h->err = c.err;
      return c.err.rfind("missing weight", 0) == 0 ? DPT_ERR_MISSING : DPT_ERR_INVALID;
    }
    if (c.ar.overflow) {
      h->err = "workspace too small: need " + std::to_string(c.ar.need()) + " bytes";
      return DPT_ERR_WORKSPACE;
    }
    pl.img = img; pl.out = depth; pl.ws = ws; pl.B = B; pl.H = H; pl.W = W;
    pl.valid = true;
  }
  cudaStream_t s = (cudaStream_t)stream;
  if (!pl.init_done) {
    // per-grid tables: once per plan, ahead of the first forward on the same stream
    for (auto& f : pl.init) {
      const cudaError_t e = f(s);
      if (e != cudaSuccess) {
        h->err = std::string("kernel launch failed (") + f.label + "): " + cudaGetErrorString(e);
        return DPT_ERR_CUDA;
      }
    }
    pl.init_done = true;
  }
  ++pl.uses;
  // first use: launch by launch (this also opts every kernel into its shared-memory size on this device); second use:
  // capture; afterwards: one cudaGraphLaunch per forward. Per-launch profiling needs the launch-by-launch path.
  if (graph_enabled() && !h->profiling && pl.uses >= 2) {
    if (!pl.graph_exec && !pl.graph_tried) {
      pl.graph_tried = true;
      capture_plan_graph(h, pl);
    }
    if (pl.graph_exec) {
      const cudaError_t e = cudaGraphLaunch(pl.graph_exec, s);
      if (e != cudaSuccess) {
        h->err = std::string("cudaGraphLaunch: ") + cudaGetErrorString(e);
        return DPT_ERR_CUDA;
      }
      h->last_launches = (int)pl.launches.size();
      return DPT_OK;
    }
  }
  return run_launches(h, pl.launches, s, h->err);
}

int dpt_forward_host(dpt_handle h, const void* host_img, void* host_depth, void* dev_img, void* dev_depth, void* ws,
                     size_t ws_bytes, int B, int H, int W, void* stream) {
  if (!h || !host_img || !host_depth || !dev_img || !dev_depth) return DPT_ERR_INVALID;
  DeviceGuard guard(h->device);
  if (guard.err != cudaSuccess) { h->err = std::string("cudaSetDevice: ") + cudaGetErrorString(guard.err); return DPT_ERR_CUDA; }
  cudaStream_t s = (cudaStream_t)stream;
  cudaError_t e = cudaMemcpyAsync(dev_img, host_img, (size_t)B * 3 * H * W * 2, cudaMemcpyHostToDevice, s);
  if (e != cudaSuccess) { h->err = std::string("H2D copy: ") + cudaGetErrorString(e); return DPT_ERR_CUDA; }
  int rc = dpt_forward(h, dev_img, dev_depth, ws, ws_bytes, B, H, W, stream);
  if (rc != DPT_OK) return rc;
  e = cudaMemcpyAsync(host_depth, dev_depth, (size_t)B * H * W * 2, cudaMemcpyDeviceToHost, s);
  if (e != cudaSuccess) { h->err = std::string("D2H copy: ") + cudaGetErrorString(e); return DPT_ERR_CUDA; }
  e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { h->err = std::string("stream sync: ") + cudaGetErrorString(e); return DPT_ERR_CUDA; }
  return DPT_OK;
}

// event attached to a device buffer (created on first use); `armed` = it has been recorded at least once
static cudaEvent_t buf_event(std::vector<dpt_model_s::BufEvent>& v, const void* ptr, bool** armed) {
  for (auto& b : v)
    if (b.ptr == ptr) { *armed = &b.armed; return b.ev; }
  dpt_model_s::BufEvent b;
  b.ptr = ptr;
  if (cudaEventCreateWithFlags(&b.ev, cudaEventDisableTiming) != cudaSuccess) return nullptr;
  v.push_back(b);
  *armed = &v.back().armed;
  return v.back().ev;
}

int dpt_forward_host_async(dpt_handle h, const void* host_img, void* host_depth, void* dev_img, void* dev_depth, void* ws,
                           size_t ws_bytes, int B, int H, int W, void* stream, void* copy_in_stream, void* copy_out_stream) {
  if (!h || !host_img || !host_depth || !dev_img || !dev_depth || !copy_in_stream || !copy_out_stream) return DPT_ERR_INVALID;
  DeviceGuard guard(h->device);
  if (guard.err != cudaSuccess) { h->err = std::string("cudaSetDevice: ") + cudaGetErrorString(guard.err); return DPT_ERR_CUDA; }
  cudaStream_t s = (cudaStream_t)stream, s_in = (cudaStream_t)copy_in_stream, s_out = (cudaStream_t)copy_out_stream;
  bool *a_read = nullptr, *a_copied = nullptr, *a_h2d = nullptr;
  cudaEvent_t ev_read = buf_event(h->img_read_done, dev_img, &a_read);      // last forward that READ dev_img is done
  cudaEvent_t ev_copied = buf_event(h->depth_copied, dev_depth, &a_copied);  // last D2H that read dev_depth is done
  cudaEvent_t ev_h2d = buf_event(h->h2d_done, dev_img, &a_h2d);              // this call's H2D into dev_img is done
  if (!ev_read || !ev_copied || !ev_h2d) { h->err = "cudaEventCreate failed"; return DPT_ERR_CUDA; }
  auto fail = [&](const char* what, cudaError_t e) { h->err = std::string(what) + ": " + cudaGetErrorString(e); return DPT_ERR_CUDA; };
  cudaError_t e;
  // 1. H2D on the copy-in stream, once the previous forward that read this image buffer has finished with it
  if (*a_read && (e = cudaStreamWaitEvent(s_in, ev_read, 0)) != cudaSuccess) return fail("cudaStreamWaitEvent", e);
  if ((e = cudaMemcpyAsync(dev_img, host_img, (size_t)B * 3 * H * W * 2, cudaMemcpyHostToDevice, s_in)) != cudaSuccess) return fail("H2D copy", e);
  if ((e = cudaEventRecord(ev_h2d, s_in)) != cudaSuccess) return fail("cudaEventRecord", e);
  *a_h2d = true;
  // 2. forward on the compute stream, after that copy and after the previous D2H out of this depth buffer
  if ((e = cudaStreamWaitEvent(s, ev_h2d, 0)) != cudaSuccess) return fail("cudaStreamWaitEvent", e);
  if (*a_copied && (e = cudaStreamWaitEvent(s, ev_copied, 0)) != cudaSuccess) return fail("cudaStreamWaitEvent", e);
  const int rc = dpt_forward(h, dev_img, dev_depth, ws, ws_bytes, B, H, W, stream);
  if (rc != DPT_OK) return rc;
  if ((e = cudaEventRecord(ev_read, s)) != cudaSuccess) return fail("cudaEventRecord", e);
  *a_read = true;
  // 3. D2H on the copy-out stream, after the forward
  if ((e = cudaStreamWaitEvent(s_out, ev_read, 0)) != cudaSuccess) return fail("cudaStreamWaitEvent", e);
  if ((e = cudaMemcpyAsync(host_depth, dev_depth, (size_t)B * H * W * 2, cudaMemcpyDeviceToHost, s_out)) != cudaSuccess) return fail("D2H copy", e);
  if ((e = cudaEventRecord(ev_copied, s_out)) != cudaSuccess) return fail("cudaEventRecord", e);
  *a_copied = true;
  return DPT_OK;
}

int dpt_patch_embed(dpt_handle h, const void* img, void* tokens, void* ws, size_t ws_bytes, int B, int H, int W,
                    void* stream) {
  return build_and_run(h, ws, ws_bytes, stream, [&](Ctx& c) { return build_patch_embed(c, img, tokens, B, H, W); });
}
int dpt_encoder(dpt_handle h, const void* tokens, void* const taps[4], void* ws, size_t ws_bytes, int B, int gh, int gw,
                void* stream) {
  return build_and_run(h, ws, ws_bytes, stream, [&](Ctx& c) { return stage_encoder(c, tokens, taps, B, gh, gw); });
}
int dpt_reassemble(dpt_handle h, const void* const taps[4], void* const maps[4], void* ws, size_t ws_bytes, int B,
                   int gh, int gw, void* stream) {
  return build_and_run(h, ws, ws_bytes, stream,
                       [&](Ctx& c) { return stage_reassemble(c, taps, maps, nullptr, B, gh, gw); });
}
int dpt_fusion(dpt_handle h, const void* const maps[4], void* fused, void* ws, size_t ws_bytes, int B, int gh, int gw,
               void* stream) {
  return build_and_run(h, ws, ws_bytes, stream,
                       [&](Ctx& c) {
                         const int k = map0_scale_num(c.m->cfg);
                         return build_fusion(c, maps, nullptr, fused, B, gh * k, gw * k);
                       });
}
int dpt_fusion_block(dpt_handle h, int level, const void* reasm_map, const void* prev_fused, void* out, void* ws,
                     size_t ws_bytes, int B, int map_h, int map_w, void* stream) {
  return build_and_run(h, ws, ws_bytes, stream, [&](Ctx& c) {
    return build_fusion_level(c, level, reasm_map, nullptr, prev_fused, out, B, map_h, map_w);
  });
}
int dpt_encoder_capture(dpt_handle h, const void* tokens, void* const taps[4], void* const* probs, void* const* block_out,
                        int num_blocks, void* ws, size_t ws_bytes, int B, int gh, int gw, void* stream) {
  return build_and_run(h, ws, ws_bytes, stream, [&](Ctx& c) {
    c.cap_probs = probs;
    c.cap_block_out = block_out;
    c.cap_blocks = num_blocks;
    return stage_encoder(c, tokens, taps, B, gh, gw);
  });
}
int dpt_head(dpt_handle h, const void* fused, void* depth, void* ws, size_t ws_bytes, int B, int gh, int gw,
             void* stream) {
  return build_and_run(h, ws, ws_bytes, stream, [&](Ctx& c) {
    const int k = 2 * map0_scale_num(c.m->cfg);
    return build_head(c, fused, depth, B, gh * k, gw * k);
  });
}

// ------------------------------------------------------- single operators ---------------------------------------

static int run_op(Ctx& c, std::vector<LaunchFn>& launches, void* stream) {
  if (!c.ok) { g_err = c.err; return DPT_ERR_INVALID; }
  std::string err;
  int rc = run_launches(nullptr, launches, (cudaStream_t)stream, err);
  if (rc != DPT_OK) g_err = err;
  return rc;
}

int dpt_op_conv_gemm(const void* A, const void* Wt, const float* bias, void* out, const void* add1, const void* add2,
                     void* out_relu, int B, int H, int W, int C, int N, int taps, int xoff, int act, int out_f32,
                     int dtype, void* stream) {
  std::vector<LaunchFn> launches;
  Ctx c = make_ctx(nullptr, nullptr, 0, &launches, false);
  c.is_bf16 = dtype == DPT_BF16;
  int dev = 0, nsm = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  c.num_sms = nsm;
  GemmOp op;
  op.A = A; op.B = B; op.Ht = H; op.Wt = W; op.C = C; op.xoff = xoff;
  op.Wt_ptr = Wt; op.N = N; op.taps = taps;
  op.bias = bias; op.act = act; op.out_kind = out_f32 ? OUT_F32 : OUT_HALF;
  op.out = out; op.add1 = add1; op.add2 = add2; op.out2_relu = out_relu;
  add_gemm(c, op);
  return run_op(c, launches, stream);
}

int dpt_op_attention(const void* qkv, const void* bias, int64_t bias_ld, int bias_wmod, void* out, int B, int N,
                     int heads, int head_dim, float scale, int dtype, void* stream) {
  std::vector<LaunchFn> launches;
  Ctx c = make_ctx(nullptr, nullptr, 0, &launches, false);
  c.is_bf16 = dtype == DPT_BF16;
  add_attention(c, qkv, bias, bias_ld, bias_wmod, out, B, N, heads, head_dim, scale);
  return run_op(c, launches, stream);
}

int dpt_op_layernorm(const float* x, const float* w, const float* b, void* y, int64_t M, int F, float eps, int dtype,
                     void* stream) {
  std::vector<LaunchFn> launches;
  Ctx c = make_ctx(nullptr, nullptr, 0, &launches, false);
  c.is_bf16 = dtype == DPT_BF16;
  add_layernorm(c, x, w, b, y, M, F, eps);
  return run_op(c, launches, stream);
}

int dpt_op_resize_bilinear(const void* in, void* out, int B, int IH, int IW, int OH, int OW, int C, int dtype,
                           void* stream) {
  std::vector<LaunchFn> launches;
  Ctx c = make_ctx(nullptr, nullptr, 0, &launches, false);
  c.is_bf16 = dtype == DPT_BF16;
  add_resize(c, in, out, B, IH, IW, OH, OW, C);
  return run_op(c, launches, stream);
}

int dpt_prepare_image(const uint8_t* bgr_hwc, int IH, int IW, void* out_chw, int OH, int OW, const float* mean_rgb,
                      const float* inv_std_rgb, int dtype, void* stream) {
  if (!bgr_hwc || !out_chw || !mean_rgb || !inv_std_rgb || IH <= 0 || IW <= 0 || OH <= 0 || OW <= 0 ||
      (dtype != DPT_BF16 && dtype != DPT_F16)) {
    g_err = "dpt_prepare_image: bad argument";
    return DPT_ERR_INVALID;
  }
  const float3 mean = make_float3(mean_rgb[0], mean_rgb[1], mean_rgb[2]);
  const float3 istd = make_float3(inv_std_rgb[0], inv_std_rgb[1], inv_std_rgb[2]);
  const dim3 grid((unsigned)((OW + 127) / 128), (unsigned)OH);
  cudaStream_t s = (cudaStream_t)stream;
  if (dtype == DPT_BF16) prepare_image_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(bgr_hwc, (__nv_bfloat16*)out_chw, IH, IW, OH, OW, mean, istd);
  else prepare_image_kernel<__half><<<grid, 128, 0, s>>>(bgr_hwc, (__half*)out_chw, IH, IW, OH, OW, mean, istd);
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_err = std::string("dpt_prepare_image: ") + cudaGetErrorString(e);
    return DPT_ERR_CUDA;
  }
  return DPT_OK;
}

int dpt_postprocess_u8(const void* depth_bhw, int B, int H, int W, uint8_t* out_u8, int OH, int OW, float* minmax,
                       int dtype, void* stream) {
  if (!depth_bhw || !out_u8 || !minmax || B <= 0 || H <= 0 || W <= 0 || OH <= 0 || OW <= 0 ||
      (dtype != DPT_BF16 && dtype != DPT_F16)) {
    g_err = "dpt_postprocess_u8: bad argument";
    return DPT_ERR_INVALID;
  }
  cudaStream_t s = (cudaStream_t)stream;
  int dev = 0, nsm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev);
  const long long total = (long long)B * OH * OW;
  const int grid = (int)std::max<long long>(1, std::min<long long>((total + 255) / 256, (long long)nsm * 8));
  minmax_init_kernel<<<1, 1, 0, s>>>(minmax);
  if (dtype == DPT_BF16) {
    post_minmax_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)depth_bhw, minmax, B, H, W, OH, OW);
    post_u8_kernel<__nv_bfloat16><<<grid, 256, 0, s>>>((const __nv_bfloat16*)depth_bhw, minmax, out_u8, B, H, W, OH, OW);
  } else {
    post_minmax_kernel<__half><<<grid, 256, 0, s>>>((const __half*)depth_bhw, minmax, B, H, W, OH, OW);
    post_u8_kernel<__half><<<grid, 256, 0, s>>>((const __half*)depth_bhw, minmax, out_u8, B, H, W, OH, OW);
  }
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_err = std::string("dpt_postprocess_u8: ") + cudaGetErrorString(e);
    return DPT_ERR_CUDA;
  }
  return DPT_OK;
}

int dpt_allgather_depth(void* nccl_comm, const void* local_depth, void* global_depth, size_t elems_per_rank, int dtype,
                        void* stream) {
  if (!nccl_comm || !local_depth || !global_depth || elems_per_rank == 0 || (dtype != DPT_BF16 && dtype != DPT_F16)) {
    g_err = "dpt_allgather_depth: bad argument";
    return DPT_ERR_INVALID;
  }
  // ncclResult_t ncclAllGather(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t)
  using AllGatherFn = int (*)(const void*, void*, size_t, int, void*, cudaStream_t);
  static AllGatherFn fn = nullptr;
  if (!fn) {
    void* sym = dlsym(RTLD_DEFAULT, "ncclAllGather");  // the NCCL the host process already loaded (e.g. PyTorch's)
    if (!sym) {
      void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
      if (h) sym = dlsym(h, "ncclAllGather");
    }
    fn = reinterpret_cast<AllGatherFn>(sym);
  }
  if (!fn) {
    g_err = "dpt_allgather_depth: no NCCL in this process (ncclAllGather not found)";
    return DPT_ERR_UNSUPPORTED;
  }
  const int nccl_dtype = dtype == DPT_BF16 ? 9 /* ncclBfloat16 */ : 6 /* ncclFloat16 */;
  const int rc = fn(local_depth, global_depth, elems_per_rank, nccl_dtype, nccl_comm, (cudaStream_t)stream);
  if (rc != 0) {
    g_err = "dpt_allgather_depth: ncclAllGather failed with ncclResult_t " + std::to_string(rc);
    return DPT_ERR_CUDA;
  }
  return DPT_OK;
}

}  // extern "C"

#ifdef ATT_TRACE
// trace build only (tools/attn_trace.py): copies the per-phase clock stamps of attn_tc_kernel to the host
extern "C" int dpt_debug_attn_trace(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, dpt::g_att_trace, sizeof(long long) * 4 * 16 * 10);
}
#endif

#ifdef GEMM_TRACE
// trace build only (tools/gemm_trace.py): copies the clock stamps of gemm_tc_kernel's epilogue warps / MMA thread
extern "C" int dpt_debug_gemm_trace(long long* host_out) {
  return (int)cudaMemcpyFromSymbol(host_out, dpt::g_gemm_trace, sizeof(long long) * 9 * 8 * 16);
}
#endif

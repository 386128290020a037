// libdrs.so -- C-ABI entry points (include/drs.h) and the orchestration of the four dilated nets.
// Reference graph: dilated_icpr_original / dilated_grsl / dilated_icpr_rate6_densely / dilated_grsl_rate8
// (/root/reference/isprs_dilated_random.py:761-788, 962-993, 914-959, 996-1033), loss_def (1089-1099),
// MomentumOptimizer (1685-1687).  No CPU fallback: every compute entry point needs an sm_100 device.
#include "conv_simt.cuh"
#include "conv_tc.cuh"
#include "drs_common.cuh"
#include "ops.cuh"
#include "scene.cuh"
#include "wgrad_tc.cuh"
#include "conv1_tc.cuh"
#include "npz_io.h"

thread_local char g_drs_err[1024] = {0};

#define API_BEGIN try {
#define API_END                                                        \
  }                                                                    \
  catch (const DrsError& e) { return e.code; }                         \
  catch (const std::exception& e) {                                    \
    snprintf(g_drs_err, sizeof(g_drs_err), "exception: %s", e.what()); \
    return 2;                                                          \
  }                                                                    \
  return 0;

static inline unsigned nblk(int64_t n, int t) { return (unsigned)ceil_div(n, t); }
#include "variants.cuh"
static inline int act_type(const Handle* h) { return h->cfg.precision == DRS_PREC_FP32 ? ET_F32 : (h->cfg.precision == DRS_PREC_F16 ? ET_F16 : ET_BF16); }

// ------------------------------------------------------------------------------------------------
// network description
// ------------------------------------------------------------------------------------------------
struct ConvSpec { int k, rate, co; };
static void build_net(NetDesc& n, const drs_config& cfg) {
  static const ConvSpec d6[] = {{5, 1, 64}, {5, 1, 64}, {4, 2, 128}, {4, 2, 128}, {3, 4, 256}, {3, 4, 256}};
  static const ConvSpec d6p[] = {{5, 1, 64}, {5, 2, 64}, {4, 3, 128}, {4, 4, 128}, {3, 5, 256}, {3, 6, 256}};
  static const ConvSpec dd6[] = {{5, 1, 32}, {5, 2, 32}, {4, 3, 64}, {4, 4, 64}, {3, 5, 128}, {3, 6, 128}};
  static const ConvSpec d8p[] = {{5, 1, 64}, {5, 2, 64}, {4, 3, 128}, {4, 4, 128}, {3, 5, 192}, {3, 6, 192}, {3, 7, 256}, {3, 8, 256}};
  static const ConvSpec r6[] = {{5, 1, 64}, {5, 2, 64}, {4, 3, 128}, {4, 4, 128}, {3, 5, 256}, {3, 6, 256}};
  static const ConvSpec r6s[] = {{5, 1, 64}, {5, 2, 64}, {4, 3, 64}, {4, 4, 128}, {3, 5, 128}, {3, 6, 128}};
  static const ConvSpec r6n[] = {{5, 1, 64}, {5, 1, 64}, {4, 1, 128}, {4, 1, 128}, {3, 1, 256}, {3, 1, 256}};
  static const ConvSpec r1[] = {{5, 1, 64}, {5, 1, 64}, {4, 1, 128}, {4, 1, 128}, {3, 1, 256}, {3, 1, 256}};    // coffee:788-813
  static const ConvSpec vr[] = {{5, 1, 64}, {5, 2, 64}, {4, 4, 128}, {4, 1, 128}, {3, 2, 256}, {3, 4, 256}};    // coffee:816-841
  static const ConvSpec old3[] = {{5, 1, 64}, {4, 2, 128}, {3, 4, 256}};                                         // contest:574-603
  static const int old3_scope[] = {1, 3, 5};        // the reference kept the scope numbers of the layers it commented out
  const int* scope_no = nullptr;
  const ConvSpec* sp = nullptr;
  int L = 0;
  n.net_type = cfg.net_type;
  n.channels = cfg.channels;
  n.classes = cfg.num_classes;
  n.pool = n.dense = false;
  n.squeeze = false;
  n.se.clear();
  const char* prefix = "conv";
  switch (cfg.net_type) {
    case DRS_NET_DILATED6: sp = d6; L = 6; n.act = ACT_RELU; if (cfg.isprs_scopes) prefix = "main_conv"; break;
    case DRS_NET_DILATED6_POOLING: sp = d6p; L = 6; n.act = ACT_LRELU; n.pool = true; break;
    case DRS_NET_DENSE_DILATED6: sp = dd6; L = 6; n.act = ACT_RELU; n.dense = true; break;
    case DRS_NET_DILATED8_POOLING: sp = d8p; L = 8; n.act = ACT_LRELU; n.pool = true; break;
    case DRS_NET_RATE6: sp = r6; L = 6; n.act = ACT_RELU; break;
    case DRS_NET_RATE6_SMALL: sp = r6s; L = 6; n.act = ACT_RELU; break;
    case DRS_NET_RATE6_NODILATION: sp = r6n; L = 6; n.act = ACT_RELU; break;
    case DRS_NET_RATE1: sp = r1; L = 6; n.act = ACT_RELU; break;
    case DRS_NET_VARY_RATE: sp = vr; L = 6; n.act = ACT_RELU; break;
    case DRS_NET_ICPR_OLD: sp = old3; L = 3; n.act = ACT_RELU; scope_no = old3_scope; break;
    case DRS_NET_RATE6_AVGPOOL: sp = r6; L = 6; n.act = ACT_RELU; break;
    case DRS_NET_RATE6_SE: sp = r6; L = 6; n.act = ACT_RELU; break;
    case DRS_NET_RATE6_SQUEEZE: sp = r6; L = 1; n.act = ACT_RELU; n.squeeze = true; break;
    default: DRS_FAIL("Error! Net type not identified: %d", cfg.net_type);
  }
  DRS_CHECK(cfg.channels >= 1 && cfg.channels <= 16, "channels=%d out of range", cfg.channels);
  DRS_CHECK(cfg.num_classes >= 2 && cfg.num_classes <= MAX_CLASSES, "num_classes=%d out of range [2,%d]", cfg.num_classes, MAX_CLASSES);
  int cin = cfg.channels;
  int64_t off = 0, boff = 0;
  int feat = 0;   // dense: running width of the concat buffer
  n.convs.clear();
  for (int i = 0; i < L; ++i) {
    ConvLayer c;
    c.scope = std::string(prefix) + std::to_string(scope_no ? scope_no[i] : i + 1);
    c.k = sp[i].k; c.rate = sp[i].rate; c.ci = cin; c.co = sp[i].co;
    const int total = (c.k - 1) * c.rate;
    c.pad_b = total / 2; c.pad_a = total - c.pad_b;
    c.in_coff = 0;
    c.out_coff = n.dense ? feat : 0;
    c.w_off = off; off += (int64_t)c.k * c.k * c.ci * c.co;
    c.b_off = off; off += c.co;
    c.mm_off = boff; boff += c.co;
    c.mv_off = boff; boff += c.co;
    if (cfg.net_type == DRS_NET_RATE6_AVGPOOL && i < 5) { c.post = 1; c.post_k = i < 3 ? 5 : 7; }     // isprs:824-837
    if (cfg.net_type == DRS_NET_RATE6_SE && (i & 1)) {                                               // isprs:1042-1051
      SeBlock sb;
      sb.name = "se" + std::to_string(i / 2 + 1);
      sb.c = c.co; sb.r = c.co / 4;
      sb.w1_off = off; off += (int64_t)sb.c * sb.r;
      sb.b1_off = off; off += sb.r;
      sb.w2_off = off; off += (int64_t)sb.r * sb.c;
      sb.b2_off = off; off += sb.c;
      c.post = 2; c.se = (int)n.se.size();
      n.se.push_back(sb);
    }
    n.convs.push_back(c);
    if (n.dense) { feat += c.co; cin = feat; } else cin = c.co;
  }
  if (n.squeeze) {
    // _squeeze_conv_layer (isprs:726-742, 1068-1075): name, in, out, k_dim, kernel, rate
    struct Sq { const char* name; int in, out, kd, ks, rate; };
    static const Sq sq[] = {{"conv2", 64, 64, 32, 5, 2}, {"conv3", 64, 128, 64, 4, 3}, {"conv4", 128, 128, 64, 4, 4},
                            {"conv5", 128, 256, 64, 3, 5}, {"conv6", 256, 256, 128, 3, 6}};
    int prev_group = 0, prev_cs = n.convs[0].co;
    for (const Sq& q : sq) {
      const int s1 = (int)n.convs.size();
      for (int part = 0; part < 3; ++part) {
        ConvLayer c;
        c.scope = std::string(q.name) + (part == 0 ? "_s1" : part == 1 ? "_s2_1" : "_s2_2");
        c.k = part == 2 ? q.ks : 1;
        c.rate = q.rate;
        c.ci = part == 0 ? q.in : q.kd;
        c.co = part == 0 ? q.kd : q.out / 2;
        const int total = (c.k - 1) * c.rate;
        c.pad_b = total / 2; c.pad_a = total - c.pad_b;
        c.in_coff = 0;
        c.in_group = part == 0 ? prev_group : s1;
        c.in_cs = part == 0 ? prev_cs : q.kd;
        c.out_group = part == 0 ? s1 : s1 + 1;
        c.out_cs = part == 0 ? q.kd : q.out;
        c.out_coff = part == 2 ? q.out / 2 : 0;
        c.w_off = off; off += (int64_t)c.k * c.k * c.ci * c.co;
        c.b_off = off; off += c.co;
        c.mm_off = boff; boff += c.co;
        c.mv_off = boff; boff += c.co;
        n.convs.push_back(c);
      }
      prev_group = s1 + 1;
      prev_cs = q.out;
      cin = q.out;
    }
  }
  n.cls_in = cin;
  n.cls_w_off = off; off += (int64_t)cin * cfg.num_classes;
  n.cls_b_off = off; off += cfg.num_classes;
  n.n_trainable = off;
  n.n_bnstat = boff;
  n.feat_stride = 0;
  for (auto& c : n.convs) n.feat_stride = std::max(n.feat_stride, std::max(c.co, c.out_cs));
  if (n.dense) n.feat_stride = feat;
}

// ------------------------------------------------------------------------------------------------
// life cycle
// ------------------------------------------------------------------------------------------------
extern "C" const char* drs_last_error(void) { return g_drs_err; }
extern "C" int drs_version(void) { return DRS_VERSION; }

struct InferLane {
  cudaStream_t stream = nullptr;
  Arena arena;
  float* x = nullptr;
  float* lg = nullptr;
  size_t x_cap = 0, lg_cap = 0;
  cudaEvent_t fwd_done = nullptr, acc_done = nullptr;
};

// One captured training step (CUDA graph) per (batch, patch size, buffers, learning rate): the step is ~65 small
// dependent launches, so replaying it as a graph removes the launch gaps between them.
struct TrainGraphKey {
  int B, crop, ignore_label, profiling;
  const void *x, *y, *mask, *acc_mask, *pred, *cm;
  uint64_t lr_bits;
  uint64_t arena_epoch;
  bool operator<(const TrainGraphKey& o) const { return memcmp(this, &o, sizeof(*this)) < 0; }
};
struct TrainGraph {
  cudaGraphExec_t exec = nullptr;
  int64_t launches = 0, conv_launches = 0;
  double conv_flops = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;   // external event-record nodes around the tensor-core launches
};

constexpr int RESULT_STRIDE = 64 + MAX_CLASSES * MAX_CLASSES;   // floats per pinned result slot: loss, then K*K+1 counts
struct HandleExtra {
  // persistent scene-pass buffers (score map, occurrence counts, cell tables, label map ...): cudaMalloc / cudaFree of
  // gigabyte-sized buffers per call costs hundreds of milliseconds, so each named slot only ever grows
  // the label map of the last scene pass stays in slot 4 for drs_scene_confusion
  int last_scene = -1, last_row_begin = 0, last_rows = 0, last_W = 0;
  void* slot_ptr[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  size_t slot_cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  std::map<TrainGraphKey, TrainGraph> graphs;
  std::map<TrainGraphKey, int> graph_seen;
  TrainGraph* capturing = nullptr;   // non-null while a training step is being captured
  cudaStream_t side_stream = nullptr;     // filter gradients run here, overlapped with the backward's HBM-bound kernels
  cudaEvent_t ev_dz[2] = {nullptr, nullptr}, ev_wgrad[2] = {nullptr, nullptr};
  // data-parallel exchange of the late layers' gradients, overlapped with the rest of the backward
  cudaStream_t comm_stream = nullptr;
  cudaEvent_t ev_bucket_ready = nullptr, ev_bucket_done = nullptr;
  cudaEvent_t ev_x8 = nullptr;          // conv1's padded input is ready: the side stream builds the im2col matrix under the forward
  bool use_graphs = false;          // opt-in (DRS_GRAPHS=1): measured 6-9 % per step in steady state, see DESIGN.md
  double conv_ms_acc = 0;            // device time of profiled launches already read back
  InferLane lanes[2];             // scene-inference lanes (drs_scene_api.cuh)
  cudaEvent_t lanes_ready = nullptr;
  // streamed scene upload (drs_scene_infer_host): copy stream + "rows uploaded" event
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t up_ev[1] = {nullptr};
  uint8_t* is_weight = nullptr;   // [n_trainable] 1 for `weights` variables
  SceneTable table;               // host copy of the device scene table
  // training scratch kept between calls
  float* mean = nullptr;          // [sum co] batch mean of the last train step (per layer, same offsets as mm_off/2)
  float* inv_std = nullptr;
  float* sums = nullptr;          // [2*256]
  float* loss_dev = nullptr;      // [4]: ce, l2, count, spare
  unsigned int* cm_dev = nullptr; // [K*K+1]
  unsigned int* count_dev = nullptr;
  unsigned int* bn_counter = nullptr;   // last-block-done counter of bn_partial_kernel (self-resetting)
  long long* bn_acc = nullptr;          // [BN_ACC_REPLICAS][2*512] fixed-point accumulators of bn_partial_kernel (self-clearing)
  float* ones = nullptr;          // [512]
  float* zeros = nullptr;         // [512]
  float* cls_w_eval = nullptr;
  // profiling
  double conv_flops = 0;
  int64_t conv_launches = 0;
  // pipelined training loop (drs_gather_plan_dev / drs_train_step_async / drs_train_result): the plan of step i+1 is
  // uploaded on plan_stream into a two-slot device staging ring while step i runs; loss + confusion counts of each step
  // land in a ring of pinned result slots behind an event, so the host never blocks on the step it has just enqueued
  void* nccl = nullptr;             // ncclComm_t of drs_comm_init (drs_comm.cuh); null: single process or callback exchange
  int rank = 0;
  cudaStream_t plan_stream = nullptr;
  void* plan_stage[2] = {nullptr, nullptr};
  size_t plan_stage_cap[2] = {0, 0};
  cudaEvent_t plan_uploaded[2] = {nullptr, nullptr}, plan_consumed[2] = {nullptr, nullptr};
  bool plan_consumed_valid[2] = {false, false};
  int64_t plan_calls = 0;
  static constexpr int RESULT_RING = 8;
  float* result_host = nullptr;     // pinned [RESULT_RING][RESULT_STRIDE]
  cudaEvent_t result_ev[RESULT_RING] = {};
  int64_t next_ticket = 0;
  // DRS_DEBUG_KEEP=1: fp32 copies of backward intermediates ("dz:<scope>", "da:<scope>")
  std::vector<float*> keep;
  bool debug_keep = false;
};
static std::map<Handle*, HandleExtra*> g_extra;
static HandleExtra* X(Handle* h) { return g_extra[h]; }
static void* slot_buf(Handle* h, int slot, size_t bytes) {
  HandleExtra* x = X(h);
  if (bytes == 0) bytes = 16;
  if (x->slot_cap[slot] < bytes) {
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    if (x->slot_ptr[slot]) CUDA_CHECK(cudaFree(x->slot_ptr[slot]));
    x->slot_ptr[slot] = nullptr;
    x->slot_cap[slot] = 0;
    CUDA_CHECK(cudaMalloc(&x->slot_ptr[slot], bytes));
    x->slot_cap[slot] = bytes;
  }
  return x->slot_ptr[slot];
}
static void lanes_release(Handle* h);
static void pass_geometry_release(Handle* h);

static void free_packed(Handle* h) {
  for (auto& c : h->net.convs) {
    if (c.w_fprop) cudaFree(c.w_fprop);
    if (c.w_dgrad) cudaFree(c.w_dgrad);
    if (c.fold_scale) cudaFree(c.fold_scale);
    if (c.fold_shift) cudaFree(c.fold_shift);
    c.w_fprop = c.w_dgrad = nullptr;
    c.fold_scale = c.fold_shift = nullptr;
  }
}

extern "C" int drs_create(drs_handle_t* out, const drs_config* cfg) {
  API_BEGIN
  DRS_CHECK(out && cfg, "drs_create: null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  DRS_CHECK(e == cudaSuccess && ndev > 0, "drs_create: no CUDA device (%s); libdrs has no CPU fallback", cudaGetErrorString(e));
  DRS_CHECK(cfg->device >= 0 && cfg->device < ndev, "drs_create: device %d out of range (%d devices)", cfg->device, ndev);
  CUDA_CHECK(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  CUDA_CHECK(cudaGetDeviceProperties(&prop, cfg->device));
  DRS_CHECK(prop.major == 10, "drs_create: device %s is sm_%d%d; libdrs is built for sm_100a only", prop.name, prop.major, prop.minor);
  Handle* h = new Handle();
  HandleExtra* x = new HandleExtra();
  g_extra[h] = x;
  h->cfg = *cfg;
  h->sm_count = prop.multiProcessorCount;
  CUDA_CHECK(cudaDriverGetVersion(&h->driver_version));
  build_net(h->net, *cfg);
  CUDA_CHECK(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
  h->stream = h->own_stream;
  const int64_t nt = h->net.n_trainable;
  CUDA_CHECK(cudaMalloc(&h->params, nt * 4));
  CUDA_CHECK(cudaMalloc(&h->grads, (nt + 1024) * 4));
  CUDA_CHECK(cudaMalloc(&h->moms, nt * 4));
  CUDA_CHECK(cudaMalloc(&h->bnstat, h->net.n_bnstat * 4));
  CUDA_CHECK(cudaMemset(h->params, 0, nt * 4));
  CUDA_CHECK(cudaMemset(h->grads, 0, (nt + 1024) * 4));
  CUDA_CHECK(cudaMemset(h->moms, 0, nt * 4));
  CUDA_CHECK(cudaMalloc(&x->is_weight, nt));
  {
    std::vector<uint8_t> isw(nt, 0);
    for (auto& c : h->net.convs)
      for (int64_t i = c.w_off; i < c.b_off; ++i) isw[i] = 1;
    for (int64_t i = h->net.cls_w_off; i < h->net.cls_b_off; ++i) isw[i] = 1;
    for (auto& sb : h->net.se) {                         // _fc_layer weights carry weight decay too (isprs:668-670)
      for (int64_t i = sb.w1_off; i < sb.b1_off; ++i) isw[i] = 1;
      for (int64_t i = sb.w2_off; i < sb.b2_off; ++i) isw[i] = 1;
    }
    CUDA_CHECK(cudaMemcpy(x->is_weight, isw.data(), nt, cudaMemcpyHostToDevice));
    // moving_mean = 0, moving_variance = 1 (Appendix B.3)
    std::vector<float> bn(h->net.n_bnstat, 0.0f);
    for (auto& c : h->net.convs)
      for (int i = 0; i < c.co; ++i) bn[c.mv_off + i] = 1.0f;
    CUDA_CHECK(cudaMemcpy(h->bnstat, bn.data(), bn.size() * 4, cudaMemcpyHostToDevice));
  }
  CUDA_CHECK(cudaMalloc(&x->mean, h->net.n_bnstat * 4));
  CUDA_CHECK(cudaMalloc(&x->inv_std, h->net.n_bnstat * 4));
  CUDA_CHECK(cudaMalloc(&x->sums, 2 * 512 * 4));
  CUDA_CHECK(cudaMalloc(&x->loss_dev, 16 * 4));
  CUDA_CHECK(cudaMalloc(&x->cm_dev, (MAX_CLASSES * MAX_CLASSES + 1) * 4));
  CUDA_CHECK(cudaMalloc(&x->count_dev, 16));
  CUDA_CHECK(cudaMemset(x->count_dev, 0, 16));
  x->bn_counter = x->count_dev + 2;
  CUDA_CHECK(cudaMalloc(&x->bn_acc, (size_t)BN_ACC_REPLICAS * BN_ACC_STRIDE * 8));
  CUDA_CHECK(cudaMemset(x->bn_acc, 0, (size_t)BN_ACC_REPLICAS * BN_ACC_STRIDE * 8));
  CUDA_CHECK(cudaMalloc(&x->ones, 512 * 4));
  CUDA_CHECK(cudaMalloc(&x->zeros, 512 * 4));
  CUDA_CHECK(cudaMemset(x->zeros, 0, 512 * 4));
  {
    std::vector<float> o(512, 1.0f);
    CUDA_CHECK(cudaMemcpy(x->ones, o.data(), 512 * 4, cudaMemcpyHostToDevice));
  }
  memset(&x->table, 0, sizeof(x->table));
  CUDA_CHECK(cudaHostAlloc(&h->diag_host, 64, cudaHostAllocMapped));
  memset(h->diag_host, 0, 64);
  CUDA_CHECK(cudaHostGetDevicePointer((void**)&h->diag_dev, h->diag_host, 0));
  // driver entry points for tensor-map encoding (no link-time dependency on libcuda)
  {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    DRS_CHECK(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeTiled not available");
    h->encodeTiled = (PFN_encodeTiled)fn;
    fn = nullptr;
    CUDA_CHECK(cudaGetDriverEntryPoint("cuTensorMapEncodeIm2col", &fn, cudaEnableDefault, &qres));
    DRS_CHECK(fn && qres == cudaDriverEntryPointSuccess, "cuTensorMapEncodeIm2col not available");
    h->encodeIm2col = (PFN_encodeIm2col)fn;
  }
  CUDA_CHECK(cudaEventCreate(&h->ev_a));
  CUDA_CHECK(cudaEventCreate(&h->ev_b));
  h->packed_dirty = true;
  h->eval_dirty = true;
  // whole-step CUDA graphs per (batch, patch size, buffers, lr): on by default (DRS_GRAPHS=0 turns them off); the drop-in
  // loops pre-capture the patch-size interval at start-up (drs_train_prepare), like TF building its graph before the loop
  x->use_graphs = getenv("DRS_GRAPHS") == nullptr || atoi(getenv("DRS_GRAPHS")) != 0;
  CUDA_CHECK(cudaStreamCreateWithFlags(&x->plan_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    CUDA_CHECK(cudaEventCreateWithFlags(&x->plan_uploaded[i], cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&x->plan_consumed[i], cudaEventDisableTiming));
  }
  CUDA_CHECK(cudaHostAlloc((void**)&x->result_host, (size_t)HandleExtra::RESULT_RING * RESULT_STRIDE * 4, cudaHostAllocDefault));
  for (int i = 0; i < HandleExtra::RESULT_RING; ++i) CUDA_CHECK(cudaEventCreateWithFlags(&x->result_ev[i], cudaEventDisableTiming));
  CUDA_CHECK(cudaStreamCreateWithFlags(&x->side_stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaStreamCreateWithFlags(&x->comm_stream, cudaStreamNonBlocking));
  CUDA_CHECK(cudaEventCreateWithFlags(&x->ev_bucket_ready, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&x->ev_x8, cudaEventDisableTiming));
  CUDA_CHECK(cudaEventCreateWithFlags(&x->ev_bucket_done, cudaEventDisableTiming));
  for (int i = 0; i < 2; ++i) {
    CUDA_CHECK(cudaEventCreateWithFlags(&x->ev_dz[i], cudaEventDisableTiming));
    CUDA_CHECK(cudaEventCreateWithFlags(&x->ev_wgrad[i], cudaEventDisableTiming));
  }
  *out = h;
  API_END
}

extern "C" int drs_destroy(drs_handle_t h) {
  API_BEGIN
  if (!h) return 0;
  cudaSetDevice(h->cfg.device);
  cudaStreamSynchronize(h->stream);
  HandleExtra* x = X(h);
  if (x) {
    if (x->nccl) { drs_comm_destroy(h); }
    pass_geometry_release(h);
    if (x->side_stream) { cudaStreamSynchronize(x->side_stream); cudaStreamDestroy(x->side_stream); }
    if (x->comm_stream) { cudaStreamSynchronize(x->comm_stream); cudaStreamDestroy(x->comm_stream); }
    if (x->ev_bucket_ready) cudaEventDestroy(x->ev_bucket_ready);
    if (x->ev_x8) cudaEventDestroy(x->ev_x8);
    if (x->ev_bucket_done) cudaEventDestroy(x->ev_bucket_done);
    if (x->copy_stream) { cudaStreamSynchronize(x->copy_stream); cudaStreamDestroy(x->copy_stream); }
    if (x->plan_stream) { cudaStreamSynchronize(x->plan_stream); cudaStreamDestroy(x->plan_stream); }
    for (int i = 0; i < 2; ++i) {
      if (x->plan_uploaded[i]) cudaEventDestroy(x->plan_uploaded[i]);
      if (x->plan_consumed[i]) cudaEventDestroy(x->plan_consumed[i]);
      if (x->plan_stage[i]) cudaFree(x->plan_stage[i]);
    }
    for (int i = 0; i < HandleExtra::RESULT_RING; ++i)
      if (x->result_ev[i]) cudaEventDestroy(x->result_ev[i]);
    if (x->result_host) cudaFreeHost(x->result_host);
    if (x->up_ev[0]) cudaEventDestroy(x->up_ev[0]);
    for (int i = 0; i < 2; ++i) {
      if (x->ev_dz[i]) cudaEventDestroy(x->ev_dz[i]);
      if (x->ev_wgrad[i]) cudaEventDestroy(x->ev_wgrad[i]);
    }
    lanes_release(h);
    for (int i = 0; i < 8; ++i)
      if (x->slot_ptr[i]) cudaFree(x->slot_ptr[i]);
    for (auto& kv : x->graphs) {
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
      for (auto& pr : kv.second.events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
    }
    x->graphs.clear();
  }
  free_packed(h);
  for (auto& kv : h->scenes) {
    if (kv.second.data) cudaFree(kv.second.data);
    if (kv.second.labels) cudaFree(kv.second.labels);
  }
  cudaFree(h->params); cudaFree(h->grads); cudaFree(h->moms); cudaFree(h->bnstat);
  if (h->arena.base) cudaFree(h->arena.base);
  if (h->dstage) cudaFree(h->dstage);
  if (h->pinned) cudaFreeHost(h->pinned);
  if (h->diag_host) cudaFreeHost(h->diag_host);
  if (x) {
    cudaFree(x->is_weight); cudaFree(x->mean); cudaFree(x->inv_std); cudaFree(x->sums); cudaFree(x->loss_dev);
    cudaFree(x->cm_dev); cudaFree(x->count_dev); cudaFree(x->bn_acc); cudaFree(x->ones); cudaFree(x->zeros);
    delete x;
    g_extra.erase(h);
  }
  for (auto& pr : h->conv_events) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
  cudaEventDestroy(h->ev_a); cudaEventDestroy(h->ev_b);
  cudaStreamDestroy(h->own_stream);
  delete h;
  API_END
}

extern "C" int drs_set_stream(drs_handle_t h, void* s) {
  API_BEGIN
  DRS_CHECK(h, "null handle");
  // (void*)-1 selects the handle's own non-blocking stream; any other value is a cudaStream_t, 0 = the legacy default stream
  h->stream = (s == (void*)(intptr_t)-1) ? h->own_stream : (cudaStream_t)s;
  API_END
}
extern "C" int drs_synchronize(drs_handle_t h) {
  API_BEGIN
  DRS_CHECK(h, "null handle");
  cudaError_t e = cudaStreamSynchronize(h->stream);
  if (e != cudaSuccess) {
    DRS_FAIL("stream failed: %s (diag %08x blk %u thr %u par %u)", cudaGetErrorString(e), h->diag_host[0], h->diag_host[1],
             h->diag_host[2], h->diag_host[3]);
  }
  API_END
}
extern "C" int64_t drs_launch_count(drs_handle_t h) { return h ? h->launches : -1; }

// ------------------------------------------------------------------------------------------------
// variables
// ------------------------------------------------------------------------------------------------
struct VarRef { float* ptr; int64_t count; int kind; };  // kind 0 device float, 1 global_step
static bool find_var(Handle* h, const std::string& name, VarRef& r, bool grad = false) {
  float* base = grad ? h->grads : h->params;
  auto slot = [&](const std::string& nm, float* p, float* m, int64_t cnt) -> bool {
    if (name == nm) { r = {p, cnt, 0}; return true; }
    if (!grad && m && name == nm + "/Momentum") { r = {m, cnt, 0}; return true; }
    return false;
  };
  for (auto& c : h->net.convs) {
    const int64_t wc = (int64_t)c.k * c.k * c.ci * c.co;
    if (slot(c.scope + "/weights", base + c.w_off, h->moms + c.w_off, wc)) return true;
    if (slot(c.scope + "/biases", base + c.b_off, h->moms + c.b_off, c.co)) return true;
    if (!grad) {
      if (slot(c.scope + "/moving_mean", h->bnstat + c.mm_off, nullptr, c.co)) return true;
      if (slot(c.scope + "/moving_variance", h->bnstat + c.mv_off, nullptr, c.co)) return true;
    }
  }
  for (auto& sb : h->net.se) {
    if (slot(sb.name + "_fc1/weights", base + sb.w1_off, h->moms + sb.w1_off, (int64_t)sb.c * sb.r)) return true;
    if (slot(sb.name + "_fc1/biases", base + sb.b1_off, h->moms + sb.b1_off, sb.r)) return true;
    if (slot(sb.name + "_fc2/weights", base + sb.w2_off, h->moms + sb.w2_off, (int64_t)sb.r * sb.c)) return true;
    if (slot(sb.name + "_fc2/biases", base + sb.b2_off, h->moms + sb.b2_off, sb.c)) return true;
  }
  if (slot("conv_classifier/weights", base + h->net.cls_w_off, h->moms + h->net.cls_w_off, (int64_t)h->net.cls_in * h->net.classes)) return true;
  if (slot("conv_classifier/biases", base + h->net.cls_b_off, h->moms + h->net.cls_b_off, h->net.classes)) return true;
  if (!grad && (name == "global_step" || name == "main_global_step")) { r = {nullptr, 1, 1}; return true; }
  return false;
}
static void list_vars(Handle* h, std::vector<std::pair<std::string, int64_t>>& v) {
  for (auto& c : h->net.convs) {
    const int64_t wc = (int64_t)c.k * c.k * c.ci * c.co;
    v.push_back({c.scope + "/weights", wc});
    v.push_back({c.scope + "/biases", c.co});
    v.push_back({c.scope + "/moving_mean", c.co});
    v.push_back({c.scope + "/moving_variance", c.co});
    v.push_back({c.scope + "/weights/Momentum", wc});
    v.push_back({c.scope + "/biases/Momentum", c.co});
  }
  for (auto& sb : h->net.se) {
    const std::pair<std::string, int64_t> vs[4] = {{sb.name + "_fc1/weights", (int64_t)sb.c * sb.r}, {sb.name + "_fc1/biases", sb.r},
                                                    {sb.name + "_fc2/weights", (int64_t)sb.r * sb.c}, {sb.name + "_fc2/biases", sb.c}};
    for (auto& pr : vs) v.push_back(pr);
    for (auto& pr : vs) v.push_back({pr.first + "/Momentum", pr.second});
  }
  v.push_back({"conv_classifier/weights", (int64_t)h->net.cls_in * h->net.classes});
  v.push_back({"conv_classifier/biases", h->net.classes});
  v.push_back({"conv_classifier/weights/Momentum", (int64_t)h->net.cls_in * h->net.classes});
  v.push_back({"conv_classifier/biases/Momentum", h->net.classes});
  v.push_back({"global_step", 1});
}
extern "C" int drs_num_variables(drs_handle_t h) {
  if (!h) return -1;
  std::vector<std::pair<std::string, int64_t>> v;
  list_vars(h, v);
  return (int)v.size();
}
extern "C" int drs_variable_name(drs_handle_t h, int index, char* name_out, int cap, int64_t* count_out) {
  API_BEGIN
  DRS_CHECK(h, "null handle");
  std::vector<std::pair<std::string, int64_t>> v;
  list_vars(h, v);
  DRS_CHECK(index >= 0 && index < (int)v.size(), "variable index %d out of range", index);
  if (name_out && cap > 0) snprintf(name_out, cap, "%s", v[index].first.c_str());
  if (count_out) *count_out = v[index].second;
  API_END
}
extern "C" int drs_set_variable(drs_handle_t h, const char* name, const float* data, int64_t count) {
  API_BEGIN
  DRS_CHECK(h && name && data, "null argument");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  VarRef r;
  DRS_CHECK(find_var(h, name, r), "unknown variable '%s'", name);
  DRS_CHECK(r.count == count, "variable '%s' has %lld elements, got %lld", name, (long long)r.count, (long long)count);
  if (r.kind == 1) { h->global_step = (int64_t)data[0]; return 0; }
  CUDA_CHECK(cudaMemcpyAsync(r.ptr, data, count * 4, cudaMemcpyHostToDevice, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  h->packed_dirty = true;
  h->eval_dirty = true;
  API_END
}
extern "C" int drs_get_variable(drs_handle_t h, const char* name, float* data, int64_t count) {
  API_BEGIN
  DRS_CHECK(h && name && data, "null argument");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  VarRef r;
  DRS_CHECK(find_var(h, name, r), "unknown variable '%s'", name);
  DRS_CHECK(r.count == count, "variable '%s' has %lld elements, got %lld", name, (long long)r.count, (long long)count);
  if (r.kind == 1) { data[0] = (float)h->global_step; return 0; }
  CUDA_CHECK(cudaMemcpyAsync(data, r.ptr, count * 4, cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  API_END
}
extern "C" int drs_get_gradient(drs_handle_t h, const char* name, float* data, int64_t count) {
  API_BEGIN
  DRS_CHECK(h && name && data, "null argument");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  VarRef r;
  DRS_CHECK(find_var(h, name, r, true), "unknown trainable variable '%s'", name);
  DRS_CHECK(r.count == count, "variable '%s' has %lld elements, got %lld", name, (long long)r.count, (long long)count);
  CUDA_CHECK(cudaMemcpyAsync(data, r.ptr, count * 4, cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  API_END
}

// ------------------------------------------------------------------------------------------------
// checkpoints: tf.train.Saver.save / restore (isprs:1693-1717, 1797-1802) as an uncompressed .npz keyed by the TF variable
// names ('/' written as '__', which is what numpy.savez keyword arguments allow), shapes as TF holds them (HWIO filters)
// ------------------------------------------------------------------------------------------------
static std::vector<int64_t> var_shape(Handle* h, const std::string& full, int64_t count) {
  std::string name = full;
  const std::string mom = "/Momentum";
  if (name.size() > mom.size() && name.compare(name.size() - mom.size(), mom.size(), mom) == 0) name.resize(name.size() - mom.size());
  for (auto& c : h->net.convs)
    if (name == c.scope + "/weights") return {c.k, c.k, c.ci, c.co};
  for (auto& sb : h->net.se) {
    if (name == sb.name + "_fc1/weights") return {sb.c, sb.r};
    if (name == sb.name + "_fc2/weights") return {sb.r, sb.c};
  }
  if (name == "conv_classifier/weights") return {1, 1, h->net.cls_in, h->net.classes};
  return {count};
}
static std::string npz_key(std::string name) {
  for (size_t i = 0; (i = name.find('/', i)) != std::string::npos;) name.replace(i, 1, "__");
  return name;
}
static std::string npz_unkey(std::string key) {
  for (size_t i = 0; (i = key.find("__", i)) != std::string::npos; ++i) key.replace(i, 2, "/");
  return key;
}

extern "C" int drs_save(drs_handle_t h, const char* path) {
  API_BEGIN
  DRS_CHECK(h && path, "null argument");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  std::vector<std::pair<std::string, int64_t>> v;
  list_vars(h, v);
  std::vector<drs_npz::Array> arrays(v.size());
  for (size_t i = 0; i < v.size(); ++i) {
    VarRef r;
    DRS_CHECK(find_var(h, v[i].first, r), "unknown variable '%s'", v[i].first.c_str());
    arrays[i].name = npz_key(v[i].first);
    arrays[i].shape = var_shape(h, v[i].first, r.count);
    arrays[i].data.resize((size_t)r.count);
    if (r.kind == 1) arrays[i].data[0] = (float)h->global_step;
    else CUDA_CHECK(cudaMemcpyAsync(arrays[i].data.data(), r.ptr, r.count * 4, cudaMemcpyDeviceToHost, h->stream));
  }
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  const std::string e = drs_npz::write(path, arrays);
  DRS_CHECK(e.empty(), "%s", e.c_str());
  API_END
}

extern "C" int drs_load(drs_handle_t h, const char* path) {
  API_BEGIN
  DRS_CHECK(h && path, "null argument");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  std::vector<drs_npz::Array> arrays;
  const std::string e = drs_npz::read(path, arrays);
  DRS_CHECK(e.empty(), "%s", e.c_str());
  // validate everything before touching the model: a checkpoint of another net must not be applied half-way
  for (auto& a : arrays) {
    VarRef r;
    const std::string name = npz_unkey(a.name);
    DRS_CHECK(find_var(h, name, r), "checkpoint '%s': unknown variable '%s'", path, name.c_str());
    DRS_CHECK(r.count == a.count(), "checkpoint '%s': variable '%s' has %lld elements, the model expects %lld", path, name.c_str(),
              (long long)a.count(), (long long)r.count);
  }
  for (auto& a : arrays) {
    VarRef r;
    find_var(h, npz_unkey(a.name), r);
    if (r.kind == 1) { h->global_step = (int64_t)a.data[0]; continue; }
    CUDA_CHECK(cudaMemcpyAsync(r.ptr, a.data.data(), r.count * 4, cudaMemcpyHostToDevice, h->stream));
  }
  CUDA_CHECK(cudaStreamSynchronize(h->stream));      // the host arrays die with this scope
  h->packed_dirty = true;
  h->eval_dirty = true;
  API_END
}

// the container alone, no handle and no device: inspection / conversion tools and the CPU tests
extern "C" int drs_npz_write(const char* path, int32_t n, const char* const* names, const float* const* data, const int32_t* ndim,
                             const int64_t* dims) {
  API_BEGIN
  DRS_CHECK(path && n >= 0 && (n == 0 || (names && data && ndim && dims)), "bad argument");
  std::vector<drs_npz::Array> arrays((size_t)n);
  for (int i = 0; i < n; ++i) {
    DRS_CHECK(names[i] && data[i] && ndim[i] >= 0 && ndim[i] <= 4, "array %d: bad argument", i);
    arrays[i].name = names[i];
    arrays[i].shape.assign(dims + 4 * i, dims + 4 * i + ndim[i]);
    for (int64_t d : arrays[i].shape) DRS_CHECK(d >= 0, "array %d: negative dimension", i);
    arrays[i].data.assign(data[i], data[i] + arrays[i].count());
  }
  const std::string e = drs_npz::write(path, arrays);
  DRS_CHECK(e.empty(), "%s", e.c_str());
  API_END
}
extern "C" int drs_npz_entry(const char* path, int32_t index, char* name_out, int32_t name_cap, int32_t* ndim_out, int64_t* dims_out,
                             int64_t* count_out, int32_t* total_out) {
  API_BEGIN
  DRS_CHECK(path, "null argument");
  std::vector<drs_npz::Array> arrays;
  const std::string e = drs_npz::read(path, arrays);
  DRS_CHECK(e.empty(), "%s", e.c_str());
  if (total_out) *total_out = (int32_t)arrays.size();
  if (index < 0) return 0;                                  // count only
  DRS_CHECK(index < (int)arrays.size(), "entry %d out of range (%d arrays)", index, (int)arrays.size());
  const drs_npz::Array& a = arrays[index];
  DRS_CHECK(a.shape.size() <= 4, "array '%s' has %d dimensions", a.name.c_str(), (int)a.shape.size());
  if (name_out && name_cap > 0) snprintf(name_out, name_cap, "%s", a.name.c_str());
  if (ndim_out) *ndim_out = (int32_t)a.shape.size();
  if (dims_out) for (size_t i = 0; i < a.shape.size(); ++i) dims_out[i] = a.shape[i];
  if (count_out) *count_out = a.count();
  API_END
}
extern "C" int drs_npz_read(const char* path, const char* name, float* out, int64_t count) {
  API_BEGIN
  DRS_CHECK(path && name && out, "null argument");
  std::vector<drs_npz::Array> arrays;
  const std::string e = drs_npz::read(path, arrays);
  DRS_CHECK(e.empty(), "%s", e.c_str());
  for (auto& a : arrays)
    if (a.name == name) {
      DRS_CHECK(a.count() == count, "array '%s' has %lld elements, got room for %lld", name, (long long)a.count(), (long long)count);
      memcpy(out, a.data.data(), (size_t)count * 4);
      return 0;
    }
  DRS_FAIL("checkpoint '%s' has no array '%s'", path, name);
  API_END
}

// host-only: the units CTA `block` of a `grid`-CTA conv_tc launch works on, in order (ConvSched is the kernel's own code)
extern "C" int drs_debug_conv_schedule(int32_t num_units, int32_t mt, int32_t grid, int32_t block, int32_t* units_out,
                                       uint8_t* pair_out, int32_t cap, int32_t* n_out) {
  API_BEGIN
  DRS_CHECK(num_units >= 0 && (mt == 1 || mt == 2) && grid >= 1 && block >= 0 && block < grid && n_out, "bad argument");
  const ConvSched sc(num_units, mt, grid, block);
  *n_out = sc.iters;
  for (int i = 0; i < sc.iters && i < cap; ++i) {
    bool two;
    const int u = sc.unit(i, two);
    if (units_out) units_out[i] = u;
    if (pair_out) pair_out[i] = two ? 1 : 0;
  }
  API_END
}

extern "C" int drs_set_ignore_label(drs_handle_t h, int32_t label) {
  API_BEGIN
  DRS_CHECK(h, "null handle");
  h->ignore_label = label;
  API_END
}

extern "C" int drs_set_allreduce(drs_handle_t h, drs_allreduce_fn fn, void* user, int32_t world, int32_t sync_bn) {
  API_BEGIN
  DRS_CHECK(h, "null handle");
  h->allreduce = fn;
  h->allreduce_user = user;
  h->world = fn ? (world > 0 ? world : 1) : 1;
  h->sync_bn = fn ? sync_bn : 0;
  API_END
}
#include "drs_comm.cuh"

// ------------------------------------------------------------------------------------------------
// packed operands / folded BN refresh (after set_variable or an optimizer step)
// ------------------------------------------------------------------------------------------------
// One launch repacks the operand matrices of every tensor-core layer after set_variable / an optimizer step:
//   kind 0  fprop operand  Wf[co][tap*ci + c]            = W[tap][c][co]
//   kind 1  dgrad operand  Wd[c][tap'*co + o]            = W[taps-1-tap'][c][o]
//   kind 2  dgrad operand of the fp32 CUDA-core path  [tap'*co + o][c]
struct RepackSeg {
  const float* w;
  void* out;
  int taps, ci, co, kind;
  long long start;
};
struct RepackTable {
  RepackSeg seg[40];       // up to 20 tensor-core layers (the squeeze net has 15), fprop + dgrad operand each
  int n, etype;
  long long total;
};
__global__ void repack_kernel(const __grid_constant__ RepackTable t) {
  pdl_sync();
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= t.total) return;
  int si = 0;
  while (si + 1 < t.n && gid >= t.seg[si + 1].start) ++si;
  const RepackSeg& s = t.seg[si];
  const long long i = gid - s.start;
  float v;
  if (s.kind == 0) {
    const int o = (int)(i / ((long long)s.taps * s.ci));
    const int kk = (int)(i - (long long)o * s.taps * s.ci);
    v = s.w[(long long)kk * s.co + o];
  } else if (s.kind == 1) {
    const int c = (int)(i / ((long long)s.taps * s.co));
    const int r = (int)(i - (long long)c * s.taps * s.co);
    const int tp = r / s.co, o = r - tp * s.co;
    v = s.w[((long long)(s.taps - 1 - tp) * s.ci + c) * s.co + o];
  } else {
    const int c = (int)(i % s.ci);
    const int r = (int)(i / s.ci);
    const int tp = r / s.co, o = r - tp * s.co;
    v = s.w[((long long)(s.taps - 1 - tp) * s.ci + c) * s.co + o];
  }
  if (s.kind == 2 || t.etype == ET_F32) reinterpret_cast<float*>(s.out)[i] = v;
  else if (t.etype == ET_F16) reinterpret_cast<__half*>(s.out)[i] = __float2half_rn(v);
  else reinterpret_cast<__nv_bfloat16*>(s.out)[i] = __float2bfloat16_rn(v);
}

// training needs only the layer operands; inference additionally the folded eval-mode BN and conv1's tensor-core operand
static void refresh_packed(Handle* h, bool training) {
  const int et = act_type(h);
  if (h->packed_dirty) {
    RepackTable t;
    memset(&t, 0, sizeof(t));
    t.etype = et;
    long long off = 0;
    for (size_t l = 1; l < h->net.convs.size(); ++l) {
      ConvLayer& c = h->net.convs[l];
      const int taps = c.k * c.k;
      const long long n = (long long)taps * c.ci * c.co;
      const size_t es = et == ET_F32 ? 4 : 2;
      if (et != ET_F32) {
        if (!c.w_fprop) CUDA_CHECK(cudaMalloc(&c.w_fprop, n * es));
        t.seg[t.n++] = RepackSeg{h->params + c.w_off, c.w_fprop, taps, c.ci, c.co, 0, off};
        off += n;
      }
      if (!c.w_dgrad) CUDA_CHECK(cudaMalloc(&c.w_dgrad, n * es));
      t.seg[t.n++] = RepackSeg{h->params + c.w_off, c.w_dgrad, taps, c.ci, c.co, et == ET_F32 ? 2 : 1, off};
      off += n;
    }
    t.total = off;
    if (off > 0) {
      repack_kernel<<<nblk(off, 256), 256, 0, h->stream>>>(t);
      LAUNCH_CHECK(h);
    }
    h->packed_dirty = false;
  }
  if (!training && h->eval_dirty) {
    for (size_t l = 0; l < h->net.convs.size(); ++l) {
      ConvLayer& c = h->net.convs[l];
      if (!c.fold_scale) {
        CUDA_CHECK(cudaMalloc(&c.fold_scale, c.co * 4));
        CUDA_CHECK(cudaMalloc(&c.fold_shift, c.co * 4));
      }
      fold_bn_kernel<<<nblk(c.co, 128), 128, 0, h->stream>>>(h->params + c.b_off, h->bnstat + c.mm_off, h->bnstat + c.mv_off,
                                                              h->cfg.bn_eps, c.fold_scale, c.fold_shift, c.co);
      LAUNCH_CHECK(h);
      // 16-bit inference runs conv1 on the tensor cores (conv1_tc.cuh) from a core-matrix-ordered operand; training and the
      // fp32 mode run it on the CUDA-core kernel straight from the HWIO weights
      if (l == 0 && et != ET_F32 && conv1_tc_supported(c.k, c.rate, c.ci, c.co)) {
        const int n = C1_SLOTS * c.co * 8;
        if (!c.w_fprop) CUDA_CHECK(cudaMalloc(&c.w_fprop, (size_t)n * 2));
        if (et == ET_F16) pack_conv1_kernel<__half><<<nblk(n, 256), 256, 0, h->stream>>>(h->params + c.w_off, (__half*)c.w_fprop, c.ci, c.co);
        else pack_conv1_kernel<__nv_bfloat16><<<nblk(n, 256), 256, 0, h->stream>>>(h->params + c.w_off, (__nv_bfloat16*)c.w_fprop, c.ci, c.co);
        LAUNCH_CHECK(h);
      }
    }
    h->eval_dirty = false;
  }
}

// ------------------------------------------------------------------------------------------------
// one convolution, any precision
// ------------------------------------------------------------------------------------------------
struct ActBuf { void* p; int cs; int co; };   // pointer, channel stride, channel offset

static void prof_begin(Handle* h, cudaEvent_t* a, cudaEvent_t* b) {
  if (!h->time_convs) return;
  HandleExtra* x = X(h);
  if (x->capturing) {
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    x->capturing->events.push_back({e0, e1});
    *a = e0; *b = e1;
    CUDA_CHECK(cudaEventRecordWithFlags(e0, h->stream, cudaEventRecordExternal));
    return;
  }
  if (h->conv_events_used == h->conv_events.size()) {
    cudaEvent_t e0, e1;
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    h->conv_events.push_back({e0, e1});
  }
  *a = h->conv_events[h->conv_events_used].first;
  *b = h->conv_events[h->conv_events_used].second;
  h->conv_events_used++;
  CUDA_CHECK(cudaEventRecord(*a, h->stream));
}
static void prof_end(Handle* h, cudaEvent_t b) {
  if (!h->time_convs) return;
  if (X(h)->capturing) { CUDA_CHECK(cudaEventRecordWithFlags(b, h->stream, cudaEventRecordExternal)); return; }
  CUDA_CHECK(cudaEventRecord(b, h->stream));
}

// generic conv over activation type TA.  w_simt: HWIO fp32 [k*k*ci][co]; w_tc: packed [co][k*k*ci].
template <typename TA>
static void run_conv(Handle* h, const ActBuf& in, int ci, const float* w_simt, const void* w_tc, const ActBuf& out, int co,
                     int B, int crop, int k, int rate, int pad_b, const float* scale, const float* shift, int act,
                     const BnFinish* stats = nullptr) {
  if (ElemTag<TA>::v == ET_F32) {
    launch_conv_simt<TA, TA>(h, (const TA*)in.p, in.cs, in.co, ci, w_simt, (TA*)out.p, out.cs, out.co, co, B, crop, k, rate,
                             pad_b, scale, shift, act);
    return;
  }
  cudaEvent_t ea = nullptr, eb = nullptr;
  prof_begin(h, &ea, &eb);
  // tensor-core path; N > 256 is split into equal column tiles
  int nsplit = 1;
  while (co / nsplit > 256 || (co % nsplit) != 0 || ((co / nsplit) % 32) != 0) {
    ++nsplit;
    DRS_CHECK(nsplit <= 8, "cannot tile N=%d", co);
  }
  DRS_CHECK(!stats || nsplit == 1, "fused BN statistics need a single N tile (Co=%d)", co);
  const int nt = co / nsplit;
  for (int s = 0; s < nsplit; ++s) {
    ConvTcArgs a;
    a.in = in.p; a.in_cstride = in.cs; a.in_coff = in.co; a.ci = ci;
    a.w = (const char*)w_tc + (size_t)s * nt * k * k * ci * 2;
    a.out = out.p; a.out_cstride = out.cs; a.out_coff = out.co + s * nt; a.co = nt;
    a.B = B; a.crop = crop; a.k = k; a.rate = rate; a.pad_b = pad_b;
    a.scale = scale + s * nt; a.shift = shift + s * nt; a.act = act;
    a.etype = ElemTag<TA>::v;
    a.stats = stats;
    launch_conv_tc(h, a);
  }
  prof_end(h, eb);
  X(h)->conv_flops += 2.0 * (double)B * crop * crop * k * k * ci * co;
  X(h)->conv_launches += nsplit;
}

// ------------------------------------------------------------------------------------------------
// inference forward (eval-mode BN folded into the conv epilogue)
// ------------------------------------------------------------------------------------------------
static size_t forward_eval_workspace(Handle* h, int B, int crop) {
  const int64_t M = (int64_t)B * crop * crop;
  const size_t es = h->cfg.precision == DRS_PREC_FP32 ? 4 : 2;
  const int nbuf = h->net.dense ? 1 : 3;
  size_t se_bytes = 0;
  for (auto& sb : h->net.se) se_bytes = std::max(se_bytes, ((size_t)B * (3 * sb.c + sb.r)) * 4 + 4096);
  return nbuf * ((size_t)M * h->net.feat_stride * es + 4096) + (size_t)M * h->net.classes * 4 + (size_t)M * 32 * es + se_bytes + 65536;
}

template <typename TA>
static void forward_eval_t(Handle* h, const float* x_dev, int B, int crop, float* logits_dev, uint8_t* pred_dev) {
  NetDesc& n = h->net;
  const int64_t M = (int64_t)B * crop * crop;
  refresh_packed(h, false);
  const int fs = n.feat_stride;
  const size_t buf_bytes = (size_t)M * fs * sizeof(TA);
  const int nbuf = n.dense ? 1 : 3;
  ensure_arena(h, forward_eval_workspace(h, B, crop));
  h->arena.reset();
  TA* bufs[3] = {nullptr, nullptr, nullptr};
  for (int i = 0; i < nbuf; ++i) bufs[i] = (TA*)arena_take(h, buf_bytes);
  float* logits = logits_dev ? logits_dev : (float*)arena_take(h, (size_t)M * n.classes * 4);
  float* se_scratch = nullptr;
  {
    size_t se_floats = 0;
    for (auto& sb : n.se) se_floats = std::max(se_floats, (size_t)B * (3 * sb.c + sb.r));
    if (se_floats) se_scratch = (float*)arena_take(h, se_floats * 4);
  }
  h->taps.clear();

  ActBuf cur{nullptr, 0, 0};
  int xi = 0;   // index of the buffer holding the current input (non-dense)
  int sq_s = 0, sq_c = 0;   // squeeze modules: buffers of the squeeze output S and of the concatenated module output
  for (size_t l = 0; l < n.convs.size(); ++l) {
    ConvLayer& c = n.convs[l];
    ActBuf out;
    const int sq_part = (n.squeeze && l >= 1) ? (int)((l - 1) % 3) : -1;    // 0: _s1, 1: _s2_1, 2: _s2_2 (isprs:726-742)
    if (n.dense) out = {bufs[0], fs, c.out_coff};
    else if (sq_part == 0) { sq_s = (xi + 1) % 3; sq_c = (xi + 2) % 3; out = {bufs[sq_s], fs, 0}; }
    else if (sq_part >= 1) { cur = {bufs[sq_s], fs, 0}; out = {bufs[sq_c], fs, c.out_coff}; }
    else out = {bufs[(xi + 1) % 3], fs, 0};
    if (l == 0 && ElemTag<TA>::v != ET_F32 && c.w_fprop && conv1_tc_supported(c.k, c.rate, c.ci, c.co)) {
      TA* x8 = (TA*)arena_take(h, (size_t)M * 8 * sizeof(TA));
      pad_cast8_kernel<TA><<<nblk(M, 256), 256, 0, h->stream>>>(x_dev, x8, c.ci, M);
      LAUNCH_CHECK(h);
      cudaEvent_t ea = nullptr, eb = nullptr;
      prof_begin(h, &ea, &eb);
      Conv1TcArgs a1;
      a1.x8 = x8; a1.wpack = c.w_fprop; a1.out = out.p; a1.out_cstride = out.cs; a1.out_coff = out.co; a1.co = c.co;
      a1.B = B; a1.crop = crop; a1.scale = c.fold_scale; a1.shift = c.fold_shift; a1.act = n.act; a1.etype = ElemTag<TA>::v;
      launch_conv1_tc(h, a1);
      prof_end(h, eb);
      X(h)->conv_flops += 2.0 * (double)M * c.k * c.k * c.ci * c.co;   // algorithmic FLOPs: the real C channels
      X(h)->conv_launches += 1;
    } else if (l == 0) {
      launch_conv_simt<float, TA>(h, x_dev, n.channels, 0, n.channels, h->params + c.w_off, (TA*)out.p, out.cs, out.co, c.co,
                                  B, crop, c.k, c.rate, c.pad_b, c.fold_scale, c.fold_shift, n.act);
    } else {
      run_conv<TA>(h, cur, c.ci, h->params + c.w_off, c.w_fprop, out, c.co, B, crop, c.k, c.rate, c.pad_b, c.fold_scale,
                   c.fold_shift, n.act);
    }
    if (n.pool && l + 1 == n.convs.size() && c.co == n.cls_in && pool_classifier_fused_supported<TA>(c.co)) {
      // last layer: pool + classifier in one pass, the pooled activations are never stored (no activation tap for this layer)
      launch_maxpool3_classifier<TA>(h, (const TA*)out.p, out.cs, out.co, c.co, B, crop, h->params + n.cls_w_off,
                                     h->params + n.cls_b_off, n.classes, logits, pred_dev);
      return;
    } else if (n.pool) {
      ActBuf pout{bufs[(xi + 2) % 3], fs, 0};
      launch_maxpool3_fwd<TA>(h, (const TA*)out.p, out.cs, out.co, (TA*)pout.p, pout.cs, pout.co, nullptr, c.co, B, crop);
      cur = pout;
      xi = (xi + 2) % 3;
    } else if (c.post == 1) {
      ActBuf pout{bufs[(xi + 2) % 3], fs, 0};
      launch_avgpool_fwd<TA>(h, (const TA*)out.p, out.cs, out.co, (TA*)pout.p, pout.cs, pout.co, c.co, B, crop, c.post_k);
      cur = pout;
      xi = (xi + 2) % 3;
    } else if (c.post == 2) {
      // squeeze-and-excitation gate, in place (per-image statistics: the result depends on the patch, as in the reference)
      const SeBlock& sb = n.se[c.se];
      float* sum = se_scratch;
      float* sv = sum + (size_t)B * sb.c;
      float* ev = sv + (size_t)B * sb.c;
      float* hv = ev + (size_t)B * sb.c;
      se_sum_kernel<TA, 0><<<dim3(B, (unsigned)ceil_div(sb.c, 64)), 256, 0, h->stream>>>((const TA*)out.p, out.cs, out.co, nullptr, 0, 0, sb.c,
                                                                                      crop * crop, sum);
      LAUNCH_CHECK(h);
      se_fc_fwd_kernel<<<B, 256, 0, h->stream>>>(sum, 1.0f / (float)(crop * crop), h->params + sb.w1_off, h->params + sb.b1_off,
                                                h->params + sb.w2_off, h->params + sb.b2_off, sb.c, sb.r, sv, hv, ev);
      LAUNCH_CHECK(h);
      se_scale_fwd_kernel<TA><<<nblk(M * (c.co / 8), 256), 256, 0, h->stream>>>((const TA*)out.p, out.cs, out.co, ev, (TA*)out.p, out.cs, out.co,
                                                                              c.co, M, crop * crop);
      LAUNCH_CHECK(h);
      cur = out;
      xi = (xi + 1) % 3;
    } else if (sq_part >= 0) {
      if (sq_part == 0) cur = out;                       // S: read by the two expand convolutions
      else if (sq_part == 2) { cur = {bufs[sq_c], fs, 0}; xi = sq_c; }    // the module's concatenated output
    } else if (n.dense) {
      cur = {bufs[0], fs, 0};
    } else {
      cur = out;
      xi = (xi + 1) % 3;
    }
    h->taps[c.scope] = {n.dense ? out.p : cur.p, ElemTag<TA>::v, fs, n.dense ? c.out_coff : 0, c.co, M};
  }
  launch_classifier_fwd<TA>(h, (const TA*)cur.p, cur.cs, cur.co, n.cls_in, h->params + n.cls_w_off, h->params + n.cls_b_off, n.classes,
                            logits, pred_dev, M);
}

static void forward_eval(Handle* h, const float* x_dev, int B, int crop, float* logits_dev, uint8_t* pred_dev) {
  DRS_CHECK(B >= 1 && crop >= 3 && crop <= 256, "forward: bad B=%d crop=%d", B, crop);
  DRS_CHECK((int64_t)B * crop * crop < (int64_t)1 << 30, "forward: too many pixels in one call");
  switch (act_type(h)) {
    case ET_F32: forward_eval_t<float>(h, x_dev, B, crop, logits_dev, pred_dev); break;
    case ET_F16: forward_eval_t<__half>(h, x_dev, B, crop, logits_dev, pred_dev); break;
    default: forward_eval_t<__nv_bfloat16>(h, x_dev, B, crop, logits_dev, pred_dev); break;
  }
}

__global__ void widen_u8_i64_kernel(const uint8_t* __restrict__ in, long long* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

extern "C" int drs_forward_dev(drs_handle_t h, const float* x_dev, int32_t B, int32_t crop, float* logits_dev, uint8_t* pred_dev) {
  API_BEGIN
  DRS_CHECK(h && x_dev, "null argument");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  forward_eval(h, x_dev, B, crop, logits_dev, pred_dev);
  API_END
}

extern "C" int drs_forward_host(drs_handle_t h, const float* x_host, int32_t B, int32_t crop, float* logits_host, int64_t* pred_host) {
  API_BEGIN
  DRS_CHECK(h && x_host, "null argument");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  const int64_t M = (int64_t)B * crop * crop;
  const int C = h->net.channels, K = h->net.classes;
  const size_t xb = (size_t)M * C * 4, lb = (size_t)M * K * 4;
  ensure_dstage(h, xb + lb + (size_t)M * 9 + 4096);
  char* d = (char*)h->dstage;
  float* x_dev = (float*)d;
  float* lg_dev = (float*)(d + round_up(xb, 256));
  long long* p64 = (long long*)((char*)lg_dev + round_up(lb, 256));
  uint8_t* p8 = (uint8_t*)(p64 + M);
  CUDA_CHECK(cudaMemcpyAsync(x_dev, x_host, xb, cudaMemcpyHostToDevice, h->stream));
  forward_eval(h, x_dev, B, crop, lg_dev, p8);
  if (logits_host) CUDA_CHECK(cudaMemcpyAsync(logits_host, lg_dev, lb, cudaMemcpyDeviceToHost, h->stream));
  if (pred_host) {
    widen_u8_i64_kernel<<<nblk(M, 256), 256, 0, h->stream>>>(p8, p64, M);
    LAUNCH_CHECK(h);
    CUDA_CHECK(cudaMemcpyAsync(pred_host, p64, (size_t)M * 8, cudaMemcpyDeviceToHost, h->stream));
  }
  int rc = drs_synchronize(h);
  if (rc) return rc;
  API_END
}

template <typename T>
static void debug_keep(Handle* h, const std::string& name, const T* ptr, int cs, int co, int C, int64_t M) {
  HandleExtra* x = X(h);
  if (!x->debug_keep) return;
  float* buf = nullptr;
  CUDA_CHECK(cudaMalloc(&buf, M * C * 4));
  x->keep.push_back(buf);
  slice_to_f32_kernel<T><<<nblk(M * C, 256), 256, 0, h->stream>>>(ptr, cs, co, C, M, buf);
  LAUNCH_CHECK(h);
  h->taps[name] = {buf, ET_F32, C, 0, C, M};
}
static void debug_keep_reset(Handle* h) {
  HandleExtra* x = X(h);
  x->debug_keep = getenv("DRS_DEBUG_KEEP") != nullptr;
  if (x->keep.empty()) return;
  cudaStreamSynchronize(h->stream);
  for (float* p : x->keep) cudaFree(p);
  x->keep.clear();
}

#include "drs_train.cuh"
#include "drs_scene_api.cuh"

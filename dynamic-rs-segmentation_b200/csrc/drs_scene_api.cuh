// Scene-resident entry points (gather, ordered accumulate + argmax, whole-scene inference, confusion)
// and the debug / profiling hooks.  Included by drs_api.cu.
#pragma once

static void scene_upload_rows(Handle* h, int32_t scene_id, const void* rows_host, int32_t H, int32_t W, int32_t C, int32_t dtype,
                              const uint8_t* labels_rows_host, int32_t row0, int32_t rows) {
  DRS_CHECK(h && rows_host, "null argument");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  DRS_CHECK(scene_id >= 0 && scene_id < MAX_SCENES, "scene_id %d out of range [0,%d)", scene_id, MAX_SCENES);
  DRS_CHECK(dtype == DRS_SCENE_F64 || dtype == DRS_SCENE_F32, "bad scene dtype %d", dtype);
  DRS_CHECK(C == h->net.channels, "scene has %d channels, net expects %d", C, h->net.channels);
  DRS_CHECK(row0 >= 0 && rows >= 1 && row0 + rows <= H, "scene rows [%d,%d) outside [0,%d)", row0, row0 + rows, H);
  Scene& s = h->scenes[scene_id];
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  const size_t bytes = (size_t)rows * W * C * (dtype == DRS_SCENE_F64 ? 8 : 4);
  if (s.data_cap < bytes) {
    if (s.data) CUDA_CHECK(cudaFree(s.data));
    s.data = nullptr; s.data_cap = 0;
    CUDA_CHECK(cudaMalloc(&s.data, bytes));
    s.data_cap = bytes;
  }
  CUDA_CHECK(cudaMemcpyAsync(s.data, rows_host, bytes, cudaMemcpyHostToDevice, h->stream));
  if (labels_rows_host) {
    if (s.labels_cap < (size_t)rows * W) {
      if (s.labels) CUDA_CHECK(cudaFree(s.labels));
      s.labels = nullptr; s.labels_cap = 0;
      CUDA_CHECK(cudaMalloc(&s.labels, (size_t)rows * W));
      s.labels_cap = (size_t)rows * W;
    }
    CUDA_CHECK(cudaMemcpyAsync(s.labels, labels_rows_host, (size_t)rows * W, cudaMemcpyHostToDevice, h->stream));
  } else if (s.labels) {
    CUDA_CHECK(cudaFree(s.labels));
    s.labels = nullptr; s.labels_cap = 0;
  }
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  s.H = H; s.W = W; s.C = C; s.dtype = dtype; s.row0 = row0; s.rows = rows;
  X(h)->table.s[scene_id] = SceneDesc{s.data, s.labels, H, W, C, dtype, row0, rows};
}

extern "C" int drs_scene_upload(drs_handle_t h, int32_t scene_id, const void* scene_host, int32_t H, int32_t W, int32_t C,
                                int32_t dtype, const uint8_t* labels_host) {
  API_BEGIN
  scene_upload_rows(h, scene_id, scene_host, H, W, C, dtype, labels_host, 0, H);
  API_END
}

extern "C" int drs_scene_upload_rows(drs_handle_t h, int32_t scene_id, const void* rows_host, int32_t H, int32_t W, int32_t C,
                                     int32_t dtype, const uint8_t* labels_rows_host, int32_t row0, int32_t rows) {
  API_BEGIN
  scene_upload_rows(h, scene_id, rows_host, H, W, C, dtype, labels_rows_host, row0, rows);
  API_END
}

extern "C" int drs_scene_free(drs_handle_t h, int32_t scene_id) {
  API_BEGIN
  DRS_CHECK(h, "null handle");
  auto it = h->scenes.find(scene_id);
  if (it == h->scenes.end()) return 0;
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (it->second.data) cudaFree(it->second.data);
  if (it->second.labels) cudaFree(it->second.labels);
  h->scenes.erase(it);
  memset(&X(h)->table.s[scene_id], 0, sizeof(SceneDesc));
  API_END
}

extern "C" int drs_set_gather_fp16(drs_handle_t h, int32_t on) {
  API_BEGIN
  DRS_CHECK(h, "null handle");
  h->gather_fp16 = on ? 1 : 0;
  API_END
}

extern "C" int drs_set_normalization(drs_handle_t h, const double* mean3, const double* std3) {
  API_BEGIN
  DRS_CHECK(h && mean3 && std3, "null argument");
  for (int i = 0; i < 3; ++i) { h->norm_mean[i] = mean3[i]; h->norm_std[i] = std3[i]; }
  API_END
}

// all pointer members of gp are device pointers
static void launch_gather(Handle* h, GatherParams gp, const GatherInline* il = nullptr) {
  for (int i = 0; i < 3; ++i) { gp.mean[i] = h->norm_mean[i]; gp.stdv[i] = h->norm_std[i]; }
  gp.C = h->net.channels;
  gp.fp16_patches = h->gather_fp16;
  const int64_t n = (int64_t)gp.B * gp.crop * gp.crop * gp.C;
  if (il) {
    gather_kernel<true><<<nblk(n, 256), 256, 0, h->stream>>>(X(h)->table, gp, *il);
  } else {
    static const GatherInline none = {};
    gather_kernel<false><<<nblk(n, 256), 256, 0, h->stream>>>(X(h)->table, gp, none);
  }
  LAUNCH_CHECK(h);
}

static void gather_dev_impl(drs_handle_t h, const int32_t* inst_host, const uint8_t* flips_host, int32_t B, int32_t crop,
                            const double* noise_host, const uint8_t* noise_on_host, const double* over_x_host,
                            const uint8_t* over_y_host, const uint8_t* over_on_host, const double* rot_host,
                            const uint8_t* rot_on_host, float* x_out_dev, float* y_out_dev, uint8_t* amask_out_dev) {
  DRS_CHECK(h && inst_host && x_out_dev, "null argument");
  DRS_CHECK(B >= 1 && crop >= 1, "gather: empty batch (B=%d, crop=%d)", B, crop);
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  const int C = h->net.channels;
  const int64_t pp = (int64_t)B * crop * crop;
  for (int b = 0; b < B; ++b) {
    const int sid = inst_host[b * 3], r = inst_host[b * 3 + 1], c = inst_host[b * 3 + 2];
    const bool over = over_on_host && over_on_host[b];
    if (over) continue;
    auto it = h->scenes.find(sid);
    DRS_CHECK(it != h->scenes.end(), "gather: scene %d not uploaded", sid);
    DRS_CHECK(r >= 0 && c >= 0 && r + crop <= it->second.H && c + crop <= it->second.W,
              "Error: Current PATCH size is out of the scene (scene %d, row %d, col %d, crop %d)", sid, r, c, crop);
    DRS_CHECK(r >= it->second.row0 && r + crop <= it->second.row0 + it->second.rows,
              "gather: rows [%d,%d) of scene %d are not resident (uploaded rows [%d,%d))", r, r + crop, sid, it->second.row0,
              it->second.row0 + it->second.rows);
  }
  const bool plain = !(noise_host && noise_on_host) && !(over_x_host && over_on_host) && !(rot_host && rot_on_host);
  if (plain && B <= GATHER_INLINE_MAX) {
    // common case (no host-made noise / rotation overrides): everything travels in the kernel parameters
    GatherInline il;
    memset(&il, 0, sizeof(il));
    memcpy(il.inst, inst_host, (size_t)B * 12);
    if (flips_host) memcpy(il.flips, flips_host, B);
    GatherParams gp;
    memset(&gp, 0, sizeof(gp));
    gp.x_out = x_out_dev;
    gp.y_out = y_out_dev;
    gp.amask_out = amask_out_dev;
    gp.B = B;
    gp.crop = crop;
    launch_gather(h, gp, &il);
    return;
  }
  size_t need = round_up((size_t)B * 3 * 4, 256) + 4 * round_up((size_t)B, 256) + round_up((size_t)B * 48, 256);
  if (noise_host) need += round_up((size_t)pp * C * 8, 256);
  if (over_x_host) need += round_up((size_t)pp * C * 8, 256) + round_up((size_t)pp, 256);
  ensure_dstage(h, need + 1024);
  char* d = (char*)h->dstage;
  GatherParams gp;
  memset(&gp, 0, sizeof(gp));
  auto put = [&](const void* src, size_t bytes) -> void* {
    void* dst = d;
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->stream));
    d += round_up(bytes, 256);
    return dst;
  };
  gp.inst = (const int32_t*)put(inst_host, (size_t)B * 12);
  if (flips_host) gp.flips = (const uint8_t*)put(flips_host, B);
  if (noise_host && noise_on_host) {
    gp.noise_on = (const uint8_t*)put(noise_on_host, B);
    gp.noise = (const double*)put(noise_host, (size_t)pp * C * 8);
  }
  if (over_x_host && over_on_host) {
    gp.over_on = (const uint8_t*)put(over_on_host, B);
    gp.over_x = (const double*)put(over_x_host, (size_t)pp * C * 8);
    if (over_y_host) gp.over_y = (const uint8_t*)put(over_y_host, (size_t)pp);
  }
  if (rot_host && rot_on_host) {
    gp.rot_on = (const uint8_t*)put(rot_on_host, B);
    gp.rot = (const double*)put(rot_host, (size_t)B * 48);
  }
  gp.x_out = x_out_dev;
  gp.y_out = y_out_dev;
  gp.amask_out = amask_out_dev;
  gp.B = B;
  gp.crop = crop;
  launch_gather(h, gp);
  // the staging buffer is reused by the next call: the copies above must have been consumed
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
}

extern "C" int drs_gather_dev(drs_handle_t h, const int32_t* inst_host, const uint8_t* flips_host, int32_t B, int32_t crop,
                              const double* noise_host, const uint8_t* noise_on_host, const double* over_x_host,
                              const uint8_t* over_y_host, const uint8_t* over_on_host, float* x_out_dev, float* y_out_dev) {
  API_BEGIN
  gather_dev_impl(h, inst_host, flips_host, B, crop, noise_host, noise_on_host, over_x_host, over_y_host, over_on_host, nullptr,
                  nullptr, x_out_dev, y_out_dev, nullptr);
  API_END
}

extern "C" int drs_gather_rot_dev(drs_handle_t h, const int32_t* inst_host, const uint8_t* flips_host, int32_t B, int32_t crop,
                                  const double* noise_host, const uint8_t* noise_on_host, const double* rot_host,
                                  const uint8_t* rot_on_host, float* x_out_dev, float* y_out_dev, uint8_t* amask_out_dev) {
  API_BEGIN
  gather_dev_impl(h, inst_host, flips_host, B, crop, noise_host, noise_on_host, nullptr, nullptr, nullptr, rot_host, rot_on_host,
                  x_out_dev, y_out_dev, amask_out_dev);
  API_END
}

// The gather of a planned batch (drs_plan_isprs_batch) without any host synchronisation: the plan's arrays (page-locked
// host memory owned by the caller, untouched until the step's result has been fetched) are copied on plan_stream into one
// of two device staging slots and the gather is ordered behind that copy with an event; a staging slot is rewritten only
// after the gather that read it two calls ago has run.
extern "C" int drs_gather_plan_dev(drs_handle_t h, const int32_t* inst_host, const uint8_t* flips_host, int32_t B, int32_t crop,
                                   const uint8_t* rot_on_host, const double* rot_host, const uint8_t* noise_on_host,
                                   const int32_t* noise_slot_host, const double* noise_host, int64_t noise_count,
                                   float* x_out_dev, float* y_out_dev, uint8_t* amask_out_dev) {
  API_BEGIN
  DRS_CHECK(h && inst_host && x_out_dev, "null argument");
  DRS_CHECK(B >= 1 && crop >= 1, "gather_plan: bad B=%d crop=%d", B, crop);
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  HandleExtra* x = X(h);
  for (int b = 0; b < B; ++b) {
    const int sid = inst_host[b * 3], r = inst_host[b * 3 + 1], c = inst_host[b * 3 + 2];
    auto it = h->scenes.find(sid);
    DRS_CHECK(it != h->scenes.end(), "gather: scene %d not uploaded", sid);
    DRS_CHECK(r >= it->second.row0 && r + crop <= it->second.row0 + it->second.rows && c >= 0 && c + crop <= it->second.W,
              "Error: Current PATCH size is out of the resident scene rows (scene %d, row %d, col %d, crop %d)", sid, r, c, crop);
  }
  const bool has_noise = noise_host && noise_on_host && noise_slot_host && noise_count > 0;
  const bool has_rot = rot_host && rot_on_host;
  const size_t small = round_up((size_t)B * 12, 256) + 3 * round_up((size_t)B, 256) + round_up((size_t)B * 4, 256) +
                       round_up((size_t)B * 48, 256);
  const size_t need = small + (has_noise ? round_up((size_t)noise_count * 8, 256) : 0) + 1024;
  const int s = (int)(x->plan_calls++ & 1);
  if (x->plan_consumed_valid[s]) CUDA_CHECK(cudaStreamWaitEvent(x->plan_stream, x->plan_consumed[s], 0));
  if (x->plan_stage_cap[s] < need) {
    CUDA_CHECK(cudaStreamSynchronize(h->stream));
    CUDA_CHECK(cudaStreamSynchronize(x->plan_stream));
    if (x->plan_stage[s]) CUDA_CHECK(cudaFree(x->plan_stage[s]));
    x->plan_stage[s] = nullptr;
    x->plan_stage_cap[s] = 0;
    const size_t want = need + (need >> 1);
    CUDA_CHECK(cudaMalloc(&x->plan_stage[s], want));
    x->plan_stage_cap[s] = want;
  }
  char* d = (char*)x->plan_stage[s];
  auto put = [&](const void* src, size_t bytes) -> void* {
    void* dst = d;
    CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, x->plan_stream));
    d += round_up(bytes, 256);
    return dst;
  };
  GatherParams gp;
  memset(&gp, 0, sizeof(gp));
  gp.inst = (const int32_t*)put(inst_host, (size_t)B * 12);
  if (flips_host) gp.flips = (const uint8_t*)put(flips_host, B);
  if (has_rot) {
    gp.rot_on = (const uint8_t*)put(rot_on_host, B);
    gp.rot = (const double*)put(rot_host, (size_t)B * 48);
  }
  if (has_noise) {
    gp.noise_on = (const uint8_t*)put(noise_on_host, B);
    gp.noise_slot = (const int32_t*)put(noise_slot_host, (size_t)B * 4);
    gp.noise = (const double*)put(noise_host, (size_t)noise_count * 8);
  }
  CUDA_CHECK(cudaEventRecord(x->plan_uploaded[s], x->plan_stream));
  CUDA_CHECK(cudaStreamWaitEvent(h->stream, x->plan_uploaded[s], 0));
  gp.x_out = x_out_dev;
  gp.y_out = y_out_dev;
  gp.amask_out = amask_out_dev;
  gp.B = B;
  gp.crop = crop;
  launch_gather(h, gp);
  CUDA_CHECK(cudaEventRecord(x->plan_consumed[s], h->stream));
  x->plan_consumed_valid[s] = true;
  API_END
}

// host-only: the visiting order of create_patches_per_map (no GPU needed)
extern "C" int drs_grid_positions(int32_t H, int32_t W, int32_t crop, int32_t batch, int32_t variant, int32_t* pos_out,
                                  int64_t cap_pairs, int64_t* n_out) {
  API_BEGIN
  std::vector<int32_t> pos;
  grid_positions(H, W, crop, batch, variant, pos);
  const int64_t n = (int64_t)pos.size() / 2;
  if (n_out) *n_out = n;
  if (pos_out) {
    DRS_CHECK(cap_pairs >= n, "grid_positions: capacity %lld < %lld", (long long)cap_pairs, (long long)n);
    memcpy(pos_out, pos.data(), pos.size() * 4);
  }
  API_END
}

// cell tables of an ordered position list (host)
struct CellTables {
  int nh, nw, stride;
  std::vector<int32_t> off, seq;
};
static void build_cells(const std::vector<int32_t>& pos, int H, int W, int crop, CellTables& ct) {
  const int stride = crop / 2;
  ct.stride = stride;
  ct.nh = grid_count(H, crop, stride);
  ct.nw = grid_count(W, crop, stride);
  const int ncell = ct.nh * ct.nw;
  const int64_t P = (int64_t)pos.size() / 2;
  std::vector<int32_t> cell_of(P);
  ct.off.assign(ncell + 1, 0);
  for (int64_t p = 0; p < P; ++p) {
    const int y0 = pos[2 * p], x0 = pos[2 * p + 1];
    int i = (y0 % stride == 0 && y0 / stride < ct.nh) ? y0 / stride : -1;
    int j = (x0 % stride == 0 && x0 / stride < ct.nw) ? x0 / stride : -1;
    if (y0 == H - crop) i = ct.nh - 1;
    if (x0 == W - crop) j = ct.nw - 1;
    DRS_CHECK(i >= 0 && j >= 0, "accumulate: position (%d,%d) is not on the sliding-window lattice", y0, x0);
    cell_of[p] = i * ct.nw + j;
    ct.off[cell_of[p] + 1]++;
  }
  for (int c = 0; c < ncell; ++c) ct.off[c + 1] += ct.off[c];
  ct.seq.assign(P, 0);
  std::vector<int32_t> fill(ct.off.begin(), ct.off.end() - 1);
  for (int64_t p = 0; p < P; ++p) ct.seq[fill[cell_of[p]]++] = (int32_t)p;   // ascending visit order per cell
}

struct ScenePass {
  float* prob = nullptr;
  uint32_t* occur = nullptr;
  int32_t* cell_off = nullptr;
  int32_t* cell_seq = nullptr;
  void release() { prob = nullptr; occur = nullptr; cell_off = cell_seq = nullptr; }   // buffers live in the handle's slots
};

static void accumulate_chunk(Handle* h, const ScenePass& sp, const CellTables& ct, const float* logits, const std::vector<int32_t>& pos,
                             int seq0, int seq1, int H, int W, int K, int crop, int row_begin, int row_end) {
  int y_lo = H, y_hi = 0;
  for (int s = seq0; s < seq1; ++s) {
    y_lo = std::min(y_lo, pos[2 * s]);
    y_hi = std::max(y_hi, pos[2 * s] + crop);
  }
  y_lo = std::max(y_lo, row_begin);
  y_hi = std::min(y_hi, row_end);
  if (y_hi <= y_lo) return;
  AccumParams ap;
  ap.logits = logits; ap.cell_off = sp.cell_off; ap.cell_seq = sp.cell_seq; ap.prob = sp.prob; ap.occur = sp.occur;
  ap.H = H; ap.W = W; ap.K = K; ap.crop = crop; ap.stride = ct.stride; ap.nh = ct.nh; ap.nw = ct.nw;
  ap.row_begin = row_begin; ap.row_end = row_end; ap.y_lo = y_lo; ap.y_hi = y_hi; ap.seq0 = seq0; ap.seq1 = seq1;
  ap.overflow = h->diag_dev + 15;
  dim3 grid(nblk(W, 128), (unsigned)(y_hi - y_lo));
  accumulate_kernel<<<grid, 128, 0, h->stream>>>(ap);
  LAUNCH_CHECK(h);
}

static void scene_pass_begin(Handle* h, ScenePass& sp, const CellTables& ct, int rows, int W, int K) {
  sp.prob = (float*)slot_buf(h, 0, (size_t)rows * W * K * 4);
  sp.occur = (uint32_t*)slot_buf(h, 1, (size_t)rows * W * 4);
  sp.cell_off = (int32_t*)slot_buf(h, 2, ct.off.size() * 4);
  sp.cell_seq = (int32_t*)slot_buf(h, 3, std::max<size_t>(ct.seq.size(), 1) * 4);
  CUDA_CHECK(cudaMemsetAsync(sp.prob, 0, (size_t)rows * W * K * 4, h->stream));
  CUDA_CHECK(cudaMemsetAsync(sp.occur, 0, (size_t)rows * W * 4, h->stream));
  CUDA_CHECK(cudaMemcpyAsync(sp.cell_off, ct.off.data(), ct.off.size() * 4, cudaMemcpyHostToDevice, h->stream));
  if (!ct.seq.empty()) CUDA_CHECK(cudaMemcpyAsync(sp.cell_seq, ct.seq.data(), ct.seq.size() * 4, cudaMemcpyHostToDevice, h->stream));
}

static void scene_pass_finish(Handle* h, ScenePass& sp, int rows, int W, int K, uint8_t* labels_host, double* mean_host) {
  const int64_t npix = (int64_t)rows * W;
  uint8_t* lab_dev = (uint8_t*)slot_buf(h, 4, npix);
  double* mean_dev = mean_host ? (double*)slot_buf(h, 5, npix * K * 8) : nullptr;
  scene_argmax_kernel<<<nblk(npix, 256), 256, 0, h->stream>>>(sp.prob, sp.occur, npix, K, lab_dev, mean_dev);
  LAUNCH_CHECK(h);
  if (labels_host) CUDA_CHECK(cudaMemcpyAsync(labels_host, lab_dev, npix, cudaMemcpyDeviceToHost, h->stream));
  if (mean_host) CUDA_CHECK(cudaMemcpyAsync(mean_host, mean_dev, npix * K * 8, cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  if (h->diag_host[15]) {
    h->diag_host[15] = 0;
    DRS_FAIL("accumulate: a pixel is covered by more than %d patch visits of one chunk; the label map would be wrong", ACCUM_MAX_CONTRIB);
  }
}

extern "C" int drs_accumulate_argmax(drs_handle_t h, const float* logits_dev, const int32_t* pos_host, int32_t P, int32_t crop,
                                     int32_t K, int32_t H, int32_t W, uint8_t* labels_out_host, double* mean_out_host) {
  API_BEGIN
  DRS_CHECK(h && logits_dev && pos_host && labels_out_host, "null argument");
  DRS_CHECK(K >= 1 && K <= MAX_CLASSES, "K=%d out of range", K);
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  std::vector<int32_t> pos(pos_host, pos_host + (size_t)P * 2);
  CellTables ct;
  build_cells(pos, H, W, crop, ct);
  ScenePass sp;
  try {
    scene_pass_begin(h, sp, ct, H, W, K);
    accumulate_chunk(h, sp, ct, logits_dev, pos, 0, P, H, W, K, crop, 0, H);
    scene_pass_finish(h, sp, H, W, K, labels_out_host, mean_out_host);
  } catch (...) { sp.release(); throw; }
  sp.release();
  API_END
}

// Two lanes (stream + workspace each) alternate over the chunks of a scene: while lane A's tensor-core convolutions
// run, lane B's HBM-bound kernels (gather, pooling, classifier) fill the rest of the machine.  The ordered
// accumulation stays on the handle's stream, chunk after chunk, so per-pixel sums keep the script's visiting order.
// Lanes are kept in the handle between calls (their workspace is several GB for a Potsdam tile).
static void lanes_release(Handle* h) {
  HandleExtra* x = X(h);
  for (auto& L : x->lanes) {
    if (L.stream) { cudaStreamSynchronize(L.stream); cudaStreamDestroy(L.stream); }
    if (L.arena.base) cudaFree(L.arena.base);
    if (L.x) cudaFree(L.x);
    if (L.lg) cudaFree(L.lg);
    if (L.fwd_done) cudaEventDestroy(L.fwd_done);
    if (L.acc_done) cudaEventDestroy(L.acc_done);
    L = InferLane();
  }
  if (x->lanes_ready) cudaEventDestroy(x->lanes_ready);
  x->lanes_ready = nullptr;
}
static void lanes_ensure(Handle* h, int n_lanes, size_t arena_bytes, size_t x_bytes, size_t lg_bytes) {
  HandleExtra* x = X(h);
  if (!x->lanes_ready) CUDA_CHECK(cudaEventCreateWithFlags(&x->lanes_ready, cudaEventDisableTiming));
  for (int l = 0; l < n_lanes; ++l) {
    InferLane& L = x->lanes[l];
    if (!L.stream) {
      CUDA_CHECK(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
      CUDA_CHECK(cudaEventCreateWithFlags(&L.fwd_done, cudaEventDisableTiming));
      CUDA_CHECK(cudaEventCreateWithFlags(&L.acc_done, cudaEventDisableTiming));
    }
    if (L.arena.cap < arena_bytes || L.x_cap < x_bytes || L.lg_cap < lg_bytes) {
      CUDA_CHECK(cudaStreamSynchronize(L.stream));
      if (L.arena.base) cudaFree(L.arena.base);
      if (L.x) cudaFree(L.x);
      if (L.lg) cudaFree(L.lg);
      L.arena = Arena(); L.x = L.lg = nullptr; L.x_cap = L.lg_cap = 0;
      CUDA_CHECK(cudaMalloc(&L.arena.base, arena_bytes));
      L.arena.cap = arena_bytes;
      CUDA_CHECK(cudaMalloc(&L.x, x_bytes));
      L.x_cap = x_bytes;
      CUDA_CHECK(cudaMalloc(&L.lg, lg_bytes));
      L.lg_cap = lg_bytes;
    }
  }
}

// Streamed upload of a host scene during a scene pass: rows are copied on a separate stream, in ~8 MB pieces, just ahead of
// the chunk that first reads them, so the host->device transfer of a fresh tile (1.44 GB for a Potsdam tile) overlaps the
// convolutions instead of preceding them.
struct HostSceneFeed {
  const char* host = nullptr;    // [H,W,C] in the scene dtype, row-major
  size_t row_bytes = 0;
  int uploaded = 0;              // rows [0, uploaded) are enqueued
  bool recorded = false;         // up_ev[0] has been recorded at least once during this pass
};

static void feed_rows(Handle* h, HostSceneFeed& f, const Scene& sc, int rows_needed, cudaStream_t consumer) {
  HandleExtra* x = X(h);
  if (rows_needed > sc.H) rows_needed = sc.H;
  bool any = false;
  const int rows_per_piece = (int)std::max<size_t>(1, ((size_t)8 << 20) / f.row_bytes);
  while (f.uploaded < rows_needed) {
    const int n = std::min(rows_per_piece, rows_needed - f.uploaded);
    // Pageable source: the driver stages the piece itself (measured faster than a memcpy into our own pinned ring) and
    // returns once the source has been read; kernels already queued keep the GPU busy meanwhile.
    CUDA_CHECK(cudaMemcpyAsync((char*)sc.data + (size_t)f.uploaded * f.row_bytes, f.host + (size_t)f.uploaded * f.row_bytes,
                               (size_t)n * f.row_bytes, cudaMemcpyHostToDevice, x->copy_stream));
    f.uploaded += n;
    any = true;
  }
  if (any) {
    CUDA_CHECK(cudaEventRecord(x->up_ev[0], x->copy_stream));
    f.recorded = true;
  }
  // Every consumer waits for the latest "rows uploaded" event, also when this call enqueued nothing: with two lanes the
  // rows a lane needs may have been enqueued by the OTHER lane's look-ahead, and that lane's stream never saw the event.
  // (Copies on one stream complete in order, so the latest event covers every earlier piece.)
  if (f.recorded) CUDA_CHECK(cudaStreamWaitEvent(consumer, x->up_ev[0], 0));
}

static void scene_infer_impl(drs_handle_t h, int32_t scene_id, int32_t crop, int32_t batch, int32_t variant, int32_t row_begin,
                             int32_t row_end, uint8_t* labels_out_host, double* mean_out_host, HostSceneFeed* feed);

extern "C" int drs_scene_infer(drs_handle_t h, int32_t scene_id, int32_t crop, int32_t batch, int32_t variant, int32_t row_begin,
                               int32_t row_end, uint8_t* labels_out_host, double* mean_out_host) {
  API_BEGIN
  scene_infer_impl(h, scene_id, crop, batch, variant, row_begin, row_end, labels_out_host, mean_out_host, nullptr);
  API_END
}

extern "C" int drs_scene_infer_host(drs_handle_t h, int32_t scene_id, const void* scene_host, int32_t H, int32_t W, int32_t C,
                                    int32_t dtype, int32_t crop, int32_t batch, int32_t variant, uint8_t* labels_out_host,
                                    double* mean_out_host) {
  API_BEGIN
  DRS_CHECK(h && scene_host && labels_out_host, "null argument");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  DRS_CHECK(scene_id >= 0 && scene_id < MAX_SCENES, "scene_id %d out of range [0,%d)", scene_id, MAX_SCENES);
  DRS_CHECK(dtype == DRS_SCENE_F64 || dtype == DRS_SCENE_F32, "bad scene dtype %d", dtype);
  DRS_CHECK(C == h->net.channels, "scene has %d channels, net expects %d", C, h->net.channels);
  HandleExtra* x = X(h);
  Scene& s = h->scenes[scene_id];
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  const size_t row_bytes = (size_t)W * C * (dtype == DRS_SCENE_F64 ? 8 : 4);
  const size_t bytes = row_bytes * H;
  if (s.data_cap < bytes) {
    if (s.data) CUDA_CHECK(cudaFree(s.data));
    s.data = nullptr; s.data_cap = 0;
    CUDA_CHECK(cudaMalloc(&s.data, bytes));
    s.data_cap = bytes;
  }
  if (s.labels) { CUDA_CHECK(cudaFree(s.labels)); s.labels = nullptr; s.labels_cap = 0; }
  s.H = H; s.W = W; s.C = C; s.dtype = dtype; s.row0 = 0; s.rows = H;
  x->table.s[scene_id] = SceneDesc{s.data, s.labels, H, W, C, dtype, 0, H};
  if (!x->copy_stream) CUDA_CHECK(cudaStreamCreateWithFlags(&x->copy_stream, cudaStreamNonBlocking));
  if (!x->up_ev[0]) CUDA_CHECK(cudaEventCreateWithFlags(&x->up_ev[0], cudaEventDisableTiming));
  HostSceneFeed feed;
  feed.host = (const char*)scene_host;
  feed.row_bytes = row_bytes;
  scene_infer_impl(h, scene_id, crop, batch, variant, 0, H, labels_out_host, mean_out_host, &feed);
  API_END
}

struct PassGeometry {
  int key[8];
  std::vector<int32_t> pos, inst;
  CellTables ct;
};
static std::map<Handle*, std::vector<PassGeometry>> g_pass_geometry;
static void pass_geometry_release(Handle* h) { g_pass_geometry.erase(h); }

static PassGeometry& pass_geometry(Handle* h, int scene_id, int H, int W, int crop, int batch, int variant, int row_begin, int row_end) {
  std::vector<PassGeometry>& cache = g_pass_geometry[h];
  const int key[8] = {scene_id, H, W, crop, batch, variant, row_begin, row_end};
  for (auto& g : cache)
    if (memcmp(g.key, key, sizeof(key)) == 0) return g;
  if (cache.size() >= 8) cache.erase(cache.begin());
  cache.emplace_back();
  PassGeometry& g = cache.back();
  memcpy(g.key, key, sizeof(key));
  std::vector<int32_t> all;
  grid_positions(H, W, crop, batch, variant, all);
  for (size_t p = 0; p < all.size() / 2; ++p)
    if (all[2 * p] < row_end && all[2 * p] + crop > row_begin) { g.pos.push_back(all[2 * p]); g.pos.push_back(all[2 * p + 1]); }
  build_cells(g.pos, H, W, crop, g.ct);
  const size_t P = g.pos.size() / 2;
  g.inst.resize(P * 3);
  for (size_t p = 0; p < P; ++p) { g.inst[3 * p] = scene_id; g.inst[3 * p + 1] = g.pos[2 * p]; g.inst[3 * p + 2] = g.pos[2 * p + 1]; }
  return g;
}

static void scene_infer_impl(drs_handle_t h, int32_t scene_id, int32_t crop, int32_t batch, int32_t variant, int32_t row_begin,
                             int32_t row_end, uint8_t* labels_out_host, double* mean_out_host, HostSceneFeed* feed) {
  // labels_out_host may be NULL: the stripe stays on the device for drs_scene_gather_labels / drs_scene_confusion
  DRS_CHECK(h, "null argument");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  auto it = h->scenes.find(scene_id);
  DRS_CHECK(it != h->scenes.end(), "scene_infer: scene %d not uploaded", scene_id);
  const Scene& sc = it->second;
  const int H = sc.H, W = sc.W, K = h->net.classes, C = h->net.channels;
  if (row_end <= 0 || row_end > H) row_end = H;
  if (row_begin < 0) row_begin = 0;
  DRS_CHECK(row_begin < row_end, "scene_infer: empty stripe");
  // visiting order of the script, restricted to patches that touch the stripe (order preserved), with its per-cell visit
  // tables: pure functions of the geometry, memoised (a validation run visits the same scenes with the same crop again
  // and again; for a 6000x6000 tile they take several milliseconds of host time per call)
  PassGeometry& pg = pass_geometry(h, scene_id, H, W, crop, batch, variant, row_begin, row_end);
  const std::vector<int32_t>& pos = pg.pos;
  const CellTables& ct = pg.ct;
  const int P = (int)(pos.size() / 2);
  for (int p = 0; p < P; ++p)
    DRS_CHECK(pos[2 * p] >= sc.row0 && pos[2 * p] + crop <= sc.row0 + sc.rows,
              "scene_infer: stripe [%d,%d) needs scene rows [%d,%d) but only [%d,%d) are resident", row_begin, row_end, pos[2 * p],
              pos[2 * p] + crop, sc.row0, sc.row0 + sc.rows);
  const int rows = row_end - row_begin;
  // chunk = a whole number of 128-pixel tiles close to a multiple of the SM count (full conv waves), ~0.75 M pixels
  const int64_t pp = (int64_t)crop * crop;
  const int waves = getenv("DRS_CHUNK_WAVES") ? std::max(1, atoi(getenv("DRS_CHUNK_WAVES"))) : 40;
  const int64_t target_tiles = (int64_t)h->sm_count * waves;
  int chunk = (int)std::max<int64_t>(1, std::min<int64_t>(P, (target_tiles * CONV_TC_BM) / pp));
  // A second lane (DRS_TWO_LANES=1) overlaps one chunk's HBM-bound kernels with the other's convolutions; measured on
  // B200 it loses (L2 thrash + power cap: 1174 ms vs 979 ms for a 6000x6000 tile), so one lane is the default.
  const int n_lanes = (P > chunk && getenv("DRS_TWO_LANES")) ? 2 : 1;
  ScenePass sp;
  int32_t* inst_dev = nullptr;
  HandleExtra* hx = X(h);
  cudaStream_t main_stream = h->stream;
  Arena main_arena = h->arena;
  auto cleanup = [&]() {
    h->stream = main_stream;
    h->arena = main_arena;
    cudaStreamSynchronize(main_stream);
    for (auto& L : hx->lanes)
      if (L.stream) cudaStreamSynchronize(L.stream);
    sp.release();
  };
  try {
    scene_pass_begin(h, sp, ct, rows, W, K);
    const std::vector<int32_t>& inst = pg.inst;
    inst_dev = (int32_t*)slot_buf(h, 6, std::max<size_t>(inst.size(), 1) * 4);
    CUDA_CHECK(cudaMemcpyAsync(inst_dev, inst.data(), inst.size() * 4, cudaMemcpyHostToDevice, h->stream));
    refresh_packed(h, false);                       // packed weights / folded BN once, before the lanes start
    lanes_ensure(h, n_lanes, forward_eval_workspace(h, chunk, crop), (size_t)chunk * pp * C * 4, (size_t)chunk * pp * K * 4);
    CUDA_CHECK(cudaEventRecord(hx->lanes_ready, main_stream));
    for (int l = 0; l < n_lanes; ++l) CUDA_CHECK(cudaStreamWaitEvent(hx->lanes[l].stream, hx->lanes_ready, 0));
    int ci = 0;
    for (int s0 = 0; s0 < P; s0 += chunk, ++ci) {
      InferLane& L = hx->lanes[ci % n_lanes];
      const int nb = std::min(chunk, P - s0);
      h->stream = L.stream;
      h->arena = L.arena;
      if (ci >= n_lanes) CUDA_CHECK(cudaStreamWaitEvent(L.stream, L.acc_done, 0));   // logits buffer consumed
      if (feed) {
        // rows this chunk's patches read, plus one more chunk's worth so that the copy runs ahead of the compute
        int need = 0;
        const int s1 = std::min(P, s0 + 2 * chunk);
        for (int s = s0; s < s1; ++s) need = std::max(need, pos[2 * s] + crop);
        feed_rows(h, *feed, sc, need, L.stream);
      }
      GatherParams gp;
      memset(&gp, 0, sizeof(gp));
      gp.inst = inst_dev + (size_t)s0 * 3;
      gp.x_out = L.x;
      gp.B = nb;
      gp.crop = crop;
      launch_gather(h, gp);
      forward_eval(h, L.x, nb, crop, L.lg, nullptr);
      CUDA_CHECK(cudaEventRecord(L.fwd_done, L.stream));
      h->stream = main_stream;
      h->arena = main_arena;
      CUDA_CHECK(cudaStreamWaitEvent(main_stream, L.fwd_done, 0));
      accumulate_chunk(h, sp, ct, L.lg, pos, s0, s0 + nb, H, W, K, crop, row_begin, row_end);
      CUDA_CHECK(cudaEventRecord(L.acc_done, main_stream));
    }
    scene_pass_finish(h, sp, rows, W, K, labels_out_host, mean_out_host);
    hx->last_scene = scene_id; hx->last_row_begin = row_begin; hx->last_rows = rows; hx->last_W = W;
    if (feed) {                                      // rows no patch reads (none for the grids of the scripts) and the tail
      feed_rows(h, *feed, sc, sc.H, main_stream);
      CUDA_CHECK(cudaStreamSynchronize(hx->copy_stream));
    }
  } catch (...) { cleanup(); throw; }
  cleanup();
}

// Scene-level confusion matrix (isprs:1289-1296, contest:944-951): the label map of the last drs_scene_infer pass over
// `scene_id` (still resident) against the ground truth uploaded with the scene; pixels whose truth is `ignore_label`
// (the eroded class 6 of isprs, the unlabelled class 7 of the contest) or >= K are skipped.
extern "C" int drs_scene_confusion(drs_handle_t h, int32_t scene_id, int32_t K, int32_t ignore_label, uint32_t* cm_out_host) {
  API_BEGIN
  DRS_CHECK(h && cm_out_host, "null argument");
  DRS_CHECK(K >= 1 && K <= MAX_CLASSES, "K=%d out of range", K);
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  HandleExtra* x = X(h);
  auto it = h->scenes.find(scene_id);
  DRS_CHECK(it != h->scenes.end(), "scene_confusion: scene %d not uploaded", scene_id);
  const Scene& sc = it->second;
  DRS_CHECK(sc.labels, "scene_confusion: scene %d was uploaded without labels", scene_id);
  DRS_CHECK(x->last_scene == scene_id && x->slot_ptr[4], "scene_confusion: no label map of scene %d is resident (run drs_scene_infer first)", scene_id);
  DRS_CHECK(x->last_W == sc.W && x->last_row_begin >= sc.row0 && x->last_row_begin + x->last_rows <= sc.row0 + sc.rows,
            "scene_confusion: label rows [%d,%d) outside the resident ground truth", x->last_row_begin, x->last_row_begin + x->last_rows);
  const int64_t n = (int64_t)x->last_rows * sc.W;
  const uint8_t* truth = sc.labels + (int64_t)(x->last_row_begin - sc.row0) * sc.W;
  CUDA_CHECK(cudaMemsetAsync(x->cm_dev, 0, (K * K + 1) * 4, h->stream));
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(std::max<int64_t>(n, 1), 256), (int64_t)h->sm_count * 8);
  confusion_kernel<<<blocks, 256, 0, h->stream>>>(truth, (const uint8_t*)x->slot_ptr[4], nullptr, n, K, ignore_label, x->cm_dev);
  LAUNCH_CHECK(h);
  CUDA_CHECK(cudaMemcpyAsync(cm_out_host, x->cm_dev, (K * K + 1) * 4, cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  API_END
}

extern "C" int drs_confusion_dev(drs_handle_t h, const uint8_t* truth_dev, const uint8_t* pred_dev, const uint8_t* mask_dev,
                                 int64_t n, int32_t K, int32_t ignore_label, uint32_t* cm_out_host) {
  API_BEGIN
  DRS_CHECK(h && truth_dev && pred_dev && cm_out_host, "null argument");
  DRS_CHECK(K >= 1 && K <= MAX_CLASSES, "K=%d out of range", K);
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  HandleExtra* x = X(h);
  CUDA_CHECK(cudaMemsetAsync(x->cm_dev, 0, (K * K + 1) * 4, h->stream));
  const unsigned blocks = (unsigned)std::min<int64_t>(ceil_div(std::max<int64_t>(n, 1), 256), (int64_t)h->sm_count * 8);
  confusion_kernel<<<blocks, 256, 0, h->stream>>>(truth_dev, pred_dev, mask_dev, n, K, ignore_label, x->cm_dev);
  LAUNCH_CHECK(h);
  CUDA_CHECK(cudaMemcpyAsync(cm_out_host, x->cm_dev, (K * K + 1) * 4, cudaMemcpyDeviceToHost, h->stream));
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  API_END
}

// ------------------------------------------------------------------------------------------------
// profiling + debug hooks
// ------------------------------------------------------------------------------------------------
extern "C" int drs_set_profiling(drs_handle_t h, int32_t on) {
  API_BEGIN
  DRS_CHECK(h, "null handle");
  h->time_convs = on != 0;
  h->conv_events_used = 0;
  X(h)->conv_flops = 0;
  X(h)->conv_launches = 0;
  X(h)->conv_ms_acc = 0;
  API_END
}
// Sum of the device time of the tensor-core convolution launches recorded since drs_set_profiling(1)
// (CUDA events on the handle's stream), their count and their algorithmic FLOPs; resets the record.
extern "C" int drs_profile_read(drs_handle_t h, float* conv_ms, int64_t* conv_launches, double* conv_flops) {
  API_BEGIN
  DRS_CHECK(h, "null handle");
  CUDA_CHECK(cudaStreamSynchronize(h->stream));
  float total = (float)X(h)->conv_ms_acc;
  X(h)->conv_ms_acc = 0;
  for (size_t i = 0; i < h->conv_events_used; ++i) {
    float ms = 0.0f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, h->conv_events[i].first, h->conv_events[i].second));
    total += ms;
  }
  if (conv_ms) *conv_ms = total;
  if (conv_launches) *conv_launches = X(h)->conv_launches;
  if (conv_flops) *conv_flops = X(h)->conv_flops;
  h->conv_events_used = 0;
  X(h)->conv_flops = 0;
  X(h)->conv_launches = 0;
  API_END
}
extern "C" int drs_last_conv_ms(drs_handle_t h, float* ms_out) { return drs_profile_read(h, ms_out, nullptr, nullptr); }

extern "C" int drs_debug_activation(drs_handle_t h, const char* name, float* out_host, int64_t count) {
  API_BEGIN
  DRS_CHECK(h && name && out_host, "null argument");
  auto it = h->taps.find(name);
  DRS_CHECK(it != h->taps.end(), "no activation recorded for scope '%s'", name);
  const Handle::Tap& t = it->second;
  DRS_CHECK(count == t.pixels * t.co, "activation '%s' has %lld elements, got %lld", name, (long long)(t.pixels * t.co), (long long)count);
  float* tmp = nullptr;
  CUDA_CHECK(cudaMalloc(&tmp, count * 4));
  if (t.type == ET_F32) slice_to_f32_kernel<float><<<nblk(count, 256), 256, 0, h->stream>>>((const float*)t.ptr, t.cstride, t.coff, t.co, t.pixels, tmp);
  else if (t.type == ET_F16) slice_to_f32_kernel<__half><<<nblk(count, 256), 256, 0, h->stream>>>((const __half*)t.ptr, t.cstride, t.coff, t.co, t.pixels, tmp);
  else slice_to_f32_kernel<__nv_bfloat16><<<nblk(count, 256), 256, 0, h->stream>>>((const __nv_bfloat16*)t.ptr, t.cstride, t.coff, t.co, t.pixels, tmp);
  h->launches++;
  cudaError_t e = cudaMemcpyAsync(out_host, tmp, count * 4, cudaMemcpyDeviceToHost, h->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
  cudaFree(tmp);
  CUDA_CHECK(e);
  API_END
}

// One convolution through the production kernels (unit tests): y = act(conv(x, w) * scale + shift)
extern "C" int drs_debug_conv(drs_handle_t h, const float* x_host, const float* w_host, const float* scale_host, const float* shift_host,
                              int32_t B, int32_t crop, int32_t k, int32_t rate, int32_t Ci, int32_t Co, int32_t act, int32_t precision,
                              float* y_host) {
  API_BEGIN
  DRS_CHECK(h && x_host && w_host && scale_host && shift_host && y_host, "null argument");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  const int64_t M = (int64_t)B * crop * crop;
  const int taps = k * k;
  const int64_t nw = (int64_t)taps * Ci * Co;
  const int pad_b = ((k - 1) * rate) / 2;
  float *x32 = nullptr, *w32 = nullptr, *sc = nullptr, *sh = nullptr, *y32 = nullptr;
  void *xa = nullptr, *wp = nullptr, *ya = nullptr;
  auto cleanup = [&]() {
    cudaFree(x32); cudaFree(w32); cudaFree(sc); cudaFree(sh); cudaFree(y32); cudaFree(xa); cudaFree(wp); cudaFree(ya);
  };
  try {
    CUDA_CHECK(cudaMalloc(&x32, M * Ci * 4));
    CUDA_CHECK(cudaMalloc(&w32, nw * 4));
    CUDA_CHECK(cudaMalloc(&sc, Co * 4));
    CUDA_CHECK(cudaMalloc(&sh, Co * 4));
    CUDA_CHECK(cudaMalloc(&y32, M * Co * 4));
    CUDA_CHECK(cudaMemcpyAsync(x32, x_host, M * Ci * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(w32, w_host, nw * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(sc, scale_host, Co * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(sh, shift_host, Co * 4, cudaMemcpyHostToDevice, h->stream));
    if (precision == DRS_PREC_FP32) {
      launch_conv_simt<float, float>(h, x32, Ci, 0, Ci, w32, y32, Co, 0, Co, B, crop, k, rate, pad_b, sc, sh, act);
    } else {
      CUDA_CHECK(cudaMalloc(&xa, M * Ci * 2));
      CUDA_CHECK(cudaMalloc(&wp, nw * 2));
      CUDA_CHECK(cudaMalloc(&ya, M * Co * 2));
      ConvTcArgs a;
      a.in = xa; a.in_cstride = Ci; a.in_coff = 0; a.ci = Ci; a.w = wp; a.out = ya; a.out_cstride = Co; a.out_coff = 0; a.co = Co;
      a.B = B; a.crop = crop; a.k = k; a.rate = rate; a.pad_b = pad_b; a.scale = sc; a.shift = sh; a.act = act;
      if (precision == DRS_PREC_F16) {
        cast_kernel<float, __half><<<nblk(M * Ci, 256), 256, 0, h->stream>>>(x32, (__half*)xa, M * Ci);
        pack_fprop_kernel<__half><<<nblk(nw, 256), 256, 0, h->stream>>>(w32, (__half*)wp, taps, Ci, Co);
        a.etype = ET_F16;
        launch_conv_tc(h, a);
        cast_kernel<__half, float><<<nblk(M * Co, 256), 256, 0, h->stream>>>((const __half*)ya, y32, M * Co);
      } else {
        cast_kernel<float, __nv_bfloat16><<<nblk(M * Ci, 256), 256, 0, h->stream>>>(x32, (__nv_bfloat16*)xa, M * Ci);
        pack_fprop_kernel<__nv_bfloat16><<<nblk(nw, 256), 256, 0, h->stream>>>(w32, (__nv_bfloat16*)wp, taps, Ci, Co);
        a.etype = ET_BF16;
        launch_conv_tc(h, a);
        cast_kernel<__nv_bfloat16, float><<<nblk(M * Co, 256), 256, 0, h->stream>>>((const __nv_bfloat16*)ya, y32, M * Co);
      }
      h->launches += 3;
    }
    CUDA_CHECK(cudaMemcpyAsync(y_host, y32, M * Co * 4, cudaMemcpyDeviceToHost, h->stream));
    int rc = drs_synchronize(h);
    if (rc) throw DrsError{rc};
  } catch (...) { cleanup(); throw; }
  cleanup();
  API_END
}

// Kernel micro-benchmark (tools/conv_bench.py): `reps` back-to-back launches of one tcgen05 convolution on pseudo-random
// resident operands, timed with events on the handle's stream.  exp_mode >= 0 overrides DRS_EXP_MODE for these launches.
__global__ void bench_fill_kernel(uint16_t* __restrict__ p, int64_t n, uint32_t seed, int bf16) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    uint32_t v = (uint32_t)i * 2654435761u + seed;
    v ^= v >> 15; v *= 2246822519u; v ^= v >> 13;
    const float f = ((float)(v & 0xffff) / 65536.0f - 0.5f) * 0.5f;
    p[i] = bf16 ? __bfloat16_as_ushort(__float2bfloat16(f)) : __half_as_ushort(__float2half(f));
  }
}
extern "C" int drs_bench_conv(drs_handle_t h, int32_t B, int32_t crop, int32_t k, int32_t rate, int32_t Ci, int32_t Co,
                              int32_t precision, int32_t exp_mode, int32_t reps, float* ms_out, uint32_t* instr_out) {
  API_BEGIN
  DRS_CHECK(h && ms_out && reps >= 1, "bad argument");
  DRS_CHECK(precision == DRS_PREC_F16 || precision == DRS_PREC_BF16, "bench_conv: tensor-core precisions only");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  const int64_t M = (int64_t)B * crop * crop;
  const int64_t nw = (int64_t)k * k * Ci * Co;
  const int pad_b = ((k - 1) * rate) / 2;
  void *xa = nullptr, *wp = nullptr, *ya = nullptr;
  float* sc = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  auto cleanup = [&]() {
    cudaFree(xa); cudaFree(wp); cudaFree(ya); cudaFree(sc);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    g_conv_exp_mode = -1;
  };
  try {
    const int bf = precision == DRS_PREC_BF16;
    CUDA_CHECK(cudaMalloc(&xa, M * Ci * 2));
    CUDA_CHECK(cudaMalloc(&wp, nw * 2));
    CUDA_CHECK(cudaMalloc(&ya, M * Co * 2));
    CUDA_CHECK(cudaMalloc(&sc, 2 * 256 * 4));
    bench_fill_kernel<<<1024, 256, 0, h->stream>>>((uint16_t*)xa, M * Ci, 1u, bf);
    bench_fill_kernel<<<256, 256, 0, h->stream>>>((uint16_t*)wp, nw, 2u, bf);
    {
      std::vector<float> one_zero(512, 0.0f);
      for (int i = 0; i < 256; ++i) one_zero[i] = 1.0f;
      CUDA_CHECK(cudaMemcpy(sc, one_zero.data(), 512 * 4, cudaMemcpyHostToDevice));
    }
    CUDA_CHECK(cudaEventCreate(&e0));
    CUDA_CHECK(cudaEventCreate(&e1));
    ConvTcArgs a;
    a.in = xa; a.in_cstride = Ci; a.in_coff = 0; a.ci = Ci; a.w = wp; a.out = ya; a.out_cstride = Co; a.out_coff = 0; a.co = Co;
    a.B = B; a.crop = crop; a.k = k; a.rate = rate; a.pad_b = pad_b; a.scale = sc; a.shift = sc + 256; a.act = ACT_RELU;
    a.etype = bf ? ET_BF16 : ET_F16;
    g_conv_exp_mode = exp_mode;
    launch_conv_tc(h, a);                       // warm-up (tensor maps, module load)
    CUDA_CHECK(cudaEventRecord(e0, h->stream));
    for (int r = 0; r < reps; ++r) launch_conv_tc(h, a);
    CUDA_CHECK(cudaEventRecord(e1, h->stream));
    int rc = drs_synchronize(h);
    if (rc) throw DrsError{rc};
    float ms = 0.0f;
    CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    *ms_out = ms / reps;
    // instrumented launches (exp_mode bit 16): cycle counters of CTA 0's last launch
    if (instr_out) for (int i = 0; i < 10; ++i) instr_out[i] = h->diag_host[4 + i];
  } catch (...) { cleanup(); throw; }
  cleanup();
  API_END
}

// Filter gradient of one convolution through the production kernels (unit tests)
extern "C" int drs_debug_wgrad(drs_handle_t h, const float* x_host, const float* dy_host, int32_t B, int32_t crop, int32_t k,
                               int32_t rate, int32_t Ci, int32_t Co, int32_t precision, float* dw_host) {
  API_BEGIN
  DRS_CHECK(h && x_host && dy_host && dw_host, "null argument");
  CUDA_CHECK(cudaSetDevice(h->cfg.device));
  const int64_t M = (int64_t)B * crop * crop;
  const int64_t nw = (int64_t)k * k * Ci * Co;
  const int pad_b = ((k - 1) * rate) / 2;
  const int max_splits = 48;
  float *x32 = nullptr, *dy32 = nullptr, *dw = nullptr, *part = nullptr;
  void *xa = nullptr, *dya = nullptr;
  auto cleanup = [&]() { cudaFree(x32); cudaFree(dy32); cudaFree(dw); cudaFree(part); cudaFree(xa); cudaFree(dya); };
  try {
    CUDA_CHECK(cudaMalloc(&x32, M * Ci * 4));
    CUDA_CHECK(cudaMalloc(&dy32, M * Co * 4));
    CUDA_CHECK(cudaMalloc(&dw, nw * 4));
    // (the tensor-core kernel runs 32-channel operands as 64: its split partials have the padded shape)
    const int64_t nw_pad = (int64_t)k * k * ((Ci + 63) / 64 * 64) * ((Co + 63) / 64 * 64);
    CUDA_CHECK(cudaMalloc(&part, nw_pad * 4 * max_splits));
    CUDA_CHECK(cudaMemcpyAsync(x32, x_host, M * Ci * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_CHECK(cudaMemcpyAsync(dy32, dy_host, M * Co * 4, cudaMemcpyHostToDevice, h->stream));
    if (precision == DRS_PREC_FP32) {
      launch_wgrad_simt<float, float>(h, x32, Ci, 0, Ci, dy32, Co, 0, Co, dw, part, max_splits, B, crop, k, rate, pad_b);
    } else {
      DRS_CHECK(precision == DRS_PREC_BF16, "debug_wgrad: precision must be FP32 or BF16");
      CUDA_CHECK(cudaMalloc(&xa, M * Ci * 2));
      CUDA_CHECK(cudaMalloc(&dya, M * Co * 2));
      cast_kernel<float, __nv_bfloat16><<<nblk(M * Ci, 256), 256, 0, h->stream>>>(x32, (__nv_bfloat16*)xa, M * Ci);
      cast_kernel<float, __nv_bfloat16><<<nblk(M * Co, 256), 256, 0, h->stream>>>(dy32, (__nv_bfloat16*)dya, M * Co);
      h->launches += 2;
      WgradTcArgs wa;
      wa.x = xa; wa.in_cstride = Ci; wa.in_coff = 0; wa.ci = Ci;
      wa.dy = dya; wa.dy_cstride = Co; wa.dy_coff = 0; wa.co = Co;
      wa.B = B; wa.crop = crop; wa.k = k; wa.rate = rate; wa.pad_b = pad_b;
      wa.dw = dw; wa.part = part; wa.part_capacity = (size_t)nw_pad * max_splits;
      launch_wgrad_tc(h, wa);
    }
    CUDA_CHECK(cudaMemcpyAsync(dw_host, dw, nw * 4, cudaMemcpyDeviceToHost, h->stream));
    int rc = drs_synchronize(h);
    if (rc) throw DrsError{rc};
  } catch (...) { cleanup(); throw; }
  cleanup();
  API_END
}

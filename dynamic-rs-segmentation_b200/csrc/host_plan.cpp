// Host-side step planner of the isprs training loop (no CUDA in this file; linked into libdrs.so).
//
// dynamically_create_patches (/root/reference/isprs_dilated_random.py:245-334) decides, per patch and in this order,
//     np.random.randint(0, 2)                    rotate?           isprs:289
//     np.random.randint(0, 2)                    additive noise?   isprs:298
//     np.random.normal(0, 0.01, patch.shape)     the noise itself  isprs:300   (only when the previous draw was 1)
//     np.random.randint(0, 3)                    flip              isprs:304
// from the legacy global NumPy stream (MT19937), which the patch-size draw of the NEXT step (isprs:1727-1737) shares.  A
// run seeded like the reference must therefore consume exactly the same 32-bit words, and the noise values must be the
// ones NumPy would have produced -- they are added to the patch in float64 before the cast to float32.
//
// This file restates the three generators involved, bit for bit, from NumPy's own sources (numpy 2.3,
// numpy/random/src/mt19937/mt19937.c, src/legacy/legacy-distributions.c, src/distributions/distributions.c):
//     next_uint32      MT19937 with the standard tempering
//     legacy double    (a >> 5, b >> 6) -> (a * 67108864 + b) / 9007199254740992
//     randint(0, n)    masked rejection on one 32-bit word per attempt (buffered_bounded_masked_uint32)
//     legacy_gauss     polar Box-Muller, returns f*x2 and caches f*x1 (the cache survives across calls)
// The caller passes np.random.get_state() in and writes the result back with np.random.set_state().
//
// Why native: the Python loop + np.random.normal cost 7-12 ms per batch of 64 against a 1.9 ms GPU step.  Here the
// stream is scanned once sequentially (word generation + the rejection test, which fixes how many words each noise
// array consumes) and the expensive part -- sqrt(-2 log(r2) / r2) per accepted pair -- is deferred and spread over a
// small worker pool, because it no longer touches the generator state.  The same libm `log` NumPy calls is used
// (no fast-math, no vector math library, no FMA contraction), so the values are bit-identical; tests/test_host_plan.py
// pins them against np.random for random seeds, odd/even element counts and a cached Gaussian carried across calls.
#include <emmintrin.h>
#include <math.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/drs.h"

namespace {

constexpr int MT_N = 624, MT_M = 397;
constexpr uint32_t MATRIX_A = 0x9908b0dfU, UPPER_MASK = 0x80000000U, LOWER_MASK = 0x7fffffffU;

inline void mt_gen(drs_mt_state* s) {
  uint32_t* mt = s->key;
  int kk = 0;
  uint32_t y;
  for (; kk < MT_N - MT_M; kk++) {
    y = (mt[kk] & UPPER_MASK) | (mt[kk + 1] & LOWER_MASK);
    mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ (-(y & 1) & MATRIX_A);
  }
  for (; kk < MT_N - 1; kk++) {
    y = (mt[kk] & UPPER_MASK) | (mt[kk + 1] & LOWER_MASK);
    mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ (-(y & 1) & MATRIX_A);
  }
  y = (mt[MT_N - 1] & UPPER_MASK) | (mt[0] & LOWER_MASK);
  mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ (-(y & 1) & MATRIX_A);
  s->pos = 0;
}

// The generator is consumed through a small look-ahead buffer of TEMPERED words: one regeneration of the 624-word state is
// tempered in one (vectorisable) pass, and the scan below reads plain words.  `st->pos` is restored from the number of
// unread words when the scan ends, so the state handed back is exactly the one NumPy would hold.
struct Stream {
  drs_mt_state* st;
  uint32_t buf[8 + MT_N];
  int head = 0, tail = 0;

  static inline void temper(const uint32_t* __restrict__ in, uint32_t* __restrict__ out, int n) {
    for (int i = 0; i < n; ++i) {
      uint32_t y = in[i];
      y ^= (y >> 11);
      y ^= (y << 7) & 0x9d2c5680U;
      y ^= (y << 15) & 0xefc60000U;
      y ^= (y >> 18);
      out[i] = y;
    }
  }
  explicit Stream(drs_mt_state* s) : st(s) {
    const int left = MT_N - s->pos;
    temper(s->key + s->pos, buf, left);
    tail = left;
  }
  // keep the unread words (fewer than one attempt's worth) and append a freshly generated block behind them
  inline void refill() {
    const int r = tail - head;
    for (int i = 0; i < r; ++i) buf[i] = buf[head + i];
    mt_gen(st);
    temper(st->key, buf + r, MT_N);
    head = 0;
    tail = r + MT_N;
  }
  inline uint32_t next32() {
    if (head == tail) refill();
    return buf[head++];
  }
  inline void finish() { st->pos = MT_N - (tail - head); }
};

// np.random.randint(0, n) for 1 <= n-1 < 2^32-1: one word per attempt, masked, rejected while > n-1
inline uint32_t mt_randint(Stream& s, uint32_t n) {
  const uint32_t rng = n - 1;
  if (rng == 0) return 0;
  uint32_t mask = rng;
  mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
  uint32_t v;
  while ((v = (s.next32() & mask)) > rng) {}
  return v;
}

// ---- deferred polar transform -------------------------------------------------------------------------------------------
struct Pair { double x1, x2; };      // an accepted attempt; r2 = x1*x1 + x2*x2 is recomputed (same IEEE operations)

// values 2p and 2p+1 of the Gaussian stream from accepted pair p, written as loc + scale * g (legacy_normal)
inline void transform_range(const Pair* pr, int64_t p0, int64_t p1, double* out, int64_t n_out, double loc, double scale) {
  for (int64_t p = p0; p < p1; ++p) {
    const double r2 = pr[p].x1 * pr[p].x1 + pr[p].x2 * pr[p].x2;
    const double f = sqrt(-2.0 * log(r2) / r2);
    const double first = f * pr[p].x2, second = f * pr[p].x1;   // legacy_gauss returns f*x2 and caches f*x1
    out[2 * p] = loc + scale * first;
    if (2 * p + 1 < n_out) out[2 * p + 1] = loc + scale * second;
  }
}

// the value a pair leaves in the generator's cache when only its first half was consumed (has_gauss = 1)
inline double cached_half(const Pair& q) {
  const double r2 = q.x1 * q.x1 + q.x2 * q.x2;
  const double f = sqrt(-2.0 * log(r2) / r2);
  return f * q.x1;
}

class Pool {
 public:
  explicit Pool(int n) {
    for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { loop(i); });
  }
  ~Pool() {
    {
      std::lock_guard<std::mutex> g(m_);
      stop_ = true;
      ++epoch_;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  // start(fn): every worker runs fn(part, parts) concurrently with the caller; join() runs the caller's own part
  // (part = parts-1) and returns when all parts are done.  run(fn) = start + join.
  template <typename F>
  void start(F&& fn) {
    const int parts = size() + 1;
    fn_ = [fn, parts](int part) { fn(part, parts); };
    pending_.store(size());
    {
      std::lock_guard<std::mutex> g(m_);
      ++epoch_;
    }
    cv_.notify_all();
  }
  void join() {
    fn_(size());
    std::unique_lock<std::mutex> l(m_);
    done_.wait(l, [this] { return pending_.load() == 0; });
  }
  template <typename F>
  void run(F&& fn) {
    start(fn);
    join();
  }

 private:
  void loop(int id) {
    uint64_t seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return epoch_ != seen; });
        seen = epoch_;
        if (stop_) return;
      }
      fn_(id);
      if (pending_.fetch_sub(1) == 1) {
        std::lock_guard<std::mutex> g(m_);
        done_.notify_all();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  std::function<void(int)> fn_;
  std::atomic<int> pending_{0};
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

// legacy_gauss's rejection loop for `n_pairs` more accepted pairs (isprs:300 draws crop*crop*C values per noisy patch):
//     do { x1 = 2*double() - 1; x2 = 2*double() - 1; r2 = x1*x1 + x2*x2; } while (r2 >= 1.0 || r2 == 0.0);
// with NumPy's legacy double = ((w0 >> 5) * 67108864 + (w1 >> 6)) / 9007199254740992.  Every attempt consumes exactly four
// words, so the loop runs over the buffered words in blocks whose length is known up front (an attempt yields at most one
// pair), stores unconditionally and advances the output index by the acceptance flag -- no data-dependent branch.
struct PairBuf {          // grows geometrically, never zero-filled
  Pair* d = nullptr;
  size_t n = 0, cap = 0;
  ~PairBuf() { free(d); }
  void clear() { n = 0; }
  size_t size() const { return n; }
  const Pair* data() const { return d; }
  const Pair& back() const { return d[n - 1]; }
  void need(size_t want) {
    if (want <= cap) return;
    size_t c = std::max<size_t>(want, cap * 2 + 4096);
    Pair* nd = (Pair*)realloc(d, c * sizeof(Pair));
    if (!nd) abort();
    d = nd;
    cap = c;
  }
};

inline void scan_pairs(Stream& s, PairBuf& pairs, int64_t n_pairs) {
  size_t np = pairs.n;
  const size_t target = np + (size_t)n_pairs;
  pairs.need(target + 2);                         // slack for the unconditional stores
  Pair* out = pairs.d;
  const __m128d two = _mm_set1_pd(2.0), one = _mm_set1_pd(1.0), k26 = _mm_set1_pd(67108864.0), inv53 = _mm_set1_pd(1.0 / 9007199254740992.0);
  while (np < target) {
    if (s.tail - s.head < 4) s.refill();
    int64_t m = std::min<int64_t>((s.tail - s.head) / 4, (int64_t)(target - np));
    const uint32_t* w = s.buf + s.head;
    s.head += (int)(4 * m);
    // two attempts (eight words) per iteration: a = w >> 5 of the even words, b = w >> 6 of the odd ones
    for (; m >= 2; m -= 2, w += 8) {
      const __m128 v0 = _mm_castsi128_ps(_mm_loadu_si128((const __m128i*)w)), v1 = _mm_castsi128_ps(_mm_loadu_si128((const __m128i*)(w + 4)));
      const __m128i a = _mm_srli_epi32(_mm_castps_si128(_mm_shuffle_ps(v0, v1, _MM_SHUFFLE(2, 0, 2, 0))), 5);
      const __m128i b = _mm_srli_epi32(_mm_castps_si128(_mm_shuffle_ps(v0, v1, _MM_SHUFFLE(3, 1, 3, 1))), 6);
      // division by 2^53 == multiplication by 2^-53 (both exact scalings of the same exactly representable sum)
      const __m128d d01 = _mm_mul_pd(_mm_add_pd(_mm_mul_pd(_mm_cvtepi32_pd(a), k26), _mm_cvtepi32_pd(b)), inv53);
      const __m128d d23 = _mm_mul_pd(_mm_add_pd(_mm_mul_pd(_mm_cvtepi32_pd(_mm_unpackhi_epi64(a, a)), k26), _mm_cvtepi32_pd(_mm_unpackhi_epi64(b, b))), inv53);
      const __m128d x01 = _mm_sub_pd(_mm_mul_pd(two, d01), one), x23 = _mm_sub_pd(_mm_mul_pd(two, d23), one);
      const __m128d q01 = _mm_mul_pd(x01, x01), q23 = _mm_mul_pd(x23, x23);
      // r2 of both attempts in one vector: [q01.lo + q01.hi, q23.lo + q23.hi]
      const __m128d r2 = _mm_add_pd(_mm_unpacklo_pd(q01, q23), _mm_unpackhi_pd(q01, q23));
      const int acc = _mm_movemask_pd(_mm_and_pd(_mm_cmplt_pd(r2, one), _mm_cmpneq_pd(r2, _mm_setzero_pd())));
      _mm_storeu_pd(&out[np].x1, x01);
      np += (size_t)(acc & 1);
      _mm_storeu_pd(&out[np].x1, x23);
      np += (size_t)(acc >> 1);
    }
    if (m == 1) {
      const int32_t a0 = (int32_t)(w[0] >> 5), b0 = (int32_t)(w[1] >> 6), a1 = (int32_t)(w[2] >> 5), b1 = (int32_t)(w[3] >> 6);
      const double d0 = (a0 * 67108864.0 + b0) / 9007199254740992.0;
      const double d1 = (a1 * 67108864.0 + b1) / 9007199254740992.0;
      const double x1 = 2.0 * d0 - 1.0, x2 = 2.0 * d1 - 1.0;
      const double r2 = x1 * x1 + x2 * x2;
      out[np] = Pair{x1, x2};
      np += (size_t)((r2 < 1.0) & (r2 != 0.0));
    }
  }
  pairs.n = target;
}

constexpr int64_t CHUNK = 2048;     // pairs per unit of deferred work

struct Planner {
  PairBuf pairs;
  Pool* pool = nullptr;
  int pool_threads = 0;
  // pipelined transform: the scan publishes how many pairs are final; workers claim chunks behind it
  std::atomic<int64_t> ready{0}, next_chunk{0};
  std::atomic<int> scan_done{0};
  ~Planner() { delete pool; }
  void ensure_pool(int threads) {
    if (pool && pool_threads == threads) return;
    delete pool;
    pool = new Pool(threads - 1);
    pool_threads = threads;
  }
};

// pairs [pb, pe) of the step -> values out[2p], out[2p+1]
void run_transform(Planner* pl, double* out, int64_t n_out, double loc, double scale, int threads, int64_t pb, int64_t pe) {
  const int64_t P = pe - pb;
  if (P <= 0) return;
  const Pair* pr = pl->pairs.data();
  if (threads <= 1 || P < 4096) {
    transform_range(pr, pb, pe, out, n_out, loc, scale);
    return;
  }
  pl->ensure_pool(threads);
  pl->pool->run([=](int part, int parts) {
    const int64_t a = pb + P * part / parts, b = pb + P * (part + 1) / parts;
    transform_range(pr, a, b, out, n_out, loc, scale);
  });
}

// worker side of the pipelined transform: claim chunk c, wait until the scan has passed its end (or finished), transform
void chunk_worker(Planner* pl, double* out, int64_t n_out_cap, double loc, double scale) {
  const Pair* pr = pl->pairs.data();      // sized for the worst case before the scan starts: never reallocated under us
  for (;;) {
    const int64_t c = pl->next_chunk.fetch_add(1);
    const int64_t begin = c * CHUNK;
    int64_t end = begin + CHUNK;
    int spins = 0;
    for (;;) {
      const int64_t r = pl->ready.load(std::memory_order_acquire);
      if (r >= end) break;
      if (pl->scan_done.load(std::memory_order_acquire)) {
        end = std::min(end, pl->ready.load(std::memory_order_acquire));
        break;
      }
      if (++spins < 2000) _mm_pause();
      else std::this_thread::yield();
    }
    if (begin >= end) return;
    transform_range(pr, begin, end, out, n_out_cap, loc, scale);
  }
}

}  // namespace

extern "C" int drs_planner_create(drs_planner_t* out) {
  if (!out) return 1;
  *out = reinterpret_cast<drs_planner_t>(new Planner());
  return 0;
}

extern "C" int drs_planner_destroy(drs_planner_t p) {
  delete reinterpret_cast<Planner*>(p);
  return 0;
}

// np.random.normal(loc, scale, n) on the caller's generator state (unit-test entry and building block)
extern "C" int drs_mt_normal(drs_planner_t p, drs_mt_state* st, double loc, double scale, double* out, int64_t n, int32_t threads) {
  Planner* pl = reinterpret_cast<Planner*>(p);
  if (!pl || !st || (!out && n > 0) || n < 0 || st->pos < 0 || st->pos > MT_N) return 1;
  if (n == 0) return 0;
  int64_t off = 0;
  if (st->has_gauss) {
    out[0] = loc + scale * st->gauss;
    st->has_gauss = 0;
    st->gauss = 0.0;
    off = 1;
  }
  const int64_t rest = n - off, P = (rest + 1) / 2;
  pl->pairs.clear();
  Stream sm(st);
  scan_pairs(sm, pl->pairs, P);
  sm.finish();
  run_transform(pl, out + off, rest, loc, scale, threads, 0, P);
  if (rest & 1) { st->has_gauss = 1; st->gauss = cached_half(pl->pairs.back()); }
  return 0;
}

extern "C" int drs_mt_randint(drs_mt_state* st, uint32_t n, int32_t* out, int64_t count) {
  if (!st || !out || n == 0 || st->pos < 0 || st->pos > MT_N) return 1;
  Stream sm(st);
  for (int64_t i = 0; i < count; ++i) out[i] = (int32_t)mt_randint(sm, n);
  sm.finish();
  return 0;
}

extern "C" int drs_plan_isprs_batch(drs_planner_t p, drs_mt_state* st, const int64_t* batch_inst, int32_t B,
                                    const int32_t* scene_hw, int32_t n_scenes, int32_t crop, int32_t C, int32_t is_train,
                                    const double* rot_table, int32_t* inst_out, uint8_t* flips_out, uint8_t* rot_on_out,
                                    double* rot_out, uint8_t* noise_on_out, int32_t* noise_slot_out, double* noise_out,
                                    int64_t noise_cap, int32_t* n_noise_out, int32_t threads, int32_t own_b0, int32_t own_b1) {
  Planner* pl = reinterpret_cast<Planner*>(p);
  if (!pl || !st || !batch_inst || !scene_hw || !inst_out || B < 0 || crop < 1 || C < 1) return 1;
  if (st->pos < 0 || st->pos > MT_N) return 1;
  if (is_train && (!flips_out || !rot_on_out || !rot_out || !noise_on_out || !noise_slot_out || !rot_table)) return 1;
  const int64_t per = (int64_t)crop * crop * C;
  // ---- pass 0: the border rule (isprs:259-269: a window that sticks out is moved back); no generator access
  for (int b = 0; b < B; ++b) {
    const int64_t* in = batch_inst + (int64_t)b * 4;
    const int64_t map = in[0];
    if (map < 0 || map >= n_scenes) return 2;
    const int h = scene_hw[map * 2], w = scene_hw[map * 2 + 1];
    int64_t cx = in[1], cy = in[2];
    const int64_t len_x = std::max<int64_t>(0, std::min<int64_t>(cx + crop, h) - cx);
    const int64_t len_y = std::max<int64_t>(0, std::min<int64_t>(cy + crop, w) - cy);
    if (len_x != crop) cx -= crop - len_x;
    if (len_y != crop) cy -= crop - len_y;
    if (cx < 0 || cy < 0 || cx + crop > h || cy + crop > w) return 3;     // the reference prints an error and returns (isprs:273-280)
    if (is_train && (in[3] < 0 || in[3] >= 360)) return 4;                 // create_rotation_distribution draws [0, 360) (isprs:489)
    inst_out[b * 3] = (int32_t)map; inst_out[b * 3 + 1] = (int32_t)cx; inst_out[b * 3 + 2] = (int32_t)cy;
  }
  if (n_noise_out) *n_noise_out = 0;
  if (!is_train) return 0;
  // ---- pass 1: every decision, scanning the generator exactly as the reference consumes it.  The Gaussian stream of
  // the step is continuous (legacy_gauss keeps its second value across calls): value v of the step lives at out index
  // v, where a cached Gaussian left by an earlier call (has_gauss) is value 0.  Patch slot k owns values
  // [k*per, (k+1)*per).  The deferred transform runs behind the scan on the worker pool (chunk_worker); data-parallel
  // ranks, which scan the whole batch to stay on the same stream but need only the noise of their own patches
  // [own_b0, own_b1), transform that share after the scan instead.
  const int64_t worst_pairs = (int64_t)B * (per / 2 + 1) + 2;
  pl->pairs.clear();
  pl->pairs.need((size_t)worst_pairs + 2);
  Stream sm(st);
  int64_t n_values = 0;       // Gaussian values consumed so far in this step
  int64_t n_have = 0;         // values available from the cache + the pairs scanned so far
  const int lead = st->has_gauss ? 1 : 0;
  const double lead_gauss = st->gauss;
  if (lead) { n_have = 1; st->has_gauss = 0; st->gauss = 0.0; }
  const bool all_own = own_b0 <= 0 && own_b1 >= B;
  const bool pipelined = threads > 1 && all_own && noise_out && noise_cap >= (int64_t)B * per + 2 && (int64_t)B * per >= 8 * CHUNK;
  if (pipelined) {
    pl->ensure_pool(threads);
    pl->ready.store(0);
    pl->next_chunk.store(0);
    pl->scan_done.store(0);
    double* o = noise_out + lead;
    pl->pool->start([pl, o](int, int) { chunk_worker(pl, o, INT64_MAX, 0.0, 0.01); });
  }
  int n_noise = 0;
  for (int b = 0; b < B; ++b) {
    const uint32_t possible_rotation = mt_randint(sm, 2);                 // isprs:289
    rot_on_out[b] = (uint8_t)possible_rotation;
    if (possible_rotation == 1) memcpy(rot_out + (int64_t)b * 6, rot_table + batch_inst[(int64_t)b * 4 + 3] * 6, 48);
    else memset(rot_out + (int64_t)b * 6, 0, 48);
    const uint32_t possible_noise = mt_randint(sm, 2);                    // isprs:298
    noise_on_out[b] = (uint8_t)possible_noise;
    noise_slot_out[b] = -1;
    if (possible_noise == 1) {
      noise_slot_out[b] = n_noise++;
      n_values += per;
      if (n_values > n_have) {
        const int64_t P = (n_values - n_have + 1) / 2;
        scan_pairs(sm, pl->pairs, P);
        n_have += 2 * P;
        if (pipelined) pl->ready.store((int64_t)pl->pairs.size(), std::memory_order_release);
      }
    }
    const uint32_t possible_flip = mt_randint(sm, 3);                     // isprs:304: 0 none, 1 flipud, 2 fliplr
    flips_out[b] = (uint8_t)possible_flip;
  }
  sm.finish();
  if (pipelined) {
    pl->scan_done.store(1, std::memory_order_release);
    pl->pool->join();                              // the caller works through the remaining chunks too
  }
  if (n_noise_out) *n_noise_out = n_noise;
  if (n_values > 0) {
    if (!noise_out || noise_cap < n_values + 2) return 5;
    if (lead) noise_out[0] = 0.0 + 0.01 * lead_gauss;
    if (!pipelined) {
      const int64_t rest = n_values - lead;          // values that come from the pairs scanned in this step
      const int64_t P = (int64_t)pl->pairs.size();
      int64_t s0 = -1, s1 = -1;                      // noise slots of the rank's patches
      for (int b = std::max(0, own_b0); b < std::min(B, own_b1); ++b)
        if (noise_slot_out[b] >= 0) { if (s0 < 0) s0 = noise_slot_out[b]; s1 = noise_slot_out[b] + 1; }
      if (s0 >= 0) {
        const int64_t v0 = std::max<int64_t>(s0 * per - lead, 0), v1 = std::min<int64_t>(s1 * per - lead, rest);
        run_transform(pl, noise_out + lead, INT64_MAX, 0.0, 0.01, threads, v0 / 2, std::min<int64_t>((v1 + 1) / 2, P));
      }
    }
    if (n_have > n_values) { st->has_gauss = 1; st->gauss = cached_half(pl->pairs.back()); }
  } else if (lead) {
    st->has_gauss = 1; st->gauss = lead_gauss;     // nothing consumed it
  }
  return 0;
}

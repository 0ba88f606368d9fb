// Updating (non-windowed) keyed aggregate on sm_100a: the GPU side of `IncrementalAggregatingFunc`
// (arroyo-worker/src/arrow/incremental_aggregator.rs), SURVEY.md 8(f) rank 2.
//
// The reference keeps one accumulator object per key and aggregate, updates them ONE ROW AT A TIME through dyn
// `Accumulator`s (:860-879), remembers for every key touched since the last flush the values it had before (:842-857)
// and, at a flush (every `flush_interval` tick, at checkpoints, at end of data), emits per touched key a retraction of
// the old values and an append of the new ones -- unless only the timestamp moved (:637-738).
//
// Here (append-only inputs: no `_updating_meta.is_retract` upstream; COUNT(*) / SUM / AVG / MIN / MAX over Int64):
//   ingest  one thread per row: dense id from the bucketed key dictionary (bdict.cuh), one RED per accumulator, the
//           trailing max(_timestamp) aggregate as a RED.max, and the key joins the touched list on its first row
//           since the last flush (atomicExch on a per-id flag);
//   flush   one thread per touched key: compares the accumulators with their values at the previous flush (kept per
//           id: "the values it had before"), writes the retraction row (old values, old timestamp) and the append row
//           (new values), and rolls the snapshot forward.
// Output rows: [key?, aggregates..., _timestamp, is_retract] -- retractions first, then appends (a key's retraction
// must precede its append; the order between keys is unspecified in the reference too: it iterates a HashMap).
// The shim wraps `is_retract` into the `_updating_meta` struct together with the row id its metadata expression
// computes (:719-729).
//
// Not restated: retractions on the input (an updating upstream), count(distinct), TTL expiry (wall clock).  Such
// plans are refused at construction (ARROYO_B200_UNSUPPORTED) and stay on the stock operator.
#include <algorithm>
#include <climits>

#include "bdict.cuh"
#include "op.h"

namespace ab {
namespace {

constexpr int U_MAX_ACC = ARROYO_B200_MAX_AGGS + 1;
enum : int { U_ROWS = 0, U_SUM = 1, U_MIN = 3, U_MAX = 4 };

struct UState {
  unsigned long long* cur;   // [n_acc][id_cap]; cur[0] = rows
  unsigned long long* prev;  // same layout: values at the previous flush (prev rows == 0: the key did not exist)
  long long* cur_ts;         // max(_timestamp)
  long long* prev_ts;
  unsigned int* touched;     // per id: in the touched list
  unsigned int* list;        // touched ids
  unsigned int* n_touched;
  unsigned long long id_cap;
  int n_acc;
  int acc_kind[U_MAX_ACC];
  int acc_val[U_MAX_ACC];
};

struct UIngest {
  const long long* key;
  const long long* ts;
  const long long* val[4];
  long long n;
  int keyed;
  BDict dict;
  UState st;
  unsigned long long* lost;
};

__global__ void __launch_bounds__(256) upd_ingest_kernel(const __grid_constant__ UIngest p) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (; i < p.n; i += stride) {
    uint32_t id = 0;
    if (p.keyed) {
      id = bd_lookup_or_insert(p.dict, __ldcs(p.key + i));
      if (id >= ID_OVERFLOW) {
        atomicAdd(p.lost, 1ull);
        continue;
      }
    }
    atomicAdd(p.st.cur + id, 1ull);
#pragma unroll
    for (int a = 1; a < U_MAX_ACC; ++a) {
      if (a >= p.st.n_acc) break;
      const long long v = __ldcs(p.val[p.st.acc_val[a]] + i);
      unsigned long long* dst = p.st.cur + (unsigned long long)a * p.st.id_cap + id;
      switch (p.st.acc_kind[a]) {
        case U_SUM: atomicAdd(dst, (unsigned long long)v); break;
        case U_MIN: atomicMin(reinterpret_cast<long long*>(dst), v); break;
        case U_MAX: atomicMax(reinterpret_cast<long long*>(dst), v); break;
      }
    }
    atomicMax(p.st.cur_ts + id, __ldcs(p.ts + i));
    if (atomicExch(p.st.touched + id, 1u) == 0u) p.st.list[atomicAdd(p.st.n_touched, 1u)] = id;
  }
}

struct UFlush {
  UState st;
  const long long* id_keys;
  unsigned int n;  // touched keys
  int keyed;
  int n_aggs;
  int agg_kind[ARROYO_B200_MAX_AGGS];
  int agg_acc[ARROYO_B200_MAX_AGGS];
  // output: retractions at [0, n_retract), appends at [n, n + n_append)
  long long* o_key;
  unsigned long long* o_agg[ARROYO_B200_MAX_AGGS];
  long long* o_ts;
  unsigned int* counts;  // [0] retractions, [1] appends
};

__device__ __forceinline__ unsigned long long finalise(int kind, unsigned long long acc, unsigned long long rows) {
  if (kind == ARROYO_B200_AGG_COUNT_STAR) return rows;
  if (kind == ARROYO_B200_AGG_AVG_I64) return (unsigned long long)__double_as_longlong((double)(long long)acc / (double)rows);
  return acc;
}

__global__ void __launch_bounds__(256) upd_flush_kernel(const __grid_constant__ UFlush p) {
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned int stride = gridDim.x * blockDim.x;
  for (; i < p.n; i += stride) {
    const unsigned int id = p.st.list[i];
    unsigned long long now[U_MAX_ACC], old[U_MAX_ACC];
    bool changed = false;
    for (int a = 0; a < p.st.n_acc; ++a) {
      now[a] = p.st.cur[(unsigned long long)a * p.st.id_cap + id];
      old[a] = p.st.prev[(unsigned long long)a * p.st.id_cap + id];
    }
    const bool had = old[0] != 0;
    // "don't bother emitting updates that just retract / append the same values (excluding the timestamp)" (:655-664):
    // compared on the OUTPUT values, like the reference compares ScalarValues
    for (int g = 0; g < p.n_aggs; ++g)
      changed = changed || finalise(p.agg_kind[g], now[p.agg_acc[g]], now[0]) != finalise(p.agg_kind[g], old[p.agg_acc[g]], old[0]);
    const long long now_ts = p.st.cur_ts[id], old_ts = p.st.prev_ts[id];
    const long long key = p.keyed ? p.id_keys[id] : 0;
    if (had && changed) {
      const unsigned int o = atomicAdd(p.counts + 0, 1u);
      if (p.keyed) p.o_key[o] = key;
      for (int g = 0; g < p.n_aggs; ++g) p.o_agg[g][o] = finalise(p.agg_kind[g], old[p.agg_acc[g]], old[0]);
      p.o_ts[o] = old_ts;
    }
    if (!had || changed) {
      const unsigned int o = p.n + atomicAdd(p.counts + 1, 1u);
      if (p.keyed) p.o_key[o] = key;
      for (int g = 0; g < p.n_aggs; ++g) p.o_agg[g][o] = finalise(p.agg_kind[g], now[p.agg_acc[g]], now[0]);
      p.o_ts[o] = now_ts;
    }
    for (int a = 0; a < p.st.n_acc; ++a) p.st.prev[(unsigned long long)a * p.st.id_cap + id] = now[a];
    p.st.prev_ts[id] = now_ts;
    p.st.touched[id] = 0;
  }
}

__global__ void upd_init_kernel(UState st, unsigned long long n) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    for (int a = 0; a < st.n_acc; ++a) {
      unsigned long long v = 0;
      if (st.acc_kind[a] == U_MIN) v = (unsigned long long)LLONG_MAX;
      if (st.acc_kind[a] == U_MAX) v = (unsigned long long)LLONG_MIN;
      st.cur[(unsigned long long)a * st.id_cap + i] = v;
      st.prev[(unsigned long long)a * st.id_cap + i] = a == 0 ? 0 : v;
    }
    st.cur_ts[i] = LLONG_MIN;
    st.prev_ts[i] = LLONG_MIN;
    st.touched[i] = 0;
  }
}

// after the dictionary grew: new[map[i]] = old[i]
__global__ void upd_permute_kernel(UState o, UState n, const uint32_t* __restrict__ map, uint32_t old_ids) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; i < old_ids; i += stride) {
    const uint32_t m = map[i];
    if (m == ID_UNSET || m >= ID_OVERFLOW) continue;
    for (int a = 0; a < o.n_acc; ++a) {
      n.cur[(unsigned long long)a * n.id_cap + m] = o.cur[(unsigned long long)a * o.id_cap + i];
      n.prev[(unsigned long long)a * n.id_cap + m] = o.prev[(unsigned long long)a * o.id_cap + i];
    }
    n.cur_ts[m] = o.cur_ts[i];
    n.prev_ts[m] = o.prev_ts[i];
    n.touched[m] = o.touched[i];
  }
}
__global__ void upd_remap_list_kernel(unsigned int* list, unsigned int n, const uint32_t* __restrict__ map) {
  unsigned int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) list[i] = map[list[i]];
}

class UpdatingAggOp final : public OpBase {
 public:
  explicit UpdatingAggOp(const ArroyoB200OpConfig& c);
  ~UpdatingAggOp() override;
  void on_start(ArrowArray*, ArrowSchema*, int64_t n, int64_t, int64_t) override {
    AB_REQUIRE(n == 0, ARROYO_B200_UNSUPPORTED, "updating aggregate: state restore (tables 'a' / 'b') is not implemented");
  }
  void process_batch(uint32_t, uint32_t, ArrowArray* batch, const ArrowSchema* schema) override;
  void process_device_batch(uint32_t, uint32_t, const uint64_t* cols, int32_t n_cols, int64_t n_rows) override;
  // watermarks pass through an updating aggregate untouched (it emits on ticks, incremental_aggregator.rs:990-1004)
  void handle_watermark(int64_t, BatchesPriv*, std::vector<ArroyoB200DeviceBatch>*) override {}
  void handle_checkpoint(int64_t, BatchesPriv* out) override { flush_to(out); }  // :951-961
  void on_close(int end_of_data, BatchesPriv* out) override {                    // :1006-1018
    if (end_of_data && out) flush_to(out);
  }
  void handle_tick(BatchesPriv* out) override { flush_to(out); }  // :994-1004
  void flush() override {
    AB_CUDA(cudaSetDevice(device_));
    AB_CUDA(cudaStreamSynchronize(stream_));
  }
  void stats(ArroyoB200Stats* out) override {
    if (keyed_) {
      AB_CUDA(cudaSetDevice(device_));
      AB_CUDA(cudaMemcpyAsync(&total_keys_, n_total_.p, 4, cudaMemcpyDeviceToHost, stream_));
      AB_CUDA(cudaStreamSynchronize(stream_));
    }
    st_.n_keys = keyed_ ? total_keys_ : 0;
    *out = st_;
  }

 private:
  int device_ = 0;
  cudaStream_t stream_ = nullptr;
  bool own_stream_ = false;
  int num_sms_ = 148;
  bool keyed_ = false;
  int key_col_ = 0, ts_col_ = 0;
  int n_vals_ = 0, val_cols_[4];
  int n_acc_ = 1, acc_kind_[U_MAX_ACC], acc_val_[U_MAX_ACC];
  int n_aggs_ = 0, agg_kind_[ARROYO_B200_MAX_AGGS], agg_acc_[ARROYO_B200_MAX_AGGS];
  std::string key_format_ = "l";
  std::vector<std::string> agg_format_;
  // dictionary + state
  uint64_t n_buckets_ = 1, id_cap_ = 0;
  uint32_t total_keys_ = 0;
  DevBuf slots_, bucket_nkeys_, id_keys_, n_total_;
  DevBuf cur_, prev_, cur_ts_, prev_ts_, touched_, list_, counters_;  // counters_: [n_touched, retractions, appends, pad] u32 + lost u64
  DevBuf staging_;
  uint64_t staging_cap_ = 0;
  uint64_t out_cap_ = 0;
  DevBuf o_key_, o_ts_, o_agg_[ARROYO_B200_MAX_AGGS];
  ArroyoB200Stats st_{};

  BDict dict_view() const;
  UState state_view() const;
  void alloc_state(uint64_t n_buckets);
  void grow();
  void ensure_room(uint64_t new_rows);
  void ingest(const long long* key, const long long* ts, const long long* const* vals, int64_t n);
  void flush_to(BatchesPriv* out);
};

UpdatingAggOp::UpdatingAggOp(const ArroyoB200OpConfig& c) {
  cfg = c;
  name = "UpdatingAggregatingFunc";
  AB_REQUIRE(c.n_key_cols == 0 || c.n_key_cols == 1, ARROYO_B200_UNSUPPORTED, "only 0 or 1 group-by key columns are supported");
  keyed_ = c.n_key_cols == 1;
  key_col_ = c.key_col;
  ts_col_ = c.timestamp_col;
  AB_REQUIRE(c.n_cols >= 1 && c.n_cols <= ARROYO_B200_MAX_COLS && ts_col_ >= 0 && ts_col_ < c.n_cols,
             ARROYO_B200_INVALID_ARGUMENT, "bad column layout");
  AB_REQUIRE(c.n_aggs >= 1 && c.n_aggs <= ARROYO_B200_MAX_AGGS, ARROYO_B200_INVALID_ARGUMENT, "bad n_aggs");
  AB_REQUIRE(!(c.flags & ARROYO_B200_FLAG_UPDATING_INPUT), ARROYO_B200_UNSUPPORTED,
             "updating aggregate over an updating input (retractions) is not supported");
  n_aggs_ = c.n_aggs;
  acc_kind_[0] = U_ROWS;
  acc_val_[0] = 0;
  for (int g = 0; g < n_aggs_; ++g) {
    const int kind = c.aggs[g].kind;
    agg_kind_[g] = kind;
    agg_acc_[g] = 0;
    if (kind == ARROYO_B200_AGG_COUNT_STAR) {
      agg_format_.push_back("l");
      continue;
    }
    const int col = c.aggs[g].input_col;
    AB_REQUIRE(col >= 0 && col < c.n_cols, ARROYO_B200_INVALID_ARGUMENT, "aggregate input column out of range");
    int vs = -1;
    for (int v = 0; v < n_vals_; ++v)
      if (val_cols_[v] == col) vs = v;
    if (vs < 0) {
      AB_REQUIRE(n_vals_ < 4, ARROYO_B200_UNSUPPORTED, "more than 4 distinct aggregate input columns");
      vs = n_vals_;
      val_cols_[n_vals_++] = col;
    }
    int ak;
    switch (kind) {
      case ARROYO_B200_AGG_SUM_I64: ak = U_SUM; agg_format_.push_back("l"); break;
      case ARROYO_B200_AGG_AVG_I64: ak = U_SUM; agg_format_.push_back("g"); break;  // exact integer sum, divided at output
      case ARROYO_B200_AGG_MIN_I64: ak = U_MIN; agg_format_.push_back("l"); break;
      case ARROYO_B200_AGG_MAX_I64: ak = U_MAX; agg_format_.push_back("l"); break;
      default: throw Error(ARROYO_B200_UNSUPPORTED, "unsupported aggregate kind");
    }
    int found = -1;
    for (int a = 1; a < n_acc_; ++a)
      if (acc_kind_[a] == ak && acc_val_[a] == vs) found = a;
    if (found < 0) {
      found = n_acc_;
      acc_kind_[n_acc_] = ak;
      acc_val_[n_acc_] = vs;
      ++n_acc_;
    }
    agg_acc_[g] = found;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0)
    throw Error(ARROYO_B200_FATAL, "no CUDA device available: libarroyo_b200 has no CPU fallback");
  device_ = c.device;
  AB_REQUIRE(device_ >= 0 && device_ < count, ARROYO_B200_INVALID_ARGUMENT, "bad device ordinal");
  AB_CUDA(cudaSetDevice(device_));
  cudaDeviceProp prop{};
  AB_CUDA(cudaGetDeviceProperties(&prop, device_));
  num_sms_ = prop.multiProcessorCount;
  if (c.stream) {
    stream_ = (cudaStream_t)c.stream;
  } else {
    AB_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    own_stream_ = true;
  }
  counters_.alloc(32);
  AB_CUDA(cudaMemsetAsync(counters_.p, 0, 32, stream_));
  n_total_.alloc(4);
  AB_CUDA(cudaMemsetAsync(n_total_.p, 0, 4, stream_));
  alloc_state(keyed_ ? bd_buckets_for(c.expected_keys ? c.expected_keys : (1ull << 16)) : 1);
  AB_CUDA(cudaStreamSynchronize(stream_));
}

UpdatingAggOp::~UpdatingAggOp() {
  cudaSetDevice(device_);
  cudaStreamSynchronize(stream_);
  if (own_stream_ && stream_) cudaStreamDestroy(stream_);
}

BDict UpdatingAggOp::dict_view() const {
  BDict d{};
  d.slots = slots_.as<BSlot>();
  d.nkeys = bucket_nkeys_.as<unsigned int>();
  d.id_keys = id_keys_.as<long long>();
  d.n_total = n_total_.as<unsigned int>();
  d.n_buckets = (uint32_t)n_buckets_;
  return d;
}

UState UpdatingAggOp::state_view() const {
  UState s{};
  s.cur = cur_.as<unsigned long long>();
  s.prev = prev_.as<unsigned long long>();
  s.cur_ts = cur_ts_.as<long long>();
  s.prev_ts = prev_ts_.as<long long>();
  s.touched = touched_.as<unsigned int>();
  s.list = list_.as<unsigned int>();
  s.n_touched = counters_.as<unsigned int>();
  s.id_cap = id_cap_;
  s.n_acc = n_acc_;
  for (int a = 0; a < n_acc_; ++a) {
    s.acc_kind[a] = acc_kind_[a];
    s.acc_val[a] = acc_val_[a];
  }
  return s;
}

void UpdatingAggOp::alloc_state(uint64_t n_buckets) {
  n_buckets_ = n_buckets;
  id_cap_ = bd_id_cap(n_buckets_);
  AB_REQUIRE(id_cap_ < (1ull << 31), ARROYO_B200_RUNTIME, "key dictionary too large");
  id_keys_.alloc(id_cap_ * 8);
  bd_fill_keys_kernel<<<num_sms_ * 4, 256, 0, stream_>>>(id_keys_.as<long long>(), id_cap_);
  AB_CUDA(cudaGetLastError());
  bucket_nkeys_.alloc(n_buckets_ * 4);
  AB_CUDA(cudaMemsetAsync(bucket_nkeys_.p, 0, n_buckets_ * 4, stream_));
  if (keyed_) {
    slots_.alloc(n_buckets_ * BD_KS * sizeof(BSlot));
    bd_init_kernel<<<num_sms_ * 4, 256, 0, stream_>>>(slots_.as<BSlot>(), n_buckets_ * BD_KS);
    AB_CUDA(cudaGetLastError());
  }
  cur_.alloc((size_t)n_acc_ * id_cap_ * 8);
  prev_.alloc((size_t)n_acc_ * id_cap_ * 8);
  cur_ts_.alloc(id_cap_ * 8);
  prev_ts_.alloc(id_cap_ * 8);
  touched_.alloc(id_cap_ * 4);
  list_.alloc(id_cap_ * 4);
  upd_init_kernel<<<num_sms_ * 4, 256, 0, stream_>>>(state_view(), id_cap_);
  AB_CUDA(cudaGetLastError());
  st_.kernel_launches += 2;
}

// Doubles the bucket count: keys are re-inserted (ids change), the per-id state and the touched list follow the map.
void UpdatingAggOp::grow() {
  const BDict old_d = dict_view();
  const UState old_s = state_view();
  const uint32_t old_ids = (uint32_t)(BD_ID_BASE + n_buckets_ * BD_CAPB);
  const uint64_t old_cap = id_cap_;
  DevBuf k_slots = std::move(slots_), k_nk = std::move(bucket_nkeys_), k_keys = std::move(id_keys_), k_cur = std::move(cur_),
         k_prev = std::move(prev_), k_cts = std::move(cur_ts_), k_pts = std::move(prev_ts_), k_t = std::move(touched_),
         k_list = std::move(list_);
  unsigned int h_touched = 0;
  AB_CUDA(cudaMemcpyAsync(&h_touched, counters_.p, 4, cudaMemcpyDeviceToHost, stream_));
  AB_CUDA(cudaMemsetAsync(n_total_.p, 0, 4, stream_));
  AB_CUDA(cudaStreamSynchronize(stream_));
  alloc_state(n_buckets_ * 2);
  DevBuf map((size_t)old_cap * 4);
  const int grid = (int)std::min<uint64_t>((old_ids + 255) / 256, (uint64_t)num_sms_ * 8);
  bd_rehash_kernel<<<std::max(grid, 1), 256, 0, stream_>>>(old_d, dict_view(), old_ids, map.as<uint32_t>());
  AB_CUDA(cudaGetLastError());
  upd_permute_kernel<<<std::max(grid, 1), 256, 0, stream_>>>(old_s, state_view(), map.as<uint32_t>(), old_ids);
  AB_CUDA(cudaGetLastError());
  if (h_touched) {
    AB_CUDA(cudaMemcpyAsync(list_.p, k_list.p, (size_t)h_touched * 4, cudaMemcpyDeviceToDevice, stream_));
    upd_remap_list_kernel<<<(h_touched + 255) / 256, 256, 0, stream_>>>(list_.as<unsigned int>(), h_touched, map.as<uint32_t>());
    AB_CUDA(cudaGetLastError());
  }
  AB_CUDA(cudaMemcpyAsync(&total_keys_, n_total_.p, 4, cudaMemcpyDeviceToHost, stream_));
  AB_CUDA(cudaStreamSynchronize(stream_));
  st_.kernel_launches += 3;
}

// every row of the batch may bring a new key: keep the mean bucket fill at or under the target
void UpdatingAggOp::ensure_room(uint64_t new_rows) {
  if (!keyed_) return;
  AB_CUDA(cudaMemcpyAsync(&total_keys_, n_total_.p, 4, cudaMemcpyDeviceToHost, stream_));
  AB_CUDA(cudaStreamSynchronize(stream_));
  while ((uint64_t)total_keys_ + new_rows > n_buckets_ * (uint64_t)BD_MEAN) grow();
}

void UpdatingAggOp::ingest(const long long* key, const long long* ts, const long long* const* vals, int64_t n) {
  if (n <= 0) return;
  UIngest p{};
  p.key = key;
  p.ts = ts;
  for (int v = 0; v < n_vals_; ++v) p.val[v] = vals[v];
  p.n = n;
  p.keyed = keyed_ ? 1 : 0;
  p.dict = dict_view();
  p.st = state_view();
  p.lost = reinterpret_cast<unsigned long long*>((char*)counters_.p + 16);
  const int grid = (int)std::min<int64_t>((n + 255) / 256, (int64_t)num_sms_ * 8);
  upd_ingest_kernel<<<std::max(grid, 1), 256, 0, stream_>>>(p);
  AB_CUDA(cudaGetLastError());
  ++st_.kernel_launches;
  ++st_.ingest_launches;
}

void UpdatingAggOp::process_batch(uint32_t, uint32_t, ArrowArray* batch, const ArrowSchema* schema) {
  AB_CUDA(cudaSetDevice(device_));
  int64_t n = 0;
  std::vector<InColumn> cols = import_batch(batch, schema, &n);
  AB_REQUIRE((int)cols.size() == cfg.n_cols, ARROYO_B200_INVALID_ARGUMENT, "batch has the wrong number of columns");
  if (keyed_) key_format_ = cols[key_col_].format;
  for (int g = 0; g < n_aggs_; ++g)
    if (agg_kind_[g] == ARROYO_B200_AGG_MIN_I64 || agg_kind_[g] == ARROYO_B200_AGG_MAX_I64)
      agg_format_[g] = cols[cfg.aggs[g].input_col].format;
  st_.rows_in += (uint64_t)n;
  if (n == 0) {
    if (batch->release) batch->release(batch);
    return;
  }
  ensure_room((uint64_t)n);
  const int n_used = 2 + n_vals_;
  if ((uint64_t)n > staging_cap_) {
    AB_CUDA(cudaStreamSynchronize(stream_));
    staging_cap_ = std::max<uint64_t>((uint64_t)n, staging_cap_ * 2);
    staging_.alloc((size_t)n_used * staging_cap_ * 8);
  }
  long long* base = staging_.as<long long>();
  const long long* vals[4] = {nullptr, nullptr, nullptr, nullptr};
  if (keyed_) AB_CUDA(cudaMemcpyAsync(base, cols[key_col_].data, (size_t)n * 8, cudaMemcpyHostToDevice, stream_));
  AB_CUDA(cudaMemcpyAsync(base + staging_cap_, cols[ts_col_].data, (size_t)n * 8, cudaMemcpyHostToDevice, stream_));
  for (int v = 0; v < n_vals_; ++v) {
    AB_CUDA(cudaMemcpyAsync(base + (size_t)(2 + v) * staging_cap_, cols[val_cols_[v]].data, (size_t)n * 8, cudaMemcpyHostToDevice,
                            stream_));
    vals[v] = base + (size_t)(2 + v) * staging_cap_;
  }
  st_.h2d_bytes += (uint64_t)n * 8 * (uint64_t)((keyed_ ? 1 : 0) + 1 + n_vals_);
  ingest(base, base + staging_cap_, vals, n);
  // the staging buffer is reused by the next batch: the copies and the kernel must have consumed the host batch
  AB_CUDA(cudaStreamSynchronize(stream_));
  if (batch->release) batch->release(batch);
  batch->release = nullptr;
}

void UpdatingAggOp::process_device_batch(uint32_t, uint32_t, const uint64_t* cols, int32_t n_cols, int64_t n_rows) {
  AB_CUDA(cudaSetDevice(device_));
  AB_REQUIRE(n_cols == cfg.n_cols, ARROYO_B200_INVALID_ARGUMENT, "batch has the wrong number of columns");
  if (n_rows <= 0) return;
  st_.rows_in += (uint64_t)n_rows;
  ensure_room((uint64_t)n_rows);
  const long long* vals[4] = {nullptr, nullptr, nullptr, nullptr};
  for (int v = 0; v < n_vals_; ++v) vals[v] = (const long long*)cols[val_cols_[v]];
  ingest(keyed_ ? (const long long*)cols[key_col_] : nullptr, (const long long*)cols[ts_col_], vals, n_rows);
}

static void* d2h_part(const void* dev, size_t off_rows, int64_t n, void* host, size_t host_off_rows, cudaStream_t s) {
  if (n > 0)
    AB_CUDA(cudaMemcpyAsync((char*)host + host_off_rows * 8, (const char*)dev + off_rows * 8, (size_t)n * 8, cudaMemcpyDeviceToHost, s));
  return host;
}

// flush (:637-738): one batch [key?, aggregates..., _timestamp, is_retract], or nothing when no key changed
void UpdatingAggOp::flush_to(BatchesPriv* out) {
  AB_CUDA(cudaSetDevice(device_));
  struct {
    unsigned int touched, retracts, appends, pad;
    unsigned long long lost;
  } h{};
  AB_CUDA(cudaMemcpyAsync(&h, counters_.p, 24, cudaMemcpyDeviceToHost, stream_));
  AB_CUDA(cudaStreamSynchronize(stream_));
  AB_REQUIRE(h.lost == 0, ARROYO_B200_RUNTIME, "updating aggregate: a dictionary bucket ran out of ids");
  const unsigned int n = h.touched;
  if (n == 0 || !out) return;
  if (2ull * n > out_cap_) {
    out_cap_ = std::max<uint64_t>(2ull * n, 1024);
    o_key_.alloc(out_cap_ * 8);
    o_ts_.alloc(out_cap_ * 8);
    for (int g = 0; g < n_aggs_; ++g) o_agg_[g].alloc(out_cap_ * 8);
  }
  UFlush p{};
  p.st = state_view();
  p.id_keys = id_keys_.as<long long>();
  p.n = n;
  p.keyed = keyed_ ? 1 : 0;
  p.n_aggs = n_aggs_;
  for (int g = 0; g < n_aggs_; ++g) {
    p.agg_kind[g] = agg_kind_[g];
    p.agg_acc[g] = agg_acc_[g];
    p.o_agg[g] = o_agg_[g].as<unsigned long long>();
  }
  p.o_key = o_key_.as<long long>();
  p.o_ts = o_ts_.as<long long>();
  p.counts = counters_.as<unsigned int>() + 1;
  const int grid = (int)std::min<unsigned int>((n + 255) / 256, (unsigned int)num_sms_ * 8);
  upd_flush_kernel<<<std::max(grid, 1), 256, 0, stream_>>>(p);
  AB_CUDA(cudaGetLastError());
  ++st_.kernel_launches;
  ++st_.emit_launches;
  AB_CUDA(cudaMemcpyAsync(&h, counters_.p, 24, cudaMemcpyDeviceToHost, stream_));
  AB_CUDA(cudaMemsetAsync(counters_.p, 0, 16, stream_));  // touched list and output counters start over
  AB_CUDA(cudaStreamSynchronize(stream_));
  const int64_t nr = h.retracts, na = h.appends, total = nr + na;
  if (total == 0) return;
  std::vector<OutColumn> cols;
  auto column = [&](const char* nm, const std::string& fmt, const void* dev) {
    OutColumn c;
    c.name = nm;
    c.format = fmt;
    void* host = PinnedPool::get().alloc((size_t)std::max<int64_t>(total, 1) * 8);
    d2h_part(dev, 0, nr, host, 0, stream_);
    d2h_part(dev, n, na, host, (size_t)nr, stream_);
    st_.d2h_bytes += (uint64_t)total * 8;
    c.data = host;
    cols.push_back(c);
  };
  if (keyed_) column("key", key_format_, o_key_.p);
  for (int g = 0; g < n_aggs_; ++g) column(("agg" + std::to_string(g)).c_str(), agg_format_[g], o_agg_[g].p);
  column("_timestamp", "tsn:", o_ts_.p);
  {
    OutColumn r;
    r.name = "is_retract";
    r.format = "b";
    unsigned char* bits = (unsigned char*)PinnedPool::get().alloc((size_t)(total + 7) / 8 + 8);
    memset(bits, 0, (size_t)(total + 7) / 8 + 8);
    for (int64_t i = 0; i < nr; ++i) bits[i >> 3] |= (unsigned char)(1u << (i & 7));
    r.data = bits;
    cols.push_back(r);
  }
  AB_CUDA(cudaStreamSynchronize(stream_));
  st_.rows_out += (uint64_t)total;
  ++st_.windows_out;
  out->arrays.emplace_back();
  out->schemas.emplace_back();
  export_batch(cols, total, &out->arrays.back(), &out->schemas.back());
}

}  // namespace

OpBase* make_updating_agg_op(const ArroyoB200OpConfig& cfg) { return new UpdatingAggOp(cfg); }

}  // namespace ab

// The Shuffle edge inside the library: partition -> control all-gather -> variable all-to-all, with NCCL called from
// C++ on the caller's stream.  OPT-IN and not yet measured (written at the end of round 1, after the N > 1 benchmark
// turned out to be bound by Python / torch.distributed call overhead: DESIGN.md section 8, item 1).  It is built as a
// separate library (libarroyo_b200_xchg.so) so that nothing on the measured paths depends on it.
//
// Replaces, like arroyo_b200/multi_gpu.py::ShuffleExchange.round_packed, ArrowCollector::collect -> repartition
// (arroyo-operator/src/context.rs:506-541) plus the network hop between the subtasks of one box, and carries each
// sender's watermark for the receiver's min-merge (WatermarkHolder, context.rs:35-86).
//
// One round:
//   1. arroyo_b200_partition_packed buckets the rows by destination (counts land in the control record on the device)
//   2. ncclAllGather of the control records {rows per destination x world, watermark, more-rounds flag}
//   3. one small D2H copy + stream synchronise: the host learns the split sizes (the only host wait of the round)
//   4. grouped ncclSend / ncclRecv of the variable-size blocks (every block = its columns back to back)
// The receive buffer alternates between two allocations: a block stays valid until the round after next.
#include <nccl.h>

#include <climits>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/arroyo_b200.h"
#include "common.cuh"

namespace {
struct NcclError {
  std::string what;
};
#define AB_NCCL(expr)                                                                          \
  do {                                                                                         \
    ncclResult_t r__ = (expr);                                                                 \
    if (r__ != ncclSuccess) throw NcclError{std::string(#expr) + ": " + ncclGetErrorString(r__)}; \
  } while (0)
}  // namespace

struct ArroyoB200Exchange {
  int device = 0, rank = 0, world = 1, n_cols = 0;
  int64_t max_rows = 0;       // rows this rank may send in one round
  int64_t max_recv_rows = 0;  // rows this rank may receive in one round
  cudaStream_t stream = nullptr;
  ncclComm_t comm = nullptr;
  ArroyoB200Partitioner* part = nullptr;
  ab::DevBuf packed, ctrl, ctrl_all, offsets, recv[2];
  ab::PinnedBuf h_ctrl_all;
  int flip = 0;
  std::string last_error;
};

extern "C" {

// 128 bytes (NCCL_UNIQUE_ID_BYTES): created by one rank, handed to every rank's arroyo_b200_xchg_create out of band
// (the benchmark broadcasts it through torch.distributed)
int32_t arroyo_b200_xchg_unique_id(void* out128) {
  if (!out128) return ARROYO_B200_INVALID_ARGUMENT;
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return ARROYO_B200_RUNTIME;
  static_assert(sizeof id == 128, "NCCL unique id size");
  memcpy(out128, &id, sizeof id);
  return ARROYO_B200_OK;
}

int32_t arroyo_b200_xchg_create(int32_t device, uint64_t stream, int32_t rank, int32_t world, const void* unique_id128,
                                int32_t n_cols, int32_t key_col, int64_t max_rows, int64_t max_recv_rows,
                                ArroyoB200Exchange** out) {
  if (!out) return ARROYO_B200_INVALID_ARGUMENT;
  *out = nullptr;
  if (!unique_id128 || !stream || world < 1 || rank < 0 || rank >= world || n_cols < 1 || max_rows < 1 || max_recv_rows < 1)
    return ARROYO_B200_INVALID_ARGUMENT;
  auto* x = new ArroyoB200Exchange();
  try {
    AB_CUDA(cudaSetDevice(device));
    x->device = device;
    x->rank = rank;
    x->world = world;
    x->n_cols = n_cols;
    x->max_rows = max_rows;
    x->max_recv_rows = max_recv_rows;
    x->stream = (cudaStream_t)stream;
    ncclUniqueId id;
    memcpy(&id, unique_id128, sizeof id);
    AB_NCCL(ncclCommInitRank(&x->comm, world, id, rank));
    int32_t st = arroyo_b200_partitioner_create(device, stream, world, n_cols, key_col, max_rows, &x->part);
    if (st != ARROYO_B200_OK) throw ab::Error(st, "partitioner_create failed");
    const size_t rec = (size_t)(world + 2) * sizeof(long long);
    x->packed.alloc((size_t)max_rows * n_cols * 8);
    x->ctrl.alloc(rec);
    x->ctrl_all.alloc(rec * world);
    x->offsets.alloc((size_t)world * 8);
    x->h_ctrl_all.alloc(rec * world);
    for (auto& r : x->recv) r.alloc((size_t)max_recv_rows * n_cols * 8);
    *out = x;
    return ARROYO_B200_OK;
  } catch (const ab::Error& e) {
    delete x;
    return e.status;
  } catch (const NcclError&) {
    delete x;
    return ARROYO_B200_RUNTIME;
  } catch (...) {
    delete x;
    return ARROYO_B200_RUNTIME;
  }
}

void arroyo_b200_xchg_destroy(ArroyoB200Exchange* x) {
  if (!x) return;
  cudaSetDevice(x->device);
  cudaStreamSynchronize(x->stream);
  if (x->part) arroyo_b200_partitioner_destroy(x->part);
  if (x->comm) ncclCommDestroy(x->comm);
  delete x;
}

const char* arroyo_b200_xchg_last_error(const ArroyoB200Exchange* x) { return x ? x->last_error.c_str() : "invalid handle"; }

// One round.  in_cols: n_cols device pointers (ignored when n_rows == 0).  watermark_ns = INT64_MIN for "none".
// Outputs: recv_cols[s * n_cols + c] = device pointer of column c of sender s's block (0 when it sent nothing),
// recv_rows[s] its rows, sender_watermarks[s] (INT64_MIN = none), *any_more = some sender has more rounds queued.
int32_t arroyo_b200_xchg_round(ArroyoB200Exchange* x, const uint64_t* in_cols, int64_t n_rows, int64_t watermark_ns,
                               int32_t more, uint64_t* recv_cols, int64_t* recv_rows, int64_t* sender_watermarks,
                               int32_t* any_more) {
  if (!x || !recv_cols || !recv_rows || !sender_watermarks || !any_more || n_rows < 0 || n_rows > x->max_rows ||
      (n_rows > 0 && !in_cols))
    return ARROYO_B200_INVALID_ARGUMENT;
  try {
    AB_CUDA(cudaSetDevice(x->device));
    const int W = x->world, nc = x->n_cols;
    long long* ctrl = x->ctrl.as<long long>();
    if (n_rows > 0) {
      int32_t st = arroyo_b200_partition_packed(x->part, in_cols, n_rows, (uint64_t)x->packed.p, (uint64_t)ctrl,
                                                (uint64_t)x->offsets.p);
      if (st != ARROYO_B200_OK) throw ab::Error(st, "partition_packed failed");
    } else {
      AB_CUDA(cudaMemsetAsync(ctrl, 0, (size_t)W * 8, x->stream));
    }
    const long long tail[2] = {(long long)watermark_ns, more ? 1ll : 0ll};
    AB_CUDA(cudaMemcpyAsync(ctrl + W, tail, sizeof tail, cudaMemcpyHostToDevice, x->stream));
    AB_NCCL(ncclAllGather(ctrl, x->ctrl_all.p, (size_t)(W + 2), ncclInt64, x->comm, x->stream));
    long long* h = x->h_ctrl_all.as<long long>();
    AB_CUDA(cudaMemcpyAsync(h, x->ctrl_all.p, (size_t)W * (W + 2) * 8, cudaMemcpyDeviceToHost, x->stream));
    AB_CUDA(cudaStreamSynchronize(x->stream));  // the one host wait of the round: split sizes
    int64_t n_recv = 0;
    *any_more = 0;
    for (int s = 0; s < W; ++s) {
      recv_rows[s] = h[(size_t)s * (W + 2) + x->rank];
      sender_watermarks[s] = h[(size_t)s * (W + 2) + W];
      if (h[(size_t)s * (W + 2) + W + 1]) *any_more = 1;
      n_recv += recv_rows[s];
    }
    AB_REQUIRE(n_recv <= x->max_recv_rows, ARROYO_B200_RUNTIME, "shuffle receive buffer too small");
    long long* rbuf = x->recv[x->flip].as<long long>();
    x->flip ^= 1;
    const long long* sbuf = x->packed.as<long long>();
    AB_NCCL(ncclGroupStart());
    int64_t soff = 0, roff = 0;
    for (int p = 0; p < W; ++p) {
      const int64_t sc = h[(size_t)x->rank * (W + 2) + p];  // rows this rank sends to p (its own control record)
      const int64_t rc = recv_rows[p];
      if (sc) AB_NCCL(ncclSend(sbuf + soff * nc, (size_t)sc * nc, ncclInt64, p, x->comm, x->stream));
      if (rc) AB_NCCL(ncclRecv(rbuf + roff * nc, (size_t)rc * nc, ncclInt64, p, x->comm, x->stream));
      for (int c = 0; c < nc; ++c) recv_cols[(size_t)p * nc + c] = rc ? (uint64_t)(rbuf + roff * nc + (int64_t)c * rc) : 0;
      soff += sc;
      roff += rc;
    }
    AB_NCCL(ncclGroupEnd());
    return ARROYO_B200_OK;
  } catch (const ab::Error& e) {
    x->last_error = e.what();
    return e.status;
  } catch (const NcclError& e) {
    x->last_error = e.what;
    return ARROYO_B200_RUNTIME;
  } catch (...) {
    x->last_error = "unknown C++ exception";
    return ARROYO_B200_RUNTIME;
  }
}

}  // extern "C"

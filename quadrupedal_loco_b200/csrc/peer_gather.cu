// peer_gather.cu -- once-per-batch gather of result rows to rank 0 WITHOUT a collective: every rank's kernels store their rows
// straight into rank 0's buffer over NVLink / NVSwitch peer memory, signalling with flags in the same memory.
//
// the batch shards with no collective inside the solve; what is left is the optional gather of (out, status) to
// rank 0.  torch.distributed.gather (NCCL) makes every batch a rendezvous of all ranks (measured at 8 GPUs: the e2e leg runs
// 25 % slower with it than without any gather).  Here rank 0 allocates `slots` result blocks [world][rows][row_doubles] plus a
// flag and an acknowledge array and exports them through CUDA IPC; a peer maps them and passes go1mpc_gather_dest() as the
// compact_d pointer of go1mpc_control_tick_host_async, so the tick's own pack kernel writes over NVLink.  Per slot:
//   peer   go1mpc_gather_acquire (device-side wait until rank 0 released the slot's previous use) -> tick -> go1mpc_gather_publish
//          (system-scope fence + store of the use's sequence number into rank 0's flag array)
//   rank 0 tick (its own rows, local) -> publish -> go1mpc_gather_wait_all (device-side wait for every rank's flag) -> read the
//          block (D2H) -> go1mpc_gather_release (acknowledge: sequence number into the ack array the peers poll over NVLink)
// Everything is stream ordered and the use counters advance on the device, so a tick with its gather calls can be captured
// into a CUDA graph and replayed; the host never blocks and no rank waits for another on the host.  The waits are single-thread
// kernels with a time-out (about 10 s of GPU clock): a rank that dies turns into an error code in `status`, not a hang.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include <new>
#include "../../include/go1mpc.h"

namespace {

struct GatherBlob {                     // what rank 0 exports (fits GO1MPC_GATHER_BLOB_BYTES)
  cudaIpcMemHandle_t mem;
  int world, slots, rows, row_doubles;
};

// The use counter of a slot lives in this rank's own memory (cnt) and is advanced ON THE DEVICE, so a captured CUDA graph of a
// tick (acquire -> tick -> publish ...) can be replayed: no sequence number is baked into a kernel argument.
// publish: advance the slot's use count, make the rows written before this kernel on this stream visible, raise the flag
__global__ void pg_publish_kernel(volatile unsigned* flag, unsigned* cnt) {
  const unsigned s = *cnt + 1;
  *cnt = s;
  __threadfence_system();
  *flag = s;
}
// release (rank 0): acknowledge the current use of the slot
__global__ void pg_release_kernel(volatile unsigned* ack, const unsigned* cnt) {
  __threadfence_system();
  *ack = *cnt;
}
// waits until every flag[0..n) has reached the slot's use count (wrap-safe); a count of 0 (first use) waits for nothing;
// on time-out sets *status = 1
__global__ void pg_wait_kernel(const volatile unsigned* flag, int n, const unsigned* cnt, int* status) {
  const unsigned seq = *cnt;
  if (seq == 0) return;
  const long long t0 = clock64();
  for (int r = 0; r < n; r++) {
    while ((int)(flag[r] - seq) < 0) {
      if (clock64() - t0 > 20000000000ll) { *status = 1; return; }
      __nanosleep(200);
    }
  }
  __threadfence_system();
}

}  // namespace

struct go1mpc_gather {
  int device = 0, rank = 0, world = 1, slots = 1, rows = 0, row_doubles = 0;
  bool root = false, mapped = false;
  char* base = nullptr;                 // rank 0's allocation (local pointer on rank 0, IPC mapping on peers)
  size_t block_bytes = 0, flags_off = 0, acks_off = 0, total = 0;
  int* status_d = nullptr;
  unsigned* cnt_d = nullptr;            // per slot: uses so far, advanced on the device by publish (identical on every rank by construction)
};

extern "C" {

int go1mpc_gather_create(int device, int rank, int world, int slots, int rows_per_rank, int row_doubles, go1mpc_gather_t** out) {
  if (!out || rank < 0 || world < 1 || rank >= world || slots < 1 || slots > 64 || rows_per_rank < 1 || row_doubles < 1) return GO1MPC_E_INVALID;
  go1mpc_gather* g = new (std::nothrow) go1mpc_gather();
  if (!g) return GO1MPC_E_CUDA;
  if (device < 0) cudaGetDevice(&device);
  g->device = device; g->rank = rank; g->world = world; g->slots = slots; g->rows = rows_per_rank; g->row_doubles = row_doubles;
  g->root = (rank == 0);
  g->block_bytes = (size_t)world * rows_per_rank * row_doubles * sizeof(double);
  g->flags_off = (size_t)slots * g->block_bytes;
  g->acks_off = g->flags_off + (size_t)slots * world * sizeof(unsigned);
  g->total = g->acks_off + (size_t)slots * sizeof(unsigned);
  if (cudaSetDevice(device) != cudaSuccess || cudaMalloc((void**)&g->status_d, sizeof(int)) != cudaSuccess ||
      cudaMemset(g->status_d, 0, sizeof(int)) != cudaSuccess || cudaMalloc((void**)&g->cnt_d, slots * sizeof(unsigned)) != cudaSuccess ||
      cudaMemset(g->cnt_d, 0, slots * sizeof(unsigned)) != cudaSuccess) { go1mpc_gather_destroy(g); return GO1MPC_E_CUDA; }
  if (g->root) {
    if (cudaMalloc((void**)&g->base, g->total) != cudaSuccess || cudaMemset(g->base, 0, g->total) != cudaSuccess) {
      go1mpc_gather_destroy(g); return GO1MPC_E_CUDA;
    }
  }
  *out = g;
  return GO1MPC_OK;
}

void go1mpc_gather_destroy(go1mpc_gather_t* g) {
  if (!g) return;
  cudaSetDevice(g->device);
  if (g->base) { if (g->root) cudaFree(g->base); else if (g->mapped) cudaIpcCloseMemHandle(g->base); }
  if (g->status_d) cudaFree(g->status_d);
  if (g->cnt_d) cudaFree(g->cnt_d);
  delete g;
}

/* rank 0: the blob every peer needs (broadcast it with whatever the host program uses, e.g. torch.distributed) */
int go1mpc_gather_export(go1mpc_gather_t* g, void* blob, int blob_bytes) {
  if (!g || !blob || !g->root || blob_bytes < (int)sizeof(GatherBlob)) return GO1MPC_E_INVALID;
  GatherBlob b;
  memset(&b, 0, sizeof b);
  if (cudaSetDevice(g->device) != cudaSuccess || cudaIpcGetMemHandle(&b.mem, g->base) != cudaSuccess) return GO1MPC_E_CUDA;
  b.world = g->world; b.slots = g->slots; b.rows = g->rows; b.row_doubles = g->row_doubles;
  memset(blob, 0, blob_bytes);
  memcpy(blob, &b, sizeof b);
  return GO1MPC_OK;
}

/* peers: map rank 0's memory (needs peer access between the two GPUs: NVLink / NVSwitch or PCIe P2P) */
int go1mpc_gather_import(go1mpc_gather_t* g, const void* blob, int blob_bytes) {
  if (!g || !blob || g->root || g->mapped || blob_bytes < (int)sizeof(GatherBlob)) return GO1MPC_E_INVALID;
  GatherBlob b;
  memcpy(&b, blob, sizeof b);
  if (b.world != g->world || b.slots != g->slots || b.rows != g->rows || b.row_doubles != g->row_doubles) return GO1MPC_E_INVALID;
  void* p = nullptr;
  if (cudaSetDevice(g->device) != cudaSuccess || cudaIpcOpenMemHandle(&p, b.mem, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return GO1MPC_E_CUDA;
  g->base = (char*)p; g->mapped = true;
  return GO1MPC_OK;
}

/* where THIS rank's rows of `slot` go: pass it as compact_d (rank 0: local memory; peers: rank 0's memory over NVLink) */
double* go1mpc_gather_dest(go1mpc_gather_t* g, int slot) {
  if (!g || !g->base || slot < 0 || slot >= g->slots) return nullptr;
  return (double*)(g->base + (size_t)slot * g->block_bytes) + (size_t)g->rank * g->rows * g->row_doubles;
}
/* rank 0: the whole block of `slot`, [world][rows][row_doubles] */
const double* go1mpc_gather_block(go1mpc_gather_t* g, int slot) {
  if (!g || !g->root || slot < 0 || slot >= g->slots) return nullptr;
  return (const double*)(g->base + (size_t)slot * g->block_bytes);
}

/* before a rank writes `slot` again: wait (on the device, in stream order) until rank 0 released its previous use */
int go1mpc_gather_acquire(go1mpc_gather_t* g, int slot, void* stream) {
  if (!g || !g->base || slot < 0 || slot >= g->slots) return GO1MPC_E_INVALID;
  if (cudaSetDevice(g->device) != cudaSuccess) return GO1MPC_E_CUDA;
  if (!g->root) {                                 // rank 0's own next write follows its release in stream order
    const volatile unsigned* ack = (const volatile unsigned*)(g->base + g->acks_off) + slot;
    pg_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ack, 1, g->cnt_d + slot, g->status_d);      // uses completed so far
  }
  return cudaGetLastError() == cudaSuccess ? GO1MPC_OK : GO1MPC_E_CUDA;
}
/* after the rows of `slot` were written on `stream`: make them visible and raise this rank's flag */
int go1mpc_gather_publish(go1mpc_gather_t* g, int slot, void* stream) {
  if (!g || !g->base || slot < 0 || slot >= g->slots) return GO1MPC_E_INVALID;
  if (cudaSetDevice(g->device) != cudaSuccess) return GO1MPC_E_CUDA;
  volatile unsigned* flag = (volatile unsigned*)(g->base + g->flags_off) + (size_t)slot * g->world + g->rank;
  pg_publish_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag, g->cnt_d + slot);
  return cudaGetLastError() == cudaSuccess ? GO1MPC_OK : GO1MPC_E_CUDA;
}
/* rank 0, after its own publish: wait (device, stream order) until every rank published this use of `slot` */
int go1mpc_gather_wait_all(go1mpc_gather_t* g, int slot, void* stream) {
  if (!g || !g->root || slot < 0 || slot >= g->slots) return GO1MPC_E_INVALID;
  if (cudaSetDevice(g->device) != cudaSuccess) return GO1MPC_E_CUDA;
  const volatile unsigned* flag = (const volatile unsigned*)(g->base + g->flags_off) + (size_t)slot * g->world;
  pg_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag, g->world, g->cnt_d + slot, g->status_d);
  return cudaGetLastError() == cudaSuccess ? GO1MPC_OK : GO1MPC_E_CUDA;
}
/* rank 0, after it consumed the block on `stream` (e.g. enqueued its D2H copy): let the peers write the slot again */
int go1mpc_gather_release(go1mpc_gather_t* g, int slot, void* stream) {
  if (!g || !g->root || slot < 0 || slot >= g->slots) return GO1MPC_E_INVALID;
  if (cudaSetDevice(g->device) != cudaSuccess) return GO1MPC_E_CUDA;
  volatile unsigned* ack = (volatile unsigned*)(g->base + g->acks_off) + slot;
  pg_release_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ack, g->cnt_d + slot);
  return cudaGetLastError() == cudaSuccess ? GO1MPC_OK : GO1MPC_E_CUDA;
}
/* 0 = fine, 1 = a device-side wait timed out (a rank is gone); synchronises the device */
int go1mpc_gather_status(go1mpc_gather_t* g, int* status) {
  if (!g || !status) return GO1MPC_E_INVALID;
  if (cudaSetDevice(g->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess ||
      cudaMemcpy(status, g->status_d, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) return GO1MPC_E_CUDA;
  return GO1MPC_OK;
}

}  // extern "C"

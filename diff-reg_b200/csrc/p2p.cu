// Peer-to-peer exchange buffers for the row-sharded Sinkhorn (one process per GPU on one NVSwitch node).
// Each rank owns one device allocation (inbox + flags + status), exports it with CUDA IPC, and maps every peer's
// allocation; skh_shard_exchange_kernel (sinkhorn.cu) then stores into the peers' inboxes directly over NVLink.
// The 64-byte handles travel between the processes by whatever the host has (torch.distributed all_gather here).
#include <string.h>

#include "common.cuh"

using namespace drg;

extern "C" size_t drg_p2p_handle_bytes(void) { return sizeof(cudaIpcMemHandle_t); }

extern "C" int drg_p2p_create(int rank, int world, size_t slot_elems, int nflags, void** comm_out, void* handle_out) {
  DRG_CHECK_ARG(comm_out && handle_out, "comm_out / handle_out is null");
  DRG_CHECK_ARG(world >= 1 && world <= P2P_MAX_RANKS && rank >= 0 && rank < world, "rank / world out of range (world <= 8)");
  DRG_CHECK_ARG(slot_elems >= 1 && nflags >= 1, "slot_elems and nflags must be >= 1");
  P2PComm* c = new P2PComm();
  memset(c, 0, sizeof(P2PComm));
  c->rank = rank;
  c->world = world;
  c->slot_elems = slot_elems;
  c->nflags = nflags;
  c->bytes = p2p_inbox_bytes(*c) + 2 * (size_t)nflags * sizeof(unsigned int) + 256;
  cudaError_t e = cudaMalloc(&c->base, c->bytes);
  if (e != cudaSuccess) {
    set_error("p2p: cudaMalloc(%zu) failed: %s", c->bytes, cudaGetErrorString(e));
    delete c;
    return DRG_ERR_CUDA;
  }
  e = cudaMemset(c->base, 0, c->bytes);
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, c->base);
  if (e != cudaSuccess) {
    set_error("p2p: cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    cudaFree(c->base);
    delete c;
    return DRG_ERR_CUDA;
  }
  memcpy(handle_out, &h, sizeof(h));
  c->peer_base[rank] = c->base;
  *comm_out = c;
  return DRG_OK;
}

// handles: world x drg_p2p_handle_bytes() bytes, rank-major (every rank's handle_out of drg_p2p_create)
extern "C" int drg_p2p_connect(void* comm, const void* handles) {
  DRG_CHECK_ARG(comm && handles, "comm / handles is null");
  P2PComm* c = reinterpret_cast<P2PComm*>(comm);
  for (int r = 0; r < c->world; ++r) {
    if (r == c->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, reinterpret_cast<const char*>(handles) + (size_t)r * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      set_error("p2p: cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
      return DRG_ERR_CUDA;
    }
    c->peer_base[r] = ptr;
    c->peer_open[r] = true;
  }
  return DRG_OK;
}

// 0: every wait so far completed; 1: a wait timed out (results of that call are invalid).  Synchronises the device.
extern "C" int drg_p2p_status(void* comm) {
  DRG_CHECK_ARG(comm != nullptr, "comm is null");
  P2PComm* c = reinterpret_cast<P2PComm*>(comm);
  int st = 0;
  const char* sp = reinterpret_cast<const char*>(c->base) + p2p_inbox_bytes(*c) + 2 * (size_t)c->nflags * sizeof(unsigned int);
  DRG_CUDA(cudaMemcpy(&st, sp, sizeof(int), cudaMemcpyDeviceToHost));
  return st;
}

extern "C" int drg_p2p_destroy(void* comm) {
  if (!comm) return DRG_OK;
  P2PComm* c = reinterpret_cast<P2PComm*>(comm);
  for (int r = 0; r < c->world; ++r)
    if (c->peer_open[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
  cudaFree(c->base);
  delete c;
  return DRG_OK;
}
